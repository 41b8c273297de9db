// patch_embed.cu — N4 ("next" row 4 of SURVEY §8f): the cost-map PatchEmbed projection.
// Replaces the conv stack of PatchEmbed.forward (core/FlowFormer/PerCostFormer3/encoder.py:36-43, :60-73;
// called for every cost map at encoder.py:263 with patch_size = 8, patch_embed = "single", embed_dim = 64):
//     x [N, 1, 64, 64] -> Conv2d(1, 16, 6, s2, p2) -> ReLU -> Conv2d(16, 32, 6, s2, p2) -> ReLU
//                      -> Conv2d(32, 64, 6, s2, p2)                                   -> [N, 64, 8, 8]
// 20 MFLOP per cost map, N = B * 4096 maps: 1.3 TFLOP per direction at batch 16 — the only other consumer that
// streams the whole fp32 volume, and tensor-bound (HBM floor: 16 KB in + 16 KB out per map).
//
// Formulation.  A 6x6 / stride-2 / pad-2 convolution is a 3x3 / stride-1 / pad-1 convolution over the
// space-to-depth (2x2 -> 4 channels) image: ky = 2 dy + py + 2, kx = 2 dx + px + 2 with dy, dx in {-1,0,1},
// py, px in {0,1}.  Activations live in shared memory as a FLAT stream of 16-byte entries (8 bf16 channels), one
// array per 8-channel chunk, rows of OW + 1 entries whose last entry is a zero column, with one zero row above and
// below:  entry(iy, ix) = 1 + (iy + 1) * (OW + 1) + ix.   Output pixel (oy, ox) reads, for tap (dy, dx), the entry
// (OW + 1) * (1 + dy) + dx + 1 - 1 after its own: a CONSTANT offset — the left / right zero padding is the zero column
// of the previous / same row, the top / bottom padding the zero rows.  So the A operand of an implicit GEMM is the
// activation buffer itself, un-swizzled K-major: a core matrix = 8 consecutive entries x 16 bytes = 8 horizontally
// adjacent pixels, LBO = the chunk stride, and SBO (the distance between consecutive core matrices of M) = ONE IMAGE
// ROW of the buffer, (OW + 1) * 16 bytes.  The 128 rows a CTA contributes to an instruction are therefore an
// 8-pixel-wide, 16-row-high strip of the output: the zero column is never an output row, no im2col copy, no per-tap
// staging.  (First version: SBO = 128 B, M = the flat index including the zero column: 9 / 3 / 1 blocks of 128
// rows with 1 junk row in OW + 1, 12 % slower.)
//   conv1 (1 -> 16):  K per tap is only 4, so an entry holds TWO horizontally adjacent space-to-depth pixels
//                     [s2d(ix), s2d(ix+1)]; a K = 16 step is one dy: entry ox-1 (dx = -1, 0) and entry ox+1 (dx = +1, pad).
//                     M = 32 * 32 outputs -> 8 strips of 16 x 8, N = 16, K = 48.
//   conv2 (16 -> 32): 64 space-to-depth channels = 8 chunks; M = 16 * 16 -> 2 strips, N = 32, K = 9 * 64 = 576.
//   conv3 (32 -> 64): 128 channels = 16 chunks; M = 8 * 8 = half a strip (rows 64..127 of the instruction are
//                     unused), N = 64, K = 9 * 128 = 1152.
// Each epilogue adds the bias, applies ReLU, rounds to bf16 and scatters straight into the next layer's flat
// space-to-depth buffer; the last one writes the fp32 [64, 8, 8] result through a staged 16 KB bulk store.
//
// Machine mapping.  The bf16 weights (186 KB) do not fit one SM next to the activations, so the kernel runs as
// clusters of two CTAs with tcgen05.mma.cta_group::2: M = 256 per instruction — each CTA supplies the 128 rows of
// ITS OWN cost map and HALF of the weight rows (N/2), 93 KB per SM.  Per CTA (16 warps): warps 0 / 2 / 3 = MMA
// issuers of conv1 / conv2 / conv3 (leader CTA; rolled loops on the uniform datapath, see there), warp 1 = TMEM
// allocator, warps 4-7 / 8-11 / 12-15 = three epilogue sets (TMEM lane quarter = warp % 4) that share conv1's
// epilogue; set A also runs conv3's, set B conv2's, set C is the loader (fp32 map -> bf16 space-to-depth entries).
// The three layers run on three consecutive maps at once (conv1(t), conv2(t-1), conv3(t-2)); the conv2 / conv3
// accumulators are double-buffered in tensor memory.  The tensor pipe executes MMAs in issue order, so the ORDER of
// issue is part of the design: conv3(t) is queued behind conv2(t+1) and runs while conv1's epilogue converts map
// t+2 (see the conv3 issuer).  tools/pe_trace.py prints the pipeline's timeline from a -DSB_PE_TRACE build.
//
// Measured dead ends (round 2, all bit-identical, tools/pe_variants.sh): double-buffering A2 at the cost of a single
// A1 and an unstaged output store (+16 %); all of E1's tcgen05.ld issued up front with conv1 released by a separate
// "D1 drained" barrier (+3 %), the same with conv3 (or its second half) held back until those loads have landed
// (+13..45 %: a tcgen05.ld completes only after the MMAs issued before it, but coupling the conv3 issue to the
// epilogue warps serialises them with E2 / E3), E1 on two dedicated sets with E2 / E3 on the loader warps (+4..25 %).
#include <cuda_bf16.h>

#include "common.cuh"
#include "ptx_sm100.cuh"

namespace sb {
namespace pe {

constexpr int kThreads = 512;
// ---- flat activation geometry (entries of 16 bytes)
constexpr int kRow1 = 33, kRow2 = 17, kRow3 = 9;          // OW + 1 of conv1 / conv2 / conv3
constexpr int kMB1 = 8, kMB2 = 2;                         // 128-row M blocks per map (conv3: one, half used)
constexpr int kA1Entries = 1220;                           // 1 + 34 * 33 data, reads reach 9*128 - 1 + 66 + 2
constexpr int kA1Bytes = kA1Entries * 16;                  // 19 520
#ifndef SB_PE_CH2
#define SB_PE_CH2 308
#define SB_PE_CH3 92
#endif
#ifndef SB_PE_EXP
#define SB_PE_EXP 0         // TIMING-ONLY experiments (results are wrong): 1 = 128-byte aligned tap offsets, 2 = no dx != 0 taps,
#endif                      // 4 = epilogues / loaders do not store to shared memory (tools/pe_variants.sh)
#ifndef SB_PE_ORDER
#define SB_PE_ORDER 1       // conv3(t) is issued after conv2(t + 1) has been queued (see the conv3 issuer)
#endif
#ifndef SB_PE_BIAS12
#define SB_PE_BIAS12 0      // 1: b1 / b2 read once per stage as LDS.128 into registers
#endif
#ifndef SB_PE_BIAS3
#define SB_PE_BIAS3 1       // 1: b3 read as LDS.128 (4 channels per load): measured -2.3 %
#endif
constexpr int kCh2 = SB_PE_CH2, kCh3 = SB_PE_CH3;          // chunk strides (entries): >= 1 + 18*17 = 307, 1 + 10*9 = 91
// (measured, round 2: strides 310 / 93 and 312 / 96 — which make the epilogue's 16-byte stores bank-conflict free —
//  change nothing; b1 / b2 hoisted into registers once per stage: +7 %; in-loop LDS.128 for b1 / b2: no change;
//  b3 as LDS.128: -2.3 % and kept.  The SB_PE_* macros select these variants at compile time.)
constexpr int kA2Bytes = (8 * kCh2 + 112) * 16;            // + tail the last chunk's junk rows read (finite zeros)
constexpr int kA3Bytes = (16 * kCh3 + 56) * 16;
// ---- weights per CTA (half of the rows), K-major un-swizzled: [k/8][n/8][n%8][k%8]
constexpr int kW1Bytes = 6 * 128;                          // 8 rows  x K 48
constexpr int kW2Bytes = 72 * 256;                         // 16 rows x K 576
constexpr int kW3Bytes = 144 * 512;                        // 32 rows x K 1152
constexpr int kWBytes = kW1Bytes + kW2Bytes + kW3Bytes;    // 92 928 per CTA rank
constexpr int kBiasFloats = 16 + 32 + 64;
constexpr int kStageBytes = 64 * 64 * 4;                   // fp32 [64 ch][64 pos] of one map
// ---- shared memory map
constexpr int oW1 = 0, oW2 = oW1 + kW1Bytes, oW3 = oW2 + kW2Bytes;
constexpr int oA1 = oW3 + kW3Bytes;                        // two buffers
constexpr int oA2 = oA1 + 2 * kA1Bytes;
constexpr int oA3 = oA2 + kA2Bytes;
constexpr int oStage = oA3 + kA3Bytes;
constexpr int oBias = oStage + kStageBytes;
constexpr int oBar = oBias + kBiasFloats * 4;
constexpr int kSmemTotal = oBar + 256;
static_assert(oA1 % 16 == 0 && oA2 % 16 == 0 && oA3 % 16 == 0 && oStage % 16 == 0 && oBar % 8 == 0, "alignment");
static_assert(kSmemTotal <= 232448, "shared memory budget");
// ---- tensor memory columns
constexpr int kD2Cols = kMB2 * 32, kD3Cols = 64;                            // one accumulator set of conv2 / conv3
constexpr int kD1Col = 0, kD2Col = kMB1 * 16, kD3Col = kD2Col + 2 * kD2Cols;  // D1 144 | D2 x2 192 | D3 x2 128 = 464 <= 512
constexpr int kTmemCols = 512;

// instruction descriptor: D = f32, A = B = bf16, both K-major, M = 256 (CTA pair), N = n
__host__ __device__ constexpr uint32_t idesc(int n) {
  return (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(256 >> 4) << 24);
}

// K-major, un-swizzled (interleaved) shared-memory matrix descriptor: rows of a core matrix 16 bytes apart,
// 8-row groups SBO apart, the two 16-byte K chunks of a K = 16 step LBO apart.
__device__ __forceinline__ uint64_t desc_nosw(uint32_t addr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
  uint64_t d = 0;
  d |= (uint64_t)((addr >> 4) & 0x3FFFu);
  d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFFu) << 16;
  d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFFu) << 32;
  d |= (uint64_t)1 << 46;                          // descriptor version (Blackwell); layout type 0 = no swizzle
  return d;
}

__device__ __forceinline__ void tmem_ld_x16(uint32_t taddr, uint32_t (&r)[16]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr)
      : "memory");
}

__device__ __forceinline__ void bulk_store_1d(void* gdst, uint32_t smem_src, uint32_t bytes) {
  asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(reinterpret_cast<uint64_t>(gdst)),
               "r"(smem_src), "r"(bytes)
               : "memory");
}

__device__ __forceinline__ uint32_t pack_bf16(float a, float b) {
  __nv_bfloat162 h = __floats2bfloat162_rn(a, b);
  return *reinterpret_cast<uint32_t*>(&h);
}
__device__ __forceinline__ float2 ldg_stream2(const float* p) {
  float2 v;
  asm volatile("ld.global.nc.L1::no_allocate.v2.f32 {%0,%1}, [%2];" : "=f"(v.x), "=f"(v.y) : "l"(p));
  return v;
}
#define SB_PE_STS_GUARD if ((SB_PE_EXP & 4) && addr != 0xFFFFFFF0u) return;
__device__ __forceinline__ void sts128(uint32_t addr, uint32_t a, uint32_t b, uint32_t c, uint32_t d) {
  SB_PE_STS_GUARD
  asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(a), "r"(b), "r"(c), "r"(d) : "memory");
}
__device__ __forceinline__ void sts64(uint32_t addr, uint32_t a, uint32_t b) {
  SB_PE_STS_GUARD
  asm volatile("st.shared.v2.b32 [%0], {%1, %2};" ::"r"(addr), "r"(a), "r"(b) : "memory");
}
__device__ __forceinline__ void sts32f(uint32_t addr, float v) {
  SB_PE_STS_GUARD
  asm volatile("st.shared.f32 [%0], %1;" ::"r"(addr), "f"(v) : "memory");
}

#ifdef SB_PE_TRACE
// Debug build only (tools/pe_variants.sh "trace:-DSB_PE_TRACE", tools/pe_trace.py): clock64() of the pipeline events of
// CTA 0 in iterations 16..23, 16 events each.
__device__ long long g_pe_trace[8 * 16 + 2 * 148];   // + per CTA: cycles of the whole kernel, of the steady loop
__device__ long long g_pe_acc[74 * 8];                // per cluster: issuer i (conv1/2/3): [2i] cycles waiting, [2i+1] issuing; [6] smid
#define PE_TRACE(ev) do { if (blockIdx.x == 0 && lane == 0 && t >= 16 && t < 24) g_pe_trace[(t - 16) * 16 + (ev)] = clock64(); } while (0)
#define PE_ACC_DECL long long acc_w = 0, acc_i = 0, acc_t = 0
#define PE_ACC_TOP acc_t = clock64()
#define PE_ACC_START do { const long long c = clock64(); acc_w += c - acc_t; acc_t = c; } while (0)
#define PE_ACC_END do { acc_i += clock64() - acc_t; } while (0)
#define PE_ACC_STORE(i) do { if (lane == 0 && cluster_id < 74) { g_pe_acc[cluster_id * 8 + 2 * (i)] = acc_w; g_pe_acc[cluster_id * 8 + 2 * (i) + 1] = acc_i; \
    if ((i) == 0) { unsigned smid; asm volatile("mov.u32 %0, %%smid;" : "=r"(smid)); g_pe_acc[cluster_id * 8 + 6] = smid; } } } while (0)
#else
#define PE_TRACE(ev) do { } while (0)
#define PE_ACC_DECL do { } while (0)
#define PE_ACC_TOP do { } while (0)
#define PE_ACC_START do { } while (0)
#define PE_ACC_END do { } while (0)
#define PE_ACC_STORE(i) do { } while (0)
#endif
struct Params {
  const float* maps;        // [NQ, 64, 64] fp32 (cost_maps [NQ, 1, 64, 64])
  float* out;               // [NQ, 64, 8, 8] fp32
  const uint8_t* wpack;     // 2 x kWBytes: the packed bf16 weight halves of CTA rank 0 and 1
  const float* bias;        // 16 + 32 + 64 fp32
  long long nq;
  int iters;                // maps per CTA (loop trip count, equal for all CTAs)
  int nclusters;
  unsigned int* dbg;
};

__global__ void __launch_bounds__(kThreads, 1)
patch_embed_umma_kernel(const Params p) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  const uint32_t sm = ptx::smem_u32(smem_raw);
#ifdef SB_PE_TRACE
  const long long trace_k0 = clock64();
#endif
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t rank = ptx::cluster_ctarank();
  const bool leader = rank == 0;
  const long long cluster_id = blockIdx.x >> 1;

  // barriers (8 bytes each, same offsets in both CTAs).  "leader": only the leader CTA's copy is used, both CTAs'
  // warps arrive on it through shared::cluster; "both": a multicast tcgen05.commit arrives on each CTA's copy.
  // Everything that is double-buffered has one barrier per buffer, so that a waiter is never more than one phase behind.
  const uint32_t bar = sm + oBar;
  const uint32_t bar_w = bar + 0;                  // weights landed (local, tx)
  const uint32_t bar_ldr_full = bar + 8;           // [2] leader: 8 loader-warp arrivals
  const uint32_t bar_a1_free = bar + 24;           // [2] both: conv1 MMAs of that A1 buffer retired
  const uint32_t bar_d1_full = bar + 40;           // both: conv1 accumulators ready
  const uint32_t bar_m2_done = bar + 48;           // [2] both: conv2 into D2[i] retired (D2[i] ready, A2 reusable)
  const uint32_t bar_m3_done = bar + 64;           // [2] both: conv3 into D3[i] retired (D3[i] ready, A3 reusable)
  const uint32_t bar_e1_done = bar + 80;           // leader: 24 warp arrivals (D1 drained, A2 written)
  const uint32_t bar_e2_done = bar + 88;           // [2] leader: 8 arrivals (D2[i] drained, A3 written)
  const uint32_t bar_e3_done = bar + 104;          // [2] leader: 8 arrivals (D3[i] drained)
  const uint32_t tmem_slot = bar + 128;
  const uint32_t bar_c2_issued = bar + 136;        // [2] leader, local: the conv2 issuer has queued conv2(t) (orders conv3 behind it)

  if (threadIdx.x == 0) {
    ptx::mbar_init(bar_w, 1);
    for (int i = 0; i < 2; ++i) {
      ptx::mbar_init(bar_ldr_full + 8 * i, 8); ptx::mbar_init(bar_a1_free + 8 * i, 1);
      ptx::mbar_init(bar_m2_done + 8 * i, 1);  ptx::mbar_init(bar_m3_done + 8 * i, 1);
      ptx::mbar_init(bar_e2_done + 8 * i, 8);  ptx::mbar_init(bar_e3_done + 8 * i, 8);
    }
    ptx::mbar_init(bar_d1_full, 1);
    ptx::mbar_init(bar_e1_done, 24);
    ptx::mbar_init(bar_c2_issued, 1); ptx::mbar_init(bar_c2_issued + 8, 1);
    ptx::fence_mbar_init();
    // this CTA's half of the weights: three bulk copies, one transaction barrier
    ptx::mbar_arrive_expect_tx(bar_w, kWBytes);
    const uint8_t* wsrc = p.wpack + (size_t)rank * kWBytes;
    ptx::bulk_load_1d(sm + oW1, wsrc, kW1Bytes, bar_w);
    ptx::bulk_load_1d(sm + oW2, wsrc + kW1Bytes, kW2Bytes, bar_w);
    ptx::bulk_load_1d(sm + oW3, wsrc + kW1Bytes + kW2Bytes, kW3Bytes, bar_w);
  }
  if (warp == 1) { ptx::tmem_alloc_2cta(tmem_slot, kTmemCols); ptx::tmem_relinquish_2cta(); }
  // zero the activation buffers once: padding rows / columns / tails are never written again
  for (int i = threadIdx.x; i < (oStage - oA1) / 16; i += kThreads)
    sts128(sm + oA1 + i * 16, 0u, 0u, 0u, 0u);
  for (int i = threadIdx.x; i < kBiasFloats; i += kThreads)
    reinterpret_cast<float*>(smem_raw + oBias)[i] = __ldg(p.bias + i);
  ptx::fence_proxy_async_smem();
  ptx::tc_fence_before_sync();
  __syncthreads();
  ptx::mbar_wait(bar_w, 0, 1, p.dbg);               // every thread: the weight bytes are visible
  ptx::cluster_sync_all();                          // peer's barriers, weights and zeroed buffers are ready
  ptx::tc_fence_after_sync();
  const uint32_t tmem_base = *reinterpret_cast<volatile uint32_t*>(smem_raw + oBar + 128);
  const int T = p.iters;
  const float* bias = reinterpret_cast<const float*>(smem_raw + oBias);
#ifdef SB_PE_TRACE
  const long long trace_loop0 = clock64();
#endif

  // ------------------------------------------------------------------ E1: D1 -> bias, ReLU, bf16 -> conv2's flat s2d buffer
  // Blocks [mb0, mb1) of map t, by the calling warp's TMEM lane quarter.  Three warp sets share the nine blocks
  // (E1 sits between conv2(t-1) and conv2(t) on the critical path: A2 is single-buffered).
  const int wq = warp & 3;
  const uint32_t lane_t = tmem_base + ((uint32_t)(wq * 32) << 16);
  const uint32_t e1_tgt = ptx::mapa_shared(bar_e1_done, 0);
  auto e1_blocks = [&](const int t, const int mb0, const int mb1) {
    ptx::mbar_wait(bar_d1_full, (uint32_t)(t & 1), 8, p.dbg);
    if (t >= 1) ptx::mbar_wait(bar_m2_done + 8 * ((t - 1) & 1), (uint32_t)(((t - 1) >> 1) & 1), 9, p.dbg);   // conv2(t-1) has read A2
    ptx::tc_fence_after_sync();
    if (wq == 0 && mb0 < 6) PE_TRACE(mb0 == 0 ? 6 : 8);
#if SB_PE_BIAS12 == 1
    float b1[16];
#pragma unroll
    for (int c = 0; c < 4; ++c) {
      const float4 v = reinterpret_cast<const float4*>(bias)[c];
      b1[4 * c] = v.x; b1[4 * c + 1] = v.y; b1[4 * c + 2] = v.z; b1[4 * c + 3] = v.w;
    }
#else
    const float* b1 = bias;
#endif
    uint32_t r[2][16];
    tmem_ld_x16(lane_t + kD1Col + mb0 * 16, r[0]);
#pragma unroll 1
    for (int mb = mb0; mb < mb1; ++mb) {
      const int cur = (mb - mb0) & 1;
      ptx::tmem_ld_wait();
      if (mb + 1 < mb1) {                          // next block's load in flight while this one is converted
        if (cur == 0) tmem_ld_x16(lane_t + kD1Col + (mb + 1) * 16, r[1]);
        else tmem_ld_x16(lane_t + kD1Col + (mb + 1) * 16, r[0]);
      }
      const int row = wq * 32 + lane;              // core matrix row / 8 = image row of the strip, row % 8 = column
      const int oy = (mb >> 2) * 16 + (row >> 3), ox = (mb & 3) * 8 + (row & 7);
      {
        uint32_t w[8];
#if SB_PE_BIAS12 == 2
#pragma unroll
        for (int c4 = 0; c4 < 4; ++c4) {
          const float4 bv = reinterpret_cast<const float4*>(bias)[c4];
          const uint32_t x0 = cur == 0 ? r[0][4 * c4] : r[1][4 * c4], x1 = cur == 0 ? r[0][4 * c4 + 1] : r[1][4 * c4 + 1];
          const uint32_t x2 = cur == 0 ? r[0][4 * c4 + 2] : r[1][4 * c4 + 2], x3 = cur == 0 ? r[0][4 * c4 + 3] : r[1][4 * c4 + 3];
          w[2 * c4] = pack_bf16(fmaxf(__uint_as_float(x0) + bv.x, 0.0f), fmaxf(__uint_as_float(x1) + bv.y, 0.0f));
          w[2 * c4 + 1] = pack_bf16(fmaxf(__uint_as_float(x2) + bv.z, 0.0f), fmaxf(__uint_as_float(x3) + bv.w, 0.0f));
        }
#else
#pragma unroll
        for (int c = 0; c < 8; ++c) {
          const uint32_t x0 = cur == 0 ? r[0][2 * c] : r[1][2 * c], x1 = cur == 0 ? r[0][2 * c + 1] : r[1][2 * c + 1];
          w[c] = pack_bf16(fmaxf(__uint_as_float(x0) + b1[2 * c], 0.0f), fmaxf(__uint_as_float(x1) + b1[2 * c + 1], 0.0f));
        }
#endif
        // s2d channel (py, px, c1) -> chunk (py*2 + px)*2 + c1/8 at entry 1 + (oy/2 + 1) * 17 + ox/2
        const uint32_t chunk = (uint32_t)(((oy & 1) * 2 + (ox & 1)) * 2);
        const uint32_t e = sm + oA2 + (chunk * kCh2 + (uint32_t)(1 + ((oy >> 1) + 1) * kRow2 + (ox >> 1))) * 16u;
        sts128(e, w[0], w[1], w[2], w[3]);
        sts128(e + kCh2 * 16u, w[4], w[5], w[6], w[7]);
      }
    }
    ptx::fence_proxy_async_smem();
    ptx::tc_fence_before_sync();
    __syncwarp();
    if (lane == 0) ptx::mbar_arrive_cluster(e1_tgt);
    if (wq == 0 && mb0 < 6) PE_TRACE(mb0 == 0 ? 7 : 9);
  };

  if (warp == 0 || warp == 2 || warp == 3) {
    // ================================================================ MMA issuers (leader CTA, one warp each)
    // One issuing warp per layer: N = 16 / 32 / 64 instructions execute in 8 / 16 / 32 cycles, less than it takes
    // one thread to issue them back to back (first version: 207 MMAs per iteration from one thread, 87 cycles each
    // with the descriptors rebuilt per instruction -> 4.2 ms per 65 536 maps).  Descriptors are a constant high word
    // and a running low word (14-bit address field + LBO): one add per operand and instruction.
    if (leader) {                    // the whole warp runs the loop (converged); one elected lane issues each instruction
      // Rolled loops (one tap per trip), the election done once per layer, and every descriptor word derived from
      // warp-uniform values: the issue code of a layer is a few hundred bytes that stay in the instruction cache
      // (fully unrolled it was 350 B per MMA, 40 KB per iteration streamed from L2 -- and the SMs of a GPC then ran
      // at visibly different speeds: 3.07 M to 4.0 M cycles for the same 443 maps, tools/pe_trace.py).
      const uint32_t tmem_u = __shfl_sync(0xffffffffu, tmem_base, 0);
      const uint32_t sm_u = __shfl_sync(0xffffffffu, sm, 0);
      // High descriptor word = SBO (bytes >> 4) + descriptor version 1.  SBO is the distance between consecutive 8-row
      // core matrices of the M dimension.  Weights: 128 B (dense).  Activations: ONE ROW of the flat buffer (OW + 1
      // entries), so that core matrix k is 8 consecutive pixels of image row k: an instruction's 128 rows per CTA are
      // an 8-pixel-wide, 16-row-high strip of the output and the zero column is never an output row.
      auto lo_of = [](uint32_t addr, uint32_t lbo_bytes) { return ((addr >> 4) & 0x3FFFu) | ((lbo_bytes >> 4) << 16); };
      auto d64 = [](uint32_t sbo_entries, uint32_t lo) { return ((uint64_t)(sbo_entries | (1u << 14)) << 32) | (uint64_t)lo; };
      if (warp == 0) {
        // ---------------- conv1(t): needs the entries of map t (loaders) and D1 drained (E1(t-1))
        const uint32_t b_lo = lo_of(sm_u + oW1, 128u);
        PE_ACC_DECL;
        for (int t = 0; t < T; ++t) {
          const int buf = t & 1;
          PE_ACC_TOP;
          if (t >= 1) ptx::mbar_wait(bar_e1_done, (uint32_t)((t - 1) & 1), 2, p.dbg);
          ptx::mbar_wait(bar_ldr_full + 8 * buf, (uint32_t)((t >> 1) & 1), 3, p.dbg);
          ptx::tc_fence_after_sync();
          PE_TRACE(0); PE_ACC_START;
          uint32_t a_lo = lo_of(sm_u + oA1 + buf * kA1Bytes, 32u);    // entries (ox - 1) and (ox + 1): LBO = 2 entries
          uint32_t d = tmem_u + kD1Col;
          const uint32_t el = ptx::elect_one();
#pragma unroll 1
          for (int mb = 0; mb < kMB1; ++mb) {     // block = rows 16 (mb >> 2) .. + 15, columns 8 (mb & 3) .. + 7
            const uint32_t a_blk = a_lo + (uint32_t)((mb >> 2) * 16 * kRow1 + (mb & 3) * 8);
#pragma unroll
            for (int ks = 0; ks < 3; ++ks)        // ks = dy + 1: input row oy + dy
              ptx::umma_f16_2cta_if(el, d, d64(kRow1, a_blk + (uint32_t)(kRow1 * ks)), d64(8, b_lo + (uint32_t)(ks * 16)), idesc(16), ks != 0);
            d += 16u;
          }
          ptx::umma_commit_2cta_elect(bar_a1_free + 8 * buf, 3);
          ptx::umma_commit_2cta_elect(bar_d1_full, 3);
          PE_TRACE(1); PE_ACC_END;
        }
        PE_ACC_STORE(0);
      } else if (warp == 2) {
        // ---------------- conv2(t) -> D2[t & 1]: needs A2 written (E1(t)) and that accumulator drained (E2(t-2))
        const uint32_t a_base = lo_of(sm_u + oA2, kCh2 * 16u), b_base = lo_of(sm_u + oW2, 256u);
        PE_ACC_DECL;
        for (int t = 0; t < T; ++t) {
          const int buf = t & 1;
          PE_ACC_TOP;
          ptx::mbar_wait(bar_e1_done, (uint32_t)(t & 1), 4, p.dbg);
          if (t >= 2) ptx::mbar_wait(bar_e2_done + 8 * buf, (uint32_t)(((t >> 1) - 1) & 1), 5, p.dbg);
          ptx::tc_fence_after_sync();
          PE_TRACE(2); PE_ACC_START;
          uint32_t a_lo = a_base, d = tmem_u + kD2Col + buf * kD2Cols;
          const uint32_t el = ptx::elect_one();
#pragma unroll 1
          for (int mb = 0; mb < kMB2; ++mb) {
            uint32_t b_lo = b_base;
#pragma unroll 1
            for (int dy = 0; dy < 3; ++dy) {
#pragma unroll 1
              for (int dx = 0; dx < 3; ++dx) {
                const uint32_t a_tap = a_lo + (uint32_t)(kRow2 * dy + dx);
#pragma unroll
                for (int j = 0; j < 4; ++j)
                  ptx::umma_f16_2cta_if(el, d, d64(kRow2, a_tap + (uint32_t)(2 * j * kCh2)), d64(8, b_lo + (uint32_t)(32 * j)), idesc(32),
                                        (uint32_t)((dy | dx | j) != 0));
                b_lo += 128u;
              }
            }
            a_lo += 8u; d += 32u;                  // block = all 16 rows, columns 8 mb .. + 7
          }
          ptx::umma_commit_2cta_elect(bar_m2_done + 8 * buf, 3);
          PE_TRACE(3); PE_ACC_END;
#if SB_PE_ORDER
          if (lane == 0) ptx::mbar_arrive(bar_c2_issued + 8 * buf);
#endif
        }
        PE_ACC_STORE(1);
      } else {
        // ---------------- conv3(t) -> D3[t & 1]: needs A3 written (E2(t)) and that accumulator drained (E3(t-2))
        const uint32_t a_base = lo_of(sm_u + oA3, kCh3 * 16u), b_base = lo_of(sm_u + oW3, 512u);
        PE_ACC_DECL;
        for (int t = 0; t < T; ++t) {
          const int buf = t & 1;
          PE_ACC_TOP;
          ptx::mbar_wait(bar_e2_done + 8 * buf, (uint32_t)((t >> 1) & 1), 6, p.dbg);
          if (t >= 2) ptx::mbar_wait(bar_e3_done + 8 * buf, (uint32_t)(((t >> 1) - 1) & 1), 13, p.dbg);
#if SB_PE_ORDER
          // Tensor-pipe order.  The pipe executes MMAs in issue order.  E1(t+1) can only start when conv1(t+1) AND
          // conv2(t) have retired, and while it runs no conv1 / conv2 work exists -- so conv3(t) is queued BEHIND
          // conv2(t+1)... i.e. behind the conv2 of the next map, and executes during that epilogue instead of competing
          // with the two layers the epilogue is waiting for (without this the pipe idles for the whole of E1).
          if (t + 1 < T) ptx::mbar_wait(bar_c2_issued + 8 * ((t + 1) & 1), (uint32_t)(((t + 1) >> 1) & 1), 14, p.dbg);
#endif
          ptx::tc_fence_after_sync();
          PE_TRACE(4); PE_ACC_START;
          const uint32_t d = tmem_u + kD3Col + buf * kD3Cols;
          const uint32_t el = ptx::elect_one();
          uint32_t b_lo = b_base;
#pragma unroll 1
          for (int dy = 0; dy < 3; ++dy) {
#pragma unroll 1
            for (int dx = 0; dx < 3; ++dx) {
              const uint32_t a_tap = a_base + (uint32_t)(kRow3 * dy + dx);
#pragma unroll
              for (int j = 0; j < 8; ++j)
                ptx::umma_f16_2cta_if(el, d, d64(kRow3, a_tap + (uint32_t)(2 * j * kCh3)), d64(8, b_lo + (uint32_t)(64 * j)), idesc(64),
                                      (uint32_t)((dy | dx | j) != 0));
              b_lo += 512u;
            }
          }
          ptx::umma_commit_2cta_elect(bar_m3_done + 8 * buf, 3);
          PE_TRACE(5); PE_ACC_END;
        }
        PE_ACC_STORE(2);
      }
    }
    __syncwarp();
  } else if (warp >= 12) {
    // ================================================================ loaders: fp32 map -> bf16 s2d pair entries,
    // and the third E1 set.  Map t + 1 is written to A1 and map t + 2 is on its way in registers before the warp
    // turns to E1(t), so conv1 never waits for its input.
    const int tid = threadIdx.x - 384;               // 0..127
    const uint32_t full_tgt0 = ptx::mapa_shared(bar_ldr_full, 0);
    float2 top[8], bot[8];
    auto fetch = [&](const int t) {                  // global loads of map t into registers
      const long long q = ((long long)t * p.nclusters + cluster_id) * 2 + rank;
      if (t < T && q < p.nq) {
        const float* src = p.maps + q * 4096;
#pragma unroll
        for (int j = 0; j < 8; ++j) {                // s2d position (r, i) = (pos >> 5, pos & 31), pos = tid + 128 j
          const int pos = tid + 128 * j, r = pos >> 5, i = pos & 31;
          top[j] = ldg_stream2(src + (2 * r) * 64 + 2 * i);
          bot[j] = ldg_stream2(src + (2 * r + 1) * 64 + 2 * i);
        }
      }
    };
    auto publish = [&](const int t) {                // registers -> A1[t & 1], then tell the conv1 issuer
      if (t >= T) return;
      const int buf = t & 1;
      const long long q = ((long long)t * p.nclusters + cluster_id) * 2 + rank;
      if (t >= 2) ptx::mbar_wait(bar_a1_free + 8 * buf, (uint32_t)(((t >> 1) - 1) & 1), 7, p.dbg);   // conv1(t-2) has read it
      if (warp == 12) PE_TRACE(14);
      if (q < p.nq) {
        const uint32_t a1 = sm + oA1 + buf * kA1Bytes;
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          const int pos = tid + 128 * j, r = pos >> 5, i = pos & 31;
          const uint32_t w0 = pack_bf16(top[j].x, top[j].y), w1 = pack_bf16(bot[j].x, bot[j].y);   // (py, px) order
          const uint32_t e = a1 + (uint32_t)(1 + (r + 1) * kRow1 + i) * 16u;
          sts64(e, w0, w1);                          // first half of entry (r, i)
          sts64(e - 16u + 8u, w0, w1);               // second half of entry (r, i - 1) (i = 0: the zero column of row r - 1)
        }
      }
      ptx::fence_proxy_async_smem();
      __syncwarp();
      if (lane == 0) ptx::mbar_arrive_cluster(full_tgt0 + 8 * buf);
      if (warp == 12) PE_TRACE(15);
    };
    fetch(0); publish(0); fetch(1);
    for (int t = 0; t < T; ++t) {
      publish(t + 1);
      fetch(t + 2);
      e1_blocks(t, 6, kMB1);
    }
  } else if (warp >= 4) {
    // ================================================================ epilogues (TMEM lane quarter = warp % 4)
    //   set A (warps 4-7):  E1 blocks 0..2 of map t, then E3 of map t-2 (+ the output store)
    //   set B (warps 8-11): E1 blocks 3..5 of map t, then E2 of map t-1
    //   (the loader warps 12-15 take E1 blocks 6, 7)
    const bool set_a = warp < 8;
    const uint32_t e2_tgt = ptx::mapa_shared(bar_e2_done, 0), e3_tgt = ptx::mapa_shared(bar_e3_done, 0);
    const int etid = threadIdx.x - 128;               // set A: 0..127
    for (int t = 0; t < T + 2; ++t) {
      if (t < T) e1_blocks(t, set_a ? 0 : 3, set_a ? 3 : 6);     // (loader warps: blocks 6, 7)
      if (!set_a) {
        // ---------------- E2(t-1): D2[(t-1) & 1] -> conv3's flat s2d buffer
        if (t >= 1 && t <= T) {
          const int u = t - 1, buf = u & 1;
          ptx::mbar_wait(bar_m2_done + 8 * buf, (uint32_t)((u >> 1) & 1), 10, p.dbg);
          if (u >= 1) ptx::mbar_wait(bar_m3_done + 8 * ((u - 1) & 1), (uint32_t)(((u - 1) >> 1) & 1), 11, p.dbg);   // conv3(u-1) has read A3
          ptx::tc_fence_after_sync();
          if (wq == 0) PE_TRACE(10);
#if SB_PE_BIAS12 == 1
          float b2[32];
#pragma unroll
          for (int c = 0; c < 8; ++c) {
            const float4 v = reinterpret_cast<const float4*>(bias)[4 + c];
            b2[4 * c] = v.x; b2[4 * c + 1] = v.y; b2[4 * c + 2] = v.z; b2[4 * c + 3] = v.w;
          }
#else
          const float* b2 = bias + 16;
#endif
#pragma unroll 1
          for (int mb = 0; mb < kMB2; ++mb) {
            uint32_t r[32];
            ptx::tmem_ld_32x32b_x32(lane_t + kD2Col + buf * kD2Cols + mb * 32, r);
            ptx::tmem_ld_wait();
            const int row = wq * 32 + lane;
            const int oy = row >> 3, ox = mb * 8 + (row & 7);
            {
              const uint32_t chunk = (uint32_t)(((oy & 1) * 2 + (ox & 1)) * 4);
              const uint32_t e = sm + oA3 + (chunk * kCh3 + (uint32_t)(1 + ((oy >> 1) + 1) * kRow3 + (ox >> 1))) * 16u;
#pragma unroll
              for (int g = 0; g < 4; ++g) {
                uint32_t w[4];
#if SB_PE_BIAS12 == 2
#pragma unroll
                for (int c2 = 0; c2 < 2; ++c2) {
                  const float4 bv = reinterpret_cast<const float4*>(bias)[4 + 2 * g + c2];
                  w[2 * c2] = pack_bf16(fmaxf(__uint_as_float(r[8 * g + 4 * c2]) + bv.x, 0.0f),
                                        fmaxf(__uint_as_float(r[8 * g + 4 * c2 + 1]) + bv.y, 0.0f));
                  w[2 * c2 + 1] = pack_bf16(fmaxf(__uint_as_float(r[8 * g + 4 * c2 + 2]) + bv.z, 0.0f),
                                            fmaxf(__uint_as_float(r[8 * g + 4 * c2 + 3]) + bv.w, 0.0f));
                }
#else
#pragma unroll
                for (int c = 0; c < 4; ++c)
                  w[c] = pack_bf16(fmaxf(__uint_as_float(r[8 * g + 2 * c]) + b2[8 * g + 2 * c], 0.0f),
                                   fmaxf(__uint_as_float(r[8 * g + 2 * c + 1]) + b2[8 * g + 2 * c + 1], 0.0f));
#endif
                sts128(e + (uint32_t)g * (kCh3 * 16u), w[0], w[1], w[2], w[3]);
              }
            }
          }
          ptx::fence_proxy_async_smem();
          ptx::tc_fence_before_sync();
          __syncwarp();
          if (lane == 0) ptx::mbar_arrive_cluster(e2_tgt + 8 * buf);
          if (wq == 0) PE_TRACE(11);
        }
      } else {
        // ---------------- E3(t-2): D3[(t-2) & 1] + bias -> fp32 [64, 8, 8] through a staged 16 KB bulk store
        if (t >= 2) {
          const int u = t - 2, buf = u & 1;
          const long long q = ((long long)u * p.nclusters + cluster_id) * 2 + rank;
          ptx::mbar_wait(bar_m3_done + 8 * buf, (uint32_t)((u >> 1) & 1), 12, p.dbg);
          ptx::tc_fence_after_sync();
          if (wq == 0) PE_TRACE(12);
          if (etid == 0) ptx::tma_store_wait_read<0>();               // the previous map's store has read the staging
          asm volatile("bar.sync 1, 128;" ::: "memory");
          const int m = wq * 32 + lane;              // rows 0..63 = (oy, ox) row-major; rows 64..127 are not outputs
          const int oy = m >> 3, ox = m & 7;
          const bool valid = m < 64;
          const uint32_t st = sm + oStage + (uint32_t)(oy * 8 + ox) * 4u;
#pragma unroll
          for (int h = 0; h < 2; ++h) {
            uint32_t r[32];
            ptx::tmem_ld_32x32b_x32(lane_t + kD3Col + buf * kD3Cols + h * 32, r);
            ptx::tmem_ld_wait();
            if (valid) {
#if SB_PE_BIAS3
#pragma unroll
              for (int c4 = 0; c4 < 8; ++c4) {
                const float4 bv = reinterpret_cast<const float4*>(bias)[12 + h * 8 + c4];
                sts32f(st + (uint32_t)(h * 32 + 4 * c4 + 0) * 256u, __uint_as_float(r[4 * c4 + 0]) + bv.x);
                sts32f(st + (uint32_t)(h * 32 + 4 * c4 + 1) * 256u, __uint_as_float(r[4 * c4 + 1]) + bv.y);
                sts32f(st + (uint32_t)(h * 32 + 4 * c4 + 2) * 256u, __uint_as_float(r[4 * c4 + 2]) + bv.z);
                sts32f(st + (uint32_t)(h * 32 + 4 * c4 + 3) * 256u, __uint_as_float(r[4 * c4 + 3]) + bv.w);
              }
#else
#pragma unroll
              for (int c = 0; c < 32; ++c)
                sts32f(st + (uint32_t)(h * 32 + c) * 256u, __uint_as_float(r[c]) + bias[48 + h * 32 + c]);
#endif
            }
          }
          ptx::tc_fence_before_sync();
          __syncwarp();
          if (lane == 0) ptx::mbar_arrive_cluster(e3_tgt + 8 * buf);
          if (wq == 0) PE_TRACE(13);
          ptx::fence_proxy_async_smem();
          asm volatile("bar.sync 1, 128;" ::: "memory");
          if (etid == 0 && q < p.nq) {
            bulk_store_1d(p.out + q * 4096, sm + oStage, kStageBytes);
            ptx::tma_store_commit();
          }
        }
      }
    }
    if (set_a && etid == 0) ptx::tma_store_wait_all<0>();
  }

  ptx::tc_fence_before_sync();
  __syncthreads();
#ifdef SB_PE_TRACE
  if (threadIdx.x == 0 && blockIdx.x < 148) {
    g_pe_trace[128 + 2 * blockIdx.x] = clock64() - trace_k0;
    g_pe_trace[128 + 2 * blockIdx.x + 1] = clock64() - trace_loop0;
  }
#endif
  ptx::cluster_sync_all();                            // neither CTA frees TMEM / exits while its peer still uses the pair
  if (warp == 1) {
    ptx::tc_fence_after_sync();
    ptx::tmem_dealloc_2cta(tmem_base, kTmemCols);
  }
}

// ------------------------------------------------------------------ weight packing (once per model)
// fp32 conv weights [16,1,6,6], [32,16,6,6], [64,32,6,6] -> the two per-CTA bf16 images described above.
__global__ void __launch_bounds__(256)
patch_embed_pack_kernel(const float* __restrict__ w1, const float* __restrict__ w2, const float* __restrict__ w3,
                        uint16_t* __restrict__ pack) {
  // one thread per packed bf16 element of both ranks
  const int per_rank = kWBytes / 2;
  for (int idx = blockIdx.x * blockDim.x + threadIdx.x; idx < 2 * per_rank; idx += gridDim.x * blockDim.x) {
    const int rank = idx / per_rank;
    int e = idx - rank * per_rank;                   // element within the rank's image
    float v = 0.0f;
    if (e < kW1Bytes / 2) {
      // [k/8][n%8][k%8], 8 rows; k = ks*16 + kk: kk 0-3 dx=-1, 4-7 dx=0, 8-11 dx=+1, 12-15 zero; (py, px) = (kk>>1)&1, kk&1
      const int kc = e / 64, n = (e % 64) / 8 + rank * 8, k = kc * 8 + (e % 8);
      const int ks = k / 16, kk = k % 16;
      if (kk < 12) {
        const int dxi = kk >> 2, py = (kk >> 1) & 1, px = kk & 1;
        const int ky = 2 * (ks - 1) + py + 2, kx = 2 * (dxi - 1) + px + 2;
        v = w1[(n * 6 + ky) * 6 + kx];
      }
    } else if (e < (kW1Bytes + kW2Bytes) / 2) {
      e -= kW1Bytes / 2;
      // [k/8][n/8][n%8][k%8], 16 rows; k = tap*64 + (py*2+px)*16 + c1
      const int kc = e / 128, nl = (e % 128) / 8, n = nl + rank * 16, k = kc * 8 + (e % 8);
      const int tap = k / 64, s = (k % 64) / 16, c1 = k % 16;
      const int ky = 2 * (tap / 3 - 1) + (s >> 1) + 2, kx = 2 * (tap % 3 - 1) + (s & 1) + 2;
      v = w2[((n * 16 + c1) * 6 + ky) * 6 + kx];
    } else {
      e -= (kW1Bytes + kW2Bytes) / 2;
      // [k/8][n/8][n%8][k%8], 32 rows; k = tap*128 + (py*2+px)*32 + c2
      const int kc = e / 256, nl = (e % 256) / 8, n = nl + rank * 32, k = kc * 8 + (e % 8);
      const int tap = k / 128, s = (k % 128) / 32, c2 = k % 32;
      const int ky = 2 * (tap / 3 - 1) + (s >> 1) + 2, kx = 2 * (tap % 3 - 1) + (s & 1) + 2;
      v = w3[((n * 32 + c2) * 6 + ky) * 6 + kx];
    }
    __nv_bfloat16 h = __float2bfloat16_rn(v);
    pack[idx] = *reinterpret_cast<uint16_t*>(&h);
  }
}

}  // namespace pe
}  // namespace sb

extern "C" size_t sb_patch_embed_pack_bytes(void) { return 2 * (size_t)sb::pe::kWBytes; }

extern "C" int sb_patch_embed_pack(const float* w1, const float* w2, const float* w3, void* pack, sb_stream_t stream) {
  using namespace sb;
  SB_ENTER();
  SB_REQUIRE(w1 && w2 && w3 && pack, SB_EINVAL, "sb_patch_embed_pack: null pointer");
  SB_REQUIRE(aligned16(pack), SB_EINVAL, "sb_patch_embed_pack: pack must be 16-byte aligned");
  pe::patch_embed_pack_kernel<<<kNumSMs, 256, 0, as_stream(stream)>>>(w1, w2, w3, reinterpret_cast<uint16_t*>(pack));
  SB_LAUNCH_CHECK("patch_embed_pack_kernel");
  return SB_OK;
}

extern "C" int sb_patch_embed_proj(const float* cost_maps, const void* pack, const float* bias, float* out,
                                   long long n_maps, int H, int W, sb_stream_t stream) {
  using namespace sb;
  SB_ENTER();
  SB_REQUIRE(n_maps >= 0, SB_EINVAL, "sb_patch_embed_proj: negative map count");
  SB_REQUIRE(H == 64 && W == 64, SB_EUNSUP,
             "sb_patch_embed_proj: cost maps of %d x %d; the kernel is specialised for 64 x 64 maps (512 x 512 images)", H, W);
  if (n_maps == 0) return SB_OK;
  SB_REQUIRE(cost_maps && pack && bias && out, SB_EINVAL, "sb_patch_embed_proj: null pointer");
  SB_REQUIRE(aligned16(cost_maps) && aligned16(pack) && aligned16(out), SB_EINVAL,
             "sb_patch_embed_proj: cost_maps, pack and out must be 16-byte aligned");
  unsigned int* dbg = debug_word_device();
  if (!dbg) return SB_ECUDA;
  static SmemOptIn opt_in;
  int opt_dev;
  if (opt_in.need(pe::kSmemTotal, &opt_dev)) {
    SB_CUDA(cudaFuncSetAttribute(pe::patch_embed_umma_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, pe::kSmemTotal));
    opt_in.done(pe::kSmemTotal, opt_dev);
  }
  pe::Params p;
  p.maps = cost_maps; p.out = out; p.wpack = static_cast<const uint8_t*>(pack); p.bias = bias; p.nq = n_maps; p.dbg = dbg;
  const long long pairs = (n_maps + 1) / 2;
  p.nclusters = (int)(pairs < kNumSMs / 2 ? pairs : kNumSMs / 2);
  p.iters = (int)((pairs + p.nclusters - 1) / p.nclusters);
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3((unsigned)(2 * p.nclusters));
  cfg.blockDim = dim3(pe::kThreads);
  cfg.dynamicSmemBytes = pe::kSmemTotal;
  cfg.stream = as_stream(stream);
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = 2; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr; cfg.numAttrs = 1;
  SB_CUDA(cudaLaunchKernelEx(&cfg, pe::patch_embed_umma_kernel, p));
  SB_LAUNCH_CHECK("patch_embed_umma_kernel");
  return SB_OK;
}

#ifdef SB_PE_TRACE
// debug builds only (not declared in include/stitch_b200.h): copies the event trace of the last launch to the host
extern "C" int sb_pe_acc_read(long long* host_out) {
  if (cudaDeviceSynchronize() != cudaSuccess) return -1;
  return cudaMemcpyFromSymbol(host_out, sb::pe::g_pe_acc, sizeof(long long) * 74 * 8) == cudaSuccess ? 0 : -1;
}
extern "C" int sb_pe_trace_read(long long* host_out) {
  if (cudaDeviceSynchronize() != cudaSuccess) return -1;
  return cudaMemcpyFromSymbol(host_out, sb::pe::g_pe_trace, sizeof(long long) * (8 * 16 + 2 * 148)) == cudaSuccess ? 0 : -1;
}
#endif
