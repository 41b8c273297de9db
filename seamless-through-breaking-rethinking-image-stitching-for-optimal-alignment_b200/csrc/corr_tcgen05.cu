// corr_tcgen05.cu — C1 (+ fused C2): FlowFormer's all-pairs cost volume.
// Replaces MemoryEncoder.corr (core/FlowFormer/PerCostFormer3/encoder.py:359-369):
//   corr[b, i, j] = sum_d fmap1[b, d, i] * fmap2[b, d, j]       heads = 1, no scale
// and builds the avg-pool pyramid over the target axes (C2; F.avg_pool2d(.,2,2)
// chained, the RAFT form hinted at encoder.py:376) in the GEMM epilogue.
//
// Two kernels:
//  (1) feat_to_tokens_bf16: fp32 NCHW [B, C, N] -> bf16 token-major [B, N, Cpad]
//      (K-major operands for the MMA; 12 B/element pre-pass, reusable for the
//      forward and the backward volume of a pair).
//  (2) corr_umma_kernel: persistent, warp-specialised tcgen05 GEMM, one CTA per
//      SM.  Work unit = (batch b, 128-query block, group of 4 consecutive
//      128-column target tiles):
//        warp 0  TMA producer : A (128 x Cpad, resident for the unit) and a
//                               2-stage ring of B tiles (128 x Cpad), SWIZZLE_128B
//        warp 1  MMA issuer   : tcgen05.mma cta_group::1 kind::f16, M=128 N=128 K=16,
//                               fp32 accumulators in TMEM, 4 x 128-column buffers
//        warp 2  TMEM alloc / dealloc (512 columns)
//        warps 4-7 epilogue   : tcgen05.ld 32x32b.x32 -> registers -> (2x2 / 4x4 /
//                               8x8 pooling in registers) -> swizzled smem -> TMA store
//      K = C <= 256 is only 16 MMA k-steps per tile, so the kernel is bounded by
//      the fp32 volume WRITE (64 KB per tile): roofline = HBM, not tensor pipe
//      (SURVEY.md D4).  TMA stores write full 128-byte lines.
//
// With W2 == 64 a 128-column tile is exactly two target rows, so each epilogue
// thread (one query row) holds everything the 2x2 pool needs; four consecutive
// tiles give the 4x4 and 8x8 levels.
#include <cuda.h>   // CUtensorMap (types only; the encoder is fetched at run time)
#include <cuda_bf16.h>

#include "common.cuh"
#define SB_MBAR_SLOW_NOINLINE 0   // setmaxnreg kernels cannot contain calls (ptxas C7600)
#include "ptx_sm100.cuh"
#include "tmap.cuh"

namespace sb {

// ------------------------------------------------------------- pre-pass
constexpr int kTokTile = 64;

__global__ void __launch_bounds__(256)
feat_to_tokens_bf16_kernel(const float* __restrict__ fmap, uint16_t* __restrict__ tok, int C,
                           int Cpad, int N) {
  // tile: 64 channels x 64 tokens
  __shared__ float s[kTokTile][kTokTile + 1];  // [token][channel]
  const int b = blockIdx.z;
  const int c0 = blockIdx.y * kTokTile, n0 = blockIdx.x * kTokTile;
  const float* src = fmap + (long long)b * C * N;
  const bool vec = ((N & 3) == 0);
  // load: 64 rows (channels) x 16 float4 (tokens)
  for (int t = threadIdx.x; t < kTokTile * (kTokTile / 4); t += blockDim.x) {
    const int cr = t / (kTokTile / 4), nq = (t % (kTokTile / 4)) * 4;
    const int c = c0 + cr, n = n0 + nq;
    float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
    if (c < C) {
      const float* p = src + (long long)c * N + n;
      if (vec && n + 3 < N) v = ldg_stream4(p);
      else {
        if (n < N) v.x = p[0];
        if (n + 1 < N) v.y = p[1];
        if (n + 2 < N) v.z = p[2];
        if (n + 3 < N) v.w = p[3];
      }
    }
    s[nq + 0][cr] = v.x; s[nq + 1][cr] = v.y; s[nq + 2][cr] = v.z; s[nq + 3][cr] = v.w;
  }
  __syncthreads();
  // store: 64 tokens x 8 chunks of 8 channels (16 bytes of bf16)
  uint16_t* dst = tok + (long long)b * N * Cpad;
  for (int t = threadIdx.x; t < kTokTile * 8; t += blockDim.x) {
    const int tr = t >> 3, ch = (t & 7) * 8;
    const int n = n0 + tr;
    if (n >= N) continue;
    uint32_t w[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      // round-to-nearest-even fp32 -> bf16 (pad channels are zeros already)
      __nv_bfloat162 h = __floats2bfloat162_rn(s[tr][ch + 2 * k], s[tr][ch + 2 * k + 1]);
      w[k] = *reinterpret_cast<uint32_t*>(&h);
    }
    *reinterpret_cast<uint4*>(dst + (long long)n * Cpad + c0 + ch) = make_uint4(w[0], w[1], w[2], w[3]);
  }
}

// ------------------------------------------------------------- GEMM kernel
constexpr int BM = 128, BN = 128, BKP = 64;       // tile, K panel (64 bf16 = 128 B)
constexpr int kMaxPanels = 4;                      // C <= 256
constexpr int kPanelBytes = BM * 128;              // 16 KB
#ifndef SB_CORR_BSTAGES
#define SB_CORR_BSTAGES 2
#endif
#ifndef SB_CORR_SBUFS
#define SB_CORR_SBUFS 2
#endif
#ifndef SB_CORR_PPS
#define SB_CORR_PPS 4
#endif
constexpr int kStages = SB_CORR_BSTAGES;
constexpr int kPPS = SB_CORR_PPS;                   // K panels per B stage (4 = a whole target tile at C = 256)
constexpr int kSBufs = SB_CORR_SBUFS;               // staging buffers per epilogue warp
#ifndef SB_CORR_FENCE_GROUP
#define SB_CORR_FENCE_GROUP 2
#endif
constexpr int kFenceGroup = SB_CORR_FENCE_GROUP;    // volume slices staged behind one fence.proxy.async (2 or 4, <= kSBufs)
static_assert((kFenceGroup == 2 || kFenceGroup == 4) && kFenceGroup <= SB_CORR_SBUFS, "fence group");
static_assert(kPPS == 1 || kPPS == 2 || kPPS == 4, "panels per stage");
static_assert(kSBufs == 2 || kSBufs == 4, "staging buffers per warp");
constexpr int kTilesPerUnit = 4;
#ifndef SB_CORR_FENCE_PAIRS
#define SB_CORR_FENCE_PAIRS 1   // the volume slices of a tile are staged two at a time behind ONE fence.proxy.async: the fence
#endif                          // alone was 400 cycles per slice, 35-52 % of the epilogue warps' time (tools/corr_trace.py):
                                // 213.0 -> 206.8 us without, 264.2 -> 260.1 us with the fused pyramid, bit-identical
#ifndef SB_CORR_ROLL_TT
#define SB_CORR_ROLL_TT 1   // the epilogue's loop over the 4 tiles of a unit is NOT unrolled: 76 -> 46 KB of SASS with the fused
#endif                      // pyramid (the unrolled body alone exceeded the 32 KB instruction-cache level): 292.8 -> 289.7 us
#ifdef SB_CORR_TRACE
// Debug build only (tools/build_variants.sh corr_tcgen05 "trace:-DSB_CORR_TRACE", tools/corr_trace.py): per CTA, the
// cycles each role spent in its waits.  [0] MMA issuer: accumulator not drained, [1] B stage not loaded, [2] A block not
// loaded; [3] epilogue warp 4: accumulator not ready, [4] staging buffer still being read by an earlier store,
// [5] its whole loop; [6] producer: no free B stage; [7] smid.
__device__ long long g_corr_acc[148 * 8];
__device__ long long g_corr_acc2[148 * 4];         // epilogue warp 4, volume slices: [0] tcgen05.ld, [1] pooling + st.shared, [2] fence + store issue
#define CT_T0 long long ct_t0 = clock64()
#define CT_ADD(slot) do { const long long ct_c = clock64(); ct_acc[slot] += ct_c - ct_t0; ct_t0 = ct_c; } while (0)
#define CT_MARK ct_t0 = clock64()
#else
#define CT_T0 do { } while (0)
#define CT_ADD(slot) do { } while (0)
#define CT_MARK do { } while (0)
#endif
constexpr int kAccBufs = 4;                        // 4 x 128 TMEM columns
constexpr int kTmemCols = 512;
constexpr int kStageBufBytes = 32 * 128;           // per-warp staging: 32 rows x 32 fp32
constexpr int kSmemA = kMaxPanels * kPanelBytes;                 // 65536
constexpr int kSmemB = kStages * kPPS * kPanelBytes;             // 131072 (2 stages x 4 panels)
constexpr int kSmemStage = 4 * kSBufs * kStageBufBytes;          // 32768 with 2 buffers per warp
constexpr int kSmemBar = 512;
// No alignment slack: the dynamic shared-memory array is declared __align__(1024) (SWIZZLE_128B tiles need it) and
// the kernel traps if the base is not aligned.  224 KB + 256 B leaves room for the 1 KB-per-CTA reservations of two
// more shared-memory-free CTAs on the SM (233 472 B per SM): the warp-stage kernels co-reside with this one.
constexpr int kSmemTotal = kSmemA + kSmemB + kSmemStage + kSmemBar;
// Register budget (SMX != 1): the kernel is registered at kRegsLaunch per thread (256 threads -> 38 912 of the SM's
// 65 536 registers), warps 0-3 (TMA / MMA / TMEM / idle) shrink to kRegsLight and the four epilogue warps grow to
// kRegsEpilogue with setmaxnreg; 4*32*(56 + 248) = 38 912.  The remaining 26 624 registers hold two 256-thread
// CTAs of the instruction-bound gather kernels (32-40 registers per thread) next to this HBM-store-bound one.
constexpr int kRegsLaunch = 152, kRegsLight = 56, kRegsEpilogue = 248;

struct CorrParams {
  int B, N1, N2, KP;            // KP = Cpad / 64
  int MB, NT, NG;               // query blocks, target tiles, tile groups per (b, mblk)
  long long n_units;
  int H2h, H2q, H2e;            // H2/2, H2/4, H2/8 (pooling)
  float* lvl1; float* lvl2; float* lvl3;
  unsigned int* dbg;
  int store_policy;             // L2 policy of the output stores (sb_tune SB_TUNE_CORR_STORE_POLICY)
  int tpu;                      // tiles per unit (kTilesPerUnit; 8 for POOL == 2; NT for the softmax statistics pass)
  float* smx;                   // SMX: per-row (max, sum of exp) [B, N1, 2]; written by pass 1, read by pass 2
  int b_rot;                    // the B operand of batch element b is token map (b + b_rot) mod B (0: the same index)
  int dyn;                      // 1: work units are handed out by the hardware (cluster launch control), grid = n_units
};

// instruction descriptor: D=f32, A=B=bf16, both K-major, N=128, M=128
constexpr uint32_t kIdesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(BN >> 3) << 17) |
                            ((uint32_t)(BM >> 4) << 24);

// instruction descriptor of the CTA-pair MMA: M = 256 across two CTAs
constexpr uint32_t kIdesc2 = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(BN >> 3) << 17) |
                             ((uint32_t)((2 * BM) >> 4) << 24);

// BF16OUT (never with POOL): the volume is written as bf16 [B, N1, N2] — half the bytes, which moves
// the kernel from the HBM-write roofline towards the tensor pipe (the SURVEY D4 option).
// TWO_CTA: launched as clusters of 2 CTAs (one SM pair); a unit is then 256 queries (CTA r owns rows
// [256*mbp + 128*r, +128)) and every target tile is ONE tcgen05.mma.cta_group::2 chain issued by the
// leader CTA: each CTA loads only its half of every B tile (64 target rows, 32 KB instead of 64 KB),
// all TMA loads signal the leader's barriers, tcgen05.commit multicasts the "stage free" /
// "accumulator ready" arrivals to both CTAs, both epilogues drain their own TMEM and report back to
// the leader's "accumulator free" barrier through shared::cluster.
// POOL: 0 = volume only; 1 = fused pyramid for W2 == 64 (a 128-column tile is two target rows);
// 2 = fused pyramid for W2 == 128 (a tile is ONE target row: the epilogue drains tile PAIRS, reading
// both rows' accumulators from TMEM slice by slice, so level 1 needs no cross-tile carry; units are
// 8 tiles = 8 target rows so that level 3 closes inside the unit).
__device__ __forceinline__ float ex2_approx(float x) {   // 2^x, 2 ulp; -inf -> 0
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

// SMX (never with POOL / BF16OUT / TWO_CTA): row softmax of the volume fused as two passes over the SAME
// contraction instead of a round trip of the fp32 logits through HBM (GMA Attention.forward):
//   1 = statistics: a unit is a whole row block (all NT tiles); every epilogue thread keeps the running
//       maximum and sum of exp of its query row and writes (max, sum) at the end — nothing else is stored;
//       launched with TWO epilogue warp quads (384 threads) that take alternate tiles, because one warp per
//       SM sub-partition cannot hide the TMEM-load and ex2 latencies on its own; their partial (max, sum)
//       are merged through shared memory;
//   2 = normalise: the contraction is recomputed and every value leaves as
//       tf32(exp(x - max) / sum) through the usual staged TMA stores.
// ATMEM (POOL 0 / 1, one CTA per tile, no SMX): the unit's A block (128 queries x C channels) is copied from
// shared memory into tensor memory ONCE per unit (tcgen05.cp, 8 columns per K = 16 step, columns 384..511;
// three accumulator buffers instead of four) and every MMA reads A from there: the shared-memory pipe, the
// busiest unit of this kernel (ncu: 75-78 %), loses 48 of its ~320 KB per tile.
template <int POOL, bool BF16OUT = false, bool TWO_CTA = false, int SMX = 0, bool ATMEM = false>
__global__ void __maxnreg__(SMX == 1 ? 168 : kRegsLaunch)
corr_umma_kernel(const __grid_constant__ CUtensorMap map_a, const __grid_constant__ CUtensorMap map_b,
                 const __grid_constant__ CUtensorMap map_v, const __grid_constant__ CUtensorMap map_l1,
                 const __grid_constant__ CUtensorMap map_l2, const CorrParams p) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  const uint32_t smem_base = ptx::smem_u32(smem_raw);
  if (smem_base & 1023u) {                       // never observed; a misaligned base would corrupt swizzled tiles
    if (threadIdx.x == 0 && p.dbg) atomicExch(p.dbg, 0xDEAD00A1u);
    __threadfence_system();
    __trap();
  }
  const uint32_t sA = smem_base;
  const uint32_t sB = sA + kSmemA;
  const uint32_t sStage = sB + kSmemB;
  const uint32_t sBar = sStage + kSmemStage;
  // barrier map (8 B each)
  const uint32_t bar_a_full = sBar + 0, bar_a_empty = sBar + 8;
  const uint32_t bar_b_full = sBar + 256;   // [<= 8 stages]
  const uint32_t bar_b_empty = sBar + 320;  // [<= 8 stages]
  const uint32_t bar_t_full = sBar + 80;    // [kAccBufs]
  const uint32_t bar_t_empty = sBar + 112;  // [kAccBufs]
  const uint32_t tmem_slot = sBar + 144;    // u32
  // dynamic unit scheduling (p.dyn): two response slots of clusterlaunchcontrol.try_cancel, each with a "response has
  // landed" barrier (transaction bytes) and an "every reader has decoded it" barrier
  const uint32_t bar_s_full = sBar + 152;   // [2]
  const uint32_t bar_s_empty = sBar + 168;  // [2]
  const uint32_t sResp = sBar + 192;        // [2] x 16 B
  // CTA-pair mode loads half-size B tiles: the same 128 KB ring holds twice as many stages
  constexpr int kStages = TWO_CTA ? 2 * sb::kStages : sb::kStages;
  static_assert(kStages <= 8, "barrier map holds 8 B stages");
  uint8_t* smem_gen = smem_raw + (smem_base - ptx::smem_u32(smem_raw));
  volatile uint32_t* tmem_slot_ptr = reinterpret_cast<volatile uint32_t*>(smem_gen + kSmemA + kSmemB + kSmemStage + 144);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t cta_rank = TWO_CTA ? ptx::cluster_ctarank() : 0u;
  const bool leader = cta_rank == 0;
  // work is enumerated per cluster in TWO_CTA mode
  const long long unit0 = TWO_CTA ? (long long)(blockIdx.x >> 1) : (long long)blockIdx.x;
  const long long unit_step = TWO_CTA ? (long long)(gridDim.x >> 1) : (long long)gridDim.x;
  // ---- work units.  Static: unit0, unit0 + unit_step, ...  Dynamic (p.dyn, one CTA per tile only): the grid has one
  // CTA per unit; a CTA works on its own blockIdx first and then asks the hardware to cancel the launch of a CTA that
  // has not started and takes that CTA's index (clusterlaunchcontrol.try_cancel) until none is left -- the SMs do not
  // run this store-bound kernel at the same speed (tools/corr_trace.py: the slowest ten are 12 % behind the median
  // with the fused pyramid), and with a static split the whole grid waits for them.  The producer thread issues the
  // request for unit k + 1 while unit k is loaded; the 16-byte response lands in shared memory (slot k & 1) and every
  // role decodes it for itself.
  const bool dyn = !TWO_CTA && p.dyn != 0;
  auto sched_decode = [&](int m, long long& u) -> bool {          // response m -> unit of iteration m + 1
    ptx::mbar_wait(bar_s_full + 8 * (m & 1), (uint32_t)((m >> 1) & 1), 20, p.dbg);
    uint32_t x, y, z, valid;
    asm volatile(
        "{\n\t.reg .pred p1;\n\t.reg .b128 resp;\n\t"
        "ld.shared.b128 resp, [%4];\n\t"
        "clusterlaunchcontrol.query_cancel.is_canceled.pred.b128 p1, resp;\n\t"
        "selp.u32 %3, 1, 0, p1;\n\t"
        "mov.u32 %0, 0; mov.u32 %1, 0; mov.u32 %2, 0;\n\t"
        "@p1 clusterlaunchcontrol.query_cancel.get_first_ctaid.v4.b32.b128 {%0, %1, %2, _}, resp;\n\t}"
        : "=r"(x), "=r"(y), "=r"(z), "=r"(valid)
        : "r"(sResp + 16u * (uint32_t)(m & 1))
        : "memory");
    (void)y; (void)z;
    u = (long long)x;
    return valid != 0;
  };
  // consumers (MMA issuer thread, epilogue warps): advance to the next unit; false when there is none
  auto next_unit = [&](long long& u, int& m, bool one_thread) -> bool {
    if (!dyn) { u += unit_step; return u < p.n_units; }
    const bool ok = sched_decode(m, u);
    ptx::fence_proxy_async_smem();                 // the slot is rewritten through the async proxy
    if (!one_thread) __syncwarp();
    if (one_thread || lane == 0) ptx::mbar_arrive(bar_s_empty + 8 * (m & 1));
    ++m;
    return ok;
  };
  constexpr int kBRows = TWO_CTA ? BN / 2 : BN;                   // B rows this CTA loads per tile
  constexpr int kNAcc = ATMEM ? 3 : kAccBufs;                     // accumulator buffers in use
  static_assert(!ATMEM || (POOL != 2 && !TWO_CTA && SMX == 0), "ATMEM: one CTA per tile, POOL 0 / 1 only");
  constexpr int kBPanelBytes = kBRows * 128;

  if (warp == 0 && lane == 0) {
    ptx::prefetch_tensormap(&map_a);
    ptx::prefetch_tensormap(&map_b);
    ptx::prefetch_tensormap(&map_v);
    if (POOL) { ptx::prefetch_tensormap(&map_l1); ptx::prefetch_tensormap(&map_l2); }
  }
  if (warp == 1 && lane == 0) {
    ptx::mbar_init(bar_a_full, 1);
    ptx::mbar_init(bar_a_empty, 1);
    for (int s = 0; s < kStages; ++s) {
      ptx::mbar_init(bar_b_full + 8 * s, 1);
      ptx::mbar_init(bar_b_empty + 8 * s, 1);
    }
    for (int i = 0; i < 2; ++i) {
      ptx::mbar_init(bar_s_full + 8 * i, 1);
      ptx::mbar_init(bar_s_empty + 8 * i, 1 + (blockDim.x >> 5) - 4);   // the MMA issuer + every epilogue warp
    }
    for (int a = 0; a < kAccBufs; ++a) {
      ptx::mbar_init(bar_t_full + 8 * a, 1);
      ptx::mbar_init(bar_t_empty + 8 * a, TWO_CTA ? 8 : 4);   // one elected lane per epilogue warp (of both CTAs)
    }
    ptx::fence_mbar_init();
  }
  if (warp == 2) {
    if (TWO_CTA) { ptx::tmem_alloc_2cta(tmem_slot, kTmemCols); ptx::tmem_relinquish_2cta(); }
    else { ptx::tmem_alloc(tmem_slot, kTmemCols); ptx::tmem_relinquish(); }
  }
  ptx::tc_fence_before_sync();
  __syncthreads();
  if (TWO_CTA) ptx::cluster_sync_all();     // the peer's barriers are initialised before anything remote arrives
  ptx::tc_fence_after_sync();
  const uint32_t tmem_base = *tmem_slot_ptr;

  const uint32_t panel_tx = (uint32_t)p.KP * kPanelBytes;                 // A: this CTA's 128 rows

  if (warp < 4) {
  // warpgroup 0 (one elected lane each for TMA and MMA issue): hand registers to the epilogue warpgroup
  if (SMX != 1) asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;\n" ::"n"(kRegsLight));
  if (warp == 0) {
    // ================================================================ producer
    if (lane == 0) {
      uint32_t a_par = 0, stage = 0, b_par = 0;
#ifdef SB_CORR_TRACE
      long long ct_acc[1] = {0};
#endif
      CT_T0;
      // experiment knob (store_policy bit 2): operand loads keep their lines in L2 (evict_last)
      const uint64_t ld_policy = (p.store_policy & 4) ? ptx::l2_policy_evict_last() : 0ull;
      // TWO_CTA: completions of both CTAs' loads are counted on the LEADER's "full" barriers
      const uint32_t a_full_tgt = TWO_CTA ? ptx::mapa_shared(bar_a_full, 0) : bar_a_full;
      int sm_i = 0;                                 // index of the next scheduler response
      bool more = unit0 < p.n_units;
      for (long long u = unit0; more;) {
        if (dyn) {
          // ask for the unit after this one: slot sm_i & 1 must have been decoded by every reader of response sm_i - 2
          if (sm_i >= 2) ptx::mbar_wait(bar_s_empty + 8 * (sm_i & 1), (uint32_t)(((sm_i >> 1) - 1) & 1), 21, p.dbg);
          ptx::mbar_arrive_expect_tx(bar_s_full + 8 * (sm_i & 1), 16);
          asm volatile("clusterlaunchcontrol.try_cancel.async.shared::cta.mbarrier::complete_tx::bytes.b128 [%0], [%1];"
                       ::"r"(sResp + 16u * (uint32_t)(sm_i & 1)), "r"(bar_s_full + 8 * (sm_i & 1)) : "memory");
        }
        const int ng = (int)(u % p.NG);
        const long long r1 = u / p.NG;
        const int mb = (int)(r1 % p.MB) * (TWO_CTA ? 2 : 1) + (int)cta_rank;
        const int b = (int)(r1 / p.MB);
        ptx::mbar_wait(bar_a_empty, a_par ^ 1, 1, p.dbg);
        if (!TWO_CTA) ptx::mbar_arrive_expect_tx(bar_a_full, panel_tx);
        else if (leader) ptx::mbar_arrive_expect_tx(bar_a_full, 2 * panel_tx);
        for (int kp = 0; kp < p.KP; ++kp) {
          if (TWO_CTA) ptx::tma_load_3d_2cta(sA + kp * kPanelBytes, &map_a, a_full_tgt, kp * BKP, mb * BM, b);
          else if (ld_policy) ptx::tma_load_3d_hint(sA + kp * kPanelBytes, &map_a, bar_a_full, kp * BKP, mb * BM, b, ld_policy);
          else ptx::tma_load_3d(sA + kp * kPanelBytes, &map_a, bar_a_full, kp * BKP, mb * BM, b);
        }
        a_par ^= 1;
        const int t0 = ng * p.tpu;
        const int t1 = min(t0 + p.tpu, p.NT);
        const int bb = (b + p.b_rot >= p.B) ? b + p.b_rot - p.B : b + p.b_rot;     // batch element of the B operand
        for (int t = t0; t < t1; ++t) {
          for (int h0 = 0; h0 < p.KP; h0 += kPPS) {          // a stage holds kPPS K panels of the tile
            const int np = min(kPPS, p.KP - h0);
            CT_MARK;
            ptx::mbar_wait(bar_b_empty + 8 * stage, b_par ^ 1, 2, p.dbg);
            CT_ADD(0);
            const uint32_t tx = (uint32_t)np * kBPanelBytes;
            if (!TWO_CTA) ptx::mbar_arrive_expect_tx(bar_b_full + 8 * stage, tx);
            else if (leader) ptx::mbar_arrive_expect_tx(bar_b_full + 8 * stage, 2 * tx);
            const uint32_t b_full_tgt = TWO_CTA ? ptx::mapa_shared(bar_b_full + 8 * stage, 0) : bar_b_full + 8 * stage;
            for (int kp = h0; kp < h0 + np; ++kp) {
              const uint32_t dstb = sB + (stage * kPPS + (kp - h0)) * kBPanelBytes;
              if (TWO_CTA) ptx::tma_load_3d_2cta(dstb, &map_b, b_full_tgt, kp * BKP, t * BN + (int)cta_rank * kBRows, bb);
              else if (ld_policy) ptx::tma_load_3d_hint(dstb, &map_b, bar_b_full + 8 * stage, kp * BKP, t * BN, bb, ld_policy);
              else ptx::tma_load_3d(dstb, &map_b, bar_b_full + 8 * stage, kp * BKP, t * BN, bb);
            }
            if (++stage == kStages) { stage = 0; b_par ^= 1; }
          }
        }
        if (dyn) { more = sched_decode(sm_i, u); ++sm_i; }
        else { u += unit_step; more = u < p.n_units; }
      }
#ifdef SB_CORR_TRACE
      if (blockIdx.x < 148) {
        g_corr_acc[blockIdx.x * 8 + 6] = ct_acc[0];
        unsigned smid; asm volatile("mov.u32 %0, %%smid;" : "=r"(smid)); g_corr_acc[blockIdx.x * 8 + 7] = smid;
      }
#endif
    }
    __syncwarp();
  } else if (warp == 1) {
    // ============================================================== MMA issuer
    if (lane == 0 && leader) {
      uint32_t a_par = 0, stage = 0, b_par = 0, acc = 0, acc_par = 0;
#ifdef SB_CORR_TRACE
      long long ct_acc[3] = {0, 0, 0};
#endif
      CT_T0;
      int sm_i = 0;
      bool more = unit0 < p.n_units;
      for (long long u = unit0; more; more = next_unit(u, sm_i, true)) {
        const int ng = (int)(u % p.NG);
        const int t0 = ng * p.tpu;
        const int t1 = min(t0 + p.tpu, p.NT);
        CT_MARK;
        ptx::mbar_wait(bar_a_full, a_par, 3, p.dbg);
        CT_ADD(2);
        a_par ^= 1;
        const uint32_t a_tmem = tmem_base + 3 * BN;      // ATMEM: columns 384..511
        if (ATMEM) {
          ptx::tc_fence_after_sync();
          for (int kp = 0; kp < p.KP; ++kp) {
            const uint64_t adesc = ptx::umma_desc_k_sw128(sA + kp * kPanelBytes);
#pragma unroll
            for (int k = 0; k < BKP / 16; ++k)
              ptx::tmem_cp_128x256b(a_tmem + (uint32_t)(kp * (BKP / 16) + k) * 8u, adesc + 2 * k);
          }
          ptx::umma_commit(bar_a_empty);                 // the A block in shared memory is free once the copies retire
        }
        for (int t = t0; t < t1; ++t) {
          CT_MARK;
          ptx::mbar_wait(bar_t_empty + 8 * acc, acc_par ^ 1, 4, p.dbg);
          CT_ADD(0);
          const uint32_t d_tmem = tmem_base + acc * BN;
          for (int h0 = 0; h0 < p.KP; h0 += kPPS) {
            const int np = min(kPPS, p.KP - h0);
            CT_MARK;
            ptx::mbar_wait(bar_b_full + 8 * stage, b_par, 5, p.dbg);
            CT_ADD(1);
            ptx::tc_fence_after_sync();
            for (int kp = h0; kp < h0 + np; ++kp) {
              const uint64_t adesc = ptx::umma_desc_k_sw128(sA + kp * kPanelBytes);
              const uint64_t bdesc =
                  ptx::umma_desc_k_sw128(sB + (stage * kPPS + (kp - h0)) * kBPanelBytes);
#pragma unroll
              for (int k = 0; k < BKP / 16; ++k) {
                // advance 16 bf16 = 32 bytes along K inside the swizzle atom: +2 in (addr >> 4)
                if (TWO_CTA) ptx::umma_f16_2cta(d_tmem, adesc + 2 * k, bdesc + 2 * k, kIdesc2, (kp | k) != 0);
                else if (ATMEM) ptx::umma_f16_ts(d_tmem, a_tmem + (uint32_t)(kp * (BKP / 16) + k) * 8u, bdesc + 2 * k, kIdesc, (kp | k) != 0);
                else ptx::umma_f16(d_tmem, adesc + 2 * k, bdesc + 2 * k, kIdesc, (kp | k) != 0);
              }
            }
            if (TWO_CTA) ptx::umma_commit_2cta(bar_b_empty + 8 * stage, 3);   // both CTAs' B stages reusable
            else ptx::umma_commit(bar_b_empty + 8 * stage);                    // B stage reusable once these MMAs retire
            if (++stage == kStages) { stage = 0; b_par ^= 1; }
          }
          if (TWO_CTA) ptx::umma_commit_2cta(bar_t_full + 8 * acc, 3);        // both CTAs' accumulators ready
          else ptx::umma_commit(bar_t_full + 8 * acc);                         // accumulator ready for the epilogue
          if (++acc == kNAcc) { acc = 0; acc_par ^= 1; }
        }
        if (TWO_CTA) ptx::umma_commit_2cta(bar_a_empty, 3);
        else if (!ATMEM) ptx::umma_commit(bar_a_empty);  // A reusable once the unit's MMAs retire
      }
#ifdef SB_CORR_TRACE
      if (blockIdx.x < 148) for (int i = 0; i < 3; ++i) g_corr_acc[blockIdx.x * 8 + i] = ct_acc[i];
#endif
    }
    __syncwarp();
  }
  } else {
    if (SMX != 1) asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;\n" ::"n"(kRegsEpilogue));
    // ================================================================ epilogue
    const int wq = (warp - 4) & 3;                 // TMEM lane quarter == warp % 4
    const int quad = (warp - 4) >> 2;              // second epilogue quad: SMX == 1 only
    const uint32_t my_stage = sStage + wq * kSBufs * kStageBufBytes;
    const uint32_t lane_taddr = tmem_base + ((uint32_t)(wq * 32) << 16);
    uint32_t acc = 0, acc_par = 0, sbuf = 0;
    // experiment knob: L2 policy of the output stores (0 none, 1 evict_first, 2 evict_last)
    const uint64_t st_policy = (p.store_policy & 3) == 1 ? ptx::l2_policy_evict_first()
                               : ((p.store_policy & 3) == 2 ? ptx::l2_policy_evict_last() : 0ull);
#define TMA_STORE_V(map_, smem_, c0_, c1_, c2_)                                             \
  do {                                                                                      \
    if (p.store_policy & 3) ptx::tma_store_3d_hint(map_, smem_, c0_, c1_, c2_, st_policy);  \
    else ptx::tma_store_3d(map_, smem_, c0_, c1_, c2_);                                     \
  } while (0)
    // pooling state (one query row per thread)
    float h1[32];   // level-1 partial sums of the current tile (target row pair)
    float h2[16];   // level-2 partial sums across tile pairs
    float h2a[16];  // level-2 row of the unit's first tile pair
    float h1a[32];  // level-1 row of the even tile of a pair
    float h3[8];    // level-3 partial sums across the unit
#ifdef SB_CORR_TRACE
    long long ct_acc[6] = {0, 0, 0, 0, 0, 0};      // + [3] tcgen05.ld round trips, [4] pooling + st.shared, [5] fence + store issue
    const long long ct_loop0 = clock64();
#endif
    CT_T0;
    int sm_i = 0;
    bool more = unit0 < p.n_units;
    for (long long u = unit0; more; more = next_unit(u, sm_i, false)) {
      const int ng = (int)(u % p.NG);
      const long long r1 = u / p.NG;
      const int mb = (int)(r1 % p.MB) * (TWO_CTA ? 2 : 1) + (int)cta_rank;
      const int b = (int)(r1 / p.MB);
      const int t0 = ng * p.tpu;
      const int t1 = min(t0 + p.tpu, p.NT);
      const int row = mb * BM + wq * 32 + lane;    // query index within the batch
      const bool row_ok = row < p.N1;
      const long long q = (long long)b * p.N1 + row;
      constexpr float kLog2e = 1.4426950408889634f;
      if (SMX == 1) {
        // ------------------------------------------------ softmax statistics of this thread's query row
        float m = -INFINITY, l = 0.0f;
        for (int t = t0; t < t1; ++t) {
          if ((int)(acc & 1u) != quad) {               // the other quad's tile
            if (++acc == kAccBufs) { acc = 0; acc_par ^= 1; }
            continue;
          }
          ptx::mbar_wait(bar_t_full + 8 * acc, acc_par, 6, p.dbg);
          ptx::tc_fence_after_sync();
#pragma unroll
          for (int sl = 0; sl < 4; ++sl) {
            uint32_t r[32];
            ptx::tmem_ld_32x32b_x32(lane_taddr + acc * BN + sl * 32, r);
            ptx::tmem_ld_wait();
            const int cbase = t * BN + sl * 32;
            if (cbase >= p.N2) continue;               // columns past the last key: zero-filled operands, not logits
            // four independent chains for the maximum and the sum: one warp per SM sub-partition drains
            // the tile, so a 32-long dependent chain would be pure latency
            float x[32];
            if (cbase + 32 <= p.N2) {
#pragma unroll
              for (int j = 0; j < 32; ++j) x[j] = __uint_as_float(r[j]);
            } else {
#pragma unroll
              for (int j = 0; j < 32; ++j) x[j] = (cbase + j < p.N2) ? __uint_as_float(r[j]) : -INFINITY;
            }
            float c4[4] = {x[0], x[1], x[2], x[3]};
#pragma unroll
            for (int j = 4; j < 32; ++j) c4[j & 3] = fmaxf(c4[j & 3], x[j]);
            const float cm = fmaxf(fmaxf(c4[0], c4[1]), fmaxf(c4[2], c4[3]));
            if (cm > m) {                              // exp2(-inf) = 0 covers the first slice
              l *= ex2_approx((m - cm) * kLog2e);
              m = cm;
            }
            const float mb2 = m * kLog2e;
            float s4[4] = {0.0f, 0.0f, 0.0f, 0.0f};
#pragma unroll
            for (int j = 0; j < 32; ++j) s4[j & 3] += ex2_approx(__fmaf_rn(x[j], kLog2e, -mb2));
            const float sacc = (s4[0] + s4[1]) + (s4[2] + s4[3]);
            l += sacc;
          }
          ptx::tc_fence_before_sync();
          __syncwarp();
          if (lane == 0) ptx::mbar_arrive(bar_t_empty + 8 * acc);
          if (++acc == kAccBufs) { acc = 0; acc_par ^= 1; }
        }
        // merge the two quads' partial statistics of the same 128 rows (staging area is unused in this mode)
        float2* s_ml = reinterpret_cast<float2*>(smem_gen + kSmemA + kSmemB) + wq * 32 + lane;
        if (quad == 1) *s_ml = make_float2(m, l);
        asm volatile("bar.sync 1, 256;" ::: "memory");
        if (quad == 0) {
          const float2 o = *s_ml;                      // (-inf, 0) when the other quad had no tile
          const float mm = fmaxf(m, o.x);
          l = l * ex2_approx((m - mm) * kLog2e) + o.y * ex2_approx((o.x - mm) * kLog2e);
          if (row_ok) *reinterpret_cast<float2*>(p.smx + 2 * q) = make_float2(mm, l);
        }
        asm volatile("bar.sync 1, 256;" ::: "memory");   // s_ml is free for the next unit
        continue;
      }
      float smx_mb2 = 0.0f;                       // max * log2(e) + log2(sum): exp(x - max) / sum = 2^(x log2(e) - smx_mb2)
      if (SMX == 2 && row_ok) {
        const float2 ml = *reinterpret_cast<const float2*>(p.smx + 2 * q);
        smx_mb2 = __fmaf_rn(ml.x, kLog2e, log2f(ml.y));
      }
      if (POOL == 2) {
        // ---------------------------------------------------------- W2 == 128: tile pairs
        // staged 32 x 128-byte block -> one TMA store (volume slice, level-1 half row, level-2 row)
        auto stage_store = [&](const CUtensorMap* map, const uint32_t* w, int c0, int r0) {
          if (lane == 0) ptx::tma_store_wait_read<kSBufs - 1>();
          __syncwarp();
          const uint32_t dst = my_stage + sbuf * kStageBufBytes + lane * 128;
#pragma unroll
          for (int c = 0; c < 8; ++c) {
            const uint32_t a = dst + (((uint32_t)c ^ ((uint32_t)lane & 7u)) << 4);
            asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(a), "r"(w[4 * c]), "r"(w[4 * c + 1]),
                         "r"(w[4 * c + 2]), "r"(w[4 * c + 3])
                         : "memory");
          }
          ptx::fence_proxy_async_smem();
          __syncwarp();
          if (lane == 0) {
            TMA_STORE_V(map, my_stage + sbuf * kStageBufBytes, c0, r0, b);
            ptx::tma_store_commit();
          }
          if (++sbuf == kSBufs) sbuf = 0;
        };
        const int r0 = mb * BM + wq * 32;
        float g2[32], g2a[32], g3[16];               // level-2 partial / first row of the unit, level-3 partial
#if SB_CORR_ROLL_TT
#pragma unroll 1
#else
#pragma unroll
#endif
        for (int pp = 0; pp < 4; ++pp) {             // pair pp = target rows t0 + 2pp, t0 + 2pp + 1
          const int t = t0 + 2 * pp;
          if (t >= t1) break;
          const uint32_t accA = acc, parA = acc_par;
          if (++acc == kAccBufs) { acc = 0; acc_par ^= 1; }
          const uint32_t accB = acc, parB = acc_par;
          if (++acc == kAccBufs) { acc = 0; acc_par ^= 1; }
          ptx::mbar_wait(bar_t_full + 8 * accA, parA, 6, p.dbg);
          ptx::mbar_wait(bar_t_full + 8 * accB, parB, 7, p.dbg);
          ptx::tc_fence_after_sync();
          uint32_t l1w[32];                          // level-1 values of two slices = 32 floats = 128 bytes
#pragma unroll
          for (int sl = 0; sl < 4; ++sl) {
            uint32_t ra[32], rb[32];
            ptx::tmem_ld_32x32b_x32(lane_taddr + accA * BN + sl * 32, ra);
            ptx::tmem_ld_32x32b_x32(lane_taddr + accB * BN + sl * 32, rb);
            ptx::tmem_ld_wait();
            // avg_pool2d order: ((a00 + a01) + a10) + a11, then * 0.25
#pragma unroll
            for (int w = 0; w < 16; ++w)
              l1w[(sl & 1) * 16 + w] = __float_as_uint(fmul(
                  fadd(fadd(fadd(__uint_as_float(ra[2 * w]), __uint_as_float(ra[2 * w + 1])), __uint_as_float(rb[2 * w])),
                       __uint_as_float(rb[2 * w + 1])), 0.25f));
            stage_store(&map_v, ra, t * BN + sl * 32, r0);
            stage_store(&map_v, rb, (t + 1) * BN + sl * 32, r0);
            if (sl & 1) {
              // level-1 row (t0/2 + pp) of every query: 64 floats, this half = columns [32*(sl>>1), +32)
              if (p.lvl1) stage_store(&map_l1, l1w, ((t >> 1)) * 64 + (sl >> 1) * 32, r0);
              // level 2: horizontal pairs of level 1 now, vertical pair with the next level-1 row later
#pragma unroll
              for (int w = 0; w < 16; ++w) {
                const float hsum = fadd(__uint_as_float(l1w[2 * w]), __uint_as_float(l1w[2 * w + 1]));
                const int i2 = (sl >> 1) * 16 + w;
                if ((pp & 1) == 0) g2[i2] = hsum;
                else g2[i2] = fmul(fadd(fadd(g2[i2], __uint_as_float(l1w[2 * w])), __uint_as_float(l1w[2 * w + 1])), 0.25f);
              }
            }
          }
          // both accumulators drained
          ptx::tc_fence_before_sync();
          __syncwarp();
          if (lane == 0) {
            if (TWO_CTA) {
              ptx::mbar_arrive_cluster(ptx::mapa_shared(bar_t_empty + 8 * accA, 0));
              ptx::mbar_arrive_cluster(ptx::mapa_shared(bar_t_empty + 8 * accB, 0));
            } else {
              ptx::mbar_arrive(bar_t_empty + 8 * accA);
              ptx::mbar_arrive(bar_t_empty + 8 * accB);
            }
          }
          if (pp & 1) {
            // a level-2 row (32 floats = one 128-byte line per query) is complete
            if (p.lvl2) {
              uint32_t w2[32];
#pragma unroll
              for (int w = 0; w < 32; ++w) w2[w] = __float_as_uint(g2[w]);
              stage_store(&map_l2, w2, (t >> 2) * 32, r0);
            }
            if (pp == 1) {
#pragma unroll
              for (int w = 0; w < 16; ++w) g3[w] = fadd(g2[2 * w], g2[2 * w + 1]);
            } else {
#pragma unroll
              for (int w = 0; w < 16; ++w) g3[w] = fmul(fadd(fadd(g3[w], g2[2 * w]), g2[2 * w + 1]), 0.25f);
              if (p.lvl3 && row_ok) {
                float* o = p.lvl3 + (q * p.H2e + (t >> 3)) * 16;
#pragma unroll
                for (int c = 0; c < 4; ++c)
                  stg_stream4(o + 4 * c, make_float4(g3[4 * c], g3[4 * c + 1], g3[4 * c + 2], g3[4 * c + 3]));
              }
            }
          }
        }
        (void)g2a;
        continue;
      }
      // a unit is p.tpu tiles (a multiple of 4 when pooling): the pooling state closes every 4 tiles, the A block stays
      for (int tb = t0; tb < t1; tb += kTilesPerUnit) {
#if SB_CORR_ROLL_TT
#pragma unroll 1
#else
#pragma unroll
#endif
      for (int tt = 0; tt < kTilesPerUnit; ++tt) {
        const int t = tb + tt;
        if (t >= t1) break;
        CT_MARK;
        ptx::mbar_wait(bar_t_full + 8 * acc, acc_par, 6, p.dbg);
        CT_ADD(0);
        ptx::tc_fence_after_sync();
        if (BF16OUT) {
#pragma unroll
          for (int s2 = 0; s2 < 2; ++s2) {          // 64 columns = one 128-byte row of bf16 per query
            uint32_t r0[32], r1[32], w[32];
            ptx::tmem_ld_32x32b_x32(lane_taddr + acc * BN + s2 * 64, r0);
            ptx::tmem_ld_32x32b_x32(lane_taddr + acc * BN + s2 * 64 + 32, r1);
            ptx::tmem_ld_wait();
#pragma unroll
            for (int j = 0; j < 16; ++j) {
              __nv_bfloat162 a = __floats2bfloat162_rn(__uint_as_float(r0[2 * j]), __uint_as_float(r0[2 * j + 1]));
              __nv_bfloat162 c = __floats2bfloat162_rn(__uint_as_float(r1[2 * j]), __uint_as_float(r1[2 * j + 1]));
              w[j] = *reinterpret_cast<uint32_t*>(&a);
              w[16 + j] = *reinterpret_cast<uint32_t*>(&c);
            }
            if (lane == 0) ptx::tma_store_wait_read<kSBufs - 1>();
            __syncwarp();
            const uint32_t dst = my_stage + sbuf * kStageBufBytes + lane * 128;
#pragma unroll
            for (int c = 0; c < 8; ++c) {
              const uint32_t a = dst + (((uint32_t)c ^ ((uint32_t)lane & 7u)) << 4);
              asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(a), "r"(w[4 * c]),
                           "r"(w[4 * c + 1]), "r"(w[4 * c + 2]), "r"(w[4 * c + 3])
                           : "memory");
            }
            ptx::fence_proxy_async_smem();
            __syncwarp();
            if (lane == 0) {
              TMA_STORE_V(&map_v, my_stage + sbuf * kStageBufBytes, t * BN + s2 * 64, mb * BM + wq * 32, b);
              ptx::tma_store_commit();
            }
            if (++sbuf == kSBufs) sbuf = 0;
          }
        } else {
#pragma unroll
        for (int sl = 0; sl < 4; ++sl) {
          uint32_t r[32];
          CT_MARK;
          ptx::tmem_ld_32x32b_x32(lane_taddr + acc * BN + sl * 32, r);
          ptx::tmem_ld_wait();
          CT_ADD(3);
          if (SMX == 2) {
#pragma unroll
            for (int j = 0; j < 32; ++j) {
              const float pr = ex2_approx(__fmaf_rn(__uint_as_float(r[j]), kLog2e, -smx_mb2));
              // cvt.rna.tf32 (round to nearest, ties away) of a non-negative finite value with integer
              // ops: the conversion instruction shares the 16-lane unit with ex2
              r[j] = (__float_as_uint(pr) + 0x1000u) & 0xffffe000u;
            }
          }
          if (POOL == 1) {
            // W2 == 64: slices 0,1 = target row 2t (w 0..31, 32..63); 2,3 = row 2t+1.
            // avg_pool2d order: ((a00 + a01) + a10) + a11, then * 0.25
            const int hb = (sl & 1) * 16;
            if (sl < 2) {
#pragma unroll
              for (int w = 0; w < 16; ++w)
                h1[hb + w] = fadd(__uint_as_float(r[2 * w]), __uint_as_float(r[2 * w + 1]));
            } else {
#pragma unroll
              for (int w = 0; w < 16; ++w)
                h1[hb + w] = fmul(fadd(fadd(h1[hb + w], __uint_as_float(r[2 * w])),
                                       __uint_as_float(r[2 * w + 1])), 0.25f);
            }
          }
          // staging buffer `sbuf` was last read by the store issued two slices ago
          CT_ADD(4);
#if SB_CORR_FENCE_PAIRS
          // slices are staged in groups of kFenceGroup and fenced once; slice g of a group overwrites the buffer read by
          // the store issued kSBufs - g stores ago (none of the group's own stores has been issued yet)
          if (lane == 0) {
            const int pend = kSBufs - 1 - (sl % kFenceGroup);
            if (pend >= 3) ptx::tma_store_wait_read<3>();
            else if (pend == 2) ptx::tma_store_wait_read<2>();
            else if (pend == 1) ptx::tma_store_wait_read<1>();
            else ptx::tma_store_wait_read<0>();
          }
#else
          if (lane == 0) ptx::tma_store_wait_read<kSBufs - 1>();
#endif
          __syncwarp();
          CT_ADD(1);
          const uint32_t dst = my_stage + sbuf * kStageBufBytes + lane * 128;
#pragma unroll
          for (int c = 0; c < 8; ++c) {
            const uint32_t a = dst + (((uint32_t)c ^ ((uint32_t)lane & 7u)) << 4);
            asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(a), "r"(r[4 * c]),
                         "r"(r[4 * c + 1]), "r"(r[4 * c + 2]), "r"(r[4 * c + 3])
                         : "memory");
          }
          CT_ADD(4);
#if SB_CORR_FENCE_PAIRS
          if ((sl % kFenceGroup) == kFenceGroup - 1) {
            ptx::fence_proxy_async_smem();           // one generic -> async proxy fence for the group's slices
            __syncwarp();
            if (lane == 0) {
#pragma unroll
              for (int i = 0; i < kFenceGroup; ++i) {
                const uint32_t bi = (sbuf + (uint32_t)(kSBufs - (kFenceGroup - 1 - i))) % (uint32_t)kSBufs;
                TMA_STORE_V(&map_v, my_stage + bi * kStageBufBytes, t * BN + (sl - (kFenceGroup - 1 - i)) * 32, mb * BM + wq * 32, b);
                ptx::tma_store_commit();
              }
            }
          }
#else
          ptx::fence_proxy_async_smem();
          __syncwarp();
          if (lane == 0) {
            TMA_STORE_V(&map_v, my_stage + sbuf * kStageBufBytes, t * BN + sl * 32,
                              mb * BM + wq * 32, b);
            ptx::tma_store_commit();
          }
#endif
          if (++sbuf == kSBufs) sbuf = 0;
          CT_ADD(5);
        }
        }
        // accumulator buffer drained
        ptx::tc_fence_before_sync();
        __syncwarp();
        if (lane == 0) {
          if (TWO_CTA) ptx::mbar_arrive_cluster(ptx::mapa_shared(bar_t_empty + 8 * acc, 0));   // the leader's barrier
          else ptx::mbar_arrive(bar_t_empty + 8 * acc);
        }
        if (++acc == kNAcc) { acc = 0; acc_par ^= 1; }

        if (POOL == 1) {
          if (p.lvl1) {
            // level 1 of a tile = 32 rows x 32 floats per warp: same swizzled staging + TMA store as a
            // volume slice. The even tile's row is held in registers (h1a) and stored right before the
            // odd tile's, so that the two adjacent 128-byte pieces of every query reach L2 / DRAM together.
            if ((tt & 1) == 0 && t + 1 < t1) {
#pragma unroll
              for (int w = 0; w < 32; ++w) h1a[w] = h1[w];
            } else {
#pragma unroll
              for (int half = 0; half < 2; ++half) {
                if (half == 0 && (tt & 1) == 0) continue;          // unpaired last tile: only its own row
                const float* src = (half == 0) ? h1a : h1;
                const int tcol = (half == 0) ? t - 1 : t;
                if (lane == 0) ptx::tma_store_wait_read<kSBufs - 1>();
                __syncwarp();
                const uint32_t dst = my_stage + sbuf * kStageBufBytes + lane * 128;
#pragma unroll
                for (int c = 0; c < 8; ++c) {
                  const uint32_t a = dst + (((uint32_t)c ^ ((uint32_t)lane & 7u)) << 4);
                  asm volatile("st.shared.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(a), "f"(src[4 * c]),
                               "f"(src[4 * c + 1]), "f"(src[4 * c + 2]), "f"(src[4 * c + 3])
                               : "memory");
                }
                ptx::fence_proxy_async_smem();
                __syncwarp();
                if (lane == 0) {
                  TMA_STORE_V(&map_l1, my_stage + sbuf * kStageBufBytes, tcol * 32, mb * BM + wq * 32, b);
                  ptx::tma_store_commit();
                }
                if (++sbuf == kSBufs) sbuf = 0;
              }
            }
          }
          // level 2: pool level-1 rows (t even, t odd)
          if ((tt & 1) == 0) {
#pragma unroll
            for (int w = 0; w < 16; ++w) h2[w] = fadd(h1[2 * w], h1[2 * w + 1]);
          } else {
#pragma unroll
            for (int w = 0; w < 16; ++w)
              h2[w] = fmul(fadd(fadd(h2[w], h1[2 * w]), h1[2 * w + 1]), 0.25f);
            if (tt == 1) {
#pragma unroll
              for (int w = 0; w < 16; ++w) h2a[w] = h2[w];       // first level-2 row of the unit: kept for one store
#pragma unroll
              for (int w = 0; w < 8; ++w) h3[w] = fadd(h2[2 * w], h2[2 * w + 1]);
            } else {  // tt == 3
              if (p.lvl2) {
                // the unit's two level-2 rows = 32 floats = one full 128-byte line per query: staged and
                // TMA-stored like a volume slice (was: 64-byte pieces straight from registers)
                if (lane == 0) ptx::tma_store_wait_read<kSBufs - 1>();
                __syncwarp();
                const uint32_t dst = my_stage + sbuf * kStageBufBytes + lane * 128;
#pragma unroll
                for (int c = 0; c < 8; ++c) {
                  const uint32_t a = dst + (((uint32_t)c ^ ((uint32_t)lane & 7u)) << 4);
                  const float* src = (c < 4) ? (h2a + 4 * c) : (h2 + 4 * (c - 4));
                  asm volatile("st.shared.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(a), "f"(src[0]), "f"(src[1]),
                               "f"(src[2]), "f"(src[3])
                               : "memory");
                }
                ptx::fence_proxy_async_smem();
                __syncwarp();
                if (lane == 0) {
                  TMA_STORE_V(&map_l2, my_stage + sbuf * kStageBufBytes, (t >> 2) * 32, mb * BM + wq * 32, b);
                  ptx::tma_store_commit();
                }
                if (++sbuf == kSBufs) sbuf = 0;
              }
#pragma unroll
              for (int w = 0; w < 8; ++w)
                h3[w] = fmul(fadd(fadd(h3[w], h2[2 * w]), h2[2 * w + 1]), 0.25f);
              if (p.lvl3 && row_ok) {
                float* o = p.lvl3 + (q * p.H2e + (t >> 2)) * 8;
                stg_stream4(o, make_float4(h3[0], h3[1], h3[2], h3[3]));
                stg_stream4(o + 4, make_float4(h3[4], h3[5], h3[6], h3[7]));
              }
            }
          }
        }
      }
      }
    }
    if (lane == 0) ptx::tma_store_wait_all<0>();
#ifdef SB_CORR_TRACE
    if (warp == 4 && lane == 0 && blockIdx.x < 148) {
      g_corr_acc[blockIdx.x * 8 + 3] = ct_acc[0];
      g_corr_acc[blockIdx.x * 8 + 4] = ct_acc[1];
      g_corr_acc2[blockIdx.x * 4 + 0] = ct_acc[3];
      g_corr_acc2[blockIdx.x * 4 + 1] = ct_acc[4];
      g_corr_acc2[blockIdx.x * 4 + 2] = ct_acc[5];
      g_corr_acc[blockIdx.x * 8 + 5] = clock64() - ct_loop0;
    }
#endif
  }

  ptx::tc_fence_before_sync();
  __syncthreads();
  if (TWO_CTA) ptx::cluster_sync_all();     // neither CTA frees TMEM / exits while its peer still uses the pair
  if (warp == 2) {
    ptx::tc_fence_after_sync();
    if (TWO_CTA) ptx::tmem_dealloc_2cta(tmem_base, kTmemCols);
    else ptx::tmem_dealloc(tmem_base, kTmemCols);
  }
}

// ------------------------------------------------------- standalone pooling
__global__ void __launch_bounds__(256)
avg_pool2x2_kernel(const float* __restrict__ in, float* __restrict__ out, int H, int W, int Ho,
                   int Wo, long long total) {
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const long long pl = i / ((long long)Ho * Wo);
    const int rem = (int)(i - pl * Ho * Wo);
    const int y = rem / Wo, x = rem - y * Wo;
    const float* s = in + pl * H * W + (long long)(2 * y) * W + 2 * x;
    const float2 a = *reinterpret_cast<const float2*>(s);
    const float2 c = *reinterpret_cast<const float2*>(s + W);
    out[i] = fmul(fadd(fadd(fadd(a.x, a.y), c.x), c.y), 0.25f);
  }
}

// --------------------------------------------------------------- host side
static int make_map_3d(CUtensorMap* m, CUtensorMapDataType dt, int elt_bytes, const void* base,
                       unsigned long long d0, unsigned long long d1, unsigned long long d2,
                       unsigned b0, unsigned b1, const char* what) {
  return make_map_3d_ex(m, dt, elt_bytes, base, d0, d1, d2, b0, b1, CU_TENSOR_MAP_SWIZZLE_128B,
                        CU_TENSOR_MAP_L2_PROMOTION_L2_256B, what);
}

static inline int round_up(int a, int b) { return (a + b - 1) / b * b; }

// Word written (system scope) just before a protocol-timeout trap. It lives in
// mapped pinned host memory so the host can still read it after the context died.
}  // namespace sb

extern "C" size_t sb_corr_workspace_bytes(int B, int C, int N1, int N2) {
  if (B < 0 || C < 0 || N1 < 0 || N2 < 0) return 0;
  const size_t cpad = (size_t)sb::round_up(C, 64);
  // 256-byte aligned halves
  const size_t a = ((size_t)B * N1 * cpad * 2 + 255) / 256 * 256;
  const size_t b = ((size_t)B * N2 * cpad * 2 + 255) / 256 * 256;
  return a + b;
}

extern "C" int sb_feat_to_tokens_bf16(const float* fmap, void* tok, int B, int C, int N,
                                      sb_stream_t stream) {
  using namespace sb;
  SB_ENTER();
  SB_REQUIRE(fmap && tok, SB_EINVAL, "sb_feat_to_tokens_bf16: null pointer");
  SB_REQUIRE(B >= 0 && C > 0 && N >= 0, SB_EINVAL, "sb_feat_to_tokens_bf16: bad size");
  SB_REQUIRE(aligned16(fmap) && aligned16(tok), SB_EINVAL,
             "sb_feat_to_tokens_bf16: pointers must be 16-byte aligned");
  SB_REQUIRE(B <= 65535, SB_EUNSUP, "sb_feat_to_tokens_bf16: B > 65535");
  if (B == 0 || N == 0) return SB_OK;
  const int Cpad = round_up(C, 64);
  dim3 grid((N + kTokTile - 1) / kTokTile, Cpad / kTokTile, B);
  feat_to_tokens_bf16_kernel<<<grid, 256, 0, as_stream(stream)>>>(
      fmap, reinterpret_cast<uint16_t*>(tok), C, Cpad, N);
  SB_LAUNCH_CHECK("feat_to_tokens_bf16_kernel");
  return SB_OK;
}

extern "C" int sb_avg_pool2x2(const float* in, float* out, long long planes, int H, int W,
                              sb_stream_t stream) {
  using namespace sb;
  SB_ENTER();
  SB_REQUIRE(in && out, SB_EINVAL, "sb_avg_pool2x2: null pointer");
  SB_REQUIRE(planes >= 0 && H >= 0 && W >= 0, SB_EINVAL, "sb_avg_pool2x2: bad size");
  SB_REQUIRE((W & 1) == 0 && (reinterpret_cast<uintptr_t>(in) & 7) == 0, SB_EUNSUP,
             "sb_avg_pool2x2: W must be even and `in` 8-byte aligned");
  const int Ho = H / 2, Wo = W / 2;
  const long long total = planes * Ho * Wo;
  if (total == 0) return SB_OK;
  long long blocks = (total + 255) / 256;
  const long long max_blocks = (long long)kNumSMs * 8 * 16;
  if (blocks > max_blocks) blocks = max_blocks;
  avg_pool2x2_kernel<<<(int)blocks, 256, 0, as_stream(stream)>>>(in, out, H, W, Ho, Wo, total);
  SB_LAUNCH_CHECK("avg_pool2x2_kernel");
  return SB_OK;
}

extern "C" int sb_corr_tokens_pitched(const void* tok1, const void* tok2, float* vol, long long vol_pitch,
                                      float* lvl1, float* lvl2, float* lvl3, int B, int C, int H1, int W1,
                                      int H2, int W2, sb_stream_t stream);

extern "C" int sb_corr_tokens(const void* tok1, const void* tok2, float* vol, float* lvl1,
                              float* lvl2, float* lvl3, int B, int C, int H1, int W1, int H2,
                              int W2, sb_stream_t stream) {
  return sb_corr_tokens_pitched(tok1, tok2, vol, (long long)H2 * W2, lvl1, lvl2, lvl3, B, C, H1, W1, H2, W2, stream);
}

// vol_pitch: row pitch of the volume in floats (>= H2*W2, a multiple of 4): rows of a volume whose
// token count is not a multiple of 4 are padded by the caller so that the TMA store strides stay
// 16-byte multiples; the pad columns are never written.
namespace sb {
static int corr_tokens_impl(const void* tok1, const void* tok2, void* vol_any, long long vol_pitch, bool bf16_out,
                            float* lvl1, float* lvl2, float* lvl3, int B, int C, int H1, int W1, int H2, int W2,
                            sb_stream_t stream, float* smx_stats = nullptr, int b_rot = 0);
}

extern "C" int sb_corr_tokens_pitched(const void* tok1, const void* tok2, float* vol, long long vol_pitch,
                                      float* lvl1, float* lvl2, float* lvl3, int B, int C, int H1, int W1,
                                      int H2, int W2, sb_stream_t stream) {
  return sb::corr_tokens_impl(tok1, tok2, vol, vol_pitch, false, lvl1, lvl2, lvl3, B, C, H1, W1, H2, W2, stream);
}

// Both directions of a batch of pairs in ONE launch: tok_both [2 * Bp, N, Cpad] = the token maps of image 1 of every
// pair followed by those of image 2; volume element b < Bp is corr(img1_b, img2_b) (forward), element Bp + b is
// corr(img2_b, img1_b) (backward): the B operand of batch element b is token map (b + Bp) mod 2 Bp.
extern "C" int sb_corr_tokens_bidir(const void* tok_both, float* vol, float* lvl1, float* lvl2, float* lvl3, int Bp,
                                    int C, int H, int W, sb_stream_t stream) {
  return sb::corr_tokens_impl(tok_both, tok_both, vol, (long long)H * W, false, lvl1, lvl2, lvl3, 2 * Bp, C, H, W, H, W,
                              stream, nullptr, Bp);
}

// softmax over the keys of q . k^T, TF32-rounded probabilities [B, Nq, Nk] (GMA Attention.forward);
// stats: workspace of B * Nq * 2 floats.
extern "C" int sb_attn_softmax_tokens(const void* tok_q, const void* tok_k, float* attn, float* stats, int B, int C,
                                      int Nq, int Nk, sb_stream_t stream) {
  if (!stats) { sb::set_error("sb_attn_softmax_tokens: null stats workspace"); return SB_EINVAL; }
  return sb::corr_tokens_impl(tok_q, tok_k, attn, (long long)Nk, false, nullptr, nullptr, nullptr, B, C, 1, Nq, 1, Nk,
                              stream, stats);
}

// bf16 volume [B, N1, N2] (N2 % 8 == 0): opt-in, not the reference's dtype
extern "C" int sb_corr_tokens_bf16out(const void* tok1, const void* tok2, void* vol_bf16, int B, int C, int H1,
                                      int W1, int H2, int W2, sb_stream_t stream) {
  return sb::corr_tokens_impl(tok1, tok2, vol_bf16, (long long)H2 * W2, true, nullptr, nullptr, nullptr, B, C, H1,
                              W1, H2, W2, stream);
}

namespace sb {
static int corr_tokens_impl(const void* tok1, const void* tok2, void* vol_any, long long vol_pitch, bool bf16_out,
                            float* lvl1, float* lvl2, float* lvl3, int B, int C, int H1, int W1, int H2, int W2,
                            sb_stream_t stream, float* smx_stats, int b_rot) {
  float* vol = static_cast<float*>(vol_any);
  SB_ENTER();
  SB_REQUIRE(tok1 && tok2 && vol, SB_EINVAL, "sb_corr_tokens: null pointer");
  SB_REQUIRE(B >= 0 && C > 0 && H1 >= 0 && W1 >= 0 && H2 >= 0 && W2 >= 0, SB_EINVAL,
             "sb_corr_tokens: bad size");
  SB_REQUIRE(C <= 64 * kMaxPanels, SB_EUNSUP, "sb_corr_tokens: C=%d > %d not supported", C,
             64 * kMaxPanels);
  const long long N1 = (long long)H1 * W1, N2 = (long long)H2 * W2;
  SB_REQUIRE(N1 < (1 << 30) && N2 < (1 << 30), SB_EUNSUP, "sb_corr_tokens: too many tokens");
  if (B == 0 || N1 == 0 || N2 == 0) return SB_OK;
  SB_REQUIRE(!bf16_out || (N2 & 7) == 0, SB_EUNSUP, "sb_corr_tokens_bf16out: H2*W2 must be a multiple of 8");
  SB_REQUIRE(vol_pitch >= N2 && (vol_pitch & 3) == 0, SB_EUNSUP,
             "sb_corr_tokens: the volume row pitch (%lld floats) must be >= H2*W2 = %lld and a multiple of 4 "
             "(TMA strides); use sb_corr_tokens_pitched with a padded pitch", vol_pitch, N2);
  SB_REQUIRE(aligned16(tok1) && aligned16(tok2) && aligned16(vol), SB_EINVAL,
             "sb_corr_tokens: pointers must be 16-byte aligned");
  const bool want_pool = lvl1 || lvl2 || lvl3;
  SB_REQUIRE(!want_pool || vol_pitch == N2, SB_EUNSUP, "sb_corr_tokens: the pyramid needs a dense volume (pitch == H2*W2)");
  if (want_pool)
    SB_REQUIRE((H2 % 8) == 0 && (W2 % 8) == 0, SB_EUNSUP,
               "sb_corr_tokens: pyramid needs H2, W2 multiples of 8 (got %d x %d)", H2, W2);
  const int pool_mode = !want_pool ? 0 : (W2 == 64 ? 1 : (W2 == 128 ? 2 : 0));
  const bool fused_pool = pool_mode != 0;
  cudaStream_t s = as_stream(stream);
  const int Cpad = round_up(C, 64);

  CUtensorMap map_a, map_b, map_v, map_l1, map_l2;
  int rc;
  rc = make_map_3d(&map_a, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, tok1, Cpad, N1, B, BKP, BM, "A");
  if (rc) return rc;
  // CTA-pair mode (tcgen05 cta_group::2): every CTA loads half of each B tile
  SB_REQUIRE(!smx_stats || (!bf16_out && !want_pool), SB_EINVAL, "sb_attn_softmax_tokens: fp32 output without pyramid only");
  SB_REQUIRE(!smx_stats || (reinterpret_cast<uintptr_t>(smx_stats) & 7) == 0, SB_EINVAL, "sb_attn_softmax_tokens: stats must be 8-byte aligned");
  const bool two_cta = smx_stats ? false : (tune_get(SB_TUNE_CORR_2CTA, 1) == 2 && !(lvl1 || lvl2 || lvl3) ? true
                       : (tune_get(SB_TUNE_CORR_2CTA, 1) == 2 && W2 == 64));   // the W2 == 128 pyramid runs one CTA per tile
  rc = make_map_3d(&map_b, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, tok2, Cpad, N2, B, BKP, two_cta ? BN / 2 : BN, "B");
  if (rc) return rc;
  if (bf16_out)
    rc = make_map_3d_ex(&map_v, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, vol_any, N2, N1, B, 64, 32, CU_TENSOR_MAP_SWIZZLE_128B,
                        CU_TENSOR_MAP_L2_PROMOTION_L2_256B, "V16", (unsigned long long)vol_pitch);
  else
    rc = make_map_3d_ex(&map_v, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, vol, N2, N1, B, 32, 32, CU_TENSOR_MAP_SWIZZLE_128B,
                        CU_TENSOR_MAP_L2_PROMOTION_L2_256B, "V", (unsigned long long)vol_pitch);
  if (rc) return rc;
  map_l1 = map_v;
  if (fused_pool && lvl1) {
    // lvl1 [B*N1, H2/2, 32] viewed as rows of (H2/2)*32 floats: tile t owns columns [32t, 32t+32)
    SB_REQUIRE(aligned16(lvl1), SB_EINVAL, "sb_corr_tokens: lvl1 must be 16-byte aligned");
    rc = make_map_3d(&map_l1, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, lvl1, (unsigned long long)(H2 / 2) * (W2 / 2), N1, B,
                     32, 32, "L1");
    if (rc) return rc;
  }

  map_l2 = map_v;
  if (fused_pool && lvl2) {
    // lvl2 [B*N1, H2/4, 16] viewed as rows of (H2/4)*16 floats: unit ng owns columns [32 ng, 32 ng + 32)
    SB_REQUIRE(aligned16(lvl2), SB_EINVAL, "sb_corr_tokens: lvl2 must be 16-byte aligned");
    rc = make_map_3d(&map_l2, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, lvl2, (unsigned long long)(H2 / 4) * (W2 / 4), N1, B,
                     32, 32, "L2");
    if (rc) return rc;
  }
  unsigned int* g_dbg = debug_word_device();
  if (!g_dbg) {
    return SB_ECUDA;
  }

  const bool a_tmem_req = tune_get(SB_TUNE_CORR_A_TMEM, 1) == 1 && !two_cta && !bf16_out && !smx_stats && pool_mode != 2;
  CorrParams p;
  p.B = B; p.N1 = (int)N1; p.N2 = (int)N2; p.KP = Cpad / 64;
  p.MB = (int)((N1 + BM - 1) / BM);
  if (two_cta) p.MB = (p.MB + 1) / 2;             // units are pairs of query blocks
  p.NT = (int)((N2 + BN - 1) / BN);
  // tiles per unit: 4 (default) keeps A for 4 B tiles; SB_TUNE_CORR_TILES_PER_UNIT = 8 / 16 halves / quarters the A reloads
  int tpu_tune = tune_get(SB_TUNE_CORR_TILES_PER_UNIT, kTilesPerUnit);
  if (tpu_tune != 8 && tpu_tune != 16) tpu_tune = kTilesPerUnit;
  const int tpu = (pool_mode == 2) ? 8 : ((two_cta || a_tmem_req) ? kTilesPerUnit : tpu_tune);
  p.tpu = tpu;
  p.smx = smx_stats;
  p.b_rot = (B > 0) ? ((b_rot % B) + B) % B : 0;
  p.NG = (p.NT + tpu - 1) / tpu;
  p.n_units = (long long)B * p.MB * p.NG;
  p.H2h = H2 / 2; p.H2q = H2 / 4; p.H2e = H2 / 8;
  p.lvl1 = fused_pool ? lvl1 : nullptr;
  p.lvl2 = fused_pool ? lvl2 : nullptr;
  p.lvl3 = fused_pool ? lvl3 : nullptr;
  p.dbg = g_dbg;
  {
    const int tp = tune_get(SB_TUNE_CORR_STORE_POLICY, 1);   // bits 0-1: stores (1 evict_first default, 2 evict_last, 3 none); bit 2: operand loads evict_last
    p.store_policy = ((tp & 3) == 3 ? 0 : (tp & 3)) | (tp & 4);
  }

  const bool a_tmem = a_tmem_req;
  // dynamic unit scheduling (cluster launch control): one CTA per unit in the grid, see the kernel
  p.dyn = (!two_cta && tune_get(SB_TUNE_CORR_DYNAMIC, 1) == 1 && p.n_units > kNumSMs && p.n_units < (1ll << 30)) ? 1 : 0;
  const int grid = p.dyn ? (int)p.n_units : (int)((p.n_units < kNumSMs) ? p.n_units : kNumSMs);
  static SmemOptIn opt_in;
  int opt_dev;
  if (opt_in.need(kSmemTotal, &opt_dev)) {
    SB_CUDA(cudaFuncSetAttribute(corr_umma_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemTotal));
    SB_CUDA(cudaFuncSetAttribute(corr_umma_kernel<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemTotal));
    SB_CUDA(cudaFuncSetAttribute(corr_umma_kernel<0, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemTotal));
    SB_CUDA(cudaFuncSetAttribute(corr_umma_kernel<1, false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemTotal));
    SB_CUDA(cudaFuncSetAttribute(corr_umma_kernel<0, false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemTotal));
    SB_CUDA(cudaFuncSetAttribute(corr_umma_kernel<0, true, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemTotal));
    SB_CUDA(cudaFuncSetAttribute(corr_umma_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemTotal));
    SB_CUDA(cudaFuncSetAttribute(corr_umma_kernel<0, false, false, 0, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemTotal));
    SB_CUDA(cudaFuncSetAttribute(corr_umma_kernel<1, false, false, 0, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemTotal));
    SB_CUDA(cudaFuncSetAttribute(corr_umma_kernel<0, false, false, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemTotal));
    SB_CUDA(cudaFuncSetAttribute(corr_umma_kernel<0, false, false, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemTotal));
    opt_in.done(kSmemTotal, opt_dev);
  }
  if (smx_stats) {
    // pass 1: one unit per row block, all NT tiles; pass 2: the usual units
    CorrParams p1 = p;
    p1.tpu = p.NT; p1.NG = 1; p1.n_units = (long long)B * p.MB;
    p1.dyn = (p.dyn && p1.n_units > kNumSMs) ? 1 : 0;
    const int grid1 = p1.dyn ? (int)p1.n_units : (int)((p1.n_units < kNumSMs) ? p1.n_units : kNumSMs);
    corr_umma_kernel<0, false, false, 1><<<grid1, 384, kSmemTotal, s>>>(map_a, map_b, map_v, map_l1, map_l2, p1);
    SB_LAUNCH_CHECK("corr_umma_kernel<softmax statistics>");
    corr_umma_kernel<0, false, false, 2><<<grid, 256, kSmemTotal, s>>>(map_a, map_b, map_v, map_l1, map_l2, p);
    SB_LAUNCH_CHECK("corr_umma_kernel<softmax normalise>");
    return SB_OK;
  }
  if (two_cta) {
    long long clusters = p.n_units < kNumSMs / 2 ? p.n_units : kNumSMs / 2;
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3((unsigned)(2 * clusters));
    cfg.blockDim = dim3(256);
    cfg.dynamicSmemBytes = kSmemTotal;
    cfg.stream = s;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = 2; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr; cfg.numAttrs = 1;
    if (bf16_out) SB_CUDA(cudaLaunchKernelEx(&cfg, corr_umma_kernel<0, true, true>, map_a, map_b, map_v, map_l1, map_l2, p));
    else if (fused_pool) SB_CUDA(cudaLaunchKernelEx(&cfg, corr_umma_kernel<1, false, true>, map_a, map_b, map_v, map_l1, map_l2, p));
    else SB_CUDA(cudaLaunchKernelEx(&cfg, corr_umma_kernel<0, false, true>, map_a, map_b, map_v, map_l1, map_l2, p));
  } else if (bf16_out)
    corr_umma_kernel<0, true><<<grid, 256, kSmemTotal, s>>>(map_a, map_b, map_v, map_l1, map_l2, p);
  else if (pool_mode == 2)
    corr_umma_kernel<2><<<grid, 256, kSmemTotal, s>>>(map_a, map_b, map_v, map_l1, map_l2, p);
  else if (fused_pool && a_tmem)
    corr_umma_kernel<1, false, false, 0, true><<<grid, 256, kSmemTotal, s>>>(map_a, map_b, map_v, map_l1, map_l2, p);
  else if (fused_pool)
    corr_umma_kernel<1><<<grid, 256, kSmemTotal, s>>>(map_a, map_b, map_v, map_l1, map_l2, p);
  else if (a_tmem)
    corr_umma_kernel<0, false, false, 0, true><<<grid, 256, kSmemTotal, s>>>(map_a, map_b, map_v, map_l1, map_l2, p);
  else
    corr_umma_kernel<0><<<grid, 256, kSmemTotal, s>>>(map_a, map_b, map_v, map_l1, map_l2, p);
  SB_LAUNCH_CHECK("corr_umma_kernel");

  if (want_pool && !fused_pool) {
    // generic target width: chained standalone pooling over the finished volume
    const long long planes = (long long)B * N1;
    float* l1 = lvl1;
    SB_REQUIRE(lvl1 != nullptr, SB_EUNSUP,
               "sb_corr_tokens: for W2 != 64 the pyramid is chained, lvl1 must be provided");
    rc = sb_avg_pool2x2(vol, l1, planes, H2, W2, stream);
    if (rc) return rc;
    if (lvl2 || lvl3) {
      SB_REQUIRE(lvl2 != nullptr, SB_EUNSUP, "sb_corr_tokens: lvl3 requires lvl2 when W2 != 64");
      rc = sb_avg_pool2x2(l1, lvl2, planes, H2 / 2, W2 / 2, stream);
      if (rc) return rc;
    }
    if (lvl3) {
      rc = sb_avg_pool2x2(lvl2, lvl3, planes, H2 / 4, W2 / 4, stream);
      if (rc) return rc;
    }
  }
  return SB_OK;
}
}  // namespace sb

extern "C" int sb_corr(const float* fmap1, const float* fmap2, float* vol, float* lvl1, float* lvl2,
                       float* lvl3, void* workspace, size_t workspace_bytes, int B, int C, int H1,
                       int W1, int H2, int W2, sb_stream_t stream) {
  using namespace sb;
  SB_ENTER();
  SB_REQUIRE(fmap1 && fmap2 && vol, SB_EINVAL, "sb_corr: null pointer");
  SB_REQUIRE(B >= 0 && C > 0 && H1 >= 0 && W1 >= 0 && H2 >= 0 && W2 >= 0, SB_EINVAL, "sb_corr: bad size");
  const long long N1 = (long long)H1 * W1, N2 = (long long)H2 * W2;
  if (B == 0 || N1 == 0 || N2 == 0) return SB_OK;
  const size_t need = sb_corr_workspace_bytes(B, C, (int)N1, (int)N2);
  SB_REQUIRE(workspace && workspace_bytes >= need, SB_EINVAL,
             "sb_corr: workspace too small (%zu < %zu bytes)", workspace_bytes, need);
  SB_REQUIRE((reinterpret_cast<uintptr_t>(workspace) & 255) == 0, SB_EINVAL,
             "sb_corr: workspace must be 256-byte aligned");
  const size_t cpad = (size_t)round_up(C, 64);
  const size_t a_bytes = ((size_t)B * N1 * cpad * 2 + 255) / 256 * 256;
  void* tok1 = workspace;
  void* tok2 = static_cast<uint8_t*>(workspace) + a_bytes;
  int rc = sb_feat_to_tokens_bf16(fmap1, tok1, B, C, (int)N1, stream);
  if (rc) return rc;
  rc = sb_feat_to_tokens_bf16(fmap2, tok2, B, C, (int)N2, stream);
  if (rc) return rc;
  return sb_corr_tokens(tok1, tok2, vol, lvl1, lvl2, lvl3, B, C, H1, W1, H2, W2, stream);
}

#ifdef SB_CORR_TRACE
// debug builds only (not declared in include/stitch_b200.h)
extern "C" int sb_corr_acc2_read(long long* host_out) {
  if (cudaDeviceSynchronize() != cudaSuccess) return -1;
  return cudaMemcpyFromSymbol(host_out, sb::g_corr_acc2, sizeof(long long) * 148 * 4) == cudaSuccess ? 0 : -1;
}
extern "C" int sb_corr_acc_read(long long* host_out) {
  if (cudaDeviceSynchronize() != cudaSuccess) return -1;
  return cudaMemcpyFromSymbol(host_out, sb::g_corr_acc, sizeof(long long) * 148 * 8) == cudaSuccess ? 0 : -1;
}
#endif
