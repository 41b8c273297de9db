// gma.cu — N1 ("next" row 1): GMA attention and aggregation.
// Replaces the dense parts of
//   Attention.forward  (core/FlowFormer/PerCostFormer3/gma.py:54-76):  attn = softmax(scale*q . k^T)
//   Aggregate.forward  (gma.py:102-115):  out = fmap + gamma * (attn @ v)   (project == None)
// called once (decoder.py:283) resp. every GRU iteration (gru.py:324 via GMAUpdateBlock).
//
//  * q.k^T is the same all-pairs contraction as the cost volume (K = dim_head = 128) and runs on
//    corr_umma_kernel (corr_tcgen05.cu); softmax_rows_kernel below normalises the rows in place and
//    rounds the probabilities to TF32 (round-to-nearest), so that
//  * attn_v_umma_kernel can feed the fp32 attention matrix straight to the tensor cores as
//    tcgen05.mma.kind::tf32 operands (no conversion pass, no bf16 copy): per (batch*head, 128-query
//    block) tile, D[128 x 128] += A[128 x 32] . B[128 x 32]^T over K = Nk keys, A = attn rows
//    (K-major, 128-byte rows = 32 fp32, SWIZZLE_128B), B = v [d, Nk] (K-major as the 1x1 conv leaves
//    it).  6-stage TMA ring (32 KB per stage), one elected MMA thread, fp32 accumulators in TMEM
//    (double-buffered), epilogue warps write D transposed ([d, Nq], the conv layout the reference
//    rearranges to) fused with the residual fmap + gamma * out.
//    Roofline: HBM — the attention matrix (Nq*Nk*4 B per batch*head) is read once per call; the
//    reference reads it with an fp32 cuBLAS bmm (FFMA pipe).
#include <cuda.h>
#include <cuda_bf16.h>

#include "common.cuh"
#include "ptx_sm100.cuh"
#include "tmap.cuh"

namespace sb {

// ------------------------------------------------------------------ softmax
__device__ __forceinline__ float round_tf32(float x) {
  uint32_t u;
  asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(u) : "f"(x));
  return __uint_as_float(u);
}

template <bool IS_MAX>
__device__ __forceinline__ float block_reduce(float v, float* s_red) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    const float t = __shfl_xor_sync(0xffffffffu, v, o);
    v = IS_MAX ? fmaxf(v, t) : v + t;
  }
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (lane == 0) s_red[warp] = v;
  __syncthreads();
  float r = s_red[0];
#pragma unroll
  for (int w = 1; w < 8; ++w) r = IS_MAX ? fmaxf(r, s_red[w]) : r + s_red[w];
  __syncthreads();
  return r;
}

// one CTA (256 threads) per row; the row lives in registers between the passes (n <= 256 * 4 * kMaxVec)
constexpr int kSoftmaxVec = 8;   // float4 per thread -> rows up to 8192
template <bool BF16_OUT>
__global__ void __launch_bounds__(256)
softmax_rows_kernel(float* __restrict__ x, long long rows, int n, long long row_stride, int to_tf32,
                    uint16_t* __restrict__ out_bf16) {
  __shared__ float s_red[8];
  for (long long row = blockIdx.x; row < rows; row += gridDim.x) {
    float* p = x + row * row_stride;
    float4 v[kSoftmaxVec];
    float mx = -INFINITY;
#pragma unroll
    for (int i = 0; i < kSoftmaxVec; ++i) {
      const int c = (i * 256 + threadIdx.x) * 4;
      if (c < n) {
        v[i] = *reinterpret_cast<const float4*>(p + c);
        mx = fmaxf(fmaxf(fmaxf(mx, v[i].x), fmaxf(v[i].y, v[i].z)), v[i].w);
      }
    }
    mx = block_reduce<true>(mx, s_red);
    float sum = 0.0f;
#pragma unroll
    for (int i = 0; i < kSoftmaxVec; ++i) {
      const int c = (i * 256 + threadIdx.x) * 4;
      if (c < n) {
        v[i].x = expf(v[i].x - mx); v[i].y = expf(v[i].y - mx);
        v[i].z = expf(v[i].z - mx); v[i].w = expf(v[i].w - mx);
        sum += (v[i].x + v[i].y) + (v[i].z + v[i].w);
      }
    }
    sum = block_reduce<false>(sum, s_red);
#pragma unroll
    for (int i = 0; i < kSoftmaxVec; ++i) {
      const int c = (i * 256 + threadIdx.x) * 4;
      if (c < n) {
        float4 o = make_float4(v[i].x / sum, v[i].y / sum, v[i].z / sum, v[i].w / sum);
        if (BF16_OUT) {      // bf16 copy [rows, n] (dense) instead of the in-place fp32 result
          __nv_bfloat162 lo = __floats2bfloat162_rn(o.x, o.y), hi = __floats2bfloat162_rn(o.z, o.w);
          *reinterpret_cast<uint2*>(out_bf16 + row * (long long)n + c) =
              make_uint2(*reinterpret_cast<uint32_t*>(&lo), *reinterpret_cast<uint32_t*>(&hi));
          continue;
        }
        if (to_tf32) { o.x = round_tf32(o.x); o.y = round_tf32(o.y); o.z = round_tf32(o.z); o.w = round_tf32(o.w); }
        *reinterpret_cast<float4*>(p + c) = o;
      }
    }
  }
}

// ------------------------------------------------------------- attn @ v
constexpr int GM = 128, GN = 128, GK = 32;           // block; K step = 32 fp32 = one 128-byte swizzle row
constexpr int kGBlkBytes = GM * 128;                 // one operand block of one K step: 128 rows x 128 B = 16 KB
constexpr int kGRingBudget = 192 * 1024;
constexpr int kGSmemTotal = kGRingBudget + 256 + 1024;
constexpr int kGTmemCols = 512;

// instruction descriptor: D = f32, A = B = TF32 (format 2), both K-major, N = 128, M = 128
constexpr uint32_t kIdescTf32 = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(GN >> 3) << 17) |
                                ((uint32_t)(GM >> 4) << 24);
// same with A = B = BF16 (format 1): a K step is then 64 elements (still one 128-byte row)
constexpr uint32_t kIdescBf16 = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(GN >> 3) << 17) |
                                ((uint32_t)(GM >> 4) << 24);

__device__ __forceinline__ void umma_tf32(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc,
                                          uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}

struct AttnVParams {
  int BH, Nq, Nk, d;            // batch*heads, queries (M), keys (K), head dim (transposed mode: == 128)
  int MU, KS;                   // units (MBLK query blocks each) per batch*head, K steps
  int NT, ncols;                // row-major mode (plain batched D = A . B^T): column tiles, total columns
  int row_major;                // 0: out[bh, n, i] (+ residual, gamma)   1: out[bh, i, n]
  long long n_units;
  const float* residual;        // [BH, d, Nq] or nullptr
  const float* gamma;           // device scalar or nullptr (1.0)
  float* out;                   // [BH, d, Nq]
  unsigned int* dbg;
};

// MBLK = 128-query blocks per CTA unit: MBLK blocks share each v stage (A: MBLK x 16 KB + B: 16 KB
// per K step, so v is re-read from L2 MBLK times less often), accumulators = MBLK x 128 TMEM columns,
// 512 / (MBLK * 128) accumulator sets (MBLK = 4: one set, the epilogue is not overlapped).
template <int MBLK, bool BF16>
__global__ void __launch_bounds__(256, 1)
attn_v_umma_kernel(const __grid_constant__ CUtensorMap map_a, const __grid_constant__ CUtensorMap map_b,
                   const AttnVParams p) {
  constexpr int kKStep = BF16 ? 64 : GK;                       // elements per 128-byte operand row
  constexpr int kStageBytes = (MBLK + 1) * kGBlkBytes;
  constexpr int kStages = kGRingBudget / kStageBytes;          // 6, 4, 2 for MBLK = 1, 2, 4
  constexpr int kAccSets = 4 / MBLK;                           // 4, 2, 1
  extern __shared__ uint8_t smem_raw[];
  const uint32_t smem_base = (ptx::smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t sBar = smem_base + kGRingBudget;
  const uint32_t bar_full = sBar;                       // [kStages]   (<= 6)
  const uint32_t bar_empty = sBar + 48;                 // [kStages]
  const uint32_t bar_t_full = sBar + 96;                // [kAccSets]  (<= 4)
  const uint32_t bar_t_empty = sBar + 128;              // [kAccSets]
  const uint32_t tmem_slot = sBar + 160;
  uint8_t* smem_gen = smem_raw + (smem_base - ptx::smem_u32(smem_raw));
  volatile uint32_t* tmem_slot_ptr = reinterpret_cast<volatile uint32_t*>(smem_gen + kGRingBudget + 160);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  if (warp == 0 && lane == 0) {
    ptx::prefetch_tensormap(&map_a);
    ptx::prefetch_tensormap(&map_b);
  }
  if (warp == 1 && lane == 0) {
    for (int s = 0; s < kStages; ++s) {
      ptx::mbar_init(bar_full + 8 * s, 1);
      ptx::mbar_init(bar_empty + 8 * s, 1);
    }
    for (int a = 0; a < kAccSets; ++a) {
      ptx::mbar_init(bar_t_full + 8 * a, 1);
      ptx::mbar_init(bar_t_empty + 8 * a, 4);           // one elected lane per epilogue warp
    }
    ptx::fence_mbar_init();
  }
  if (warp == 2) {
    ptx::tmem_alloc(tmem_slot, kGTmemCols);
    ptx::tmem_relinquish();
  }
  ptx::tc_fence_before_sync();
  __syncthreads();
  ptx::tc_fence_after_sync();
  const uint32_t tmem_base = *tmem_slot_ptr;

  if (warp == 0) {
    // ================================================================ producer
    if (lane == 0) {
      uint32_t stage = 0, par = 0;
      const uint64_t pol_keep = ptx::l2_policy_evict_last(), pol_stream = ptx::l2_policy_evict_first();
      for (long long u = blockIdx.x; u < p.n_units; u += gridDim.x) {
        const int nt = (int)(u % p.NT);
        const long long u2 = u / p.NT;
        const int mu = (int)(u2 % p.MU), bh = (int)(u2 / p.MU);
        for (int ks = 0; ks < p.KS; ++ks) {
          ptx::mbar_wait(bar_empty + 8 * stage, par ^ 1, 0x21, p.dbg);
          ptx::mbar_arrive_expect_tx(bar_full + 8 * stage, kStageBytes);   // rows past Nq / keys past Nk arrive as zeros
          const uint32_t sa = smem_base + stage * kStageBytes;
          // B (v / the second operand) is re-read by every unit of its batch element: keep it in L2;
          // A (the attention matrix) streams through once: evict first
          ptx::tma_load_3d_hint(sa + MBLK * kGBlkBytes, &map_b, bar_full + 8 * stage, ks * kKStep, nt * GN, bh, pol_keep);
#pragma unroll
          for (int m = 0; m < MBLK; ++m)
            ptx::tma_load_3d_hint(sa + m * kGBlkBytes, &map_a, bar_full + 8 * stage, ks * kKStep, (mu * MBLK + m) * GM, bh,
                                  p.row_major ? pol_keep : pol_stream);
          if (++stage == kStages) { stage = 0; par ^= 1; }
        }
      }
    }
    __syncwarp();
  } else if (warp == 1) {
    // ============================================================== MMA issuer
    if (lane == 0) {
      uint32_t stage = 0, par = 0, acc = 0, acc_par = 0;
      for (long long u = blockIdx.x; u < p.n_units; u += gridDim.x) {
        ptx::mbar_wait(bar_t_empty + 8 * acc, acc_par ^ 1, 0x22, p.dbg);
        ptx::tc_fence_after_sync();
        const uint32_t d_tmem = tmem_base + acc * (MBLK * GN);
        for (int ks = 0; ks < p.KS; ++ks) {
          ptx::mbar_wait(bar_full + 8 * stage, par, 0x23, p.dbg);
          ptx::tc_fence_after_sync();
          const uint32_t sa = smem_base + stage * kStageBytes;
          const uint64_t bdesc = ptx::umma_desc_k_sw128(sa + MBLK * kGBlkBytes);
#pragma unroll
          for (int m = 0; m < MBLK; ++m) {
            const uint64_t adesc = ptx::umma_desc_k_sw128(sa + m * kGBlkBytes);
#pragma unroll
            for (int k = 0; k < 4; ++k) {      // 8 tf32 / 16 bf16 = 32 bytes along K inside the swizzle atom: +2 in (addr >> 4)
              if (BF16) ptx::umma_f16(d_tmem + m * GN, adesc + 2 * k, bdesc + 2 * k, kIdescBf16, (ks | k) != 0);
              else umma_tf32(d_tmem + m * GN, adesc + 2 * k, bdesc + 2 * k, kIdescTf32, (ks | k) != 0);
            }
          }
          ptx::umma_commit(bar_empty + 8 * stage);      // stage reusable once these MMAs retire
          if (++stage == kStages) { stage = 0; par ^= 1; }
        }
        ptx::umma_commit(bar_t_full + 8 * acc);         // accumulators ready for the epilogue
        if (++acc == kAccSets) { acc = 0; acc_par ^= 1; }
      }
    }
    __syncwarp();
  } else if (warp >= 4) {
    // ================================================================ epilogue
    const int wq = warp - 4;                            // TMEM lane quarter == warp % 4
    const uint32_t lane_taddr = tmem_base + ((uint32_t)(wq * 32) << 16);
    const float g = p.gamma ? __ldg(p.gamma) : 1.0f;
    uint32_t acc = 0, acc_par = 0;
    for (long long u = blockIdx.x; u < p.n_units; u += gridDim.x) {
      const int nt = (int)(u % p.NT);
      const long long u2 = u / p.NT;
      const int mu = (int)(u2 % p.MU), bh = (int)(u2 / p.MU);
      ptx::mbar_wait(bar_t_full + 8 * acc, acc_par, 0x24, p.dbg);
      ptx::tc_fence_after_sync();
#pragma unroll 1
      for (int m = 0; m < MBLK; ++m) {
        const int i = (mu * MBLK + m) * GM + wq * 32 + lane;   // query index
        const bool ok = i < p.Nq;
#pragma unroll 1
        for (int sl = 0; sl < 4; ++sl) {
          uint32_t r[32];
          ptx::tmem_ld_32x32b_x32(lane_taddr + acc * (MBLK * GN) + m * GN + sl * 32, r);
          ptx::tmem_ld_wait();
          if (p.row_major) {
            // D[i, n] -> out[bh, i, nt*128 + n]: every thread writes its row's 32 floats (one 128-byte line)
            const int n0 = nt * GN + sl * 32;
            if (ok && n0 < p.ncols) {
              float* orow = p.out + ((size_t)bh * p.Nq + i) * p.ncols + n0;
              if (n0 + 32 <= p.ncols && (p.ncols & 3) == 0) {
#pragma unroll
                for (int c = 0; c < 8; ++c)
                  stg_stream4(orow + 4 * c, make_float4(__uint_as_float(r[4 * c]), __uint_as_float(r[4 * c + 1]),
                                                        __uint_as_float(r[4 * c + 2]), __uint_as_float(r[4 * c + 3])));
              } else {
                for (int c = 0; c < 32 && n0 + c < p.ncols; ++c) orow[c] = __uint_as_float(r[c]);
              }
            }
            continue;
          }
          // D[i, n] -> out[bh, n, i]: for every n the warp writes 32 consecutive queries (128 bytes)
#pragma unroll
          for (int c = 0; c < 32; ++c) {
            const int n = sl * 32 + c;
            if (ok) {
              const size_t o = ((size_t)bh * p.d + n) * p.Nq + i;
              const float a = __uint_as_float(r[c]);
              // out = fmap + gamma * out  (gma.py:113): one rounding per op like the reference
              const float val = p.residual ? fadd(__ldg(p.residual + o), fmul(g, a)) : (p.gamma ? fmul(g, a) : a);
              stg_stream(p.out + o, val);
            }
          }
        }
      }
      ptx::tc_fence_before_sync();
      __syncwarp();
      if (lane == 0) ptx::mbar_arrive(bar_t_empty + 8 * acc);
      if (++acc == kAccSets) { acc = 0; acc_par ^= 1; }
    }
  }

  ptx::tc_fence_before_sync();
  __syncthreads();
  if (warp == 2) {
    ptx::tc_fence_after_sync();
    ptx::tmem_dealloc(tmem_base, kGTmemCols);
  }
}

}  // namespace sb

extern "C" int sb_softmax_rows(float* x, long long rows, int n, long long row_stride, int to_tf32,
                               sb_stream_t stream) {
  using namespace sb;
  SB_ENTER();
  SB_REQUIRE(rows >= 0 && n >= 0 && row_stride >= n, SB_EINVAL, "sb_softmax_rows: bad size");
  if (rows == 0 || n == 0) return SB_OK;
  SB_REQUIRE(x != nullptr, SB_EINVAL, "sb_softmax_rows: null pointer");
  SB_REQUIRE((n & 3) == 0 && (row_stride & 3) == 0 && aligned16(x), SB_EUNSUP,
             "sb_softmax_rows: n and row_stride must be multiples of 4 floats, x 16-byte aligned");
  SB_REQUIRE(n <= 256 * 4 * kSoftmaxVec, SB_EUNSUP, "sb_softmax_rows: rows longer than %d", 256 * 4 * kSoftmaxVec);
  long long blocks = rows < (long long)kNumSMs * 8 ? rows : (long long)kNumSMs * 8;
  softmax_rows_kernel<false><<<(int)blocks, 256, 0, as_stream(stream)>>>(x, rows, n, row_stride, to_tf32, nullptr);
  SB_LAUNCH_CHECK("softmax_rows_kernel");
  return SB_OK;
}

extern "C" int sb_softmax_rows_bf16(const float* x, void* out_bf16, long long rows, int n, long long row_stride,
                                    sb_stream_t stream) {
  using namespace sb;
  SB_ENTER();
  SB_REQUIRE(rows >= 0 && n >= 0 && row_stride >= n, SB_EINVAL, "sb_softmax_rows_bf16: bad size");
  if (rows == 0 || n == 0) return SB_OK;
  SB_REQUIRE(x && out_bf16, SB_EINVAL, "sb_softmax_rows_bf16: null pointer");
  SB_REQUIRE((n & 3) == 0 && (row_stride & 3) == 0 && aligned16(x) && (reinterpret_cast<uintptr_t>(out_bf16) & 7) == 0,
             SB_EUNSUP, "sb_softmax_rows_bf16: n and row_stride must be multiples of 4, x 16-byte / out 8-byte aligned");
  SB_REQUIRE(n <= 256 * 4 * kSoftmaxVec, SB_EUNSUP, "sb_softmax_rows_bf16: rows longer than %d", 256 * 4 * kSoftmaxVec);
  long long blocks = rows < (long long)kNumSMs * 8 ? rows : (long long)kNumSMs * 8;
  softmax_rows_kernel<true><<<(int)blocks, 256, 0, as_stream(stream)>>>(const_cast<float*>(x), rows, n, row_stride, 0,
                                                                  static_cast<uint16_t*>(out_bf16));
  SB_LAUNCH_CHECK("softmax_rows_kernel");
  return SB_OK;
}

namespace sb {
static int launch_attn_v(const CUtensorMap& map_a, const CUtensorMap& map_b, AttnVParams& p, int mblk, bool bf16,
                         cudaStream_t stream) {
  p.dbg = debug_word_device();
  if (!p.dbg) return SB_ECUDA;
  static SmemOptIn opt_in;
  int opt_dev;
  if (opt_in.need(kGSmemTotal, &opt_dev)) {
    SB_CUDA(cudaFuncSetAttribute(attn_v_umma_kernel<1, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, kGSmemTotal));
    SB_CUDA(cudaFuncSetAttribute(attn_v_umma_kernel<2, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, kGSmemTotal));
    SB_CUDA(cudaFuncSetAttribute(attn_v_umma_kernel<4, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, kGSmemTotal));
    SB_CUDA(cudaFuncSetAttribute(attn_v_umma_kernel<1, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, kGSmemTotal));
    SB_CUDA(cudaFuncSetAttribute(attn_v_umma_kernel<2, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, kGSmemTotal));
    opt_in.done(kGSmemTotal, opt_dev);
  }
  const int grid = (int)(p.n_units < kNumSMs ? p.n_units : kNumSMs);
  if (bf16) {
    if (mblk >= 2) attn_v_umma_kernel<2, true><<<grid, 256, kGSmemTotal, stream>>>(map_a, map_b, p);
    else attn_v_umma_kernel<1, true><<<grid, 256, kGSmemTotal, stream>>>(map_a, map_b, p);
  } else if (mblk == 4) attn_v_umma_kernel<4, false><<<grid, 256, kGSmemTotal, stream>>>(map_a, map_b, p);
  else if (mblk == 2) attn_v_umma_kernel<2, false><<<grid, 256, kGSmemTotal, stream>>>(map_a, map_b, p);
  else attn_v_umma_kernel<1, false><<<grid, 256, kGSmemTotal, stream>>>(map_a, map_b, p);
  SB_LAUNCH_CHECK("attn_v_umma_kernel");
  return SB_OK;
}
}  // namespace sb

// D[bh] = A[bh] . B[bh]^T on the TF32 tensor cores: A [BH, M, K], B [BH, N, K] (both K-major, fp32),
// D [BH, M, N] fp32 row-major. Used by the CCL correlation (ccl.cu).
extern "C" int sb_gemm_nt_tf32(const float* A, const float* B, float* D, int BH, int M, int N, int K,
                               sb_stream_t stream) {
  using namespace sb;
  SB_ENTER();
  SB_REQUIRE(BH >= 0 && M >= 0 && N >= 0 && K > 0, SB_EINVAL, "sb_gemm_nt_tf32: bad size");
  if (BH == 0 || M == 0 || N == 0) return SB_OK;
  SB_REQUIRE(A && B && D, SB_EINVAL, "sb_gemm_nt_tf32: null pointer");
  SB_REQUIRE((K & 3) == 0 && aligned16(A) && aligned16(B) && aligned16(D), SB_EUNSUP,
             "sb_gemm_nt_tf32: K must be a multiple of 4 and A / B / D 16-byte aligned");
  CUtensorMap map_a, map_b;
  int rc = make_map_3d_ex(&map_a, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, A, (unsigned long long)K,
                          (unsigned long long)M, (unsigned long long)BH, GK, GM, CU_TENSOR_MAP_SWIZZLE_128B,
                          CU_TENSOR_MAP_L2_PROMOTION_L2_256B, "A");
  if (rc) return rc;
  rc = make_map_3d_ex(&map_b, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, B, (unsigned long long)K,
                      (unsigned long long)N, (unsigned long long)BH, GK, GN, CU_TENSOR_MAP_SWIZZLE_128B,
                      CU_TENSOR_MAP_L2_PROMOTION_L2_256B, "B");
  if (rc) return rc;
  AttnVParams p;
  p.BH = BH; p.Nq = M; p.Nk = K; p.d = N;
  p.MU = (M + GM - 1) / GM;
  p.KS = (K + GK - 1) / GK;
  p.NT = (N + GN - 1) / GN; p.ncols = N; p.row_major = 1;
  p.n_units = (long long)BH * p.MU * p.NT;
  p.residual = nullptr; p.gamma = nullptr; p.out = D;
  return launch_attn_v(map_a, map_b, p, 1, false, as_stream(stream));
}

namespace sb {
static int attn_aggregate_impl(const void* attn, const void* v, const float* residual, const float* gamma,
                               float* out, int BH, int Nq, int Nk, int d, bool bf16, sb_stream_t stream) {
  SB_ENTER();
  SB_REQUIRE(BH >= 0 && Nq >= 0 && Nk >= 0 && d >= 0, SB_EINVAL, "sb_attn_aggregate: bad size");
  if (BH == 0 || Nq == 0 || d == 0) return SB_OK;
  SB_REQUIRE(attn && v && out, SB_EINVAL, "sb_attn_aggregate: null pointer");
  SB_REQUIRE(Nk > 0, SB_EINVAL, "sb_attn_aggregate: no keys");
  SB_REQUIRE(d == GN, SB_EUNSUP, "sb_attn_aggregate: dim_head must be %d (got %d)", GN, d);
  SB_REQUIRE((Nk & (bf16 ? 7 : 3)) == 0 && aligned16(attn) && aligned16(v), SB_EUNSUP,
             "sb_attn_aggregate: Nk must be a multiple of %d and attn / v 16-byte aligned (TMA strides)", bf16 ? 8 : 4);
  const CUtensorMapDataType dt = bf16 ? CU_TENSOR_MAP_DATA_TYPE_BFLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT32;
  const int esz = bf16 ? 2 : 4, kstep = bf16 ? 64 : GK;
  CUtensorMap map_a, map_b;
  int rc = make_map_3d_ex(&map_a, dt, esz, attn, (unsigned long long)Nk,
                          (unsigned long long)Nq, (unsigned long long)BH, kstep, GM, CU_TENSOR_MAP_SWIZZLE_128B,
                          CU_TENSOR_MAP_L2_PROMOTION_L2_256B, "attn");
  if (rc) return rc;
  rc = make_map_3d_ex(&map_b, dt, esz, v, (unsigned long long)Nk,
                      (unsigned long long)d, (unsigned long long)BH, kstep, GN, CU_TENSOR_MAP_SWIZZLE_128B,
                      CU_TENSOR_MAP_L2_PROMOTION_L2_256B, "v");
  if (rc) return rc;
  const int MB = (Nq + GM - 1) / GM;
  // query blocks per CTA: 2 share every v stage when that still leaves a unit per SM (measured at
  // B=16, N=4096: MBLK 1 / 2 / 4 = 217 / 215 / 237 us — 4 has a single accumulator set, so its
  // epilogue is exposed)
  int mblk = tune_get(SB_TUNE_AGG_MBLK, 0);
  if ((mblk != 1 && mblk != 2 && mblk != 4) || (bf16 && mblk == 4)) mblk = ((long long)BH * ((MB + 1) / 2) >= kNumSMs) ? 2 : 1;
  AttnVParams p;
  p.BH = BH; p.Nq = Nq; p.Nk = Nk; p.d = d;
  p.MU = (MB + mblk - 1) / mblk;
  p.KS = (Nk + kstep - 1) / kstep;
  p.NT = 1; p.ncols = d; p.row_major = 0;
  p.n_units = (long long)BH * p.MU;
  p.residual = residual; p.gamma = gamma; p.out = out;
  return launch_attn_v(map_a, map_b, p, mblk, bf16, as_stream(stream));
}
}  // namespace sb

extern "C" int sb_attn_aggregate(const float* attn, const float* v, const float* residual,
                                 const float* gamma, float* out, int BH, int Nq, int Nk, int d,
                                 sb_stream_t stream) {
  return sb::attn_aggregate_impl(attn, v, residual, gamma, out, BH, Nq, Nk, d, false, stream);
}

extern "C" int sb_attn_aggregate_bf16(const void* attn_bf16, const void* v_bf16, const float* residual,
                                      const float* gamma, float* out, int BH, int Nq, int Nk, int d,
                                      sb_stream_t stream) {
  return sb::attn_aggregate_impl(attn_bf16, v_bf16, residual, gamma, out, BH, Nq, Nk, d, true, stream);
}
