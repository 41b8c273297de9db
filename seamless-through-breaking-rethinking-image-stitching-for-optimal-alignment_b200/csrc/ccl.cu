// ccl.cu — N3 ("next" row 3): the contextual correlation layer of the UDIS2 homography net.
// Replaces UDIS2Network.CCL(feature_1, feature_2)  (core/UDIS2/Homography/network.py:147-199):
//   nf = F.normalize(f, p=2, dim=1)                                                (:150-151)
//   match[q, p] = conv2d(nf1, 3x3 patches of nf2 as filters, padding=1)            (:153-164)
//               = sum_{c, d in 3x3} nf1[c, p + d] * nf2[c, q + d]   (zero outside either map)
//   prob = softmax(10 * match, over q)                                              (:166-168)
//   flow_h[p] = sum_q prob[q, p] * (q // w - p_y),  flow_w[p] = sum_q prob * (q % w - p_x)   (:170-197)
//   out = cat([flow_w, flow_h], 1)                                                  (:199)
//
// The reference evaluates the h*w x (9*c) x h*w contraction as a Python loop of per-sample conv2d
// calls (19.3 GFLOP per pair at 32 x 32 x 1024) and materialises match / softmax / three index
// volumes.  Here:  match[q, p] = sum_{d in 3x3} C0[p + d, q + d]  with the plain all-pairs
// correlation C0 = nf1^T nf2 (2.1 GFLOP: 9x fewer) computed once on the TF32 tensor cores
// (sb_gemm_nt_tf32, gma.cu) from normalised token-major copies; one kernel then sums the nine
// shifted diagonals, takes the softmax over q and the expected displacement per position p —
// neither the match volume nor the softmax ever exist in memory.
#include "common.cuh"

extern "C" int sb_gemm_nt_tf32(const float* A, const float* B, float* D, int BH, int M, int N, int K,
                               sb_stream_t stream);

namespace sb {

// denom[b, n] = max(||f[b, :, n]||_2, 1e-12)   (F.normalize's clamp_min(eps))
__global__ void __launch_bounds__(256)
ccl_norm_kernel(const float* __restrict__ f, float* __restrict__ denom, int C, int N) {
  const int n = blockIdx.x * 32 + (threadIdx.x & 31), b = blockIdx.y;
  const int cw = threadIdx.x >> 5;                 // 8 channel slices
  __shared__ float s_part[8][33];
  float acc = 0.0f;
  if (n < N)
    for (int c = cw; c < C; c += 8) {
      const float v = __ldg(f + ((size_t)b * C + c) * N + n);
      acc = __fmaf_rn(v, v, acc);
    }
  s_part[cw][threadIdx.x & 31] = acc;
  __syncthreads();
  if (cw == 0 && n < N) {
    float t = 0.0f;
#pragma unroll
    for (int k = 0; k < 8; ++k) t += s_part[k][threadIdx.x];
    denom[(size_t)b * N + n] = fmaxf(sqrtf(t), 1e-12f);
  }
}

__device__ __forceinline__ float to_tf32(float x) {
  uint32_t u;
  asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(u) : "f"(x));
  return __uint_as_float(u);
}

// tok[b, n, c] = tf32(f[b, c, n] / denom[b, n]): tiles of 128 channels x 32 positions; reads are
// 128-byte rows of one channel, writes 512-byte rows of one position (float4 per lane)
__global__ void __launch_bounds__(256)
ccl_tokens_kernel(const float* __restrict__ f, const float* __restrict__ denom, float* __restrict__ tok,
                  int C, int N) {
  __shared__ float s[128][33];
  const int b = blockIdx.z, c0 = blockIdx.y * 128, n0 = blockIdx.x * 32;
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
#pragma unroll 4
  for (int r = ty; r < 128; r += 8) {
    const int c = c0 + r, n = n0 + tx;
    s[r][tx] = (c < C && n < N) ? ldg_stream(f + ((size_t)b * C + c) * N + n) : 0.0f;
  }
  __syncthreads();
#pragma unroll
  for (int r = ty; r < 32; r += 8) {
    const int n = n0 + r, c = c0 + tx * 4;
    if (n < N && c < C) {                       // C % 4 == 0: a float4 is entirely in or out
      const float d = __ldg(denom + (size_t)b * N + n);
      float4 v;
      v.x = to_tf32(fdiv(s[tx * 4 + 0][r], d)); v.y = to_tf32(fdiv(s[tx * 4 + 1][r], d));
      v.z = to_tf32(fdiv(s[tx * 4 + 2][r], d)); v.w = to_tf32(fdiv(s[tx * 4 + 3][r], d));
      *reinterpret_cast<float4*>(tok + ((size_t)b * N + n) * C + c) = v;
    }
  }
}

// One WARP per position p (8 positions per CTA, no block-level synchronisation): the nine shifted
// rows of C0 are summed into a shared-memory row (lanes stride q, so every load is a coalesced
// 128-byte segment and neighbouring warps re-use each other's rows from L1), then softmax over q
// and the expected displacement with warp shuffles.
__global__ void __launch_bounds__(256)
ccl_flow_kernel(const float* __restrict__ c0, float* __restrict__ out, int H, int W, float scale) {
  extern __shared__ float s_rows[];             // [8][N]
  const int N = H * W;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int p = blockIdx.x * 8 + warp, b = blockIdx.y;
  if (p >= N) return;
  float* row = s_rows + (size_t)warp * N;
  const float inv_w = 1.0f / (float)W;
  const int py = __float2int_rd(((float)p + 0.5f) * inv_w), px = p - py * W;   // exact for p < 2^22
  const float* cb = c0 + (size_t)b * N * N;
  float mx = -INFINITY;
  for (int q = lane; q < N; q += 32) {
    const int qy = __float2int_rd(((float)q + 0.5f) * inv_w), qx = q - qy * W;
    float acc = 0.0f;
#pragma unroll
    for (int dy = -1; dy <= 1; ++dy)
#pragma unroll
      for (int dx = -1; dx <= 1; ++dx) {
        const bool ok = (unsigned)(py + dy) < (unsigned)H && (unsigned)(px + dx) < (unsigned)W &&
                        (unsigned)(qy + dy) < (unsigned)H && (unsigned)(qx + dx) < (unsigned)W;
        if (ok) acc += __ldg(cb + (size_t)(p + dy * W + dx) * N + (q + dy * W + dx));
      }
    acc *= scale;
    row[q] = acc;
    mx = fmaxf(mx, acc);
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
  float se = 0.0f, sh = 0.0f, sw = 0.0f;
  for (int q = lane; q < N; q += 32) {
    const int qy = __float2int_rd(((float)q + 0.5f) * inv_w), qx = q - qy * W;
    const float e = expf(row[q] - mx);
    se += e;
    sh = __fmaf_rn(e, (float)(qy - py), sh);
    sw = __fmaf_rn(e, (float)(qx - px), sw);
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    se += __shfl_xor_sync(0xffffffffu, se, o);
    sh += __shfl_xor_sync(0xffffffffu, sh, o);
    sw += __shfl_xor_sync(0xffffffffu, sw, o);
  }
  if (lane == 0) {
    out[((size_t)b * 2 + 0) * N + p] = sw / se;   // channel 0 = flow_w, 1 = flow_h  (:199)
    out[((size_t)b * 2 + 1) * N + p] = sh / se;
  }
}

static inline size_t align256(size_t x) { return (x + 255) / 256 * 256; }

}  // namespace sb

extern "C" size_t sb_ccl_workspace_bytes(int B, int C, int H, int W) {
  if (B < 0 || C < 0 || H < 0 || W < 0) return 0;
  const size_t N = (size_t)H * W;
  return 2 * sb::align256((size_t)B * N * C * 4) + 2 * sb::align256((size_t)B * N * 4) + sb::align256((size_t)B * N * N * 4);
}

extern "C" int sb_ccl(const float* feature_1, const float* feature_2, float* flow, void* workspace,
                      size_t workspace_bytes, int B, int C, int H, int W, float softmax_scale,
                      sb_stream_t stream) {
  using namespace sb;
  SB_ENTER();
  SB_REQUIRE(B >= 0 && C > 0 && H >= 0 && W >= 0, SB_EINVAL, "sb_ccl: bad size");
  const long long N = (long long)H * W;
  if (B == 0 || N == 0) return SB_OK;
  SB_REQUIRE(feature_1 && feature_2 && flow, SB_EINVAL, "sb_ccl: null pointer");
  SB_REQUIRE(N <= 4096, SB_EUNSUP, "sb_ccl: more than 4096 positions (8 rows of N floats must fit shared memory)");
  SB_REQUIRE((C & 3) == 0, SB_EUNSUP, "sb_ccl: C must be a multiple of 4");
  SB_REQUIRE(B <= 65535, SB_EUNSUP, "sb_ccl: B > 65535");
  const size_t need = sb_ccl_workspace_bytes(B, C, H, W);
  SB_REQUIRE(workspace && workspace_bytes >= need && (reinterpret_cast<uintptr_t>(workspace) & 255) == 0, SB_EINVAL,
             "sb_ccl: workspace must be 256-byte aligned and >= %zu bytes", need);
  char* ws = static_cast<char*>(workspace);
  float* tok1 = reinterpret_cast<float*>(ws); ws += align256((size_t)B * N * C * 4);
  float* tok2 = reinterpret_cast<float*>(ws); ws += align256((size_t)B * N * C * 4);
  float* den1 = reinterpret_cast<float*>(ws); ws += align256((size_t)B * N * 4);
  float* den2 = reinterpret_cast<float*>(ws); ws += align256((size_t)B * N * 4);
  float* c0 = reinterpret_cast<float*>(ws);
  cudaStream_t s = as_stream(stream);
  const dim3 gn((unsigned)((N + 31) / 32), B);
  const dim3 gt((unsigned)((N + 31) / 32), (unsigned)((C + 127) / 128), B);
  ccl_norm_kernel<<<gn, 256, 0, s>>>(feature_1, den1, C, (int)N);
  SB_LAUNCH_CHECK("ccl_norm_kernel");
  ccl_norm_kernel<<<gn, 256, 0, s>>>(feature_2, den2, C, (int)N);
  SB_LAUNCH_CHECK("ccl_norm_kernel");
  ccl_tokens_kernel<<<gt, 256, 0, s>>>(feature_1, den1, tok1, C, (int)N);
  SB_LAUNCH_CHECK("ccl_tokens_kernel");
  ccl_tokens_kernel<<<gt, 256, 0, s>>>(feature_2, den2, tok2, C, (int)N);
  SB_LAUNCH_CHECK("ccl_tokens_kernel");
  const int rc = sb_gemm_nt_tf32(tok1, tok2, c0, B, (int)N, (int)N, C, stream);
  if (rc) return rc;
  const size_t flow_smem = (size_t)8 * N * sizeof(float);
  static size_t flow_smem_set = 0;
  if (flow_smem > 48 * 1024 && flow_smem > flow_smem_set) {
    SB_CUDA(cudaFuncSetAttribute(ccl_flow_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)flow_smem));
    flow_smem_set = flow_smem;
  }
  ccl_flow_kernel<<<dim3((unsigned)((N + 7) / 8), B), 256, flow_smem, s>>>(c0, flow, H, W, softmax_scale);
  SB_LAUNCH_CHECK("ccl_flow_kernel");
  return SB_OK;
}
