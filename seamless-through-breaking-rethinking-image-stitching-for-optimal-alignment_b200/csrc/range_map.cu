// range_map.cu — W4: forward bilinear splat count ("range map") + occlusion.
// Replaces compute_range_map (core/warp_utils.py:114-175) and the 'wang' branch
// of compute_occlusion (:185-221).
//
// Each source pixel p adds its four bilinear weights at floor(p + flow) + {0,1}^2
// when that target is inside the image.  The reference accumulates with
// scatter_add_ (fp32 atomics on a GPU: order-dependent).  Here the weights are
// accumulated as 2^-32 fixed-point integers with 64-bit integer atomics, which
// is exact per weight down to 2^-33 and ORDER-INDEPENDENT, so the result is
// deterministic; a second pass converts to fp32 and applies the caller's
// clamp / invert / threshold.
//
// HBM + L2 atomics: 8 B/px flow read, 4 RED.64 per px (L2-resident accumulator),
// 8 B/px accumulator read + 4 B/px write in the finalise pass.
#include <limits.h>

#include "common.cuh"

namespace sb {

__global__ void __launch_bounds__(256)
range_splat_kernel(const float* __restrict__ flow, unsigned long long* __restrict__ accum, int H,
                   int W) {
  // grid (W/32, H/8, B), block (32, 8): no index divisions.  No early return: every lane takes part in the shuffles.
  const int px = blockIdx.x * 32 + threadIdx.x, py = blockIdx.y * 8 + threadIdx.y;
  const bool in_image = px < W && py < H;
  const int plane = H * W;
  const size_t boff = (size_t)blockIdx.z * plane;
  float cx = 0.0f, cy = 0.0f;
  if (in_image) {
    const float* fl = flow + 2 * boff + py * W + px;
    // coords = grid + flow (flow_to_warp, :54-69)
    cx = fadd((float)px, ldg_stream(fl));
    cy = fadd((float)py, ldg_stream(fl + plane));
  }
  const float fx = floorf(cx), fy = floorf(cy);
  const float ox = fsub(cx, fx), oy = fsub(cy, fy);        // coords_offset (:121)
  const bool live = in_image && (fx >= -1.0f && fx <= (float)W && fy >= -1.0f && fy <= (float)H);   // also NaN
  const int ix = live ? (int)fx : INT_MIN / 2, iy = live ? (int)fy : INT_MIN / 2;
  unsigned long long* acc = accum + boff;
  // Weights are 2^-32 fixed-point integers, so the accumulation is order independent and partial sums may be formed
  // anywhere.  With a smooth flow the lane to the left splats its east column (di = 1) exactly where this lane splats
  // its west column (di = 0): it hands the two values over by shuffle and skips its own atomics — two instead of
  // four L2 atomics per pixel (the kernel is bound by L2 atomic throughput: ncu round 1, lts 57 %, DRAM 19 %).
  const int ix_l = __shfl_up_sync(0xffffffffu, ix, 1), iy_l = __shfl_up_sync(0xffffffffu, iy, 1);
  const bool take = live && threadIdx.x > 0 && ix_l + 1 == ix && iy_l == iy;      // I add my left neighbour's east column
  const bool given = __shfl_down_sync(0xffffffffu, take ? 1 : 0, 1) != 0 && threadIdx.x < 31;   // my east column is taken over
#pragma unroll
  for (int dj = 0; dj < 2; ++dj) {
    const int ty = iy + dj;
    const bool row_ok = live && ty >= 0 && ty < H;
    // weights_i = (1 - di) - (-1)^di * off_x ; weights_j likewise (:158-160)
    const float wj = dj ? fsub(0.0f, fmul(-1.0f, oy)) : fsub(1.0f, oy);
    const float w0 = fmul(fsub(1.0f, ox), wj), w1 = fmul(fsub(0.0f, fmul(-1.0f, ox)), wj);
    // w in [0, 1]: scale by 2^32 exactly (power of two), round to integer
    unsigned long long q0 = (row_ok && ix >= 0 && ix < W) ? __float2ull_rn(w0 * 4294967296.0f) : 0ull;
    const unsigned long long q1 = (row_ok && ix + 1 >= 0 && ix + 1 < W) ? __float2ull_rn(w1 * 4294967296.0f) : 0ull;
    const unsigned long long q1_l = __shfl_up_sync(0xffffffffu, q1, 1);
    if (take) q0 += q1_l;
    if (q0) atomicAdd(acc + ty * W + ix, q0);
    if (q1 && !given) atomicAdd(acc + ty * W + ix + 1, q1);
  }
}

__device__ __forceinline__ float range_finalize_one(unsigned long long a, int mode) {
  // exact: (double)a is exact below 2^53, the scale is a power of two, one rounding to fp32
  float v = (float)((double)a * (1.0 / 4294967296.0));
  if (mode >= 1) {
    const float c = fminf(fmaxf(v, 0.0f), 1.0f);
    // compute_occlusion: occ = 1 - clamp(range); occlusion_are_zeros: 1 - occ  (:213-220)
    const float occ = fsub(1.0f, c);
    if (mode == 1) v = fsub(1.0f, occ);
    else if (mode == 2) v = occ;
    else v = (fsub(1.0f, occ) >= 0.5f) ? 1.0f : 0.0f;
  }
  return v;
}

// Four accumulators per thread (two 16-byte loads, one 16-byte store): the one-element-per-thread version moved 50 MB
// in 19 us.  `total` need not be a multiple of 4 (tail handled by the last thread's scalar loop).
__global__ void __launch_bounds__(256)
range_finalize_kernel(const unsigned long long* __restrict__ accum, float* __restrict__ out,
                      int mode, long long total) {
  const long long nvec = total >> 2;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < nvec;
       i += (long long)gridDim.x * blockDim.x) {
    const ulonglong2 a0 = *reinterpret_cast<const ulonglong2*>(accum + 4 * i);
    const ulonglong2 a1 = *reinterpret_cast<const ulonglong2*>(accum + 4 * i + 2);
    float4 v;
    v.x = range_finalize_one(a0.x, mode); v.y = range_finalize_one(a0.y, mode);
    v.z = range_finalize_one(a1.x, mode); v.w = range_finalize_one(a1.y, mode);
    *reinterpret_cast<float4*>(out + 4 * i) = v;
  }
  if (blockIdx.x == 0 && threadIdx.x == 0)
    for (long long i = nvec << 2; i < total; ++i) out[i] = range_finalize_one(accum[i], mode);
}

}  // namespace sb

extern "C" int sb_range_map(const float* flow, unsigned long long* accum, float* range_map, int B,
                            int H, int W, int mode, sb_stream_t stream) {
  using namespace sb;
  SB_ENTER();

  SB_REQUIRE(B >= 0 && H >= 0 && W >= 0, SB_EINVAL, "sb_range_map: bad size");
  SB_REQUIRE(mode >= 0 && mode <= 3, SB_EINVAL, "sb_range_map: mode %d", mode);
  SB_REQUIRE((long long)H * W < (1ll << 31), SB_EUNSUP, "sb_range_map: plane too large");
  const long long total = (long long)B * H * W;
  if (total == 0) return SB_OK;
  SB_REQUIRE(flow && accum && range_map, SB_EINVAL, "sb_range_map: null pointer");
  cudaStream_t s = as_stream(stream);
  SB_CUDA(cudaMemsetAsync(accum, 0, (size_t)total * sizeof(unsigned long long), s));
  long long blocks = ((total >> 2) + 255) / 256;            // four elements per thread in the finalise pass
  const long long max_blocks = (long long)kNumSMs * 8 * 16;
  if (blocks > max_blocks) blocks = max_blocks;
  if (blocks < 1) blocks = 1;
  SB_REQUIRE(aligned16(accum) && aligned16(range_map), SB_EINVAL, "sb_range_map: accum and range_map must be 16-byte aligned");
  SB_REQUIRE(B <= 65535 && (H + 7) / 8 <= 65535, SB_EUNSUP, "sb_range_map: B or H too large for one launch");
  range_splat_kernel<<<dim3((W + 31) / 32, (H + 7) / 8, B), dim3(32, 8), 0, s>>>(flow, accum, H, W);
  SB_LAUNCH_CHECK("range_splat_kernel");
  range_finalize_kernel<<<(int)blocks, 256, 0, s>>>(accum, range_map, mode, total);
  SB_LAUNCH_CHECK("range_finalize_kernel");
  return SB_OK;
}
