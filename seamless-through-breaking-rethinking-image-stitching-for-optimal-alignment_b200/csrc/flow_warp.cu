// flow_warp.cu — W1: dense-flow backward warp.
// Replaces warp(x, flo) (core/warp_utils.py:71-80) = flow_to_warp (:54-69) +
// normalise (:74-75) + F.grid_sample(bilinear, zeros, align_corners=True) (:77).
//
// HBM-bound gather: per output pixel 2 flow floats are read once, C planes are
// gathered at 4 taps each (neighbouring pixels share cache lines, so DRAM
// traffic ~ 1x the source), C floats are written. Algorithmic bytes:
// (2*C + 2) * 4 per pixel.
#include "bilinear.cuh"

namespace sb {

template <int C_T>
__global__ void __launch_bounds__(256)
flow_warp_kernel(const float* __restrict__ x, const float* __restrict__ flo,
                 const float* __restrict__ mul_mask, float* __restrict__ out,
                 float* __restrict__ overlap, int C_rt, int H, int W,
                 long long total /* B*H*W */) {
  const int C = (C_T > 0) ? C_T : C_rt;
  const long long plane = (long long)H * W;
  const float denx = (float)max(W - 1, 1), deny = (float)max(H - 1, 1);
  // ATen: scaling_factor = float(size - 1) / 2
  const float halfx = fdiv((float)(W - 1), 2.0f), halfy = fdiv((float)(H - 1), 2.0f);

  for (long long p = blockIdx.x * (long long)blockDim.x + threadIdx.x; p < total;
       p += (long long)gridDim.x * blockDim.x) {
    const long long b = p / plane;
    const long long rem = p - b * plane;
    const int py = (int)(rem / W), px = (int)(rem - (long long)py * W);
    const float* fl = flo + b * 2 * plane + rem;
    const float fx = ldg_stream(fl), fy = ldg_stream(fl + plane);
    // grid + flow (coords_grid is exact integers as float), then the round trip.
    const float ix = grid_roundtrip(fadd((float)px, fx), denx, halfx);
    const float iy = grid_roundtrip(fadd((float)py, fy), deny, halfy);
    GridTap tap;
    tap.setup(ix, iy, H, W);
    const float m = mul_mask ? ldg_stream(mul_mask + b * plane + rem) : 1.0f;
    const float* src = x + b * C * plane;
    float* dst = out + b * C * plane + rem;
    if (C_T > 0) {
      float v[C_T > 0 ? C_T : 1];
#pragma unroll
      for (int c = 0; c < C_T; ++c) v[c] = tap.sample(src + c * plane, W);
      if (C_T == 6 && overlap) {
        // flowHomoAdpater.py:171-174 on the UNMASKED warp: where(mean_c(mask) < 0.9, 1, 0)
        const float mean = fdiv(fadd(fadd(v[3 % C_T], v[4 % C_T]), v[5 % C_T]), 3.0f);
        stg_stream(overlap + p, mean < 0.9f ? 1.0f : 0.0f);
      }
#pragma unroll
      for (int c = 0; c < C_T; ++c) stg_stream(dst + c * plane, mul_mask ? fmul(v[c], m) : v[c]);
    } else {
      for (int c = 0; c < C; ++c) {
        float v = tap.sample(src + c * plane, W);
        stg_stream(dst + c * plane, mul_mask ? fmul(v, m) : v);
      }
    }
  }
}

}  // namespace sb

extern "C" int sb_flow_warp(const float* x, const float* flo, const float* mul_mask, float* out,
                            float* overlap, int B, int C, int H, int W, sb_stream_t stream) {
  using namespace sb;
  SB_ENTER();
  
  SB_REQUIRE(B >= 0 && C >= 0 && H >= 0 && W >= 0, SB_EINVAL, "sb_flow_warp: negative size");
  SB_REQUIRE((long long)H * W < (1ll << 31), SB_EUNSUP, "sb_flow_warp: plane too large");
  SB_REQUIRE(!overlap || C == 6, SB_EINVAL, "sb_flow_warp: overlap output needs C == 6 (image | mask)");
  const long long total = (long long)B * H * W;
  if (total == 0 || C == 0) return SB_OK;
  SB_REQUIRE(x && flo && out, SB_EINVAL, "sb_flow_warp: null pointer");
  const int threads = 256;
  long long blocks = (total + threads - 1) / threads;
  const long long max_blocks = (long long)kNumSMs * 8 * 16;
  if (blocks > max_blocks) blocks = max_blocks;
  cudaStream_t s = as_stream(stream);
  switch (C) {
    case 1: flow_warp_kernel<1><<<(int)blocks, threads, 0, s>>>(x, flo, mul_mask, out, overlap, C, H, W, total); break;
    case 2: flow_warp_kernel<2><<<(int)blocks, threads, 0, s>>>(x, flo, mul_mask, out, overlap, C, H, W, total); break;
    case 3: flow_warp_kernel<3><<<(int)blocks, threads, 0, s>>>(x, flo, mul_mask, out, overlap, C, H, W, total); break;
    case 6: flow_warp_kernel<6><<<(int)blocks, threads, 0, s>>>(x, flo, mul_mask, out, overlap, C, H, W, total); break;
    default: flow_warp_kernel<0><<<(int)blocks, threads, 0, s>>>(x, flo, mul_mask, out, overlap, C, H, W, total); break;
  }
  SB_LAUNCH_CHECK("flow_warp_kernel");
  return SB_OK;
}
