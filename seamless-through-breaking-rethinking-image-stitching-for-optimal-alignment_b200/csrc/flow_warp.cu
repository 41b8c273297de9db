// flow_warp.cu — W1: dense-flow backward warp.
// Replaces warp(x, flo) (core/warp_utils.py:71-80) = flow_to_warp (:54-69) +
// normalise (:74-75) + F.grid_sample(bilinear, zeros, align_corners=True) (:77).
//
// HBM-bound gather: per output pixel 2 flow floats are read once, C planes are
// gathered at 4 taps each (neighbouring pixels share cache lines, so DRAM
// traffic ~ 1x the source), C floats are written. Algorithmic bytes:
// (2*C + 2) * 4 per pixel.
//
// Launch shape: grid (W/32, H/8, B), block (32, 8): a warp is 32 consecutive
// pixels of one row (coalesced flow reads / output writes), a CTA a 32x8 tile
// (vertical tap reuse in L1), and no index divisions are needed (the first
// version spent ~450 instructions per pixel, mostly 64-bit div/mod).
#include "bilinear.cuh"

namespace sb {

template <int C_T, bool EXACT_RCP>
__global__ void __launch_bounds__(256)
flow_warp_kernel(const float* __restrict__ x, const float* __restrict__ flo,
                 const float* __restrict__ mul_mask, float* __restrict__ out,
                 float* __restrict__ overlap, int C_rt, int H, int W) {
  const int C = (C_T > 0) ? C_T : C_rt;
  const int px = blockIdx.x * 32 + threadIdx.x, py = blockIdx.y * 8 + threadIdx.y;
  if (px >= W || py >= H) return;
  const int b = blockIdx.z;
  const int plane = H * W;
  const int rem = py * W + px;
  const float denx = (float)max(W - 1, 1), deny = (float)max(H - 1, 1);
  // ATen: scaling_factor = float(size - 1) / 2   (x * 0.5 == x / 2 exactly)
  const float halfx = fmul((float)(W - 1), 0.5f), halfy = fmul((float)(H - 1), 0.5f);

  const float* fl = flo + (size_t)b * 2 * plane + rem;
  const float fx = ldg_stream(fl), fy = ldg_stream(fl + plane);
  const float m = mul_mask ? ldg_stream(mul_mask + (size_t)b * plane + rem) : 1.0f;
  // grid + flow (coords_grid is exact integers as float), then the round trip.
  GridTap tap;
  if (EXACT_RCP)
    tap.setup(grid_roundtrip_rcp(fadd((float)px, fx), denx, __frcp_rn(denx), halfx),
              grid_roundtrip_rcp(fadd((float)py, fy), deny, __frcp_rn(deny), halfy), H, W);
  else
    tap.setup(grid_roundtrip(fadd((float)px, fx), denx, halfx),
              grid_roundtrip(fadd((float)py, fy), deny, halfy), H, W);
  // two running 64-bit row pointers stepped by the plane stride: one add each per channel
  const float* pn = x + (size_t)b * C * plane + tap.off_nw;
  const float* ps = pn + W;
  float* po = out + (size_t)b * C * plane + rem;
  if (C_T > 0) {
    float v[C_T > 0 ? C_T : 1];
#pragma unroll
    for (int c = 0; c < C_T; ++c) {
      v[c] = tap.sample_rows(pn, ps);
      pn += plane; ps += plane;
    }
    if (C_T == 6 && overlap) {
      // flowHomoAdpater.py:171-174 on the UNMASKED warp: where(mean_c(mask) < 0.9, 1, 0)
      const float mean = fdiv(fadd(fadd(v[3 % C_T], v[4 % C_T]), v[5 % C_T]), 3.0f);
      stg_stream(overlap + (size_t)b * plane + rem, mean < 0.9f ? 1.0f : 0.0f);
    }
#pragma unroll
    for (int c = 0; c < C_T; ++c) {
      stg_stream(po, mul_mask ? fmul(v[c], m) : v[c]);
      po += plane;
    }
  } else {
    for (int c = 0; c < C; ++c) {
      const float v = tap.sample_rows(pn, ps);
      stg_stream(po, mul_mask ? fmul(v, m) : v);
      pn += plane; ps += plane; po += plane;
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
  SB_REQUIRE(B <= 65535 && (H + 7) / 8 <= 65535, SB_EUNSUP, "sb_flow_warp: B or H too large for one launch");
  const dim3 block(32, 8), grid((W + 31) / 32, (H + 7) / 8, B);
  cudaStream_t s = as_stream(stream);
  // the reciprocal restatement of the division by (size-1) is proven for integer sizes up to 2048
  const int exact_rcp = (W >= 2 && W <= 2048 && H >= 2 && H <= 2048) ? 1 : 0;
#define SB_FLOW_LAUNCH(CT)   do {                                                                                  \
    if (exact_rcp) flow_warp_kernel<CT, true><<<grid, block, 0, s>>>(x, flo, mul_mask, out, overlap, C, H, W);  \
    else flow_warp_kernel<CT, false><<<grid, block, 0, s>>>(x, flo, mul_mask, out, overlap, C, H, W);           \
  } while (0)
  switch (C) {
    case 1: SB_FLOW_LAUNCH(1); break;
    case 2: SB_FLOW_LAUNCH(2); break;
    case 3: SB_FLOW_LAUNCH(3); break;
    case 6: SB_FLOW_LAUNCH(6); break;
    default: SB_FLOW_LAUNCH(0); break;
  }
#undef SB_FLOW_LAUNCH
  SB_LAUNCH_CHECK("flow_warp_kernel");
  return SB_OK;
}
