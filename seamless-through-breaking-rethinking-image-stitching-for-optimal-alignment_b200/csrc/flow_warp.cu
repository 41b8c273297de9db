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

// Per-launch constants computed once on the host (IEEE fp32, the same values the device would compute).
struct FlowWarpConst {
  float denx, deny;        // max(size - 1, 1)
  float rcpx, rcpy;        // RN(1 / den)           (EXACT_RCP only)
  float halfx, halfy;      // (size - 1) * 0.5      ATen: scaling_factor = float(size - 1) / 2
};

// The kernel was issue-bound (338 instructions per pixel, ncu round 1): the per-thread reciprocal, a predicate and a
// zero-initialised destination per tap, and re-derived 64-bit addresses per tap and channel cost more issue slots
// than the 24 loads themselves.  Now: launch constants arrive as arguments; a warp whose 32 pixels have all four
// taps inside the image (the interior: > 99 % of the warps) takes a path with plain loads; both paths step two
// 64-bit BYTE pointers by the plane stride (two adds per channel) and address the east taps as +4 immediates.
template <int C_T, bool EXACT_RCP>
__global__ void __launch_bounds__(256)
flow_warp_kernel(const float* __restrict__ x, const float* __restrict__ flo,
                 const float* __restrict__ mul_mask, float* __restrict__ out,
                 float* __restrict__ overlap, int C_rt, int H, int W, const FlowWarpConst k) {
  const int C = (C_T > 0) ? C_T : C_rt;
  const int px = blockIdx.x * 32 + threadIdx.x, py = blockIdx.y * 8 + threadIdx.y;
  if (px >= W || py >= H) return;
  const int b = blockIdx.z;
  const int plane = H * W;
  const int rem = py * W + px;

  const float* fl = flo + (size_t)b * 2 * plane + rem;
  const float fx = ldg_stream(fl), fy = ldg_stream(fl + plane);
  const float m = mul_mask ? ldg_stream(mul_mask + (size_t)b * plane + rem) : 1.0f;
  // grid + flow (coords_grid is exact integers as float), then the round trip.
  GridTap tap;
  if (EXACT_RCP)
    tap.setup(grid_roundtrip_rcp(fadd((float)px, fx), k.denx, k.rcpx, k.halfx),
              grid_roundtrip_rcp(fadd((float)py, fy), k.deny, k.rcpy, k.halfy), H, W);
  else
    tap.setup(grid_roundtrip(fadd((float)px, fx), k.denx, k.halfx),
              grid_roundtrip(fadd((float)py, fy), k.deny, k.halfy), H, W);
  // two running 64-bit byte pointers stepped by the plane stride: one 64-bit add each per channel
  const size_t plane_bytes = (size_t)plane * sizeof(float);
  const char* pn = reinterpret_cast<const char*>(x) + ((size_t)b * C * plane + tap.off_nw) * sizeof(float);
  const char* ps = pn + (size_t)W * sizeof(float);
  char* po = reinterpret_cast<char*>(out) + ((size_t)b * C * plane + rem) * sizeof(float);
  const bool interior = __all_sync(__activemask(), tap.m_nw & tap.m_ne & tap.m_sw & tap.m_se);
  auto ld = [](const char* p, int byte_off) { return __ldg(reinterpret_cast<const float*>(p + byte_off)); };
  if (C_T > 0) {
    float v[C_T > 0 ? C_T : 1];
    if (interior) {
#pragma unroll
      for (int c = 0; c < C_T; ++c) {
        v[c] = tap.combine(ld(pn, 0), ld(pn, 4), ld(ps, 0), ld(ps, 4));
        pn += plane_bytes; ps += plane_bytes;
      }
    } else {
#pragma unroll
      for (int c = 0; c < C_T; ++c) {
        v[c] = tap.sample_rows(reinterpret_cast<const float*>(pn), reinterpret_cast<const float*>(ps));
        pn += plane_bytes; ps += plane_bytes;
      }
    }
    if (C_T == 6 && overlap) {
      // flowHomoAdpater.py:171-174 on the UNMASKED warp: where(mean_c(mask) < 0.9, 1, 0); the IEEE division by 3 is
      // the exact reciprocal restatement (tests/test_div_restatement.py covers every divisor <= 2047)
      const float mean = div_small_int(fadd(fadd(v[3 % C_T], v[4 % C_T]), v[5 % C_T]), 3.0f, 0.3333333432674407958984375f);
      stg_stream(overlap + (size_t)b * plane + rem, mean < 0.9f ? 1.0f : 0.0f);
    }
#pragma unroll
    for (int c = 0; c < C_T; ++c) {
      stg_stream(reinterpret_cast<float*>(po), mul_mask ? fmul(v[c], m) : v[c]);
      po += plane_bytes;
    }
  } else {
    for (int c = 0; c < C; ++c) {
      const float v = interior ? tap.combine(ld(pn, 0), ld(pn, 4), ld(ps, 0), ld(ps, 4))
                               : tap.sample_rows(reinterpret_cast<const float*>(pn), reinterpret_cast<const float*>(ps));
      stg_stream(reinterpret_cast<float*>(po), mul_mask ? fmul(v, m) : v);
      pn += plane_bytes; ps += plane_bytes; po += plane_bytes;
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
  FlowWarpConst k;
  k.denx = (float)(W - 1 > 1 ? W - 1 : 1); k.deny = (float)(H - 1 > 1 ? H - 1 : 1);
  k.rcpx = 1.0f / k.denx; k.rcpy = 1.0f / k.deny;                       // IEEE round-to-nearest, == __frcp_rn
  k.halfx = (float)(W - 1) * 0.5f; k.halfy = (float)(H - 1) * 0.5f;
#define SB_FLOW_LAUNCH(CT)   do {                                                                                     \
    if (exact_rcp) flow_warp_kernel<CT, true><<<grid, block, 0, s>>>(x, flo, mul_mask, out, overlap, C, H, W, k);  \
    else flow_warp_kernel<CT, false><<<grid, block, 0, s>>>(x, flo, mul_mask, out, overlap, C, H, W, k);           \
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
