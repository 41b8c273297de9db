// bilinear.cuh — the two bilinear samplers of the reference, as device code.
//
//  (1) GridTap: F.grid_sample(bilinear, zeros padding, align_corners=True) as
//      evaluated by ATen's vectorised CPU kernel (the oracle's arithmetic):
//        x_w = floor(x); w = x - x_w; e = 1 - w; n = y - y_n; s = 1 - n
//        nw = s*e, ne = s*w, sw = n*e, se = n*w
//        out = fma(se_v, se, fma(sw_v, sw, fma(ne_v, ne, nw_v*nw)))   (OOB taps read 0)
//      used by warp (core/warp_utils.py:77) and bilinear_sampler
//      (core/utils/utils.py:70).
//
//  (2) UdisTap: the hand-rolled sampler of core/udis_utils/torch_homo_transform.py:17-92
//      (identical copy in torch_tps_transform.py:18-94): floor, +1, clamp both
//      corners to [0, size-1], weights from the CLAMPED corners, sum
//      ((wa*Ia + wb*Ib) + wc*Ic) + wd*Id.
#pragma once
#include "common.cuh"

namespace sb {

struct GridTap {
  float nw, ne, sw, se;     // weights
  int off_nw;               // y_n * W + x_w (may be out of range; use masks)
  int xi, yi;               // x_w, y_n saturated into [-2, W + 1] / [-2, H + 1] (NaN -> -2)
  bool m_nw, m_ne, m_sw, m_se;

  __device__ __forceinline__ void setup(float ix, float iy, int H, int W) {
    float x_w = floorf(ix), y_n = floorf(iy);
    float w = fsub(ix, x_w), e = fsub(1.0f, w);
    float n = fsub(iy, y_n), s = fsub(1.0f, n);
    nw = fmul(s, e); ne = fmul(s, w); sw = fmul(n, e); se = fmul(n, w);
    // Saturating float->int: anything outside [-1, size] is masked anyway.
    float xc = fminf(fmaxf(x_w, -2.0f), (float)W + 1.0f);
    float yc = fminf(fmaxf(y_n, -2.0f), (float)H + 1.0f);
    xi = (x_w == x_w) ? (int)xc : -2;   // NaN -> fully masked (weights stay NaN)
    yi = (y_n == y_n) ? (int)yc : -2;
    bool mw = (xi >= 0) & (xi < W), me = (xi + 1 >= 0) & (xi + 1 < W);
    bool mn = (yi >= 0) & (yi < H), ms = (yi + 1 >= 0) & (yi + 1 < H);
    m_nw = mn & mw; m_ne = mn & me; m_sw = ms & mw; m_se = ms & me;
    off_nw = yi * W + xi;
  }

  // plane: pointer to the [H, W] channel plane.
  __device__ __forceinline__ float sample(const float* __restrict__ plane, int W) const {
    const float* pn = plane + off_nw;     // one 64-bit address per row pair, +1 as an immediate
    const float* ps = pn + W;
    float v_nw = 0.0f, v_ne = 0.0f, v_sw = 0.0f, v_se = 0.0f;
    if (m_nw) v_nw = __ldg(pn);
    if (m_ne) v_ne = __ldg(pn + 1);
    if (m_sw) v_sw = __ldg(ps);
    if (m_se) v_se = __ldg(ps + 1);
    return combine(v_nw, v_ne, v_sw, v_se);
  }
  // pn / ps: the north / south tap rows of this pixel in one channel plane (plane + off_nw, + W);
  // callers step both by the plane stride from channel to channel (one 64-bit add each).
  __device__ __forceinline__ float sample_rows(const float* __restrict__ pn, const float* __restrict__ ps) const {
    float v_nw = 0.0f, v_ne = 0.0f, v_sw = 0.0f, v_se = 0.0f;
    if (m_nw) v_nw = __ldg(pn);
    if (m_ne) v_ne = __ldg(pn + 1);
    if (m_sw) v_sw = __ldg(ps);
    if (m_se) v_se = __ldg(ps + 1);
    return combine(v_nw, v_ne, v_sw, v_se);
  }
  __device__ __forceinline__ float combine(float v_nw, float v_ne, float v_sw, float v_se) const {
    // ATen's CPU build contracts the mul/add chain into FMAs (bit-exact vs F.grid_sample)
    return __fmaf_rn(v_se, se, __fmaf_rn(v_sw, sw, __fmaf_rn(v_ne, ne, fmul(v_nw, nw))));
  }
};

struct UdisTap {
  float wa, wb, wc, wd;
  int x0, x1, y0, y1;       // clamped integer grid indices

  // x, y: normalised sample position in [-1,1] (torch_homo_transform.py:29-30).
  __device__ __forceinline__ void setup(float xn, float yn, int H, int W) {
    float x = fmul(fmul(fadd(xn, 1.0f), (float)W), 0.5f);   // * 0.5 == / 2.0 exactly
    float y = fmul(fmul(fadd(yn, 1.0f), (float)H), 0.5f);
    // torch: floor(x).int() — saturate so that garbage coordinates (|x| huge,
    // NaN) still clamp into the image like the reference's int32 cast + clamp.
    float xf = floorf(x), yf = floorf(y);
    int xi = f2i(xf), yi = f2i(yf);
    x0 = min(max(xi, 0), W - 1);
    x1 = min(max(xi + 1, 0), W - 1);   // xi < 2^31 - 127, so xi + 1 cannot overflow
    y0 = min(max(yi, 0), H - 1);
    y1 = min(max(yi + 1, 0), H - 1);
    float x0f = (float)x0, x1f = (float)x1, y0f = (float)y0, y1f = (float)y1;
    wa = fmul(fsub(x1f, x), fsub(y1f, y));
    wb = fmul(fsub(x1f, x), fsub(y, y0f));
    wc = fmul(fsub(x, x0f), fsub(y1f, y));
    wd = fmul(fsub(x, x0f), fsub(y, y0f));
  }
  static __device__ __forceinline__ int f2i(float f) {
    // x86 cvttps2dq semantics of the oracle: NaN / out-of-range -> INT_MIN.
    if (!(f >= -2147483648.0f && f < 2147483648.0f)) return INT_MIN;
    return (int)f;
  }
  // Ia=(y0,x0) Ib=(y1,x0) Ic=(y0,x1) Id=(y1,x1)   (torch_homo_transform.py:54-59)
  __device__ __forceinline__ float sample(const float* __restrict__ plane, int W) const {
    float Ia = __ldg(plane + y0 * W + x0);
    float Ib = __ldg(plane + y1 * W + x0);
    float Ic = __ldg(plane + y0 * W + x1);
    float Id = __ldg(plane + y1 * W + x1);
    return fadd(fadd(fadd(fmul(wa, Ia), fmul(wb, Ib)), fmul(wc, Ic)), fmul(wd, Id));
  }
};

}  // namespace sb
