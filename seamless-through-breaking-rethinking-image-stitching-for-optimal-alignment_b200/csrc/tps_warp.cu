// tps_warp.cu — W3: thin-plate-spline backward warp with the UDIS sampler.
// Replaces the dense part of transformer(U, source, target, out_size)
// (core/udis_utils/torch_tps_transform.py):
//   _meshgrid (:96-125): basis (1, x, y, r_1..r_pn), r_k = d2 * log(d2 + 1e-6),
//                        d2 = (x - px_k)^2 + (y - py_k)^2
//   _transform (:127-147): (x_s, y_s) = T @ basis
//   _interpolate (:18-94): UdisTap (bilinear.cuh)
// The reference materialises the [B, pn+3, H*W] basis (180 MB at 512^2, pn = 169)
// and multiplies with BLAS; here the basis is evaluated on the fly per pixel and
// never touches memory.  FP32-ALU/MUFU-bound for pn >~ 30 (one logf per pixel
// and control point), HBM traffic is only the 2*C*4 B/px of the sampler.
// The dot product is accumulated in fp64 (the BLAS summation order of the
// reference is unspecified; fp64 is the order-independent answer both orders
// approximate).
#include "bilinear.cuh"

namespace sb {

constexpr int kTpsMaxPn = 1024;

template <int C_T>
__global__ void __launch_bounds__(256)
tps_warp_kernel(const float* __restrict__ U, const float* __restrict__ T,
                const float* __restrict__ source, const float* __restrict__ xs,
                const float* __restrict__ ys, float* __restrict__ out,
                int32_t* __restrict__ idx_dbg, int C_rt, int H, int W, int Hout, int Wout,
                int pn, int blocks_per_image) {
  extern __shared__ float s_tps[];  // px[pn], py[pn], tx[pn+3], ty[pn+3]
  float* s_px = s_tps;
  float* s_py = s_px + pn;
  float* s_tx = s_py + pn;
  float* s_ty = s_tx + pn + 3;
  const int C = (C_T > 0) ? C_T : C_rt;
  const int b = blockIdx.x / blocks_per_image;
  const int blk = blockIdx.x - b * blocks_per_image;
  for (int k = threadIdx.x; k < pn; k += blockDim.x) {
    s_px[k] = __ldg(source + ((long long)b * pn + k) * 2);
    s_py[k] = __ldg(source + ((long long)b * pn + k) * 2 + 1);
  }
  for (int k = threadIdx.x; k < pn + 3; k += blockDim.x) {
    s_tx[k] = __ldg(T + ((long long)b * 2) * (pn + 3) + k);
    s_ty[k] = __ldg(T + ((long long)b * 2 + 1) * (pn + 3) + k);
  }
  __syncthreads();
  const long long oplane = (long long)Hout * Wout, iplane = (long long)H * W;
  for (long long rem = (long long)blk * blockDim.x + threadIdx.x; rem < oplane;
       rem += (long long)blocks_per_image * blockDim.x) {
    const int r = (int)(rem / Wout), c = (int)(rem - (long long)r * Wout);
    const float gx = __ldg(xs + c), gy = __ldg(ys + r);
    // basis order (ones, x, y, r_1 .. r_pn)   (:123)
    double ax = (double)s_tx[0] + (double)s_tx[1] * (double)gx + (double)s_tx[2] * (double)gy;
    double ay = (double)s_ty[0] + (double)s_ty[1] * (double)gx + (double)s_ty[2] * (double)gy;
#pragma unroll 4
    for (int k = 0; k < pn; ++k) {
      const float dx = fsub(gx, s_px[k]), dy = fsub(gy, s_py[k]);
      const float d2 = fadd(fmul(dx, dx), fmul(dy, dy));            // square + square (:115)
      const float rk = fmul(d2, logf(fadd(d2, 1e-6f)));             // (:116)
      ax = fma((double)s_tx[3 + k], (double)rk, ax);
      ay = fma((double)s_ty[3 + k], (double)rk, ay);
    }
    UdisTap tap;
    tap.setup((float)ax, (float)ay, H, W);
    if (idx_dbg) {
      int32_t* d = idx_dbg + (long long)b * 4 * oplane + rem;
      d[0] = tap.x0; d[oplane] = tap.x1; d[2 * oplane] = tap.y0; d[3 * oplane] = tap.y1;
    }
    const float* src = U + (long long)b * C * iplane;
    float* dst = out + (long long)b * C * oplane + rem;
    if (C_T > 0) {
      float v[C_T > 0 ? C_T : 1];
#pragma unroll
      for (int ch = 0; ch < C_T; ++ch) v[ch] = tap.sample(src + ch * iplane, W);
#pragma unroll
      for (int ch = 0; ch < C_T; ++ch) stg_stream(dst + ch * oplane, v[ch]);
    } else {
      for (int ch = 0; ch < C; ++ch) stg_stream(dst + ch * oplane, tap.sample(src + ch * iplane, W));
    }
  }
}

}  // namespace sb

extern "C" int sb_tps_warp(const float* U, const float* T, const float* source, const float* xs,
                           const float* ys, float* out, int32_t* idx_dbg, int B, int C, int H,
                           int W, int Hout, int Wout, int pn, sb_stream_t stream) {
  using namespace sb;
  SB_ENTER();
  
  SB_REQUIRE(B >= 0 && C >= 0 && H > 0 && W > 0 && Hout >= 0 && Wout >= 0 && pn >= 0, SB_EINVAL,
             "sb_tps_warp: bad size");
  SB_REQUIRE(pn <= kTpsMaxPn, SB_EUNSUP, "sb_tps_warp: pn=%d > %d control points", pn, kTpsMaxPn);
  SB_REQUIRE((long long)H * W < (1ll << 31) && (long long)Hout * Wout < (1ll << 31), SB_EUNSUP,
             "sb_tps_warp: plane too large");
  const long long oplane = (long long)Hout * Wout;
  if ((long long)B * oplane == 0 || C == 0) return SB_OK;
  SB_REQUIRE(U && T && source && xs && ys && out, SB_EINVAL, "sb_tps_warp: null pointer");
  const int threads = 256;
  long long bpi = (oplane + threads - 1) / threads;
  const long long cap = (long long)kNumSMs * 8 * 4 / (B > 0 ? B : 1) + 1;
  if (bpi > cap) bpi = cap;
  const size_t smem = (size_t)(4 * pn + 6) * sizeof(float);
  const int grid = (int)(bpi * B);
  cudaStream_t s = as_stream(stream);
#define SB_TPS_LAUNCH(CT)                                                                       \
  tps_warp_kernel<CT><<<grid, threads, smem, s>>>(U, T, source, xs, ys, out, idx_dbg, C, H, W, \
                                                  Hout, Wout, pn, (int)bpi)
  switch (C) {
    case 1: SB_TPS_LAUNCH(1); break;
    case 3: SB_TPS_LAUNCH(3); break;
    case 6: SB_TPS_LAUNCH(6); break;
    default: SB_TPS_LAUNCH(0); break;
  }
#undef SB_TPS_LAUNCH
  SB_LAUNCH_CHECK("tps_warp_kernel");
  return SB_OK;
}
