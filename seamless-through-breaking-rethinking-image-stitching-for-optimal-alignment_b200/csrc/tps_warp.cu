// tps_warp.cu — W3: thin-plate-spline backward warp with the UDIS sampler.
// Replaces the dense part of transformer(U, source, target, out_size)
// (core/udis_utils/torch_tps_transform.py):
//   _meshgrid (:96-125): basis (1, x, y, r_1..r_pn), r_k = d2 * log(d2 + 1e-6),
//                        d2 = (x - px_k)^2 + (y - py_k)^2
//   _transform (:127-147): (x_s, y_s) = T @ basis
//   _interpolate (:18-94): UdisTap (bilinear.cuh)
// The reference materialises the [B, pn+3, H*W] basis (180 MB at 512^2, pn = 169)
// and multiplies with BLAS; here the basis is evaluated on the fly per pixel and
// never touches memory: HBM traffic is only the 2*C*4 B/px of the sampler and the
// kernel is bound by the SM's special-function unit (one lg2 per pixel and control point;
// 16 lanes/clk/SM = 8 issue cycles per warp).
//
// Inner-loop budget per pixel and control point (ncu, profiles/): the first version spent three
// F2F.F64.F32 conversions (the same 16-lane unit as lg2) and the ~22-instruction libdevice logf
// and ran at 1.0 ms for 16 x 512^2 x 169.  Now
//   * each thread carries TWO pixels as the lanes of packed fp32 instructions (sm_100 FADD2 /
//     FMUL2 / FFMA2): every operation of the reference's elementwise chain keeps its own
//     round-to-nearest, two pixels per issue slot;
//   * control points and weights sit in shared memory already negated / duplicated per lane,
//     so one LDS.128 feeds each packed operand pair;
//   * log() is lg2.approx * ln2 (|err| <= 2^-21.4 around 1, 2 ulp elsewhere: below the
//     rounding of d2 * log itself; SB_TUNE_TPS_LOG = 1 selects libdevice logf instead);
//   * the sum over control points runs in fp32 FFMA over groups of 8 and the group sums are
//     added in fp64 (the reference's BLAS sums all pn + 3 terms in fp32 in an unspecified order;
//     this bounds the order-dependent part to 8 terms and costs 1/4 conversion per evaluation).
#include "bilinear.cuh"

namespace sb {

constexpr int kTpsMaxPn = 1024;
constexpr int kTpsThreads = 256;
constexpr int kTpsPix = 2;                        // pixels per thread = lanes of the packed ops
constexpr int kTpsChunk = kTpsThreads * kTpsPix;  // pixels per CTA step
constexpr int kTpsGroup = 8;                      // control points per fp32 partial sum

__host__ __device__ constexpr int tps_padded(int pn) { return (pn + kTpsGroup - 1) / kTpsGroup * kTpsGroup; }

// a + b where a is the result of a packed multiply.  ptxas 12.9 contracts mul.rn.f32x2 + add.rn.f32x2
// into FFMA2 even under --fmad=false (the scalar forms are left alone), and it also folds a literal
// fma(a, 1, b) back into that.  `one` is therefore a kernel ARGUMENT (always 1.0f): fma(a, one, b) is the
// same single rounding as the add, one issue slot like FADD2, and nothing can be contracted into it.
__device__ __forceinline__ float2 add2_unfused(float2 a, float2 b, float2 one) { return __ffma2_rn(a, one, b); }

template <int LOGMODE>
__device__ __forceinline__ float2 tps_log2(float2 x) {   // x >= 1e-8: positive, normal
  if (LOGMODE == 1) return make_float2(logf(x.x), logf(x.y));
  float2 l2;
  asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(l2.x) : "f"(x.x));
  asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(l2.y) : "f"(x.y));
  return __fmul2_rn(l2, make_float2(0.693147180559945f, 0.693147180559945f));
}

template <int C_T, int LOGMODE>
__global__ void __launch_bounds__(kTpsThreads)
tps_warp_kernel(const float* __restrict__ U, const float* __restrict__ T,
                const float* __restrict__ source, const float* __restrict__ xs,
                const float* __restrict__ ys, float* __restrict__ out,
                int32_t* __restrict__ idx_dbg, int C_rt, int H, int W, int Hout, int Wout,
                int pn, int blocks_per_image, float one_arg) {
  // per control point (padded to a multiple of 8 with zero weights):
  //   s_p[k] = (-px, -px, -py, -py)     s_w[k] = (tx, tx, ty, ty)        [lane-duplicated]
  // then the affine part of T (ones, x, y rows) as fp64 pairs.
  extern __shared__ float4 s_tps[];
  const int pnp = tps_padded(pn);
  float4* s_p = s_tps;
  float4* s_w = s_tps + pnp;
  double2* s_aff = reinterpret_cast<double2*>(s_tps + 2 * pnp);
  const int C = (C_T > 0) ? C_T : C_rt;
  const int b = blockIdx.x / blocks_per_image;
  const int blk = blockIdx.x - b * blocks_per_image;
  for (int k = threadIdx.x; k < pnp; k += blockDim.x) {
    float px = 0.f, py = 0.f, tx = 0.f, ty = 0.f;
    if (k < pn) {
      px = __ldg(source + ((long long)b * pn + k) * 2);
      py = __ldg(source + ((long long)b * pn + k) * 2 + 1);
      tx = __ldg(T + ((long long)b * 2) * (pn + 3) + 3 + k);
      ty = __ldg(T + ((long long)b * 2 + 1) * (pn + 3) + 3 + k);
    }
    s_p[k] = make_float4(-px, -px, -py, -py);
    s_w[k] = make_float4(tx, tx, ty, ty);
  }
  if (threadIdx.x < 3)
    s_aff[threadIdx.x] = make_double2((double)__ldg(T + ((long long)b * 2) * (pn + 3) + threadIdx.x),
                                      (double)__ldg(T + ((long long)b * 2 + 1) * (pn + 3) + threadIdx.x));
  __syncthreads();
  const long long oplane = (long long)Hout * Wout, iplane = (long long)H * W;
  const float2 eps = make_float2(1e-6f, 1e-6f), one = make_float2(one_arg, one_arg);
  for (long long base = (long long)blk * kTpsChunk; base < oplane; base += (long long)blocks_per_image * kTpsChunk) {
    long long rem[kTpsPix];
    float g[2][kTpsPix];
    double ax[kTpsPix], ay[kTpsPix];
#pragma unroll
    for (int i = 0; i < kTpsPix; ++i) {
      rem[i] = base + i * kTpsThreads + threadIdx.x;
      const long long rr = rem[i] < oplane ? rem[i] : oplane - 1;    // tail: compute a valid pixel, skip the store
      const int r = (int)(rr / Wout), c = (int)(rr - (long long)r * Wout);
      g[0][i] = __ldg(xs + c); g[1][i] = __ldg(ys + r);
      // basis order (ones, x, y, r_1 .. r_pn)   (:123)
      ax[i] = s_aff[0].x + s_aff[1].x * (double)g[0][i] + s_aff[2].x * (double)g[1][i];
      ay[i] = s_aff[0].y + s_aff[1].y * (double)g[0][i] + s_aff[2].y * (double)g[1][i];
    }
    const float2 GX = make_float2(g[0][0], g[0][1]), GY = make_float2(g[1][0], g[1][1]);
    for (int k0 = 0; k0 < pnp; k0 += kTpsGroup) {
      float2 sx = make_float2(0.f, 0.f), sy = make_float2(0.f, 0.f);
#pragma unroll
      for (int j = 0; j < kTpsGroup; ++j) {
        const float4 p = s_p[k0 + j], w = s_w[k0 + j];
        const float2 dx = __fadd2_rn(GX, make_float2(p.x, p.y)), dy = __fadd2_rn(GY, make_float2(p.z, p.w));
        const float2 d2 = add2_unfused(__fmul2_rn(dx, dx), __fmul2_rn(dy, dy), one);        // square + square (:115)
        const float2 rk = __fmul2_rn(d2, tps_log2<LOGMODE>(__fadd2_rn(d2, eps)));      // (:116)
        sx = __ffma2_rn(make_float2(w.x, w.y), rk, sx);
        sy = __ffma2_rn(make_float2(w.z, w.w), rk, sy);
      }
      ax[0] += (double)sx.x; ax[1] += (double)sx.y;
      ay[0] += (double)sy.x; ay[1] += (double)sy.y;
    }
#pragma unroll
    for (int i = 0; i < kTpsPix; ++i) {
      if (rem[i] >= oplane) continue;
      UdisTap tap;
      tap.setup((float)ax[i], (float)ay[i], H, W);
      if (idx_dbg) {
        int32_t* d = idx_dbg + (long long)b * 4 * oplane + rem[i];
        d[0] = tap.x0; d[oplane] = tap.x1; d[2 * oplane] = tap.y0; d[3 * oplane] = tap.y1;
      }
      const float* src = U + (long long)b * C * iplane;
      float* dst = out + (long long)b * C * oplane + rem[i];
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
}

// ---------------------------------------------------------------------------
// W3k: kornia-style TPS image warp, warp_image_tps(image, kernel_centers, kernel_weights,
// affine_weights, align_corners) (core/inference/tps_methods/kornia_tps.py:105-176):
//   coords  = create_meshgrid(h, w) in [-1, 1]                       (tables xs / ys passed in)
//   d2_k    = clamp(-2 * p.c_k + |p|^2 + |c_k|^2, min=0)             (_pair_square_euclidean :26-36)
//   U_k     = 0.5 * d2_k * log(d2_k + 1e-8)                          (_kernel_distance :38-45)
//   warped  = sum_k U_k * w_k + (p.x * a_1 + p.y * a_2) + a_0        (kornia warp_points_tps)
//   out     = F.grid_sample(image, warped, bilinear, zeros, align_corners)   (:170-174)
// Same structure as the kernel above: the [B, H*W, K] kernel matrix is never materialised, two
// pixels per thread in packed fp32, fp32 group sums of 8 control points added in fp64.
template <int C_T, int LOGMODE>
__global__ void __launch_bounds__(kTpsThreads)
tps_kornia_warp_kernel(const float* __restrict__ image, const float* __restrict__ centers,
                       const float* __restrict__ kweights, const float* __restrict__ affine,
                       const float* __restrict__ xs, const float* __restrict__ ys,
                       float* __restrict__ out, float* __restrict__ coords_dbg, int C_rt, int H, int W,
                       int K, int align_corners, int blocks_per_image, float one_arg) {
  // per control point (padded with zero weights), lane-duplicated:
  //   s_c[k] = (cx, cx, cy, cy)   s_w[k] = (wx, wx, wy, wy)   s_q[k] = (|c|^2, |c|^2)
  extern __shared__ float4 s_tps[];
  const int Kp = tps_padded(K);
  float4* s_c = s_tps;
  float4* s_w = s_tps + Kp;
  float2* s_q = reinterpret_cast<float2*>(s_tps + 2 * Kp);
  const int C = (C_T > 0) ? C_T : C_rt;
  const int b = blockIdx.x / blocks_per_image;
  const int blk = blockIdx.x - b * blocks_per_image;
  for (int k = threadIdx.x; k < Kp; k += blockDim.x) {
    float cx = 0.f, cy = 0.f, wx = 0.f, wy = 0.f;
    if (k < K) {
      cx = __ldg(centers + ((long long)b * K + k) * 2); cy = __ldg(centers + ((long long)b * K + k) * 2 + 1);
      wx = __ldg(kweights + ((long long)b * K + k) * 2); wy = __ldg(kweights + ((long long)b * K + k) * 2 + 1);
    }
    const float c2 = fadd(fmul(cx, cx), fmul(cy, cy));               // t2_sq = sum(c * c)
    s_c[k] = make_float4(cx, cx, cy, cy);
    s_w[k] = make_float4(wx, wx, wy, wy);
    s_q[k] = make_float2(c2, c2);
  }
  __syncthreads();
  const float* A = affine + (long long)b * 6;                       // [3, 2]: rows a_0, a_1 (x), a_2 (y)
  const float a0x = __ldg(A), a0y = __ldg(A + 1), a1x = __ldg(A + 2), a1y = __ldg(A + 3),
              a2x = __ldg(A + 4), a2y = __ldg(A + 5);
  const long long plane = (long long)H * W;
  // ATen CPU grid_sample: align_corners ? (g + 1) * ((size-1)/2) : fma(g + 1, size/2, -0.5)
  const float sfx = align_corners ? fmul((float)(W - 1), 0.5f) : fmul((float)W, 0.5f);
  const float sfy = align_corners ? fmul((float)(H - 1), 0.5f) : fmul((float)H, 0.5f);
  const float2 eps = make_float2(1e-8f, 1e-8f), m2 = make_float2(-2.0f, -2.0f), half = make_float2(0.5f, 0.5f),
               one = make_float2(one_arg, one_arg);
  for (long long base = (long long)blk * kTpsChunk; base < plane; base += (long long)blocks_per_image * kTpsChunk) {
    long long rem[kTpsPix];
    float g[2][kTpsPix];
    double ax[kTpsPix], ay[kTpsPix];
#pragma unroll
    for (int i = 0; i < kTpsPix; ++i) {
      rem[i] = base + i * kTpsThreads + threadIdx.x;
      const long long rr = rem[i] < plane ? rem[i] : plane - 1;
      const int r = (int)(rr / W), c = (int)(rr - (long long)r * W);
      g[0][i] = __ldg(xs + c); g[1][i] = __ldg(ys + r);
      ax[i] = 0.0; ay[i] = 0.0;
    }
    const float2 GX = make_float2(g[0][0], g[0][1]), GY = make_float2(g[1][0], g[1][1]);
    const float2 P2 = add2_unfused(__fmul2_rn(GX, GX), __fmul2_rn(GY, GY), one);         // t1_sq
    for (int k0 = 0; k0 < Kp; k0 += kTpsGroup) {
      float2 sx = make_float2(0.f, 0.f), sy = make_float2(0.f, 0.f);
#pragma unroll
      for (int j = 0; j < kTpsGroup; ++j) {
        const float4 cc = s_c[k0 + j], w = s_w[k0 + j];
        const float2 q = s_q[k0 + j];
        const float2 dot = add2_unfused(__fmul2_rn(GX, make_float2(cc.x, cc.y)), __fmul2_rn(GY, make_float2(cc.z, cc.w)), one);
        float2 d2 = __fadd2_rn(add2_unfused(__fmul2_rn(m2, dot), P2, one), q);
        d2.x = fmaxf(d2.x, 0.0f); d2.y = fmaxf(d2.y, 0.0f);
        const float2 u = __fmul2_rn(__fmul2_rn(half, d2), tps_log2<LOGMODE>(__fadd2_rn(d2, eps)));
        sx = __ffma2_rn(make_float2(w.x, w.y), u, sx);
        sy = __ffma2_rn(make_float2(w.z, w.w), u, sy);
      }
      ax[0] += (double)sx.x; ax[1] += (double)sx.y;
      ay[0] += (double)sy.x; ay[1] += (double)sy.y;
    }
#pragma unroll
    for (int i = 0; i < kTpsPix; ++i) {
      if (rem[i] >= plane) continue;
      const float gx = g[0][i], gy = g[1][i];
      // + points . affine[1:] + affine[0]
      const float wxs = fadd(fadd((float)ax[i], fadd(fmul(gx, a1x), fmul(gy, a2x))), a0x);
      const float wys = fadd(fadd((float)ay[i], fadd(fmul(gx, a1y), fmul(gy, a2y))), a0y);
      if (coords_dbg) {
        coords_dbg[((long long)b * plane + rem[i]) * 2] = wxs;
        coords_dbg[((long long)b * plane + rem[i]) * 2 + 1] = wys;
      }
      const float ix = align_corners ? fmul(fadd(wxs, 1.0f), sfx) : __fmaf_rn(fadd(wxs, 1.0f), sfx, -0.5f);
      const float iy = align_corners ? fmul(fadd(wys, 1.0f), sfy) : __fmaf_rn(fadd(wys, 1.0f), sfy, -0.5f);
      GridTap tap;
      tap.setup(ix, iy, H, W);
      const float* src = image + (long long)b * C * plane;
      float* dst = out + (long long)b * C * plane + rem[i];
      if (C_T > 0) {
        float v[C_T > 0 ? C_T : 1];
#pragma unroll
        for (int ch = 0; ch < C_T; ++ch) v[ch] = tap.sample(src + ch * plane, W);
#pragma unroll
        for (int ch = 0; ch < C_T; ++ch) stg_stream(dst + ch * plane, v[ch]);
      } else {
        for (int ch = 0; ch < C; ++ch) stg_stream(dst + ch * plane, tap.sample(src + ch * plane, W));
      }
    }
  }
}

// Plain F.grid_sample(bilinear, zeros) on a normalised grid [N, Ho, Wo, 2], both align_corners modes.
__global__ void __launch_bounds__(256)
grid_sample_kernel(const float* __restrict__ img, const float* __restrict__ grid, float* __restrict__ out,
                   int C, int H, int W, long long HoWo, long long total, int align_corners) {
  const float sfx = align_corners ? fmul((float)(W - 1), 0.5f) : fmul((float)W, 0.5f);
  const float sfy = align_corners ? fmul((float)(H - 1), 0.5f) : fmul((float)H, 0.5f);
  const long long plane = (long long)H * W;
  for (long long p = blockIdx.x * (long long)blockDim.x + threadIdx.x; p < total;
       p += (long long)gridDim.x * blockDim.x) {
    const long long n = p / HoWo, rem = p - n * HoWo;
    const float gx = __ldg(grid + p * 2), gy = __ldg(grid + p * 2 + 1);
    const float ix = align_corners ? fmul(fadd(gx, 1.0f), sfx) : __fmaf_rn(fadd(gx, 1.0f), sfx, -0.5f);
    const float iy = align_corners ? fmul(fadd(gy, 1.0f), sfy) : __fmaf_rn(fadd(gy, 1.0f), sfy, -0.5f);
    GridTap tap;
    tap.setup(ix, iy, H, W);
    for (int c = 0; c < C; ++c) out[(n * C + c) * HoWo + rem] = tap.sample(img + (n * C + c) * plane, W);
  }
}

}  // namespace sb

extern "C" int sb_tps_kornia_warp(const float* image, const float* centers, const float* kweights,
                                  const float* affine, const float* xs, const float* ys, float* out,
                                  float* coords_dbg, int B, int C, int H, int W, int K,
                                  int align_corners, sb_stream_t stream) {
  using namespace sb;
  SB_ENTER();
  SB_REQUIRE(B >= 0 && C >= 0 && H > 0 && W > 0 && K >= 0, SB_EINVAL, "sb_tps_kornia_warp: bad size");
  SB_REQUIRE(K <= kTpsMaxPn, SB_EUNSUP, "sb_tps_kornia_warp: K=%d > %d control points", K, kTpsMaxPn);
  SB_REQUIRE((long long)H * W < (1ll << 31), SB_EUNSUP, "sb_tps_kornia_warp: plane too large");
  const long long plane = (long long)H * W;
  if ((long long)B * plane == 0 || C == 0) return SB_OK;
  SB_REQUIRE(image && centers && kweights && affine && xs && ys && out, SB_EINVAL,
             "sb_tps_kornia_warp: null pointer");
  const int threads = kTpsThreads;
  long long bpi = (plane + kTpsChunk - 1) / kTpsChunk;
  const long long cap = (long long)kNumSMs * 8 * 4 / (B > 0 ? B : 1) + 1;
  if (bpi > cap) bpi = cap;
  const size_t smem = (size_t)tps_padded(K) * (2 * sizeof(float4) + sizeof(float2));
  const int grid = (int)(bpi * B);
  cudaStream_t s = as_stream(stream);
  const bool libm_log = tune_get(SB_TUNE_TPS_LOG, 0) == 1;
#define SB_TPSK_LAUNCH(CT)                                                                            \
  do {                                                                                                \
    if (libm_log)                                                                                     \
      tps_kornia_warp_kernel<CT, 1><<<grid, threads, smem, s>>>(image, centers, kweights, affine, xs, ys, out, \
                                                                coords_dbg, C, H, W, K, align_corners, (int)bpi, 1.0f); \
    else                                                                                              \
      tps_kornia_warp_kernel<CT, 0><<<grid, threads, smem, s>>>(image, centers, kweights, affine, xs, ys, out, \
                                                                coords_dbg, C, H, W, K, align_corners, (int)bpi, 1.0f); \
  } while (0)
  switch (C) {
    case 1: SB_TPSK_LAUNCH(1); break;
    case 3: SB_TPSK_LAUNCH(3); break;
    case 6: SB_TPSK_LAUNCH(6); break;
    default: SB_TPSK_LAUNCH(0); break;
  }
#undef SB_TPSK_LAUNCH
  SB_LAUNCH_CHECK("tps_kornia_warp_kernel");
  return SB_OK;
}

extern "C" int sb_grid_sample(const float* img, const float* grid, float* out, int N, int C, int H, int W,
                              int Ho, int Wo, int align_corners, sb_stream_t stream) {
  using namespace sb;
  SB_ENTER();
  SB_REQUIRE(N >= 0 && C >= 0 && H > 0 && W > 0 && Ho >= 0 && Wo >= 0, SB_EINVAL, "sb_grid_sample: bad size");
  SB_REQUIRE((long long)H * W < (1ll << 31), SB_EUNSUP, "sb_grid_sample: plane too large");
  const long long HoWo = (long long)Ho * Wo, total = (long long)N * HoWo;
  if (total == 0 || C == 0) return SB_OK;
  SB_REQUIRE(img && grid && out, SB_EINVAL, "sb_grid_sample: null pointer");
  long long blocks = (total + 255) / 256;
  const long long max_blocks = (long long)kNumSMs * 8 * 16;
  if (blocks > max_blocks) blocks = max_blocks;
  grid_sample_kernel<<<(int)blocks, 256, 0, as_stream(stream)>>>(img, grid, out, C, H, W, HoWo, total,
                                                                 align_corners);
  SB_LAUNCH_CHECK("grid_sample_kernel");
  return SB_OK;
}

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
  const int threads = kTpsThreads;
  long long bpi = (oplane + kTpsChunk - 1) / kTpsChunk;
  const long long cap = (long long)kNumSMs * 8 * 4 / (B > 0 ? B : 1) + 1;
  if (bpi > cap) bpi = cap;
  const size_t smem = (size_t)tps_padded(pn) * 2 * sizeof(float4) + 3 * sizeof(double2);
  const int grid = (int)(bpi * B);
  cudaStream_t s = as_stream(stream);
  const bool libm_log = tune_get(SB_TUNE_TPS_LOG, 0) == 1;
#define SB_TPS_LAUNCH(CT)                                                                             \
  do {                                                                                                \
    if (libm_log)                                                                                     \
      tps_warp_kernel<CT, 1><<<grid, threads, smem, s>>>(U, T, source, xs, ys, out, idx_dbg, C, H, W, Hout, \
                                                         Wout, pn, (int)bpi, 1.0f);                         \
    else                                                                                              \
      tps_warp_kernel<CT, 0><<<grid, threads, smem, s>>>(U, T, source, xs, ys, out, idx_dbg, C, H, W, Hout, \
                                                         Wout, pn, (int)bpi, 1.0f);                         \
  } while (0)
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
