// homo_warp.cu — W2: homography backward warp with the UDIS sampler.
// Replaces transformer(U, theta, out_size) (core/udis_utils/torch_homo_transform.py:5-151):
//   _meshgrid (:94-112)  -> (xs[c], ys[r], 1)          (linspace tables passed in)
//   _transform (:114-145)-> T_g = theta @ grid; t += 1e-6 where |t| < 1e-7; x = X/t, y = Y/t
//   _interpolate (:17-92)-> UdisTap (bilinear.cuh)
//
// The 3-term dot product is evaluated as the oracle's BLAS does for K = 3:
//   acc = t0*gx ; acc = fma(t1, gy, acc) ; acc = fma(t2, 1, acc)
// (verified bit-for-bit against torch.matmul on CPU, see tests/golden).
//
// HBM-bound: coordinates are computed, not loaded; per output pixel C planes
// are gathered (4 taps) and C floats written: 2*C*4 algorithmic bytes / px.
//
// n_ones: the callers always warp cat(image, ones) (flowHomoAdpater.py:110-113,
// :292,:310,:314) — the trailing all-ones planes are synthesised here
// (((wa*1 + wb*1) + wc*1) + wd*1, the same fp32 sequence as sampling a stored
// 1.0) instead of being materialised, concatenated and read back.
//
// Launch shape: grid (Wout/32, Hout/8, B), block (32, 8) — no index divisions.
#include <cuda.h>   // CUtensorMap (types only)
#include <limits.h>

#include "bilinear.cuh"
#include "ptx_sm100.cuh"
#include "tmap.cuh"

namespace sb {

__device__ __forceinline__ float dot3(const float* t, float gx, float gy) {
  float acc = fmul(t[0], gx);
  acc = __fmaf_rn(t[1], gy, acc);
  acc = __fmaf_rn(t[2], 1.0f, acc);
  return acc;
}

template <int C_T>
__global__ void __launch_bounds__(256)
homo_warp_kernel(const float* __restrict__ U, const float* __restrict__ theta,
                 const float* __restrict__ xs, const float* __restrict__ ys,
                 float* __restrict__ out, int32_t* __restrict__ idx_dbg,
                 int C_rt, int n_ones, int H, int W, int Hout, int Wout, int theta_batch) {
  const int C = (C_T > 0) ? C_T : C_rt;
  const int Cout = C + n_ones;
  const int c = blockIdx.x * 32 + threadIdx.x, r = blockIdx.y * 8 + threadIdx.y;
  if (c >= Wout || r >= Hout) return;
  const int b = blockIdx.z;
  const int oplane = Hout * Wout, iplane = H * W;
  const int rem = r * Wout + c;
  const float* th = theta + (theta_batch > 1 ? b * 9 : 0);
  float t[9];
#pragma unroll
  for (int i = 0; i < 9; ++i) t[i] = __ldg(th + i);
  const float gx = __ldg(xs + c), gy = __ldg(ys + r);
  const float X = dot3(t, gx, gy), Y = dot3(t + 3, gx, gy);
  float T = dot3(t + 6, gx, gy);
  // smallers = 1e-6 * (1 - float(|t| >= 1e-7)); t = t + smallers   (:133-137)
  const float ge = (fabsf(T) >= 1e-7f) ? 1.0f : 0.0f;
  T = fadd(T, fmul(1e-6f, fsub(1.0f, ge)));
  UdisTap tap;
  tap.setup(fdiv(X, T), fdiv(Y, T), H, W);
  if (idx_dbg) {
    int32_t* d = idx_dbg + (size_t)b * 4 * oplane + rem;
    d[0] = tap.x0; d[oplane] = tap.x1; d[2 * (size_t)oplane] = tap.y0; d[3 * (size_t)oplane] = tap.y1;
  }
  // Channel loop: byte pointers stepped by the plane stride (one 64-bit add per pointer and channel).  A warp whose
  // 32 pixels all have unclamped corners (x1 == x0 + 1, y1 == y0 + 1: everything but the image border) reads the east
  // taps as +4-byte immediates of the two row pointers; the first version re-derived y*W + x and a 64-bit address for
  // every tap of every channel (327 issued instructions per pixel for 12 loads and 6 stores, ncu round 1).
  const size_t iplane_b = (size_t)iplane * sizeof(float), oplane_b = (size_t)oplane * sizeof(float);
  const char* src = reinterpret_cast<const char*>(U) + (size_t)b * C * iplane_b;
  char* dst = reinterpret_cast<char*>(out) + ((size_t)b * Cout * oplane + rem) * sizeof(float);
  if (n_ones > 0) {
    const float one = fadd(fadd(fadd(fmul(tap.wa, 1.0f), fmul(tap.wb, 1.0f)), fmul(tap.wc, 1.0f)),
                           fmul(tap.wd, 1.0f));
    for (int ch = 0; ch < n_ones; ++ch) stg_stream(reinterpret_cast<float*>(dst + (size_t)(C + ch) * oplane_b), one);
  }
  const bool interior = __all_sync(__activemask(), (tap.x1 == tap.x0 + 1) & (tap.y1 == tap.y0 + 1));
  const char* p0 = src + ((size_t)tap.y0 * W + tap.x0) * sizeof(float);      // Ia; Ic = +4 when interior
  const char* p1 = src + ((size_t)tap.y1 * W + tap.x0) * sizeof(float);      // Ib; Id = +4 when interior
  const int east = (tap.x1 - tap.x0) * (int)sizeof(float);                   // 0 or 4 (per lane) on the border path
  auto ld = [](const char* p, int byte_off) { return __ldg(reinterpret_cast<const float*>(p + byte_off)); };
  auto mix = [&](float Ia, float Ib, float Ic, float Id) {
    return fadd(fadd(fadd(fmul(tap.wa, Ia), fmul(tap.wb, Ib)), fmul(tap.wc, Ic)), fmul(tap.wd, Id));
  };
  if (C_T > 0) {
    float v[C_T > 0 ? C_T : 1];
    if (interior) {
#pragma unroll
      for (int ch = 0; ch < C_T; ++ch) {
        v[ch] = mix(ld(p0, 0), ld(p1, 0), ld(p0, 4), ld(p1, 4));
        p0 += iplane_b; p1 += iplane_b;
      }
    } else {
#pragma unroll
      for (int ch = 0; ch < C_T; ++ch) {
        v[ch] = mix(ld(p0, 0), ld(p1, 0), ld(p0, east), ld(p1, east));
        p0 += iplane_b; p1 += iplane_b;
      }
    }
#pragma unroll
    for (int ch = 0; ch < C_T; ++ch) {
      stg_stream(reinterpret_cast<float*>(dst), v[ch]);
      dst += oplane_b;
    }
  } else {
    for (int ch = 0; ch < C; ++ch) {
      stg_stream(reinterpret_cast<float*>(dst), mix(ld(p0, 0), ld(p1, 0), ld(p0, east), ld(p1, east)));
      p0 += iplane_b; p1 += iplane_b; dst += oplane_b;
    }
  }
}

// ---------------------------------------------------------------------------------------------------------
// Tiled form (W % 4 == 0, C in {1, 2, 3, 6}): the same shared-memory staging as flow_warp_tiled_kernel.  A CTA
// owns a 64 x 16 OUTPUT tile; its threads compute the source positions of their 4 pixels (the same fp32 sequence:
// dot products, the small-T fix-up, two IEEE divisions, clamped corners), the CTA reduces the bounding box of the
// clamped corner indices, one thread fetches that box of every channel plane with the TMA, and the 4 taps of a
// channel are shared-memory reads.  A tile whose source footprint does not fit 80 x 25 floats (strong zoom-out or
// rotation) stages the box under the tile and its outliers gather from global memory as before.
constexpr int kHTW = 64, kHTH = 16, kHBoxW = 80, kHBoxH = 25;
constexpr int kHPlane = (kHBoxW * kHBoxH + 31) / 32 * 32;      // planes 8064 B apart (128-byte aligned TMA destinations)

template <int C_T>
__global__ void __launch_bounds__(256)
homo_warp_tiled_kernel(const __grid_constant__ CUtensorMap map_u, const float* __restrict__ U,
                       const float* __restrict__ theta, const float* __restrict__ xs, const float* __restrict__ ys,
                       float* __restrict__ out, int32_t* __restrict__ idx_dbg, int n_ones, int H, int W, int Hout,
                       int Wout, int theta_batch, unsigned int* dbg) {
  extern __shared__ __align__(128) uint8_t tile_raw[];
  float* tile = reinterpret_cast<float*>(tile_raw);                     // [C_T][kHBoxH][kHBoxW]
  int* red = reinterpret_cast<int*>(tile_raw + C_T * kHPlane * 4);
  const uint32_t bar = ptx::smem_u32(tile_raw) + C_T * kHPlane * 4 + 192;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int c = blockIdx.x * kHTW + (tid & 63);
  const int r0 = blockIdx.y * kHTH + (tid >> 6);                         // rows r0 + 4 j
  const int b = blockIdx.z;
  const int Cout = C_T + n_ones;
  const int oplane = Hout * Wout, iplane = H * W;
  if (tid == 0) { ptx::mbar_init(bar, 1); ptx::fence_mbar_init(); }
  const float* th = theta + (theta_batch > 1 ? b * 9 : 0);
  float t[9];
#pragma unroll
  for (int i = 0; i < 9; ++i) t[i] = __ldg(th + i);
  const float gx = (c < Wout) ? __ldg(xs + c) : 0.0f;

  UdisTap tap[4];
  bool ok[4];
  int bx0 = INT_MAX, bx1 = INT_MIN, by0 = INT_MAX, by1 = INT_MIN;
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    const int r = r0 + 4 * j;
    ok[j] = c < Wout && r < Hout;
    const float gy = ok[j] ? __ldg(ys + r) : 0.0f;
    const float X = dot3(t, gx, gy), Y = dot3(t + 3, gx, gy);
    float T = dot3(t + 6, gx, gy);
    const float ge = (fabsf(T) >= 1e-7f) ? 1.0f : 0.0f;
    T = fadd(T, fmul(1e-6f, fsub(1.0f, ge)));
    tap[j].setup(fdiv(X, T), fdiv(Y, T), H, W);
    if (ok[j]) {
      bx0 = min(bx0, tap[j].x0); bx1 = max(bx1, tap[j].x1);
      by0 = min(by0, tap[j].y0); by1 = max(by1, tap[j].y1);
    }
  }
  bx0 = __reduce_min_sync(0xffffffffu, bx0); bx1 = __reduce_max_sync(0xffffffffu, bx1);
  by0 = __reduce_min_sync(0xffffffffu, by0); by1 = __reduce_max_sync(0xffffffffu, by1);
  if (lane == 0) { red[warp * 4 + 0] = bx0; red[warp * 4 + 1] = bx1; red[warp * 4 + 2] = by0; red[warp * 4 + 3] = by1; }
  __syncthreads();
  if (tid == 0) {
    for (int w2 = 1; w2 < 8; ++w2) {
      bx0 = min(bx0, red[w2 * 4 + 0]); bx1 = max(bx1, red[w2 * 4 + 1]);
      by0 = min(by0, red[w2 * 4 + 2]); by1 = max(by1, red[w2 * 4 + 3]);
    }
    int ox = bx0 & ~3, oy = by0;
    if (bx1 - ox >= kHBoxW || by1 - oy >= kHBoxH) {            // footprint too large: box under the tile itself
      ox = min(max((int)(blockIdx.x * kHTW) - 8, 0), max(W - kHBoxW, 0)) & ~3;
      oy = min(max((int)(blockIdx.y * kHTH) - 4, 0), max(H - kHBoxH, 0));
    }
    red[32] = ox; red[33] = oy;
    ptx::mbar_arrive_expect_tx(bar, (uint32_t)(C_T * kHBoxW * kHBoxH * 4));
#pragma unroll
    for (int ch = 0; ch < C_T; ++ch)
      ptx::tma_load_3d(ptx::smem_u32(tile) + ch * kHPlane * 4, &map_u, bar, ox, oy, b * C_T + ch);
  }
  __syncthreads();
  const int ox = red[32], oy = red[33];
  ptx::mbar_wait(bar, 0, 22, dbg);

  const size_t iplane_b = (size_t)iplane * sizeof(float), oplane_b = (size_t)oplane * sizeof(float);
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    if (!ok[j]) continue;
    const int r = r0 + 4 * j;
    const int rem = r * Wout + c;
    const UdisTap& tp = tap[j];
    if (idx_dbg) {
      int32_t* d = idx_dbg + (size_t)b * 4 * oplane + rem;
      d[0] = tp.x0; d[oplane] = tp.x1; d[2 * (size_t)oplane] = tp.y0; d[3 * (size_t)oplane] = tp.y1;
    }
    char* dst = reinterpret_cast<char*>(out) + ((size_t)b * Cout * oplane + rem) * sizeof(float);
    if (n_ones > 0) {
      const float one = fadd(fadd(fadd(fmul(tp.wa, 1.0f), fmul(tp.wb, 1.0f)), fmul(tp.wc, 1.0f)), fmul(tp.wd, 1.0f));
      for (int ch = 0; ch < n_ones; ++ch) stg_stream(reinterpret_cast<float*>(dst + (size_t)(C_T + ch) * oplane_b), one);
    }
    auto mix = [&](float Ia, float Ib, float Ic, float Id) {
      return fadd(fadd(fadd(fmul(tp.wa, Ia), fmul(tp.wb, Ib)), fmul(tp.wc, Ic)), fmul(tp.wd, Id));
    };
    float v[C_T];
    const int lx0 = tp.x0 - ox, lx1 = tp.x1 - ox, ly0 = tp.y0 - oy, ly1 = tp.y1 - oy;
    if (lx0 >= 0 && lx1 < kHBoxW && ly0 >= 0 && ly1 < kHBoxH) {
      const float* pa = tile + ly0 * kHBoxW + lx0;     // Ia (y0, x0)
      const float* pb = tile + ly1 * kHBoxW + lx0;     // Ib (y1, x0)
      const int east = lx1 - lx0;                      // 0 on a clamped border, else 1
#pragma unroll
      for (int ch = 0; ch < C_T; ++ch)
        v[ch] = mix(pa[ch * kHPlane], pb[ch * kHPlane], pa[ch * kHPlane + east], pb[ch * kHPlane + east]);
    } else {
      const char* src = reinterpret_cast<const char*>(U) + (size_t)b * C_T * iplane_b;
      const char* p0 = src + ((size_t)tp.y0 * W + tp.x0) * sizeof(float);
      const char* p1 = src + ((size_t)tp.y1 * W + tp.x0) * sizeof(float);
      const int east = (tp.x1 - tp.x0) * (int)sizeof(float);
#pragma unroll
      for (int ch = 0; ch < C_T; ++ch) {
        v[ch] = mix(__ldg(reinterpret_cast<const float*>(p0)), __ldg(reinterpret_cast<const float*>(p1)),
                    __ldg(reinterpret_cast<const float*>(p0 + east)), __ldg(reinterpret_cast<const float*>(p1 + east)));
        p0 += iplane_b; p1 += iplane_b;
      }
    }
#pragma unroll
    for (int ch = 0; ch < C_T; ++ch) {
      stg_stream(reinterpret_cast<float*>(dst), v[ch]);
      dst += oplane_b;
    }
  }
}

template <int C_T>
static int launch_homo_tiled(const float* U, const float* theta, const float* xs, const float* ys, float* out,
                             int32_t* idx_dbg, int B, int n_ones, int H, int W, int Hout, int Wout, int theta_batch,
                             cudaStream_t s) {
  CUtensorMap map_u;
  int rc = make_map_3d_ex(&map_u, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, U, (unsigned long long)W, (unsigned long long)H,
                          (unsigned long long)B * C_T, kHBoxW, kHBoxH, CU_TENSOR_MAP_SWIZZLE_NONE,
                          CU_TENSOR_MAP_L2_PROMOTION_NONE, "homography warp source");
  if (rc != SB_OK) return rc;
  unsigned int* dbg = debug_word_device();
  if (!dbg) return SB_ECUDA;
  const size_t smem = (size_t)C_T * kHPlane * 4 + 256;
  static SmemOptIn opt_in;
  int opt_dev;
  if (smem > 48 * 1024 && opt_in.need(smem, &opt_dev)) {
    SB_CUDA(cudaFuncSetAttribute(homo_warp_tiled_kernel<C_T>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    opt_in.done(smem, opt_dev);
  }
  const dim3 grid((Wout + kHTW - 1) / kHTW, (Hout + kHTH - 1) / kHTH, B);
  homo_warp_tiled_kernel<C_T><<<grid, 256, smem, s>>>(map_u, U, theta, xs, ys, out, idx_dbg, n_ones, H, W, Hout, Wout,
                                                      theta_batch, dbg);
  return SB_OK;
}

}  // namespace sb

extern "C" int sb_homo_warp(const float* U, const float* theta, const float* xs, const float* ys,
                            float* out, int32_t* idx_dbg, int B, int C, int n_ones, int H, int W,
                            int Hout, int Wout, int theta_batch, sb_stream_t stream) {
  using namespace sb;
  SB_ENTER();
  SB_REQUIRE(B >= 0 && C >= 0 && n_ones >= 0 && H > 0 && W > 0 && Hout >= 0 && Wout >= 0, SB_EINVAL,
             "sb_homo_warp: bad size");
  SB_REQUIRE(theta_batch == 1 || theta_batch == B, SB_EINVAL,
             "sb_homo_warp: theta batch %d must be 1 or B=%d", theta_batch, B);
  SB_REQUIRE((long long)H * W < (1ll << 31) && (long long)Hout * Wout < (1ll << 31), SB_EUNSUP,
             "sb_homo_warp: plane too large");
  const long long total = (long long)B * Hout * Wout;
  if (total == 0 || C + n_ones == 0) return SB_OK;
  SB_REQUIRE((U || C == 0) && theta && xs && ys && out, SB_EINVAL, "sb_homo_warp: null pointer");
  SB_REQUIRE(B <= 65535 && (Hout + 7) / 8 <= 65535, SB_EUNSUP, "sb_homo_warp: B or Hout too large for one launch");
  const dim3 block(32, 8), grid((Wout + 31) / 32, (Hout + 7) / 8, B);
  cudaStream_t s = as_stream(stream);
#define SB_HOMO_LAUNCH(CT)                                                                      \
  homo_warp_kernel<CT><<<grid, block, 0, s>>>(U, theta, xs, ys, out, idx_dbg, C, n_ones, H, W, \
                                              Hout, Wout, theta_batch)
  const bool tiled = (W & 3) == 0 && aligned16(U) && (C == 1 || C == 2 || C == 3 || C == 6) && W >= 32 && H >= 8 &&
                     tune_get(SB_TUNE_WARP_TILED, 0) == 1;
  if (tiled) {
    int rc = SB_OK;
    switch (C) {
      case 1: rc = launch_homo_tiled<1>(U, theta, xs, ys, out, idx_dbg, B, n_ones, H, W, Hout, Wout, theta_batch, s); break;
      case 2: rc = launch_homo_tiled<2>(U, theta, xs, ys, out, idx_dbg, B, n_ones, H, W, Hout, Wout, theta_batch, s); break;
      case 3: rc = launch_homo_tiled<3>(U, theta, xs, ys, out, idx_dbg, B, n_ones, H, W, Hout, Wout, theta_batch, s); break;
      default: rc = launch_homo_tiled<6>(U, theta, xs, ys, out, idx_dbg, B, n_ones, H, W, Hout, Wout, theta_batch, s); break;
    }
    if (rc != SB_OK) return rc;
    SB_LAUNCH_CHECK("homo_warp_tiled_kernel");
    return SB_OK;
  }
  switch (C) {
    case 1: SB_HOMO_LAUNCH(1); break;
    case 2: SB_HOMO_LAUNCH(2); break;
    case 3: SB_HOMO_LAUNCH(3); break;
    case 6: SB_HOMO_LAUNCH(6); break;
    default: SB_HOMO_LAUNCH(0); break;
  }
#undef SB_HOMO_LAUNCH
  SB_LAUNCH_CHECK("homo_warp_kernel");
  return SB_OK;
}
