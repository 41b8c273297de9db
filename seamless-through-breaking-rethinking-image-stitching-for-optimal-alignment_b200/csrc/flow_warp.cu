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
#include <cuda.h>   // CUtensorMap (types only)
#include <limits.h>

#include "bilinear.cuh"
#include "ptx_sm100.cuh"
#include "tmap.cuh"

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

// mode = 'nearest' of the reference's warp (core/warp_utils.py:74-79).  The grid is normalised as for the bilinear mode
// (2 v / max(size-1, 1) - 1) but the reference then calls F.grid_sample(mode='nearest') WITHOUT align_corners (:79), i.e.
// align_corners=False: ATen's CPU kernel un-normalises as fma(g + 1, size / 2, -0.5) (one rounding), rounds half to even
// (nearbyint) and copies that pixel if it lies inside the image, else 0.  No caller in the reference uses this mode; it
// exists so that the mirrored signature is complete.
__global__ void __launch_bounds__(256)
flow_warp_nearest_kernel(const float* __restrict__ x, const float* __restrict__ flo, float* __restrict__ out, int C,
                         int H, int W, const FlowWarpConst k) {
  const int px = blockIdx.x * 32 + threadIdx.x, py = blockIdx.y * 8 + threadIdx.y;
  if (px >= W || py >= H) return;
  const int b = blockIdx.z;
  const int plane = H * W;
  const int rem = py * W + px;
  const float* fl = flo + (size_t)b * 2 * plane + rem;
  const float fx = ldg_stream(fl), fy = ldg_stream(fl + plane);
  const float gx = fsub(fdiv(fmul(2.0f, fadd((float)px, fx)), k.denx), 1.0f);
  const float gy = fsub(fdiv(fmul(2.0f, fadd((float)py, fy)), k.deny), 1.0f);
  const float ix = nearbyintf(__fmaf_rn(fadd(gx, 1.0f), fmul((float)W, 0.5f), -0.5f));      // round half to even
  const float iy = nearbyintf(__fmaf_rn(fadd(gy, 1.0f), fmul((float)H, 0.5f), -0.5f));
  const bool inside = ix >= 0.0f && ix <= (float)(W - 1) && iy >= 0.0f && iy <= (float)(H - 1);   // false for NaN
  const size_t off = inside ? (size_t)((int)iy) * W + (int)ix : 0;
  const float* src = x + (size_t)b * C * plane + off;
  float* po = out + (size_t)b * C * plane + rem;
  for (int c = 0; c < C; ++c) {
    stg_stream(po, inside ? __ldg(src) : 0.0f);
    src += plane; po += plane;
  }
}

// ---------------------------------------------------------------------------------------------------------
// Tiled form (W % 4 == 0, C in {1, 2, 3, 6}): shared-memory staging of the source tile.
// The per-pixel kernel above is bound by the L1 data stage, not by HBM or issue slots: 24 four-byte gathers per
// pixel, each warp-level one an unaligned 128-byte access (two tag / data passes), measured 59.6 us at batch 16 /
// 512^2 = 60 % of HBM peak with DRAM at 45 % (and unchanged when the instruction count was cut by a third).
// Here a CTA owns a 64 x 16 output tile: every thread computes the sample positions of its 4 pixels (the same
// fp32 sequence), the CTA reduces the integer bounding box of all taps, and ONE thread asks the TMA for that box
// of every channel plane (<= 80 x 25 floats; the corner may be negative or past the edge: the hardware zero-fills
// what lies outside the image, which IS grid_sample's zeros padding).  The 24 taps of a pixel are then
// conflict-free shared-memory reads at immediate offsets of one address.  A pixel whose taps fall outside the staged
// box (a flow field that is not smooth at the scale of a tile) takes the masked global gathers of the first kernel.
// MEASURED (round 2, B200, batch 16 / 512^2, C = 6): bit-identical to the per-pixel kernel and SLOWER — 74.8 us vs
// 62.9 us (homography form: 55.8 vs 53.9 us).  The four phases of a CTA (flow load, bounding-box reduction, TMA
// round trip, sampling) run back to back and 78 registers x 256 threads leave three CTAs per SM to overlap them.
// Kept as an opt-in (sb_tune(SB_TUNE_WARP_TILED, 1)) with its tests; the per-pixel kernels stay the default.
// Variants measured on the same workload (per-pixel kernel: 63.4 us): 64 x 16 tiles capped at 64 registers (four
// CTAs per SM) 65.0 us; 32 x 16 tiles with 128 threads 72.7 - 81.0 us; 128 x 16 tiles 117.9 us.  ncu on the per-pixel
// kernel: 76 % of the warp cycles wait on the L1TEX scoreboard (ptxas keeps it at 32 registers and consumes the 24
// taps in small dependent batches), but neither issuing the 24 loads as ordered asm statements behind a warp-sync
// fence (ptxas still interleaves them) nor an L1 prefetch (CCTL.PF1) per tap row ahead of the loads changes the
// time (66.5 us; 138 us instead of 91 us for a noise flow), so both were dropped again.
constexpr int kTW = 64, kTH = 16, kBoxW = 80, kBoxH = 25;
constexpr int kPlaneFloats = (kBoxW * kBoxH + 31) / 32 * 32;   // 8000 B box, planes 8064 B apart: TMA destinations are 128-byte aligned

template <int C_T, bool EXACT_RCP>
__global__ void __launch_bounds__(256)
flow_warp_tiled_kernel(const __grid_constant__ CUtensorMap map_x, const float* __restrict__ x,
                       const float* __restrict__ flo, const float* __restrict__ mul_mask, float* __restrict__ out,
                       float* __restrict__ overlap, int H, int W, const FlowWarpConst k, unsigned int* dbg) {
  extern __shared__ __align__(128) uint8_t tile_raw[];
  float* tile = reinterpret_cast<float*>(tile_raw);                     // [C_T][kBoxH][kBoxW]
  int* red = reinterpret_cast<int*>(tile_raw + C_T * kPlaneFloats * 4);  // [8 warps][4] bbox partials, then origin
  const uint32_t bar = ptx::smem_u32(tile_raw) + C_T * kPlaneFloats * 4 + 192;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int px = blockIdx.x * kTW + (tid & 63);
  const int py0 = blockIdx.y * kTH + (tid >> 6);                         // rows py0 + 4 j, j = 0..3
  const int b = blockIdx.z;
  const int plane = H * W;
  if (tid == 0) { ptx::mbar_init(bar, 1); ptx::fence_mbar_init(); }

  GridTap tap[4];
  float m[4];
  bool ok[4];
  int bx0 = INT_MAX, bx1 = INT_MIN, by0 = INT_MAX, by1 = INT_MIN;
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    const int py = py0 + 4 * j;
    ok[j] = px < W && py < H;
    float fx = 0.0f, fy = 0.0f;
    m[j] = 1.0f;
    if (ok[j]) {
      const float* fl = flo + (size_t)b * 2 * plane + (size_t)py * W + px;
      fx = ldg_stream(fl); fy = ldg_stream(fl + plane);
      if (mul_mask) m[j] = ldg_stream(mul_mask + (size_t)b * plane + (size_t)py * W + px);
    }
    if (EXACT_RCP)
      tap[j].setup(grid_roundtrip_rcp(fadd((float)px, fx), k.denx, k.rcpx, k.halfx),
                   grid_roundtrip_rcp(fadd((float)py, fy), k.deny, k.rcpy, k.halfy), H, W);
    else
      tap[j].setup(grid_roundtrip(fadd((float)px, fx), k.denx, k.halfx),
                   grid_roundtrip(fadd((float)py, fy), k.deny, k.halfy), H, W);
    if (ok[j]) {
      bx0 = min(bx0, tap[j].xi); bx1 = max(bx1, tap[j].xi);
      by0 = min(by0, tap[j].yi); by1 = max(by1, tap[j].yi);
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
    int ox = bx0 & ~3, oy = by0;                                          // TMA: 16-byte aligned inner coordinate
    if (bx1 + 1 - ox >= kBoxW || by1 + 1 - oy >= kBoxH) {
      // the taps of this tile do not fit one box: stage the box around the tile itself (identity flow); pixels
      // that reach outside take the global path
      ox = (blockIdx.x * kTW - 8) & ~3; oy = blockIdx.y * kTH - 4;
    }
    red[32] = ox; red[33] = oy;
    ptx::mbar_arrive_expect_tx(bar, (uint32_t)(C_T * kBoxW * kBoxH * 4));
#pragma unroll
    for (int c = 0; c < C_T; ++c)
      ptx::tma_load_3d(ptx::smem_u32(tile) + c * kPlaneFloats * 4, &map_x, bar, ox, oy, b * C_T + c);
  }
  __syncthreads();
  const int ox = red[32], oy = red[33];
  ptx::mbar_wait(bar, 0, 21, dbg);

  const size_t plane_bytes = (size_t)plane * sizeof(float);
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    if (!ok[j]) continue;
    const int py = py0 + 4 * j;
    const int rem = py * W + px;
    const int lx = tap[j].xi - ox, ly = tap[j].yi - oy;
    float v[C_T];
    if (lx >= 0 && lx + 1 < kBoxW && ly >= 0 && ly + 1 < kBoxH) {
      const float* t0 = tile + ly * kBoxW + lx;
#pragma unroll
      for (int c = 0; c < C_T; ++c)
        v[c] = tap[j].combine(t0[c * kPlaneFloats], t0[c * kPlaneFloats + 1], t0[c * kPlaneFloats + kBoxW],
                              t0[c * kPlaneFloats + kBoxW + 1]);
    } else {
      const char* pn = reinterpret_cast<const char*>(x) + ((size_t)b * C_T * plane + tap[j].off_nw) * sizeof(float);
      const char* ps = pn + (size_t)W * sizeof(float);
#pragma unroll
      for (int c = 0; c < C_T; ++c) {
        v[c] = tap[j].sample_rows(reinterpret_cast<const float*>(pn), reinterpret_cast<const float*>(ps));
        pn += plane_bytes; ps += plane_bytes;
      }
    }
    if (C_T == 6 && overlap) {
      const float mean = div_small_int(fadd(fadd(v[3 % C_T], v[4 % C_T]), v[5 % C_T]), 3.0f, 0.3333333432674407958984375f);
      stg_stream(overlap + (size_t)b * plane + rem, mean < 0.9f ? 1.0f : 0.0f);
    }
    char* po = reinterpret_cast<char*>(out) + ((size_t)b * C_T * plane + rem) * sizeof(float);
#pragma unroll
    for (int c = 0; c < C_T; ++c) {
      stg_stream(reinterpret_cast<float*>(po), mul_mask ? fmul(v[c], m[j]) : v[c]);
      po += plane_bytes;
    }
  }
}

template <int C_T>
static int launch_flow_tiled(const float* x, const float* flo, const float* mul_mask, float* out, float* overlap, int B,
                             int H, int W, const FlowWarpConst& k, bool exact_rcp, cudaStream_t s) {
  CUtensorMap map_x;
  int rc = make_map_3d_ex(&map_x, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, x, (unsigned long long)W, (unsigned long long)H,
                          (unsigned long long)B * C_T, kBoxW, kBoxH, CU_TENSOR_MAP_SWIZZLE_NONE,
                          CU_TENSOR_MAP_L2_PROMOTION_NONE, "flow warp source");
  if (rc != SB_OK) return rc;
  unsigned int* dbg = debug_word_device();
  if (!dbg) return SB_ECUDA;
  const size_t smem = (size_t)C_T * kPlaneFloats * 4 + 256;
  static SmemOptIn opt_in;
  int opt_dev;
  if (smem > 48 * 1024 && opt_in.need(smem, &opt_dev)) {
    SB_CUDA(cudaFuncSetAttribute(flow_warp_tiled_kernel<C_T, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    SB_CUDA(cudaFuncSetAttribute(flow_warp_tiled_kernel<C_T, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    opt_in.done(smem, opt_dev);
  }
  const dim3 grid((W + kTW - 1) / kTW, (H + kTH - 1) / kTH, B);
  if (exact_rcp) flow_warp_tiled_kernel<C_T, true><<<grid, 256, smem, s>>>(map_x, x, flo, mul_mask, out, overlap, H, W, k, dbg);
  else flow_warp_tiled_kernel<C_T, false><<<grid, 256, smem, s>>>(map_x, x, flo, mul_mask, out, overlap, H, W, k, dbg);
  return SB_OK;
}

}  // namespace sb

extern "C" int sb_flow_warp_nearest(const float* x, const float* flo, float* out, int B, int C, int H, int W,
                                    sb_stream_t stream) {
  using namespace sb;
  SB_ENTER();
  SB_REQUIRE(B >= 0 && C >= 0 && H >= 0 && W >= 0, SB_EINVAL, "sb_flow_warp_nearest: negative size");
  SB_REQUIRE((long long)H * W < (1ll << 31), SB_EUNSUP, "sb_flow_warp_nearest: plane too large");
  if ((long long)B * H * W == 0 || C == 0) return SB_OK;
  SB_REQUIRE(x && flo && out, SB_EINVAL, "sb_flow_warp_nearest: null pointer");
  SB_REQUIRE(B <= 65535 && (H + 7) / 8 <= 65535, SB_EUNSUP, "sb_flow_warp_nearest: B or H too large for one launch");
  const dim3 block(32, 8), grid((W + 31) / 32, (H + 7) / 8, B);
  FlowWarpConst k;
  k.denx = (float)(W - 1 > 1 ? W - 1 : 1); k.deny = (float)(H - 1 > 1 ? H - 1 : 1);
  k.rcpx = 1.0f / k.denx; k.rcpy = 1.0f / k.deny;
  k.halfx = (float)(W - 1) * 0.5f; k.halfy = (float)(H - 1) * 0.5f;
  flow_warp_nearest_kernel<<<grid, block, 0, as_stream(stream)>>>(x, flo, out, C, H, W, k);
  SB_LAUNCH_CHECK("flow_warp_nearest_kernel");
  return SB_OK;
}

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
  // tiled form (opt-in, sb_tune(SB_TUNE_WARP_TILED, 1)): bit-identical, measured SLOWER than the per-pixel kernel
  // (74.8 vs 62.9 us at batch 16 / 512^2: flow load -> bbox reduction -> TMA -> sample serialise inside a CTA)
  const bool tiled = (W & 3) == 0 && aligned16(x) && (C == 1 || C == 2 || C == 3 || C == 6) && W >= 32 && H >= 8 &&
                     tune_get(SB_TUNE_WARP_TILED, 0) == 1;
  if (tiled) {
    int rc = SB_OK;
    switch (C) {
      case 1: rc = launch_flow_tiled<1>(x, flo, mul_mask, out, overlap, B, H, W, k, exact_rcp != 0, s); break;
      case 2: rc = launch_flow_tiled<2>(x, flo, mul_mask, out, overlap, B, H, W, k, exact_rcp != 0, s); break;
      case 3: rc = launch_flow_tiled<3>(x, flo, mul_mask, out, overlap, B, H, W, k, exact_rcp != 0, s); break;
      default: rc = launch_flow_tiled<6>(x, flo, mul_mask, out, overlap, B, H, W, k, exact_rcp != 0, s); break;
    }
    if (rc != SB_OK) return rc;
    SB_LAUNCH_CHECK("flow_warp_tiled_kernel");
    return SB_OK;
  }
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
