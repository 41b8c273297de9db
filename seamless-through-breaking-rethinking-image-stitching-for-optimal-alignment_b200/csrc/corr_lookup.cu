// corr_lookup.cu — C3 / C3p: per-iteration bilinear correlation lookup.
// Replaces MemoryDecoder.encode_flow_token (core/FlowFormer/PerCostFormer3/decoder.py:242-260)
// + bilinear_sampler (core/utils/utils.py:62-76); pyramid convention from the
// dead code core/FlowFormer/common.py:245-248 (centroid / 2**i + delta).
//
// One warp per query. The (2r+4) x (2r+4) window of that query's H2 x W2 cost
// map that the (2r+1)^2 taps can touch is staged in shared memory with
// row-aligned 16-byte loads (each cache line of the window is requested once),
// out-of-range rows/columns are filled with zeros (= grid_sample zeros
// padding), then every lane evaluates taps k = lane, lane+32, ... with the
// reference's exact per-tap coordinate arithmetic and writes the
// [q, (2r+1)^2] row coalesced.
//
// HBM-bound gather. Algorithmic bytes per query (r = 4): (2r+2)^2*4 = 400 read
// + 8 coords + (2r+1)^2*4 = 324 written = 732 B.
#include "bilinear.cuh"

namespace sb {

constexpr int kLookupWarps = 8;
constexpr int kMaxR = 7;
constexpr int kWinRowsMax = 2 * kMaxR + 4;           // 18
constexpr int kWinChunksMax = (2 * kMaxR + 4 + 3 + 3) / 4;  // 6 float4 per row
constexpr int kWinFloatsMax = kWinRowsMax * kWinChunksMax * 4;

__device__ __forceinline__ int sat_floor_to_int(float f, int lo, int hi) {
  // floor(f) as int, saturated into [lo, hi]; NaN -> lo.
  if (!(f == f)) return lo;
  f = floorf(f);
  f = fminf(fmaxf(f, (float)lo), (float)hi);
  return (int)f;
}

__global__ void __launch_bounds__(kLookupWarps * 32)
corr_lookup_kernel(const float* __restrict__ cost_maps, const float* __restrict__ coords,
                   float* __restrict__ out, long long n_query, int HW1, int H2, int W2, int r,
                   float coord_scale, int out_stride, int out_offset) {
  __shared__ __align__(16) float s_win[kLookupWarps][kWinFloatsMax];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int side = 2 * r + 1, ntap = side * side;
  const int win_rows = 2 * r + 4;
  const int win_chunks = (2 * r + 4 + 3 + 3) / 4;
  const int win_cols = win_chunks * 4;
  const float denx = (float)(W2 - 1), deny = (float)(H2 - 1);
  const float halfx = fdiv((float)(W2 - 1), 2.0f), halfy = fdiv((float)(H2 - 1), 2.0f);
  const long long map_sz = (long long)H2 * W2;
  const bool vec_ok = (W2 & 3) == 0;  // rows keep 16-byte alignment
  float* win = s_win[warp];

  for (long long q = (long long)blockIdx.x * kLookupWarps + warp; q < n_query;
       q += (long long)gridDim.x * kLookupWarps) {
    const long long b = q / HW1;
    const long long pos = q - b * HW1;
    // coords [B, 2, H1, W1]: channel 0 = x, channel 1 = y
    const float cx = fmul(__ldg(coords + (b * 2) * HW1 + pos), coord_scale);
    const float cy = fmul(__ldg(coords + (b * 2 + 1) * HW1 + pos), coord_scale);
    // Window origin from the first tap (i = j = 0), one pixel of slack for the
    // ulp-level wobble of the per-tap round trips.
    const float ix0 = grid_roundtrip(fadd(cx, (float)(-r)), denx, halfx);
    const float iy0 = grid_roundtrip(fadd(cy, (float)(-r)), deny, halfy);
    const int wx0 = sat_floor_to_int(ix0, -64, W2 + 64) - 1;
    const int wy0 = sat_floor_to_int(iy0, -64, H2 + 64) - 1;
    const int ax = wx0 & ~3;  // 16-byte aligned window start column (floor to multiple of 4)
    const float* map = cost_maps + q * map_sz;

    __syncwarp();  // previous query's reads of `win` are done
    for (int t = lane; t < win_rows * win_chunks; t += 32) {
      const int wr = t / win_chunks, ch = t - wr * win_chunks;
      const int gy = wy0 + wr, gx = ax + ch * 4;
      float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
      if (vec_ok && gy >= 0 && gy < H2 && gx >= 0 && gx + 3 < W2)
        v = ldg_stream4(map + (long long)gy * W2 + gx);
      else if (gy >= 0 && gy < H2 && gx + 3 >= 0 && gx < W2) {  // ragged edge (W2 % 4 != 0)
        const float* rowp = map + (long long)gy * W2;
        if (gx >= 0 && gx < W2) v.x = __ldg(rowp + gx);
        if (gx + 1 >= 0 && gx + 1 < W2) v.y = __ldg(rowp + gx + 1);
        if (gx + 2 >= 0 && gx + 2 < W2) v.z = __ldg(rowp + gx + 2);
        if (gx + 3 >= 0 && gx + 3 < W2) v.w = __ldg(rowp + gx + 3);
      }
      *reinterpret_cast<float4*>(win + wr * win_cols + ch * 4) = v;
    }
    __syncwarp();

    float* orow = out + q * out_stride + out_offset;
    for (int k = lane; k < ntap; k += 32) {
      const int i = k / side, j = k - i * side;
      // coords = centroid + delta; delta[i][j] = (dy[i], dx[j]) -> x += i-r, y += j-r
      const float ix = grid_roundtrip(fadd(cx, (float)(i - r)), denx, halfx);
      const float iy = grid_roundtrip(fadd(cy, (float)(j - r)), deny, halfy);
      GridTap tap;
      tap.setup(ix, iy, H2, W2);
      // tap.off_nw = y_n * W2 + x_w  -> recover (x_w, y_n) relative to the window
      float v_nw = 0.f, v_ne = 0.f, v_sw = 0.f, v_se = 0.f;
      const float xwf = floorf(ix), ynf = floorf(iy);
      const int xw = sat_floor_to_int(xwf, -100000, 100000);
      const int yn = sat_floor_to_int(ynf, -100000, 100000);
      const int lx = xw - ax, ly = yn - wy0;
      if (lx >= 0 && lx + 1 < win_cols && ly >= 0 && ly + 1 < win_rows) {
        const float* wp = win + ly * win_cols + lx;
        v_nw = wp[0]; v_ne = wp[1]; v_sw = wp[win_cols]; v_se = wp[win_cols + 1];
      } else {  // window missed (non-finite or wildly inconsistent coordinates): direct gather
        if (tap.m_nw) v_nw = __ldg(map + tap.off_nw);
        if (tap.m_ne) v_ne = __ldg(map + tap.off_nw + 1);
        if (tap.m_sw) v_sw = __ldg(map + tap.off_nw + W2);
        if (tap.m_se) v_se = __ldg(map + tap.off_nw + W2 + 1);
      }
      orow[k] = tap.combine(v_nw, v_ne, v_sw, v_se);
    }
  }
}

// ---------------------------------------------------------------------------
// Fast path: r = 4 (the only radius the reference uses), W2 % 4 == 0.
// The generic kernel above is instruction-bound (~1000 issued instructions per
// query: per-tap divisions, runtime index divisions, masks).  Here
//   * the 9 x- and 9 y-coordinates of a query are round-tripped ONCE by lanes
//     0..17 (the 81 taps are their outer product) and shared with shuffles;
//   * the window is zero-filled in shared memory, which IS grid_sample's zeros
//     padding (an out-of-range tap multiplies its weight by 0.0 in the same fma
//     chain), so the taps need no masks;
//   * lane l < 27 owns window row pair j = l % 9 and the three x-positions
//     i = 3*(l/9) + {0,1,2}: 8 shared-memory loads feed 3 taps;
//   * all index math is 32-bit and compile-time where possible.
// A query whose taps do not fit the staged window (non-finite coordinates) is
// handled by the generic per-tap gather.
constexpr int kFastR = 4, kFastSide = 9;
constexpr int kFastRows = 12, kFastPitch = 20;   // 12 x 16 floats, rows padded to 20 words
constexpr int kFastQ = 4;                        // queries per warp, all loads issued up front

// grid = (ceil(HW1 / 32), B): a CTA owns 32 consecutive queries of one batch element,
// warp w the 4 queries [4w, 4w+4) — no index divisions, and 4 x (1 + 2) independent
// global loads in flight per warp hide the two dependent DRAM round trips.
__global__ void __launch_bounds__(kLookupWarps * 32)
corr_lookup_r4_kernel(const float* __restrict__ cost_maps, const float* __restrict__ coords,
                      float* __restrict__ out, int HW1, int H2, int W2, float coord_scale,
                      int out_stride, int out_offset) {
  __shared__ __align__(16) float s_win[kLookupWarps][kFastQ][kFastRows * kFastPitch];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int b = blockIdx.y;
  const int pos0 = blockIdx.x * (kLookupWarps * kFastQ) + warp * kFastQ;
  if (pos0 >= HW1) return;
  const float denx = (float)(W2 - 1), deny = (float)(H2 - 1);
  const float halfx = fmul((float)(W2 - 1), 0.5f), halfy = fmul((float)(H2 - 1), 0.5f);  // == /2 exactly
  const int map_sz = H2 * W2;
  // lane roles
  const bool is_x = lane < kFastSide;
  const int cidx = is_x ? lane : (lane < 2 * kFastSide ? lane - kFastSide : 0);
  const float cden = is_x ? denx : deny, chalf = is_x ? halfx : halfy;
  const float coff = (float)(cidx - kFastR);
  const int tj = lane % kFastSide, tig = (lane / kFastSide) % 3;   // tap row / x-group (lanes >= 27 idle)
  const int wrow0 = lane >> 2, wch = lane & 3;                     // window chunk owned in pass 0
  const float* cbase = coords + ((size_t)b * 2 + (is_x ? 0 : 1)) * HW1 + pos0;
  const size_t q0 = (size_t)b * HW1 + pos0;

  // ---- phase A: the 4 centre coordinates (independent loads)
  float c_raw[kFastQ];
#pragma unroll
  for (int k = 0; k < kFastQ; ++k) c_raw[k] = (pos0 + k < HW1) ? __ldg(cbase + k) : 0.0f;

  // ---- phase B: per-axis round trips, window origins, window loads (8 x LDG.128 in flight)
  int fi[kFastQ], wy0[kFastQ], ax[kFastQ];
  float wfrac[kFastQ];
  float4 v0[kFastQ], v1[kFastQ];
#pragma unroll
  for (int k = 0; k < kFastQ; ++k) {
    const float t = grid_roundtrip(fadd(fmul(c_raw[k], coord_scale), coff), cden, chalf);
    const float fl = floorf(t);
    wfrac[k] = fsub(t, fl);
    fi[k] = sat_floor_to_int(fl, -100000, 100000);
    const int x0i = __shfl_sync(0xffffffffu, fi[k], 0), y0i = __shfl_sync(0xffffffffu, fi[k], kFastSide);
    wy0[k] = y0i - 1;
    ax[k] = (x0i - 1) & ~3;
    const float* map = cost_maps + (q0 + k) * map_sz;
    const int gx = ax[k] + wch * 4;
    const bool xin = (gx >= 0) & (gx + 3 < W2) & (pos0 + k < HW1);
    const int gy = wy0[k] + wrow0;
    v0[k] = make_float4(0.f, 0.f, 0.f, 0.f);
    v1[k] = make_float4(0.f, 0.f, 0.f, 0.f);
    if (xin & (gy >= 0) & (gy < H2)) v0[k] = ldg_stream4(map + gy * W2 + gx);
    if (xin & (lane < 16) & (gy + 8 >= 0) & (gy + 8 < H2)) v1[k] = ldg_stream4(map + (gy + 8) * W2 + gx);
  }
#pragma unroll
  for (int k = 0; k < kFastQ; ++k) {
    float* win = s_win[warp][k];
    *reinterpret_cast<float4*>(win + wrow0 * kFastPitch + wch * 4) = v0[k];
    if (lane < 16) *reinterpret_cast<float4*>(win + (wrow0 + 8) * kFastPitch + wch * 4) = v1[k];
  }
  __syncwarp();

  // ---- phase C: taps
#pragma unroll
  for (int k = 0; k < kFastQ; ++k) {
    if (pos0 + k >= HW1) break;                                    // warp-uniform
    const float* win = s_win[warp][k];
    const int yn = __shfl_sync(0xffffffffu, fi[k], kFastSide + tj);
    const float n = __shfl_sync(0xffffffffu, wfrac[k], kFastSide + tj);
    int xw[3];
    float wv[3];
#pragma unroll
    for (int m = 0; m < 3; ++m) {
      xw[m] = __shfl_sync(0xffffffffu, fi[k], 3 * tig + m);
      wv[m] = __shfl_sync(0xffffffffu, wfrac[k], 3 * tig + m);
    }
    const int ly = yn - wy0[k];
    bool fits = (ly >= 0) & (ly + 1 < kFastRows);
#pragma unroll
    for (int m = 0; m < 3; ++m) fits &= (xw[m] - ax[k] >= 0) & (xw[m] - ax[k] + 1 < 16);
    const bool all_fit = __all_sync(0xffffffffu, fits | (lane >= 27));
    float* orow = out + (q0 + k) * out_stride + out_offset;
    if (all_fit) {
      if (lane < 27) {
        const float s = fsub(1.0f, n);
#pragma unroll
        for (int m = 0; m < 3; ++m) {
          const float w = wv[m], e = fsub(1.0f, w);
          const float* wp = win + ly * kFastPitch + (xw[m] - ax[k]);
          const float v_nw = wp[0], v_ne = wp[1], v_sw = wp[kFastPitch], v_se = wp[kFastPitch + 1];
          const float nw = fmul(s, e), ne = fmul(s, w), sw = fmul(n, e), se = fmul(n, w);
          orow[(3 * tig + m) * kFastSide + tj] =
              __fmaf_rn(v_se, se, __fmaf_rn(v_sw, sw, __fmaf_rn(v_ne, ne, fmul(v_nw, nw))));
        }
      }
    } else {
      // cold path (non-finite coordinates): per-tap masked gather straight from global memory
      const float* cb = coords + (size_t)b * 2 * HW1 + pos0 + k;
      const float cx = fmul(__ldg(cb), coord_scale), cy = fmul(__ldg(cb + HW1), coord_scale);
      const float* map = cost_maps + (q0 + k) * map_sz;
      for (int t = lane; t < kFastSide * kFastSide; t += 32) {
        const int i = t / kFastSide, j = t - i * kFastSide;
        GridTap tap;
        tap.setup(grid_roundtrip(fadd(cx, (float)(i - kFastR)), denx, halfx),
                  grid_roundtrip(fadd(cy, (float)(j - kFastR)), deny, halfy), H2, W2);
        orow[t] = tap.sample(map, W2);
      }
    }
  }
}

// Generic bilinear_sampler: one thread per output location, loops channels.
__global__ void __launch_bounds__(256)
bilinear_sampler_kernel(const float* __restrict__ img, const float* __restrict__ coords,
                        float* __restrict__ out, int C, int H, int W, long long HoWo,
                        long long total) {
  const float denx = (float)(W - 1), deny = (float)(H - 1);
  const float halfx = fdiv((float)(W - 1), 2.0f), halfy = fdiv((float)(H - 1), 2.0f);
  const long long plane = (long long)H * W;
  for (long long p = blockIdx.x * (long long)blockDim.x + threadIdx.x; p < total;
       p += (long long)gridDim.x * blockDim.x) {
    const long long n = p / HoWo, rem = p - n * HoWo;
    const float x = __ldg(coords + p * 2), y = __ldg(coords + p * 2 + 1);
    GridTap tap;
    tap.setup(grid_roundtrip(x, denx, halfx), grid_roundtrip(y, deny, halfy), H, W);
    for (int c = 0; c < C; ++c)
      out[(n * C + c) * HoWo + rem] = tap.sample(img + (n * C + c) * plane, W);
  }
}

}  // namespace sb

extern "C" int sb_corr_lookup(const float* cost_maps, const float* coords, float* out, int B,
                              int H1, int W1, int H2, int W2, int r, float coord_scale,
                              int out_stride, int out_offset, sb_stream_t stream) {
  using namespace sb;
  SB_ENTER();
  
  SB_REQUIRE(B >= 0 && H1 >= 0 && W1 >= 0 && H2 > 0 && W2 > 0, SB_EINVAL,
             "sb_corr_lookup: bad size");
  SB_REQUIRE(r >= 0 && r <= kMaxR, SB_EUNSUP, "sb_corr_lookup: r=%d outside [0,%d]", r, kMaxR);
  const int ntap = (2 * r + 1) * (2 * r + 1);
  SB_REQUIRE(out_stride >= ntap && out_offset >= 0 && out_offset + ntap <= out_stride, SB_EINVAL,
             "sb_corr_lookup: out_stride/out_offset inconsistent with (2r+1)^2=%d", ntap);
  SB_REQUIRE(aligned16(cost_maps), SB_EINVAL, "sb_corr_lookup: cost_maps must be 16-byte aligned");
  SB_REQUIRE((long long)H2 * W2 < (1ll << 31), SB_EUNSUP, "sb_corr_lookup: map too large");
  const long long nq = (long long)B * H1 * W1;
  if (nq == 0) return SB_OK;
  SB_REQUIRE(cost_maps && coords && out, SB_EINVAL, "sb_corr_lookup: null pointer");
  long long blocks = (nq + kLookupWarps - 1) / kLookupWarps;
  const long long max_blocks = (long long)kNumSMs * 8 * 8;
  if (blocks > max_blocks) blocks = max_blocks;
  if (r == kFastR && (W2 & 3) == 0 && B <= 65535) {
    const int hw1 = H1 * W1, per_cta = kLookupWarps * kFastQ;
    dim3 grid((hw1 + per_cta - 1) / per_cta, B);
    corr_lookup_r4_kernel<<<grid, kLookupWarps * 32, 0, as_stream(stream)>>>(
        cost_maps, coords, out, hw1, H2, W2, coord_scale, out_stride, out_offset);
    SB_LAUNCH_CHECK("corr_lookup_r4_kernel");
    return SB_OK;
  }
  corr_lookup_kernel<<<(int)blocks, kLookupWarps * 32, 0, as_stream(stream)>>>(
      cost_maps, coords, out, nq, H1 * W1, H2, W2, r, coord_scale, out_stride, out_offset);
  SB_LAUNCH_CHECK("corr_lookup_kernel");
  return SB_OK;
}

extern "C" int sb_bilinear_sampler(const float* img, const float* coords, float* out, int N, int C,
                                   int H, int W, int Ho, int Wo, sb_stream_t stream) {
  using namespace sb;
  SB_ENTER();
  
  SB_REQUIRE(N >= 0 && C >= 0 && H > 0 && W > 0 && Ho >= 0 && Wo >= 0, SB_EINVAL,
             "sb_bilinear_sampler: bad size");
  SB_REQUIRE((long long)H * W < (1ll << 31), SB_EUNSUP, "sb_bilinear_sampler: plane too large");
  const long long HoWo = (long long)Ho * Wo, total = (long long)N * HoWo;
  if (total == 0 || C == 0) return SB_OK;
  SB_REQUIRE(img && coords && out, SB_EINVAL, "sb_bilinear_sampler: null pointer");
  long long blocks = (total + 255) / 256;
  const long long max_blocks = (long long)kNumSMs * 8 * 16;
  if (blocks > max_blocks) blocks = max_blocks;
  bilinear_sampler_kernel<<<(int)blocks, 256, 0, as_stream(stream)>>>(img, coords, out, C, H, W,
                                                                      HoWo, total);
  SB_LAUNCH_CHECK("bilinear_sampler_kernel");
  return SB_OK;
}
