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
#include "ptx_sm100.cuh"
#include "tmap.cuh"

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
// query) and exposes two dependent DRAM round trips (coords -> window) per CTA.
// Here the window fetch is handed to the TMA and the warps only do arithmetic:
//   * cost_maps is described to the TMA as a 3-D tensor {W2, H2, B*H1*W1}; the
//     window of a query is ONE cp.async.bulk.tensor of a 12|16 x 10|12 box whose
//     corner is the query's exact minimum tap (xmin & ~3, ymin) — negative or
//     past-the-edge coordinates included: the hardware zero-fills
//     what lies outside the map, which IS grid_sample's zeros padding (an
//     out-of-range tap multiplies its weight by 0.0 in the same fma chain), so
//     neither the fetch nor the taps need masks or address arithmetic;
//   * persistent warps walk groups of 4 consecutive queries through a software
//     pipeline: the boxes of groups i+1..i+D are in flight into a ring of D+1
//     shared-memory buffers (one mbarrier each) while group i is sampled;
//   * a "superblock" is 8/16/32 consecutive queries of one batch element; a warp
//     loads its x- and y-centres with ONE coalesced load each (lane l holds
//     query l) and keeps the next superblock's in registers too, so the only
//     dependent DRAM round trip left in the loop is the window;
//   * the 9 x- and 9 y-coordinates of a query are round-tripped ONCE by lanes
//     0..17 (the 81 taps are their outer product) and parked next to the window
//     in shared memory; the round trip is monotone, so lanes 0 and 9 hold the
//     exact minimum floors. The IEEE division by (size-1) is done with a
//     correctly-rounded reciprocal and one exact-residual fma step (same
//     result as div.rn, no range check / slow-path call);
//   * lane l < 27 owns x-position i = l % 9 and the three y-positions
//     j = 3*(l/9) + {0,1,2} (lanes of a row group read 9 different columns:
//     at most 2-way bank conflicts on the dense TMA rows);
//   * the 4 x 81 results of a group are staged in shared memory and leave as
//     81 coalesced 16-byte stores.
// A query whose taps do not fit the 16 x 12 box (cannot happen for finite
// coordinates) is handled by the generic per-tap gather.
constexpr int kFastR = 4, kFastSide = 9, kFastTaps = 81;
constexpr int kFastQ = 4;                        // queries per group
constexpr int kBoxWMax = 16, kBoxHMax = 12;      // largest TMA box: 16 floats x 12 rows
constexpr int kWinBytes = kBoxWMax * kBoxHMax * 4;   // 768 (a multiple of 128: TMA destination alignment)
constexpr int kAxisBytes = 160;                  // int2 {floor, frac bits} x 18, then {fits, row pitch}
constexpr int kAxisOff = kFastQ * kWinBytes;     // the 4 axis tables follow the 4 windows
constexpr int kGroupBytes = kFastQ * (kWinBytes + kAxisBytes);   // 3712 = 29 * 128
constexpr int kStageBytes = 1408;                // 4 * 81 * 4 = 1296 -> next multiple of 128
// The TMA wants 16-byte aligned box corners, so a window starts at ax = xmin & ~3 and is
// 12 floats wide when the 10-11 columns the taps touch end before ax + 12 (3 times out of 4),
// else 16; 10 rows when the y-floors are the regular ymin..ymin+8, else 12. One tensor map each.
struct LookupMaps {
  CUtensorMap m[4];                              // [wide + 2 * tall]
};
__host__ __device__ constexpr int fast_warp_bytes(int depth) {
  return (depth + 1) * kGroupBytes + kStageBytes + 128;   // ring + output stage + mbarriers
}
__host__ __device__ constexpr int fast_smem_bytes(int depth) {
  return kLookupWarps * fast_warp_bytes(depth) + 128;     // + alignment slack
}
static_assert(kGroupBytes % 128 == 0 && kWinBytes % 128 == 0 && 19 * 8 <= kAxisBytes, "slot layout");

struct FastLane {   // per-lane constants of the r = 4 kernel
  float den, rcp, half, coff, scale;
};

// Per-axis position of this lane's coordinate (lanes 0..8: x + lane-4, lanes 9..17: y + lane-13):
// grid_roundtrip() with the division restated as a correctly rounded reciprocal and one
// exact-residual fma correction (== div.rn for operands in the normal range, proven exhaustively
// on the CPU for every mantissa and every integer denominator up to 2047). Returns true
// when the operand is outside that range and fast_axis_exact() must be used instead.
__device__ __forceinline__ bool fast_axis(const FastLane& L, float c_raw, int& fi, float& frac) {
  const float a = fmul(2.0f, fadd(fmul(c_raw, L.scale), L.coff));
  float q = fmul(a, L.rcp);
  q = __fmaf_rn(__fmaf_rn(-q, L.den, a), L.rcp, q);   // Markstein: exact for integer den <= 2047 (tests/test_div_restatement.py)
  const float t = fmul(fadd(fsub(q, 1.0f), 1.0f), L.half);
  const float fl = floorf(t);
  frac = fsub(t, fl);
  fi = sat_floor_to_int(fl, -100000, 100000);
  const float aa = fabsf(a);
  return !((aa > 1e-30f && aa < 1e30f) || aa == 0.0f);   // denormal range / huge / non-finite
}
__device__ __noinline__ void fast_axis_exact(const FastLane& L, float c_raw, int& fi, float& frac) {
  const float t = grid_roundtrip(fadd(fmul(c_raw, L.scale), L.coff), L.den, L.half);
  const float fl = floorf(t);
  frac = fsub(t, fl);
  fi = sat_floor_to_int(fl, -100000, 100000);
}

struct FastSb {     // a superblock: this lane's query centre + where the superblock starts
  float x, y;
  int b, pos;
};

template <int kFastDepth>
__global__ void __launch_bounds__(kLookupWarps * 32, kFastDepth == 1 ? 3 : 2)
corr_lookup_r4_kernel(const __grid_constant__ LookupMaps maps, const float* __restrict__ cost_maps,
                      const float* __restrict__ coords, float* __restrict__ out, int B, int HW1,
                      int H2, int W2, float coord_scale, int out_stride, int out_offset, int vec_out,
                      int sbq, unsigned int* dbg) {
  asm volatile("griddepcontrol.launch_dependents;");   // the next kernel's CTAs may take freed SMs early (see the wait below)
  extern __shared__ __align__(128) unsigned char s_raw[];
  constexpr int kBufs = kFastDepth + 1;
  constexpr unsigned kFull = 0xffffffffu;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t raw_u32 = ptx::smem_u32(s_raw);
  const uint32_t w_u32 = ((raw_u32 + 127u) & ~127u) + warp * fast_warp_bytes(kFastDepth);
  unsigned char* w_gen = s_raw + (w_u32 - raw_u32);
  float* s_stage = reinterpret_cast<float*>(w_gen + kBufs * kGroupBytes);
  const uint32_t bar0 = w_u32 + kBufs * kGroupBytes + kStageBytes;

  if (lane == 0) {
#pragma unroll
    for (int i = 0; i < 4; ++i) ptx::prefetch_tensormap(&maps.m[i]);
#pragma unroll
    for (int i = 0; i < kBufs; ++i) ptx::mbar_init(bar0 + 8 * i, 1);
    ptx::fence_mbar_init();
  }
  __syncwarp();

  FastLane L;
  const bool is_x = lane < kFastSide;
  const int cidx = is_x ? lane : (lane < 2 * kFastSide ? lane - kFastSide : 0);
  L.den = is_x ? (float)(W2 - 1) : (float)(H2 - 1);
  L.rcp = __frcp_rn(L.den);
  L.half = fmul(L.den, 0.5f);                                     // == (size-1)/2 exactly
  L.coff = (float)(cidx - kFastR);
  L.scale = coord_scale;
  const int map_sz = H2 * W2;
  const int ti = lane % kFastSide, tjg = (lane / kFastSide) % 3;  // tap x-position / y-group (lanes >= 27 idle)

  const int sbq_flags = sbq;
  sbq &= 0xff;
  const int l2_eighths = ((sbq_flags >> 12) & 0xf) <= 8 ? (sbq_flags >> 12) & 0xf : 0;   // 0 = no cache hint
  const uint64_t l2_policy = l2_eighths > 0 ? ptx::l2_policy_evict_last_fraction(l2_eighths) : 0ull;
  const int gps = sbq / kFastQ;                                   // groups per superblock (>= kFastDepth)
  const int sb_per_b = (HW1 + sbq - 1) / sbq;
  const int n_sb = B * sb_per_b;
  const int n_warps = gridDim.x * kLookupWarps;
  const int sb0 = warp * gridDim.x + blockIdx.x;                  // CTA-interleaved: small problems spread over SMs
  if (sb0 >= n_sb) return;
  const int n_my_sb = (n_sb - sb0 + n_warps - 1) / n_warps;
  const int n_my = n_my_sb * gps;                                 // groups (some may be empty at a ragged HW1 tail)

  auto load_sb = [&](int t) {
    FastSb S;
    S.x = 0.0f; S.y = 0.0f; S.b = 0; S.pos = HW1;
    if (t < n_my_sb) {
      const int sbi = sb0 + t * n_warps;
      S.b = sbi / sb_per_b;
      S.pos = (sbi - S.b * sb_per_b) * sbq;
      if (lane < sbq && S.pos + lane < HW1) {
        const float* cb = coords + (size_t)S.b * 2 * HW1 + S.pos + lane;
        S.x = __ldg(cb);
        S.y = __ldg(cb + HW1);
      }
    }
    return S;
  };

  // Programmatic dependent launch (SB_TUNE_LOOKUP_PDL): everything above — barrier init, tensor-map
  // prefetch, index setup — may run while the previous kernel on the stream drains; the first read of
  // data a predecessor may have produced (coords; the volume through TMA later) waits here.  A no-op
  // when the kernel is launched without the attribute.
  asm volatile("griddepcontrol.wait;" ::: "memory");
  FastSb cur = load_sb(0), nxt = load_sb(1);
  int t_cur = 0;                               // superblock of the group being sampled
  int f_t = 0, f_in = 0, f_buf = 0;            // front cursor: superblock, group in it, ring slot

  // ---- front: axis tables + TMA boxes of one group (the 4 queries in lock step for ILP)
  auto front = [&]() {
    if (f_t < n_my_sb) {
      const bool use_cur = (f_t == t_cur);
      const float sx = use_cur ? cur.x : nxt.x, sy = use_cur ? cur.y : nxt.y;
      const int b = use_cur ? cur.b : nxt.b;
      const int pos0 = (use_cur ? cur.pos : nxt.pos) + f_in * kFastQ;
      unsigned char* gbuf = w_gen + f_buf * kGroupBytes;
      const uint32_t gbuf_u32 = w_u32 + f_buf * kGroupBytes;
      const uint32_t bar = bar0 + 8 * f_buf;
      const int q0 = b * HW1 + pos0;
      const int nvalid = HW1 - pos0;                              // queries k < nvalid exist
      float c[kFastQ];
#pragma unroll
      for (int k = 0; k < kFastQ; ++k) {
        const int src = f_in * kFastQ + k;
        const float vx = __shfl_sync(kFull, sx, src), vy = __shfl_sync(kFull, sy, src);
        c[k] = is_x ? vx : vy;
      }
      int fi[kFastQ];
      float fr[kFastQ];
      bool odd = false;
#pragma unroll
      for (int k = 0; k < kFastQ; ++k) odd |= fast_axis(L, c[k], fi[k], fr[k]);
      if (__any_sync(kFull, odd)) {
#pragma unroll
        for (int k = 0; k < kFastQ; ++k) fast_axis_exact(L, c[k], fi[k], fr[k]);
      }
      int ax[kFastQ], ymin[kFastQ];
      unsigned big[kFastQ], over[kFastQ];
      int hflag[kFastQ];                                          // bit 0 fits, bit 1 regular, bits 2-3 xmin & 3
#pragma unroll
      for (int k = 0; k < kFastQ; ++k) {
        const int xmin = __shfl_sync(kFull, fi[k], 0);
        ymin[k] = __shfl_sync(kFull, fi[k], kFastSide);
        ax[k] = xmin & ~3;
        // taps reach floor+1: span = last column / row touched, relative to the box corner
        const int span = (lane < 2 * kFastSide) ? fi[k] + 1 - (is_x ? ax[k] : ymin[k]) : 0;
        big[k] = __ballot_sync(kFull, span >= (is_x ? 12 : 10));
        over[k] = __ballot_sync(kFull, span >= (is_x ? kBoxWMax : kBoxHMax));
        // regular = every floor is min + tap index (no ulp wobble across an integer): rows / columns are consecutive
        const bool irregular = (lane < 2 * kFastSide) & (fi[k] - (is_x ? xmin : ymin[k]) != cidx);
        hflag[k] = (over[k] == 0u ? 1 : 0) | (__any_sync(kFull, irregular) ? 0 : 2) | ((xmin & 3) << 2);
      }
#pragma unroll
      for (int k = 0; k < kFastQ; ++k) {
        int2* axis = reinterpret_cast<int2*>(gbuf + kAxisOff + k * kAxisBytes);
        const int bw = (big[k] & 0x1ffu) ? 16 : 12;
        if (lane < 2 * kFastSide) axis[lane] = make_int2(fi[k], __float_as_int(fr[k]));
        else if (lane == 2 * kFastSide) axis[lane] = make_int2(hflag[k], bw);
      }
      if (lane == 0) {
        uint32_t tx = 0;
#pragma unroll
        for (int k = 0; k < kFastQ; ++k) {
          if (k < nvalid && over[k] == 0u) {
            const int wide = (big[k] & 0x1ffu) ? 1 : 0, tall = (big[k] >> kFastSide) ? 1 : 0;
            if (l2_eighths > 0)
              ptx::tma_load_3d_hint(gbuf_u32 + k * kWinBytes, &maps.m[wide + 2 * tall], bar, ax[k], ymin[k], q0 + k,
                                    l2_policy);
            else
              ptx::tma_load_3d(gbuf_u32 + k * kWinBytes, &maps.m[wide + 2 * tall], bar, ax[k], ymin[k], q0 + k);
            tx += (wide ? 16 : 12) * (tall ? 12 : 10) * 4;
          }
        }
        ptx::mbar_arrive_expect_tx(bar, tx);
      }
      __syncwarp();
    }
    if (++f_in == gps) { f_in = 0; ++f_t; }
    if (++f_buf == kBufs) f_buf = 0;
  };

#pragma unroll
  for (int d = 0; d < kFastDepth; ++d) front();

  int b_in = 0, b_buf = 0;
  uint32_t b_par = 0;
  for (int i = 0; i < n_my; ++i) {
    front();                                                      // group i + D
    ptx::mbar_wait(bar0 + 8 * b_buf, b_par, 0x40 + b_buf, dbg);   // boxes of group i have landed
    const int pos0 = cur.pos + b_in * kFastQ;
    const size_t q0 = (size_t)cur.b * HW1 + pos0;
    const int nvalid = HW1 - pos0;
    const unsigned char* gbuf = w_gen + b_buf * kGroupBytes;
    int2 hdr[kFastQ];                                             // {fits, row pitch}
#pragma unroll
    for (int k = 0; k < kFastQ; ++k)
      hdr[k] = reinterpret_cast<const int2*>(gbuf + kAxisOff + k * kAxisBytes)[2 * kFastSide];
    if (sbq_flags & 0x100) {
      // experiment knob (tools/lookup_exp.py): fetch only, no sampling
    } else if (nvalid >= kFastQ && ((hdr[0].x & hdr[1].x & hdr[2].x & hdr[3].x) & 3) == 3) {
      // ---- hot path: all 4 queries fit and are regular -> window rows 3*tjg .. 3*tjg+3, columns
      // (xmin & 3) + ti, +1; 4 queries in lock step, 48 independent shared-memory loads per lane
      if (lane < 27) {
        float v[kFastQ][4][2], w[kFastQ], n[kFastQ][3];
#pragma unroll
        for (int k = 0; k < kFastQ; ++k) {
          const int pitch = hdr[k].y;
          const float* base = reinterpret_cast<const float*>(gbuf + k * kWinBytes) + (hdr[k].x >> 2) + ti +
                              3 * tjg * pitch;
          const int2* axis = reinterpret_cast<const int2*>(gbuf + kAxisOff + k * kAxisBytes);
          w[k] = __int_as_float(axis[ti].y);
#pragma unroll
          for (int m = 0; m < 3; ++m) n[k][m] = __int_as_float(axis[kFastSide + 3 * tjg + m].y);
#pragma unroll
          for (int r = 0; r < 4; ++r) {
            v[k][r][0] = base[r * pitch];
            v[k][r][1] = base[r * pitch + 1];
          }
        }
#pragma unroll
        for (int k = 0; k < kFastQ; ++k) {
          const float e = fsub(1.0f, w[k]);
#pragma unroll
          for (int m = 0; m < 3; ++m) {
            const float s = fsub(1.0f, n[k][m]);
            const float nw = fmul(s, e), ne = fmul(s, w[k]), sw = fmul(n[k][m], e), se = fmul(n[k][m], w[k]);
            s_stage[k * kFastTaps + ti * kFastSide + 3 * tjg + m] = __fmaf_rn(
                v[k][m + 1][1], se,
                __fmaf_rn(v[k][m + 1][0], sw, __fmaf_rn(v[k][m][1], ne, fmul(v[k][m][0], nw))));
          }
        }
      }
    } else {
      // ---- ragged tail of a batch element, or a query whose taps do not fit the box
      for (int k = 0; k < kFastQ && k < nvalid; ++k) {
        const float* win = reinterpret_cast<const float*>(gbuf + k * kWinBytes);
        const int2* axis = reinterpret_cast<const int2*>(gbuf + kAxisOff + k * kAxisBytes);
        const int2 h = axis[2 * kFastSide];
        float* srow = s_stage + k * kFastTaps;
        if (h.x & 1) {
          if (lane < 27) {
            const int ax = axis[0].x & ~3, ymin = axis[kFastSide].x;
            const int2 axx = axis[ti];
            const float wq = __int_as_float(axx.y), e = fsub(1.0f, wq);
            for (int m = 0; m < 3; ++m) {
              const int2 ay = axis[kFastSide + 3 * tjg + m];
              const float nq = __int_as_float(ay.y), s = fsub(1.0f, nq);
              const float* p = win + (ay.x - ymin) * h.y + (axx.x - ax);
              const float v_nw = p[0], v_ne = p[1], v_sw = p[h.y], v_se = p[h.y + 1];
              const float nw = fmul(s, e), ne = fmul(s, wq), sw = fmul(nq, e), se = fmul(nq, wq);
              srow[ti * kFastSide + 3 * tjg + m] =
                  __fmaf_rn(v_se, se, __fmaf_rn(v_sw, sw, __fmaf_rn(v_ne, ne, fmul(v_nw, nw))));
            }
          }
        } else {
          // cold path: per-tap masked gather straight from global memory
          const float* cb = coords + (size_t)cur.b * 2 * HW1 + pos0 + k;
          const float cx = fmul(__ldg(cb), coord_scale), cy = fmul(__ldg(cb + HW1), coord_scale);
          const float* map = cost_maps + (q0 + k) * map_sz;
          const float denx = (float)(W2 - 1), deny = (float)(H2 - 1);
          const float halfx = fmul(denx, 0.5f), halfy = fmul(deny, 0.5f);
          for (int t = lane; t < kFastTaps; t += 32) {
            const int tx = t / kFastSide, ty = t - tx * kFastSide;
            GridTap tap;
            tap.setup(grid_roundtrip(fadd(cx, (float)(tx - kFastR)), denx, halfx),
                      grid_roundtrip(fadd(cy, (float)(ty - kFastR)), deny, halfy), H2, W2);
            srow[t] = tap.sample(map, W2);
          }
        }
      }
    }
    __syncwarp();
    // ---- results leave coalesced
    if (vec_out && nvalid >= kFastQ) {
      float4* o4 = reinterpret_cast<float4*>(out + q0 * kFastTaps);
      const float4* s4 = reinterpret_cast<const float4*>(s_stage);
#pragma unroll
      for (int t = lane; t < kFastTaps; t += 32) stg_stream4(reinterpret_cast<float*>(o4 + t), s4[t]);
    } else {
      for (int k = 0; k < kFastQ && k < nvalid; ++k) {
        float* orow = out + (q0 + k) * out_stride + out_offset;
        for (int t = lane; t < kFastTaps; t += 32) orow[t] = s_stage[k * kFastTaps + t];
      }
    }
    if (++b_buf == kBufs) { b_buf = 0; b_par ^= 1u; }
    if (++b_in == gps) {                                          // next superblock becomes current
      b_in = 0;
      ++t_cur;
      cur = nxt;
      nxt = load_sb(t_cur + 1);
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

// Launch of the r = 4 kernel with D window groups in flight per warp: persistent grid sized for
// `per_sm` resident CTAs per SM, superblock size by problem size, tuning knobs folded into `sbq`.
template <int D>
static int launch_r4(const LookupMaps& maps, const float* cost_maps, const float* coords, float* out, int B, int hw1,
                     int H2, int W2, float coord_scale, int out_stride, int out_offset, int vec_out, int per_sm,
                     long long nq, unsigned int* dbg_word, sb_stream_t stream) {
  static SmemOptIn opt_in;   // one per instantiation D
  int opt_dev;
  if (opt_in.need(fast_smem_bytes(D), &opt_dev)) {
    SB_CUDA(cudaFuncSetAttribute(corr_lookup_r4_kernel<D>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                 fast_smem_bytes(D)));
    opt_in.done(fast_smem_bytes(D), opt_dev);
  }
  if (per_sm <= 0) per_sm = (D == 1) ? 3 : 2;
  const long long resident_warps = (long long)kNumSMs * per_sm * kLookupWarps;
  int sbq = tune_get(SB_TUNE_LOOKUP_SUPERBLOCK, 0);
  if (sbq != 8 && sbq != 16 && sbq != 32)
    sbq = (nq >= resident_warps * 24) ? 32 : (nq >= resident_warps * 12 ? 16 : 8);
  if (sbq < D * kFastQ) sbq = 16;
  const long long n_sb = (long long)B * ((hw1 + sbq - 1) / sbq);
  const long long ctas = n_sb < (long long)kNumSMs * per_sm ? n_sb : (long long)kNumSMs * per_sm;
  const int sbq_arg = sbq | (tune_get(SB_TUNE_LOOKUP_FETCH_ONLY, 0) ? 0x100 : 0) |
                      ((tune_get(SB_TUNE_LOOKUP_L2_KEEP_EIGHTHS, 3) & 0xf) << 12);
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3((unsigned)ctas);
  cfg.blockDim = dim3(kLookupWarps * 32);
  cfg.dynamicSmemBytes = fast_smem_bytes(D);
  cfg.stream = as_stream(stream);
  cudaLaunchAttribute pdl_attr[1];
  pdl_attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  pdl_attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = pdl_attr;
  cfg.numAttrs = tune_get(SB_TUNE_LOOKUP_PDL, 0) ? 1 : 0;
  SB_CUDA(cudaLaunchKernelEx(&cfg, corr_lookup_r4_kernel<D>, maps, cost_maps, coords, out, B, hw1, H2, W2, coord_scale,
                             out_stride, out_offset, vec_out, sbq_arg, dbg_word));
  return SB_OK;
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
  if (r == kFastR && (W2 & 3) == 0 && nq < (1ll << 30) && W2 <= 2048 && H2 <= 2048 && !tune_get(SB_TUNE_LOOKUP_GENERIC, 0)) {
    const int hw1 = H1 * W1;
    const int vec_out = (out_stride == kFastTaps && out_offset == 0 && (hw1 % kFastQ) == 0 && aligned16(out)) ? 1 : 0;
    const int depth = tune_get(SB_TUNE_LOOKUP_DEPTH, 2);
    int per_sm = tune_get(SB_TUNE_LOOKUP_CTAS_PER_SM, 0);
    unsigned int* dbg_word = debug_word_device();
    if (!dbg_word) return SB_ECUDA;
    LookupMaps maps;
    for (int v = 0; v < 4; ++v) {
      const int rc = make_map_3d_ex(&maps.m[v], CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, cost_maps,
                                    (unsigned long long)W2, (unsigned long long)H2, (unsigned long long)nq,
                                    (v & 1) ? 16 : 12, (v & 2) ? 12 : 10, CU_TENSOR_MAP_SWIZZLE_NONE,
                                    CU_TENSOR_MAP_L2_PROMOTION_NONE, "cost_maps");
      if (rc != SB_OK) return rc;
    }
    const int rc4 = depth == 1 ? launch_r4<1>(maps, cost_maps, coords, out, B, hw1, H2, W2, coord_scale, out_stride,
                                               out_offset, vec_out, per_sm, nq, dbg_word, stream)
                               : launch_r4<2>(maps, cost_maps, coords, out, B, hw1, H2, W2, coord_scale, out_stride,
                                               out_offset, vec_out, per_sm, nq, dbg_word, stream);
    if (rc4 != SB_OK) return rc4;
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
