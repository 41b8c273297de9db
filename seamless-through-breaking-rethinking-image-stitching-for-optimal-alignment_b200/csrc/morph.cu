// morph.cu — W5: binary morphological open (erode then dilate with a box).
// Replaces preprocess_occlusion_mask (core/flowHomoAdpater.py:18-35):
//   m = (mask >= .5); e = (conv(m, ones kh x kw, zero pad) == kh*kw);
//   d = (conv(e, ones, zero pad) >= 1); out = (d >= .5)
// and, with border_is_zero = 0, the cv2.erode/cv2.dilate 11x11 open of
// core/inference/tps_pipline.py:143-147 (cv2 ignores out-of-image pixels when
// eroding).
//
// Bit-parallel: a CTA packs a (64 + 2*(kh-1)) x 256 pixel tile into one bit per
// pixel with warp ballots (coalesced 128-byte row loads), runs the four
// separable min/max passes as log-step shift-AND / shift-OR on 32-bit words in
// shared memory, and unpacks the central 64 x 192 block.  Exact (boolean).
// HBM-bound: 4 B/px read (+ halo re-reads served by L2) + 4 B/px written.
#include "common.cuh"

namespace sb {

constexpr int kMorphTW = 8;                  // words per tile row (256 px)
constexpr int kMorphOutW = (kMorphTW - 2) * 32;  // 192 output columns, 32 px halo each side
constexpr int kMorphTY = 64;                 // output rows per tile
constexpr int kMorphMaxK = 33;               // kernel side <= 33 (halo 32 >= kw - 1)
constexpr int kMorphRowsMax = kMorphTY + 2 * (kMorphMaxK - 1);  // 128

// out[x] = in[x + s] (towards lower x), zero fill past the tile.
__device__ __forceinline__ uint32_t row_shift_down(const uint32_t* row, int w, int s) {
  const uint32_t lo = row[w], hi = (w + 1 < kMorphTW) ? row[w + 1] : 0u;
  return __funnelshift_r(lo, hi, s);
}
// out[x] = in[x - s]
__device__ __forceinline__ uint32_t row_shift_up(const uint32_t* row, int w, int s) {
  const uint32_t hi = row[w], lo = (w > 0) ? row[w - 1] : 0u;
  return __funnelshift_l(lo, hi, s);
}

// One separable pass along x over `rows` rows: window [x - r, x + r], AND or OR.
template <bool IS_AND>
__device__ void pass_x(uint32_t*& cur, uint32_t*& nxt, int rows, int k) {
  const int n = rows * kMorphTW;
  int win = 1;
  while (win < k) {
    const int s = min(win, k - win);
    for (int i = threadIdx.x; i < n; i += blockDim.x) {
      const int y = i / kMorphTW, w = i - y * kMorphTW;
      const uint32_t a = cur[i], b = row_shift_down(cur + y * kMorphTW, w, s);
      nxt[i] = IS_AND ? (a & b) : (a | b);
    }
    __syncthreads();
    uint32_t* t = cur; cur = nxt; nxt = t;
    win += s;
  }
  // centre: window [x, x+k-1] -> [x-r, x+r]
  const int r = k / 2;
  if (r > 0) {
    for (int i = threadIdx.x; i < n; i += blockDim.x) {
      const int y = i / kMorphTW, w = i - y * kMorphTW;
      nxt[i] = row_shift_up(cur + y * kMorphTW, w, r);
    }
    __syncthreads();
    uint32_t* t = cur; cur = nxt; nxt = t;
  }
}

// One separable pass along y: out[y] = op over rows [y, y + k - 1] (caller
// accounts for the k/2 row offset). Rows past `rows` read as `fill`.
template <bool IS_AND>
__device__ void pass_y(uint32_t*& cur, uint32_t*& nxt, int rows, int k) {
  const int n = rows * kMorphTW;
  int win = 1;
  while (win < k) {
    const int s = min(win, k - win);
    for (int i = threadIdx.x; i < n; i += blockDim.x) {
      const int y = i / kMorphTW;
      const uint32_t a = cur[i];
      const uint32_t b = (y + s < rows) ? cur[i + s * kMorphTW] : (IS_AND ? 0u : 0u);
      nxt[i] = IS_AND ? (a & b) : (a | b);
    }
    __syncthreads();
    uint32_t* t = cur; cur = nxt; nxt = t;
    win += s;
  }
}

__global__ void __launch_bounds__(256)
morph_open_kernel(const float* __restrict__ mask, float* __restrict__ out, int H, int W, int kh,
                  int kw, int border_is_zero, int tiles_x, int tiles_y) {
  __shared__ uint32_t s_a[kMorphRowsMax * kMorphTW];
  __shared__ uint32_t s_b[kMorphRowsMax * kMorphTW];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int tile = blockIdx.x;
  const int tx = tile % tiles_x, ty = (tile / tiles_x) % tiles_y;
  const long long pl = tile / (tiles_x * tiles_y);
  const int rh = kh / 2, rw = kw / 2;
  const int rows = kMorphTY + 4 * rh;              // erosion halo + dilation halo
  const int x0 = tx * kMorphOutW - 32;             // image column of tile bit 0
  const int y0 = ty * kMorphTY - 2 * rh;           // image row of tile row 0
  const float* src = mask + pl * (long long)H * W;

  // ---- pack: bit = (mask >= 0.5) inside the image, `border` outside
  const uint32_t border = border_is_zero ? 0u : 1u;
  // 8 row-words per warp and step: 8 independent 128-byte loads in flight before the first ballot
  // (one load per step left the pack phase latency-bound: 20 warps x 128 B outstanding per SM)
  constexpr int kPackU = 8;
  const int n_words = rows * kMorphTW, n_warps = blockDim.x >> 5;
  for (int i0 = warp; i0 < n_words; i0 += n_warps * kPackU) {
    float v[kPackU];
    bool ok[kPackU];
#pragma unroll
    for (int u = 0; u < kPackU; ++u) {
      const int i = i0 + u * n_warps;
      const int y = i / kMorphTW, w = i - y * kMorphTW;
      const int gy = y0 + y, gx = x0 + w * 32 + lane;
      ok[u] = (i < n_words) && gy >= 0 && gy < H && gx >= 0 && gx < W;
      v[u] = ok[u] ? ldg_stream(src + (long long)gy * W + gx) : 0.0f;
    }
#pragma unroll
    for (int u = 0; u < kPackU; ++u) {
      const int i = i0 + u * n_warps;
      const uint32_t bit = ok[u] ? ((v[u] >= 0.5f) ? 1u : 0u) : border;
      const uint32_t word = __ballot_sync(0xffffffffu, bit);
      if (lane == 0 && i < n_words) s_a[i] = word;
    }
  }
  __syncthreads();
  uint32_t* cur = s_a;
  uint32_t* nxt = s_b;

  // ---- erosion: AND over [x-rw, x+rw] then over rows [y, y+kh-1]
  pass_x<true>(cur, nxt, rows, kw);
  pass_y<true>(cur, nxt, rows, kh);        // cur[y] now holds erosion at image row y0 + y + rh
  // eroded pixels outside the image never feed the dilation (zero padding / cv2 -inf border)
  for (int i = threadIdx.x; i < rows * kMorphTW; i += blockDim.x) {
    const int y = i / kMorphTW, w = i - y * kMorphTW;
    const int gy = y0 + y + rh;
    uint32_t inside = 0u;
    if (gy >= 0 && gy < H) {
      const int gx = x0 + w * 32;
      // bits b with 0 <= gx + b < W
      const int lo = max(0, -gx), hi = min(32, W - gx);
      if (hi > lo) inside = ((hi - lo) == 32) ? 0xffffffffu : (((1u << (hi - lo)) - 1u) << lo);
    }
    cur[i] &= inside;
  }
  __syncthreads();
  // ---- dilation: OR over [x-rw, x+rw], then rows [y, y+kh-1]
  pass_x<false>(cur, nxt, rows, kw);
  pass_y<false>(cur, nxt, rows, kh);       // cur[y] = dilation at image row y0 + y + 2*rh = ty*TY + y

  // ---- unpack the central block
  for (int i = threadIdx.x; i < kMorphTY * kMorphOutW; i += blockDim.x) {
    const int y = i / kMorphOutW, x = i - y * kMorphOutW;
    const int gy = ty * kMorphTY + y, gx = tx * kMorphOutW + x;
    if (gy < H && gx < W) {
      const int bx = x + 32;
      const uint32_t word = cur[y * kMorphTW + (bx >> 5)];
      stg_stream(out + pl * (long long)H * W + (long long)gy * W + gx, ((word >> (bx & 31)) & 1u) ? 1.0f : 0.0f);
    }
  }
}

}  // namespace sb

extern "C" int sb_morph_open(const float* mask, float* out, int P, int H, int W, int kh, int kw,
                             int border_is_zero, sb_stream_t stream) {
  using namespace sb;
  SB_ENTER();

  SB_REQUIRE(P >= 0 && H >= 0 && W >= 0, SB_EINVAL, "sb_morph_open: bad size");
  SB_REQUIRE(kh >= 1 && kw >= 1 && (kh & 1) && (kw & 1), SB_EUNSUP,
             "sb_morph_open: kernel %dx%d must be odd", kh, kw);
  SB_REQUIRE(kh <= kMorphMaxK && kw <= kMorphMaxK, SB_EUNSUP, "sb_morph_open: kernel %dx%d > %d", kh,
             kw, kMorphMaxK);
  if ((long long)P * H * W == 0) return SB_OK;
  SB_REQUIRE(mask && out, SB_EINVAL, "sb_morph_open: null pointer");
  const int tiles_x = (W + kMorphOutW - 1) / kMorphOutW, tiles_y = (H + kMorphTY - 1) / kMorphTY;
  const long long tiles = (long long)P * tiles_x * tiles_y;
  SB_REQUIRE(tiles < (1ll << 31), SB_EUNSUP, "sb_morph_open: too many tiles");
  morph_open_kernel<<<(int)tiles, 256, 0, as_stream(stream)>>>(mask, out, H, W, kh, kw,
                                                               border_is_zero, tiles_x, tiles_y);
  SB_LAUNCH_CHECK("morph_open_kernel");
  return SB_OK;
}
