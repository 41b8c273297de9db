// upsample_flow.cu — convex 8x flow upsampling (RAFT / FlowFormer).
// Replaces MemoryDecoder.upsample_flow(flow, mask)  (core/FlowFormer/PerCostFormer3/decoder.py:214-225):
//   mask [N, 9*8*8, H, W] -> view [N, 1, 9, 8, 8, H, W], softmax over the 9 taps        (:217-218)
//   up   = unfold(8 * flow, 3x3, padding=1) -> [N, 2, 9, 1, 1, H, W]                     (:220-221)
//   out[n, c, 8h+dy, 8w+dx] = sum_k softmax_k(mask[n, k, dy, dx, h, w]) * 8*flow[n, c, h+ky-1, w+kx-1]   (:223-225)
// (k = ky*3 + kx, zero padding outside the coarse map).
//
// HBM-bound: 576 mask floats read and 128 flow floats written per coarse pixel (2.8 KB); the 9
// coarse neighbours come from L1/L2.  A CTA owns 32 consecutive coarse pixels of one row: every
// mask channel row is read as one coalesced 128-byte segment per warp, the 2 x 8 x 256 output
// tile is transposed through shared memory and leaves as full 128-byte lines.
#include "common.cuh"

namespace sb {

constexpr int kUpW = 32;   // coarse pixels per CTA

__global__ void __launch_bounds__(256)
upsample_flow_kernel(const float* __restrict__ flow, const float* __restrict__ mask,
                     float* __restrict__ out, int H, int W) {
  __shared__ float s_out[2][8][kUpW * 8 + 4];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int w0 = blockIdx.x * kUpW, h = blockIdx.y, n = blockIdx.z;
  const int w = w0 + lane;
  const size_t plane = (size_t)H * W;
  const bool in = w < W;
  // the 9 coarse neighbours of (h, w), both components, times 8 (:220)
  float f[2][9];
#pragma unroll
  for (int c = 0; c < 2; ++c)
#pragma unroll
    for (int k = 0; k < 9; ++k) {
      const int hy = h + k / 3 - 1, wx = w + k % 3 - 1;
      const bool ok = in && hy >= 0 && hy < H && wx >= 0 && wx < W;
      f[c][k] = ok ? fmul(8.0f, __ldg(flow + ((size_t)n * 2 + c) * plane + (size_t)hy * W + wx)) : 0.0f;
    }
  // warp `warp` handles dy = warp (8 warps), lanes = coarse x, loop over dx
  const int dy = warp;
  // channel (k, dy, dx) = k*64 + dy*8 + dx: tap k is 64 planes further, dx one plane further
  const float* mrow = mask + ((size_t)n * 576 + dy * 8) * plane + (size_t)h * W + w;
  const size_t tap_stride = 64 * plane;
#pragma unroll 2
  for (int dx = 0; dx < 8; ++dx) {
    float m[9];
    const float* mp = mrow;
#pragma unroll
    for (int k = 0; k < 9; ++k) {
      m[k] = in ? ldg_stream(mp) : 0.0f;
      mp += tap_stride;
    }
    mrow += plane;
    // softmax over k as ATen evaluates it on a non-last dim: max, exp(x - max), sum, normalise
    float mx = m[0];
#pragma unroll
    for (int k = 1; k < 9; ++k) mx = fmaxf(mx, m[k]);
    float e[9], sum = 0.0f;
#pragma unroll
    for (int k = 0; k < 9; ++k) {
      e[k] = expf(fsub(m[k], mx));
      sum = fadd(sum, e[k]);
    }
    const float inv = fdiv(1.0f, sum);            // one division per output pixel; e * inv is within 1 ulp of e / sum
    float a0 = 0.0f, a1 = 0.0f;
#pragma unroll
    for (int k = 0; k < 9; ++k) {
      const float p = fmul(e[k], inv);
      a0 = fadd(a0, fmul(p, f[0][k]));
      a1 = fadd(a1, fmul(p, f[1][k]));
    }
    s_out[0][dy][lane * 8 + dx] = a0;
    s_out[1][dy][lane * 8 + dx] = a1;
  }
  __syncthreads();
  // write 2 x 8 rows of (kUpW * 8) floats, coalesced
  const int Wo = 8 * W, Ho = 8 * H;
  const int ncols = min(kUpW, W - w0) * 8;
  for (int r = warp; r < 16; r += 8) {
    const int c = r >> 3, ry = r & 7;
    float* orow = out + (((size_t)n * 2 + c) * Ho + (size_t)(8 * h + ry)) * Wo + (size_t)w0 * 8;
    for (int x = lane; x < ncols; x += 32) stg_stream(orow + x, s_out[c][ry][x]);
  }
}

}  // namespace sb

extern "C" int sb_upsample_flow(const float* flow, const float* mask, float* out, int N, int H, int W,
                                sb_stream_t stream) {
  using namespace sb;
  SB_ENTER();
  SB_REQUIRE(N >= 0 && H >= 0 && W >= 0, SB_EINVAL, "sb_upsample_flow: bad size");
  SB_REQUIRE((long long)H * W * 64 < (1ll << 31), SB_EUNSUP, "sb_upsample_flow: plane too large");
  if ((long long)N * H * W == 0) return SB_OK;
  SB_REQUIRE(flow && mask && out, SB_EINVAL, "sb_upsample_flow: null pointer");
  SB_REQUIRE(N <= 65535 && H <= 65535, SB_EUNSUP, "sb_upsample_flow: N or H too large for one launch");
  const dim3 grid((W + kUpW - 1) / kUpW, H, N);
  upsample_flow_kernel<<<grid, 256, 0, as_stream(stream)>>>(flow, mask, out, H, W);
  SB_LAUNCH_CHECK("upsample_flow_kernel");
  return SB_OK;
}
