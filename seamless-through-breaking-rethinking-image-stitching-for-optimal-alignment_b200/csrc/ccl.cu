// ccl.cu — N3 ("next" row 3): the contextual correlation layer of the UDIS2 homography net.
// Replaces UDIS2Network.CCL(feature_1, feature_2)  (core/UDIS2/Homography/network.py:147-199):
//   nf = F.normalize(f, p=2, dim=1)                                                (:150-151)
//   match[q, p] = conv2d(nf1, 3x3 patches of nf2 as filters, padding=1)            (:153-164)
//               = sum_{c, d in 3x3} nf1[c, p + d] * nf2[c, q + d]   (zero outside either map)
//   prob = softmax(10 * match, over q)                                              (:166-168)
//   flow_h[p] = sum_q prob[q, p] * (q // w - p_y),  flow_w[p] = sum_q prob * (q % w - p_x)   (:170-197)
//   out = cat([flow_w, flow_h], 1)                                                  (:199)
//
// The reference evaluates the h*w x (9*c) x h*w contraction as a Python loop of per-sample conv2d
// calls (19.3 GFLOP per pair at 32 x 32 x 1024) and materialises match / softmax / three index
// volumes.  Here:  match[q, p] = sum_{d in 3x3} C0[p + d, q + d]  with the plain all-pairs
// correlation C0 = nf1^T nf2 (2.1 GFLOP: 9x fewer) computed once on the TF32 tensor cores
// (sb_gemm_nt_tf32, gma.cu) from normalised token-major copies; one kernel then sums the nine
// shifted diagonals, takes the softmax over q and the expected displacement per position p —
// neither the match volume nor the softmax ever exist in memory.
#include "common.cuh"
#include "ptx_sm100.cuh"
#include "tmap.cuh"

extern "C" int sb_gemm_nt_tf32(const float* A, const float* B, float* D, int BH, int M, int N, int K,
                               sb_stream_t stream);

namespace sb {

// denom[b, n] = max(||f[b, :, n]||_2, 1e-12)   (F.normalize's clamp_min(eps))
__global__ void __launch_bounds__(256)
ccl_norm_kernel(const float* __restrict__ f, float* __restrict__ denom, int C, int N) {
  const int n = blockIdx.x * 32 + (threadIdx.x & 31), b = blockIdx.y;
  const int cw = threadIdx.x >> 5;                 // 8 channel slices
  __shared__ float s_part[8][33];
  float acc = 0.0f;
  if (n < N)
    for (int c = cw; c < C; c += 8) {
      const float v = __ldg(f + ((size_t)b * C + c) * N + n);
      acc = __fmaf_rn(v, v, acc);
    }
  s_part[cw][threadIdx.x & 31] = acc;
  __syncthreads();
  if (cw == 0 && n < N) {
    float t = 0.0f;
#pragma unroll
    for (int k = 0; k < 8; ++k) t += s_part[k][threadIdx.x];
    denom[(size_t)b * N + n] = fmaxf(sqrtf(t), 1e-12f);
  }
}

__device__ __forceinline__ float to_tf32(float x) {
  uint32_t u;
  asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(u) : "f"(x));
  return __uint_as_float(u);
}

// tok[b, n, c] = tf32(f[b, c, n] / denom[b, n]): tiles of 128 channels x 32 positions; reads are
// 128-byte rows of one channel, writes 512-byte rows of one position (float4 per lane)
__global__ void __launch_bounds__(256)
ccl_tokens_kernel(const float* __restrict__ f, const float* __restrict__ denom, float* __restrict__ tok,
                  int C, int N) {
  __shared__ float s[128][33];
  const int b = blockIdx.z, c0 = blockIdx.y * 128, n0 = blockIdx.x * 32;
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
#pragma unroll 4
  for (int r = ty; r < 128; r += 8) {
    const int c = c0 + r, n = n0 + tx;
    s[r][tx] = (c < C && n < N) ? ldg_stream(f + ((size_t)b * C + c) * N + n) : 0.0f;
  }
  __syncthreads();
#pragma unroll
  for (int r = ty; r < 32; r += 8) {
    const int n = n0 + r, c = c0 + tx * 4;
    if (n < N && c < C) {                       // C % 4 == 0: a float4 is entirely in or out
      const float d = __ldg(denom + (size_t)b * N + n);
      float4 v;
      v.x = to_tf32(fdiv(s[tx * 4 + 0][r], d)); v.y = to_tf32(fdiv(s[tx * 4 + 1][r], d));
      v.z = to_tf32(fdiv(s[tx * 4 + 2][r], d)); v.w = to_tf32(fdiv(s[tx * 4 + 3][r], d));
      *reinterpret_cast<float4*>(tok + ((size_t)b * N + n) * C + c) = v;
    }
  }
}

// Fused normalise + transpose (C <= kTokMaxC, N % 4 == 0): a CTA holds ALL channels of 16 positions in
// shared memory, so the feature map is read from HBM once.  The [C x 16] tile arrives as TMA boxes of
// 256 channel rows x 64 bytes with the 64-byte swizzle (the whole tile is in flight at once and three
// CTAs per SM overlap each other's load / compute / store phases; a plain load loop was latency-bound
// at 158 us, one 32-position tile per SM took 62 us).  Sum of squares per position from 8 channel
// slices, then lanes = channels read four positions as one conflict-free LDS.128 and write the
// normalised TF32 tokens as 128-byte pieces of four token rows.
// Swizzle<2,4,3>: the 16-byte chunk index of row c (64-byte rows) is XORed with (c >> 1) & 3.
constexpr int kTokMaxC = 1024;                       // 1024 rows * 64 B = 64 KB: three CTAs per SM
constexpr int kTokPos = 16;
constexpr int kTokThreads = 256, kTokWarps = kTokThreads / 32;

struct CclTokSmem {
  static constexpr int kTail = kTokWarps * kTokPos * 4 + 2 * kTokPos * 4 + 16; // partial sums, norms, reciprocals, mbarrier
  __host__ __device__ static size_t bytes(int C) { return (size_t)((C + 255) / 256 * 256) * 64 + kTail + 1024; }
};

// a / d from r = RN(1/d): q = RN(a r), one Newton residual step.  Correctly rounded except in rare
// double-rounding cases (then 1 ulp off), well below the TF32 rounding that follows; 3 instructions
// instead of the ~20 of an IEEE division per element.
__device__ __forceinline__ float div_by(float a, float d, float r) {
  const float q = fmul(a, r);
  return __fmaf_rn(__fmaf_rn(-q, d, a), r, q);
}

__device__ __forceinline__ int tok_word(int c, int pos) {               // word index of (channel, position) in the tile
  return c * kTokPos + ((((pos >> 2) ^ ((c >> 1) & 3)) << 2) | (pos & 3));
}

__global__ void __launch_bounds__(kTokThreads)
ccl_norm_tokens_kernel(const __grid_constant__ CUtensorMap map_f, float* __restrict__ tok, int C, int N,
                       unsigned int* dbg) {
  extern __shared__ uint8_t s_tok_raw[];
  const uint32_t raw = ptx::smem_u32(s_tok_raw);
  uint8_t* base = s_tok_raw + ((1024u - (raw & 1023u)) & 1023u);     // swizzle atoms start at aligned addresses
  const int Cp = (C + 255) / 256 * 256;
  const float* s_feat = reinterpret_cast<const float*>(base);          // [Cp][16] floats, swizzled
  float* s_part = reinterpret_cast<float*>(base + (size_t)Cp * 64);    // [kTokWarps][16]
  float* s_den = s_part + kTokWarps * kTokPos;                         // [16] norms (16-byte aligned)
  float* s_rcp = s_den + kTokPos;                                      // [16] their reciprocals
  const uint32_t bar = ptx::smem_u32(s_rcp + kTokPos);
  const int b = blockIdx.y, n0 = blockIdx.x * kTokPos;
  const int lane = threadIdx.x & 31, ty = threadIdx.x >> 5;
  if (threadIdx.x == 0) {
    ptx::mbar_init(bar, 1);
    ptx::fence_mbar_init();
    ptx::mbar_arrive_expect_tx(bar, (uint32_t)Cp * 64u);
    for (int c0 = 0; c0 < Cp; c0 += 256)                              // rows >= C and columns >= N are zero-filled
      ptx::tma_load_3d(ptx::smem_u32(base) + (uint32_t)c0 * 64u, &map_f, bar, n0, c0, b);
  }
  __syncthreads();
  ptx::mbar_wait(bar, 0, 0x70, dbg);
  {                                                  // lanes 0-15: even channels, 16-31: odd channels (32 distinct banks)
    const int pos = lane & 15, sub = lane >> 4;
    float acc = 0.0f;
    for (int c = 2 * ty + sub; c < C; c += 2 * kTokWarps) {
      const float v = s_feat[tok_word(c, pos)];
      acc = __fmaf_rn(v, v, acc);
    }
    acc += __shfl_xor_sync(0xffffffffu, acc, 16);
    if (sub == 0) s_part[ty * kTokPos + pos] = acc;
  }
  __syncthreads();
  if (threadIdx.x < kTokPos) {
    float t = 0.0f;
#pragma unroll
    for (int k = 0; k < kTokWarps; ++k) t += s_part[k * kTokPos + threadIdx.x];
    const float d = fmaxf(sqrtf(t), 1e-12f);
    s_den[threadIdx.x] = d;
    s_rcp[threadIdx.x] = fdiv(1.0f, d);
  }
  __syncthreads();
  const int items = (C / 32) * (kTokPos / 4);        // (32-channel block, 4-position group); C % 32 tail below
  for (int it = ty; it < items; it += kTokWarps) {
    const int g = it & 3, c = (it >> 2) * 32 + lane;
    const float4 v = *reinterpret_cast<const float4*>(s_feat + tok_word(c, g * 4));
    const float4 d = *reinterpret_cast<const float4*>(s_den + g * 4);
    const float vv[4] = {v.x, v.y, v.z, v.w}, dd[4] = {d.x, d.y, d.z, d.w};
    float* dst = tok + ((size_t)b * N + n0 + g * 4) * C + c;
    const int nleft = N - (n0 + g * 4);
#pragma unroll
    for (int k = 0; k < 4; ++k)
      if (k < nleft) dst[(size_t)k * C] = to_tf32(div_by(vv[k], dd[k], s_rcp[g * 4 + k]));
  }
  for (int c = (C / 32) * 32 + lane; c < C; c += 32)   // C % 32 != 0: the last partial channel block
    for (int r = ty; r < kTokPos; r += kTokWarps)
      if (n0 + r < N) tok[((size_t)b * N + n0 + r) * C + c] = to_tf32(div_by(s_feat[tok_word(c, r)], s_den[r], s_rcp[r]));
}

// One WARP per position p (8 positions per CTA, no block-level synchronisation): the nine shifted
// rows of C0 are summed into a shared-memory row (lanes stride q, so every load is a coalesced
// 128-byte segment and neighbouring warps re-use each other's rows from L1), then softmax over q
// and the expected displacement with warp shuffles.
__global__ void __launch_bounds__(256)
ccl_flow_kernel(const float* __restrict__ c0, float* __restrict__ out, int H, int W, float scale) {
  extern __shared__ float s_rows[];             // [8][N]
  const int N = H * W;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int p = blockIdx.x * 8 + warp, b = blockIdx.y;
  if (p >= N) return;
  float* row = s_rows + (size_t)warp * N;
  const float inv_w = 1.0f / (float)W;
  const int py = __float2int_rd(((float)p + 0.5f) * inv_w), px = p - py * W;   // exact for p < 2^22
  const float* cb = c0 + (size_t)b * N * N;
  float mx = -INFINITY;
  for (int q = lane; q < N; q += 32) {
    const int qy = __float2int_rd(((float)q + 0.5f) * inv_w), qx = q - qy * W;
    float acc = 0.0f;
#pragma unroll
    for (int dy = -1; dy <= 1; ++dy)
#pragma unroll
      for (int dx = -1; dx <= 1; ++dx) {
        const bool ok = (unsigned)(py + dy) < (unsigned)H && (unsigned)(px + dx) < (unsigned)W &&
                        (unsigned)(qy + dy) < (unsigned)H && (unsigned)(qx + dx) < (unsigned)W;
        if (ok) acc += __ldg(cb + (size_t)(p + dy * W + dx) * N + (q + dy * W + dx));
      }
    acc *= scale;
    row[q] = acc;
    mx = fmaxf(mx, acc);
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
  float se = 0.0f, sh = 0.0f, sw = 0.0f;
  for (int q = lane; q < N; q += 32) {
    const int qy = __float2int_rd(((float)q + 0.5f) * inv_w), qx = q - qy * W;
    const float e = expf(row[q] - mx);
    se += e;
    sh = __fmaf_rn(e, (float)(qy - py), sh);
    sw = __fmaf_rn(e, (float)(qx - px), sw);
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    se += __shfl_xor_sync(0xffffffffu, se, o);
    sh += __shfl_xor_sync(0xffffffffu, sh, o);
    sw += __shfl_xor_sync(0xffffffffu, sw, o);
  }
  if (lane == 0) {
    out[((size_t)b * 2 + 0) * N + p] = sw / se;   // channel 0 = flow_w, 1 = flow_h  (:199)
    out[((size_t)b * 2 + 1) * N + p] = sh / se;
  }
}

// Staged variant (the one the network's shapes take): a CTA owns kFlowG consecutive positions p.  The
// rows of C0 they need are three bands of kFlowG + 2 consecutive rows (p + dy*W + dx); the CTA copies
// those bands into shared memory once (one cp.async.bulk per row) and every warp then sums its nine shifted
// rows from there, so C0 is read 4.5 times per position instead of 9 times through L1.  q is walked in
// chunks of kFlowQC columns with a running (max, sums) update, so N is not limited by shared memory.
constexpr int kFlowG = 4;
constexpr int kFlowWarpsPerP = 2;                   // warps sharing one position (each takes every other 32 q)
constexpr int kFlowQC = 1024;
constexpr int kFlowBandRows = kFlowG + 2;

__host__ __device__ inline int ccl_stage_cols(int N, int W) {
  const int c = kFlowQC + 2 * W + 2 + 6;            // + alignment slack (col0 is rounded down to 4)
  return ((N < c ? N : c) + 3) & ~3;
}

__global__ void __launch_bounds__(kFlowG * kFlowWarpsPerP * 32)
ccl_flow_staged_kernel(const float* __restrict__ c0, float* __restrict__ out, int H, int W, float scale,
                       unsigned int* dbg) {
  extern __shared__ float4 s_flow4[];
  const int N = H * W;
  const int SC = ccl_stage_cols(N, W);              // staged columns per row (multiple of 4)
  float* s_band = reinterpret_cast<float*>(s_flow4);                  // [3][kFlowBandRows][SC]
  float* s_val = s_band + 3 * kFlowBandRows * SC;                     // [kFlowG][kFlowQC]
  float2* s_qf = reinterpret_cast<float2*>(s_val + kFlowG * kFlowQC); // [kFlowQC] (qy, qx) of the chunk's q as floats
  uint32_t* s_qm = reinterpret_cast<uint32_t*>(s_qf + kFlowQC);       // [kFlowQC] taps with (qy+dy, qx+dx) in the map
  float* s_red = reinterpret_cast<float*>(s_qm + kFlowQC);            // [warps][4]
  const uint32_t bar = ptx::smem_u32(s_red + kFlowG * kFlowWarpsPerP * 4);
  uint32_t phase = 0;
  if (threadIdx.x == 0) {
    ptx::mbar_init(bar, 1);
    ptx::fence_mbar_init();
  }
  __syncthreads();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int slot = warp / kFlowWarpsPerP, part = warp - slot * kFlowWarpsPerP;   // position slot, share of q
  const int p0 = blockIdx.x * kFlowG, b = blockIdx.y;
  const int p = p0 + slot;
  const bool live = p < N;
  const float inv_w = 1.0f / (float)W;
  const int py = __float2int_rd(((float)p + 0.5f) * inv_w), px = p - py * W;   // exact for p < 2^22
  // taps whose (py + dy, px + dx) lies in the map: bit (dy+1)*3 + dx+1
  unsigned pmask = 0;
#pragma unroll
  for (int t = 0; t < 9; ++t)
    if ((unsigned)(py + t / 3 - 1) < (unsigned)H && (unsigned)(px + t % 3 - 1) < (unsigned)W) pmask |= 1u << t;
  const float* cb = c0 + (size_t)b * N * N;
  float* val = s_val + slot * kFlowQC;
  constexpr int kStep = 32 * kFlowWarpsPerP;
  float mx = -INFINITY, se = 0.0f, sh = 0.0f, sw = 0.0f;
  for (int q0 = 0; q0 < N; q0 += kFlowQC) {
    const int q1 = min(q0 + kFlowQC, N);
    const int col0 = max(q0 - W - 1, 0) & ~3;
    const int col1 = min((q1 + W + 1 + 3) & ~3, N);  // N % 4 == 0 (checked by the host)
    // ---- stage rows r = p0 + dy*W - 1 + j, j < kFlowBandRows, columns [col0, col1): one bulk copy per row
    if (q0 > 0) __syncthreads();                    // previous chunk fully consumed
    if (threadIdx.x == 0) {
      if (q0 > 0) ptx::fence_proxy_async_smem();
      const uint32_t row_bytes = (uint32_t)(col1 - col0) * 4u;
      int rows = 0;
      for (int rj = 0; rj < 3 * kFlowBandRows; ++rj) {
        const int r = p0 + (rj / kFlowBandRows - 1) * W - 1 + rj % kFlowBandRows;
        rows += (r >= 0 && r < N);
      }
      ptx::mbar_arrive_expect_tx(bar, row_bytes * (uint32_t)rows);
      for (int rj = 0; rj < 3 * kFlowBandRows; ++rj) {
        const int r = p0 + (rj / kFlowBandRows - 1) * W - 1 + rj % kFlowBandRows;
        if (r < 0 || r >= N) continue;              // never read: every use is guarded by pmask
        ptx::bulk_load_1d(ptx::smem_u32(s_band + (size_t)rj * SC), cb + (size_t)r * N + col0, row_bytes, bar);
      }
    }
    // per-q tables of the chunk (independent of p): coordinates as floats, in-map taps as a 9-bit mask
    for (int q = q0 + (int)threadIdx.x; q < q1; q += (int)blockDim.x) {
      const int qy = __float2int_rd(((float)q + 0.5f) * inv_w), qx = q - qy * W;
      const unsigned rows_ok = (qy > 0 ? 0x007u : 0u) | 0x038u | (qy < H - 1 ? 0x1c0u : 0u);
      const unsigned cols_ok = (qx > 0 ? 0x049u : 0u) | 0x092u | (qx < W - 1 ? 0x124u : 0u);
      s_qf[q - q0] = make_float2((float)qy, (float)qx);
      s_qm[q - q0] = rows_ok & cols_ok;
    }
    __syncthreads();
    ptx::mbar_wait(bar, phase, 0x71, dbg);
    phase ^= 1u;
    float cmx = -INFINITY;
    if (live) {
      const float* sb0 = s_band + (size_t)(slot + 1) * SC - col0;     // row (dy = -1, dx = 0) of this position
#pragma unroll 2
      for (int q = q0 + part * 32 + lane; q < q1; q += kStep) {
        const unsigned m = pmask & s_qm[q - q0];
        float acc = 0.0f;
#pragma unroll
        for (int dy = -1; dy <= 1; ++dy)
#pragma unroll
          for (int dx = -1; dx <= 1; ++dx)
            if (m & (1u << ((dy + 1) * 3 + dx + 1)))
              acc += sb0[((dy + 1) * kFlowBandRows + dx) * SC + (q + dy * W + dx)];
        acc *= scale;
        val[q - q0] = acc;
        cmx = fmaxf(cmx, acc);
      }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) cmx = fmaxf(cmx, __shfl_xor_sync(0xffffffffu, cmx, o));
    if (lane == 0) s_red[warp * 4] = cmx;
    __syncthreads();
#pragma unroll
    for (int k = 0; k < kFlowWarpsPerP; ++k) cmx = fmaxf(cmx, s_red[(slot * kFlowWarpsPerP + k) * 4]);
    if (live) {
      if (cmx > mx) {                               // rescale what earlier chunks accumulated
        const float f = expf(mx - cmx);             // exp(-inf) = 0 on the first chunk
        se *= f; sh *= f; sw *= f;
        mx = cmx;
      }
      const float fpy = (float)py, fpx = (float)px;
#pragma unroll 4
      for (int q = q0 + part * 32 + lane; q < q1; q += kStep) {       // the values this warp wrote itself
        const float2 qf = s_qf[q - q0];
        const float e = exp2f(fmul(fsub(val[q - q0], mx), 1.4426950408889634f));
        se += e;
        sh = __fmaf_rn(e, fsub(qf.x, fpy), sh);       // small integers: the differences are exact
        sw = __fmaf_rn(e, fsub(qf.y, fpx), sw);
      }
    }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    se += __shfl_xor_sync(0xffffffffu, se, o);
    sh += __shfl_xor_sync(0xffffffffu, sh, o);
    sw += __shfl_xor_sync(0xffffffffu, sw, o);
  }
  __syncthreads();                                  // s_red: the chunk maxima have been read
  if (lane == 0) { s_red[warp * 4] = se; s_red[warp * 4 + 1] = sh; s_red[warp * 4 + 2] = sw; }
  __syncthreads();
  if (live && part == 0 && lane == 0) {
    float tse = 0.0f, tsh = 0.0f, tsw = 0.0f;
#pragma unroll
    for (int k = 0; k < kFlowWarpsPerP; ++k) {
      const float* r = s_red + (slot * kFlowWarpsPerP + k) * 4;
      tse += r[0]; tsh += r[1]; tsw += r[2];
    }
    out[((size_t)b * 2 + 0) * N + p] = tsw / tse;   // channel 0 = flow_w, 1 = flow_h  (:199)
    out[((size_t)b * 2 + 1) * N + p] = tsh / tse;
  }
}

static inline size_t align256(size_t x) { return (x + 255) / 256 * 256; }

}  // namespace sb

extern "C" size_t sb_ccl_workspace_bytes(int B, int C, int H, int W) {
  if (B < 0 || C < 0 || H < 0 || W < 0) return 0;
  const size_t N = (size_t)H * W;
  return 2 * sb::align256((size_t)B * N * C * 4) + 2 * sb::align256((size_t)B * N * 4) + sb::align256((size_t)B * N * N * 4);
}

extern "C" int sb_ccl(const float* feature_1, const float* feature_2, float* flow, void* workspace,
                      size_t workspace_bytes, int B, int C, int H, int W, float softmax_scale,
                      sb_stream_t stream) {
  using namespace sb;
  SB_ENTER();
  SB_REQUIRE(B >= 0 && C > 0 && H >= 0 && W >= 0, SB_EINVAL, "sb_ccl: bad size");
  const long long N = (long long)H * W;
  if (B == 0 || N == 0) return SB_OK;
  SB_REQUIRE(feature_1 && feature_2 && flow, SB_EINVAL, "sb_ccl: null pointer");
  SB_REQUIRE((C & 3) == 0, SB_EUNSUP, "sb_ccl: C must be a multiple of 4");
  SB_REQUIRE(B <= 65535, SB_EUNSUP, "sb_ccl: B > 65535");
  const size_t need = sb_ccl_workspace_bytes(B, C, H, W);
  SB_REQUIRE(workspace && workspace_bytes >= need && (reinterpret_cast<uintptr_t>(workspace) & 255) == 0, SB_EINVAL,
             "sb_ccl: workspace must be 256-byte aligned and >= %zu bytes", need);
  char* ws = static_cast<char*>(workspace);
  float* tok1 = reinterpret_cast<float*>(ws); ws += align256((size_t)B * N * C * 4);
  float* tok2 = reinterpret_cast<float*>(ws); ws += align256((size_t)B * N * C * 4);
  float* den1 = reinterpret_cast<float*>(ws); ws += align256((size_t)B * N * 4);
  float* den2 = reinterpret_cast<float*>(ws); ws += align256((size_t)B * N * 4);
  float* c0 = reinterpret_cast<float*>(ws);
  cudaStream_t s = as_stream(stream);
  const dim3 gn((unsigned)((N + 31) / 32), B);
  const dim3 gt((unsigned)((N + 31) / 32), (unsigned)((C + 127) / 128), B);
  unsigned int* dbg_word = debug_word_device();
  if (!dbg_word) return SB_ECUDA;
  if (C <= kTokMaxC && (N & 3) == 0 && aligned16(feature_1) && aligned16(feature_2)) {
    const size_t smem = CclTokSmem::bytes(C);
    static SmemOptIn tok_opt_in;
    int opt_dev;
    if (tok_opt_in.need(smem, &opt_dev)) {
      SB_CUDA(cudaFuncSetAttribute(ccl_norm_tokens_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
      tok_opt_in.done(smem, opt_dev);
    }
    const float* feats[2] = {feature_1, feature_2};
    float* toks[2] = {tok1, tok2};
    for (int i = 0; i < 2; ++i) {
      CUtensorMap map_f;
      const int rc_map = make_map_3d_ex(&map_f, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, feats[i], (unsigned long long)N,
                                        (unsigned long long)C, (unsigned long long)B, kTokPos, 256,
                                        CU_TENSOR_MAP_SWIZZLE_64B, CU_TENSOR_MAP_L2_PROMOTION_NONE, "ccl features");
      if (rc_map != SB_OK) return rc_map;
      ccl_norm_tokens_kernel<<<dim3((unsigned)((N + kTokPos - 1) / kTokPos), B), kTokThreads, smem, s>>>(map_f, toks[i], C,
                                                                                                      (int)N, dbg_word);
      SB_LAUNCH_CHECK("ccl_norm_tokens_kernel");
    }
  } else {
    ccl_norm_kernel<<<gn, 256, 0, s>>>(feature_1, den1, C, (int)N);
    SB_LAUNCH_CHECK("ccl_norm_kernel");
    ccl_norm_kernel<<<gn, 256, 0, s>>>(feature_2, den2, C, (int)N);
    SB_LAUNCH_CHECK("ccl_norm_kernel");
    ccl_tokens_kernel<<<gt, 256, 0, s>>>(feature_1, den1, tok1, C, (int)N);
    SB_LAUNCH_CHECK("ccl_tokens_kernel");
    ccl_tokens_kernel<<<gt, 256, 0, s>>>(feature_2, den2, tok2, C, (int)N);
    SB_LAUNCH_CHECK("ccl_tokens_kernel");
  }
  const int rc = sb_gemm_nt_tf32(tok1, tok2, c0, B, (int)N, (int)N, C, stream);
  if (rc) return rc;
  if ((N & 3) == 0 && W <= 256) {
    const size_t smem = ((size_t)3 * kFlowBandRows * ccl_stage_cols((int)N, W) + (size_t)kFlowG * kFlowQC + 3 * kFlowQC + kFlowG * kFlowWarpsPerP * 4) * sizeof(float) + 16;
    static SmemOptIn staged_opt_in;
    int opt_dev;
    if (staged_opt_in.need(200 * 1024, &opt_dev)) {
      SB_CUDA(cudaFuncSetAttribute(ccl_flow_staged_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
      staged_opt_in.done(200 * 1024, opt_dev);
    }
    ccl_flow_staged_kernel<<<dim3((unsigned)((N + kFlowG - 1) / kFlowG), B), kFlowG * kFlowWarpsPerP * 32, smem, s>>>(c0, flow, H, W, softmax_scale, dbg_word);
    SB_LAUNCH_CHECK("ccl_flow_staged_kernel");
    return SB_OK;
  }
  SB_REQUIRE(N <= 4096, SB_EUNSUP, "sb_ccl: more than 4096 positions with H*W %% 4 != 0 or W > 256 (8 rows of N floats must fit shared memory)");
  const size_t flow_smem = (size_t)8 * N * sizeof(float);
  static SmemOptIn flow_opt_in;
  int flow_dev;
  if (flow_smem > 48 * 1024 && flow_opt_in.need(flow_smem, &flow_dev)) {
    SB_CUDA(cudaFuncSetAttribute(ccl_flow_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)flow_smem));
    flow_opt_in.done(flow_smem, flow_dev);
  }
  ccl_flow_kernel<<<dim3((unsigned)((N + 7) / 8), B), 256, flow_smem, s>>>(c0, flow, H, W, softmax_scale);
  SB_LAUNCH_CHECK("ccl_flow_kernel");
  return SB_OK;
}
