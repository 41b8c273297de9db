/*
 * oracle.c — CPU restatement of the reference's hot-path arithmetic.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing under oracle/ is imported, linked or
 * executed by the product (package stitch_b200); only tests/, __graft_entry__.smoke()
 * and bench.py's cpu_baseline / --impl reference legs use it, as the checker
 * and as the timed CPU baseline.
 *
 * The reference (gargatik/Seamless-Through-Breaking-...) is pure Python over
 * ATen; this file restates, in plain C with one fp32 rounding per reference
 * op (compiled with -ffp-contract=off), what those ops compute ON THE CPU —
 * including the vectorised-CPU grid_sample arithmetic of ATen that
 * F.grid_sample resolves to.  It is pinned against golden vectors produced by
 * importing and running the reference's own functions (tests/golden/).
 * Every function cites the reference file:line it follows (paths relative to
 * the reference root).
 */
#include <limits.h>
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

typedef long long i64;

/* x86 cvttss2si: out-of-range / NaN -> INT_MIN ("integer indefinite"), which is
 * what torch's .int() / convert_to_int produce on the CPU. */
static inline int f2i_x86(float f) {
  if (!(f >= -2147483648.0f && f < 2147483648.0f)) return INT_MIN;
  return (int)f;
}

/* ---------------------------------------------------------------------------
 * grid_sample(bilinear, zeros, align_corners=True) as used by
 *   warp            core/warp_utils.py:71-80
 *   bilinear_sampler core/utils/utils.py:62-76
 * Python side:  g = 2*v / max(size-1,1) - 1
 * ATen (aten/src/ATen/native/cpu/GridSamplerKernel.cpp, ComputeLocation<align_corners=true>):
 *   x = (g + 1) * ((size-1)/2);  x_w = floor(x); w = x - x_w; e = 1 - w; ...
 *   nw = s*e, ne = s*w, sw = n*e, se = n*w;
 *   out = fma(se_v, se, fma(sw_v, sw, fma(ne_v, ne, nw_v*nw)))   (as compiled: FMA-contracted)
 * ------------------------------------------------------------------------- */
static inline float grid_roundtrip(float v, float den, float half) {
  float g = (2.0f * v) / den - 1.0f;
  return (g + 1.0f) * half;
}

typedef struct {
  float nw, ne, sw, se;
  int xw, yn;
  int m_nw, m_ne, m_sw, m_se;
} gtap_t;

static inline void gtap_setup(gtap_t* t, float ix, float iy, int H, int W) {
  float x_w = floorf(ix), y_n = floorf(iy);
  float w = ix - x_w, e = 1.0f - w, n = iy - y_n, s = 1.0f - n;
  t->nw = s * e; t->ne = s * w; t->sw = n * e; t->se = n * w;
  int xi = f2i_x86(x_w), yi = f2i_x86(y_n);
  /* masks: (i > -1) & (i < size), east = xi + 1 (wraps like the SIMD add) */
  int xe = (int)((unsigned)xi + 1u), ys = (int)((unsigned)yi + 1u);
  int mw = xi > -1 && xi < W, me = xe > -1 && xe < W;
  int mn = yi > -1 && yi < H, ms = ys > -1 && ys < H;
  t->m_nw = mn && mw; t->m_ne = mn && me; t->m_sw = ms && mw; t->m_se = ms && me;
  t->xw = xi; t->yn = yi;
}

static inline float gtap_sample(const gtap_t* t, const float* plane, int W) {
  float v_nw = t->m_nw ? plane[(i64)t->yn * W + t->xw] : 0.0f;
  float v_ne = t->m_ne ? plane[(i64)t->yn * W + t->xw + 1] : 0.0f;
  float v_sw = t->m_sw ? plane[(i64)(t->yn + 1) * W + t->xw] : 0.0f;
  float v_se = t->m_se ? plane[(i64)(t->yn + 1) * W + t->xw + 1] : 0.0f;
  /* ATen's Vectorized mul/add chain is contracted to FMAs by its compiler:
   * verified bit-for-bit against F.grid_sample on the CPU (tests/golden). */
  return fmaf(v_se, t->se, fmaf(v_sw, t->sw, fmaf(v_ne, t->ne, v_nw * t->nw)));
}

/* W1 — warp(x, flo): core/warp_utils.py:54-80.  mul_mask / overlap restate the
 * caller's follow-up ops (core/flowHomoAdpater.py:171-174, :182, :317). */
void o_flow_warp(const float* x, const float* flo, const float* mul_mask, float* out,
                 float* overlap, int B, int C, int H, int W) {
  const i64 plane = (i64)H * W;
  const float denx = (float)(W - 1 > 1 ? W - 1 : 1), deny = (float)(H - 1 > 1 ? H - 1 : 1);
  const float halfx = (float)(W - 1) / 2.0f, halfy = (float)(H - 1) / 2.0f;
#pragma omp parallel for collapse(2) schedule(static)
  for (int b = 0; b < B; ++b)
    for (int py = 0; py < H; ++py)
      for (int px = 0; px < W; ++px) {
        const i64 rem = (i64)py * W + px;
        const float fx = flo[((i64)b * 2) * plane + rem], fy = flo[((i64)b * 2 + 1) * plane + rem];
        gtap_t t;
        gtap_setup(&t, grid_roundtrip((float)px + fx, denx, halfx),
                   grid_roundtrip((float)py + fy, deny, halfy), H, W);
        float v[64];
        for (int c = 0; c < C; ++c) {
          float s = gtap_sample(&t, x + ((i64)b * C + c) * plane, W);
          if (c < 64) v[c] = s;
          out[((i64)b * C + c) * plane + rem] = mul_mask ? s * mul_mask[(i64)b * plane + rem] : s;
        }
        if (overlap && C == 6) {
          float mean = ((v[3] + v[4]) + v[5]) / 3.0f;
          overlap[(i64)b * plane + rem] = mean < 0.9f ? 1.0f : 0.0f;
        }
      }
}

/* W1, mode='nearest' — warp(x, flo, mode='nearest'): core/warp_utils.py:74-79.  The grid is normalised exactly as
 * for the bilinear mode (2 v / max(size-1, 1) - 1), but the reference then calls F.grid_sample(mode='nearest') WITHOUT
 * align_corners (:79), i.e. align_corners=False: ATen's CPU kernel un-normalises as fma(g + 1, size / 2, -0.5) (one
 * rounding), rounds half to even (nearbyint) and copies that pixel if it lies inside the image, else 0. */
void o_flow_warp_nearest(const float* x, const float* flo, float* out, int B, int C, int H, int W) {
  const i64 plane = (i64)H * W;
  const float denx = (float)(W - 1 > 1 ? W - 1 : 1), deny = (float)(H - 1 > 1 ? H - 1 : 1);
  const float hw = (float)W / 2.0f, hh = (float)H / 2.0f;
#pragma omp parallel for collapse(2) schedule(static)
  for (int b = 0; b < B; ++b)
    for (int py = 0; py < H; ++py)
      for (int px = 0; px < W; ++px) {
        const i64 rem = (i64)py * W + px;
        const float fx = flo[((i64)b * 2) * plane + rem], fy = flo[((i64)b * 2 + 1) * plane + rem];
        const float gx = 2.0f * ((float)px + fx) / denx - 1.0f, gy = 2.0f * ((float)py + fy) / deny - 1.0f;
        const float ix = nearbyintf(fmaf(gx + 1.0f, hw, -0.5f)), iy = nearbyintf(fmaf(gy + 1.0f, hh, -0.5f));
        const int inside = ix >= 0.0f && ix <= (float)(W - 1) && iy >= 0.0f && iy <= (float)(H - 1);   /* false for NaN */
        const i64 off = inside ? (i64)iy * W + (i64)ix : 0;
        for (int c = 0; c < C; ++c)
          out[((i64)b * C + c) * plane + rem] = inside ? x[((i64)b * C + c) * plane + off] : 0.0f;
      }
}

/* bilinear_sampler(img, coords): core/utils/utils.py:62-76 */
void o_bilinear_sampler(const float* img, const float* coords, float* out, int N, int C, int H,
                        int W, int Ho, int Wo) {
  const i64 plane = (i64)H * W, HoWo = (i64)Ho * Wo;
  const float denx = (float)(W - 1), deny = (float)(H - 1);
  const float halfx = (float)(W - 1) / 2.0f, halfy = (float)(H - 1) / 2.0f;
#pragma omp parallel for schedule(static)
  for (i64 p = 0; p < (i64)N * HoWo; ++p) {
    const i64 n = p / HoWo, rem = p - n * HoWo;
    gtap_t t;
    gtap_setup(&t, grid_roundtrip(coords[p * 2], denx, halfx),
               grid_roundtrip(coords[p * 2 + 1], deny, halfy), H, W);
    for (int c = 0; c < C; ++c) out[(n * C + c) * HoWo + rem] = gtap_sample(&t, img + (n * C + c) * plane, W);
  }
}

/* C3 / C3p — MemoryDecoder.encode_flow_token: core/FlowFormer/PerCostFormer3/decoder.py:242-260
 *   delta = stack(meshgrid(dy, dx), -1); coords = centroid + delta   (:250-256)
 *   -> tap k = i*(2r+1)+j samples (cx + dy[i], cy + dx[j])
 * pyramid level: centroid / 2**l + delta  (core/FlowFormer/common.py:245-248). */
void o_corr_lookup(const float* cost_maps, const float* coords, float* out, int B, int H1, int W1,
                   int H2, int W2, int r, float coord_scale, int out_stride, int out_offset) {
  const i64 HW1 = (i64)H1 * W1, nq = (i64)B * HW1, map = (i64)H2 * W2;
  const int side = 2 * r + 1;
  const float denx = (float)(W2 - 1), deny = (float)(H2 - 1);
  const float halfx = (float)(W2 - 1) / 2.0f, halfy = (float)(H2 - 1) / 2.0f;
#pragma omp parallel for schedule(static)
  for (i64 q = 0; q < nq; ++q) {
    const i64 b = q / HW1, pos = q - b * HW1;
    const float cx = coords[(b * 2) * HW1 + pos] * coord_scale;
    const float cy = coords[(b * 2 + 1) * HW1 + pos] * coord_scale;
    for (int i = 0; i < side; ++i)
      for (int j = 0; j < side; ++j) {
        gtap_t t;
        gtap_setup(&t, grid_roundtrip(cx + (float)(i - r), denx, halfx),
                   grid_roundtrip(cy + (float)(j - r), deny, halfy), H2, W2);
        out[q * out_stride + out_offset + i * side + j] = gtap_sample(&t, cost_maps + q * map, W2);
      }
  }
}

/* ---------------------------------------------------------------------------
 * UDIS sampler: core/udis_utils/torch_homo_transform.py:17-92 (same code in
 * torch_tps_transform.py:18-94).
 * ------------------------------------------------------------------------- */
typedef struct {
  float wa, wb, wc, wd;
  int x0, x1, y0, y1;
} utap_t;

static inline int clampi(int v, int lo, int hi) { return v < lo ? lo : (v > hi ? hi : v); }

static inline void utap_setup(utap_t* t, float xn, float yn, int H, int W) {
  float x = (xn + 1.0f) * (float)W / 2.0f;      /* :29 */
  float y = (yn + 1.0f) * (float)H / 2.0f;      /* :30 */
  int xi = f2i_x86(floorf(x)), yi = f2i_x86(floorf(y));   /* :33-36 */
  t->x0 = clampi(xi, 0, W - 1);
  t->x1 = clampi((int)((unsigned)xi + 1u), 0, W - 1);
  t->y0 = clampi(yi, 0, H - 1);
  t->y1 = clampi((int)((unsigned)yi + 1u), 0, H - 1);     /* :38-41 */
  float x0f = (float)t->x0, x1f = (float)t->x1, y0f = (float)t->y0, y1f = (float)t->y1;
  t->wa = (x1f - x) * (y1f - y);                /* :86-89 */
  t->wb = (x1f - x) * (y - y0f);
  t->wc = (x - x0f) * (y1f - y);
  t->wd = (x - x0f) * (y - y0f);
}

static inline float utap_sample(const utap_t* t, const float* plane, int W) {
  float Ia = plane[(i64)t->y0 * W + t->x0], Ib = plane[(i64)t->y1 * W + t->x0];
  float Ic = plane[(i64)t->y0 * W + t->x1], Id = plane[(i64)t->y1 * W + t->x1];
  return ((t->wa * Ia + t->wb * Ib) + t->wc * Ic) + t->wd * Id;   /* :90 */
}

/* 3-term dot product as the CPU BLAS evaluates torch.matmul(theta, grid) for
 * K = 3 (:126): acc = a0*b0; acc = fma(a1,b1,acc); acc = fma(a2,b2,acc).
 * Verified bit-for-bit against torch on this image (tests/golden). */
static inline float dot3(const float* t, float gx, float gy) {
  float acc = t[0] * gx;
  acc = fmaf(t[1], gy, acc);
  acc = fmaf(t[2], 1.0f, acc);
  return acc;
}

/* W2 — transformer(U, theta, out_size): core/udis_utils/torch_homo_transform.py:5-151 */
void o_homo_warp(const float* U, const float* theta, const float* xs, const float* ys, float* out,
                 int32_t* idx, int B, int C, int H, int W, int Hout, int Wout, int theta_batch) {
  const i64 oplane = (i64)Hout * Wout, iplane = (i64)H * W;
#pragma omp parallel for collapse(2) schedule(static)
  for (int b = 0; b < B; ++b)
    for (int r = 0; r < Hout; ++r) {
      const float* th = theta + (theta_batch > 1 ? (i64)b * 9 : 0);
      for (int c = 0; c < Wout; ++c) {
        const float gx = xs[c], gy = ys[r];
        float X = dot3(th, gx, gy), Y = dot3(th + 3, gx, gy), T = dot3(th + 6, gx, gy);
        float ge = fabsf(T) >= 1e-7f ? 1.0f : 0.0f;      /* :133-137 */
        T = T + 1e-6f * (1.0f - ge);
        utap_t t;
        utap_setup(&t, X / T, Y / T, H, W);
        const i64 rem = (i64)r * Wout + c;
        if (idx) {
          int32_t* d = idx + (i64)b * 4 * oplane + rem;
          d[0] = t.x0; d[oplane] = t.x1; d[2 * oplane] = t.y0; d[3 * oplane] = t.y1;
        }
        for (int ch = 0; ch < C; ++ch)
          out[((i64)b * C + ch) * oplane + rem] = utap_sample(&t, U + ((i64)b * C + ch) * iplane, W);
      }
    }
}

/* W3 — dense part of transformer(U, source, target, out_size):
 * core/udis_utils/torch_tps_transform.py:96-147.  T [B,2,pn+3] comes from the
 * fp64 solve (:149-185, done with numpy in stitch_oracle.py).  The reference
 * sums the pn+3 fp32 products with BLAS in an unspecified order; the oracle
 * accumulates them in fp64 (the value every order approximates). */
void o_tps_warp(const float* U, const float* T, const float* source, const float* xs,
                const float* ys, float* out, int32_t* idx, float* coords_dbg, int B, int C, int H,
                int W, int Hout, int Wout, int pn) {
  const i64 oplane = (i64)Hout * Wout, iplane = (i64)H * W;
#pragma omp parallel for collapse(2) schedule(static)
  for (int b = 0; b < B; ++b)
    for (int r = 0; r < Hout; ++r) {
      const float* tx = T + ((i64)b * 2) * (pn + 3);
      const float* ty = T + ((i64)b * 2 + 1) * (pn + 3);
      const float* sp = source + (i64)b * pn * 2;
      for (int c = 0; c < Wout; ++c) {
        const float gx = xs[c], gy = ys[r];
        double ax = (double)tx[0] + (double)tx[1] * gx + (double)tx[2] * gy;
        double ay = (double)ty[0] + (double)ty[1] * gx + (double)ty[2] * gy;
        for (int k = 0; k < pn; ++k) {
          float dx = gx - sp[2 * k], dy = gy - sp[2 * k + 1];
          float d2 = dx * dx + dy * dy;                  /* :115 */
          float rk = d2 * logf(d2 + 1e-6f);              /* :116 */
          ax += (double)tx[3 + k] * (double)rk;
          ay += (double)ty[3 + k] * (double)rk;
        }
        utap_t t;
        utap_setup(&t, (float)ax, (float)ay, H, W);
        const i64 rem = (i64)r * Wout + c;
        if (coords_dbg) {
          coords_dbg[((i64)b * 2) * oplane + rem] = (float)ax;
          coords_dbg[((i64)b * 2 + 1) * oplane + rem] = (float)ay;
        }
        if (idx) {
          int32_t* d = idx + (i64)b * 4 * oplane + rem;
          d[0] = t.x0; d[oplane] = t.x1; d[2 * oplane] = t.y0; d[3 * oplane] = t.y1;
        }
        for (int ch = 0; ch < C; ++ch)
          out[((i64)b * C + ch) * oplane + rem] = utap_sample(&t, U + ((i64)b * C + ch) * iplane, W);
      }
    }
}

/* W3k — warp_image_tps (core/inference/tps_methods/kornia_tps.py:105-176) = kornia's
 * create_meshgrid + warp_points_tps (kornia is absent from the image and unpinned by the
 * reference: restated from its published source) + F.grid_sample(bilinear, zeros, align_corners):
 *   d2_k = clamp(-2 p.c_k + |p|^2 + |c_k|^2, 0)   (_pair_square_euclidean :26-36)
 *   U_k  = 0.5 * d2_k * log(d2_k + 1e-8)          (_kernel_distance :38-45)
 *   q    = sum_k U_k w_k + (p.x a_1 + p.y a_2) + a_0
 * torch reduces the K products with a vectorised cascade sum in an unspecified order; the oracle
 * accumulates them in fp64.  ATen CPU un-normalisation (verified bit-for-bit against
 * F.grid_sample here): align_corners ? (g + 1) * ((size-1)/2) : fma(g + 1, size/2, -0.5). */
static inline float gs_unnormalize(float g, int size, int align_corners) {
  if (align_corners) return (g + 1.0f) * ((float)(size - 1) / 2.0f);
  return fmaf(g + 1.0f, (float)size / 2.0f, -0.5f);
}

void o_grid_sample(const float* img, const float* grid, float* out, int N, int C, int H, int W,
                   int Ho, int Wo, int align_corners) {
  const i64 plane = (i64)H * W, HoWo = (i64)Ho * Wo;
#pragma omp parallel for schedule(static)
  for (i64 p = 0; p < (i64)N * HoWo; ++p) {
    const i64 n = p / HoWo, rem = p - n * HoWo;
    gtap_t t;
    gtap_setup(&t, gs_unnormalize(grid[p * 2], W, align_corners),
               gs_unnormalize(grid[p * 2 + 1], H, align_corners), H, W);
    for (int c = 0; c < C; ++c) out[(n * C + c) * HoWo + rem] = gtap_sample(&t, img + (n * C + c) * plane, W);
  }
}

void o_tps_kornia_grid(const float* centers, const float* kweights, const float* affine,
                       const float* xs, const float* ys, float* grid, int B, int H, int W, int K) {
#pragma omp parallel for collapse(2) schedule(static)
  for (int b = 0; b < B; ++b)
    for (int r = 0; r < H; ++r) {
      const float* cp = centers + (i64)b * K * 2;
      const float* wp = kweights + (i64)b * K * 2;
      const float* A = affine + (i64)b * 6;
      for (int c = 0; c < W; ++c) {
        const float gx = xs[c], gy = ys[r];
        const float p2 = gx * gx + gy * gy;
        double ax = 0.0, ay = 0.0;
        for (int k = 0; k < K; ++k) {
          const float cx = cp[2 * k], cy = cp[2 * k + 1];
          const float c2 = cx * cx + cy * cy;
          const float dot = gx * cx + gy * cy;
          float d2 = (-2.0f * dot + p2) + c2;
          d2 = d2 > 0.0f ? d2 : 0.0f;
          const float u = (0.5f * d2) * logf(d2 + 1e-8f);
          ax += (double)wp[2 * k] * (double)u;
          ay += (double)wp[2 * k + 1] * (double)u;
        }
        float* g = grid + (((i64)b * H + r) * W + c) * 2;
        g[0] = ((float)ax + (gx * A[2] + gy * A[4])) + A[0];
        g[1] = ((float)ay + (gx * A[3] + gy * A[5])) + A[1];
      }
    }
}

/* N2 — MemoryDecoder.upsample_flow: core/FlowFormer/PerCostFormer3/decoder.py:214-225.
 * softmax over the 9 taps (ATen, non-last dim: max, exp(x - max), running sum, divide), then
 * sum_k p_k * (8 * flow)[neighbour k] in tap order (torch.sum over a 9-long dim). */
void o_upsample_flow(const float* flow, const float* mask, float* out, int N, int H, int W) {
  const i64 plane = (i64)H * W;
  const int Ho = 8 * H, Wo = 8 * W;
#pragma omp parallel for collapse(2) schedule(static)
  for (int n = 0; n < N; ++n)
    for (int h = 0; h < H; ++h)
      for (int w = 0; w < W; ++w) {
        float f[2][9];
        for (int c = 0; c < 2; ++c)
          for (int k = 0; k < 9; ++k) {
            const int hy = h + k / 3 - 1, wx = w + k % 3 - 1;
            const int ok = hy >= 0 && hy < H && wx >= 0 && wx < W;
            f[c][k] = ok ? 8.0f * flow[((i64)n * 2 + c) * plane + (i64)hy * W + wx] : 0.0f;
          }
        for (int dy = 0; dy < 8; ++dy)
          for (int dx = 0; dx < 8; ++dx) {
            float m[9], e[9], mx, sum = 0.0f, a0 = 0.0f, a1 = 0.0f;
            for (int k = 0; k < 9; ++k) m[k] = mask[((i64)n * 576 + k * 64 + dy * 8 + dx) * plane + (i64)h * W + w];
            mx = m[0];
            for (int k = 1; k < 9; ++k) mx = m[k] > mx ? m[k] : mx;
            for (int k = 0; k < 9; ++k) { e[k] = expf(m[k] - mx); sum += e[k]; }
            for (int k = 0; k < 9; ++k) {
              const float p = e[k] / sum;
              a0 += p * f[0][k];
              a1 += p * f[1][k];
            }
            out[(((i64)n * 2 + 0) * Ho + 8 * h + dy) * Wo + 8 * w + dx] = a0;
            out[(((i64)n * 2 + 1) * Ho + 8 * h + dy) * Wo + 8 * w + dx] = a1;
          }
      }
}

/* W4 — compute_range_map: core/warp_utils.py:114-175, and the 'wang' branch of
 * compute_occlusion (:185-221).  scatter_add_ on the CPU adds the weights in
 * list order: (di, dj) outer (:142-143), flattened [B,H,W] pixels inner.
 * mode 0 raw, 1 = 1-(1-clamp) (occlusion_are_zeros), 2 = 1-clamp, 3 = mode 1 >= 0.5 */
void o_range_map(const float* flow, float* out, int B, int H, int W, int mode) {
  const i64 plane = (i64)H * W, total = (i64)B * plane;
  memset(out, 0, (size_t)total * sizeof(float));
  for (int di = 0; di < 2; ++di)
    for (int dj = 0; dj < 2; ++dj)
      for (i64 p = 0; p < total; ++p) {
        const i64 b = p / plane, rem = p - b * plane;
        const int py = (int)(rem / W), px = (int)(rem - (i64)py * W);
        const float cx = (float)px + flow[(b * 2) * plane + rem];
        const float cy = (float)py + flow[(b * 2 + 1) * plane + rem];
        const float fx = floorf(cx), fy = floorf(cy);
        const float ox = cx - fx, oy = cy - fy;              /* :121 */
        const int ix = (int)((unsigned)f2i_x86(fx) + (unsigned)di);
        const int iy = (int)((unsigned)f2i_x86(fy) + (unsigned)dj);
        if (!(ix >= 0 && ix < W && iy >= 0 && iy < H)) continue;   /* :150-153 */
        const float wi = (1.0f - (float)di) - (di ? -1.0f : 1.0f) * ox;   /* :158 */
        const float wj = (1.0f - (float)dj) - (dj ? -1.0f : 1.0f) * oy;   /* :159 */
        out[b * plane + (i64)iy * W + ix] += wi * wj;                      /* :160,173 */
      }
  if (mode >= 1)
    for (i64 p = 0; p < total; ++p) {
      float c = out[p] < 0.0f ? 0.0f : (out[p] > 1.0f ? 1.0f : out[p]);
      float occ = 1.0f - c;                                  /* :213 */
      if (mode == 1) out[p] = 1.0f - occ;                    /* :219-220 */
      else if (mode == 2) out[p] = occ;
      else out[p] = (1.0f - occ) >= 0.5f ? 1.0f : 0.0f;      /* flowHomoAdpater.py:181 */
    }
}

/* W5 — preprocess_occlusion_mask: core/flowHomoAdpater.py:18-35 (border_is_zero = 1);
 * cv2.erode / cv2.dilate open of core/inference/tps_pipline.py:143-147 (border_is_zero = 0). */
void o_morph_open(const float* mask, float* out, int P, int H, int W, int kh, int kw,
                  int border_is_zero) {
  const int rh = kh / 2, rw = kw / 2;
  const i64 plane = (i64)H * W;
  unsigned char* bin = (unsigned char*)malloc((size_t)plane);
  unsigned char* ero = (unsigned char*)malloc((size_t)plane);
  for (int p = 0; p < P; ++p) {
    const float* m = mask + p * plane;
    for (i64 i = 0; i < plane; ++i) bin[i] = m[i] >= 0.5f;
#pragma omp parallel for schedule(static)
    for (int y = 0; y < H; ++y)
      for (int x = 0; x < W; ++x) {
        int all = 1;
        for (int dy = -rh; dy <= rh && all; ++dy)
          for (int dx = -rw; dx <= rw; ++dx) {
            int yy = y + dy, xx = x + dx;
            int v = (yy >= 0 && yy < H && xx >= 0 && xx < W) ? bin[(i64)yy * W + xx] : !border_is_zero;
            if (!v) { all = 0; break; }
          }
        ero[(i64)y * W + x] = (unsigned char)all;     /* conv == kernel.numel() (:27) */
      }
#pragma omp parallel for schedule(static)
    for (int y = 0; y < H; ++y)
      for (int x = 0; x < W; ++x) {
        int any = 0;
        for (int dy = -rh; dy <= rh && !any; ++dy)
          for (int dx = -rw; dx <= rw; ++dx) {
            int yy = y + dy, xx = x + dx;
            if (yy >= 0 && yy < H && xx >= 0 && xx < W && ero[(i64)yy * W + xx]) { any = 1; break; }
          }
        out[p * plane + (i64)y * W + x] = any ? 1.0f : 0.0f;   /* conv >= 1 (:30) */
      }
  }
  free(bin);
  free(ero);
}

static inline float clipf(float v, float lo, float hi) {
  if (v != v) return v;
  return v < lo ? lo : (v > hi ? hi : v);
}
static inline uint8_t to_u8(float v) {
  if (v != v) return 0;          /* cvttss2si(NaN) = INT_MIN, low byte 0 */
  return (uint8_t)(int)v;
}
static inline float mean3(float a, float b, float c) { return ((a + b) + c) / 3.0f; }

/* W6 — compositing of test_out_forward: core/flowHomoAdpater.py:337-360 */
void o_composite_test_out(const float* homo1, const float* homo2, const float* fw_in,
                          const float* occ, float* final_warp, float* output2, float* mask1,
                          float* mask2, uint8_t* blend, int B, int H, int W) {
  const i64 plane = (i64)H * W;
#pragma omp parallel for schedule(static)
  for (i64 p = 0; p < (i64)B * plane; ++p) {
    const i64 b = p / plane, rem = p - b * plane;
    float a1[6], a2[6], f[6];
    for (int c = 0; c < 6; ++c) {
      a1[c] = homo1[(b * 6 + c) * plane + rem];
      a2[c] = homo2[(b * 6 + c) * plane + rem];
      f[c] = fw_in[(b * 6 + c) * plane + rem];
      if (occ) f[c] = f[c] * occ[b * plane + rem];                       /* :337 */
      final_warp[(b * 6 + c) * plane + rem] = f[c];
    }
    float o2[3], m2n[3];
    for (int c = 0; c < 3; ++c) {
      const float m1c = a1[3 + c], m2c = f[3 + c];
      if (occ) {
        const float non_ov = 1.0f - m1c;                                  /* :341 */
        o2[c] = a2[c] * (1.0f - m2c) * non_ov + f[c] * m2c;               /* :343 */
        m2n[c] = a2[3 + c] * (1.0f - m2c) * non_ov + m2c * m2c;           /* :344 */
      } else {
        o2[c] = a2[c] * (1.0f - m2c) + f[c] * m2c;                        /* :348 */
        m2n[c] = a2[3 + c] * (1.0f - m2c) + m2c * m2c;                    /* :349 */
      }
    }
    const float mm1 = clipf(mean3(a1[3], a1[4], a1[5]), 0.0f, 1.0f);      /* :359 */
    const float mm2 = clipf(mean3(m2n[0], m2n[1], m2n[2]), 0.0f, 1.0f);   /* :360 */
    for (int c = 0; c < 3; ++c) {
      const i64 o3 = (b * 3 + c) * plane + rem;
      output2[o3] = o2[c];
      mask1[o3] = mm1;
      mask2[o3] = mm2;
      const float num = a1[c] * a1[3 + c] + o2[c] * m2n[c];               /* :355 */
      const float den = a1[3 + c] + m2n[c];
      blend[o3] = to_u8(clipf(num / den, 0.0f, 255.0f));                  /* :356 */
    }
  }
}

/* W7 — build_model arithmetic: core/UDIS2/Composition/network.py:12-14 */
void o_build_model(const float* w1, const float* w2, const float* m1, const float* m2,
                   const float* net_out, float* lm1, float* lm2, float* st, int B, int H, int W) {
  const i64 plane = (i64)H * W;
#pragma omp parallel for schedule(static)
  for (i64 i = 0; i < (i64)B * 3 * plane; ++i) {
    const i64 bc = i / plane, rem = i - bc * plane, b = bc / 3;
    const float a = m1[i], c = m2[i], o = net_out[b * plane + rem];
    const float mm = a * c;
    const float l1 = (a - mm) + mm * o;
    const float l2 = (c - mm) + mm * (1.0f - o);
    lm1[i] = l1;
    lm2[i] = l2;
    st[i] = ((w1[i] + 1.0f) * l1 + (w2[i] + 1.0f) * l2) - 1.0f;
  }
}

/* W8 — TPS-stage mix + blend: core/inference/tps_pipline.py:150-170 */
void o_tps_mix_blend(const float* final_warp, const float* tps_warp, const float* tps_mask,
                     const float* output1, const float* mask1, float* output2, float* mask2,
                     uint8_t* blend, int B, int H, int W) {
  const i64 plane = (i64)H * W;
#pragma omp parallel for schedule(static)
  for (i64 p = 0; p < (i64)B * plane; ++p) {
    const i64 b = p / plane, rem = p - b * plane;
    float fw[3], tp[3], o1[3], m1[3];
    for (int c = 0; c < 3; ++c) {
      const i64 o3 = (b * 3 + c) * plane + rem;
      fw[c] = final_warp[o3]; tp[c] = tps_warp[o3]; o1[c] = output1[o3]; m1[c] = mask1[o3];
    }
    const float tm = tps_mask[b * plane + rem];
    const float fm = mean3(fw[0] >= 3.0f, fw[1] >= 3.0f, fw[2] >= 3.0f) >= 0.5f ? 1.0f : 0.0f;  /* :151-152 */
    const float inv1 = mean3(1.0f - m1[0], 1.0f - m1[1], 1.0f - m1[2]) >= 0.5f ? 1.0f : 0.0f;   /* :154-155 */
    const float tfm = fm + (1.0f - fm) * tm * inv1;                                                /* :157 */
    mask2[b * plane + rem] = tfm;
    for (int c = 0; c < 3; ++c) {
      const i64 o3 = (b * 3 + c) * plane + rem;
      const float tfw = fw[c] * fm + tp[c] * (1.0f - fm) * inv1;                                   /* :156 */
      const float o2 = tfw * tfm;                                                                  /* :162 */
      output2[o3] = o2;
      blend[o3] = to_u8(clipf((o1[c] * m1[c] + o2 * tfm) / (m1[c] + tfm), 0.0f, 255.0f));          /* :168-169 */
    }
  }
}

/* flowHomoAdpater.py:171-174 */
void o_overlap_mask(const float* final_warp, float* overlap, int B, int H, int W) {
  const i64 plane = (i64)H * W;
  for (i64 p = 0; p < (i64)B * plane; ++p) {
    const i64 b = p / plane, rem = p - b * plane;
    const float* m = final_warp + (b * 6 + 3) * plane + rem;
    overlap[p] = mean3(m[0], m[plane], m[2 * plane]) < 0.9f ? 1.0f : 0.0f;
  }
}

/* C2 — F.avg_pool2d(x, 2, stride=2): sum in (kh, kw) order, divide by 4. */
void o_avg_pool2x2(const float* in, float* out, i64 planes, int H, int W) {
  const int Ho = H / 2, Wo = W / 2;
#pragma omp parallel for schedule(static)
  for (i64 pl = 0; pl < planes; ++pl)
    for (int y = 0; y < Ho; ++y)
      for (int x = 0; x < Wo; ++x) {
        const float* s = in + pl * H * W + (i64)(2 * y) * W + 2 * x;
        out[pl * Ho * Wo + (i64)y * Wo + x] = (((s[0] + s[1]) + s[W]) + s[W + 1]) * 0.25f;
      }
}

/* C1 — MemoryEncoder.corr: core/FlowFormer/PerCostFormer3/encoder.py:359-369 (heads folded
 * into B by the caller).  Plain fp32 triple loop (k innermost, ascending); used for
 * small cases, stitch_oracle.py uses numpy's BLAS matmul for the large ones. */
void o_corr(const float* f1, const float* f2, float* vol, int B, int C, int N1, int N2) {
#pragma omp parallel for collapse(2) schedule(static)
  for (int b = 0; b < B; ++b)
    for (int i = 0; i < N1; ++i)
      for (int j = 0; j < N2; ++j) {
        float acc = 0.0f;
        for (int d = 0; d < C; ++d)
          acc += f1[((i64)b * C + d) * N1 + i] * f2[((i64)b * C + d) * N2 + j];
        vol[((i64)b * N1 + i) * N2 + j] = acc;
      }
}
