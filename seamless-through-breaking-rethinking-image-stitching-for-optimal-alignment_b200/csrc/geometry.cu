// geometry.cu — G1 fused: 4-point DLT + normalised-coordinate conjugation.
// Replaces, for the adapter's hot loop, tensor_DLT (core/udis_utils/torch_DLT.py:17-45)
// followed by the M^-1 H M / M^-1 H^-1 M products of core/flowHomoAdpater.py:96-113:
//   H         = DLT(src_p, dst_p)                  8x8 solve per pair
//   theta     = L * H * R                          (L = M^-1, R = M in train_eval)
//   theta_inv = L * H^-1 * R
// The reference does this with ~40 tiny ATen launches and torch.inverse, whose
// singularity check synchronises the host every call; here it is ONE launch
// (one thread per pair), capturable in a CUDA graph.  The solve runs in fp64
// with partial pivoting and is rounded once to fp32 (the reference's fp32 LU
// agrees to ~1e-6 relative).
#include "common.cuh"

namespace sb {

struct Mat3 { float m[9]; };

__device__ inline void mul3(const double* a, const double* b, double* c) {
  for (int i = 0; i < 3; ++i)
    for (int j = 0; j < 3; ++j)
      c[i * 3 + j] = a[i * 3] * b[j] + a[i * 3 + 1] * b[3 + j] + a[i * 3 + 2] * b[6 + j];
}

__device__ inline bool inv3(const double* a, double* o) {
  const double c0 = a[4] * a[8] - a[5] * a[7], c1 = a[5] * a[6] - a[3] * a[8], c2 = a[3] * a[7] - a[4] * a[6];
  const double det = a[0] * c0 + a[1] * c1 + a[2] * c2;
  const double id = 1.0 / det;
  o[0] = c0 * id; o[1] = (a[2] * a[7] - a[1] * a[8]) * id; o[2] = (a[1] * a[5] - a[2] * a[4]) * id;
  o[3] = c1 * id; o[4] = (a[0] * a[8] - a[2] * a[6]) * id; o[5] = (a[2] * a[3] - a[0] * a[5]) * id;
  o[6] = c2 * id; o[7] = (a[1] * a[6] - a[0] * a[7]) * id; o[8] = (a[0] * a[4] - a[1] * a[3]) * id;
  return det != 0.0;
}

__global__ void dlt_theta_kernel(const float* __restrict__ src_p, const float* __restrict__ dst_p,
                                 const Mat3 L, const Mat3 R, float* __restrict__ H_out,
                                 float* __restrict__ theta, float* __restrict__ theta_inv, int B) {
  const int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= B) return;
  // Four correspondences in general position determine H up to scale, so the 8x8 DLT system
  // (torch_DLT.py:8-15, h33 = 1) has the closed-form solution H = Q(dst) * Q(src)^-1 / [.]_33,
  // where Q(p) is the projective map of the unit square onto the quad p (Heckbert 1989).
  // ~150 fp64 flops instead of a serial 8x8 elimination.
  double qs[9], qd[9];
  {
    const float* pts[2] = {src_p + b * 8, dst_p + b * 8};
    double* qq[2] = {qs, qd};
#pragma unroll
    for (int w = 0; w < 2; ++w) {
      // point order of the reference: p0 = (0,0), p1 = (1,0), p2 = (0,1), p3 = (1,1) of the square
      const double x0 = pts[w][0], y0 = pts[w][1], x1 = pts[w][2], y1 = pts[w][3];
      const double x3 = pts[w][4], y3 = pts[w][5], x2 = pts[w][6], y2 = pts[w][7];
      const double dx1 = x1 - x2, dx2 = x3 - x2, sx = x0 - x1 + x2 - x3;
      const double dy1 = y1 - y2, dy2 = y3 - y2, sy = y0 - y1 + y2 - y3;
      const double den = dx1 * dy2 - dx2 * dy1;
      const double g = (sx * dy2 - dx2 * sy) / den;
      const double hh = (dx1 * sy - sx * dy1) / den;
      double* q = qq[w];
      q[0] = x1 - x0 + g * x1; q[1] = x3 - x0 + hh * x3; q[2] = x0;
      q[3] = y1 - y0 + g * y1; q[4] = y3 - y0 + hh * y3; q[5] = y0;
      q[6] = g;                q[7] = hh;                q[8] = 1.0;
    }
  }
  double qsi[9], h[9];
  inv3(qs, qsi);
  mul3(qd, qsi, h);
  {
    const double n = 1.0 / h[8];
#pragma unroll
    for (int i = 0; i < 9; ++i) h[i] *= n;
  }
  h[8] = 1.0;
  // the reference rounds H to fp32 before using it further
  for (int i = 0; i < 9; ++i) h[i] = (double)(float)h[i];
  if (H_out)
    for (int i = 0; i < 9; ++i) H_out[b * 9 + i] = (float)h[i];
  double Ld[9], Rd[9], t1[9], t2[9];
  for (int i = 0; i < 9; ++i) { Ld[i] = L.m[i]; Rd[i] = R.m[i]; }
  if (theta) {
    mul3(Ld, h, t1); mul3(t1, Rd, t2);
    for (int i = 0; i < 9; ++i) theta[b * 9 + i] = (float)t2[i];
  }
  if (theta_inv) {
    double hi[9];
    inv3(h, hi);
    for (int i = 0; i < 9; ++i) hi[i] = (double)(float)hi[i];
    mul3(Ld, hi, t1); mul3(t1, Rd, t2);
    for (int i = 0; i < 9; ++i) theta_inv[b * 9 + i] = (float)t2[i];
  }
}

}  // namespace sb

extern "C" int sb_dlt_theta(const float* src_p, const float* dst_p, const float* L_host,
                            const float* R_host, float* H, float* theta, float* theta_inv, int B,
                            sb_stream_t stream) {
  using namespace sb;
  SB_ENTER();
  SB_REQUIRE(B >= 0, SB_EINVAL, "sb_dlt_theta: bad size");
  if (B == 0) return SB_OK;
  SB_REQUIRE(src_p && dst_p && L_host && R_host, SB_EINVAL, "sb_dlt_theta: null pointer");
  Mat3 L, R;
  for (int i = 0; i < 9; ++i) { L.m[i] = L_host[i]; R.m[i] = R_host[i]; }
  dlt_theta_kernel<<<(B + 63) / 64, 64, 0, as_stream(stream)>>>(src_p, dst_p, L, R, H, theta, theta_inv, B);
  SB_LAUNCH_CHECK("dlt_theta_kernel");
  return SB_OK;
}
