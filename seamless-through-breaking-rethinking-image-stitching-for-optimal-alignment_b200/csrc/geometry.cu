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
  // A h = rhs, rows alternate (x y 1 0 0 0 -x*x' -y*x') / (0 0 0 x y 1 -x*y' -y*y')  (torch_DLT.py:8-15)
  double A[8][9];
  for (int k = 0; k < 4; ++k) {
    const double x = src_p[(b * 4 + k) * 2], y = src_p[(b * 4 + k) * 2 + 1];
    const double u = dst_p[(b * 4 + k) * 2], v = dst_p[(b * 4 + k) * 2 + 1];
    double* r0 = A[2 * k];
    double* r1 = A[2 * k + 1];
    r0[0] = x; r0[1] = y; r0[2] = 1; r0[3] = 0; r0[4] = 0; r0[5] = 0; r0[6] = -u * x; r0[7] = -u * y; r0[8] = u;
    r1[0] = 0; r1[1] = 0; r1[2] = 0; r1[3] = x; r1[4] = y; r1[5] = 1; r1[6] = -v * x; r1[7] = -v * y; r1[8] = v;
  }
  for (int c = 0; c < 8; ++c) {            // Gaussian elimination, partial pivoting
    int piv = c;
    double best = fabs(A[c][c]);
    for (int r = c + 1; r < 8; ++r)
      if (fabs(A[r][c]) > best) { best = fabs(A[r][c]); piv = r; }
    if (piv != c)
      for (int k = c; k < 9; ++k) { const double t = A[c][k]; A[c][k] = A[piv][k]; A[piv][k] = t; }
    const double inv = 1.0 / A[c][c];
    for (int r = c + 1; r < 8; ++r) {
      const double f = A[r][c] * inv;
      for (int k = c; k < 9; ++k) A[r][k] -= f * A[c][k];
    }
  }
  double h[9];
  for (int r = 7; r >= 0; --r) {
    double acc = A[r][8];
    for (int k = r + 1; k < 8; ++k) acc -= A[r][k] * h[k];
    h[r] = acc / A[r][r];
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
