/* Exhaustive check of the division restatement used by corr_lookup_r4_kernel (fast_axis):
 *     r = RN(1/d);  q = RN(a*r);  q = fma(fma(-q, d, a), r, q)   ==   RN(a/d)
 * for every float32 mantissa of `a` in two adjacent binades (correct rounding of a quotient is
 * scale-invariant in the normal range, and two binades cover both normalisations of a/d) and
 * every integer denominator d in [d_lo, d_hi] (d = size-1 of a cost map axis).
 * Returns the number of mismatches.  Test infrastructure only.
 * build: gcc -O2 -mfma -ffp-contract=off -fopenmp -shared -fPIC */
#include <math.h>
#include <stdint.h>
#include <string.h>

static inline float as_float(uint32_t u) { float f; memcpy(&f, &u, 4); return f; }

long long div_trick_mismatches(int d_lo, int d_hi) {
  long long bad = 0;
#pragma omp parallel for schedule(dynamic, 1) reduction(+ : bad)
  for (int di = d_lo; di <= d_hi; ++di) {
    const float d = (float)di;
    const float r = 1.0f / d;                      /* IEEE: correctly rounded, like __frcp_rn */
    for (uint32_t e = 127; e <= 128; ++e) {
      for (uint32_t m = 0; m < (1u << 23); ++m) {
        const float a = as_float((e << 23) | m);
        float q = a * r;
        q = fmaf(fmaf(-q, d, a), r, q);
        if (q != a / d) ++bad;
      }
    }
  }
  return bad;
}
