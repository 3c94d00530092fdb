// sclane_fastmath.cuh -- table-driven fp64 exp / log1p for the per-point and per-cell loops of the NB fits.
//
// Why: ncu showed the fit kernels FP64-pipe / issue bound, and the CUDA math library's log1p alone is ~64 fp64
// instructions (+ ~100 integer / predicate ones) and exp ~16 (+ ~25): together 3/4 of a point evaluation.  Both
// arguments are benign here -- log1p is only ever taken of x = sigma * mu >= 0 next to t = 1 + x and 1/t, which the
// caller already has, and exp of a linear predictor -- so two small tables in shared memory replace them:
//
//   fast_log_ge1(t), t >= 1:  t = 2^e m, m in [1, 2); j = top 7 mantissa bits; r = m * rc_j - 1 with rc_j = fl(1/(1 + j/128))
//                             (|r| < 2^-7, one fma, exact to 2^-60); log t = e ln2 + lc_j + log1p(r), lc_j = -log(rc_j)
//                             tabulated to 0.5 ulp, log1p(r) by its degree-7 Taylor polynomial (remainder < 2^-59).
//                             j = 0 has rc = 1, lc = 0: full RELATIVE accuracy as t -> 1.
//   fast_log1p_pos(x, t, inv) = fast_log_ge1(t) + (x - (t - 1)) * inv     (the rounding error of t = fl(1 + x) put back:
//                             relative accuracy for tiny x, i.e. for the near-Poisson fits with theta ~ 1e6)
//   fast_exp(x):              k = rint(64 x / ln2), r = x - k ln2/64 (two-term Cody-Waite, |r| <= 0.0055),
//                             exp x = 2^(k >> 6) * 2^((k & 63)/64) * (1 + r + r^2/2 + ... + r^6/720)
//
// Errors (checked against 60-digit arithmetic in tests/test_host_logic.py through the host twin below): <= 1 ulp.
// 11 + 3 fp64 instructions for the log1p, 10 for the exp.  Host code (test harness) can call the same functions: the bit
// manipulation has a portable branch.
#pragma once
#include <string.h>

#include "sclane_math.cuh"
#include "sclane_fastmath_tables.inc"

namespace sclane {

struct FastTabs {
  double lg[128][2];     // {rc_j, lc_j}
  double ex[64];         // 2^(j/64)
};

#ifdef __CUDACC__
__device__ const double g_fast_lg[128][2] = SCLANE_LOGTAB_INIT;
__device__ const double g_fast_ex[64] = SCLANE_EXPTAB_INIT;
// cooperative copy of the tables into a CTA's shared memory (call before the first pass; the caller syncs)
__device__ __forceinline__ void fast_tabs_load(FastTabs* ft, int tid, int nthreads) {
  for (int i = tid; i < 256; i += nthreads) (&ft->lg[0][0])[i] = (&g_fast_lg[0][0])[i];
  for (int i = tid; i < 64; i += nthreads) ft->ex[i] = g_fast_ex[i];
}
#endif
inline void fast_tabs_host(FastTabs* ft) {
  static const double lg[128][2] = SCLANE_LOGTAB_INIT;
  static const double ex[64] = SCLANE_EXPTAB_INIT;
  memcpy(ft->lg, lg, sizeof(lg));
  memcpy(ft->ex, ex, sizeof(ex));
}

SC_HD int fm_hi(double x) {
#ifdef __CUDA_ARCH__
  return __double2hiint(x);
#else
  long long b; memcpy(&b, &x, 8); return (int)(b >> 32);
#endif
}
SC_HD int fm_lo(double x) {
#ifdef __CUDA_ARCH__
  return __double2loint(x);
#else
  long long b; memcpy(&b, &x, 8); return (int)(b & 0xffffffffLL);
#endif
}
SC_HD double fm_make(int hi, int lo) {
#ifdef __CUDA_ARCH__
  return __hiloint2double(hi, lo);
#else
  const long long b = ((long long)hi << 32) | (long long)(unsigned int)lo;
  double x; memcpy(&x, &b, 8); return x;
#endif
}

// log(t) for t >= 1.  t > 1e300, +inf and NaN are returned unchanged (they stay non-finite for the callers' checks).
SC_HD double fast_log_ge1(double t, const FastTabs& ft) {
  const int hi = fm_hi(t), lo = fm_lo(t);
  const int e = (hi >> 20) - 1023;
  const int j = (hi >> 13) & 127;
  const double m = fm_make((hi & 0x000fffff) | 0x3ff00000, lo);
  const double rc = ft.lg[j][0], lc = ft.lg[j][1];
  const double r = fma(m, rc, -1.0);
  double q = fma(r, 1.0 / 7.0, -1.0 / 6.0);
  q = fma(r, q, 0.2);
  q = fma(r, q, -0.25);
  q = fma(r, q, 1.0 / 3.0);
  q = fma(r, q, -0.5);
  const double p = fma(r * r, q, r);
  const double ed = fm_make(0x43300000, e) - 4503599627370496.0;      // (double)e for 0 <= e < 2^31, no conversion instruction
  const double res = fma(ed, SCLANE_LN2, lc + p);
  return t < 1e300 ? res : t;
}

// log1p(x) for x >= 0 given t = fl(1 + x) and inv = 1/t (to working precision)
SC_HD double fast_log1p_pos(double x, double t, double inv, const FastTabs& ft) {
  return fma(x - (t - 1.0), inv, fast_log_ge1(t, ft));
}

// Branch free (a rare-path call would split the loop body into basic blocks and stop the compiler from interleaving the
// independent cells / points of an unrolled pass loop): the argument is clamped to +-700 for the arithmetic and the IEEE
// results of the out-of-range cases are put back with selects: x > 700 -> +inf (exp overflows at 709.78; mu > 1e304 is a
// diverged fit either way), x < -700 -> 0, NaN -> NaN.
SC_HD double fast_exp(double x, const FastTabs& ft) {
  const double xc = fmin(fmax(x, -700.0), 700.0);
  double kd = fma(xc, SCLANE_64_LN2, 6755399441055744.0);               // 1.5 * 2^52: round to nearest integer in the low word
  const int k = fm_lo(kd);
  kd -= 6755399441055744.0;
  double r = fma(kd, -SCLANE_LN2_64_HI, xc);
  r = fma(kd, -SCLANE_LN2_64_LO, r);
  double q = fma(r, 1.0 / 720.0, 1.0 / 120.0);
  q = fma(r, q, 1.0 / 24.0);
  q = fma(r, q, 1.0 / 6.0);
  q = fma(r, q, 0.5);
  const double p = fma(r * r, q, r);
  const double s0 = ft.ex[k & 63];
  const double s = fm_make(fm_hi(s0) + ((k >> 6) << 20), fm_lo(s0));     // * 2^(k >> 6): exponent stays in the normal range for |x| <= 700
  double res = fma(s, p, s);
  res = x > 700.0 ? (double)INFINITY : res;
  res = x < -700.0 ? 0.0 : res;
  return x != x ? x : res;
}

// The math policy of the pass bodies: with tables (device kernels, host tests of the tables) the fast versions, without
// (ft == nullptr: the serial host harness) the C library.
#ifdef __CUDA_ARCH__
// device: always the tables (ft is never null in a kernel) -- the library versions are not even instantiated, which keeps
// the pass loops small in the instruction cache
SC_HD double sc_exp(double x, const FastTabs* ft) { return fast_exp(x, *ft); }
SC_HD double sc_log1p_pos(double x, double t, double inv, const FastTabs* ft) { return fast_log1p_pos(x, t, inv, *ft); }
#else
SC_HD double sc_exp(double x, const FastTabs* ft) { return ft ? fast_exp(x, *ft) : exp(x); }
SC_HD double sc_log1p_pos(double x, double t, double inv, const FastTabs* ft) { return ft ? fast_log1p_pos(x, t, inv, *ft) : log1p(x); }
#endif

}  // namespace sclane
