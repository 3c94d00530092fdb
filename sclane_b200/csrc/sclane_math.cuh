// sclane_math.cuh -- scalar fp64 helpers shared by every kernel of the path.
// R semantics (round half-even on the decimal grid), special-function differences for the
// negative-binomial likelihood, and tiny dense algebra (n <= 8) kept in registers/local memory.
#pragma once
#include <math.h>
#include <stdint.h>

#ifdef __CUDACC__
#define SC_HD __host__ __device__ __forceinline__
#define SC_HDN __host__ __device__ __noinline__
#else
#define SC_HD inline
#define SC_HDN inline
#endif

namespace sclane {

// ---- R's round(x, digits) for digits > 0 (nmath/fround.c, R >= 4.0.0): nearest of the two
// decimal neighbours, ties to the even scaled integer.  Used for round(score, 6) / round(score2, 4)
// (marge2.R:374-576, :681) and round(GCVq1, 10) (stat_out.R:43).
SC_HD double r_round(double x, double p10) {
  if (!(x == x) || isinf(x) || x == 0.0) return x;
  double s = x < 0.0 ? -1.0 : 1.0;
  double v = fabs(x);
  double x10 = p10 * v;
  if (x10 >= 4503599627370496.0) return x;  // 2^52: already an integer multiple
  double i10 = floor(x10);
  {
    // fast path: the scaled value is clearly on one side of the half-way point (the product p10 * v carries a relative error
    // of 2^-53, far below the 1e-3 margin), so the nearer of the two decimal neighbours is known without comparing distances
    const double fr = x10 - i10;
    if (fr < 0.499 && x10 < 4.0e12) return s * (i10 / p10);
    if (fr > 0.501 && x10 < 4.0e12) return s * ((i10 + 1.0) / p10);
  }
  double xd = i10 / p10, xu = ceil(x10) / p10;
  double du = xu - v, dd = v - xd;
  bool even = fmod(i10, 2.0) == 0.0;
  return s * ((dd < du || (dd == du && even)) ? xd : xu);
}

// ---- digamma / trigamma for x > 0: upward recurrence to x >= 10, then the asymptotic series.
SC_HD double digamma_pos(double x) {
  double r = 0.0;
  while (x < 10.0) { r -= 1.0 / x; x += 1.0; }
  double f = 1.0 / (x * x);
  double t = f * (-1.0 / 12.0 + f * (1.0 / 120.0 + f * (-1.0 / 252.0 + f * (1.0 / 240.0 +
             f * (-1.0 / 132.0 + f * (691.0 / 32760.0 + f * (-1.0 / 12.0)))))));
  return r + log(x) - 0.5 / x + t;
}
SC_HD double trigamma_pos(double x) {
  double r = 0.0;
  while (x < 10.0) { r += 1.0 / (x * x); x += 1.0; }
  double f = 1.0 / (x * x);
  double t = (1.0 / 6.0 + f * (-1.0 / 30.0 + f * (1.0 / 42.0 + f * (-1.0 / 30.0 + f * (5.0 / 66.0 +
             f * (-691.0 / 2730.0 + f * (7.0 / 6.0))))))) * f / x;
  return r + 1.0 / x + 0.5 * f + t;
}

// psi(th + y) - psi(th) = sum_{k<y} 1/(th+k) and trigamma(th) - trigamma(th + y) = sum_{k<y} 1/(th+k)^2 for an
// integer count y >= 0.  Small counts (the bulk of single-cell data) use the exact finite sums, accumulated as
// one rational each (numerator / denominator) so that a cell costs one division per sum instead of y.
constexpr int kSmallCount = 12;
// large counts (rare): out of line so that the streaming loops stay small in the instruction cache
SC_HDN void psi_tri_diff_big(int y, double th, double& psi, double& tri) {
  const double yd = (double)y;
  if (th >= 1e4) {
    // large theta: the differences of the asymptotic series, term by term (no cancellation); the omitted
    // terms are below 1e-17 relative for th >= 1e4
    const double x1 = th, x2 = th + yd, pr = x1 * x2;
    psi = log1p(yd / th) + yd / (2.0 * pr) + yd * (x1 + x2) / (12.0 * pr * pr);
    tri = yd / pr + 0.5 * yd * (x1 + x2) / (pr * pr) + yd * (x1 * x1 + pr + x2 * x2) / (6.0 * pr * pr * pr);
    return;
  }
  psi = digamma_pos(th + yd) - digamma_pos(th);
  tri = trigamma_pos(th) - trigamma_pos(th + yd);
}
SC_HDN void psi_tri_diff_sums(int y, double th, double& psi, double& tri) {   // products would overflow: plain sums
  double s1 = 0.0, s2 = 0.0;
#pragma unroll 1
  for (int k = 0; k < y; ++k) { const double d = th + (double)k; s1 += 1.0 / d; s2 += 1.0 / (d * d); }
  psi = s1; tri = s2;
}
SC_HD void psi_tri_diff(int y, double th, double& psi, double& tri) {
  if (y == 0) { psi = 0.0; tri = 0.0; return; }
  if (y < kSmallCount) {
    if (th < 1e10) {
      double n1 = 1.0, d1 = th, n2 = 1.0, d2 = th * th;      // k = 0
#pragma unroll 1
      for (int k = 1; k < y; ++k) {
        const double x = th + (double)k, x2 = x * x;
        n1 = n1 * x + d1; d1 *= x;
        n2 = n2 * x2 + d2; d2 *= x2;
      }
      psi = n1 / d1;
      tri = n2 / d2;
    } else {
      psi_tri_diff_sums(y, th, psi, tri);
    }
    return;
  }
  psi_tri_diff_big(y, th, psi, tri);
}

// Stirling remainder of lgamma: lgamma(x) = (x - 1/2) log x - x + log(2 pi)/2 + stirling_corr(x), x >= 20.
SC_HD double stirling_corr(double x) {
  double f = 1.0 / (x * x);
  return (1.0 / 12.0 + f * (-1.0 / 360.0 + f * (1.0 / 1260.0 + f * (-1.0 / 1680.0 + f * (1.0 / 1188.0))))) / x;
}
// Lambda(y, sigma) = lgamma(th + y) - lgamma(th) + y log(sigma) = sum_{k=1}^{y-1} log1p(k sigma), th = 1/sigma.
// The mu-free part of the NB log-likelihood; cancellation-free for any th.  Small counts: one log of the product.
SC_HDN double nb_lambda_big(int y, double sigma, double th) {
  double yd = (double)y;
  if (th >= 20.0) {
    return (th + yd - 0.5) * log1p(yd * sigma) - yd + stirling_corr(th + yd) - stirling_corr(th);
  }
  return lgamma(th + yd) - lgamma(th) + yd * log(sigma);
}
SC_HD double nb_lambda(int y, double sigma, double th) {
  if (y <= 1) return 0.0;
  if (y < kSmallCount && sigma < 1e20) {
    double prod = 1.0 + sigma;
#pragma unroll 1
    for (int k = 2; k < y; ++k) prod *= 1.0 + (double)k * sigma;
    return log(prod);
  }
  return nb_lambda_big(y, sigma, th);
}

// lower triangle of the leading n x n block (all that chol_factor / chol_solve / chol_inverse read): n (n + 1) / 2 moves instead of NS^2
template <int NS>
SC_HD void copy_lower(int n, const double* src, double* dst) {
#pragma unroll 1
  for (int i = 0; i < n; ++i) {
#pragma unroll 1
    for (int j = 0; j <= i; ++j) dst[i * NS + j] = src[i * NS + j];
  }
}

// ---- tiny dense symmetric algebra, n <= NS, row-major a[NS*NS].
// Cholesky A = L L^T in place (lower); returns false on a non-positive pivot.
template <int NS>
SC_HDN bool chol_factor(int n, double* a) {
#pragma unroll 1
  for (int k = 0; k < n; ++k) {
    double x = a[k * NS + k];
#pragma unroll 1
    for (int j = 0; j < k; ++j) x -= a[k * NS + j] * a[k * NS + j];
    if (!(x > 0.0)) return false;
    x = sqrt(x);
    a[k * NS + k] = x;
#pragma unroll 1
    for (int i = k + 1; i < n; ++i) {
      double s = a[i * NS + k];
#pragma unroll 1
      for (int j = 0; j < k; ++j) s -= a[i * NS + j] * a[k * NS + j];
      a[i * NS + k] = s / x;
    }
  }
  return true;
}
template <int NS>
SC_HDN void chol_solve(int n, const double* l, const double* b, double* x) {
#pragma unroll 1
  for (int i = 0; i < n; ++i) {
    double s = b[i];
#pragma unroll 1
    for (int j = 0; j < i; ++j) s -= l[i * NS + j] * x[j];
    x[i] = s / l[i * NS + i];
  }
#pragma unroll 1
  for (int i = n - 1; i >= 0; --i) {
    double s = x[i];
#pragma unroll 1
    for (int j = i + 1; j < n; ++j) s -= l[j * NS + i] * x[j];
    x[i] = s / l[i * NS + i];
  }
}
// inverse from the Cholesky factor: inv (row-major, full symmetric)
template <int NS>
SC_HDN void chol_inverse(int n, const double* l, double* inv) {
  double e[NS], col[NS];
#pragma unroll 1
  for (int c = 0; c < n; ++c) {
#pragma unroll 1
    for (int i = 0; i < n; ++i) e[i] = (i == c) ? 1.0 : 0.0;
    chol_solve<NS>(n, l, e, col);
#pragma unroll 1
    for (int i = 0; i < n; ++i) inv[i * NS + c] = col[i];
  }
}

// Eigen's unblocked LLT as the reference uses it (src/eigenMapInvert.cpp:10, A.llt().solve(I)) for
// c x c Schur complements, c <= 2: the factorisation stops at the first non-positive pivot and leaves
// the remainder untouched; info() is never checked, so solve() runs on whatever is there.
// Returns b' S^-1 b for c = 1 or 2 (s is the lower triangle: s00, s10, s11).
SC_HD double llt_quadform(int c, double s00, double s10, double s11, double b0, double b1) {
  if (c == 1) {
    double l00 = (s00 <= 0.0) ? s00 : sqrt(s00);   // NaN goes through sqrt and stays NaN
    double inv = 1.0 / (l00 * l00);
    return b0 * inv * b0;
  }
  double l00, l10, l11;
  if (s00 <= 0.0) { l00 = s00; l10 = s10; l11 = s11; }
  else {
    l00 = sqrt(s00);
    l10 = s10 / l00;
    double x = s11 - l10 * l10;
    l11 = (x <= 0.0) ? s11 : sqrt(x);
  }
  // solve L y = b ; L^T x = y ; score = b.x
  double y0 = b0 / l00;
  double y1 = (b1 - l10 * y0) / l11;
  double x1 = y1 / l11;
  double x0 = (y0 - l10 * x1) / l00;
  return b0 * x0 + b1 * x1;
}

}  // namespace sclane
