// sclane_comp.cuh -- the X-compressed form of the GLM passes (shared host/device).
//
// marge2() rounds pseudotime to 4 decimals (marge2.R:154), so a lineage with N cells has only D <= ~10^4 distinct
// abscissae when pseudotime lives in [0, 1] (D = 10,000 at N = 200,000).  Without an offset the fitted mean of a cell
// depends on its abscissa only, and every sum the NB fits and the score sweep need is either
//   * linear in y at fixed abscissa  -> needs per point d:  n_d (cells), s_d = sum y, q_d = sum y^2, or
//   * free of mu                      -> needs the gene's count histogram only:
//       sum_i [psi(th + y_i) - psi(th)]        = sum_k C_k / (th + k),       C_k = #{i : y_i > k}
//       sum_i [trigamma(th) - trigamma(th+y_i)] = sum_k C_k / (th + k)^2
//       sum_i [lgamma(th+y_i) - lgamma(th) + y_i log sigma] = sum_{k>=1} C_k log1p(k sigma)
//       sum_{y_i>0} (y_i + th) log1p(sigma y_i) = sum_{y>=1} hist[y] (y + th) log1p(sigma y)
// (glm.fit's mustart pass is the one place that is non-linear in y at fixed abscissa; its two sums are functions of y
// alone because init.theta = 1, so they are aggregated once per gene: a0_d, b0_d.)
// A pass therefore costs D point evaluations + ymax histogram terms instead of N cell evaluations, and a bucket in which
// the linear predictor is flat is one closed-form evaluation from its bucket totals.  The executor that implements
// pass<MODE>() this way produces the same W.acc / W.sc as the per-cell executor (to rounding), so the whole algorithm in
// sclane_core.cuh runs unchanged on top of it.  Used for every fit without offset (all forward / backward fits, and the
// final fits when size.factor.offset is NULL); fits with an offset stay on the per-cell path.
#pragma once
#include "sclane_core.cuh"

namespace sclane {

constexpr int kHistCap = 4096;          // counts below this go through the gene's histogram (tail counts)
constexpr int kBigCap = 8192;           // counts >= kHistCap are kept as a sorted per-gene list of at most kBigCap cells, whose
                                        // mu-free terms are evaluated directly; a gene with more such cells takes the per-cell path

// per-gene aggregates
struct CompGeneHdr {
  int ymax;                 // largest count
  int n_zero;               // cells with y == 0
  int route;                // 1 = compressed path usable (at most kBigCap cells with a count >= kHistCap)
  int n_big;                // cells with a count >= kHistCap (listed in ascending order)
  double SB[NBMAX], SBu[NBMAX], SBuu[NBMAX], QB[NBMAX];     // per bucket: sum y, sum y u, sum y u^2, sum y^2
};

// per-gene aggregates as the device holds them (k_aggregate)
struct CompGeneDev {
  CompGeneHdr hdr;
  double ST[NBMAX][5];          // glm.fit mustart pass per bucket: sum w, sum w u, sum w u^2, sum w z, sum w z u
};

// One point of the compressed layout: n cells, s = sum y, q = sum y^2, all with the same mean mu = exp(a + c u).
template <int MODE>
SC_HD void point_contrib(double n, double s, double q, double u, double a, double c, double sigma, double theta, bool want_ll,
                         const FastTabs* ft, double* acc, double* sc, double off = 0.0) {
  // off: offset of the point (only the per-cell hybrid passes of the fits with a size-factor offset pass one: n = 1, s = y)
  const double eta = a + c * u + off;
  const double mu = sc_exp(eta, ft);
  if (MODE == PASS_SWEEP) {
    const double inv = rcp_pos(1.0 + sigma * mu);
    const double w = n * mu * inv, r = (s - n * mu) * inv;
    acc[0] += w; acc[1] += w * u; acc[2] += w * u * u;
    acc[3] += r; acc[4] += r * u;
    return;
  }
  if (MODE == PASS_FIT) {
    if (sigma == 0.0) {
      const double m = n * mu, r = s - m;
      acc[0] += m; acc[1] += m * u; acc[2] += m * u * u;
      acc[3] += r; acc[4] += r * u;
      sc[2] += s * eta - m;
      return;
    }
    const double sm = sigma * mu, t1 = 1.0 + sm;
    const double inv = rcp_pos(t1);
    const double r = (s - n * mu) * inv;
    const double ns = n + sigma * s;
    const double o = mu * ns * inv * inv;
    const double cc = sm * r * inv;
    const double lg = sc_log1p_pos(sm, t1, inv, ft);
    acc[0] += o; acc[1] += o * u; acc[2] += o * u * u;
    acc[3] += r; acc[4] += r * u;
    acc[5] += cc; acc[6] += cc * u;
    sc[0] += -n * lg - sigma * r;
    sc[1] += n * sigma - 2.0 * sigma * n * inv + sigma * ns * inv * inv;
    sc[2] += s * eta - (s + n * theta) * lg;
    return;
  }
  if (MODE == PASS_IRLS) {
    const double sm = sigma * mu, t1 = 1.0 + sm;
    const double inv = rcp_pos(t1);
    const double w = n * mu * inv;
    const double r = (s - n * mu) * inv;
    const double wz = w * (eta - off) + r;                 // w z, z = (eta - offset) + (y - mu)/mu
    acc[0] += w; acc[1] += w * u; acc[2] += w * u * u;
    acc[3] += wz; acc[4] += wz * u;
    const double lg = sc_log1p_pos(sm, t1, inv, ft);
    const double t = s * eta - (s + n * theta) * lg;
    sc[0] -= t;
    const double rmu = rcp_pos(mu);
    sc[1] += (q - 2.0 * s * mu + n * mu * mu) * rmu * rmu;
    if (want_ll) sc[2] += t;
    return;
  }
  if (MODE == PASS_THETA) {
    const double sm = sigma * mu, t1 = 1.0 + sm;
    const double inv = rcp_pos(t1);
    const double r = (s - n * mu) * inv;
    const double lg = sc_log1p_pos(sm, t1, inv, ft);
    sc[0] += -n * lg - sigma * r;
    sc[1] += -n * sigma + 2.0 * sigma * n * inv - sigma * (n + sigma * s) * inv * inv;
    return;
  }
}

// A whole bucket with a flat linear predictor (c_b == 0): the point sums collapse to the bucket totals
// (N_b, U1, U2 from the lineage; SB, SBu, SBuu, QB from the gene).
template <int MODE>
SC_HD void flat_bucket_contrib(double Nb, double U1, double U2, double SB, double SBu, double SBuu, double QB, double a,
                               double sigma, double theta, bool want_ll, const FastTabs* ft, double* acc, double* sc) {
  if (Nb == 0.0) return;
  const double eta = a;
  const double mu = sc_exp(eta, ft);
  if (MODE == PASS_FIT && sigma == 0.0) {
    acc[0] += mu * Nb; acc[1] += mu * U1; acc[2] += mu * U2;
    acc[3] += SB - Nb * mu; acc[4] += SBu - U1 * mu;
    sc[2] += SB * eta - Nb * mu;
    return;
  }
  const double sm = sigma * mu, t1 = 1.0 + sm;
  const double inv = rcp_pos(t1);
  if (MODE == PASS_SWEEP) {
    const double w = mu * inv;
    acc[0] += w * Nb; acc[1] += w * U1; acc[2] += w * U2;
    acc[3] += (SB - Nb * mu) * inv; acc[4] += (SBu - U1 * mu) * inv;
    return;
  }
  const double lg = sc_log1p_pos(sm, t1, inv, ft);
  if (MODE == PASS_FIT) {
    const double f = mu * inv * inv;
    const double r0 = (SB - Nb * mu) * inv, r1 = (SBu - U1 * mu) * inv;
    const double g = sm * inv;
    acc[0] += f * (Nb + sigma * SB); acc[1] += f * (U1 + sigma * SBu); acc[2] += f * (U2 + sigma * SBuu);
    acc[3] += r0; acc[4] += r1;
    acc[5] += g * r0; acc[6] += g * r1;
    sc[0] += -Nb * lg - sigma * r0;
    sc[1] += Nb * sigma - 2.0 * sigma * Nb * inv + sigma * (Nb + sigma * SB) * inv * inv;
    sc[2] += SB * eta - (SB + Nb * theta) * lg;
    return;
  }
  if (MODE == PASS_IRLS) {
    const double w = mu * inv;
    const double r0 = (SB - Nb * mu) * inv, r1 = (SBu - U1 * mu) * inv;
    acc[0] += w * Nb; acc[1] += w * U1; acc[2] += w * U2;
    acc[3] += w * eta * Nb + r0; acc[4] += w * eta * U1 + r1;
    const double t = SB * eta - (SB + Nb * theta) * lg;
    sc[0] -= t;
    const double rmu = rcp_pos(mu);
    sc[1] += (QB - 2.0 * SB * mu + Nb * mu * mu) * rmu * rmu;
    if (want_ll) sc[2] += t;
    return;
  }
  if (MODE == PASS_THETA) {
    const double r0 = (SB - Nb * mu) * inv;
    sc[0] += -Nb * lg - sigma * r0;
    sc[1] += -Nb * sigma + 2.0 * sigma * Nb * inv - sigma * (Nb + sigma * SB) * inv * inv;
    return;
  }
}

// The mu-free part of a pass, term k of the sum over the gene's tail counts: Ck = #{y > k}, Ck1 = #{y > k + 1}.
template <int MODE>
SC_HD void hist_contrib(int k, double Ck, double Ck1, double sigma, double theta, bool want_ll, double* sc) {
  if (MODE == PASS_FIT) {
    if (sigma == 0.0) return;
    const double d = rcp_pos(theta + (double)k);
    sc[0] += Ck * d;
    sc[1] -= Ck * d * d;
    if (k >= 1) sc[2] += Ck * log1p((double)k * sigma);
    return;
  }
  if (MODE == PASS_IRLS) {
    const double y = (double)(k + 1);
    sc[0] -= (Ck - Ck1) * (y + theta) * log1p(sigma * y);
    if (want_ll && k >= 1) sc[2] += Ck * log1p((double)k * sigma);
    return;
  }
  if (MODE == PASS_THETA) {
    const double d = rcp_pos(theta + (double)k);
    sc[0] += Ck * d;
    sc[1] += Ck * d * d;
    return;
  }
}

// The mu-free part of a pass for one cell with a large count (not in the histogram): direct special-function evaluation.
template <int MODE>
SC_HD void big_contrib(int yi, double sigma, double theta, bool want_ll, double* sc) {
  const double y = (double)yi;
  if (MODE == PASS_FIT) {
    if (sigma == 0.0) return;
    double dpsi, dtri;
    psi_tri_diff_big(yi, theta, dpsi, dtri);
    sc[0] += dpsi;
    sc[1] -= dtri;
    sc[2] += nb_lambda_big(yi, sigma, theta);
    return;
  }
  if (MODE == PASS_IRLS) {
    sc[0] -= (y + theta) * log1p(sigma * y);
    if (want_ll) sc[2] += nb_lambda_big(yi, sigma, theta);
    return;
  }
  if (MODE == PASS_THETA) {
    double dpsi, dtri;
    psi_tri_diff_big(yi, theta, dpsi, dtri);
    sc[0] += dpsi;
    sc[1] += dtri;
    return;
  }
}

// ---------------------------------------------------------------------------------------------
// Offset quadrature: the intercept-only glm.nb WITH a size-factor offset (testDynamic.R:228-260; also an intercept-only
// alternative model, marge2.R:826-847) without a pass over the cells.
//
// With the offset o_i = log(1/sf_i) the null model's mean is mu_i = exp(b0 + o_i): every sum of every IRLS / theta.ml /
// log-likelihood pass is  sum_i [A(b0 + o_i) + y_i B(b0 + o_i) + y_i^2 C(b0 + o_i)]  with A, B, C analytic in o (their
// only singularities sit at Im o = +-pi, where 1 + sigma mu = 0).  The o axis is cut into K bins of width w <= 1/2 and
// inside each bin the integrand is replaced by its degree-15 interpolant at the kOffJ = 16 Chebyshev nodes of the bin.
// Interpolation is linear in the function values, so the sums over cells collapse ONCE per gene into node weights
//     n_kj = sum_{i in bin k} l_j(xi_i),   s_kj = sum y_i l_j(xi_i),   q_kj = sum y_i^2 l_j(xi_i)
// (l_j = Lagrange basis at the Chebyshev nodes = (1/J) [1 + 2 sum_{m>=1} T_m(xi_j) T_m(xi)], i.e. Chebyshev moments of the
// cells followed by a J x J cosine transform), and a pass is K*J "node" evaluations -- the same algebra as a point of the
// X-compressed layout with (n, s, q) replaced by the node weights.  Error: for f analytic inside the Bernstein ellipse
// of parameter rho the interpolation error is <= 4 M rho^-J / (rho - 1); with half-width h = 1/4 and the ellipse through
// Im o = pi/2 (where |1 + sigma mu| >= 1, so M is of the size of f on the real axis) rho = 12.5: 4 * 12.5^-16 / 11.5 ~ 1e-18
// relative -- below fp64 rounding.  Measured: records equal to the per-cell fits to ~1e-13 (tests/test_host_logic.py,
// tests/test_gpu_parity.py).
// ---------------------------------------------------------------------------------------------
constexpr int kOffJ = 16;               // Chebyshev nodes per bin
constexpr int kOffKMax = 64;            // bins: offsets may span e^32
constexpr double kOffBinWidth = 0.5;

struct OffsetNodes {
  int on;                               // 0 = not built (no offset, or offsets out of range): per-cell fits
  int K;
  double omin, w;
  double sum_o;                         // sum of the offsets over the lineage's cells
  int kstart[kOffKMax + 1];             // range of bin k in the bin-ordered cell list
  double ctab[kOffJ][kOffJ];            // T_m(xi_j) = cos(m (2j + 1) pi / (2J))
  double node_o[kOffKMax * kOffJ];      // node abscissae
  double node_n[kOffKMax * kOffJ];      // gene independent node weights (cells)
};

// Chebyshev moments of one bin (cm[m] = sum_i v_i T_m(xi_i)) -> node weights (out[j] = sum_i v_i l_j(xi_i))
SC_HD void off_moments_to_nodes(const OffsetNodes& O, const double* cm, double* out) {
  for (int j = 0; j < kOffJ; ++j) {
    double s = 0.0;
    for (int m = kOffJ - 1; m >= 1; --m) s += O.ctab[m][j] * cm[m];
    out[j] = (cm[0] + 2.0 * s) / (double)kOffJ;
  }
}

// One node of the offset quadrature: weights (n, s, q), offset o, intercept a (the fits that run here have no other column).
// Same sums as point_contrib() at u = 0 with eta = a + o and the working response taken off the offset.
template <int MODE>
SC_HD void offnode_contrib(double n, double s, double q, double o, double a, double sigma, double theta, bool want_ll,
                           const FastTabs* ft, double* acc, double* sc) {
  const double eta = a + o;
  const double mu = sc_exp(eta, ft);
  const double sm = sigma * mu, t1 = 1.0 + sm;
  const double inv = rcp_pos(t1);
  const double r = (s - n * mu) * inv;
  const double lg = sc_log1p_pos(sm, t1, inv, ft);
  if (MODE == PASS_IRLS) {
    const double w = n * mu * inv;
    acc[0] += w;
    acc[3] += w * a + r;                                   // w z, z = (eta - offset) + (y - mu)/mu
    const double t = s * eta - (s + n * theta) * lg;
    sc[0] -= t;
    const double rmu = rcp_pos(mu);
    sc[1] += (q - 2.0 * s * mu + n * mu * mu) * rmu * rmu;
    if (want_ll) sc[2] += t;
    return;
  }
  if (MODE == PASS_THETA) {
    sc[0] += -n * lg - sigma * r;
    sc[1] += -n * sigma + 2.0 * sigma * n * inv - sigma * (n + sigma * s) * inv * inv;
    return;
  }
}

// ---------------------------------------------------------------------------------------------
// theta.ml on Chebyshev nodes (compressed route).
//
// MASS::theta.ml (glm.nb: once per alternation, restarted from the moment estimate every time) runs up to 25 Newton steps on
// theta AT FIXED mu -- ~70 % of the passes of the final stage.  With mu fixed, both sums of a step are
//     sum_d [ n_d A(eta_d; theta) + s_d B(eta_d; theta) ],   eta_d = a_b + c_b u_d  inside bucket b,
// with A, B analytic in eta (singularities at Im eta = +-pi).  Per theta.ml call the points of every sloped bucket are
// therefore condensed ONCE: the bucket's eta range is cut into bins of width <= kThBinWidth and inside a bin the integrand
// is replaced by its degree-15 interpolant at the kOffJ = 16 Chebyshev nodes, which turns the point sums into node weights
//     n_kj = sum_{d in bin k} n_d l_j(xi_d),   s_kj = sum s_d l_j(xi_d)
// (Chebyshev moments of the points + a 16 x 16 cosine transform, as for the offset quadrature above).  A Newton step then
// costs ~16 evaluations per bin (a few hundred per gene) instead of one per abscissa (up to 10,001).  Flat buckets are one
// entry (their totals); buckets with too few points to gain keep their points.  Error: half-width 0.375 in eta, ellipse
// through Im eta = pi/2: rho = 8.5, 4 rho^-16 / (rho - 1) ~ 7e-16 relative.  The layout is fixed by (a, c, the point
// table) alone and every sum runs in a fixed order, so records stay bit-reproducible.
// ---------------------------------------------------------------------------------------------
constexpr int kThCap = 1024;             // node-list entries (shared memory: 3 doubles each)
constexpr double kThBinWidth = 0.75;

// Bucket nodes: the points of a knot bucket condensed ONCE per gene into 16 Chebyshev nodes along u (fit independent).
// Every sum of every pass is  sum_d [n_d A(u_d) + s_d B(u_d) + q_d C(u_d)]  with A, B, C analytic functions of z = a_b + c_b u
// (times a power of u); point_contrib() is linear in (n, s, q), so evaluating it at the 16 nodes of the bucket with the node
// weights (N_j, S_j, Q_j) = sum_d (n_d, s_d, q_d) l_j(xi_d) gives the sums of the degree-15 interpolant -- as for the offset
// quadrature above and with the same bound: while |c_b| x (width of the bucket in u) <= kThBinWidth the error is below fp64
// rounding; a bucket whose slope is steeper in some pass walks its points in that pass.  A pass over a 385-point bucket
// (200,000 cells) becomes 16 evaluations.
struct BucketNodes {                     // per lineage (gene independent part)
  int elig[NBMAX];                       // 1 = the bucket has a node set (enough points to pay)
  double ulo[NBMAX], uw[NBMAX];          // smallest coordinate / width in u of the bucket's points
  double un[NBMAX][16];                  // node coordinates
  double nn[NBMAX][16];                  // node weights of the cells
};
constexpr int kBnMinPoints = 48;

struct ThetaNodes {
  int on;                                // 0: the list does not fit / would not pay -> theta passes walk the points
  int skip;                              // 1: every sloped bucket runs on its bucket nodes in this theta.ml call, no list needed
  int n;                                 // entries
  int first[NBMAX + 1];                  // first entry of bucket b
  int bins[NBMAX + 1];                   // 0 = flat or empty bucket (one entry), -1 = the bucket's points as they are, K >= 1 = K bins
  double ulo[NBMAX + 1], h[NBMAX + 1];   // binned buckets: smallest coordinate and bin width (in u)
  double en[kThCap], es[kThCap], eeta[kThCap];
};

// Leader only.  ufirst / ulast: coordinate of the first / last point of bucket b (points are sorted inside a bucket).
template <class UF>
SC_HD void theta_nodes_layout(ThetaNodes& tn, int T, const double* c, const int* pbstart, UF u_of) {
  int n = 0, saved = 0;
  for (int b = 0; b <= T; ++b) {
    const int p0 = pbstart[b], p1 = pbstart[b + 1], P = p1 - p0;
    tn.first[b] = n;
    if (c[b] == 0.0 || P <= 0) { tn.bins[b] = 0; n += 1; continue; }
    const double u0 = u_of(b, p0), u1 = u_of(b, p1 - 1);
    const double span = fabs(c[b]) * (u1 - u0);
    int K = (int)ceil(span / kThBinWidth);
    if (K < 1) K = 1;
    if (!(span == span) || K > kThCap || 3 * K * kOffJ > 2 * P) { tn.bins[b] = -1; n += P; }       // not worth it: keep the points
    else { tn.bins[b] = K; tn.ulo[b] = u0; tn.h[b] = (u1 - u0) / (double)K; n += K * kOffJ; saved += P - K * kOffJ; }
    if (n > kThCap) { tn.on = 0; tn.n = 0; return; }
  }
  tn.n = n;
  tn.on = (n <= kThCap && saved > 0) ? 1 : 0;
}

// One bin, serially (host twin of the warp-parallel build in sclane_comp_device.cuh): points [lo, hi) -> the bin's 16 entries.
template <class UF>
SC_HD void theta_nodes_bin_serial(ThetaNodes& tn, int b, int k, int lo, int hi, const double* pn, const double* s1, double a, double c,
                                  const double (*ctab)[kOffJ], UF u_of) {
  const double u0 = tn.ulo[b] + (double)k * tn.h[b], h = tn.h[b];
  double cn[kOffJ], cs[kOffJ];
  for (int m = 0; m < kOffJ; ++m) cn[m] = cs[m] = 0.0;
  for (int d = lo; d < hi; ++d) {
    double xi = h > 0.0 ? 2.0 * (u_of(b, d) - u0) / h - 1.0 : 0.0;
    xi = xi < -1.0 ? -1.0 : (xi > 1.0 ? 1.0 : xi);
    double t0 = 1.0, t1 = xi;
    cn[0] += pn[d]; cs[0] += s1[d];
    cn[1] += pn[d] * xi; cs[1] += s1[d] * xi;
    for (int m = 2; m < kOffJ; ++m) { const double t2 = 2.0 * xi * t1 - t0; cn[m] += pn[d] * t2; cs[m] += s1[d] * t2; t0 = t1; t1 = t2; }
  }
  for (int j = 0; j < kOffJ; ++j) {
    double sn = 0.0, ss = 0.0;
    for (int m = kOffJ - 1; m >= 1; --m) { sn += ctab[m][j] * cn[m]; ss += ctab[m][j] * cs[m]; }
    const int e = tn.first[b] + k * kOffJ + j;
    tn.en[e] = (cn[0] + 2.0 * sn) / (double)kOffJ;
    tn.es[e] = (cs[0] + 2.0 * ss) / (double)kOffJ;
    tn.eeta[e] = a + c * (u0 + 0.5 * (ctab[1][j] + 1.0) * h);
  }
}

// negative.binomial(theta = 1)$initialize start of glm.fit: per cell mu = y + (y == 0)/6, eta = log(mu), sigma = 1.
// w = mu / (1 + mu), w z = w eta + (y - mu)/(1 + mu).
SC_HD void start_cell(int yi, double& w, double& wz) {
  const double y = (double)yi;
  const double mu = y + (yi == 0 ? 1.0 / 6.0 : 0.0);
  const double inv = 1.0 / (1.0 + mu);
  w = mu * inv;
  wz = w * log(mu) + (y - mu) * inv;
}

}  // namespace sclane
