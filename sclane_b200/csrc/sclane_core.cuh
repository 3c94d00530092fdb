// sclane_core.cuh -- the per-gene algorithm of the path, written once for host and device.
//
// Everything a gene needs after its cells have been streamed is a function of PER-BUCKET
// MOMENTS: the <= T candidate knots cut the (sorted) cells into T+1 buckets, and inside bucket b
// every dictionary column (intercept, tp1(x,t_k), tp2(x,t_k); tp1.R:17, tp2.R:17) is affine in the
// bucket-local coordinate u = x - ref_b:   column = alpha + gamma * u.  So X'WX, X'Wz, X'r of ANY
// basis drawn from the dictionary are small sums over buckets of (sum w, sum w u, sum w u^2, ...).
// A "pass" streams the gene's cells once and produces those moments; the fits below are
// sequences of passes with O(T p^2) algebra in between.
//
// The code is templated on an executor E that owns the pass:
//   E::pass<MODE>()  cooperative on device (one CTA per gene, sclane_device.cuh), a plain loop in
//                    the host test harness (tests/host_emu) -- never in the product library.
//   E::leader()/sync()  leader-only sections for the small algebra; all threads follow the same
//                    control flow by reading the shared workspace after a sync.
#pragma once
#include "../../include/sclane_b200.h"
#include "sclane_math.cuh"
#include "sclane_fastmath.cuh"

namespace sclane {

constexpr int PMAX = SCLANE_PMAX;     // 7 columns
constexpr int TMAX = SCLANE_TMAX;     // 32 knots
constexpr int NBMAX = TMAX + 1;       // buckets
constexpr int NACC = 7;               // per-bucket accumulators of a pass
constexpr int NSC = 4;                // scalar accumulators of a pass
constexpr int NSYS = PMAX + 1;        // joint (beta, log sigma) system
constexpr double kSigmaPoisson = 1e-4;   // gamlss.dist dNBI -> Poisson switch (DESIGN.md rule 2)
constexpr double kSigmaMax = 1e8;
constexpr double kEpsGlm = 1e-8;         // glm.control(epsilon)
constexpr int kMaxitGlm = 25;            // glm.control(maxit)
constexpr int kMaxNewton = 60;

enum PassMode : int {
  PASS_FIT = 0,        // NBI joint-Newton pass: observed-information moments, score, sigma derivatives, loglik
  PASS_IRLS = 1,       // glm.fit pass at eta = X beta + offset: Fisher moments, deviance, t0-sum, loglik
  PASS_IRLS_START = 2, // glm.fit pass at family$initialize: mustart = y + (y == 0)/6
  PASS_THETA = 3,      // theta.ml pass: score and information of theta at fixed mu
  PASS_SWEEP = 4,      // score sweep pass: reads (y, mu) -> weighted/unweighted bucket moments
  PASS_EMIT = 5        // writes mu = exp(eta) for the sweep
};

// Lineage geometry (gene independent; marge2.R:152-166 computed once per lineage, K0).
struct Lineage {
  int N;                 // cells in the lineage
  int Npad;              // row stride of the sorted count matrix (multiple of 128)
  int T;                 // candidate knots
  double knots[TMAX];    // ascending
  double ref[NBMAX];     // bucket reference point: ref[0] = knots[0], ref[b] = knots[b-1]
  int bstart[NBMAX + 1]; // cell index range of bucket b in sorted order: [bstart[b], bstart[b+1])
  double U0[NBMAX], U1[NBMAX], U2[NBMAX];   // unweighted moments: count, sum u, sum u^2
};

// alpha/gamma of dictionary column (kind, knot k) in bucket b.
SC_HD void term_ab(const Lineage& L, int kind, int k, int b, double& al, double& ga) {
  if (kind == SCLANE_TERM_INTERCEPT) { al = 1.0; ga = 0.0; }
  else if (kind == SCLANE_TERM_TP1) {
    if (b >= k + 1) { al = L.ref[b] - L.knots[k]; ga = 1.0; } else { al = 0.0; ga = 0.0; }
  } else {
    if (b <= k) { al = L.knots[k] - L.ref[b]; ga = -1.0; } else { al = 0.0; ga = 0.0; }
  }
}

// A model: ordered list of terms (intercept first).
struct Model {
  int n;
  int kind[PMAX];
  int knot[PMAX];
};

// Per-gene constants gathered while the counts are sorted (k_gather_stats).
struct GeneStats {
  double sum_y, sum_y2, sum_ylogy, sum_lgamma_y1;
  int ymax;
  int pad;
};

// Count-indexed tables of the mu-free special-function terms of the NB likelihood at the pass's theta / sigma.
// Most cells of single-cell data have small counts, so a pass looks these up instead of looping / dividing per cell;
// counts >= kTab fall back to the (out-of-line) direct evaluation.
constexpr int kTab = 64;
struct PassTabs {
  double psi[kTab];    // psi(th + y) - psi(th)            = sum_{k<y} 1/(th+k)
  double tri[kTab];    // trigamma(th) - trigamma(th + y)  = sum_{k<y} 1/(th+k)^2
  double lam[kTab];    // lgamma(th+y) - lgamma(th) + y log(sigma) = sum_{k=1}^{y-1} log1p(k sigma)
  double l1p[kTab];    // log1p(sigma y)
  double stage[3 * kTab];   // scratch of coop_set_tables
};

// Shared workspace of one gene (shared memory on device).  The leader thread mutates it between
// syncs; every thread reads it after a sync to take the same branches.
struct Work {
  // ---- pass inputs
  double a[NBMAX], c[NBMAX];   // eta = a[b] + c[b] * u (+ offset)
  double sigma;                // NB sigma = 1/theta; 0 => Poisson
  double theta;                // 1/sigma (unused when sigma == 0)
  int use_offset;
  // ---- pass outputs
  double acc[NBMAX][NACC];
  double sc[NSC];
  // ---- fit state
  double beta[PMAX];
  double s;                    // log sigma
  int flag;                    // loop control broadcast
  int flag2;
  int passes;
  // leader algebra scratch
  double H[NSYS * NSYS];
  double grad[NSYS];
  double step[NSYS];
  double cov[NSYS * NSYS];
  // nb_mle (Newton) state
  double nt_prev[NSYS];        // last accepted point (beta.., s)
  double nt_ell_prev;
  double nt_scale;
  int nt_have_prev, nt_poisson, nt_converged;
  // glm.fit2 state
  double gf_devold, gf_dev, gf_coefold[PMAX], gf_G[PMAX * PMAX];
  int gf_have_coefold, gf_ii, gf_conv;
  // theta.ml state
  double tm_theta, tm_del, tm_info;
  int tm_notes;
  // glm.nb alternation state
  double nb_th, nb_Lm, nb_Lm0, nb_del, nb_dev, nb_se_th, nb_t0sum;
  double nb_beta_stale[PMAX];
  int nb_iter, nb_notes;
  // backward pass state
  Model bw_cur, bw_models[PMAX];
  double bw_wic[PMAX], bw_cum;
  // sweep scores (NaN = NA)
  double sw_one[TMAX], sw_both[TMAX];
  double sw_one_r[TMAX], sw_both_r[TMAX];   // round(score, 6), computed alongside the scores
  // alpha/gamma tables of the current model's terms (coop_set_model): storage is provided by the kernel / harness
  // (TermTabs) or absent (nullptr: the warp-per-gene kernels compute term_ab() on the fly and keep Work small)
  double (*tal)[NBMAX];
  double (*tga)[NBMAX];
  int want_ll;                 // PASS_IRLS: also accumulate the log-likelihood (sc[2])
  int gee_p;                   // GEE passes: columns of the current model
  int sw_trunc, sw_kidx;       // forward selection broadcast
  double sw_score2;
  // misc broadcast
  double bc[4];
};

// storage of the alpha/gamma tables (Work.tal / Work.tga point here)
struct TermTabs {
  double tal[PMAX][NBMAX], tga[PMAX][NBMAX];
};
SC_HD void work_attach_tabs(Work& W, TermTabs* tt) {
  W.tal = tt ? tt->tal : nullptr;
  W.tga = tt ? tt->tga : nullptr;
}
// alpha/gamma of term j of model m in bucket b: from the tables (executors with E::kTabs: the CTA-per-gene kernels, where the
// small algebra between two passes is serial time every warp of the CTA waits for) or on the fly (warp-per-gene kernels)
template <bool TABS>
SC_HD void term_coef(const Lineage& L, const Model& m, const Work& W, int j, int b, double& al, double& ga) {
  if (TABS) { al = W.tal[j][b]; ga = W.tga[j][b]; }
  else term_ab(L, m.kind[j], m.knot[j], b, al, ga);
}

// ---- cooperative small algebra: every thread of the CTA calls these; work is split with E::par_for ----
// alpha/gamma tables of the model's terms.
struct SetModelFn {
  const Lineage* L; const Model* m; Work* W;
  SC_HD void operator()(int i) const {
    const int nb = L->T + 1;
    const int j = i / nb, b = i - j * nb;
    double al, ga;
    term_ab(*L, m->kind[j], m->knot[j], b, al, ga);
    W->tal[j][b] = al;
    W->tga[j][b] = ga;
  }
};
template <class E>
SC_HD void coop_set_model(E& ex, const Lineage& L, const Model& m, Work& W) {
  ex.sync();
  if (!E::kTabs) return;                          // no tables: term_coef() computes on the fly
  SetModelFn fn{&L, &m, &W};
  ex.par_for(m.n * (L.T + 1), fn);
  ex.sync();
}
// eta parameters per bucket from beta (beta must live in shared memory).
template <bool TABS>
struct SetEtaFn {
  const Lineage* L; const Model* m; Work* W; const double* beta;
  SC_HD void operator()(int b) const {
    double a = 0.0, c = 0.0;
    for (int j = 0; j < m->n; ++j) { double al, ga; term_coef<TABS>(*L, *m, *W, j, b, al, ga); a += beta[j] * al; c += beta[j] * ga; }
    W->a[b] = a;
    W->c[b] = c;
  }
};
template <class E>
SC_HD void coop_set_eta(E& ex, const Lineage& L, const Model& m, Work& W, const double* beta) {
  ex.sync();
  SetEtaFn<E::kTabs> fn{&L, &m, &W, beta};
  ex.par_for(L.T + 1, fn);
  ex.sync();
}
// Tables for the current (W.sigma, W.theta): terms in parallel (W.H is free scratch at this point), then prefix sums
// in ascending k (the order a serial loop would use).
struct TabTermFn {
  const Work* W; PassTabs* tb;
  SC_HD void operator()(int k) const {
    const double th = W->theta, sg = W->sigma;
    const double d = th + (double)k;
    tb->stage[k] = 1.0 / d;
    tb->stage[kTab + k] = 1.0 / (d * d);
    tb->stage[2 * kTab + k] = k >= 1 ? log1p((double)k * sg) : 0.0;
    tb->l1p[k] = log1p(sg * (double)k);
  }
};
struct TabScanFn {
  PassTabs* tb;
  SC_HD void operator()(int y) const {
    double a = 0.0, b = 0.0, c = 0.0;
    for (int k = 0; k < y; ++k) { a += tb->stage[k]; b += tb->stage[kTab + k]; c += tb->stage[2 * kTab + k]; }
    tb->psi[y] = a; tb->tri[y] = b; tb->lam[y] = c;
  }
};
template <class E>
SC_HD void coop_set_tables(E& ex, Work& W) {
  ex.sync();
  if (W.sigma == 0.0) return;                 // Poisson passes do not use the tables (uniform branch)
  PassTabs* tb = ex.tabs();
  if (tb == nullptr) return;                  // executors without per-cell special functions (X-compressed, GEE)
  TabTermFn t{&W, tb};
  ex.par_for(kTab, t);
  ex.sync();
  TabScanFn sc{tb};
  ex.par_for(kTab, sc);
  ex.sync();
}

// Symmetric system from a pass: H[j][l] = <d_j, d_l> with accumulators I0..I0+2, grad[j] = <d_j, .> with J0, J0+1,
// and (K0 >= 0) the extra column H[j][p] = <d_j, .> with K0, K0+1.  Results in W.H (NSYS stride) / W.grad.
template <bool TABS>
struct SystemFn {
  const Lineage* L; const Model* m; Work* W; int I0, J0, K0;
  SC_HD void operator()(int t) const {
    const int p = m->n, nb = L->T + 1;
    const int ntri = p * (p + 1) / 2;
    if (t < ntri) {
      int j = 0;
      while ((j + 1) * (j + 2) / 2 <= t) ++j;
      const int l = t - j * (j + 1) / 2;
      double s = 0.0;
      for (int b = 0; b < nb; ++b) {
        double aj, gj, al, gl;
        term_coef<TABS>(*L, *m, *W, j, b, aj, gj);
        term_coef<TABS>(*L, *m, *W, l, b, al, gl);
        s += aj * al * W->acc[b][I0] + (aj * gl + al * gj) * W->acc[b][I0 + 1] + gj * gl * W->acc[b][I0 + 2];
      }
      W->H[j * NSYS + l] = s;
      W->H[l * NSYS + j] = s;
    } else if (t < ntri + p) {
      const int j = t - ntri;
      double s = 0.0;
      for (int b = 0; b < nb; ++b) { double al, ga; term_coef<TABS>(*L, *m, *W, j, b, al, ga); s += al * W->acc[b][J0] + ga * W->acc[b][J0 + 1]; }
      W->grad[j] = s;
    } else {
      const int j = t - ntri - p;
      double s = 0.0;
      for (int b = 0; b < nb; ++b) { double al, ga; term_coef<TABS>(*L, *m, *W, j, b, al, ga); s += al * W->acc[b][K0] + ga * W->acc[b][K0 + 1]; }
      W->H[j * NSYS + p] = s;
      W->H[p * NSYS + j] = s;
    }
  }
};
template <class E>
SC_HD void coop_system(E& ex, const Lineage& L, const Model& m, Work& W, int I0, int J0, int K0) {
  SystemFn<E::kTabs> fn{&L, &m, &W, I0, J0, K0};
  const int p = m.n;
  ex.par_for(p * (p + 1) / 2 + (J0 >= 0 ? p : 0) + (K0 >= 0 ? p : 0), fn);
  ex.sync();
}

// <d_j, d_l>_w from accumulators I0..I0+2 and <d_j, v> from I0, I0+1.
SC_HD double gram2(const Lineage& L, const double (*acc)[NACC], int I0, int kj, int tj, int kl, int tl) {
  double s = 0.0;
  for (int b = 0; b <= L.T; ++b) {
    double aj, gj, al, gl;
    term_ab(L, kj, tj, b, aj, gj);
    term_ab(L, kl, tl, b, al, gl);
    s += aj * al * acc[b][I0] + (aj * gl + al * gj) * acc[b][I0 + 1] + gj * gl * acc[b][I0 + 2];
  }
  return s;
}
SC_HD double lin2(const Lineage& L, const double (*acc)[NACC], int I0, int kj, int tj) {
  double s = 0.0;
  for (int b = 0; b <= L.T; ++b) {
    double aj, gj;
    term_ab(L, kj, tj, b, aj, gj);
    s += aj * acc[b][I0] + gj * acc[b][I0 + 1];
  }
  return s;
}
// unweighted versions from the lineage tables
SC_HD double gram_u(const Lineage& L, int kj, int tj, int kl, int tl) {
  double s = 0.0;
  for (int b = 0; b <= L.T; ++b) {
    double aj, gj, al, gl;
    term_ab(L, kj, tj, b, aj, gj);
    term_ab(L, kl, tl, b, al, gl);
    s += aj * al * L.U0[b] + (aj * gl + al * gj) * L.U1[b] + gj * gl * L.U2[b];
  }
  return s;
}

// ---------------------------------------------------------------------------------------------
// Per-cell contributions.  y: count, u: bucket-local coordinate, off: offset (0 if none),
// a/c: eta parameters of the cell's bucket.  acc[NACC], sc[NSC] are the running sums.
// ---------------------------------------------------------------------------------------------
// reciprocal of a positive, normal-range double: hardware seed + two Newton steps on device (1 ulp), a division on host
SC_HD double rcp_pos(double x) {
#ifdef __CUDA_ARCH__
  double r;
  asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(x));
  double e = fma(-x, r, 1.0);
  r = fma(r, e, r);
  e = fma(-x, r, 1.0);
  return fma(r, e, r);
#else
  return 1.0 / x;
#endif
}
SC_HD void tab_psi_tri(const PassTabs& tb, int y, double th, double& psi, double& tri) {
  if (y < kTab) { psi = tb.psi[y]; tri = tb.tri[y]; }
  else psi_tri_diff_big(y, th, psi, tri);
}
SC_HD double tab_lambda(const PassTabs& tb, int y, double sigma, double th) {
  return y < kTab ? tb.lam[y] : nb_lambda_big(y, sigma, th);
}

template <int MODE>
SC_HD void cell_contrib(int yi, double u, double off, double a, double c, double sigma, double theta, const PassTabs& tb,
                        const FastTabs* ft, double mu_in, bool want_ll, double* acc, double* sc, double& mu_out) {
  const double y = (double)yi;
  if (MODE == PASS_EMIT) {
    mu_out = sc_exp(a + c * u, ft);
    return;
  }
  if (MODE == PASS_SWEEP) {
    // stat_out_score_glm_null.R:27-30 / score_fun_glm.R:40,44: w = mu^2/V, r = mu*VS = (y-mu)/(1+mu sigma)
    const double inv = 1.0 / (1.0 + sigma * mu_in);
    const double w = mu_in * inv, r = (y - mu_in) * inv;
    acc[0] += w; acc[1] += w * u; acc[2] += w * u * u;
    acc[3] += r; acc[4] += r * u;
    acc[5] += y; acc[6] += y * u;
    return;
  }
  if (MODE == PASS_FIT) {
    const double eta = a + c * u;
    const double mu = sc_exp(eta, ft);
    if (sigma == 0.0) {   // Poisson
      const double r = y - mu;
      acc[0] += mu; acc[1] += mu * u; acc[2] += mu * u * u;
      acc[3] += r; acc[4] += r * u;
      sc[2] += y * eta - mu;
      return;
    }
    const double sm = sigma * mu, t1 = 1.0 + sm;
    const double inv = rcp_pos(t1);
    const double r = (y - mu) * inv;                       // d l / d eta
    const double o = mu * (1.0 + sigma * y) * inv * inv;   // - d2 l / d eta2
    const double cc = sm * r * inv;                        // - d2 l / d eta d(log sigma)
    const double lg = sc_log1p_pos(sm, t1, inv, ft);
    double dpsi, dtri;
    tab_psi_tri(tb, yi, theta, dpsi, dtri);
    const double g = dpsi - lg - sigma * r;                                                  // d l / d theta
    const double h = -dtri + sigma - 2.0 * sigma * inv + sigma * (1.0 + sigma * y) * inv * inv;  // d2 l / d theta2
    acc[0] += o; acc[1] += o * u; acc[2] += o * u * u;
    acc[3] += r; acc[4] += r * u;
    acc[5] += cc; acc[6] += cc * u;
    sc[0] += g; sc[1] += h;
    sc[2] += tab_lambda(tb, yi, sigma, theta) + y * eta - (y + theta) * lg;
    return;
  }
  if (MODE == PASS_IRLS || MODE == PASS_IRLS_START) {
    double eta, mu;
    if (MODE == PASS_IRLS_START) { mu = y + (yi == 0 ? 1.0 / 6.0 : 0.0); eta = log(mu); }   // negative.binomial()$initialize
    else { eta = a + c * u + off; mu = sc_exp(eta, ft); }
    const double sm = sigma * mu, t1 = 1.0 + sm;
    const double inv = rcp_pos(t1);
    const double w = mu * inv;                             // mu.eta^2 / variance(mu)
    const double r = (y - mu) * inv;
    const double wz = w * (eta - off) + r;                 // w * z, z = (eta - offset) + (y - mu)/mu.eta
    acc[0] += w; acc[1] += w * u; acc[2] += w * u * u;
    acc[3] += wz; acc[4] += wz * u;
    const double lg = sc_log1p_pos(sm, t1, inv, ft);
    // dev.resids: 2 (y log(max(1,y)/mu) - (y+th) log((y+th)/(mu+th))); the y log y part is a gene constant
    const double l1py = yi < kTab ? tb.l1p[yi] : log1p(sigma * y);
    sc[0] += yi > 0 ? (-y * eta - (y + theta) * (l1py - lg)) : theta * lg;
    const double e = (y - mu) * rcp_pos(mu);
    sc[1] += e * e;                                        // theta.ml start: n / sum((y/mu - 1)^2)
    if (want_ll) sc[2] += tab_lambda(tb, yi, sigma, theta) + y * eta - (y + theta) * lg;   // glm.nb loglik() minus sum lgamma(y+1)
    return;
  }
  if (MODE == PASS_THETA) {
    const double eta = a + c * u + off;
    const double mu = sc_exp(eta, ft);
    const double sm = sigma * mu, t1 = 1.0 + sm;
    const double inv = rcp_pos(t1);
    const double r = (y - mu) * inv;
    const double lg = sc_log1p_pos(sm, t1, inv, ft);
    double dpsi, dtri;
    tab_psi_tri(tb, yi, theta, dpsi, dtri);
    sc[0] += dpsi - lg - sigma * r;                                                           // theta.ml score
    sc[1] += dtri - sigma + 2.0 * sigma * inv - sigma * (1.0 + sigma * y) * inv * inv;        // theta.ml info
    return;
  }
}

// Per-cell contributions inside a bucket where the linear predictor is constant (gamma-coefficient c_b == 0 and no
// offset: intercept-only fits, and every bucket on the flat side of the model's hinges).  mu, 1/(1+sigma mu),
// log1p(sigma mu) and 1/mu are computed ONCE per bucket (FlatBucket) -- no transcendental per cell.
struct FlatBucket {
  double eta, mu, inv, lg, rmu;
};
SC_HD FlatBucket make_flat_bucket(double a, double sigma) {
  FlatBucket f;
  f.eta = a;
  f.mu = exp(a);
  f.inv = sigma == 0.0 ? 1.0 : 1.0 / (1.0 + sigma * f.mu);
  f.lg = sigma == 0.0 ? 0.0 : log1p(sigma * f.mu);
  f.rmu = 1.0 / f.mu;
  return f;
}
template <int MODE>
SC_HD void cell_contrib_flat(int yi, double u, const FlatBucket& f, double sigma, double theta, const PassTabs& tb,
                             bool want_ll, double* acc, double* sc, double& mu_out) {
  const double y = (double)yi;
  if (MODE == PASS_EMIT) { mu_out = f.mu; return; }
  if (MODE == PASS_FIT) {
    if (sigma == 0.0) {
      const double r = y - f.mu;
      acc[0] += f.mu; acc[1] += f.mu * u; acc[2] += f.mu * u * u;
      acc[3] += r; acc[4] += r * u;
      sc[2] += y * f.eta - f.mu;
      return;
    }
    const double r = (y - f.mu) * f.inv;
    const double o = f.mu * (1.0 + sigma * y) * f.inv * f.inv;
    const double cc = sigma * f.mu * r * f.inv;
    double dpsi, dtri;
    tab_psi_tri(tb, yi, theta, dpsi, dtri);
    const double g = dpsi - f.lg - sigma * r;
    const double h = -dtri + sigma - 2.0 * sigma * f.inv + sigma * (1.0 + sigma * y) * f.inv * f.inv;
    acc[0] += o; acc[1] += o * u; acc[2] += o * u * u;
    acc[3] += r; acc[4] += r * u;
    acc[5] += cc; acc[6] += cc * u;
    sc[0] += g; sc[1] += h;
    sc[2] += tab_lambda(tb, yi, sigma, theta) + y * f.eta - (y + theta) * f.lg;
    return;
  }
  if (MODE == PASS_IRLS) {
    const double w = f.mu * f.inv;
    const double r = (y - f.mu) * f.inv;
    const double wz = w * f.eta + r;
    acc[0] += w; acc[1] += w * u; acc[2] += w * u * u;
    acc[3] += wz; acc[4] += wz * u;
    const double l1py = yi < kTab ? tb.l1p[yi] : log1p(sigma * y);
    sc[0] += yi > 0 ? (-y * f.eta - (y + theta) * (l1py - f.lg)) : theta * f.lg;
    const double e = (y - f.mu) * f.rmu;
    sc[1] += e * e;
    if (want_ll) sc[2] += tab_lambda(tb, yi, sigma, theta) + y * f.eta - (y + theta) * f.lg;
    return;
  }
  if (MODE == PASS_THETA) {
    const double r = (y - f.mu) * f.inv;
    double dpsi, dtri;
    tab_psi_tri(tb, yi, theta, dpsi, dtri);
    sc[0] += dpsi - f.lg - sigma * r;
    sc[1] += dtri - sigma + 2.0 * sigma * f.inv - sigma * (1.0 + sigma * y) * f.inv * f.inv;
    return;
  }
}

// ---------------------------------------------------------------------------------------------
// NB maximum-likelihood fit of y ~ B - 1 (no offset): stands in for gamlss NBI
// (stat_out_score_glm_null.R:22-25, backward_sel_WIC.R:44-47).  Joint Newton on (beta, log sigma)
// with the observed information and a log-likelihood backtracking guard; sigma is clamped at
// kSigmaPoisson, and if the optimum sits on that bound with the likelihood still increasing towards
// sigma -> 0 the fit is Poisson (sigma := 0, DESIGN.md rule 2).
// On return W.acc / W.sc hold the last pass (taken at the pre-final-step point), W.beta / W.sigma /
// W.s the estimate.
// ---------------------------------------------------------------------------------------------
struct FitResult {
  int converged;
  int poisson;
};

// Joint observed information -H and the gradient from a PASS_FIT result: the beta blocks come from
// coop_system(I0 = 0, J0 = 3, K0 = 5); this adds the log-sigma entries.  Leader only.
SC_HD void finish_newton_system(const Model& m, Work& W, bool sigma_free) {
  const int p = m.n;
  if (sigma_free) {
    const double th = W.theta;
    const double G = W.sc[0], Hh = W.sc[1];
    W.H[p * NSYS + p] = -(th * G + th * th * Hh);   // - d2 l / d s2,  s = log sigma
    W.grad[p] = -th * G;                             //   d l / d s
  }
}

// One leader-side Newton decision after a PASS_FIT.  Sets W.flag = 1 when the iteration is over.
SC_HDN void nb_mle_step(const Model& m, Work& W, double smin, double smax) {
  const int p = m.n;
  const bool pois = W.nt_poisson != 0;
  const double ell = W.sc[2];
  bool accept;
  if (W.nt_have_prev) accept = (ell >= W.nt_ell_prev - 1e-10 * (fabs(W.nt_ell_prev) + 1.0));   // false for NaN
  else accept = (ell == ell) && !isinf(ell);
  if (!accept) {
    if (!W.nt_have_prev || W.nt_scale < 1e-6) {
      W.flag = 1;                                   // cannot improve: stay at the last accepted point
      if (W.nt_have_prev) {
        _Pragma("unroll 1") for (int j = 0; j < p; ++j) W.beta[j] = W.nt_prev[j];
        if (!pois) W.s = W.nt_prev[p];
      }
    } else {                                        // halve the last step and retry
      W.nt_scale *= 0.5;
      _Pragma("unroll 1") for (int j = 0; j < p; ++j) W.beta[j] = W.nt_prev[j] + W.nt_scale * W.step[j];
      if (!pois) {
        double s = W.nt_prev[p] + W.nt_scale * W.step[p];
        W.s = s < smin ? smin : (s > smax ? smax : s);
      }
    }
    return;
  }
  W.nt_have_prev = 1;
  W.nt_ell_prev = ell;
  const bool sigma_free = !pois;
  finish_newton_system(m, W, sigma_free);
  // active-set treatment of the bounds on s
  const bool at_lo = sigma_free && (W.s <= smin) && (W.grad[p] < 0.0);
  const bool at_hi = sigma_free && (W.s >= smax) && (W.grad[p] > 0.0);
  const bool s_moves = sigma_free && !at_lo && !at_hi;
  double Lm[NSYS * NSYS];
  bool ok = false;
  if (s_moves) {
    copy_lower<NSYS>(p + 1, W.H, Lm);
    ok = chol_factor<NSYS>(p + 1, Lm);
    if (ok) chol_solve<NSYS>(p + 1, Lm, W.grad, W.step);
  }
  if (!ok) {
    // beta block alone (always used when s is fixed); safeguarded 1-D Newton on s otherwise
    copy_lower<NSYS>(p, W.H, Lm);
    const bool okb = chol_factor<NSYS>(p, Lm);
    if (okb) chol_solve<NSYS>(p, Lm, W.grad, W.step);
    else { _Pragma("unroll 1") for (int j = 0; j < p; ++j) W.step[j] = 0.0; W.flag = 1; }   // singular information: stop here
    W.step[p] = 0.0;
    if (s_moves) {
      const double d2 = W.H[p * NSYS + p];
      W.step[p] = d2 > 0.0 ? W.grad[p] / d2 : (W.grad[p] > 0.0 ? 1.0 : -1.0);
    }
  }
  // step limits: |delta s| <= 2, |delta beta_j| <= 5
  double scale = 1.0;
  if (s_moves && fabs(W.step[p]) > 2.0) scale = 2.0 / fabs(W.step[p]);
  _Pragma("unroll 1") for (int j = 0; j < p; ++j) if (fabs(W.step[j]) * scale > 5.0) scale = 5.0 / fabs(W.step[j]);
  if (scale < 1.0) _Pragma("unroll 1") for (int j = 0; j <= p; ++j) W.step[j] *= scale;
  double gain = 0.0, dmax = 0.0;
  _Pragma("unroll 1") for (int j = 0; j < p; ++j) {
    gain += W.grad[j] * W.step[j];
    const double d = fabs(W.step[j]) / (1.0 + fabs(W.beta[j]));
    if (d > dmax) dmax = d;
  }
  if (s_moves) { gain += W.grad[p] * W.step[p]; if (fabs(W.step[p]) > dmax) dmax = fabs(W.step[p]); }
  _Pragma("unroll 1") for (int j = 0; j < p; ++j) W.nt_prev[j] = W.beta[j];
  W.nt_prev[p] = W.s;
  W.nt_scale = 1.0;
  const bool conv = (dmax <= 1e-8) || (0.5 * gain <= 1e-13 * (1.0 + fabs(ell)));
  _Pragma("unroll 1") for (int j = 0; j < p; ++j) W.beta[j] += W.step[j];
  if (s_moves) {
    const double s = W.s + W.step[p];
    W.s = s < smin ? smin : (s > smax ? smax : s);
  }
  if (conv && W.flag == 0) {
    if (at_lo) {            // optimum on the sigma = kSigmaPoisson bound, d l/d s < 0: Poisson regime
      W.nt_poisson = 1;
      W.nt_have_prev = 0;
    } else {
      W.flag = 1;
      W.nt_converged = 1;
    }
  }
}

template <class E>
SC_HDN FitResult nb_mle(E& ex, const Lineage& L, const Model& m, const GeneStats& gs, Work& W, bool warm) {
  const int p = m.n;
  const double smin = log(kSigmaPoisson), smax = log(kSigmaMax);
  coop_set_model(ex, L, m, W);
  if (ex.leader()) {
    if (!warm) {
      const double ybar = gs.sum_y / (double)L.N;
      for (int j = 0; j < p; ++j) W.beta[j] = 0.0;
      W.beta[0] = log(ybar);
      const double var = (gs.sum_y2 - gs.sum_y * ybar) / (double)(L.N > 1 ? L.N - 1 : 1);
      double sg = (var - ybar) / (ybar * ybar);
      if (!(sg > 0.1)) sg = 0.1;               // gamlss NBI initial sigma: max(.., 0.1)
      W.s = log(sg);
    }
    if (!(W.s >= smin)) W.s = smin;
    if (W.s > smax) W.s = smax;
    W.flag = 0;
    W.nt_poisson = 0;
    W.nt_have_prev = 0;
    W.nt_converged = 0;
    W.nt_scale = 1.0;
    W.use_offset = 0;
  }
  ex.sync();
  for (int it = 0; it < kMaxNewton; ++it) {
    if (ex.leader()) {
      if (W.nt_poisson) { W.sigma = 0.0; W.theta = 0.0; }
      else { W.sigma = exp(W.s); W.theta = 1.0 / W.sigma; }
      W.passes++;
    }
    coop_set_eta(ex, L, m, W, W.beta);
    coop_set_tables(ex, W);
    ex.template pass<PASS_FIT>();
    coop_system(ex, L, m, W, 0, 3, W.nt_poisson ? -1 : 5);
    if (ex.leader()) nb_mle_step(m, W, smin, smax);
    ex.sync();
    if (W.flag) break;
  }
  FitResult res;
  res.converged = W.nt_converged;
  res.poisson = W.nt_poisson;
  ex.sync();
  if (ex.leader()) {
    if (W.nt_poisson) { W.sigma = 0.0; W.theta = 0.0; }
    else { W.sigma = exp(W.s); W.theta = 1.0 / W.sigma; }
  }
  ex.sync();
  return res;
}

// Wald statistics of the non-intercept columns from the last PASS_FIT of nb_mle
// (backward_sel_WIC.R:47-52).  Leader only.  out[0..p-2].
SC_HDN void wald_from_fit(const Model& m, Work& W, bool poisson, double* out) {
  const int p = m.n;
  const int n = poisson ? p : p + 1;      // W.H / W.grad still hold the system of the last pass
  double Lm[NSYS * NSYS];
  copy_lower<NSYS>(n, W.H, Lm);
  if (!chol_factor<NSYS>(n, Lm)) {
    for (int j = 1; j < p; ++j) out[j - 1] = nan("");
    return;
  }
  chol_inverse<NSYS>(n, Lm, W.cov);
  for (int j = 1; j < p; ++j) {
    const double var = W.cov[j * NSYS + j];
    if (poisson) out[j - 1] = W.beta[j] / var;                  // fallback branch: coef / sqrt(diag)^2
    else out[j - 1] = (W.beta[j] * W.beta[j]) / var;            // (coef / se)^2
  }
}

// ---------------------------------------------------------------------------------------------
// Score sweep algebra (score_fun_glm.R:32-50 for all T knots x {one, both}) from PASS_SWEEP moments.
// ---------------------------------------------------------------------------------------------
// Rule 1: exact structural rank check standing in for fastLmPure's NA coefficients.
SC_HD bool cand_collinear(const Lineage& L, const Model& m, int k, bool both) {
  bool has_pair = false;
  for (int j = 1; j < m.n; ++j) {
    const double tj = L.knots[m.knot[j]];
    if (tj == L.knots[k] && (m.kind[j] == SCLANE_TERM_TP1 || both)) return true;   // duplicated column
    if (m.kind[j] == SCLANE_TERM_TP1)
      for (int l = 1; l < m.n; ++l)
        if (m.kind[l] == SCLANE_TERM_TP2 && L.knots[m.knot[l]] == tj) has_pair = true;
  }
  return both && has_pair;
}

// A = (B'WB)^-1 (stat_out_score_glm_null.R:33-40; eigenMapMatrixInvert on a PD matrix = plain Cholesky
// inverse).  W.H holds B'WB (coop_system with I0 = 0); the leader inverts it into W.cov.  false if not PD.
SC_HDN bool sweep_invert(const Model& m, Work& W) {
  const int p = m.n;
  if (p == 1) { W.cov[0] = 1.0 / W.H[0]; return true; }      // stat_out_score_glm_null.R:33-34
  double Lm[NSYS * NSYS];
  copy_lower<NSYS>(p, W.H, Lm);
  if (!chol_factor<NSYS>(p, Lm)) return false;
  chol_inverse<NSYS>(p, Lm, W.cov);
  return true;
}

// score of candidate knot k; both = pair (tp1, tp2) else tp1 only.  Any thread may call it once the
// model tables (coop_set_model) and W.cov (sweep_invert) are in place.  NaN = NA.
template <bool TABS>
SC_HDN double sweep_score(const Lineage& L, const Model& m, const Work& W, int k, bool both) {
  if (cand_collinear(L, m, k, both)) return nan("");
  const int p = m.n, nb = L.T + 1;
  const int c = both ? 2 : 1;
  double C[2][PMAX], D[2], bvec[2];
  for (int q = 0; q < c; ++q) {
    const int kind = q == 0 ? SCLANE_TERM_TP1 : SCLANE_TERM_TP2;
    for (int j = 0; j < p; ++j) C[q][j] = 0.0;
    double d = 0.0, bv = 0.0;
    // tp1(t_k) lives on buckets > k, tp2(t_k) on buckets <= k
    const int b0 = q == 0 ? k + 1 : 0, b1 = q == 0 ? nb : k + 1;
    for (int b = b0; b < b1; ++b) {
      double al, ga;
      term_ab(L, kind, k, b, al, ga);
      const double s0 = W.acc[b][0], s1 = W.acc[b][1], s2 = W.acc[b][2];
      const double w0 = al * s0 + ga * s1;      // <cand, 1>_w  restricted to the bucket
      const double w1 = al * s1 + ga * s2;      // <cand, u>_w
      for (int j = 0; j < p; ++j) { double tj, gj; term_coef<TABS>(L, m, W, j, b, tj, gj); C[q][j] += tj * w0 + gj * w1; }
      d += al * w0 + ga * w1;
      bv += al * W.acc[b][3] + ga * W.acc[b][4];
    }
    D[q] = d;
    bvec[q] = bv;
  }
  // S = D - C' A C  (the 2x2 D is diagonal: tp1 * tp2 == 0)
  double S[2][2] = {{0, 0}, {0, 0}};
  for (int q = 0; q < c; ++q)
    for (int r = 0; r <= q; ++r) {
      double t = 0.0;
      for (int j = 0; j < p; ++j) {
        double aj = 0.0;
        for (int l = 0; l < p; ++l) aj += W.cov[j * NSYS + l] * C[r][l];
        t += C[q][j] * aj;
      }
      S[q][r] = (q == r ? D[q] : 0.0) - t;
    }
  return llt_quadform(c, S[0][0], S[1][0], S[1][1], bvec[0], bvec[1]);
}

// R helpers over score vectors (NaN = NA)
SC_HD double rmax_na(const double* v, int n) {
  double m = -INFINITY;
  for (int i = 0; i < n; ++i) if (v[i] == v[i] && v[i] > m) m = v[i];
  return m;
}
SC_HD bool all_na(const double* v, int n) {
  for (int i = 0; i < n; ++i) if (v[i] == v[i]) return false;
  return true;
}
// which.max over round(score, 6) (vr holds the rounded values, NaN = NA): first maximum, or the last one for the
// `tail(which(max == round(score, 6)), 1)` idiom.
SC_HD int which_max_round6(const double* vr, int n, bool last) {
  double m = -INFINITY;
  int idx = -1;
  for (int i = 0; i < n; ++i) {
    const double r = vr[i];
    if (!(r == r)) continue;
    if (r > m || idx < 0) { m = r; idx = i; }
    else if (last && r == m) idx = i;
  }
  return idx;
}

// The selection tree of marge2.R:368-587 (q = 1) on per-vector SUMMARIES: all the tree looks at is, per score vector, whether
// it is all NA and its maximum (NA dropped); the winner's index is then the first or the last maximum of round(score, 6).
// Vectors: additive one / both, and the "interaction" ones -- the constant -1e5 in the first iteration (:244), duplicates of the
// additive ones afterwards (the "interaction with the intercept", :247-307).  Returns false for the breakFlag exit; else
// trunc_type (1 | 2), the vector the index is taken from (SEL_*) and whether the LAST maximum is meant.
struct VecSummary {
  bool na;             // every entry NA
  double mx;           // max over the non-NA entries
  int first, last;     // first / last index of the maximum of round(score, 6); -1 when all NA
};
enum SelVec : int { SEL_ONE = 0, SEL_BOTH = 1, SEL_ONE_INT = 2, SEL_BOTH_INT = 3 };
SC_HDN bool select_tree(const VecSummary& bi, const VecSummary& oi, const VecSummary& ba, const VecSummary& oa, int& trunc_type, int& vec,
                        int& last) {
  const bool bi_na = bi.na, oi_na = oi.na, ba_na = ba.na, oa_na = oa.na;
  const double mx_bi = bi.mx, mx_oi = oi.mx, mx_ba = ba.mx, mx_oa = oa.mx;
#define SEL(tt, v, l) do { trunc_type = (tt); vec = (v); last = (l) ? 1 : 0; return true; } while (0)
  if (bi_na && oi_na) {                                            // :368
    if (!ba_na && !oa_na) { if (mx_ba > mx_oa) SEL(2, SEL_BOTH, true); SEL(1, SEL_ONE, true); }
    if (ba_na && !oa_na) SEL(1, SEL_ONE, false);
    if (!ba_na && oa_na) SEL(2, SEL_ONE, false);                   // sic (:386): all-NA vector -> error
    return false;
  }
  if (bi_na && !oi_na) {                                           // :392
    if (ba_na) {
      if (oa_na) SEL(1, SEL_ONE_INT, true);
      if (mx_oi > mx_oa) SEL(1, SEL_ONE_INT, true);
      SEL(1, SEL_ONE, false);
    }
    if (mx_oi > mx_ba) { if (mx_oi > mx_oa) SEL(1, SEL_ONE_INT, true); SEL(1, SEL_ONE, true); }
    if (mx_ba <= mx_oa) SEL(1, SEL_ONE, false);
    SEL(2, SEL_BOTH, false);
  }
  if (!bi_na && oi_na) {                                           // :443
    if (ba_na) { if (mx_bi > mx_oa) SEL(2, SEL_BOTH_INT, true); SEL(1, SEL_ONE, false); }
    if (mx_bi > mx_ba) SEL(2, SEL_BOTH_INT, true);
    if (mx_ba <= mx_oa) SEL(1, SEL_ONE, false);
    SEL(2, SEL_BOTH, false);
  }
  if (mx_bi > mx_oi) {                                             // :494
    if (!ba_na) {
      if (mx_ba >= mx_bi) { if (mx_ba > mx_oa) SEL(2, SEL_BOTH, false); SEL(1, SEL_ONE, false); }
      if (mx_bi > mx_oa) SEL(2, SEL_BOTH_INT, true);
      SEL(1, SEL_ONE, false);
    }
    if (mx_oa >= mx_bi) SEL(1, SEL_ONE, false);
    SEL(2, SEL_BOTH_INT, true);
  }
  if (!ba_na) {                                                    // :538
    if (mx_ba >= mx_oi) { if (mx_ba <= mx_oa) SEL(1, SEL_ONE, false); SEL(2, SEL_BOTH, false); }
    if (mx_oi > mx_oa) SEL(1, SEL_ONE_INT, true);
    SEL(1, SEL_ONE, false);
  }
  if (oa_na) SEL(1, SEL_ONE_INT, true);
  if (mx_oa >= mx_oi) SEL(1, SEL_ONE, false);
  SEL(1, SEL_ONE_INT, true);
#undef SEL
}

SC_HD VecSummary summarize_scores(const double* v, const double* vr, int n) {
  VecSummary s;
  s.na = all_na(v, n);
  s.mx = rmax_na(v, n);
  s.first = which_max_round6(vr, n, false);
  s.last = which_max_round6(vr, n, true);
  return s;
}

// The same on explicit score vectors (T <= TMAX candidate knots): one/both = additive scores, *_r = round(score, 6); in_set: 0 for
// the first iteration, 1 afterwards.  kidx may come back as -1 (R subscript error).
SC_HDN bool select_candidate(const double* one, const double* both, const double* one_r, const double* both_r, int T, int in_set,
                             int& trunc_type, int& kidx) {
  const VecSummary oa = summarize_scores(one, one_r, T), ba = summarize_scores(both, both_r, T);
  VecSummary ci;                                                   // the constant -1e5 vector (rounds to itself)
  ci.na = false; ci.mx = -1e5; ci.first = 0; ci.last = T - 1;
  const VecSummary& oi = in_set ? oa : ci;
  const VecSummary& bi = in_set ? ba : ci;
  int vec = 0, last = 0;
  if (!select_tree(bi, oi, ba, oa, trunc_type, vec, last)) return false;
  const VecSummary& sv = vec == SEL_ONE ? oa : (vec == SEL_BOTH ? ba : (vec == SEL_ONE_INT ? oi : bi));
  kidx = last ? sv.last : sv.first;
  return true;
}

// stat_out() (stat_out.R:23-47) for a candidate basis from the unweighted moments.  The Gram B'B and B'y are
// assembled cooperatively into W.H / W.grad (StatOutFn); the leader solves and returns GCVq1 rounded to 10 dp
// (NaN when B'B is not PD).
struct StatOutFn {
  const Lineage* L; const Model* m; Work* W;
  SC_HD void operator()(int t) const {
    const int p = m->n;
    const int ntri = p * (p + 1) / 2;
    if (t < ntri) {
      int j = 0;
      while ((j + 1) * (j + 2) / 2 <= t) ++j;
      const int l = t - j * (j + 1) / 2;
      const double v = gram_u(*L, m->kind[j], m->knot[j], m->kind[l], m->knot[l]);
      W->H[j * NSYS + l] = v;
      W->H[l * NSYS + j] = v;
    } else {
      const int j = t - ntri;
      W->grad[j] = lin2(*L, W->acc, 5, m->kind[j], m->knot[j]);
    }
  }
};
SC_HDN double stat_out_gcvq1(const Lineage& L, const Model& m, const Work& W, const GeneStats& gs) {
  const int p = m.n;
  double Lm[NSYS * NSYS], coef[NSYS];
  copy_lower<NSYS>(p, W.H, Lm);
  if (!chol_factor<NSYS>(p, Lm)) return nan("");
  chol_solve<NSYS>(p, Lm, W.grad, coef);
  double fit2 = 0.0;
  for (int j = 0; j < p; ++j) fit2 += coef[j] * W.grad[j];
  const double N = (double)L.N;
  const double ybar = gs.sum_y / N;
  const double TSS = gs.sum_y2 - gs.sum_y * ybar;
  const double GCVnull = TSS / (N * (1.0 - 1.0 / N) * (1.0 - 1.0 / N));
  const double RSS = gs.sum_y2 - fit2;
  const double df1a = (double)p + 2.0 * (double)(p - 1) / 2.0;
  const double GCV1 = RSS / (N * (1.0 - df1a / N) * (1.0 - df1a / N));
  return r_round(1.0 - GCV1 / GCVnull, 1e10);
}

// ---------------------------------------------------------------------------------------------
// Per-gene state carried between the kernels of the path.
// ---------------------------------------------------------------------------------------------
struct GeneState {
  Model fwd;            // forward-pass basis so far
  double beta[PMAX];    // last forward null fit (warm start)
  double sigma;         // sigma of the last forward null fit (0 = Poisson)
  double s;             // log sigma for warm starts
  int fwd_done;         // forward pass finished (stop reason in fwd_stop)
  int fwd_stop;
  int m_count;          // marge2's m (1, 3, 5, ...)
  int error;            // SCLANE_NOTE_* bits accumulated
  int fwd_iters;
  int passes;
  double fwd_score[4];
  double fwd_sigma[4];
  Model fin;            // final model after the backward pass
  double wic[PMAX];
  int n_wic;
  long long cycles[3];  // profiling: clock64 ticks spent in the forward / backward / final kernels
  // approx.knot = FALSE: the gene's own knot table (the knots its forward pass selected, ascending); Model.knot indexes it
  int ex_T;
  int ex_pidx[PMAX];    // point index (abscissa) of each knot
  double ex_knots[PMAX];
  int done;             // bit 0: backward pass done, bit 1: final fits done (set by the X-compressed kernels; the per-cell
                        // kernels skip genes that already went through a stage)
};

// Forward-iteration bookkeeping after a sweep (marge2.R:368-758), in two leader steps around the cooperative
// stat_out Gram.  Step 1: selection tree -> candidate basis in W.bw_cur (W.flag = 1 when there is one).
SC_HDN void forward_select(const Lineage& L, GeneState& st, Work& W) {
  const int it = st.fwd_iters;
  st.fwd_iters++;
  W.flag = 0;
  int trunc_type = 0, kidx = -1;
  const int in_set = st.fwd.n > 1 ? 1 : 0;
  if (!select_candidate(W.sw_one, W.sw_both, W.sw_one_r, W.sw_both_r, L.T, in_set, trunc_type, kidx)) { st.fwd_done = 1; st.fwd_stop = 1; return; }
  if (kidx < 0) { st.fwd_done = 1; st.fwd_stop = 5; st.error |= SCLANE_NOTE_NUMERIC; return; }
  const bool pair = trunc_type == 2;
  const double score2 = pair ? W.sw_both[kidx] : W.sw_one[kidx];   // the re-scoring at :665-675 recomputes this same number
  if (it < 4) st.fwd_score[it] = score2;
  if (!(score2 == score2)) { st.fwd_done = 1; st.fwd_stop = 5; st.error |= SCLANE_NOTE_NUMERIC; return; }   // if (NA) -> error
  Model cand = st.fwd;
  cand.kind[cand.n] = SCLANE_TERM_TP1; cand.knot[cand.n] = kidx; cand.n++;
  if (pair) { cand.kind[cand.n] = SCLANE_TERM_TP2; cand.knot[cand.n] = kidx; cand.n++; }
  W.bw_cur = cand;
  W.sw_score2 = score2;
  W.flag = 1;
}
// Step 2: stat_out check and acceptance (marge2.R:676-758).
SC_HDN void forward_accept(const Lineage& L, GeneState& st, Work& W, const GeneStats& gstat, int M, double tols_score) {
  const Model cand = W.bw_cur;
  const double score2 = W.sw_score2;
  const double gcvq1 = stat_out_gcvq1(L, cand, W, gstat);
  if (!(gcvq1 == gcvq1)) { st.fwd_done = 1; st.fwd_stop = 5; st.error |= SCLANE_NOTE_NUMERIC; return; }
  if (gcvq1 < -10.0 || r_round(score2, 1e4) <= 0.0) { st.fwd_done = 1; st.fwd_stop = 2; return; }   // :681
  if (score2 < tols_score) { st.fwd_done = 1; st.fwd_stop = 3; return; }                              // :737
  for (int j = st.fwd.n; j < cand.n; ++j) st.beta[j] = 0.0;
  st.fwd = cand;
  st.m_count += 2;
  if (L.N <= st.fwd.n + 2 || st.m_count >= M || st.fwd.n + 2 > PMAX) { st.fwd_done = 1; st.fwd_stop = 4; }   // :756
}

// ---------------------------------------------------------------------------------------------
// Backward pass (marge2.R:763-812) -- a chain of NB refits + Wald statistics.
// ---------------------------------------------------------------------------------------------
SC_HD double nanmin(const double* v, int n) {
  double m = INFINITY;
  for (int i = 0; i < n; ++i) if (v[i] == v[i] && v[i] < m) m = v[i];
  return m;
}

template <class E>
SC_HDN void backward_pass(E& ex, const Lineage& L, const GeneStats& gstat, GeneState& st, Work& W) {
  // st and W are visible to all threads; the leader mutates them between syncs.
  const int ncol = st.fwd.n;
  ex.sync();
  if (ncol <= 2) {            // marge2.R:777 / :788 subscript error -> "MARGE model error"
    if (ex.leader()) { st.error |= SCLANE_NOTE_FWD_EMPTY; st.fin.n = 0; }
    ex.sync();
    return;
  }
  if (ex.leader()) {
    W.bw_cur = st.fwd;
    W.bw_models[0] = st.fwd;
    W.bw_cum = 0.0;
    for (int j = 0; j < ncol; ++j) W.beta[j] = st.beta[j];   // warm start: last forward null fit, new columns at 0
    W.s = st.s;
  }
  ex.sync();
  for (int i = 1; i <= ncol - 1; ++i) {
    FitResult fr = nb_mle(ex, L, W.bw_cur, gstat, W, true);
    if (ex.leader()) {
      double wald[PMAX];
#ifdef SCLANE_EMU_DEBUG
      printf("bw step %d: n=%d conv=%d pois=%d sigma=%g beta0=%g passes=%d\n", i, W.bw_cur.n, fr.converged, fr.poisson, W.sigma, W.beta[0], W.passes);
#endif
      const Model cur = W.bw_cur;
      wald_from_fit(cur, W, fr.poisson != 0, wald);
      const int nw = cur.n - 1;
      const double mn = nanmin(wald, nw);
      W.bw_cum += mn;
      W.bw_wic[i - 1] = W.bw_cum + 2.0 * (double)cur.n;
      if (i != ncol - 1) {
        // variable.lowest: first column attaining the minimum (:781,:791); drop it, keep the rest as warm start
        int low = -1;
        for (int j = 0; j < nw; ++j) if (wald[j] == mn) { low = j; break; }
        if (low < 0) { st.error |= SCLANE_NOTE_NUMERIC; low = 0; }         // every Wald statistic NA: R fails inside try()
        const int drop = low + 1;
        Model nxt;
        nxt.n = 0;
        for (int j = 0; j < cur.n; ++j) if (j != drop) {
          nxt.kind[nxt.n] = cur.kind[j]; nxt.knot[nxt.n] = cur.knot[j]; nxt.n++;
        }
        W.bw_cur = nxt;
        // the reduced model restarts from the intercept-only point (hinge coefficients of the larger
        // model are a poor start once a correlated column is gone); log sigma carries over
        for (int j = 0; j < PMAX; ++j) W.beta[j] = 0.0;
        W.beta[0] = log(gstat.sum_y / (double)L.N);
        if (fr.poisson) W.s = log(kSigmaPoisson);
      } else {
        W.bw_cur.n = 1;                       // survivor renamed "Intercept" (:798-799)
      }
      W.bw_models[i] = W.bw_cur;
    }
    ex.sync();
  }
  if (ex.leader()) {
    // which.min(WIC_vec_2): entry 1 is NA (the full model can never win); first minimum, NaN skipped
    int best = -1;
    double bm = INFINITY;
    for (int i = 0; i < ncol - 1; ++i)
      if (W.bw_wic[i] == W.bw_wic[i] && (best < 0 || W.bw_wic[i] < bm)) { bm = W.bw_wic[i]; best = i; }
    st.n_wic = ncol - 1;
    for (int i = 0; i < ncol - 1; ++i) st.wic[i] = W.bw_wic[i];
    if (best < 0) { st.error |= SCLANE_NOTE_NUMERIC; st.fin.n = 0; }
    else {
      const Model sel = W.bw_models[best + 1];
      // B[, colnames(B) %in% cnames]: forward order, duplicates dropped (:805-812)
      Model fin;
      fin.n = 0;
      for (int j = 0; j < st.fwd.n; ++j) {
        bool in = false;
        for (int l = 0; l < sel.n; ++l) if (sel.kind[l] == st.fwd.kind[j] && sel.knot[l] == st.fwd.knot[j]) in = true;
        for (int l = 0; l < fin.n; ++l) if (fin.kind[l] == st.fwd.kind[j] && fin.knot[l] == st.fwd.knot[j]) in = false;
        if (in) { fin.kind[fin.n] = st.fwd.kind[j]; fin.knot[fin.n] = st.fwd.knot[j]; fin.n++; }
      }
      st.fin = fin;
    }
  }
  ex.sync();
}

// ---------------------------------------------------------------------------------------------
// MASS::glm.nb(..., method = "glm.fit2", init.theta = 1, link = log) (marge2.R:841-847,
// testDynamic.R:254-260): schedule restated from MASS::glm.nb / theta.ml / stats::glm.fit + the glm2
// step-halving addition (third-party sources, not in the reference tree; pinned by the golden).
// ---------------------------------------------------------------------------------------------
struct NbResult {
  double coef[PMAX];
  double cov[PMAX * PMAX];   // (X'WX)^-1 of the last WLS solve, row-major p x p (stride PMAX)
  double theta, se_theta, loglik, deviance;
  int notes;
  int alternations, irls_iters;
  int ok;
};

// Solve the WLS normal equations of a PASS_IRLS result: G beta = X'Wz, with G / X'Wz assembled by
// coop_system(I0 = 0, J0 = 3) into W.H / W.grad.  Leader only.  beta -> W.step, G -> W.gf_G.
// false if G is not positive definite or the solution is not finite.
SC_HDN bool wls_solve(const Model& m, Work& W) {
  const int p = m.n;
  double Lm[NSYS * NSYS];
  for (int j = 0; j < p; ++j)
    for (int l = 0; l < p; ++l) { Lm[j * NSYS + l] = W.H[j * NSYS + l]; W.gf_G[j * PMAX + l] = W.H[j * NSYS + l]; }
  if (!chol_factor<NSYS>(p, Lm)) return false;
  chol_solve<NSYS>(p, Lm, W.grad, W.step);
  for (int j = 0; j < p; ++j) if (!(W.step[j] == W.step[j]) || isinf(W.step[j])) return false;
  return true;
}

SC_HD double dev_from_pass(const GeneStats& gstat, const Work& W) { return 2.0 * (gstat.sum_ylogy + W.sc[0]); }

// stats::glm.fit (+ glm2 step-halving) with family negative.binomial(theta, link = log).
// On entry W.acc / W.sc hold the pass at the starting point (etastart or mustart) under this theta.
// On exit: W.beta = coefficients, W.acc / W.sc = pass at the final point, W.gf_dev = deviance,
// W.gf_G = X'WX of the last solve, W.gf_conv = converged.  Returns the number of iterations, or -1
// on a hard failure ("no valid set of coefficients", non-PD normal equations).
template <class E>
SC_HDN int glm_fit2(E& ex, const Lineage& L, const Model& m, const GeneStats& gstat, Work& W, double theta) {
  const int p = m.n;
  ex.sync();
  if (ex.leader()) {
    W.gf_devold = dev_from_pass(gstat, W);
    W.gf_have_coefold = 0;        // glm.fit: coefold <- NULL (etastart/mustart start, no `start`)
    W.gf_conv = 0;
    W.flag = 0;
    W.sigma = 1.0 / theta; W.theta = theta;
  }
  ex.sync();
  coop_set_tables(ex, W);
  int it;
  for (it = 1; it <= kMaxitGlm; ++it) {
    coop_system(ex, L, m, W, 0, 3, -1);
    if (ex.leader()) {
      if (!wls_solve(m, W)) W.flag = 2;
      else {
        for (int j = 0; j < p; ++j) W.beta[j] = W.step[j];
        W.sigma = 1.0 / theta; W.theta = theta;
        W.want_ll = 0;
        W.passes++;
      }
    }
    ex.sync();
    if (W.flag == 2) return -1;
    coop_set_eta(ex, L, m, W, W.beta);
    ex.template pass<PASS_IRLS>();
    // (1) non-finite deviance: halve towards coefold until finite
    if (ex.leader()) { W.gf_dev = dev_from_pass(gstat, W); W.gf_ii = 1; }
    ex.sync();
    for (;;) {
      if (ex.leader()) {
        const bool finite = (W.gf_dev == W.gf_dev) && !isinf(W.gf_dev);
        W.flag2 = 0;
        if (!finite) {
          if (!W.gf_have_coefold || W.gf_ii > kMaxitGlm) W.flag = 2;
          else {
            W.gf_ii++;
            for (int j = 0; j < p; ++j) W.beta[j] = 0.5 * (W.beta[j] + W.gf_coefold[j]);
            W.passes++;
            W.flag2 = 1;
          }
        }
      }
      ex.sync();
      if (W.flag == 2) return -1;
      if (!W.flag2) break;
      coop_set_eta(ex, L, m, W, W.beta);
      ex.template pass<PASS_IRLS>();
      if (ex.leader()) W.gf_dev = dev_from_pass(gstat, W);
      ex.sync();
    }
    // (2) glm2: deviance went up -> halve until it goes down
    if (ex.leader()) {
      W.gf_ii = 1;
      W.bc[0] = ((W.gf_dev - W.gf_devold) / (0.1 + fabs(W.gf_dev)) >= kEpsGlm && it > 1 && W.gf_have_coefold) ? 1.0 : 0.0;
    }
    ex.sync();
    if (W.bc[0] != 0.0) {
      for (;;) {
        if (ex.leader()) {
          W.flag2 = 0;
          if ((W.gf_dev - W.gf_devold) / (0.1 + fabs(W.gf_dev)) > -kEpsGlm && W.gf_ii <= kMaxitGlm) {
            W.gf_ii++;
            for (int j = 0; j < p; ++j) W.beta[j] = 0.5 * (W.beta[j] + W.gf_coefold[j]);
            W.passes++;
            W.flag2 = 1;
          }
        }
        ex.sync();
        if (!W.flag2) break;
        coop_set_eta(ex, L, m, W, W.beta);
        ex.template pass<PASS_IRLS>();
        if (ex.leader()) W.gf_dev = dev_from_pass(gstat, W);
        ex.sync();
      }
    }
    // (3) convergence
    if (ex.leader()) {
      if (fabs(W.gf_dev - W.gf_devold) / (fabs(W.gf_dev) + 0.1) < kEpsGlm) { W.gf_conv = 1; W.flag = 1; }
      else {
        W.gf_devold = W.gf_dev;
        for (int j = 0; j < p; ++j) W.gf_coefold[j] = W.beta[j];
        W.gf_have_coefold = 1;
      }
    }
    ex.sync();
    if (W.flag == 1) break;
  }
  ex.sync();
  return it > kMaxitGlm ? kMaxitGlm : it;
}

// Executors that can condense the cells / points for the Newton steps of theta.ml (mu is fixed while theta.ml iterates)
// implement theta_prepare(); for the others this is a no-op.
template <class E>
SC_HD auto theta_prepare_if(E& ex, int) -> decltype(ex.theta_prepare(), void()) { ex.theta_prepare(); }
template <class E>
SC_HD void theta_prepare_if(E&, long) {}

// MASS::theta.ml(y, mu, n, limit = 25, eps = .Machine$double.eps^0.25).  mu is defined by the eta
// parameters currently in W.a / W.c (+ offset).  Result: W.tm_theta, W.tm_info (-> SE), W.tm_notes.
template <class E>
SC_HDN void theta_ml(E& ex, const Lineage& L, Work& W, double t0sum) {
  const double eps = 1.220703125e-4;     // 2^-13 = .Machine$double.eps^0.25
  ex.sync();
  theta_prepare_if(ex, 0);
  if (ex.leader()) {
    W.tm_theta = (double)L.N / t0sum;
    W.tm_info = nan("");
    W.tm_del = 1.0;
    W.flag = 0;
  }
  ex.sync();
  int it = 0;
  for (;;) {
    ++it;
    if (ex.leader()) {
      if (!(it < kMaxitGlm && fabs(W.tm_del) > eps)) W.flag = 1;
      else {
        W.tm_theta = fabs(W.tm_theta);
        W.theta = W.tm_theta;
        W.sigma = 1.0 / W.theta;
        W.passes++;
      }
    }
    ex.sync();
    if (W.flag) break;
    coop_set_tables(ex, W);
    ex.template pass<PASS_THETA>();
    if (ex.leader()) {
      W.tm_del = W.sc[0] / W.sc[1];
      W.tm_theta += W.tm_del;
      W.tm_info = W.sc[1];
    }
    ex.sync();
  }
  if (ex.leader()) {
    int notes = 0;
    if (W.tm_theta < 0.0) { W.tm_theta = 0.0; notes |= SCLANE_NOTE_TH_TRUNC; }
    if (it == kMaxitGlm) notes |= SCLANE_NOTE_TH_ITER_LIMIT;
    W.tm_notes = notes;
  }
  ex.sync();
}

// glm.nb.  `out` is written by the leader only.
template <class E>
SC_HDN void glm_nb(E& ex, const Lineage& L, const Model& m, const GeneStats& gstat, Work& W, bool use_offset, NbResult* out) {
  const int p = m.n;
  const double N = (double)L.N;
  const double d1 = sqrt(2.0 * fmax(1.0, N - (double)p));
  coop_set_model(ex, L, m, W);
  if (ex.leader()) {
    W.use_offset = use_offset ? 1 : 0;
    W.sigma = 1.0; W.theta = 1.0;       // init.theta = 1
    W.want_ll = 0;
    W.passes++;
    out->ok = 0;
    out->notes = 0;
  }
  ex.sync();
  coop_set_tables(ex, W);
  ex.template pass<PASS_IRLS_START>();
  int irls = glm_fit2(ex, L, m, gstat, W, 1.0);
  if (irls < 0) return;
  // th <- theta.ml(Y, mu) at the fresh mu of the initial fit (W.a / W.c still describe it)
  const double t0sum0 = W.sc[1];
  theta_ml(ex, L, W, t0sum0);
  if (ex.leader()) {
    W.nb_th = W.tm_theta;
    W.nb_se_th = W.tm_info;
    W.nb_notes = W.tm_notes;
    W.nb_dev = W.gf_dev;
    for (int j = 0; j < p * PMAX; ++j) out->cov[j] = W.gf_G[j];
    // pass at (beta, th): gives Lm and doubles as the starting pass of the next glm.fit
    W.theta = W.nb_th; W.sigma = 1.0 / W.nb_th;
    W.want_ll = 1;
    W.passes++;
  }
  ex.sync();
  coop_set_tables(ex, W);
  ex.template pass<PASS_IRLS>();
  if (ex.leader()) {
    W.nb_Lm = W.sc[2] - gstat.sum_lgamma_y1;
    W.nb_Lm0 = W.nb_Lm + 2.0 * d1;
    W.nb_del = 1.0;
    W.nb_iter = 0;
  }
  ex.sync();
  for (;;) {
    if (ex.leader()) {
      W.nb_iter++;
      W.flag = !(W.nb_iter <= kMaxitGlm && (fabs(W.nb_Lm0 - W.nb_Lm) / d1 + fabs(W.nb_del)) > kEpsGlm) ? 1 : 0;
      if (!W.flag) {
        for (int j = 0; j < p; ++j) W.nb_beta_stale[j] = W.beta[j];
        W.nb_t0sum = W.sc[1];
      }
    }
    ex.sync();
    if (W.flag) break;
    const double th_cur = W.nb_th;
    const int r = glm_fit2(ex, L, m, gstat, W, th_cur);
    if (r < 0) return;
    irls += r;
    if (ex.leader()) {
      W.nb_dev = W.gf_dev;
      for (int j = 0; j < p * PMAX; ++j) out->cov[j] = W.gf_G[j];
    }
    coop_set_eta(ex, L, m, W, W.nb_beta_stale);     // theta.ml at the PREVIOUS mu (MASS quirk)
    theta_ml(ex, L, W, W.nb_t0sum);
    if (ex.leader()) {
      const double t0 = W.nb_th;
      W.nb_th = W.tm_theta;
      W.nb_se_th = W.tm_info;
      W.nb_notes = W.tm_notes;
      W.nb_del = t0 - W.nb_th;
      W.theta = W.nb_th; W.sigma = 1.0 / W.nb_th;
      W.want_ll = 1;
      W.passes++;
    }
    coop_set_eta(ex, L, m, W, W.beta);
    coop_set_tables(ex, W);
    ex.template pass<PASS_IRLS>();
    if (ex.leader()) {
      W.nb_Lm0 = W.nb_Lm;
      W.nb_Lm = W.sc[2] - gstat.sum_lgamma_y1;
    }
    ex.sync();
  }
  if (ex.leader()) {
    out->ok = 1;
    out->irls_iters = irls;
    for (int j = 0; j < p; ++j) out->coef[j] = W.beta[j];
    out->theta = W.nb_th;
    const double info = W.nb_se_th;
    out->se_theta = (info == info && info > 0.0) ? sqrt(1.0 / info) : nan("");
    out->loglik = W.nb_Lm;
    out->deviance = W.nb_dev;
    int notes = W.nb_notes;
    if (W.nb_iter > kMaxitGlm) notes = SCLANE_NOTE_ALT_LIMIT;
    if (!W.gf_conv) notes |= SCLANE_NOTE_NOT_CONVERGED;
    out->notes = notes;
    out->alternations = W.nb_iter - 1;
    // out->cov holds G = X'WX of the last solve: invert for (X'WX)^-1
    double Lm[NSYS * NSYS], inv[NSYS * NSYS];
    for (int j = 0; j < p; ++j) for (int l = 0; l < p; ++l) Lm[j * NSYS + l] = out->cov[j * PMAX + l];
    if (chol_factor<NSYS>(p, Lm)) {
      chol_inverse<NSYS>(p, Lm, inv);
      for (int j = 0; j < p; ++j) for (int l = 0; l < p; ++l) out->cov[j * PMAX + l] = inv[j * NSYS + l];
    } else {
      for (int j = 0; j < p * PMAX; ++j) out->cov[j] = nan("");
    }
  }
  ex.sync();
}

// ---------------------------------------------------------------------------------------------
// The four per-gene stages as the kernels run them (K1 forward null fit, K2 score sweep,
// K4 backward pass, K5 final + null glm.nb).  `st` and `W` are CTA-shared on device.
// ---------------------------------------------------------------------------------------------
template <bool TABS>
struct ScoreFn {
  const Lineage* L;
  const Model* m;
  Work* W;
  SC_HD void operator()(int i) const {
    const int T = L->T;
    if (i < T) { const double v = sweep_score<TABS>(*L, *m, *W, i, false); W->sw_one[i] = v; W->sw_one_r[i] = r_round(v, 1e6); }
    else if (i < 2 * T) { const double v = sweep_score<TABS>(*L, *m, *W, i - T, true); W->sw_both[i - T] = v; W->sw_both_r[i - T] = r_round(v, 1e6); }
  }
};

// K1: stat_out_score_glm_null() (stat_out_score_glm_null.R:18-48) for the current forward basis:
// NB fit, then mu is written out for the sweep.
template <class E>
SC_HDN void gene_forward_fit(E& ex, const Lineage& L, const GeneStats& gstat, GeneState& st, Work& W) {
  ex.sync();
  if (st.fwd_done || st.error) return;
  const bool warm = st.fwd_iters > 0;   // later iterations start from the previous null fit (new columns at 0)
  if (ex.leader()) {
    for (int j = 0; j < st.fwd.n; ++j) W.beta[j] = st.beta[j];
    W.s = st.s;
  }
  ex.sync();
  FitResult fr = nb_mle(ex, L, st.fwd, gstat, W, warm);
  if (ex.leader()) {
    for (int j = 0; j < st.fwd.n; ++j) st.beta[j] = W.beta[j];
    st.sigma = fr.poisson ? 0.0 : W.sigma;
    st.s = fr.poisson ? log(kSigmaPoisson) : W.s;
    if (st.fwd_iters < 4) st.fwd_sigma[st.fwd_iters] = st.sigma;
    W.passes++;
  }
  coop_set_eta(ex, L, st.fwd, W, W.beta);
  ex.template pass<PASS_EMIT>();
}

// K2: the score sweep (score_fun_glm.R:21-54 for every knot and both candidate shapes), the
// selection tree, the stat_out check and the forward bookkeeping (marge2.R:182-758).
// The part of K2 after the moments are in W.acc (rows 0..4: weighted moments of this iteration, rows 5..6:
// the y moments): B'WB inverse, all 2T scores, selection tree, stat_out check, bookkeeping.  Runs on a CTA
// (host harness / k_stage) or on a single warp (k_sweep_select).
template <class E>
SC_HDN void gene_sweep_select(E& ex, const Lineage& L, const GeneStats& gstat, GeneState& st, Work& W, int M, double tols_score) {
  coop_set_model(ex, L, st.fwd, W);
  coop_system(ex, L, st.fwd, W, 0, -1, -1);                 // B'WB
  if (ex.leader()) W.flag = sweep_invert(st.fwd, W) ? 1 : 0;
  ex.sync();
  if (!W.flag) {
    if (ex.leader()) { st.error |= SCLANE_NOTE_NUMERIC; st.fwd_done = 1; st.fwd_stop = 5; }
    ex.sync();
    return;
  }
  ScoreFn<E::kTabs> fn{&L, &st.fwd, &W};
  ex.par_for(2 * L.T, fn);
  ex.sync();
  if (ex.leader()) forward_select(L, st, W);
  ex.sync();
  if (!W.flag) return;
  StatOutFn so{&L, &W.bw_cur, &W};
  ex.par_for(W.bw_cur.n * (W.bw_cur.n + 1) / 2 + W.bw_cur.n, so);
  ex.sync();
  if (ex.leader()) forward_accept(L, st, W, gstat, M, tols_score);
  ex.sync();
}

template <class E>
SC_HDN void gene_sweep(E& ex, const Lineage& L, const GeneStats& gstat, GeneState& st, Work& W, int M, double tols_score) {
  ex.sync();
  if (st.fwd_done || st.error) return;
  if (ex.leader()) {
    W.sigma = st.sigma;
    W.theta = st.sigma > 0.0 ? 1.0 / st.sigma : 0.0;
    W.use_offset = 0;
    W.passes++;
  }
  ex.sync();
  ex.template pass<PASS_SWEEP>();
  gene_sweep_select(ex, L, gstat, st, W, M, tols_score);
}

// K4
template <class E>
SC_HDN void gene_backward(E& ex, const Lineage& L, const GeneStats& gstat, GeneState& st, Work& W) {
  ex.sync();
  if (st.error) { if (ex.leader()) st.fin.n = 0; ex.sync(); return; }
  backward_pass(ex, L, gstat, st, W);
}

// K5: final glm.nb on the selected basis (marge2.R:841-847) and the intercept-only null glm.nb
// (testDynamic.R:254-260), both with the size-factor offset when one is given.
// parts: bit 0 = the intercept-only fits (the null model, copied to the alternative when only the intercept survived the
// backward pass), bit 1 = an alternative with hinges.  The two parts may run in different kernels / executors (with an
// offset: part 1 on the offset quadrature, part 2 per cell); 3 = both here.
template <class E>
SC_HDN void gene_final(E& ex, const Lineage& L, const GeneStats& gstat, GeneState& st, Work& W, bool use_offset,
                       NbResult* alt, NbResult* null, int parts) {
  ex.sync();
  if (parts & 1) {
    if (ex.leader()) { alt->ok = 0; null->ok = 0; alt->notes = 0; null->notes = 0; }
    ex.sync();
  }
  if (st.error & (SCLANE_NOTE_ALL_ZERO | SCLANE_NOTE_BAD_COUNTS)) return;
  const bool alt_usable = !(st.fin.n < 1 || (st.error & (SCLANE_NOTE_FWD_EMPTY | SCLANE_NOTE_NUMERIC)));
  if (parts & 1) {
    if (ex.leader()) { W.bw_cur.n = 1; W.bw_cur.kind[0] = SCLANE_TERM_INTERCEPT; W.bw_cur.knot[0] = 0; }
    ex.sync();
    glm_nb(ex, L, W.bw_cur, gstat, W, use_offset, null);
    ex.sync();
    if (alt_usable && st.fin.n == 1) {
      // intercept-only alternative: the same fit as the null model (LRT is exactly 0, modelLRT.R:48-50)
      if (ex.leader()) *alt = *null;
      ex.sync();
    }
  }
  if ((parts & 2) && alt_usable && st.fin.n > 1) glm_nb(ex, L, st.fin, gstat, W, use_offset, alt);
}

}  // namespace sclane
