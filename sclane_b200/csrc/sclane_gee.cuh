// sclane_gee.cuh -- GEE mode of the path (is.gee = TRUE), shared host/device algorithm.
//
// Reference: R/stat_out_score_gee_null.R:22-112, R/score_fun_gee.R:28-85, R/backward_sel_WIC.R:36-42,
// R/marge2.R:831-838, R/testDynamic.R:235-244, R/waldTestGEE.R:24-93; the fitter geeM::geem (third party, restated
// from its published algorithm -- see oracle/gee.py; PARITY UNPINNED for this mode).
//
// The reference builds dense n_i x n_i working covariances per subject (O(n_i^3)).  Here everything is O(n_i):
//   * AR-1:          R^-1 is tridiagonal: (R^-1 t)_j = (c_j t_j - alpha (t_{j-1} + t_{j+1})) / (1 - alpha^2),
//                    c_j = 1 at the subject's ends, 1 + alpha^2 inside (n_i = 1: identity);
//   * exchangeable:  R^-1 = (I - kappa_i J) / (1 - alpha), kappa_i = alpha / (1 + (n_i - 1) alpha);
//   * AWA_i = V_i^-1 S_i S_i' V_i^-1 is rank one, so with u_i = V_i^-1 S_i, a_i = Dn_i' u_i, b_i = Dx_i' u_i:
//       Sigma11 = sum a_i a_i',  Sigma21 = sum b_i a_i',  Sigma22 = sum b_i b_i',  B.est = sum b_i,
//       J11 = sum Dn_i' V_i^-1 Dn_i,  J21 = sum Dx_i' V_i^-1 Dn_i.
// Cells stay in the caller's order inside each subject (the AR-1 coupling is between neighbouring ROWS, the reference
// never sorts by pseudotime); hinge columns are evaluated per cell from its knot bucket: column = alpha + gamma * u.
#pragma once
#include "sclane_core.cuh"

namespace sclane {

constexpr int GEE_PMAX = 5;                 // M <= 5 in GEE mode
constexpr int GEE_SMAX = 64;                // subjects per lineage
constexpr int GEE_NSYM = GEE_PMAX * (GEE_PMAX + 1) / 2;
constexpr double kGeeTheta = 50.0;          // family = MASS::negative.binomial(50)
constexpr double kGeeTol = 1e-5;            // geem(tol)
constexpr int kGeeMaxit = 20;               // geem(maxit)

enum GeeCor : int { GEE_INDEP = 0, GEE_EXCH = 1, GEE_AR1 = 2 };
enum GeePass : int {
  GPASS_RESID = 0,    // Pearson residuals at beta: sum e^2, AR-1 neighbour products, per-subject sums
  GPASS_BETA = 1,     // updateBeta ingredients at beta: Z'R^-1 Z and Z'R^-1 r pieces
  GPASS_PREP1 = 2,    // exchangeable only: per-subject sums of t = sv * (S, mu B_1..)
  GPASS_PREP2 = 3,    // one subject: u = V^-1 S, Q = V^-1 Dn -> J11, a_i, bucket moments for the candidates
  GPASS_YMOM = 4,     // per-bucket sum y, sum y u into W.acc[b][5..6] (stat_out's B'y), once per gene
  GPASS_SAND = 5      // per-subject pieces of Z_i' R_i^-1 rr_i at beta (geeM getSandwich): ssum[i][0..p) = sum c_j z_j r_j,
                      // [p..2p) = AR-1 neighbour sums / exchangeable sum z, [2p] = exchangeable sum r
};

struct GeeLineage {
  int S;
  int sstart[GEE_SMAX + 1];     // subject i = cells [sstart[i], sstart[i+1]) of the lineage, caller's order
};

SC_HD int sym_idx(int j, int l) { return j >= l ? j * (j + 1) / 2 + l : l * (l + 1) / 2 + j; }

// Shared state of a GEE fit / sweep (in addition to Work, whose a/c/tal/tga/beta/H/grad/step/cov scratch is reused).
struct GeeWork {
  int cor;
  int subject;                 // GPASS_PREP2: the subject being processed
  double alpha, phi;           // working correlation and scale (fit passes) / NB dispersion (prep passes)
  // ---- pass outputs
  double A1[GEE_NSYM];         // sum c_j z_j z_j'            (PREP2: J11)
  double A2[GEE_NSYM];         // sum_adj (z_j z_{j+1}' + z_{j+1} z_j')
  double e1[GEE_PMAX];         // sum c_j z_j r_j             (PREP2: a_i of this subject)
  double e2[GEE_PMAX];         // sum_adj (z_j r_{j+1} + z_{j+1} r_j)
  double sc[4];                // RESID: [0] sum e^2, [1] sum_adj e_j e_{j+1}
  double ssum[GEE_SMAX][2 * GEE_PMAX + 1];   // per-subject sums: RESID [.][0] = sum e; BETA (exch) z (p) and r; PREP1 t (1 + p); SAND
  double Gm[NBMAX][2];         // PREP2, this subject: sum g, sum g u per bucket, g = mu * u_vec
  double Hm[NBMAX][2 * GEE_PMAX];   // PREP2, all subjects so far: sum h_l, sum h_l u per bucket, h_l = mu * Q_l
  // ---- geem state
  double beta_old[GEE_PMAX];
  double hess[GEE_PMAX * GEE_PMAX];
  int converged, niter;
  // ---- sweep state (per candidate column: tp1 columns 0..T-1, tp2 columns T..2T-1)
  double J11inv[GEE_PMAX * GEE_PMAX], JS11[GEE_PMAX * GEE_PMAX], S11[GEE_NSYM];
  double cb[2 * TMAX];                     // B.est per column
  double cS21[2 * TMAX][GEE_PMAX];         // Sigma21 row per column
  double cS22[2 * TMAX];                   // sum b_i^2 per column
  double cS22x[TMAX];                      // sum b_i(tp1_k) b_i(tp2_k)
  double cJ21[2 * TMAX][GEE_PMAX];         // J21 row per column
};

// ---------------------------------------------------------------------------------------------
// per-cell quantities
// ---------------------------------------------------------------------------------------------
struct GeeCell {
  double mu, s, r;             // mean, 1/sqrt(variance), s * (y - mu)
  double z[GEE_PMAX];          // s * mu * x_l
};
// fit passes: variance function of negative.binomial(50); prep passes: scLANE's V = mu (1 + mu phi)
template <bool PREP>
SC_HD void gee_cell(const Work& W, const GeeWork& G, int p, int yi, int b, double u, double off, GeeCell& c) {
  const double eta = W.a[b] + W.c[b] * u + off;
  c.mu = exp(eta);
  const double var = PREP ? c.mu * (1.0 + c.mu * G.phi) : c.mu + c.mu * c.mu / kGeeTheta;
  c.s = 1.0 / sqrt(var);
  c.r = c.s * ((double)yi - c.mu);
  const double sm = c.s * c.mu;
  for (int l = 0; l < p; ++l) c.z[l] = sm * (W.tal[l][b] + W.tga[l][b] * u);
}

// One cell of GPASS_PREP2: u_vec = (V^-1 S)_j, Q_l = (V^-1 Dn)_jl from the cell and (AR-1) its neighbours cp / cn
// (null at the subject's ends) or (exchangeable) the subject sums of GPASS_PREP1.
struct GeePrepOut {
  double uvec, g;              // g = mu * u_vec
  double q[GEE_PMAX];          // Q_l
  double h[GEE_PMAX];          // mu * Q_l
  double xm[GEE_PMAX];         // Dn_jl = mu * x_l
};
SC_HD void gee_prep_cell(const Work& W, const GeeWork& G, int p, int subj, int n, int qpos, int b, double u, const GeeCell& c,
                         const GeeCell* cp, const GeeCell* cn, GeePrepOut& o) {
  const double al = G.alpha;
  double rt[GEE_PMAX + 1];
  if (G.cor == GEE_AR1 && n > 1) {
    const double cj = (qpos == 0 || qpos == n - 1) ? 1.0 : 1.0 + al * al;
    const double f = 1.0 / (1.0 - al * al);
    rt[0] = f * (cj * c.r - al * ((cp ? cp->r : 0.0) + (cn ? cn->r : 0.0)));
    for (int l = 0; l < p; ++l) rt[1 + l] = f * (cj * c.z[l] - al * ((cp ? cp->z[l] : 0.0) + (cn ? cn->z[l] : 0.0)));
  } else if (G.cor == GEE_EXCH) {
    const double kap = al / (1.0 + (double)(n - 1) * al);
    const double f = 1.0 / (1.0 - al);
    rt[0] = f * (c.r - kap * G.ssum[subj][0]);
    for (int l = 0; l < p; ++l) rt[1 + l] = f * (c.z[l] - kap * G.ssum[subj][1 + l]);
  } else {
    rt[0] = c.r;
    for (int l = 0; l < p; ++l) rt[1 + l] = c.z[l];
  }
  o.uvec = c.s * rt[0];
  o.g = c.mu * o.uvec;
  for (int l = 0; l < p; ++l) {
    o.q[l] = c.s * rt[1 + l];
    o.h[l] = c.mu * o.q[l];
    o.xm[l] = c.mu * (W.tal[l][b] + W.tga[l][b] * u);
  }
}

// AR-1 weight of the diagonal term for position q of a subject with n cells
SC_HD double ar1_cj(int q, int n, double alpha) {
  if (n == 1) return 1.0 - alpha * alpha;
  return (q == 0 || q == n - 1) ? 1.0 : 1.0 + alpha * alpha;
}

// ---------------------------------------------------------------------------------------------
// geeM::geem.  On entry: model tables set (coop_set_model), W.use_offset set.  On exit: W.beta = coefficients,
// G.alpha / G.phi, G.hess = hessian of the last inner Newton step (naiv.var = hess^-1), G.converged / G.niter.
// Returns false on a singular system.
// ---------------------------------------------------------------------------------------------
// assemble hess (p x p, row-major in G.hess) and esteq (W.grad) from a GPASS_BETA result.  Leader only.
SC_HD void gee_assemble_beta(const GeeLineage& GL, const Model& m, Work& W, GeeWork& G) {
  const int p = m.n;
  const double al = G.alpha, ph = G.phi;
  if (G.cor == GEE_AR1) {
    const double f = 1.0 / ((1.0 - al * al) * ph);
    for (int j = 0; j < p; ++j) {
      for (int l = 0; l < p; ++l) G.hess[j * GEE_PMAX + l] = f * (G.A1[sym_idx(j, l)] - al * G.A2[sym_idx(j, l)]);
      W.grad[j] = f * (G.e1[j] - al * G.e2[j]);
    }
  } else if (G.cor == GEE_EXCH) {
    const double f = 1.0 / ((1.0 - al) * ph);
    for (int j = 0; j < p; ++j) {
      for (int l = 0; l < p; ++l) {
        double corr = 0.0;
        for (int i = 0; i < GL.S; ++i) {
          const int n = GL.sstart[i + 1] - GL.sstart[i];
          const double kap = al / (1.0 + (double)(n - 1) * al);
          corr += kap * G.ssum[i][j] * G.ssum[i][l];
        }
        G.hess[j * GEE_PMAX + l] = f * (G.A1[sym_idx(j, l)] - corr);
      }
      double corr = 0.0;
      for (int i = 0; i < GL.S; ++i) {
        const int n = GL.sstart[i + 1] - GL.sstart[i];
        const double kap = al / (1.0 + (double)(n - 1) * al);
        corr += kap * G.ssum[i][j] * G.ssum[i][p];
      }
      W.grad[j] = f * (G.e1[j] - corr);
    }
  } else {
    const double f = 1.0 / ph;
    for (int j = 0; j < p; ++j) {
      for (int l = 0; l < p; ++l) G.hess[j * GEE_PMAX + l] = f * G.A1[sym_idx(j, l)];
      W.grad[j] = f * G.e1[j];
    }
  }
}

// plain Gaussian elimination with partial pivoting (R's solve(hess, esteq)); n <= GEE_PMAX.  false if singular.
SC_HDN bool gee_solve(int n, const double* A, const double* b, double* x) {
  double M[GEE_PMAX][GEE_PMAX + 1];
  for (int i = 0; i < n; ++i) { for (int j = 0; j < n; ++j) M[i][j] = A[i * GEE_PMAX + j]; M[i][n] = b[i]; }
  for (int k = 0; k < n; ++k) {
    int piv = k;
    for (int i = k + 1; i < n; ++i) if (fabs(M[i][k]) > fabs(M[piv][k])) piv = i;
    if (!(fabs(M[piv][k]) > 0.0)) return false;
    if (piv != k) for (int j = 0; j <= n; ++j) { const double t = M[k][j]; M[k][j] = M[piv][j]; M[piv][j] = t; }
    for (int i = k + 1; i < n; ++i) {
      const double f = M[i][k] / M[k][k];
      for (int j = k; j <= n; ++j) M[i][j] -= f * M[k][j];
    }
  }
  for (int i = n - 1; i >= 0; --i) {
    double s = M[i][n];
    for (int j = i + 1; j < n; ++j) s -= M[i][j] * x[j];
    x[i] = s / M[i][i];
  }
  for (int i = 0; i < n; ++i) if (!(x[i] == x[i]) || isinf(x[i])) return false;
  return true;
}
// inverse of a general small matrix (solve(hess)); out row-major GEE_PMAX stride.  R's solve() refuses a system whose
// 1-norm reciprocal condition number is below .Machine$double.eps ("computationally singular"); try() turns that into a
// model failure.  rcond is computed exactly here (LAPACK estimates it) -- DESIGN.md section 7.
SC_HDN bool gee_invert(int n, const double* A, double* out) {
  double e[GEE_PMAX], col[GEE_PMAX];
  for (int i = 0; i < n; ++i) for (int j = 0; j < n; ++j) { const double v = A[i * GEE_PMAX + j]; if (!(v == v) || isinf(v)) return false; }
  for (int c = 0; c < n; ++c) {
    for (int i = 0; i < n; ++i) e[i] = i == c ? 1.0 : 0.0;
    if (!gee_solve(n, A, e, col)) return false;
    for (int i = 0; i < n; ++i) out[i * GEE_PMAX + c] = col[i];
  }
  double na = 0.0, ni = 0.0;
  for (int c = 0; c < n; ++c) {
    double sa = 0.0, si = 0.0;
    for (int i = 0; i < n; ++i) { sa += fabs(A[i * GEE_PMAX + c]); si += fabs(out[i * GEE_PMAX + c]); }
    if (sa > na) na = sa;
    if (si > ni) ni = si;
  }
  const double prod = na * ni;
  if (!(prod == prod) || isinf(prod) || 1.0 / prod < 2.220446049250313e-16) return false;
  return true;
}

// Robust (sandwich) variance of geeM::geem(sandwich = TRUE): var = hess^-1 (sum_i e_i e_i') hess^-T with
// e_i = Z_i' (R_i^-1 / phi) rr_i from a GPASS_SAND pass at the final beta and hess the hessian of the last inner step.
// out: row-major p x p (stride GEE_PMAX).  Leader only.  false if hess is computationally singular.
SC_HDN bool gee_sandwich(const GeeLineage& GL, const Model& m, const GeeWork& G, double* out) {
  const int p = m.n;
  const double al = G.alpha, ph = G.phi;
  double num[GEE_PMAX * GEE_PMAX], Hinv[GEE_PMAX * GEE_PMAX], tmp[GEE_PMAX * GEE_PMAX];
  for (int j = 0; j < p * GEE_PMAX; ++j) num[j] = 0.0;
  for (int i = 0; i < GL.S; ++i) {
    double e[GEE_PMAX];
    const int n = GL.sstart[i + 1] - GL.sstart[i];
    for (int j = 0; j < p; ++j) {
      if (G.cor == GEE_AR1) e[j] = (G.ssum[i][j] - al * G.ssum[i][p + j]) / ((1.0 - al * al) * ph);
      else if (G.cor == GEE_EXCH) {
        const double kap = al / (1.0 + (double)(n - 1) * al);
        e[j] = (G.ssum[i][j] - kap * G.ssum[i][p + j] * G.ssum[i][2 * p]) / ((1.0 - al) * ph);
      } else e[j] = G.ssum[i][j] / ph;
    }
    for (int j = 0; j < p; ++j) for (int l = 0; l < p; ++l) num[j * GEE_PMAX + l] += e[j] * e[l];
  }
  if (!gee_invert(p, G.hess, Hinv)) return false;
  for (int j = 0; j < p; ++j) for (int l = 0; l < p; ++l) {
    double s = 0.0;
    for (int q = 0; q < p; ++q) s += Hinv[j * GEE_PMAX + q] * num[q * GEE_PMAX + l];
    tmp[j * GEE_PMAX + l] = s;
  }
  for (int j = 0; j < p; ++j) for (int l = 0; l < p; ++l) {
    double s = 0.0;
    for (int q = 0; q < p; ++q) s += tmp[j * GEE_PMAX + q] * Hinv[l * GEE_PMAX + q];     // x hess^-T
    out[j * GEE_PMAX + l] = s;
  }
  return true;
}

template <class E>
SC_HDN bool geem_fit(E& ex, const Lineage& L, const GeeLineage& GL, const Model& m, const GeneStats& gs, Work& W, GeeWork& G,
                     double mean_offset, bool sandwich = false) {
  const int p = m.n;
  const double N = (double)L.N;
  coop_set_model(ex, L, m, W);
  if (ex.leader()) {
    // init.beta: zero except the intercept column = linkfun(mean(Y)) - mean(offset)
    for (int j = 0; j < p; ++j) W.beta[j] = 0.0;
    W.beta[0] = log(gs.sum_y / N) - mean_offset;
    for (int j = 0; j < p; ++j) G.beta_old[j] = W.beta[j];
    G.phi = 1.0; G.alpha = 0.0;
    G.converged = 0; G.niter = 0;
    W.flag = 0;
    W.gee_p = p;
  }
  ex.sync();
  for (int count = 1; count <= kGeeMaxit; ++count) {
    coop_set_eta(ex, L, m, W, W.beta);
    if (ex.leader()) W.passes++;
    ex.template gee_pass<GPASS_RESID>();
    if (ex.leader()) {
      G.phi = G.sc[0] / (N - (double)p);                                              // updatePhi (useP = TRUE)
      if (G.cor == GEE_AR1) {
        double npairs = 0.0;
        for (int i = 0; i < GL.S; ++i) npairs += (double)(GL.sstart[i + 1] - GL.sstart[i] - 1);
        G.alpha = G.sc[1] / (G.phi * (npairs - (double)p));
      } else if (G.cor == GEE_EXCH) {
        double num = 0.0, npairs = 0.0;
        for (int i = 0; i < GL.S; ++i) {
          const double n = (double)(GL.sstart[i + 1] - GL.sstart[i]);
          num += G.ssum[i][0] * G.ssum[i][0];
          npairs += n * (n - 1.0) / 2.0;
        }
        G.alpha = 0.5 * (num - G.sc[0]) / (G.phi * (npairs - (double)p));
      } else G.alpha = 0.0;
      G.niter = count;
    }
    ex.sync();
    // updateBeta: up to 10 Newton steps
    for (int inner = 0; inner < 10; ++inner) {
      coop_set_eta(ex, L, m, W, W.beta);
      if (ex.leader()) W.passes++;
      ex.template gee_pass<GPASS_BETA>();
      if (ex.leader()) {
        gee_assemble_beta(GL, m, W, G);
        W.flag2 = 0;
        double inv_scratch[GEE_PMAX * GEE_PMAX];
        if (!gee_invert(p, G.hess, inv_scratch) || !gee_solve(p, G.hess, W.grad, W.step)) W.flag = 2;
        else {
          double mx = -INFINITY;
          for (int j = 0; j < p; ++j) {
            W.beta[j] += W.step[j];
            const double ratio = fabs(W.step[j]) / W.beta[j];          // sic: no abs() on beta.new
            if (ratio > mx || !(ratio == ratio)) mx = ratio;            // NaN propagates through max() in R
          }
          if (mx < 100.0 * kGeeTol) W.flag2 = 1;
        }
      }
      ex.sync();
      if (W.flag == 2) return false;
      if (W.flag2) break;
    }
    if (ex.leader()) {
      double mx = -INFINITY;
      for (int j = 0; j < p; ++j) {
        const double d = fabs((W.beta[j] - G.beta_old[j]) / (G.beta_old[j] + 2.220446049250313e-16));
        if (d > mx || !(d == d)) mx = d;
      }
      if (mx < kGeeTol) { G.converged = 1; W.flag = 1; }
      for (int j = 0; j < p; ++j) G.beta_old[j] = W.beta[j];
    }
    ex.sync();
    if (W.flag == 1) break;
  }
  ex.sync();
  if (sandwich) {
    coop_set_eta(ex, L, m, W, W.beta);
    if (ex.leader()) W.passes++;
    ex.template gee_pass<GPASS_SAND>();
  }
  return true;
}

// ---------------------------------------------------------------------------------------------
// GEE score sweep: stat_out_score_gee_null (minus the geem fit) + score_fun_gee for all knots and both shapes.
// On entry: geem_fit done for model m without offset (W.beta, G.alpha, G.phi).  Scores -> W.sw_one / W.sw_both (+ rounded).
// ---------------------------------------------------------------------------------------------
struct GeeCandFn {        // after each subject: fold b_i(k) of every candidate column into the running sums
  const Lineage* L; Work* W; GeeWork* G; int p;
  SC_HD void operator()(int col) const {
    const int T = L->T;
    if (col >= 2 * T) return;
    const int kind = col < T ? SCLANE_TERM_TP1 : SCLANE_TERM_TP2;
    const int k = col < T ? col : col - T;
    double bi = 0.0;
    for (int b = 0; b <= T; ++b) {
      double al, ga;
      term_ab(*L, kind, k, b, al, ga);
      bi += al * G->Gm[b][0] + ga * G->Gm[b][1];
    }
    G->cb[col] += bi;
    G->cS22[col] += bi * bi;
    for (int l = 0; l < p; ++l) G->cS21[col][l] += bi * G->e1[l];      // e1 = a_i of this subject
    if (col < T) {
      // cross term with the tp2 column of the same knot
      double bj = 0.0;
      for (int b = 0; b <= T; ++b) {
        double al, ga;
        term_ab(*L, SCLANE_TERM_TP2, k, b, al, ga);
        bj += al * G->Gm[b][0] + ga * G->Gm[b][1];
      }
      G->cS22x[k] += bi * bj;
    }
  }
};
struct GeeJ21Fn {         // J21 rows from the bucket moments of h_l
  const Lineage* L; GeeWork* G; int p;
  SC_HD void operator()(int col) const {
    const int T = L->T;
    if (col >= 2 * T) return;
    const int kind = col < T ? SCLANE_TERM_TP1 : SCLANE_TERM_TP2;
    const int k = col < T ? col : col - T;
    for (int l = 0; l < p; ++l) {
      double s = 0.0;
      for (int b = 0; b <= T; ++b) {
        double al, ga;
        term_ab(*L, kind, k, b, al, ga);
        s += al * G->Hm[b][2 * l] + ga * G->Hm[b][2 * l + 1];
      }
      G->cJ21[col][l] = s;
    }
  }
};
struct GeeScoreFn {
  const Lineage* L; const Model* m; Work* W; GeeWork* G;
  SC_HD void operator()(int i) const {
    const int T = L->T, p = m->n;
    if (i >= 2 * T) return;
    const bool both = i >= T;
    const int k = both ? i - T : i;
    double score;
    if (cand_collinear(*L, *m, k, both)) score = nan("");
    else {
      const int c = both ? 2 : 1;
      const int cols[2] = {k, T + k};
      double Sg[2][2];
      for (int q = 0; q < c; ++q)
        for (int r = 0; r <= q; ++r) {
          const double s22 = (q == r) ? G->cS22[cols[q]] : G->cS22x[k];
          double t1 = 0.0, t2 = 0.0, t3 = 0.0;
          for (int a = 0; a < p; ++a)
            for (int bq = 0; bq < p; ++bq) {
              t1 += G->cJ21[cols[q]][a] * G->J11inv[a * GEE_PMAX + bq] * G->cS21[cols[r]][bq];   // J21 J11^-1 Sigma21'
              t2 += G->cS21[cols[q]][a] * G->J11inv[a * GEE_PMAX + bq] * G->cJ21[cols[r]][bq];   // Sigma21 J11^-1 J21'
              t3 += G->cJ21[cols[q]][a] * G->JS11[a * GEE_PMAX + bq] * G->cJ21[cols[r]][bq];     // J21 JSigma11 J21'
            }
          Sg[q][r] = s22 - t1 - t2 + t3;
        }
      score = llt_quadform(c, Sg[0][0], c == 2 ? Sg[1][0] : 0.0, c == 2 ? Sg[1][1] : 0.0, G->cb[cols[0]], c == 2 ? G->cb[cols[1]] : 0.0);
    }
    if (both) { W->sw_both[k] = score; W->sw_both_r[k] = r_round(score, 1e6); }
    else { W->sw_one[k] = score; W->sw_one_r[k] = r_round(score, 1e6); }
  }
};

// Eigen LLT inverse semantics for the small J11 (eigenMapMatrixInvert; stat_out_score_gee_null.R:93-100); p == 1 -> 1/J11.
SC_HDN void gee_llt_inverse(int n, const double* A, double* out) {
  double Lm[GEE_PMAX][GEE_PMAX];
  for (int i = 0; i < n; ++i) for (int j = 0; j < n; ++j) Lm[i][j] = A[i * GEE_PMAX + j];
  for (int k = 0; k < n; ++k) {
    double x = Lm[k][k];
    for (int j = 0; j < k; ++j) x -= Lm[k][j] * Lm[k][j];
    if (x <= 0.0) break;                               // Eigen stops here and leaves the rest untouched
    x = sqrt(x);
    Lm[k][k] = x;
    for (int i = k + 1; i < n; ++i) {
      double s = Lm[i][k];
      for (int j = 0; j < k; ++j) s -= Lm[i][j] * Lm[k][j];
      Lm[i][k] = s / x;
    }
  }
  for (int c = 0; c < n; ++c) {
    double y[GEE_PMAX];
    for (int i = 0; i < n; ++i) {
      double s = i == c ? 1.0 : 0.0;
      for (int j = 0; j < i; ++j) s -= Lm[i][j] * y[j];
      y[i] = s / Lm[i][i];
    }
    for (int i = n - 1; i >= 0; --i) {
      double s = y[i];
      for (int j = i + 1; j < n; ++j) s -= Lm[j][i] * y[j];
      y[i] = s / Lm[i][i];
    }
    for (int i = 0; i < n; ++i) out[i * GEE_PMAX + c] = y[i];
  }
}

template <class E>
SC_HDN void gee_sweep(E& ex, const Lineage& L, const GeeLineage& GL, const Model& m, Work& W, GeeWork& G) {
  const int p = m.n, T = L.T;
  // eta at the fitted beta (no offset in the forward pass); G.phi is used as the NB dispersion from here on
  coop_set_eta(ex, L, m, W, W.beta);
  if (ex.leader()) {
    for (int i = 0; i < GEE_NSYM; ++i) { G.A1[i] = 0.0; G.S11[i] = 0.0; }
    for (int c = 0; c < 2 * TMAX; ++c) { G.cb[c] = 0.0; G.cS22[c] = 0.0; for (int l = 0; l < GEE_PMAX; ++l) { G.cS21[c][l] = 0.0; G.cJ21[c][l] = 0.0; } }
    for (int k = 0; k < TMAX; ++k) G.cS22x[k] = 0.0;
    for (int b = 0; b < NBMAX; ++b) for (int l = 0; l < 2 * GEE_PMAX; ++l) G.Hm[b][l] = 0.0;
    W.passes++;
    W.gee_p = p;
  }
  ex.sync();
  if (G.cor == GEE_EXCH) ex.template gee_pass<GPASS_PREP1>();
  for (int i = 0; i < GL.S; ++i) {
    if (ex.leader()) G.subject = i;
    ex.sync();
    ex.template gee_pass<GPASS_PREP2>();          // adds J11 into A1, fills e1 = a_i, Gm (this subject), adds into Hm
    if (ex.leader()) for (int j = 0; j < p; ++j) for (int l = 0; l <= j; ++l) G.S11[sym_idx(j, l)] += G.e1[j] * G.e1[l];
    GeeCandFn cf{&L, &W, &G, p};
    ex.par_for(2 * T, cf);
    ex.sync();
  }
  if (ex.leader()) {
    double J11[GEE_PMAX * GEE_PMAX], S11[GEE_PMAX * GEE_PMAX], tmp[GEE_PMAX * GEE_PMAX];
    for (int j = 0; j < p; ++j) for (int l = 0; l < p; ++l) { J11[j * GEE_PMAX + l] = G.A1[sym_idx(j, l)]; S11[j * GEE_PMAX + l] = G.S11[sym_idx(j, l)]; }
    if (p == 1) G.J11inv[0] = 1.0 / J11[0];
    else gee_llt_inverse(p, J11, G.J11inv);
    for (int j = 0; j < p; ++j) for (int l = 0; l < p; ++l) {
      double s = 0.0;
      for (int q = 0; q < p; ++q) s += G.J11inv[j * GEE_PMAX + q] * S11[q * GEE_PMAX + l];
      tmp[j * GEE_PMAX + l] = s;
    }
    for (int j = 0; j < p; ++j) for (int l = 0; l < p; ++l) {
      double s = 0.0;
      for (int q = 0; q < p; ++q) s += tmp[j * GEE_PMAX + q] * G.J11inv[q * GEE_PMAX + l];
      G.JS11[j * GEE_PMAX + l] = s;
    }
  }
  GeeJ21Fn jf{&L, &G, p};
  ex.par_for(2 * T, jf);
  ex.sync();
  GeeScoreFn sf{&L, &m, &W, &G};
  ex.par_for(2 * T, sf);
  ex.sync();
}

// ---------------------------------------------------------------------------------------------
// result of the GEE path for one gene
// ---------------------------------------------------------------------------------------------
struct GeeResult {
  double coef[GEE_PMAX];
  double cov[GEE_PMAX * GEE_PMAX];     // naiv.var, row-major p x p (stride GEE_PMAX)
  double cov_robust[GEE_PMAX * GEE_PMAX];   // geem $var (sandwich), only when the fit was made with sandwich = TRUE
  double alpha, phi;
  double score_stat;                   // scoreTestGEE statistic (alt only, gee_test = 1); NaN when not computed
  int converged, niter, ok;
};

template <class E>
SC_HDN void gee_store_fit(E& ex, const GeeLineage& GL, const Model& m, Work& W, GeeWork& G, bool sandwich, GeeResult* out) {
  if (ex.leader()) {
    out->ok = gee_invert(m.n, G.hess, out->cov) ? 1 : 0;
    if (sandwich && out->ok) out->ok = gee_sandwich(GL, m, G, out->cov_robust) ? 1 : 0;
    for (int j = 0; j < m.n; ++j) out->coef[j] = W.beta[j];
    out->alpha = G.alpha; out->phi = G.phi; out->converged = G.converged; out->niter = G.niter;
    out->score_stat = nan("");
  }
  ex.sync();
}

// marge2(is.gee = TRUE) + null geem for one gene: forward pass, backward WIC pass, final fits (with offset).
// Needs, per subject-ordered cell: y, bucket id, u, offset.  OLS / stat_out uses unweighted y-moments per bucket, which the
// executor provides in W.acc[b][5..6] (sum y, sum y u) before the call.
template <class E>
SC_HDN void gene_gee(E& ex, const Lineage& L, const GeeLineage& GL, const GeneStats& gstat, GeneState& st, Work& W, GeeWork& G,
                     int M, double tols_score, bool use_offset, double mean_offset, int gee_test, bool sandwich, GeeResult* alt,
                     GeeResult* null) {
  ex.sync();
  if (ex.leader()) { alt->ok = 0; null->ok = 0; st.fin.n = 0; W.use_offset = 0; }
  ex.sync();
  if (st.error) return;
  ex.template gee_pass<GPASS_YMOM>();
  ex.sync();
  // ---- forward pass (marge2.R:109-759)
  for (int it = 0; it < 4 && !st.fwd_done; ++it) {
    if (!geem_fit(ex, L, GL, st.fwd, gstat, W, G, 0.0)) {
      if (ex.leader()) { st.error |= SCLANE_NOTE_NUMERIC; st.fwd_done = 1; }
      ex.sync();
      break;
    }
    if (ex.leader() && st.fwd_iters < 4) st.fwd_sigma[st.fwd_iters] = G.phi;
    gee_sweep(ex, L, GL, st.fwd, W, G);
    if (ex.leader()) forward_select(L, st, W);
    ex.sync();
    if (W.flag) {
      StatOutFn so{&L, &W.bw_cur, &W};
      ex.par_for(W.bw_cur.n * (W.bw_cur.n + 1) / 2 + W.bw_cur.n, so);
      ex.sync();
      if (ex.leader()) forward_accept(L, st, W, gstat, M, tols_score);
      ex.sync();
    }
  }
  if (st.error) return;
  // ---- backward pass (marge2.R:763-812, GEE branch of backward_sel_WIC.R:36-42)
  const int ncol = st.fwd.n;
  if (ncol <= 2) {
    if (ex.leader()) st.error |= SCLANE_NOTE_FWD_EMPTY;
    ex.sync();
  } else {
    if (ex.leader()) { W.bw_cur = st.fwd; W.bw_models[0] = st.fwd; W.bw_cum = 0.0; }
    ex.sync();
    for (int i = 1; i <= ncol - 1; ++i) {
      const bool ok = geem_fit(ex, L, GL, W.bw_cur, gstat, W, G, 0.0, sandwich);
      if (ex.leader()) {
        double wald[PMAX], cov[GEE_PMAX * GEE_PMAX];
        const Model cur = W.bw_cur;
        const int nw = cur.n - 1;
        // summary(fit)$wald.test: beta / se.model, or beta / se.robust for a fit with a sandwich variance
        const bool inv_ok = ok && (sandwich ? gee_sandwich(GL, cur, G, cov) : gee_invert(cur.n, G.hess, cov));
        if (!inv_ok) st.error |= SCLANE_NOTE_NUMERIC;                   // geem / solve() error inside try() -> MARGE model error
        for (int j = 1; j < cur.n; ++j) {
          const double se = sqrt(cov[j * GEE_PMAX + j]);
          wald[j - 1] = inv_ok ? (W.beta[j] / se) * (W.beta[j] / se) : nan("");
        }
        const double mn = nanmin(wald, nw);
        W.bw_cum += mn;
        W.bw_wic[i - 1] = W.bw_cum + 2.0 * (double)cur.n;
        if (i != ncol - 1) {
          int low = -1;
          for (int j = 0; j < nw; ++j) if (wald[j] == mn) { low = j; break; }
          if (low < 0) { st.error |= SCLANE_NOTE_NUMERIC; low = 0; }       // every Wald statistic NA: R fails inside try()
          Model nxt;
          nxt.n = 0;
          for (int j = 0; j < cur.n; ++j) if (j != low + 1) { nxt.kind[nxt.n] = cur.kind[j]; nxt.knot[nxt.n] = cur.knot[j]; nxt.n++; }
          W.bw_cur = nxt;
        } else W.bw_cur.n = 1;
        W.bw_models[i] = W.bw_cur;
      }
      ex.sync();
    }
    if (ex.leader()) {
      int best = -1;
      double bm = INFINITY;
      for (int i = 0; i < ncol - 1; ++i)
        if (W.bw_wic[i] == W.bw_wic[i] && (best < 0 || W.bw_wic[i] < bm)) { bm = W.bw_wic[i]; best = i; }
      st.n_wic = ncol - 1;
      for (int i = 0; i < ncol - 1; ++i) st.wic[i] = W.bw_wic[i];
      if (best < 0) st.error |= SCLANE_NOTE_NUMERIC;
      else {
        const Model sel = W.bw_models[best + 1];
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
  // ---- null model (testDynamic.R:235-244) and final model (marge2.R:831-838), both with the offset
  if (ex.leader()) { W.use_offset = use_offset ? 1 : 0; W.bw_cur.n = 1; W.bw_cur.kind[0] = SCLANE_TERM_INTERCEPT; W.bw_cur.knot[0] = 0; }
  ex.sync();
  if (geem_fit(ex, L, GL, W.bw_cur, gstat, W, G, use_offset ? mean_offset : 0.0, sandwich)) gee_store_fit(ex, GL, W.bw_cur, W, G, sandwich, null);
  if (st.fin.n >= 1 && !(st.error & (SCLANE_NOTE_FWD_EMPTY | SCLANE_NOTE_NUMERIC))) {
    if (st.fin.n == 1) { if (ex.leader()) *alt = *null; ex.sync(); }
    else if (geem_fit(ex, L, GL, st.fin, gstat, W, G, use_offset ? mean_offset : 0.0, sandwich)) gee_store_fit(ex, GL, st.fin, W, G, sandwich, alt);
    ex.sync();
    // scoreTestGEE (scoreTestGEE.R:50-128): with K = diag(mu), V = A^1/2 R A^1/2 at the NULL fit, U = X'K V^-1 r / phi and
    // V_U = X'K V^-1 K X are exactly geem's estimating equation and (phi x) hessian for the alternative design evaluated at
    // beta = (null intercept, 0, ...), alpha = rho_null, phi = phi_null: one GPASS_BETA over the cells.
    if (gee_test == 1 && st.fin.n > 1 && alt->ok && null->ok) {
      const int p = st.fin.n;
      coop_set_model(ex, L, st.fin, W);
      if (ex.leader()) {
        for (int j = 0; j < p; ++j) W.beta[j] = 0.0;
        W.beta[0] = null->coef[0];
        G.alpha = null->alpha; G.phi = null->phi;
        W.gee_p = p;
        W.passes++;
      }
      coop_set_eta(ex, L, st.fin, W, W.beta);
      ex.template gee_pass<GPASS_BETA>();
      if (ex.leader()) {
        gee_assemble_beta(GL, st.fin, W, G);                      // G.hess = V_U / phi, W.grad = U
        double VU[GEE_PMAX * GEE_PMAX], VarM[GEE_PMAX * GEE_PMAX], mid[GEE_PMAX * GEE_PMAX], midinv[GEE_PMAX * GEE_PMAX], v[GEE_PMAX];
        for (int j = 0; j < p; ++j) for (int l = 0; l < p; ++l) VU[j * GEE_PMAX + l] = G.phi * G.hess[j * GEE_PMAX + l];
        gee_llt_inverse(p, VU, VarM);
        for (int j = 0; j < p; ++j) for (int l = 0; l < p; ++l) VarM[j * GEE_PMAX + l] *= G.phi;
        for (int j = 1; j < p; ++j) for (int l = 1; l < p; ++l) mid[(j - 1) * GEE_PMAX + (l - 1)] = VarM[j * GEE_PMAX + l];
        gee_llt_inverse(p - 1, mid, midinv);
        for (int j = 0; j < p; ++j) { double s = 0.0; for (int l = 0; l < p; ++l) s += VarM[j * GEE_PMAX + l] * W.grad[l]; v[j] = s; }
        double S = 0.0;
        for (int j = 1; j < p; ++j) for (int l = 1; l < p; ++l) S += v[j] * midinv[(j - 1) * GEE_PMAX + (l - 1)] * v[l];
        alt->score_stat = S;
      }
      ex.sync();
    }
  }
  ex.sync();
}

}  // namespace sclane
