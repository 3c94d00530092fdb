// sclane_gee_host.hpp -- host-side pieces of the GEE mode: per-lineage subject layout and record finalisation
// (testDynamic.R:107-118 subject ordering check, :296-301 waldTestGEE, pullMARGESummary.R:76-84 GEE branch).
#pragma once
#include "sclane_gee.cuh"
#include "sclane_host.hpp"

namespace sclane {

// Cells of the lineage in the CALLER's order (subject-contiguous), with the knot bucket of each cell.
struct GeePlan {
  GeeLineage GL;
  std::vector<int32_t> cells;     // lineage position -> original cell index
  std::vector<uint8_t> bid;       // [Npad] knot bucket of the cell
  std::vector<double> u;          // [Npad] x - ref[bucket]
  std::vector<double> off;        // [Npad] log(1/size_factor) or empty
  double mean_off = 0.0;
  std::string error;
};

// P must come from build_lineage_plan() on the same pt column (knots, ref, U moments are order independent).
inline bool build_gee_plan(const LineagePlan& P, const double* pt_col, int64_t n_cells, const int32_t* subject_id,
                           const double* size_factor, GeePlan& G) {
  const Lineage& L = P.L;
  G.cells.clear();
  std::vector<double> x;
  for (int64_t i = 0; i < n_cells; ++i)
    if (!std::isnan(pt_col[i])) { G.cells.push_back((int32_t)i); x.push_back(r_round(pt_col[i], 1e4)); }
  const size_t N = x.size();
  G.bid.assign((size_t)L.Npad, 0);
  G.u.assign((size_t)L.Npad, 0.0);
  if (size_factor) G.off.assign((size_t)L.Npad, 0.0); else G.off.clear();
  std::memset(&G.GL, 0, sizeof(G.GL));
  int S = 0;
  double so = 0.0;
  std::vector<int32_t> seen;
  for (size_t i = 0; i < N; ++i) {
    const int32_t sid = subject_id[G.cells[i]];
    if (i == 0 || sid != subject_id[G.cells[i - 1]]) {
      if (std::find(seen.begin(), seen.end(), sid) != seen.end()) { G.error = "data must be ordered by subject (testDynamic.R:115)"; return false; }
      if (S >= GEE_SMAX) { G.error = "more than SCLANE_GEE_SMAX subjects in a lineage"; return false; }
      seen.push_back(sid);
      G.GL.sstart[S++] = (int)i;
    }
    const int b = (int)(std::upper_bound(L.knots, L.knots + L.T, x[i]) - L.knots);
    G.bid[i] = (uint8_t)b;
    G.u[i] = x[i] - L.ref[b];
    if (size_factor) { G.off[i] = std::log(1.0 / size_factor[G.cells[i]]); so += G.off[i]; }
  }
  G.GL.S = S;
  for (int i = S; i <= GEE_SMAX; ++i) G.GL.sstart[i] = (int)N;
  G.mean_off = size_factor ? so / (double)N : 0.0;
  return true;
}

struct GeeGeneOut {
  GeneState st;
  GeeResult alt, null;
};

inline void finalize_record_gee(const GeeGeneOut& g, const Lineage& L, int gee_test, int bias_correction, int n_subjects, sclane_record& r) {
  const double NaN = std::numeric_limits<double>::quiet_NaN();
  GeneOut dummy;
  std::memset(&dummy, 0, sizeof(dummy));
  dummy.st = g.st;
  dummy.st.fin.n = 0;
  finalize_record(dummy, L, r);                 // trace fields + NaN defaults
  const GeneState& st = g.st;
  const bool marge_ok = (st.fin.n >= 1) && g.alt.ok && !(st.error & (SCLANE_NOTE_FWD_EMPTY | SCLANE_NOTE_NUMERIC | SCLANE_NOTE_ALL_ZERO | SCLANE_NOTE_BAD_COUNTS));
  const bool null_ok = g.null.ok && !(st.error & (SCLANE_NOTE_ALL_ZERO | SCLANE_NOTE_BAD_COUNTS));
  r.status = (marge_ok ? SCLANE_MARGE_OK : 0) | (null_ok ? SCLANE_NULL_OK : 0);
  r.marge_notes = st.error | ((marge_ok && !g.alt.converged) ? SCLANE_NOTE_NOT_CONVERGED : 0);
  r.null_notes = null_ok ? (g.null.converged ? 0 : SCLANE_NOTE_NOT_CONVERGED) : (st.error & (SCLANE_NOTE_ALL_ZERO | SCLANE_NOTE_BAD_COUNTS));
  const bool sandwich = bias_correction != 0;          // testDynamic.R:195,243,286: sandwich.var = !is.null(gee.bias.correction.method)
  if (marge_ok) {
    const int p = st.fin.n;
    const double* V = sandwich ? g.alt.cov_robust : g.alt.cov;      // pullMARGESummary.R:33-37,46: $var / se.robust, else naive
    r.n_terms = p;
    for (int j = 0; j < p; ++j) {
      r.term_kind[j] = st.fin.kind[j];
      r.term_knot_idx[j] = st.fin.kind[j] == SCLANE_TERM_INTERCEPT ? -1 : st.fin.knot[j];
      if (st.fin.kind[j] != SCLANE_TERM_INTERCEPT) {
        r.term_knot[j] = L.knots[st.fin.knot[j]];
        r.term_label[j] = r_signif(r.term_knot[j], 4);
      }
      r.coef[j] = g.alt.coef[j];
      r.se[j] = std::sqrt(V[j * GEE_PMAX + j]);
      r.z[j] = r.coef[j] / r.se[j];
      r.p[j] = r_round(pnorm_two_sided(r.z[j]), 1e8);         // summary.geem() rounds its p-values to 8 digits
      for (int l = 0; l < p; ++l) r.cov[j * p + l] = V[j * GEE_PMAX + l];
    }
    r.gee_alpha = g.alt.alpha; r.gee_phi = g.alt.phi;
    r.gee_converged = g.alt.converged; r.gee_iters = g.alt.niter;
    // waldTestGEE(): W = b' (L V L')^-1 b over the non-intercept coefficients, naive variance.  Both tests return NA for
    // Wald_Stat / DF / P_Val when EITHER model is a try-error (waldTestGEE.R:31-37, scoreTestGEE.R:34-39)
    if (!null_ok) { /* test_stat, df, p_val stay NaN */ }
    else if (p <= 1) { r.test_stat = 0.0; r.df = 0.0; r.p_val = 1.0; }
    else if (gee_test == 1) {
      // scoreTestGEE(): statistic from the device; NA when the null model failed (scoreTestGEE.R:34-39)
      r.test_stat = g.alt.score_stat; r.df = (double)(p - 1); r.p_val = 1.0 - (1.0 - pchisq_upper(r.test_stat, r.df));
    } else {
      // biasCorrectGEE.R: "df" n_s / (n_s - p) (unchanged when p >= n_s); "kc" (n_s / (n_s - 1)) / (1 - tr(H) / n) with
      // tr(H) = p (H is a rank-p projection); both scale the sandwich variance
      double factor = 1.0;
      if (bias_correction == 2 && p < n_subjects) factor = (double)n_subjects / (double)(n_subjects - p);
      if (bias_correction == 1) factor = ((double)n_subjects / ((double)n_subjects - 1.0)) / (1.0 - (double)p / (double)L.N);
      double mid[GEE_PMAX * GEE_PMAX], inv[GEE_PMAX * GEE_PMAX];
      for (int j = 1; j < p; ++j) for (int l = 1; l < p; ++l) mid[(j - 1) * GEE_PMAX + (l - 1)] = factor * V[j * GEE_PMAX + l];
      gee_llt_inverse(p - 1, mid, inv);
      double w = 0.0;
      for (int j = 1; j < p; ++j) for (int l = 1; l < p; ++l) w += g.alt.coef[j] * inv[(j - 1) * GEE_PMAX + (l - 1)] * g.alt.coef[l];
      r.test_stat = w;
      r.df = (double)(p - 1);
      r.p_val = 1.0 - (1.0 - pchisq_upper(w, r.df));      // 1 - pchisq(W, df) as waldTestGEE.R:78 computes it
    }
  }
  if (null_ok) {
    r.null_coef = g.null.coef[0];
    r.null_se = std::sqrt(sandwich ? g.null.cov_robust[0] : g.null.cov[0]);
    r.null_z = r.null_coef / r.null_se;
    r.null_p = r_round(pnorm_two_sided(r.null_z), 1e8);
    r.null_gee_alpha = g.null.alpha; r.null_gee_phi = g.null.phi;
    r.null_gee_converged = g.null.converged; r.null_gee_iters = g.null.niter;
  }
  if (!marge_ok) { r.test_stat = r.df = r.p_val = NaN; }
}

}  // namespace sclane
