/* sclane_b200_shim.c -- .Call glue between R and libsclane_b200.so, and the C half of the result assembly.
 *
 * R is not available in the build image, so this file is never compiled against a real R here; it IS compiled and run
 * against the minimal stand-in for R's C API in tests/c/rstub/ (tests/test_r_shim.py): the conformance program calls
 * C_sclane_test_dynamic() below on the reference's bundled data, dumps the returned SEXP tree as JSON and the test diffs
 * the 21-field structure (testDynamic.R:320-342) against the reference's own bundled output.  Only functions of R's
 * public C API are used (Rinternals.h), so the same source builds inside the scLANE package:
 *   PKG_CPPFLAGS = -I<prefix>/include      PKG_LIBS = -L<prefix>/lib -lsclane_b200
 *
 * It follows the reference's own native boundary pattern (src/RcppExports.cpp:14-63: one `SEXP f(SEXP...)` per entry
 * point, registered with R_registerRoutines, errors raised with Rf_error) but needs no Rcpp.
 */
#include <R.h>
#include <Rinternals.h>
#include <R_ext/Rdynload.h>
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "sclane_b200.h"

static const char* LETTERS_ = "ABCDEFGHIJKLMNOPQRSTUVWXYZ";

/* ---- small helpers ------------------------------------------------------------------------------------------- */
static double na_if_nan(double x) { return isnan(x) ? NA_REAL : x; }

static SEXP mk_names(int n, const char** names) {
  SEXP s = PROTECT(allocVector(STRSXP, n));
  for (int i = 0; i < n; ++i) SET_STRING_ELT(s, i, mkChar(names[i]));
  UNPROTECT(1);
  return s;
}

/* a data.frame from already built columns: names, class, compact row names c(NA_integer_, -nrow) */
static SEXP mk_data_frame(SEXP cols, int ncol, const char** names, int nrow) {
  SEXP nm = PROTECT(mk_names(ncol, names));
  setAttrib(cols, R_NamesSymbol, nm);
  SEXP rn = PROTECT(allocVector(INTSXP, 2));
  INTEGER(rn)[0] = NA_INTEGER;
  INTEGER(rn)[1] = -nrow;
  setAttrib(cols, R_RowNamesSymbol, rn);
  SEXP cl = PROTECT(mkString("data.frame"));
  setAttrib(cols, R_ClassSymbol, cl);
  UNPROTECT(3);
  return cols;
}

static SEXP rep_string(const char* s, int n) {
  SEXP v = PROTECT(allocVector(STRSXP, n));
  for (int i = 0; i < n; ++i) SET_STRING_ELT(v, i, s ? mkChar(s) : NA_STRING);
  UNPROTECT(1);
  return v;
}

static SEXP rep_real(double x, int n) {
  SEXP v = PROTECT(allocVector(REALSXP, n));
  for (int i = 0; i < n; ++i) REAL(v)[i] = x;
  UNPROTECT(1);
  return v;
}

/* as.character(<double>) as paste() applies it to signif(knot, 4) (marge2.R:593-594): 15 significant digits */
static void fmt_num(double x, char* buf, size_t n) { snprintf(buf, n, "%.15g", x); }

/* coefficient names: "Intercept", "h_<Lineage>_<knot>" (tp1) / "h_<knot>_<Lineage>" (tp2) -- marge2.R:593-594 + :817-821 */
static void term_name(const sclane_record* r, int j, char letter, char* buf, size_t n) {
  if (j == 0) { snprintf(buf, n, "Intercept"); return; }
  char lab[64];
  fmt_num(r->term_label[j], lab, sizeof(lab));
  if (r->term_kind[j] == SCLANE_TERM_TP1) snprintf(buf, n, "h_Lineage_%c_%s", letter, lab);
  else snprintf(buf, n, "h_%s_Lineage_%c", lab, letter);
}

static const char* fit_note(int bits, int ok) {
  if (ok) return NULL;
  if (bits & SCLANE_NOTE_FWD_EMPTY) return "Error in wic_mat_2[...] : subscript out of bounds (forward pass added fewer than two basis functions)";
  if (bits & SCLANE_NOTE_ALL_ZERO) return "all counts are zero";
  if (bits & SCLANE_NOTE_BAD_COUNTS) return "negative values not allowed for the 'Negative Binomial' family";
  if (bits & SCLANE_NOTE_NUMERIC) return "missing value where TRUE/FALSE needed";
  return "model fit failed";
}

/* ---- MARGE_Summary / Null_Summary (pullMARGESummary.R:70-98, pullNullSummary.R:66-80): term, estimate, std.error,
 *      statistic, p.value ------------------------------------------------------------------------------------------- */
static SEXP coef_table(const sclane_record* r, char letter, int null_model) {
  static const char* cn[5] = {"term", "estimate", "std.error", "statistic", "p.value"};
  const int n = null_model ? 1 : r->n_terms;
  SEXP df = PROTECT(allocVector(VECSXP, 5));
  SEXP term = PROTECT(allocVector(STRSXP, n));
  SEXP est = PROTECT(allocVector(REALSXP, n)), se = PROTECT(allocVector(REALSXP, n));
  SEXP z = PROTECT(allocVector(REALSXP, n)), p = PROTECT(allocVector(REALSXP, n));
  for (int j = 0; j < n; ++j) {
    char nm[96];
    if (null_model) snprintf(nm, sizeof(nm), "Intercept"); else term_name(r, j, letter, nm, sizeof(nm));
    SET_STRING_ELT(term, j, mkChar(nm));
    REAL(est)[j] = null_model ? r->null_coef : r->coef[j];
    REAL(se)[j] = null_model ? r->null_se : r->se[j];
    REAL(z)[j] = null_model ? r->null_z : r->z[j];
    REAL(p)[j] = null_model ? r->null_p : r->p[j];
  }
  SET_VECTOR_ELT(df, 0, term); SET_VECTOR_ELT(df, 1, est); SET_VECTOR_ELT(df, 2, se); SET_VECTOR_ELT(df, 3, z); SET_VECTOR_ELT(df, 4, p);
  mk_data_frame(df, 5, cn, n);
  UNPROTECT(6);
  return df;
}

/* ---- MARGE_Slope_Data (createSlopeTestData.R:15-74 + Gene / Lineage columns of testDynamic.R:295-298) ------------------ */
static SEXP slope_data(const sclane_record* r, const double* pt_lin, int64_t n_lin_cells, const char* gene, char letter) {
  static const char* cn[9] = {"Gene", "Lineage", "Breakpoint", "Rounded_Breakpoint", "Direction", "Beta", "Test_Stat", "P_Val", "Notes"};
  const int marge_ok = (r->status & SCLANE_MARGE_OK) != 0;
  const int k = (marge_ok && r->n_terms > 1) ? r->n_terms - 1 : 1;
  char lin[2] = {letter, 0};
  SEXP df = PROTECT(allocVector(VECSXP, 9));
  SET_VECTOR_ELT(df, 0, rep_string(gene, k));
  SET_VECTOR_ELT(df, 1, rep_string(lin, k));
  if (!marge_ok || r->n_terms == 1) {
    for (int c = 2; c < 8; ++c) SET_VECTOR_ELT(df, c, c == 4 ? rep_string(NULL, 1) : rep_real(NA_REAL, 1));
    SET_VECTOR_ELT(df, 8, rep_string(marge_ok ? "No non-intercept coefficients" : "MARGE model error", 1));
  } else {
    SEXP brk = PROTECT(allocVector(REALSXP, k)), rb = PROTECT(allocVector(REALSXP, k)), dir = PROTECT(allocVector(STRSXP, k));
    SEXP beta = PROTECT(allocVector(REALSXP, k)), ts = PROTECT(allocVector(REALSXP, k)), pv = PROTECT(allocVector(REALSXP, k));
    for (int j = 1; j < r->n_terms; ++j) {
      const double lab = r->term_label[j];
      /* Breakpoint = the pseudotime value nearest to the rounded breakpoint (createSlopeTestData.R:46: which.min(abs(pt - x))) */
      int64_t best = 0;
      double bd = INFINITY;
      for (int64_t i = 0; i < n_lin_cells; ++i) { const double d = fabs(pt_lin[i] - lab); if (d < bd) { bd = d; best = i; } }
      REAL(brk)[j - 1] = pt_lin[best];
      REAL(rb)[j - 1] = lab;
      SET_STRING_ELT(dir, j - 1, mkChar(r->term_kind[j] == SCLANE_TERM_TP1 ? "Right" : "Left"));
      REAL(beta)[j - 1] = r->coef[j];
      REAL(ts)[j - 1] = r->z[j];
      REAL(pv)[j - 1] = r->p[j];
    }
    SET_VECTOR_ELT(df, 2, brk); SET_VECTOR_ELT(df, 3, rb); SET_VECTOR_ELT(df, 4, dir);
    SET_VECTOR_ELT(df, 5, beta); SET_VECTOR_ELT(df, 6, ts); SET_VECTOR_ELT(df, 7, pv);
    SET_VECTOR_ELT(df, 8, rep_string(NULL, k));
    UNPROTECT(6);
  }
  mk_data_frame(df, 9, cn, k);
  UNPROTECT(1);
  return df;
}

/* ---- Gene_Dynamics (summarizeModel.R:28-140, GLM branch, incl. its duplicate-breakpoint indexing; the wide one-row
 *      data.frame of testDynamic.R:313-316): Gene, Lineage, Breakpoint<i>, Slope.Segment<i>, Trend.Segment<i> ---------------- */
typedef struct { double bp; int right; double coef; } hinge_row;
static int hinge_cmp(const void* a, const void* b) {
  const hinge_row* x = (const hinge_row*)a; const hinge_row* y = (const hinge_row*)b;
  if (x->bp != y->bp) return x->bp < y->bp ? -1 : 1;
  return x->right - y->right;                       /* "Left" < "Right" in R's sort of the Direction strings */
}
static SEXP gene_dynamics(const sclane_record* r, double pt_min, double pt_max, const char* gene, char letter) {
  const int marge_ok = (r->status & SCLANE_MARGE_OK) != 0;
  char lin[2] = {letter, 0};
  double bps[SCLANE_PMAX], slopes[SCLANE_PMAX + 1], trends[SCLANE_PMAX + 1];
  int n_bp = 0, n_seg = 0;
  if (!marge_ok) { n_bp = 1; n_seg = 1; bps[0] = NA_REAL; slopes[0] = NA_REAL; trends[0] = NA_REAL; }
  else if (r->n_terms == 1) { n_bp = 1; n_seg = 1; bps[0] = NA_REAL; slopes[0] = 0.0; trends[0] = 0.0; }
  else {
    hinge_row rows[SCLANE_PMAX];
    const int k = r->n_terms - 1;
    for (int j = 1; j < r->n_terms; ++j) { rows[j - 1].bp = r->term_label[j]; rows[j - 1].right = r->term_kind[j] == SCLANE_TERM_TP1; rows[j - 1].coef = r->coef[j]; }
    qsort(rows, (size_t)k, sizeof(hinge_row), hinge_cmp);
    double lo[SCLANE_PMAX], hi[SCLANE_PMAX];
    for (int j = 0; j < k; ++j) { lo[j] = rows[j].right ? rows[j].bp : pt_min; hi[j] = rows[j].right ? pt_max : rows[j].bp; }
    for (int j = 0; j < k; ++j) { int dup = 0; for (int l = 0; l < n_bp; ++l) if (bps[l] == rows[j].bp) dup = 1; if (!dup) bps[n_bp++] = rows[j].bp; }
    n_seg = n_bp + 1;
    for (int i = 1; i <= n_seg; ++i) {
      /* summarizeModel.R:92-104 indexes the SORTED (not the unique) breakpoints: bp(i) = rows[i - 1].bp, NA beyond the end */
      const double b_prev = (i - 1 >= 1 && i - 1 <= k) ? rows[i - 2].bp : NAN;
      const double b_cur = (i >= 1 && i <= k) ? rows[i - 1].bp : NAN;
      const double s_lo = i == 1 ? pt_min : b_prev, s_hi = i == 1 ? b_cur : (i == n_seg ? pt_max : b_cur);
      double s = 0.0;
      for (int j = 0; j < k; ++j) if (s_lo < hi[j] && s_hi > lo[j]) s += rows[j].right ? rows[j].coef : -rows[j].coef;
      slopes[i - 1] = s;
      trends[i - 1] = s > 0.0 ? 1.0 : (s < 0.0 ? -1.0 : 0.0);
    }
  }
  const int ncol = 2 + n_bp + 2 * n_seg;
  SEXP df = PROTECT(allocVector(VECSXP, ncol));
  SEXP nm = PROTECT(allocVector(STRSXP, ncol));
  int c = 0;
  char buf[48];
  SET_VECTOR_ELT(df, c, rep_string(gene, 1)); SET_STRING_ELT(nm, c++, mkChar("Gene"));
  SET_VECTOR_ELT(df, c, rep_string(lin, 1)); SET_STRING_ELT(nm, c++, mkChar("Lineage"));
  for (int i = 0; i < n_bp; ++i) {
    if (n_bp == 1) snprintf(buf, sizeof(buf), "Breakpoint"); else snprintf(buf, sizeof(buf), "Breakpoint%d", i + 1);
    SET_VECTOR_ELT(df, c, rep_real(bps[i], 1)); SET_STRING_ELT(nm, c++, mkChar(buf));
  }
  for (int i = 0; i < n_seg; ++i) {
    if (n_seg == 1) snprintf(buf, sizeof(buf), "Slope.Segment"); else snprintf(buf, sizeof(buf), "Slope.Segment%d", i + 1);
    SET_VECTOR_ELT(df, c, rep_real(slopes[i], 1)); SET_STRING_ELT(nm, c++, mkChar(buf));
  }
  for (int i = 0; i < n_seg; ++i) {
    if (n_seg == 1) snprintf(buf, sizeof(buf), "Trend.Segment"); else snprintf(buf, sizeof(buf), "Trend.Segment%d", i + 1);
    SET_VECTOR_ELT(df, c, rep_real(trends[i], 1)); SET_STRING_ELT(nm, c++, mkChar(buf));
  }
  setAttrib(df, R_NamesSymbol, nm);
  SEXP rn = PROTECT(allocVector(INTSXP, 2));
  INTEGER(rn)[0] = NA_INTEGER; INTEGER(rn)[1] = -1;
  setAttrib(df, R_RowNamesSymbol, rn);
  SEXP cl = PROTECT(mkString("data.frame"));
  setAttrib(df, R_ClassSymbol, cl);
  UNPROTECT(4);
  return df;
}

/* ---- one (gene, lineage) record -> the 21-field list of testDynamic.R:320-342 --------------------------------- */
static SEXP record_to_list(const sclane_record* r, const char* gene, int lin_idx, const sclane_lineage_info* info,
                           const double* pt_lin, int64_t n_lin_cells, int is_gee, int gee_test) {
  static const char* fn[21] = {"Gene", "Lineage", "Test_Stat", "Test_Stat_Type", "Test_Stat_Note", "Degrees_Freedom", "P_Val",
                               "LogLik_MARGE", "LogLik_Null", "Dev_MARGE", "Dev_Null", "Model_Status", "MARGE_Fit_Notes", "Null_Fit_Notes",
                               "Gene_Time", "MARGE_Summary", "Null_Summary", "MARGE_Preds", "Null_Preds", "MARGE_Slope_Data", "Gene_Dynamics"};
  const char letter = LETTERS_[lin_idx % 26];
  const int marge_ok = (r->status & SCLANE_MARGE_OK) != 0, null_ok = (r->status & SCLANE_NULL_OK) != 0;
  char lin[2] = {letter, 0}, status[64];
  snprintf(status, sizeof(status), "MARGE model %s, null model %s", marge_ok ? "OK" : "error", null_ok ? "OK" : "error");
  SEXP out = PROTECT(allocVector(VECSXP, 21));
  SET_VECTOR_ELT(out, 0, mkString(gene));
  SET_VECTOR_ELT(out, 1, mkString(lin));
  SET_VECTOR_ELT(out, 2, ScalarReal(na_if_nan(r->test_stat)));
  SET_VECTOR_ELT(out, 3, mkString(is_gee ? (gee_test == 1 ? "Score" : "Wald") : "LRT"));
  SET_VECTOR_ELT(out, 4, (marge_ok && null_ok) ? ScalarString(NA_STRING) : mkString("No test performed due to model failure."));
  SET_VECTOR_ELT(out, 5, ScalarReal(na_if_nan(r->df)));
  SET_VECTOR_ELT(out, 6, ScalarReal(na_if_nan(r->p_val)));
  SET_VECTOR_ELT(out, 7, ScalarReal(na_if_nan(r->loglik)));
  SET_VECTOR_ELT(out, 8, ScalarReal(na_if_nan(r->null_loglik)));
  SET_VECTOR_ELT(out, 9, ScalarReal(na_if_nan(r->deviance)));
  SET_VECTOR_ELT(out, 10, ScalarReal(na_if_nan(r->null_deviance)));
  SET_VECTOR_ELT(out, 11, mkString(status));
  {
    const char* n1 = fit_note(r->marge_notes, marge_ok);
    const char* n0 = fit_note(r->null_notes, null_ok);
    SET_VECTOR_ELT(out, 12, n1 ? mkString(n1) : ScalarString(NA_STRING));
    SET_VECTOR_ELT(out, 13, n0 ? mkString(n0) : ScalarString(NA_STRING));
  }
  SET_VECTOR_ELT(out, 14, ScalarReal(r->gene_time));
  SET_VECTOR_ELT(out, 15, marge_ok ? coef_table(r, letter, 0) : R_NilValue);
  SET_VECTOR_ELT(out, 16, null_ok ? coef_table(r, letter, 1) : R_NilValue);
  /* MARGE_Preds / Null_Preds (N x 2 data.frames per gene and lineage: 128 GB at 20,000 x 200,000) are produced on demand by
   * C_sclane_link_preds() from the record kept in attr(, "sclane_record"); the R wrapper installs them lazily */
  SET_VECTOR_ELT(out, 17, R_NilValue);
  SET_VECTOR_ELT(out, 18, R_NilValue);
  SET_VECTOR_ELT(out, 19, slope_data(r, pt_lin, n_lin_cells, gene, letter));
  SET_VECTOR_ELT(out, 20, gene_dynamics(r, info->pt_min, info->pt_max, gene, letter));
  SEXP nm = PROTECT(mk_names(21, fn));
  setAttrib(out, R_NamesSymbol, nm);
  SEXP raw = PROTECT(allocVector(RAWSXP, (R_xlen_t)sizeof(sclane_record)));
  memcpy(RAW(raw), r, sizeof(sclane_record));
  setAttrib(out, install("sclane_record"), raw);
  UNPROTECT(3);
  return out;
}

/* opts: integer vector c(M, approx_knot, n_gpus, cor_structure [0 independence, 1 exchangeable, 2 ar1], gee_test [0 wald, 1 score],
 *       bias_correction [0 NULL, 1 kc, 2 df], n_knot_max, minspan [-1 = NULL]); shorter vectors leave the defaults */
static void fill_opts(sclane_opts* o, SEXP opts_) {
  sclane_default_opts(o);
  const int n = opts_ == R_NilValue ? 0 : LENGTH(opts_);
  const int* v = n ? INTEGER(opts_) : NULL;
  if (n >= 1) o->M = v[0];
  if (n >= 2) o->approx_knot = v[1];
  if (n >= 3) { o->n_gpus = v[2]; for (int i = 0; i < o->n_gpus && i < SCLANE_MAX_GPUS; ++i) o->gpu_ids[i] = i; }
  if (n >= 4) o->gee_cor = v[3];
  if (n >= 5) o->gee_test = v[4];
  if (n >= 6) o->gee_bias_correction = v[5];
  if (n >= 7) o->n_knot_max = v[6];
  if (n >= 8) o->minspan = v[7];
}

/* records + lineage infos -> list (one element per gene, named) of lists (one per lineage, "Lineage_A"...) of 21-field lists,
 * class "scLANE" (testDynamic.R:373,387,446) */
static SEXP assemble(const sclane_record* recs, const sclane_lineage_info* infos, SEXP genes, int64_t n_genes, int n_lin,
                     const double* pt, int64_t n_cells, int is_gee, int gee_test) {
  SEXP res = PROTECT(allocVector(VECSXP, (R_xlen_t)n_genes));
  double* pt_lin = (double*)R_alloc((size_t)n_cells, sizeof(double));
  SEXP lin_names = PROTECT(allocVector(STRSXP, n_lin));
  for (int l = 0; l < n_lin; ++l) { char b[16]; snprintf(b, sizeof(b), "Lineage_%c", LETTERS_[l % 26]); SET_STRING_ELT(lin_names, l, mkChar(b)); }
  for (int l = 0; l < n_lin; ++l) {
    int64_t m = 0;
    for (int64_t i = 0; i < n_cells; ++i) { const double v = pt[(size_t)l * (size_t)n_cells + (size_t)i]; if (!isnan(v)) pt_lin[m++] = v; }
    for (int64_t g = 0; g < n_genes; ++g) {
      SEXP per_l;
      if (l == 0) { per_l = PROTECT(allocVector(VECSXP, n_lin)); setAttrib(per_l, R_NamesSymbol, lin_names); SET_VECTOR_ELT(res, (R_xlen_t)g, per_l); UNPROTECT(1); }
      per_l = VECTOR_ELT(res, (R_xlen_t)g);
      SET_VECTOR_ELT(per_l, l, record_to_list(&recs[g * n_lin + l], CHAR(STRING_ELT(genes, (R_xlen_t)g)), l, &infos[l], pt_lin, m, is_gee, gee_test));
    }
  }
  setAttrib(res, R_NamesSymbol, genes);
  SEXP cl = PROTECT(mkString("scLANE"));
  setAttrib(res, R_ClassSymbol, cl);
  UNPROTECT(3);
  return res;
}

static void check_common(SEXP pt, SEXP sf, SEXP genes, int64_t n_cells, int64_t n_genes) {
  if (!isReal(pt) || !isMatrix(pt)) Rf_error("pt must be a double matrix (cells x lineages)");
  if (nrows(pt) != n_cells) Rf_error("pt must have one row per cell");
  if (sf != R_NilValue && (!isReal(sf) || XLENGTH(sf) != n_cells)) Rf_error("size.factor.offset must have one entry per cell");
  if (!isString(genes) || XLENGTH(genes) != n_genes) Rf_error("genes must name every column of expr.mat");
}

/* GLM mode.  counts: integer matrix cells x genes (column-major => gene-major, the bytes of the FBM at testDynamic.R:145-147);
 * pt: double matrix cells x lineages, NA_real_ (a NaN) = cell not in lineage (:183); sf: size factors or NULL (:228-231). */
SEXP C_sclane_test_dynamic(SEXP counts, SEXP pt, SEXP sf, SEXP genes, SEXP opts_) {
  if (!isInteger(counts) || !isMatrix(counts)) Rf_error("expr.mat must be an integer matrix (cells x genes)");
  const int64_t n_cells = nrows(counts), n_genes = ncols(counts);
  check_common(pt, sf, genes, n_cells, n_genes);
  const int n_lin = ncols(pt);
  sclane_opts o;
  fill_opts(&o, opts_);
  sclane_record* recs = (sclane_record*)R_alloc((size_t)(n_genes * n_lin), sizeof(sclane_record));
  sclane_lineage_info* infos = (sclane_lineage_info*)R_alloc((size_t)n_lin, sizeof(sclane_lineage_info));
  char err[1024]; err[0] = 0;
  /* NA_integer_ never occurs in counts; NA_real_ in pt is a NaN and is tested with isnan() by the library */
  const int rc = sclane_test_dynamic(INTEGER(counts), n_cells, n_genes, REAL(pt), n_lin, sf == R_NilValue ? NULL : REAL(sf), &o, recs, infos, err, sizeof(err));
  if (rc != SCLANE_OK) Rf_error("sclane_b200: %s", err);
  return assemble(recs, infos, genes, n_genes, n_lin, REAL(pt), n_cells, 0, 0);
}

/* Sparse input: the genes x cells dgCMatrix of a SingleCellExperiment / Seurat object (its @p, @i, @x slots and dims); replaces
 * the densify + transpose of testDynamic.R:80-98. */
SEXP C_sclane_test_dynamic_csc(SEXP p_, SEXP i_, SEXP x_, SEXP dims, SEXP pt, SEXP sf, SEXP genes, SEXP opts_) {
  if (!isInteger(p_) || !isInteger(i_) || !isReal(x_) || !isInteger(dims) || LENGTH(dims) != 2) Rf_error("expected the @p, @i, @x and @Dim slots of a dgCMatrix");
  const int64_t n_genes = INTEGER(dims)[0], n_cells = INTEGER(dims)[1];
  if (XLENGTH(p_) != n_cells + 1) Rf_error("@p must have ncol + 1 entries");
  check_common(pt, sf, genes, n_cells, n_genes);
  const int n_lin = ncols(pt);
  sclane_opts o;
  fill_opts(&o, opts_);
  sclane_record* recs = (sclane_record*)R_alloc((size_t)(n_genes * n_lin), sizeof(sclane_record));
  sclane_lineage_info* infos = (sclane_lineage_info*)R_alloc((size_t)n_lin, sizeof(sclane_lineage_info));
  char err[1024]; err[0] = 0;
  const int rc = sclane_test_dynamic_csc(INTEGER(p_), INTEGER(i_), REAL(x_), n_cells, n_genes, REAL(pt), n_lin, sf == R_NilValue ? NULL : REAL(sf), &o, recs, infos, err, sizeof(err));
  if (rc != SCLANE_OK) Rf_error("sclane_b200: %s", err);
  return assemble(recs, infos, genes, n_genes, n_lin, REAL(pt), n_cells, 0, 0);
}

/* GEE mode (testDynamic(is.gee = TRUE), testDynamic.R:66-70,235-244,296-301).  ids: as.integer(factor(id.vec, levels = unique(id.vec))),
 * cells ordered by subject (:115). */
SEXP C_sclane_test_dynamic_gee(SEXP counts, SEXP pt, SEXP ids, SEXP sf, SEXP genes, SEXP opts_) {
  if (!isInteger(counts) || !isMatrix(counts)) Rf_error("expr.mat must be an integer matrix (cells x genes)");
  const int64_t n_cells = nrows(counts), n_genes = ncols(counts);
  check_common(pt, sf, genes, n_cells, n_genes);
  if (!isInteger(ids) || XLENGTH(ids) != n_cells) Rf_error("id.vec must be an integer vector with one entry per cell");
  const int n_lin = ncols(pt);
  sclane_opts o;
  fill_opts(&o, opts_);
  o.mode = 1;
  sclane_record* recs = (sclane_record*)R_alloc((size_t)(n_genes * n_lin), sizeof(sclane_record));
  sclane_lineage_info* infos = (sclane_lineage_info*)R_alloc((size_t)n_lin, sizeof(sclane_lineage_info));
  char err[1024]; err[0] = 0;
  const int rc = sclane_test_dynamic_gee(INTEGER(counts), n_cells, n_genes, REAL(pt), n_lin, INTEGER(ids), sf == R_NilValue ? NULL : REAL(sf), &o, recs, infos,
                                         err, sizeof(err));
  if (rc != SCLANE_OK) Rf_error("sclane_b200: %s", err);
  return assemble(recs, infos, genes, n_genes, n_lin, REAL(pt), n_cells, 1, o.gee_test);
}

/* marge2(is.glmm = TRUE) for every (gene, pt column): the selected basis only (marge2.R:871-879; fitGLMM.R:54-64 passes one pt
 * column per subject).  Returns a list (per gene) of lists (per column) of character vectors: the coef_names. */
SEXP C_sclane_marge_basis(SEXP counts, SEXP pt, SEXP opts_) {
  if (!isInteger(counts) || !isMatrix(counts)) Rf_error("Y must be an integer matrix (cells x genes)");
  if (!isReal(pt) || !isMatrix(pt) || nrows(pt) != nrows(counts)) Rf_error("pt must be a double matrix with one row per cell");
  const int64_t n_cells = nrows(counts), n_genes = ncols(counts);
  const int n_lin = ncols(pt);
  sclane_opts o;
  fill_opts(&o, opts_);
  sclane_record* recs = (sclane_record*)R_alloc((size_t)(n_genes * n_lin), sizeof(sclane_record));
  char err[1024]; err[0] = 0;
  if (sclane_marge_basis(INTEGER(counts), n_cells, n_genes, REAL(pt), n_lin, &o, recs, NULL, err, sizeof(err))) Rf_error("sclane_b200: %s", err);
  SEXP res = PROTECT(allocVector(VECSXP, (R_xlen_t)n_genes));
  for (int64_t g = 0; g < n_genes; ++g) {
    SEXP per = PROTECT(allocVector(VECSXP, n_lin));
    for (int l = 0; l < n_lin; ++l) {
      const sclane_record* r = &recs[g * n_lin + l];
      if (!(r->status & SCLANE_MARGE_OK)) { SET_VECTOR_ELT(per, l, R_NilValue); continue; }
      SEXP nm = PROTECT(allocVector(STRSXP, r->n_terms));
      for (int j = 0; j < r->n_terms; ++j) { char b[96]; term_name(r, j, 'A', b, sizeof(b)); SET_STRING_ELT(nm, j, mkChar(b)); }
      SET_VECTOR_ELT(per, l, nm);
      UNPROTECT(1);
    }
    SET_VECTOR_ELT(res, (R_xlen_t)g, per);
    UNPROTECT(1);
  }
  UNPROTECT(1);
  return res;
}

/* MARGE_Preds / Null_Preds of one record on demand (pullMARGESummary.R:66-69, pullNullSummary.R:69-72): rec = the raw vector in
 * attr(<21-field list>, "sclane_record"); x = the lineage's pseudotime; returns list(marge_link_fit, marge_link_se, null_link_fit,
 * null_link_se). */
SEXP C_sclane_link_preds(SEXP rec, SEXP x, SEXP sf) {
  static const char* cn[4] = {"marge_link_fit", "marge_link_se", "null_link_fit", "null_link_se"};
  if (TYPEOF(rec) != RAWSXP || XLENGTH(rec) != (R_xlen_t)sizeof(sclane_record)) Rf_error("rec must be the raw sclane_record of a testDynamic() result");
  if (!isReal(x)) Rf_error("x must be a double vector");
  const int64_t n = XLENGTH(x);
  if (sf != R_NilValue && (!isReal(sf) || XLENGTH(sf) != n)) Rf_error("size factors must have one entry per cell of the lineage");
  SEXP out = PROTECT(allocVector(VECSXP, 4));
  for (int k = 0; k < 4; ++k) SET_VECTOR_ELT(out, k, allocVector(REALSXP, (R_xlen_t)n));
  char err[512]; err[0] = 0;
  if (sclane_link_preds((const sclane_record*)RAW(rec), REAL(x), n, sf == R_NilValue ? NULL : REAL(sf), 0, REAL(VECTOR_ELT(out, 0)), REAL(VECTOR_ELT(out, 1)),
                        REAL(VECTOR_ELT(out, 2)), REAL(VECTOR_ELT(out, 3)), err, sizeof(err))) Rf_error("sclane_b200: %s", err);
  SEXP nm = PROTECT(mk_names(4, cn));
  setAttrib(out, R_NamesSymbol, nm);
  UNPROTECT(2);
  return out;
}

/* smoothedCountsMatrix() for one lineage (smoothedCountsMatrix.R:27-74): recs = list of the raw records (one per gene, NULL
 * where absent); returns a cells x genes double matrix (NaN columns for failed models; the R wrapper drops them, :50-52). */
SEXP C_sclane_smoothed_counts(SEXP recs, SEXP x, SEXP sf, SEXP is_gee, SEXP log1p_norm) {
  if (TYPEOF(recs) != VECSXP || !isReal(x)) Rf_error("bad arguments");
  const int64_t n_genes = XLENGTH(recs), n = XLENGTH(x);
  sclane_record* buf = (sclane_record*)R_alloc((size_t)n_genes, sizeof(sclane_record));
  memset(buf, 0, (size_t)n_genes * sizeof(sclane_record));
  for (int64_t g = 0; g < n_genes; ++g) {
    SEXP r = VECTOR_ELT(recs, (R_xlen_t)g);
    if (TYPEOF(r) == RAWSXP && XLENGTH(r) == (R_xlen_t)sizeof(sclane_record)) memcpy(&buf[g], RAW(r), sizeof(sclane_record));
  }
  SEXP out = PROTECT(allocMatrix(REALSXP, (int)n, (int)n_genes));
  char err[512]; err[0] = 0;
  if (sclane_smoothed_counts(buf, n_genes, 1, 0, REAL(x), n, sf == R_NilValue ? NULL : REAL(sf), asInteger(is_gee), asInteger(log1p_norm), 0, REAL(out), err, sizeof(err))) {
    Rf_error("sclane_b200: %s", err);
  }
  UNPROTECT(1);
  return out;
}

/* Drop-ins with the signatures of the reference's Rcpp exports (R/RcppExports.R:4-14).  NOT general: the library kernels take
 * n <= 64 (LLT inverse) and m, n <= 48 (pseudo-inverse), the sizes the hot path needs (p x p systems); larger inputs raise an R
 * error (the reference calls these helpers on n_i x n_i matrices in biasCorrectGEE.R:123-143 -- that code path is replaced by the
 * library's O(n_i) closed forms and never reaches them). */
SEXP C_eigenMapMatMult(SEXP A, SEXP B, SEXP n_cores) {
  (void)n_cores;
  const int m = nrows(A), k = ncols(A), n = ncols(B);
  if (nrows(B) != k) Rf_error("non-conformable arguments");
  SEXP C = PROTECT(allocMatrix(REALSXP, m, n));
  char err[512]; err[0] = 0;
  if (sclane_matmul(REAL(A), REAL(B), REAL(C), m, k, n, 0, err, sizeof(err))) Rf_error("sclane_b200: %s", err);
  UNPROTECT(1);
  return C;
}
SEXP C_eigenMapMatrixInvert(SEXP A, SEXP n_cores) {
  (void)n_cores;
  const int n = nrows(A);
  SEXP B = PROTECT(allocMatrix(REALSXP, n, n));
  char err[512]; err[0] = 0;
  if (sclane_llt_inverse(REAL(A), REAL(B), n, 0, err, sizeof(err))) Rf_error("sclane_b200: %s", err);
  UNPROTECT(1);
  return B;
}
SEXP C_eigenMapPseudoInverse(SEXP A, SEXP tol, SEXP n_cores) {
  (void)n_cores;
  const int m = nrows(A), n = ncols(A);
  SEXP B = PROTECT(allocMatrix(REALSXP, n, m));
  char err[512]; err[0] = 0;
  if (sclane_pinv(REAL(A), REAL(B), m, n, asReal(tol), 0, err, sizeof(err))) Rf_error("sclane_b200: %s", err);
  UNPROTECT(1);
  return B;
}

static const R_CallMethodDef CallEntries[] = {
  {"C_sclane_test_dynamic", (DL_FUNC)&C_sclane_test_dynamic, 5},
  {"C_sclane_test_dynamic_csc", (DL_FUNC)&C_sclane_test_dynamic_csc, 8},
  {"C_sclane_test_dynamic_gee", (DL_FUNC)&C_sclane_test_dynamic_gee, 6},
  {"C_sclane_marge_basis", (DL_FUNC)&C_sclane_marge_basis, 3},
  {"C_sclane_link_preds", (DL_FUNC)&C_sclane_link_preds, 3},
  {"C_sclane_smoothed_counts", (DL_FUNC)&C_sclane_smoothed_counts, 5},
  {"C_eigenMapMatMult", (DL_FUNC)&C_eigenMapMatMult, 3},
  {"C_eigenMapMatrixInvert", (DL_FUNC)&C_eigenMapMatrixInvert, 2},
  {"C_eigenMapPseudoInverse", (DL_FUNC)&C_eigenMapPseudoInverse, 3},
  {NULL, NULL, 0}
};
void R_init_sclaneB200(DllInfo* dll) {
  R_registerRoutines(dll, NULL, CallEntries, NULL, NULL);
  R_useDynamicSymbols(dll, FALSE);
}
