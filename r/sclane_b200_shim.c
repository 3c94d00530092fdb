/* sclane_b200_shim.c -- .Call glue between R and libsclane_b200.so.
 *
 * Source only: R is not available in the build image, so this file is NOT compiled or tested here.  It follows the
 * reference's own native boundary pattern (src/RcppExports.cpp:14-63: one `SEXP f(SEXP...)` per entry point, registered
 * with R_registerRoutines, errors raised with Rf_error) but needs no Rcpp.  Build inside the scLANE package:
 *   PKG_CPPFLAGS = -I<prefix>/include      PKG_LIBS = -L<prefix>/lib -lsclane_b200
 */
#include <R.h>
#include <Rinternals.h>
#include <R_ext/Rdynload.h>
#include <string.h>

#include "sclane_b200.h"

/* counts: integer matrix cells x genes (column-major => gene-major, the layout of the FBM at testDynamic.R:145-147)
 * pt:     double matrix cells x lineages, NA_real_ (a NaN) = cell not in lineage (testDynamic.R:183)
 * sf:     double vector of size factors or NULL (testDynamic.R:228-231)
 * opts:   integer vector c(M, approx_knot, n_gpus)
 * returns a raw vector holding n_genes * n_lineages sclane_record structs + an attribute with the lineage infos;
 * R/testDynamic_b200.R unpacks it into the 21-field list of testDynamic.R:320-342. */
SEXP C_sclane_test_dynamic(SEXP counts, SEXP pt, SEXP sf, SEXP opts_) {
  if (!isInteger(counts) || !isMatrix(counts)) Rf_error("expr.mat must be an integer matrix (cells x genes)");
  if (!isReal(pt) || !isMatrix(pt)) Rf_error("pt must be a double matrix (cells x lineages)");
  const int64_t n_cells = nrows(counts), n_genes = ncols(counts);
  const int n_lin = ncols(pt);
  if (nrows(pt) != n_cells) Rf_error("pt must have one row per cell");
  if (sf != R_NilValue && (!isReal(sf) || XLENGTH(sf) != n_cells)) Rf_error("size.factor.offset must have one entry per cell");
  sclane_opts o;
  sclane_default_opts(&o);
  if (LENGTH(opts_) >= 3) { o.M = INTEGER(opts_)[0]; o.approx_knot = INTEGER(opts_)[1]; o.n_gpus = INTEGER(opts_)[2];
                            for (int i = 0; i < o.n_gpus && i < SCLANE_MAX_GPUS; ++i) o.gpu_ids[i] = i; }
  SEXP out = PROTECT(allocVector(RAWSXP, (R_xlen_t)(n_genes * n_lin) * (R_xlen_t)sizeof(sclane_record)));
  SEXP info = PROTECT(allocVector(RAWSXP, (R_xlen_t)n_lin * (R_xlen_t)sizeof(sclane_lineage_info)));
  char err[1024]; err[0] = 0;
  /* R's NA_integer_ never occurs in counts; NA_real_ in pt is a NaN and is tested with isnan() by the library */
  const int rc = sclane_test_dynamic(INTEGER(counts), n_cells, n_genes, REAL(pt), n_lin,
                                     sf == R_NilValue ? NULL : REAL(sf), &o,
                                     (sclane_record*)RAW(out), (sclane_lineage_info*)RAW(info), err, sizeof(err));
  if (rc != SCLANE_OK) { UNPROTECT(2); Rf_error("sclane_b200: %s", err); }
  setAttrib(out, install("lineage_info"), info);
  UNPROTECT(2);
  return out;
}

/* GEE mode (testDynamic(is.gee = TRUE), testDynamic.R:66-70,235-244,296-301).
 * ids:   integer vector, one subject id per cell (as.integer(factor(id.vec))); cells ordered by subject (testDynamic.R:115)
 * opts:  integer vector c(M, approx_knot, n_gpus, cor_structure [0 independence, 1 exchangeable, 2 ar1], gee_test [0 wald, 1 score],
 *        bias_correction [0 NULL, 1 kc, 2 df]) */
SEXP C_sclane_test_dynamic_gee(SEXP counts, SEXP pt, SEXP ids, SEXP sf, SEXP opts_) {
  if (!isInteger(counts) || !isMatrix(counts)) Rf_error("expr.mat must be an integer matrix (cells x genes)");
  if (!isReal(pt) || !isMatrix(pt)) Rf_error("pt must be a double matrix (cells x lineages)");
  const int64_t n_cells = nrows(counts), n_genes = ncols(counts);
  const int n_lin = ncols(pt);
  if (nrows(pt) != n_cells) Rf_error("pt must have one row per cell");
  if (!isInteger(ids) || XLENGTH(ids) != n_cells) Rf_error("id.vec must be an integer vector with one entry per cell");
  if (sf != R_NilValue && (!isReal(sf) || XLENGTH(sf) != n_cells)) Rf_error("size.factor.offset must have one entry per cell");
  sclane_opts o;
  sclane_default_opts(&o);
  o.mode = 1;
  if (LENGTH(opts_) >= 5) { o.M = INTEGER(opts_)[0]; o.approx_knot = INTEGER(opts_)[1]; o.n_gpus = INTEGER(opts_)[2];
                            o.gee_cor = INTEGER(opts_)[3]; o.gee_test = INTEGER(opts_)[4];
                            if (LENGTH(opts_) >= 6) o.gee_bias_correction = INTEGER(opts_)[5];
                            for (int i = 0; i < o.n_gpus && i < SCLANE_MAX_GPUS; ++i) o.gpu_ids[i] = i; }
  SEXP out = PROTECT(allocVector(RAWSXP, (R_xlen_t)(n_genes * n_lin) * (R_xlen_t)sizeof(sclane_record)));
  SEXP info = PROTECT(allocVector(RAWSXP, (R_xlen_t)n_lin * (R_xlen_t)sizeof(sclane_lineage_info)));
  char err[1024]; err[0] = 0;
  const int rc = sclane_test_dynamic_gee(INTEGER(counts), n_cells, n_genes, REAL(pt), n_lin, INTEGER(ids),
                                         sf == R_NilValue ? NULL : REAL(sf), &o,
                                         (sclane_record*)RAW(out), (sclane_lineage_info*)RAW(info), err, sizeof(err));
  if (rc != SCLANE_OK) { UNPROTECT(2); Rf_error("sclane_b200: %s", err); }
  setAttrib(out, install("lineage_info"), info);
  UNPROTECT(2);
  return out;
}

/* Drop-ins with the signatures of the reference's Rcpp exports (R/RcppExports.R:4-14). */
SEXP C_eigenMapMatMult(SEXP A, SEXP B, SEXP n_cores) {
  const int m = nrows(A), k = ncols(A), n = ncols(B);
  if (nrows(B) != k) Rf_error("non-conformable arguments");
  SEXP C = PROTECT(allocMatrix(REALSXP, m, n));
  char err[512];
  if (sclane_matmul(REAL(A), REAL(B), REAL(C), m, k, n, 0, err, sizeof(err))) { UNPROTECT(1); Rf_error("sclane_b200: %s", err); }
  UNPROTECT(1);
  return C;
}
SEXP C_eigenMapMatrixInvert(SEXP A, SEXP n_cores) {
  const int n = nrows(A);
  SEXP B = PROTECT(allocMatrix(REALSXP, n, n));
  char err[512];
  if (sclane_llt_inverse(REAL(A), REAL(B), n, 0, err, sizeof(err))) { UNPROTECT(1); Rf_error("sclane_b200: %s", err); }
  UNPROTECT(1);
  return B;
}
SEXP C_eigenMapPseudoInverse(SEXP A, SEXP tol, SEXP n_cores) {
  const int m = nrows(A), n = ncols(A);
  SEXP B = PROTECT(allocMatrix(REALSXP, n, m));
  char err[512];
  if (sclane_pinv(REAL(A), REAL(B), m, n, asReal(tol), 0, err, sizeof(err))) { UNPROTECT(1); Rf_error("sclane_b200: %s", err); }
  UNPROTECT(1);
  return B;
}

static const R_CallMethodDef CallEntries[] = {
  {"C_sclane_test_dynamic", (DL_FUNC)&C_sclane_test_dynamic, 4},
  {"C_sclane_test_dynamic_gee", (DL_FUNC)&C_sclane_test_dynamic_gee, 5},
  {"C_eigenMapMatMult", (DL_FUNC)&C_eigenMapMatMult, 3},
  {"C_eigenMapMatrixInvert", (DL_FUNC)&C_eigenMapMatrixInvert, 2},
  {"C_eigenMapPseudoInverse", (DL_FUNC)&C_eigenMapPseudoInverse, 3},
  {NULL, NULL, 0}
};
void R_init_sclaneB200(DllInfo* dll) {
  R_registerRoutines(dll, NULL, CallEntries, NULL, NULL);
  R_useDynamicSymbols(dll, FALSE);
}
