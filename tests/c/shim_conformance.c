/* shim_conformance.c -- TEST HARNESS: drives the C half of the R shim (r/sclane_b200_shim.c, compiled against the R API stand-in
 * of tests/c/rstub/) on inputs written by tests/test_r_shim.py and prints the returned R object as JSON.
 * usage: shim_conformance <prefix> <n_cells> <n_genes> <n_lineages> <with_offset>
 *   <prefix>.counts  int32  [n_genes][n_cells]  (column-major cells x genes)     <prefix>.pt  double [n_lineages][n_cells]
 *   <prefix>.sf      double [n_cells]                                             <prefix>.genes  one name per line */
#include <R.h>
#include <Rinternals.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

SEXP C_sclane_test_dynamic(SEXP counts, SEXP pt, SEXP sf, SEXP genes, SEXP opts_);

static void slurp(const char* prefix, const char* ext, void* dst, size_t bytes) {
  char path[1024];
  snprintf(path, sizeof(path), "%s.%s", prefix, ext);
  FILE* f = fopen(path, "rb");
  if (!f || fread(dst, 1, bytes, f) != bytes) { fprintf(stderr, "cannot read %s\n", path); exit(2); }
  fclose(f);
}

int main(int argc, char** argv) {
  if (argc < 6) { fprintf(stderr, "usage: %s prefix n_cells n_genes n_lineages with_offset\n", argv[0]); return 2; }
  rstub_init();
  const int n_cells = atoi(argv[2]), n_genes = atoi(argv[3]), n_lin = atoi(argv[4]), with_off = atoi(argv[5]);
  SEXP counts = allocMatrix(INTSXP, n_cells, n_genes);
  slurp(argv[1], "counts", INTEGER(counts), (size_t)n_cells * n_genes * 4);
  SEXP pt = allocMatrix(REALSXP, n_cells, n_lin);
  slurp(argv[1], "pt", REAL(pt), (size_t)n_cells * n_lin * 8);
  SEXP sf = R_NilValue;
  if (with_off) { sf = allocVector(REALSXP, n_cells); slurp(argv[1], "sf", REAL(sf), (size_t)n_cells * 8); }
  SEXP genes = allocVector(STRSXP, n_genes);
  {
    char path[1024], line[256];
    snprintf(path, sizeof(path), "%s.genes", argv[1]);
    FILE* f = fopen(path, "r");
    if (!f) { fprintf(stderr, "cannot read %s\n", path); return 2; }
    for (int g = 0; g < n_genes; ++g) {
      if (!fgets(line, sizeof(line), f)) { fprintf(stderr, "too few gene names\n"); return 2; }
      line[strcspn(line, "\r\n")] = 0;
      SET_STRING_ELT(genes, g, mkChar(line));
    }
    fclose(f);
  }
  SEXP opts = allocVector(INTSXP, 3);
  INTEGER(opts)[0] = 5; INTEGER(opts)[1] = 1; INTEGER(opts)[2] = 1;
  SEXP res = C_sclane_test_dynamic(counts, pt, sf, genes, opts);
  rstub_dump_json(res, stdout);
  return 0;
}
