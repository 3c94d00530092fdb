/* abi_smoke.c -- plain C99 caller of libsclane_b200.so (no Python, no C++): what the R shim does, minus R.
 * Built and run by tests/test_c_abi.py.  Usage: abi_smoke [n_genes n_cells]
 * Generates Poisson-ish counts with a hinge at 0.5 for half of the genes, runs sclane_test_dynamic(), prints one line per
 * gene, and exits 0 iff every record is sane (hinge genes dynamic with a knot near 0.5, flat genes not). */
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "sclane_b200.h"

static unsigned long long rng_state = 88172645463325252ULL;
static double urand(void) {
  rng_state ^= rng_state << 13; rng_state ^= rng_state >> 7; rng_state ^= rng_state << 17;
  return (double)(rng_state >> 11) * (1.0 / 9007199254740992.0);
}
static int poisson(double lambda) {           /* Knuth; lambda is small here */
  const double L = exp(-lambda);
  int k = 0;
  double p = 1.0;
  do { ++k; p *= urand(); } while (p > L);
  return k - 1;
}

int main(int argc, char** argv) {
  const long n_genes = argc > 2 ? atol(argv[1]) : 8, n_cells = argc > 2 ? atol(argv[2]) : 4000;
  if (sclane_abi_version() != SCLANE_ABI_VERSION) { fprintf(stderr, "ABI version mismatch\n"); return 2; }
  if (sclane_record_size() != sizeof(sclane_record)) { fprintf(stderr, "record size mismatch\n"); return 2; }
  int32_t* counts = (int32_t*)malloc(sizeof(int32_t) * (size_t)n_genes * (size_t)n_cells);
  double* pt = (double*)malloc(sizeof(double) * (size_t)n_cells);
  sclane_record* out = (sclane_record*)malloc(sizeof(sclane_record) * (size_t)n_genes);
  sclane_lineage_info info;
  for (long i = 0; i < n_cells; ++i) pt[i] = ((double)((i * 7919L) % n_cells) + 0.5) / (double)n_cells;
  for (long g = 0; g < n_genes; ++g)
    for (long i = 0; i < n_cells; ++i) {
      const double x = pt[i];
      const double eta = (g % 2 == 0) ? 0.7 + 3.0 * (x > 0.5 ? x - 0.5 : 0.0) : 0.9;
      counts[g * n_cells + i] = poisson(exp(eta));
    }
  sclane_opts o;
  sclane_default_opts(&o);
  o.n_gpus = 1; o.gpu_ids[0] = 0;
  char err[512]; err[0] = 0;
  if (sclane_device_count() < 1) { printf("no CUDA device: the library has no CPU fallback (expected on a CPU box)\n"); return 3; }
  const int rc = sclane_test_dynamic(counts, n_cells, n_genes, pt, 1, NULL, &o, out, &info, err, sizeof(err));
  if (rc != SCLANE_OK) { fprintf(stderr, "sclane_test_dynamic failed (%d): %s\n", rc, err); return 1; }
  int bad = 0;
  for (long g = 0; g < n_genes; ++g) {
    const sclane_record* r = &out[g];
    printf("gene %ld status %d terms %d LRT %.6g df %.0f p %.3g theta %.4g knots:", g, r->status, r->n_terms, r->test_stat, r->df, r->p_val, r->theta);
    for (int j = 1; j < r->n_terms; ++j) printf(" %s%.4g", r->term_kind[j] == SCLANE_TERM_TP1 ? "+" : "-", r->term_knot[j]);
    printf("\n");
    if (r->status != (SCLANE_MARGE_OK | SCLANE_NULL_OK)) bad++;
    else if (g % 2 == 0) {
      int near = 0;
      for (int j = 1; j < r->n_terms; ++j) if (fabs(r->term_knot[j] - 0.5) < 0.08) near = 1;
      if (!(r->p_val < 1e-6) || !near) bad++;
    } else if (r->p_val < 1e-4) bad++;
  }
  printf("launches %lld, knots %d, lineage cells %lld, bad %d\n", (long long)sclane_launch_count(), info.n_knots, (long long)info.n_cells, bad);
  sclane_release();
  free(counts); free(pt); free(out);
  return bad ? 1 : 0;
}
