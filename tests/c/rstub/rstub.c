/* Implementation of the R API stand-in + a JSON dump of a SEXP tree (TEST HARNESS ONLY, see Rinternals.h). */
#define _POSIX_C_SOURCE 200809L
#include "Rinternals.h"

#include <math.h>
#include <stdarg.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

static struct rstub_sexp nil_obj = {NILSXP, 0, NULL, NULL};
SEXP R_NilValue = &nil_obj, R_NamesSymbol, R_ClassSymbol, R_RowNamesSymbol, R_DimSymbol, NA_STRING;
double R_NaReal;

static SEXP new_sexp(int type, R_xlen_t n, size_t elt) {
  SEXP s = (SEXP)calloc(1, sizeof(*s));
  s->type = type; s->len = n;
  s->data = calloc((size_t)(n > 0 ? n : 1), elt);
  return s;
}
SEXP Rf_allocVector(int type, R_xlen_t n) {
  switch (type) {
    case INTSXP: case LGLSXP: return new_sexp(type, n, sizeof(int));
    case REALSXP: return new_sexp(type, n, sizeof(double));
    case RAWSXP: return new_sexp(type, n, 1);
    case STRSXP: { SEXP s = new_sexp(type, n, sizeof(SEXP)); for (R_xlen_t i = 0; i < n; ++i) ((SEXP*)s->data)[i] = NA_STRING; return s; }
    case VECSXP: { SEXP s = new_sexp(type, n, sizeof(SEXP)); for (R_xlen_t i = 0; i < n; ++i) ((SEXP*)s->data)[i] = R_NilValue; return s; }
    default: Rf_error("rstub: allocVector of type %d", type); return R_NilValue;
  }
}
SEXP Rf_allocMatrix(int type, int nrow, int ncol) {
  SEXP m = Rf_allocVector(type, (R_xlen_t)nrow * ncol);
  SEXP d = Rf_allocVector(INTSXP, 2);
  INTEGER(d)[0] = nrow; INTEGER(d)[1] = ncol;
  Rf_setAttrib(m, R_DimSymbol, d);
  return m;
}
SEXP Rf_mkChar(const char* s) { SEXP c = new_sexp(CHARSXP, (R_xlen_t)strlen(s), 1); free(c->data); c->data = strdup(s); return c; }
SEXP Rf_mkString(const char* s) { SEXP v = Rf_allocVector(STRSXP, 1); SET_STRING_ELT(v, 0, Rf_mkChar(s)); return v; }
SEXP Rf_ScalarReal(double x) { SEXP v = Rf_allocVector(REALSXP, 1); REAL(v)[0] = x; return v; }
SEXP Rf_ScalarInteger(int x) { SEXP v = Rf_allocVector(INTSXP, 1); INTEGER(v)[0] = x; return v; }
SEXP Rf_ScalarString(SEXP ch) { SEXP v = Rf_allocVector(STRSXP, 1); SET_STRING_ELT(v, 0, ch); return v; }
static SEXP symbols[64];
static int n_symbols = 0;
SEXP Rf_install(const char* name) {
  for (int i = 0; i < n_symbols; ++i) if (strcmp((const char*)symbols[i]->data, name) == 0) return symbols[i];
  SEXP s = new_sexp(SYMSXP, (R_xlen_t)strlen(name), 1);
  free(s->data); s->data = strdup(name);
  if (n_symbols < 64) symbols[n_symbols++] = s;
  return s;
}
SEXP Rf_setAttrib(SEXP x, SEXP tag, SEXP val) {
  for (struct rstub_attr* a = x->attr; a; a = a->next) if (a->tag == tag) { a->val = val; return val; }
  struct rstub_attr* a = (struct rstub_attr*)calloc(1, sizeof(*a));
  a->tag = tag; a->val = val; a->next = x->attr; x->attr = a;
  return val;
}
SEXP Rf_getAttrib(SEXP x, SEXP tag) { for (struct rstub_attr* a = x->attr; a; a = a->next) if (a->tag == tag) return a->val; return R_NilValue; }
void Rf_error(const char* fmt, ...) {
  va_list ap; va_start(ap, fmt);
  fprintf(stderr, "Error: "); vfprintf(stderr, fmt, ap); fprintf(stderr, "\n");
  va_end(ap);
  exit(3);
}
char* R_alloc(size_t n, int size) { return (char*)calloc(n ? n : 1, (size_t)size); }
int Rf_isInteger(SEXP x) { return x->type == INTSXP; }
int Rf_isReal(SEXP x) { return x->type == REALSXP; }
int Rf_isString(SEXP x) { return x->type == STRSXP; }
int Rf_isMatrix(SEXP x) { SEXP d = Rf_getAttrib(x, R_DimSymbol); return d != R_NilValue && d->len == 2; }
int Rf_nrows(SEXP x) { SEXP d = Rf_getAttrib(x, R_DimSymbol); return d != R_NilValue ? INTEGER(d)[0] : (int)x->len; }
int Rf_ncols(SEXP x) { SEXP d = Rf_getAttrib(x, R_DimSymbol); return d != R_NilValue && d->len >= 2 ? INTEGER(d)[1] : 1; }
int Rf_asInteger(SEXP x) { return x->type == INTSXP || x->type == LGLSXP ? INTEGER(x)[0] : (int)REAL(x)[0]; }
double Rf_asReal(SEXP x) { return x->type == REALSXP ? REAL(x)[0] : (double)INTEGER(x)[0]; }

void rstub_init(void) {
  /* R's NA_real_ is a NaN with payload 1954 */
  union { double d; uint64_t u; } v; v.u = 0x7FF00000000007A2ULL; R_NaReal = v.d;
  NA_STRING = new_sexp(CHARSXP, 2, 1); free(NA_STRING->data); NA_STRING->data = strdup("NA");
  R_NamesSymbol = Rf_install("names"); R_ClassSymbol = Rf_install("class"); R_RowNamesSymbol = Rf_install("row.names"); R_DimSymbol = Rf_install("dim");
}

/* ---- JSON: named VECSXP -> object, unnamed -> array, atomic vectors -> arrays, NA / NaN -> null, NULL -> null.  Raw vectors and
 *      the "row.names" / "sclane_record" attributes are skipped; the class attribute is kept as "_class". ---- */
static void dump_str(const char* s, FILE* f) {
  fputc('"', f);
  for (; *s; ++s) { if (*s == '"' || *s == '\\') fputc('\\', f); fputc(*s, f); }
  fputc('"', f);
}
static void dump(SEXP x, FILE* f) {
  if (x == R_NilValue) { fputs("null", f); return; }
  switch (x->type) {
    case REALSXP:
      fputc('[', f);
      for (R_xlen_t i = 0; i < x->len; ++i) { if (i) fputc(',', f); const double v = REAL(x)[i]; if (isnan(v) || isinf(v)) fputs("null", f); else fprintf(f, "%.17g", v); }
      fputc(']', f);
      return;
    case INTSXP: case LGLSXP:
      fputc('[', f);
      for (R_xlen_t i = 0; i < x->len; ++i) { if (i) fputc(',', f); if (INTEGER(x)[i] == NA_INTEGER) fputs("null", f); else fprintf(f, "%d", INTEGER(x)[i]); }
      fputc(']', f);
      return;
    case STRSXP:
      fputc('[', f);
      for (R_xlen_t i = 0; i < x->len; ++i) { if (i) fputc(',', f); SEXP c = STRING_ELT(x, i); if (c == NA_STRING) fputs("null", f); else dump_str(CHAR(c), f); }
      fputc(']', f);
      return;
    case VECSXP: {
      SEXP nm = Rf_getAttrib(x, R_NamesSymbol);
      if (nm != R_NilValue) {
        fputc('{', f);
        for (R_xlen_t i = 0; i < x->len; ++i) { if (i) fputc(',', f); dump_str(CHAR(STRING_ELT(nm, i)), f); fputc(':', f); dump(VECTOR_ELT(x, i), f); }
        SEXP cl = Rf_getAttrib(x, R_ClassSymbol);
        if (cl != R_NilValue) { if (x->len) fputc(',', f); fputs("\"_class\":", f); dump(cl, f); }
        fputc('}', f);
      } else {
        fputc('[', f);
        for (R_xlen_t i = 0; i < x->len; ++i) { if (i) fputc(',', f); dump(VECTOR_ELT(x, i), f); }
        fputc(']', f);
      }
      return;
    }
    default: fputs("null", f); return;
  }
}
void rstub_dump_json(SEXP x, void* file) { dump(x, (FILE*)file); fputc('\n', (FILE*)file); }
