/* Minimal stand-in for R's C API (Rinternals.h) -- TEST HARNESS ONLY.
 * Just enough of the public API for r/sclane_b200_shim.c to compile and run without R, so that the shim's C half can be driven
 * on the GPU box and its result tree diffed against the reference's golden output (tests/test_r_shim.py).  Semantics follow
 * "Writing R Extensions": SEXP vectors with attributes; PROTECT is a no-op (nothing is ever collected); Rf_error prints and exits. */
#ifndef RSTUB_RINTERNALS_H
#define RSTUB_RINTERNALS_H
#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef ptrdiff_t R_xlen_t;
typedef unsigned char Rbyte;
enum { NILSXP = 0, SYMSXP = 1, CHARSXP = 9, LGLSXP = 10, INTSXP = 13, REALSXP = 14, STRSXP = 16, VECSXP = 19, RAWSXP = 24 };

typedef struct rstub_sexp {
  int type;
  R_xlen_t len;
  void* data;                      /* int / double / Rbyte / SEXP[] / char[] */
  struct rstub_attr* attr;
} *SEXP;
struct rstub_attr { SEXP tag; SEXP val; struct rstub_attr* next; };

extern SEXP R_NilValue, R_NamesSymbol, R_ClassSymbol, R_RowNamesSymbol, R_DimSymbol, NA_STRING;
extern double R_NaReal;
#define NA_REAL R_NaReal
#define NA_INTEGER INT32_MIN

SEXP Rf_allocVector(int type, R_xlen_t n);
SEXP Rf_allocMatrix(int type, int nrow, int ncol);
SEXP Rf_mkChar(const char* s);
SEXP Rf_mkString(const char* s);
SEXP Rf_ScalarReal(double x);
SEXP Rf_ScalarInteger(int x);
SEXP Rf_ScalarString(SEXP ch);
SEXP Rf_install(const char* name);
SEXP Rf_setAttrib(SEXP x, SEXP tag, SEXP val);
SEXP Rf_getAttrib(SEXP x, SEXP tag);
void Rf_error(const char* fmt, ...);
char* R_alloc(size_t n, int size);
int Rf_isInteger(SEXP x);
int Rf_isReal(SEXP x);
int Rf_isString(SEXP x);
int Rf_isMatrix(SEXP x);
int Rf_nrows(SEXP x);
int Rf_ncols(SEXP x);
int Rf_asInteger(SEXP x);
double Rf_asReal(SEXP x);

#define allocVector Rf_allocVector
#define allocMatrix Rf_allocMatrix
#define mkChar Rf_mkChar
#define mkString Rf_mkString
#define ScalarReal Rf_ScalarReal
#define ScalarInteger Rf_ScalarInteger
#define ScalarString Rf_ScalarString
#define install Rf_install
#define setAttrib Rf_setAttrib
#define getAttrib Rf_getAttrib
#define isInteger Rf_isInteger
#define isReal Rf_isReal
#define isString Rf_isString
#define isMatrix Rf_isMatrix
#define nrows Rf_nrows
#define ncols Rf_ncols
#define asInteger Rf_asInteger
#define asReal Rf_asReal

#define PROTECT(x) (x)
#define UNPROTECT(n) ((void)(n))
#define TYPEOF(x) ((x)->type)
#define LENGTH(x) ((int)(x)->len)
#define XLENGTH(x) ((x)->len)
#define INTEGER(x) ((int*)(x)->data)
#define REAL(x) ((double*)(x)->data)
#define RAW(x) ((Rbyte*)(x)->data)
#define CHAR(x) ((const char*)(x)->data)
#define STRING_ELT(x, i) (((SEXP*)(x)->data)[i])
#define VECTOR_ELT(x, i) (((SEXP*)(x)->data)[i])
#define SET_STRING_ELT(x, i, v) (((SEXP*)(x)->data)[i] = (v))
#define SET_VECTOR_ELT(x, i, v) (((SEXP*)(x)->data)[i] = (v))

/* test-side helpers (not part of R) */
void rstub_init(void);
void rstub_dump_json(SEXP x, void* file);

#ifdef __cplusplus
}
#endif
#endif
