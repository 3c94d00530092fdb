/* stand-in for R_ext/Rdynload.h (TEST HARNESS ONLY, see ../Rinternals.h) */
#ifndef RSTUB_RDYNLOAD_H
#define RSTUB_RDYNLOAD_H
typedef void* (*DL_FUNC)(void);
typedef struct { const char* name; DL_FUNC fun; int numArgs; } R_CallMethodDef;
typedef struct rstub_dll { int unused; } DllInfo;
static inline int R_registerRoutines(DllInfo* d, const void* a, const R_CallMethodDef* c, const void* f, const void* e) { (void)d; (void)a; (void)c; (void)f; (void)e; return 1; }
static inline int R_useDynamicSymbols(DllInfo* d, int v) { (void)d; (void)v; return 0; }
#define FALSE 0
#define TRUE 1
#endif
