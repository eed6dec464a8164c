/*
 * tests/_rstub/Rinternals.h — a small FAKE of R's C API, just large enough to compile and drive
 * r_package/src/epx_rshim.c where R itself cannot be installed (TEST INFRASTRUCTURE; not R, not shipped).
 * What it models, because the shim's correctness depends on it:
 *   - SEXP records with type, length, payload, `dim` / `names` attributes;
 *   - a PROTECT stack, and a garbage collector that runs at EVERY allocation (gctorture): records reachable neither
 *     from the protect stack nor from the driver's roots are poisoned, so a missing PROTECT shows up as garbage;
 *   - Rf_error / R_CheckUserInterrupt as longjmps to the driver's top level (rstub_toplevel), which unwinds the
 *     protect stack and the R_alloc stack exactly as R does;
 *   - R_alloc with its vmax stack; external pointers with C finalizers; NA_real_ as the NaN with payload 1954.
 */
#ifndef RSTUB_RINTERNALS_H
#define RSTUB_RINTERNALS_H
#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef ptrdiff_t R_xlen_t;
typedef enum { FALSE = 0, TRUE = 1 } Rboolean;
typedef struct SEXPREC *SEXP;

enum { NILSXP = 0, CHARSXP = 9, LGLSXP = 10, INTSXP = 13, REALSXP = 14, STRSXP = 16, VECSXP = 19, EXTPTRSXP = 22 };

extern SEXP R_NilValue, R_DimSymbol, R_NamesSymbol;
extern double R_NaReal;
extern int R_NaInt;
#define NA_REAL R_NaReal
#define NA_INTEGER R_NaInt

int TYPEOF(SEXP x);
R_xlen_t XLENGTH(SEXP x);
int Rf_length(SEXP x);
R_xlen_t Rf_xlength(SEXP x);
double *REAL(SEXP x);
int *INTEGER(SEXP x);
int *LOGICAL(SEXP x);
SEXP VECTOR_ELT(SEXP x, R_xlen_t i);
SEXP SET_VECTOR_ELT(SEXP x, R_xlen_t i, SEXP v);
SEXP STRING_ELT(SEXP x, R_xlen_t i);
void SET_STRING_ELT(SEXP x, R_xlen_t i, SEXP v);
const char *CHAR(SEXP x);
SEXP Rf_mkChar(const char *s);
SEXP Rf_allocVector(unsigned type, R_xlen_t n);
SEXP Rf_allocMatrix(unsigned type, int nrow, int ncol);
SEXP Rf_protect(SEXP x);
void Rf_unprotect(int n);
#define PROTECT(x) Rf_protect(x)
#define UNPROTECT(n) Rf_unprotect(n)
SEXP Rf_getAttrib(SEXP x, SEXP name);
SEXP Rf_setAttrib(SEXP x, SEXP name, SEXP v);
void Rf_error(const char *fmt, ...) __attribute__((noreturn, format(printf, 1, 2)));
char *R_alloc(size_t n, int size);
void R_CheckUserInterrupt(void);
int R_IsNA(double x);

SEXP R_MakeExternalPtr(void *p, SEXP tag, SEXP prot);
void *R_ExternalPtrAddr(SEXP s);
void R_ClearExternalPtr(SEXP s);
typedef void (*R_CFinalizer_t)(SEXP);
void R_RegisterCFinalizerEx(SEXP s, R_CFinalizer_t fun, Rboolean onexit);

/* ---- driver side (not part of R): tests/_rstub/rstub_drive.c ---- */
typedef SEXP (*rstub_body)(void *arg);
int rstub_toplevel(rstub_body body, void *arg, SEXP *result, char *msg, size_t msglen); /* 0 ok, 1 Rf_error, 2 interrupt */
void rstub_root(SEXP x);            /* keep x (and what it references) alive: the caller's variables */
void rstub_unroot_all(void);
void rstub_run_finalizers(void);    /* what R does at exit / at a full collection of unreachable external pointers */
int rstub_protect_depth(void);
int rstub_live_finalizers(void);
long rstub_collections(void);
void rstub_interrupt_after(int checks); /* the n-th R_CheckUserInterrupt from now raises an interrupt; 0 = never */
size_t rstub_ralloc_live(void);

#ifdef __cplusplus
}
#endif
#endif
