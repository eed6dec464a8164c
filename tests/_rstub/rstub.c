/*
 * tests/_rstub/rstub.c — runtime of the fake R C API declared in Rinternals.h (TEST INFRASTRUCTURE).
 * Collector: runs at every allocation; marks from the PROTECT stack, the driver's roots and the record being built;
 * everything else is POISONED (payload overwritten, type invalidated) but never freed, so a shim that forgot a PROTECT
 * reads 0xDD garbage or trips an accessor check instead of crashing the test process.
 */
#include "Rinternals.h"
#include "R_ext/Rdynload.h"

#include <math.h>
#include <setjmp.h>
#include <stdarg.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

struct SEXPREC {
    int type;
    R_xlen_t length;
    void *data;          /* double / int / SEXP / char payload */
    SEXP dim, names;
    void *ext_addr;
    R_CFinalizer_t fin;
    int mark, poisoned, is_symbol;
    struct SEXPREC *next;
};

static struct SEXPREC nil_rec = { NILSXP, 0, NULL, NULL, NULL, NULL, NULL, 0, 0, 0, NULL };
static struct SEXPREC dim_rec = { CHARSXP, 0, NULL, NULL, NULL, NULL, NULL, 0, 0, 1, NULL };
static struct SEXPREC names_rec = { CHARSXP, 0, NULL, NULL, NULL, NULL, NULL, 0, 0, 1, NULL };
SEXP R_NilValue = &nil_rec, R_DimSymbol = &dim_rec, R_NamesSymbol = &names_rec;
double R_NaReal;
int R_NaInt = INT32_MIN;

static SEXP all_records = NULL;
#define MAX_STACK 10000
static SEXP protect_stack[MAX_STACK];
static int protect_top = 0;
static SEXP roots[MAX_STACK];
static int n_roots = 0;
static long n_collections = 0;

typedef struct ralloc_block { struct ralloc_block *next; } ralloc_block;
static ralloc_block *ralloc_stack = NULL;
static size_t ralloc_count = 0;

static jmp_buf *top_jmp = NULL;
static char top_msg[1024];
static int interrupt_countdown = 0;

static void __attribute__((constructor)) rstub_init(void)
{
    union { double d; uint32_t w[2]; } u;
    u.w[1] = 0x7ff00000u; u.w[0] = 1954u;   /* little endian: low word carries R's NA payload */
    R_NaReal = u.d;
}

static void fatal(const char *what)
{
    /* an API misuse by the shim (poisoned record, bad index, unbalanced stack): reported as an R error when a top
     * level is active so the test sees the message, else abort */
    if (top_jmp) Rf_error("rstub: %s", what);
    fprintf(stderr, "rstub fatal: %s\n", what);
    abort();
}

static void check(SEXP x, const char *fn)
{
    char b[160];
    if (!x) { snprintf(b, sizeof b, "%s on a NULL SEXP", fn); fatal(b); }
    if (x->poisoned) { snprintf(b, sizeof b, "%s on a record the collector has reclaimed (missing PROTECT)", fn); fatal(b); }
}

/* ---------------------------------------------------------------- collector */
static void mark(SEXP x)
{
    if (!x || x == R_NilValue || x->is_symbol || x->mark) return;
    x->mark = 1;
    mark(x->dim);
    mark(x->names);
    if ((x->type == VECSXP || x->type == STRSXP) && x->data)
        for (R_xlen_t i = 0; i < x->length; i++) mark(((SEXP *)x->data)[i]);
}

static void collect(SEXP building)
{
    n_collections++;
    for (SEXP r = all_records; r; r = r->next) r->mark = 0;
    for (int i = 0; i < protect_top; i++) mark(protect_stack[i]);
    for (int i = 0; i < n_roots; i++) mark(roots[i]);
    mark(building);
    for (SEXP r = all_records; r; r = r->next) {
        if (r->mark || r->poisoned) continue;
        if (r->type == EXTPTRSXP && r->fin && r->ext_addr) continue;  /* finalizable: stays until its finalizer ran */
        size_t bytes = 0;
        switch (r->type) {
        case REALSXP: bytes = (size_t)r->length * 8; break;
        case INTSXP: case LGLSXP: bytes = (size_t)r->length * 4; break;
        case VECSXP: case STRSXP: bytes = (size_t)r->length * sizeof(SEXP); break;
        case CHARSXP: bytes = (size_t)r->length; break;
        default: break;
        }
        if (r->data && bytes) memset(r->data, 0xDD, bytes);
        r->poisoned = 1;
    }
}

static SEXP new_record(int type, R_xlen_t n, size_t elt)
{
    SEXP r = (SEXP)calloc(1, sizeof(struct SEXPREC));
    if (!r) fatal("out of memory");
    r->type = type; r->length = n; r->dim = R_NilValue; r->names = R_NilValue;
    if (n > 0 && elt) {
        r->data = calloc((size_t)n, elt);
        if (!r->data) fatal("out of memory");
        if (type == REALSXP || type == INTSXP || type == LGLSXP) memset(r->data, 0xAB, (size_t)n * elt); /* R does not zero */
        if (type == VECSXP || type == STRSXP) for (R_xlen_t i = 0; i < n; i++) ((SEXP *)r->data)[i] = R_NilValue;
    }
    collect(r);                     /* gctorture: a collection at every allocation */
    r->next = all_records;
    all_records = r;
    return r;
}

/* ---------------------------------------------------------------- API */
int TYPEOF(SEXP x) { check(x, "TYPEOF"); return x->type; }
R_xlen_t XLENGTH(SEXP x) { check(x, "XLENGTH"); return x->length; }
R_xlen_t Rf_xlength(SEXP x) { check(x, "xlength"); return x->length; }
int Rf_length(SEXP x) { check(x, "length"); return (int)x->length; }
double *REAL(SEXP x) { check(x, "REAL"); if (x->type != REALSXP) fatal("REAL() on a non-double"); return (double *)x->data; }
int *INTEGER(SEXP x) { check(x, "INTEGER"); if (x->type != INTSXP && x->type != LGLSXP) fatal("INTEGER() on a non-integer"); return (int *)x->data; }
int *LOGICAL(SEXP x) { check(x, "LOGICAL"); if (x->type != LGLSXP) fatal("LOGICAL() on a non-logical"); return (int *)x->data; }
SEXP VECTOR_ELT(SEXP x, R_xlen_t i)
{
    check(x, "VECTOR_ELT");
    if (x->type != VECSXP || i < 0 || i >= x->length) fatal("VECTOR_ELT out of range / not a list");
    return ((SEXP *)x->data)[i];
}
SEXP SET_VECTOR_ELT(SEXP x, R_xlen_t i, SEXP v)
{
    check(x, "SET_VECTOR_ELT"); check(v, "SET_VECTOR_ELT value");
    if (x->type != VECSXP || i < 0 || i >= x->length) fatal("SET_VECTOR_ELT out of range / not a list");
    ((SEXP *)x->data)[i] = v;
    return v;
}
SEXP STRING_ELT(SEXP x, R_xlen_t i)
{
    check(x, "STRING_ELT");
    if (x->type != STRSXP || i < 0 || i >= x->length) fatal("STRING_ELT out of range / not a character vector");
    return ((SEXP *)x->data)[i];
}
void SET_STRING_ELT(SEXP x, R_xlen_t i, SEXP v)
{
    check(x, "SET_STRING_ELT"); check(v, "SET_STRING_ELT value");
    if (x->type != STRSXP || v->type != CHARSXP || i < 0 || i >= x->length) fatal("SET_STRING_ELT misuse");
    ((SEXP *)x->data)[i] = v;
}
const char *CHAR(SEXP x) { check(x, "CHAR"); if (x->type != CHARSXP) fatal("CHAR() on a non-CHARSXP"); return (const char *)x->data; }
SEXP Rf_mkChar(const char *s)
{
    SEXP r = new_record(CHARSXP, (R_xlen_t)strlen(s) + 1, 1);
    memcpy(r->data, s, strlen(s) + 1);
    return r;
}
SEXP Rf_allocVector(unsigned type, R_xlen_t n)
{
    if (n < 0) fatal("allocVector: negative length");
    switch (type) {
    case REALSXP: return new_record(REALSXP, n, 8);
    case INTSXP: return new_record(INTSXP, n, 4);
    case LGLSXP: return new_record(LGLSXP, n, 4);
    case VECSXP: return new_record(VECSXP, n, sizeof(SEXP));
    case STRSXP: return new_record(STRSXP, n, sizeof(SEXP));
    default: fatal("allocVector: type not modelled by the stub"); return R_NilValue;
    }
}
SEXP Rf_allocMatrix(unsigned type, int nrow, int ncol)
{
    if (nrow < 0 || ncol < 0) fatal("allocMatrix: negative extent");
    SEXP m = PROTECT(Rf_allocVector(type, (R_xlen_t)nrow * ncol));
    SEXP d = Rf_allocVector(INTSXP, 2);
    INTEGER(d)[0] = nrow; INTEGER(d)[1] = ncol;
    m->dim = d;
    UNPROTECT(1);
    return m;
}
SEXP Rf_protect(SEXP x)
{
    check(x, "PROTECT");
    if (protect_top >= MAX_STACK) fatal("protect(): protection stack overflow");
    protect_stack[protect_top++] = x;
    return x;
}
void Rf_unprotect(int n)
{
    if (n > protect_top) fatal("unprotect(): only fewer protected items");
    protect_top -= n;
}
SEXP Rf_getAttrib(SEXP x, SEXP name)
{
    check(x, "getAttrib");
    if (name == R_DimSymbol) return x->dim;
    if (name == R_NamesSymbol) return x->names;
    return R_NilValue;
}
SEXP Rf_setAttrib(SEXP x, SEXP name, SEXP v)
{
    check(x, "setAttrib"); check(v, "setAttrib value");
    if (name == R_DimSymbol) x->dim = v;
    else if (name == R_NamesSymbol) x->names = v;
    else fatal("setAttrib: attribute not modelled by the stub");
    return v;
}
void Rf_error(const char *fmt, ...)
{
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(top_msg, sizeof top_msg, fmt, ap);
    va_end(ap);
    if (!top_jmp) { fprintf(stderr, "Rf_error outside a top level: %s\n", top_msg); abort(); }
    longjmp(*top_jmp, 1);
}
char *R_alloc(size_t n, int size)
{
    ralloc_block *b = (ralloc_block *)malloc(sizeof(ralloc_block) + n * (size_t)size + 16);
    if (!b) Rf_error("cannot allocate memory block of size %.1f Gb", (double)n * size / 1073741824.0);
    memset(b + 1, 0xCD, n * (size_t)size);  /* R_alloc does not zero either */
    b->next = ralloc_stack;
    ralloc_stack = b;
    ralloc_count++;
    return (char *)(b + 1);
}
void R_CheckUserInterrupt(void)
{
    if (interrupt_countdown > 0 && --interrupt_countdown == 0) {
        snprintf(top_msg, sizeof top_msg, "interrupted");
        if (!top_jmp) abort();
        longjmp(*top_jmp, 2);
    }
}
int R_IsNA(double x)
{
    union { double d; uint32_t w[2]; } u;
    u.d = x;
    return isnan(x) && u.w[0] == 1954u;
}
SEXP R_MakeExternalPtr(void *p, SEXP tag, SEXP prot)
{
    (void)tag; (void)prot;
    SEXP r = new_record(EXTPTRSXP, 0, 0);
    r->ext_addr = p;
    return r;
}
void *R_ExternalPtrAddr(SEXP s) { check(s, "R_ExternalPtrAddr"); if (s->type != EXTPTRSXP) fatal("not an external pointer"); return s->ext_addr; }
void R_ClearExternalPtr(SEXP s) { check(s, "R_ClearExternalPtr"); s->ext_addr = NULL; }
void R_RegisterCFinalizerEx(SEXP s, R_CFinalizer_t fun, Rboolean onexit) { (void)onexit; check(s, "R_RegisterCFinalizerEx"); s->fin = fun; }

int R_registerRoutines(DllInfo *info, const R_CMethodDef *c, const R_CallMethodDef *call, const void *f, const void *e)
{
    (void)c; (void)f; (void)e;
    info->call = call;
    return 1;
}
Rboolean R_useDynamicSymbols(DllInfo *info, Rboolean value) { info->dynamic = value; return TRUE; }

/* ---------------------------------------------------------------- driver side */
int rstub_toplevel(rstub_body body, void *arg, SEXP *result, char *msg, size_t msglen)
{
    jmp_buf jb, *outer = top_jmp;
    const int depth = protect_top;
    ralloc_block *vmax = ralloc_stack;
    volatile int rc;
    top_jmp = &jb;
    rc = setjmp(jb);
    if (rc == 0) {
        SEXP r = body(arg);
        if (result) *result = r;
        if (protect_top != depth) {
            rc = 1;
            snprintf(top_msg, sizeof top_msg, "stack imbalance in .Call: %d then %d", depth, protect_top);
        }
    }
    if (rc != 0 && msg && msglen) snprintf(msg, msglen, "%s", top_msg);
    protect_top = depth;                       /* R unwinds the protect stack to the context's depth */
    while (ralloc_stack != vmax) {             /* vmaxset(): R_alloc memory of the call goes away */
        ralloc_block *b = ralloc_stack;
        ralloc_stack = b->next;
        free(b);
        ralloc_count--;
    }
    top_jmp = outer;
    return rc;
}
void rstub_root(SEXP x) { if (n_roots < MAX_STACK) roots[n_roots++] = x; }
void rstub_unroot_all(void) { n_roots = 0; }
void rstub_run_finalizers(void)
{
    for (SEXP r = all_records; r; r = r->next)
        if (r->type == EXTPTRSXP && r->fin && r->ext_addr && !r->poisoned) {
            R_CFinalizer_t f = r->fin;
            r->fin = NULL;
            f(r);
        }
}
int rstub_protect_depth(void) { return protect_top; }
int rstub_live_finalizers(void)
{
    int n = 0;
    for (SEXP r = all_records; r; r = r->next) n += (r->type == EXTPTRSXP && r->fin && r->ext_addr);
    return n;
}
long rstub_collections(void) { return n_collections; }
void rstub_interrupt_after(int checks) { interrupt_countdown = checks; }
size_t rstub_ralloc_live(void) { return ralloc_count; }
