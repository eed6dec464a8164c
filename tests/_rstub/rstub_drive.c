/*
 * tests/_rstub/rstub_drive.c — plain-C entry points (for ctypes) that build fake SEXPs from numpy buffers, call the
 * shim's .Call routines inside a top-level context and copy the results back (TEST INFRASTRUCTURE).
 */
#include "Rinternals.h"
#include "R_ext/Rdynload.h"

#include <math.h>
#include <stdio.h>
#include <string.h>

SEXP C_epx_upload(SEXP methCols, SEXP device);
SEXP C_epx_analyze(SEXP session, SEXP expr, SEXP rowPtr, SEXP probeIdx, SEXP opts);
SEXP C_epx_groups(SEXP session, SEXP groupPtr, SEXP groupPairs);
SEXP C_epx_release(SEXP session);
SEXP C_epx_device_info(void);
void R_init_EpiMethEx(DllInfo *dll);

typedef struct {
    int rc;                 /* 0 ok, 1 Rf_error, 2 interrupt */
    int protect_depth;      /* after the call: must equal the depth before it */
    int live_finalizers;
    long collections;
    long ralloc_live;       /* R_alloc blocks still allocated after the call: must be 0 */
    char msg[512];
} rshim_report;

static void fill(rshim_report *rep, int rc)
{
    rep->rc = rc;
    rep->protect_depth = rstub_protect_depth();
    rep->live_finalizers = rstub_live_finalizers();
    rep->collections = rstub_collections();
    rep->ralloc_live = (long)rstub_ralloc_live();
}

static SEXP rooted(SEXP x) { rstub_root(x); return x; }

static SEXP real_vec(const double *v, long n)
{
    SEXP r = rooted(Rf_allocVector(REALSXP, n));
    if (n) memcpy(REAL(r), v, (size_t)n * 8);
    return r;
}
static SEXP int_vec(const int *v, long n)
{
    SEXP r = rooted(Rf_allocVector(INTSXP, n));
    if (n) memcpy(INTEGER(r), v, (size_t)n * 4);
    return r;
}

typedef struct { SEXP a, b, c, d, e; } args5;
static SEXP body_upload(void *p) { args5 *a = (args5 *)p; return C_epx_upload(a->a, a->b); }
static SEXP body_analyze(void *p) { args5 *a = (args5 *)p; return C_epx_analyze(a->a, a->b, a->c, a->d, a->e); }
static SEXP body_groups(void *p) { args5 *a = (args5 *)p; return C_epx_groups(a->a, a->b, a->c); }
static SEXP body_release(void *p) { args5 *a = (args5 *)p; return C_epx_release(a->a); }
static SEXP body_info(void *p) { (void)p; return C_epx_device_info(); }

/* methylation as n column pointers of P doubles; as_int[j] != 0: hand column j over as an INTSXP (NaN -> NA_integer_);
 * bad_kind: 0 well-formed, 1 one column too short, 2 a character column, 3 devices is a double vector */
void *rshim_upload(const double *const *cols, int n, long P, const int *as_int, const int *devices, int nd, int bad_kind,
                   rshim_report *rep)
{
    SEXP frame = rooted(Rf_allocVector(VECSXP, n));
    for (int j = 0; j < n; j++) {
        SEXP c;
        long len = (bad_kind == 1 && j == n - 1) ? P - 1 : P;
        if (bad_kind == 2 && j == 1) c = Rf_allocVector(STRSXP, len);
        else if (as_int && as_int[j]) {
            c = Rf_allocVector(INTSXP, len);
            for (long i = 0; i < len; i++) INTEGER(c)[i] = isnan(cols[j][i]) ? NA_INTEGER : (int)cols[j][i];
        } else {
            c = Rf_allocVector(REALSXP, len);
            memcpy(REAL(c), cols[j], (size_t)len * 8);
        }
        SET_VECTOR_ELT(frame, j, c);
    }
    SEXP dev;
    if (bad_kind == 3) { double d = 0; dev = real_vec(&d, 1); } else dev = int_vec(devices, nd);
    args5 a = { frame, dev, NULL, NULL, NULL };
    SEXP out = NULL;
    int rc = rstub_toplevel(body_upload, &a, &out, rep->msg, sizeof rep->msg);
    if (rc == 0) rstub_root(out);       /* the R variable `session` */
    fill(rep, rc);
    return rc == 0 ? (void *)out : NULL;
}

/* probe_idx: INT32_MIN = NA_integer_. row_ptr as doubles (what the R wrapper passes). interrupt_at: raise an interrupt at
 * the k-th R_CheckUserInterrupt of the call (0 = never). Outputs may be NULL. */
int rshim_analyze(void *session, const double *expr, int n, int G, const double *row_ptr, long n_row_ptr,
                  const int *probe_idx, long n_probe_idx, const int *opts, int n_opts, int interrupt_at,
                  double *pair, double *gene, int *perm, int *dnum, int *counts, int *na_pair_mask, rshim_report *rep)
{
    SEXP x = rooted(Rf_allocMatrix(REALSXP, n, G));
    memcpy(REAL(x), expr, (size_t)n * G * 8);
    args5 a = { (SEXP)session, x, real_vec(row_ptr, n_row_ptr), int_vec(probe_idx, n_probe_idx), int_vec(opts, n_opts) };
    SEXP out = NULL;
    rstub_interrupt_after(interrupt_at);
    int rc = rstub_toplevel(body_analyze, &a, &out, rep->msg, sizeof rep->msg);
    rstub_interrupt_after(0);
    if (rc == 0) {
        rstub_root(out);
        const long Np = (long)row_ptr[n_row_ptr - 1];
        /* list(pair, gene, perm, ksDnum, counts) with names */
        SEXP names = Rf_getAttrib(out, R_NamesSymbol);
        static const char *want[5] = { "pair", "gene", "perm", "ksDnum", "counts" };
        for (int i = 0; i < 5; i++)
            if (strcmp(CHAR(STRING_ELT(names, i)), want[i]) != 0) { snprintf(rep->msg, sizeof rep->msg, "name %d is %s", i, CHAR(STRING_ELT(names, i))); rc = 3; }
        SEXP p = VECTOR_ELT(out, 0), g = VECTOR_ELT(out, 1), pm = VECTOR_ELT(out, 2), dn = VECTOR_ELT(out, 3), ct = VECTOR_ELT(out, 4);
        SEXP pd = Rf_getAttrib(p, R_DimSymbol);
        if (INTEGER(pd)[0] != 11 || INTEGER(pd)[1] != Np || XLENGTH(g) != 15L * G || XLENGTH(pm) != (long)n * G || XLENGTH(dn) != 3 * Np) {
            snprintf(rep->msg, sizeof rep->msg, "result shapes are wrong"); rc = 3;
        }
        if (rc == 0) {
            if (pair) memcpy(pair, REAL(p), (size_t)Np * 11 * 8);
            if (na_pair_mask)  /* which cells hold R's NA (payload 1954), as opposed to a plain NaN */
                for (long i = 0; i < Np * 11; i++) na_pair_mask[i] = R_IsNA(REAL(p)[i]);
            if (gene) memcpy(gene, REAL(g), (size_t)G * 15 * 8);
            if (perm) memcpy(perm, INTEGER(pm), (size_t)n * G * 4);
            if (dnum) memcpy(dnum, INTEGER(dn), (size_t)Np * 3 * 4);
            if (counts) memcpy(counts, INTEGER(ct), 4 * 4);
        }
    }
    fill(rep, rc);
    return rc;
}

int rshim_groups(void *session, const double *group_ptr, long n_group_ptr, const int *group_pairs, long n_pairs,
                 double *out11, rshim_report *rep)
{
    args5 a = { (SEXP)session, real_vec(group_ptr, n_group_ptr), int_vec(group_pairs, n_pairs), NULL, NULL };
    SEXP out = NULL;
    int rc = rstub_toplevel(body_groups, &a, &out, rep->msg, sizeof rep->msg);
    if (rc == 0) {
        rstub_root(out);
        const long Q = n_group_ptr - 1;
        if (XLENGTH(out) != 11 * Q) { snprintf(rep->msg, sizeof rep->msg, "group table has %ld cells", (long)XLENGTH(out)); rc = 3; }
        else if (out11) memcpy(out11, REAL(out), (size_t)Q * 11 * 8);
    }
    fill(rep, rc);
    return rc;
}

int rshim_release(void *session, rshim_report *rep)
{
    args5 a = { (SEXP)session, NULL, NULL, NULL, NULL };
    int rc = rstub_toplevel(body_release, &a, NULL, rep->msg, sizeof rep->msg);
    fill(rep, rc);
    return rc;
}

int rshim_device_info(char *name, size_t len, rshim_report *rep)
{
    SEXP out = NULL;
    int rc = rstub_toplevel(body_info, NULL, &out, rep->msg, sizeof rep->msg);
    if (rc == 0) snprintf(name, len, "%s", CHAR(STRING_ELT(out, 0)));
    fill(rep, rc);
    return rc;
}

/* end of the "R session": drop the driver's variables and run the finalizers of what is left */
void rshim_finish(rshim_report *rep)
{
    rstub_unroot_all();
    rstub_run_finalizers();
    fill(rep, 0);
}

/* R_init_EpiMethEx: "name/nargs;..." of the registered .Call routines */
int rshim_registered(char *buf, size_t len)
{
    DllInfo info = { NULL, 1 };
    R_init_EpiMethEx(&info);
    size_t k = 0;
    buf[0] = 0;
    for (const R_CallMethodDef *m = info.call; m && m->name; m++)
        k += (size_t)snprintf(buf + k, k < len ? len - k : 0, "%s/%d;", m->name, m->numArgs);
    return info.dynamic;
}
