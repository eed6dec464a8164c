/*
 * epx_rshim.c — the `.Call` boundary between R and libepx (include/epx.h).
 *
 * R, Rinternals.h and libR are absent from the build environment.  This file is therefore compiled with
 * -Wall -Wextra -Werror against tests/_rstub/ (a small fake of R's C API: SEXP records, a PROTECT stack, a garbage
 * collector that runs at EVERY allocation and poisons whatever is unreachable, Rf_error as a longjmp, R_alloc with
 * its vmax stack, external pointers with finalizers) and driven on the GPU by tests/test_gpu_rshim.py; the same C ABI
 * underneath is what every other GPU test and bench.py call through ctypes.  Rules it follows:
 *   - arguments belong to R and are read-only; nothing is written through REAL()/INTEGER() of an argument;
 *   - every argument's type and length is validated before it is indexed;
 *   - Rf_error, R_CheckUserInterrupt and every allocator of R can longjmp out of the function: no malloc'ed memory
 *     is ever live across them.  Scratch memory comes from R_alloc (released by R when .Call returns or unwinds), the
 *     device sessions hang off an external pointer whose finalizer is registered BEFORE the session is created;
 *   - every SEXP allocation is PROTECTed while another allocation can happen; UNPROTECT is balanced on every path
 *     that returns (an Rf_error unwinds the protect stack itself);
 *   - libepx never calls R and never longjmps: errors come back as code + message, and only then Rf_error is raised;
 *   - no CPU fallback: a missing GPU is an R error.
 *
 * Replaces the numeric part of R/EpimethexAnalysis.R:117-206 and :299-453 and R/Functions.R:2-132,192-220.
 */
#include <R.h>
#include <Rinternals.h>
#include <R_ext/Rdynload.h>
#include <limits.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#include "epx.h"

/* The handle behind the external pointer: one session (device = one integer) or an epx_multi (device = several
 * integers: one process drives all GPUs of the box, replacing the PSOCK cluster of R/EpimethexAnalysis.R:333-352). */
typedef struct { epx_session *one; epx_multi *many; } handle;

static void session_finalizer(SEXP ptr)
{
    handle *h = (handle *)R_ExternalPtrAddr(ptr);
    if (h) {
        if (h->one) epx_session_destroy(h->one);
        if (h->many) epx_multi_destroy(h->many);
        free(h);
        R_ClearExternalPtr(ptr);
    }
}

static handle *get_handle(SEXP session)
{
    if (TYPEOF(session) != EXTPTRSXP) Rf_error("libepx: session must be the external pointer returned by C_epx_upload");
    handle *h = (handle *)R_ExternalPtrAddr(session);
    if (!h || (!h->one && !h->many)) Rf_error("libepx: the session has been released");
    return h;
}

/* a REALSXP of whole numbers (offsets: pair counts may pass 2^31 in principle) -> int64 in R_alloc memory */
static int64_t *offsets_from(SEXP v, R_xlen_t len, const char *what)
{
    if (TYPEOF(v) != REALSXP && TYPEOF(v) != INTSXP) Rf_error("libepx: %s must be a numeric vector", what);
    if (XLENGTH(v) != len) Rf_error("libepx: %s has %ld elements, expected %ld", what, (long)XLENGTH(v), (long)len);
    int64_t *out = (int64_t *)R_alloc((size_t)len, sizeof(int64_t));
    for (R_xlen_t i = 0; i < len; i++) {
        const double d = TYPEOF(v) == REALSXP ? REAL(v)[i] : (INTEGER(v)[i] == NA_INTEGER ? -1.0 : (double)INTEGER(v)[i]);
        if (!(d >= 0.0) || d > 9007199254740992.0 || d != (double)(int64_t)d) Rf_error("libepx: %s[%ld] is not a non-negative whole number", what, (long)i + 1);
        out[i] = (int64_t)d;
        if (i > 0 && out[i] < out[i - 1]) Rf_error("libepx: %s is not non-decreasing at element %ld", what, (long)i + 1);
    }
    if (len > 0 && out[0] != 0) Rf_error("libepx: %s must start at 0", what);
    return out;
}

/* .Call(C_epx_upload, methylation_data_frame_columns, device) -> external pointer holding the device copy/copies.
 * methCols: VECSXP of n numeric columns of length P (the data.frame itself, minus its name column): no as.matrix(),
 * so the 9,000-sample shape never needs an R long vector. Integer / logical columns (read.csv gives integer columns
 * for 0/1 data) are converted. device: INTSXP, one or several GPUs. */
SEXP C_epx_upload(SEXP methCols, SEXP device)
{
    char err[512] = "";
    if (TYPEOF(methCols) != VECSXP || XLENGTH(methCols) < 3) Rf_error("libepx: methylation must be a list of at least 3 numeric columns");
    if (TYPEOF(device) != INTSXP || XLENGTH(device) < 1 || XLENGTH(device) > 64) Rf_error("libepx: device must be an integer vector");
    const R_xlen_t n = XLENGTH(methCols);
    const int nd = (int)XLENGTH(device);
    for (int i = 0; i < nd; i++)
        if (INTEGER(device)[i] == NA_INTEGER || INTEGER(device)[i] < 0) Rf_error("libepx: device[%d] is not a GPU number", i + 1);
    const R_xlen_t P = XLENGTH(VECTOR_ELT(methCols, 0));
    if (P < 1) Rf_error("libepx: the methylation frame has no rows");
    const double **cols = (const double **)R_alloc((size_t)n, sizeof(double *));
    for (R_xlen_t j = 0; j < n; j++) {
        SEXP c = VECTOR_ELT(methCols, j);
        if (XLENGTH(c) != P) Rf_error("libepx: column %ld has %ld rows, column 1 has %ld", (long)j + 1, (long)XLENGTH(c), (long)P);
        if (TYPEOF(c) == REALSXP) cols[j] = REAL(c);
        else if (TYPEOF(c) == INTSXP || TYPEOF(c) == LGLSXP) {
            double *d = (double *)R_alloc((size_t)P, sizeof(double));  /* as.double(), NA_integer_ -> NA_real_ */
            const int *v = INTEGER(c);
            for (R_xlen_t i = 0; i < P; i++) d[i] = v[i] == NA_INTEGER ? NA_REAL : (double)v[i];
            cols[j] = d;
        } else Rf_error("libepx: column %ld is not numeric", (long)j + 1);
    }
    /* the external pointer and its finalizer exist before any device state does: whatever unwinds from here on
     * (an R error, an interrupt) leaves the clean-up to the finalizer */
    handle *h = (handle *)calloc(1, sizeof(handle));
    if (!h) Rf_error("libepx: out of memory");
    SEXP ptr = PROTECT(R_MakeExternalPtr(h, R_NilValue, R_NilValue));
    R_RegisterCFinalizerEx(ptr, session_finalizer, TRUE);
    int rc;
    if (nd == 1) {
        rc = epx_session_create(INTEGER(device)[0], &h->one, err, sizeof err);
        if (rc == EPX_OK) rc = epx_session_set_methylation_cols(h->one, cols, (int64_t)n, (int64_t)P, err, sizeof err);
    } else {
        rc = epx_multi_create(INTEGER(device), nd, &h->many, err, sizeof err);
        if (rc == EPX_OK) rc = epx_multi_set_methylation_cols(h->many, cols, (int64_t)n, (int64_t)P, err, sizeof err);
    }
    if (rc != EPX_OK) {
        session_finalizer(ptr); /* release the GPU memory now rather than at the next garbage collection */
        Rf_error("libepx: %s", err);
    }
    UNPROTECT(1);
    return ptr;
}

/* .Call(C_epx_analyze, session, expr, rowPtr, probeIdx, opts)
 * expr: numeric matrix n x G (column-major = one gene after the other, n samples each, original sample order);
 * rowPtr: numeric G+1 offsets; probeIdx: INTSXP of rowPtr[G+1] entries, 0-based row of the methylation frame or
 * NA_INTEGER (annotated probe without data); opts: INTSXP c(linear, testExpr, testMeth, fc_from_means, strict_n3).
 * Returns list(pair = 11 x Np, gene = 15 x G, perm = n x G (1-based), ksDnum = 3 x Np, counts = c(nUP, nM, nD, refGene)). */
SEXP C_epx_analyze(SEXP session, SEXP expr, SEXP rowPtr, SEXP probeIdx, SEXP opts)
{
    char err[512] = "";
    handle *h = get_handle(session);
    if (TYPEOF(expr) != REALSXP) Rf_error("libepx: expr must be a double matrix (use storage.mode(x) <- 'double')");
    SEXP dim = Rf_getAttrib(expr, R_DimSymbol);
    if (TYPEOF(dim) != INTSXP || XLENGTH(dim) != 2) Rf_error("libepx: expr must be a numeric matrix");
    const int64_t n = INTEGER(dim)[0], G = INTEGER(dim)[1];
    if (n < 3 || G < 1) Rf_error("libepx: expr must have at least 3 rows (samples) and 1 column (gene)");
    if (TYPEOF(opts) != INTSXP && TYPEOF(opts) != LGLSXP) Rf_error("libepx: opts must be an integer vector");
    if (XLENGTH(opts) != 5) Rf_error("libepx: opts must hold 5 values (linear, testExpr, testMeth, fc_from_means, strict)");
    for (int i = 0; i < 5; i++)
        if (INTEGER(opts)[i] == NA_INTEGER) Rf_error("libepx: opts[%d] is NA", i + 1);
    const int64_t *rp = offsets_from(rowPtr, (R_xlen_t)G + 1, "rowPtr");
    const int64_t Np = rp[G];
    if (Np > INT_MAX / 2) Rf_error("libepx: %.0f pairs are more than one call can take", (double)Np);
    if (TYPEOF(probeIdx) != INTSXP) Rf_error("libepx: probeIdx must be an integer vector");
    if ((int64_t)XLENGTH(probeIdx) != Np) Rf_error("libepx: probeIdx has %ld entries, rowPtr ends at %.0f", (long)XLENGTH(probeIdx), (double)Np);
    int32_t *pi = (int32_t *)R_alloc((size_t)(Np > 0 ? Np : 1), sizeof(int32_t));
    for (int64_t j = 0; j < Np; j++) {
        const int v = INTEGER(probeIdx)[j];
        if (v != NA_INTEGER && v < 0) Rf_error("libepx: probeIdx[%ld] is negative", (long)j + 1);
        pi[j] = v == NA_INTEGER ? -1 : v;
    }
    uint16_t *pna = (uint16_t *)R_alloc((size_t)(Np > 0 ? Np : 1), sizeof(uint16_t));
    uint16_t *gna = (uint16_t *)R_alloc((size_t)G, sizeof(uint16_t));
    uint16_t *perm16 = (uint16_t *)R_alloc((size_t)(G * n), sizeof(uint16_t));

    SEXP out = PROTECT(Rf_allocVector(VECSXP, 5)); /* the parts are protected by being stored in it right away */
    SEXP pair = Rf_allocMatrix(REALSXP, EPX_PAIR_COLS, (int)Np);   /* 11 x Np, column = pair */
    SET_VECTOR_ELT(out, 0, pair);
    SEXP gene = Rf_allocMatrix(REALSXP, EPX_GENE_COLS, (int)G);
    SET_VECTOR_ELT(out, 1, gene);
    SEXP perm = Rf_allocMatrix(INTSXP, (int)n, (int)G);
    SET_VECTOR_ELT(out, 2, perm);
    SEXP dnum = Rf_allocMatrix(INTSXP, 3, (int)Np);
    SET_VECTOR_ELT(out, 3, dnum);
    SEXP counts = Rf_allocVector(INTSXP, 4);
    SET_VECTOR_ELT(out, 4, counts);
    SEXP names = PROTECT(Rf_allocVector(STRSXP, 5));
    static const char *const nm[5] = { "pair", "gene", "perm", "ksDnum", "counts" };
    for (int i = 0; i < 5; i++) SET_STRING_ELT(names, i, Rf_mkChar(nm[i]));
    Rf_setAttrib(out, R_NamesSymbol, names);
    UNPROTECT(1); /* names: reachable from out now */

    epx_options o;
    epx_options_default(&o);
    o.data_linear = INTEGER(opts)[0] != 0; o.test_expr_t = INTEGER(opts)[1] != 0; o.test_meth_t = INTEGER(opts)[2] != 0;
    o.fc_from_means = INTEGER(opts)[3] != 0; o.strict_n3 = INTEGER(opts)[4] != 0;
    epx_result r;
    memset(&r, 0, sizeof r);
    r.pair = REAL(pair); r.pair_na = pna; r.gene = REAL(gene); r.gene_na = gna;
    r.pair_ks_dnum = INTEGER(dnum); r.perm = perm16;

    /* Ctrl-C: R_CheckUserInterrupt longjmps when an interrupt is pending. Nothing malloc'ed is live here, the
     * session survives in its external pointer, and kernels still in flight only write device memory it owns. */
    R_CheckUserInterrupt();
    int rc;
    if (h->one) {
        /* the pair kernel in slices, with a look at the interrupt flag between them (threading row) */
        const int n_chunks = Np > 200000 ? 8 : 1;
        rc = epx_session_set_expression(h->one, REAL(expr), G, n, 0, err, sizeof err);
        if (rc == EPX_OK) rc = epx_session_set_index(h->one, rp, pi, G, err, sizeof err);
        if (rc == EPX_OK) rc = epx_session_run_prepare(h->one, &o, NULL, err, sizeof err);
        for (int c = 0; rc == EPX_OK && c < n_chunks; c++) {
            rc = epx_session_run_pairs(h->one, c, n_chunks, NULL, err, sizeof err);
            if (rc == EPX_OK && c + 1 < n_chunks) {
                rc = epx_session_wait(h->one, NULL, err, sizeof err);
                if (rc == EPX_OK) R_CheckUserInterrupt();
            }
        }
        if (rc == EPX_OK) rc = epx_session_fetch(h->one, &r, NULL, err, sizeof err);
    } else {
        rc = epx_multi_analyze(h->many, REAL(expr), G, n, rp, pi, &o, &r, err, sizeof err);
    }
    if (rc != EPX_OK) Rf_error("libepx: %s", err);
    R_CheckUserInterrupt();

    /* NA convention: NaN + mask from the engine -> R's NA_real_ (a genuine NaN, e.g. a 0/0 fold change, stays NaN) */
    for (int64_t j = 0; j < Np; j++)
        for (int c = 0; c < EPX_PAIR_COLS; c++)
            if ((pna[j] >> c) & 1) REAL(pair)[j * EPX_PAIR_COLS + c] = NA_REAL;
    for (int64_t g = 0; g < G; g++)
        for (int c = 0; c < EPX_GENE_COLS; c++)
            if ((gna[g] >> c) & 1) REAL(gene)[g * EPX_GENE_COLS + c] = NA_REAL;
    for (int64_t i = 0; i < G * n; i++) INTEGER(perm)[i] = (int)perm16[i] + 1; /* 1-based for R */
    for (int64_t i = 0; i < Np * 3; i++)
        if (INTEGER(dnum)[i] < 0) INTEGER(dnum)[i] = NA_INTEGER;               /* no KS test ran */
    for (int c = 0; c < 3; c++) INTEGER(counts)[c] = r.counts[c];
    INTEGER(counts)[3] = r.ref_gene + 1;
    UNPROTECT(1);
    return out;
}

/* .Call(C_epx_groups, session, groupPtr, groupPairs): pooled tables (Loops B, C, D of the reference,
 * R/EpimethexAnalysis.R:476-620). groupPtr: numeric Q+1 offsets; groupPairs: INTSXP, 0-based rows of dfCGunique
 * (same numbering as probeIdx), each group inside one gene and with data. Returns an 11 x Q matrix. */
SEXP C_epx_groups(SEXP session, SEXP groupPtr, SEXP groupPairs)
{
    char err[512] = "";
    handle *h = get_handle(session);
    if (XLENGTH(groupPtr) < 1) Rf_error("libepx: groupPtr must hold at least one offset");
    const int64_t Q = (int64_t)XLENGTH(groupPtr) - 1;
    if (Q > INT_MAX / 2) Rf_error("libepx: too many groups");
    const int64_t *gp = offsets_from(groupPtr, (R_xlen_t)Q + 1, "groupPtr");
    if (TYPEOF(groupPairs) != INTSXP) Rf_error("libepx: groupPairs must be an integer vector");
    if ((int64_t)XLENGTH(groupPairs) != gp[Q]) Rf_error("libepx: groupPairs has %ld entries, groupPtr ends at %.0f", (long)XLENGTH(groupPairs), (double)gp[Q]);
    for (int64_t j = 0; j < gp[Q]; j++)
        if (INTEGER(groupPairs)[j] == NA_INTEGER || INTEGER(groupPairs)[j] < 0) Rf_error("libepx: groupPairs[%ld] is not a row of dfCGunique", (long)j + 1);
    uint16_t *na = (uint16_t *)R_alloc((size_t)(Q > 0 ? Q : 1), sizeof(uint16_t));
    SEXP grp = PROTECT(Rf_allocMatrix(REALSXP, EPX_PAIR_COLS, (int)Q));
    epx_group_result r;
    memset(&r, 0, sizeof r);
    r.group = REAL(grp); r.group_na = na;
    R_CheckUserInterrupt();
    int rc;
    if (h->one) {
        rc = epx_session_run_groups(h->one, gp, (const int32_t *)INTEGER(groupPairs), Q, NULL, err, sizeof err);
        if (rc == EPX_OK) rc = epx_session_fetch_groups(h->one, &r, NULL, err, sizeof err);
    } else {
        rc = epx_multi_groups(h->many, gp, (const int32_t *)INTEGER(groupPairs), Q, &r, err, sizeof err);
    }
    if (rc != EPX_OK) Rf_error("libepx: %s", err);
    for (int64_t q = 0; q < Q; q++)
        for (int c = 0; c < EPX_PAIR_COLS; c++)
            if ((na[q] >> c) & 1) REAL(grp)[q * EPX_PAIR_COLS + c] = NA_REAL;
    UNPROTECT(1);
    return grp;
}

/* .Call(C_epx_release, session): free the device copy now instead of at garbage collection */
SEXP C_epx_release(SEXP session)
{
    if (TYPEOF(session) != EXTPTRSXP) Rf_error("libepx: session must be the external pointer returned by C_epx_upload");
    session_finalizer(session);
    return R_NilValue;
}

SEXP C_epx_device_info(void)
{
    char err[512] = "", name[256] = "";
    int sm = 0; int64_t mem = 0;
    if (epx_device_info(0, name, sizeof name, &sm, &mem, err, sizeof err) != EPX_OK) Rf_error("libepx: %s", err);
    SEXP out = PROTECT(Rf_allocVector(STRSXP, 1));
    SET_STRING_ELT(out, 0, Rf_mkChar(name));
    UNPROTECT(1);
    return out;
}

/* R's DL_FUNC is `void *(*)(void)`; going through void (*)(void) is the cast -Wcast-function-type accepts */
#define EPX_CALLDEF(fn, nargs) { #fn, (DL_FUNC)(void (*)(void))&fn, nargs }
static const R_CallMethodDef call_methods[] = {
    EPX_CALLDEF(C_epx_upload, 2),
    EPX_CALLDEF(C_epx_analyze, 5),
    EPX_CALLDEF(C_epx_groups, 3),
    EPX_CALLDEF(C_epx_release, 1),
    EPX_CALLDEF(C_epx_device_info, 0),
    { NULL, NULL, 0 }
};

void R_init_EpiMethEx(DllInfo *dll)
{
    R_registerRoutines(dll, NULL, call_methods, NULL, NULL);
    R_useDynamicSymbols(dll, FALSE);
}
