/*
 * epx_rshim.c — the `.Call` boundary between R and libepx (include/epx.h).
 *
 * NOT COMPILED IN THE BUILD ENVIRONMENT: R, Rinternals.h and libR are absent there, so this
 * file is exercised only by reading; the identical C ABI is driven from Python ctypes in tests/ and
 * bench.py. It is kept short on purpose. Rules it follows:
 *   - inputs belong to R and are read-only; nothing is written through REAL()/INTEGER() of an argument;
 *   - every SEXP allocation is PROTECTed, UNPROTECT is balanced before return;
 *   - libepx never calls R and never longjmps: errors come back as codes + message, everything owned by C
 *     is released, and only THEN Rf_error() is raised;
 *   - no CPU fallback: a missing GPU is an R error.
 *
 * Replaces the numeric part of R/EpimethexAnalysis.R:117-206 and :299-453 and R/Functions.R:2-132,192-220.
 */
#include <R.h>
#include <Rinternals.h>
#include <R_ext/Rdynload.h>
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
    handle *h = (handle *)R_ExternalPtrAddr(session);
    if (!h) Rf_error("libepx: the session has been released");
    return h;
}

/* .Call(C_epx_upload, methylation_data_frame_columns, device) -> external pointer holding the device copy/copies.
 * methCols: VECSXP of n REALSXP columns of length P (the data.frame itself, minus its name column):
 * no as.matrix(), so the 9,000-sample shape never needs an R long vector. device: INTSXP, one or several GPUs. */
SEXP C_epx_upload(SEXP methCols, SEXP device)
{
    char err[512] = "";
    const R_xlen_t n = XLENGTH(methCols);
    const int nd = Rf_length(device);
    if (TYPEOF(methCols) != VECSXP || n < 3) Rf_error("methylation must be a list of at least 3 numeric columns");
    if (TYPEOF(device) != INTSXP || nd < 1) Rf_error("device must be an integer vector");
    const R_xlen_t P = XLENGTH(VECTOR_ELT(methCols, 0));
    const double **cols = (const double **)malloc((size_t)n * sizeof(double *));
    handle *h = (handle *)calloc(1, sizeof(handle));
    if (!cols || !h) { free(cols); free(h); Rf_error("out of memory"); }
    for (R_xlen_t j = 0; j < n; j++) {
        SEXP c = VECTOR_ELT(methCols, j);
        if (TYPEOF(c) != REALSXP || XLENGTH(c) != P) { free(cols); free(h); Rf_error("column %ld is not a numeric vector of length %ld", (long)j + 1, (long)P); }
        cols[j] = REAL(c);
    }
    int rc;
    if (nd == 1) {
        rc = epx_session_create(INTEGER(device)[0], &h->one, err, sizeof err);
        if (rc == EPX_OK) rc = epx_session_set_methylation_cols(h->one, cols, (int64_t)n, (int64_t)P, err, sizeof err);
    } else {
        rc = epx_multi_create(INTEGER(device), nd, &h->many, err, sizeof err);
        if (rc == EPX_OK) rc = epx_multi_set_methylation_cols(h->many, cols, (int64_t)n, (int64_t)P, err, sizeof err);
    }
    free(cols);
    if (rc != EPX_OK) {
        if (h->one) epx_session_destroy(h->one);
        if (h->many) epx_multi_destroy(h->many);
        free(h);
        Rf_error("libepx: %s", err);
    }
    SEXP ptr = PROTECT(R_MakeExternalPtr(h, R_NilValue, R_NilValue));
    R_RegisterCFinalizerEx(ptr, session_finalizer, TRUE);
    UNPROTECT(1);
    return ptr;
}

/* .Call(C_epx_analyze, session, expr, rowPtr, probeIdx, opts)
 * expr: REALSXP n x G column-major = gene-major rows of n samples (original sample order);
 * rowPtr: REALSXP G+1 (doubles: pair counts can pass 2^31 in principle); probeIdx: INTSXP, 0-based row of the
 * methylation frame or NA_INTEGER; opts: INTSXP c(linear, testExpr, testMeth, fc_from_means, strict_n3). */
SEXP C_epx_analyze(SEXP session, SEXP expr, SEXP rowPtr, SEXP probeIdx, SEXP opts)
{
    char err[512] = "";
    handle *h = get_handle(session);
    SEXP dim = Rf_getAttrib(expr, R_DimSymbol);
    if (TYPEOF(expr) != REALSXP || Rf_length(dim) != 2) Rf_error("expr must be a numeric matrix");
    const int64_t n = INTEGER(dim)[0], G = INTEGER(dim)[1];
    const int64_t Np = (int64_t)REAL(rowPtr)[G];
    int64_t *rp = (int64_t *)malloc((size_t)(G + 1) * sizeof(int64_t));
    int32_t *pi = (int32_t *)malloc((size_t)(Np > 0 ? Np : 1) * sizeof(int32_t));
    if (!rp || !pi) { free(rp); free(pi); Rf_error("out of memory"); }
    for (int64_t g = 0; g <= G; g++) rp[g] = (int64_t)REAL(rowPtr)[g];
    for (int64_t j = 0; j < Np; j++) pi[j] = INTEGER(probeIdx)[j] == NA_INTEGER ? -1 : INTEGER(probeIdx)[j];

    SEXP pair = PROTECT(Rf_allocMatrix(REALSXP, EPX_PAIR_COLS, (int)Np));   /* 11 x Np, column = pair */
    SEXP gene = PROTECT(Rf_allocMatrix(REALSXP, EPX_GENE_COLS, (int)G));
    SEXP perm = PROTECT(Rf_allocMatrix(INTSXP, (int)n, (int)G));
    SEXP dnum = PROTECT(Rf_allocMatrix(INTSXP, 3, (int)Np));
    SEXP counts = PROTECT(Rf_allocVector(INTSXP, 4));
    uint16_t *pna = (uint16_t *)malloc((size_t)(Np > 0 ? Np : 1) * 2), *gna = (uint16_t *)malloc((size_t)G * 2);
    uint16_t *perm16 = (uint16_t *)malloc((size_t)(G * n) * 2);

    epx_options o;
    epx_options_default(&o);
    o.data_linear = INTEGER(opts)[0]; o.test_expr_t = INTEGER(opts)[1]; o.test_meth_t = INTEGER(opts)[2];
    o.fc_from_means = INTEGER(opts)[3]; o.strict_n3 = INTEGER(opts)[4];
    epx_result r;
    memset(&r, 0, sizeof r);
    r.pair = REAL(pair); r.pair_na = pna; r.gene = REAL(gene); r.gene_na = gna;
    r.pair_ks_dnum = INTEGER(dnum); r.perm = perm16;
    int rc = EPX_ENOMEM;
    if (pna && gna && perm16)
        rc = h->one ? epx_analyze(h->one, REAL(expr), G, n, rp, pi, &o, &r, err, sizeof err)
                    : epx_multi_analyze(h->many, REAL(expr), G, n, rp, pi, &o, &r, err, sizeof err);
    if (rc == EPX_OK) {
        /* NA convention: NaN + mask from the engine -> R's NA_real_ (genuine NaN, e.g. 0/0 fold change, stays NaN) */
        for (int64_t j = 0; j < Np; j++)
            for (int c = 0; c < EPX_PAIR_COLS; c++)
                if ((pna[j] >> c) & 1) REAL(pair)[j * EPX_PAIR_COLS + c] = NA_REAL;
        for (int64_t g = 0; g < G; g++)
            for (int c = 0; c < EPX_GENE_COLS; c++)
                if ((gna[g] >> c) & 1) REAL(gene)[g * EPX_GENE_COLS + c] = NA_REAL;
        for (int64_t i = 0; i < G * n; i++) INTEGER(perm)[i] = (int)perm16[i] + 1; /* 1-based for R */
        for (int c = 0; c < 3; c++) INTEGER(counts)[c] = r.counts[c];
        INTEGER(counts)[3] = r.ref_gene + 1;
    }
    free(rp); free(pi); free(pna); free(gna); free(perm16);
    if (rc != EPX_OK) { UNPROTECT(5); Rf_error("libepx: %s", err); }

    SEXP out = PROTECT(Rf_allocVector(VECSXP, 5));
    SEXP names = PROTECT(Rf_allocVector(STRSXP, 5));
    const char *nm[5] = { "pair", "gene", "perm", "ksDnum", "counts" };
    SEXP vals[5] = { pair, gene, perm, dnum, counts };
    for (int i = 0; i < 5; i++) { SET_VECTOR_ELT(out, i, vals[i]); SET_STRING_ELT(names, i, Rf_mkChar(nm[i])); }
    Rf_setAttrib(out, R_NamesSymbol, names);
    UNPROTECT(7);
    return out;
}

/* .Call(C_epx_groups, session, groupPtr, groupPairs): pooled tables (Loops B, C, D of the reference,
 * R/EpimethexAnalysis.R:476-620). groupPtr: REALSXP Q+1 offsets; groupPairs: INTSXP, 0-based rows of dfCGunique
 * (same numbering as probeIdx), each group inside one gene and with data. Returns an 11 x Q matrix. */
SEXP C_epx_groups(SEXP session, SEXP groupPtr, SEXP groupPairs)
{
    char err[512] = "";
    handle *h = get_handle(session);
    const int64_t Q = (int64_t)Rf_xlength(groupPtr) - 1;
    if (Q < 0) Rf_error("groupPtr must hold at least one offset");
    int64_t *gp = (int64_t *)malloc((size_t)(Q + 1) * sizeof(int64_t));
    uint16_t *na = (uint16_t *)malloc((size_t)(Q > 0 ? Q : 1) * 2);
    if (!gp || !na) { free(gp); free(na); Rf_error("out of memory"); }
    for (int64_t q = 0; q <= Q; q++) gp[q] = (int64_t)REAL(groupPtr)[q];
    SEXP grp = PROTECT(Rf_allocMatrix(REALSXP, EPX_PAIR_COLS, (int)Q));
    epx_group_result r;
    memset(&r, 0, sizeof r);
    r.group = REAL(grp); r.group_na = na;
    int rc;
    if (h->one) {
        rc = epx_session_run_groups(h->one, gp, (const int32_t *)INTEGER(groupPairs), Q, NULL, err, sizeof err);
        if (rc == EPX_OK) rc = epx_session_fetch_groups(h->one, &r, NULL, err, sizeof err);
    } else {
        rc = epx_multi_groups(h->many, gp, (const int32_t *)INTEGER(groupPairs), Q, &r, err, sizeof err);
    }
    if (rc == EPX_OK)
        for (int64_t q = 0; q < Q; q++)
            for (int c = 0; c < EPX_PAIR_COLS; c++)
                if ((na[q] >> c) & 1) REAL(grp)[q * EPX_PAIR_COLS + c] = NA_REAL;
    free(gp); free(na);
    UNPROTECT(1);
    if (rc != EPX_OK) Rf_error("libepx: %s", err);
    return grp;
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

static const R_CallMethodDef call_methods[] = {
    { "C_epx_upload", (DL_FUNC)&C_epx_upload, 2 },
    { "C_epx_analyze", (DL_FUNC)&C_epx_analyze, 5 },
    { "C_epx_groups", (DL_FUNC)&C_epx_groups, 3 },
    { "C_epx_device_info", (DL_FUNC)&C_epx_device_info, 0 },
    { NULL, NULL, 0 }
};

void R_init_EpiMethEx(DllInfo *dll)
{
    R_registerRoutines(dll, NULL, call_methods, NULL, NULL);
    R_useDynamicSymbols(dll, FALSE);
}
