/*
 * epx.h — C ABI of libepx.so, the B200-native engine behind
 *
 *     epimethex.analysis(dfExpressions, dfGpl, dfMethylation, minRangeGene, maxRangeGene,
 *                        numCores, dataGenesLinear, testExpression, testMethylation)
 *
 * (reference: R/EpimethexAnalysis.R:70-72, exported at NAMESPACE:3).  The reference has no native
 * boundary at all (pure R; no src/, no useDynLib), so every entry point below is NEW: it is what an R
 * `.Call` shim (r_package/src/epx_rshim.c) binds to replace the R-level loops.
 * Plain C: pointers, sizes and POD structs only; no torch / R / C++ types.
 *
 * What each entry point replaces in the reference:
 *   epx_session_set_methylation_*   R/EpimethexAnalysis.R:267-296  (methylation frame -> lookup table)
 *   epx_session_set_expression      R/EpimethexAnalysis.R:103-108  (transpose to samples x genes)
 *   epx_session_set_index           R/EpimethexAnalysis.R:287-291  (dfCGunique as a CSR gene->probe index)
 *   epx_session_run (_prepare/_pairs) R/EpimethexAnalysis.R:117-206  (order/quantiles/bin.var/FC/gene tests),
 *                                   :299-311 (gather), :351-453 (Loop A: Analysis + corr.test),
 *                                   R/Functions.R:2-22,26-80,85-132,192-220
 *   epx_session_fetch               the `.combine = cbind` fold of foreach (R/EpimethexAnalysis.R:352)
 *   epx_analyze                     all of the above in one blocking call with host buffers
 *   epx_analyze_host                the same plus R/EpimethexAnalysis.R:278-282 (only the annotated probes' rows are uploaded)
 *
 * Error convention: every function returns 0 on success or a negative EPX_E* code and, when `err` is
 * non-NULL, writes a NUL-terminated message into err[0..errlen).  The library never calls into R, never
 * aborts, and has NO CPU fallback: a missing GPU / failed CUDA call is an error (EPX_ECUDA).
 *
 * NA convention: values the reference would report as NA are returned as quiet NaN AND flagged in a
 * per-row bit mask (bit c set = column c is NA), so a caller can tell NA from a genuine NaN
 * (e.g. linear fold change 0/0, R/Functions.R:75).
 */
#ifndef EPX_H
#define EPX_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define EPX_ABI_VERSION 1

enum {
    EPX_OK = 0,
    EPX_EINVAL = -1,   /* bad argument / unsupported shape */
    EPX_ECUDA = -2,    /* CUDA runtime failure or no usable device */
    EPX_ENOMEM = -3,   /* host or device allocation failed */
    EPX_ESTATE = -4,   /* call sequence error (e.g. run before set_index) */
    EPX_EBREAKS = -5,  /* bin.var: 'breaks' are not unique (the reference stops with this error) */
    EPX_ESTRICT = -6   /* strict_n3 set and the sample count is not a multiple of 3 */
};

/* per-pair table columns: rows 1..11 of setRowNames (R/Functions.R:224-226) */
enum {
    EPX_MEDIAN_DOWN = 0, EPX_MEDIAN_MEDIUM, EPX_MEDIAN_UP,
    EPX_BD_UP_MID, EPX_BD_UP_DOWN, EPX_BD_MID_DOWN,
    EPX_P_UP_MID, EPX_P_UP_DOWN, EPX_P_MID_DOWN,
    EPX_PEARSON_R, EPX_PEARSON_P,
    EPX_PAIR_COLS = 11
};

/* per-gene table columns; the first six are valExprGene (R/EpimethexAnalysis.R:313-315), i.e. rows
 * 12..17 of setRowNames (R/Functions.R:227-228) */
enum {
    EPX_G_FC_UP_MID = 0, EPX_G_FC_UP_DOWN, EPX_G_FC_MID_DOWN,
    EPX_G_P_UP_MID, EPX_G_P_UP_DOWN, EPX_G_P_MID_DOWN,
    EPX_G_STAT_UP_MID, EPX_G_STAT_UP_DOWN, EPX_G_STAT_MID_DOWN, /* Welch t, or KS D */
    EPX_G_Q33, EPX_G_Q66, EPX_G_Q99,
    EPX_G_MEAN_DOWN, EPX_G_MEAN_MEDIUM, EPX_G_MEAN_UP,
    EPX_GENE_COLS = 15
};

typedef struct epx_options {
    int32_t struct_size;    /* = sizeof(epx_options); lets the struct grow compatibly */
    int32_t data_linear;    /* dataGenesLinear:  1 linear fold change, 0 log2 (R/Functions.R:34-42) */
    int32_t test_expr_t;    /* testExpression:   1 Welch t-test, 0 Kolmogorov-Smirnov */
    int32_t test_meth_t;    /* testMethylation:  1 Welch t-test, 0 Kolmogorov-Smirnov */
    int32_t fc_from_means;  /* 0 (default) = reference behaviour, FC of q99/q66/q33 (calcFC reads the
                               quantile rows, R/Functions.R:29-31 after R/EpimethexAnalysis.R:159);
                               1 = FC of the group means (the presumably intended behaviour) */
    int32_t strict_n3;      /* 1 = fail with EPX_ESTRICT when n %% 3 != 0, as the reference's cbind does
                               (R/EpimethexAnalysis.R:323-325,381); 0 = extension: first floor(n/3) ranks UP,
                               next Medium, next Down, the n mod 3 lowest-expression samples unlabelled */
    int32_t reserved[2];
} epx_options;

/* host-side result tables; any pointer may be NULL (that table is then not copied back) */
typedef struct epx_result {
    double   *pair;          /* [n_pairs * 11]  pair-major, columns EPX_MEDIAN_DOWN.. */
    uint16_t *pair_na;       /* [n_pairs]       NA bit mask over the 11 columns */
    double   *pair_stat;     /* [n_pairs * 3]   Welch t or KS D per comparison (UPvsMID, UPvsDOWN, MIDvsDOWN) */
    int32_t  *pair_ks_dnum;  /* [n_pairs * 3]   KS numerator max|c_a*n_b - c_b*n_a| (-1 when no KS ran) */
    double   *gene;          /* [n_genes * 15]  gene-major, columns EPX_G_* */
    uint16_t *gene_na;       /* [n_genes]       NA bit mask over the 15 columns */
    int32_t  *gene_ks_dnum;  /* [n_genes * 3] */
    uint16_t *perm;          /* [n_genes * n_samples] sample indices in descending-expression order */
    int32_t   counts[3];     /* expression-side label sizes {nUP, nM, nD} from bin.var on the reference gene */
    int32_t   ref_gene;      /* index of that gene (first column with the most distinct values) */
} epx_result;

/* device-time breakdown of the last epx_session_run, milliseconds from CUDA events on the run's stream */
typedef struct epx_timing {
    float total_ms;
    float gene_sort_ms;   /* per-gene order + centring          (stratification kernel family) */
    float gene_tests_ms;  /* reference pick + quantiles/FC/tests */
    float probe_rank_ms;  /* per-probe ascending rank            (segmented sort) */
    float pair_stats_ms;  /* gather + moments + medians + KS     (the HBM-roofline kernel) */
    int64_t n_pairs, n_unique_probes, n_genes, n_samples;
    int64_t algorithmic_bytes; /* N_p*(8n+88) + G*(10n+48), SURVEY 8(d) */
    int32_t kernel_launches;   /* kernels launched by the last run */
    int32_t sort_full_rows;    /* rows (since the session was created) whose bucket placement failed its check and that
                                * went through the full sorting network instead: a performance counter, results are the same */
} epx_timing;

typedef struct epx_session epx_session;

/* ---- library / device ---- */
int  epx_abi_version(void);
int  epx_device_count(int *count, char *err, size_t errlen);
/* name[0..namelen), SM count, global memory bytes of device `device` */
int  epx_device_info(int device, char *name, size_t namelen, int *sm_count, int64_t *mem_bytes,
                     char *err, size_t errlen);
void epx_options_default(epx_options *opt);

/* pinned host memory for callers that want full-speed H2D (R's own vectors are pageable) */
int  epx_host_alloc(void **ptr, size_t bytes, char *err, size_t errlen);
void epx_host_free(void *ptr);

/* ---- session: device-resident state that survives across gene-range calls ---- */
int  epx_session_create(int device, epx_session **out, char *err, size_t errlen);
void epx_session_destroy(epx_session *s);

/* Methylation, probe-major: row p = rows + p*ld, n_samples doubles (host memory). No NaN allowed
 * (R/EpimethexAnalysis.R:278 na.omit already dropped incomplete probes). */
int  epx_session_set_methylation_rows(epx_session *s, const double *rows, int64_t n_probes,
                                      int64_t n_samples, int64_t ld, char *err, size_t errlen);
/* Methylation, sample-major: cols[j] points at n_probes doubles = one data.frame column (one sample).
 * This is what the R shim passes (no as.matrix: pan-cancer would need an R long vector). Transposed on
 * the device. */
int  epx_session_set_methylation_cols(epx_session *s, const double *const *cols, int64_t n_samples,
                                      int64_t n_probes, char *err, size_t errlen);
/* Methylation that already lives on the device (probe-major, ld >= n_samples). The session borrows the
 * pointer; the caller keeps ownership and must keep it alive. */
int  epx_session_set_methylation_device(epx_session *s, const double *d_rows, int64_t n_probes,
                                        int64_t n_samples, int64_t ld, char *err, size_t errlen);

/* Expression: gene-major, row g = expr + g*n_samples in ORIGINAL sample order, after the host filters
 * of R/EpimethexAnalysis.R:86-114. is_device != 0: `expr` is a device pointer (copied device-to-device). */
int  epx_session_set_expression(epx_session *s, const double *expr, int64_t n_genes, int64_t n_samples,
                                int is_device, char *err, size_t errlen);

/* CSR gene->probe index in dfCGunique order: row_ptr[n_genes+1], probe_idx[row_ptr[n_genes]];
 * probe_idx = methylation row, or -1 for an annotated probe with no data (-> all-NA pair row). */
int  epx_session_set_index(epx_session *s, const int64_t *row_ptr, const int32_t *probe_idx,
                           int64_t n_genes, char *err, size_t errlen);

/* Launch the whole device pipeline asynchronously on `stream` (a cudaStream_t; NULL = the session's own
 * stream). No host<->device copies, no host synchronisation. */
int  epx_session_run(epx_session *s, const epx_options *opt, void *stream, char *err, size_t errlen);
/* The same pipeline in two phases, for callers that overlap the result gather with the computation (multi-GPU):
 * run_prepare launches everything but the pair kernel; run_pairs(chunk, n_chunks) launches the pair kernel for
 * the chunk-th of n_chunks contiguous slices of the pairs; pairs_chunk tells which rows of the pair tables that
 * slice writes. epx_session_run == run_prepare + run_pairs(0, 1). */
int  epx_session_run_prepare(epx_session *s, const epx_options *opt, void *stream, char *err, size_t errlen);
int  epx_session_run_pairs(epx_session *s, int chunk, int n_chunks, void *stream, char *err, size_t errlen);
int  epx_session_pairs_chunk(epx_session *s, int chunk, int n_chunks, int64_t *first_pair, int64_t *n_pairs,
                             char *err, size_t errlen);
/* Block until everything launched so far on `stream` (NULL = the session's own stream) has finished. For callers that
 * cut a run into slices (run_pairs chunks) and want to look at something between them: the R shim polls
 * R_CheckUserInterrupt() there (threading row). */
int  epx_session_wait(epx_session *s, void *stream, char *err, size_t errlen);
/* Wait for `stream` and copy the result tables to the host. */
int  epx_session_fetch(epx_session *s, epx_result *out, void *stream, char *err, size_t errlen);
/* Timing of the last run (synchronises the events). */
int  epx_session_timing(epx_session *s, epx_timing *out, char *err, size_t errlen);
/* Device pointers of the result tables (for a caller that gathers them itself, e.g. NCCL):
 * which: 0 pair, 1 pair_na, 2 pair_stat, 3 pair_ks_dnum, 4 gene, 5 gene_na, 6 gene_ks_dnum, 7 perm */
int  epx_session_result_device_ptr(epx_session *s, int which, void **ptr, int64_t *bytes,
                                   char *err, size_t errlen);

/* ---- several GPUs from ONE process (R is a single process; replaces the PSOCK cluster of
 * R/EpimethexAnalysis.R:333-352 and its `.combine = cbind`) ----
 * One session per device, each driven by its own host thread. Genes are cut into contiguous ranges with equal
 * numbers of pairs; every device holds a replica of the methylation matrix (or, with epx_multi_set_methylation_sliced,
 * only the rows its genes reference) and the whole expression matrix (the
 * gene-level kernels run everywhere, so the bin.var reference gene needs no exchange) and computes its own pairs,
 * which are copied straight into the caller's result tables (the destination is host memory, so no device-to-device
 * gather is needed in this mode; the one-process-per-GPU mode of bench.py gathers over NCCL instead). The expression
 * matrix goes over PCIe to the first device only; the others copy it from there (peer copies behind an event). Result
 * tables in page-locked memory should come from epx_host_alloc (portable: pinned for every device). With EPX_TRACE set
 * in the environment, epx_multi_analyze prints the host-side phase times of every device thread to stderr. */
typedef struct epx_multi epx_multi;
int  epx_multi_create(const int *devices, int n_devices, epx_multi **out, char *err, size_t errlen);
void epx_multi_destroy(epx_multi *m);
int  epx_multi_set_methylation_rows(epx_multi *m, const double *rows, int64_t n_probes, int64_t n_samples,
                                    int64_t ld, char *err, size_t errlen);
/* the same from sample-major columns (what the R shim holds: the data.frame's columns) */
int  epx_multi_set_methylation_cols(epx_multi *m, const double *const *cols, int64_t n_samples, int64_t n_probes,
                                    char *err, size_t errlen);
/* Probe-sliced copies instead of replicas (pan-cancer: 35 GB of methylation, of which a device needs only the rows
 * its own genes reference): cuts the genes of the given index exactly as epx_multi_analyze will, and uploads to each
 * device only the distinct rows of `rows` that its pairs name, so the total host-to-device traffic is about one
 * matrix instead of one per device. epx_multi_analyze must then be called with the same index (a pair naming a row
 * that is not in its device's slice is EPX_ESTATE); it translates probe_idx to slice rows itself. */
int  epx_multi_set_methylation_sliced(epx_multi *m, const double *rows, int64_t n_probes, int64_t n_samples,
                                      int64_t ld, const int64_t *row_ptr, const int32_t *probe_idx,
                                      int64_t n_genes, char *err, size_t errlen);
int  epx_multi_analyze(epx_multi *m, const double *expr, int64_t n_genes, int64_t n_samples,
                       const int64_t *row_ptr, const int32_t *probe_idx, const epx_options *opt,
                       epx_result *out, char *err, size_t errlen);

/* ---- pooled probe groups: the rows of CG_of_a_genes.csv, CG_by_position.csv, CG_by_islands.csv ----
 * Replaces Loops B, C and D of the reference (R/EpimethexAnalysis.R:476-545 and :549-620 with
 * AnalysisIslands_PositionsCG, R/Functions.R:136-188): the probe columns of a group are stacked, the
 * stratification is recycled down the stack, then the same Analysis + corr.test as for one probe.
 * Group q = pairs group_pairs[group_ptr[q] .. group_ptr[q+1]) (indices into the probe_idx of the current
 * index), all of ONE gene and all with a data row (probe_idx >= 0); an empty group gives an all-NA row
 * (R/EpimethexAnalysis.R:536-541). Which groups exist (every gene; (gene, position) / (gene, island) with more
 * than one data-bearing probe, R/Functions.R:153; islands "NA_NA" / "_*" skipped, R/EpimethexAnalysis.R:321) is
 * decided by the caller. Call after epx_session_run / run_prepare of the same expression + index. */
typedef struct epx_group_result {
    double   *group;        /* [n_groups * 11] columns as epx_result.pair */
    uint16_t *group_na;     /* [n_groups] NA bit mask */
    double   *group_stat;   /* [n_groups * 3] Welch t or KS D */
    int64_t  *group_ks_dnum;/* [n_groups * 3] KS numerator over (k*m)^2, -1 when no KS test ran */
} epx_group_result;
int  epx_session_run_groups(epx_session *s, const int64_t *group_ptr, const int32_t *group_pairs,
                            int64_t n_groups, void *stream, char *err, size_t errlen);
int  epx_session_fetch_groups(epx_session *s, epx_group_result *out, void *stream, char *err, size_t errlen);

/* The pooled tables on several GPUs (one process): call after epx_multi_analyze with pair indices of THAT index. A
 * group lies inside one gene, so it runs on the device that owns the gene; the rows come back in group order. Blocking:
 * runs and fetches. */
int  epx_multi_groups(epx_multi *m, const int64_t *group_ptr, const int32_t *group_pairs, int64_t n_groups,
                      epx_group_result *out, char *err, size_t errlen);

/* ---- result files: R's write.csv, formatted on several host threads (no GPU involved) ----
 * Replaces write.csv(...) at R/EpimethexAnalysis.R:470, :545, :577, :615. values [n_rows][n_cols] row-major;
 * na (optional) [n_rows][n_cols] non-zero = NA, NULL = every NaN is NA; text_col (optional) one more, quoted,
 * column of strings (the `island` row of CG_data_individually, which also makes R quote the numbers:
 * quote_numbers = 1). Numbers are rendered like R does with 15 significant digits (epx_format_real15). */
int  epx_write_csv(const char *path, int64_t n_rows, int n_cols, const char *const *row_names,
                   const char *const *col_names, const double *values, const uint8_t *na,
                   const char *const *text_col, const char *text_col_name, int quote_numbers, int threads,
                   char *err, size_t errlen);
int  epx_format_real15(double x, char *buf, size_t buflen);

/* ---- one-shot: what `.Call(C_epx_analyze, ...)` needs ----
 * Uploads expression + index, runs, fetches. The session must already hold the methylation matrix. */
int  epx_analyze(epx_session *s, const double *expr, int64_t n_genes, int64_t n_samples,
                 const int64_t *row_ptr, const int32_t *probe_idx, const epx_options *opt,
                 epx_result *out, char *err, size_t errlen);

/* The same with the methylation matrix still in HOST memory (probe-major rows as in epx_session_set_methylation_rows):
 * only the rows the index names cross the bus. Replaces, on top of epx_analyze, the methylation filter of
 * R/EpimethexAnalysis.R:278-282 (`subset(dfMethylation, sample %in% dfAnnotations$cg)`: the reference keeps the rows of
 * the annotated probes of the gene range and nothing else) together with the upload: when `meth_rows` is page-locked
 * (epx_host_alloc, cudaHostAlloc, cudaHostRegister) a kernel fetches exactly those rows over PCIe from the caller's
 * memory into the session's device matrix -- a 1,000-gene range of the SKCM shape moves ~5 % of the matrix, the whole
 * genome 75 % (the probes no gene names stay behind). Pageable `meth_rows`: the device cannot read them in place, the
 * call is epx_session_set_methylation_rows + epx_analyze (same tables). NaN / infinite values are refused (EPX_EINVAL)
 * in the rows the index names; other rows are never looked at, as in the reference, where they are dropped before any
 * arithmetic. Afterwards the session holds a PARTIAL matrix: epx_session_run* / run_groups / fetch work on this index,
 * epx_session_set_index / epx_analyze with another index answer EPX_ESTATE until a whole matrix is uploaded
 * (epx_session_set_methylation_*) -- or call epx_analyze_host again, which is the intended use: one call per gene
 * range (README.md:90), nothing kept in between but the allocations. */
int  epx_analyze_host(epx_session *s, const double *meth_rows, int64_t n_probes, int64_t ld,
                      const double *expr, int64_t n_genes, int64_t n_samples,
                      const int64_t *row_ptr, const int32_t *probe_idx, const epx_options *opt,
                      epx_result *out, char *err, size_t errlen);
/* The same on several GPUs from one process (epx_multi_analyze with the matrix still in host memory): the genes are cut
 * into ranges of equal pair counts as in epx_multi_analyze, and every device fetches, over its own PCIe link, only the
 * rows ITS range names -- the whole job moves about one copy of the named rows, spread over the links. `meth_rows` must
 * be page-locked for every device (epx_host_alloc is portable); pageable memory: replicas through
 * epx_multi_set_methylation_rows, then epx_multi_analyze. EPX_NO_ZEROCOPY in the environment forces the pageable path
 * of both calls. */
int  epx_multi_analyze_host(epx_multi *m, const double *meth_rows, int64_t n_probes, int64_t ld,
                            const double *expr, int64_t n_genes, int64_t n_samples,
                            const int64_t *row_ptr, const int32_t *probe_idx, const epx_options *opt,
                            epx_result *out, char *err, size_t errlen);

#ifdef __cplusplus
}
#endif
#endif /* EPX_H */
