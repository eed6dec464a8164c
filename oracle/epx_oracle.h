/*
 * epx_oracle.h — CPU restatement ("oracle") of EpiMethEx's per-gene methylation-vs-expression
 * association path.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing under oracle/ is part of the product: only tests/,
 * __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may build, load or call
 * it, and only as the checker (or as the timed CPU baseline), never as the thing shipped.
 *
 * PARITY UNPINNED.  The reference (giupardeb/EpiMethEx) is pure R, has no tests and no golden
 * vectors, and R is not installable in the build environment, so this restatement could not be
 * pinned against output of the reference itself.  It follows the reference source line by line
 * (citations below are into /root/reference) and restates the published algorithms of the base-R
 * functions the reference calls (stats::t.test, stats::ks.test [R 3.4-4.0 "classic" dialect],
 * stats::quantile type 7, stats::median, stats::sd, stats::cor, psych::corr.test,
 * RcmdrMisc::bin.var).  Every statistic is cross-checked against scipy/mpmath in tests/.
 *
 * Table layouts are shared with include/epx.h (same column order, same NA-mask convention).
 */
#ifndef EPX_ORACLE_H
#define EPX_ORACLE_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* columns of the per-pair table (R/Functions.R:224-226, first 11 names of setRowNames) */
enum {
    EPO_MEDIAN_DOWN = 0, EPO_MEDIAN_MEDIUM, EPO_MEDIAN_UP,
    EPO_BD_UP_MID, EPO_BD_UP_DOWN, EPO_BD_MID_DOWN,
    EPO_P_UP_MID, EPO_P_UP_DOWN, EPO_P_MID_DOWN,
    EPO_PEARSON_R, EPO_PEARSON_P,
    EPO_PAIR_COLS = 11
};
/* columns of the per-gene table: the 6 values of valExprGene (R/EpimethexAnalysis.R:313-315) first */
enum {
    EPO_G_FC_UP_MID = 0, EPO_G_FC_UP_DOWN, EPO_G_FC_MID_DOWN,
    EPO_G_P_UP_MID, EPO_G_P_UP_DOWN, EPO_G_P_MID_DOWN,
    EPO_G_STAT_UP_MID, EPO_G_STAT_UP_DOWN, EPO_G_STAT_MID_DOWN, /* Welch t, or KS D */
    EPO_G_Q33, EPO_G_Q66, EPO_G_Q99,
    EPO_G_MEAN_DOWN, EPO_G_MEAN_MEDIUM, EPO_G_MEAN_UP,
    EPO_GENE_COLS = 15
};

typedef struct {
    int32_t data_linear;   /* dataGenesLinear  (R/EpimethexAnalysis.R:71) */
    int32_t test_expr_t;   /* testExpression: 1 = Welch t, 0 = KS */
    int32_t test_meth_t;   /* testMethylation: 1 = Welch t, 0 = KS */
    int32_t fc_from_means; /* 0 = reference behaviour (FC of q99/q66/q33); 1 = group means */
    int32_t threads;       /* OpenMP threads over genes (the reference's numCores) */
} epo_options;

/* ---- scalar building blocks (each is also exported for unit tests) ---- */

/* stable descending order: perm[i] = index of the i-th largest; ties keep ascending index
 * (R/EpimethexAnalysis.R:117-122: order(decreasing=TRUE) / sort(decreasing=TRUE)); -0 == +0. */
void epo_order_desc(const double *x, int n, int32_t *perm);

/* stats::quantile type 7 on an ASCENDING sorted vector (matrixStats::colQuantiles default,
 * R/EpimethexAnalysis.R:125-126; also used by bin.var's cut points). */
double epo_quantile7(const double *asc, int n, double p);

/* RcmdrMisc::bin.var(x, bins=3, method="proportions") on the reference gene's sorted column
 * (R/EpimethexAnalysis.R:134-139).  x_desc is sorted descending.  counts = {nUP, nM, nD} (row blocks in
 * that order).  Returns 0, or -1 when cut() would fail ("'breaks' are not unique"). */
int epo_binvar_counts(const double *x_desc, int n, int32_t counts[3]);

/* R/Functions.R:59-63 and :68-79 */
double epo_log_fc(double first, double second);
double epo_linear_fc(double first, double second);

/* guard of calculateTtest (R/Functions.R:193): sd(mapply('-', a, b)).  1 = sd != 0, 0 = sd == 0,
 * -1 = sd is NA (fewer than two differences). mapply recycles the shorter argument. */
int epo_guard(const double *a, int la, const double *b, int lb);

/* stats::t.test(a, b, var.equal=FALSE): 0 ok; 1 = not enough observations; 2 = "data are essentially
 * constant".  p = 2*pt(-|t|, df). */
int epo_welch(const double *a, int la, const double *b, int lb, double *t, double *df, double *p);

/* Student t lower tail P(T <= q) for real df>0, and the regularised incomplete beta it is built on. */
double epo_pt(double q, double df);
double epo_pbeta(double x, double a, double b);

/* stats::ks.test(a, b) two-sided, classic dialect.  dnum = max |c_a*lb - c_b*la| evaluated at the end of
 * each tie run (integer numerator of D = dnum/(la*lb)); exact = 1 when psmirnov2x was used
 * (la*lb < 10000 and no ties in the pooled sample), else asymptotic pkstwo. */
int epo_ks(const double *a, int la, const double *b, int lb, int64_t *dnum, double *p, int *exact);
double epo_psmirnov2x(double statistic, int m, int n);  /* P(D < statistic) exact */
double epo_pkstwo(double x);                            /* asymptotic Kolmogorov cdf, tol 1e-6 */

/* stats::median */
double epo_median(const double *x, int n);

/* psych::corr.test(x, y, adjust="none") for one column pair: r (clamped to [-1,1]) and
 * p = 2*pt(-|t|, n-2), t = r*sqrt(n-2)/sqrt(1-r^2).  Returns 1 (r, p = NaN) when either sd is 0 or n<3. */
int epo_pearson(const double *x, const double *y, int n, double *r, double *p);

/* ---- the path ---- */

/*
 * expr      : G rows of n doubles (row g = gene g in ORIGINAL sample order), after the host filters of
 *             R/EpimethexAnalysis.R:86-114 (range slice, all-zero genes dropped, NA samples dropped).
 * meth      : probe-major methylation rows, row p at meth + p*meth_ld, n doubles, no NA
 *             (R/EpimethexAnalysis.R:278 na.omit has dropped incomplete probes).
 * row_ptr   : CSR over genes, G+1 entries; pairs of gene g are [row_ptr[g], row_ptr[g+1]).
 * probe_idx : N_p entries, methylation row of each pair in dfCGunique order
 *             (R/EpimethexAnalysis.R:287-291), or -1 for an annotated probe with no data row
 *             (-> all-NA pair columns, R/EpimethexAnalysis.R:375-394).
 * outputs   : pair[N_p*11], pair_na[N_p] (bit c set = column c is NA), pair_stat[N_p*3] (Welch t or KS D),
 *             ks_dnum[N_p*3] (-1 when no KS ran), gene[G*15], gene_na[G], gene_ks_dnum[G*3],
 *             perm[G*n] (descending-expression sample order), counts[3] = {nUP,nM,nD} of the expression-side
 *             labels, *ref_gene = index of the bin.var reference gene.  Any output pointer may be NULL.
 * returns 0, or a negative error code (-1 bad argument, -2 bin.var breaks not unique).
 */
int epo_analyze(const double *expr, int64_t n, int64_t G,
                const double *meth, int64_t meth_ld,
                const int64_t *row_ptr, const int32_t *probe_idx,
                const epo_options *opt,
                double *pair, uint16_t *pair_na, double *pair_stat, int32_t *ks_dnum,
                double *gene, uint16_t *gene_na, int32_t *gene_ks_dnum,
                int32_t *perm, int32_t counts[3], int32_t *ref_gene);

/*
 * Pooled analyses of probe groups: CG_of_a_genes (R/EpimethexAnalysis.R:476-545), CG_by_position and
 * CG_by_islands (:549-620, R/Functions.R:136-188).  Group q = pairs group_pairs[group_ptr[q] .. group_ptr[q+1]),
 * indices into probe_idx, all of ONE gene and all with a data row (probe_idx >= 0); an empty group yields an
 * all-NA row.  Output tables like the pair tables: grp [n_groups][11], grp_na, grp_stat [.][3] (Welch t or KS D),
 * grp_dnum [.][3] (64-bit KS numerator over (k*m)^2).  Returns -3 when a group mixes genes or names a pair
 * without data.
 */
int epo_analyze_groups(const double *expr, int64_t n, int64_t G,
                       const double *meth, int64_t meth_ld,
                       const int64_t *row_ptr, const int32_t *probe_idx,
                       const epo_options *opt,
                       int64_t n_groups, const int64_t *group_ptr, const int32_t *group_pairs,
                       double *grp, uint16_t *grp_na, double *grp_stat, int64_t *grp_dnum);

#ifdef __cplusplus
}
#endif
#endif
