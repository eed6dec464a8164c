/*
 * epx_oracle.c — CPU restatement of EpiMethEx's association path.  TEST INFRASTRUCTURE ONLY;
 * PARITY UNPINNED (see epx_oracle.h for both statements).  Plain C99 + optional OpenMP.
 *
 * Accumulations use long double because R's mean/var/cor accumulate in LDOUBLE (x87 extended on the
 * x86-64 hosts the reference was tested on, README.md:85-90).
 *
 * Citations "R/..." are into /root/reference.
 */
#include "epx_oracle.h"

#include <math.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

/* ------------------------------------------------------------------ ordering */

/* stable merge sort of indices; cmp_gt(a,b) says "a must come before b" strictly */
static void msort_idx(int32_t *idx, int32_t *tmp, int n, const double *key, int descending)
{
    for (int w = 1; w < n; w *= 2) {
        for (int lo = 0; lo < n; lo += 2 * w) {
            int mid = lo + w < n ? lo + w : n, hi = lo + 2 * w < n ? lo + 2 * w : n;
            int i = lo, j = mid, k = lo;
            while (i < mid && j < hi) {
                double a = key[idx[i]], b = key[idx[j]];
                /* take from the right run only when it is strictly "better": keeps ties stable */
                int take_right = descending ? (b > a) : (b < a);
                tmp[k++] = take_right ? idx[j++] : idx[i++];
            }
            while (i < mid) tmp[k++] = idx[i++];
            while (j < hi) tmp[k++] = idx[j++];
        }
        memcpy(idx, tmp, (size_t)n * sizeof(int32_t));
    }
}

void epo_order_desc(const double *x, int n, int32_t *perm)
{
    int32_t *tmp = (int32_t *)malloc((size_t)(n > 0 ? n : 1) * sizeof(int32_t));
    for (int i = 0; i < n; i++) perm[i] = i;
    msort_idx(perm, tmp, n, x, 1);
    free(tmp);
}

static int cmp_double_asc(const void *a, const void *b)
{
    double x = *(const double *)a, y = *(const double *)b;
    return (x > y) - (x < y);
}

/* ------------------------------------------------------------------ quantiles, bin.var, FC */

double epo_quantile7(const double *asc, int n, double p)
{
    /* stats:::quantile.default, type 7: index <- 1 + (n-1)*p; lo/hi; qs <- x[lo];
     * where index > lo & x[hi] != qs: qs <- (1-h)*qs + h*x[hi] */
    if (n <= 0) return NAN;
    double index = 1.0 + (double)(n - 1) * p;
    double lo = floor(index), hi = ceil(index);
    double qs = asc[(int)lo - 1];
    double xh = asc[(int)hi - 1];
    if (index > lo && xh != qs) {
        double h = index - lo;
        qs = (1.0 - h) * qs + h * xh;
    }
    return qs;
}

int epo_binvar_counts(const double *x_desc, int n, int32_t counts[3])
{
    /* RcmdrMisc::bin.var(method="proportions"):
     *   cut(x, quantile(x, probs = seq(0, 1, 1/bins)), include.lowest = TRUE, labels = c('D','M','UP'))
     * seq(0,1,1/3) = 0, 1/3, 2*(1/3), 1  (3*(1/3) rounds to 1). */
    double *asc = (double *)malloc((size_t)n * sizeof(double));
    for (int i = 0; i < n; i++) asc[i] = x_desc[n - 1 - i];
    const double third = 1.0 / 3.0;
    double b0 = epo_quantile7(asc, n, 0.0);
    double b1 = epo_quantile7(asc, n, third);
    double b2 = epo_quantile7(asc, n, 2.0 * third);
    double b3 = epo_quantile7(asc, n, 1.0);
    free(asc);
    if (!(b0 < b1 && b1 < b2 && b2 < b3)) return -1; /* cut(): 'breaks' are not unique */
    int nup = 0, nm = 0, nd = 0;
    for (int i = 0; i < n; i++) {
        double v = x_desc[i];
        /* .bincode right=TRUE, include.lowest=TRUE: (b_i, b_{i+1}], lowest closed */
        if (v > b2) nup++;
        else if (v > b1) nm++;
        else nd++;
    }
    counts[0] = nup; counts[1] = nm; counts[2] = nd;
    return 0;
}

static double r_sign(double v) { return (v > 0) - (v < 0); }

double epo_log_fc(double a, double b)
{
    /* R/Functions.R:59-63 */
    if (isnan(a) || isnan(b)) return NAN;
    double fold = pow(2.0, fabs(a - b));
    return (b < a) ? fold : -fold;
}

double epo_linear_fc(double a, double b)
{
    /* R/Functions.R:68-79 */
    if (isnan(a) || isnan(b)) return NAN;
    double aa = fabs(a), ab = fabs(b);
    double maxabs = aa > ab ? aa : ab, minabs = aa < ab ? aa : ab;
    double mx = a > b ? a : b, mn = a < b ? a : b;
    double fc = (r_sign(a) == r_sign(b)) ? maxabs / minabs : mx - mn;
    return (b < a) ? fc : -fc;
}

/* ------------------------------------------------------------------ moments (R's LDOUBLE two-pass) */

static double r_mean(const double *x, int n)
{
    /* summary.c real_mean: s = sum/n; then s += sum(x - s)/n */
    long double s = 0.0L;
    for (int i = 0; i < n; i++) s += x[i];
    s /= n;
    long double t = 0.0L;
    for (int i = 0; i < n; i++) t += (x[i] - s);
    s += t / n;
    return (double)s;
}

static double r_var(const double *x, int n, double *mean_out)
{
    /* cov.c cov_complete1: mean as above (stored as double), then sum((x-m)^2)/(n-1) in LDOUBLE */
    double m = r_mean(x, n);
    if (mean_out) *mean_out = m;
    if (n < 2) return NAN;
    long double ss = 0.0L;
    for (int i = 0; i < n; i++) {
        long double d = (long double)x[i] - m;
        ss += d * d;
    }
    return (double)(ss / (n - 1));
}

int epo_guard(const double *a, int la, const double *b, int lb)
{
    /* R/Functions.R:193: sd(mapply('-', array1, array2), na.rm=TRUE); mapply recycles the shorter one.
     * sd == 0 exactly iff all differences are identical (the mean of identical doubles is exact in
     * LDOUBLE for any realistic n, so every deviation is 0). */
    int len = la > lb ? la : lb;
    if (la == 0 || lb == 0 || len < 2) return -1;
    double d0 = a[0] - b[0];
    for (int i = 1; i < len; i++) {
        double d = a[i % la] - b[i % lb];
        if (d != d0) return 1;
    }
    return 0;
}

/* ------------------------------------------------------------------ Student t / incomplete beta */

/* continued fraction for I_x(a,b) (modified Lentz), long double */
static long double betacf_l(long double a, long double b, long double x)
{
    const long double tiny = 1e-4000L, eps = 3e-19L;
    long double qab = a + b, qap = a + 1.0L, qam = a - 1.0L;
    long double c = 1.0L, d = 1.0L - qab * x / qap;
    if (fabsl(d) < tiny) d = tiny;
    d = 1.0L / d;
    long double h = d;
    for (int m = 1; m <= 200000; m++) {
        long double m2 = 2.0L * m;
        long double aa = m * (b - m) * x / ((qam + m2) * (a + m2));
        d = 1.0L + aa * d; if (fabsl(d) < tiny) d = tiny;
        c = 1.0L + aa / c; if (fabsl(c) < tiny) c = tiny;
        d = 1.0L / d; h *= d * c;
        aa = -(a + m) * (qab + m) * x / ((a + m2) * (qap + m2));
        d = 1.0L + aa * d; if (fabsl(d) < tiny) d = tiny;
        c = 1.0L + aa / c; if (fabsl(c) < tiny) c = tiny;
        d = 1.0L / d;
        long double del = d * c;
        h *= del;
        if (fabsl(del - 1.0L) < eps) break;
    }
    return h;
}

/* I_x(a,b) given both x and y = 1-x to full precision */
static long double pbeta_xy_l(long double x, long double y, long double a, long double b)
{
    if (x <= 0.0L) return 0.0L;
    if (y <= 0.0L) return 1.0L;
    long double lbeta = lgammal(a) + lgammal(b) - lgammal(a + b);
    long double front = expl(a * logl(x) + b * logl(y) - lbeta);
    if (x < (a + 1.0L) / (a + b + 2.0L))
        return front * betacf_l(a, b, x) / a;
    return 1.0L - front * betacf_l(b, a, y) / b;
}

double epo_pbeta(double x, double a, double b)
{
    return (double)pbeta_xy_l((long double)x, 1.0L - (long double)x, a, b);
}

double epo_pt(double q, double df)
{
    /* nmath pt(): P(T<=q) = 0.5*I_{df/(df+q^2)}(df/2, 1/2) for q<0, mirrored for q>0 */
    if (isnan(q) || isnan(df) || df <= 0) return NAN;
    if (isinf(q)) return q < 0 ? 0.0 : 1.0;
    long double t2 = (long double)q * q;
    long double x = df / (df + t2), y = t2 / (df + t2);
    long double tail = 0.5L * pbeta_xy_l(x, y, 0.5L * df, 0.5L);
    return (double)(q <= 0 ? tail : 1.0L - tail);
}

static double two_sided_p(double t, double df)
{
    if (isnan(t) || isnan(df)) return NAN;
    if (isinf(t)) return 0.0;
    long double t2 = (long double)t * t;
    long double x = df / (df + t2), y = t2 / (df + t2);
    long double p = pbeta_xy_l(x, y, 0.5L * df, 0.5L);
    if (p > 1.0L) p = 1.0L;
    return (double)p;
}

int epo_welch(const double *a, int la, const double *b, int lb, double *t, double *df, double *p)
{
    /* stats:::t.test.default, var.equal = FALSE, mu = 0, two.sided */
    *t = NAN; *df = NAN; *p = NAN;
    if (la < 2 || lb < 2) return 1; /* "not enough 'x' observations" */
    double mx, my;
    double vx = r_var(a, la, &mx), vy = r_var(b, lb, &my);
    double stderrx = sqrt(vx / la), stderry = sqrt(vy / lb);
    double se = sqrt(stderrx * stderrx + stderry * stderry);
    double d = se * se * se * se /
               (stderrx * stderrx * stderrx * stderrx / (la - 1) + stderry * stderry * stderry * stderry / (lb - 1));
    double amx = fabs(mx), amy = fabs(my);
    if (se < 10.0 * 2.220446049250313e-16 * (amx > amy ? amx : amy)) return 2; /* essentially constant */
    *t = (mx - my) / se;
    *df = d;
    *p = two_sided_p(*t, d);
    return 0;
}

/* ------------------------------------------------------------------ Kolmogorov-Smirnov */

double epo_psmirnov2x(double statistic, int m, int n)
{
    /* stats ks.c psmirnov2x (R 3.x): exact P(D < statistic) for sizes m, n without ties */
    if (m > n) { int t = n; n = m; m = t; }
    double md = (double)m, nd = (double)n;
    double q = (0.5 + floor(statistic * md * nd - 1e-7)) / (md * nd);
    double *u = (double *)malloc((size_t)(n + 1) * sizeof(double));
    for (int j = 0; j <= n; j++) u[j] = ((j / nd) > q) ? 0 : 1;
    for (int i = 1; i <= m; i++) {
        double w = (double)i / ((double)(i + n));
        if ((i / md) > q) u[0] = 0; else u[0] = w * u[0];
        for (int j = 1; j <= n; j++) {
            if (fabs(i / md - j / nd) > q) u[j] = 0;
            else u[j] = w * u[j] + u[j - 1];
        }
    }
    double r = u[n];
    free(u);
    return r;
}

double epo_pkstwo(double x)
{
    /* stats ks.c pkstwo with tol = 1e-6 (as called from ks.test).  For x < 1 only the k = 1 term of
     * the alternative series is summed, because k_max = (int)sqrt(2 - log(tol)) = 3 and k steps by 2.
     * Replicated as is. */
    const double tol = 1e-6;
    if (isnan(x)) return NAN;
    if (x <= 0) return 0.0;
    int k_max = (int)sqrt(2.0 - log(tol));
    if (x < 1.0) {
        const double M_PI_2_ = 1.570796326794896619231321691640, M_PI_4_ = 0.785398163397448309615660845820;
        const double M_1_SQRT_2PI_ = 0.398942280401432677939946059934;
        double z = -(M_PI_2_ * M_PI_4_) / (x * x);
        double w = log(x);
        double s = 0;
        for (int k = 1; k < k_max; k += 2) s += exp(k * k * z - w);
        return s / M_1_SQRT_2PI_;
    }
    double z = -2.0 * x * x, s = -1.0, old = 0.0, nw = 1.0;
    int k = 1;
    while (fabs(old - nw) > tol) {
        old = nw;
        nw += 2.0 * s * exp(z * k * k);
        s *= -1.0;
        k++;
    }
    return nw;
}

typedef struct { double v; int src; } ks_item;
static int cmp_ks_item(const void *a, const void *b)
{
    double x = ((const ks_item *)a)->v, y = ((const ks_item *)b)->v;
    return (x > y) - (x < y);
}

int epo_ks(const double *a, int la, const double *b, int lb, int64_t *dnum, double *p, int *exact)
{
    /* stats:::ks.test two-sample, alternative two.sided, exact = NULL (classic R 3.4-4.0) */
    *dnum = -1; *p = NAN; if (exact) *exact = 0;
    if (la < 1 || lb < 1) return 1; /* "not enough 'x' data" */
    int N = la + lb;
    ks_item *w = (ks_item *)malloc((size_t)N * sizeof(ks_item));
    for (int i = 0; i < la; i++) { w[i].v = a[i]; w[i].src = 0; }
    for (int i = 0; i < lb; i++) { w[la + i].v = b[i]; w[la + i].src = 1; }
    qsort(w, (size_t)N, sizeof(ks_item), cmp_ks_item);
    int64_t ca = 0, cb = 0, best = 0;
    int ties = 0;
    for (int k = 0; k < N; k++) {
        if (w[k].src == 0) ca++; else cb++;
        int run_end = (k == N - 1) || (w[k + 1].v != w[k].v);
        if (!run_end) { ties = 1; continue; }
        /* z = cumsum(+1/n.x | -1/n.y) at the end of each tie run; numerator over n.x*n.y */
        int64_t z = ca * lb - cb * la;
        if (z < 0) z = -z;
        if (z > best) best = z;
    }
    free(w);
    *dnum = best;
    double D = (double)best / ((double)la * (double)lb);
    int use_exact = ((double)la * (double)lb < 10000.0) && !ties;
    double pv;
    if (use_exact) {
        pv = 1.0 - epo_psmirnov2x(D, la, lb);
    } else {
        double nn = (double)la * (double)lb / ((double)la + (double)lb);
        pv = 1.0 - epo_pkstwo(sqrt(nn) * D);
    }
    if (pv > 1.0) pv = 1.0;
    if (pv < 0.0) pv = 0.0;
    *p = pv;
    if (exact) *exact = use_exact;
    return 0;
}

/* ------------------------------------------------------------------ median, Pearson */

double epo_median(const double *x, int n)
{
    if (n <= 0) return NAN;
    double *s = (double *)malloc((size_t)n * sizeof(double));
    memcpy(s, x, (size_t)n * sizeof(double));
    qsort(s, (size_t)n, sizeof(double), cmp_double_asc);
    double r;
    if (n % 2) r = s[n / 2];
    else {
        double two[2] = { s[n / 2 - 1], s[n / 2] };
        r = r_mean(two, 2); /* median.default: mean(sort(x, partial=half+0:1)[half+0:1]) */
    }
    free(s);
    return r;
}

int epo_pearson(const double *x, const double *y, int n, double *r, double *p)
{
    /* cov.c cov_pairwise2 (cor): one-pass LDOUBLE means, then centred sums;
     * psych::corr.test: t = r*sqrt(n-2)/sqrt(1-r^2), p two-sided on n-2 df */
    *r = NAN; *p = NAN;
    if (n < 3) return 1;
    long double xm = 0, ym = 0;
    for (int i = 0; i < n; i++) { xm += x[i]; ym += y[i]; }
    xm /= n; ym /= n;
    long double sxx = 0, syy = 0, sxy = 0;
    for (int i = 0; i < n; i++) {
        long double dx = x[i] - xm, dy = y[i] - ym;
        sxx += dx * dx; syy += dy * dy; sxy += dx * dy;
    }
    if (sxx == 0 || syy == 0) return 1; /* "the standard deviation is zero" -> NA */
    long double n1 = n - 1;
    long double rr = (sxy / n1) / (sqrtl(sxx / n1) * sqrtl(syy / n1));
    if (rr > 1.0L) rr = 1.0L;
    if (rr < -1.0L) rr = -1.0L;
    double rd = (double)rr;
    *r = rd;
    double t = (rd * sqrt((double)(n - 2))) / sqrt(1.0 - rd * rd);
    *p = two_sided_p(t, (double)(n - 2));
    return 0;
}

/* ------------------------------------------------------------------ calculateTtest */

/* R/Functions.R:192-220.  Returns stat (t or D), p, and dnum; na=1 when the reference yields NA
 * (or would stop with an error: defined as NA here). */
static void calc_test(const double *a, int la, const double *b, int lb, int t_mode,
                      double *stat, double *p, int32_t *dnum, int *na)
{
    *stat = NAN; *p = NAN; *dnum = -1; *na = 1;
    int g = epo_guard(a, la, b, lb);
    if (t_mode) {
        if (g != 1) return;                 /* sd == 0 -> c(NA, NA); sd NA -> `if (NA)` error */
        double t, df, pv;
        if (epo_welch(a, la, b, lb, &t, &df, &pv) != 0) return;
        *stat = t; *p = pv; *na = 0;
    } else {
        if (g == 0) return;                 /* difference == 0 and not NA -> NA */
        int64_t dn; double pv;
        if (epo_ks(a, la, b, lb, &dn, &pv, NULL) != 0) return; /* tryCatch -> NA */
        *dnum = (int32_t)dn;
        *stat = (double)dn / ((double)la * (double)lb);
        *p = pv; *na = 0;
    }
}

/* ------------------------------------------------------------------ the path */

int epo_analyze(const double *expr, int64_t n64, int64_t G,
                const double *meth, int64_t meth_ld,
                const int64_t *row_ptr, const int32_t *probe_idx,
                const epo_options *opt,
                double *pair, uint16_t *pair_na, double *pair_stat, int32_t *ks_dnum,
                double *gene, uint16_t *gene_na, int32_t *gene_ks_dnum,
                int32_t *perm_out, int32_t counts_out[3], int32_t *ref_gene_out)
{
    if (!expr || !opt || n64 < 3 || n64 > 32767 || G < 1) return -1;
    const int n = (int)n64;
    int threads = opt->threads > 0 ? opt->threads : 1;
#ifndef _OPENMP
    (void)threads;
#endif

    /* --- R/EpimethexAnalysis.R:117-122: per-gene descending order and sorted values --- */
    int32_t *perm = (int32_t *)malloc((size_t)G * n * sizeof(int32_t));
    double *xs = (double *)malloc((size_t)G * n * sizeof(double)); /* sorted descending */
    int32_t *nuniq = (int32_t *)malloc((size_t)G * sizeof(int32_t));
#pragma omp parallel for num_threads(threads) schedule(dynamic, 16)
    for (int64_t g = 0; g < G; g++) {
        const double *x = expr + g * n;
        int32_t *pg = perm + g * n;
        double *s = xs + g * n;
        epo_order_desc(x, n, pg);
        int u = 1;
        for (int i = 0; i < n; i++) {
            s[i] = x[pg[i]];
            if (i > 0 && s[i] != s[i - 1]) u++;
        }
        nuniq[g] = u;
    }
    if (perm_out) memcpy(perm_out, perm, (size_t)G * n * sizeof(int32_t));

    /* --- :134-139: reference gene = first column with the most distinct values; bin.var on it --- */
    int64_t ref = 0;
    for (int64_t g = 1; g < G; g++) if (nuniq[g] > nuniq[ref]) ref = g;
    int32_t counts[3];
    int rc = epo_binvar_counts(xs + ref * n, n, counts);
    if (ref_gene_out) *ref_gene_out = (int32_t)ref;
    if (rc != 0) { free(perm); free(xs); free(nuniq); return -2; }
    if (counts_out) memcpy(counts_out, counts, sizeof(counts));
    const int nUP = counts[0], nM = counts[1], nD = counts[2];

    /* --- gene level: :125-129 quantiles, :146-155 means, :172 calcFC, :184-206 tests --- */
#pragma omp parallel for num_threads(threads) schedule(dynamic, 16)
    for (int64_t g = 0; g < G; g++) {
        const double *s = xs + g * n;
        double *asc = (double *)malloc((size_t)n * sizeof(double));
        for (int i = 0; i < n; i++) asc[i] = s[n - 1 - i];
        double q33 = epo_quantile7(asc, n, 0.33), q66 = epo_quantile7(asc, n, 0.66),
               q99 = epo_quantile7(asc, n, 0.99);
        free(asc);
        const double *up = s, *mid = s + nUP, *down = s + nUP + nM;
        double mUP = nUP ? r_mean(up, nUP) : NAN, mM = nM ? r_mean(mid, nM) : NAN,
               mD = nD ? r_mean(down, nD) : NAN;
        /* calcFC reads the LAST three rows = perc99 / perc66 / perc33 (R/Functions.R:29-31 after
         * R/EpimethexAnalysis.R:155,159): "meanUp"=q99, "meanMid"=q66, "meanDown"=q33. */
        double fu = opt->fc_from_means ? mUP : q99, fm = opt->fc_from_means ? mM : q66,
               fd = opt->fc_from_means ? mD : q33;
        double fc[3];
        if (opt->data_linear) {
            fc[0] = epo_linear_fc(fu, fm); fc[1] = epo_linear_fc(fu, fd); fc[2] = epo_linear_fc(fm, fd);
        } else {
            fc[0] = epo_log_fc(fu, fm); fc[1] = epo_log_fc(fu, fd); fc[2] = epo_log_fc(fm, fd);
        }
        double st[3], pv[3]; int32_t dn[3]; int na[3];
        calc_test(up, nUP, mid, nM, opt->test_expr_t, &st[0], &pv[0], &dn[0], &na[0]);
        calc_test(up, nUP, down, nD, opt->test_expr_t, &st[1], &pv[1], &dn[1], &na[1]);
        calc_test(mid, nM, down, nD, opt->test_expr_t, &st[2], &pv[2], &dn[2], &na[2]);
        if (gene) {
            double *o = gene + g * EPO_GENE_COLS;
            for (int c = 0; c < 3; c++) { o[EPO_G_FC_UP_MID + c] = fc[c]; o[EPO_G_P_UP_MID + c] = pv[c]; o[EPO_G_STAT_UP_MID + c] = st[c]; }
            o[EPO_G_Q33] = q33; o[EPO_G_Q66] = q66; o[EPO_G_Q99] = q99;
            o[EPO_G_MEAN_DOWN] = mD; o[EPO_G_MEAN_MEDIUM] = mM; o[EPO_G_MEAN_UP] = mUP;
        }
        if (gene_na) {
            uint16_t mask = 0;
            for (int c = 0; c < 3; c++) if (na[c]) mask |= (uint16_t)((1u << (EPO_G_P_UP_MID + c)) | (1u << (EPO_G_STAT_UP_MID + c)));
            gene_na[g] = mask;
        }
        if (gene_ks_dnum) for (int c = 0; c < 3; c++) gene_ks_dnum[g * 3 + c] = dn[c];
    }

    /* --- pair level: :299-302 gather, :323-325 stratification, :351-453 Loop A --- */
    if (row_ptr && probe_idx && (pair || pair_na || pair_stat || ks_dnum)) {
        const int m = n / 3; /* rep(c("UP","Medium","Down"), each = n/3): first m ranks UP, ... */
#pragma omp parallel for num_threads(threads) schedule(dynamic, 4)
        for (int64_t g = 0; g < G; g++) {
            const int32_t *pg = perm + g * n;
            const double *s = xs + g * n;
            double *y = (double *)malloc((size_t)n * sizeof(double));
            for (int64_t j = row_ptr[g]; j < row_ptr[g + 1]; j++) {
                double o[EPO_PAIR_COLS], st[3] = { NAN, NAN, NAN };
                int32_t dn[3] = { -1, -1, -1 };
                uint16_t mask = 0;
                int32_t pi = probe_idx[j];
                if (pi < 0 || !meth) {
                    for (int c = 0; c < EPO_PAIR_COLS; c++) o[c] = NAN;
                    mask = (uint16_t)((1u << EPO_PAIR_COLS) - 1);
                } else {
                    const double *row = meth + (int64_t)pi * meth_ld;
                    for (int i = 0; i < n; i++) y[i] = row[pg[i]]; /* the gather, in the gene's order */
                    const double *up = y, *mid = y + m, *down = y + 2 * m;
                    /* R/Functions.R:86-96 medians, :2-22 beta differences */
                    o[EPO_MEDIAN_DOWN] = epo_median(down, m);
                    o[EPO_MEDIAN_MEDIUM] = epo_median(mid, m);
                    o[EPO_MEDIAN_UP] = epo_median(up, m);
                    o[EPO_BD_UP_MID] = o[EPO_MEDIAN_UP] - o[EPO_MEDIAN_MEDIUM];
                    o[EPO_BD_UP_DOWN] = o[EPO_MEDIAN_UP] - o[EPO_MEDIAN_DOWN];
                    o[EPO_BD_MID_DOWN] = o[EPO_MEDIAN_MEDIUM] - o[EPO_MEDIAN_DOWN];
                    /* R/Functions.R:115-127 */
                    int na[3];
                    calc_test(up, m, mid, m, opt->test_meth_t, &st[0], &o[EPO_P_UP_MID], &dn[0], &na[0]);
                    calc_test(up, m, down, m, opt->test_meth_t, &st[1], &o[EPO_P_UP_DOWN], &dn[1], &na[1]);
                    calc_test(mid, m, down, m, opt->test_meth_t, &st[2], &o[EPO_P_MID_DOWN], &dn[2], &na[2]);
                    for (int c = 0; c < 3; c++) if (na[c]) mask |= (uint16_t)(1u << (EPO_P_UP_MID + c));
                    /* R/EpimethexAnalysis.R:404-417 */
                    if (epo_pearson(s, y, n, &o[EPO_PEARSON_R], &o[EPO_PEARSON_P]) != 0)
                        mask |= (uint16_t)((1u << EPO_PEARSON_R) | (1u << EPO_PEARSON_P));
                }
                if (pair) memcpy(pair + j * EPO_PAIR_COLS, o, sizeof(o));
                if (pair_na) pair_na[j] = mask;
                if (pair_stat) memcpy(pair_stat + j * 3, st, sizeof(st));
                if (ks_dnum) memcpy(ks_dnum + j * 3, dn, sizeof(dn));
            }
            free(y);
        }
    }
    free(perm); free(xs); free(nuniq);
    return 0;
}

/* ------------------------------------------------------------------ pooled groups (Loops B, C, D) */

/*
 * R/EpimethexAnalysis.R:476-545 (CG_of_a_genes: all probes of a gene), :549-620 with R/Functions.R:136-188
 * (CG_by_position / CG_by_islands: the probes of one gene that share a RefGene group / island relation).
 * A group is a list of data-bearing pairs of ONE gene.  Their columns, each in the gene's sample order, are
 * stacked (`stack()`, :504 / R/Functions.R:156) and the n-row stratification frame is recycled down the k*n rows
 * (`cbind`, :508), so probe j contributes its first m = n/3 ranks to UP, the next m to Medium, the next m to Down;
 * then Analysis (R/Functions.R:85-132) and psych::corr.test of rep(sorted expression, k) against the stack
 * (:517-524, R/Functions.R:166-171).
 */
int epo_analyze_groups(const double *expr, int64_t n64, int64_t G,
                       const double *meth, int64_t meth_ld,
                       const int64_t *row_ptr, const int32_t *probe_idx,
                       const epo_options *opt,
                       int64_t n_groups, const int64_t *group_ptr, const int32_t *group_pairs,
                       double *grp, uint16_t *grp_na, double *grp_stat, int64_t *grp_dnum)
{
    if (!expr || !meth || !opt || !row_ptr || !probe_idx || !group_ptr || n64 < 3 || n64 > 32767 || G < 1) return -1;
    const int n = (int)n64, m = n / 3;
    int threads = opt->threads > 0 ? opt->threads : 1;
#ifndef _OPENMP
    (void)threads;
#endif
    int bad = 0;
#pragma omp parallel for num_threads(threads) schedule(dynamic, 1)
    for (int64_t q = 0; q < n_groups; q++) {
        const int64_t k64 = group_ptr[q + 1] - group_ptr[q];
        double o[EPO_PAIR_COLS], st[3] = { NAN, NAN, NAN };
        int64_t dn[3] = { -1, -1, -1 };
        uint16_t mask = 0;
        if (k64 <= 0) { /* :536-541: no probe with data -> a column of NA */
            for (int c = 0; c < EPO_PAIR_COLS; c++) o[c] = NAN;
            mask = (uint16_t)((1u << EPO_PAIR_COLS) - 1);
        } else {
            const int k = (int)k64;
            const int32_t *pl = group_pairs + group_ptr[q];
            /* the gene of the group = the CSR row that holds its first pair */
            int64_t lo = 0, hi = G;
            while (hi - lo > 1) { int64_t mid = (lo + hi) / 2; if (row_ptr[mid] <= pl[0]) lo = mid; else hi = mid; }
            const int64_t g = lo;
            int ok = 1;
            for (int j = 0; j < k; j++)
                if (pl[j] < row_ptr[g] || pl[j] >= row_ptr[g + 1] || probe_idx[pl[j]] < 0) ok = 0;
            if (!ok) {
#pragma omp atomic write
                bad = 1;
                continue;
            }
            int32_t *pg = (int32_t *)malloc((size_t)n * sizeof(int32_t));
            epo_order_desc(expr + g * n, n, pg);
            const size_t M = (size_t)k * m, N = (size_t)k * n;
            double *up = (double *)malloc(3 * M * sizeof(double) + 8), *mid = up + M, *down = mid + M;
            double *xs = (double *)malloc(N * sizeof(double)), *ys = (double *)malloc(N * sizeof(double));
            for (int j = 0; j < k; j++) {
                const double *row = meth + (int64_t)probe_idx[pl[j]] * meth_ld;
                for (int i = 0; i < n; i++) {
                    const double v = row[pg[i]];
                    xs[(size_t)j * n + i] = expr[g * n + pg[i]];
                    ys[(size_t)j * n + i] = v;
                    if (i < m) up[(size_t)j * m + i] = v;
                    else if (i < 2 * m) mid[(size_t)j * m + i - m] = v;
                    else if (i < 3 * m) down[(size_t)j * m + i - 2 * m] = v;
                }
            }
            const int Mi = (int)M;
            o[EPO_MEDIAN_DOWN] = epo_median(down, Mi);
            o[EPO_MEDIAN_MEDIUM] = epo_median(mid, Mi);
            o[EPO_MEDIAN_UP] = epo_median(up, Mi);
            o[EPO_BD_UP_MID] = o[EPO_MEDIAN_UP] - o[EPO_MEDIAN_MEDIUM];
            o[EPO_BD_UP_DOWN] = o[EPO_MEDIAN_UP] - o[EPO_MEDIAN_DOWN];
            o[EPO_BD_MID_DOWN] = o[EPO_MEDIAN_MEDIUM] - o[EPO_MEDIAN_DOWN];
            int na[3]; int32_t d32;
            const double *A[3] = { up, up, mid }, *B[3] = { mid, down, down };
            for (int c = 0; c < 3; c++) {
                /* calc_test keeps a 32-bit numerator; pooled groups need 64 bits, so KS is called directly */
                st[c] = NAN; o[EPO_P_UP_MID + c] = NAN; na[c] = 1;
                const int gd = epo_guard(A[c], Mi, B[c], Mi);
                if (opt->test_meth_t) {
                    calc_test(A[c], Mi, B[c], Mi, 1, &st[c], &o[EPO_P_UP_MID + c], &d32, &na[c]);
                } else if (gd != 0) {
                    int64_t d; double pv;
                    if (epo_ks(A[c], Mi, B[c], Mi, &d, &pv, NULL) == 0) {
                        dn[c] = d; st[c] = (double)d / ((double)Mi * (double)Mi); o[EPO_P_UP_MID + c] = pv; na[c] = 0;
                    }
                }
                if (na[c]) mask |= (uint16_t)(1u << (EPO_P_UP_MID + c));
            }
            if (epo_pearson(xs, ys, (int)N, &o[EPO_PEARSON_R], &o[EPO_PEARSON_P]) != 0)
                mask |= (uint16_t)((1u << EPO_PEARSON_R) | (1u << EPO_PEARSON_P));
            free(pg); free(up); free(xs); free(ys);
        }
        if (grp) memcpy(grp + q * EPO_PAIR_COLS, o, sizeof(o));
        if (grp_na) grp_na[q] = mask;
        if (grp_stat) memcpy(grp_stat + q * 3, st, sizeof(st));
        if (grp_dnum) memcpy(grp_dnum + q * 3, dn, sizeof(dn));
    }
    return bad ? -3 : 0;
}
