/*
 * epx_math.h — scalar fp64 math shared by the CUDA kernels and the host part of libepx:
 * Student-t tail (regularised incomplete beta), Kolmogorov-Smirnov p-values as stats::ks.test
 * computes them (classic R 3.4-4.0 dialect), type-7 quantile, fold changes.
 *
 * Everything is EPX_HD (__host__ __device__ under nvcc, plain inline under gcc) so the very same
 * source is compiled into the kernels and into tests/_mathtest (a CPU unit-test build of this header;
 * it is not a fallback path — libepx never computes statistics on the host).
 *
 * Reference call sites (into /root/reference): stats::t.test R/Functions.R:198; stats::ks.test
 * R/Functions.R:210; psych::corr.test R/EpimethexAnalysis.R:414; colQuantiles :125; calcFC R/Functions.R:26-80.
 */
#ifndef EPX_MATH_H
#define EPX_MATH_H

#include <math.h>
#include <stdint.h>

#if defined(__CUDACC__)
#define EPX_HD __host__ __device__ __forceinline__
#else
#define EPX_HD static inline
#endif

#define EPX_NAN (__builtin_nan(""))

/* log B(a, 1/2) = lgamma(a) + lgamma(1/2) - lgamma(a + 1/2).  For a >= 20 the difference of the two
 * lgammas is taken from its asymptotic series (error < 2e-17) instead of cancelling two ~a*log(a) terms. */
EPX_HD double epx_lbeta_half(double a)
{
    const double half_log_pi = 0.57236494292470008707; /* lgamma(1/2) */
    if (a >= 20.0) {
        double r = 1.0 / a, r2 = r * r;
        /* lgamma(a+1/2) - lgamma(a) = log(a)/2 - 1/(8a) + 1/(192a^3) - 1/(640a^5) + 17/(14336a^7) - 31/(18432a^9) */
        double s = r * (-0.125 + r2 * (1.0 / 192.0 + r2 * (-1.0 / 640.0 + r2 * (17.0 / 14336.0 + r2 * (-31.0 / 18432.0)))));
        return half_log_pi - (0.5 * log(a) + s);
    }
    return lgamma(a) + half_log_pi - lgamma(a + 0.5);
}

/* Continued fraction of the incomplete beta function, I_x(a,b) = front/a * betacf(a,b,x), evaluated by the
 * forward (Wallis) recurrence.  Both partial numerators of one double step share the factor 1/(a+2m), so
 * a step costs ONE reciprocal; numerator/denominator are renormalised and the convergence test is run only
 * every 4th step (the recurrences are linear, so scaling both by the same factor changes nothing).
 * fp64 division is ~15 instructions on the GPU: this form is ~5x cheaper than modified Lentz (6 divisions
 * per step) at the same accuracy. */
EPX_HD double epx_betacf(double a, double b, double x)
{
    /* 9 ulp: successive 4-step convergents agree to this => the truncation error is far below it; a tighter
     * bound makes the loop chase rounding noise for hundreds of steps */
    const double eps = 2e-15;
    const double qab = a + b, qap = a + 1.0, qam = a - 1.0;
    double am = 1.0, bm = 1.0, az = 1.0, bz = 1.0 - qab * x / qap;
    double fold = az / bz;
    for (int m = 1; m <= 40000; m++) {
        const double em = (double)m, tem = em + em;
        const double t0 = qam + tem, t1 = a + tem, t2 = qap + tem;
        const double r = x / (t0 * t1 * t2);
        const double de = em * (b - em) * t2 * r;            /* d_{2m}   = m(b-m)x / ((a+2m-1)(a+2m)) */
        const double dodd = -(a + em) * (qab + em) * t0 * r; /* d_{2m+1} = -(a+m)(a+b+m)x / ((a+2m)(a+2m+1)) */
        const double ap = az + de * am, bp = bz + de * bm;
        const double app = ap + dodd * az, bpp = bp + dodd * bz;
        am = ap; bm = bp; az = app; bz = bpp;
        if ((m & 3) == 0) {
            const double rb = 1.0 / bz;
            am *= rb; bm *= rb; az *= rb; bz = 1.0;
            if (fabs(az - fold) <= eps * fabs(az)) return az;
            fold = az;
        }
    }
    return az / bz;
}

/* two-sided Student-t p-value 2*pt(-|t|, df) = I_{df/(df+t^2)}(df/2, 1/2), real df > 0 */
EPX_HD double epx_t_two_sided_p(double t, double df)
{
    if (isnan(t) || isnan(df) || !(df > 0.0)) return EPX_NAN;
    if (isinf(t)) return 0.0;
    double t2 = t * t;
    if (t2 == 0.0) return 1.0;
    double a = 0.5 * df, b = 0.5;
    double den = df + t2;
    double x = df / den, y = t2 / den;           /* y = 1 - x without cancellation */
    double logx = -log1p(t2 / df);               /* log(x) accurately when x is close to 1 */
    double front = exp(a * logx + b * log(y) - epx_lbeta_half(a));
    double p;
    /* Which continued fraction: the textbook switch is y = 1 - x > (b+1)/(a+b+2). For large df the fraction in
     * y (the complement, p = 1 - ...) still converges several times faster up to 4x that point, and p is
     * > 1e-4 there, so the subtraction costs nothing against the 1e-8 relative tolerance. */
    const double ysw = (a >= 20.0 ? 4.0 : 1.0) * (b + 1.0) / (a + b + 2.0);
    if (y > ysw) p = front * epx_betacf(a, b, x) / a;
    else p = 1.0 - front * epx_betacf(b, a, y) / b;
    if (p > 1.0) p = 1.0;
    if (p < 0.0) p = 0.0;
    return p;
}

/* stats::t.test(var.equal=FALSE) from group moments. returns 0 ok, 1 too few observations,
 * 2 "data are essentially constant" (the reference would stop; reported as NA) */
EPX_HD int epx_welch(double mx, double vx, double nx, double my, double vy, double ny,
                     double *t, double *df)
{
    if (nx < 2.0 || ny < 2.0) return 1;
    double sx = sqrt(vx / nx), sy = sqrt(vy / ny);
    double se = sqrt(sx * sx + sy * sy);
    double amx = fabs(mx), amy = fabs(my);
    if (se < 10.0 * 2.220446049250313e-16 * (amx > amy ? amx : amy)) return 2;
    double se2 = se * se, sx2 = sx * sx, sy2 = sy * sy;
    *df = se2 * se2 / (sx2 * sx2 / (nx - 1.0) + sy2 * sy2 / (ny - 1.0));
    *t = (mx - my) / se;
    return 0;
}

/* stats ks.c pkstwo(x, tol=1e-6).  For x < 1 only k = 1 of the alternative series is summed
 * (k_max = (int)sqrt(2 - log(tol)) = 3, k += 2): replicated, not "fixed" (SURVEY 8c-3). */
EPX_HD double epx_pkstwo(double x)
{
    const double tol = 1e-6;
    if (isnan(x)) return EPX_NAN;
    if (x <= 0.0) return 0.0;
    if (x < 1.0) {
        const double pi2_8 = 1.570796326794896619231321691640 * 0.785398163397448309615660845820;
        const double inv_sqrt_2pi = 0.398942280401432677939946059934;
        int k_max = (int)sqrt(2.0 - log(tol));
        double z = -pi2_8 / (x * x), w = log(x), s = 0.0;
        for (int k = 1; k < k_max; k += 2) s += exp(k * k * z - w);
        return s / inv_sqrt_2pi;
    }
    double z = -2.0 * x * x, s = -1.0, old = 0.0, nw = 1.0;
    int k = 1;
    while (fabs(old - nw) > tol) {
        old = nw;
        nw += 2.0 * s * exp(z * k * k);
        s = -s;
        k++;
    }
    return nw;
}

/* asymptotic two-sided KS p-value for numerator dnum = max|c_a*lb - c_b*la| */
EPX_HD double epx_ks_p_asym(double dnum, double la, double lb)
{
    double D = dnum / (la * lb);
    double nn = la * lb / (la + lb);
    double p = 1.0 - epx_pkstwo(sqrt(nn) * D);
    return p > 1.0 ? 1.0 : (p < 0.0 ? 0.0 : p);
}

/* stats ks.c psmirnov2x: exact P(D < statistic), no ties. u = scratch of max(m,n)+1 doubles. */
EPX_HD double epx_psmirnov2x(double statistic, int m, int n, double *u)
{
    if (m > n) { int t = n; n = m; m = t; }
    double md = (double)m, nd = (double)n;
    double q = (0.5 + floor(statistic * md * nd - 1e-7)) / (md * nd);
    for (int j = 0; j <= n; j++) u[j] = ((j / nd) > q) ? 0.0 : 1.0;
    for (int i = 1; i <= m; i++) {
        double w = (double)i / ((double)(i + n));
        if ((i / md) > q) u[0] = 0.0; else u[0] = w * u[0];
        for (int j = 1; j <= n; j++) {
            if (fabs(i / md - j / nd) > q) u[j] = 0.0;
            else u[j] = w * u[j] + u[j - 1];
        }
    }
    return u[n];
}

EPX_HD double epx_ks_p_exact(double dnum, int la, int lb, double *u)
{
    double D = dnum / ((double)la * (double)lb);
    double p = 1.0 - epx_psmirnov2x(D, la, lb, u);
    return p > 1.0 ? 1.0 : (p < 0.0 ? 0.0 : p);
}

/* R/Functions.R:59-63 */
EPX_HD double epx_log_fc(double a, double b)
{
    if (isnan(a) || isnan(b)) return EPX_NAN;
    double fold = exp2(fabs(a - b));
    return (b < a) ? fold : -fold;
}

/* R/Functions.R:68-79 (sign(0) = 0, so a zero and a positive value count as "different sign") */
EPX_HD double epx_linear_fc(double a, double b)
{
    if (isnan(a) || isnan(b)) return EPX_NAN;
    double aa = fabs(a), ab = fabs(b);
    double maxabs = aa > ab ? aa : ab, minabs = aa < ab ? aa : ab;
    double mx = a > b ? a : b, mn = a < b ? a : b;
    int sa = (a > 0.0) - (a < 0.0), sb = (b > 0.0) - (b < 0.0);
    double fc = (sa == sb) ? maxabs / minabs : mx - mn;
    return (b < a) ? fc : -fc;
}

/* h(r) = log p(r) - (df/2) log(1 - r^2), where p(r) = I_{1-r^2}(df/2, 1/2) is the two-sided p-value of a
 * Pearson correlation r on df = n - 2 degrees of freedom (t = r sqrt(df)/sqrt(1-r^2) => t^2/(df+t^2) = r^2).
 * h is smooth and bounded on [0,1] (h(0) = 0, h(1) = -log(a B(a,1/2))), which makes it a good thing to
 * tabulate: libepx builds the table once per sample count on the host and the pair kernel interpolates it
 * (4-point Lagrange, relative error < 1e-10 with the grid of epx_pearson_grid) instead of running a continued
 * fraction per pair. */
EPX_HD double epx_pearson_h(double r, double df)
{
    const double a = 0.5 * df, b = 0.5;
    if (r <= 0.0) return 0.0;
    if (r >= 1.0) return -(log(a) + epx_lbeta_half(a));
    const double y = r * r, x = 1.0 - y;
    if (y > (a >= 20.0 ? 4.0 : 1.0) * (b + 1.0) / (a + b + 2.0)) return b * log(y) - epx_lbeta_half(a) + log(epx_betacf(a, b, x) / a);
    const double logx = log1p(-y);
    const double front = exp(a * logx + b * log(y) - epx_lbeta_half(a));
    return log(1.0 - front * epx_betacf(b, a, y) / b) - a * logx;
}

/* number of grid intervals on r in [0,1]: step in t of at most 0.004 */
EPX_HD int epx_pearson_grid(double df)
{
    double v = ceil(sqrt(df) / 0.004);
    if (v < 1024.0) v = 1024.0;
    if (v > 65536.0) v = 65536.0;
    return (int)v;
}

/* p-value from the table h[0..N] */
EPX_HD double epx_pearson_p_tab(double r, const double *h, int N, double df)
{
    const double ar = fabs(r);
    if (!(ar < 1.0)) return ar == 1.0 ? 0.0 : EPX_NAN;
    const double u = ar * (double)N;
    int i = (int)u;
    if (i < 1) i = 1;
    if (i > N - 2) i = N - 2;
    const double t = u - (double)i;
    const double ym = h[i - 1], y0 = h[i], y1 = h[i + 1], y2 = h[i + 2];
    const double tm = t - 1.0, tp = t + 1.0, t2 = t - 2.0;
    const double c = (-t * tm * t2 * (1.0 / 6.0)) * ym + (tp * tm * t2 * 0.5) * y0 + (-tp * t * t2 * 0.5) * y1 +
                     (tp * t * tm * (1.0 / 6.0)) * y2;
    const double p = exp(c + 0.5 * df * log1p(-ar * ar));
    return p > 1.0 ? 1.0 : p;
}

/* ---- Welch p-values of the pair kernel's t mode by a two-dimensional table -------------------------------------
 * Two groups of m observations each: the Welch-Satterthwaite df lies in [m - 1, 2(m - 1)].  p(t, df) = I_{1-r^2}(df/2, 1/2)
 * with r^2 = t^2 / (df + t^2), i.e. exactly the Pearson p-value function above with a df that varies per pair.  So the
 * same smooth h(r; df) is tabulated on a (df, r) grid -- uniform in r as for the Pearson table, uniform in df with a step
 * of 1/160 of the smallest df, where the 4-point Lagrange error in df, ~0.07 (step/df)^4 (h ~ -(1/2) log df + ...), is
 * about 1e-10 -- and a p-value is a 4 x 4 Lagrange interpolation plus one log1p and one exp: ~300 instructions per
 * value with no loop, where the continued fraction took thousands and made the warp wait for its slowest lane.
 * Built on the device once per sample count (k_build_welch_table). */
EPX_HD void epx_welch_grid(int m, int *N, int *K, double *df0, double *ddf)
{
    const double lo = (double)m - 1.0, hi = 2.0 * ((double)m - 1.0);
    *ddf = lo / 160.0;
    *df0 = lo - *ddf;                                /* one node below the range: the stencil is k-1 .. k+2 */
    *K = (int)ceil((hi - *df0) / *ddf) + 3;
    *N = epx_pearson_grid(hi + 2.0 * *ddf);
}

/* tab[k * ld + i] = epx_pearson_h(i / N, df0 + k * ddf), k < K, i <= N (ld >= N + 1) */
EPX_HD double epx_welch_p_tab(double t, double df, const double *tab, int N, int ld, int K, double df0, double inv_ddf)
{
    if (isnan(t) || isnan(df)) return EPX_NAN;
    if (isinf(t)) return 0.0;
    const double t2 = t * t, y = t2 / (df + t2);          /* r^2 */
    const double u = sqrt(y) * (double)N;
    int i = (int)u;
    if (i < 1) i = 1;
    if (i > N - 2) i = N - 2;
    const double a = u - (double)i;
    double v = (df - df0) * inv_ddf;
    int k = (int)v;
    if (k < 1) k = 1;
    if (k > K - 3) k = K - 3;
    const double b = v - (double)k;
    const double am = a - 1.0, ap = a + 1.0, a2 = a - 2.0, bm = b - 1.0, bp = b + 1.0, b2 = b - 2.0;
    const double wa0 = -a * am * a2 * (1.0 / 6.0), wa1 = ap * am * a2 * 0.5, wa2 = -ap * a * a2 * 0.5, wa3 = ap * a * am * (1.0 / 6.0);
    const double wb0 = -b * bm * b2 * (1.0 / 6.0), wb1 = bp * bm * b2 * 0.5, wb2 = -bp * b * b2 * 0.5, wb3 = bp * b * bm * (1.0 / 6.0);
    const double *q = tab + (long long)(k - 1) * ld + (i - 1);
    const double r0 = wa0 * q[0] + wa1 * q[1] + wa2 * q[2] + wa3 * q[3];
    q += ld;
    const double r1 = wa0 * q[0] + wa1 * q[1] + wa2 * q[2] + wa3 * q[3];
    q += ld;
    const double r2 = wa0 * q[0] + wa1 * q[1] + wa2 * q[2] + wa3 * q[3];
    q += ld;
    const double r3 = wa0 * q[0] + wa1 * q[1] + wa2 * q[2] + wa3 * q[3];
    const double c = wb0 * r0 + wb1 * r1 + wb2 * r2 + wb3 * r3;
    const double p = exp(c + 0.5 * df * log1p(-y));
    return p > 1.0 ? 1.0 : p;
}

/* psych::corr.test: p of r on n-2 df */
EPX_HD double epx_pearson_p(double r, double n)
{
    double t = (r * sqrt(n - 2.0)) / sqrt(1.0 - r * r);
    return epx_t_two_sided_p(t, n - 2.0);
}

#endif /* EPX_MATH_H */
