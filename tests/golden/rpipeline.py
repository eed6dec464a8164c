"""Second, independent restatement of the reference path, used ONLY to generate and re-check the golden tables in
tests/golden/ (make_golden.py).  It is written from the R sources, statement by statement, in the reference's own
data layout (the samples x genes frame sorted column by column, the `variable` / `stratification` label columns, the
n-row probe columns) with numpy, fractions, big integers and mpmath -- it imports neither oracle/ nor libepx nor any
statistic from scipy, and it shares no code with them.  A misreading of R therefore has to be made twice, in two
languages and two program structures, to slip through.  It is NOT output of R (R cannot be installed here): parity
stays "unpinned", this is the strongest pin available without R.

Reference statements followed (citations into /root/reference):
  R/EpimethexAnalysis.R:117-122  indexTCGA = order(column, decreasing = TRUE); columns sorted decreasingly
  :125-129                       matrixStats::colQuantiles(probs = c(.33, .66, .99))           (type 7)
  :134-139                       which.max(#unique) and RcmdrMisc::bin.var(bins = 3, 'proportions')
  :146-172                       group means; calcFC reads the LAST three rows (= the quantile rows)
  :184-206, R/Functions.R:192-220 calculateTtest on the expression groups (guard, t.test / ks.test)
  :299-302                       probe values in the gene's sample order
  :313-315                       valExprGene
  :323-325                       stratification = rep(c("UP","Medium","Down"), each = n/3)
  :351-453, R/Functions.R:85-132 Loop A: Analysis (medians, beta differences, three tests) + psych::corr.test
  :476-545                       Loop B: the stacked probes of a gene (CG_of_a_genes)
Third-party arithmetic restated from its published definition: stats::t.test (Welch), stats::ks.test in its
R 3.4-4.0 form (exact unless ties / n.x*n.y >= 10000; C_pSmirnov2x, C_pKS2 with tol = 1e-6), stats::quantile type 7,
psych::corr.test (t = r sqrt(n-2)/sqrt(1-r^2), two-sided p on n-2 df).
"""
from __future__ import annotations

import math
from fractions import Fraction

import mpmath as mp
import numpy as np

mp.mp.dps = 40
NA = float("nan")


# ---------------------------------------------------------------- base R pieces

def r_order_decreasing(col):
    """order(x, decreasing = TRUE): ties keep their original order (radix / xtfrm-rank order is stable)"""
    return sorted(range(len(col)), key=lambda i: (-col[i], i))


def r_quantile7(x_sorted_increasing, p):
    """stats::quantile(type = 7) as quantile.default writes it (matrixStats::colQuantiles does the same):
    index <- 1 + (n - 1) * p; lo <- floor(index); hi <- ceiling(index); qs <- x[lo];
    where (index > lo & x[hi] != qs): qs <- (1 - h) * qs + h * x[hi] with h = index - lo"""
    x = x_sorted_increasing
    n = len(x)
    index = 1 + (n - 1) * p
    lo, hi = math.floor(index), math.ceil(index)
    qs = x[lo - 1]
    if index > lo and x[hi - 1] != qs:
        h = index - lo
        qs = (1 - h) * qs + h * x[hi - 1]
    return qs


def r_sign(v):
    return (v > 0) - (v < 0)


def r_median(v):
    s = sorted(v)
    k = len(s)
    if k == 0:
        return NA
    return s[k // 2] if k % 2 else (s[k // 2 - 1] + s[k // 2]) / 2.0     # mean(sort(x)[half + 0:1])


def r_mean(v):
    """mean(): long-double accumulation with a second pass (summary.c); math.fsum is exact, which agrees to < 1 ulp"""
    return math.fsum(v) / len(v)


def r_var(v):
    m = r_mean(v)
    return math.fsum((a - m) * (a - m) for a in v) / (len(v) - 1)


def r_sd(v):
    return math.sqrt(r_var(v)) if len(v) >= 2 else NA


def pt_two_sided(t, df):
    """2 * pt(-|t|, df) = I_{df/(df+t^2)}(df/2, 1/2)"""
    t, df = mp.mpf(t), mp.mpf(df)
    if t == 0:
        return 1.0
    return float(mp.betainc(df / 2, mp.mpf(1) / 2, 0, df / (df + t * t), regularized=True))


def t_test_welch(x, y):
    """stats::t.test(x, y, var.equal = FALSE)[c('statistic','p.value')]; None = R stops with an error"""
    nx, ny = len(x), len(y)
    if nx < 2 or ny < 2:
        return None                       # "not enough 'x' observations"
    mx, my, vx, vy = r_mean(x), r_mean(y), r_var(x), r_var(y)
    sx, sy = math.sqrt(vx / nx), math.sqrt(vy / ny)
    se = math.sqrt(sx * sx + sy * sy)
    if se < 10 * np.finfo(float).eps * max(abs(mx), abs(my)):
        return None                       # "data are essentially constant"
    df = se ** 4 / (sx ** 4 / (nx - 1) + sy ** 4 / (ny - 1))
    t = (mx - my) / se
    return t, pt_two_sided(t, df), df


def psmirnov_exact_upper(dnum, m, n):
    """1 - C_pSmirnov2x(D, m, n) for D = dnum / (m n), no ties: P(D_mn >= D) by counting the lattice paths from (0,0) to
    (m,n) that stay inside |i n - j m| <= dnum - 1 (big integers, exact), over choose(m+n, n).  ks.c's recursion
    u[j] <- w u[j] + u[j-1] with q = (0.5 + floor(D m n - 1e-7)) / (m n) is this count divided step by step."""
    lim = dnum - 1
    row = [1 if j * m <= lim else 0 for j in range(n + 1)]          # i = 0
    for i in range(1, m + 1):
        new = [0] * (n + 1)
        new[0] = row[0] if i * n <= lim else 0
        for j in range(1, n + 1):
            new[j] = (row[j] + new[j - 1]) if abs(i * n - j * m) <= lim else 0
        row = new
    return float(1 - Fraction(row[n], math.comb(m + n, n)))


def pkstwo_r(x, tol=1e-6):
    """ks.c pkstwo: for x < 1 the series sqrt(2 pi)/x * sum_{k odd} exp(-k^2 pi^2 / (8 x^2)) stops at
    k_max = (int) sqrt(2 - log(tol)) = 3 with k += 2, i.e. after its FIRST term; else the alternating series until two
    partial sums differ by <= tol"""
    if x <= 0:
        return 0.0
    if x < 1:
        k_max = int(math.sqrt(2 - math.log(tol)))
        z = -(math.pi / 2) * (math.pi / 4) / (x * x)
        w = math.log(x)
        s = 0.0
        k = 1
        while k < k_max:
            s += math.exp(k * k * z - w)
            k += 2
        return s * math.sqrt(2 * math.pi)
    z = -2 * x * x
    s, k, old, new = -1.0, 1, 0.0, 1.0
    while abs(old - new) > tol:
        old = new
        new += 2 * s * math.exp(z * k * k)
        s = -s
        k += 1
    return new


def ks_test(x, y):
    """stats::ks.test(x, y) two-sided, R 3.4-4.0: returns (p.value, D numerator over n.x n.y, exact?)"""
    nx, ny = len(x), len(y)
    if nx < 1 or ny < 1:
        return None                       # "not enough 'x' data"
    w = list(x) + list(y)
    o = sorted(range(nx + ny), key=lambda i: (w[i], i))              # order(w)
    z, cx, cy = [], 0, 0
    for i in o:                                                        # cumsum(ifelse(order(w) <= n.x, 1/n.x, -1/n.y))
        if i < nx:
            cx += 1
        else:
            cy += 1
        z.append((cx, cy))
    ws = [w[i] for i in o]
    ties = len(set(w)) < nx + ny
    if ties:                                                           # z[c(which(diff(sort(w)) != 0), n.x + n.y)]
        z = [z[k] for k in range(nx + ny) if k == nx + ny - 1 or ws[k + 1] != ws[k]]
    dnum = max(abs(a * ny - b * nx) for a, b in z)                      # max|z| * n.x * n.y, exactly
    exact = (nx * ny < 10000) and not ties
    if exact:
        p = psmirnov_exact_upper(dnum, min(nx, ny), max(nx, ny))
    else:
        D = dnum / (nx * ny)
        p = 1 - pkstwo_r(math.sqrt(nx * ny / (nx + ny)) * D)
    return min(1.0, max(0.0, p)), dnum, exact


def calculate_ttest(a1, a2, flag):
    """R/Functions.R:192-220.  Returns (statistic or NA, p or NA, KS numerator or -1)."""
    L = max(len(a1), len(a2))
    diff = [a1[i % len(a1)] - a2[i % len(a2)] for i in range(L)] if len(a1) and len(a2) else []   # mapply recycles
    difference = r_sd(diff)
    if flag:
        if difference != 0 and not math.isnan(difference):
            res = t_test_welch(a1, a2)
            if res is None:
                return NA, NA, -1
            return res[0], res[1], -1
        return NA, NA, -1           # sd == 0 -> c(NA, NA); sd NA -> `if (NA)` stops the reference: reported as NA
    if difference != 0 or math.isnan(difference):
        res = ks_test(a1, a2)
        if res is None:
            return NA, NA, -1       # tryCatch -> p.value NA
        p, dnum, _ = res
        return dnum / (len(a1) * len(a2)), p, dnum
    return NA, NA, -1


def linear_fc(a, b):
    """calculateLinearFC, R/Functions.R:68-79"""
    if math.isnan(a) or math.isnan(b):
        return NA
    same = r_sign(a) == r_sign(b)
    if same:
        fc = max(abs(a), abs(b)) / min(abs(a), abs(b)) if min(abs(a), abs(b)) != 0 else (NA if max(abs(a), abs(b)) == 0 else math.inf)
    else:
        fc = max(a, b) - min(a, b)
    return fc if b < a else -fc


def log_fc(a, b):
    """calculateLogFC, R/Functions.R:59-63"""
    if math.isnan(a) or math.isnan(b):
        return NA
    fold = 2.0 ** abs(a - b)
    return fold if b < a else -fold


def corr_test(x, y):
    """psych::corr.test(x, y, adjust = 'none') for two columns: r = cor(x, y), t = r sqrt(n-2)/sqrt(1-r^2),
    p = 2 (1 - pt(|t|, n-2))"""
    n = len(x)
    mx, my = r_mean(x), r_mean(y)
    sxx = math.fsum((a - mx) ** 2 for a in x)
    syy = math.fsum((b - my) ** 2 for b in y)
    if n < 3 or sxx == 0 or syy == 0:
        return NA, NA                     # sd is zero -> cor NA
    sxy = math.fsum((a - mx) * (b - my) for a, b in zip(x, y))
    r = sxy / math.sqrt(sxx) / math.sqrt(syy)
    r = max(-1.0, min(1.0, r))
    if abs(r) == 1.0:
        return r, 0.0
    rr = mp.mpf(r)
    df = mp.mpf(n - 2)
    p = float(mp.betainc(df / 2, mp.mpf(1) / 2, 0, 1 - rr * rr, regularized=True))
    return r, p


# ---------------------------------------------------------------- the path

def analysis_block(values, labels, test_t):
    """Analysis(), R/Functions.R:85-132, for ONE value column with its stratification column:
    medianDown/Medium/UP, the three beta differences, the three tests."""
    grp = {k: [v for v, l in zip(values, labels) if l == k] for k in ("UP", "Medium", "Down")}
    med_up, med_mid, med_down = r_median(grp["UP"]), r_median(grp["Medium"]), r_median(grp["Down"])
    out = [med_down, med_mid, med_up, med_up - med_mid, med_up - med_down, med_mid - med_down]
    stats3, nums = [], []
    for a, b in (("UP", "Medium"), ("UP", "Down"), ("Medium", "Down")):
        st, p, dn = calculate_ttest(grp[a], grp[b], test_t)
        out.append(p)
        stats3.append(st)
        nums.append(dn)
    return out, stats3, nums


def run(expr, meth, row_ptr, probe_idx, *, data_linear=True, test_expr_t=True, test_meth_t=False, fc_from_means=False,
        pooled_by_gene=False):
    """expr [G][n] (gene-major, original sample order), meth [P][n]; CSR gene -> methylation rows (-1 = no data).
    Returns a dict of python lists: pair rows (11 values), gene rows (15 values as include/epx.h orders them), KS
    numerators, sample order, bin.var group sizes, reference gene; optionally the CG_of_a_genes rows."""
    G, n = len(expr), len(expr[0])
    # :117-122
    perm = [r_order_decreasing(list(expr[g])) for g in range(G)]
    sorted_cols = [[expr[g][i] for i in perm[g]] for g in range(G)]
    # :134-139
    nuniq = [len(set(c)) for c in sorted_cols]
    index = max(range(G), key=lambda g: (nuniq[g], -g))                # which.max: first maximum
    ref_inc = sorted_cols[index][::-1]
    br = [r_quantile7(ref_inc, k / 3.0) for k in range(4)]             # quantile(x, seq(0, 1, 1/3))
    if not (br[0] < br[1] < br[2] < br[3]):
        raise ValueError("'breaks' are not unique")
    variable = []
    for v in sorted_cols[index]:                                       # cut(..., include.lowest = TRUE), labels D, M, UP
        variable.append("D" if v <= br[1] else ("M" if v <= br[2] else "UP"))
    counts = (variable.count("UP"), variable.count("M"), variable.count("D"))
    # :323-325
    m = n // 3
    strat = ["UP"] * m + ["Medium"] * m + ["Down"] * m + [None] * (n - 3 * m)   # rep(each = n/3) truncates; extension: unlabelled

    gene_rows, gene_nums = [], []
    for g in range(G):
        col = sorted_cols[g]
        grp = {k: [v for v, l in zip(col, variable) if l == k] for k in ("UP", "M", "D")}
        mean_up, mean_mid, mean_down = (r_mean(grp[k]) if grp[k] else NA for k in ("UP", "M", "D"))
        inc = col[::-1]
        q33, q66, q99 = (r_quantile7(inc, p) for p in (0.33, 0.66, 0.99))
        a_up, a_mid, a_down = (mean_up, mean_mid, mean_down) if fc_from_means else (q99, q66, q33)   # calcFC: last 3 rows
        fcf = linear_fc if data_linear else log_fc
        fc = [fcf(a_up, a_mid), fcf(a_up, a_down), fcf(a_mid, a_down)]
        st3, p3, n3 = [], [], []
        for a, b in (("UP", "M"), ("UP", "D"), ("M", "D")):
            st, p, dn = calculate_ttest(grp[a], grp[b], test_expr_t)
            st3.append(st); p3.append(p); n3.append(dn)
        gene_rows.append(fc + p3 + st3 + [q33, q66, q99, mean_down, mean_mid, mean_up])
        gene_nums.append(n3)

    pair_rows, pair_stats, pair_nums = [], [], []
    for g in range(G):
        for j in range(row_ptr[g], row_ptr[g + 1]):
            pi = probe_idx[j]
            if pi < 0:                                                  # all-NA column, :375-394
                pair_rows.append([NA] * 11); pair_stats.append([NA] * 3); pair_nums.append([-1] * 3)
                continue
            y = [meth[pi][i] for i in perm[g]]                          # :299-302
            out, st3, n3 = analysis_block(y, strat, test_meth_t)
            r, p = corr_test(sorted_cols[g], y)                         # :404-417
            pair_rows.append(out + [r, p]); pair_stats.append(st3); pair_nums.append(n3)

    res = dict(pair=pair_rows, pair_stat=pair_stats, ks_dnum=pair_nums, gene=gene_rows, gene_ks_dnum=gene_nums,
               perm=perm, counts=counts, ref_gene=index)
    if pooled_by_gene:                                                   # Loop B, :476-545
        rows, nums = [], []
        for g in range(G):
            cols = [probe_idx[j] for j in range(row_ptr[g], row_ptr[g + 1]) if probe_idx[j] >= 0]
            if not cols:
                rows.append([NA] * 11); nums.append([-1] * 3)
                continue
            stacked, labels = [], []
            for pi in cols:                                             # stack(): column after column; cbind recycles strat
                stacked += [meth[pi][i] for i in perm[g]]
                labels += strat
            out, _, n3 = analysis_block(stacked, labels, test_meth_t)
            r, p = corr_test(sorted_cols[g] * len(cols), stacked)       # rep(expression, num_CG)
            rows.append(out + [r, p]); nums.append(n3)
        res["group"] = rows
        res["group_dnum"] = nums
    return res
