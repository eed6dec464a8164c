/*
 * epx_group.cuh — pooled analyses of probe groups: the rows of CG_of_a_genes.csv (every probe of a gene,
 * R/EpimethexAnalysis.R:476-545), CG_by_position.csv and CG_by_islands.csv (the probes of a gene that share a
 * RefGene group / island relation, R/EpimethexAnalysis.R:549-620 with R/Functions.R:136-188).
 *
 * The reference stacks the k probe columns of the group (each in the gene's sample order) into one k*n vector,
 * recycles the stratification column down it and runs the same Analysis + corr.test as for a single probe.  So a
 * group needs (1) moments over the stack, (2) the stack in ascending order with its labels, for the three pooled
 * medians and the three Kolmogorov-Smirnov statistics.  Four kernels:
 *
 *   k_group_moments    one CTA per group: two-pass centred sums straight from the methylation rows (Pearson r of
 *                      rep(expression, k) against the stack, Welch t from the pooled group moments) and the
 *                      calculateTtest guard over all k*m rank-paired differences
 *   k_group_blocksort  one CTA per block of <= rpb probe rows: rows + labels into shared memory; every row is already
 *                      ranked (k_probe_rank's order rows), so log2(rpb) merge-path passes in shared memory order
 *                      the block; write (value, label) runs
 *   k_group_merge      one pass of pairwise merges of adjacent sorted runs of a group in global memory (merge
 *                      path, one contiguous output slice per thread); ceil(log2(blocks)) launches, ping-pong buffers
 *   k_group_walk       one CTA per group: running label counts over the ordered stack -> medians, KS numerators at
 *                      tie-run ends, p-values
 */
#pragma once
#include "epx_ctasort.cuh"
#include "epx_internal.h"

namespace epx {

constexpr int GROUP_THREADS = 256;
constexpr int BLOCKSORT_THREADS = 512;
constexpr int MERGE_SLICE = 16;                         /* outputs per thread of k_group_merge */
constexpr int MERGE_TILE = GROUP_THREADS * MERGE_SLICE; /* outputs per CTA */

__host__ __device__ inline size_t group_sort_smem(int cap) { return (size_t)cap * 8 + (size_t)cap * 2 * 2 + (size_t)cap; }

/* block-wide sum of three doubles; every thread gets the results */
__device__ __forceinline__ void block_sum3(double &a, double &b, double &c, double *red)
{
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5, nw = blockDim.x >> 5;
    a = warp_sum(a); b = warp_sum(b); c = warp_sum(c);
    __syncthreads();
    if (lane == 0) { red[w] = a; red[32 + w] = b; red[64 + w] = c; }
    __syncthreads();
    double ta = 0, tb = 0, tc = 0;
    for (int i = 0; i < nw; i++) { ta += red[i]; tb += red[32 + i]; tc += red[64 + i]; }
    a = ta; b = tb; c = tc;
}

/* ---------------------------------------------------------------- moments, guard, Pearson, Welch */
/* row_of(j): pointer to the n values (original sample order) of the group's j-th probe */
template <typename RowOf>
__device__ __forceinline__ void group_moments_body(const GroupParams &p, const int q, const GroupDesc &gd, RowOf row_of)
{
    __shared__ double red[96];
    const int k = gd.k, n = p.n, m = p.m;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = blockDim.x >> 5;
    if (k == 0) { /* no probe with data: a column of NA (R/EpimethexAnalysis.R:536-541) */
        if (threadIdx.x < PAIR_COLS) p.grp[(int64_t)q * PAIR_COLS + threadIdx.x] = EPX_NAN;
        if (threadIdx.x < 3) { p.grp_stat[(int64_t)q * 3 + threadIdx.x] = EPX_NAN; p.grp_dnum[(int64_t)q * 3 + threadIdx.x] = -1; }
        if (threadIdx.x == 0) p.grp_na[q] = (uint16_t)((1u << PAIR_COLS) - 1);
        return;
    }
    const double *xc = p.xc + (int64_t)gd.gene * p.ldx;
    const uint8_t *lab = p.lab8 + (int64_t)gd.gene * p.ldl;
    const uint16_t *perm = p.perm + (int64_t)gd.gene * n;
    const double dM = (double)k * (double)m, dN = (double)k * (double)n;

    /* pass 1: sums -> means. Shifted by the first value so that a constant stack gives exactly zero deviations
     * (R's long-double mean is exact there and sd == 0 decides NA). */
    const double K = row_of(0)[0];
    double s0 = 0, s1 = 0, s2 = 0, s3 = 0;
    for (int j = warp; j < k; j += nw) {
        const double *row = row_of(j);
        for (int s = lane; s < n; s += 32) {
            const double y = row[s] - K;
            const unsigned l = lab[s];
            s0 += l == 0 ? y : 0.0; s1 += l == 1 ? y : 0.0; s2 += l == 2 ? y : 0.0; s3 += l == 3 ? y : 0.0;
        }
    }
    block_sum3(s0, s1, s2, red);
    double dummy1 = 0, dummy2 = 0;
    block_sum3(s3, dummy1, dummy2, red);
    const double mean_all = K + (s0 + s1 + s2 + s3) / dN;
    const double mg0 = K + s0 / dM, mg1 = K + s1 / dM, mg2 = K + s2 / dM;

    /* pass 2: centred sums + the guard (are all rank-paired differences of the stack identical?) */
    double syy = 0, sxy = 0, q0 = 0, q1 = 0, q2 = 0;
    int differs = 0;
    double ref[3] = { 0, 0, 0 };
    if (m >= 1) {
        const double *row0 = row_of(0);
        const double a = row0[perm[0]], b = row0[perm[m]], c = row0[perm[2 * m]];
        ref[0] = a - b; ref[1] = a - c; ref[2] = b - c;
    }
    for (int j = warp; j < k; j += nw) {
        const double *row = row_of(j);
        for (int s = lane; s < n; s += 32) {
            const double y = row[s];
            const unsigned l = lab[s];
            const double d = y - mean_all;
            syy = fma(d, d, syy);
            sxy = fma(xc[s], d, sxy);
            const double dg = y - (l == 0 ? mg0 : (l == 1 ? mg1 : mg2));
            q0 = fma(l == 0 ? dg : 0.0, dg, q0); q1 = fma(l == 1 ? dg : 0.0, dg, q1); q2 = fma(l == 2 ? dg : 0.0, dg, q2);
        }
        for (int i = lane; i < m; i += 32) {
            const double a = row[perm[i]], b = row[perm[m + i]], c = row[perm[2 * m + i]];
            if (a - b != ref[0]) differs |= 1;
            if (a - c != ref[1]) differs |= 2;
            if (b - c != ref[2]) differs |= 4;
        }
    }
    double zero = 0;
    block_sum3(syy, sxy, zero, red);
    block_sum3(q0, q1, q2, red);
    { /* __syncthreads_or is a logical OR: one call per comparison */
        const int f0 = __syncthreads_or(differs & 1), f1 = __syncthreads_or(differs & 2), f2 = __syncthreads_or(differs & 4);
        differs = (f0 ? 1 : 0) | (f1 ? 2 : 0) | (f2 ? 4 : 0);
    }

    if (threadIdx.x == 0) {
        unsigned na = 0;
        int guard[3];
        for (int c = 0; c < 3; c++) guard[c] = dM < 2.0 ? -1 : ((differs >> c) & 1);
        const double *ga = p.gene_aux + 4 * (int64_t)gd.gene;
        double r = EPX_NAN, pr = EPX_NAN;
        if (ga[0] != 0.0 && syy != 0.0 && dN >= 3.0) {
            r = sxy * ga[3] / sqrt((double)k * syy);
            r = r > 1.0 ? 1.0 : (r < -1.0 ? -1.0 : r);
            pr = epx_pearson_p(r, dN);
        } else na |= (1u << EPX_PEARSON_R) | (1u << EPX_PEARSON_P);
        p.grp[(int64_t)q * PAIR_COLS + EPX_PEARSON_R] = r;
        p.grp[(int64_t)q * PAIR_COLS + EPX_PEARSON_P] = pr;
        if (p.test_t) {
            const double mg[3] = { mg0, mg1, mg2 };
            const double var[3] = { q0 / (dM - 1.0), q1 / (dM - 1.0), q2 / (dM - 1.0) };
            const int ga_[3] = { 0, 0, 1 }, gb_[3] = { 1, 2, 2 };
            for (int c = 0; c < 3; c++) {
                double t = EPX_NAN, df = EPX_NAN, pv = EPX_NAN;
                if (guard[c] == 1 && epx_welch(mg[ga_[c]], var[ga_[c]], dM, mg[gb_[c]], var[gb_[c]], dM, &t, &df) == 0)
                    pv = epx_t_two_sided_p(t, df);
                else { t = EPX_NAN; na |= 1u << (EPX_P_UP_MID + c); }
                p.grp[(int64_t)q * PAIR_COLS + EPX_P_UP_MID + c] = pv;
                p.grp_stat[(int64_t)q * 3 + c] = t;
                p.grp_dnum[(int64_t)q * 3 + c] = -1;
            }
        }
        p.grp_na[q] = (uint16_t)na;
        for (int c = 0; c < 3; c++) p.guard[(int64_t)q * 3 + c] = guard[c];
    }
}

/* groups that need more than one sorted block (and the empty ones): rows straight from global memory */
__global__ void __launch_bounds__(GROUP_THREADS) k_group_moments(GroupParams p)
{
    const int q = p.multi[blockIdx.x];
    const GroupDesc gd = p.groups[q];
    const int32_t *pairs = p.group_pairs + gd.first;
    group_moments_body(p, q, gd, [&](int j) { return p.meth + (int64_t)p.probe_idx[pairs[j]] * p.ld; });
}

/* ---------------------------------------------------------------- one pass of pairwise run merges */
__global__ void __launch_bounds__(GROUP_THREADS) k_group_merge(GroupParams p, int pass, int64_t run_len)
{
    const TileDesc td = p.tiles[blockIdx.x];
    const GroupDesc gd = p.groups[td.group];
    const int64_t N = (int64_t)gd.k * p.n;
    const double *sv = p.val[pass & 1] + gd.elem;
    const uint8_t *sl = p.lab[pass & 1] + gd.elem;
    double *dv = p.val[(pass + 1) & 1] + gd.elem;
    uint8_t *dl = p.lab[(pass + 1) & 1] + gd.elem;
    int64_t o = (int64_t)td.tile * MERGE_TILE + (int64_t)threadIdx.x * MERGE_SLICE;
    const int64_t o_end = min(o + MERGE_SLICE, N);
    while (o < o_end) {
        const int64_t pbase = (o / (2 * run_len)) * (2 * run_len);
        /* 32-bit offsets inside the pair of runs (a group's stack is far below 2^31 entries) */
        const int la = (int)min(run_len, N - pbase), lb = (int)min(run_len, N - pbase - la);
        const double *A = sv + pbase;    /* B = A + la */
        const uint8_t *LA = sl + pbase;
        double *D = dv + pbase;
        uint8_t *DL = dl + pbase;
        const int seg_end = (int)(min(o_end, pbase + la + lb) - pbase);
        const int d = (int)(o - pbase);
        /* merge path: i = how many of the first d outputs of this pair of runs come from A (A wins ties) */
        int lo = max(0, d - lb), hi = min(d, la);
        while (lo < hi) {
            const int mid = (lo + hi) >> 1;
            if (!(A[la + d - 1 - mid] < A[mid])) lo = mid + 1; else hi = mid;
        }
        /* branch-free sequential merge: pa/pb = next unread entry of A/B, (va, ua)/(vb, ub) = the current heads;
         * heads are clamped into the pair of runs, an absent B aliases the last entry of A */
        int pa = lo, pb = d - lo;
        const int lastA = la - 1, lastB = la + lb - 1;
        int ia = min(pa, lastA), ib = min(la + pb, lastB);
        double va = A[ia], vb = A[ib];
        unsigned ua = LA[ia], ub = LA[ib];
        for (int t = d; t < seg_end; t++) {
            const bool take_a = pb >= lb || (pa < la && !(vb < va));
            D[t] = take_a ? va : vb;
            DL[t] = (uint8_t)(take_a ? ua : ub);
            pa += take_a ? 1 : 0;
            pb += take_a ? 0 : 1;
            const int nx = take_a ? min(pa, lastA) : min(la + pb, lastB);
            const double nv = A[nx];
            const unsigned nl = LA[nx];
            va = take_a ? nv : va; ua = take_a ? nl : ua;
            vb = take_a ? vb : nv; ub = take_a ? ub : nl;
        }
        o = pbase + seg_end;
    }
}

/* ---------------------------------------------------------------- walk of the ordered stack */
/* val_at(i), lab_at(i): value and label of the i-th smallest entry of the stack */
template <typename ValAt, typename LabAt>
__device__ __forceinline__ void group_walk_body(const GroupParams &p, const int q, const GroupDesc &gd, ValAt val_at, LabAt lab_at)
{
    __shared__ int s_cnt[3][32];
    __shared__ int s_mx[3][32];
    __shared__ double s_med[3][2];
    __shared__ double s_u[128];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarp = blockDim.x >> 5;
    const int64_t N = (int64_t)gd.k * p.n, M = (int64_t)gd.k * p.m;
    const int64_t per = (N + blockDim.x - 1) / blockDim.x;
    const int64_t b = min((int64_t)threadIdx.x * per, N), e = min(b + per, N);

    /* pass 1: label counts of this thread's slice, exclusive scan over the CTA */
    int c0 = 0, c1 = 0, c2 = 0;
    for (int64_t i = b; i < e; i++) {
        const unsigned l = lab_at(i);
        c0 += l == 0; c1 += l == 1; c2 += l == 2;
    }
    int i0 = c0, i1 = c1, i2 = c2;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        const int t0 = __shfl_up_sync(FULL, i0, d), t1 = __shfl_up_sync(FULL, i1, d), t2 = __shfl_up_sync(FULL, i2, d);
        if (lane >= d) { i0 += t0; i1 += t1; i2 += t2; }
    }
    if (lane == 31) { s_cnt[0][warp] = i0; s_cnt[1][warp] = i1; s_cnt[2][warp] = i2; }
    __syncthreads();
    int r0 = i0 - c0, r1 = i1 - c1, r2 = i2 - c2;
    for (int w = 0; w < warp; w++) { r0 += s_cnt[0][w]; r1 += s_cnt[1][w]; r2 += s_cnt[2][w]; }

    /* pass 2: running counts; middle elements of each label group; KS differences at tie-run ends */
    const int64_t t1 = (M - 1) / 2 + 1, t2 = M / 2 + 1;
    int mx01 = 0, mx02 = 0, mx12 = 0;
    if (b < e) {
        double v = val_at(b);
        for (int64_t i = b; i < e; i++) {
            const double nx = i + 1 < N ? val_at(i + 1) : 0.0;
            const unsigned l = lab_at(i);
            if (l < 3) {
                const int cl = l == 0 ? ++r0 : (l == 1 ? ++r1 : ++r2);
                if (cl == t1) s_med[l][0] = v;
                if (cl == t2) s_med[l][1] = v;
            }
            if (i + 1 == N || nx != v) {
                mx01 = max(mx01, abs(r0 - r1)); mx02 = max(mx02, abs(r0 - r2)); mx12 = max(mx12, abs(r1 - r2));
            }
            v = nx;
        }
    }
    mx01 = warp_max_i(mx01); mx02 = warp_max_i(mx02); mx12 = warp_max_i(mx12);
    if (lane == 0) { s_mx[0][warp] = mx01; s_mx[1][warp] = mx02; s_mx[2][warp] = mx12; }
    __syncthreads();
    if (threadIdx.x != 0) return;

    int dd[3] = { 0, 0, 0 };
    for (int w = 0; w < nwarp; w++)
        for (int c = 0; c < 3; c++) dd[c] = max(dd[c], s_mx[c][w]);
    double *o = p.grp + (int64_t)q * PAIR_COLS;
    double medUP = EPX_NAN, medMID = EPX_NAN, medDOWN = EPX_NAN;
    if (M >= 1) {
        medUP = (s_med[0][0] + s_med[0][1]) * 0.5; medMID = (s_med[1][0] + s_med[1][1]) * 0.5;
        medDOWN = (s_med[2][0] + s_med[2][1]) * 0.5;
    }
    o[EPX_MEDIAN_DOWN] = medDOWN; o[EPX_MEDIAN_MEDIUM] = medMID; o[EPX_MEDIAN_UP] = medUP;
    o[EPX_BD_UP_MID] = medUP - medMID; o[EPX_BD_UP_DOWN] = medUP - medDOWN; o[EPX_BD_MID_DOWN] = medMID - medDOWN;
    if (p.test_t) return;

    const double dM = (double)M;
    const bool exact_ok = dM * dM < 10000.0;
    int ties[3] = { 0, 0, 0 };
    if (exact_ok) {
        /* ks.test uses psmirnov2x only when the two groups being compared hold no tie (tiny stacks only) */
        int a0 = 0, a1 = 0, a2 = 0;
        for (int64_t i = 0; i < N; i++) {
            const unsigned l = lab_at(i);
            a0 += l == 0; a1 += l == 1; a2 += l == 2;
            if (i + 1 == N || val_at(i + 1) != val_at(i)) {
                if (a0 + a1 >= 2) ties[0] = 1;
                if (a0 + a2 >= 2) ties[1] = 1;
                if (a1 + a2 >= 2) ties[2] = 1;
                a0 = a1 = a2 = 0;
            }
        }
    }
    unsigned na = p.grp_na[q];
    for (int c = 0; c < 3; c++) {
        double pv = EPX_NAN, stat = EPX_NAN;
        int64_t dnum = -1;
        if (p.guard[(int64_t)q * 3 + c] != 0 && M >= 1) { /* difference != 0 || is.na(difference) */
            dnum = (int64_t)dd[c] * M;
            stat = (double)dnum / (dM * dM);
            pv = (exact_ok && !ties[c]) ? epx_ks_p_exact((double)dnum, (int)M, (int)M, s_u) : epx_ks_p_asym((double)dnum, dM, dM);
        } else na |= 1u << (EPX_P_UP_MID + c);
        o[EPX_P_UP_MID + c] = pv;
        p.grp_stat[(int64_t)q * 3 + c] = stat;
        p.grp_dnum[(int64_t)q * 3 + c] = dnum;
    }
    p.grp_na[q] = (uint16_t)na;
}

__global__ void __launch_bounds__(GROUP_THREADS) k_group_walk(GroupParams p)
{
    const int q = p.multi[blockIdx.x];
    const GroupDesc gd = p.groups[q];
    if (gd.k == 0) return;
    int fin = 0;
    while ((1 << fin) < gd.nblocks) fin++;
    const double *val = p.val[fin & 1] + gd.elem;
    const uint8_t *lab = p.lab[fin & 1] + gd.elem;
    group_walk_body(p, q, gd, [&](int64_t i) { return val[i]; }, [&](int64_t i) { return (unsigned)lab[i]; });
}

/* ---------------------------------------------------------------- sorted blocks of <= rpb probe rows */
__global__ void __launch_bounds__(BLOCKSORT_THREADS) k_group_blocksort(GroupParams p)
{
    extern __shared__ __align__(16) unsigned char smem[];
    const BlockDesc bd = p.blocks[blockIdx.x];
    const GroupDesc gd = p.groups[bd.group];
    const int n = p.n, nb = bd.runs * n, cap = p.cap;
    double *s_val = (double *)smem;
    unsigned short *s_a = (unsigned short *)(smem + (size_t)cap * 8);
    unsigned short *s_b = s_a + cap;
    unsigned char *s_lab = (unsigned char *)(s_b + cap);
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = blockDim.x >> 5;
    const uint8_t *lab = p.lab8 + (int64_t)gd.gene * p.ldl;
    const int32_t *pairs = p.group_pairs + gd.first + bd.r0;
    /* every probe row was ranked once by k_probe_rank: start from sorted runs of n and only merge */
    for (int j = warp; j < bd.runs; j += nw) {
        const int pr = pairs[j];
        const double *row = p.meth + (int64_t)p.probe_idx[pr] * p.ld;
        const uint16_t *po = p.ord + (int64_t)p.pair_slot[pr] * p.ldo;
        for (int s = lane; s < n; s += 32) {
            s_val[j * n + s] = row[s];
            s_lab[j * n + s] = lab[s];
            s_a[j * n + s] = (unsigned short)(j * n + (po[s] & 0x7fffu));
        }
    }
    __syncthreads();
    const bool whole = gd.nblocks == 1; /* the whole group is in shared memory: finish here, nothing goes to HBM */
    if (whole) group_moments_body(p, bd.group, gd, [&](int j) { return (const double *)(s_val + j * n); });
    const unsigned short *ord = cta_merge_runs<false>(s_val, nb, n, s_a, s_b);
    if (whole) {
        group_walk_body(p, bd.group, gd, [&](int64_t i) { return s_val[ord[i]]; }, [&](int64_t i) { return (unsigned)s_lab[ord[i]]; });
        return;
    }
    double *ov = p.val[0] + gd.elem + (int64_t)bd.r0 * n;
    uint8_t *ol = p.lab[0] + gd.elem + (int64_t)bd.r0 * n;
    for (int t = threadIdx.x; t < nb; t += blockDim.x) {
        const int i = ord[t];
        ov[t] = s_val[i];
        ol[t] = s_lab[i];
    }
}

} // namespace epx
