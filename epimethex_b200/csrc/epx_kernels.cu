/*
 * epx_kernels.cu — sm_100a kernels of the EpiMethEx association path.
 *
 *   k_gene_sort       per gene: stable descending order of the samples (R/EpimethexAnalysis.R:117-122),
 *                     methylation-side labels by rank (:323-325), centred expression for Pearson (:404-417)
 *   k_pick_reference  reference gene + RcmdrMisc::bin.var thirds (:134-139)
 *   k_gene_tests      quantiles (:125-129), group means (:146-155), calcFC (:172, R/Functions.R:26-80),
 *                     the three expression tests (:184-206, R/Functions.R:192-220)
 *   k_probe_rank      per referenced probe: ascending order of its methylation row with tie marks
 *                     (the sort that stats::median / stats::ks.test perform per call in the reference)
 *   k_pair_stats      per (gene, probe) pair: the gather (:299-302) fused with Analysis
 *                     (R/Functions.R:85-132) and corr.test (:404-417)
 */
#include <atomic>
#include "epx_device.cuh"
#include "epx_internal.h"
#include "epx_warpsort.cuh"
#include "epx_ctasort.cuh"
#include "epx_pair.cuh"
#include "epx_ctabucket.cuh"
#include "epx_group.cuh"

namespace epx {

/* ============================================================ k_gene_sort */

__global__ void __maxnreg__(80) k_gene_sort(GeneParams p)
{
    extern __shared__ __align__(16) unsigned char smem[];
    const int g = blockIdx.x, n = p.n;
    double *s_val = (double *)smem;
    unsigned *region = (unsigned *)(smem + ctab_row_bytes(n)); /* bucket placement; holds the merge sort's buffers on a fallback */
    __shared__ double red[40];
    __shared__ int s_ties;

    const double *x = p.expr + (int64_t)g * n;
    if (threadIdx.x == 0) s_ties = 0;
    double sum = 0.0;
    for (int i = threadIdx.x; i < n; i += blockDim.x) {
        const double v = x[i];
        s_val[i] = v;
        sum += v;
    }
    __syncthreads();
    /* value descending, ties by sample index: bucket placement (epx_ctabucket.cuh), else the merge sort */
    const unsigned short *ord = cta_bucket_order<true>(s_val, x, n, region, p.one, nullptr);
    bool marked = ord != nullptr;
    if (!ord) { /* the placement used the row's place in shared memory: stage it again for the merge sort */
        unsigned short *s_a = (unsigned short *)region, *s_b = s_a + ctasort_entries(n);
        __syncthreads();
        for (int i = threadIdx.x; i < n; i += blockDim.x) s_val[i] = x[i];
        __syncthreads();
        ord = cta_sort_row<true>(s_val, n, s_a, s_b, p.one);
    }

    int ties = 0; /* positions whose successor holds an equal value */
    for (int i = threadIdx.x; i < n; i += blockDim.x) {
        const int s = ord[i] & 0x7fff;
        p.perm[(int64_t)g * n + i] = (uint16_t)s;
        const int lab = i / p.m;
        p.lab8[(int64_t)g * p.ldl + s] = (uint8_t)(lab > 3 ? 3 : lab);
        if (marked) ties += ord[i] >> 15;
        else if (i > 0 && x[s] == x[ord[i - 1] & 0x7fff]) ties++;
    }
    if (ties) atomicAdd(&s_ties, ties);

    double mean = block_sum(sum, red) / (double)n;
    double sxx = 0.0, sxc = 0.0;
    for (int i = threadIdx.x; i < n; i += blockDim.x) {
        double d = x[i] - mean; /* from global memory (L2): the staged row is gone */
        p.xc[(int64_t)g * p.ldx + i] = d;
        sxx += d * d;
        sxc += d;
    }
    sxx = block_sum(sxx, red);
    sxc = block_sum(sxc, red);
    if (threadIdx.x == 0) {
        int u = n - s_ties;
        p.nuniq[g] = u;
        double *ga = p.gene_aux + 4 * (int64_t)g;
        ga[0] = (u == 1) ? 0.0 : sxx; /* constant gene: sd is exactly 0 in R */
        ga[1] = mean; ga[2] = sxc; ga[3] = (u == 1 || sxx == 0.0) ? 0.0 : rsqrt(sxx);
    }
}

/* ============================================================ k_pick_reference */

__device__ __forceinline__ double quantile7(const double *x, const uint16_t *perm, int n, double prob)
{
    /* type 7 on the ascending order; ascending element k (0-based) = x[perm[n-1-k]] */
    double index = 1.0 + (double)(n - 1) * prob;
    double lo = floor(index), hi = ceil(index);
    double qs = x[perm[n - (int)lo]];
    double xh = x[perm[n - (int)hi]];
    if (index > lo && xh != qs) {
        double h = index - lo;
        qs = (1.0 - h) * qs + h * xh;
    }
    return qs;
}

/* the same on a column already sorted in descending order (shared memory): xs[i] = x[perm[i]] */
__device__ __forceinline__ double quantile7_desc(const double *xs, int n, double prob)
{
    double index = 1.0 + (double)(n - 1) * prob;
    double lo = floor(index), hi = ceil(index);
    double qs = xs[n - (int)lo];
    double xh = xs[n - (int)hi];
    if (index > lo && xh != qs) {
        double h = index - lo;
        qs = (1.0 - h) * qs + h * xh;
    }
    return qs;
}

__global__ void __launch_bounds__(1024) k_pick_reference(GeneTestParams p)
{
    __shared__ int s_best[32], s_arg[32];
    __shared__ double s_b[4];
    __shared__ int s_cnt[2];
    /* which.max(apply(..., function(x) length(unique(x)))): first gene with the most distinct values */
    int best = -1, arg = 0x7fffffff;
    for (int g = threadIdx.x; g < p.G; g += blockDim.x) {
        int u = p.nuniq[g];
        if (u > best) { best = u; arg = g; }
    }
    for (int o = 16; o > 0; o >>= 1) {
        int b2 = __shfl_xor_sync(FULL, best, o), a2 = __shfl_xor_sync(FULL, arg, o);
        if (b2 > best || (b2 == best && a2 < arg)) { best = b2; arg = a2; }
    }
    int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    if (lane == 0) { s_best[w] = best; s_arg[w] = arg; }
    __syncthreads();
    if (w == 0) {
        best = s_best[lane]; arg = s_arg[lane];
        for (int o = 16; o > 0; o >>= 1) {
            int b2 = __shfl_xor_sync(FULL, best, o), a2 = __shfl_xor_sync(FULL, arg, o);
            if (b2 > best || (b2 == best && a2 < arg)) { best = b2; arg = a2; }
        }
        if (lane == 0) { s_arg[0] = arg; s_cnt[0] = 0; s_cnt[1] = 0; }
    }
    __syncthreads();
    const int ref = s_arg[0], n = p.n;
    const double *x = p.expr + (int64_t)ref * n;
    const uint16_t *perm = p.perm + (int64_t)ref * n;
    if (threadIdx.x == 0) {
        /* bin.var: cut(x, quantile(x, seq(0,1,1/3)), include.lowest=TRUE) */
        const double third = 1.0 / 3.0;
        s_b[0] = quantile7(x, perm, n, 0.0);
        s_b[1] = quantile7(x, perm, n, third);
        s_b[2] = quantile7(x, perm, n, 2.0 * third);
        s_b[3] = quantile7(x, perm, n, 1.0);
    }
    __syncthreads();
    int nup = 0, nm = 0;
    for (int i = threadIdx.x; i < n; i += blockDim.x) {
        double v = x[i];
        if (v > s_b[2]) nup++;
        else if (v > s_b[1]) nm++;
    }
    if (nup) atomicAdd(&s_cnt[0], nup);
    if (nm) atomicAdd(&s_cnt[1], nm);
    __syncthreads();
    if (threadIdx.x == 0) {
        p.counts[0] = s_cnt[0];
        p.counts[1] = s_cnt[1];
        p.counts[2] = n - s_cnt[0] - s_cnt[1];
        *p.ref_gene = ref;
        *p.status = (s_b[0] < s_b[1] && s_b[1] < s_b[2] && s_b[2] < s_b[3]) ? ST_OK : ST_BREAKS_NOT_UNIQUE;
    }
}

/* ============================================================ k_gene_tests */

__global__ void __launch_bounds__(128) k_gene_tests(GeneTestParams p)
{
    extern __shared__ __align__(16) unsigned char smem[];
    double *xs = (double *)smem; /* sorted descending; later reused as psmirnov scratch */
    __shared__ double red[40];
    __shared__ int s_dmax[3];
    if (*p.status != ST_OK) return;

    const int g = blockIdx.x, n = p.n, tid = threadIdx.x;
    const double *x = p.expr + (int64_t)g * n;
    const uint16_t *perm = p.perm + (int64_t)g * n;
    const int nUP = p.counts[0], nM = p.counts[1], nD = p.counts[2];
    /* row blocks of the expression-side labels: UP, then M, then D */
    const int off[3] = { 0, nUP, nUP + nM };
    const int len[3] = { nUP, nM, nD };

    for (int i = tid; i < n; i += blockDim.x) xs[i] = x[perm[i]];
    if (tid < 3) s_dmax[tid] = 0;
    __syncthreads();

    /* group means and variances (two-pass) */
    double mean[3], var[3];
    for (int c = 0; c < 3; c++) {
        double s = 0.0;
        for (int i = tid; i < len[c]; i += blockDim.x) s += xs[off[c] + i];
        mean[c] = len[c] > 0 ? block_sum(s, red) / (double)len[c] : EPX_NAN;
        double ss = 0.0;
        for (int i = tid; i < len[c]; i += blockDim.x) { double d = xs[off[c] + i] - mean[c]; ss += d * d; }
        var[c] = len[c] > 1 ? block_sum(ss, red) / (double)(len[c] - 1) : EPX_NAN;
    }

    /* the three comparisons of R/EpimethexAnalysis.R:184-206: (UP,M) (UP,D) (M,D) */
    const int ca[3] = { 0, 0, 1 }, cb[3] = { 1, 2, 2 };
    int guard[3], ties[3];
    for (int c = 0; c < 3; c++) {
        const int la = len[ca[c]], lb = len[cb[c]];
        const double *A = xs + off[ca[c]], *B = xs + off[cb[c]];
        /* guard, R/Functions.R:193: sd(mapply('-', a, b)) with recycling of the shorter argument */
        const int L = la > lb ? la : lb;
        int differs = 0;
        if (la > 0 && lb > 0 && L >= 2) {
            const double d0 = A[0] - B[0];
            for (int i = 1 + tid; i < L; i += blockDim.x)
                if (A[i % la] - B[i % lb] != d0) differs = 1;
        }
        differs = __syncthreads_or(differs);
        guard[c] = (la == 0 || lb == 0 || L < 2) ? -1 : (differs ? 1 : 0);

        /* KS numerator: the column is globally sorted, so counts are closed forms of the position;
         * evaluate at the end of every tie run of the whole column (extra thresholds cannot exceed the sup) */
        int dmax = 0, tie = 0;
        if (!p.test_t && la > 0 && lb > 0) {
            const int a0 = off[ca[c]], b0 = off[cb[c]];
            for (int i = tid; i < n; i += blockDim.x) {
                bool run_end = (i == n - 1) || (xs[i] != xs[i + 1]);
                if (run_end) {
                    int cA = min(max(i + 1 - a0, 0), la), cB = min(max(i + 1 - b0, 0), lb);
                    int z = cA * lb - cB * la;
                    dmax = max(dmax, z < 0 ? -z : z);
                } else {
                    /* xs[i] == xs[i+1]: a tie inside the pooled sample iff both lie in A u B, or they
                     * bracket a fully tied third group between A and B */
                    bool in_i = (i >= a0 && i < a0 + la) || (i >= b0 && i < b0 + lb);
                    bool in_n = (i + 1 >= a0 && i + 1 < a0 + la) || (i + 1 >= b0 && i + 1 < b0 + lb);
                    if (in_i && in_n) tie = 1;
                }
            }
            /* UP vs D with the whole M block tied to both neighbours */
            if (c == 1 && tid == 0 && nM > 0 && xs[off[1] - 1] == xs[off[2]]) tie = 1;
            dmax = warp_max_i(dmax);
            if ((tid & 31) == 0 && dmax) atomicMax(&s_dmax[c], dmax);
        }
        ties[c] = __syncthreads_or(tie);
    }
    __syncthreads();

    if (tid == 0) {
        double *o = p.gene + (int64_t)g * GENE_COLS;
        const double q33 = quantile7(x, perm, n, 0.33), q66 = quantile7(x, perm, n, 0.66),
                     q99 = quantile7(x, perm, n, 0.99);
        /* calcFC reads the last three rows = perc99/perc66/perc33 (reference quirk) */
        const double fu = p.fc_from_means ? mean[0] : q99, fm = p.fc_from_means ? mean[1] : q66,
                     fd = p.fc_from_means ? mean[2] : q33;
        double fc0, fc1, fc2;
        if (p.data_linear) { fc0 = epx_linear_fc(fu, fm); fc1 = epx_linear_fc(fu, fd); fc2 = epx_linear_fc(fm, fd); }
        else { fc0 = epx_log_fc(fu, fm); fc1 = epx_log_fc(fu, fd); fc2 = epx_log_fc(fm, fd); }
        o[0] = fc0; o[1] = fc1; o[2] = fc2;
        o[9] = q33; o[10] = q66; o[11] = q99;
        o[12] = mean[2]; o[13] = mean[1]; o[14] = mean[0];
        uint16_t na = 0;
        for (int c = 0; c < 3; c++) {
            const int la = len[ca[c]], lb = len[cb[c]];
            double stat = EPX_NAN, pv = EPX_NAN;
            int dn = -1;
            bool ok = false;
            if (p.test_t) {
                if (guard[c] == 1) {
                    double t, df;
                    if (epx_welch(mean[ca[c]], var[ca[c]], (double)la, mean[cb[c]], var[cb[c]], (double)lb, &t, &df) == 0) {
                        stat = t; pv = epx_t_two_sided_p(t, df); ok = true;
                    }
                }
            } else if (guard[c] != 0 && la > 0 && lb > 0) {
                dn = s_dmax[c];
                stat = (double)dn / ((double)la * (double)lb);
                if ((double)la * (double)lb < 10000.0 && !ties[c]) pv = epx_ks_p_exact((double)dn, la, lb, xs);
                else pv = epx_ks_p_asym((double)dn, (double)la, (double)lb);
                ok = true;
            }
            o[3 + c] = pv; o[6 + c] = stat;
            p.gene_dnum[(int64_t)g * 3 + c] = dn;
            if (!ok) na |= (uint16_t)((1u << (3 + c)) | (1u << (6 + c)));
        }
        p.gene_na[g] = na;
    }
}

/* ============================================================ k_probe_rank */

/* long rows (n > 1024): ctab_threads(n) threads, bucket placement (epx_ctabucket.cuh); a row whose placement fails its
 * check goes through the merge sort of epx_ctasort.cuh in the same buffers */
__global__ void __maxnreg__(80) k_probe_rank(RankParams p)
{
    extern __shared__ __align__(16) unsigned char smem[];
    const int n = p.n;
    double *s_val = (double *)smem;
    unsigned *region = (unsigned *)(smem + ctab_row_bytes(n));
    /* the row comes in as ONE bulk copy (cp.async.bulk + mbarrier) when it is 16-byte aligned: a per-thread loop of
     * loads and shared stores ran as ~35 dependent round trips to DRAM and took a quarter of the kernel */
    __shared__ __align__(8) uint64_t bar;
    const uint32_t rowbytes = (uint32_t)((n + 1) & ~1) * 8u;
    const bool bulk = ((((uintptr_t)p.meth) | ((uintptr_t)p.ld * 8u)) & 15u) == 0;
    if (threadIdx.x == 0) { mbar_init(&bar, 1); fence_barrier_init(); }
    uint32_t phase = 0;
    for (int64_t slot = blockIdx.x; slot < p.U; slot += gridDim.x) {
        const double *row = p.meth + (int64_t)p.uniq_probe[slot] * p.ld;
        __syncthreads(); /* the previous row is done with the buffers (and the barrier is initialised) */
        if (bulk) {
            if (threadIdx.x == 0) { mbar_expect_tx(&bar, rowbytes); bulk_g2s(s_val, row, rowbytes, &bar); }
            mbar_wait(&bar, phase);
            phase ^= 1;
        } else {
            for (int i = threadIdx.x; i < n; i += blockDim.x) s_val[i] = __ldg(row + i);
        }
        __syncthreads();
        const unsigned short *ord = cta_bucket_order<false>(s_val, row, n, region, p.one, p.full_rows);
        if (!ord) { /* the placement used the row's place in shared memory: stage it again for the merge sort */
            unsigned short *s_a = (unsigned short *)region, *s_b = s_a + ctasort_entries(n);
            __syncthreads();
            if (bulk) {
                if (threadIdx.x == 0) { mbar_expect_tx(&bar, rowbytes); bulk_g2s(s_val, row, rowbytes, &bar); }
                mbar_wait(&bar, phase);
                phase ^= 1;
            } else {
                for (int i = threadIdx.x; i < n; i += blockDim.x) s_val[i] = __ldg(row + i);
            }
            __syncthreads();
            /* n > 1024 here, so there is a last merge pass and it leaves the tie marks in place */
            ord = cta_sort_row<false, true>(s_val, n, s_a, s_b, p.one);
        }
        uint16_t *o = p.ord + slot * p.ldo;
        if ((p.ldo & 7) == 0 && (((uintptr_t)p.ord) & 15u) == 0) { /* 16 bytes per load and store */
            for (int k0 = threadIdx.x * 8; k0 < p.ldo; k0 += blockDim.x * 8) {
                uint4 q;
                if (k0 + 8 <= n) q = *reinterpret_cast<const uint4 *>(ord + k0);
                else {
                    unsigned w[8];
#pragma unroll
                    for (int i = 0; i < 8; i++) w[i] = k0 + i < n ? (unsigned)ord[k0 + i] : (unsigned)n; /* pad */
                    q.x = w[0] | (w[1] << 16); q.y = w[2] | (w[3] << 16); q.z = w[4] | (w[5] << 16); q.w = w[6] | (w[7] << 16);
                }
                *reinterpret_cast<uint4 *>(o + k0) = q;
            }
        } else {
            for (int k = threadIdx.x; k < p.ldo; k += blockDim.x) o[k] = k < n ? ord[k] : (unsigned short)n;
        }
    }
}

/* ============================================================ warp-per-row variants (n <= 1024) */

constexpr int SORT_WARPS = 8;

/* k_gene_sort, one warp per gene */
template <int EPS>
__global__ void __launch_bounds__(SORT_WARPS * 32, (EPS <= 16 ? 3 : 2)) k_gene_sort_w(GeneParams p)
{
    __shared__ unsigned short s_ord[SORT_WARPS][32 * EPS];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, n = p.n;
    const int nw = gridDim.x * SORT_WARPS;
    for (int g = blockIdx.x * SORT_WARPS + warp; g < p.G; g += nw) {
        const double *x = p.expr + (int64_t)g * n;
        unsigned khi[EPS], klo[EPS];
        double sum = 0.0;
#pragma unroll
        for (int e = 0; e < EPS; e++) {
            const int s = lane + 32 * e;
            const double v = __ldg(x + min(s, n - 1));
            key_halves(v, khi[e], klo[e]);
            khi[e] = ~khi[e]; klo[e] = ~klo[e]; /* descending by value */
            sum += s < n ? v : 0.0;
        }
        const double mean = warp_sum(sum) / (double)n;
        unsigned short out[EPS];
        WarpSort<EPS>::template order<true>(khi, klo, n, lane, s_ord[warp], [&](int i) { return __ldg(x + i); },
                                            p.one, out);
        int ties = 0;
        uint16_t *perm = p.perm + (int64_t)g * n;
        uint8_t *lab8 = p.lab8 + (int64_t)g * p.ldl;
#pragma unroll
        for (int r = 0; r < EPS; r++) {
            const int k = lane * EPS + r;
            if (k < n) {
                const int s = out[r] & 0x7fff;
                ties += (out[r] >> 15) & 1;
                perm[k] = (uint16_t)s;
                const int lab = k / p.m;
                lab8[s] = (uint8_t)(lab > 3 ? 3 : lab);
            }
        }
        ties = __reduce_add_sync(FULL, ties);
        double sxx = 0.0, sxc = 0.0;
#pragma unroll
        for (int e = 0; e < EPS; e++) {
            const int s = lane + 32 * e;
            if (s < n) {
                const double d = __ldg(x + s) - mean; /* L1 hit: the row was just read */
                p.xc[(int64_t)g * p.ldx + s] = d;
                sxx += d * d;
                sxc += d;
            }
        }
        sxx = warp_sum(sxx);
        sxc = warp_sum(sxc);
        if (lane == 0) {
            const int u = n - ties;
            p.nuniq[g] = u;
            double *ga = p.gene_aux + 4 * (int64_t)g;
            ga[0] = (u == 1) ? 0.0 : sxx; /* constant gene: sd is exactly 0 in R */
            ga[1] = mean; ga[2] = sxc; ga[3] = (u == 1 || sxx == 0.0) ? 0.0 : rsqrt(sxx);
        }
    }
}

/* k_gene_tests, one warp per gene: the column in the gene's descending order is staged in shared memory.
 * gene_tests_row is the work of one gene.
 * xs: 32*EPS + 2 doubles of per-warp shared memory (sorted column, later psmirnov scratch). */
constexpr int GT_WARPS = 4;

template <int EPS, bool DEFER_T>
__device__ __noinline__ void gene_tests_row(const GeneTestParams &p, const int g, const int lane, double *xs)
{
    const int n = p.n;
    const int nUP = p.counts[0], nM = p.counts[1], nD = p.counts[2];
    const int off[3] = { 0, nUP, nUP + nM };
    const int len[3] = { nUP, nM, nD };
    const int ca[3] = { 0, 0, 1 }, cb[3] = { 1, 2, 2 };
    const double *x = p.expr + (int64_t)g * n;
    const uint16_t *perm = p.perm + (int64_t)g * n;
    __syncwarp();
    /* position k = lane + 32*e of the sorted column; group by position block */
    double s0 = 0, s1 = 0, s2 = 0;
#pragma unroll
    for (int e = 0; e < EPS; e++) {
        const int k = lane + 32 * e;
        if (k < n) {
            const double v = __ldg(x + perm[k]);
            xs[k] = v;
            if (k < off[1]) s0 += v; else if (k < off[2]) s1 += v; else s2 += v;
        }
    }
    __syncwarp();
    double mean[3], var[3];
    mean[0] = len[0] > 0 ? warp_sum(s0) / (double)len[0] : EPX_NAN;
    mean[1] = len[1] > 0 ? warp_sum(s1) / (double)len[1] : EPX_NAN;
    mean[2] = len[2] > 0 ? warp_sum(s2) / (double)len[2] : EPX_NAN;
    double q0 = 0, q1 = 0, q2 = 0;
    int dmaxv[3] = { 0, 0, 0 }, tiev[3] = { 0, 0, 0 }, differs[3] = { 0, 0, 0 };
#pragma unroll
    for (int e = 0; e < EPS; e++) {
        const int k = lane + 32 * e;
        if (k < n) {
            const double v = xs[k];
            if (k < off[1]) { const double d = v - mean[0]; q0 += d * d; }
            else if (k < off[2]) { const double d = v - mean[1]; q1 += d * d; }
            else { const double d = v - mean[2]; q2 += d * d; }
            if (!p.test_t) {
                /* KS: the column is globally sorted, so counts are closed forms of the position; evaluate at
                 * the end of every tie run of the whole column (extra thresholds cannot exceed the sup) */
                const bool run_end = (k == n - 1) || (v != xs[k + 1]);
#pragma unroll
                for (int c = 0; c < 3; c++) {
                    const int la = len[ca[c]], lb = len[cb[c]], a0 = off[ca[c]], b0 = off[cb[c]];
                    if (run_end) {
                        const int cA = min(max(k + 1 - a0, 0), la), cB = min(max(k + 1 - b0, 0), lb);
                        const int z = cA * lb - cB * la;
                        dmaxv[c] = max(dmaxv[c], z < 0 ? -z : z);
                    } else {
                        const bool in_i = (k >= a0 && k < a0 + la) || (k >= b0 && k < b0 + lb);
                        const bool in_n = (k + 1 >= a0 && k + 1 < a0 + la) || (k + 1 >= b0 && k + 1 < b0 + lb);
                        if (in_i && in_n) tiev[c] = 1;
                    }
                }
            }
        }
    }
    var[0] = len[0] > 1 ? warp_sum(q0) / (double)(len[0] - 1) : EPX_NAN;
    var[1] = len[1] > 1 ? warp_sum(q1) / (double)(len[1] - 1) : EPX_NAN;
    var[2] = len[2] > 1 ? warp_sum(q2) / (double)(len[2] - 1) : EPX_NAN;
    /* guard, R/Functions.R:193: sd(mapply('-', a, b)) with recycling of the shorter argument */
    int guard[3];
#pragma unroll
    for (int c = 0; c < 3; c++) {
        const int la = len[ca[c]], lb = len[cb[c]];
        const double *A = xs + off[ca[c]], *B = xs + off[cb[c]];
        const int L = la > lb ? la : lb;
        if (la > 0 && lb > 0 && L >= 2) {
            const double d0 = A[0] - B[0];
            if (la == lb) {
                for (int i = 1 + lane; i < L; i += 32)
                    if (A[i] - B[i] != d0) differs[c] = 1;
            } else {
                for (int i = 1 + lane; i < L; i += 32)
                    if (A[i % la] - B[i % lb] != d0) differs[c] = 1;
            }
        }
        guard[c] = (la == 0 || lb == 0 || L < 2) ? -1 : (__any_sync(FULL, differs[c]) ? 1 : 0);
    }
    if (!p.test_t) {
        /* UP vs D with the whole M block tied to both neighbours */
        if (lane == 0 && nM > 0 && nUP > 0 && nD > 0 && xs[off[1] - 1] == xs[off[2]]) tiev[1] = 1;
#pragma unroll
        for (int c = 0; c < 3; c++) { dmaxv[c] = warp_max_i(dmaxv[c]); tiev[c] = __any_sync(FULL, tiev[c]); }
    }
    __syncwarp(); /* all reads of xs are done: lane 0 may reuse it as psmirnov scratch */
    if (lane == 0) {
        double *o = p.gene + (int64_t)g * GENE_COLS;
        /* from the staged sorted column: the global x[perm[.]] form is six dependent DRAM round trips on one lane */
        const double q33 = quantile7_desc(xs, n, 0.33), q66 = quantile7_desc(xs, n, 0.66),
                     q99 = quantile7_desc(xs, n, 0.99);
        /* calcFC reads the last three rows = perc99/perc66/perc33 (reference quirk) */
        const double fu = p.fc_from_means ? mean[0] : q99, fm = p.fc_from_means ? mean[1] : q66,
                     fd = p.fc_from_means ? mean[2] : q33;
        if (p.data_linear) { o[0] = epx_linear_fc(fu, fm); o[1] = epx_linear_fc(fu, fd); o[2] = epx_linear_fc(fm, fd); }
        else { o[0] = epx_log_fc(fu, fm); o[1] = epx_log_fc(fu, fd); o[2] = epx_log_fc(fm, fd); }
        o[9] = q33; o[10] = q66; o[11] = q99;
        o[12] = mean[2]; o[13] = mean[1]; o[14] = mean[0];
    }
    /* the three p-values on three lanes at once */
    uint16_t na = 0;
    if (lane < 3) {
        const int c = lane;
        const int la = len[ca[c]], lb = len[cb[c]];
        double stat = EPX_NAN, pv = EPX_NAN;
        int dn = -1;
        bool ok = false;
        const int gd = c == 0 ? guard[0] : (c == 1 ? guard[1] : guard[2]);
        if (p.test_t) {
            if (gd == 1) {
                double t, df;
                const double ma = c == 2 ? mean[1] : mean[0], va = c == 2 ? var[1] : var[0];
                const double mb = c == 0 ? mean[1] : mean[2], vb = c == 0 ? var[1] : var[2];
                if (epx_welch(ma, va, (double)la, mb, vb, (double)lb, &t, &df) == 0) {
                    /* DEFER_T: the caller turns (t, df) into the p-value later, one test per lane for a block of genes
                     * (the continued fraction is a long dependent chain: three lanes of a warp per gene left it to
                     * latency); df waits in the p-value's own cell */
                    stat = t; pv = DEFER_T ? df : epx_t_two_sided_p(t, df); ok = true;
                }
            }
        } else if (gd != 0 && la > 0 && lb > 0) {
            dn = c == 0 ? dmaxv[0] : (c == 1 ? dmaxv[1] : dmaxv[2]);
            const int tv = c == 0 ? tiev[0] : (c == 1 ? tiev[1] : tiev[2]);
            stat = (double)dn / ((double)la * (double)lb);
            ok = true;
            if ((double)la * (double)lb < 10000.0 && !tv) pv = EPX_NAN; /* exact: done serially below */
            else pv = epx_ks_p_asym((double)dn, (double)la, (double)lb);
        }
        double *o = p.gene + (int64_t)g * GENE_COLS;
        const bool exact_pending = !p.test_t && ok && (double)la * (double)lb < 10000.0 &&
                                   !(c == 0 ? tiev[0] : (c == 1 ? tiev[1] : tiev[2]));
        if (!exact_pending) o[3 + c] = pv; /* otherwise lane 0 writes it below */
        o[6 + c] = stat;
        p.gene_dnum[(int64_t)g * 3 + c] = dn;
        if (!ok) na = (uint16_t)((1u << (3 + c)) | (1u << (6 + c)));
    }
    if (!p.test_t) {
        /* exact KS p-values (small groups, no pooled ties): psmirnov2x needs the shared scratch -> serial */
#pragma unroll
        for (int c = 0; c < 3; c++) {
            const int la = len[ca[c]], lb = len[cb[c]];
            if (guard[c] != 0 && la > 0 && lb > 0 && (double)la * (double)lb < 10000.0 && !tiev[c]) {
                if (lane == 0) p.gene[(int64_t)g * GENE_COLS + 3 + c] = epx_ks_p_exact((double)dmaxv[c], la, lb, xs);
            }
        }
    }
    na |= (uint16_t)__shfl_xor_sync(FULL, (int)na, 1);
    na |= (uint16_t)__shfl_xor_sync(FULL, (int)na, 2);
    if (lane == 0) p.gene_na[g] = na;
}

template <int EPS>
__global__ void __launch_bounds__(GT_WARPS * 32) k_gene_tests_w(GeneTestParams p)
{
    extern __shared__ __align__(16) unsigned char smem[];
    if (*p.status != ST_OK) return;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    double *xs = (double *)smem + (size_t)warp * (32 * EPS + 2);
    const int nw = gridDim.x * GT_WARPS;
    for (int g = blockIdx.x * GT_WARPS + warp; g < p.G; g += nw) gene_tests_row<EPS, true>(p, g, lane, xs);
}

/* the Student-t tails k_gene_tests_w left behind as (t in the statistic column, df in the p-value column): one test per
 * THREAD. The continued fraction is a long dependent chain; on three lanes of a warp per gene it left the kernel to
 * latency (0.18 ms for 20 k genes), here 3 G of them run side by side. */
__global__ void __launch_bounds__(128) k_gene_tails(GeneTestParams p)
{
    if (*p.status != ST_OK) return;
    const int q = blockIdx.x * blockDim.x + threadIdx.x;
    if (q >= 3 * p.G) return;
    double *o = p.gene + (int64_t)(q / 3) * GENE_COLS;
    const int c = q % 3;
    const double t = o[EPX_G_STAT_UP_MID + c];
    if (!isnan(t)) o[EPX_G_P_UP_MID + c] = epx_t_two_sided_p(t, o[EPX_G_P_UP_MID + c]);
}

/* k_probe_rank, one warp per referenced probe */
template <int EPS> __host__ __device__ constexpr bool rank_by_buckets();
template <int EPS> __host__ __device__ constexpr int rank_warps() { return rank_by_buckets<EPS>() ? 4 : SORT_WARPS; }
/* per-warp shared memory: 2 row stages of 32*EPS doubles | 2 mbarriers | 32*EPS uint16 of sort scratch */
/* rows of 257..512 values are placed by buckets before a short network (WarpSort::order_bucket) */
template <int EPS> __host__ __device__ constexpr bool rank_by_buckets() { return EPS == 16; }
/* 513..1024 values: ONE row stage per warp (8 KB); it is refilled as soon as the row's words are built, and the few
 * values finish() needs for ties come from global memory (L2: the row was just read) */
template <int EPS> __host__ __device__ constexpr int rank_stages() { return rank_by_buckets<EPS>() && EPS == 32 ? 1 : 2; }
template <int EPS> __host__ __device__ constexpr size_t rank_warp_smem()
{
    return (size_t)rank_stages<EPS>() * 32 * EPS * 8 + 16 + (rank_by_buckets<EPS>() ? WarpSort<EPS>::BUCKET_SMEM : (size_t)32 * EPS * 2);
}

/* the bin.var reference gene and its group sizes (R/EpimethexAnalysis.R:134-139), by ONE warp: the last warp to leave
 * k_probe_rank_w, which by then sees the order and the distinct-value count of every gene (same result as
 * k_pick_reference, which the other sample-count ranges still launch) */
__device__ __noinline__ void pick_reference_warp(const RankParams &p, const int lane)
{
    const GeneParams &gp = p.genes;
    const int n = gp.n;
    /* which.max(#distinct values): every gene row left (count << 32 | ~gene) in *p.best with an atomicMax, so the
     * first gene with the largest count is one load away (a scan of 20 k counts by one warp cost 0.2 ms of tail) */
    const unsigned long long key = __ldcg(p.best);
    const int ref = (int)(0xffffffffu - (unsigned)(key & 0xffffffffull));
    if (lane == 0) *p.best = 0ull; /* re-armed for the next launch */
    const double *x = gp.expr + (int64_t)ref * n;
    const uint16_t *perm = gp.perm + (int64_t)ref * n;
    /* bin.var: cut(x, quantile(x, seq(0,1,1/3)), include.lowest=TRUE); type 7 on the ascending order, whose element k
     * is x[perm[n-1-k]] */
    double b[4] = { 0, 0, 0, 0 };
    if (lane < 4) {
        const double third = 1.0 / 3.0;
        const double prob = lane == 0 ? 0.0 : (lane == 1 ? third : (lane == 2 ? 2.0 * third : 1.0));
        const double index = 1.0 + (double)(n - 1) * prob;
        const double lo = floor(index), hi = ceil(index);
        double qs = x[__ldcg(perm + n - (int)lo)];
        const double xh = x[__ldcg(perm + n - (int)hi)];
        if (index > lo && xh != qs) { const double h = index - lo; qs = (1.0 - h) * qs + h * xh; }
        b[0] = qs;
    }
    b[3] = __shfl_sync(FULL, b[0], 3); b[2] = __shfl_sync(FULL, b[0], 2); b[1] = __shfl_sync(FULL, b[0], 1);
    b[0] = __shfl_sync(FULL, b[0], 0);
    int nup = 0, nm = 0;
    for (int i = lane; i < n; i += 32) {
        const double v = x[i];
        if (v > b[2]) nup++;
        else if (v > b[1]) nm++;
    }
    nup = __reduce_add_sync(FULL, nup);
    nm = __reduce_add_sync(FULL, nm);
    if (lane == 0) {
        p.counts[0] = nup; p.counts[1] = nm; p.counts[2] = n - nup - nm;
        *p.ref_gene = ref;
        *p.status = (b[0] < b[1] && b[1] < b[2] && b[2] < b[3]) ? ST_OK : ST_BREAKS_NOT_UNIQUE;
    }
}

/* GENES = false: the referenced methylation rows, ascending (the order rows of the pair kernel). GENES = true: the
 * expression rows, descending, with everything k_gene_sort_w wrote, and the last warp picks the bin.var reference gene.
 * Two instantiations of one loop rather than one kernel for both: with both row kinds in one kernel its code outgrew
 * the instruction cache and every row paid for it (measured: 0.93 -> 1.15 ms). */
template <int EPS, bool GENES>
__global__ void __launch_bounds__(rank_warps<EPS>() * 32) __maxnreg__(EPS >= 32 ? 128 : (rank_by_buckets<EPS>() ? 96 : 80)) k_probe_rank_w(RankParams p)
{
    constexpr int RW = rank_warps<EPS>();
    constexpr int N = 32 * EPS;
    constexpr bool BUCKETS = rank_by_buckets<EPS>();
    constexpr int STAGES = rank_stages<EPS>();
    extern __shared__ __align__(128) unsigned char smem[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, n = p.n;
    unsigned char *base = smem + (size_t)warp * rank_warp_smem<EPS>();
    double *stage0 = (double *)base;                         /* stage s at stage0 + s * N */
    uint64_t *bar = (uint64_t *)(base + (size_t)STAGES * N * 8);
    unsigned short *s_ord = (unsigned short *)(base + (size_t)STAGES * N * 8 + 16);
    const int nw = (int)gridDim.x * RW;
    const int64_t G = GENES ? p.genes.G : 0, total = GENES ? G : p.U;
    int32_t *const queue = GENES ? p.gene_queue : p.queue;
    /* The row comes in as ONE bulk copy (cp.async.bulk, SASS UBLKCP) into a two-stage ring: the copy of the NEXT row is
     * in flight while this one is sorted, the 16 loads of a lane become conflict-free LDS with immediate offsets, and
     * the values the exact path needs (ties) are already on chip. Rows must be 16-byte aligned with a 16-byte multiple
     * length (the session's methylation layout: ld % 2 == 0); expression rows (n doubles apart: not aligned when n is
     * odd) and unaligned borrowed matrices are copied by the lanes. */
    const bool bulk = !GENES && ((((uintptr_t)p.meth) | ((uintptr_t)p.ld * 8u)) & 15u) == 0 && p.ld <= N;
    const uint32_t rowb = (uint32_t)p.ld * 8u;
    if (lane == 0) { mbar_init(&bar[0], 1); mbar_init(&bar[1], 1); fence_barrier_init(); }
    unsigned *s_bkt = reinterpret_cast<unsigned *>(s_ord); /* BUCKETS: counters, then slots */
    if constexpr (BUCKETS) s_bkt[WarpSort<EPS>::NB + lane] = 0u;
    __syncwarp();
    uint32_t ph0 = 0, ph1 = 0;
    int fill = 0, use = 0;
    unsigned by_bulk = 0;    /* bit s: stage s was filled by a bulk copy (its mbarrier has to be waited for) */
    auto issue = [&](int64_t slot) {
        double *dst = stage0 + (fill ? N : 0);
        if (!GENES && bulk) {
            const double *row = p.meth + (int64_t)__ldg(p.uniq_probe + slot) * p.ld;
            if (lane == 0) { mbar_expect_tx(&bar[fill], rowb); bulk_g2s(dst, row, rowb, &bar[fill]); }
            by_bulk |= 1u << fill;
        } else {
            const double *row = GENES ? p.genes.expr + slot * n : p.meth + (int64_t)__ldg(p.uniq_probe + slot) * p.ld;
#pragma unroll
            for (int e = 0; e < EPS; e++) if (lane + 32 * e < n) dst[lane + 32 * e] = __ldg(row + lane + 32 * e);
            by_bulk &= ~(1u << fill);
        }
        if (STAGES == 2) fill ^= 1;
    };
    /* Rows are handed out by a ticket counter (tickets 0 .. G-1: expression rows, then the referenced probes): rows
     * with tied or nearly tied values take the longer exact path, and a static split of a one-wave grid lost 6 % to
     * that imbalance. The ticket of the row after next is drawn while this row is sorted, so neither the counter
     * nor the copy is ever waited for. */
    int64_t slot = (int64_t)blockIdx.x * RW + warp;
    if (slot < total) issue(slot);
    while (slot < total) {
        int nslot = 0;
        if (lane == 0) nslot = nw + atomicAdd(queue, 1);
        const int64_t next = __shfl_sync(FULL, nslot, 0);
        __syncwarp(); /* every lane is done with the stage about to be refilled */
        if (STAGES == 2 && next < total) issue(next);
        const double *row = stage0 + (use ? N : 0);
        if ((by_bulk >> use) & 1u) { if (use == 0) { mbar_wait(&bar[0], ph0); ph0 ^= 1; } else { mbar_wait(&bar[1], ph1); ph1 ^= 1; } }
        else __syncwarp();
        if (STAGES == 2) use ^= 1;
        constexpr bool gene = GENES; /* descending order for expression rows (R/EpimethexAnalysis.R:117-122) */
        unsigned short out[EPS];
        const double *vals = row; /* the row's values after the sort (expression rows: sums, centring) */
        if constexpr (BUCKETS && STAGES == 1) {
            /* one stage: the row is read twice from shared memory while its words are built, then the stage goes to
             * the next row's copy; ties are settled on values fetched from global memory */
            const double *grow = GENES ? p.genes.expr + slot * n : p.meth + (int64_t)__ldg(p.uniq_probe + slot) * p.ld;
            WarpSort<EPS>::order_bucket([&](int e) { const int s = lane + 32 * e; return row[(2 * e < EPS || s < n) ? s : 0]; },
                                        [&]() { __syncwarp(); if (next < total) issue(next); }, n, lane, s_bkt,
                                        [&](int i) { return __ldg(grow + i); }, GENES, p.one, out, p.full_rows);
            vals = grow;
        } else if constexpr (BUCKETS) {
            double x[EPS];
#pragma unroll
            for (int e = 0; e < EPS; e++) {
                const int s = lane + 32 * e;
                x[e] = row[(2 * e < EPS || s < n) ? s : 0];
            }
            WarpSort<EPS>::order_bucket([&](int e) { return x[e]; }, []() {}, n, lane, s_bkt, [&](int i) { return row[i]; }, GENES,
                                        p.one, out, p.full_rows);
        } else {
        unsigned khi[EPS], klo[EPS];
        if (!gene) {
#pragma unroll
            for (int e = 0; e < EPS; e++) {
                /* slots past the row (upper half of the slots only) take element 0: any real key of the row will do */
                const int s = lane + 32 * e;
                key_halves(row[(2 * e < EPS || s < n) ? s : 0], khi[e], klo[e]);
            }
        } else {
#pragma unroll
            for (int e = 0; e < EPS; e++) {
                const int s = lane + 32 * e;
                key_halves(row[(2 * e < EPS || s < n) ? s : 0], khi[e], klo[e]);
                khi[e] = ~khi[e]; klo[e] = ~klo[e];
            }
        }
        /* pad positions (k >= n) carry index n: the pair kernel's inert table entry */
        WarpSort<EPS>::template order<GENES>(khi, klo, n, lane, s_ord, [&](int i) { return row[i]; }, p.one, out);
        }
        if (!gene) {
            uint16_t *o = p.ord + slot * p.ldo + lane * EPS;
            if (EPS >= 8) {
#pragma unroll
                for (int c = 0; c < EPS / 8; c++) {
                    if (p.ldo == N || lane * EPS + c * 8 < p.ldo) {
                        uint4 q;
                        q.x = out[c * 8 + 0] | ((unsigned)out[c * 8 + 1] << 16);
                        q.y = out[c * 8 + 2] | ((unsigned)out[c * 8 + 3] << 16);
                        q.z = out[c * 8 + 4] | ((unsigned)out[c * 8 + 5] << 16);
                        q.w = out[c * 8 + 6] | ((unsigned)out[c * 8 + 7] << 16);
                        *reinterpret_cast<uint4 *>(o + c * 8) = q;
                    }
                }
            } else {
#pragma unroll
                for (int r = 0; r < EPS; r++)
                    if (lane * EPS + r < p.ldo) o[r] = out[r];
            }
        } else {
            /* what k_gene_sort_w writes: the order, the methylation-side labels by rank (:323-325), the centred row and
             * its sums for Pearson (:404-417), the distinct-value count for the reference pick */
            const GeneParams &gp = p.genes;
            const int64_t g = slot;
            const int m1 = gp.m, m2 = 2 * gp.m, m3 = 3 * gp.m;
            int ties = 0;
            uint16_t *perm = gp.perm + g * n;
            uint8_t *lab8 = gp.lab8 + g * gp.ldl;
            double sum = 0.0;
#pragma unroll
            for (int e = 0; e < EPS; e++) { const int s = lane + 32 * e; if (s < n) sum += vals[s]; }
            if constexpr (BUCKETS) {
                /* order and labels go through shared memory (the sort's scratch is free): the labels are indexed by
                 * SAMPLE, and 473 one-byte stores to random places of a global row cost a sector each; from shared
                 * memory the label row leaves as one 16-byte store per lane and the order as coalesced 2-byte stores */
                unsigned short *s16 = reinterpret_cast<unsigned short *>(s_bkt);
                unsigned char *s8 = reinterpret_cast<unsigned char *>(s_bkt + N / 2);
                __syncwarp();
#pragma unroll
                for (int r = 0; r < EPS; r++) {
                    const int k = lane * EPS + r;
                    if (k < n) {
                        const int s = out[r] & 0x7fff;
                        ties += (out[r] >> 15) & 1;
                        s16[k] = (unsigned short)s;
                        s8[s] = (unsigned char)((k >= m1) + (k >= m2) + (k >= m3)); /* = min(k / m, 3); m = 0 only for n < 3 (refused) */
                    }
                }
                __syncwarp();
                for (int i = lane; i < n; i += 32) perm[i] = s16[i];
                if (lane * 16 < (int)gp.ldl) /* the bytes behind the n-th label are never looked at */
                    *reinterpret_cast<uint4 *>(lab8 + lane * 16) = *reinterpret_cast<const uint4 *>(s8 + lane * 16);
                __syncwarp();
            } else {
#pragma unroll
            for (int r = 0; r < EPS; r++) {
                const int k = lane * EPS + r;
                if (k < n) {
                    const int s = out[r] & 0x7fff;
                    ties += (out[r] >> 15) & 1;
                    perm[k] = (uint16_t)s;
                    lab8[s] = (uint8_t)((k >= m1) + (k >= m2) + (k >= m3)); /* = min(k / m, 3); m = 0 only for n < 3 (refused) */
                }
            }
            }
            ties = __reduce_add_sync(FULL, ties);
            const double mean = warp_sum(sum) / (double)n;
            double sxx = 0.0, sxc = 0.0;
#pragma unroll
            for (int e = 0; e < EPS; e++) {
                const int s = lane + 32 * e;
                if (s < n) {
                    const double d = vals[s] - mean;
                    gp.xc[g * gp.ldx + s] = d;
                    sxx += d * d;
                    sxc += d;
                }
            }
            sxx = warp_sum(sxx);
            sxc = warp_sum(sxc);
            if (lane == 0) {
                const int u = n - ties;
                gp.nuniq[g] = u;
                atomicMax(p.best, ((unsigned long long)(unsigned)u << 32) | (unsigned long long)(0xffffffffu - (unsigned)g));
                double *ga = gp.gene_aux + 4 * g;
                ga[0] = (u == 1) ? 0.0 : sxx; /* constant gene: sd is exactly 0 in R */
                ga[1] = mean; ga[2] = sxc; ga[3] = (u == 1 || sxx == 0.0) ? 0.0 : rsqrt(sxx);
            }
        }
        slot = next;
    }
    /* the last warp to leave re-arms the queue for the next launch and, with every gene ordered, picks the reference */
    int last = 0;
    if (lane == 0) {
        __threadfence();
        last = atomicAdd(queue + 1, 1) == nw - 1;
        if (last) { queue[0] = 0; queue[1] = 0; }
    }
    if (GENES && G > 0 && __shfl_sync(FULL, last, 0)) {
        __threadfence();
        pick_reference_warp(p, lane);
    }
}

/* k_probe_rank_w2: rows of 513..1024 values (EPIC-array cohorts: n = 1,000). One warp per row: the row is staged in
 * shared memory, its two halves are ordered by the 512-value register sort (a 1,024-value network costs 3.4x as
 * many instructions as a 512-value one and needs 128 registers), and one merge-path pass — lane L produces outputs
 * [32L, 32L + 32) after a binary search for its split point; the left half wins ties, which keeps equal values in
 * ascending sample index — writes the order row with its tie marks. Same output as k_probe_rank_w<32>. */
constexpr int RANK2_WARPS = 4;
/* values | run indices | merged; the bucket scratch of the two half sorts lies over `merged` and the tail behind it */
constexpr int RANK2_WARP_SMEM = 1024 * 8 + 1024 * 2 + (int)WarpSort<16>::BUCKET_SMEM;
static_assert(WarpSort<16>::BUCKET_SMEM >= 1024 * 2, "the merged row must fit the scratch it shares");

__global__ void __launch_bounds__(RANK2_WARPS * 32, 4) k_probe_rank_w2(RankParams p)
{
    extern __shared__ __align__(16) unsigned char smem[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, n = p.n;
    double *vals = (double *)(smem + (size_t)warp * RANK2_WARP_SMEM);
    unsigned short *runs = (unsigned short *)(vals + 1024);
    unsigned short *merged = runs + 1024;
    unsigned short *scratch = merged; /* WarpSort's scratch is dead before the merge starts */
    const int nw = (int)gridDim.x * RANK2_WARPS;
    const int la = (n + 1) / 2, lb = n - la; /* two halves of 257..512 values (n = 513: 257 + 256) */
    if (lane < 32) reinterpret_cast<unsigned *>(scratch)[WarpSort<16>::NB + lane] = 0u; /* idle counters of order_bucket */
    int64_t slot = (int64_t)blockIdx.x * RANK2_WARPS + warp;
    while (slot < p.U) {
        const double *row = p.meth + (int64_t)p.uniq_probe[slot] * p.ld;
        __syncwarp();
#pragma unroll
        for (int e = 0; e < 32; e++) {
            const int s = lane + 32 * e;
            if (s < n) vals[s] = __ldg(row + s);
        }
        __syncwarp();
#pragma unroll 1
        for (int c = 0; c < 2; c++) {
            const int base = c == 0 ? 0 : la, nc = c == 0 ? la : lb;
            unsigned short out[16];
            if (nc > 256) { /* bucket placement + short network (WarpSort::order_bucket) */
                double x[16];
#pragma unroll
                for (int e = 0; e < 16; e++) x[e] = vals[base + min(lane + 32 * e, nc - 1)];
                WarpSort<16>::order_bucket([&](int e) { return x[e]; }, []() {}, nc, lane, reinterpret_cast<unsigned *>(scratch),
                                           [&](int i) { return vals[base + i]; }, false, p.one, out, p.full_rows);
            } else { /* n = 513 only: the 256-value half */
                unsigned khi[16], klo[16];
#pragma unroll
                for (int e = 0; e < 16; e++) key_halves(vals[base + min(lane + 32 * e, nc - 1)], khi[e], klo[e]);
                WarpSort<16>::template order<false>(khi, klo, nc, lane, scratch, [&](int i) { return vals[base + i]; },
                                                    p.one, out);
            }
#pragma unroll
            for (int r = 0; r < 16; r++) {
                const int k = lane * 16 + r;
                if (k < nc) runs[base + k] = (unsigned short)(base + (out[r] & 0x7fff));
            }
        }
        __syncwarp();
        /* merge path: of the first d outputs, i come from A (A wins ties) */
        const unsigned short *A = runs, *B = runs + la;
        /* 33 outputs per lane, not 32: with 32 the lanes' 16-bit stores (and their reads of the runs) start 64 bytes
         * apart and fall on two banks */
        const int d = min(lane * 33, n), d_end = min(d + 33, n);
        int lo = max(0, d - lb), hi = min(d, la);
        while (lo < hi) {
            const int mid = (lo + hi) >> 1;
            if (!(vals[B[d - 1 - mid]] < vals[A[mid]])) lo = mid + 1; else hi = mid;
        }
        int pa = lo, pb = d - lo;
        unsigned short xa = A[min(pa, la - 1)], xb = B[min(pb, lb - 1)];
        double va = vals[xa], vb = vals[xb];
        /* the tie mark of an output (bit 15: the next output holds an equal value) is known one step later, so every
         * store is delayed by one step; the last output of a lane's slice is compared with the first of the next lane's */
        unsigned short prev_x = 0;
        double prev_v = 0.0, first_v = 0.0;
        for (int k = d; k < d_end; k++) {
            const bool take_a = pb >= lb || (pa < la && !(vb < va));
            const unsigned short cur_x = take_a ? xa : xb;
            const double cur_v = take_a ? va : vb;
            if (k > d) merged[k - 1] = (unsigned short)(prev_x | (cur_v == prev_v ? 0x8000u : 0u));
            else first_v = cur_v;
            prev_x = cur_x; prev_v = cur_v;
            pa += take_a ? 1 : 0;
            pb += take_a ? 0 : 1;
            const unsigned short nx = take_a ? A[min(pa, la - 1)] : B[min(pb, lb - 1)];
            const double nv = vals[nx];
            xa = take_a ? nx : xa; va = take_a ? nv : va;
            xb = take_a ? xb : nx; vb = take_a ? vb : nv;
        }
        const double next_first = __shfl_down_sync(FULL, first_v, 1);
        if (d < d_end) merged[d_end - 1] = (unsigned short)(prev_x | ((d_end < n && prev_v == next_first) ? 0x8000u : 0u));
        __syncwarp();
        /* the padded row (pad positions carry index n), 16 bytes per store */
        uint16_t *o = p.ord + slot * p.ldo;
#pragma unroll
        for (int c = 0; c < 4; c++) {
            const int k0 = (lane + 32 * c) * 8;
            if (k0 < p.ldo) {
                uint4 q;
                if (k0 + 8 <= n) q = *reinterpret_cast<const uint4 *>(merged + k0);
                else {
                    unsigned w[8];
#pragma unroll
                    for (int i = 0; i < 8; i++) w[i] = k0 + i < n ? (unsigned)merged[k0 + i] : (unsigned)n;
                    q.x = w[0] | (w[1] << 16); q.y = w[2] | (w[3] << 16); q.z = w[4] | (w[5] << 16); q.w = w[6] | (w[7] << 16);
                }
                *reinterpret_cast<uint4 *>(o + k0) = q;
            }
        }
        int nslot = 0;
        if (lane == 0) nslot = nw + atomicAdd(p.queue, 1);
        slot = __shfl_sync(FULL, nslot, 0);
    }
    if (lane == 0) { /* the last warp to leave re-arms the queue for the next launch */
        __threadfence();
        if (atomicAdd(p.queue + 1, 1) == nw - 1) { p.queue[0] = 0; p.queue[1] = 0; }
    }
}

/* ============================================================ k_pair_stats (warp per pair, n <= 1024) */

constexpr int PAIR_WARPS = 8;

__host__ __device__ inline size_t pair_warp_smem(int n)
{
    const int npad = (n + 1) & ~1;
    /* xc | row | 8 median slots | deferred Student-t arguments (r, 3 t, 3 df per pair of the item) | labels */
    return (size_t)16 * npad + 64 + (size_t)ITEM_PAIRS * 7 * 8 + (size_t)((n + 15) & ~15);
}

/* EPL: elements per lane of the sample-major pass (ceil(n/32) rounded up to an instantiated value);
 * EPW: positions per lane of the sorted walk = the power of two k_probe_rank_w blocks its output by. */
template <int EPL, int EPW, bool TMODE>
__global__ void __launch_bounds__(PAIR_WARPS * 32) k_pair_stats(PairParams p)
{
    extern __shared__ __align__(16) unsigned char smem[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int n = p.n, m = p.m;
    const int npad = (n + 1) & ~1;
    unsigned char *base = smem + (size_t)warp * pair_warp_smem(n);
    double *s_xc = (double *)base;
    double *s_row = s_xc + npad;
    double *s_med = s_row + npad;                 /* 6 used */
    double *s_r = s_med + 8;                      /* [ITEM_PAIRS] Pearson r (NaN = NA) */
    double *s_t = s_r + ITEM_PAIRS;               /* [ITEM_PAIRS][3] Welch t (NaN = NA) */
    double *s_df = s_t + 3 * ITEM_PAIRS;          /* [ITEM_PAIRS][3] */
    uint8_t *s_lab = (uint8_t *)(s_df + 3 * ITEM_PAIRS);

    const int gw = blockIdx.x * PAIR_WARPS + warp, nw = gridDim.x * PAIR_WARPS;
    const double dm = (double)m, dn_ = (double)n;
    const int t1 = (m - 1) / 2 + 1, t2 = m / 2 + 1; /* within-group counts at which the two middle values pass */
    int cur_gene = -1;
    double sxx = 0.0;
    int pm[6] = { 0, 0, 0, 0, 0, 0 };
    const uint16_t *perm_g = nullptr;

    for (int it = gw; it < p.n_items; it += nw) {
        const Item item = p.items[it];
        if (item.gene != cur_gene) {
            cur_gene = item.gene;
            __syncwarp();
            const double *xc = p.xc + (int64_t)cur_gene * p.ldx;
            const uint8_t *lab = p.lab8 + (int64_t)cur_gene * p.ldl;
            for (int s = lane; s < n; s += 32) { s_xc[s] = xc[s]; s_lab[s] = lab[s]; }
            sxx = p.gene_aux[4 * (int64_t)cur_gene];
            perm_g = p.perm + (int64_t)cur_gene * n;
            if (m >= 2) {
                pm[0] = perm_g[0]; pm[1] = perm_g[1]; pm[2] = perm_g[m]; pm[3] = perm_g[m + 1];
                pm[4] = perm_g[2 * m]; pm[5] = perm_g[2 * m + 1];
            }
            __syncwarp();
        }
        for (int jj = 0; jj < item.count; jj++) {
            const int j = item.begin + jj;
            const int pi = p.probe_idx[j];
            if (pi < 0) { /* annotated probe without data: all-NA column (R/EpimethexAnalysis.R:375-394) */
                if (lane < PAIR_COLS) p.pair[(int64_t)j * PAIR_COLS + lane] = EPX_NAN;
                if (lane < 3) { p.pair_stat[(int64_t)j * 3 + lane] = EPX_NAN; p.pair_dnum[(int64_t)j * 3 + lane] = -1; }
                if (lane == 0) {
                    p.pair_na[j] = (uint16_t)((1u << PAIR_COLS) - 1);
                    s_r[jj] = EPX_NAN;
                    if (TMODE) { s_t[jj * 3] = EPX_NAN; s_t[jj * 3 + 1] = EPX_NAN; s_t[jj * 3 + 2] = EPX_NAN; }
                }
                continue;
            }
            const double *row = p.meth + (int64_t)pi * p.ld;
            const uint16_t *ord = p.ord + (int64_t)p.pair_slot[j] * p.ldo;

            /* ---- pass A: the gather. sample-major, coalesced; label = the gene's rank third ---- */
            double y[EPL];
            unsigned long long labs = 0; /* 2 bits per element */
            double s_all = 0.0, sg0 = 0.0, sg1 = 0.0, sg2 = 0.0;
            __syncwarp(); /* previous pair's readers of s_row are done */
#pragma unroll
            for (int e = 0; e < EPL; e++) {
                const int s = lane + 32 * e;
                double v = 0.0;
                unsigned lab = 3;
                if (s < n) {
                    v = __ldg(row + s);
                    lab = s_lab[s];
                    s_row[s] = v;
                    s_all += v;
                    if (TMODE) { sg0 += lab == 0 ? v : 0.0; sg1 += lab == 1 ? v : 0.0; sg2 += lab == 2 ? v : 0.0; }
                }
                y[e] = v;
                labs |= (unsigned long long)lab << (2 * e);
            }
            /* the probe's ascending order, EPW consecutive positions per lane (as k_probe_rank_w wrote it) */
            unsigned short o[EPW];
            if (EPW >= 8) {
#pragma unroll
                for (int c = 0; c < EPW / 8; c++) {
                    uint4 q = make_uint4(0xffffffffu, 0xffffffffu, 0xffffffffu, 0xffffffffu);
                    if (lane * EPW + c * 8 < p.ldo) q = __ldg(reinterpret_cast<const uint4 *>(ord + lane * EPW + c * 8));
                    o[c * 8 + 0] = (unsigned short)q.x; o[c * 8 + 1] = (unsigned short)(q.x >> 16);
                    o[c * 8 + 2] = (unsigned short)q.y; o[c * 8 + 3] = (unsigned short)(q.y >> 16);
                    o[c * 8 + 4] = (unsigned short)q.z; o[c * 8 + 5] = (unsigned short)(q.z >> 16);
                    o[c * 8 + 6] = (unsigned short)q.w; o[c * 8 + 7] = (unsigned short)(q.w >> 16);
                }
            } else {
#pragma unroll
                for (int e = 0; e < EPW; e++) o[e] = lane * EPW + e < n ? __ldg(ord + lane * EPW + e) : (unsigned short)0xffffu;
            }
            s_all = warp_sum(s_all);
            double mg0 = 0, mg1 = 0, mg2 = 0;
            if (TMODE) { mg0 = warp_sum(sg0) / dm; mg1 = warp_sum(sg1) / dm; mg2 = warp_sum(sg2) / dm; }
            const double mean_all = s_all / dn_;
            const double y_first = __shfl_sync(FULL, y[0], 0);
            bool same = true;
            double syy = 0.0, sxy = 0.0, ss0 = 0.0, ss1 = 0.0, ss2 = 0.0;
#pragma unroll
            for (int e = 0; e < EPL; e++) {
                const int s = lane + 32 * e;
                if (s < n) {
                    const double d = y[e] - mean_all;
                    syy += d * d;
                    sxy += s_xc[s] * d;
                    same = same && (y[e] == y_first);
                    if (TMODE) {
                        const unsigned lab = (unsigned)(labs >> (2 * e)) & 3u;
                        const double mgl = lab == 0 ? mg0 : (lab == 1 ? mg1 : mg2);
                        const double dg = y[e] - mgl;
                        const double q = dg * dg;
                        ss0 += lab == 0 ? q : 0.0; ss1 += lab == 1 ? q : 0.0; ss2 += lab == 2 ? q : 0.0;
                    }
                }
            }
            syy = warp_sum(syy);
            sxy = warp_sum(sxy);
            if (TMODE) { ss0 = warp_sum(ss0); ss1 = warp_sum(ss1); ss2 = warp_sum(ss2); }
            if (__all_sync(FULL, same)) syy = 0.0; /* constant row: sd is exactly 0 in R -> NA */
            __syncwarp(); /* s_row complete */

            /* ---- sorted walk ---- */
            unsigned long long wl = 0; /* labels along the walk, 2 bits each */
            unsigned cnt = 0;          /* 3 x 10-bit label counts */
            bool anytie = false;
#pragma unroll
            for (int e = 0; e < EPW; e++) {
                const int k = lane * EPW + e;
                unsigned lab = 3;
                if (k < n) {
                    lab = s_lab[o[e] & 0x7fffu];
                    anytie = anytie || (o[e] & 0x8000u);
                }
                wl |= (unsigned long long)lab << (2 * e);
                if (lab < 3) cnt += 1u << (10 * lab);
            }
            unsigned inc = cnt;
#pragma unroll
            for (int d = 1; d < 32; d <<= 1) {
                unsigned t = __shfl_up_sync(FULL, inc, d);
                if (lane >= d) inc += t;
            }
            const unsigned excl = inc - cnt;
            int c0 = excl & 1023, c1 = (excl >> 10) & 1023, c2 = (excl >> 20) & 1023;
            int d01 = 0, d02 = 0, d12 = 0;
#pragma unroll
            for (int e = 0; e < EPW; e++) {
                const int k = lane * EPW + e;
                if (k < n) {
                    const unsigned lab = (unsigned)(wl >> (2 * e)) & 3u;
                    const int s = o[e] & 0x7fffu;
                    int c = -1;
                    if (lab == 0) c = ++c0; else if (lab == 1) c = ++c1; else if (lab == 2) c = ++c2;
                    if (c == t1) s_med[lab * 2] = s_row[s];
                    if (c == t2) s_med[lab * 2 + 1] = s_row[s];
                    if (!TMODE && !(o[e] & 0x8000u)) { /* end of a tie run */
                        d01 = max(d01, abs(c0 - c1)); d02 = max(d02, abs(c0 - c2)); d12 = max(d12, abs(c1 - c2));
                    }
                }
            }
            if (!TMODE) { d01 = warp_max_i(d01); d02 = warp_max_i(d02); d12 = warp_max_i(d12); }
            __syncwarp();
            /* labels: 0 UP, 1 Medium, 2 Down */
            const double medUP = (s_med[0] + s_med[1]) * 0.5, medMID = (s_med[2] + s_med[3]) * 0.5,
                         medDOWN = (s_med[4] + s_med[5]) * 0.5;

            /* ---- guard of calculateTtest (R/Functions.R:193): are all rank-paired differences identical? ---- */
            int guard[3] = { -1, -1, -1 };
            if (m >= 2) {
                const double a0 = s_row[pm[0]], a1 = s_row[pm[1]], b0 = s_row[pm[2]], b1 = s_row[pm[3]],
                             e0 = s_row[pm[4]], e1 = s_row[pm[5]];
                const double dd[3] = { a0 - b0, a0 - e0, b0 - e0 };
                const bool eq[3] = { dd[0] == a1 - b1, dd[1] == a1 - e1, dd[2] == b1 - e1 };
                const int offA[3] = { 0, 0, m }, offB[3] = { m, 2 * m, 2 * m };
#pragma unroll
                for (int c = 0; c < 3; c++) {
                    if (!eq[c]) { guard[c] = 1; continue; }
                    int differs = 0; /* rare: walk the whole group */
                    for (int i = 2 + lane; i < m; i += 32)
                        if (s_row[perm_g[offA[c] + i]] - s_row[perm_g[offB[c] + i]] != dd[c]) differs = 1;
                    guard[c] = __any_sync(FULL, differs) ? 1 : 0;
                }
            }

            /* ---- tests. Student-t tails (Welch, Pearson) are deferred to the end of the item so that up to
             * 32 of them run on 32 lanes at once instead of one lane carrying a continued fraction ---- */
            double stat[3] = { EPX_NAN, EPX_NAN, EPX_NAN }, pv[3] = { EPX_NAN, EPX_NAN, EPX_NAN };
            int dnum[3] = { -1, -1, -1 };
            unsigned na = 0;
            double r = EPX_NAN;
            const bool r_ok = (sxx != 0.0 && syy != 0.0 && n >= 3);
            if (r_ok) {
                r = sxy / (sqrt(sxx) * sqrt(syy));
                r = r > 1.0 ? 1.0 : (r < -1.0 ? -1.0 : r);
            } else na |= (1u << EPX_PEARSON_R) | (1u << EPX_PEARSON_P);
            if (TMODE) {
                const double mg[3] = { mg0, mg1, mg2 };
                const double vg[3] = { ss0 / (dm - 1.0), ss1 / (dm - 1.0), ss2 / (dm - 1.0) };
                const int ga[3] = { 0, 0, 1 }, gb[3] = { 1, 2, 2 };
                double df[3] = { EPX_NAN, EPX_NAN, EPX_NAN };
#pragma unroll
                for (int c = 0; c < 3; c++) {
                    bool ok = false;
                    if (guard[c] == 1)
                        ok = epx_welch(mg[ga[c]], vg[ga[c]], dm, mg[gb[c]], vg[gb[c]], dm, &stat[c], &df[c]) == 0;
                    if (!ok) { na |= 1u << (EPX_P_UP_MID + c); stat[c] = EPX_NAN; }
                }
                if (lane < 3) {
                    s_t[jj * 3 + lane] = lane == 0 ? stat[0] : (lane == 1 ? stat[1] : stat[2]);
                    s_df[jj * 3 + lane] = lane == 0 ? df[0] : (lane == 1 ? df[1] : df[2]);
                }
            } else {
                int ties[3] = { 0, 0, 0 };
                if (p.exact_ok && __any_sync(FULL, anytie)) {
                    /* small groups with tied values: ks.test drops to the asymptotic p-value when the POOLED
                     * pair of groups holds a tie. Rare and tiny (m < 100): one lane walks the order. */
                    if (lane == 0) {
                        int r0 = 0, r1 = 0, r2 = 0;
                        for (int k = 0; k < n; k++) {
                            const unsigned short v = ord[k];
                            const unsigned lab = s_lab[v & 0x7fffu];
                            r0 += lab == 0; r1 += lab == 1; r2 += lab == 2;
                            if (!(v & 0x8000u)) {
                                if (r0 + r1 >= 2) ties[0] = 1;
                                if (r0 + r2 >= 2) ties[1] = 1;
                                if (r1 + r2 >= 2) ties[2] = 1;
                                r0 = r1 = r2 = 0;
                            }
                        }
                    }
#pragma unroll
                    for (int c = 0; c < 3; c++) ties[c] = __shfl_sync(FULL, ties[c], 0);
                }
                const int dd[3] = { d01, d02, d12 };
#pragma unroll
                for (int c = 0; c < 3; c++) {
                    if (guard[c] != 0) { /* difference != 0 || is.na(difference) */
                        dnum[c] = dd[c] * m;
                        stat[c] = (double)dnum[c] / (dm * dm);
                        pv[c] = (p.exact_ok && !ties[c]) ? p.lut_exact[dd[c]] : p.lut_asym[dd[c]];
                    } else na |= 1u << (EPX_P_UP_MID + c);
                }
            }
            if (lane == 0) s_r[jj] = r;

            /* ---- write the columns (R/Functions.R:224-226 order); deferred p-values come later ---- */
            double out = 0.0;
            bool wr = lane < PAIR_COLS - 1; /* column 10 (Pearson p) is always deferred */
            switch (lane) {
            case 0: out = medDOWN; break;
            case 1: out = medMID; break;
            case 2: out = medUP; break;
            case 3: out = medUP - medMID; break;
            case 4: out = medUP - medDOWN; break;
            case 5: out = medMID - medDOWN; break;
            case 6: out = pv[0]; wr = !TMODE; break;
            case 7: out = pv[1]; wr = !TMODE; break;
            case 8: out = pv[2]; wr = !TMODE; break;
            case 9: out = r; break;
            default: break;
            }
            if (wr) p.pair[(int64_t)j * PAIR_COLS + lane] = out;
            if (lane < 3) {
                p.pair_stat[(int64_t)j * 3 + lane] = lane == 0 ? stat[0] : (lane == 1 ? stat[1] : stat[2]);
                p.pair_dnum[(int64_t)j * 3 + lane] = lane == 0 ? dnum[0] : (lane == 1 ? dnum[1] : dnum[2]);
            }
            if (lane == 0) p.pair_na[j] = (uint16_t)na;
        }

        /* ---- deferred Student-t tails of the whole item, one per lane ---- */
        __syncwarp();
        constexpr int PER = TMODE ? 4 : 1;
        for (int q = lane; q < item.count * PER; q += 32) {
            const int jj = q / PER, slot = q % PER;
            double pq = EPX_NAN;
            int col = EPX_PEARSON_P;
            if (TMODE && slot < 3) {
                const double t = s_t[jj * 3 + slot];
                if (!isnan(t)) pq = welch_p(p, t, s_df[jj * 3 + slot]);
                col = EPX_P_UP_MID + slot;
            } else {
                const double rr = s_r[jj];
                if (!isnan(rr)) pq = epx_pearson_p_tab(rr, p.pt_tab, p.pt_n, dn_ - 2.0);
            }
            p.pair[(int64_t)(item.begin + jj) * PAIR_COLS + col] = pq;
        }
        __syncwarp();
    }
}


/* ============================================================ k_pair_stats_cta (CTA per pair, n > 1024) */

/* Same statistics as k_pair_stats2 for rows too long for one warp (pan-cancer: n = 9,000). One CTA per pair:
 * row, labels and the probe's order are staged in shared memory (11n bytes); the gene's centred expression is
 * read from global memory (it stays in L2 across the gene's probes). Correct and general (n <= 16,384), not yet
 * tuned: no bulk-copy pipeline, one pair per CTA at a time. */
constexpr int BIG_THREADS = 512;

__device__ __forceinline__ int block_max_i(int v, int *red)
{
    v = warp_max_i(v);
    __syncthreads();
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = v;
    __syncthreads();
    int r = red[0];
    for (int w = 1; w < (int)(blockDim.x >> 5); w++) r = max(r, red[w]);
    return r;
}

/* three sums over the block at once; every thread gets the results. red: >= 52 doubles of shared memory */
__device__ __forceinline__ void block_sum3w(double &a, double &b, double &c, double *red)
{
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5, nw = (blockDim.x + 31) >> 5;
    a = warp_sum(a); b = warp_sum(b); c = warp_sum(c);
    __syncthreads(); /* protect red[] from a previous use */
    if (lane == 0) { red[w] = a; red[16 + w] = b; red[32 + w] = c; }
    __syncthreads();
    if (w == 0) {
        double x = lane < nw ? red[lane] : 0.0, y = lane < nw ? red[16 + lane] : 0.0, z = lane < nw ? red[32 + lane] : 0.0;
        x = warp_sum(x); y = warp_sum(y); z = warp_sum(z);
        if (lane == 0) { red[48] = x; red[49] = y; red[50] = z; }
    }
    __syncthreads();
    a = red[48]; b = red[49]; c = red[50];
}

/* bytes of one stage: row (ld doubles) + order (ldo uint16) */
__host__ __device__ inline size_t big_stage_bytes(int64_t ld, int64_t ldo) { return (size_t)ld * 8 + (size_t)ldo * 2; }

/* STAGES = 2: the next pair's row and order are bulk-copied (cp.async.bulk + mbarrier) while this pair is
 * computed; STAGES = 1 when two stages do not fit in shared memory (n > ~11,000). */
template <bool TMODE, int STAGES>
__global__ void __launch_bounds__(BIG_THREADS) k_pair_stats_cta(PairParams p)
{
    extern __shared__ __align__(128) unsigned char smem[];
    const int n = p.n, m = p.m, tid = threadIdx.x, nt = blockDim.x;
    const uint32_t rowb = (uint32_t)p.ld * 8u, ordb = (uint32_t)p.ldo * 2u, stageb = rowb + ordb;
    uint8_t *s_lab = (uint8_t *)(smem + (size_t)STAGES * stageb);
    __shared__ __align__(8) uint64_t bar[2];
    __shared__ double red[64];
    __shared__ int s_dmax[BIG_THREADS / 32][3];
    __shared__ int s_wtot[BIG_THREADS / 32][3];
    __shared__ double s_med[8];
    __shared__ uint2 s_inc[4]; /* what a position of label 0 / 1 / 2 / none adds to the two packed count words of the walk */
    /* scalar results of up to TAIL_PAIRS finished pairs: their tails (Pearson p, tests, p-values, 17 result values --
     * serial work with dependent global loads) are done 32 at a time, one pair per lane of warp 0, instead of by one
     * thread per pair while the other 511 wait at the next barrier (12 % of the stall samples) */
    constexpr int TAIL_PAIRS = 32;
    __shared__ double t_d[TAIL_PAIRS][TMODE ? 11 : 5]; /* syy, sxy, medUP, medMID, medDOWN, then mg[3], ss[3] (t mode) */
    __shared__ int t_i[TAIL_PAIRS][6];                 /* pair, gene, guard bits (2 per comparison, biased by 1), d01, d02, d12 */
    int npend = 0;
    const double dm = (double)m, dn_ = (double)n;
    const int t1 = (m - 1) / 2 + 1, t2 = m / 2 + 1;
    const int lane = tid & 31, warp = tid >> 5;

    if (tid == 0) {
        mbar_init(&bar[0], 1); mbar_init(&bar[1], 1); fence_barrier_init();
        s_inc[0] = make_uint2(WALK_A0, WALK_B0); s_inc[1] = make_uint2(WALK_A1, WALK_B1);
        s_inc[2] = make_uint2(WALK_A2, WALK_B2); s_inc[3] = make_uint2(0u, 0u);
    }
    __syncthreads();
    uint32_t ph[2] = { 0, 0 };
    auto issue = [&](int j, int stage) { /* thread 0 only */
        const int pi = p.probe_idx[j];
        if (pi < 0) return;
        unsigned char *dst = smem + (size_t)stage * stageb;
        mbar_expect_tx(&bar[stage], stageb);
        bulk_g2s(dst, p.meth + (int64_t)pi * p.ld, rowb, &bar[stage]);
        bulk_g2s(dst + rowb, p.ord + (int64_t)p.pair_slot[j] * p.ldo, ordb, &bar[stage]);
    };
    auto flush_tails = [&](int count) { /* all threads call; ends on a barrier-free note: slots are rewritten by tid 0 only */
        __syncthreads();
        if (warp == 0) {
            if (lane < count) {
                const double syy = t_d[lane][0], sxy = t_d[lane][1];
                const double medUP = t_d[lane][2], medMID = t_d[lane][3], medDOWN = t_d[lane][4];
                const int j = t_i[lane][0];
                const double *ga = p.gene_aux + 4 * (int64_t)t_i[lane][1];
                const int gbits = t_i[lane][2];
                double out[PAIR_COLS], stat[3] = { EPX_NAN, EPX_NAN, EPX_NAN }, pv[3] = { EPX_NAN, EPX_NAN, EPX_NAN };
                int dnum[3] = { -1, -1, -1 };
                unsigned na = 0;
                double r = EPX_NAN, pr = EPX_NAN;
                if (ga[0] != 0.0 && syy != 0.0 && n >= 3) {
                    r = sxy * ga[3] * rsqrt(syy);
                    r = r > 1.0 ? 1.0 : (r < -1.0 ? -1.0 : r);
                    pr = epx_pearson_p_tab(r, p.pt_tab, p.pt_n, dn_ - 2.0);
                } else na |= (1u << EPX_PEARSON_R) | (1u << EPX_PEARSON_P);
                if (TMODE) {
                    const double mg[3] = { t_d[lane][TMODE ? 5 : 0], t_d[lane][TMODE ? 6 : 0], t_d[lane][TMODE ? 7 : 0] };
                    const double vg[3] = { t_d[lane][TMODE ? 8 : 0] / (dm - 1.0), t_d[lane][TMODE ? 9 : 0] / (dm - 1.0),
                                           t_d[lane][TMODE ? 10 : 0] / (dm - 1.0) };
                    const int ga_[3] = { 0, 0, 1 }, gb_[3] = { 1, 2, 2 };
                    for (int c = 0; c < 3; c++) {
                        double df = EPX_NAN;
                        bool ok = false;
                        if (((gbits >> (2 * c)) & 3) - 1 == 1)
                            ok = epx_welch(mg[ga_[c]], vg[ga_[c]], dm, mg[gb_[c]], vg[gb_[c]], dm, &stat[c], &df) == 0;
                        if (ok) pv[c] = welch_p(p, stat[c], df);
                        else { na |= 1u << (EPX_P_UP_MID + c); stat[c] = EPX_NAN; }
                    }
                } else {
                    for (int c = 0; c < 3; c++) {
                        const int ddc = t_i[lane][3 + c];
                        if (((gbits >> (2 * c)) & 3) - 1 != 0) {
                            dnum[c] = ddc * m;
                            stat[c] = (double)dnum[c] / (dm * dm);
                            pv[c] = p.lut_asym[ddc]; /* m > 341: n.x*n.y >= 10000, always asymptotic */
                        } else na |= 1u << (EPX_P_UP_MID + c);
                    }
                }
                out[0] = medDOWN; out[1] = medMID; out[2] = medUP;
                out[3] = medUP - medMID; out[4] = medUP - medDOWN; out[5] = medMID - medDOWN;
                out[6] = pv[0]; out[7] = pv[1]; out[8] = pv[2]; out[9] = r; out[10] = pr;
                for (int c = 0; c < PAIR_COLS; c++) p.pair[(int64_t)j * PAIR_COLS + c] = out[c];
                for (int c = 0; c < 3; c++) { p.pair_stat[(int64_t)j * 3 + c] = stat[c]; p.pair_dnum[(int64_t)j * 3 + c] = dnum[c]; }
                p.pair_na[j] = (uint16_t)na;
            }
            __syncwarp(); /* lane 0 refills the slots only after every lane has read its own */
        }
    };
    int cur_gene = -1;
    /* every CTA takes a CONTIGUOUS run of pairs: consecutive pairs share their gene (about 23 pairs per gene), so the
     * label bytes are reloaded once per gene instead of once per pair and the gene's centred expression stays in L1/L2
     * (a stride of gridDim.x pairs changed gene at every step: 11 % of the stall samples sat in the label reload) */
    const int np_all = p.n_pairs - p.pair_begin;
    const int per_cta = (np_all + (int)gridDim.x - 1) / (int)gridDim.x;
    const int j_begin = p.pair_begin + (int)blockIdx.x * per_cta;
    const int j_end = min(j_begin + per_cta, p.n_pairs);
    if (tid == 0 && j_begin < j_end) issue(j_begin, 0);

    int it = 0;
    for (int j = j_begin; j < j_end; j++, it++) {
        const int stage = STAGES == 2 ? (it & 1) : 0;
        const int pi = p.probe_idx[j];
        const int gene = p.pair_gene[j];
        __syncthreads(); /* everybody is done with the other stage (and with s_lab / s_med of the previous pair) */
        if (STAGES == 2 && tid == 0 && j + 1 < j_end) issue(j + 1, stage ^ 1);
        if (pi < 0) {
            if (tid < PAIR_COLS) p.pair[(int64_t)j * PAIR_COLS + tid] = EPX_NAN;
            if (tid < 3) { p.pair_stat[(int64_t)j * 3 + tid] = EPX_NAN; p.pair_dnum[(int64_t)j * 3 + tid] = -1; }
            if (tid == 0) p.pair_na[j] = (uint16_t)((1u << PAIR_COLS) - 1);
            if (STAGES == 1 && tid == 0 && j + 1 < j_end) issue(j + 1, 0);
            continue;
        }
        const double *s_row = (const double *)(smem + (size_t)stage * stageb);
        const unsigned short *s_ord = (const unsigned short *)(smem + (size_t)stage * stageb + rowb);
        const double *xc = p.xc + (int64_t)gene * p.ldx;
        const uint16_t *perm_g = p.perm + (int64_t)gene * n;
        const double *ga = p.gene_aux + 4 * (int64_t)gene;
        if (gene != cur_gene) { /* uniform: labels of the gene, reused by its following probes */
            const uint8_t *lab = p.lab8 + (int64_t)gene * p.ldl;
            for (int s = tid; s < n; s += nt) s_lab[s] = lab[s];
            cur_gene = gene;
        }
        mbar_wait(&bar[stage], ph[stage]);
        ph[stage] ^= 1;
        __syncthreads();

        const double K = s_row[0];
        double sd = 0, sdd = 0, sxd = 0, g0 = 0, g1 = 0, g2 = 0, q0 = 0, q1 = 0, q2 = 0;
        for (int s = tid; s < n; s += nt) {
            const double d = s_row[s] - K;
            sd += d; sdd = fma(d, d, sdd); sxd = fma(__ldg(xc + s), d, sxd);
            if (TMODE) {
                const unsigned l = s_lab[s];
                if (l == 0) { g0 += d; q0 = fma(d, d, q0); } else if (l == 1) { g1 += d; q1 = fma(d, d, q1); }
                else if (l == 2) { g2 += d; q2 = fma(d, d, q2); }
            }
        }
        block_sum3w(sd, sdd, sxd, red); /* three sums behind one pair of barriers */
        double syy = sdd - sd * sd / dn_, sxy = sxd - (sd / dn_) * ga[2];
        double mg0 = 0, mg1 = 0, mg2 = 0, ss0 = 0, ss1 = 0, ss2 = 0;
        bool redo = sdd > 0.0 && syy < 1e-7 * sdd;
        if (TMODE) {
            g0 = block_sum(g0, red); g1 = block_sum(g1, red); g2 = block_sum(g2, red);
            q0 = block_sum(q0, red); q1 = block_sum(q1, red); q2 = block_sum(q2, red);
            ss0 = q0 - g0 * g0 / dm; ss1 = q1 - g1 * g1 / dm; ss2 = q2 - g2 * g2 / dm;
            mg0 = K + g0 / dm; mg1 = K + g1 / dm; mg2 = K + g2 / dm;
            redo = redo || (q0 > 0.0 && ss0 < 1e-7 * q0) || (q1 > 0.0 && ss1 < 1e-7 * q1) || (q2 > 0.0 && ss2 < 1e-7 * q2);
        }
        if (sdd == 0.0) syy = 0.0;
        if (redo) {
            const double mean_all = K + sd / dn_;
            double a0 = 0, a1 = 0, b0 = 0, b1 = 0, b2 = 0;
            for (int s = tid; s < n; s += nt) {
                const double yv = s_row[s], d = yv - mean_all;
                a0 = fma(d, d, a0); a1 = fma(__ldg(xc + s), d, a1);
                if (TMODE) {
                    const unsigned l = s_lab[s];
                    const double dg = yv - (l == 0 ? mg0 : (l == 1 ? mg1 : mg2));
                    if (l == 0) b0 = fma(dg, dg, b0); else if (l == 1) b1 = fma(dg, dg, b1); else if (l == 2) b2 = fma(dg, dg, b2);
                }
            }
            syy = block_sum(a0, red); sxy = block_sum(a1, red);
            if (TMODE) { ss0 = block_sum(b0, red); ss1 = block_sum(b1, red); ss2 = block_sum(b2, red); }
        }

        /* sorted walk, ONE pass: thread t owns positions [t*chunk, (t+1)*chunk). As in k_pair_stats2 the label counts run as
         * two words of two biased 16-bit fields, ra = [cUP-cMID | cUP-cDOWN], rb = [cMID-cDOWN | cUP], advanced by the
         * increment pair of each position's label (a 4-entry table) and folded into running packed maxima / minima at
         * tie-run ends; the counts in front of the thread come from a scan afterwards (max(prefix + x) = prefix + max x),
         * and only the threads whose positions hold a group's middle members look at their positions a second time.
         * (Two passes with absolute counts -- count, scan, walk -- cost 66 instructions per position; this costs ~15.) */
        int chunk = (n + nt - 1) / nt;
        chunk += (chunk & 7) == 0 ? 1 : 0; /* lanes start `chunk` 16-bit entries apart: a multiple of 8 would share few banks */
        const int k0 = min(tid * chunk, n), k1 = min(k0 + chunk, n);
        unsigned ra = WALK_BIAS, rb = WALK_BIAS, mxa = 0u, mna = 0xffffffffu, mxb = 0u, mnb = 0xffffffffu;
        for (int k = k0; k < k1; k++) {
            const unsigned short o = s_ord[k];
            const uint2 inc = s_inc[s_lab[o & 0x7fffu]];
            ra += inc.x; rb += inc.y;
            if (!TMODE && !(o & 0x8000u)) { /* inside a tie run nothing is evaluated */
                mxa = __vmaxu2(mxa, ra); mna = __vminu2(mna, ra);
                mxb = __vmaxu2(mxb, rb); mnb = __vminu2(mnb, rb);
            }
        }
        const unsigned da = ra - WALK_BIAS, db = rb - WALK_BIAS; /* this thread's totals (wrap-around sums of the fields) */
        unsigned ia = da, ib = db;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            const unsigned ta = __shfl_up_sync(FULL, ia, d), tb = __shfl_up_sync(FULL, ib, d);
            if (lane >= d) { ia += ta; ib += tb; }
        }
        __syncthreads(); /* s_wtot of the previous pair has been consumed */
        if (lane == 31) { s_wtot[warp][0] = (int)ia; s_wtot[warp][1] = (int)ib; }
        __syncthreads();
        unsigned wa = 0, wb = 0;
        for (int w = 0; w < warp; w++) { wa += (unsigned)s_wtot[w][0]; wb += (unsigned)s_wtot[w][1]; }
        ia += wa; ib += wb;
        const unsigned ea = ia - da, eb = ib - db; /* counts in front of this thread */
        int d01 = 0, d02 = 0, d12 = 0;
        if (!TMODE && mxa != 0u) { /* the thread evaluated at least one position */
            const unsigned xa = mxa + ea, na_ = mna + ea, xb = mxb + eb, nb = mnb + eb;
            d01 = max((int)(xa >> 16) - 0x4000, 0x4000 - (int)(na_ >> 16));
            d02 = max((int)(xa & 0xffffu) - 0x4000, 0x4000 - (int)(na_ & 0xffffu));
            d12 = max((int)(xb >> 16) - 0x4000, 0x4000 - (int)(nb >> 16));
        }
        if (m >= 1 && k0 < k1) { /* medians: the t1-th and t2-th member of every group */
            const unsigned eab = ea + WALK_BIAS, iab = ia + WALK_BIAS;
            const int ce[3] = { (int)(eb & 0xffffu), (int)(eb & 0xffffu) - ((int)(eab >> 16) - 0x4000),
                                (int)(eb & 0xffffu) - ((int)(eab & 0xffffu) - 0x4000) };
            const int ci[3] = { (int)(ib & 0xffffu), (int)(ib & 0xffffu) - ((int)(iab >> 16) - 0x4000),
                                (int)(ib & 0xffffu) - ((int)(iab & 0xffffu) - 0x4000) };
#pragma unroll
            for (int g = 0; g < 3; g++) {
                if (ce[g] < t2 && t1 <= ci[g]) { /* at least one of the two lies in this thread's positions (rare) */
                    int c = ce[g];
                    for (int k = k0; k < k1; k++) {
                        const int s = s_ord[k] & 0x7fffu;
                        if (s_lab[s] == (unsigned)g) {
                            ++c;
                            if (c == t1) s_med[g * 2] = s_row[s];
                            if (c == t2) s_med[g * 2 + 1] = s_row[s];
                        }
                    }
                }
            }
        }
        if (!TMODE) { /* per-warp maxima; thread 0 folds them when it files the pair's scalars (nobody else needs them) */
            d01 = warp_max_i(d01); d02 = warp_max_i(d02); d12 = warp_max_i(d12);
            if (lane == 0) { s_dmax[warp][0] = d01; s_dmax[warp][1] = d02; s_dmax[warp][2] = d12; }
        }
        __syncthreads();
        const double medUP = (s_med[0] + s_med[1]) * 0.5, medMID = (s_med[2] + s_med[3]) * 0.5,
                     medDOWN = (s_med[4] + s_med[5]) * 0.5;

        /* guard (R/Functions.R:193) */
        int guard[3] = { -1, -1, -1 };
        if (m >= 2) {
            const int offA[3] = { 0, 0, m }, offB[3] = { m, 2 * m, 2 * m };
            for (int c = 0; c < 3; c++) {
                const double dd = s_row[perm_g[offA[c]]] - s_row[perm_g[offB[c]]];
                int differs = 0;
                if (s_row[perm_g[offA[c] + 1]] - s_row[perm_g[offB[c] + 1]] != dd) differs = 1;
                else
                    for (int i = 2 + tid; i < m; i += nt)
                        if (s_row[perm_g[offA[c] + i]] - s_row[perm_g[offB[c] + i]] != dd) differs = 1;
                guard[c] = __syncthreads_or(differs) ? 1 : 0;
            }
        } else __syncthreads();
        /* single stage: nobody reads the row or the order any more (the barrier above), so the next pair's copy
         * starts now and runs under the scalar tail below */
        if (STAGES == 1 && tid == 0 && j + 1 < j_end) issue(j + 1, 0);

        if (tid == 0) {
            t_d[npend][0] = syy; t_d[npend][1] = sxy;
            t_d[npend][2] = medUP; t_d[npend][3] = medMID; t_d[npend][4] = medDOWN;
            if (TMODE) {
                t_d[npend][TMODE ? 5 : 0] = mg0; t_d[npend][TMODE ? 6 : 0] = mg1; t_d[npend][TMODE ? 7 : 0] = mg2;
                t_d[npend][TMODE ? 8 : 0] = ss0; t_d[npend][TMODE ? 9 : 0] = ss1; t_d[npend][TMODE ? 10 : 0] = ss2;
            }
            t_i[npend][0] = j; t_i[npend][1] = gene;
            t_i[npend][2] = (guard[0] + 1) | ((guard[1] + 1) << 2) | ((guard[2] + 1) << 4);
            if (!TMODE)
                for (int w = 1; w < (nt >> 5); w++) { d01 = max(d01, s_dmax[w][0]); d02 = max(d02, s_dmax[w][1]); d12 = max(d12, s_dmax[w][2]); }
            t_i[npend][3] = d01; t_i[npend][4] = d02; t_i[npend][5] = d12;
        }
        if (++npend == TAIL_PAIRS) { flush_tails(npend); npend = 0; }
    }
    if (npend > 0) flush_tails(npend);
}

/* ============================================================ transpose (sample-major upload) */

/* Every upload path also looks at the values it moves: a NaN (the reference drops such probes / samples with na.omit,
 * R/EpimethexAnalysis.R:114,278) or an infinity has no place in the sort keys (key_of) or the moments, so the C ABI
 * refuses the matrix instead of returning undefined medians. first_bad: smallest (row * n + sample) that is not finite. */
__device__ __forceinline__ void note_nonfinite(double v, unsigned long long where, unsigned long long *first_bad)
{
    if ((__double2hiint(v) & 0x7ff00000) == 0x7ff00000) atomicMin(first_bad, where);
}

__global__ void __launch_bounds__(256) k_transpose_cols(const double *__restrict__ src, double *__restrict__ dst,
                                                        int64_t P, int ns, int64_t s0, int64_t ld, int64_t n_all,
                                                        unsigned long long *first_bad)
{
    __shared__ double tile[32][33];
    const int64_t p0 = (int64_t)blockIdx.x * 32;
    const int j0 = blockIdx.y * 32;
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5; /* 32 x 8 */
    for (int r = ty; r < 32; r += 8) {
        int j = j0 + r;
        int64_t pp = p0 + tx;
        if (j < ns && pp < P) {
            const double v = src[(int64_t)j * P + pp];
            tile[r][tx] = v;
            note_nonfinite(v, (unsigned long long)(pp * n_all + s0 + j), first_bad);
        }
    }
    __syncthreads();
    for (int r = ty; r < 32; r += 8) {
        int64_t pp = p0 + r;
        int j = j0 + tx;
        if (j < ns && pp < P) dst[pp * ld + s0 + j] = tile[tx][r];
    }
}

__global__ void __launch_bounds__(256) k_respace_rows(const double *__restrict__ src, double *__restrict__ dst, int64_t rows,
                                                      int n, int64_t ld, int64_t row0, unsigned long long *first_bad)
{
    const int64_t total = rows * n;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
        const int64_t r = i / n;
        const double v = src[i];
        dst[r * ld + (i - r * n)] = v;
        note_nonfinite(v, (unsigned long long)(row0 * n + i), first_bad);
    }
}

/* the same test for matrices that arrive without a re-spacing pass (padded host rows, device pointers, expression) */
__global__ void __launch_bounds__(256) k_check_finite(const double *__restrict__ a, int64_t rows, int n, int64_t ld,
                                                      unsigned long long *first_bad)
{
    const int64_t total = rows * n;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
        const int64_t r = i / n;
        note_nonfinite(a[r * ld + (i - r * n)], (unsigned long long)i, first_bad);
    }
}

/* Rows of a page-locked HOST matrix fetched over PCIe by the kernel itself -- only the rows the index names
 * (uniq[0..U), ascending) -- into their places in the padded device matrix: the upload, the subset() of
 * R/EpimethexAnalysis.R:281-282 (methylation rows of the annotated probes of the gene range) and the re-spacing in one
 * pass. A warp takes a row; every load instruction of the warp covers one 256-byte block of host memory that starts
 * on a 128-byte boundary (whole read requests on the bus but for the two ends of the row: rows of 473 doubles start
 * at any multiple of 8 bytes), eight loads in flight per lane. Lanes in front of the row or behind it load nothing. */
__global__ void __launch_bounds__(256) k_gather_host_rows(const double *__restrict__ src, int64_t lds,
                                                          const int32_t *__restrict__ uniq, int64_t U, int n,
                                                          double *__restrict__ dst, int64_t ld,
                                                          unsigned long long *first_bad)
{
    const int lane = threadIdx.x & 31;
    const int64_t nw = ((int64_t)gridDim.x * blockDim.x) >> 5;
    for (int64_t u = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5; u < U; u += nw) {
        const int64_t p = uniq[u];
        const double *s = src + p * lds;
        double *d = dst + p * ld;
        const int lead = (int)(((uintptr_t)s & 127) >> 3); /* doubles between the 128-byte boundary below the row and the row */
        for (int j0 = -lead; j0 < n; j0 += 256) {
            double v[8];
#pragma unroll
            for (int k = 0; k < 8; k++) {
                const int j = j0 + k * 32 + lane;
                v[k] = (j >= 0 && j < n) ? __ldcs(s + j) : 0.0;
            }
#pragma unroll
            for (int k = 0; k < 8; k++) {
                const int j = j0 + k * 32 + lane;
                if (j >= 0 && j < n) {
                    d[j] = v[k];
                    note_nonfinite(v[k], (unsigned long long)(p * n + j), first_bad);
                }
            }
        }
    }
}

/* ============================================================ launchers */

cudaError_t launch_gather_host_rows(const double *src, int64_t lds, const int32_t *uniq, int64_t U, int64_t n, double *dst,
                                    int64_t ld, unsigned long long *first_bad, cudaStream_t st)
{
    if (U <= 0) return cudaSuccess;
    k_gather_host_rows<<<148 * 8, 256, 0, st>>>(src, lds, uniq, U, (int)n, dst, ld, first_bad);
    return cudaGetLastError();
}

cudaError_t launch_respace_rows(const double *src, double *dst, int64_t rows, int64_t n, int64_t ld, int64_t row0,
                                unsigned long long *first_bad, cudaStream_t st)
{
    if (rows <= 0) return cudaSuccess;
    k_respace_rows<<<148 * 8, 256, 0, st>>>(src, dst, rows, (int)n, ld, row0, first_bad);
    return cudaGetLastError();
}

cudaError_t launch_check_finite(const double *a, int64_t rows, int64_t n, int64_t ld, unsigned long long *first_bad,
                                cudaStream_t st)
{
    if (rows <= 0) return cudaSuccess;
    k_check_finite<<<148 * 8, 256, 0, st>>>(a, rows, (int)n, ld, first_bad);
    return cudaGetLastError();
}



template <typename F> static cudaError_t allow_max_dynamic_smem(F *func)
{
    /* opt in to the full per-CTA shared memory: device limit minus the kernel's static shared memory */
    int dev = 0, optin = 0;
    cudaError_t e = cudaGetDevice(&dev);
    if (e != cudaSuccess) return e;
    e = cudaDeviceGetAttribute(&optin, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev);
    if (e != cudaSuccess) return e;
    cudaFuncAttributes fa;
    e = cudaFuncGetAttributes(&fa, (const void *)func);
    if (e != cudaSuccess) return e;
    return cudaFuncSetAttribute((const void *)func, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                optin - (int)fa.sharedSizeBytes);
}

#define EPX_FOR_PAIR_VARIANTS(X) X(1, 1) X(2, 2) X(4, 4) X(8, 8) X(15, 16) X(16, 16) X(24, 32) X(32, 32)
#define EPX_FOR_PAIR2_VARIANTS(X) X(1) X(2) X(3) X(4) X(5) X(6) X(7) X(8) X(9) X(10) X(11) X(12) X(13) X(14) X(15) X(16) X(32)
#define EPX_FOR_SORT_VARIANTS(X) X(1) X(2) X(4) X(8) X(16) X(32)

cudaError_t kernels_init()
{
    cudaError_t e;
    if ((e = allow_max_dynamic_smem(k_gene_sort)) != cudaSuccess) return e;
    if ((e = allow_max_dynamic_smem(k_probe_rank)) != cudaSuccess) return e;
    if ((e = allow_max_dynamic_smem(k_gene_tests)) != cudaSuccess) return e;
#define EPX_PAIR_ATTR(L, W)                                                                  \
    if ((e = allow_max_dynamic_smem(k_pair_stats<L, W, true>)) != cudaSuccess) return e;     \
    if ((e = allow_max_dynamic_smem(k_pair_stats<L, W, false>)) != cudaSuccess) return e;
    EPX_FOR_PAIR_VARIANTS(EPX_PAIR_ATTR)
#undef EPX_PAIR_ATTR
    if ((e = allow_max_dynamic_smem(k_pair_stats_cta<true, 1>)) != cudaSuccess) return e;
    if ((e = allow_max_dynamic_smem(k_pair_stats_cta<false, 1>)) != cudaSuccess) return e;
    if ((e = allow_max_dynamic_smem(k_pair_stats_cta<true, 2>)) != cudaSuccess) return e;
    if ((e = allow_max_dynamic_smem(k_pair_stats_cta<false, 2>)) != cudaSuccess) return e;
#define EPX_PAIR2_ATTR(E)                                                                 \
    if ((e = allow_max_dynamic_smem(k_pair_stats2<E, true>)) != cudaSuccess) return e;    \
    if ((e = allow_max_dynamic_smem(k_pair_stats2<E, false>)) != cudaSuccess) return e;
    EPX_FOR_PAIR2_VARIANTS(EPX_PAIR2_ATTR)
#undef EPX_PAIR2_ATTR
    if ((e = allow_max_dynamic_smem(k_group_blocksort)) != cudaSuccess) return e;
    if ((e = allow_max_dynamic_smem(k_probe_rank_w2)) != cudaSuccess) return e;
#define EPX_RANK_ATTR(E)                                                                      \
    if ((e = allow_max_dynamic_smem(k_probe_rank_w<E, false>)) != cudaSuccess) return e;      \
    if ((e = allow_max_dynamic_smem(k_probe_rank_w<E, true>)) != cudaSuccess) return e;
    EPX_FOR_SORT_VARIANTS(EPX_RANK_ATTR)
#undef EPX_RANK_ATTR
    return cudaSuccess;
}

/* ---- pooled probe groups ---- */
int group_block_cap(int n) { return n <= 8192 ? 8192 : ctasort_entries(n); }

cudaError_t launch_group_moments(const GroupParams &p, cudaStream_t st)
{
    if (p.n_multi <= 0) return cudaSuccess;
    k_group_moments<<<p.n_multi, GROUP_THREADS, 0, st>>>(p);
    return cudaGetLastError();
}
cudaError_t launch_group_blocksort(const GroupParams &p, cudaStream_t st)
{
    if (p.n_blocks <= 0) return cudaSuccess;
    k_group_blocksort<<<p.n_blocks, BLOCKSORT_THREADS, group_sort_smem(p.cap), st>>>(p);
    return cudaGetLastError();
}
cudaError_t launch_group_merge(const GroupParams &p, int pass, int64_t run_len, int n_tiles, cudaStream_t st)
{
    if (n_tiles <= 0) return cudaSuccess;
    k_group_merge<<<n_tiles, GROUP_THREADS, 0, st>>>(p, pass, run_len);
    return cudaGetLastError();
}
cudaError_t launch_group_walk(const GroupParams &p, cudaStream_t st)
{
    if (p.n_multi <= 0) return cudaSuccess;
    k_group_walk<<<p.n_multi, GROUP_THREADS, 0, st>>>(p);
    return cudaGetLastError();
}

static int sort_eps(int n)
{
    int eps = 1;
    while (32 * eps < n) eps <<= 1;
    return eps;
}

cudaError_t launch_gene_sort(const GeneParams &p, int sm_count, cudaStream_t st)
{
    if (p.G <= 0) return cudaSuccess;
    if (p.n > MAX_WARP_SAMPLES) {
        k_gene_sort<<<p.G, ctab_threads(p.n), ctab_row_bytes(p.n) + ctab_smem(p.n), st>>>(p);
        return cudaGetLastError();
    }
    int grid = (p.G + SORT_WARPS - 1) / SORT_WARPS; /* one gene per warp: G is small, favour parallelism */
    (void)sm_count;
    switch (sort_eps(p.n)) {
#define EPX_CASE(E) case E: k_gene_sort_w<E><<<grid, SORT_WARPS * 32, 0, st>>>(p); break;
        EPX_FOR_SORT_VARIANTS(EPX_CASE)
#undef EPX_CASE
    default: return cudaErrorInvalidValue;
    }
    return cudaGetLastError();
}

cudaError_t launch_pick_reference(const GeneTestParams &p, cudaStream_t st)
{
    k_pick_reference<<<1, 1024, 0, st>>>(p);
    return cudaGetLastError();
}

cudaError_t launch_gene_tests(const GeneTestParams &p, int sm_count, cudaStream_t st)
{
    if (p.G <= 0) return cudaSuccess;
    if (p.n > MAX_WARP_SAMPLES) {
        k_gene_tests<<<p.G, 128, (size_t)(p.n + 2) * sizeof(double), st>>>(p);
        return cudaGetLastError();
    }
    int grid = (p.G + GT_WARPS - 1) / GT_WARPS; /* one gene per warp */
    (void)sm_count;
    const int eps = sort_eps(p.n);
    const size_t smem = (size_t)GT_WARPS * (32 * eps + 2) * sizeof(double);
    switch (eps) {
#define EPX_CASE(E) case E: k_gene_tests_w<E><<<grid, GT_WARPS * 32, smem, st>>>(p); break;
        EPX_FOR_SORT_VARIANTS(EPX_CASE)
#undef EPX_CASE
    default: return cudaErrorInvalidValue;
    }
    if (p.test_t) k_gene_tails<<<(3 * p.G + 127) / 128, 128, 0, st>>>(p);
    return cudaGetLastError();
}

bool probe_rank_takes_genes(int n) { return n <= 512; }

template <int E, bool GENES> static cudaError_t launch_rows_w(const RankParams &p, int sm_count, cudaStream_t st)
{
    constexpr int RW = rank_warps<E>();
    const size_t smem = rank_warp_smem<E>() * RW;
    /* persistent grid of exactly one wave: a grid of 8 CTAs per SM with 3 resident ran 2.67 waves */
    static std::atomic<int> per_sm{0};
    if (per_sm == 0) {
        int nb = 0;
        if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, k_probe_rank_w<E, GENES>, RW * 32, smem) != cudaSuccess || nb < 1) nb = 1;
        per_sm = nb;
    }
    const int64_t rows = GENES ? p.genes.G : p.U;
    if (rows <= 0) return cudaSuccess;
    int64_t grid = (rows + RW - 1) / RW;
    if (grid > (int64_t)sm_count * per_sm) grid = (int64_t)sm_count * per_sm;
    k_probe_rank_w<E, GENES><<<(unsigned)grid, RW * 32, smem, st>>>(p);
    return cudaGetLastError();
}

/* the expression rows through the row kernel (n <= 512): order, labels, centred rows, reference pick; replaces
 * k_gene_sort_w + k_pick_reference for those sample counts */
cudaError_t launch_gene_rows(const RankParams &p, int sm_count, cudaStream_t st)
{
    switch (sort_eps(p.n)) {
#define EPX_CASE(E) case E: return launch_rows_w<E, true>(p, sm_count, st);
        EPX_FOR_SORT_VARIANTS(EPX_CASE)
#undef EPX_CASE
    default: return cudaErrorInvalidValue;
    }
}

cudaError_t launch_probe_rank(const RankParams &p, int sm_count, cudaStream_t st)
{
    if (p.U <= 0) return cudaSuccess;
    if (p.n > MAX_WARP_SAMPLES) {
        {
            const int threads = ctab_threads(p.n);
            const size_t smem = ctab_row_bytes(p.n) + ctab_smem(p.n);
            int per_sm = (int)((226 * 1024) / (smem + 1024));
            const int by_regs = 65536 / (80 * ((threads + 127) / 128 * 128)), by_threads = 2048 / threads; /* registers come in units of four warps */
            if (per_sm > by_regs) per_sm = by_regs;
            if (per_sm > by_threads) per_sm = by_threads;
            if (per_sm < 1) per_sm = 1;
            int64_t grid = (int64_t)sm_count * per_sm;
            if (grid > p.U) grid = p.U;
            k_probe_rank<<<(unsigned)grid, threads, smem, st>>>(p);
        }
        return cudaGetLastError();
    }
    if (p.n > 512 && !rank_by_buckets<32>() && (p.ldo & 7) == 0 && ((uintptr_t)p.ord & 15) == 0) { /* two 512-value halves + one merge per warp */
        static std::atomic<int> per_sm2{0}; /* written by one host thread per device (epx_multi_*): the value is the same */
        const size_t smem = (size_t)RANK2_WARPS * RANK2_WARP_SMEM;
        if (per_sm2 == 0) {
            int nb = 0;
            if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, k_probe_rank_w2, RANK2_WARPS * 32, smem) != cudaSuccess || nb < 1) nb = 1;
            per_sm2 = nb;
        }
        int64_t grid = (p.U + RANK2_WARPS - 1) / RANK2_WARPS;
        if (grid > (int64_t)sm_count * per_sm2) grid = (int64_t)sm_count * per_sm2;
        k_probe_rank_w2<<<(unsigned)grid, RANK2_WARPS * 32, smem, st>>>(p);
        return cudaGetLastError();
    }
    switch (sort_eps(p.n)) {
#define EPX_CASE(E) case E: return launch_rows_w<E, false>(p, sm_count, st);
        EPX_FOR_SORT_VARIANTS(EPX_CASE)
#undef EPX_CASE
    default: return cudaErrorInvalidValue;
    }
}

template <int EPL, int EPW>
static cudaError_t launch_pair_epl(const PairParams &p, int sm_count, cudaStream_t st)
{
    const size_t smem = pair_warp_smem(p.n) * PAIR_WARPS;
    int per_sm = (int)((227 * 1024) / (smem + 1024));
    if (per_sm < 1) per_sm = 1;
    if (per_sm > 8) per_sm = 8;
    int grid = sm_count * per_sm;
    const int need = (p.n_items + PAIR_WARPS - 1) / PAIR_WARPS;
    if (grid > need) grid = need;
    if (grid < 1) grid = 1;
    if (p.test_t) k_pair_stats<EPL, EPW, true><<<grid, PAIR_WARPS * 32, smem, st>>>(p);
    else k_pair_stats<EPL, EPW, false><<<grid, PAIR_WARPS * 32, smem, st>>>(p);
    return cudaGetLastError();
}

template <int EPL>
static cudaError_t launch_pair2(PairParams p, int sm_count, cudaStream_t st)
{
    p.warp_smem = (unsigned)pair2_warp_smem(p.n, p.ld, p.ldo, p.test_t != 0);
    const size_t smem = (size_t)p.warp_smem * PAIR2_WARPS;
    int per_sm = (int)((227 * 1024) / (smem + 1024));
    if (per_sm < 1) per_sm = 1;
    if (per_sm > 16) per_sm = 16;
    int grid = sm_count * per_sm;
    const int need = (p.n_items + PAIR2_WARPS - 1) / PAIR2_WARPS;
    if (grid > need) grid = need;
    if (grid < 1) grid = 1;
    if (p.test_t) k_pair_stats2<EPL, true><<<grid, PAIR2_WARPS * 32, smem, st>>>(p);
    else k_pair_stats2<EPL, false><<<grid, PAIR2_WARPS * 32, smem, st>>>(p);
    return cudaGetLastError();
}

static bool pair2_eligible(const PairParams &p)
{
    if (p.n > MAX_WARP_SAMPLES) return false;
    const int epl = (p.n + 31) / 32;
    if (!(epl <= 16 || epl == 32)) return false;
    /* bulk-copy path: rows and order rows must be 16-byte aligned with a 16-byte multiple length */
    const bool aligned = (((uintptr_t)p.meth) & 15) == 0 && (p.ld & 1) == 0 && (((uintptr_t)p.ord) & 15) == 0 && (p.ldo & 7) == 0;
    return aligned && pair2_warp_smem(p.n, p.ld, p.ldo, p.test_t != 0) * PAIR2_WARPS <= 200 * 1024;
}

cudaError_t launch_pair_stats(const PairParams &p, int sm_count, cudaStream_t st)
{
    if (p.n_items <= 0) return cudaSuccess;
    if (p.n > MAX_WARP_SAMPLES) {
        const size_t stage = big_stage_bytes(p.ld, p.ldo), labb = (size_t)((p.n + 15) & ~15);
        /* two single-stage CTAs per SM when they fit (while one sits in its barriers and scalar tail the other
         * computes), else one CTA with two stages, else one stage */
        const bool pair_up = 2 * (stage + labb + 2048) <= 227 * 1024;
        const bool two = !pair_up && 2 * stage + labb + 4096 <= 227 * 1024;
        const size_t smem = (two ? 2 : 1) * stage + labb;
        int grid = sm_count * (pair_up ? 2 : 1); /* 512-thread CTAs, persistent */
        if (grid > p.n_pairs - p.pair_begin) grid = p.n_pairs - p.pair_begin;
        if (grid < 1) return cudaSuccess;
        if (two) {
            if (p.test_t) k_pair_stats_cta<true, 2><<<grid, BIG_THREADS, smem, st>>>(p);
            else k_pair_stats_cta<false, 2><<<grid, BIG_THREADS, smem, st>>>(p);
        } else {
            if (p.test_t) k_pair_stats_cta<true, 1><<<grid, BIG_THREADS, smem, st>>>(p);
            else k_pair_stats_cta<false, 1><<<grid, BIG_THREADS, smem, st>>>(p);
        }
        return cudaGetLastError();
    }
    const int epl = (p.n + 31) / 32;
    if (pair2_eligible(p)) {
        switch (epl) {
#define EPX_CASE2(E) case E: return launch_pair2<E>(p, sm_count, st);
            EPX_FOR_PAIR2_VARIANTS(EPX_CASE2)
#undef EPX_CASE2
        default: break;
        }
    }
#define EPX_CASE(L, W) if (epl <= L) return launch_pair_epl<L, W>(p, sm_count, st);
    EPX_FOR_PAIR_VARIANTS(EPX_CASE)
#undef EPX_CASE
    return cudaErrorInvalidValue;
}

cudaError_t launch_transpose_cols(const double *src, double *dst, int64_t P, int ns, int64_t s0, int64_t ld,
                                  int64_t n_all, unsigned long long *first_bad, cudaStream_t st)
{
    dim3 grid((unsigned)((P + 31) / 32), (unsigned)((ns + 31) / 32));
    k_transpose_cols<<<grid, 256, 0, st>>>(src, dst, P, ns, s0, ld, n_all, first_bad);
    return cudaGetLastError();
}

} // namespace epx
