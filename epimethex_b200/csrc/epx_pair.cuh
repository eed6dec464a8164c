/*
 * epx_pair.cuh — k_pair_stats2: the gather/statistics kernel of the path, one warp per (gene, probe) pair.
 *
 * Reference code replaced: the per-pair gather dfMethylation[cg, samples-in-the-gene's-order]
 * (R/EpimethexAnalysis.R:299-302), Analysis (R/Functions.R:85-132: three group medians, beta differences,
 * three two-sample tests with the calculateTtest guard, R/Functions.R:192-220) and psych::corr.test
 * (R/EpimethexAnalysis.R:404-417).
 *
 * Data movement: each warp streams its pairs through a two-stage ring in shared memory.  One elected lane
 * issues cp.async.bulk (SASS UBLKCP) for the probe's methylation row (8n bytes) and its ascending order
 * (2n bytes, from k_probe_rank) of the NEXT pair while all lanes compute the current one; completion is
 * tracked by an mbarrier per stage (expect_tx / try_wait.parity).  The gene's centred expression vector
 * lives in registers and its labels in a small shared table, both reused by every probe of the gene.
 *
 * Arithmetic per pair: one pass over the row accumulates pivot-shifted moments (sum d, sum d^2, sum xc*d with
 * d = y - y[0]) for Pearson r (+ per-group sums for Welch); one walk over the probe's sorted order with
 * running label counts yields the three medians and the three KS numerators at tie-run ends.  Student-t
 * tails are deferred to the end of the work item so that 32 of them run on 32 lanes at once.
 */
#pragma once
#include "epx_device.cuh"
#include "epx_internal.h"

namespace epx {

/* 8-byte load from a shared-window address */
__device__ __forceinline__ uint2 lds_u2(uint32_t addr)
{
    uint2 v;
    asm volatile("ld.shared.v2.u32 {%0, %1}, [%2];" : "=r"(v.x), "=r"(v.y) : "r"(addr));
    return v;
}

/* out of line on purpose: the continued fraction + lgamma code is big, and inlining it into the pair loop pushes
 * the kernel past the instruction cache */
__device__ __noinline__ double t_two_sided_p_call(double t, double df) { return epx_t_two_sided_p(t, df); }

/* Welch p-value of the t mode: the (df, r) table when the session built one (m >= 30), else the continued fraction */
__device__ __forceinline__ double welch_p(const PairParams &p, double t, double df)
{
    if (p.wt_tab) return epx_welch_p_tab(t, df, p.wt_tab, p.wt_N, p.wt_ld, p.wt_K, p.wt_df0, p.wt_inv_ddf);
    return t_two_sided_p_call(t, df);
}

constexpr int PAIR2_WARPS = 4;


/* sorted walk: two running words of two biased 16-bit fields each, ra = [cUP-cMID | cUP-cDOWN],
 * rb = [cMID-cDOWN | cUP]; the table holds, per sample, what its label adds to them (32-bit wrap-around sums:
 * a field value stays within bias +- 1024, so no carry ever crosses a field in the decoded value) */
constexpr unsigned WALK_BIAS = 0x40004000u;
constexpr unsigned WALK_A0 = 0x00010001u, WALK_B0 = 0x00000001u; /* UP:     +1 +1 |  0 +1 */
constexpr unsigned WALK_A1 = 0xffff0000u, WALK_B1 = 0x00010000u; /* Medium: -1  0 | +1  0 */
constexpr unsigned WALK_A2 = 0xffffffffu, WALK_B2 = 0xffff0000u; /* Down:    0 -1 | -1  0 */

__host__ __device__ inline size_t round16(size_t v) { return (v + 15) & ~(size_t)15; }

/* per-warp shared memory: 2 mbarriers | 2 x (row, order) | label table | median owners [32][6] u16 (+ pad).
 * The t mode needs no more: the Welch t and df of a pair wait for the item's second stage in the pair's own result
 * rows (pair_stat and the p-value columns of `pair`), so both modes keep four CTAs = 16 warps per SM. */
__host__ __device__ inline size_t pair2_warp_smem(int n, int64_t ld, int64_t ldo, bool tmode)
{
    (void)tmode;
    return round16(16 + 2 * ((size_t)ld * 8 + (size_t)ldo * 2) + round16((size_t)(n + 1) * 8) + ITEM_PAIRS * 6 * 2 + 128);
}

template <int EPL, bool TMODE>
__global__ void __launch_bounds__(PAIR2_WARPS * 32, (EPL <= 16 ? 4 : 2)) k_pair_stats2(PairParams p)
{
    constexpr int EPW = EPL <= 1 ? 1 : EPL <= 2 ? 2 : EPL <= 4 ? 4 : EPL <= 8 ? 8 : EPL <= 16 ? 16 : 32;
    static_assert(EPL <= 32, "warp-per-pair kernel handles n <= 1024");
    extern __shared__ __align__(128) unsigned char smem[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int n = p.n, m = p.m;
    const uint32_t rowb = (uint32_t)p.ld * 8u, ordb = (uint32_t)p.ldo * 2u;
    unsigned char *base = smem + (size_t)warp * p.warp_smem;
    uint64_t *bar = (uint64_t *)base;
    const uint32_t stageb = rowb + ordb;
    unsigned char *buf0 = base + 16; /* stage s lives at buf0 + s*stageb */
    uint2 *tab = (uint2 *)(base + 16 + 2 * (size_t)stageb);
    const uint32_t tab_sa = smem_u32(tab);
    /* per pair of the item: (lane << 6 | rank inside the lane) of the lower/upper middle element of each group */
    unsigned short *s_own = (unsigned short *)((unsigned char *)tab + round16((size_t)(n + 1) * 8));

    if (lane == 0) {
        mbar_init(&bar[0], 1);
        mbar_init(&bar[1], 1);
        fence_barrier_init();
    }
    __syncwarp();

    const int gw = blockIdx.x * PAIR2_WARPS + warp, nw = gridDim.x * PAIR2_WARPS;
    const double dm = (double)m, dn_ = (double)n, inv_n = 1.0 / dn_, inv_m = 1.0 / dm;
    const int t1 = (m - 1) / 2 + 1, t2 = m / 2 + 1;
    const bool m_even = (m & 1) == 0;
    const bool has_unl = 3 * m != n; /* n mod 3 != 0: the lowest-expression samples carry no label */

    int fill = 0, use = 0;
    uint32_t ph0 = 0, ph1 = 0;
    auto issue = [&](int pi, int slot) {
        if (lane == 0) {
            uint64_t *b = &bar[fill];
            unsigned char *dst = buf0 + (fill ? stageb : 0u);
            mbar_expect_tx(b, stageb);
            bulk_g2s(dst, p.meth + (int64_t)pi * p.ld, rowb, b);
            bulk_g2s(dst + rowb, p.ord + (int64_t)slot * p.ldo, ordb, b);
        }
        fill ^= 1;
    };

    /* gene state */
    int cur_gene = -1;
    double xc_r[EPL];
    unsigned labs = 0; /* t-mode: 2-bit label of this lane's element e (EPL <= 16) */
    unsigned long long labs64 = 0;
    double rsx = 0.0, sxc = 0.0;
    bool x_const = true;
    int pm[6] = { 0, 0, 0, 0, 0, 0 };
    const uint16_t *perm_g = nullptr;

    /* sums of pair jj of the item, kept by lane jj for the epilogue */
    double st_syy = 0.0, st_sxy = 0.0;
    unsigned st_ks = 0; /* KS numerators / m: 3 x 9 bits, then 3 bits "the pooled pair of groups holds a tie" */

    /* A gene change is split in two: the centred expression vector (registers) and the label bytes are requested
     * before the epilogue of the previous item, which needs neither; the label table is rebuilt after it. */
    constexpr int NLQ = (EPL * 32 + 127) / 128; /* 4 labels per 32-bit load */
    unsigned lraw[NLQ];
    int pend_gene = -1;
    auto load_gene = [&](int g) {
        const double *xc = p.xc + (int64_t)g * p.ldx;
        const uint32_t *lab = (const uint32_t *)(p.lab8 + (int64_t)g * p.ldl);
#pragma unroll
        for (int e = 0; e < EPL; e++) {
            const int s = lane + 32 * e;
            xc_r[e] = s < n ? xc[s] : 0.0;
        }
#pragma unroll
        for (int q = 0; q < NLQ; q++) lraw[q] = 4 * (lane + 32 * q) < (int)p.ldl ? lab[lane + 32 * q] : 0x03030303u;
        pend_gene = g;
    };

    int it = gw;
    Item item, nitem;
    item.gene = -1; item.begin = 0; item.count = 0; item.pad = 0;
    nitem = item;
    int my_pi = -1, my_slot = -1, nmy_pi = -1, nmy_slot = -1;
    if (it < p.n_items) {
        item = p.items[it];
        if (lane < item.count) { my_pi = p.probe_idx[item.begin + lane]; my_slot = p.pair_slot[item.begin + lane]; }
        const int fpi = __shfl_sync(FULL, my_pi, 0), fsl = __shfl_sync(FULL, my_slot, 0);
        if (fpi >= 0) issue(fpi, fsl);
        load_gene(item.gene);
    }

    while (it < p.n_items) {
        /* Work items are handed out dynamically (they hold 1..32 pairs: a static round-robin leaves the busiest
         * warp with 1.75x the mean load on the SKCM index). The first item of a warp is its own number, the
         * following ones come from a global ticket counter. Descriptor + probe list of the NEXT item are fetched
         * now so that the latency hides behind this item. */
        int nit = 0;
        if (lane == 0) nit = nw + atomicAdd(p.queue, 1);
        nit = __shfl_sync(FULL, nit, 0);
        const bool has_next = nit < p.n_items;
        nmy_pi = -1; nmy_slot = -1;
        if (has_next) {
            nitem = p.items[nit];
            if (lane < nitem.count) { nmy_pi = p.probe_idx[nitem.begin + lane]; nmy_slot = p.pair_slot[nitem.begin + lane]; }
        }
        if (pend_gene >= 0) {
            /* second half of the gene change (the loads were issued before the previous item's epilogue) */
            cur_gene = pend_gene;
            pend_gene = -1;
            if (lane == 0) tab[n] = make_uint2(0u, 0u);
#pragma unroll
            for (int q = 0; q < NLQ; q++) {
#pragma unroll
                for (int b4 = 0; b4 < 4; b4++) {
                    const int s4 = 4 * (lane + 32 * q) + b4;
                    const unsigned l = (lraw[q] >> (8 * b4)) & 0xffu;
                    if (s4 < n)
                        tab[s4] = make_uint2(l == 0 ? WALK_A0 : (l == 1 ? WALK_A1 : (l == 2 ? WALK_A2 : 0u)),
                                             l == 0 ? WALK_B0 : (l == 1 ? WALK_B1 : (l == 2 ? WALK_B2 : 0u)));
                }
            }
            const double *ga = p.gene_aux + 4 * (int64_t)cur_gene;
            x_const = ga[0] == 0.0;
            sxc = ga[2];
            rsx = ga[3];
            perm_g = p.perm + (int64_t)cur_gene * n;
            if (m >= 2) {
                pm[0] = perm_g[0]; pm[1] = perm_g[1]; pm[2] = perm_g[m]; pm[3] = perm_g[m + 1];
                pm[4] = perm_g[2 * m]; pm[5] = perm_g[2 * m + 1];
            }
            __syncwarp();
            if (TMODE) { /* label of this lane's element e, for the per-group sums */
                labs = 0; labs64 = 0;
#pragma unroll
                for (int e = 0; e < EPL; e++) {
                    const int s = lane + 32 * e;
                    unsigned l = 3;
                    if (s < n) { const unsigned a = tab[s].x; l = a == WALK_A0 ? 0u : (a == WALK_A1 ? 1u : (a == WALK_A2 ? 2u : 3u)); }
                    if (EPL <= 16) labs |= l << (2 * e); else labs64 |= (unsigned long long)l << (2 * e);
                }
            }
        }

        for (int jj = 0; jj < item.count; jj++) {
            const int pi = __shfl_sync(FULL, my_pi, jj);
            /* prefetch the next pair's row + order into the other stage */
            int npi, nsl;
            if (jj + 1 < item.count) { npi = __shfl_sync(FULL, my_pi, jj + 1); nsl = __shfl_sync(FULL, my_slot, jj + 1); }
            else { npi = __shfl_sync(FULL, nmy_pi, 0); nsl = __shfl_sync(FULL, nmy_slot, 0); }
            __syncwarp(); /* every lane has finished reading the stage about to be refilled */
            if (npi >= 0) issue(npi, nsl);

            if (pi < 0) continue; /* annotated probe without data: the epilogue writes its all-NA row */
            const double *row = (const double *)(buf0 + (use ? stageb : 0u));
            const uint32_t *ordw = (const uint32_t *)(buf0 + (use ? stageb : 0u) + rowb);
            if (use == 0) { mbar_wait(&bar[0], ph0); ph0 ^= 1; } else { mbar_wait(&bar[1], ph1); ph1 ^= 1; }
            use ^= 1;

            /* ---- pass A: pivot-shifted moments, sample-major (conflict-free LDS.64) ---- */
            const double K = row[0];
            double sd = 0.0, sdd = 0.0, sxd = 0.0;
            double g0 = 0.0, g1 = 0.0, g2 = 0.0, q0 = 0.0, q1 = 0.0, q2 = 0.0, g3 = 0.0, q3 = 0.0;
#pragma unroll
            for (int e = 0; e < EPL; e++) {
                const int s = lane + 32 * e;
                /* out-of-range lanes of the last element read row[0]: d = 0 contributes nothing */
                const double d = row[(e < EPL - 1 || s < n) ? s : 0] - K;
                sd += d;
                sdd = fma(d, d, sdd);
                sxd = fma(xc_r[e], d, sxd);
                if (TMODE) {
                    /* groups 0 and 1 (and the <= 2 unlabelled samples) explicitly, group 2 by subtraction */
                    const unsigned l = EPL <= 16 ? (labs >> (2 * e)) & 3u : (unsigned)(labs64 >> (2 * e)) & 3u;
                    const double d0 = l == 0 ? d : 0.0, d1 = l == 1 ? d : 0.0;
                    g0 += d0; q0 = fma(d0, d0, q0);
                    g1 += d1; q1 = fma(d1, d1, q1);
                    if (has_unl) { const double d3 = (l == 3 && (e < EPL - 1 || s < n)) ? d : 0.0; g3 += d3; q3 = fma(d3, d3, q3); }
                }
            }
            sd = warp_sum(sd); sdd = warp_sum(sdd); sxd = warp_sum(sxd);
            double syy = sdd - sd * sd * inv_n;
            double sxy = sxd - (sd * inv_n) * sxc;
            double mg0 = 0, mg1 = 0, mg2 = 0, ss0 = 0, ss1 = 0, ss2 = 0;
            bool redo = sdd > 0.0 && syy < 1e-7 * sdd;
            if (TMODE) {
                g0 = warp_sum(g0); g1 = warp_sum(g1); q0 = warp_sum(q0); q1 = warp_sum(q1);
                if (has_unl) { g3 = warp_sum(g3); q3 = warp_sum(q3); }
                g2 = sd - g0 - g1 - g3; q2 = sdd - q0 - q1 - q3;
                ss0 = q0 - g0 * g0 * inv_m; ss1 = q1 - g1 * g1 * inv_m; ss2 = q2 - g2 * g2 * inv_m;
                mg0 = K + g0 * inv_m; mg1 = K + g1 * inv_m; mg2 = K + g2 * inv_m;
                redo = redo || (q0 > 0.0 && ss0 < 1e-7 * q0) || (q1 > 0.0 && ss1 < 1e-7 * q1) || (q2 > 0.0 && ss2 < 1e-7 * q2);
                /* the third group's sum of squares is a difference: an outlier in another group (|d| 300 times the
                 * rest and up) leaves it with few good digits */
                redo = redo || q2 < 1e-5 * sdd;
            }
            if (sdd == 0.0) syy = 0.0; /* constant row: sd is exactly 0 in R -> NA */
            if (redo) {
                /* the pivot sat far outside a very tight row: redo with centred two-pass sums (rare) */
                const double mean_all = K + sd * inv_n;
                double a0 = 0, a1 = 0, b0 = 0, b1 = 0, b2 = 0;
#pragma unroll
                for (int e = 0; e < EPL; e++) {
                    const int s = lane + 32 * e;
                    if (s < n) {
                        const double yv = row[s];
                        const double d = yv - mean_all;
                        a0 = fma(d, d, a0);
                        a1 = fma(xc_r[e], d, a1);
                        if (TMODE) {
                            const unsigned l = EPL <= 16 ? (labs >> (2 * e)) & 3u : (unsigned)(labs64 >> (2 * e)) & 3u;
                            const double dg = yv - (l == 0 ? mg0 : (l == 1 ? mg1 : mg2));
                            if (l == 0) b0 = fma(dg, dg, b0); else if (l == 1) b1 = fma(dg, dg, b1); else if (l == 2) b2 = fma(dg, dg, b2);
                        }
                    }
                }
                syy = warp_sum(a0); sxy = warp_sum(a1);
                if (TMODE) { ss0 = warp_sum(b0); ss1 = warp_sum(b1); ss2 = warp_sum(b2); }
            }

            /* ---- sorted walk: lane owns EPW consecutive positions of the probe's ascending order ---- */
            unsigned ow[(EPW + 1) / 2];
            if (EPW >= 8) {
#pragma unroll
                for (int c = 0; c < EPW / 8; c++) {
                    const uint4 q = *reinterpret_cast<const uint4 *>(ordw + lane * (EPW / 2) + c * 4);
                    ow[c * 4 + 0] = q.x; ow[c * 4 + 1] = q.y; ow[c * 4 + 2] = q.z; ow[c * 4 + 3] = q.w;
                }
            } else if (EPW >= 2) {
#pragma unroll
                for (int c = 0; c < EPW / 2; c++) ow[c] = ordw[lane * (EPW / 2) + c];
            } else {
                ow[0] = ((const unsigned short *)ordw)[lane];
            }
            bool anytie = false;
            if (!TMODE) {
                unsigned tb = 0;
#pragma unroll
                for (int c = 0; c < (EPW + 1) / 2; c++) tb |= ow[c];
                anytie = __any_sync(FULL, (tb & 0x80008000u) != 0u);
            }
            /* One pass, lane-local: ra = [cUP-cMID | cUP-cDOWN], rb = [cMID-cDOWN | cUP] as biased 16-bit halves
             * of one word each, advanced by the per-sample increment pair of the gene's table (pad positions
             * k >= n carry index n, whose entry is 0). The ECDF differences are evaluated with packed 16x2
             * max/min (VIMNMX.U16x2) at the END of every tie run; the counts in front of the lane are added
             * afterwards (max(prefix + x) = prefix + max(x)). */
            unsigned ra = WALK_BIAS, rb = WALK_BIAS;
            unsigned mxa = 0u, mna = 0xffffffffu, mxb = 0u, mnb = 0xffffffffu;
            /* table entry of the sample in the low / high half of an order word whose tie marks are cleared:
             * indices are <= 1024, so bits 13-15 of the low half are 0 and (w >> 13) is 8 * the high index */
#define EPX_TAB_OF(w, e) lds_u2(((e) & 1) ? tab_sa + ((w) >> 13) : tab_sa + 8u * __byte_perm((w), 0u, 0x4410u))
            if (TMODE || !anytie) {
#pragma unroll
                for (int e = 0; e < EPW; e++) {
                    const unsigned w = TMODE ? (ow[e / 2] & 0x7fff7fffu) : ow[e / 2];
                    const uint2 inc = EPX_TAB_OF(w, e);
                    ra += inc.x; rb += inc.y;
                    if (!TMODE) {
                        mxa = __vmaxu2(mxa, ra); mna = __vminu2(mna, ra);
                        mxb = __vmaxu2(mxb, rb); mnb = __vminu2(mnb, rb);
                    }
                }
            } else {
#pragma unroll
                for (int e = 0; e < EPW; e++) {
                    const unsigned w = ow[e / 2] & 0x7fff7fffu;
                    const uint2 inc = EPX_TAB_OF(w, e);
                    ra += inc.x; rb += inc.y;
                    if (!(ow[e / 2] & ((e & 1) ? 0x80000000u : 0x8000u))) { /* inside a tie run nothing is evaluated */
                        mxa = __vmaxu2(mxa, ra); mna = __vminu2(mna, ra);
                        mxb = __vmaxu2(mxb, rb); mnb = __vminu2(mnb, rb);
                    }
                }
            }
#undef EPX_TAB_OF
            /* inclusive scan of the lane totals (32-bit wrap-around sums of the two 16-bit fields) */
            const unsigned da = ra - WALK_BIAS, db = rb - WALK_BIAS;
            unsigned ia = da, ib = db;
#pragma unroll
            for (int d = 1; d < 32; d <<= 1) {
                const unsigned ta = __shfl_up_sync(FULL, ia, d), tb2 = __shfl_up_sync(FULL, ib, d);
                if (lane >= d) { ia += ta; ib += tb2; }
            }
            const unsigned ea = ia - da, eb = ib - db; /* counts in front of this lane */
            int d01 = 0, d02 = 0, d12 = 0;
            if (!TMODE) {
                int k01 = 0, k02 = 0, k12 = 0;
                if (mxa != 0u) { /* the lane evaluated at least one position */
                    const unsigned xa = mxa + ea, na_ = mna + ea, xb = mxb + eb, nb = mnb + eb;
                    k01 = max((int)(xa >> 16) - 0x4000, 0x4000 - (int)(na_ >> 16));
                    k02 = max((int)(xa & 0xffffu) - 0x4000, 0x4000 - (int)(na_ & 0xffffu));
                    k12 = max((int)(xb >> 16) - 0x4000, 0x4000 - (int)(nb >> 16));
                }
                d01 = warp_max_i(k01); d02 = warp_max_i(k02); d12 = warp_max_i(k12);
            }
            /* medians: the lane whose positions hold the t-th member of a group knows it from the scanned counts
             * and leaves (lane, rank inside the lane) for the item's epilogue, which does the search */
            if (m >= 1) {
                const unsigned eab = ea + WALK_BIAS, iab = ia + WALK_BIAS;
                const int ce0 = (int)(eb & 0xffffu), ci0 = (int)(ib & 0xffffu);
                const int ce1 = ce0 - ((int)(eab >> 16) - 0x4000), ci1 = ci0 - ((int)(iab >> 16) - 0x4000);
                const int ce2 = ce0 - ((int)(eab & 0xffffu) - 0x4000), ci2 = ci0 - ((int)(iab & 0xffffu) - 0x4000);
                unsigned short *own = s_own + jj * 6;
                const unsigned lane6 = (unsigned)lane << 6;
                if (ce0 < t1 && t1 <= ci0) own[0] = (unsigned short)(lane6 | (unsigned)(t1 - ce0));
                if (ce1 < t1 && t1 <= ci1) own[2] = (unsigned short)(lane6 | (unsigned)(t1 - ce1));
                if (ce2 < t1 && t1 <= ci2) own[4] = (unsigned short)(lane6 | (unsigned)(t1 - ce2));
                if (m_even) {
                    if (ce0 < t2 && t2 <= ci0) own[1] = (unsigned short)(lane6 | (unsigned)(t2 - ce0));
                    if (ce1 < t2 && t2 <= ci1) own[3] = (unsigned short)(lane6 | (unsigned)(t2 - ce1));
                    if (ce2 < t2 && t2 <= ci2) own[5] = (unsigned short)(lane6 | (unsigned)(t2 - ce2));
                }
            }
            unsigned tiebits = 0;
            if (!TMODE && p.exact_ok && anytie) {
                /* small groups with tied values: ks.test drops to the asymptotic p-value when the POOLED pair of
                 * groups holds a tie. Rare and tiny (m < 100): one lane walks the order. */
                if (lane == 0) {
                    const unsigned short *ord16 = (const unsigned short *)ordw;
                    int r0 = 0, r1 = 0, r2 = 0;
                    for (int k = 0; k < n; k++) {
                        const unsigned short v = ord16[k];
                        const unsigned w = tab[v & 0x7fffu].x;
                        r0 += w == WALK_A0; r1 += w == WALK_A1; r2 += w == WALK_A2;
                        if (!(v & 0x8000u)) {
                            if (r0 + r1 >= 2) tiebits |= 1u;
                            if (r0 + r2 >= 2) tiebits |= 2u;
                            if (r1 + r2 >= 2) tiebits |= 4u;
                            r0 = r1 = r2 = 0;
                        }
                    }
                }
                tiebits = __shfl_sync(FULL, tiebits, 0);
            }
            /* ---- everything else of the pair is scalar work: lane jj keeps the sums for the item's epilogue ---- */
            if (lane == jj) {
                st_syy = syy; st_sxy = sxy;
                st_ks = (unsigned)d01 | ((unsigned)d02 << 9) | ((unsigned)d12 << 18) | (tiebits << 27);
            }
            if (TMODE && lane < 3) {
                /* stats::t.test(var.equal = FALSE) of comparison `lane` (UPvsMID, UPvsDOWN, MIDvsDOWN): se^2 = vx/nx + vy/ny,
                 * Welch-Satterthwaite df; "data are essentially constant" -> NaN. t waits in pair_stat and df in the
                 * p-value column it belongs to; the item's second stage turns them into p-values, 32 at a time. */
                const double sc = 1.0 / (dm * (dm - 1.0));
                const double ma = lane == 2 ? mg1 : mg0, mb = lane == 0 ? mg1 : mg2;
                const double A = (lane == 2 ? ss1 : ss0) * sc, B = (lane == 0 ? ss1 : ss2) * sc, se2 = A + B;
                const double thr = 10.0 * 2.220446049250313e-16 * fmax(fabs(ma), fabs(mb));
                const bool ok = m >= 2 && !(se2 < thr * thr);
                const int64_t j = (int64_t)item.begin + jj;
                p.pair_stat[j * 3 + lane] = ok ? (ma - mb) / sqrt(se2) : EPX_NAN;
                p.pair[j * PAIR_COLS + EPX_P_UP_MID + lane] = se2 * se2 * (dm - 1.0) / (A * A + B * B);
            }
        }

        if (has_next && nitem.gene != cur_gene) load_gene(nitem.gene);

        /* ---- epilogue of the item, one pair per lane: medians, guard, tests, p-values, result rows ---- */
        __syncwarp();
        double r_keep = EPX_NAN; /* t mode: this lane's pair's Pearson r, for the second stage */
        if (lane < item.count) {
            const int64_t j = (int64_t)item.begin + lane;
            double *orow = p.pair + j * PAIR_COLS;
            if (my_pi < 0) { /* annotated probe without data: all-NA column (R/EpimethexAnalysis.R:375-394) */
#pragma unroll
                for (int c = 0; c < PAIR_COLS; c++) orow[c] = EPX_NAN;
#pragma unroll
                for (int c = 0; c < 3; c++) { p.pair_stat[j * 3 + c] = EPX_NAN; p.pair_dnum[j * 3 + c] = -1; }
                p.pair_na[j] = (uint16_t)((1u << PAIR_COLS) - 1);
            } else {
                const double *grow = p.meth + (int64_t)my_pi * p.ld;
                const uint16_t *gord = p.ord + (int64_t)my_slot * p.ldo;
                /* group medians: walk the EPW positions of the owning lane (labels: 0 UP, 1 Medium, 2 Down) */
                double med[3] = { EPX_NAN, EPX_NAN, EPX_NAN };
                if (m >= 1) {
#pragma unroll
                    for (int g = 0; g < 3; g++) {
                        const unsigned labc = g == 0 ? WALK_A0 : (g == 1 ? WALK_A1 : WALK_A2);
                        double v0 = 0.0, v1 = 0.0;
#pragma unroll
                        for (int hs = 0; hs < 2; hs++) {
                            if (hs == 1 && !m_even) { v1 = v0; continue; }
                            const unsigned o = s_own[lane * 6 + g * 2 + hs];
                            const int L = (int)(o >> 6), k = (int)(o & 63u);
                            const uint16_t *src = gord + L * EPW;
                            unsigned hit = 0;
                            int c = 0;
                            if (EPW >= 8) {
#pragma unroll
                                for (int q4 = 0; q4 < EPW / 8; q4++) {
                                    const uint4 q = __ldg(reinterpret_cast<const uint4 *>(src) + q4);
                                    const unsigned w4[4] = { q.x, q.y, q.z, q.w };
#pragma unroll
                                    for (int e = 0; e < 8; e++) {
                                        const unsigned h = ((e & 1) ? (w4[e / 2] >> 16) : w4[e / 2]) & 0x7fffu;
                                        const bool is = lds_u2(tab_sa + 8u * h).x == labc;
                                        c += is;
                                        if (is && c == k) hit = h;
                                    }
                                }
                            } else {
#pragma unroll
                                for (int e = 0; e < EPW; e++) {
                                    const unsigned h = __ldg(src + e) & 0x7fffu;
                                    const bool is = lds_u2(tab_sa + 8u * h).x == labc;
                                    c += is;
                                    if (is && c == k) hit = h;
                                }
                            }
                            const double v = __ldg(grow + hit);
                            if (hs == 0) v0 = v; else v1 = v;
                        }
                        med[g] = (v0 + v1) * 0.5;
                    }
                }
                const double medUP = med[0], medMID = med[1], medDOWN = med[2];

                /* guard of calculateTtest (R/Functions.R:193): are all rank-paired differences identical? */
                int guard[3] = { -1, -1, -1 };
                if (m >= 2) {
                    const double a0 = __ldg(grow + pm[0]), a1 = __ldg(grow + pm[1]), b0 = __ldg(grow + pm[2]),
                                 b1 = __ldg(grow + pm[3]), e0 = __ldg(grow + pm[4]), e1 = __ldg(grow + pm[5]);
                    const double dd[3] = { a0 - b0, a0 - e0, b0 - e0 };
                    const bool eq[3] = { dd[0] == a1 - b1, dd[1] == a1 - e1, dd[2] == b1 - e1 };
                    const int offA[3] = { 0, 0, m }, offB[3] = { m, 2 * m, 2 * m };
#pragma unroll
                    for (int c = 0; c < 3; c++) {
                        guard[c] = 1;
                        if (eq[c]) { /* rare: walk the whole group */
                            int differs = 0;
                            for (int i = 2; i < m && !differs; i++)
                                if (__ldg(grow + perm_g[offA[c] + i]) - __ldg(grow + perm_g[offB[c] + i]) != dd[c]) differs = 1;
                            guard[c] = differs;
                        }
                    }
                }

                unsigned na = 0;
                double r = EPX_NAN;
                if (!x_const && st_syy != 0.0 && n >= 3) {
                    r = st_sxy * rsx * rsqrt(st_syy);
                    r = r > 1.0 ? 1.0 : (r < -1.0 ? -1.0 : r);
                } else na |= (1u << EPX_PEARSON_R) | (1u << EPX_PEARSON_P);
                orow[0] = medDOWN; orow[1] = medMID; orow[2] = medUP;
                orow[3] = medUP - medMID; orow[4] = medUP - medDOWN; orow[5] = medMID - medDOWN;
                orow[9] = r;
                if (TMODE) {
                    /* the Welch p-values follow in the second stage (one comparison per lane); here only what decides
                     * NA: the guard (t.test's own stops were folded into t = NaN when it was computed) */
#pragma unroll
                    for (int c = 0; c < 3; c++) {
                        const bool ok = guard[c] == 1 && !isnan(__ldcg(p.pair_stat + j * 3 + c));
                        if (!ok) { na |= 1u << (EPX_P_UP_MID + c); p.pair_stat[j * 3 + c] = EPX_NAN; }
                        p.pair_dnum[j * 3 + c] = -1;
                    }
                    r_keep = r;
                } else {
#pragma unroll
                    for (int c = 0; c < 3; c++) {
                        const int dk = (int)((st_ks >> (9 * c)) & 511u);
                        double stat = EPX_NAN, pv = EPX_NAN;
                        int dnum = -1;
                        if (guard[c] != 0) { /* difference != 0 || is.na(difference) */
                            dnum = dk * m;
                            stat = (double)dk * inv_m; /* = dnum / m^2 to the last bit or one ulp */
                            pv = (p.exact_ok && !((st_ks >> (27 + c)) & 1u)) ? __ldg(p.lut_exact + dk) : __ldg(p.lut_asym + dk);
                        } else na |= 1u << (EPX_P_UP_MID + c);
                        orow[6 + c] = pv;
                        p.pair_stat[j * 3 + c] = stat;
                        p.pair_dnum[j * 3 + c] = dnum;
                    }
                    orow[10] = isnan(r) ? EPX_NAN : epx_pearson_p_tab(r, p.pt_tab, p.pt_n, dn_ - 2.0);
                }
                p.pair_na[j] = (uint16_t)na;
            }
        }

        if (TMODE) {
            /* ---- second stage: the p-values of the whole item, one per lane (3 Welch + 1 Pearson per pair) ---- */
            __syncwarp();
            for (int q0 = 0; q0 < item.count * 4; q0 += 32) {
                const int q = q0 + lane, jj = q >> 2, slot = q & 3;
                const double rr = __shfl_sync(FULL, r_keep, jj & 31);
                if (q < item.count * 4) {
                    const int64_t j = (int64_t)item.begin + jj;
                    double pq = EPX_NAN;
                    int col = EPX_PEARSON_P;
                    if (slot < 3) {
                        col = EPX_P_UP_MID + slot;
                        const double t = __ldcg(p.pair_stat + j * 3 + slot);
                        if (!isnan(t)) pq = welch_p(p, t, __ldcg(p.pair + j * PAIR_COLS + col));
                    } else if (!isnan(rr)) pq = epx_pearson_p_tab(rr, p.pt_tab, p.pt_n, dn_ - 2.0);
                    p.pair[j * PAIR_COLS + col] = pq;
                }
            }
        }
        __syncwarp();
        item = nitem;
        my_pi = nmy_pi; my_slot = nmy_slot;
        it = nit;
    }
    /* the last warp to leave re-arms the queue for the next launch */
    if (lane == 0) {
        __threadfence();
        if (atomicAdd(p.queue + 1, 1) == nw - 1) { p.queue[0] = 0; p.queue[1] = 0; }
    }
}

} // namespace epx
