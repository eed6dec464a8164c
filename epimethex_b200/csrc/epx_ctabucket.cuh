/*
 * epx_ctabucket.cuh — one CTA orders one long row (1024 < n <= 16384 values, e.g. the 9,000-sample pan-cancer
 * shape) by BUCKET PLACEMENT: the CTA-wide form of WarpSort::order_bucket (epx_warpsort.cuh).
 *
 *   1. range of the row (exact minimum and maximum, from the order-preserving keys of epx_device.cuh);
 *   2. every value is counted into bucket Q >> 15, Q = round((x - first) * scale) a 30-bit fixed-point position in
 *      the row's range, ~1.03 buckets per slot, shared-memory atomics (ATOMS ~2.7 cycles per warp instruction when
 *      the addresses differ, profiles/micro/hist_micro.cu);
 *   3. exclusive scan of the counters; every value draws the next free slot of its bucket and drops the word
 *      [pad flag | 17 leading bits of Q | 14-bit sample index] there, so the row is ordered but for the values
 *      that share a bucket;
 *   4. those are settled by the 32-value network inside each thread (WarpSort<32>::sort_thread) and one merge of
 *      the windows across the thread boundaries (runs of up to 17 values are put right);
 *   5. whether that was enough is CHECKED (one comparison per neighbour pair, __syncthreads_or); a row that fails
 *      is ordered by the merge sort of epx_ctasort.cuh instead, so the result never depends on the data's shape;
 *   6. neighbours with equal 17-bit keys (exact ties, or values closer than 2^-17 of ~one bucket) are settled on
 *      the full values: odd-even transposition over exactly those pairs, then the tie marks.
 *
 * Values equal to the row's minimum or maximum (beta values clipped at a detection limit, zero counts: piles of
 * hundreds of equal values) do not draw slots in arrival order -- CTA-wide atomics have none -- but by sample
 * index: a membership bitmap per pile, prefix pop-counts, rank = members with a smaller index.
 *
 * Replaces: the per-call sort() inside median()/ks.test() (R/Functions.R:85-132) and order() of
 * R/EpimethexAnalysis.R:117-122 for long rows; same output as cta_sort_row (epx_ctasort.cuh).
 */
#pragma once
#include "epx_ctasort.cuh"

namespace epx {

constexpr int CTAB_W = 32;       /* slots per thread */
constexpr int CTAB_STRIDE = 33;  /* words between the slot blocks of two threads (odd: conflict-free 32-bit accesses) */
constexpr int CTAB_IDXB = 14;    /* sample index bits (n <= 16384) */
constexpr unsigned CTAB_IDXM = (1u << CTAB_IDXB) - 1u;
constexpr int CTAB_SH = 15;      /* Q >> 15 = bucket; Q < 16896 << 15 < 2^30: bits 30 and 31 are free for the pile flags */
constexpr int CTAB_MAXPASS = 48; /* transposition rounds before the row is handed to the merge sort instead */

__host__ __device__ inline int ctab_threads(int n) { return 32 * ((n + 1023) / 1024); }
/* shared memory besides the row: counters/slots (33 words per thread; also holds the merge sort's index buffers when a
 * row falls back) | two pile bitmaps | their prefix counts | scratch of the reductions */
__host__ __device__ inline size_t ctab_smem(int n)
{
    const size_t T = (size_t)ctab_threads(n);
    return T * CTAB_STRIDE * 4 + 3 * T * 4 + 64 * 4;
}

/* exclusive scan of one unsigned per thread over the CTA; *total = the sum. red: >= 33 words. Ends on a barrier. */
__device__ __forceinline__ unsigned ctab_scan(unsigned v, unsigned *red, unsigned *total)
{
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarps = blockDim.x >> 5;
    unsigned incl = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const unsigned y = __shfl_up_sync(FULL, incl, o);
        if (lane >= o) incl += y;
    }
    __syncthreads(); /* red[] free */
    if (lane == 31) red[warp] = incl;
    __syncthreads();
    if (warp == 0) {
        const unsigned w = lane < nwarps ? red[lane] : 0u;
        unsigned wi = w;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const unsigned y = __shfl_up_sync(FULL, wi, o);
            if (lane >= o) wi += y;
        }
        red[lane] = wi - w; /* exclusive prefix of the warp totals; wi of the last lane = the total */
        if (lane == 31) red[32] = wi;
    }
    __syncthreads();
    *total = red[32];
    return red[warp] + incl - v;
}

/*
 * s_val: the row (n doubles, shared). region: ctab_smem(n) bytes. All ctab_threads(n) threads of the CTA must call.
 * Returns the order row (n uint16: sample index of the k-th value, bit 15 = "the next value is equal"), which lies in
 * `region`; ends on a barrier.
 */
template <bool DESC>
__device__ const unsigned short *cta_bucket_order(const double *s_val, const int n, unsigned *region, const unsigned one,
                                                  int32_t *full_rows)
{
    const int T = (int)blockDim.x, t = (int)threadIdx.x, lane = t & 31, warp = t >> 5, nwarps = T >> 5;
    const int NB = CTAB_STRIDE * T;
    unsigned *bm = region + (size_t)CTAB_STRIDE * T; /* [2][T] membership of the two piles, one bit per sample */
    unsigned *pre = bm + 2 * T;                      /* [T] members in front of each bitmap word: first pile | last pile << 16 */
    unsigned *red = pre + T;                         /* 64 words */
    const int qw = ((unsigned)NB << 3) > (1u << (31 - CTAB_IDXB)) ? 13 : 12; /* Q >> qw: the (at most 17) bits of Q kept in the word */

    /* ---- clean counters and bitmaps; exact range of the row ---- */
    {   /* 33 counter words and 2 bitmap words per thread, contiguous: 16 bytes per store */
        uint4 *z = reinterpret_cast<uint4 *>(region);
        const int nz = (CTAB_STRIDE + 2) * T / 4;
        for (int i = t; i < nz; i += T) z[i] = make_uint4(0u, 0u, 0u, 0u);
    }
    unsigned long long kmin = ~0ull, kmax = 0ull;
#pragma unroll 4
    for (int e = 0; e < CTAB_W; e++) {
        const int i = t + T * e;
        const unsigned long long k = key_of(s_val[i < n ? i : 0]);
        kmin = k < kmin ? k : kmin;
        kmax = k > kmax ? k : kmax;
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        const unsigned long long a = __shfl_xor_sync(FULL, kmin, o), b = __shfl_xor_sync(FULL, kmax, o);
        kmin = a < kmin ? a : kmin;
        kmax = b > kmax ? b : kmax;
    }
    unsigned long long *red64 = reinterpret_cast<unsigned long long *>(red);
    if (lane == 0) { red64[warp] = kmin; red64[16 + warp] = kmax; }
    __syncthreads();
    for (int w = 0; w < nwarps; w++) {
        const unsigned long long a = red64[w], b = red64[16 + w];
        kmin = a < kmin ? a : kmin;
        kmax = b > kmax ? b : kmax;
    }
    auto unkey = [](unsigned long long k) {
        return __longlong_as_double((long long)((k >> 63) ? (k ^ 0x8000000000000000ull) : ~k));
    };
    const double xlo = unkey(kmin), xhi = unkey(kmax);
    const double first = DESC ? xhi : xlo, last = DESC ? xlo : xhi, span = last - first;
    /* the row's first value sits near the top of its bucket and the last one near the bottom of its bucket */
    const unsigned QMAX = (unsigned)NB << CTAB_SH;
    double scale = (double)(((unsigned)(NB - 2) << CTAB_SH) + 128u) / span;
    if (!(fabs(scale) < 1e300)) scale = 0.0; /* constant row, or a span beyond the exponent range */
    const double c0 = fma(-first, scale, 6755399441055744.0 + (double)((1u << CTAB_SH) - 64u));
    unsigned qf = (unsigned)__double2loint(fma(first, scale, c0));
    qf = min(qf, QMAX - 1u);
    const int b_first = (int)(qf >> CTAB_SH); /* its pile draws no slots from this bucket's counter */

    /* ---- counting pass ---- */
    unsigned Q[CTAB_W]; /* per element: Q; after the placement pass: leading bits of Q | slot */
#pragma unroll
    for (int e = 0; e < CTAB_W; e++) {
        const int i = t + T * e;
        Q[e] = 0xffffffffu;
        if (i < n) {
            const double x = s_val[i];
            unsigned q = (unsigned)__double2loint(fma(x, scale, c0));
            q = min(q, QMAX - 1u); /* memory safety only: a q out of range means the row fails the check below */
            atomicAdd(region + (q >> CTAB_SH), 1u);
            /* bits 31/30 of the kept word: member of the first / last pile */
            if (x == first) { atomicOr(bm + (i >> 5), 1u << (i & 31)); q |= 0x80000000u; }
            else if (x == last) { atomicOr(bm + T + (i >> 5), 1u << (i & 31)); q |= 0x40000000u; }
            Q[e] = q;
        }
    }
    __syncthreads();
    /* ---- scans: pile members in front of every bitmap word, then the bucket counters ---- */
    unsigned pile_tot;
    {
        const unsigned pc = (unsigned)__popc(bm[t]) | ((unsigned)__popc(bm[T + t]) << 16);
        pre[t] = ctab_scan(pc, red, &pile_tot);
    }
    const unsigned pile0 = pile_tot & 0xffffu, pile1 = pile_tot >> 16;
    {   /* the counters are read twice rather than kept: Q[] is live across this scan */
        unsigned *mine = region + CTAB_STRIDE * t;
        unsigned tot = 0, dummy;
#pragma unroll
        for (int i = 0; i < CTAB_STRIDE; i++) tot += mine[i];
        unsigned run = ctab_scan(tot, red, &dummy);
#pragma unroll
        for (int i = 0; i < CTAB_STRIDE; i++) {
            const unsigned cc = mine[i];
            mine[i] = run;
            run += cc;
        }
        /* the first pile takes the first pile0 slots of its bucket by sample index; the counter starts behind them */
        const int own = b_first - CTAB_STRIDE * t;
        if (own >= 0 && own < CTAB_STRIDE) mine[own] += pile0;
    }
    __syncthreads();
    /* ---- placement: slot of every value ---- */
#pragma unroll
    for (int e = 0; e < CTAB_W; e++) {
        const int i = t + T * e;
        const unsigned q = Q[e];
        unsigned pos, top;
        if (i >= n) { pos = (unsigned)i; top = 0x80000000u; }
        else {
            const unsigned qq = q & 0x3fffffffu;
            top = (qq >> qw) << CTAB_IDXB;
            if (q & 0xc0000000u) {
                const bool lastp = !(q & 0x80000000u);
                const unsigned w = bm[(lastp ? T : 0) + (i >> 5)], pr = pre[i >> 5];
                const unsigned rank = (lastp ? (pr >> 16) : (pr & 0xffffu)) + (unsigned)__popc(w & ((1u << (i & 31)) - 1u));
                pos = lastp ? (unsigned)n - pile1 + rank : rank; /* no value lies below the first pile, none above the last */
            } else pos = atomicAdd(region + (qq >> CTAB_SH), 1u);
        }
        Q[e] = top | pos;
    }
    __syncthreads();
    /* ---- the words move to their slots, over the counters: slot k lives at word k + k / 32 ---- */
#pragma unroll
    for (int e = 0; e < CTAB_W; e++) {
        const int i = t + T * e;
        const unsigned pos = Q[e] & CTAB_IDXM;
        region[pos + (pos >> 5)] = (Q[e] & ~CTAB_IDXM) | (unsigned)(i < n ? i : (int)pos);
    }
    __syncthreads();
    /* ---- short network: 32 values inside each thread, then the windows across the thread boundaries ---- */
    const unsigned minus1 = 0u - one;
    unsigned v[CTAB_W];
    unsigned *mine = region + CTAB_STRIDE * t;
#pragma unroll
    for (int r = 0; r < CTAB_W; r++) v[r] = mine[r];
    WarpSort<CTAB_W>::sort_thread(v, one, minus1);
#pragma unroll
    for (int r = 0; r < CTAB_W; r++) mine[r] = v[r];
    __syncthreads();
    if (t + 1 < T) {
#pragma unroll
        for (int r = 0; r < 16; r++) { v[r] = mine[16 + r]; v[16 + r] = mine[CTAB_STRIDE + r]; }
#pragma unroll
        for (int r = 0; r < 16; r++) { /* both halves ascend: mirror step, then half-cleaners inside each half */
            const unsigned a = v[r], b = v[31 - r];
            v[r] = min(a, b);
            v[31 - r] = max(a, b);
        }
        WarpSort<CTAB_W>::thread_tail(v, 8, one, minus1);
#pragma unroll
        for (int r = 0; r < 16; r++) { mine[16 + r] = v[r]; mine[CTAB_STRIDE + r] = v[16 + r]; }
    }
    __syncthreads();
    unsigned nx0 = t + 1 < T ? mine[CTAB_STRIDE] : 0xffffffffu;
#pragma unroll
    for (int r = 0; r < CTAB_W; r++) v[r] = mine[r];
    bool bad = v[CTAB_W - 1] > nx0;
#pragma unroll
    for (int r = 0; r + 1 < CTAB_W; r++) bad = bad || (v[r] > v[r + 1]);
    if (__syncthreads_or(bad)) { /* the barrier also means: every thread holds its words */
        if (t == 0 && full_rows) atomicAdd(full_rows, 1);
        return nullptr;
    }
    /* ---- neighbours with equal leading bits: settle on the full values ---- */
    unsigned marks = 0; /* bit r: positions 32t + r and 32t + r + 1 share their leading bits (pads never do) */
#pragma unroll
    for (int r = 0; r < CTAB_W; r++) {
        const unsigned nxt = r + 1 < CTAB_W ? v[r + 1] : nx0;
        if (((v[r] ^ nxt) & ~CTAB_IDXM) == 0u && 32 * t + r + 1 < n) marks |= 1u << r;
    }
    unsigned short *out = reinterpret_cast<unsigned short *>(region);
    {
        unsigned *o32 = region + 16 * t; /* two positions per word */
#pragma unroll
        for (int r = 0; r < CTAB_W; r += 2) o32[r / 2] = (v[r] & CTAB_IDXM) | ((v[r + 1] & CTAB_IDXM) << 16);
    }
    __syncthreads();
    if (__syncthreads_or(marks != 0u)) {
        for (int pass = 0;; pass++) {
            int swapped = 0;
#pragma unroll
            for (int parity = 0; parity < 2; parity++) {
                unsigned m = marks & (parity ? 0xaaaaaaaau : 0x55555555u);
                while (m) {
                    const int k = 32 * t + (__ffs(m) - 1);
                    m &= m - 1;
                    const unsigned short ia = out[k], ib = out[k + 1];
                    const double a = s_val[ia], b = s_val[ib];
                    if ((DESC ? a < b : a > b) || (a == b && ia > ib)) { out[k] = ib; out[k + 1] = ia; swapped = 1; }
                }
                __syncthreads();
            }
            if (!__syncthreads_or(swapped)) break;
            if (pass >= CTAB_MAXPASS) { /* a long run of near-equal values: not worth bubbling */
                if (t == 0 && full_rows) atomicAdd(full_rows, 1);
                return nullptr;
            }
        }
        unsigned m = marks;
        while (m) { /* tie marks; readers of a neighbour's entry mask them off */
            const int k = 32 * t + (__ffs(m) - 1);
            m &= m - 1;
            const unsigned short ia = out[k] & 0x7fffu, ib = out[k + 1] & 0x7fffu;
            if (s_val[ia] == s_val[ib]) out[k] = (unsigned short)(ia | 0x8000u);
        }
        __syncthreads();
    }
    return out;
}

} // namespace epx
