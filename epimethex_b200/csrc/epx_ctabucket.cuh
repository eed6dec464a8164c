/*
 * epx_ctabucket.cuh — one CTA orders one long row (1024 < n <= 16384 values, e.g. the 9,000-sample pan-cancer
 * shape) by BUCKET PLACEMENT: the CTA-wide form of WarpSort::order_bucket (epx_warpsort.cuh).
 *
 *   1. range of the row (exact minimum and maximum, from the order-preserving keys of epx_device.cuh);
 *   2. every value is counted into bucket Q >> 15, Q = round((x - first) * scale) a 29-bit fixed-point position in
 *      the row's range, ~1.03 buckets per slot, shared-memory atomics (ATOMS ~2.7 cycles per warp instruction when
 *      the addresses differ, profiles/micro/hist_micro.cu);
 *   3. exclusive scan of the counters; every value draws the next free slot of its bucket and drops the word
 *      [pad flag | 25 leading bits of Q | slot mod 64] there, and its sample index into a slot table: the row is
 *      ordered but for the values that share a bucket. The staged row is dead by then (every Q is in a register), so
 *      slots and slot table take its place in shared memory;
 *   4. those are settled by the 32-value network inside each thread (WarpSort<32>::sort_thread) and one merge of
 *      the windows across the thread boundaries (runs of up to 17 values are put right);
 *   5. whether that was enough is CHECKED (one comparison per neighbour pair, __syncthreads_or); a row that fails,
 *      or whose fullest bucket holds more than 31 values, is left to the caller (merge sort of epx_ctasort.cuh), so
 *      the result never depends on the data's shape;
 *   6. a value has moved by fewer than 32 slots (it stays inside its bucket's run), so the 6-bit slot number in its
 *      word names its slot of arrival, and the slot table its sample;
 *   7. neighbours with equal 25-bit keys (exact ties; values closer than 2^-25 of ~1000 buckets: a handful per row)
 *      are settled on the full values, read from global memory: odd-even transposition over exactly those pairs,
 *      then the tie marks. (With the 14-bit sample index in the word there was room for 17 key bits only, a fifth
 *      of the neighbour pairs at n = 9,000 had to be settled, and the row had to stay in shared memory for it.)
 *
 * Values equal to the row's minimum or maximum (beta values clipped at a detection limit, zero counts: piles of
 * hundreds of equal values) do not draw slots in arrival order -- CTA-wide atomics have none -- but by sample
 * index: a membership bitmap per pile, prefix pop-counts, rank = members with a smaller index. They are ties by
 * construction and take their tie marks without a look at the values.
 *
 * Replaces: the per-call sort() inside median()/ks.test() (R/Functions.R:85-132) and order() of
 * R/EpimethexAnalysis.R:117-122 for long rows; same output as cta_sort_row (epx_ctasort.cuh).
 */
#pragma once
#include "epx_ctasort.cuh"

namespace epx {

constexpr int CTAB_W = 32;       /* slots per thread */
constexpr int CTAB_STRIDE = 33;  /* words between the slot blocks of two threads (odd: conflict-free 32-bit accesses) */
constexpr int CTAB_LIDB = 6;     /* bits of the slot number kept in the word */
constexpr unsigned CTAB_LIDM = (1u << CTAB_LIDB) - 1u;
constexpr int CTAB_SH = 15;      /* Q >> 15 = bucket; Q < 16896 << 15 < 2^30: bits 30 and 31 are free for the pile flags */
constexpr int CTAB_MAXPASS = 48; /* transposition rounds before the row is handed to the merge sort instead */

__host__ __device__ inline int ctab_threads(int n) { return 32 * ((n + 1023) / 1024); }
/* bytes of the row area: the staged row first, slots (33 words per thread) + slot table (32 uint16 per thread) later */
__host__ __device__ inline size_t ctab_row_bytes(int n)
{
    const size_t T = (size_t)ctab_threads(n), a = (size_t)((n + 1) & ~1) * 8, b = T * CTAB_STRIDE * 4 + T * CTAB_W * 2;
    return (a > b ? a : b + 15) & ~(size_t)15;
}
/* shared memory besides the row area: counters, later the order row (33 words per thread; also the merge sort's index
 * buffers when a row falls back) | two pile bitmaps | their prefix counts | scratch of the reductions */
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
 * s_val: the row staged in shared memory, in a row area of ctab_row_bytes(n) bytes -- OVERWRITTEN (slots, slot table);
 * gval: the same row in global memory (read for the few neighbours that have to be settled on their values);
 * region: ctab_smem(n) bytes. All ctab_threads(n) threads of the CTA must call.
 * Returns the order row (n uint16: sample index of the k-th value, bit 15 = "the next value is equal"), which lies in
 * `region`, or NULL when the row has to be ordered by other means (the row area is garbage then); ends on a barrier.
 */
template <bool DESC>
__device__ const unsigned short *cta_bucket_order(double *s_val, const double *__restrict__ gval, const int n, unsigned *region,
                                                  const unsigned one, int32_t *full_rows)
{
    const int T = (int)blockDim.x, t = (int)threadIdx.x, lane = t & 31, warp = t >> 5, nwarps = T >> 5;
    const int NB = CTAB_STRIDE * T;
    unsigned *bm = region + (size_t)CTAB_STRIDE * T; /* [2][T] membership of the two piles, one bit per sample */
    unsigned *pre = bm + 2 * T;                      /* [T] members in front of each bitmap word: first pile | last pile << 16 */
    unsigned *red = pre + T;                         /* 64 words */
    unsigned *slots = reinterpret_cast<unsigned *>(s_val);                           /* slot k at word k + k / 32 */
    unsigned short *idxtab = reinterpret_cast<unsigned short *>(slots + (size_t)CTAB_STRIDE * T); /* sample of the value that arrived in slot k */
    const unsigned QMAX = (unsigned)NB << CTAB_SH;
    int qw = 0; /* Q >> qw: the 25 bits of Q kept in the word; every such word lies below the last pile's 0x7fff0000 + rank */
    while ((QMAX >> qw) >= (0x7fff0000u >> CTAB_LIDB)) qw++;

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
    /* the row's first value sits near the top of bucket 0 and the last one near the bottom of the last bucket */
    double scale = (double)(((unsigned)(NB - 2) << CTAB_SH) + 128u) / span;
    if (!(fabs(scale) < 1e300)) scale = 0.0; /* constant row, or a span beyond the exponent range */
    const double c0 = fma(-first, scale, 6755399441055744.0 + (double)((1u << CTAB_SH) - 64u));

    /* ---- counting pass: the two piles into their bitmaps, everything else into its bucket's counter ---- */
    unsigned qf = (unsigned)__double2loint(fma(first, scale, c0)), ql = (unsigned)__double2loint(fma(last, scale, c0));
    qf = min(qf, QMAX - 1u); ql = min(ql, QMAX - 1u);
    unsigned Q[CTAB_W]; /* per element: Q, bit 31 / 30: member of the first / last pile; 0xffffffff: no such element */
    /* n > 32 T - 1024 >= 16 T: the elements e < 16 of every thread exist */
#pragma unroll
    for (int e = 0; e < CTAB_W; e++) {
        const int i = t + T * e;
        Q[e] = 0xffffffffu;
        if (e < CTAB_W / 2 || i < n) {
            const double x = s_val[i];
            unsigned q = (unsigned)__double2loint(fma(x, scale, c0));
            q = min(q, QMAX - 1u); /* memory safety only: a q out of range means the row fails the check below */
            bool pile = false;
            if (q == qf || q == ql) { /* equal Q is necessary for membership; the values decide */
                if (x == first) { atomicOr(bm + (i >> 5), 1u << (i & 31)); q |= 0x80000000u; pile = true; }
                else if (x == last) { atomicOr(bm + T + (i >> 5), 1u << (i & 31)); q |= 0x40000000u; pile = true; }
            }
            if (!pile) atomicAdd(region + (q >> CTAB_SH), 1u);
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
    bool too_full = false;
    {   /* the counters are read twice rather than kept: Q[] is live across this scan */
        unsigned *mine = region + CTAB_STRIDE * t;
        unsigned tot = 0, mx = 0, dummy;
#pragma unroll
        for (int i = 0; i < CTAB_STRIDE; i++) { const unsigned c = mine[i]; tot += c; mx = max(mx, c); }
        too_full = mx >= 32u; /* a run of 32 or more: its values could move further than the word's slot number can tell */
        unsigned run = pile0 + ctab_scan(tot, red, &dummy); /* the first pile takes the first pile0 slots */
#pragma unroll
        for (int i = 0; i < CTAB_STRIDE; i++) {
            const unsigned cc = mine[i];
            mine[i] = run;
            run += cc;
        }
    }
    if (__syncthreads_or(too_full)) { /* the barrier also means: every Q is in its register, the staged row is dead */
        if (t == 0 && full_rows) atomicAdd(full_rows, 1);
        return nullptr;
    }
    /* ---- placement: every value draws its slot and moves there ---- */
#pragma unroll
    for (int e = 0; e < CTAB_W; e++) {
        const int i = t + T * e;
        const unsigned q = Q[e];
        if (e < CTAB_W / 2 || i < n) {
            const unsigned qq = q & 0x3fffffffu;
            unsigned pos, word;
            if (q & 0xc0000000u) {
                const bool lastp = !(q & 0x80000000u);
                const unsigned w = bm[(lastp ? T : 0) + (i >> 5)], pr = pre[i >> 5];
                const unsigned rank = (lastp ? (pr >> 16) : (pr & 0xffffu)) + (unsigned)__popc(w & ((1u << (i & 31)) - 1u));
                /* no value lies below the first pile, none above the last; their words ascend with the slot and lie
                 * below / above every other word of the row */
                pos = lastp ? (unsigned)n - pile1 + rank : rank;
                word = lastp ? 0x7fff0000u + rank : rank;
            } else {
                pos = atomicAdd(region + (qq >> CTAB_SH), 1u);
                word = ((qq >> qw) << CTAB_LIDB) | (pos & CTAB_LIDM);
            }
            slots[pos + (pos >> 5)] = word;
            idxtab[pos] = (unsigned short)i;
        }
    }
    for (int k = n + t; k < CTAB_W * T; k += T) slots[k + (k >> 5)] = 0x80000000u | (unsigned)k; /* pads behind the row */
    __syncthreads();
    /* ---- short network: 32 values inside each thread, then the windows across the thread boundaries ---- */
    const unsigned minus1 = 0u - one;
    unsigned v[CTAB_W];
    unsigned *mine = slots + CTAB_STRIDE * t;
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
    if (__syncthreads_or(bad)) {
        if (t == 0 && full_rows) atomicAdd(full_rows, 1);
        return nullptr;
    }
    /* ---- the order row: the sample of every position; tie marks of the piles; neighbours with equal keys ---- */
    unsigned short *out = reinterpret_cast<unsigned short *>(region); /* the counters are no longer needed */
    unsigned marks = 0; /* bit r: positions 32t + r and 32t + r + 1 hold equal keys and have to be settled on their values */
    const int lo_end = (int)pile0, hi_begin = n - (int)pile1;
    {
        unsigned *o32 = region + 16 * t; /* two positions per word */
        unsigned pend = 0;
#pragma unroll
        for (int r = 0; r < CTAB_W; r++) {
            const int k = 32 * t + r;
            unsigned o = 0;
            if (k < n) {
                const bool in_pile = k < lo_end || k >= hi_begin;
                /* the slot of arrival: within 31 of k, and its low six bits are in the word */
                const int p0 = in_pile ? k : k - 32 + (int)((v[r] - (unsigned)(k - 32)) & CTAB_LIDM);
                o = idxtab[p0];
                const unsigned nxt = r + 1 < CTAB_W ? v[r + 1] : nx0;
                if (in_pile) { if (k + 1 != lo_end && k + 1 != n) o |= 0x8000u; } /* equal by construction, in sample order already */
                else if (((v[r] ^ nxt) >> CTAB_LIDB) == 0u && k + 1 < hi_begin) marks |= 1u << r;
            }
            if (r & 1) o32[r / 2] = pend | (o << 16); else pend = o;
        }
    }
    bm[t] = marks; /* for the right-hand neighbour (the bitmaps are free by now) */
    __syncthreads();
    if (__syncthreads_or(marks != 0u)) {
        auto settle = [&](unsigned pairs) { /* one odd-even round over the given pairs of this thread; 1 if anything moved */
            int moved = 0;
            while (pairs) {
                const int k = 32 * t + (__ffs(pairs) - 1);
                pairs &= pairs - 1;
                const unsigned short ia = out[k], ib = out[k + 1];
                const double a = __ldg(gval + ia), b = __ldg(gval + ib);
                if ((DESC ? a < b : a > b) || (a == b && ia > ib)) { out[k] = ib; out[k + 1] = ia; moved = 1; }
            }
            return moved;
        };
        /* runs of marked pairs that stay inside this thread's 32 positions (not connected to pair 31, nor to the left
         * neighbour's pair 31 through position 0) are settled by the thread alone: nobody else touches their entries */
        const unsigned left_in = t > 0 ? (bm[t - 1] >> 31) : 0u;
        unsigned lead = 0, trail = 0;
        if (left_in) lead = marks & ~(marks + 1u);                                     /* marked pairs 0, 1, .. up to the first gap */
        if (marks >> 31) { const int z = __clz(~marks); trail = z >= 32 ? 0xffffffffu : ~(0xffffffffu >> z); } /* down from 31 */
        const unsigned shared_pairs = lead | trail, own_pairs = marks & ~shared_pairs;
        if (own_pairs) {
            for (int guard = 0; guard < 64; guard++) {
                const int a = settle(own_pairs & 0x55555555u), b = settle(own_pairs & 0xaaaaaaaau);
                if (!(a | b)) break;
            }
        }
        if (__syncthreads_or(shared_pairs != 0u)) { /* runs across block boundaries: CTA-wide rounds */
            for (int pass = 0;; pass++) {
                int swapped = settle(shared_pairs & 0x55555555u);
                __syncthreads();
                swapped |= settle(shared_pairs & 0xaaaaaaaau);
                __syncthreads();
                if (!__syncthreads_or(swapped)) break;
                if (pass >= CTAB_MAXPASS) { /* a long run of near-equal values: not worth bubbling */
                    if (t == 0 && full_rows) atomicAdd(full_rows, 1);
                    return nullptr;
                }
            }
        }
        unsigned m = marks;
        while (m) { /* tie marks; readers of a neighbour's entry mask them off */
            const int k = 32 * t + (__ffs(m) - 1);
            m &= m - 1;
            const unsigned short ia = out[k] & 0x7fffu, ib = out[k + 1] & 0x7fffu;
            if (__ldg(gval + ia) == __ldg(gval + ib)) out[k] = (unsigned short)(ia | 0x8000u);
        }
        __syncthreads();
    }
    return out;
}

} // namespace epx
