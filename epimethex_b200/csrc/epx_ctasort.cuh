/*
 * epx_ctasort.cuh — one CTA (8 warps) orders one long row (1024 < n <= 16384 values, e.g. the 9,000-sample
 * pan-cancer shape) held in shared memory: every warp sorts 512-value chunks with the register bitonic sort of
 * epx_warpsort.cuh, then log2(chunks) merge passes (merge path: each thread produces a contiguous slice of the
 * output after a binary search for its split point) combine them. Exact and stable: equal values keep ascending
 * sample index, because chunks are contiguous index ranges and merges prefer the left run on ties.
 */
#pragma once
#include "epx_warpsort.cuh"

namespace epx {

constexpr int CTASORT_THREADS = 256;               /* launch bound, see ctasort_threads() */
constexpr int CTASORT_EPS = 16;                    /* values per lane in the register sort of a chunk */
constexpr int CTASORT_CHUNK = 32 * CTASORT_EPS;    /* 512: the 1,024-value network costs 3.4x the 512-value one, a merge pass far less */

/* Threads of the sorting CTA. Measured on the 9,000-sample shape (18 chunks of 512 values): 8 warps 7.36 ms, 9 warps
 * (two full rounds of chunk sorts instead of 8 + 8 + 2) 7.94 ms, 12 warps 7.40 ms -- the kernel is bound by
 * shared-memory traffic in the merge passes, not by the number of rounds, so the smallest CTA stays. */
inline int ctasort_threads(int) { return 256; }

/* entries of each of the two index buffers: a whole number of chunks, so that the second buffer can double as
 * the warp sorts' scratch (one chunk of entries per chunk being sorted) while the first one is being filled */
__host__ __device__ inline int ctasort_entries(int n) { return (n + CTASORT_CHUNK - 1) / CTASORT_CHUNK * CTASORT_CHUNK; }

/* shared memory needed besides the row itself */
__host__ __device__ inline size_t ctasort_smem(int n) { return (size_t)ctasort_entries(n) * 2 * 2; }

template <bool DESC> __device__ __forceinline__ bool comes_before(double a, double b) { return DESC ? a > b : a < b; }

/* merge passes: s_a holds sorted runs of run_len entries (the last one may be shorter); every pass merges adjacent
 * pairs of runs (merge path: each thread produces a contiguous slice of the output after a binary search for its split
 * point; the left run wins ties). Ends on a __syncthreads(); returns the buffer that holds the result. */
/* MARK: the LAST pass also sets bit 15 of every output whose successor holds an equal value (the tie mark of the order
 * rows). The mark of an output is known one step later (the merge state after an output names the next one, also
 * across the slices of two threads), so the stores of that pass are delayed by one step; this replaces a separate sweep
 * of two 8-byte value gathers per element. Needs at least one pass (n > run_len). */
template <bool DESC, bool MARK = false>
__device__ unsigned short *cta_merge_runs(const double *s_val, const int n, const int run_len, unsigned short *s_a,
                                          unsigned short *s_b)
{
    unsigned short *src = s_a, *dst = s_b;
    /* entries per thread and pass; never a multiple of 8: the lanes of a warp start `per` 16-bit entries apart, and a
     * multiple of 8 puts them on at most eight banks (n = 9,000 on 288 threads: per = 32, two banks, measured 26 %
     * slower than per = 36 on 256 threads; 37 instead of 36 measured 6 % slower again, so only the bad strides move) */
    int per = (n + blockDim.x - 1) / blockDim.x;
    per += (per & 7) == 0 ? 1 : 0;
    for (int len = run_len; len < n; len <<= 1) {
        int o = min((int)threadIdx.x * per, n);
        const int o_end = min(o + per, n);
        while (o < o_end) {
            const int pbase = (o / (2 * len)) * (2 * len);
            const int la = min(len, n - pbase), lb = min(len, n - pbase - la);
            const unsigned short *A = src + pbase, *B = src + pbase + la;
            const int seg_end = min(o_end, pbase + la + lb);
            const int d = o - pbase;
            /* merge path: i = how many of the first d outputs come from A (A wins ties) */
            int lo = max(0, d - lb), hi = min(d, la);
            while (lo < hi) {
                const int mid = (lo + hi) >> 1;
                if (!comes_before<DESC>(s_val[B[d - 1 - mid]], s_val[A[mid]])) lo = mid + 1; else hi = mid;
            }
            /* branch-free sequential merge: pa/pb = next unread entry of A/B, (xa, va)/(xb, vb) = the current heads */
            int pa = lo, pb = d - lo;
            /* heads clamped into the pair of runs; an absent B (odd run out) aliases the last entry of A */
            const int lastA = la - 1, lastB = la + lb - 1;
            unsigned short xa = A[min(pa, lastA)], xb = A[min(la + pb, lastB)];
            double va = s_val[xa], vb = s_val[xb];
            const bool mark = MARK && 2 * len >= n; /* one pair of runs is left: this thread has one segment */
            unsigned short prev_x = 0;
            double prev_v = 0.0;
            for (int k = o; k < seg_end; k++) {
                const bool take_a = pb >= lb || (pa < la && !comes_before<DESC>(vb, va));
                if (mark) {
                    const unsigned short cur_x = take_a ? xa : xb;
                    const double cur_v = take_a ? va : vb;
                    if (k > o) dst[k - 1] = (unsigned short)(prev_x | (cur_v == prev_v ? 0x8000u : 0u));
                    prev_x = cur_x; prev_v = cur_v;
                } else dst[k] = take_a ? xa : xb;
                pa += take_a ? 1 : 0;
                pb += take_a ? 0 : 1;
                const unsigned short nx = A[take_a ? min(pa, lastA) : min(la + pb, lastB)];
                const double nv = s_val[nx];
                xa = take_a ? nx : xa; va = take_a ? nv : va;
                xb = take_a ? xb : nx; vb = take_a ? vb : nv;
            }
            if (mark && seg_end > o) {
                const bool more = pa < la || pb < lb; /* the heads now name output seg_end, if there is one */
                const bool take_a = pb >= lb || (pa < la && !comes_before<DESC>(vb, va));
                dst[seg_end - 1] = (unsigned short)(prev_x | ((more && (take_a ? va : vb) == prev_v) ? 0x8000u : 0u));
            }
            o = seg_end;
        }
        __syncthreads();
        unsigned short *t = src; src = dst; dst = t;
    }
    return src;
}

/* s_val: the row (n doubles, shared); s_a, s_b: ctasort_entries(n) uint16 each. All threads of the CTA must call.
 * Returns the buffer (s_a or s_b) that holds the sample indices in sorted order. */
template <bool DESC, bool MARK = false>
__device__ unsigned short *cta_sort_row(const double *s_val, const int n, unsigned short *s_a, unsigned short *s_b,
                                        const unsigned one)
{
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarps = blockDim.x >> 5;
    const int nchunks = (n + CTASORT_CHUNK - 1) / CTASORT_CHUNK;
    for (int c = warp; c < nchunks; c += nwarps) {
        const int base = c * CTASORT_CHUNK;
        const int nc = min(CTASORT_CHUNK, n - base);
        unsigned khi[CTASORT_EPS], klo[CTASORT_EPS];
#pragma unroll
        for (int e = 0; e < CTASORT_EPS; e++) {
            key_halves(s_val[base + min(lane + 32 * e, nc - 1)], khi[e], klo[e]);
            if (DESC) { khi[e] = ~khi[e]; klo[e] = ~klo[e]; }
        }
        unsigned short out[CTASORT_EPS];
        WarpSort<CTASORT_EPS>::template order<DESC>(khi, klo, nc, lane, s_b + c * CTASORT_CHUNK,
                                           [&](int i) { return s_val[base + i]; }, one, out);
#pragma unroll
        for (int r = 0; r < CTASORT_EPS; r++) {
            const int k = lane * CTASORT_EPS + r;
            if (k < nc) s_a[base + k] = (unsigned short)(base + (out[r] & 0x7fff));
        }
    }
    __syncthreads();

    return cta_merge_runs<DESC, MARK>(s_val, n, CTASORT_CHUNK, s_a, s_b);
}

} // namespace epx
