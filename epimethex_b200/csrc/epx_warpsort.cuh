/*
 * epx_warpsort.cuh — one warp orders one row (n <= 32*EPS values) entirely in registers and its own shared memory.
 *
 * Two paths, same result (ascending by (value, sample index), tie marks):
 *
 * order_bucket (round 2; rows of 257..512 values, and the two halves of 513..1024-value rows): every value is dropped
 * near its final position by a counting pass over a linear map of the row's range (shared-memory atomics), a
 * 16-value network inside each thread and one merge across the thread boundaries settle what shares a bucket, a
 * check decides whether that was enough, and a row that fails takes the full network below. See the comment at
 * order_bucket.
 *
 * order (round 1; shorter rows, chunks of the merge sort, the fallback): the row's doubles are mapped to
 * order-preserving 64-bit keys, the bits all keys share are stripped, and the leading (31 - IDXB) remaining bits
 * ("coarse key") are packed above the sample index, below a pad flag, into ONE 32-bit word per element.  A
 * direction-free bitonic network then sorts the words with one VIMNMX per compare-exchange half: strides below EPS
 * stay inside a thread, larger strides are one SHFL each.
 *
 * Either way only rows in which two neighbours of the sorted order share a coarse key (exact ties, or values closer
 * than ~2^-22 of the row's range) take the exact path of finish(): an odd-even transposition on the full values,
 * restricted to those neighbours, which also yields the tie marks.  The result is exact for any input.
 */
#pragma once
#include "epx_device.cuh"

namespace epx {

/* Compare-exchange with the maximum computed on the FMA pipe: max = a + b - min as two IMADs whose multipliers
 * (+1, -1) are run-time values, so ptxas cannot fold them back into integer-ALU adds. The sort kernels are bound
 * by the integer ALU pipe (VIMNMX); moving ~40 % of the max() halves there balances the two pipes. */
__device__ __forceinline__ void ce_fma(unsigned &a, unsigned &b, const unsigned one, const unsigned minus1)
{
    const unsigned mn = min(a, b);
    unsigned s, mx;
    asm("mad.lo.u32 %0, %1, %2, %3;" : "=r"(s) : "r"(a), "r"(one), "r"(b));
    asm("mad.lo.u32 %0, %1, %2, %3;" : "=r"(mx) : "r"(mn), "r"(minus1), "r"(s));
    a = mn;
    b = mx;
}

/* Cross-lane half of a compare-exchange, in place: v = upper ? max(v, t) : min(v, t) as two predicated VIMNMX
 * (the C form compiles to a min into a temporary, a predicated max over it and, inside loops, a move back). */
__device__ __forceinline__ void ce_half(unsigned &v, const unsigned t, const unsigned upper)
{
    asm("{\n\t.reg .pred p;\n\tsetp.ne.u32 p, %2, 0;\n\t@!p min.u32 %0, %0, %1;\n\t@p max.u32 %0, %0, %1;\n\t}" : "+r"(v) : "r"(t), "r"(upper));
}

/* shared-memory counter bumps under a predicate instead of a branch (the C form costs a BSSY/BRA/BSYNC per element) */
__device__ __forceinline__ void smem_inc_if(const uint32_t addr, const bool p)
{
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.u32 p, %1, 0;\n\t@p red.shared.add.u32 [%0], 1;\n\t}" ::"r"(addr), "r"((unsigned)p) : "memory");
}
__device__ __forceinline__ unsigned smem_take_if(const uint32_t addr, const bool p, unsigned otherwise)
{
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.u32 p, %2, 0;\n\t@p atom.shared.add.u32 %0, [%1], 1;\n\t}" : "+r"(otherwise) : "r"(addr), "r"((unsigned)p) : "memory");
    return otherwise;
}

template <int EPS> struct WarpSort {
    static constexpr int N = 32 * EPS;
    static constexpr int IDXB = (EPS == 1 ? 5 : EPS == 2 ? 6 : EPS == 4 ? 7 : EPS == 8 ? 8 : EPS == 16 ? 9 : 10);
    static constexpr unsigned IDXM = (1u << IDXB) - 1u;
    static constexpr int FMA_CE_OF_5 = 4; /* of every 5 in-thread compare-exchanges, this many use ce_fma */
    static_assert(EPS == 1 || EPS == 2 || EPS == 4 || EPS == 8 || EPS == 16 || EPS == 32, "EPS must be a power of two <= 32");

    /* in-thread half-cleaners with strides EPS/2 .. 1 (static register indices) */
    __device__ __forceinline__ static void thread_tail(unsigned (&v)[EPS], int from, const unsigned one, const unsigned minus1)
    {
#pragma unroll
        for (int j = EPS / 2; j >= 1; j >>= 1) {
            if (j <= from) {
#pragma unroll
                for (int r = 0; r < EPS; r++) {
                    if ((r & j) == 0) {
                        if (FMA_CE_OF_5 > 0 && (r % 5) < FMA_CE_OF_5) ce_fma(v[r], v[r | j], one, minus1);
                        else {
                            const unsigned a = v[r], b = v[r | j];
                            v[r] = min(a, b);
                            v[r | j] = max(a, b);
                        }
                    }
                }
            }
        }
    }

    /* ascending sort of the 32*EPS words held as v[r] by 32 lanes; on return position k = lane*EPS + r.
     * Direction-free bitonic network: every merge phase starts with a mirror step (i <-> i ^ (k-1), the lower
     * index keeps the minimum) followed by half-cleaners (i <-> i ^ j). Phases that fit a thread are unrolled;
     * the five cross-lane phases run as a loop (same code for every phase: the body stays in the instruction
     * cache) with the lane masks as run-time values. */
    /* the stages that stay inside a thread: on return every lane's EPS words ascend */
    __device__ __forceinline__ static void sort_thread(unsigned (&v)[EPS], const unsigned one, const unsigned minus1)
    {
#pragma unroll
        for (int k = 2; k <= EPS; k <<= 1) {
#pragma unroll
            for (int r = 0; r < EPS; r++) {
                const int q = r ^ (k - 1);
                if (r < q) {
                    const unsigned a = v[r], b = v[q];
                    v[r] = min(a, b);
                    v[q] = max(a, b);
                }
            }
            thread_tail(v, k / 4, one, minus1);
        }
    }

    /* 16 words inside a thread by the 60-comparator, 10-layer network (optimal size for 16 inputs; the bitonic stages above
     * spend 80), min/max only. Checked on all 2^16 zero-one inputs (profiles/README.md, round 2). Used where the 16-value
     * sort is the whole network (order_bucket); sort() keeps the bitonic stages its cross-lane merges continue from. */
    __device__ __forceinline__ static void sort16_thread(unsigned (&v)[EPS])
    {
        static_assert(EPS == 16, "the 60-comparator network is for 16 inputs");
#define EPX_CE(a, b) { const unsigned lo_ = min(v[a], v[b]), hi_ = max(v[a], v[b]); v[a] = lo_; v[b] = hi_; }
        EPX_CE(0, 13) EPX_CE(1, 12) EPX_CE(2, 15) EPX_CE(3, 14) EPX_CE(4, 8) EPX_CE(5, 6) EPX_CE(7, 11) EPX_CE(9, 10)
        EPX_CE(0, 5) EPX_CE(1, 7) EPX_CE(2, 9) EPX_CE(3, 4) EPX_CE(6, 13) EPX_CE(8, 14) EPX_CE(10, 15) EPX_CE(11, 12)
        EPX_CE(0, 1) EPX_CE(2, 3) EPX_CE(4, 5) EPX_CE(6, 8) EPX_CE(7, 9) EPX_CE(10, 11) EPX_CE(12, 13) EPX_CE(14, 15)
        EPX_CE(0, 2) EPX_CE(1, 3) EPX_CE(4, 10) EPX_CE(5, 11) EPX_CE(6, 7) EPX_CE(8, 9) EPX_CE(12, 14) EPX_CE(13, 15)
        EPX_CE(1, 2) EPX_CE(3, 12) EPX_CE(4, 6) EPX_CE(5, 7) EPX_CE(8, 10) EPX_CE(9, 11) EPX_CE(13, 14)
        EPX_CE(1, 4) EPX_CE(2, 6) EPX_CE(5, 8) EPX_CE(7, 10) EPX_CE(9, 13) EPX_CE(11, 14)
        EPX_CE(2, 4) EPX_CE(3, 6) EPX_CE(9, 12) EPX_CE(11, 13)
        EPX_CE(3, 5) EPX_CE(6, 8) EPX_CE(7, 9) EPX_CE(10, 12)
        EPX_CE(3, 4) EPX_CE(5, 6) EPX_CE(7, 8) EPX_CE(9, 10) EPX_CE(11, 12)
        EPX_CE(6, 7) EPX_CE(8, 9)
#undef EPX_CE
    }

    __device__ __forceinline__ static void sort(unsigned (&v)[EPS], const int lane, const unsigned one, const unsigned minus1)
    {
        sort_thread(v, one, minus1);
#pragma unroll 1
        for (int q = 1; q < 32; q <<= 1) { /* phase k = 2*EPS*q */
            {
                const int lm = 2 * q - 1;
                const unsigned upper = (unsigned)(lane & q);
                unsigned t[EPS];
#pragma unroll
                for (int r = 0; r < EPS; r++) t[r] = __shfl_xor_sync(FULL, v[r], lm);
#pragma unroll
                for (int r = 0; r < EPS; r++) ce_half(v[r], t[r ^ (EPS - 1)], upper);
            }
            /* the up to four cross-lane half-cleaners as straight-line code behind warp-uniform tests: only the
             * phase loop has a back edge (ptxas pays ~16 register moves per back edge of a loop over v[]) */
#pragma unroll
            for (int j = 8; j >= 1; j >>= 1) {
                if (j <= (q >> 1)) {
                    const unsigned upper = (unsigned)(lane & j);
#pragma unroll
                    for (int r = 0; r < EPS; r++) {
                        const unsigned t = __shfl_xor_sync(FULL, v[r], j);
                        ce_half(v[r], t, upper);
                    }
                }
            }
            thread_tail(v, EPS / 2, one, minus1);
        }
    }

    /*
     * Orders one row. (khi[e], klo[e]) = key_halves() of the element with sample index lane + 32*e; slots with an
     * index >= n may hold ANY real key of the row (callers clamp the load index or substitute element 0) so that the
     * common-prefix scan needs no validity test. Ascending by (key, index); pass complemented halves and DESC = true
     * for a descending order. Returns in out[r] the sample index at sorted position lane*EPS + r, with bit 15 set when
     * the NEXT position holds an equal value (tie mark); positions >= n hold n. s_ord: N uint16 of per-warp shared
     * scratch; value(idx) returns the double of a sample (only called for positions whose neighbour shares their
     * coarse key).
     *
     * Word layout: [pad flag | CB coarse-key bits | IDXB index bits]. Pads carry the flag, a running number in the
     * coarse field and n in the index field: they sort behind every real word, no two words of a row are ever equal,
     * and two neighbours of the sorted order share a coarse key exactly when their words differ by less than 2^IDXB
     * ... or are one coarse step apart with a falling index, which only sends a row down the exact path for nothing.
     * So the common case (no ties, no near-ties) is decided by ONE subtraction and half a min3 per element.
     */
    template <bool DESC, typename ValueOf>
    __device__ __forceinline__ static void order(const unsigned (&khi)[EPS], const unsigned (&klo)[EPS], const int n,
                                                 const int lane, unsigned short *s_ord, ValueOf value, const unsigned one,
                                                 unsigned short (&out)[EPS])
    {
        order_rt(khi, klo, n, lane, s_ord, value, DESC, one, out);
    }

    /* the same with the direction as a (warp-uniform) run-time value: one copy of the network for kernels that order
     * rows both ways (k_probe_rank_w: expression rows descending, methylation rows ascending) */
    template <typename ValueOf>
    __device__ __forceinline__ static void order_rt(const unsigned (&khi)[EPS], const unsigned (&klo)[EPS], const int n,
                                                    const int lane, unsigned short *s_ord, ValueOf value, const bool DESC,
                                                    const unsigned one, unsigned short (&out)[EPS])
    {
        constexpr unsigned CMASK = 0x7fffffffu & ~IDXM;
        /* bits shared by all keys: OR of (key ^ key0) tells where they start to differ */
        const unsigned h0 = __shfl_sync(FULL, khi[0], 0), l0 = __shfl_sync(FULL, klo[0], 0);
        unsigned dhi = 0, dlo = 0;
#pragma unroll
        for (int e = 0; e < EPS; e++) {
            dhi |= khi[e] ^ h0;
            dlo |= klo[e] ^ l0;
        }
        dhi = __reduce_or_sync(FULL, dhi);
        dlo = __reduce_or_sync(FULL, dlo);
        const int lz = dhi ? __clz(dhi) : 32 + (dlo ? __clz(dlo) : 32); /* 64: all keys equal */
        unsigned v[EPS];
        /* the coarse key = the CB bits from the first differing bit on, placed below the pad flag: a funnel shift by
         * lz - 1 leaves one (common) bit on top, which the mask clears */
        if (lz >= 1 && lz < 32) {
#pragma unroll
            for (int e = 0; e < EPS; e++) v[e] = (__funnelshift_l(klo[e], khi[e], lz - 1) & CMASK) | (unsigned)(lane + 32 * e);
        } else if (lz == 0) {
#pragma unroll
            for (int e = 0; e < EPS; e++) v[e] = ((khi[e] >> 1) & CMASK) | (unsigned)(lane + 32 * e);
        } else {
            const int sh = lz - 33; /* -1 (lz = 32: the low word from its top bit) .. 31 */
#pragma unroll
            for (int e = 0; e < EPS; e++) {
                const unsigned w = sh < 0 ? (klo[e] >> 1) : (sh < 32 ? (klo[e] << sh) : 0u);
                v[e] = (w & CMASK) | (unsigned)(lane + 32 * e);
            }
        }
        /* slots past the row: normally only the upper half of the slots can be (a row uses the smallest instantiation
         * that holds it); short chunks of the CTA-level sorts take the warp-uniform branch */
#pragma unroll
        for (int e = EPS / 2; e < EPS; e++) {
            const int s = lane + 32 * e;
            if (s >= n) v[e] = 0x80000000u | ((unsigned)(s - n) << IDXB) | (unsigned)n;
        }
        if (n <= 16 * EPS) {
#pragma unroll
            for (int e = 0; e < (EPS + 1) / 2; e++) {
                const int s = lane + 32 * e;
                if (s >= n) v[e] = 0x80000000u | ((unsigned)(s - n) << IDXB) | (unsigned)n;
            }
        }
        sort(v, lane, one, 0u - one);
        finish(v, n, lane, s_ord, value, DESC, out);
    }

    /* ---------------------------------------------------------------------------------------------------------
     * Bucket placement in front of a SHORT network (order_bucket). The full network above costs 45 compare-exchange
     * phases per row; most of them only move values over long distances. Here every value is first dropped near its
     * final position by a counting pass in shared memory -- bucket = linear map of the value over the row's range,
     * 1.25 buckets per slot, counters bumped with shared-memory atomics (ATOMS: ~2.7 cycles per warp instruction on
     * B200 when the 32 addresses differ, profiles/micro/hist_micro.cu) -- so that what is left out of order are the
     * few values that share a bucket. Those are settled by the EPS-value network inside each thread plus one merge
     * across the thread boundaries (runs of up to EPS/2 + 1 values are put right). Whether that was enough is CHECKED
     * (one comparison per neighbour pair); a row that fails (a bucket that was too full: heavy ties in the middle of
     * the range, extreme outliers) goes through the full network, so the result never depends on the data's shape.
     * The word sorted is [pad flag | q | sample index] with q = round((x - first) * scale): monotone in x, ~2^21
     * levels over the row's range; neighbours with equal q are settled on the full values by finish() as before.
     * -------------------------------------------------------------------------------------------------------- */
    static constexpr int BK = EPS + 4;             /* counters (and padded slots) per lane: 16-byte chunks at an ODD chunk stride, so the
                                                    * LDS.128 / STS.128 of 8 consecutive lanes touch 8 different bank groups */
    static constexpr int NB = 32 * BK;             /* buckets */
    static constexpr int QSH = 31 - IDXB - (EPS == 8 ? 9 : EPS == 16 ? 10 : 11); /* q >> QSH = bucket; NB << QSH < 2^(31 - IDXB) */
    static constexpr unsigned QMAX = (unsigned)NB << QSH;
    static constexpr size_t BUCKET_SMEM = (size_t)(NB + 32) * 4; /* counters, later the padded slots, later the 16-bit scratch of
                                                                  * finish(); + one idle counter per lane for the pad slots */

    /* xof(e): value of sample lane + 32*e (any real value of the row for the slots past n), asked for twice per slot
     * (registers, or the staged row); x_done() is called once the last of them has been read (a caller with one row
     * stage refills it there). value(i): the value of sample i, for the few positions finish() has to settle.
     * 16*EPS < n <= 32*EPS (a row takes the smallest instantiation that holds it), so only the upper half of the slots
     * can be pads. */
    template <typename XOf, typename XDone, typename ValueOf>
    __device__ __forceinline__ static void order_bucket(XOf xof, XDone x_done, const int n, const int lane, unsigned *s_bkt,
                                                        ValueOf value, const bool DESC, const unsigned one,
                                                        unsigned short (&out)[EPS], int32_t *full_rows)
    {
        static_assert(EPS >= 8, "bucket placement is instantiated for 8, 16 and 32 values per lane");
        static_assert(((unsigned long long)NB << QSH) < (1ull << (31 - IDXB)), "q must fit below the pad flag");
        {   /* clean counters (the previous row left its slots, or finish() its scratch, in them) */
            uint4 *c4 = reinterpret_cast<uint4 *>(s_bkt + lane * BK);
#pragma unroll
            for (int i = 0; i < BK / 4; i++) c4[i] = make_uint4(0u, 0u, 0u, 0u);
        }
        /* range of the row from the high words of its keys (order-preserving transform of key_halves) */
        unsigned tmin = 0xffffffffu, tmax = 0u;
#pragma unroll
        for (int e = 0; e < EPS; e++) {
            const int h = __double2hiint(xof(e));
            const unsigned t = (unsigned)h ^ ((unsigned)(h >> 31) | 0x80000000u);
            tmin = min(tmin, t);
            tmax = max(tmax, t);
        }
        tmin = __reduce_min_sync(FULL, tmin);
        tmax = __reduce_max_sync(FULL, tmax);
        unsigned lmin = 0u, lmax = 0xffffffffu;
        if (tmax - tmin < 64u) { /* narrow row (rare): the low words matter */
            lmin = 0xffffffffu; lmax = 0u;
#pragma unroll
            for (int e = 0; e < EPS; e++) {
                unsigned hi, lo;
                key_halves(xof(e), hi, lo);
                if (hi <= tmin) lmin = min(lmin, lo);
                if (hi >= tmax) lmax = max(lmax, lo);
            }
            lmin = __reduce_min_sync(FULL, lmin);
            lmax = __reduce_max_sync(FULL, lmax);
        }
        double xlo, xhi;
        {
            const unsigned m0 = (tmin & 0x80000000u) ? 0u : 0xffffffffu, m1 = (tmax & 0x80000000u) ? 0u : 0xffffffffu;
            xlo = __hiloint2double((int)(tmin ^ (m0 | 0x80000000u)), (int)(lmin ^ m0));
            xhi = __hiloint2double((int)(tmax ^ (m1 | 0x80000000u)), (int)(lmax ^ m1));
        }
        /* q = round((x - first) * scale) + (2^QSH - 8): the row's first value lands at the TOP of bucket 0 and its last
         * value at the BOTTOM of bucket NB - 1, so a pile of values clipped to the row's minimum or maximum (beta values
         * at a detection limit, zero expression counts) has a bucket of its own: its members arrive in index order,
         * which is their final order, and the values next to the pile do not share its bucket */
        constexpr unsigned QOFF = (1u << QSH) - 8u;
        const double first = DESC ? xhi : xlo, span = DESC ? xlo - xhi : xhi - xlo;
        double scale = (double)(((unsigned)(NB - 2) << QSH) + 16u) / span;
        if (!(fabs(scale) < 1e300)) scale = 0.0; /* constant row (span 0), or a span beyond the exponent range: one bucket, settled by finish() */
        const double c0 = fma(-first, scale, 6755399441055744.0 + (double)QOFF); /* 2^52 + 2^51: the low word of the sum is the integer */

        const uint32_t cnt_a = smem_u32(s_bkt), idle_a = cnt_a + 4u * (unsigned)(NB + lane);
        unsigned v[EPS];
        uint32_t ba[EPS]; /* shared-memory address of each value's counter; later its slot */
        __syncwarp();
#pragma unroll
        for (int e = 0; e < EPS; e++) {
            const int s = lane + 32 * e;
            unsigned q = (unsigned)__double2loint(fma(xof(e), scale, c0));
            q = min(q, QMAX - 1u); /* memory safety only: a q out of range means the row fails the check below */
            v[e] = (q << IDXB) | (unsigned)s;
            ba[e] = cnt_a + ((q >> (QSH - 2)) & ~3u);
            unsigned inc = 1u;
            if (2 * e >= EPS && s >= n) { /* a pad: its word, and a counter of its own that nobody reads */
                v[e] = 0x80000000u | ((unsigned)(s - n) << IDXB) | (unsigned)n;
                ba[e] = idle_a;
                inc = 0u;
            }
            asm volatile("red.shared.add.u32 [%0], %1;" ::"r"(ba[e]), "r"(inc) : "memory");
        }
        x_done();
        __syncwarp();
        /* exclusive scan of the counters: BK consecutive ones per lane */
        {
            uint4 *c4 = reinterpret_cast<uint4 *>(s_bkt + lane * BK);
            unsigned c[BK];
#pragma unroll
            for (int i = 0; i < BK / 4; i++) {
                const uint4 w = c4[i];
                c[4 * i] = w.x; c[4 * i + 1] = w.y; c[4 * i + 2] = w.z; c[4 * i + 3] = w.w;
            }
            unsigned tot = 0;
#pragma unroll
            for (int i = 0; i < BK; i++) tot += c[i];
            unsigned incl = tot;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const unsigned y = __shfl_up_sync(FULL, incl, o);
                if (lane >= o) incl += y;
            }
            unsigned run = incl - tot;
#pragma unroll
            for (int i = 0; i < BK; i++) { const unsigned t = c[i]; c[i] = run; run += t; }
#pragma unroll
            for (int i = 0; i < BK / 4; i++) c4[i] = make_uint4(c[4 * i], c[4 * i + 1], c[4 * i + 2], c[4 * i + 3]);
        }
        __syncwarp();
        /* every value draws the next free slot of its bucket (pads keep their own slot number) ... */
#pragma unroll
        for (int e = 0; e < EPS; e++) {
            const int s = lane + 32 * e;
            unsigned pos;
            asm volatile("atom.shared.add.u32 %0, [%1], 1;" : "=r"(pos) : "r"(ba[e]) : "memory");
            if (2 * e >= EPS && s >= n) pos = (unsigned)s;
            ba[e] = cnt_a + 4u * pos + 16u * (pos / EPS);
        }
        __syncwarp();
        /* ... and moves there, over the counters, which are no longer needed. Slot k lives at word k + 4 * (k / EPS):
         * the same padded spacing as the counters, for the 16-byte reads below */
#pragma unroll
        for (int e = 0; e < EPS; e++) asm volatile("st.shared.u32 [%0], %1;" ::"r"(ba[e]), "r"(v[e]) : "memory");
        __syncwarp();
        {
            const uint4 *w4 = reinterpret_cast<const uint4 *>(s_bkt + lane * BK);
#pragma unroll
            for (int i = 0; i < EPS / 4; i++) {
                const uint4 w = w4[i];
                v[4 * i] = w.x; v[4 * i + 1] = w.y; v[4 * i + 2] = w.z; v[4 * i + 3] = w.w;
            }
        }
        __syncwarp();
        const unsigned minus1 = 0u - one;
        if constexpr (EPS == 16) sort16_thread(v); else sort_thread(v, one, minus1);
        { /* merge across the thread boundaries: upper half of lane L with the lower half of lane L + 1 */
            constexpr int H = EPS / 2;
            unsigned tdn[H], tup[H];
#pragma unroll
            for (int j = 0; j < H; j++) {
                tdn[j] = __shfl_down_sync(FULL, v[j], 1);
                tup[j] = __shfl_up_sync(FULL, v[H + j], 1);
            }
            if (lane != 31) {
#pragma unroll
                for (int i = 0; i < H; i++) v[H + i] = min(v[H + i], tdn[H - 1 - i]);
            }
            if (lane != 0) {
#pragma unroll
                for (int j = 0; j < H; j++) v[j] = max(v[j], tup[H - 1 - j]);
            }
            thread_tail(v, H / 2, one, minus1);
        }
        { /* the check: ascending everywhere (no two words of a row are equal) */
            unsigned nx0 = __shfl_down_sync(FULL, v[0], 1);
            if (lane == 31) nx0 = 0xffffffffu;
            bool bad = v[EPS - 1] > nx0;
#pragma unroll
            for (int r = 0; r + 1 < EPS; r++) bad = bad || (v[r] > v[r + 1]);
            if (__any_sync(FULL, bad)) {
                if (lane == 0 && full_rows) atomicAdd(full_rows, 1);
                sort(v, lane, one, minus1);
            }
        }
        finish(v, n, lane, reinterpret_cast<unsigned short *>(s_bkt), value, DESC, out);
    }

    /* v[] sorted (position k = lane*EPS + r) -> out[]: sample indices with tie marks; neighbours that share a coarse
     * key are settled on the full values (see order()) */
    template <typename ValueOf>
    __device__ __forceinline__ static void finish(const unsigned (&v)[EPS], const int n, const int lane, unsigned short *s_ord,
                                                  ValueOf value, const bool DESC, unsigned short (&out)[EPS])
    {
        /* smallest difference between neighbours of the sorted order */
        unsigned nx0 = __shfl_down_sync(FULL, v[0], 1);
        if (lane == 31) nx0 = 0xffffffffu;
        unsigned gap = nx0 - v[EPS - 1];
#pragma unroll
        for (int r = 0; r + 1 < EPS; r++) gap = min(gap, v[r + 1] - v[r]);
        if (__reduce_min_sync(FULL, gap) > IDXM) {
#pragma unroll
            for (int r = 0; r < EPS; r++) out[r] = (unsigned short)(v[r] & IDXM);
            return;
        }
        unsigned eqm = 0;
#pragma unroll
        for (int r = 0; r < EPS; r++) {
            const unsigned nxt = r + 1 < EPS ? v[r + 1] : nx0;
            if (((v[r] ^ nxt) & ~IDXM) == 0) eqm |= 1u << r;   /* a pad never shares its upper bits with any word */
        }
        /* Neighbours that share a coarse key are either exact ties (common: rounded or clipped data) or
         * values closer than the coarse resolution (rare). Fetch the values of just those positions:
         * equal -> tie mark (the packed index already orders ties by sample index); inverted -> the rare
         * transposition pass below. */
        if (!__any_sync(FULL, eqm != 0)) { /* one coarse step apart with a falling index: no collision after all */
#pragma unroll
            for (int r = 0; r < EPS; r++) out[r] = (unsigned short)(v[r] & IDXM);
            return;
        }
        {
            const unsigned prev = __shfl_up_sync(FULL, eqm >> (EPS - 1), 1);
            const unsigned inv = eqm | (eqm << 1) | (lane > 0 ? (prev & 1u) : 0u); /* positions whose value is needed */
            double fv[EPS];
#pragma unroll
            for (int r = 0; r < EPS; r++) fv[r] = ((inv >> r) & 1u) ? value((int)(v[r] & IDXM)) : 0.0;
            const double fnx = __shfl_down_sync(FULL, fv[0], 1);
            bool bad = false;
            unsigned ties = 0;
#pragma unroll
            for (int r = 0; r < EPS; r++) {
                const double fn = r + 1 < EPS ? fv[r + 1] : fnx;
                if ((eqm >> r) & 1u) {
                    bad = bad || (DESC ? fv[r] < fn : fv[r] > fn);
                    ties |= (fv[r] == fn ? 1u : 0u) << r;
                }
            }
            if (!__any_sync(FULL, bad)) {
#pragma unroll
                for (int r = 0; r < EPS; r++)
                    out[r] = (unsigned short)((v[r] & IDXM) | (((ties >> r) & 1u) << 15));
                return;
            }
        }

        /* rare: two values closer than the coarse resolution came out inverted. Odd-even transposition on the
         * full values, over the marked neighbour pairs only: a pair is marked when its two POSITIONS share a coarse
         * key, which swaps inside such a run do not change, and there are only a handful of them per row. Each lane
         * owns the marked pairs that start at its own positions. */
        __syncwarp();
#pragma unroll
        for (int r = 0; r < EPS; r++) s_ord[lane * EPS + r] = (unsigned short)(v[r] & IDXM);
        __syncwarp();
        for (;;) {
            int swapped = 0;
#pragma unroll
            for (int parity = 0; parity < 2; parity++) {
                /* position k = lane*EPS + r: its parity is that of r (EPS even), or of the lane when EPS == 1 */
                unsigned m = EPS > 1 ? (eqm & (parity ? 0xaaaaaaaau : 0x55555555u)) : (((lane & 1) == parity) ? eqm : 0u);
                while (m) {
                    const int k = lane * EPS + (__ffs(m) - 1);
                    m &= m - 1;
                    const unsigned short ia = s_ord[k], ib = s_ord[k + 1];
                    const double ka = value(ia), kb = value(ib);
                    if ((DESC ? ka < kb : ka > kb) || (ka == kb && ia > ib)) {
                        s_ord[k] = ib; s_ord[k + 1] = ia;
                        swapped = 1;
                    }
                }
                __syncwarp();
            }
            if (!__any_sync(FULL, swapped)) break;
        }
#pragma unroll
        for (int r = 0; r < EPS; r++) {
            const int k = lane * EPS + r;
            unsigned short o = (unsigned short)n;
            if (k < n) {
                o = s_ord[k];
                if (((eqm >> r) & 1u) && value(o) == value(s_ord[k + 1])) o |= 0x8000u;
            }
            out[r] = o;
        }
        __syncwarp();
    }
};

} // namespace epx
