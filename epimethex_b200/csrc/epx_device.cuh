/*
 * epx_device.cuh — device helpers shared by the libepx kernels (sm_100a).
 */
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include "epx_math.h"

namespace epx {

constexpr unsigned FULL = 0xffffffffu;

/* order-preserving map double -> uint64 (ascending); -0.0 and +0.0 map to the same key, as R's
 * comparisons treat them (R/EpimethexAnalysis.R:117-122 order()/sort()). Inputs carry no NaN. */
__device__ __forceinline__ unsigned long long key_of(double v)
{
    v = v + 0.0;
    long long b = __double_as_longlong(v);
    unsigned long long u = (unsigned long long)b;
    return (b < 0) ? ~u : (u | 0x8000000000000000ull);
}

/* the same key as two 32-bit halves (what the warp sort consumes) */
__device__ __forceinline__ void key_halves(double v, unsigned &hi, unsigned &lo)
{
    v = v + 0.0;
    const int h = __double2hiint(v);
    const unsigned m = (unsigned)(h >> 31);
    hi = (unsigned)h ^ (m | 0x80000000u);
    lo = (unsigned)__double2loint(v) ^ m;
}

__device__ __forceinline__ double warp_sum(double v)
{
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(FULL, v, o);
    return v;
}

__device__ __forceinline__ int warp_max_i(int v) { return __reduce_max_sync(FULL, v); }

/* sum over the whole block; every thread gets the result. red: >= 33 doubles of shared memory */
__device__ __forceinline__ double block_sum(double v, double *red)
{
    int lane = threadIdx.x & 31, w = threadIdx.x >> 5, nw = (blockDim.x + 31) >> 5;
    v = warp_sum(v);
    __syncthreads(); /* protect red[] from a previous use */
    if (lane == 0) red[w] = v;
    __syncthreads();
    if (w == 0) {
        double t = lane < nw ? red[lane] : 0.0;
        t = warp_sum(t);
        if (lane == 0) red[32] = t;
    }
    __syncthreads();
    return red[32];
}

/* ---- mbarrier + bulk-copy (TMA, non-tensor) wrappers ---- */
__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t *bar, int count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void fence_barrier_init()
{
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t *bar, uint32_t bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_g2s(void *dst, const void *src, uint32_t bytes, uint64_t *bar)
{
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                     smem_u32(dst)),
                 "l"(src), "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t *bar, uint32_t parity)
{
    uint32_t ok;
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                 : "=r"(ok)
                 : "r"(smem_u32(bar)), "r"(parity)
                 : "memory");
    return ok != 0;
}
/* bounded wait: a lost transaction traps (kernel error) instead of hanging the GPU */
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity)
{
    for (unsigned spin = 0; !mbar_try_wait(bar, parity); ++spin)
        if (spin > (1u << 26)) __trap();
}

/* In-place ascending bitonic sort of npad (power of two) (key, idx) records in shared memory, ordered
 * by key then idx. All threads of the block must call. */
__device__ __forceinline__ void block_bitonic_sort(unsigned long long *keys, unsigned short *idx, int npad)
{
    const int half = npad >> 1;
    for (int k = 2; k <= npad; k <<= 1) {
        for (int j = k >> 1; j > 0; j >>= 1) {
            for (int t = threadIdx.x; t < half; t += blockDim.x) {
                int i = ((t & ~(j - 1)) << 1) | (t & (j - 1));
                int p = i | j;
                bool up = ((i & k) == 0);
                unsigned long long ki = keys[i], kp = keys[p];
                unsigned short ii = idx[i], ip = idx[p];
                bool gt = (ki > kp) || (ki == kp && ii > ip);
                if (gt == up) {
                    keys[i] = kp; keys[p] = ki;
                    idx[i] = ip; idx[p] = ii;
                }
            }
            __syncthreads();
        }
    }
}

} // namespace epx
