// Microbenchmark (B200): cost of the warp-level histogram primitives a bucket sort would need.
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o hist_micro hist_micro.cu ; results on B200: profiles/r02m_hist_micro.txt
// Each warp owns 512 u32 counters in shared memory and issues 16 operations per "row" on pseudo-random buckets.
// Prints ns per row per SM-resident warp set and derived cycles per warp-instruction per SM.
#include <cstdio>
#include <cuda_runtime.h>
#define FULL 0xffffffffu
template <int KIND> __global__ void __launch_bounds__(256) k(unsigned *out, int rows, int nb)
{
    __shared__ unsigned cnt[8][1024];
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    unsigned *c = cnt[w];
    for (int i = lane; i < 1024; i += 32) c[i] = 0;
    __syncwarp();
    unsigned x = (blockIdx.x * 256 + threadIdx.x) * 2654435761u + 12345u, acc = 0;
    for (int r = 0; r < rows; r++) {
#pragma unroll
        for (int e = 0; e < 16; e++) {
            x = x * 1664525u + 1013904223u;
            const unsigned b = (x >> 9) & (nb - 1);
            if (KIND == 0) acc += atomicAdd(&c[b], 1u);
            else if (KIND == 1) atomicAdd(&c[b], 1u);
            else if (KIND == 2) { const unsigned m = __match_any_sync(FULL, b); acc += __popc(m & ((1u << lane) - 1u)); }
            else if (KIND == 3) { acc += c[b]; c[b ^ 1] = acc; }
            else if (KIND == 4) { /* match + leader RMW + broadcast: stable slots */
                const unsigned m = __match_any_sync(FULL, b);
                const int leader = __ffs(m) - 1;
                unsigned old = 0;
                if (lane == leader) { old = c[b]; c[b] = old + __popc(m); }
                old = __shfl_sync(FULL, old, leader);
                acc += old + __popc(m & ((1u << lane) - 1u));
            } else if (KIND == 5) { /* tagged write retry */
                bool done = false;
                unsigned slot = 0;
                while (!__all_sync(FULL, done)) {
                    if (!done) {
                        const unsigned old = c[b] & 0x07ffffffu;
                        const unsigned mine = ((unsigned)lane << 27) | (old + 1u);
                        c[b] = mine;
                        __syncwarp(__activemask());
                        if (c[b] == mine) { done = true; slot = old; }
                    }
                }
                acc += slot;
            }
        }
        __syncwarp();
    }
    if (acc == 0xdeadbeef) out[0] = acc + c[lane];
}
template <int KIND> void run(const char *name, int nb)
{
    unsigned *d; cudaMalloc(&d, 4);
    const int rows = 2000, grid = 148 * 2;   // 16 warps per SM
    cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
    k<KIND><<<grid, 256>>>(d, 10, nb);
    cudaEventRecord(a);
    k<KIND><<<grid, 256>>>(d, rows, nb);
    cudaEventRecord(b); cudaEventSynchronize(b);
    float ms; cudaEventElapsedTime(&ms, a, b);
    // per SM: 16 warps x rows x 16 ops
    const double ops_per_sm = 16.0 * rows * 16.0, cyc = ms * 1e-3 * 1.965e9;
    printf("%-34s nb=%4d  %.3f ms  %.1f cycles per warp-op per SM  (%.0f cycles per 16-op row per SM)\n", name, nb, ms, cyc / ops_per_sm, cyc / ops_per_sm * 16);
    cudaError_t e = cudaGetLastError(); if (e != cudaSuccess) printf("  error %s\n", cudaGetErrorString(e));
    cudaFree(d);
}
int main()
{
    for (int nb : {512, 32, 1}) {
        run<0>("ATOMS.ADD with return", nb);
        run<1>("ATOMS.ADD no return", nb);
        run<2>("MATCH.ANY + popc", nb);
        run<3>("LDS + STS (baseline)", nb);
        run<4>("MATCH + leader RMW + SHFL", nb);
        run<5>("tagged write retry", nb);
    }
    return 0;
}
