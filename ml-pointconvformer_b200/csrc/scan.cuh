// Single-pass exclusive scan (decoupled look-back), shared by the inverse map and grid subsampling.
#pragma once
#include "common.cuh"

namespace pcfb {

// Exclusive scan of counts[0..n) into out[0..n] (out[n] = total), single pass, decoupled look-back.
// state[b] packs (flag << 32 | value): flag 1 = aggregate available, 2 = inclusive prefix available.
constexpr int SCAN_THREADS = 256;
constexpr int SCAN_ITEMS = 8;
constexpr int SCAN_TILE = SCAN_THREADS * SCAN_ITEMS;

static __global__ void __launch_bounds__(SCAN_THREADS)
inv_scan_kernel(const int32_t *__restrict__ counts, int n, int32_t *__restrict__ out,
                unsigned long long *__restrict__ state, unsigned int *__restrict__ ticket)
{
    pdl_wait();
    __shared__ int s_tile;
    __shared__ int s_warp[SCAN_THREADS / 32];
    __shared__ int s_prefix;
    if (threadIdx.x == 0) s_tile = (int)atomicAdd(ticket, 1u);
    __syncthreads();
    const int tile = s_tile;
    const int base = tile * SCAN_TILE + threadIdx.x * SCAN_ITEMS;
    int v[SCAN_ITEMS];
    int sum = 0;
#pragma unroll
    for (int i = 0; i < SCAN_ITEMS; ++i) {
        v[i] = (base + i < n) ? counts[base + i] : 0;
        sum += v[i];
    }
    // block exclusive scan of per-thread sums
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    int incl = sum;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const int t = __shfl_up_sync(0xffffffffu, incl, o);
        if (lane >= o) incl += t;
    }
    if (lane == 31) s_warp[warp] = incl;
    __syncthreads();
    if (warp == 0) {
        int w = lane < SCAN_THREADS / 32 ? s_warp[lane] : 0;
#pragma unroll
        for (int o = 1; o < SCAN_THREADS / 32; o <<= 1) {
            const int t = __shfl_up_sync(0xffffffffu, w, o);
            if (lane >= o) w += t;
        }
        if (lane < SCAN_THREADS / 32) s_warp[lane] = w;      // inclusive warp totals
    }
    __syncthreads();
    const int warp_excl = warp == 0 ? 0 : s_warp[warp - 1];
    const int block_total = s_warp[SCAN_THREADS / 32 - 1];
    int thread_excl = warp_excl + incl - sum;

    if (threadIdx.x == 0) {
        int prefix = 0;
        if (tile == 0) {
            atomicExch(&state[0], (2ull << 32) | (unsigned int)block_total);
        } else {
            atomicExch(&state[tile], (1ull << 32) | (unsigned int)block_total);
            int look = tile - 1;
            while (true) {
                unsigned long long s = atomicAdd(&state[look], 0ull);
                const unsigned int flag = (unsigned int)(s >> 32);
                if (flag == 0) continue;                      // predecessor not published yet
                prefix += (int)(unsigned int)(s & 0xffffffffu);
                if (flag == 2) break;
                --look;
            }
            atomicExch(&state[tile], (2ull << 32) | (unsigned int)(prefix + block_total));
        }
        s_prefix = prefix;
    }
    __syncthreads();
    thread_excl += s_prefix;
#pragma unroll
    for (int i = 0; i < SCAN_ITEMS; ++i) {
        if (base + i < n) out[base + i] = thread_excl;
        thread_excl += v[i];
    }
    if (base <= n - 1 && n - 1 < base + SCAN_ITEMS) out[n] = thread_excl;   // grand total
    if (n == 0 && tile == 0 && threadIdx.x == 0) out[0] = 0;
}


}  // namespace pcfb
