// scan.cuh -- single-CTA exclusive scan of a u64 array (per-block / per-tile counts).
#pragma once

#include <cuda_runtime.h>

#include <cstdint>

namespace sw {
namespace {

// In-place exclusive scan of counts[0..n); the grand total goes to *total.  Launch <<<1, 1024>>>.
__global__ void __launch_bounds__(1024) scan_counts_kernel(unsigned long long* counts, uint64_t n,
                                                           unsigned long long* total)
{
    __shared__ unsigned long long s_warp[32];
    __shared__ unsigned long long s_carry;
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    if (threadIdx.x == 0) s_carry = 0;
    __syncthreads();
    // each thread owns 4 consecutive entries per round: 4096 entries per round
    for (uint64_t base = 0; base < n; base += 4096) {
        const uint64_t i0 = base + (uint64_t)threadIdx.x * 4;
        unsigned long long v[4];
        unsigned long long sum4 = 0;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            v[j] = (i0 + j < n) ? counts[i0 + j] : 0;
            sum4 += v[j];
        }
        unsigned long long inc = sum4;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            const unsigned long long t = __shfl_up_sync(0xffffffffu, inc, d);
            if (lane >= d) inc += t;
        }
        if (lane == 31) s_warp[wid] = inc;
        __syncthreads();
        unsigned long long wbase = 0, sum = 0;
#pragma unroll
        for (int w = 0; w < 32; ++w) {
            const unsigned long long t = s_warp[w];
            if (w < wid) wbase += t;
            sum += t;
        }
        const unsigned long long carry = s_carry;
        unsigned long long run = carry + wbase + inc - sum4;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            if (i0 + j < n) counts[i0 + j] = run;
            run += v[j];
        }
        __syncthreads();
        if (threadIdx.x == 0) s_carry = carry + sum;
        __syncthreads();
    }
    if (threadIdx.x == 0) *total = s_carry;
}

}  // namespace
}  // namespace sw
