// scan.cuh -- exclusive scan of a u64 array on the device (per-block / per-tile counts, offsets).
#pragma once

#include <cuda_runtime.h>

#include <cstdint>

#include "device.h"

namespace sw {
namespace {

constexpr int kScanChunk = 4096;  // entries per CTA round: 1024 threads x 4

// Exclusive scan of up to kScanChunk entries held as v[4] per thread; returns the thread's
// exclusive prefix (of its 4 entries) and the chunk total.  Contains two barriers.
__device__ __forceinline__ unsigned long long scan_chunk(unsigned long long sum4, unsigned long long* s_warp,
                                                         unsigned long long* chunk_total)
{
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
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
    __syncthreads();
    *chunk_total = sum;
    return wbase + inc - sum4;
}

// In-place exclusive scan of counts[0..n) by ONE CTA; the grand total goes to *total.
__global__ void __launch_bounds__(1024) scan_counts_kernel(unsigned long long* counts, uint64_t n,
                                                           unsigned long long* total)
{
    __shared__ unsigned long long s_warp[32];
    unsigned long long carry = 0;
    for (uint64_t base = 0; base < n; base += kScanChunk) {
        const uint64_t i0 = base + (uint64_t)threadIdx.x * 4;
        unsigned long long v[4], sum4 = 0;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            v[j] = (i0 + j < n) ? counts[i0 + j] : 0;
            sum4 += v[j];
        }
        unsigned long long chunk;
        unsigned long long run = carry + scan_chunk(sum4, s_warp, &chunk);
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            if (i0 + j < n) counts[i0 + j] = run;
            run += v[j];
        }
        carry += chunk;
    }
    if (threadIdx.x == 0) *total = carry;
}

// Multi-CTA variant, phase 1: sum of each chunk.
__global__ void __launch_bounds__(1024) scan_chunk_sums_kernel(const unsigned long long* __restrict__ in, uint64_t n,
                                                               unsigned long long* __restrict__ sums)
{
    __shared__ unsigned long long s_warp[32];
    const uint64_t i0 = (uint64_t)blockIdx.x * kScanChunk + (uint64_t)threadIdx.x * 4;
    unsigned long long sum4 = 0;
#pragma unroll
    for (int j = 0; j < 4; ++j) sum4 += (i0 + j < n) ? in[i0 + j] : 0;
    unsigned long long chunk;
    scan_chunk(sum4, s_warp, &chunk);
    if (threadIdx.x == 0) sums[blockIdx.x] = chunk;
}

// phase 3: scan each chunk locally and add the chunk's exclusive offset.
__global__ void __launch_bounds__(1024) scan_apply_kernel(unsigned long long* __restrict__ data, uint64_t n,
                                                          const unsigned long long* __restrict__ chunk_off)
{
    __shared__ unsigned long long s_warp[32];
    const uint64_t i0 = (uint64_t)blockIdx.x * kScanChunk + (uint64_t)threadIdx.x * 4;
    unsigned long long v[4], sum4 = 0;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        v[j] = (i0 + j < n) ? data[i0 + j] : 0;
        sum4 += v[j];
    }
    unsigned long long chunk;
    unsigned long long run = chunk_off[blockIdx.x] + scan_chunk(sum4, s_warp, &chunk);
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        if (i0 + j < n) data[i0 + j] = run;
        run += v[j];
    }
}

// In-place exclusive scan of data[0..n) with the total written to *total (device pointer).
// Returns the number of kernels launched.
inline uint32_t exclusive_scan_u64(unsigned long long* data, uint64_t n, unsigned long long* total, cudaStream_t s)
{
    if (n <= 16 * (uint64_t)kScanChunk) {
        scan_counts_kernel<<<1, 1024, 0, s>>>(data, n, total);
        return 1;
    }
    const uint64_t n_chunks = (n + kScanChunk - 1) / kScanChunk;
    DevBuf<unsigned long long> sums(n_chunks, s, true);
    scan_chunk_sums_kernel<<<(uint32_t)n_chunks, 1024, 0, s>>>(data, n, sums.p);
    uint32_t launches = 1 + exclusive_scan_u64(sums.p, n_chunks, total, s);
    scan_apply_kernel<<<(uint32_t)n_chunks, 1024, 0, s>>>(data, n, sums.p);
    return launches + 1;
}

}  // namespace
}  // namespace sw
