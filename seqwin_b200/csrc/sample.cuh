// sample.cuh -- hash-range sample of a key stream: how many items per distinct key?  (sizes the buckets of
// the aggregation, agg.cuh).  Device helpers only, so that kernels which see every key anyway (the sketch's
// reorder pass, the edge-record emitter) can take the sample on the way.
#pragma once

#include <cstdint>

namespace sw {
namespace agg {

constexpr unsigned long long kEmptyKey = ~0ull;

// A hash-range sample: every item whose mixed key has its top `sbits` bits zero -- i.e. ALL occurrences of
// about 1 in 2^sbits distinct keys -- is counted (out[0]) and inserted into a small global set (out[1] =
// distinct keys inserted).  out[0] / out[1] estimates items per distinct key without bias.
constexpr int kSampleSetBits = 17;
__device__ __forceinline__ unsigned long long mix64(unsigned long long x)
{
    x ^= x >> 31;
    x *= 0x9E3779B97F4A7C15ull;
    x ^= x >> 29;
    return x;
}
// one item of the sample (see distinct_sample_kernel): callable from kernels that see every key anyway
__device__ __forceinline__ void distinct_sample_item(unsigned long long k, int sbits, unsigned long long* __restrict__ set,
                                                     unsigned long long* out)
{
    const unsigned long long m = mix64(k);
    if (sbits && (m >> (64 - sbits)) != 0) return;
    atomicAdd(&out[0], 1ull);
    uint32_t s = (uint32_t)(m >> 8) & ((1u << kSampleSetBits) - 1);
    const unsigned long long tag = k == kEmptyKey ? k - 1 : k;   // the empty marker cannot be stored (off by one key at most)
    for (int probes = 0; probes < 4096; ++probes) {
        const unsigned long long prev = atomicCAS(&set[s], kEmptyKey, tag);
        if (prev == kEmptyKey) atomicAdd(&out[1], 1ull);
        if (prev == kEmptyKey || prev == tag) break;
        s = (s + 1) & ((1u << kSampleSetBits) - 1);
    }
}

// sample 1 key in 2^sbits so that about 2^15 items are looked at
inline int sample_bits(uint64_t n)
{
    int b = 0;
    while (b < 40 && (n >> b) > (1ull << 15)) ++b;
    return b;
}

}  // namespace agg
}  // namespace sw
