// nbr.cuh -- which stream neighbours a minimizer "owns": shared by the partition that generates the neighbour
// arrays (radix.cu) and the kernels that consume them (agg.cuh); also compiled by the CPU emulator.
#pragma once

#include <cstdint>

namespace sw {
namespace agg {

// Owned neighbours of stream item g (key h, record rec): the pair (g-1, g) belongs to the item with the SMALLER
// hash, the item before wins a tie (a self-loop is counted once).  0 = not owned / no such neighbour.
__device__ __forceinline__ uint64_t owned_prev(const uint64_t* __restrict__ keys, const uint64_t* __restrict__ vals, uint64_t g,
                                               uint64_t h, uint32_t rec)
{
    if (g == 0 || (uint32_t)(vals[g - 1] >> 32) != rec) return 0;
    const uint64_t hp = keys[g - 1];
    return hp > h ? hp : 0;
}
__device__ __forceinline__ uint64_t owned_next(const uint64_t* __restrict__ keys, const uint64_t* __restrict__ vals, uint64_t g,
                                               uint64_t n_total, uint64_t h, uint32_t rec)
{
    if (g + 1 >= n_total || (uint32_t)(vals[g + 1] >> 32) != rec) return 0;
    const uint64_t hn = keys[g + 1];
    return hn >= h ? hn : 0;
}

}  // namespace agg
}  // namespace sw
