// agg.cuh -- bucketed group-by: the aggregation kernels of the graph stage.
//
// The minimizer stream (h1, pos | record << 32) and the adjacent-pair records (rank pair, assembly) are
// both aggregated the same way: a STABLE partition on the top P bits of the key (radix.cu, 1-3 passes)
// leaves every bucket holding a few hundred to a few thousand items in input order; one CTA then groups a
// bucket by key in shared memory -- an open-addressing table of its distinct keys.
//
//   nodes   group_count_kernel  counts the items of every distinct hash and writes the bucket's distinct hashes in
//                               ascending order (rank by counting) next to their sizes,
//           group_place_kernel  after an exclusive scan of the per-bucket distinct counts: moves every k-mer to its
//                               final place -- group offset + stable rank inside the group, so a node's k-mers
//                               stay in (record, pos) order --, counts the distinct assemblies of every node by
//                               class (get_penalty, cpp/src/seqwin/filter.cpp:92-136) and writes the nodes.
//   edges   edge_group_kernel   distinct pairs of the bucket in ascending order with their weights (distinct
//                               assemblies, cpp/src/seqwin/build.cpp:177-189) in one pass: no record is moved,
//           edge_out_kernel     after the scan: pairs of ranks -> pairs of hashes, the finished edges.
//
// This is what build_worker + merge_thread_graphs do with two hash maps and a CPU radix sort
// (cpp/src/seqwin/build.cpp:152-241, build_internals.cpp:159-291); the data is touched twice after the
// partition instead of once per radix digit.  A bucket never holds more distinct NODE keys than the table
// takes (P is chosen from the item count and h1 is a 64-bit mix); an EDGE bucket can -- a hub node with
// thousands of distinct neighbours -- and is then marked and handed to the sort-based path (graph.cu).
//
// The file compiles under nvcc and, through tests/emul/cuda_emul.h, as plain C++ (CPU tests).
#pragma once

#include <cstdint>

#include "common.h"
#include "nbr.cuh"
#include "sample.cuh"

namespace sw {
namespace agg {

constexpr int kNT = 256;
constexpr int kNW = kNT / 32;
constexpr int kSlotBits = 11;
constexpr int kSlots = 1 << kSlotBits;      // open-addressing table of one bucket's distinct keys
constexpr int kMaxDistinct = 1024;           // distinct keys a bucket may hold (table load <= 0.5: linear probing stays short)
constexpr int kItems = 8;                    // items per thread and placement chunk
constexpr int kChunk = kNT * kItems;
static_assert(kMaxDistinct == 4 * kNT, "group_place_kernel scans 4 groups per thread");
// kEmptyKey (sample.cuh): in-bucket keys have their top P >= 1 bits cleared, so ~0 never occurs
constexpr uint32_t kOverflow = 0xFFFFFFFFu;       // bucket_d value of a bucket left to the sort-based path

// smallest P >= 1 with n / 2^P <= per_bucket
inline int partition_bits(uint64_t n, uint32_t per_bucket)
{
    int p = 1;
    while (p < 40 && (n >> p) > per_bucket) ++p;
    return p;
}

// start[b] = first item whose key >> shift is >= b, for b in [0, n_buckets]  (keys ascending in those bits)
__global__ void __launch_bounds__(256) bucket_bounds_kernel(const uint64_t* __restrict__ keys, uint64_t n, int shift,
                                                            uint64_t n_buckets, uint32_t* __restrict__ start)
{
    const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    for (uint64_t j = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; j <= n; j += stride) {
        const uint64_t lo = j ? (keys[j - 1] >> shift) + 1 : 0;
        const uint64_t hi = j < n ? (keys[j] >> shift) : n_buckets;
        for (uint64_t b = lo; b <= hi; ++b) start[b] = (uint32_t)j;
    }
}

// The same bounds by binary search, one thread per bucket: cheaper than the scan above when a bucket holds many
// items (the keys only need to be partitioned on key >> shift, not sorted).
// bucket0: the keys belong to one hash slice of a sliced build, whose first bucket has this (global) number.
__global__ void __launch_bounds__(256) bucket_search_kernel(const uint64_t* __restrict__ keys, uint64_t n, int shift,
                                                            uint64_t n_buckets, uint32_t* __restrict__ start, uint64_t bucket0 = 0)
{
    const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    for (uint64_t b = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; b <= n_buckets; b += stride) {
        uint64_t lo = 0, hi = n;
        while (lo < hi) {
            const uint64_t mid = (lo + hi) >> 1;
            if ((keys[mid] >> shift) < bucket0 + b) lo = mid + 1; else hi = mid;
        }
        start[b] = (uint32_t)lo;
    }
}

// Shared state too large for a static __shared__ array lives in dynamic shared memory on the device and in a
// static object in the CPU emulator (which runs one block at a time).
#if defined(__CUDACC__)
#define SW_DYN_SMEM(T, name)                                         \
    extern __shared__ __align__(16) unsigned char name##_raw[];      \
    T& name = *reinterpret_cast<T*>(name##_raw)
#else
#define SW_DYN_SMEM(T, name) \
    static T name##_obj;     \
    T& name = name##_obj
#endif

// Table slot of an in-bucket key.  Multiplicative hashing: node keys are uniform in every bit, but the edge
// keys of one bucket share most of `first` -- a hub node's pairs differ only further down, in `second`.
__device__ __forceinline__ uint32_t first_slot(uint64_t kb, int = 0)
{
    return (uint32_t)((kb * 0x9E3779B97F4A7C15ull) >> (64 - kSlotBits));
}

__global__ void __launch_bounds__(256) distinct_sample_kernel(const uint64_t* __restrict__ keys, uint64_t n, int sbits,
                                                              unsigned long long* __restrict__ set, unsigned long long* out)
{
    const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride)
        distinct_sample_item(keys[i], sbits, set, out);
}
// bucket bits from the estimate: `distinct_target` distinct keys per bucket on average, at most max_items items
inline int partition_bits_for(uint64_t n, double items_per_key, double distinct_target, uint32_t max_items)
{
    double per_bucket = distinct_target * (items_per_key < 1.0 ? 1.0 : items_per_key);
    if (per_bucket > (double)max_items) per_bucket = (double)max_items;
    if (per_bucket < 64.0) per_bucket = 64.0;
    return partition_bits(n, (uint32_t)per_bucket);
}

// ---- pass 1: distinct keys of every bucket, ascending, with their sizes ---------------------------------
// grp_keys / grp_cnt are indexed like the items: bucket b owns [start[b], start[b] + D_b) of them.
// item_rank[i] = rank of item i's key among the distinct keys of its bucket (what pass 2 groups by).

// slot of key kb in the table, inserting it if absent; kSlots if the table is full
__device__ __forceinline__ uint32_t table_upsert(unsigned long long* t_key, unsigned long long kb, bool* inserted)
{
    uint32_t s = first_slot(kb, 0);
    *inserted = false;
    for (int probes = 0; probes < kSlots; ++probes) {
        unsigned long long cur = t_key[s];
        if (cur == kEmptyKey) {
            cur = atomicCAS(&t_key[s], kEmptyKey, kb);
            if (cur == kEmptyKey) {
                *inserted = true;
                return s;
            }
        }
        if (cur == kb) return s;
        s = (s + 1) & (kSlots - 1);
    }
    return (uint32_t)kSlots;
}

__global__ void __launch_bounds__(kNT) group_count_kernel(const uint64_t* __restrict__ keys, const uint32_t* __restrict__ start,
                                                          int key_bits, uint32_t max_distinct, uint64_t* __restrict__ grp_keys,
                                                          uint32_t* __restrict__ grp_cnt, uint32_t* __restrict__ bucket_d,
                                                          uint16_t* __restrict__ item_rank, uint64_t bucket0 = 0)
{
    __shared__ unsigned long long t_key[kSlots];
    __shared__ uint32_t t_cnt[kSlots];              // items per slot; afterwards: rank of the slot's key
    __shared__ unsigned long long dk[kMaxDistinct];
    __shared__ uint16_t dslot[kMaxDistinct];
    __shared__ uint32_t s_sub[kNT], s_sub_start[kNT + 1], s_wsum[kNW];
    __shared__ uint32_t s_n;
    static_assert(kNT == 256, "one thread per 8-bit sub-range");
    const uint32_t b = blockIdx.x, tid = threadIdx.x;
    const uint32_t bs = start[b], n = start[b + 1] - bs;
    if (n == 0) {
        if (tid == 0) bucket_d[b] = 0;
        return;
    }
    for (uint32_t s = tid; s < (uint32_t)kSlots; s += kNT) {
        t_key[s] = kEmptyKey;
        t_cnt[s] = 0;
    }
    if (tid == 0) s_n = 0;
    __syncthreads();
    const uint64_t lowmask = (1ull << key_bits) - 1;
    for (uint32_t i0 = tid; i0 < n; i0 += 4 * kNT) {
        unsigned long long kq[4];   // four independent loads in flight per thread
#pragma unroll
        for (int q = 0; q < 4; ++q) kq[q] = i0 + q * kNT < n ? keys[bs + i0 + q * kNT] : 0;
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            if (i0 + q * kNT >= n) break;
            if (*(volatile uint32_t*)&s_n > max_distinct) break;   // more distinct keys than allowed: the other path takes the bucket
            bool inserted;
            const uint32_t s = table_upsert(t_key, kq[q] & lowmask, &inserted);
            if (s == (uint32_t)kSlots) {
                atomicAdd(&s_n, (uint32_t)kSlots);
                break;
            }
            if (inserted) atomicAdd(&s_n, 1u);
            atomicAdd(&t_cnt[s], 1u);
        }
    }
    __syncthreads();
    const uint32_t D = s_n;
    if (D > max_distinct) {
        if (tid == 0) bucket_d[b] = kOverflow;
        return;
    }
    // distinct keys in ascending order: first by their next 8 key bits (a counting sort into dk[]), then each
    // key is ranked against the few keys that share those bits
    const int sshift = key_bits > 8 ? key_bits - 8 : 0;
    s_sub[tid] = 0;
    __syncthreads();
    for (uint32_t s = tid; s < (uint32_t)kSlots; s += kNT)
        if (t_key[s] != kEmptyKey) atomicAdd(&s_sub[(uint32_t)(t_key[s] >> sshift) & 255u], 1u);
    __syncthreads();
    {
        const uint32_t lane = tid & 31, wid = tid >> 5;
        const uint32_t c = s_sub[tid];
        uint32_t inc = c;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            const uint32_t t = __shfl_up_sync(0xffffffffu, inc, d);
            if (lane >= (uint32_t)d) inc += t;
        }
        if (lane == 31) s_wsum[wid] = inc;
        __syncthreads();
        uint32_t ex = inc - c;
#pragma unroll
        for (int w = 0; w < kNW; ++w)
            if ((uint32_t)w < wid) ex += s_wsum[w];
        s_sub_start[tid] = ex;
        s_sub[tid] = ex;           // now: cursor of the sub-range
        if (tid == kNT - 1) s_sub_start[kNT] = ex + c;
    }
    __syncthreads();
    for (uint32_t s = tid; s < (uint32_t)kSlots; s += kNT) {
        if (t_key[s] != kEmptyKey) {
            const uint32_t i = atomicAdd(&s_sub[(uint32_t)(t_key[s] >> sshift) & 255u], 1u);
            dk[i] = t_key[s];
            dslot[i] = (uint16_t)s;
        }
    }
    __syncthreads();
    const uint64_t prefix = (bucket0 + b) << key_bits;
    for (uint32_t i = tid; i < D; i += kNT) {
        const unsigned long long k = dk[i];
        const uint32_t sb = (uint32_t)(k >> sshift) & 255u;
        const uint32_t lo = s_sub_start[sb], hi = s_sub_start[sb + 1];
        uint32_t r = lo;
        for (uint32_t j = lo; j < hi; ++j) r += dk[j] < k ? 1u : 0u;
        const uint32_t sl = dslot[i];
        grp_keys[bs + r] = k | prefix;
        grp_cnt[bs + r] = t_cnt[sl];
        t_cnt[sl] = r;          // only this thread touches the slot: its count has just been read
    }
    if (tid == 0) bucket_d[b] = D;
    __syncthreads();
    // every item's rank: second look at the (cached) keys
    for (uint32_t i0 = tid; i0 < n; i0 += 4 * kNT) {
        unsigned long long kq[4];
#pragma unroll
        for (int q = 0; q < 4; ++q) kq[q] = i0 + q * kNT < n ? keys[bs + i0 + q * kNT] : 0;
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            if (i0 + q * kNT >= n) break;
            const unsigned long long kb = kq[q] & lowmask;
            uint32_t s = first_slot(kb, 0);
            while (t_key[s] != kb) s = (s + 1) & (kSlots - 1);
            item_rank[bs + i0 + q * kNT] = (uint16_t)t_cnt[s];
        }
    }
}

// bucket_d -> 64-bit counts for the scans: distinct keys of the buckets grouped here (0 for the others), and
// the item counts of the buckets that were not (tot[0] += such buckets, tot[1] += their items)
__global__ void __launch_bounds__(256) bucket_counts_kernel(const uint32_t* __restrict__ bucket_d, const uint32_t* __restrict__ start,
                                                            uint64_t n_buckets, unsigned long long* __restrict__ d64,
                                                            unsigned long long* __restrict__ ovf_items64, unsigned long long* tot)
{
    const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    for (uint64_t b = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; b < n_buckets; b += stride) {
        const uint32_t d = bucket_d[b];
        const bool ovf = d == kOverflow;
        d64[b] = ovf ? 0ull : (unsigned long long)d;
        if (ovf_items64) ovf_items64[b] = ovf ? (unsigned long long)(start[b + 1] - start[b]) : 0ull;
        if (ovf) {
            atomicAdd(&tot[0], 1ull);
            atomicAdd(&tot[1], (unsigned long long)(start[b + 1] - start[b]));
        }
    }
}

// ---- pass 2: placement, distinct-assembly counts, output ---------------------------------------------------
struct NodeOut {
    static constexpr bool kNodes = true;
    using Val = unsigned long long;                 // pos | record << 32 == sw_kmer
    const Val* vals;                                // partitioned like the keys
    Val* placed;                                    // kmers (final order)
    uint32_t* placed_asm;                           // [items] assembly of every placed k-mer (scoring only)
    sw_node* nodes;
    uint64_t* node_hash;                            // optional compact copy of nodes[].hash (null: not written)
    const uint32_t* rec_asm;                        // [records] assembly of a record   (scoring only)
    uint32_t rec_base;
    const uint8_t* is_target;                       // [assemblies]                      (scoring only)
    double inv_t, inv_n;
    int counts_only;                                // a shard of a multi-GPU build: penalty left at 0
    uint64_t kmer_base;                             // sliced build: k-mers placed by earlier slices (nodes[].start / stop are global)
};

struct PlaceArgs {
    const uint16_t* item_rank;           // rank of every item's key inside its bucket (group_count_kernel)
    const uint32_t* start;               // bucket bounds
    int key_bits;                        // 64 - P
    const uint64_t* grp_keys;            // from group_count_kernel
    const uint32_t* grp_cnt;
    const uint32_t* bucket_d;
    const unsigned long long* grp_base;  // exclusive scan of the distinct counts: first output group of a bucket
};

// what placement writes for one item: nodes -- the k-mer, and its assembly when the groups are counted;
// edges -- the assembly only
template <bool COUNT>
__device__ __forceinline__ void place_item(const NodeOut& o, uint64_t dst, unsigned long long v)
{
    o.placed[dst] = v;
    if (COUNT) o.placed_asm[dst] = o.rec_asm[(uint32_t)(v >> 32) - o.rec_base];
}
__device__ __forceinline__ bool is_class_a(const NodeOut& o, uint32_t as) { return o.is_target[as] != 0; }



__device__ __forceinline__ void write_group(const NodeOut& no, unsigned long long key, unsigned long long idx, uint64_t start,
                                            uint64_t stop, uint32_t ca, uint32_t cb, bool counted)
{
    sw_node nd;
    nd.hash = key;
    nd.start = start + no.kmer_base;
    nd.stop = stop + no.kmer_base;
    nd.n_tar = ca;
    nd.n_neg = cb;
    nd.penalty = 0.0;
    if (counted && !no.counts_only) {
        // filter.cpp:132-134 evaluated without fused multiply-add (x86-64 baseline)
        const double ft = __dmul_rn((double)nd.n_tar, no.inv_t);
        const double fn = __dmul_rn((double)nd.n_neg, no.inv_n);
        const double d1 = __dsub_rn(1.0, ft);
        nd.penalty = __dsqrt_rn(__dadd_rn(__dmul_rn(d1, d1), __dmul_rn(fn, fn)));
    }
    no.nodes[idx] = nd;
    if (no.node_hash) no.node_hash[idx] = key;
}

template <class Out, bool COUNT>
__global__ void __launch_bounds__(kNT, 4) group_place_kernel(PlaceArgs a, Out o)
{
    __shared__ uint32_t goff[kMaxDistinct + 1];     // first item of every group, relative to the bucket
    __shared__ uint32_t cursor[kMaxDistinct];       // items of the group placed by earlier chunks
    __shared__ uint32_t cbase[kMaxDistinct];        // cursor at the start of the current chunk
    __shared__ uint16_t whist[kNW][kMaxDistinct];   // per chunk: items of the group held by each warp -> exclusive over warps
    __shared__ uint32_t c_a[kMaxDistinct], c_b[kMaxDistinct];
    __shared__ uint32_t s_warp[kNW];
    const uint32_t b = blockIdx.x, tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    const uint32_t bs = a.start[b], n = a.start[b + 1] - bs;
    const uint32_t D = a.bucket_d[b];
    if (n == 0 || D == kOverflow) return;
    const unsigned long long base = a.grp_base[b];

    // group sizes -> exclusive scan (4 consecutive groups per thread)
    {
        uint32_t cnt[4];
        uint32_t sum4 = 0;
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            const uint32_t r = tid * 4 + q;
            cnt[q] = r < D ? a.grp_cnt[bs + r] : 0u;
            sum4 += cnt[q];
        }
        uint32_t inc = sum4;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            const uint32_t t = __shfl_up_sync(0xffffffffu, inc, d);
            if (lane >= (uint32_t)d) inc += t;
        }
        if (lane == 31) s_warp[wid] = inc;
        __syncthreads();
        uint32_t run = inc - sum4;
#pragma unroll
        for (int w = 0; w < kNW; ++w)
            if ((uint32_t)w < wid) run += s_warp[w];
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            const uint32_t r = tid * 4 + q;
            if (r < D) {
                goff[r] = run;
                cursor[r] = run;
                c_a[r] = 0;
                c_b[r] = 0;
            }
            run += cnt[q];
        }
        if (tid == 0) goff[D] = n;
    }
    __syncthreads();

    // placement, one chunk of kChunk items at a time; inside a chunk warp w owns a contiguous piece
    const uint32_t lt_mask = (1u << lane) - 1u;
    for (uint32_t c0 = 0; c0 < n; c0 += kChunk) {
        const uint32_t nc = n - c0 < (uint32_t)kChunk ? n - c0 : (uint32_t)kChunk;
        const uint32_t per_warp = (((nc + kNW - 1) / kNW) + 31u) & ~31u;
        const uint32_t w_lo = wid * per_warp;
        const uint32_t w_hi = w_lo + per_warp < nc ? w_lo + per_warp : nc;
        for (uint32_t r = lane; r < D; r += 32) whist[wid][r] = 0;
        __syncwarp();
        uint32_t rk[kItems], lr[kItems];
        typename Out::Val vq[kItems];
#pragma unroll
        for (int q = 0; q < kItems; ++q) {   // all loads of the chunk in flight before the first is used
            const uint32_t i = w_lo + q * 32 + lane;
            const bool have = (uint32_t)q * 32 < per_warp && i < w_hi;
            rk[q] = have ? (uint32_t)a.item_rank[bs + c0 + i] : 0xFFFFFFFFu;
            if (have) vq[q] = o.vals[bs + c0 + i];
        }
        if (w_lo < nc) {   // warp-uniform: warps beyond the chunk's last item have nothing to rank
#pragma unroll
            for (int q = 0; q < kItems; ++q) {
                lr[q] = 0;
                if ((uint32_t)q * 32 < w_hi - w_lo) {    // warp-uniform
                    const uint32_t peers = __match_any_sync(0xffffffffu, rk[q]);
                    const int leader = __ffs(peers) - 1;
                    uint32_t old = 0;
                    if ((int)lane == leader && rk[q] != 0xFFFFFFFFu) {
                        old = whist[wid][rk[q]];
                        whist[wid][rk[q]] = (uint16_t)(old + __popc(peers));
                    }
                    old = __shfl_sync(0xffffffffu, old, leader);
                    lr[q] = old + __popc(peers & lt_mask);
                    __syncwarp();
                }
            }
        }
        __syncthreads();
        for (uint32_t r = tid; r < D; r += kNT) {
            uint32_t run = 0;
#pragma unroll
            for (int w = 0; w < kNW; ++w) {
                const uint32_t t = whist[w][r];
                whist[w][r] = (uint16_t)run;
                run += t;
            }
            cbase[r] = cursor[r];
            cursor[r] += run;
        }
        __syncthreads();
#pragma unroll
        for (int q = 0; q < kItems; ++q) {
            if (rk[q] != 0xFFFFFFFFu) {
                const uint32_t pos = cbase[rk[q]] + whist[wid][rk[q]] + lr[q];
                place_item<COUNT>(o, (uint64_t)bs + pos, vq[q]);
            }
        }
        __syncthreads();
    }

    // distinct assemblies per group: the items of a group are in input order, hence sorted by assembly
    if (COUNT) {
        const uint32_t n_up = (n + 31u) & ~31u;
        for (uint32_t j0 = wid * 32; j0 < n_up; j0 += kNT) {
            const uint32_t j = j0 + lane;
            const bool valid = j < n;
            uint32_t as = 0xFFFFFFFFu, r = 0xFFFFFFFFu;
            if (valid) {
                as = o.placed_asm[bs + j];
                uint32_t lo = 0, hi = D;      // largest r with goff[r] <= j
                while (hi - lo > 1) {
                    const uint32_t mid = (lo + hi) >> 1;
                    if (goff[mid] <= j) lo = mid; else hi = mid;
                }
                r = lo;
            }
            uint32_t prev = __shfl_up_sync(0xffffffffu, as, 1);
            if (lane == 0) prev = (valid && j > 0) ? o.placed_asm[bs + j - 1] : 0xFFFFFFFFu;
            const bool fresh = valid && (j == goff[r] || as != prev);
            const bool cls_a = fresh ? is_class_a(o, as) : true;
            const uint32_t peers = __match_any_sync(0xffffffffu, r);
            const uint32_t fa = __ballot_sync(0xffffffffu, fresh && cls_a) & peers;
            const uint32_t fb = __ballot_sync(0xffffffffu, fresh && !cls_a) & peers;
            if (valid && (int)lane == __ffs(peers) - 1) {
                if (fa) atomicAdd(&c_a[r], (uint32_t)__popc(fa));
                if (fb) atomicAdd(&c_b[r], (uint32_t)__popc(fb));
            }
        }
        __syncthreads();
    }

    for (uint32_t r = tid; r < D; r += kNT) write_group(o, a.grp_keys[bs + r], base + r, (uint64_t)bs + goff[r], (uint64_t)bs + goff[r + 1],
                                                       COUNT ? c_a[r] : 0u, COUNT ? c_b[r] : 0u, COUNT);
}
// ---- edges: grouped inside the NODE buckets ---------------------------------------------------------------------
// An adjacent pair of the stream belongs to the minimizer with the smaller hash (the earlier one on a tie), so
// every edge {first <= second} is owned by items of node `first` -- which a node bucket already holds together.
// The node partition therefore carries, per item, the hashes of the (at most two) neighbours whose pair it owns
// (radix.cu, NbrArrays; 0 = none), and one CTA per bucket groups the records (node rank in the bucket, second
// hash, assembly) in shared memory:
//   * a table of the distinct (rank, second) pairs.  Its 64-bit key is second << 10 | rank; the 10 bits of
//     `second` that do not fit are kept beside it and checked after the barrier -- two pairs that agree in
//     everything else send the bucket to the side path, like a full table does;
//   * weight = distinct (pair, assembly) combinations (cpp/src/seqwin/build.cpp:177-189).  The items of a bucket
//     are in stream order, so its assemblies never decrease: inside a chunk of items a small set of
//     (slot, assembly) tags decides (whichever of several equal records comes first counts), across chunks the
//     largest assembly seen so far per pair does;
//   * the distinct pairs leave in (rank, second) order -- counting sort on the rank, then a rank by counting
//     among the few pairs of one node -- so that concatenating the buckets gives the edges in (first, second)
//     order with no sort at all.
// No node ranks, no lookups, no second partition: this replaces merge_edges' sort (build_internals.cpp:253-291).
constexpr int kRankBits = 10;                    // node rank inside a bucket
static_assert((1 << kRankBits) == kMaxDistinct, "ranks of a bucket's nodes fit kRankBits");

template <int ESB, int EI, int NTE = kNT>
struct BucketEdgeSmem {
    static constexpr int kNWE = NTE / 32;
    static constexpr int kES = 1 << ESB;                 // pair table slots
    static constexpr int kEMax = kES / 2;                // distinct pairs a bucket may hold
    static constexpr int kChunkItems = NTE * EI;
    static constexpr int kSetSlots = 4 * kChunkItems;    // <= 2 records per item, load <= 0.5
    unsigned long long t_key[kES];                       // second << 10 | rank
    uint32_t t_w[kES];                                   // distinct assemblies of the pair
    uint16_t t_hi[kES];                                  // second >> 54
    uint16_t dlist[kEMax];                               // slots in insertion order
    union {
        struct {
            unsigned long long set[kSetSlots];           // (slot << 32 | assembly) tags of the current chunk
            uint32_t t_last[kES];                        // 1 + the largest assembly earlier chunks showed for the pair
        } a;
        struct {                                         // afterwards: ordering
            uint32_t r_start[kMaxDistinct + 1], cursor[kMaxDistinct];
            unsigned long long dk[kEMax];
            uint16_t dslot[kEMax];
        } b;
    } u;
    // the chunk's records, dense per warp (half of the 2 * EI slots of an item are empty on average)
    unsigned long long st_sec[kNWE][2 * EI * 32];         // second
    unsigned long long st_ra[kNWE][2 * EI * 32];          // node rank << 32 | assembly
    uint16_t st_slot[kNWE][2 * EI * 32];                  // table slot (kES: none)
    uint32_t wsum[kNWE];
    uint32_t n_distinct, n_records, bad;
};

template <int BITS>
__device__ __forceinline__ uint32_t upsert_t(unsigned long long* t_key, unsigned long long key, bool* inserted)
{
    constexpr uint32_t kN = 1u << BITS;
    uint32_t s = (uint32_t)((key * 0x9E3779B97F4A7C15ull) >> (64 - BITS));
    *inserted = false;
    for (uint32_t probes = 0; probes < kN; ++probes) {
        unsigned long long cur = t_key[s];
        if (cur == kEmptyKey) {
            cur = atomicCAS(&t_key[s], kEmptyKey, key);
            if (cur == kEmptyKey) {
                *inserted = true;
                return s;
            }
        }
        if (cur == key) return s;
        s = (s + 1) & (kN - 1);
    }
    return kN;
}

// true if tag was not in the set yet (and now is); the set never fills up (load <= 0.5)
template <int SLOTS>
__device__ __forceinline__ bool set_insert_t(unsigned long long* set, unsigned long long tag)
{
    static_assert((SLOTS & (SLOTS - 1)) == 0, "power of two");
    uint32_t h = (uint32_t)((tag * 0x9E3779B97F4A7C15ull) >> 40) & (SLOTS - 1);
    for (;;) {
        unsigned long long cur = set[h];
        if (cur == kEmptyKey) {
            cur = atomicCAS(&set[h], kEmptyKey, tag);
            if (cur == kEmptyKey) return true;
        }
        if (cur == tag) return false;
        h = (h + 1) & (SLOTS - 1);
    }
}

struct BucketEdgeArgs {
    const uint16_t* item_rank;            // rank of every item's hash inside its bucket (group_count_kernel)
    const uint64_t* nb_prev;              // owned neighbour hashes, partitioned like the items (0 = none)
    const uint64_t* nb_next;
    const unsigned long long* vals;       // pos | record << 32
    const uint32_t* start;                // bucket bounds
    const uint32_t* rec_asm;
    uint32_t rec_base;
    uint32_t max_distinct;                // <= kEMax (tests lower it to reach the side path)
    // bucket b's distinct pairs, in (rank, second) order, at [2 * start[b], 2 * start[b] + bucket_e[b])
    uint64_t* te_second;
    uint32_t* te_w;
    uint16_t* te_r;
    uint32_t* bucket_e;                   // distinct pairs, kOverflow: the side path takes the bucket
    uint32_t* bucket_rec;                 // records (owned pairs) of the bucket
};

template <int ESB, int EI, int NTE = kNT>
__global__ void __launch_bounds__(NTE) bucket_edges_kernel(const BucketEdgeArgs a)
{
    using Smem = BucketEdgeSmem<ESB, EI, NTE>;
    constexpr int kNWE = NTE / 32;
    static_assert(kMaxDistinct % NTE == 0, "the rank counts are scanned in equal shares");
    SW_DYN_SMEM(Smem, sm);
    constexpr uint32_t kES = Smem::kES;
    const uint32_t b = blockIdx.x, tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    const uint32_t bs = a.start[b], n = a.start[b + 1] - bs;
    if (n == 0) {
        if (tid == 0) {
            a.bucket_e[b] = 0;
            a.bucket_rec[b] = 0;
        }
        return;
    }
    const bool multi = n > (uint32_t)Smem::kChunkItems;
    for (uint32_t s = tid; s < kES; s += NTE) {
        sm.t_key[s] = kEmptyKey;
        sm.t_w[s] = 0;
        if (multi) sm.u.a.t_last[s] = 0;
    }
    for (uint32_t s = tid; s < (uint32_t)Smem::kSetSlots; s += NTE) sm.u.a.set[s] = kEmptyKey;
    if (tid == 0) sm.n_distinct = sm.n_records = sm.bad = 0;
    __syncthreads();
    static_assert(Smem::kES < 65536, "slots are staged as 16-bit values");
    uint32_t my_records = 0;   // of the warp (every lane holds the same count)
    bool skip = false;
    const uint32_t lt_mask = (1u << lane) - 1u;
    unsigned long long* st_sec = sm.st_sec[wid];
    unsigned long long* st_ra = sm.st_ra[wid];
    uint16_t* st_slot = sm.st_slot[wid];
    // the items of a chunk are loaded while the chunk before is being grouped (the assembly is a dependent load)
    unsigned long long nsec[2 * EI];
    uint32_t nrk[EI], nas[EI];
    auto load_chunk = [&](uint32_t c0) {
#pragma unroll
        for (int q = 0; q < EI; ++q) {
            const uint32_t i = c0 + tid + q * NTE;
            const bool have = i < n;
            nsec[2 * q] = have ? a.nb_prev[bs + i] : 0;
            nsec[2 * q + 1] = have ? a.nb_next[bs + i] : 0;
            nrk[q] = have ? (uint32_t)a.item_rank[bs + i] : 0;
            nas[q] = have ? a.rec_asm[(uint32_t)(a.vals[bs + i] >> 32) - a.rec_base] : 0;
        }
    };
    load_chunk(0);
    for (uint32_t c0 = 0; c0 < n; c0 += Smem::kChunkItems) {
        unsigned long long sec[2 * EI];
        uint32_t rk[EI], as[EI];
#pragma unroll
        for (int q = 0; q < EI; ++q) {
            sec[2 * q] = nsec[2 * q];
            sec[2 * q + 1] = nsec[2 * q + 1];
            rk[q] = nrk[q];
            as[q] = nas[q];
        }
        if (c0 + (uint32_t)Smem::kChunkItems < n) load_chunk(c0 + Smem::kChunkItems);
        // the warp's records, dense, in its staging area
        uint32_t cnt = 0;
#pragma unroll
        for (int e = 0; e < 2 * EI; ++e) {
            const bool act = sec[e] != 0;
            const uint32_t bal = __ballot_sync(0xffffffffu, act);
            if (act) {
                const uint32_t at = cnt + (uint32_t)__popc(bal & lt_mask);
                st_sec[at] = sec[e];
                st_ra[at] = ((unsigned long long)rk[e >> 1] << 32) | as[e >> 1];
            }
            cnt += (uint32_t)__popc(bal);
        }
        my_records += cnt;
        if (skip) continue;   // uniform: the bucket already belongs to the side path, which needs the record count
        __syncwarp();
        for (uint32_t j0 = 0; j0 < cnt; j0 += 32) {
            const uint32_t j = j0 + lane;
            const bool act = j < cnt;
            const uint32_t mask = __ballot_sync(0xffffffffu, act);
            if (act) {
                const unsigned long long second = st_sec[j], ra = st_ra[j];
                const unsigned long long key = (second << kRankBits) | (ra >> 32);
                uint32_t s = kES;
                bool inserted = false;
                if (key != kEmptyKey) s = upsert_t<ESB>(sm.t_key, key, &inserted);
                if (inserted) {
                    sm.t_hi[s] = (uint16_t)(second >> (64 - kRankBits));
                    const uint32_t at = atomicAdd(&sm.n_distinct, 1u);
                    if (at < (uint32_t)Smem::kEMax) sm.dlist[at] = (uint16_t)s;
                }
                __syncwarp(mask);
                if (s != kES) {
                    bool fresh = set_insert_t<Smem::kSetSlots>(sm.u.a.set, ((unsigned long long)s << 32) | (uint32_t)ra);
                    if (fresh && multi) fresh = (uint32_t)ra + 1u > sm.u.a.t_last[s];
                    if (fresh) atomicAdd(&sm.t_w[s], 1u);
                } else {
                    sm.bad = 1;   // the empty marker as a key, or a full table
                }
                st_slot[j] = (uint16_t)s;
                __syncwarp(mask);
            }
        }
        __syncthreads();
        // the bits of `second` the key leaves out: the slot's first record wrote them, everybody checks
        const bool more = c0 + (uint32_t)Smem::kChunkItems < n;   // uniform: another chunk follows
        for (uint32_t j = lane; j < cnt; j += 32) {
            const uint32_t s = st_slot[j];
            if (s == kES) continue;
            if (sm.t_hi[s] != (uint16_t)(st_sec[j] >> (64 - kRankBits))) sm.bad = 1;
            if (more) atomicMax(&sm.u.a.t_last[s], (uint32_t)st_ra[j] + 1u);
        }
        if (sm.n_distinct > a.max_distinct) skip = true;   // uniform (read between the barriers)
        if (more) {
            for (uint32_t s = tid; s < (uint32_t)Smem::kSetSlots; s += NTE) sm.u.a.set[s] = kEmptyKey;
            __syncthreads();
        }
    }
    if (lane == 0 && my_records) atomicAdd(&sm.n_records, my_records);
    __syncthreads();
    const uint32_t D = sm.n_distinct;
    if (D > a.max_distinct || sm.bad) {
        if (tid == 0) {
            a.bucket_e[b] = kOverflow;
            a.bucket_rec[b] = sm.n_records;
        }
        return;
    }
    // (rank, second) order: counting sort on the rank into dk[] (over the dead set), then every pair is ranked
    // against the other pairs of its node
    for (uint32_t r = tid; r <= (uint32_t)kMaxDistinct; r += NTE) sm.u.b.r_start[r] = 0;
    __syncthreads();
    for (uint32_t i = tid; i < D; i += NTE)
        atomicAdd(&sm.u.b.r_start[(uint32_t)sm.t_key[sm.dlist[i]] & (kMaxDistinct - 1)], 1u);
    __syncthreads();
    {   // exclusive scan of the kMaxDistinct counts, kPer consecutive ranks per thread
        constexpr int kPer = kMaxDistinct / NTE;
        uint32_t cnt[kPer], sum4 = 0;
#pragma unroll
        for (int q = 0; q < kPer; ++q) {
            cnt[q] = sm.u.b.r_start[tid * kPer + q];
            sum4 += cnt[q];
        }
        uint32_t inc = sum4;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            const uint32_t t = __shfl_up_sync(0xffffffffu, inc, d);
            if (lane >= (uint32_t)d) inc += t;
        }
        if (lane == 31) sm.wsum[wid] = inc;
        __syncthreads();
        uint32_t run = inc - sum4;
#pragma unroll
        for (int w = 0; w < kNWE; ++w)
            if ((uint32_t)w < wid) run += sm.wsum[w];
#pragma unroll
        for (int q = 0; q < kPer; ++q) {
            sm.u.b.r_start[tid * kPer + q] = run;
            sm.u.b.cursor[tid * kPer + q] = run;
            run += cnt[q];
        }
        if (tid == NTE - 1) sm.u.b.r_start[kMaxDistinct] = run;
    }
    __syncthreads();
    for (uint32_t i = tid; i < D; i += NTE) {
        const uint32_t s = sm.dlist[i];
        const unsigned long long key = sm.t_key[s];
        const uint32_t at = atomicAdd(&sm.u.b.cursor[(uint32_t)key & (kMaxDistinct - 1)], 1u);
        sm.u.b.dk[at] = (key >> kRankBits) | ((unsigned long long)sm.t_hi[s] << (64 - kRankBits));
        sm.u.b.dslot[at] = (uint16_t)s;
    }
    __syncthreads();
    const uint64_t out0 = 2ull * bs;
    for (uint32_t i = tid; i < D; i += NTE) {
        const unsigned long long second = sm.u.b.dk[i];
        const uint32_t s = sm.u.b.dslot[i];
        const uint32_t r = (uint32_t)sm.t_key[s] & (kMaxDistinct - 1);
        const uint32_t lo = sm.u.b.r_start[r], hi = sm.u.b.r_start[r + 1];
        uint32_t at = lo;
        for (uint32_t j = lo; j < hi; ++j) at += sm.u.b.dk[j] < second ? 1u : 0u;
        a.te_second[out0 + at] = second;
        a.te_w[out0 + at] = sm.t_w[s];
        a.te_r[out0 + at] = (uint16_t)r;
    }
    if (tid == 0) {
        a.bucket_e[b] = D;
        a.bucket_rec[b] = sm.n_records;
    }
}

// the grouped pairs of every bucket -> edges (a warp per bucket; edge_base = exclusive scan of the distinct counts;
// grp_keys[start[b] + r] = hash of the bucket's node of rank r, from group_count_kernel)
__global__ void __launch_bounds__(256) bucket_edges_out_kernel(const uint64_t* __restrict__ te_second, const uint32_t* __restrict__ te_w,
                                                               const uint16_t* __restrict__ te_r, const uint32_t* __restrict__ start,
                                                               const uint32_t* __restrict__ bucket_e,
                                                               const unsigned long long* __restrict__ edge_base, uint64_t n_buckets,
                                                               const uint64_t* __restrict__ grp_keys, sw_edge* __restrict__ edges)
{
    const uint32_t lane = threadIdx.x & 31;
    const uint64_t warp = ((uint64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5, n_warps = ((uint64_t)gridDim.x * blockDim.x) >> 5;
    for (uint64_t b = warp; b < n_buckets; b += n_warps) {
        const uint32_t D = bucket_e[b];
        if (D == kOverflow) continue;
        const uint32_t bs = start[b];
        const unsigned long long base = edge_base[b];
        for (uint32_t j = lane; j < D; j += 32) {
            sw_edge e;
            e.first = grp_keys[bs + te_r[2ull * bs + j]];
            e.second = te_second[2ull * bs + j];
            e.weight = te_w[2ull * bs + j];
            edges[base + j] = e;
        }
    }
}

// ---- buckets left to the sort-based path (hub nodes, or a clash of the checked key bits) ----------------------------
// Their records (global index of the owning node, second hash, assembly) are written out in bucket order, sorted
// by (node, second) with two stable radix sorts (radix.cu) and run-length encoded (side_final_kernel); the
// finished edges are copied to their place among the others.  Like the bucket kernel this needs no rank of
// `second`, so it also works when the node that hash belongs to has not been built yet (sliced builds).

// per-bucket 64-bit counts for the scans: distinct pairs of the buckets grouped above (0 for the others), the
// record counts of the others; tot[0] += such buckets, tot[1] += their records
__global__ void __launch_bounds__(256) bucket_edge_counts_kernel(const uint32_t* __restrict__ bucket_e, const uint32_t* __restrict__ bucket_rec,
                                                                 uint64_t n_buckets, unsigned long long* __restrict__ e64,
                                                                 unsigned long long* __restrict__ side_rec64, unsigned long long* tot)
{
    const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    for (uint64_t b = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; b < n_buckets; b += stride) {
        const bool ovf = bucket_e[b] == kOverflow;
        e64[b] = ovf ? 0ull : (unsigned long long)bucket_e[b];
        side_rec64[b] = ovf ? (unsigned long long)bucket_rec[b] : 0ull;
        if (ovf) {
            atomicAdd(&tot[0], 1ull);
            atomicAdd(&tot[1], (unsigned long long)bucket_rec[b]);
        }
    }
}

// the records of those buckets in bucket order (side_off = exclusive scan of their record counts): items in
// stream order, so the assemblies of one pair never decrease, which the run-length encoding relies on.
// node_base[b] = index of the bucket's first node among the nodes of this build (slice).
__global__ void __launch_bounds__(kNT) side_emit_kernel(const BucketEdgeArgs a, const uint32_t* __restrict__ bucket_e,
                                                        const unsigned long long* __restrict__ side_off,
                                                        const unsigned long long* __restrict__ node_base,
                                                        uint32_t* __restrict__ side_node, uint64_t* __restrict__ side_second,
                                                        uint32_t* __restrict__ side_asm)
{
    __shared__ uint32_t s_warp[kNW];
    const uint32_t b = blockIdx.x, tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    if (bucket_e[b] != kOverflow) return;
    const uint32_t bs = a.start[b], n = a.start[b + 1] - bs;
    unsigned long long running = side_off[b];
    const unsigned long long nbase = node_base[b];
    for (uint32_t c0 = 0; c0 < n; c0 += kNT) {
        const uint32_t i = c0 + tid;
        unsigned long long sp = 0, sn = 0;
        uint32_t r = 0, as = 0;
        if (i < n) {
            sp = a.nb_prev[bs + i];
            sn = a.nb_next[bs + i];
            r = a.item_rank[bs + i];
            as = a.rec_asm[(uint32_t)(a.vals[bs + i] >> 32) - a.rec_base];
        }
        const uint32_t cnt = (sp ? 1u : 0u) + (sn ? 1u : 0u);
        uint32_t inc = cnt;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            const uint32_t t = __shfl_up_sync(0xffffffffu, inc, d);
            if (lane >= (uint32_t)d) inc += t;
        }
        if (lane == 31) s_warp[wid] = inc;
        __syncthreads();
        uint32_t before = inc - cnt, total = 0;
#pragma unroll
        for (int w = 0; w < kNW; ++w) {
            const uint32_t t = s_warp[w];
            if ((uint32_t)w < wid) before += t;
            total += t;
        }
        __syncthreads();
        unsigned long long slot = running + before;
        if (sp) {
            side_node[slot] = (uint32_t)(nbase + r);
            side_second[slot] = sp;
            side_asm[slot] = as;
            ++slot;
        }
        if (sn) {
            side_node[slot] = (uint32_t)(nbase + r);
            side_second[slot] = sn;
            side_asm[slot] = as;
        }
        running += total;
    }
}

// gather for the second of the two sorts: key2[j] = node of the record that the first sort (by `second`) put at j
__global__ void __launch_bounds__(256) side_gather_node_kernel(const uint32_t* __restrict__ perm, const uint32_t* __restrict__ side_node,
                                                               uint64_t n, uint64_t* __restrict__ key2)
{
    const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    for (uint64_t j = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; j < n; j += stride) key2[j] = side_node[perm[j]];
}

// Run-length encode the records in (node, second) order (perm = the order after both sorts).
// Pass 1 (scanned == nullptr): flags[j] = (record starts a new pair) | (new pair or new assembly) << 32, so that one
// exclusive scan numbers the pairs (low word) -- the high word is not needed afterwards.
// Pass 2: a run start writes its edge (the host zeroed the weights), every record that shows a new assembly
// adds one to the weight of its pair.
__global__ void __launch_bounds__(256) side_final_kernel(const uint32_t* __restrict__ perm, const uint32_t* __restrict__ side_node,
                                                         const uint64_t* __restrict__ side_second, const uint32_t* __restrict__ side_asm,
                                                         uint64_t n, unsigned long long* __restrict__ flags,
                                                         const unsigned long long* __restrict__ scanned, const sw_node* __restrict__ nodes,
                                                         sw_edge* __restrict__ side_edges)
{
    const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    for (uint64_t j = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; j < n; j += stride) {
        const uint32_t i = perm[j];
        const uint32_t nd = side_node[i];
        const uint64_t sec = side_second[i];
        bool first = j == 0, fresh = j == 0;
        if (!first) {
            const uint32_t ip = perm[j - 1];
            first = side_node[ip] != nd || side_second[ip] != sec;
            fresh = first || side_asm[ip] != side_asm[i];
        }
        if (!scanned) {
            flags[j] = (first ? 1ull : 0ull) | (fresh ? 1ull << 32 : 0ull);
            continue;
        }
        const unsigned long long pair = (scanned[j] & 0xFFFFFFFFull) + (first ? 1u : 0u) - 1u;
        if (first) {
            side_edges[pair].first = nodes[nd].hash;
            side_edges[pair].second = sec;
        }
        if (fresh) atomicAdd(reinterpret_cast<unsigned long long*>(&side_edges[pair].weight), 1ull);
    }
}

// distinct pairs of every such bucket: its records are [side_off[b], side_off[b + 1]) of the sorted order, whose
// flags have been scanned (scanned[n] = totals)
__global__ void __launch_bounds__(256) side_count_kernel(const uint32_t* __restrict__ bucket_e, const unsigned long long* __restrict__ side_off,
                                                         const unsigned long long* __restrict__ scanned, uint64_t n_buckets,
                                                         unsigned long long* __restrict__ e64, unsigned long long* __restrict__ ovf_e64)
{
    const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    for (uint64_t b = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; b < n_buckets; b += stride) {
        unsigned long long c = 0;
        if (bucket_e[b] == kOverflow) {
            c = (scanned[side_off[b + 1]] & 0xFFFFFFFFull) - (scanned[side_off[b]] & 0xFFFFFFFFull);
            e64[b] = c;
        }
        ovf_e64[b] = c;
    }
}

// their finished edges (side_edges, in key order) go to their place among the others
__global__ void __launch_bounds__(kNT) overflow_copy_kernel(const sw_edge* __restrict__ side_edges, const uint32_t* __restrict__ bucket_d,
                                                            const unsigned long long* __restrict__ grp_base,
                                                            const unsigned long long* __restrict__ ovf_base,
                                                            sw_edge* __restrict__ edges)
{
    const uint32_t b = blockIdx.x;
    if (bucket_d[b] != kOverflow) return;
    const unsigned long long n = grp_base[b + 1] - grp_base[b];
    for (unsigned long long i = threadIdx.x; i < n; i += kNT) edges[grp_base[b] + i] = side_edges[ovf_base[b] + i];
}

}  // namespace agg
}  // namespace sw
