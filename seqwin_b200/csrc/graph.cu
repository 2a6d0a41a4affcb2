// graph.cu -- minimizer pan-genome graph aggregation on the device.
//
// Input: the ordered minimizer stream of sketch.cu (h1, pos | record << 32), (record, pos) order.
// Output (reference layouts, cpp/include/seqwin/graph.hpp:15-53):
//   kmers  stream entries stably sorted by h1        -> (h1, record, pos) order
//   nodes  one per distinct h1: {hash, start, stop, 0, 0, 0.0}
//   edges  unordered pairs of stream-adjacent minimizers of one record, weight = number of
//          distinct assemblies showing the pair, sorted by (first, second)
// This is what build_worker + merge_thread_graphs produce with two hash maps and a CPU radix
// sort (cpp/src/seqwin/build.cpp:152-241, cpp/src/seqwin/build_internals.cpp:159-291); here both
// aggregations are sort + run-length encode, which is insensitive to hot keys.
//
// Edge keys are pairs of node RANKS (rank order == hash order), so the edge sort handles one
// 64-bit key whose unused high digits are skipped, instead of a 128-bit hash pair.
#include <cuda_runtime.h>

#include <cstdio>

#include <algorithm>

#include <cstdlib>
#include <cstring>

#include "agg.cuh"
#include "device.h"
#include "scan.cuh"

namespace sw {

namespace {

constexpr int kNT = 256;
constexpr int kRounds = 8;
constexpr int kBlockItems = kNT * kRounds;

struct EventTimer {
    cudaEvent_t a, b;
    cudaStream_t s;
    explicit EventTimer(cudaStream_t st) : s(st)
    {
        cudaEventCreate(&a);
        cudaEventCreate(&b);
    }
    ~EventTimer()
    {
        cudaEventDestroy(a);
        cudaEventDestroy(b);
    }
    void start() { cudaEventRecord(a, s); }
    void mark() { cudaEventRecord(b, s); }   // end of the interval, read later with read()
    float read()
    {
        cudaEventSynchronize(b);
        float ms = 0;
        cudaEventElapsedTime(&ms, a, b);
        return ms;
    }
    float stop()
    {
        cudaEventRecord(b, s);
        cudaEventSynchronize(b);
        float ms = 0;
        cudaEventElapsedTime(&ms, a, b);
        return ms;
    }
};

// Ordered rank of a flag inside one 256-thread round; `running` carries over rounds.
__device__ __forceinline__ uint32_t round_rank(bool flag, uint32_t* s_warp, uint32_t& running)
{
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    const uint32_t ballot = __ballot_sync(0xffffffffu, flag);
    const uint32_t prefix = __popc(ballot & ((1u << lane) - 1u));
    if (lane == 0) s_warp[wid] = __popc(ballot);
    __syncthreads();
    uint32_t wbase = 0, total = 0;
#pragma unroll
    for (int i = 0; i < kNT / 32; ++i) {
        const uint32_t t = s_warp[i];
        if (i < wid) wbase += t;
        total += t;
    }
    __syncthreads();
    const uint32_t r = running + wbase + prefix;
    running += total;
    return r;
}

// Ordered ranks of kRounds flags per thread (item r of thread t is element r * kNT + t of the
// block) with ONE barrier: ballots of all rounds first, then every thread sums the (round, warp)
// counts that precede it.  rank[r] = number of set flags before the item in block order; returns the
// block's total.
__device__ __forceinline__ uint32_t block_ranks(const bool (&flag)[kRounds], uint32_t (&rank)[kRounds],
                                                uint32_t (*s_cnt)[kNT / 32])
{
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    uint32_t prefix[kRounds];
#pragma unroll
    for (int r = 0; r < kRounds; ++r) {
        const uint32_t ballot = __ballot_sync(0xffffffffu, flag[r]);
        prefix[r] = __popc(ballot & ((1u << lane) - 1u));
        if (lane == 0) s_cnt[r][wid] = __popc(ballot);
    }
    __syncthreads();
    uint32_t running = 0;
#pragma unroll
    for (int r = 0; r < kRounds; ++r) {
        uint32_t before = 0, total = 0;
#pragma unroll
        for (int i = 0; i < kNT / 32; ++i) {
            const uint32_t t = s_cnt[r][i];
            if (i < wid) before += t;
            total += t;
        }
        rank[r] = running + before + prefix[r];
        running += total;
    }
    return running;   // set flags in the whole block
}

__device__ __forceinline__ uint32_t block_sum(uint32_t v, uint32_t* s_warp)
{
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
#pragma unroll
    for (int d = 16; d; d >>= 1) v += __shfl_xor_sync(0xffffffffu, v, d);
    if (lane == 0) s_warp[wid] = v;
    __syncthreads();
    uint32_t t = 0;
#pragma unroll
    for (int i = 0; i < kNT / 32; ++i) t += s_warp[i];
    return t;
}

__global__ void widen_u32_kernel(const uint32_t* __restrict__ in, uint32_t n, unsigned long long* __restrict__ out)
{
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) out[i] = in[i];
}

__global__ void iota_kernel(uint32_t* v, uint64_t n)
{
    const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) v[i] = (uint32_t)i;
}

// ---- nodes ------------------------------------------------------------------------------------

// run starts of a sorted key array (nodes and edges)
__global__ void __launch_bounds__(kNT) key_run_count_kernel(const uint64_t* __restrict__ ks, uint64_t n,
                                                            unsigned long long* counts)
{
    __shared__ uint32_t s_warp[kNT / 32];
    const uint64_t base = (uint64_t)blockIdx.x * kBlockItems;
    uint32_t c = 0;
#pragma unroll
    for (int r = 0; r < kRounds; ++r) {
        const uint64_t j = base + (uint64_t)r * kNT + threadIdx.x;
        if (j < n) c += (j == 0 || ks[j] != ks[j - 1]) ? 1u : 0u;
    }
    const uint32_t t = block_sum(c, s_warp);
    if (threadIdx.x == 0) counts[blockIdx.x] = t;
}


// SCORE: also count, per node, the distinct target / non-target assemblies among its k-mers (they
// are sorted by record, hence by assembly, inside a node): a warp-segmented sum over the node rank,
// one 64-bit atomic per (warp, node) on the adjacent {n_tar, n_neg} fields, which the host zeroed.
template <bool SCORE>
__global__ void __launch_bounds__(kNT) node_write_kernel(
    const uint64_t* __restrict__ ks, const uint32_t* __restrict__ idx, const uint64_t* __restrict__ stream_vals,
    uint64_t n, const unsigned long long* __restrict__ block_off, sw_kmer* __restrict__ kmers,
    sw_node* __restrict__ nodes, uint64_t* __restrict__ node_hash, uint32_t* __restrict__ rank_of_stream,
    const uint32_t* __restrict__ rec_asm, uint32_t rec_base, const uint8_t* __restrict__ is_target)
{
    __shared__ uint32_t s_cnt[kRounds][kNT / 32];
    __shared__ uint32_t s_asm[SCORE ? kBlockItems : 1];
    const uint64_t base = (uint64_t)blockIdx.x * kBlockItems;
    const unsigned long long off = block_off[blockIdx.x];
    // all loads of the block's items are issued before anything depends on them
    uint64_t key[kRounds], val[kRounds];
    uint32_t src[kRounds], rank[kRounds];
    bool flag[kRounds];
#pragma unroll
    for (int r = 0; r < kRounds; ++r) {
        const uint64_t j = base + (uint64_t)r * kNT + threadIdx.x;
        key[r] = j < n ? __ldcs(ks + j) : 0;
        src[r] = j < n ? __ldcs(idx + j) : 0;
    }
    // everything but rank_of_stream is touched once: streaming hints keep the scattered array in L2
#pragma unroll
    for (int r = 0; r < kRounds; ++r) {
        const uint64_t j = base + (uint64_t)r * kNT + threadIdx.x;
        val[r] = j < n ? __ldcs(stream_vals + src[r]) : 0;
        flag[r] = j < n && (j == 0 || key[r] != ks[j - 1]);
    }
    if (SCORE) {
#pragma unroll
        for (int r = 0; r < kRounds; ++r) {
            const uint64_t j = base + (uint64_t)r * kNT + threadIdx.x;
            s_asm[r * kNT + threadIdx.x] = j < n ? rec_asm[(uint32_t)(val[r] >> 32) - rec_base] : 0u;
        }
    }
    block_ranks(flag, rank, s_cnt);   // contains the barrier that publishes s_asm
#pragma unroll
    for (int r = 0; r < kRounds; ++r) {
        const uint64_t j = base + (uint64_t)r * kNT + threadIdx.x;
        unsigned long long nr = ~0ull;
        if (j < n) {
            // node rank of element j = (#run starts up to and including j) - 1
            nr = off + rank[r] + (flag[r] ? 1u : 0u) - 1u;
            __stcs(reinterpret_cast<unsigned long long*>(kmers + j), (unsigned long long)val[r]);
            rank_of_stream[src[r]] = (uint32_t)nr;
            if (flag[r]) {
                sw_node* nd = nodes + nr;
                nd->hash = key[r];
                node_hash[nr] = key[r];   // compact copy: the edge stage gathers hashes by rank (8 B, not a 40 B stride)
                nd->start = j;
                if (!SCORE) {
                    nd->n_tar = 0;
                    nd->n_neg = 0;
                    nd->penalty = 0.0;
                }
                if (nr > 0) nodes[nr - 1].stop = j;
            }
            if (j == n - 1) nodes[nr].stop = n;
        }
        if (SCORE) {
            unsigned long long v = 0;
            if (j < n) {
                const uint32_t li = (uint32_t)r * kNT + threadIdx.x, a = s_asm[li];
                bool fresh = flag[r];
                if (!fresh) {
                    // previous element in sorted order: one slot back, or the previous block's last element
                    const uint32_t pa = li ? s_asm[li - 1] : rec_asm[(uint32_t)(stream_vals[idx[j - 1]] >> 32) - rec_base];
                    fresh = pa != a;
                }
                if (fresh) v = is_target[a] ? 1ull : (1ull << 32);
            }
            const int lane = threadIdx.x & 31;
#pragma unroll
            for (int d = 1; d < 32; d <<= 1) {
                const unsigned long long k2 = __shfl_up_sync(0xffffffffu, nr, d), v2 = __shfl_up_sync(0xffffffffu, v, d);
                if (lane >= d && k2 == nr) v += v2;
            }
            const unsigned long long next = __shfl_down_sync(0xffffffffu, nr, 1);
            if (j < n && v && (lane == 31 || next != nr))
                atomicAdd(reinterpret_cast<unsigned long long*>(&nodes[nr].n_tar), v);
        }
    }
}

// penalty from the finished counts (filter.cpp:132-134), evaluated without fused multiply-add
__global__ void __launch_bounds__(256) penalty_finish_kernel(sw_node* __restrict__ nodes, uint64_t n_nodes, double inv_t,
                                                             double inv_n)
{
    const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_nodes) return;
    const double ft = __dmul_rn((double)nodes[i].n_tar, inv_t);
    const double fn = __dmul_rn((double)nodes[i].n_neg, inv_n);
    const double a = __dsub_rn(1.0, ft);
    nodes[i].penalty = __dsqrt_rn(__dadd_rn(__dmul_rn(a, a), __dmul_rn(fn, fn)));
}

// ---- node sort on the high word of h1 + tie fix-up ----------------------------------------------
// h1 is a 64-bit mix, so sorting on its high 32 bits (4 radix passes instead of 8) already separates
// almost every pair of nodes: with U_n distinct hashes about U_n^2 / 2^33 pairs share a high word.
// Those few groups are found and stably re-sorted on the full key afterwards, which makes the result
// identical to the 64-bit sort.

constexpr uint32_t kTieListCap = 1u << 20;   // boundaries recorded; more -> the caller sorts all 64 bits
constexpr uint32_t kTieMaxGroup = 1u << 14;  // largest group re-sorted in place (rank counting, O(G^2 / 32))

// j with equal high words but different keys at j-1, j
__global__ void __launch_bounds__(256) tie_detect_kernel(const uint64_t* __restrict__ ks, uint64_t n, int shift,
                                                         uint32_t* __restrict__ list, unsigned int* counters)
{
    const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    for (uint64_t j = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x + 1; j < n; j += stride) {
        const uint64_t a = ks[j - 1], b = ks[j];
        if ((a >> shift) == (b >> shift) && a != b) {
            const unsigned int c = atomicAdd(counters, 1u);
            if (c < kTieListCap) list[c] = (uint32_t)j;
        }
    }
}

// One warp per recorded boundary: find the group of equal high words around it.  Only the leftmost
// boundary of a group keeps it (groups[e] = [g0, g1), else an empty range).  A separate launch does
// the re-sorting, so no warp ever looks at a group while another one rewrites it.
__global__ void __launch_bounds__(256) tie_group_kernel(const uint64_t* __restrict__ ks, uint64_t n, int shift,
                                                        const uint32_t* __restrict__ list, uint2* __restrict__ groups,
                                                        unsigned int* counters)
{
    const int lane = threadIdx.x & 31;
    const uint32_t warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, n_warps = (gridDim.x * blockDim.x) >> 5;
    const uint32_t n_list = min(counters[0], kTieListCap);
    for (uint32_t e = warp; e < n_list; e += n_warps) {
        const long long j = list[e];
        const uint64_t hi = ks[j] >> shift, ref = ks[j - 1];
        long long g0 = j;
        bool earlier = false;   // another boundary lies to the left: that one owns the group
        for (;;) {
            const long long q = g0 - 1 - lane;
            const uint64_t kq = q >= 0 ? ks[q] : 0;
            const bool in = q >= 0 && (kq >> shift) == hi;
            const unsigned out_mask = __ballot_sync(0xffffffffu, !in);
            const unsigned below = out_mask ? ((1u << (__ffs(out_mask) - 1)) - 1u) : 0xffffffffu;  // lanes still in the group
            if (__ballot_sync(0xffffffffu, in && kq != ref) & below) earlier = true;
            if (out_mask) { g0 -= __ffs(out_mask) - 1; break; }
            g0 -= 32;
        }
        long long g1 = j;
        if (!earlier) {
            for (;;) {
                const long long q = g1 + lane;
                const bool in = q < (long long)n && (ks[q] >> shift) == hi;
                const unsigned out_mask = __ballot_sync(0xffffffffu, !in);
                if (out_mask) { g1 += __ffs(out_mask) - 1; break; }
                g1 += 32;
            }
            if (g1 - g0 > (long long)kTieMaxGroup) {
                if (lane == 0) counters[1] = 1u;
                earlier = true;
            }
        }
        if (lane == 0) groups[e] = earlier ? make_uint2(0u, 0u) : make_uint2((uint32_t)g0, (uint32_t)g1);
    }
}

// One warp per group: stable re-sort by (key, original order) through the scratch arrays.
__global__ void __launch_bounds__(256) tie_sort_kernel(uint64_t* ks, uint32_t* vs, uint64_t* ks_tmp, uint32_t* vs_tmp,
                                                       const uint2* __restrict__ groups, const unsigned int* counters)
{
    const int lane = threadIdx.x & 31;
    const uint32_t warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, n_warps = (gridDim.x * blockDim.x) >> 5;
    const uint32_t n_list = min(counters[0], kTieListCap);
    for (uint32_t e = warp; e < n_list; e += n_warps) {
        const uint32_t g0 = groups[e].x, g1 = groups[e].y;
        for (uint32_t i = g0 + lane; i < g1; i += 32) {
            const uint64_t ki = ks[i];
            uint32_t r = 0;
            for (uint32_t t = g0; t < g1; ++t) {
                const uint64_t kt = ks[t];
                r += (kt < ki || (kt == ki && t < i)) ? 1u : 0u;
            }
            ks_tmp[g0 + r] = ki;
            vs_tmp[g0 + r] = vs[i];
        }
        __syncwarp();
        for (uint32_t i = g0 + lane; i < g1; i += 32) {
            ks[i] = ks_tmp[i];
            vs[i] = vs_tmp[i];
        }
        __syncwarp();
    }
}

// Enqueue a stable sort of sp on key bits [full_begin_bit, 64): either all of those passes, or only
// the passes from begin_bit up followed by the tie fix-up.  The caller reads `counters` back at its
// next synchronisation point and, if gave_up(), refills sp and calls again with full = true.
struct TieFix {
    DevBuf<uint32_t> list;
    DevBuf<uint2> groups;
    DevBuf<unsigned long long> counters;   // low word: boundaries found, high word: a group was too large
    explicit TieFix(cudaStream_t s) : list(kTieListCap, s, true), groups(kTieListCap, s, true), counters(1, s, true) {}

    uint32_t sort(SortPairs& sp, int begin_bit, int full_begin_bit, bool full, cudaStream_t s)
    {
        SW_CUDA(cudaMemsetAsync(counters.p, 0, sizeof(unsigned long long), s));
        if (full || begin_bit <= full_begin_bit) return radix_sort_pairs(sp, 64, s, full_begin_bit);
        const uint32_t launches = radix_sort_pairs(sp, 64, s, begin_bit);
        unsigned int* tc = reinterpret_cast<unsigned int*>(counters.p);
        tie_detect_kernel<<<(uint32_t)sm_count() * 8, 256, 0, s>>>(sp.keys.p, sp.n, begin_bit, list.p, tc);
        tie_group_kernel<<<(uint32_t)sm_count() * 2, 256, 0, s>>>(sp.keys.p, sp.n, begin_bit, list.p, groups.p, tc);
        tie_sort_kernel<<<(uint32_t)sm_count() * 2, 256, 0, s>>>(sp.keys.p, sp.vals.p, sp.keys_alt.p, sp.vals_alt.p,
                                                                 groups.p, tc);
        SW_CUDA(cudaGetLastError());
        return launches + 3;
    }
    static bool gave_up(unsigned long long c) { return (uint32_t)c > kTieListCap || (c >> 32) != 0; }
};

// first key bit the reduced sorts look at (SEQWIN_SORT_BEGIN_BIT = 0, 8, .. 56; default 32).  0 sorts
// every significant bit; larger values leave more to the fix-up (tests use them to exercise it).
int sort_begin_bit()
{
    if (const char* e = getenv("SEQWIN_SORT_BEGIN_BIT")) return std::max(0, std::min(56, atoi(e) & ~7));
    return 32;
}

// ---- edges ------------------------------------------------------------------------------------

__global__ void __launch_bounds__(kNT) edge_count_kernel(const uint64_t* __restrict__ stream_vals, uint64_t n,
                                                         unsigned long long* counts)
{
    __shared__ uint32_t s_warp[kNT / 32];
    const uint64_t base = (uint64_t)blockIdx.x * kBlockItems;
    uint32_t c = 0;
#pragma unroll
    for (int r = 0; r < kRounds; ++r) {
        const uint64_t i = base + (uint64_t)r * kNT + threadIdx.x;
        if (i + 1 < n) c += ((stream_vals[i] >> 32) == (stream_vals[i + 1] >> 32)) ? 1u : 0u;
    }
    const uint32_t t = block_sum(c, s_warp);
    if (threadIdx.x == 0) counts[blockIdx.x] = t;
}

__global__ void __launch_bounds__(kNT) edge_write_kernel(
    const uint64_t* __restrict__ stream_vals, const uint32_t* __restrict__ rank_of_stream, uint64_t n,
    const uint32_t* __restrict__ rec_asm, uint32_t rec_base, const unsigned long long* __restrict__ block_off,
    int rank_bits, uint64_t* __restrict__ ekey, uint32_t* __restrict__ easm)
{
    __shared__ uint32_t s_cnt[kRounds][kNT / 32];
    const uint64_t base = (uint64_t)blockIdx.x * kBlockItems;
    const unsigned long long off = block_off[blockIdx.x];
    bool flag[kRounds];
    uint32_t rec[kRounds], rank[kRounds];
#pragma unroll
    for (int r = 0; r < kRounds; ++r) {
        const uint64_t i = base + (uint64_t)r * kNT + threadIdx.x;
        flag[r] = false;
        rec[r] = 0;
        if (i + 1 < n) {
            rec[r] = (uint32_t)(stream_vals[i] >> 32);
            flag[r] = rec[r] == (uint32_t)(stream_vals[i + 1] >> 32);
        }
    }
    block_ranks(flag, rank, s_cnt);
#pragma unroll
    for (int r = 0; r < kRounds; ++r) {
        const uint64_t i = base + (uint64_t)r * kNT + threadIdx.x;
        if (flag[r]) {
            uint32_t u = rank_of_stream[i], v = rank_of_stream[i + 1];
            if (v < u) { const uint32_t t = u; u = v; v = t; }
            const unsigned long long slot = off + rank[r];
            // (u, v) left-aligned: u in the top rank_bits bits, v right below (sorts like u << 32 | v)
            ekey[slot] = ((uint64_t)u << (64 - rank_bits)) | ((uint64_t)v << (64 - 2 * rank_bits));
            easm[slot] = rec_asm[rec[r] - rec_base];
        }
    }
}

// Run-length encode the sorted (pair, assembly) records: one edge per distinct pair, weight = number
// of distinct assemblies in its run.  Runs inside a block get their weight with one plain store (the
// difference of two prefix counts held in shared memory); only the run a block ends with -- it may
// continue in the next block -- and the tail of a run inherited from the previous block use an
// atomic add on the weight, which the host zeroed.
__global__ void __launch_bounds__(kNT) edge_final_kernel(
    const uint64_t* __restrict__ ekey, const uint32_t* __restrict__ easm, uint64_t n,
    const unsigned long long* __restrict__ block_off, int rank_bits, const uint64_t* __restrict__ node_hash,
    sw_edge* __restrict__ edges)
{
    __shared__ uint32_t s_cnt[kRounds][kNT / 32];
    __shared__ uint32_t s_cnt2[kRounds][kNT / 32];
    __shared__ uint32_t s_start[kBlockItems + 1];   // fresh assemblies before each run start of the block
    const uint64_t base = (uint64_t)blockIdx.x * kBlockItems;
    const unsigned long long off = block_off[blockIdx.x];
    uint64_t key[kRounds];
    bool new_pair[kRounds], new_asm[kRounds];
    uint32_t rank[kRounds], arank[kRounds];
#pragma unroll
    for (int r = 0; r < kRounds; ++r) {
        const uint64_t j = base + (uint64_t)r * kNT + threadIdx.x;
        key[r] = 0;
        new_pair[r] = new_asm[r] = false;
        if (j < n) {
            key[r] = ekey[j];
            new_pair[r] = (j == 0) || key[r] != ekey[j - 1];
            new_asm[r] = new_pair[r] || easm[j] != easm[j - 1];
        }
    }
    const uint32_t n_runs = block_ranks(new_pair, rank, s_cnt);
    const uint32_t n_fresh = block_ranks(new_asm, arank, s_cnt2);
#pragma unroll
    for (int r = 0; r < kRounds; ++r)
        if (new_pair[r]) s_start[rank[r]] = arank[r];
    if (threadIdx.x == 0) s_start[n_runs] = n_fresh;
    __syncthreads();
#pragma unroll
    for (int r = 0; r < kRounds; ++r) {
        if (new_pair[r]) {
            const unsigned long long e = off + rank[r];
            edges[e].first = node_hash[key[r] >> (64 - rank_bits)];
            edges[e].second = node_hash[(key[r] >> (64 - 2 * rank_bits)) & ((1ull << rank_bits) - 1)];
            const unsigned long long wgt = s_start[rank[r] + 1] - arank[r];
            if (rank[r] + 1 == n_runs) atomicAdd(reinterpret_cast<unsigned long long*>(&edges[e].weight), wgt);
            else edges[e].weight = wgt;
        }
    }
    // records before the block's first run start continue the previous block's last run
    if (threadIdx.x == 0 && s_start[0] != 0)
        atomicAdd(reinterpret_cast<unsigned long long*>(&edges[off - 1].weight), (unsigned long long)s_start[0]);
}

// ---- penalty ----------------------------------------------------------------------------------

// kLanes lanes per node: a node holds a handful of k-mers on average (M / U_n), hot nodes loop.
template <int kLanes>
__global__ void __launch_bounds__(256) penalty_kernel(
    const sw_kmer* __restrict__ kmers, uint64_t n_kmers, sw_node* __restrict__ nodes, uint64_t n_nodes,
    const uint32_t* __restrict__ rec_asm, uint32_t n_records, const uint8_t* __restrict__ is_target,
    double inv_t, double inv_n, uint32_t* err)
{
    const int lane = threadIdx.x & (kLanes - 1);
    const uint64_t grp = ((uint64_t)blockIdx.x * blockDim.x + threadIdx.x) / kLanes;
    const uint64_t n_grp = ((uint64_t)gridDim.x * blockDim.x) / kLanes;
    // every group of a warp runs the same number of iterations, so the shuffles below stay convergent
    const uint64_t n_iter = (n_nodes + n_grp - 1) / n_grp;
    for (uint64_t it = 0; it < n_iter; ++it) {
        const uint64_t i = grp + it * n_grp;
        const bool live = i < n_nodes;
        uint64_t start = 0, stop = 0;
        if (live) { start = nodes[i].start; stop = nodes[i].stop; }
        const bool range_bad = live && (start > stop || stop > n_kmers);
        uint32_t nt = 0, nn = 0, bad = 0;
        if (live && !range_bad) {
            for (uint64_t j = start + lane; j < stop; j += kLanes) {
                const uint32_t r = kmers[j].record_idx;
                if (r >= n_records) { bad |= 1u; continue; }
                bool fresh = (j == start);
                if (!fresh) {
                    const uint32_t pr = kmers[j - 1].record_idx;
                    if (r < pr) bad |= 2u;
                    fresh = pr >= n_records || rec_asm[pr] != rec_asm[r];
                }
                if (fresh) {
                    if (is_target[rec_asm[r]]) ++nt; else ++nn;
                }
            }
        }
#pragma unroll
        for (int d = kLanes / 2; d; d >>= 1) {
            nt += __shfl_xor_sync(0xffffffffu, nt, d);
            nn += __shfl_xor_sync(0xffffffffu, nn, d);
            bad |= __shfl_xor_sync(0xffffffffu, bad, d);
        }
        if (lane == 0 && live) {
            if (range_bad) {
                atomicOr(err, 4u);
            } else if (start == stop) {
                nodes[i].n_tar = 0;
                nodes[i].n_neg = 0;
                nodes[i].penalty = 1.0;
            } else {
                if (bad) atomicOr(err, bad);
                nodes[i].n_tar = nt;
                nodes[i].n_neg = nn;
                // filter.cpp:132-134 evaluated without fused multiply-add (x86-64 baseline)
                const double ft = __dmul_rn((double)nt, inv_t);
                const double fn = __dmul_rn((double)nn, inv_n);
                const double a = __dsub_rn(1.0, ft);
                nodes[i].penalty = __dsqrt_rn(__dadd_rn(__dmul_rn(a, a), __dmul_rn(fn, fn)));
            }
        }
    }
}

uint32_t blocks_for(uint64_t n) { return (uint32_t)((n + kBlockItems - 1) / kBlockItems); }

uint32_t env_u32(const char* name, uint32_t dflt)
{
    const char* e = getenv(name);
    if (!e || !*e) return dflt;
    const long v = atol(e);
    return v > 0 ? (uint32_t)v : dflt;
}

uint32_t stride_grid(uint64_t n) { return (uint32_t)std::min<uint64_t>(std::max<uint64_t>((n + 255) / 256, 1), (uint64_t)sm_count() * 16); }

void build_graph_legacy(SketchStream& st, const uint32_t* d_rec_asm, uint32_t rec_base, cudaStream_t s, DevGraph& g,
                        GraphTimes* times, const std::function<void()>* after_nodes, const ScoreArgs* score);

// items per distinct key of keys[0..n), from a hash-range sample (agg.cuh); one host synchronisation
double estimate_items_per_key(const uint64_t* keys, uint64_t n, unsigned long long* set, unsigned long long* out, cudaStream_t s)
{
    SW_CUDA(cudaMemsetAsync(set, 0xFF, sizeof(unsigned long long) << agg::kSampleSetBits, s));
    SW_CUDA(cudaMemsetAsync(out, 0, 2 * sizeof(unsigned long long), s));
    agg::distinct_sample_kernel<<<stride_grid(n), 256, 0, s>>>(keys, n, agg::sample_bits(n), set, out);
    SW_CUDA(cudaGetLastError());
    const unsigned long long* h = readback_u64(out, 2, s);
    SW_CUDA(cudaStreamSynchronize(s));
    return h[1] ? (double)h[0] / (double)h[1] : 1.0;
}

// Bucketed aggregation (agg.cuh): stable partition of the stream on the top bits of h1 -- every item taking along the
// neighbour hashes whose adjacent pair it owns --, then one CTA per bucket groups the nodes, and a second one the
// edges those nodes own, in shared memory.  Returns false -- nothing of `g` touched -- if a bucket holds more
// distinct hashes than its table takes (not expected: the bucket count follows the item count and h1 is a 64-bit
// mix) or a hash is 0 (the "no neighbour" marker); the caller then runs the sort-based path.
// table geometry of the edge kernel: (slot bits, items per thread and chunk); SEQWIN_AGG_EDGE_GEOM picks one
struct EdgeGeom {
    int slot_bits, items, threads;
    uint32_t e_max;
    size_t smem;
    void (*kernel)(const agg::BucketEdgeArgs);
};
template <int ESB, int EI, int NTE>
EdgeGeom edge_geom()
{
    return EdgeGeom{ESB, EI, NTE, (uint32_t)agg::BucketEdgeSmem<ESB, EI, NTE>::kEMax, sizeof(agg::BucketEdgeSmem<ESB, EI, NTE>),
                    agg::bucket_edges_kernel<ESB, EI, NTE>};
}
// [0] the small table, [1] the large one (choose_bucket_bits); the rest are tuning variants (SEQWIN_AGG_EDGE_GEOM = index + 1)
const EdgeGeom& edge_geom_at(uint32_t i)
{
    static const EdgeGeom geoms[] = {edge_geom<11, 1, 512>(), edge_geom<12, 1, 512>(), edge_geom<11, 2, 256>(),
                                     edge_geom<12, 2, 256>()};
    return geoms[i < 4 ? i : 0];
}

// Bucket bits and edge-table geometry.  A bucket must hold few enough distinct hashes for the node table (about
// `node_target` on average) and few enough distinct pairs for the edge table: a bucket owns the pairs whose smaller
// hash falls into it, so the low buckets hold twice the average, and the fill aimed at is half the table's limit.
// The small table is the faster one; the large one is taken when it saves a whole radix pass (8 bucket bits).
int choose_bucket_bits(uint64_t M, double per_node, double per_edge, const EdgeGeom** geom)
{
    const int p_nodes = agg::partition_bits_for(M, per_node, (double)env_u32("SEQWIN_AGG_NODE_TARGET", 300),
                                                env_u32("SEQWIN_AGG_MAX_ITEMS", 4096));
    const double pairs = (double)M / (per_edge < 1.0 ? 1.0 : per_edge);
    auto bits_for = [&](const EdgeGeom& g) {
        const double target = (double)env_u32("SEQWIN_AGG_EDGE_TARGET", g.e_max / 2);
        int p = p_nodes;
        while (p < 40 && 2.0 * pairs / (double)(1ull << p) > target) ++p;
        return p;
    };
    const uint32_t forced = env_u32("SEQWIN_AGG_EDGE_GEOM", 0);
    if (forced) {
        *geom = &edge_geom_at(forced - 1);
        return bits_for(**geom);
    }
    const int p_small = bits_for(edge_geom_at(0)), p_large = bits_for(edge_geom_at(1));
    const bool large = (p_large + 7) / 8 < (p_small + 7) / 8;
    *geom = &edge_geom_at(large ? 1 : 0);
    return large ? p_large : p_small;
}

// ---- hash slices (reduced-footprint plan) -------------------------------------------------------------------
// The scratch of the bucketed aggregation is 64 bytes per minimizer.  When that does not fit (or low_memory is
// asked for, cpp/src/seqwin/build.cpp:264-325), the build runs over H = 2^hb hash slices one after the other: the
// items of a slice are cut out of the stream -- in stream order, with their owned neighbours --, aggregated like
// a whole stream, and their k-mers / nodes / edges appended to the outputs.  Slices are independent because a
// node owns its k-mers AND the edges towards larger hashes; concatenating them in hash order is the graph.

// items per (slice, block of the stream): counts[slice * n_blocks + block]
__global__ void __launch_bounds__(kNT) slice_count_kernel(const uint64_t* __restrict__ keys, uint64_t n, int hshift, uint32_t H,
                                                          uint32_t n_blocks, unsigned long long* __restrict__ counts)
{
    __shared__ uint32_t s_cnt[64];
    if (threadIdx.x < 64) s_cnt[threadIdx.x] = 0;
    __syncthreads();
    const uint64_t base = (uint64_t)blockIdx.x * kBlockItems;
#pragma unroll
    for (int r = 0; r < kRounds; ++r) {
        const uint64_t j = base + (uint64_t)r * kNT + threadIdx.x;
        if (j < n) atomicAdd(&s_cnt[(uint32_t)(keys[j] >> hshift)], 1u);
    }
    __syncthreads();
    if (threadIdx.x < H) counts[(uint64_t)threadIdx.x * n_blocks + blockIdx.x] = s_cnt[threadIdx.x];
}

// the items of slice h, in stream order, with the neighbour hashes they own (what the first partition pass
// generates for a whole stream, radix.cu); off = exclusive scan of the counts above
__global__ void __launch_bounds__(kNT) slice_extract_kernel(const uint64_t* __restrict__ keys, const uint64_t* __restrict__ vals,
                                                            uint64_t n, int hshift, uint32_t h, uint32_t n_blocks,
                                                            const unsigned long long* __restrict__ off, NbrBuffers out,
                                                            unsigned int* zero_key)
{
    __shared__ uint32_t s_cnt[kRounds][kNT / 32];
    const uint64_t base = (uint64_t)blockIdx.x * kBlockItems;
    const unsigned long long first = off[(uint64_t)h * n_blocks];
    const unsigned long long dst0 = off[(uint64_t)h * n_blocks + blockIdx.x] - first;
    uint64_t key[kRounds];
    bool flag[kRounds];
    uint32_t rank[kRounds];
#pragma unroll
    for (int r = 0; r < kRounds; ++r) {
        const uint64_t j = base + (uint64_t)r * kNT + threadIdx.x;
        key[r] = j < n ? keys[j] : 0;
        flag[r] = j < n && (uint32_t)(key[r] >> hshift) == h;
    }
    block_ranks(flag, rank, s_cnt);
#pragma unroll
    for (int r = 0; r < kRounds; ++r) {
        if (!flag[r]) continue;
        const uint64_t j = base + (uint64_t)r * kNT + threadIdx.x;
        const uint64_t v = vals[j];
        const uint32_t rec = (uint32_t)(v >> 32);
        const unsigned long long d = dst0 + rank[r];
        out.keys[d] = key[r];
        out.vals[d] = v;
        out.prev[d] = agg::owned_prev(keys, vals, j, key[r], rec);
        out.next[d] = agg::owned_next(keys, vals, j, n, key[r], rec);
        if (key[r] == 0) *zero_key = 1u;
    }
}

struct CapacityExceeded {};   // sliced build: an output array sized from the estimates is too small

// What one pass of the aggregation needs to know about where it is: the whole stream, or one hash slice of it.
struct SliceJob {
    const NbrBuffers* R;        // partitioned items (keys | vals | prev | next)
    const NbrBuffers* Dd;       // the free set of the same size
    uint64_t n;                 // items
    uint64_t bucket0, n_buckets;
    int key_bits;
    uint64_t kmer_base, node_base, edge_base;   // outputs of the earlier slices
    bool exact;                 // whole stream: the outputs are allocated here, with their exact sizes
    uint64_t node_cap, edge_cap;
};

// nodes + k-mers (+ scoring) and the edges of the buckets of one job; the counters of `tot` were zeroed by the caller
void aggregate_buckets(const SliceJob& J, uint16_t* item_rank, const uint32_t* d_rec_asm, uint32_t rec_base, cudaStream_t s,
                       DevGraph& g, GraphTimes& tm, EventTimer& timer, const EdgeGeom& eg, const ScoreArgs* score,
                       const std::function<void()>* after_nodes, unsigned long long* tot, uint64_t* n_nodes_out,
                       uint64_t* n_edges_out, bool* fallback)
{
    using namespace agg;
    const NbrBuffers& R = *J.R;
    const NbrBuffers& Dd = *J.Dd;
    const uint64_t n_buckets = J.n_buckets;
    uint64_t* grp_keys = Dd.keys;                      // distinct hashes of every bucket (live until the edges are written)
    uint32_t* grp_cnt = reinterpret_cast<uint32_t*>(Dd.vals);
    DevBuf<uint32_t> start(n_buckets + 1, s, true), bucket_d(n_buckets, s, true);
    DevBuf<unsigned long long> d64(n_buckets + 1, s, true);
    bucket_search_kernel<<<stride_grid(n_buckets + 1), 256, 0, s>>>(R.keys, J.n, J.key_bits, n_buckets, start.p, J.bucket0);
    group_count_kernel<<<(uint32_t)n_buckets, kNT, 0, s>>>(R.keys, start.p, J.key_bits, (uint32_t)kMaxDistinct, grp_keys, grp_cnt,
                                                           bucket_d.p, item_rank, J.bucket0);
    SW_CUDA(cudaMemsetAsync(d64.p + n_buckets, 0, sizeof(unsigned long long), s));
    bucket_counts_kernel<<<stride_grid(n_buckets), 256, 0, s>>>(bucket_d.p, start.p, n_buckets, d64.p, nullptr, tot);
    tm.launches += 3 + exclusive_scan_u64(d64.p, n_buckets + 1, tot + 2, s);
    SW_CUDA(cudaGetLastError());
    const unsigned long long* tot_p = readback_u64(tot, 4, s);
    SW_CUDA(cudaStreamSynchronize(s));
    tm.sort_nodes_ms += timer.stop();
    if (tot_p[0] != 0 || tot_p[3] != 0) {
        *fallback = true;
        return;
    }
    const unsigned long long n_nodes = tot_p[2];
    *n_nodes_out = n_nodes;

    // -- nodes + kmers (+ scoring) -----------------------------------------------------------------------------
    timer.start();
    if (J.exact) {
        g.n_kmers = J.n;
        g.n_nodes = n_nodes;
        g.kmers.alloc(J.n, s);
        g.nodes.alloc(n_nodes, s);
    } else if (J.node_base + n_nodes > J.node_cap) {
        throw CapacityExceeded();
    }
    const ArenaMark node_mark = arena_mark();   // what follows is dead once the placement kernel has run
    DevBuf<uint32_t> node_asm(score ? J.n : 0, s, true);
    const PlaceArgs pa{item_rank, start.p, J.key_bits, grp_keys, grp_cnt, bucket_d.p, d64.p};
    NodeOut no{};
    no.vals = reinterpret_cast<const unsigned long long*>(R.vals);
    no.placed = reinterpret_cast<unsigned long long*>(g.kmers.p + J.kmer_base);
    no.placed_asm = node_asm.p;
    no.nodes = g.nodes.p + J.node_base;
    no.node_hash = nullptr;
    no.rec_asm = d_rec_asm;
    no.rec_base = rec_base;
    no.kmer_base = J.kmer_base;
    if (score) {
        no.is_target = score->d_is_target;
        no.inv_t = score->inv_t;
        no.inv_n = score->inv_n;
        no.counts_only = score->counts_only ? 1 : 0;
        group_place_kernel<NodeOut, true><<<(uint32_t)n_buckets, kNT, 0, s>>>(pa, no);
    } else {
        group_place_kernel<NodeOut, false><<<(uint32_t)n_buckets, kNT, 0, s>>>(pa, no);
    }
    SW_CUDA(cudaGetLastError());
    ++tm.launches;
    timer.mark();
    arena_release(node_mark);

    // -- edges: grouped inside the node buckets (the group sizes in Dd.vals and the keys of R are dead now) ------------
    EventTimer etimer(s);
    etimer.start();
    DevBuf<uint32_t> bucket_e(n_buckets, s, true), bucket_rec(n_buckets, s, true);
    DevBuf<unsigned long long> e64(n_buckets + 1, s, true), ebase(n_buckets + 1, s, true), side_rec(n_buckets + 1, s, true),
        etot(4, s, true);
    BucketEdgeArgs ea{};
    ea.item_rank = item_rank;
    ea.nb_prev = R.prev;
    ea.nb_next = R.next;
    ea.vals = reinterpret_cast<const unsigned long long*>(R.vals);
    ea.start = start.p;
    ea.rec_asm = d_rec_asm;
    ea.rec_base = rec_base;
    ea.max_distinct = std::min<uint32_t>(env_u32("SEQWIN_AGG_EDGE_DISTINCT", eg.e_max), eg.e_max);
    ea.te_second = Dd.prev;                                   // 2 n words: Dd.prev | Dd.next
    ea.te_w = reinterpret_cast<uint32_t*>(Dd.vals);           // 2 n u32
    ea.te_r = reinterpret_cast<uint16_t*>(R.keys);            // 2 n u16
    ea.bucket_e = bucket_e.p;
    ea.bucket_rec = bucket_rec.p;
    SW_CUDA(cudaFuncSetAttribute(eg.kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)eg.smem));
    eg.kernel<<<(uint32_t)n_buckets, eg.threads, eg.smem, s>>>(ea);
    SW_CUDA(cudaMemsetAsync(etot.p, 0, 4 * sizeof(unsigned long long), s));
    SW_CUDA(cudaMemsetAsync(e64.p + n_buckets, 0, sizeof(unsigned long long), s));
    SW_CUDA(cudaMemsetAsync(side_rec.p + n_buckets, 0, sizeof(unsigned long long), s));
    bucket_edge_counts_kernel<<<stride_grid(n_buckets), 256, 0, s>>>(bucket_e.p, bucket_rec.p, n_buckets, e64.p, side_rec.p, etot.p);
    SW_CUDA(cudaMemcpyAsync(ebase.p, e64.p, (n_buckets + 1) * sizeof(unsigned long long), cudaMemcpyDeviceToDevice, s));
    tm.launches += 2 + exclusive_scan_u64(ebase.p, n_buckets + 1, etot.p + 2, s);
    SW_CUDA(cudaGetLastError());
    if (after_nodes) (*after_nodes)();
    tm.nodes_ms += timer.read();
    const unsigned long long* etot_p = readback_u64(etot.p, 4, s);
    SW_CUDA(cudaStreamSynchronize(s));
    unsigned long long n_edges = etot_p[2];
    const unsigned long long n_ovf = etot_p[0], n_side = etot_p[1];
    if (getenv("SEQWIN_DEBUG_AGG"))
        fprintf(stderr, "[agg] items %llu buckets %llu (first %llu) nodes %llu edges %llu, %llu buckets (%llu records) to the sort path\n",
                (unsigned long long)J.n, (unsigned long long)n_buckets, (unsigned long long)J.bucket0, n_nodes, n_edges, n_ovf, n_side);
    DevBuf<sw_edge> side_edges;
    DevBuf<unsigned long long> ovf_e64;
    if (n_ovf) {
        // buckets with more distinct pairs than a table takes (hub nodes): their records are sorted by (owner, second
        // hash) -- two stable radix sorts, `second` first -- and run-length encoded
        if (n_side > 0xFFFFFFFFull || n_nodes > 0xFFFFFFFFull) fail_runtime("more than 2^32-1 edge records on the sort path");
        exclusive_scan_u64(side_rec.p, n_buckets + 1, etot.p + 3, s);     // -> first side slot of every such bucket
        DevBuf<uint32_t> side_node(n_side, s, true), side_asm(n_side, s, true);
        DevBuf<uint64_t> side_second(n_side, s, true);
        side_emit_kernel<<<(uint32_t)n_buckets, kNT, 0, s>>>(ea, bucket_e.p, side_rec.p, d64.p, side_node.p, side_second.p, side_asm.p);
        SortPairs sp;
        sp.n = n_side;
        sp.keys.alloc(n_side, s, true);
        sp.vals.alloc(n_side, s, true);
        SW_CUDA(cudaMemcpyAsync(sp.keys.p, side_second.p, n_side * sizeof(uint64_t), cudaMemcpyDeviceToDevice, s));
        iota_kernel<<<(uint32_t)std::min<uint64_t>((n_side + 255) / 256, 65535), 256, 0, s>>>(sp.vals.p, n_side);
        SW_CUDA(cudaGetLastError());
        tm.launches += 3 + radix_sort_pairs(sp, 64, s, 0);
        side_gather_node_kernel<<<stride_grid(n_side), 256, 0, s>>>(sp.vals.p, side_node.p, n_side, sp.keys.p);
        SW_CUDA(cudaGetLastError());
        int node_bits = 1;
        while (node_bits < 32 && (1ull << node_bits) < n_nodes) ++node_bits;
        tm.launches += 1 + radix_sort_pairs(sp, node_bits, s, 0);
        DevBuf<unsigned long long> flags(n_side + 1, s, true);
        SW_CUDA(cudaMemsetAsync(flags.p + n_side, 0, sizeof(unsigned long long), s));
        side_final_kernel<<<stride_grid(n_side), 256, 0, s>>>(sp.vals.p, side_node.p, side_second.p, side_asm.p, n_side, flags.p,
                                                              nullptr, nullptr, nullptr);
        exclusive_scan_u64(flags.p, n_side + 1, etot.p + 3, s);
        ovf_e64.alloc(n_buckets + 1, s, true);
        SW_CUDA(cudaMemsetAsync(ovf_e64.p + n_buckets, 0, sizeof(unsigned long long), s));
        side_count_kernel<<<stride_grid(n_buckets), 256, 0, s>>>(bucket_e.p, side_rec.p, flags.p, n_buckets, e64.p, ovf_e64.p);
        exclusive_scan_u64(ovf_e64.p, n_buckets + 1, etot.p + 3, s);
        SW_CUDA(cudaMemcpyAsync(ebase.p, e64.p, (n_buckets + 1) * sizeof(unsigned long long), cudaMemcpyDeviceToDevice, s));
        exclusive_scan_u64(ebase.p, n_buckets + 1, etot.p + 2, s);
        SW_CUDA(cudaGetLastError());
        const unsigned long long* e2 = readback_u64(etot.p, 4, s);
        SW_CUDA(cudaStreamSynchronize(s));
        n_edges = e2[2];
        const unsigned long long n_side_edges = e2[3];
        side_edges.alloc(n_side_edges, s, true);
        SW_CUDA(cudaMemsetAsync(side_edges.p, 0, n_side_edges * sizeof(sw_edge), s));
        side_final_kernel<<<stride_grid(n_side), 256, 0, s>>>(sp.vals.p, side_node.p, side_second.p, side_asm.p, n_side, nullptr,
                                                              flags.p, g.nodes.p + J.node_base, side_edges.p);
        SW_CUDA(cudaGetLastError());
        tm.launches += 9;
    }
    *n_edges_out = n_edges;
    if (J.exact) {
        g.n_edges = n_edges;
        g.edges.alloc(n_edges, s);
    } else if (J.edge_base + n_edges > J.edge_cap) {
        throw CapacityExceeded();
    }
    if (n_edges) {
        sw_edge* edges_out = g.edges.p + J.edge_base;
        bucket_edges_out_kernel<<<(uint32_t)std::min<uint64_t>((n_buckets + 7) / 8, (uint64_t)sm_count() * 16), 256, 0, s>>>(
            ea.te_second, ea.te_w, ea.te_r, start.p, bucket_e.p, ebase.p, n_buckets, grp_keys, edges_out);
        ++tm.launches;
        if (n_ovf) {
            overflow_copy_kernel<<<(uint32_t)n_buckets, kNT, 0, s>>>(side_edges.p, bucket_e.p, ebase.p, ovf_e64.p, edges_out);
            ++tm.launches;
        }
        SW_CUDA(cudaGetLastError());
    }
    tm.edges_ms += etimer.stop();
}

// how many hash slices (a power of two, <= 64) the aggregation needs so that its scratch and the outputs fit
uint32_t choose_slices(uint64_t M, double per_node, double per_edge, bool scored)
{
    if (const uint32_t forced = env_u32("SEQWIN_AGG_SLICES", 0)) {
        uint32_t h = 1;
        while (h < forced && h < 64) h <<= 1;
        return h;
    }
    const double outputs = 8.0 * (double)M + 1.1 * (40.0 * (double)M / per_node + 24.0 * (double)M / per_edge);
    const double per_item = (64.0 + 2.0 + (scored ? 4.0 : 0.0) + 2.0) * 1.25;
    uint32_t h = low_memory() ? 4 : 1;
    // builds that need less than 40 % of the device do not ask the driver (cudaMemGetInfo costs milliseconds next to a
    // large memory pool): that much is assumed to fit beside the stream and the packed input
    static const double device_bytes = [] {
        size_t f = 0, t = 0;
        return cudaMemGetInfo(&f, &t) == cudaSuccess ? (double)t : 0.0;
    }();
    if (outputs + per_item * (double)M / (double)h <= 0.4 * device_bytes) return h;
    // a large build: what the stream-ordered pool still caches from earlier (differently shaped) builds goes back to the
    // driver first, so that the free memory reported is what this build can really have
    {
        cudaDeviceSynchronize();
        int dev = 0;
        cudaMemPool_t pool;
        if (cudaGetDevice(&dev) == cudaSuccess && cudaDeviceGetDefaultMemPool(&pool, dev) == cudaSuccess) cudaMemPoolTrimTo(pool, 0);
    }
    size_t free_b = 0, total_b = 0;
    if (cudaMemGetInfo(&free_b, &total_b) != cudaSuccess) return h;
    const double avail = 0.9 * ((double)free_b + (double)arena_free_bytes());
    for (; h < 64; h <<= 1)
        if (outputs + per_item * (double)M / (double)h <= avail) break;
    return h;
}

// Bucketed aggregation (agg.cuh): stable partition of the stream on the top bits of h1 -- every item taking along the
// neighbour hashes whose adjacent pair it owns --, then one CTA per bucket groups the nodes, and a second one the
// edges those nodes own, in shared memory.  Returns false -- `g` left empty -- if a bucket holds more distinct hashes
// than its table takes (not expected: the bucket count follows the item count and h1 is a 64-bit mix) or a hash is
// 0 (the "no neighbour" marker); the caller then runs the sort-based path.
bool build_graph_bucketed(SketchStream& st, const uint32_t* d_rec_asm, uint32_t rec_base, cudaStream_t s, DevGraph& g,
                          GraphTimes* times, const std::function<void()>* after_nodes, const ScoreArgs* score)
{
    using namespace agg;
    const uint64_t M = st.n;
    GraphTimes tm;
    EventTimer timer(s);
    timer.start();

    // bucket size: about `target` distinct hashes each, which takes an estimate of the k-mers per distinct hash
    double per_node = st.items_per_key;   // taken by the sketch's reorder pass; a stream from elsewhere is sampled here
    if (per_node <= 0) {
        DevBuf<unsigned long long> sample_set(1ull << kSampleSetBits, s, true), sample_out(2, s, true);
        per_node = estimate_items_per_key(st.keys.p, M, sample_set.p, sample_out.p, s);
        tm.launches += 1;
    }
    const double per_edge = st.pairs_per_edge > 0 ? st.pairs_per_edge : 1.0;   // a stream from elsewhere: no pair assumed twice
    const EdgeGeom* egp = nullptr;
    const uint32_t fixed_nb = env_u32("SEQWIN_AGG_NODE_BUCKET", 0);
    int P = choose_bucket_bits(M, per_node, per_edge, &egp);
    if (fixed_nb) P = partition_bits(M, fixed_nb);
    const EdgeGeom& eg = *egp;
    const int key_bits = 64 - P;
    uint32_t H = choose_slices(M, per_node, per_edge, score != nullptr);
    while (H > 1 && (1ull << P) < H) H >>= 1;   // at least one bucket per slice
    int hb = 0;
    while ((1u << hb) < H) ++hb;
    DevBuf<unsigned long long> tot(4, s, true);   // [0] overflowing node buckets [1] their items [2] nodes; [3] low word: a key is 0
    SW_CUDA(cudaMemsetAsync(tot.p, 0, 4 * sizeof(unsigned long long), s));
    bool fallback = false;
    uint64_t n_nodes = 0, n_edges = 0;

    if (H == 1) {
        if (M > 0xFFFFFFFFull) fail_runtime("more than 2^32-1 minimizers in one aggregation pass");
        // scratch: two sets of four M-word arrays (keys | k-mers | owned previous | owned next neighbour), each one block
        DevBuf<uint64_t> setA(4 * M, s, true), setB(4 * M, s, true);
        const NbrBuffers A{setA.p, setA.p + M, setA.p + 2 * M, setA.p + 3 * M};
        const NbrBuffers B{setB.p, setB.p + M, setB.p + 2 * M, setB.p + 3 * M};
        DevBuf<uint16_t> item_rank(M, s, true);   // rank of every item's hash inside its bucket
        const NbrBuffers* R = nullptr;
        tm.launches += radix_partition_nbr(st.keys.p, st.vals.p, nullptr, M, key_bits, P, A, B, s, &R,
                                           reinterpret_cast<unsigned int*>(tot.p + 3));
        SliceJob J{};
        J.R = R;
        J.Dd = R == &A ? &B : &A;
        J.n = M;
        J.bucket0 = 0;
        J.n_buckets = 1ull << P;
        J.key_bits = key_bits;
        J.exact = true;
        aggregate_buckets(J, item_rank.p, d_rec_asm, rec_base, s, g, tm, timer, eg, score, after_nodes, tot.p, &n_nodes, &n_edges,
                          &fallback);
        if (fallback) return false;
        if (times) *times = tm;
        return true;
    }

    // ---- sliced ----
    const int hshift = 64 - hb;
    const uint32_t nb = blocks_for(M);
    DevBuf<unsigned long long> counts((uint64_t)H * nb + 1, s, true);
    slice_count_kernel<<<nb, kNT, 0, s>>>(st.keys.p, M, hshift, H, nb, counts.p);
    tm.launches += 1 + exclusive_scan_u64(counts.p, (uint64_t)H * nb, counts.p + (uint64_t)H * nb, s);
    SW_CUDA(cudaGetLastError());
    std::vector<unsigned long long> slice_start(H + 1, M);
    {
        DevBuf<unsigned long long> firsts(H, s, true);
        SW_CUDA(cudaMemcpy2DAsync(firsts.p, sizeof(unsigned long long), counts.p, (size_t)nb * sizeof(unsigned long long),
                                  sizeof(unsigned long long), H, cudaMemcpyDeviceToDevice, s));
        const unsigned long long* hp = readback_u64(firsts.p, H, s);
        SW_CUDA(cudaStreamSynchronize(s));
        std::copy(hp, hp + H, slice_start.begin());
    }
    uint64_t max_items = 0;
    for (uint32_t h = 0; h < H; ++h) max_items = std::max<uint64_t>(max_items, slice_start[h + 1] - slice_start[h]);
    if (max_items > 0xFFFFFFFFull) fail_runtime("more than 2^32-1 minimizers in one hash slice");
    DevBuf<uint64_t> setA(4 * max_items, s, true), setB(4 * max_items, s, true);
    const NbrBuffers A{setA.p, setA.p + max_items, setA.p + 2 * max_items, setA.p + 3 * max_items};
    const NbrBuffers B{setB.p, setB.p + max_items, setB.p + 2 * max_items, setB.p + 3 * max_items};
    DevBuf<uint16_t> item_rank(max_items, s, true);
    double slack = (double)env_u32("SEQWIN_AGG_SLACK_PCT", 110) / 100.0;   // head room over the estimated node / edge counts
    for (int attempt = 0;; ++attempt) {
        const uint64_t node_cap = std::min<uint64_t>(M, (uint64_t)(slack * (double)M / per_node) + env_u32("SEQWIN_AGG_SLACK_ITEMS", 1u << 20));
        const uint64_t edge_cap = std::min<uint64_t>(M, (uint64_t)(slack * (double)M / per_edge) + env_u32("SEQWIN_AGG_SLACK_ITEMS", 1u << 20));
        g.n_kmers = M;
        g.kmers.alloc(M, s);
        g.nodes.alloc(node_cap, s);
        g.edges.alloc(edge_cap, s);
        uint64_t node_base = 0, edge_base = 0;
        try {
            for (uint32_t h = 0; h < H; ++h) {
                const uint64_t n_h = slice_start[h + 1] - slice_start[h];
                if (n_h == 0) continue;
                if (h) timer.start();
                const ArenaMark slice_mark = arena_mark();
                slice_extract_kernel<<<nb, kNT, 0, s>>>(st.keys.p, st.vals.p, M, hshift, h, nb, counts.p, A,
                                                        reinterpret_cast<unsigned int*>(tot.p + 3));
                SW_CUDA(cudaGetLastError());
                const NbrBuffers* R = nullptr;
                tm.launches += 1 + radix_partition_nbr(nullptr, nullptr, &A, n_h, key_bits, P - hb, A, B, s, &R, nullptr);
                SliceJob J{};
                J.R = R;
                J.Dd = R == &A ? &B : &A;
                J.n = n_h;
                J.n_buckets = 1ull << (P - hb);
                J.bucket0 = (uint64_t)h * J.n_buckets;
                J.key_bits = key_bits;
                J.kmer_base = slice_start[h];
                J.node_base = node_base;
                J.edge_base = edge_base;
                J.exact = false;
                J.node_cap = node_cap;
                J.edge_cap = edge_cap;
                uint64_t nn = 0, ne = 0;
                SW_CUDA(cudaMemsetAsync(tot.p, 0, 3 * sizeof(unsigned long long), s));
                aggregate_buckets(J, item_rank.p, d_rec_asm, rec_base, s, g, tm, timer, eg, score, nullptr, tot.p, &nn, &ne, &fallback);
                if (fallback) break;
                node_base += nn;
                edge_base += ne;
                arena_release(slice_mark);
            }
        } catch (const CapacityExceeded&) {
            if (attempt >= 3) fail_runtime("sliced aggregation: output estimate exceeded repeatedly");
            slack *= 2.0;
            SW_CUDA(cudaStreamSynchronize(s));
            continue;
        }
        n_nodes = node_base;
        n_edges = edge_base;
        break;
    }
    if (fallback) {
        SW_CUDA(cudaStreamSynchronize(s));
        g.kmers.release();
        g.nodes.release();
        g.edges.release();
        return false;
    }
    g.n_nodes = n_nodes;
    g.n_edges = n_edges;
    if (after_nodes) (*after_nodes)();
    if (getenv("SEQWIN_DEBUG_AGG"))
        fprintf(stderr, "[agg] %u hash slices: M %llu nodes %llu edges %llu\n", H, (unsigned long long)M, (unsigned long long)n_nodes,
                (unsigned long long)n_edges);
    if (times) *times = tm;
    return true;
}

}  // namespace

void route_stream(SketchStream& st, cudaStream_t s, RoutedStream& out)
{
    using namespace agg;
    const uint64_t M = st.n;
    out.n = M;
    out.items_per_key = st.items_per_key;
    out.pairs_per_edge = st.pairs_per_edge;
    out.set.alloc(4 * M, s);
    std::fill(out.byte_off, out.byte_off + 257, 0ull);
    if (M == 0) return;
    if (M > 0xFFFFFFFFull) fail_runtime("more than 2^32-1 minimizers on one device");
    const NbrBuffers A{out.set.p, out.set.p + M, out.set.p + 2 * M, out.set.p + 3 * M};
    DevBuf<unsigned long long> flag(1, s, true);
    SW_CUDA(cudaMemsetAsync(flag.p, 0, sizeof(unsigned long long), s));
    const NbrBuffers* R = nullptr;
    out.launches += radix_partition_nbr(st.keys.p, st.vals.p, nullptr, M, 56, 8, A, A, s, &R, reinterpret_cast<unsigned int*>(flag.p));
    DevBuf<uint32_t> start(257, s, true);
    DevBuf<unsigned long long> start64(257, s, true);
    bucket_search_kernel<<<2, 256, 0, s>>>(R->keys, M, 56, 256, start.p);
    widen_u32_kernel<<<2, 256, 0, s>>>(start.p, 257, start64.p);
    SW_CUDA(cudaGetLastError());
    out.launches += 2;
    const unsigned long long* h = readback_u64(start64.p, 257, s);
    const unsigned long long* hz = readback_u64(flag.p, 1, s);
    SW_CUDA(cudaStreamSynchronize(s));
    std::copy(h, h + 257, out.byte_off);
    out.zero_key = (*hz & 0xFFFFFFFFull) != 0;
}

void aggregate_range(const NbrBuffers& in, uint64_t n, uint32_t byte_lo, uint32_t byte_hi, const uint32_t* d_rec_asm,
                     cudaStream_t s, DevGraph& g, GraphTimes* times, const ScoreArgs* score, double pairs_per_edge,
                     const uint64_t* byte_off, int range_bits)
{
    using namespace agg;
    GraphTimes tm;
    g.n_kmers = n;
    g.n_nodes = g.n_edges = 0;
    if (n == 0 || byte_hi <= byte_lo) {
        g.kmers.alloc(0, s);
        g.nodes.alloc(0, s);
        g.edges.alloc(0, s);
        if (times) *times = tm;
        return;
    }
    if (n > 0xFFFFFFFFull) fail_runtime("more than 2^32-1 minimizers in one hash range");
    EventTimer timer(s);
    timer.start();
    DevBuf<unsigned long long> sample_set(1ull << kSampleSetBits, s, true), sample_out(2, s, true);
    const double per_node = estimate_items_per_key(in.keys, n, sample_set.p, sample_out.p, s);
    // bucket bits as if the whole hash space were this dense; at least the bits the range is cut on
    const uint32_t width = byte_hi - byte_lo;
    const EdgeGeom* egp = nullptr;
    int P = choose_bucket_bits((uint64_t)((double)n * (double)(1u << range_bits) / (double)width), per_node, pairs_per_edge > 0 ? pairs_per_edge : 1.0, &egp);
    if (const uint32_t fixed_nb = env_u32("SEQWIN_AGG_NODE_BUCKET", 0)) P = partition_bits((uint64_t)((double)n * (double)(1u << range_bits) / (double)width), fixed_nb);
    P = std::max(P, range_bits);
    const int key_bits = 64 - P;
    int top_bits = 0;
    while ((1u << top_bits) < width) ++top_bits;
    DevBuf<uint64_t> setA(4 * n, s, true), setB(4 * n, s, true);
    const NbrBuffers A{setA.p, setA.p + n, setA.p + 2 * n, setA.p + 3 * n};
    const NbrBuffers B{setB.p, setB.p + n, setB.p + 2 * n, setB.p + 3 * n};
    DevBuf<uint16_t> item_rank(n, s, true);
    DevBuf<unsigned long long> tot(4, s, true);
    SW_CUDA(cudaMemsetAsync(tot.p, 0, 4 * sizeof(unsigned long long), s));
    const NbrBuffers* R = nullptr;
    if (byte_off) {
        // partitioned on the top byte already: the lower bucket bits, one top-byte segment after the other (every
        // segment takes the same number of passes, so all of them end up in the same set)
        R = &in;
        for (uint32_t b = 0; b < width; ++b) {
            const uint64_t o = byte_off[b], nb_items = byte_off[b + 1] - o;
            if (nb_items == 0) continue;
            const NbrBuffers inV{in.keys + o, in.vals + o, in.prev + o, in.next + o};
            const NbrBuffers AV{A.keys + o, A.vals + o, A.prev + o, A.next + o}, BV{B.keys + o, B.vals + o, B.prev + o, B.next + o};
            const NbrBuffers* Rv = nullptr;
            tm.launches += radix_partition_nbr(nullptr, nullptr, &inV, nb_items, key_bits, P - range_bits, AV, BV, s, &Rv, nullptr);
            R = Rv == &inV ? &in : (Rv == &AV ? &A : &B);
        }
    } else {
        // the bits below the top byte and the top byte itself, taken relative to the range's first value so that it
        // needs no more digits than the range is wide
        tm.launches += 1 + radix_partition_nbr(nullptr, nullptr, &in, n, key_bits, P - range_bits + top_bits, A, B, s, &R, nullptr,
                                               (uint64_t)byte_lo << (64 - range_bits));
    }
    const NbrBuffers* Dd = R == &A ? &B : &A;
    if (R == &in) Dd = &A;   // no pass ran (one bucket): the input is the result, any set is free
    SliceJob J{};
    J.R = R;
    J.Dd = Dd;
    J.n = n;
    J.n_buckets = (uint64_t)width << (P - range_bits);
    J.bucket0 = (uint64_t)byte_lo << (P - range_bits);
    J.key_bits = key_bits;
    J.exact = true;
    bool fallback = false;
    uint64_t nn = 0, ne = 0;
    aggregate_buckets(J, item_rank.p, d_rec_asm, 0, s, g, tm, timer, *egp, score, nullptr, tot.p, &nn, &ne, &fallback);
    if (fallback) fail_runtime("a node bucket of the hash range overflowed (more distinct hashes than its table takes)");
    if (times) *times = tm;
}

void build_graph(SketchStream& st, const uint32_t* d_rec_asm, uint32_t rec_base, cudaStream_t s, DevGraph& g,
                 GraphTimes* times, const std::function<void()>* after_nodes, const ScoreArgs* score)
{
    const char* mode = getenv("SEQWIN_AGG");
    const bool legacy = mode && !strcmp(mode, "legacy");
    if (!legacy && st.n != 0 && st.n <= 0xFFFFFFFFull) {
        if (build_graph_bucketed(st, d_rec_asm, rec_base, s, g, times, after_nodes, score)) return;
        if (getenv("SEQWIN_DEBUG_AGG")) fprintf(stderr, "[agg] a node bucket overflowed: sort-based path\n");
    }
    build_graph_legacy(st, d_rec_asm, rec_base, s, g, times, after_nodes, score);
}

namespace {

void build_graph_legacy(SketchStream& st, const uint32_t* d_rec_asm, uint32_t rec_base, cudaStream_t s, DevGraph& g,
                        GraphTimes* times, const std::function<void()>* after_nodes, const ScoreArgs* score)
{
    const uint64_t M = st.n;
    g.n_kmers = M;
    g.n_nodes = 0;
    g.n_edges = 0;
    GraphTimes tm;
    EventTimer timer(s);
    if (M == 0) {
        g.kmers.alloc(0, s);
        g.nodes.alloc(0, s);
        g.edges.alloc(0, s);
        if (times) *times = tm;
        return;
    }
    if (M > 0xFFFFFFFFull) fail_runtime("more than 2^32-1 minimizers on one device");

    // The number of adjacent-pair records only depends on the stream, so it is counted now and read
    // at the node stage's synchronisation point: the edge stage can then start without one of its own.
    const uint32_t nb = blocks_for(M);
    DevBuf<unsigned long long> ecnt((size_t)nb + 1, s, true);
    edge_count_kernel<<<nb, kNT, 0, s>>>(st.vals.p, M, ecnt.p);
    exclusive_scan_u64(ecnt.p, nb, ecnt.p + nb, s);
    SW_CUDA(cudaGetLastError());
    const unsigned long long* n_raw_p = readback_u64(ecnt.p + nb, 1, s);
    tm.launches += 2;

    // -- sort (h1, stream index): high word in 4 passes + tie fix-up (all 64 bits if that gives up) --
    SortPairs sp;
    DevBuf<unsigned long long> counts((size_t)nb + 1, s, true);
    TieFix tie(s);
    const int begin_bit = sort_begin_bit();
    unsigned long long n_nodes = 0;
    for (bool full = begin_bit == 0;; full = true) {
        timer.start();
        sp.n = M;
        sp.keys.alloc(M, s, true);
        sp.vals.alloc(M, s, true);
        SW_CUDA(cudaMemcpyAsync(sp.keys.p, st.keys.p, M * sizeof(uint64_t), cudaMemcpyDeviceToDevice, s));
        iota_kernel<<<(uint32_t)std::min<uint64_t>((M + 255) / 256, 65535), 256, 0, s>>>(sp.vals.p, M);
        SW_CUDA(cudaGetLastError());
        tm.launches += 1 + tie.sort(sp, begin_bit, 0, full, s);
        tm.sort_nodes_ms += timer.stop();

        // -- nodes + kmers -----------------------------------------------------------------------
        timer.start();
        key_run_count_kernel<<<nb, kNT, 0, s>>>(sp.keys.p, M, counts.p);
        exclusive_scan_u64(counts.p, nb, counts.p + nb, s);
        SW_CUDA(cudaGetLastError());
        const unsigned long long* n_nodes_p = readback_u64(counts.p + nb, 1, s);
        const unsigned long long* tie_p = readback_u64(tie.counters.p, 1, s);
        SW_CUDA(cudaStreamSynchronize(s));
        n_nodes = *n_nodes_p;
        if (getenv("SEQWIN_DEBUG_TIE")) fprintf(stderr, "[tie] nodes: M %llu boundaries %u too_large %u\n", (unsigned long long)M, (unsigned)*tie_p, (unsigned)(*tie_p >> 32));
        if (full || !TieFix::gave_up(*tie_p)) break;
        tm.nodes_ms += timer.stop();   // too many / too large groups of equal high words: sort everything
    }
    g.n_nodes = n_nodes;
    g.kmers.alloc(M, s);
    g.nodes.alloc(n_nodes, s);
    DevBuf<uint32_t> rank_of_stream(M, s, true);
    DevBuf<uint64_t> node_hash(n_nodes, s, true);
    if (score) {
        SW_CUDA(cudaMemsetAsync(g.nodes.p, 0, n_nodes * sizeof(sw_node), s));
        node_write_kernel<true><<<nb, kNT, 0, s>>>(sp.keys.p, sp.vals.p, st.vals.p, M, counts.p, g.kmers.p, g.nodes.p,
                                                   node_hash.p, rank_of_stream.p, d_rec_asm, rec_base, score->d_is_target);
        if (!score->counts_only) {
            finish_penalty(g.nodes.p, n_nodes, score->inv_t, score->inv_n, s);
            ++tm.launches;
        }
    } else {
        node_write_kernel<false><<<nb, kNT, 0, s>>>(sp.keys.p, sp.vals.p, st.vals.p, M, counts.p, g.kmers.p, g.nodes.p,
                                                    node_hash.p, rank_of_stream.p, nullptr, 0u, nullptr);
    }
    SW_CUDA(cudaGetLastError());
    tm.launches += 3;
    timer.mark();

    // -- edges -----------------------------------------------------------------------------------
    // The first kernel of the edge stage is enqueued before the nodes-ready hook runs, so that the GPU
    // has work while a multi-GPU caller cuts and ships the node arrays from the host.
    EventTimer etimer(s);
    etimer.start();
    const unsigned long long n_raw = *n_raw_p;   // valid since the node stage's synchronisation
    int rank_bits = 1;
    while (rank_bits < 32 && (1ull << rank_bits) < n_nodes) ++rank_bits;
    if (n_raw) {
        sp.n = n_raw;
        edge_write_kernel<<<nb, kNT, 0, s>>>(st.vals.p, rank_of_stream.p, M, d_rec_asm, rec_base, ecnt.p, rank_bits,
                                             sp.keys.p, sp.vals.p);
        SW_CUDA(cudaGetLastError());
        ++tm.launches;
    }
    if (after_nodes) (*after_nodes)();
    tm.nodes_ms += timer.read();
    if (n_raw == 0) {
        g.edges.alloc(0, s);
    } else {
        // reuse the node sort's buffers for the (rank pair, assembly) sort.  The key holds the two
        // ranks left-aligned (2 * rank_bits significant bits); the radix passes look at its high word
        // and the tie fix-up finishes the few pairs that agree there.
        const int full_begin_bit = (64 - 2 * rank_bits) & ~7;
        // first + 16 bits of second: pairs that still agree there are rare enough for the fix-up (with
        // only the high word, 10 bits of `second` at 4 M nodes, the C2 workload leaves 228 k boundaries)
        const int edge_begin_bit = getenv("SEQWIN_SORT_BEGIN_BIT") ? begin_bit
                                                                    : std::min(begin_bit, std::max(0, 64 - rank_bits - 16) & ~7);
        const uint32_t eb = blocks_for(n_raw);
        DevBuf<unsigned long long> ecounts((size_t)eb + 1, s, true);
        unsigned long long n_edges = 0;
        bool refill = false;
        for (bool full = edge_begin_bit <= full_begin_bit;; full = true, refill = true) {
            sp.n = n_raw;
            if (refill) {   // the fix-up gave up: write the records again and sort every significant bit
                edge_write_kernel<<<nb, kNT, 0, s>>>(st.vals.p, rank_of_stream.p, M, d_rec_asm, rec_base, ecnt.p, rank_bits,
                                                     sp.keys.p, sp.vals.p);
                SW_CUDA(cudaGetLastError());
                ++tm.launches;
            }
            tm.launches += tie.sort(sp, edge_begin_bit, full_begin_bit, full, s);
            key_run_count_kernel<<<eb, kNT, 0, s>>>(sp.keys.p, n_raw, ecounts.p);
            exclusive_scan_u64(ecounts.p, eb, ecounts.p + eb, s);
            SW_CUDA(cudaGetLastError());
            const unsigned long long* n_edges_p = readback_u64(ecounts.p + eb, 1, s);
            const unsigned long long* tie_p = readback_u64(tie.counters.p, 1, s);
            SW_CUDA(cudaStreamSynchronize(s));
            n_edges = *n_edges_p;
            if (getenv("SEQWIN_DEBUG_TIE")) fprintf(stderr, "[tie] edges: n_raw %llu boundaries %u too_large %u rank_bits %d\n", n_raw, (unsigned)*tie_p, (unsigned)(*tie_p >> 32), rank_bits);
            if (full || !TieFix::gave_up(*tie_p)) break;
        }
        g.n_edges = n_edges;
        g.edges.alloc(n_edges, s);
        SW_CUDA(cudaMemsetAsync(g.edges.p, 0, n_edges * sizeof(sw_edge), s));
        edge_final_kernel<<<eb, kNT, 0, s>>>(sp.keys.p, sp.vals.p, n_raw, ecounts.p, rank_bits, node_hash.p, g.edges.p);
        SW_CUDA(cudaGetLastError());
        tm.launches += 3;
    }
    tm.edges_ms = etimer.stop();
    if (times) *times = tm;
}

}  // namespace

void finish_penalty(sw_node* d_nodes, uint64_t n_nodes, double inv_t, double inv_n, cudaStream_t s)
{
    if (n_nodes == 0) return;
    penalty_finish_kernel<<<(uint32_t)((n_nodes + 255) / 256), 256, 0, s>>>(d_nodes, n_nodes, inv_t, inv_n);
    SW_CUDA(cudaGetLastError());
}

uint32_t run_penalty(const sw_kmer* d_kmers, uint64_t n_kmers, sw_node* d_nodes, uint64_t n_nodes,
                     const uint32_t* d_rec_asm, uint32_t n_records, const uint8_t* d_is_target, double inv_t,
                     double inv_n, cudaStream_t s)
{
    if (n_nodes == 0) return 0;
    DevBuf<unsigned long long> err64(1, s, true);
    SW_CUDA(cudaMemsetAsync(err64.p, 0, sizeof(unsigned long long), s));
    uint32_t* err_p = reinterpret_cast<uint32_t*>(err64.p);
    constexpr int kLanes = 8;
    const uint32_t grid = (uint32_t)std::min<uint64_t>((n_nodes * kLanes + 255) / 256, (uint64_t)sm_count() * 16);
    penalty_kernel<kLanes><<<grid, 256, 0, s>>>(d_kmers, n_kmers, d_nodes, n_nodes, d_rec_asm, n_records, d_is_target,
                                        inv_t, inv_n, err_p);
    SW_CUDA(cudaGetLastError());
    const unsigned long long* h = readback_u64(err64.p, 1, s);
    SW_CUDA(cudaStreamSynchronize(s));
    return (uint32_t)*h;
}

}  // namespace sw
