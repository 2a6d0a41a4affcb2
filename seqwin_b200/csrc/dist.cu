// dist.cu -- owner-side merge of per-GPU shard graphs (multi-GPU build).
//
// Every rank builds the graph of its own assemblies with the single-GPU kernels.  Because an
// assembly lives on exactly one rank, a node's k-mer list is the concatenation of the per-rank lists
// in rank order and an edge's weight is the sum of the per-rank weights -- the same rule the reference
// uses to merge its per-thread graphs (cpp/src/seqwin/build_internals.cpp:159-291, merge_nodes /
// merge_kmers / merge_edges).  Ranks exchange hash-range slices of their sorted node / k-mer / edge
// arrays (NCCL all-to-all, seqwin_b200/dist.py); this file merges what a range owner received:
//   nodes   sort the (much shorter) node list by hash, run-length merge, segment-copy the k-mers
//   edges   stable two-key LSD sort by (first, second), run-length merge, sum the weights
#include <cuda_runtime.h>

#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <memory>

#include "device.h"
#include "scan.cuh"

namespace sw {

namespace {

constexpr int kNT = 256;
constexpr int kMergeVT = 8;  // outputs per thread in the merge kernel

// ---- pairwise merge of sorted runs (merge path) ------------------------------------------------
// Each rank's slice arrives sorted, so the owner merges P sorted runs in ceil(log2 P) rounds of
// pairwise merges instead of re-sorting: one read + one write of the data per round.

struct KeyIdx {
    uint64_t key;  // node hash
    uint64_t idx;  // index into the concatenated received nodes
};
struct LessKey {
    __device__ __forceinline__ bool operator()(const KeyIdx& a, const KeyIdx& b) const { return a.key < b.key; }
};
struct LessEdge {
    __device__ __forceinline__ bool operator()(const sw_edge& a, const sw_edge& b) const
    {
        return a.first < b.first || (a.first == b.first && a.second < b.second);
    }
};

// number of elements taken from A among the first `diag` outputs of merge(A, B); ties take A first,
// which keeps equal keys in source-rank order
template <typename T, typename Less>
__device__ __forceinline__ uint64_t merge_path(const T* __restrict__ A, uint64_t na, const T* __restrict__ B,
                                               uint64_t nb, uint64_t diag, Less less)
{
    uint64_t lo = diag > nb ? diag - nb : 0, hi = diag < na ? diag : na;
    while (lo < hi) {
        const uint64_t mid = (lo + hi) >> 1;
        if (!less(B[diag - 1 - mid], A[mid])) lo = mid + 1; else hi = mid;
    }
    return lo;
}

template <typename T, typename Less>
__global__ void __launch_bounds__(kNT) merge_pair_kernel(const T* __restrict__ A, uint64_t na, const T* __restrict__ B,
                                                         uint64_t nb, T* __restrict__ out)
{
    Less less;
    const uint64_t n = na + nb;
    const uint64_t i0 = ((uint64_t)blockIdx.x * blockDim.x + threadIdx.x) * kMergeVT;
    if (i0 >= n) return;
    uint64_t a = merge_path(A, na, B, nb, i0, less);
    uint64_t b = i0 - a;
#pragma unroll
    for (int t = 0; t < kMergeVT; ++t) {
        if (i0 + t >= n) break;
        const bool take_a = b >= nb || (a < na && !less(B[b], A[a]));
        out[i0 + t] = take_a ? A[a] : B[b];
        if (take_a) ++a; else ++b;
    }
}

// Merge the runs [seg[i], seg[i+1]) of `src` pairwise until one run is left.  buf0 / buf1 are
// scratch of the same total size; returns the buffer holding the result.
template <typename T, typename Less>
const T* merge_runs(const T* src, std::vector<unsigned long long> seg, T* buf0, T* buf1, cudaStream_t s, uint32_t* launches)
{
    const T* cur = src;
    T* dst = buf0;
    while (seg.size() > 2) {
        std::vector<unsigned long long> next{seg.front()};
        for (size_t i = 0; i + 1 < seg.size(); i += 2) {
            const unsigned long long a0 = seg[i], a1 = seg[i + 1];
            const unsigned long long b1 = i + 2 < seg.size() ? seg[i + 2] : a1;
            const uint64_t na = a1 - a0, nb = b1 - a1, n = na + nb;
            if (n) {
                if (nb == 0 || na == 0) {
                    SW_CUDA(cudaMemcpyAsync(dst + a0, cur + a0, n * sizeof(T), cudaMemcpyDeviceToDevice, s));
                } else {
                    const uint64_t threads = (n + kMergeVT - 1) / kMergeVT;
                    merge_pair_kernel<T, Less><<<(uint32_t)((threads + kNT - 1) / kNT), kNT, 0, s>>>(cur + a0, na, cur + a1, nb,
                                                                                                 dst + a0);
                    SW_CUDA(cudaGetLastError());
                    ++*launches;
                }
            }
            next.push_back(b1);
        }
        seg.swap(next);
        cur = dst;
        dst = (dst == buf0) ? buf1 : buf0;
    }
    return cur;
}

// concatenated received nodes -> (hash, index) merge records and the absolute source offset of each
// node's k-mers
__global__ void node_prepare_kernel(const sw_node* __restrict__ nodes, uint64_t n, const unsigned long long* __restrict__ seg,
                                    uint32_t n_src, KeyIdx* __restrict__ rec, unsigned long long* __restrict__ abs_start)
{
    // seg layout: [0, n_src] node offsets, then n_src received-k-mer offsets, then n_src sender k-mer bases
    const unsigned long long* node_off = seg;
    const unsigned long long* kmer_off = seg + n_src + 1;
    const unsigned long long* kmer_base = kmer_off + n_src;
    const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
        uint32_t s = 0;
        while (s + 1 < n_src && node_off[s + 1] <= i) ++s;
        rec[i] = KeyIdx{nodes[i].hash, i};
        abs_start[i] = kmer_off[s] + (nodes[i].start - kmer_base[s]);
    }
}

// sorted order j -> (run start flag, k-mer count) as one packed u64: count in the low 40 bits,
// flag in bit 40, so that ONE exclusive scan yields both the output k-mer offset and the node rank
__global__ void node_pack_kernel(const KeyIdx* __restrict__ sorted, const sw_node* __restrict__ nodes, uint64_t n,
                                 unsigned long long* __restrict__ packed)
{
    const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    for (uint64_t j = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; j < n; j += stride) {
        const sw_node nd = nodes[sorted[j].idx];
        const unsigned long long flag = (j == 0 || sorted[j].key != sorted[j - 1].key) ? 1ULL : 0ULL;
        packed[j] = (nd.stop - nd.start) | (flag << 40);
    }
}

// 8 lanes per received node segment (segments hold a handful of k-mers): copy its k-mers to the
// merged position, fill merged nodes
__global__ void node_merge_kernel(const KeyIdx* __restrict__ sorted, const sw_node* __restrict__ nodes,
                                  const unsigned long long* __restrict__ abs_start,
                                  const unsigned long long* __restrict__ scanned, uint64_t n, unsigned long long total_packed,
                                  const sw_kmer* __restrict__ recv_kmers, sw_kmer* __restrict__ out_kmers,
                                  sw_node* __restrict__ out_nodes)
{
    constexpr int G = 8;
    const int sub = threadIdx.x & (G - 1);
    const uint64_t grp = ((uint64_t)blockIdx.x * blockDim.x + threadIdx.x) / G;
    const uint64_t n_grp = ((uint64_t)gridDim.x * blockDim.x) / G;
    const unsigned long long mask40 = (1ULL << 40) - 1;
    for (uint64_t j = grp; j < n; j += n_grp) {
        const uint64_t i = sorted[j].idx;
        const sw_node nd = nodes[i];
        const unsigned long long cnt = nd.stop - nd.start;
        const unsigned long long ex = scanned[j];
        const unsigned long long off = ex & mask40;
        const bool flag = (j == 0) || sorted[j].key != sorted[j - 1].key;
        // node rank = (#flags up to and including j) - 1
        const unsigned long long rank = (ex >> 40) + (flag ? 1 : 0) - 1;
        const unsigned long long src = abs_start[i];
        for (unsigned long long t = sub; t < cnt; t += G) out_kmers[off + t] = recv_kmers[src + t];
        if (sub == 0) {
            if (flag) {
                sw_node* o = out_nodes + rank;
                o->hash = nd.hash;
                o->start = off;
                if (rank > 0) out_nodes[rank - 1].stop = off;
            }
            // an assembly lives on one rank, so the shards' distinct-assembly counts add up (the host
            // zeroed the merged nodes; shards that were not scored carry zeros)
            const unsigned long long c = (unsigned long long)nd.n_tar | ((unsigned long long)nd.n_neg << 32);
            if (c) atomicAdd(reinterpret_cast<unsigned long long*>(&out_nodes[rank].n_tar), c);
            if (j == n - 1) out_nodes[rank].stop = total_packed & mask40;
        }
    }
}

__global__ void edge_flag_kernel(const sw_edge* __restrict__ sorted, uint64_t n, unsigned long long* __restrict__ flags)
{
    const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    for (uint64_t j = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; j < n; j += stride) {
        const bool f = j == 0 || sorted[j].first != sorted[j - 1].first || sorted[j].second != sorted[j - 1].second;
        flags[j] = f ? 1ULL : 0ULL;
    }
}

__global__ void edge_merge_kernel(const sw_edge* __restrict__ sorted, const unsigned long long* __restrict__ scanned,
                                  uint64_t n, sw_edge* __restrict__ out)
{
    const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    for (uint64_t j = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; j < n; j += stride) {
        const sw_edge e = sorted[j];
        const bool f = j == 0 || e.first != sorted[j - 1].first || e.second != sorted[j - 1].second;
        const unsigned long long rank = scanned[j] + (f ? 1 : 0) - 1;
        if (f) {
            out[rank].first = e.first;
            out[rank].second = e.second;
        }
        atomicAdd(reinterpret_cast<unsigned long long*>(&out[rank].weight), (unsigned long long)e.weight);
    }
}

// lower_bound of the P-1 interior hash boundaries in the sorted node / edge arrays
// (bounds: [0, P) node boundaries, [P, 2P) edge boundaries; entry 0 of each is unused)
__global__ void split_kernel(const sw_node* __restrict__ nodes, uint64_t n_nodes, uint64_t n_kmers,
                             const sw_edge* __restrict__ edges, uint64_t n_edges, uint32_t P,
                             const unsigned long long* __restrict__ bounds,
                             unsigned long long* __restrict__ out /* 3*(P+1) */)
{
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i > P) return;
    uint64_t lo_n = 0, hi_n = n_nodes, lo_e = 0, hi_e = n_edges;
    if (i == P) {
        lo_n = n_nodes;
        lo_e = n_edges;
    } else if (i > 0) {
        const uint64_t bn = bounds[i], be = bounds[P + i];
        while (lo_n < hi_n) {
            const uint64_t m = (lo_n + hi_n) >> 1;
            if (nodes[m].hash < bn) lo_n = m + 1; else hi_n = m;
        }
        while (lo_e < hi_e) {
            const uint64_t m = (lo_e + hi_e) >> 1;
            if (edges[m].first < be) lo_e = m + 1; else hi_e = m;
        }
    }
    out[i] = lo_n;
    out[(P + 1) + i] = lo_n < n_nodes ? nodes[lo_n].start : n_kmers;
    out[2 * (P + 1) + i] = lo_e;
}

uint32_t grid_for(uint64_t n) { return (uint32_t)std::min<uint64_t>(std::max<uint64_t>((n + kNT - 1) / kNT, 1), 148 * 16); }

}  // namespace

void graph_split(const DevGraph& g, uint32_t P, unsigned long long* host_out, cudaStream_t s)
{
    // Node hashes (h1, a multiplicative mix) are uniform: node range i starts at i * 2^64 / P.
    // An edge is owned through first = min(u, v), whose distribution is 1 - (1-x)^2: the matching
    // quantiles 1 - sqrt(1 - i/P) keep the edge ranges balanced.  Any monotone boundaries are
    // correct; every rank computes the same ones.
    std::vector<unsigned long long> hb(2 * (size_t)P, 0);
    for (uint32_t i = 1; i < P; ++i) {
        const unsigned __int128 full = (unsigned __int128)1 << 64;
        hb[i] = (unsigned long long)((full * i) / P);
        const long double q = 1.0L - sqrtl(1.0L - (long double)i / (long double)P);
        hb[P + i] = (unsigned long long)(q * 18446744073709551616.0L);
    }
    DevBuf<unsigned long long> d_bounds(hb.size(), s, true);
    SW_CUDA(cudaMemcpyAsync(d_bounds.p, hb.data(), hb.size() * sizeof(unsigned long long), cudaMemcpyHostToDevice, s));
    DevBuf<unsigned long long> d(3 * (size_t)(P + 1), s, true);
    split_kernel<<<(P + 1 + 63) / 64, 64, 0, s>>>(g.nodes.p, g.n_nodes, g.n_kmers, g.edges.p, g.n_edges, P, d_bounds.p, d.p);
    SW_CUDA(cudaGetLastError());
    const unsigned long long* h = readback_u64(d.p, d.n, s);
    SW_CUDA(cudaStreamSynchronize(s));
    std::copy(h, h + d.n, host_out);
}

void dist_merge_nodes(const sw_node* recv_nodes, const uint64_t* node_counts, const sw_kmer* recv_kmers,
                      const uint64_t* kmer_counts, const uint64_t* kmer_base, uint32_t n_src, cudaStream_t s, DevGraph& out,
                      uint32_t* launches)
{
    uint64_t Nn = 0, Nk = 0;
    std::vector<unsigned long long> seg(3 * (size_t)n_src + 1);
    for (uint32_t i = 0; i < n_src; ++i) {
        seg[i] = Nn;
        seg[n_src + 1 + i] = Nk;
        seg[2 * n_src + 1 + i] = kmer_base[i];
        Nn += node_counts[i];
        Nk += kmer_counts[i];
    }
    seg[n_src] = Nn;
    if (Nk >= (1ULL << 40)) fail_runtime("more than 2^40 k-mers in one hash range");
    uint32_t nl = 0;
    const bool prof = getenv("SEQWIN_DIST_PROFILE") != nullptr;
    cudaEvent_t pe[3];
    if (prof) for (auto& e : pe) cudaEventCreate(&e);
    auto mark = [&](int i) { if (prof) cudaEventRecord(pe[i], s); };
    mark(0);
    out.n_kmers = Nk;
    out.n_nodes = 0;
    out.kmers.alloc(Nk, s);
    if (Nn == 0) {
        out.nodes.alloc(0, s);
        mark(1);
    } else {
        DevBuf<unsigned long long> d_seg(seg.size(), s, true);
        SW_CUDA(cudaMemcpyAsync(d_seg.p, seg.data(), seg.size() * sizeof(unsigned long long), cudaMemcpyHostToDevice, s));
        DevBuf<KeyIdx> rec0(Nn, s, true), rec1(Nn, s, true), rec2(Nn, s, true);
        DevBuf<unsigned long long> abs_start(Nn, s, true);
        node_prepare_kernel<<<grid_for(Nn), kNT, 0, s>>>(recv_nodes, Nn, d_seg.p, n_src, rec0.p, abs_start.p);
        SW_CUDA(cudaGetLastError());
        ++nl;
        // every rank's slice is sorted by hash: merge the runs (ties keep source-rank order)
        const std::vector<unsigned long long> node_seg(seg.begin(), seg.begin() + n_src + 1);
        const KeyIdx* sorted = merge_runs<KeyIdx, LessKey>(rec0.p, node_seg, rec1.p, rec2.p, s, &nl);
        mark(1);
        DevBuf<unsigned long long> packed(Nn + 1, s, true);
        node_pack_kernel<<<grid_for(Nn), kNT, 0, s>>>(sorted, recv_nodes, Nn, packed.p);
        nl += exclusive_scan_u64(packed.p, Nn, packed.p + Nn, s);
        SW_CUDA(cudaGetLastError());
        const unsigned long long* total_p = readback_u64(packed.p + Nn, 1, s);
        SW_CUDA(cudaStreamSynchronize(s));
        const unsigned long long total = *total_p;
        out.n_nodes = total >> 40;
        out.nodes.alloc(out.n_nodes, s);
        SW_CUDA(cudaMemsetAsync(out.nodes.p, 0, out.n_nodes * sizeof(sw_node), s));
        const uint32_t grid = (uint32_t)std::min<uint64_t>((Nn + 31) / 32, 148 * 32);
        node_merge_kernel<<<grid, 256, 0, s>>>(sorted, recv_nodes, abs_start.p, packed.p, Nn, total, recv_kmers,
                                               out.kmers.p, out.nodes.p);
        SW_CUDA(cudaGetLastError());
        nl += 2;
    }
    mark(2);
    SW_CUDA(cudaStreamSynchronize(s));
    if (prof) {
        float a, b;
        cudaEventElapsedTime(&a, pe[0], pe[1]);
        cudaEventElapsedTime(&b, pe[1], pe[2]);
        fprintf(stderr, "[dist_merge] Nn=%llu Nk=%llu  node_sort %.2f  node_merge %.2f ms\n", (unsigned long long)Nn,
                (unsigned long long)Nk, a, b);
        for (auto& e : pe) cudaEventDestroy(e);
    }
    if (launches) *launches += nl;
}

void dist_merge_edges(const sw_edge* recv_edges, const uint64_t* edge_counts, uint32_t n_src, cudaStream_t s, DevGraph& out,
                      uint32_t* launches)
{
    uint64_t Ne = 0;
    for (uint32_t i = 0; i < n_src; ++i) Ne += edge_counts[i];
    uint32_t nl = 0;
    const bool prof = getenv("SEQWIN_DIST_PROFILE") != nullptr;
    cudaEvent_t pe[3];
    if (prof) for (auto& e : pe) cudaEventCreate(&e);
    auto mark = [&](int i) { if (prof) cudaEventRecord(pe[i], s); };
    mark(0);
    out.n_edges = 0;
    if (Ne == 0) {
        out.edges.alloc(0, s);
        mark(1);
    } else {
        // every rank's slice is sorted by (first, second): merge the runs, then sum equal pairs
        std::vector<unsigned long long> edge_seg(n_src + 1, 0);
        for (uint32_t i = 0; i < n_src; ++i) edge_seg[i + 1] = edge_seg[i] + edge_counts[i];
        DevBuf<sw_edge> e0(Ne, s, true), e1(Ne, s, true);
        const sw_edge* sorted = merge_runs<sw_edge, LessEdge>(recv_edges, edge_seg, e0.p, e1.p, s, &nl);
        mark(1);
        DevBuf<unsigned long long> flags(Ne + 1, s, true);
        edge_flag_kernel<<<grid_for(Ne), kNT, 0, s>>>(sorted, Ne, flags.p);
        nl += exclusive_scan_u64(flags.p, Ne, flags.p + Ne, s);
        SW_CUDA(cudaGetLastError());
        const unsigned long long* n_edges_p = readback_u64(flags.p + Ne, 1, s);
        SW_CUDA(cudaStreamSynchronize(s));
        const unsigned long long n_edges = *n_edges_p;
        out.n_edges = n_edges;
        out.edges.alloc(n_edges, s);
        SW_CUDA(cudaMemsetAsync(out.edges.p, 0, n_edges * sizeof(sw_edge), s));
        edge_merge_kernel<<<grid_for(Ne), kNT, 0, s>>>(sorted, flags.p, Ne, out.edges.p);
        SW_CUDA(cudaGetLastError());
        nl += 2;
    }
    mark(2);
    SW_CUDA(cudaStreamSynchronize(s));
    if (prof) {
        float a, b;
        cudaEventElapsedTime(&a, pe[0], pe[1]);
        cudaEventElapsedTime(&b, pe[1], pe[2]);
        fprintf(stderr, "[dist_merge] Ne=%llu  edge_sort %.2f  edge_merge %.2f ms\n", (unsigned long long)Ne, a, b);
        for (auto& e : pe) cudaEventDestroy(e);
    }
    if (launches) *launches += nl;
}

}  // namespace sw
