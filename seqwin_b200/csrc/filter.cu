// filter.cu -- the first consumers of the graph, on the device-resident arrays.
//
//   graph_filter_edges   remove low-weight edges and the nodes that are left without an edge
//                        (src/seqwin/kmers.py:132-162, _filter_edges_and_nodes: edges[weight > th],
//                        np.unique of the endpoints, nodes[searchsorted(hash, endpoints)])
//   graph_filter_kmers   keep the nodes whose hash is in a given set, compact their k-mers and
//                        rewrite start / stop (cpp/src/seqwin/filter.cpp:139-201, filter_kmers)
//
//   graph_count_sums     the three integer sums behind the automatic penalty threshold
//                        (src/seqwin/kmers.py:424-429: expected k-mer absence in targets / presence in
//                        non-targets, weighted by n_tar)
//
// The filters are stream compactions: flag, exclusive scan, ordered copy.  Node and edge order and every
// field that is not a k-mer range stay as they were.
#include <cuda_runtime.h>

#include <algorithm>
#include <vector>

#include "device.h"
#include "scan.cuh"

namespace sw {

namespace {

constexpr int kFT = 256;

uint32_t grid_of(uint64_t n) { return (uint32_t)std::min<uint64_t>(std::max<uint64_t>((n + kFT - 1) / kFT, 1), 148 * 32); }

// flag[j] = edge j survives; flag has n + 1 entries so that the scanned array ends with the total
__global__ void __launch_bounds__(kFT) edge_flag_kernel(const sw_edge* __restrict__ edges, uint64_t n, uint64_t th,
                                                        unsigned long long* __restrict__ flag)
{
    const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    for (uint64_t j = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; j <= n; j += stride)
        flag[j] = (j < n && edges[j].weight > th) ? 1ull : 0ull;
}

template <typename T>
__global__ void __launch_bounds__(kFT) compact_kernel(const T* __restrict__ in, uint64_t n,
                                                      const unsigned long long* __restrict__ pos, T* __restrict__ out)
{
    const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    for (uint64_t j = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; j < n; j += stride)
        if (pos[j + 1] != pos[j]) out[pos[j]] = in[j];
}

__global__ void __launch_bounds__(kFT) node_hash_kernel(const sw_node* __restrict__ nodes, uint64_t n,
                                                        uint64_t* __restrict__ hash)
{
    const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) hash[i] = nodes[i].hash;
}

// table[b] = first node whose hash >> (64 - bits) is >= b, for b in [0, 2^bits]  (hashes ascending)
__global__ void __launch_bounds__(kFT) bucket_fill_kernel(const uint64_t* __restrict__ hash, uint64_t n, int bits,
                                                          uint32_t* __restrict__ table)
{
    const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    const uint64_t n_buckets = 1ull << bits;
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i <= n; i += stride) {
        const uint64_t lo = i ? (hash[i - 1] >> (64 - bits)) + 1 : 0;          // buckets after the previous node's
        const uint64_t hi = i < n ? (hash[i] >> (64 - bits)) : n_buckets;      // ... up to this node's (or the end)
        for (uint64_t b = lo; b <= hi; ++b) table[b] = (uint32_t)i;
    }
}

__device__ __forceinline__ uint64_t find_node(uint64_t h, const uint64_t* __restrict__ hash, uint64_t n, int bits,
                                              const uint32_t* __restrict__ table)
{
    uint64_t i = table[h >> (64 - bits)];
    while (i < n && hash[i] < h) ++i;
    return (i < n && hash[i] == h) ? i : n;
}

// mark the endpoints of the surviving edges (mark has n_nodes + 1 entries, zeroed)
__global__ void __launch_bounds__(kFT) mark_endpoints_kernel(const sw_edge* __restrict__ edges, uint64_t n_edges,
                                                             const uint64_t* __restrict__ hash, uint64_t n_nodes, int bits,
                                                             const uint32_t* __restrict__ table,
                                                             unsigned long long* __restrict__ mark, unsigned int* err)
{
    const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    for (uint64_t j = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; j < n_edges; j += stride) {
        const uint64_t a = find_node(edges[j].first, hash, n_nodes, bits, table);
        const uint64_t b = find_node(edges[j].second, hash, n_nodes, bits, table);
        if (a == n_nodes || b == n_nodes) { *err = 1u; continue; }
        mark[a] = 1ull;
        mark[b] = 1ull;
    }
}

// keep[i] = node i's hash is in the sorted set; size[i] = its k-mer count if kept (both n + 1 long)
__global__ void __launch_bounds__(kFT) node_used_kernel(const sw_node* __restrict__ nodes, uint64_t n, uint64_t n_kmers,
                                                        const uint64_t* __restrict__ used, uint64_t n_used,
                                                        unsigned long long* __restrict__ keep,
                                                        unsigned long long* __restrict__ size, unsigned int* err)
{
    const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i <= n; i += stride) {
        unsigned long long k = 0, sz = 0;
        if (i < n) {
            const uint64_t h = nodes[i].hash;
            uint64_t lo = 0, hi = n_used;
            while (lo < hi) {
                const uint64_t m = (lo + hi) >> 1;
                if (used[m] < h) lo = m + 1; else hi = m;
            }
            if (lo < n_used && used[lo] == h) {
                const uint64_t a = nodes[i].start, b = nodes[i].stop;
                if (a > b || b > n_kmers) *err = 1u;
                else { k = 1; sz = b - a; }
            }
        }
        keep[i] = k;
        size[i] = sz;
    }
}

// 8 lanes per node: write the kept node with its new range and copy its k-mers
__global__ void __launch_bounds__(kFT) kmer_gather_kernel(const sw_node* __restrict__ nodes, uint64_t n,
                                                          const unsigned long long* __restrict__ keep_pos,
                                                          const unsigned long long* __restrict__ size_pos,
                                                          const sw_kmer* __restrict__ kmers_in,
                                                          sw_node* __restrict__ nodes_out, sw_kmer* __restrict__ kmers_out)
{
    constexpr int kLanes = 8;
    const int lane = threadIdx.x & (kLanes - 1);
    const uint64_t n_grp = ((uint64_t)gridDim.x * blockDim.x) / kLanes;
    for (uint64_t i = ((uint64_t)blockIdx.x * blockDim.x + threadIdx.x) / kLanes; i < n; i += n_grp) {
        if (keep_pos[i + 1] == keep_pos[i]) continue;
        const sw_node nd = nodes[i];
        const uint64_t dst = size_pos[i], cnt = nd.stop - nd.start;
        if (lane == 0) {
            sw_node out = nd;
            out.start = dst;
            out.stop = dst + cnt;
            nodes_out[keep_pos[i]] = out;
        }
        for (uint64_t t = lane; t < cnt; t += kLanes) kmers_out[dst + t] = kmers_in[nd.start + t];
    }
}

// sums[0..2] += sum n_tar, sum n_tar^2, sum n_tar * n_neg over the nodes
__global__ void __launch_bounds__(kFT) count_sums_kernel(const sw_node* __restrict__ nodes, uint64_t n,
                                                         unsigned long long* __restrict__ sums)
{
    __shared__ unsigned long long s_part[3][kFT / 32];
    unsigned long long a = 0, b = 0, c = 0;
    const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
        const unsigned long long t = nodes[i].n_tar, g = nodes[i].n_neg;
        a += t;
        b += t * t;
        c += t * g;
    }
#pragma unroll
    for (int d = 16; d; d >>= 1) {
        a += __shfl_xor_sync(0xffffffffu, a, d);
        b += __shfl_xor_sync(0xffffffffu, b, d);
        c += __shfl_xor_sync(0xffffffffu, c, d);
    }
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    if (lane == 0) { s_part[0][wid] = a; s_part[1][wid] = b; s_part[2][wid] = c; }
    __syncthreads();
    if (threadIdx.x < 3) {
        unsigned long long t = 0;
        for (int i = 0; i < kFT / 32; ++i) t += s_part[threadIdx.x][i];
        atomicAdd(&sums[threadIdx.x], t);
    }
}

unsigned long long read_total(const unsigned long long* d, cudaStream_t s)
{
    const unsigned long long* h = readback_u64(d, 1, s);
    SW_CUDA(cudaStreamSynchronize(s));
    return *h;
}

}  // namespace

void graph_filter_edges(DevGraph& g, uint64_t weight_th, cudaStream_t s)
{
    const uint64_t n_e = g.n_edges, n_n = g.n_nodes;
    DevBuf<unsigned long long> total(1, s, true);
    // edges with weight > th, order kept
    DevBuf<unsigned long long> eflag(n_e + 1, s, true);
    edge_flag_kernel<<<grid_of(n_e + 1), kFT, 0, s>>>(g.edges.p, n_e, weight_th, eflag.p);
    exclusive_scan_u64(eflag.p, n_e + 1, total.p, s);
    SW_CUDA(cudaGetLastError());
    const uint64_t n_keep = read_total(total.p, s);
    DevBuf<sw_edge> edges_new;
    edges_new.alloc(n_keep, s);
    if (n_e) compact_kernel<sw_edge><<<grid_of(n_e), kFT, 0, s>>>(g.edges.p, n_e, eflag.p, edges_new.p);
    // nodes that are an endpoint of a surviving edge
    DevBuf<unsigned long long> mark(n_n + 1, s, true);
    SW_CUDA(cudaMemsetAsync(mark.p, 0, mark.bytes(), s));
    DevBuf<unsigned long long> err64(1, s, true);
    SW_CUDA(cudaMemsetAsync(err64.p, 0, sizeof(unsigned long long), s));
    unsigned int* err = reinterpret_cast<unsigned int*>(err64.p);
    if (n_keep && n_n) {
        int bits = 1;
        while (bits < 24 && (1ull << bits) < n_n) ++bits;
        DevBuf<uint64_t> hash(n_n, s, true);
        DevBuf<uint32_t> table((1ull << bits) + 1, s, true);
        node_hash_kernel<<<grid_of(n_n), kFT, 0, s>>>(g.nodes.p, n_n, hash.p);
        bucket_fill_kernel<<<grid_of(n_n + 1), kFT, 0, s>>>(hash.p, n_n, bits, table.p);
        mark_endpoints_kernel<<<grid_of(n_keep), kFT, 0, s>>>(edges_new.p, n_keep, hash.p, n_n, bits, table.p, mark.p, err);
    } else if (n_keep) {
        fail_value("edge endpoint is not a node hash");
    }
    exclusive_scan_u64(mark.p, n_n + 1, total.p, s);
    SW_CUDA(cudaGetLastError());
    const unsigned long long* h_total = readback_u64(total.p, 1, s);
    const unsigned long long* h_err = readback_u64(err64.p, 1, s);
    SW_CUDA(cudaStreamSynchronize(s));
    if (*h_err) fail_value("edge endpoint is not a node hash");
    const uint64_t n_nodes_keep = *h_total;
    DevBuf<sw_node> nodes_new;
    nodes_new.alloc(n_nodes_keep, s);
    if (n_n) compact_kernel<sw_node><<<grid_of(n_n), kFT, 0, s>>>(g.nodes.p, n_n, mark.p, nodes_new.p);
    SW_CUDA(cudaGetLastError());
    SW_CUDA(cudaStreamSynchronize(s));   // the scratch arrays go out of scope
    g.edges = std::move(edges_new);
    g.nodes = std::move(nodes_new);
    g.n_edges = n_keep;
    g.n_nodes = n_nodes_keep;
}

void graph_count_sums(const DevGraph& g, unsigned long long out[3], cudaStream_t s)
{
    DevBuf<unsigned long long> sums(3, s, true);
    SW_CUDA(cudaMemsetAsync(sums.p, 0, sums.bytes(), s));
    if (g.n_nodes) count_sums_kernel<<<grid_of(g.n_nodes), kFT, 0, s>>>(g.nodes.p, g.n_nodes, sums.p);
    SW_CUDA(cudaGetLastError());
    const unsigned long long* h = readback_u64(sums.p, 3, s);
    SW_CUDA(cudaStreamSynchronize(s));
    out[0] = h[0];
    out[1] = h[1];
    out[2] = h[2];
}

void graph_filter_kmers(DevGraph& g, const uint64_t* h_used, size_t n_used, cudaStream_t s)
{
    const uint64_t n_n = g.n_nodes;
    std::vector<uint64_t> used(h_used, h_used + n_used);
    std::sort(used.begin(), used.end());
    used.erase(std::unique(used.begin(), used.end()), used.end());
    DevBuf<uint64_t> d_used(used.size(), s, true);
    if (!used.empty())
        SW_CUDA(cudaMemcpyAsync(d_used.p, used.data(), used.size() * sizeof(uint64_t), cudaMemcpyHostToDevice, s));
    DevBuf<unsigned long long> keep(n_n + 1, s, true), size(n_n + 1, s, true), totals(2, s, true);
    DevBuf<unsigned long long> err64(1, s, true);
    SW_CUDA(cudaMemsetAsync(err64.p, 0, sizeof(unsigned long long), s));
    node_used_kernel<<<grid_of(n_n + 1), kFT, 0, s>>>(g.nodes.p, n_n, g.n_kmers, d_used.p, used.size(), keep.p, size.p,
                                                      reinterpret_cast<unsigned int*>(err64.p));
    exclusive_scan_u64(keep.p, n_n + 1, totals.p, s);
    exclusive_scan_u64(size.p, n_n + 1, totals.p + 1, s);
    SW_CUDA(cudaGetLastError());
    const unsigned long long* h_tot = readback_u64(totals.p, 2, s);
    const unsigned long long* h_err = readback_u64(err64.p, 1, s);
    SW_CUDA(cudaStreamSynchronize(s));   // also: `used` may be freed after this
    if (*h_err) fail_value("node [start, stop) range lies outside kmers");
    const uint64_t nn = h_tot[0], nk = h_tot[1];
    DevBuf<sw_node> nodes_new;
    DevBuf<sw_kmer> kmers_new;
    nodes_new.alloc(nn, s);
    kmers_new.alloc(nk, s);
    if (n_n)
        kmer_gather_kernel<<<grid_of(n_n * 8), kFT, 0, s>>>(g.nodes.p, n_n, keep.p, size.p, g.kmers.p, nodes_new.p, kmers_new.p);
    SW_CUDA(cudaGetLastError());
    SW_CUDA(cudaStreamSynchronize(s));
    g.nodes = std::move(nodes_new);
    g.kmers = std::move(kmers_new);
    g.n_nodes = nn;
    g.n_kmers = nk;
}

}  // namespace sw
