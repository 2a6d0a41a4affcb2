// device.h -- internal device-side interfaces of libseqwin_b200 (not part of the C ABI).
#pragma once

#include <cuda_runtime.h>

#include <functional>
#include <string>
#include <vector>

#include "common.h"
#include "ingest.h"

struct sw_dev_batch;

namespace sw {

#define SW_CUDA(expr)                                                                        \
    do {                                                                                     \
        cudaError_t err__ = (expr);                                                          \
        if (err__ != cudaSuccess)                                                            \
            ::sw::fail_runtime(std::string("CUDA error: ") + cudaGetErrorString(err__) +    \
                               " at " __FILE__ ":" + std::to_string(__LINE__));            \
    } while (0)

// Scratch arena: one grow-only device slab per host thread, bump-allocated and reset at the start
// of every build, so the steady state performs no cudaMalloc / cudaFree at all.
void* arena_alloc(size_t bytes);
void arena_reset();
void arena_trim();   // give every slab back to the driver
// Stack discipline inside one build: everything allocated after arena_mark() is handed back by
// arena_release(mark) (stream order keeps a later owner of the bytes behind the kernels of the earlier one).
struct ArenaMark { size_t slab, off; };
ArenaMark arena_mark();
void arena_release(const ArenaMark& m);
size_t arena_free_bytes();   // unused bytes of the slabs the arena already owns
// the caller asked for the reduced-footprint plan (thread-local; cpp/src/seqwin/build.cpp:264-325)
void set_low_memory(bool on);
bool low_memory();

// Small device -> host readbacks (counts, histograms) go through a pinned, device-mapped buffer
// written by a tiny kernel instead of cudaMemcpyAsync: a copy-engine transfer would queue behind
// the bulk graph export that the end-to-end path overlaps with the kernels.  The returned host
// pointer is valid once `s` has been synchronized and until the next arena_reset().
const unsigned long long* readback_u64(const unsigned long long* d_src, size_t count, cudaStream_t s);

// Device buffer.  Persistent buffers (graph outputs, uploaded batches) come from the stream-ordered
// cudaMallocAsync pool; temporaries (tmp = true) come from the scratch arena and are never freed
// individually.
template <typename T>
struct DevBuf {
    T* p = nullptr;
    size_t n = 0;
    cudaStream_t stream = nullptr;
    bool tmp = false;
    DevBuf() = default;
    DevBuf(size_t count, cudaStream_t s, bool temporary = false) { alloc(count, s, temporary); }
    DevBuf(const DevBuf&) = delete;
    DevBuf& operator=(const DevBuf&) = delete;
    DevBuf(DevBuf&& o) noexcept : p(o.p), n(o.n), stream(o.stream), tmp(o.tmp) { o.p = nullptr; o.n = 0; }
    DevBuf& operator=(DevBuf&& o) noexcept
    {
        if (this != &o) {
            release();
            p = o.p; n = o.n; stream = o.stream; tmp = o.tmp;
            o.p = nullptr; o.n = 0;
        }
        return *this;
    }
    ~DevBuf() { release(); }
    void alloc(size_t count, cudaStream_t s, bool temporary = false)
    {
        release();
        stream = s;
        n = count;
        tmp = temporary;
        if (tmp) p = static_cast<T*>(arena_alloc((count ? count : 1) * sizeof(T)));
        else SW_CUDA(cudaMallocAsync((void**)&p, (count ? count : 1) * sizeof(T), s));
    }
    void release()
    {
        if (p && !tmp) cudaFreeAsync(p, stream);
        p = nullptr;
        n = 0;
    }
    size_t bytes() const { return n * sizeof(T); }
};

// A (pinned) host buffer owned by the HostPool cache.
struct HostBuf {
    void* p = nullptr;
    size_t bytes = 0;
    bool pinned = false;
};
HostBuf host_pool_get(size_t bytes);
void host_pool_put(HostBuf& b);

void init_device_once();
int sm_count();

// ---- sketch stage ---------------------------------------------------------------------------
struct DevPlan {
    DevBuf<Tile> tiles;
    DevBuf<Piece> pieces;
    uint32_t n_tiles = 0;
    uint64_t n_windows = 0, n_kmers = 0;
    std::vector<unsigned long long> rec_tile_off;  // [R+1] first tile of each record (host, optional)
    uint32_t tk = 0;  // tile capacity the plan was cut for
    int config = 0;   // dense kernel configuration index
    bool sparse = false;  // sparse kernel first, the dense configuration only for the tiles it hands over
};

struct SketchStream {
    DevBuf<uint64_t> keys;  // h1, (record, pos) order
    DevBuf<uint64_t> vals;  // pos | record << 32
    uint64_t n = 0;
    uint32_t launches = 0;
    float kernel_ms = 0, reorder_ms = 0;  // CUDA-event times of the last run
    double items_per_key = 0;             // minimizers per distinct hash, estimated from a hash-range sample (0: unknown)
    double pairs_per_edge = 0;            // adjacent pairs per distinct pair, same kind of estimate (0: unknown)
};

int sketch_pick_config(uint32_t w, uint32_t* tk_out, bool* sparse_out);
// Device-side tile planner: cuts every record's valid-k-mer stream into tiles (see ingest.h:
// plan_tiles is the host statement of the same rule, used by the test emulator).
DevPlan make_plan(const sw_dev_batch& d, uint32_t k, uint32_t w, cudaStream_t s, bool want_rec_tile_off = false);
// A slice of tiles whose packed bases arrive with their own H2D copy (pipelined upload).
struct SketchChunk {
    uint32_t tile_lo, tile_hi;
    cudaEvent_t ready;  // recorded on the copy stream after the slice's bases landed (may be null)
};
void run_sketch(const uint32_t* d_words, const uint64_t* d_rec_word_off, const DevPlan& plan,
                uint32_t k, uint32_t w, uint32_t rec_base, cudaStream_t s, SketchStream& out,
                const std::vector<SketchChunk>* chunks = nullptr);

// ---- radix sort -------------------------------------------------------------------------------
// Stable LSD radix sort of (u64 key, u32 value) pairs over key bits [0, end_bit).
// On return keys/vals hold the sorted data (buffers may have been swapped with the alternates).
struct SortPairs {
    DevBuf<uint64_t> keys, keys_alt;
    DevBuf<uint32_t> vals, vals_alt;
    uint64_t n = 0;
};
// stable LSD sort of sp on key bits [begin_bit, end_bit) (begin_bit a multiple of 8); returns #launches
uint32_t radix_sort_pairs(SortPairs& sp, int end_bit, cudaStream_t s, int begin_bit = 0);

// Stable partition of (keys, vals) on the top `top_bits` bits of the key (1 <= top_bits <= 64): ceil(top_bits / 8)
// onesweep passes, lowest digit first.  The input is read only; the passes ping-pong between (ka, va) and
// (kb, vb) (each n entries) and *out_k / *out_v point at whichever holds the result.  V = uint32_t or
// unsigned long long.  Returns the number of kernels launched (one host synchronisation inside).
template <typename V>
uint32_t radix_partition_top(const uint64_t* keys, const V* vals, uint64_t n, int top_bits, uint64_t* ka, V* va, uint64_t* kb,
                             V* vb, cudaStream_t s, const uint64_t** out_k, const V** out_v);

// Node-stage partition: (h1, k-mer) plus, per minimizer, the hashes of the stream neighbours whose adjacent pair
// it OWNS (the pair belongs to the item with the smaller hash; 0 = none) -- what the edge grouping needs
// (agg.cuh).  Stable, on the key bits [lo_bit, lo_bit + n_bits).  in == nullptr: (keys, vals) is the stream and the
// first pass derives the two arrays from it (*d_zero_key is set if a key is 0: the marker would be ambiguous and
// the caller takes the sort-based path); otherwise `in` already holds all four arrays and (keys, vals) are not
// read.  A / B are the ping-pong sets (n entries per array; `in` may be one of them) and *out points at the set
// holding the result.  key_bias is subtracted from every key before its digits are taken (the items of a key range:
// its first key).  Returns the number of kernels launched; no host synchronisation.
struct NbrBuffers {
    uint64_t* keys;
    uint64_t* vals;
    uint64_t* prev;
    uint64_t* next;
};
uint32_t radix_partition_nbr(const uint64_t* keys, const uint64_t* vals, const NbrBuffers* in, uint64_t n, int lo_bit, int n_bits,
                             const NbrBuffers& A, const NbrBuffers& B, cudaStream_t s, const NbrBuffers** out,
                             unsigned int* d_zero_key, uint64_t key_bias = 0);

// Fused routing pass of the multi-GPU build (radix.cu): items per top byte of h1; then one generating pass whose
// scatter writes land in the arrays of the hash-range owners (d_route: 4 * 256 device pointers -- keys, vals, prev,
// next of the owner of every top byte, peer memory -- and d_base[256]: first position of this shard's items there).
void route_histogram(const uint64_t* keys, uint64_t n, int route_bits, unsigned long long* byte_counts_out, cudaStream_t s);
void route_scatter(const uint64_t* keys, const uint64_t* vals, uint64_t n, int route_bits, const unsigned long long* d_base,
                   uint64_t* const* d_route, unsigned int* d_zero_key, cudaStream_t s);

// ---- graph stage ------------------------------------------------------------------------------
struct DevGraph {
    DevBuf<sw_kmer> kmers;
    DevBuf<sw_node> nodes;
    DevBuf<sw_edge> edges;
    uint64_t n_kmers = 0, n_nodes = 0, n_edges = 0;
};
struct GraphTimes {
    float sort_nodes_ms = 0, nodes_ms = 0, edges_ms = 0;
    uint32_t launches = 0;
};
// d_rec_asm: assembly index of every global record id used in the stream.
// after_nodes (optional) is called once kmers + nodes are final and enqueued on s, before the edge
// stage starts: the end-to-end path uses it to begin their D2H copies on a second stream.
// score (optional): the classes of the assemblies are known at build time, so n_tar / n_neg are
// counted while the nodes are written and the penalty follows in one pass over the nodes
// (get_penalty, cpp/src/seqwin/filter.cpp:15-137, fused into the build).
struct ScoreArgs {
    const uint8_t* d_is_target;   // [n_assemblies], indexed by the values of d_rec_asm
    double inv_t, inv_n;          // 1 / #targets, 1 / #non-targets
    bool counts_only = false;     // a shard of a multi-GPU build: n_tar / n_neg only, the penalty follows the merge
};
// penalty of every node from its n_tar / n_neg (filter.cpp:132-134)
void finish_penalty(sw_node* d_nodes, uint64_t n_nodes, double inv_t, double inv_n, cudaStream_t s);
void build_graph(SketchStream& st, const uint32_t* d_rec_asm, uint32_t rec_base, cudaStream_t s, DevGraph& g,
                 GraphTimes* times, const std::function<void()>* after_nodes = nullptr,
                 const ScoreArgs* score = nullptr);

// ---- multi-GPU: route the stream to the owners of its hash ranges, aggregate a range (graph.cu) -------------------
// The minimizer stream of a shard with the owned neighbour hashes of every item (NbrBuffers), stably partitioned on
// the top byte of h1: byte_off[b] = first item whose top byte is >= b.  The four arrays are one pool allocation of
// 4 n words (keys | vals | prev | next), alive until the caller frees it: slices of it travel to the owners.
struct RoutedStream {
    DevBuf<uint64_t> set;
    uint64_t n = 0;
    unsigned long long byte_off[257] = {};
    bool zero_key = false;          // a hash is 0: the "no neighbour" marker is ambiguous, the caller must not go on
    double items_per_key = 0, pairs_per_edge = 0;
    uint32_t launches = 0;
};
void route_stream(SketchStream& st, cudaStream_t s, RoutedStream& out);
// Nodes, k-mers and edges of the items whose h1 has its top byte in [byte_lo, byte_hi) -- `in` holds exactly those,
// in global stream order (ties between sources: source order) -- with the single-GPU bucket kernels.  Fails loudly
// if a node bucket overflows (there is no sort-based path for a range).
// byte_off (host, byte_hi - byte_lo + 1 entries, may be null): the items arrive stably partitioned on the top byte
// (the fused routing pass delivers them so) and byte_off[i] is the first item of top byte byte_lo + i; the lower
// bucket bits are then partitioned segment by segment and the pass on the top byte is saved.
// range_bits: the hash space is cut into 2^range_bits bins (8: top bytes) and [byte_lo, byte_hi) are bin numbers.
void aggregate_range(const NbrBuffers& in, uint64_t n, uint32_t byte_lo, uint32_t byte_hi, const uint32_t* d_rec_asm,
                     cudaStream_t s, DevGraph& g, GraphTimes* times, const ScoreArgs* score, double pairs_per_edge,
                     const uint64_t* byte_off = nullptr, int range_bits = 8);

// ---- consumers of a device-resident graph (filter.cu), in place on g ---------------------------------
// edges with weight > weight_th and the nodes that keep an edge (kmers.py:132-162); k-mers untouched
void graph_filter_edges(DevGraph& g, uint64_t weight_th, cudaStream_t s);
// nodes whose hash is in used_hashes (host array, any order), their k-mers compacted (filter.cpp:139-201)
void graph_filter_kmers(DevGraph& g, const uint64_t* used_hashes, size_t n_used, cudaStream_t s);
// out = {sum n_tar, sum n_tar^2, sum n_tar * n_neg} over the nodes (kmers.py:424-429)
void graph_count_sums(const DevGraph& g, unsigned long long out[3], cudaStream_t s);

// ---- multi-GPU merge (dist.cu) -------------------------------------------------------------------
void graph_split(const DevGraph& g, uint32_t P, unsigned long long* host_out /* 3*(P+1) */, cudaStream_t s);
// the two halves of the owner-side merge: nodes + k-mers (their slices arrive first), then edges
void dist_merge_nodes(const sw_node* recv_nodes, const uint64_t* node_counts, const sw_kmer* recv_kmers,
                      const uint64_t* kmer_counts, const uint64_t* kmer_base, uint32_t n_src, cudaStream_t s, DevGraph& out,
                      uint32_t* launches);
void dist_merge_edges(const sw_edge* recv_edges, const uint64_t* edge_counts, uint32_t n_src, cudaStream_t s, DevGraph& out,
                      uint32_t* launches);

// ---- diagnostics (diag.cu) -----------------------------------------------------------------------
// lane-operations per second of dependent LOP3 / SHF / IADD chains and of their 1:1:1 mix
void measure_int_peak(double out[4], cudaStream_t s);

// ---- penalty ----------------------------------------------------------------------------------
// Fills n_tar / n_neg / penalty of device-resident nodes; returns an error bit mask
// (1: record_idx out of range, 2: record_idx decreasing, 4: node range outside kmers).
uint32_t run_penalty(const sw_kmer* d_kmers, uint64_t n_kmers, sw_node* d_nodes, uint64_t n_nodes,
                     const uint32_t* d_rec_asm, uint32_t n_records, const uint8_t* d_is_target,
                     double inv_t, double inv_n, cudaStream_t s);

}  // namespace sw

struct sw_dev_batch {
    sw::DevBuf<uint32_t> words;
    sw::DevBuf<uint64_t> rec_word_off;
    sw::DevBuf<uint32_t> rec_asm;   // [R] assembly of each record
    sw::DevBuf<uint32_t> rec_len;      // [R] bases per record
    sw::DevBuf<uint32_t> rec_inv_off;  // [R+1] slice of inv_* owned by each record
    sw::DevBuf<uint32_t> inv_start, inv_len;  // unhashable runs (record-relative)
    sw_batch meta;                  // host tables (no packed words) for the planner + ids
    cudaStream_t stream = nullptr;
    float h2d_ms = 0;
};

struct sw_graph {
    sw::DevGraph dev;
    bool on_device = false;
    // host copies live in pinned buffers drawn from a process-wide cache (api.cu: HostPool)
    sw::HostBuf h_kmers, h_nodes, h_edges;
    uint64_t n_kmers = 0, n_nodes = 0, n_edges = 0;  // sizes (valid for device and host copies)
    bool on_host = false;
    ~sw_graph();
    std::vector<uint32_t> record_offsets;
    std::vector<std::string> ids;
    cudaStream_t stream = nullptr;
};
