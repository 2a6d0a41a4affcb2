/*
 * seqwin_b200.h -- C ABI of libseqwin_b200.so (B200 / sm_100a).
 *
 * Drop-in boundary for the native extension of treangenlab/Seqwin, `seqwin.graph._core`
 * (reference: cpp/src/bindings/python_bindings.cpp:43-169).  The three pybind11 entry
 * points of that module map onto this header as follows (citations into /root/reference):
 *
 *   _build_native        python_bindings.cpp:50-90   -> sw_build + sw_graph_size /
 *        (seqwin::build, cpp/src/seqwin/build.cpp:330-394)  sw_graph_export / sw_graph_record_id
 *   _get_penalty_native  python_bindings.cpp:92-135  -> sw_get_penalty
 *        (seqwin::get_penalty, cpp/src/seqwin/filter.cpp:15-137)
 *   _filter_kmers_native python_bindings.cpp:137-168 -> sw_filter_kmers
 *        (seqwin::filter_kmers, cpp/src/seqwin/filter.cpp:139-201)
 *
 * Plain pointers and sizes only; no C++ exception crosses the ABI.  Every function that
 * returns int returns SW_OK or an error class that the Python shim turns into the exception
 * type the reference raises (RuntimeError / ValueError); the message is sw_last_error().
 * All compute runs on the current CUDA device; there is no CPU fallback: without a usable
 * device the compute entry points return SW_ERR_RUNTIME ("no CUDA device").
 *
 * Output record layouts are the reference's (cpp/include/seqwin/graph.hpp:15-53):
 *   kmer  { u32 pos; u32 record_idx; }                                   8 bytes
 *   node  { u64 hash; u64 start; u64 stop; u32 n_tar; u32 n_neg; f64 penalty; }  40 bytes
 *   edge  { u64 first; u64 second; u64 weight; }                        24 bytes
 */
#ifndef SEQWIN_B200_H
#define SEQWIN_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define SW_OK 0
#define SW_ERR_RUNTIME 1 /* -> RuntimeError (I/O, CUDA, limits; build.cpp:136-147,337-339) */
#define SW_ERR_VALUE 2   /* -> ValueError   (std::invalid_argument in filter.cpp:33-60,106-125) */

typedef struct sw_kmer { uint32_t pos, record_idx; } sw_kmer;
typedef struct sw_node { uint64_t hash, start, stop; uint32_t n_tar, n_neg; double penalty; } sw_node;
typedef struct sw_edge { uint64_t first, second, weight; } sw_edge;

typedef struct sw_graph sw_graph;   /* a built graph (device and/or host resident) */
typedef struct sw_batch sw_batch;   /* parsed + 2-bit packed assemblies in pinned host memory */
typedef struct sw_dev_batch sw_dev_batch; /* the same, resident in HBM */

/* which= argument of sw_graph_size */
enum { SW_KMERS = 0, SW_NODES = 1, SW_EDGES = 2, SW_OFFSETS = 3, SW_RECORDS = 4 };

/* Thread-local message of the last failing call on this thread. */
const char* sw_last_error(void);

/* ---- the reference boundary ------------------------------------------------------------ */

/* Replaces seqwin::build (build.cpp:330-394): read the FASTA / .gz assemblies with
 * n_host_threads parser threads, pack to 2 bit, sketch and aggregate on the GPU.
 * low_memory is accepted for signature parity (build.cpp:380-391 produces identical output). */
int sw_build(const char* const* paths, size_t n_paths, uint32_t k, uint32_t w,
             uint32_t n_host_threads, int low_memory, sw_graph** out);

size_t sw_graph_size(const sw_graph* g, int which);
/* Copy the graph into caller-owned (e.g. numpy) buffers sized by sw_graph_size. */
int sw_graph_export(sw_graph* g, void* kmers, void* nodes, void* edges, uint32_t* record_offsets);
size_t sw_graph_n_records(const sw_graph* g, size_t assembly);
const char* sw_graph_record_id(const sw_graph* g, size_t assembly, size_t i);
void sw_graph_free(sw_graph* g);

/* Replaces seqwin::get_penalty (filter.cpp:15-137): fills n_tar / n_neg / penalty in place.
 * Unlike the reference it is told n_kmers and rejects node ranges beyond it (SW_ERR_VALUE). */
int sw_get_penalty(const sw_kmer* kmers, size_t n_kmers, sw_node* nodes, size_t n_nodes,
                   const uint32_t* record_offsets, size_t n_offsets, const uint8_t* is_targets,
                   size_t n_assemblies, uint32_t n_threads);

/* Replaces seqwin::filter_kmers (filter.cpp:139-201).  Two-call protocol: with kmers_out ==
 * NULL only the two counts are written; then call again with buffers of that size. */
int sw_filter_kmers(const sw_kmer* kmers, size_t n_kmers, const sw_node* nodes, size_t n_nodes,
                    const uint64_t* used_hashes, size_t n_used, sw_kmer* kmers_out,
                    sw_node* nodes_out, size_t* n_kmers_out, size_t* n_nodes_out);

/* ---- staged entry points (host ingest | H2D | device build), used by bench / multi-GPU -- */

/* Host ingest (replaces read_fasta, cpp/src/utils/fasta_reader.cpp:207-213, plus the 2-bit
 * packer): FASTA -> pinned 2-bit words + invalid-base runs + record table. */
int sw_batch_from_fasta(const char* const* paths, size_t n_paths, uint32_t n_host_threads,
                        sw_batch** out);
/* Same, from in-memory ASCII records (synthetic data): seqs[i] has lens[i] bytes and belongs to
 * assembly asm_of[i] (non-decreasing).  ids may be NULL. */
int sw_batch_from_memory(const uint8_t* const* seqs, const uint32_t* lens, const uint32_t* asm_of,
                         const char* const* ids, size_t n_records, size_t n_assemblies,
                         uint32_t n_host_threads, sw_batch** out);
/* One batch out of several (assemblies of parts[0] first, then parts[1], ...): lets a caller that
 * produces its input piecewise build a large batch without holding more than one piece unpacked.
 * The parts stay valid and are still the caller's to free. */
int sw_batch_concat(const sw_batch* const* parts, size_t n_parts, sw_batch** out);
size_t sw_batch_n_bases(const sw_batch* b);
size_t sw_batch_n_records(const sw_batch* b);
size_t sw_batch_packed_bytes(const sw_batch* b);
/* cumulative records per assembly, n = assemblies + 1 entries */
int sw_batch_record_offsets(const sw_batch* b, int64_t* out, size_t n);
void sw_batch_free(sw_batch* b);

int sw_dev_upload(const sw_batch* b, sw_dev_batch** out); /* cudaMemcpyAsync from pinned */
void sw_dev_batch_free(sw_dev_batch* d);

/* Per-stage device times of one build, CUDA events on the launching stream (ms). */
typedef struct sw_stage_times {
    float h2d_ms, sketch_ms, sort_nodes_ms, nodes_ms, edges_ms, d2h_ms, total_ms;
    float plan_ms, sketch_kernel_ms, reorder_ms; /* parts of sketch_ms: host tile plan + upload, kernel, reorder */
    float penalty_ms;                            /* scoring kernel, when the call includes it */
    uint64_t n_bases, n_kmers, n_nodes, n_edges, n_tiles, sketch_launches, total_launches;
} sw_stage_times;

/* Device-resident build: input already in HBM, graph left in HBM (exported lazily). */
int sw_dev_build(const sw_dev_batch* d, uint32_t k, uint32_t w, sw_graph** out, sw_stage_times* t);
/* End-to-end from pinned host memory: H2D + build + D2H into the graph's host arrays. */
int sw_build_from_batch(const sw_batch* b, uint32_t k, uint32_t w, sw_graph** out, sw_stage_times* t);

/* Build + score in one call: the classes of the batch's assemblies are known up front, so get_penalty
 * (filter.cpp:15-137) is fused into the node stage -- n_tar / n_neg are counted while the nodes are
 * written, the penalty follows in one pass -- and the nodes come back with the three fields filled in.
 * is_targets has one byte per assembly of the batch and needs at least one target and one non-target
 * (SW_ERR_VALUE otherwise, as in filter.cpp:33-60). */
int sw_build_from_batch_scored(const sw_batch* b, uint32_t k, uint32_t w, const uint8_t* is_targets,
                               size_t n_assemblies, sw_graph** out, sw_stage_times* t);
/* The same on a device-resident batch; the graph stays in HBM. */
int sw_dev_build_scored(const sw_dev_batch* d, uint32_t k, uint32_t w, const uint8_t* is_targets, size_t n_assemblies,
                        sw_graph** out, sw_stage_times* t);
/* get_penalty on a device-resident graph (record_offsets == NULL: the graph's own offsets; pass the
 * global offsets for a multi-GPU hash-range graph). kernel_ms may be NULL. */
int sw_graph_penalty(sw_graph* g, const uint32_t* record_offsets, size_t n_offsets, const uint8_t* is_targets,
                     size_t n_assemblies, float* kernel_ms);

/* Minimizer stream of a device batch in (record, position) order -- the sketch stage alone.
 * Two-call protocol like sw_filter_kmers (h1_out == NULL -> count only). Test / multi-GPU hook. */
int sw_dev_sketch(const sw_dev_batch* d, uint32_t k, uint32_t w, uint64_t* h1_out,
                  uint32_t* pos_out, uint32_t* record_out, size_t capacity, size_t* n_out);

/* ---- multi-GPU building blocks (one process per GPU; seqwin_b200/dist.py drives them) --------- */

/* Run every later call of this host thread on `stream` (a cudaStream_t, e.g. torch's current
 * stream) instead of the library's own stream; NULL restores the default. */
int sw_set_stream(void* stream);
/* One rank's shard of a multi-GPU build: sw_dev_build with the shard's global record base
 * (record_idx = rec_base + local record index).  With is_targets (one byte per assembly of the shard,
 * any mix of classes; NULL = no scoring) the shard's nodes come with n_tar / n_neg counted; an assembly
 * lives on one rank, so sw_dist_merge adds the shards' counts and sw_graph_finish_penalty turns the
 * sums into penalties -- get_penalty (filter.cpp:15-137) without a pass over the merged k-mers. */
int sw_dev_build_ex(const sw_dev_batch* d, uint32_t k, uint32_t w, uint32_t rec_base, const uint8_t* is_targets,
                    size_t n_assemblies, sw_graph** out, sw_stage_times* t);
/* sw_build_from_batch with a record base; to_host = 0 leaves the graph in HBM (copies overlapped). */
int sw_build_from_batch_ex(const sw_batch* b, uint32_t k, uint32_t w, uint32_t rec_base, int to_host,
                           const uint8_t* is_targets, size_t n_assemblies, sw_graph** out, sw_stage_times* t);
/* penalty = sqrt((1 - n_tar / n_targets)^2 + (n_neg / n_non_targets)^2) for every node of a
 * device-resident graph whose counts are already in place. */
int sw_graph_finish_penalty(sw_graph* g, uint64_t n_targets, uint64_t n_non_targets);
/* Bring a device-resident graph into (pooled, pinned) host memory; sw_graph_export then memcpy's. */
int sw_graph_fetch(sw_graph* g);
/* Device pointers of a device-resident graph (valid until sw_graph_free). */
int sw_graph_device_ptrs(sw_graph* g, void** kmers, void** nodes, void** edges);
/* Cut the sorted node / k-mer / edge arrays at the n_parts+1 hash-range boundaries (nodes and
 * k-mers at i * 2^64 / n_parts, edges by `first` at the matching quantiles); each output array has
 * n_parts + 1 entries and may be NULL when that cut is not wanted. */
int sw_graph_split(sw_graph* g, uint32_t n_parts, uint64_t* node_split, uint64_t* kmer_split,
                   uint64_t* edge_split);
/* Host callback fired by the next builds (sw_dev_build*, sw_build_from_batch_ex) once the k-mer and
 * node arrays of `g` are final on the device while the edge stage is still to run: a multi-GPU
 * caller cuts them (sw_graph_split with edge_split = NULL) and starts their exchange from here, so
 * that the transfer overlaps the edge stage -- the role the per-thread hand-off to
 * merge_thread_graphs plays in the reference (cpp/src/seqwin/build.cpp:352-378).  Inside the
 * callback only sw_graph_size / sw_graph_device_ptrs / sw_graph_split may be called on `g`.
 * fn = NULL clears the hook.  The hook is per host thread (like sw_set_stream): it fires for builds
 * started by the thread that set it. */
typedef void (*sw_nodes_ready_fn)(void* user, sw_graph* g);
int sw_set_nodes_ready(sw_nodes_ready_fn fn, void* user);
/* Merge the slices a hash-range owner received from n_src ranks (device pointers, concatenated in
 * rank order; kmer_base[i] = first k-mer index of rank i's slice in rank i's own array). Mirrors
 * merge_thread_graphs (cpp/src/seqwin/build_internals.cpp:295-392); n_tar / n_neg of equal nodes add up.
 * recv_edges == NULL merges nodes + k-mers only and leaves the edges to sw_dist_merge_edges, so that
 * the edge slices can still be in flight while the nodes are merged. */
int sw_dist_merge(const void* recv_nodes, const uint64_t* node_counts, const void* recv_kmers,
                  const uint64_t* kmer_counts, const uint64_t* kmer_base, const void* recv_edges,
                  const uint64_t* edge_counts, uint32_t n_src, sw_graph** out, uint32_t* launches);
int sw_dist_merge_edges(sw_graph* g, const void* recv_edges, const uint64_t* edge_counts, uint32_t n_src,
                        uint32_t* launches);

/* ---- multi-GPU, routed: minimizer records travel to the owner of their hash range BEFORE they are aggregated --------
 * ("all-to-all of minimizer records range-partitioned by the top bits of h1 ... the owner then reduces
 * its key range").  A record carries the hashes of the stream neighbours whose adjacent pair it owns -- the pair
 * belongs to the minimizer with the smaller hash --, so the owner of a hash range has everything it needs for the
 * nodes of the range AND the edges they own: no merge of per-shard graphs (cpp/src/seqwin/build_internals.cpp:295-392)
 * is left. */
typedef struct sw_routed sw_routed;
/* Sketch a device-resident shard (global record indices start at rec_base) and partition its records, stably, on the
 * top byte of h1. */
int sw_dev_sketch_route(const sw_dev_batch* d, uint32_t k, uint32_t w, uint32_t rec_base, sw_routed** out, sw_stage_times* t);
/* Device pointers of the four record arrays (n 64-bit words each: h1 | pos + record << 32 | owned previous | owned
 * next neighbour hash, 0 = none), byte_off[257] = first record of every top-byte value, stats = {k-mers per distinct
 * hash, adjacent pairs per distinct pair} as sampled on this shard.  Any output may be NULL. */
int sw_routed_info(const sw_routed* r, void** keys, void** vals, void** prev, void** next, uint64_t* n, uint64_t* byte_off,
                   double* stats);
void sw_routed_free(sw_routed* r);
/* Nodes, k-mers and edges (scored when is_targets != NULL) of the records whose h1 has its top byte in
 * [byte_lo, byte_hi): the four device arrays hold exactly those records of ALL shards, concatenated in shard order
 * (each shard's in its own stream order).  record_offsets / is_targets describe all assemblies of all shards
 * (global record indices).  The key array is used as scratch.  Concatenating the owners' graphs in range order gives
 * the reference graph. */
int sw_dev_aggregate(const void* keys, const void* vals, const void* prev, const void* next, uint64_t n, uint32_t byte_lo,
                     uint32_t byte_hi, const uint32_t* record_offsets, size_t n_offsets, const uint8_t* is_targets,
                     size_t n_assemblies, double pairs_per_edge, const uint64_t* byte_off, uint32_t range_bits, sw_graph** out,
                     sw_stage_times* t);
/* range_bits (1..8): the hash space is cut into 2^range_bits bins and byte_lo / byte_hi are bin numbers (8: top bytes).
 * byte_off != NULL (byte_hi - byte_lo + 1 entries from 0 to n): the records arrive stably partitioned on the bins,
 * byte_off[i] = first record of top byte byte_lo + i -- what the fused routing below delivers; the owner then skips
 * the partition pass on that byte.
 *
 * Fused routing: the exchange is part of the routing pass.  Every rank allocates its receive arrays with
 * sw_peer_alloc (plain cudaMalloc memory with a CUDA IPC handle) and opens the other ranks' with sw_peer_open
 * (peer access over NVLink); sw_dev_sketch_hist sketches the shard and counts its records per top byte; the ranks
 * all-gather those counts (the only collective) and derive where every shard's records of every top byte go in the
 * owner's arrays -- (top byte, source rank) order, i.e. global stream order inside a byte --; sw_routed_scatter runs
 * the ONE pass that generates the owned neighbours and scatters keys / k-mers / neighbours straight into those
 * arrays (route_ptrs: 4 x 256 device pointers, array-major: the owner's keys, vals, prev, next base for every top
 * byte; byte_base[256]: first position of this shard's records there).  After a barrier the owner calls
 * sw_dev_aggregate on its own arrays. */
int sw_peer_alloc(size_t bytes, void** dev_ptr, void* ipc_handle /* 64 bytes out */);
int sw_peer_open(const void* ipc_handle, void** dev_ptr);
int sw_peer_close(void* dev_ptr);
int sw_peer_free(void* dev_ptr);
int sw_dev_sketch_hist(const sw_dev_batch* d, uint32_t k, uint32_t w, uint32_t rec_base, uint32_t route_bits, sw_routed** out,
                       uint64_t* byte_counts /* 256, the first 2^route_bits used */, sw_stage_times* t);
int sw_routed_scatter(sw_routed* r, const void* const* route_ptrs, const uint64_t* byte_base, float* kernel_ms);

/* ---- first consumers of the graph, on the device-resident arrays ------------ */
/* Edge-weight filter + isolated-node removal, the array part of _filter_edges_and_nodes
 * (src/seqwin/kmers.py:132-162): keeps edges with weight > weight_th and the nodes that are an
 * endpoint of a kept edge; order and all fields unchanged (node ranges still refer to the k-mer array).
 * In place on a device-resident graph ... */
int sw_graph_filter_edges(sw_graph* g, uint64_t weight_th);
/* ... or from / to host arrays (nodes_out / edges_out sized like the inputs; the counts come back).
 * SW_ERR_VALUE if an edge endpoint is not a node hash. */
int sw_filter_edges_and_nodes(const sw_node* nodes, size_t n_nodes, const sw_edge* edges, size_t n_edges,
                              uint64_t weight_th, sw_node* nodes_out, sw_edge* edges_out, size_t* n_nodes_out,
                              size_t* n_edges_out);
/* filter_kmers (cpp/src/seqwin/filter.cpp:139-201) in place on a device-resident graph: keeps the
 * nodes whose hash is in used_hashes (any order, duplicates allowed), compacts their k-mers and
 * rewrites start / stop; edges untouched.  Saves the D2H / H2D round trip of the k-mer array between
 * build, scoring and filtering. */
int sw_graph_filter_kmers(sw_graph* g, const uint64_t* used_hashes, size_t n_used);

/* the reductions behind the automatic penalty threshold (src/seqwin/kmers.py:424-429,
 * the branch without Mash) on a scored device-resident graph: sums = {sum n_tar, sum n_tar^2,
 * sum n_tar * n_neg} over the nodes, exact integers.  With T targets and N non-targets:
 * e_absence_tar = 1 - sums[1] / (T * sums[0]), e_presence_neg = sums[2] / (N * sums[0]). */
int sw_graph_count_sums(sw_graph* g, uint64_t sums[3]);

/* Device memory of this process: out[0] = scratch arena of the calling thread (high-water: it only grows),
 * out[1] / out[2] = used / reserved high-water marks of the stream-ordered pool (batches, graphs),
 * out[3] = bytes in use on the device right now (all processes). */
int sw_mem_stats(uint64_t out[4]);

/* The reduced-footprint plan for the device-resident entry points (sw_dev_build*, sw_build_from_batch*), per host
 * thread: what sw_build's low_memory argument selects (cpp/src/seqwin/build.cpp:264-325).  The aggregation runs
 * over hash slices of the minimizer stream, one after the other (csrc/graph.cu): identical output, the scratch a
 * fraction of the single-pass plan's.  The same plan is taken without being asked for when the single pass would
 * not fit into the free device memory. */
int sw_set_low_memory(int on);
/* Hand the calling thread's scratch arena and the unused part of the stream-ordered pool back to the driver
 * (the arena otherwise stays at its high-water mark between builds). */
int sw_trim_memory(void);

/* Measured peak of the INT32 ALU pipe, the roofline that bounds the sketch kernels: lane-operations per
 * second of dependent-chain LOP3, SHF, IADD and of their 1:1:1 mix (CUDA-event timed, ~10 ms). */
int sw_measure_int_peak(double lane_ops_per_s[4]);

/* Device properties the host side sizes grids with. */
int sw_device_info(int* sm_count, int* cc_major, int* cc_minor, size_t* hbm_bytes);

#ifdef __cplusplus
}
#endif
#endif /* SEQWIN_B200_H */
