// sketch.cu -- persistent sketch kernels: ntHash2 + (w,k) window minimizers + emission.
//
// Replaces the reference's per-record btllib::minimize_sequence loop
// (cpp/vendor/btllib/minimizer.cpp:53-90, called from cpp/src/seqwin/build.cpp:152).
//
// One CTA processes one tile (<= TK valid k-mers of one record, see sketch_tile.h) at a time and
// loops, taking tile tickets from a global counter.  A tile appends its minimizers (already in
// position order) at a slot range claimed with ONE atomicAdd on a global cursor, so CTAs never
// wait for each other; a segment-copy pass (reorder_kernel) then moves each tile's range to its
// place in (record, window) order, which is the order the stable graph sort relies on.
//   sketch_fast_kernel     w-1 >= C1: chunk == thread, fully unrolled, flags fused (sketch_tile.h)
//   sketch_generic_kernel  any w: runtime chunking / direct scan for tiny windows
// Output: out_key[i] = h1, out_val[i] = pos | record_idx << 32.
#include <cuda_runtime.h>

#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <map>
#include <memory>
#include <mutex>

#include "device.h"
#include "sample.cuh"
#include "nthash.h"
#include "scan.cuh"
#include "sketch_tile.h"

namespace sw {

namespace {

// Exclusive block scan of one u32 per thread; *total = block sum. Contains two barriers.
template <int NT>
__device__ __forceinline__ uint32_t block_excl_scan(uint32_t v, uint32_t* warp_sums, uint32_t* total)
{
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    uint32_t inc = v;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        const uint32_t t = __shfl_up_sync(0xffffffffu, inc, d);
        if (lane >= d) inc += t;
    }
    if (lane == 31) warp_sums[wid] = inc;
    __syncthreads();
    uint32_t base = 0, sum = 0;
#pragma unroll
    for (int i = 0; i < NT / 32; ++i) {
        const uint32_t t = warp_sums[i];
        if (i < wid) base += t;
        sum += t;
    }
    __syncthreads();
    *total = sum;
    return base + inc - v;
}

// thread 0: claim `total` output slots for this tile and record where they are
__device__ __forceinline__ unsigned long long claim_slots(const SketchParams& P, uint32_t tile_id, uint32_t total)
{
    const unsigned long long slot = total ? atomicAdd(P.cursor, (unsigned long long)total) : 0ULL;
    P.tile_count[tile_id] = total;
    P.tile_slot[tile_id] = slot;
    return slot;
}

// ticket -> tile: a contiguous range of the plan, or (dense kernels finishing what the sparse kernel
// handed over) an explicit list whose length is only known on the device
__device__ __forceinline__ bool next_tile(const SketchParams& P, uint32_t ticket, uint32_t* tile_id)
{
    if (P.tile_list) {
        if (ticket >= *P.tile_list_n) return false;
        *tile_id = P.tile_list[ticket];
        return true;
    }
    *tile_id = P.tile_lo + ticket;
    return *tile_id < P.n_tiles;
}

template <int NT, int C1>
__global__ void __launch_bounds__(NT) sketch_fast_kernel(const __grid_constant__ SketchParams P)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    __shared__ uint32_t s_tile;
    __shared__ uint32_t s_warp_sums[NT / 32];
    __shared__ unsigned long long s_gbase;
    __shared__ uint16_t s_active[NT];

    constexpr uint32_t TK = NT * C1;
    const TileSmem S = carve_tile_smem(smem_raw, TK, NT);
    const int tid = threadIdx.x;
    if (tid < 20) S.tab[tid] = P.table.e[tid];

    for (;;) {
        if (tid == 0) s_tile = atomicAdd(P.tile_counter, 1u);
        __syncthreads();  // also: table visible / previous tile's smem reads done
        uint32_t tile_id;
        if (!next_tile(P, s_tile, &tile_id)) break;
        const Tile T = P.tiles[tile_id];

        fastA_hash_prefix<NT, C1>(tid, P, T, S);
        __syncthreads();
        fastB1_boundary<NT, C1>(tid, P, T, S);
        __syncthreads();
        // compact the chunks whose windows do not all select the same k-mer
        const bool act = fast_chunk_active<NT, C1>(tid, P, T, S);
        uint32_t n_active;
        const uint32_t a_slot = block_excl_scan<NT>(act ? 1u : 0u, s_warp_sums, &n_active);
        if (act) s_active[a_slot] = (uint16_t)tid;
        __syncthreads();
        FastState st;
        uint32_t cnt = 0;
        if ((uint32_t)tid < n_active) cnt = fastB2_windows<NT, C1>(s_active[tid], P, T, S, st);
        uint32_t total;
        const uint32_t excl = block_excl_scan<NT>(cnt, s_warp_sums, &total);
        if (tid == 0) s_gbase = claim_slots(P, tile_id, total);
        __syncthreads();
        if (cnt) fastD_write<NT, C1>(P, T, S, st, s_gbase + excl);
    }
}

template <int NT, int C1>
__global__ void __launch_bounds__(NT) sketch_generic_kernel(const __grid_constant__ SketchParams P)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    __shared__ uint32_t s_tile;
    __shared__ uint32_t s_warp_sums[NT / 32];
    __shared__ unsigned long long s_gbase;

    constexpr uint32_t TK = NT * C1;
    const TileSmem S = carve_tile_smem(smem_raw, TK, TK / 9 + 2);
    const int tid = threadIdx.x;
    if (tid < 20) S.tab[tid] = P.table.e[tid];

    for (;;) {
        if (tid == 0) s_tile = atomicAdd(P.tile_counter, 1u);
        __syncthreads();
        uint32_t tile_id;
        if (!next_tile(P, s_tile, &tile_id)) break;
        const Tile T = P.tiles[tile_id];

        phase1_hash<NT, C1>(tid, P, T, S);
        __syncthreads();
        if (P.c2) {
            phase2a_prefix<NT>(tid, P, T, S);
            __syncthreads();
            phase2b_windows<NT>(tid, P, T, S);
        } else {
            phase2_direct<NT>(tid, P, T, S);
        }
        __syncthreads();

        const uint32_t n_eval = T.n_kmers - P.w + 1;
        const uint32_t c3 = (n_eval + NT - 1) / NT;
        const uint64_t mask = phase3a_flags(tid, c3, T, P.w, S);
        uint32_t total;
        const uint32_t excl = block_excl_scan<NT>((uint32_t)__popcll(mask), s_warp_sums, &total);
        phase3b_stage(tid, c3, mask, excl, S);
        if (tid == 0) s_gbase = claim_slots(P, tile_id, total);
        __syncthreads();
        const unsigned long long gbase = s_gbase;
        for (uint32_t i = tid; i < total; i += NT) phase3c_write(i, gbase, P, T, S);
    }
}

// Sparse path (sketch_tile.h): candidates only, no per-k-mer shared memory.
// GAPS = false: the variant without the in-place handling of candidate-free stretches -- on ordinary sequence it is
// the faster one (the extra code costs about 7 % even when it never runs); tiles with such a stretch go to the list.
template <int NT, int C1, int CAP, bool GAPS>
__global__ void __launch_bounds__(NT) sketch_sparse_kernel(const __grid_constant__ SketchParams P)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    __shared__ uint32_t s_tile;
    __shared__ uint32_t s_warp_sums[NT / 32];
    __shared__ unsigned long long s_gbase;
    __shared__ GapList s_gaps;

    const SparseSmem S = carve_sparse_smem(smem_raw, NT, CAP);
    const int tid = threadIdx.x;
    if (tid < 20) S.tab[tid] = P.table.e[tid];
    constexpr uint32_t MC = sparse_mc(NT), MA = sparse_ma(NT);
    static_assert(NT * CAP < 65536, "candidate counts are scanned as 16-bit halves");

    for (;;) {
        if (tid == 0) s_tile = atomicAdd(P.tile_counter, 1u);
        __syncthreads();  // also: table visible / previous tile's smem reads done
        uint32_t tile_id;
        if (!next_tile(P, s_tile, &tile_id)) break;
        const Tile T = P.tiles[tile_id];
        // every thread is past the previous tile's write phase (which reads the stretch list); the selection
        // pass that fills it again comes several barriers later
        if (GAPS && tid == 0) s_gaps.n = 0;

        uint64_t mask = 0;
        bool ok = T.n_pieces == 1;
        if (ok) ok = sparseA_hash<NT, C1, CAP>(tid, P, T, S, &mask);
        bool hand_over = __syncthreads_or(!ok) != 0;
        uint32_t m = 0, ma = 0, off = 0, aoff = 0;
        if (!hand_over) {
            // one scan for both arrays: all candidates in the low half, the small ones in the high half
            const uint32_t cnt_all = popcount64(mask);
            uint32_t tot;
            const uint32_t ex = block_excl_scan<NT>(cnt_all | (sparse_count_small<NT>(tid, cnt_all, P, S) << 16),
                                                    s_warp_sums, &tot);
            m = tot & 0xFFFFu, ma = tot >> 16, off = ex & 0xFFFFu, aoff = ex >> 16;
            hand_over = m == 0 || m > MC || ma > MA;
        }
        uint32_t flags = 0;
        const uint32_t per = (m + NT - 1) / NT;
        if (!hand_over) {
            sparseC_compact<NT, C1>(tid, mask, off, aoff, m, ma, P, T, S);
            __syncthreads();
            bool bad;
            flags = sparseS_main<NT, GAPS>(tid, m, per, P, T, S, &s_gaps, &bad);
            sparseS_small<NT>(tid, ma, P, T, S);
            hand_over = __syncthreads_or(bad) != 0;
            if (GAPS && !hand_over && s_gaps.n != 0) {
                // stretches of more than w k-mers without a candidate (low-complexity sequence): their windows
                // are evaluated directly, in the shared memory of the private lists
                if (tid == 0) sparseG_sort(&s_gaps);
                __syncthreads();
                const uint32_t n_gaps = s_gaps.n;
                for (uint32_t gi = 0; gi < n_gaps; ++gi) {
                    sparseG_hash<NT>(tid, gi, s_gaps, P, T, S);
                    __syncthreads();
                    sparseG_windows<NT>(tid, gi, s_gaps, P, S);
                    __syncthreads();
                    if (tid == 0) sparseG_emit(gi, &s_gaps, P, T, S);
                    __syncthreads();
                }
                hand_over = (s_gaps.n & kGapOverflow) != 0;
            }
        }
        if (hand_over) {   // uniform: a dense kernel recomputes the tile
            if (tid == 0) P.fallback_tiles[atomicAdd(P.fallback_count, 1u)] = tile_id;
            continue;
        }
        flags = sparse_merge_flags(tid, m, per, flags, S);
        const uint32_t n_gaps = GAPS ? s_gaps.n : 0u;
        const uint32_t cnt = (uint32_t)__popc(flags) + (n_gaps ? sparse_gap_count(tid, m, per, s_gaps) : 0u);
        uint32_t total;
        const uint32_t excl = block_excl_scan<NT>(cnt, s_warp_sums, &total);
        if (tid == 0) s_gbase = claim_slots(P, tile_id, total);
        __syncthreads();
        if (cnt) {
            if (n_gaps) sparseD_write_gaps<NT>(tid, m, per, flags, s_gbase + excl, s_gaps, P, T, S);
            else sparseD_write<NT>(tid, per, flags, s_gbase + excl, P, T, S);
        }
    }
}

// One warp per tile: move the tile's slot range to its place in (record, window) order.
__global__ void __launch_bounds__(256) reorder_kernel(const uint64_t* __restrict__ ukey, const uint64_t* __restrict__ uval,
                                                      const unsigned long long* __restrict__ tile_off,
                                                      const unsigned long long* __restrict__ tile_slot, uint32_t n_tiles,
                                                      unsigned long long total, uint64_t* __restrict__ okey,
                                                      uint64_t* __restrict__ oval, int sbits,
                                                      unsigned long long* __restrict__ sample_set, unsigned long long* sample_out)
{
    const int lane = threadIdx.x & 31;
    const uint32_t warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const uint32_t n_warps = (gridDim.x * blockDim.x) >> 5;
    for (uint32_t t = warp; t < n_tiles; t += n_warps) {
        const unsigned long long dst = tile_off[t];
        const unsigned long long end = (t + 1 < n_tiles) ? tile_off[t + 1] : total;
        const unsigned long long src = tile_slot[t];
        const uint32_t n = (uint32_t)(end - dst);
        for (uint32_t i = lane; i < n; i += 32) {
            const uint64_t key = ukey[src + i];
            okey[dst + i] = key;
            agg::distinct_sample_item(key, sbits, sample_set, sample_out);   // k-mers per distinct hash: sizes the node buckets
            if (i + 1 < n) {
                // adjacent pairs per distinct pair (those inside a tile will do): the edges a bucket has to expect
                const uint64_t nxt = ukey[src + i + 1];
                const uint64_t lo = key < nxt ? key : nxt, hi = key < nxt ? nxt : key;
                agg::distinct_sample_item(agg::mix64(lo) + hi, sbits, sample_set + (1ull << agg::kSampleSetBits), sample_out + 2);
            }
            oval[dst + i] = uval[src + i];
        }
    }
}

// the k-independent 4-base seed tables, uploaded once per process / device
const TetraTable* device_tetra_table(cudaStream_t s)
{
    static std::mutex mu;
    static std::map<int, TetraTable*> per_device;
    int dev = 0;
    SW_CUDA(cudaGetDevice(&dev));
    std::lock_guard<std::mutex> lk(mu);
    auto it = per_device.find(dev);
    if (it != per_device.end()) return it->second;
    auto host = std::make_unique<TetraTable>();
    make_tetra_table(*host);
    TetraTable* d = nullptr;
    SW_CUDA(cudaMalloc((void**)&d, sizeof(TetraTable)));
    SW_CUDA(cudaMemcpyAsync(d, host.get(), sizeof(TetraTable), cudaMemcpyHostToDevice, s));
    SW_CUDA(cudaStreamSynchronize(s));
    per_device[dev] = d;
    return d;
}

struct KernelConfig {
    int nt, c1;
    void (*fast)(const SketchParams);
    void (*generic)(const SketchParams);
};

#define SW_CFG(NT, C1) {NT, C1, sketch_fast_kernel<NT, C1>, sketch_generic_kernel<NT, C1>}
const KernelConfig kConfigs[] = {
    SW_CFG(128, 27),  // default: 5 CTAs/SM (20 warps); measured best of the sweep in profiles/
    SW_CFG(128, 33),
    SW_CFG(256, 45),
    SW_CFG(256, 61),  // large windows (w up to ~15k)
    SW_CFG(128, 45),
    SW_CFG(128, 21),
    SW_CFG(128, 23),
    SW_CFG(64, 33),
    SW_CFG(256, 33),  // holds the sparse kernel's 8192-k-mer tiles
};
constexpr int kNumConfigs = sizeof(kConfigs) / sizeof(kConfigs[0]);
constexpr int kLargeWindowConfig = 3;
constexpr int kSparseDenseConfig = 8;
// the sparse kernel: 128 threads x 64 k-mers, 20 private candidate slots per thread
constexpr int kSparseNT = 128, kSparseC1 = 64, kSparseCap = 20;
constexpr uint32_t kSparseTK = kSparseNT * kSparseC1;

}  // namespace

int sketch_pick_config(uint32_t w, uint32_t* tk_out, bool* sparse_out)
{
    *sparse_out = false;
    // large windows: the sparse kernel (SEQWIN_SKETCH_DENSE=1 keeps every tile on the dense kernels)
    if (w >= kSparseMinW && (uint64_t)w * 4 <= kSparseTK && !getenv("SEQWIN_SKETCH_DENSE") &&
        !getenv("SEQWIN_SKETCH_CONFIG") && !getenv("SEQWIN_SKETCH_GENERIC")) {
        static_assert(kSparseTK <= 256 * 33, "the hand-over configuration must hold a sparse tile");
        *sparse_out = true;
        *tk_out = kSparseTK;
        return kSparseDenseConfig;
    }
    int cfg = 0;
    if (const char* e = getenv("SEQWIN_SKETCH_CONFIG")) cfg = std::max(0, std::min(kNumConfigs - 1, atoi(e)));
    // keep the halo (w k-mers re-hashed per tile) under ~25 % of the tile
    if ((uint64_t)w * 4 > (uint64_t)kConfigs[cfg].nt * kConfigs[cfg].c1) cfg = kLargeWindowConfig;
    const uint32_t tk = (uint32_t)kConfigs[cfg].nt * kConfigs[cfg].c1;
    if (w + 64 > tk)
        fail_runtime("windowsize " + std::to_string(w) + " exceeds the sketch kernel's limit of " +
                     std::to_string(tk - 64));
    *tk_out = tk;
    return cfg;
}

// ---- device-side tile planner -------------------------------------------------------------------
// Same rule as plan_tiles() in ingest.cpp, one thread per record (records have tens of tiles).

struct RecordRuns {
    const uint32_t* rec_len;
    const uint32_t* rec_inv_off;
    const uint32_t* inv_start;
    const uint32_t* inv_len;
};

// Walk the hashable runs of record r; f(kidx, pos, n) for every run with >= k bases.
template <typename F>
__device__ __forceinline__ uint64_t for_each_piece(const RecordRuns& rr, uint32_t r, uint32_t k, F&& f)
{
    const uint32_t L = rr.rec_len[r];
    uint64_t n_valid = 0;
    uint32_t s = 0;
    for (uint32_t iv = rr.rec_inv_off[r]; iv <= rr.rec_inv_off[r + 1]; ++iv) {
        const bool last = iv == rr.rec_inv_off[r + 1];
        const uint32_t stop = last ? L : rr.inv_start[iv];
        if (stop - s >= k) {
            f((uint32_t)n_valid, s, stop - s - k + 1);
            n_valid += stop - s - k + 1;
        }
        if (!last) s = rr.inv_start[iv] + rr.inv_len[iv];
    }
    return n_valid;
}

__global__ void plan_count_kernel(RecordRuns rr, uint32_t R, uint32_t k, uint32_t w, uint32_t tw,
                                  unsigned long long* tiles_per_rec, unsigned long long* pieces_per_rec,
                                  unsigned long long* totals)
{
    const uint32_t r = blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= R) return;
    uint32_t n_pieces = 0;
    const uint64_t n_valid = for_each_piece(rr, r, k, [&](uint32_t, uint32_t, uint32_t) { ++n_pieces; });
    unsigned long long n_tiles = 0;
    if (n_valid >= w) {
        const uint64_t n_win = n_valid - w + 1;
        n_tiles = (n_win + tw - 1) / tw;
        atomicAdd(totals + 1, (unsigned long long)n_win);
    } else {
        n_pieces = 0;
    }
    if (n_valid) atomicAdd(totals, (unsigned long long)n_valid);
    tiles_per_rec[r] = n_tiles;
    pieces_per_rec[r] = n_pieces;
}

__global__ void plan_fill_kernel(RecordRuns rr, uint32_t R, uint32_t k, uint32_t w, uint32_t tw,
                                 const unsigned long long* tile_off, const unsigned long long* piece_off,
                                 unsigned long long n_tiles_total, Tile* tiles, Piece* pieces)
{
    const uint32_t r = blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= R) return;
    const unsigned long long t0 = tile_off[r];
    const unsigned long long t1 = (r + 1 < R) ? tile_off[r + 1] : n_tiles_total;
    if (t1 == t0) return;
    const unsigned long long p0 = piece_off[r];
    uint32_t np = 0;
    const uint64_t n_valid = for_each_piece(rr, r, k, [&](uint32_t kidx, uint32_t pos, uint32_t n) {
        pieces[p0 + np] = Piece{kidx, pos, n};
        ++np;
    });
    const uint64_t n_win = n_valid - w + 1;
    uint32_t pc = 0;
    unsigned long long t = t0;
    for (uint64_t a0 = 0; a0 < n_win; a0 += tw, ++t) {
        const uint64_t a1 = a0 + tw < n_win ? a0 + tw : n_win;
        const uint64_t e0 = a0 ? a0 - 1 : 0;
        const uint64_t nk = a1 + w - 1 - e0;
        while ((uint64_t)pieces[p0 + pc].kidx + pieces[p0 + pc].n <= e0) ++pc;
        uint32_t pe = pc;
        while (pe + 1 < np && pieces[p0 + pe + 1].kidx < e0 + nk) ++pe;
        tiles[t] = Tile{r, (uint32_t)e0, (uint32_t)nk, (uint32_t)(p0 + pc), pe - pc + 1, a0 == 0 ? 1u : 0u};
    }
}

DevPlan make_plan(const sw_dev_batch& d, uint32_t k, uint32_t w, cudaStream_t s, bool want_rec_tile_off)
{
    DevPlan dp;
    dp.config = sketch_pick_config(w, &dp.tk, &dp.sparse);
    const uint32_t R = (uint32_t)d.meta.rec_len.size();
    if (R == 0) return dp;
    const uint32_t tw = dp.tk - w;
    RecordRuns rr{d.rec_len.p, d.rec_inv_off.p, d.inv_start.p, d.inv_len.p};
    // [0,R) tiles per record, [R] total tiles, [R+1, 2R+1) pieces per record, [2R+1] total pieces,
    // [2R+2] valid k-mers, [2R+3] windows
    DevBuf<unsigned long long> buf((size_t)2 * R + 4, s, true);
    SW_CUDA(cudaMemsetAsync(buf.p + 2 * (size_t)R + 2, 0, 2 * sizeof(unsigned long long), s));
    const uint32_t grid = (R + 127) / 128;
    plan_count_kernel<<<grid, 128, 0, s>>>(rr, R, k, w, tw, buf.p, buf.p + R + 1, buf.p + 2 * (size_t)R + 2);
    exclusive_scan_u64(buf.p, R, buf.p + R, s);
    exclusive_scan_u64(buf.p + R + 1, R, buf.p + 2 * (size_t)R + 1, s);
    SW_CUDA(cudaGetLastError());
    const unsigned long long* hb = readback_u64(buf.p, want_rec_tile_off ? (size_t)2 * R + 4 : 0, s);
    const unsigned long long* h0p = readback_u64(buf.p + R, 1, s);
    const unsigned long long* h1p = readback_u64(buf.p + 2 * (size_t)R + 1, 3, s);
    SW_CUDA(cudaStreamSynchronize(s));
    const unsigned long long h[4] = {h0p[0], h1p[0], h1p[1], h1p[2]};
    if (want_rec_tile_off) dp.rec_tile_off.assign(hb, hb + R + 1);
    if (h[0] > 0x7FFFFFFFull) fail_runtime("too many sketch tiles");
    if (h[1] > 0xFFFFFFF0ull) fail_runtime("too many valid runs");
    dp.n_tiles = (uint32_t)h[0];
    dp.n_kmers = h[2];
    dp.n_windows = h[3];
    dp.tiles.alloc(h[0], s, true);
    dp.pieces.alloc(h[1], s, true);
    if (dp.n_tiles)
        plan_fill_kernel<<<grid, 128, 0, s>>>(rr, R, k, w, tw, buf.p, buf.p + R + 1, h[0], dp.tiles.p, dp.pieces.p);
    SW_CUDA(cudaGetLastError());
    return dp;
}

void run_sketch(const uint32_t* d_words, const uint64_t* d_rec_word_off, const DevPlan& plan,
                uint32_t k, uint32_t w, uint32_t rec_base, cudaStream_t s, SketchStream& out,
                const std::vector<SketchChunk>* chunks)
{
    out.n = 0;
    out.launches = 0;
    out.items_per_key = 0;
    out.pairs_per_edge = 0;
    if (plan.n_tiles == 0) {
        out.keys.alloc(0, s, true);
        out.vals.alloc(0, s, true);
        return;
    }
    const KernelConfig& kc = kConfigs[plan.config];
    const bool fast = (w - 1 >= (uint32_t)kc.c1) && !getenv("SEQWIN_SKETCH_GENERIC");
    auto dense_kernel = fast ? kc.fast : kc.generic;
    const size_t dense_smem = tile_smem_bytes((uint32_t)kc.nt * kc.c1, fast ? (uint32_t)kc.nt : (uint32_t)kc.nt * kc.c1 / 9 + 2);
    SW_CUDA(cudaFuncSetAttribute(dense_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)dense_smem));
    int dense_ctas = 0;
    SW_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&dense_ctas, dense_kernel, kc.nt, dense_smem));
    if (dense_ctas < 1) fail_runtime("sketch kernel does not fit on this device");
    const uint32_t dense_grid = (uint32_t)std::min<uint64_t>(plan.n_tiles, (uint64_t)sm_count() * dense_ctas);
    // first pass: the sparse kernel when the plan was cut for it, else the dense kernel itself
    auto kernel = dense_kernel;
    size_t smem = dense_smem;
    int nt = kc.nt;
    uint32_t grid = dense_grid;
    if (plan.sparse) {
        kernel = sketch_sparse_kernel<kSparseNT, kSparseC1, kSparseCap, true>;   // (the first pass picks its variant below)
        smem = sparse_smem_bytes(kSparseNT, kSparseCap);
        nt = kSparseNT;
        SW_CUDA(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        int ctas = 0;
        SW_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&ctas, kernel, nt, smem));
        if (ctas < 1) fail_runtime("sketch kernel does not fit on this device");
        grid = (uint32_t)std::min<uint64_t>(plan.n_tiles, (uint64_t)sm_count() * ctas);
    }
    // tiles the sparse kernels cannot finish: list 1 (from the first pass) is retried by the variant that settles
    // candidate-free stretches in place, list 2 (what that one hands over) goes to the dense kernel
    DevBuf<uint32_t> fallback_tiles(plan.sparse ? plan.n_tiles : 0, s, true), fallback_tiles2(plan.sparse ? plan.n_tiles : 0, s, true);

    // [0, n_tiles): per-tile counts (scanned in place into ordered offsets); then the slots
    DevBuf<unsigned long long> tile_info((size_t)plan.n_tiles * 2, s, true);
    // [0] cursor, [1] scan total, [2] / [3] tiles on list 1 / 2 | the ticket counter of the launch reading it << 32,
    // [4] ticket counter of the probe launch, [5..] one ticket counter (as u32) per launch
    const size_t n_launch = chunks && !chunks->empty() ? chunks->size() : 1;
    constexpr size_t kFirstTicket = 5;
    DevBuf<unsigned long long> counters(kFirstTicket + n_launch, s, true);

    // expected density 2/(w+1); leave 50 % headroom and re-run with the exact size on overflow
    uint64_t capacity = (uint64_t)((double)plan.n_kmers * 3.0 / ((double)w + 1.0)) + 4096;
    capacity = std::min<uint64_t>(capacity, plan.n_windows);
    // The ordered stream is allocated first, with the capacity of the unordered one, so that the unordered
    // buffers (dead after the reorder) sit on top of the scratch arena and can be handed back.
    DevBuf<unsigned long long> sample_set(2ull << agg::kSampleSetBits, s, true), sample_out(4, s, true);   // hashes | pairs
    DevBuf<uint64_t> ukeys, uvals;
    const ArenaMark stream_mark = arena_mark();
    ArenaMark unordered_mark = stream_mark;
    unsigned long long total = 0;
    cudaEvent_t ev[4];
    for (auto& e : ev) cudaEventCreate(&e);
    struct EvGuard {
        cudaEvent_t* e;
        ~EvGuard() { for (int i = 0; i < 4; ++i) cudaEventDestroy(e[i]); }
    } ev_guard{ev};
    for (int attempt = 0;; ++attempt) {
        arena_release(stream_mark);
        out.keys.alloc(capacity, s, true);
        out.vals.alloc(capacity, s, true);
        unordered_mark = arena_mark();
        ukeys.alloc(capacity, s, true);
        uvals.alloc(capacity, s, true);
        SW_CUDA(cudaMemsetAsync(counters.p, 0, counters.bytes(), s));
        SketchParams P;
        P.words = d_words;
        P.rec_word_off = d_rec_word_off;
        P.tiles = plan.tiles.p;
        P.pieces = plan.pieces.p;
        P.n_tiles = plan.n_tiles;
        P.k = k;
        P.w = w;
        P.c2 = choose_c2(w, (uint32_t)kc.c1);
        P.rec_base = rec_base;
        P.h1_mult = h1_multiplier(k);
        P.out_key = ukeys.p;
        P.out_val = uvals.p;
        P.capacity = capacity;
        P.cursor = counters.p;
        P.tile_count = tile_info.p;
        P.tile_slot = tile_info.p + plan.n_tiles;
        P.table = make_roll_table(k);
        P.tetra = device_tetra_table(s);
        {   // SEQWIN_SPARSE_CPW / SEQWIN_SPARSE_SMALL: tuning overrides (any value gives the same output)
            const char* e1 = getenv("SEQWIN_SPARSE_CPW");
            const char* e2 = getenv("SEQWIN_SPARSE_SMALL");
            P.cand_hi = sparse_threshold(w, e1 ? atof(e1) : sparse_cand_per_window(w));
            P.cand_hi_a = sparse_threshold(w, e2 ? atof(e2) : sparse_small_per_window(w));
        }
        P.fallback_count = reinterpret_cast<unsigned int*>(counters.p + 2);
        P.fallback_tiles = fallback_tiles.p;
        P.tile_list = nullptr;
        P.tile_list_n = nullptr;
        cudaEventRecord(ev[0], s);
        // Which sparse variant runs the first pass: the one without in-place stretch handling is about 6 % faster on
        // ordinary sequence, but every tile with a candidate-free stretch then costs a second attempt.  A probe over
        // the first tiles decides (SEQWIN_SPARSE_GAPS = 0 / 1 overrides).
        uint32_t probe_hi = 0;
        if (plan.sparse) {
            const char* force = getenv("SEQWIN_SPARSE_GAPS");
            bool gaps = force ? atoi(force) != 0 : false;
            const uint32_t n_probe = std::min<uint32_t>(plan.n_tiles / 16, 8192);
            if (!force && n_probe >= 1024) {
                probe_hi = n_probe;
                if (chunks && !chunks->empty()) {
                    probe_hi = std::min(probe_hi, (*chunks)[0].tile_hi);
                    if ((*chunks)[0].ready && attempt == 0) SW_CUDA(cudaStreamWaitEvent(s, (*chunks)[0].ready, 0));
                }
                P.tile_lo = 0;
                P.n_tiles = probe_hi;
                P.tile_counter = reinterpret_cast<unsigned int*>(counters.p + 4);
                sketch_sparse_kernel<kSparseNT, kSparseC1, kSparseCap, false><<<std::min(probe_hi, grid), nt, smem, s>>>(P);
                SW_CUDA(cudaGetLastError());
                ++out.launches;
                const unsigned long long* fp = readback_u64(counters.p + 2, 1, s);
                SW_CUDA(cudaStreamSynchronize(s));
                gaps = (uint32_t)(*fp & 0xFFFFFFFFull) * 16u > probe_hi;   // more than 1 tile in 16 handed over
                if (getenv("SEQWIN_DEBUG_SKETCH"))
                    fprintf(stderr, "[sketch] probe: %u of %u tiles handed over -> %s\n", (unsigned)(*fp & 0xFFFFFFFFull), probe_hi,
                            gaps ? "stretches settled in the first pass" : "plain first pass");
            }
            kernel = gaps ? sketch_sparse_kernel<kSparseNT, kSparseC1, kSparseCap, true>
                          : sketch_sparse_kernel<kSparseNT, kSparseC1, kSparseCap, false>;
            SW_CUDA(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        }
        for (size_t c = 0; c < n_launch; ++c) {
            P.tile_lo = probe_hi;
            P.n_tiles = plan.n_tiles;
            if (chunks && !chunks->empty()) {
                // tiles of one uploaded slice of the packed stream: wait for its H2D copy only
                const SketchChunk& ch = (*chunks)[c];
                if (ch.tile_hi <= ch.tile_lo) continue;
                if (ch.ready && attempt == 0) SW_CUDA(cudaStreamWaitEvent(s, ch.ready, 0));
                P.tile_lo = std::max(ch.tile_lo, probe_hi);
                P.n_tiles = ch.tile_hi;
                if (P.n_tiles <= P.tile_lo) continue;
            }
            if (P.n_tiles <= P.tile_lo) continue;
            P.tile_counter = reinterpret_cast<unsigned int*>(counters.p + kFirstTicket + c);
            const uint32_t g = (uint32_t)std::min<uint64_t>(P.n_tiles - P.tile_lo, grid);
            kernel<<<g, nt, smem, s>>>(P);
            SW_CUDA(cudaGetLastError());
            ++out.launches;
        }
        if (plan.sparse) {
            // list 1 (few tiles; the counts stay on the device) -> the variant that settles stretches in place ...
            P.tile_lo = 0;
            P.n_tiles = plan.n_tiles;
            P.tile_list = fallback_tiles.p;
            P.tile_list_n = reinterpret_cast<unsigned int*>(counters.p + 2);
            P.tile_counter = reinterpret_cast<unsigned int*>(counters.p + 2) + 1;
            P.fallback_tiles = fallback_tiles2.p;
            P.fallback_count = reinterpret_cast<unsigned int*>(counters.p + 3);
            auto retry = sketch_sparse_kernel<kSparseNT, kSparseC1, kSparseCap, true>;
            SW_CUDA(cudaFuncSetAttribute(retry, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
            retry<<<grid, nt, smem, s>>>(P);
            // ... and what that one hands over (tiles across runs of unhashable bases, very long stretches) -> dense kernel
            P.tile_list = fallback_tiles2.p;
            P.tile_list_n = reinterpret_cast<unsigned int*>(counters.p + 3);
            P.tile_counter = reinterpret_cast<unsigned int*>(counters.p + 3) + 1;
            dense_kernel<<<dense_grid, kc.nt, dense_smem, s>>>(P);
            SW_CUDA(cudaGetLastError());
            out.launches += 2;
        }
        cudaEventRecord(ev[1], s);
        const unsigned long long* tp = readback_u64(counters.p, 4, s);
        SW_CUDA(cudaStreamSynchronize(s));
        total = *tp;
        if (plan.sparse && getenv("SEQWIN_DEBUG_SKETCH"))
            fprintf(stderr, "[sketch] %u tiles, %u to the second sparse pass, %u to the dense kernel, %llu minimizers\n", plan.n_tiles,
                    (unsigned)(tp[2] & 0xFFFFFFFFull), (unsigned)(tp[3] & 0xFFFFFFFFull), total);
        if (total <= capacity) break;
        if (attempt) fail_runtime("sketch output overflow after resize");
        capacity = total;  // low-complexity input: more minimizers than the density estimate
    }
    out.n = total;
    cudaEventElapsedTime(&out.kernel_ms, ev[0], ev[1]);
    if (total == 0) {
        arena_release(unordered_mark);
        return;
    }
    cudaEventRecord(ev[2], s);
    exclusive_scan_u64(tile_info.p, plan.n_tiles, counters.p + 1, s);
    const uint32_t rgrid = (uint32_t)std::min<uint64_t>(((uint64_t)plan.n_tiles + 7) / 8, (uint64_t)sm_count() * 8);
    SW_CUDA(cudaMemsetAsync(sample_set.p, 0xFF, sample_set.bytes(), s));
    SW_CUDA(cudaMemsetAsync(sample_out.p, 0, sample_out.bytes(), s));
    reorder_kernel<<<rgrid, 256, 0, s>>>(ukeys.p, uvals.p, tile_info.p, tile_info.p + plan.n_tiles, plan.n_tiles,
                                         total, out.keys.p, out.vals.p, agg::sample_bits(total), sample_set.p, sample_out.p);
    const unsigned long long* sample_h = readback_u64(sample_out.p, 4, s);
    cudaEventRecord(ev[3], s);
    SW_CUDA(cudaGetLastError());
    out.launches += 2;
    SW_CUDA(cudaEventSynchronize(ev[3]));
    cudaEventElapsedTime(&out.reorder_ms, ev[2], ev[3]);
    SW_CUDA(cudaStreamSynchronize(s));   // the readback kernel follows the event
    out.items_per_key = sample_h[1] ? (double)sample_h[0] / (double)sample_h[1] : 1.0;
    out.pairs_per_edge = sample_h[3] ? (double)sample_h[2] / (double)sample_h[3] : 1.0;
    arena_release(unordered_mark);   // the reorder has finished: the unordered buffers go back to the arena
}

}  // namespace sw
