// sketch.cu -- persistent sketch kernel: ntHash2 + (w,k) window minimizers + ordered emission.
//
// Replaces the reference's per-record btllib::minimize_sequence loop
// (cpp/vendor/btllib/minimizer.cpp:53-90, called from cpp/src/seqwin/build.cpp:152).
//
// One CTA processes one tile (<= TK valid k-mers of one record, see sketch_tile.h) at a time and
// loops, taking tile tickets from a global counter.  Tiles are numbered in (record, window)
// order and tickets are handed out in that order, so a decoupled look-back over per-tile
// minimizer counts gives every tile its slot range in the globally ordered output stream
// without a second pass.  Output: out_key[i] = h1, out_val[i] = pos | record_idx << 32.
#include <cuda_runtime.h>

#include <algorithm>
#include <cstdlib>

#include "device.h"
#include "nthash.h"
#include "sketch_tile.h"

namespace sw {

namespace {

constexpr unsigned long long kStAgg = 1ULL << 62;
constexpr unsigned long long kStInc = 2ULL << 62;
constexpr unsigned long long kStMask = (1ULL << 62) - 1;

__device__ __forceinline__ unsigned long long ld_status(const unsigned long long* p)
{
    unsigned long long v;
    asm volatile("ld.relaxed.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_status(unsigned long long* p, unsigned long long v)
{
    asm volatile("st.relaxed.gpu.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}

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

template <int NT, int C1>
__global__ void __launch_bounds__(NT) sketch_kernel(const __grid_constant__ SketchParams P)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    __shared__ uint32_t s_tile;
    __shared__ uint32_t s_warp_sums[NT / 32];
    __shared__ unsigned long long s_gbase;

    constexpr uint32_t TK = NT * C1;
    const TileSmem S = carve_tile_smem(smem_raw, TK);
    const int tid = threadIdx.x;

    if (tid < 20) S.tab[tid] = P.table.e[tid];

    for (;;) {
        if (tid == 0) s_tile = atomicAdd(P.tile_counter, 1u);
        __syncthreads();  // also orders the table / previous tile's smem reuse
        const uint32_t tile_id = s_tile;
        if (tile_id >= P.n_tiles) break;
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

        if (tid == 0) {
            unsigned long long before = 0;
            if (tile_id == 0) {
                st_status(P.tile_status, kStInc | total);
            } else {
                st_status(P.tile_status + tile_id, kStAgg | total);
                uint32_t j = tile_id - 1;
                for (;;) {
                    const unsigned long long st = ld_status(P.tile_status + j);
                    if ((st >> 62) == 0) { __nanosleep(40); continue; }
                    before += st & kStMask;
                    if ((st >> 62) == 2) break;
                    --j;
                }
                st_status(P.tile_status + tile_id, kStInc | (before + total));
            }
            if (tile_id == P.n_tiles - 1) *P.total_out = before + total;
            s_gbase = before;
        }
        __syncthreads();
        const unsigned long long gbase = s_gbase;
        for (uint32_t i = tid; i < total; i += NT) phase3c_write(i, gbase, P, T, S);
        // the barrier at the top of the loop separates these reads from the next tile's writes
    }
}

struct KernelConfig {
    int nt, c1;
    void (*kernel)(const SketchParams);
};

const KernelConfig kConfigs[] = {
    {128, 45, sketch_kernel<128, 45>},  // default: 3 warps per scheduler at 2 CTAs/SM
    {256, 21, sketch_kernel<256, 21>},  // more warps, more warm-up
    {256, 45, sketch_kernel<256, 45>},
    {256, 61, sketch_kernel<256, 61>},  // large windows (w up to ~15k)
};
constexpr int kNumConfigs = sizeof(kConfigs) / sizeof(kConfigs[0]);

}  // namespace

int sketch_pick_config(uint32_t w, uint32_t* tk_out)
{
    int cfg = 0;
    if (const char* e = getenv("SEQWIN_SKETCH_CONFIG")) cfg = std::max(0, std::min(kNumConfigs - 1, atoi(e)));
    // keep the halo (w k-mers re-hashed per tile) under ~25 % of the tile
    if ((uint64_t)w * 4 > (uint64_t)kConfigs[cfg].nt * kConfigs[cfg].c1) cfg = 3;
    const uint32_t tk = (uint32_t)kConfigs[cfg].nt * kConfigs[cfg].c1;
    if (w + 64 > tk)
        fail_runtime("windowsize " + std::to_string(w) + " exceeds the sketch kernel's limit of " +
                     std::to_string(tk - 64));
    *tk_out = tk;
    return cfg;
}

DevPlan make_plan(const sw_batch& meta, uint32_t k, uint32_t w, cudaStream_t s)
{
    DevPlan dp;
    dp.config = sketch_pick_config(w, &dp.tk);
    Plan plan = plan_tiles(meta, k, w, dp.tk);
    dp.n_tiles = (uint32_t)plan.tiles.size();
    dp.n_windows = plan.n_windows;
    dp.n_kmers = plan.n_kmers;
    dp.tiles.alloc(plan.tiles.size(), s);
    dp.pieces.alloc(plan.pieces.size(), s);
    if (!plan.tiles.empty())
        SW_CUDA(cudaMemcpyAsync(dp.tiles.p, plan.tiles.data(), plan.tiles.size() * sizeof(Tile),
                                cudaMemcpyHostToDevice, s));
    if (!plan.pieces.empty())
        SW_CUDA(cudaMemcpyAsync(dp.pieces.p, plan.pieces.data(), plan.pieces.size() * sizeof(Piece),
                                cudaMemcpyHostToDevice, s));
    // the host vectors die at return: the copies above must have consumed them
    SW_CUDA(cudaStreamSynchronize(s));
    return dp;
}

void run_sketch(const uint32_t* d_words, const uint64_t* d_rec_word_off, const DevPlan& plan,
                uint32_t k, uint32_t w, uint32_t rec_base, cudaStream_t s, SketchStream& out)
{
    out.n = 0;
    out.launches = 0;
    if (plan.n_tiles == 0) {
        out.keys.alloc(0, s);
        out.vals.alloc(0, s);
        return;
    }
    const KernelConfig& kc = kConfigs[plan.config];
    const uint32_t tk = (uint32_t)kc.nt * kc.c1;
    const size_t smem = tile_smem_bytes(tk);
    SW_CUDA(cudaFuncSetAttribute(kc.kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    int ctas_per_sm = 0;
    SW_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&ctas_per_sm, kc.kernel, kc.nt, smem));
    if (ctas_per_sm < 1) fail_runtime("sketch kernel does not fit on this device");
    const uint32_t grid = (uint32_t)std::min<uint64_t>(plan.n_tiles, (uint64_t)sm_count() * ctas_per_sm);

    DevBuf<unsigned long long> status(plan.n_tiles, s);
    DevBuf<unsigned long long> counters(2, s);  // [0] ticket (as u32), [1] total

    // expected density 2/(w+1); leave 50 % headroom and re-run with the exact size on overflow
    uint64_t capacity = (uint64_t)((double)plan.n_kmers * 3.0 / ((double)w + 1.0)) + 4096;
    capacity = std::min<uint64_t>(capacity, plan.n_windows);
    for (int attempt = 0; attempt < 2; ++attempt) {
        out.keys.alloc(capacity, s);
        out.vals.alloc(capacity, s);
        SW_CUDA(cudaMemsetAsync(status.p, 0, status.bytes(), s));
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
        P.out_key = out.keys.p;
        P.out_val = out.vals.p;
        P.capacity = capacity;
        P.tile_status = status.p;
        P.tile_counter = reinterpret_cast<unsigned int*>(counters.p);
        P.total_out = counters.p + 1;
        P.table = make_roll_table(k);
        kc.kernel<<<grid, kc.nt, smem, s>>>(P);
        SW_CUDA(cudaGetLastError());
        ++out.launches;
        unsigned long long total = 0;
        SW_CUDA(cudaMemcpyAsync(&total, counters.p + 1, sizeof(total), cudaMemcpyDeviceToHost, s));
        SW_CUDA(cudaStreamSynchronize(s));
        out.n = total;
        if (total <= capacity) return;
        capacity = total;  // low-complexity input: more minimizers than the density estimate
    }
    fail_runtime("sketch output overflow after resize");
}

}  // namespace sw
