// radix.cu -- stable LSD radix sort of (u64 key, u32 value) pairs, 8-bit digits, one
// read + one write of the data per digit ("onesweep": per-tile digit counts are chained
// through a decoupled look-back so the scatter pass needs no separate per-tile scan).
//
// Plays the role of the reference's 16-bit-digit CPU LSD sort
// (cpp/src/seqwin/build_internals.cpp:76-144) and, because the sort is what groups equal
// minimizer hashes / equal edges, of its two ankerl hash maps (cpp/src/seqwin/build.cpp:66-89).
#include <cuda_pipeline.h>
#include <cuda_runtime.h>

#include <algorithm>

#include "device.h"
#include "nbr.cuh"

namespace sw {

namespace {

constexpr int kRadixBits = 8;
constexpr int kRadix = 1 << kRadixBits;
constexpr int kSortThreads = 256;
constexpr int kSortItems = 16;
constexpr int kSortTile = kSortThreads * kSortItems;
constexpr int kMaxPasses = 8;
#ifndef SW_SORT_MINB
#define SW_SORT_MINB 3
#endif

// digit positions of the passes of one sort, lowest digit first
struct RadixPasses {
    int n;
    int shift[kMaxPasses];
    int bits[kMaxPasses];
    uint64_t bias;   // subtracted from every key before its digits are taken (a partition of a key RANGE)
};

constexpr unsigned long long kStAgg = 1ULL << 62;
constexpr unsigned long long kStInc = 2ULL << 62;
constexpr unsigned long long kStMask = (1ULL << 62) - 1;

__device__ __forceinline__ unsigned long long ld_relaxed(const unsigned long long* p)
{
    unsigned long long v;
    asm volatile("ld.relaxed.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_relaxed(unsigned long long* p, unsigned long long v)
{
    asm volatile("st.relaxed.gpu.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}

// Digit histograms of every pass in one read of the keys.
__global__ void __launch_bounds__(256) radix_hist_kernel(const uint64_t* __restrict__ keys, uint64_t n,
                                                         RadixPasses ps, unsigned long long* ghist)
{
    __shared__ uint32_t sh[kMaxPasses][kRadix];
    for (int i = threadIdx.x; i < kMaxPasses * kRadix; i += blockDim.x) (&sh[0][0])[i] = 0;
    __syncthreads();
    const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
        const uint64_t key = keys[i] - ps.bias;
        for (int p = 0; p < ps.n; ++p)
            atomicAdd(&sh[p][(key >> ps.shift[p]) & ((1u << ps.bits[p]) - 1u)], 1u);
    }
    __syncthreads();
    for (int i = threadIdx.x; i < ps.n * kRadix; i += blockDim.x) {
        const uint32_t c = (&sh[0][0])[i];
        if (c) atomicAdd(&ghist[i], (unsigned long long)c);
    }
}

// Per pass (one CTA each): exclusive scan of the 256 bin counts in place; out[pass] = largest bin.
__global__ void __launch_bounds__(kRadix) radix_offsets_kernel(unsigned long long* ghist, unsigned long long* max_bin)
{
    __shared__ unsigned long long s_warp[kRadix / 32];
    __shared__ unsigned long long s_max[kRadix / 32];
    unsigned long long* h = ghist + (size_t)blockIdx.x * kRadix;
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    const unsigned long long c = h[threadIdx.x];
    unsigned long long inc = c, mx = c;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        const unsigned long long t = __shfl_up_sync(0xffffffffu, inc, d);
        if (lane >= d) inc += t;
        const unsigned long long m = __shfl_xor_sync(0xffffffffu, mx, d);
        mx = m > mx ? m : mx;
    }
    if (lane == 31) s_warp[wid] = inc;
    if (lane == 0) s_max[wid] = mx;
    __syncthreads();
    unsigned long long base = 0, gmx = 0;
#pragma unroll
    for (int i = 0; i < kRadix / 32; ++i) {
        if (i < wid) base += s_warp[i];
        gmx = s_max[i] > gmx ? s_max[i] : gmx;
    }
    h[threadIdx.x] = base + inc - c;
    if (threadIdx.x == 0) max_bin[blockIdx.x] = gmx;
}

// Shared state of one onesweep CTA.
template <int NT, int ITEMS>
struct OnesweepSmem {
    uint32_t whist[NT / 32][kRadix];    // per-warp digit counts -> tile-sorted position of the warp's first item
    unsigned long long dbase[kRadix];    // global slot of tile-sorted position 0, per digit
    uint32_t scan[NT / 32];
    uint32_t tile;
};

// One tile of one pass.  FULL = the tile holds NT * ITEMS items (no bounds checks).
// Two more 64-bit value arrays that travel with the pairs: the OWNED neighbour hashes of every minimizer
// (agg.cuh, bucket_edges_kernel).  gen: the first pass of a partition derives them from the stream itself
// (in.a / in.b are not read); the later passes carry them.
struct NbrArrays {
    const uint64_t* in_a = nullptr;
    const uint64_t* in_b = nullptr;
    uint64_t* out_a = nullptr;
    uint64_t* out_b = nullptr;
    uint64_t key_bias = 0;           // subtracted from a key before its digit is taken
    uint64_t n_total = 0;            // gen: items of the whole stream (neighbours across tile borders)
    unsigned int* zero_key = nullptr;   // gen: set when a key is 0 (the "not owned" marker would be ambiguous)
    // Routed scatter (multi-GPU): the four output arrays of digit d start at route[a * 256 + d] (a = 0..3: keys, vals,
    // prev, next) -- memory of the GPU that owns the digit's hash range, mapped over NVLink -- and `goff` holds the
    // position of this shard's first item of every digit inside those arrays.  null: the ordinary outputs.
    uint64_t* const* route = nullptr;
};

template <int NT, int ITEMS, bool FULL, typename V, int NBR>   // NBR: 0 = pairs only, 1 = carry the neighbour arrays, 2 = generate them
__device__ __forceinline__ void onesweep_tile(OnesweepSmem<NT, ITEMS>& sm, uint64_t* s_buf, V* s_vin,
                                              const uint64_t* __restrict__ kin, uint64_t* __restrict__ kout,
                                              const V* __restrict__ vin, V* __restrict__ vout,
                                              uint32_t n_valid, uint32_t tile, int shift, uint32_t dmask,
                                              const unsigned long long* __restrict__ goff, unsigned long long* status,
                                              uint32_t n_tiles, const NbrArrays& nb, uint64_t tile_base)
{
    constexpr int NW = NT / 32;
    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    const uint32_t wofs = (uint32_t)wid * (32 * ITEMS) + lane;  // this thread's first item inside the tile
    const uint64_t* kin_t = kin + wofs;                         // (kin / vin already point at the tile)
    const V* vin_t = vin + wofs;

    uint64_t key[ITEMS];
    uint32_t rank[ITEMS];
#pragma unroll
    for (int i = 0; i < ITEMS; ++i) key[i] = (FULL || wofs + i * 32 < n_valid) ? kin_t[i * 32] : ~0ULL;
    // the values are not needed before the keys have been written out: fetch them into shared
    // memory asynchronously (cp.async), each thread into its own slots
    V* my_vin = s_vin + wofs;
#pragma unroll
    for (int i = 0; i < ITEMS; ++i)
        if (FULL || wofs + i * 32 < n_valid) __pipeline_memcpy_async(my_vin + i * 32, vin_t + i * 32, sizeof(V));
    __pipeline_commit();

    // warp-local stable ranking: items of one warp are ordered (item, lane)
    const uint32_t lt_mask = (1u << lane) - 1u;
#pragma unroll
    for (int i = 0; i < ITEMS; ++i) {
        const bool valid = FULL || wofs + i * 32 < n_valid;
        const uint32_t d = (uint32_t)((key[i] - nb.key_bias) >> shift) & dmask;
        const uint32_t peers = __match_any_sync(0xffffffffu, valid ? d : 0x1FFu);
        const int leader = __ffs(peers) - 1;
        uint32_t old = 0;
        if (lane == leader && valid) {
            old = sm.whist[wid][d];
            sm.whist[wid][d] = old + __popc(peers);
        }
        old = __shfl_sync(0xffffffffu, old, leader);
        rank[i] = old + __popc(peers & lt_mask);
        __syncwarp();
    }
    __syncthreads();

    // per digit: exclusive scan over warps; exclusive scan over digits gives the digit's first
    // tile-sorted position; the tile's count is chained through the decoupled look-back
    uint32_t run = 0;
    if (tid < kRadix) {
#pragma unroll
        for (int wv = 0; wv < NW; ++wv) {
            const uint32_t t = sm.whist[wv][tid];
            sm.whist[wv][tid] = run;
            run += t;
        }
    }
    uint32_t inc = run;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        const uint32_t t = __shfl_up_sync(0xffffffffu, inc, d);
        if (lane >= d) inc += t;
    }
    if (lane == 31) sm.scan[wid] = inc;
    __syncthreads();
    uint32_t digit_first = inc - run;
#pragma unroll
    for (int i = 0; i < NW; ++i)
        if (i < wid) digit_first += sm.scan[i];
    // tile-major layout: a tile's 256 words share no sector with another tile's (no false sharing
    // between the CTA that publishes and the CTAs that poll)
    unsigned long long* my_status = status + (uint64_t)tile * kRadix + tid;
    if (tid < kRadix) {
        // publish the tile's digit count right away; the look-back itself runs after the keys have
        // been staged, when the predecessors have most likely published theirs
        st_relaxed(my_status, (tile == 0 ? kStInc : kStAgg) | run);
#pragma unroll
        for (int wv = 0; wv < NW; ++wv) sm.whist[wv][tid] += digit_first;  // now: tile-sorted position
    }
    __syncthreads();

    // keys: registers -> tile-sorted shared memory -> coalesced runs in global memory
    // (rank[] becomes the item's tile-sorted position)
#pragma unroll
    for (int i = 0; i < ITEMS; ++i) {
        if (FULL || wofs + i * 32 < n_valid) {
            const uint32_t d = (uint32_t)((key[i] - nb.key_bias) >> shift) & dmask;
            rank[i] += sm.whist[wid][d];
            s_buf[rank[i]] = key[i];
        }
    }
    if (tid < kRadix) {
        unsigned long long before = 0;
        if (tile != 0) {
            // walk the predecessors kLook at a time: the loads of one batch are independent
            constexpr int kLook = 8;
            long long p = (long long)tile - 1;
            bool done = false;
            while (!done) {
                unsigned long long st[kLook];
#pragma unroll
                for (int q = 0; q < kLook; ++q) st[q] = (p - q >= 0) ? ld_relaxed(my_status - (long long)(tile - (p - q)) * kRadix) : kStInc;
#pragma unroll
                for (int q = 0; q < kLook; ++q) {
                    if (done) break;
                    if ((st[q] >> 62) == 0) break;   // not published yet: re-read from here
                    before += st[q] & kStMask;
                    --p;
                    if ((st[q] >> 62) == 2) done = true;
                }
            }
            st_relaxed(my_status, kStInc | (before + run));
        }
        sm.dbase[tid] = goff[tid] + before - digit_first;
    }
    __syncthreads();
    uint32_t dig[ITEMS / 4];  // digits of the tile-sorted items this thread writes, 4 per register
#pragma unroll
    for (int i = 0; i < ITEMS; ++i) {
        const uint32_t p = (uint32_t)i * NT + tid;
        if ((i & 3) == 0) dig[i >> 2] = 0;
        if (FULL || p < n_valid) {
            const uint64_t k2 = s_buf[p];
            const uint32_t d = (uint32_t)((k2 - nb.key_bias) >> shift) & dmask;
            dig[i >> 2] |= d << (8 * (i & 3));
            (nb.route ? nb.route[d] : kout)[sm.dbase[d] + p] = k2;
        }
    }
    __pipeline_wait_prior(0);
    __syncthreads();
    // values take the same route through the (re-used) staging buffer
    V* s_val = reinterpret_cast<V*>(s_buf);
#pragma unroll
    for (int i = 0; i < ITEMS; ++i)
        if (FULL || wofs + i * 32 < n_valid) s_val[rank[i]] = my_vin[i * 32];
    __syncthreads();
#pragma unroll
    for (int i = 0; i < ITEMS; ++i) {
        const uint32_t p = (uint32_t)i * NT + tid;
        if (FULL || p < n_valid) {
            const uint32_t d = (dig[i >> 2] >> (8 * (i & 3))) & 255u;
            (nb.route ? reinterpret_cast<V*>(nb.route[256 + d]) : vout)[sm.dbase[d] + p] = s_val[p];
        }
    }
    if (NBR != 0) {
        // the two neighbour arrays take the same route, one after the other; their loads are issued together
        static_assert(NBR == 0 || sizeof(V) == 8, "neighbour arrays travel with 64-bit values");
#pragma unroll
        for (int a = 0; a < 2; ++a) {
            uint64_t tmp[ITEMS];
            if (NBR == 2) {
                const uint64_t* keys = kin - tile_base;    // whole-stream views (kin / vin point at the tile)
                const uint64_t* vals = reinterpret_cast<const uint64_t*>(vin) - tile_base;
#pragma unroll
                for (int i = 0; i < ITEMS; ++i) {
                    tmp[i] = 0;
                    if (FULL || wofs + i * 32 < n_valid) {
                        const uint64_t g = tile_base + wofs + i * 32;
                        const uint64_t h = keys[g];
                        const uint32_t rec = (uint32_t)(vals[g] >> 32);
                        tmp[i] = a == 0 ? agg::owned_prev(keys, vals, g, h, rec) : agg::owned_next(keys, vals, g, nb.n_total, h, rec);
                        if (a == 0 && h == 0) *nb.zero_key = 1u;
                    }
                }
            } else {
                const uint64_t* src = (a == 0 ? nb.in_a : nb.in_b) + tile_base + wofs;
#pragma unroll
                for (int i = 0; i < ITEMS; ++i) tmp[i] = (FULL || wofs + i * 32 < n_valid) ? src[i * 32] : 0;
            }
            __syncthreads();   // the previous array has left the staging buffer
#pragma unroll
            for (int i = 0; i < ITEMS; ++i)
                if (FULL || wofs + i * 32 < n_valid) s_buf[rank[i]] = tmp[i];
            __syncthreads();
            uint64_t* dst = a == 0 ? nb.out_a : nb.out_b;
#pragma unroll
            for (int i = 0; i < ITEMS; ++i) {
                const uint32_t p = (uint32_t)i * NT + tid;
                if (FULL || p < n_valid) {
                    const uint32_t d = (dig[i >> 2] >> (8 * (i & 3))) & 255u;
                    (nb.route ? nb.route[(2 + a) * 256 + d] : dst)[sm.dbase[d] + p] = s_buf[p];
                }
            }
        }
    }
}

template <int NT, int ITEMS, typename V, int NBR = 0>
__global__ void __launch_bounds__(NT, SW_SORT_MINB) radix_onesweep_kernel(
    const uint64_t* __restrict__ kin, uint64_t* __restrict__ kout, const V* __restrict__ vin,
    V* __restrict__ vout, uint64_t n, int shift, uint32_t dmask, const unsigned long long* __restrict__ goff,
    unsigned long long* status, unsigned int* ticket, const NbrArrays nb = NbrArrays())
{
    constexpr int NW = NT / 32;
    constexpr int TILE = NT * ITEMS;
    static_assert(ITEMS % 4 == 0 && NT >= kRadix, "one thread per digit");
    __shared__ OnesweepSmem<NT, ITEMS> sm;
    extern __shared__ __align__(16) unsigned char radix_smem[];
    uint64_t* s_buf = reinterpret_cast<uint64_t*>(radix_smem);              // [TILE] tile-sorted staging
    V* s_vin = reinterpret_cast<V*>(radix_smem + (size_t)TILE * 8);         // [TILE] prefetched values

    const int tid = threadIdx.x;
    if (tid == 0) sm.tile = atomicAdd(ticket, 1u);
    for (int i = tid; i < NW * kRadix; i += NT) (&sm.whist[0][0])[i] = 0;
    __syncthreads();
    const uint32_t tile = sm.tile;
    const uint64_t tile_base = (uint64_t)tile * TILE;
    const uint32_t n_valid = (uint32_t)(n - tile_base < (uint64_t)TILE ? n - tile_base : (uint64_t)TILE);
    if (n_valid == (uint32_t)TILE)
        onesweep_tile<NT, ITEMS, true, V, NBR>(sm, s_buf, s_vin, kin + tile_base, kout, vin + tile_base, vout, n_valid, tile,
                                               shift, dmask, goff, status, gridDim.x, nb, tile_base);
    else
        onesweep_tile<NT, ITEMS, false, V, NBR>(sm, s_buf, s_vin, kin + tile_base, kout, vin + tile_base, vout, n_valid, tile,
                                                shift, dmask, goff, status, gridDim.x, nb, tile_base);
}

// The passes of `ps` over (kin, vin): pass 0 reads the input (left untouched), the others ping-pong between
// (ka, va) and (kb, vb).  Returns 0 / 1: the result is in (ka, va) / (kb, vb); -1: no pass ran (the input IS the
// result).  Passes whose keys all share one digit are skipped.
template <typename V>
int run_passes(const uint64_t* kin, const V* vin, uint64_t n, const RadixPasses& ps, uint64_t* ka, V* va, uint64_t* kb,
               V* vb, cudaStream_t s, uint32_t* launches_out)
{
    uint32_t launches = 0;
    DevBuf<unsigned long long> ghist((size_t)kMaxPasses * kRadix, s, true);
    SW_CUDA(cudaMemsetAsync(ghist.p, 0, ghist.bytes(), s));
    const uint32_t hist_grid = (uint32_t)std::min<uint64_t>((n + 4095) / 4096, (uint64_t)sm_count() * 8);
    radix_hist_kernel<<<hist_grid, 256, 0, s>>>(kin, n, ps, ghist.p);
    SW_CUDA(cudaGetLastError());
    ++launches;
    // exclusive digit offsets per pass, computed in place on the device; a pass whose keys all share
    // one digit is a no-op and is skipped (needs only the per-pass maximum on the host)
    DevBuf<unsigned long long> max_bin(kMaxPasses, s, true);
    radix_offsets_kernel<<<ps.n, kRadix, 0, s>>>(ghist.p, max_bin.p);
    SW_CUDA(cudaGetLastError());
    ++launches;
    const unsigned long long* h_max = readback_u64(max_bin.p, ps.n, s);
    SW_CUDA(cudaStreamSynchronize(s));

    const uint64_t n_tiles = (n + kSortTile - 1) / kSortTile;
    DevBuf<unsigned long long> status(n_tiles * kRadix, s, true);
    DevBuf<unsigned int> ticket(1, s, true);
    constexpr size_t kSortSmem = (size_t)kSortTile * (8 + sizeof(V));
    SW_CUDA(cudaFuncSetAttribute(radix_onesweep_kernel<kSortThreads, kSortItems, V>,
                                 cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kSortSmem));
    int where = -1;
    const uint64_t* ksrc = kin;
    const V* vsrc = vin;
    for (int p = 0; p < ps.n; ++p) {
        if (h_max[p] == n) continue;
        uint64_t* kdst = where == 0 ? kb : ka;
        V* vdst = where == 0 ? vb : va;
        SW_CUDA(cudaMemsetAsync(status.p, 0, status.bytes(), s));
        SW_CUDA(cudaMemsetAsync(ticket.p, 0, sizeof(unsigned int), s));
        radix_onesweep_kernel<kSortThreads, kSortItems, V><<<(uint32_t)n_tiles, kSortThreads, kSortSmem, s>>>(
            ksrc, kdst, vsrc, vdst, n, ps.shift[p], (1u << ps.bits[p]) - 1u, ghist.p + (size_t)p * kRadix, status.p, ticket.p);
        SW_CUDA(cudaGetLastError());
        ++launches;
        where = where == 0 ? 1 : 0;
        ksrc = kdst;
        vsrc = vdst;
    }
    if (launches_out) *launches_out += launches;
    return where;
}

}  // namespace

uint32_t radix_sort_pairs(SortPairs& sp, int end_bit, cudaStream_t s, int begin_bit)
{
    const uint64_t n = sp.n;
    if (n < 2 || end_bit <= begin_bit) return 0;
    const int n_passes = std::min(kMaxPasses, (end_bit + kRadixBits - 1) / kRadixBits);
    RadixPasses ps{};
    for (int p = begin_bit / kRadixBits; p < n_passes; ++p) {   // stable LSD passes over digits [first_pass, n_passes)
        ps.shift[ps.n] = p * kRadixBits;
        ps.bits[ps.n] = kRadixBits;
        ++ps.n;
    }
    if (!sp.keys_alt.p || sp.keys_alt.n < n) sp.keys_alt.alloc(n, s, true);
    if (!sp.vals_alt.p || sp.vals_alt.n < n) sp.vals_alt.alloc(n, s, true);
    // in place from the caller's point of view: the first pass reads (keys, vals) and writes the alternates
    uint32_t launches = 0;
    const int where = run_passes<uint32_t>(sp.keys.p, sp.vals.p, n, ps, sp.keys_alt.p, sp.vals_alt.p, sp.keys.p, sp.vals.p, s,
                                           &launches);
    if (where == 0) {
        std::swap(sp.keys, sp.keys_alt);
        std::swap(sp.vals, sp.vals_alt);
    }
    return launches;
}

// Stable partition of (keys, vals) on the top `top_bits` bits of the key: ceil(top_bits / 8) onesweep passes,
// the lowest (and narrowest) digit first.  The input arrays are read only; out_k / out_v point at the result.
template <typename V>
uint32_t radix_partition_top(const uint64_t* keys, const V* vals, uint64_t n, int top_bits, uint64_t* ka, V* va, uint64_t* kb,
                             V* vb, cudaStream_t s, const uint64_t** out_k, const V** out_v)
{
    RadixPasses ps{};
    const int np = (top_bits + kRadixBits - 1) / kRadixBits;
    int lo = 64 - top_bits;
    for (int p = 0; p < np; ++p) {
        const int bits = p == 0 ? top_bits - kRadixBits * (np - 1) : kRadixBits;
        ps.shift[ps.n] = lo;
        ps.bits[ps.n] = bits;
        ++ps.n;
        lo += bits;
    }
    uint32_t launches = 0;
    const int where = n ? run_passes<V>(keys, vals, n, ps, ka, va, kb, vb, s, &launches) : -1;
    *out_k = where < 0 ? keys : (where == 0 ? ka : kb);
    *out_v = where < 0 ? vals : (where == 0 ? va : vb);
    return launches;
}
// The partition of the node stage: (key, value) plus the two owned-neighbour arrays, stable, on the key bits
// [lo_bit, lo_bit + n_bits), lowest digit first.  Either the input is the stream itself (in == nullptr): the first
// pass derives the neighbour arrays from it (and runs even for n_bits == 0); or it is a set that already holds
// them (a hash slice cut out of the stream, graph.cu): every pass carries them.  Every pass runs, also one whose
// keys share a digit.
uint32_t radix_partition_nbr(const uint64_t* keys, const uint64_t* vals, const NbrBuffers* in, uint64_t n, int lo_bit, int n_bits,
                             const NbrBuffers& A, const NbrBuffers& B, cudaStream_t s, const NbrBuffers** out,
                             unsigned int* d_zero_key, uint64_t key_bias)
{
    using V = unsigned long long;
    RadixPasses ps{};
    ps.bias = key_bias;
    const int np = n_bits > 0 ? (n_bits + kRadixBits - 1) / kRadixBits : (in ? 0 : 1);
    int lo = lo_bit;
    for (int p = 0; p < np; ++p) {
        const int bits = p == 0 ? n_bits - kRadixBits * (np - 1) : kRadixBits;
        ps.shift[ps.n] = lo;
        ps.bits[ps.n] = bits;
        ++ps.n;
        lo += bits;
    }
    if (ps.n == 0 || n == 0) {
        *out = in ? in : &A;
        return 0;
    }
    const uint64_t* hist_keys = in ? in->keys : keys;
    uint32_t launches = 0;
    DevBuf<unsigned long long> ghist((size_t)kMaxPasses * kRadix, s, true);
    SW_CUDA(cudaMemsetAsync(ghist.p, 0, ghist.bytes(), s));
    const uint32_t hist_grid = (uint32_t)std::min<uint64_t>((n + 4095) / 4096, (uint64_t)sm_count() * 8);
    radix_hist_kernel<<<hist_grid, 256, 0, s>>>(hist_keys, n, ps, ghist.p);
    DevBuf<unsigned long long> max_bin(kMaxPasses, s, true);
    radix_offsets_kernel<<<ps.n, kRadix, 0, s>>>(ghist.p, max_bin.p);
    SW_CUDA(cudaGetLastError());
    launches += 2;
    const uint64_t n_tiles = (n + kSortTile - 1) / kSortTile;
    DevBuf<unsigned long long> status(n_tiles * kRadix, s, true);
    DevBuf<unsigned int> ticket(1, s, true);
    constexpr size_t kSortSmem = (size_t)kSortTile * (8 + sizeof(V));
    SW_CUDA(cudaFuncSetAttribute(radix_onesweep_kernel<kSortThreads, kSortItems, V, 1>,
                                 cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kSortSmem));
    SW_CUDA(cudaFuncSetAttribute(radix_onesweep_kernel<kSortThreads, kSortItems, V, 2>,
                                 cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kSortSmem));
    const NbrBuffers* src = in;
    for (int p = 0; p < ps.n; ++p) {
        const NbrBuffers* dst = src == &A ? &B : &A;
        SW_CUDA(cudaMemsetAsync(status.p, 0, status.bytes(), s));
        SW_CUDA(cudaMemsetAsync(ticket.p, 0, sizeof(unsigned int), s));
        NbrArrays nb;
        nb.key_bias = key_bias;
        nb.out_a = dst->prev;
        nb.out_b = dst->next;
        if (src == nullptr) {
            nb.n_total = n;
            nb.zero_key = d_zero_key;
            radix_onesweep_kernel<kSortThreads, kSortItems, V, 2><<<(uint32_t)n_tiles, kSortThreads, kSortSmem, s>>>(
                keys, dst->keys, reinterpret_cast<const V*>(vals), reinterpret_cast<V*>(dst->vals), n, ps.shift[p],
                (1u << ps.bits[p]) - 1u, ghist.p + (size_t)p * kRadix, status.p, ticket.p, nb);
        } else {
            nb.in_a = src->prev;
            nb.in_b = src->next;
            radix_onesweep_kernel<kSortThreads, kSortItems, V, 1><<<(uint32_t)n_tiles, kSortThreads, kSortSmem, s>>>(
                src->keys, dst->keys, reinterpret_cast<const V*>(src->vals), reinterpret_cast<V*>(dst->vals), n, ps.shift[p],
                (1u << ps.bits[p]) - 1u, ghist.p + (size_t)p * kRadix, status.p, ticket.p, nb);
        }
        SW_CUDA(cudaGetLastError());
        ++launches;
        src = dst;
    }
    *out = src;
    return launches;
}

// The routing pass of the fused multi-GPU build: ONE generating pass over the stream on the top route_bits (<= 8)
// bits of h1 ("bins"; tables are sized for 256) whose
// scatter writes go straight to the owners of the hash ranges.  byte_counts_out (device, 256 words): items per top
// byte, filled by route_histogram; d_base (device, 256 words): where this shard's items of every byte start in the
// owner's arrays; d_route (device, 4 * 256 pointers): see NbrArrays.
void route_histogram(const uint64_t* keys, uint64_t n, int route_bits, unsigned long long* byte_counts_out, cudaStream_t s)
{
    RadixPasses ps{};
    ps.n = 1;
    ps.shift[0] = 64 - route_bits;
    ps.bits[0] = route_bits;
    SW_CUDA(cudaMemsetAsync(byte_counts_out, 0, kRadix * sizeof(unsigned long long), s));
    if (n == 0) return;
    const uint32_t hist_grid = (uint32_t)std::min<uint64_t>((n + 4095) / 4096, (uint64_t)sm_count() * 8);
    radix_hist_kernel<<<hist_grid, 256, 0, s>>>(keys, n, ps, byte_counts_out);
    SW_CUDA(cudaGetLastError());
}

void route_scatter(const uint64_t* keys, const uint64_t* vals, uint64_t n, int route_bits, const unsigned long long* d_base,
                   uint64_t* const* d_route, unsigned int* d_zero_key, cudaStream_t s)
{
    using V = unsigned long long;
    if (n == 0) return;
    const uint64_t n_tiles = (n + kSortTile - 1) / kSortTile;
    DevBuf<unsigned long long> status(n_tiles * kRadix, s, true);
    DevBuf<unsigned int> ticket(1, s, true);
    constexpr size_t kSortSmem = (size_t)kSortTile * (8 + sizeof(V));
    SW_CUDA(cudaFuncSetAttribute(radix_onesweep_kernel<kSortThreads, kSortItems, V, 2>,
                                 cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kSortSmem));
    SW_CUDA(cudaMemsetAsync(status.p, 0, status.bytes(), s));
    SW_CUDA(cudaMemsetAsync(ticket.p, 0, sizeof(unsigned int), s));
    NbrArrays nb;
    nb.n_total = n;
    nb.zero_key = d_zero_key;
    nb.route = d_route;
    radix_onesweep_kernel<kSortThreads, kSortItems, V, 2><<<(uint32_t)n_tiles, kSortThreads, kSortSmem, s>>>(
        keys, nullptr, reinterpret_cast<const V*>(vals), nullptr, n, 64 - route_bits, (1u << route_bits) - 1u, d_base, status.p,
        ticket.p, nb);
    SW_CUDA(cudaGetLastError());
}

template uint32_t radix_partition_top<uint32_t>(const uint64_t*, const uint32_t*, uint64_t, int, uint64_t*, uint32_t*, uint64_t*,
                                                uint32_t*, cudaStream_t, const uint64_t**, const uint32_t**);
template uint32_t radix_partition_top<unsigned long long>(const uint64_t*, const unsigned long long*, uint64_t, int, uint64_t*,
                                                          unsigned long long*, uint64_t*, unsigned long long*, cudaStream_t,
                                                          const uint64_t**, const unsigned long long**);

}  // namespace sw
