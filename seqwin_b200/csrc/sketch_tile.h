// sketch_tile.h -- the per-tile sketch algorithm, written as per-thread phase functions that
// are separated by block barriers.  The CUDA kernel (sketch.cu) calls them with
// tid = threadIdx.x; the test-only host emulator (tests/emul) calls them in a loop over tid, so
// the tiling / halo / tie-break logic can be checked against the oracle without a GPU.
//
// What a tile computes (reference semantics: cpp/vendor/btllib/minimizer.cpp:14-90):
//   h0[j]   canonical ntHash of the j-th valid k-mer held by the tile          (phase 1)
//   A[a]    index of the RIGHTMOST minimum of h0[a .. a+w-1], for every window  (phase 2)
//   emit    window a emits k-mer A[a] iff A[a] != A[a-1] (or a is the record's first window)
//           and h0[A[a]] != 2^64-1                                              (phase 3)
//
// Window minima use a block decomposition: the tile's k-mers are cut into logical chunks of c2
// (c2 <= w-1, odd so that chunk-strided shared-memory accesses are conflict free).  A window
// [a, e] = suffix of a's chunk + whole chunks + prefix of e's chunk, so
//   A[a] = rightmost-min( suffix_min(a), chunk_min(range), prefix_min(e) ).
// For w <= kDirectW each window is scanned directly.
#pragma once

#include "common.h"
#include "nthash.h"

namespace sw {

constexpr uint32_t kDirectW = 9;   // windows of <= 9 k-mers are scanned directly
constexpr uint32_t kMaxC2 = 64;

// Logical chunk length for the window-minimum decomposition (0 = direct scan).
inline uint32_t choose_c2(uint32_t w, uint32_t c1)
{
    if (w <= kDirectW) return 0;
    uint32_t c = w - 1 < c1 ? w - 1 : c1;
    if (c > kMaxC2) c = kMaxC2;
    if ((c & 1) == 0) --c;
    return c;
}

struct SketchParams {
    const uint32_t* words;         // packed stream
    const uint64_t* rec_word_off;  // [R] first word of each record
    const Tile* tiles;
    const Piece* pieces;
    uint32_t tile_lo;              // this launch handles tiles [tile_lo, n_tiles)
    uint32_t n_tiles;
    uint32_t k, w, c2;
    uint32_t rec_base;             // global index of batch record 0
    uint64_t h1_mult;
    uint64_t* out_key;             // [capacity] h1 of each minimizer, (record, pos) order
    uint64_t* out_val;             // [capacity] pos | record_idx << 32
    unsigned long long capacity;
    // Tiles emit into [cursor, cursor + count) in completion order; tile_count / tile_slot let a
    // segment-copy pass (sketch.cu: reorder_kernel) restore (record, window) order afterwards.
    unsigned long long* cursor;       // next free slot of out_key / out_val
    unsigned long long* tile_count;   // [n_tiles] minimizers emitted by each tile
    unsigned long long* tile_slot;    // [n_tiles] first slot of each tile
    unsigned int* tile_counter;       // ticket dispenser
    // sparse path (below): candidate threshold and the list of tiles it hands to the dense kernels
    uint32_t cand_hi;                 // candidates are k-mers with (h0 >> 32) < cand_hi
    uint32_t cand_hi_a;               // ... and the "small" ones those below cand_hi_a
    unsigned int* fallback_count;     // sparse kernel: number of tiles appended to fallback_tiles
    uint32_t* fallback_tiles;         // [n_tiles]
    const uint32_t* tile_list;        // dense kernels: if set, ticket i handles tile tile_list[i] ...
    const unsigned int* tile_list_n;  // ... for i < *tile_list_n
    const TetraTable* tetra;          // 4-base warm-up tables (device global)
    RollTable table;
};

// Shared-memory view of one tile (TK = NT * C1 k-mers).
struct TileSmem {
    RollEntry* tab;   // [20]
    uint64_t* h0;     // [TK]
    uint64_t* cm_h;   // [TK / 9 + 2]  chunk minima (value)
    uint16_t* cm_i;   // [TK / 9 + 2]  chunk minima (tile-local index)
    uint16_t* amin;   // [TK] A[a] (fast path: only for emitting windows)
    uint16_t* pidx;   // [TK] tile-local index of the prefix argmin of the element's chunk up to it;
                      //      the generic path reuses it as the emit staging list
    uint16_t* first_a;  // [TK / 9 + 2] fast path: A of each thread's first window
};

// nc = number of chunk-minimum slots: TK / 9 + 2 for the generic kernel (logical chunks of >= 9),
// NT for the specialised kernel (chunk == thread).
SW_HD size_t tile_smem_bytes(uint32_t tk, uint32_t nc)
{
    size_t b = sizeof(RollEntry) * 20 + sizeof(uint64_t) * tk + sizeof(uint64_t) * nc;
    b += sizeof(uint16_t) * (2 * (size_t)nc + 2 * (size_t)tk);
    return (b + 15) & ~(size_t)15;
}

SW_HD TileSmem carve_tile_smem(unsigned char* base, uint32_t tk, uint32_t nc)
{
    TileSmem s;
    s.tab = reinterpret_cast<RollEntry*>(base);
    s.h0 = reinterpret_cast<uint64_t*>(base + sizeof(RollEntry) * 20);
    s.cm_h = s.h0 + tk;
    s.cm_i = reinterpret_cast<uint16_t*>(s.cm_h + nc);
    s.amin = s.cm_i + nc;
    s.pidx = s.amin + tk;
    s.first_a = s.pidx + tk;
    return s;
}

// 16 bases starting at record position p (W = first word of the record).
SW_HD uint32_t fetch16(const uint32_t* W, uint64_t p)
{
    const uint64_t wi = p >> 4;
    const uint32_t sh = (uint32_t)(p & 15) * 2;
    const uint32_t lo = W[wi], hi = W[wi + 1];
#if defined(__CUDA_ARCH__)
    return __funnelshift_r(lo, hi, sh);
#else
    return sh ? (lo >> sh) | (hi << (32 - sh)) : lo;
#endif
}

SW_HD uint32_t base_at(const uint32_t* W, uint64_t p)
{
    return (W[p >> 4] >> ((uint32_t)(p & 15) * 2)) & 3u;
}

// ---- phase 1: hash --------------------------------------------------------------------------

// Fast path: the whole tile lies in one run of hashable bases.  Thread hashes C1 consecutive
// k-mers starting at record position p0: k warm-up steps, then C1-1 rolling steps.  With PREFIX
// it also tracks the running rightmost argmin of its chunk (tile-local index, base j0) into pidx
// and returns the chunk minimum.
SW_HD uint64_t tetra_load(const uint64_t* p)
{
#if defined(__CUDA_ARCH__)
    return __ldg(reinterpret_cast<const unsigned long long*>(p));
#else
    return *p;
#endif
}

// fwd / rev hash of the k-mer at record position p0, four bases per table step (TetraTable).
SW_HD void seed_kmer(const uint32_t* W, uint64_t p0, uint32_t k, const RollEntry* tab, const TetraTable* tt,
                     uint64_t* fwd_out, uint64_t* rev_out)
{
    const uint32_t G = k >> 2, rem = k & 3;
    uint64_t fwd = 0, rev = 0;
    for (uint32_t m = 0; 4 * m < G; ++m) {
        const uint32_t x = fetch16(W, p0 + 16 * m);
#pragma unroll
        for (uint32_t b = 0; b < 4; ++b)
            if (4 * m + b < G) fwd = srol4(fwd) ^ tetra_load(&tt->fwd4[(x >> (8 * b)) & 255u]);
    }
    if (rem) {
        const uint32_t x = fetch16(W, p0 + 4 * (uint64_t)G);
        for (uint32_t j = 0; j < rem; ++j) fwd = srol1(fwd) ^ tab[16 + ((x >> (2 * j)) & 3u)].f;
        for (uint32_t j = rem; j-- > 0;) rev = srol1(rev) ^ tab[16 + (3u - ((x >> (2 * j)) & 3u))].f;
    }
    for (uint32_t m = (G + 3) / 4; m-- > 0;) {
        const uint32_t x = fetch16(W, p0 + 16 * m);
#pragma unroll
        for (uint32_t b = 4; b-- > 0;)
            if (4 * m + b < G) rev = srol4(rev) ^ tetra_load(&tt->rev4[(x >> (8 * b)) & 255u]);
    }
    *fwd_out = fwd;
    *rev_out = rev;
}

// Fast path: the whole tile lies in one run of hashable bases.  Thread hashes C1 consecutive
// k-mers starting at record position p0: a table-driven seed of the first k-mer, then C1-1 rolling
// steps.  With PREFIX it also tracks the running rightmost argmin of its chunk (tile-local index,
// base j0) into pidx and returns the chunk minimum.
template <int C1, bool PREFIX>
SW_HD void hash_chunk_fast(const uint32_t* W, uint64_t p0, uint32_t k, const RollEntry* tab, const TetraTable* tt,
                           uint64_t* h0_out, uint16_t* pidx_out, uint32_t j0, uint64_t* min_h,
                           uint32_t* min_i)
{
    uint64_t fwd, rev;
    seed_kmer(W, p0, k, tab, tt, &fwd, &rev);
    uint64_t bh = fwd + rev;
    uint32_t bi = j0;
    h0_out[0] = bh;
    if (PREFIX) pidx_out[0] = (uint16_t)bi;
#pragma unroll
    for (int b = 0; b < (C1 - 1 + 15) / 16; ++b) {
        const uint32_t in = fetch16(W, p0 + k + 16 * b);
        const uint32_t out = fetch16(W, p0 + 16 * b);
        // nibble i of x / y = (out << 2 | in) of base 2i / 2i+1 of this block of 16
        const uint32_t x = (in & 0x33333333u) | ((out & 0x33333333u) << 2);
        const uint32_t y = ((in >> 2) & 0x33333333u) | (out & 0xCCCCCCCCu);
#pragma unroll
        for (int s = 0; s < 16; ++s) {
            const int n = 16 * b + s + 1;
            if (n < C1) {
                const uint32_t idx = (((s & 1) ? y : x) >> (4 * (s >> 1))) & 15u;
                roll_step(fwd, rev, tab[idx]);
                const uint64_t h = fwd + rev;
                h0_out[n] = h;
                if (PREFIX) {
                    take_if_le(h, j0 + n, bh, bi);
                    pidx_out[n] = (uint16_t)bi;
                }
            }
        }
    }
    if (PREFIX) { *min_h = bh; *min_i = bi; }
}

// Generic path: k-mers [j0, j1) of the tile may cross gaps between runs (pieces); every run
// entered is warmed up from scratch.
SW_HD void hash_range_generic(const uint32_t* W, const Piece* pieces, uint32_t e0, uint32_t j0,
                              uint32_t j1, uint32_t k, const RollEntry* tab, uint64_t* h0)
{
    uint32_t pi = 0;
    {
        const uint64_t g = (uint64_t)e0 + j0;
        while ((uint64_t)pieces[pi].kidx + pieces[pi].n <= g) ++pi;
    }
    uint32_t j = j0;
    while (j < j1) {
        const Piece pc = pieces[pi];
        const uint32_t off = (uint32_t)((uint64_t)e0 + j - pc.kidx);
        const uint32_t room = pc.n - off;
        const uint32_t cnt = (j1 - j) < room ? (j1 - j) : room;
        const uint64_t p = (uint64_t)pc.pos + off;
        uint64_t fwd = 0, rev = 0;
        for (uint32_t i = 0; i < k; ++i) roll_step(fwd, rev, tab[16 + base_at(W, p + i)]);
        h0[j] = fwd + rev;
        for (uint32_t n = 1; n < cnt; ++n) {
            const uint32_t idx = base_at(W, p + k - 1 + n) | (base_at(W, p + n - 1) << 2);
            roll_step(fwd, rev, tab[idx]);
            h0[j + n] = fwd + rev;
        }
        j += cnt;
        ++pi;
    }
}

template <int NT, int C1>
SW_HD void phase1_hash(int tid, const SketchParams& P, const Tile& T, const TileSmem& S)
{
    const uint32_t j0 = (uint32_t)tid * C1;
    if (j0 >= T.n_kmers) return;
    const uint32_t* W = P.words + P.rec_word_off[T.rec];
    const Piece* pcs = P.pieces + T.piece_lo;
    if (T.n_pieces == 1) {
        // reads (and stores into the padded tail of h0) may run past n_kmers; the packed
        // stream has kTailPadWords of slack and window evaluation never looks there
        const uint64_t p0 = (uint64_t)pcs[0].pos + ((uint64_t)T.e0 + j0 - pcs[0].kidx);
        hash_chunk_fast<C1, false>(W, p0, P.k, S.tab, P.tetra, S.h0 + j0, nullptr, j0, nullptr, nullptr);
    } else {
        const uint32_t j1 = j0 + C1 < T.n_kmers ? j0 + C1 : T.n_kmers;
        hash_range_generic(W, pcs, T.e0, j0, j1, P.k, S.tab, S.h0);
    }
}

// ---- phase 2 (generic, any c2 <= w-1): window minima -------------------------------------------

// 2a: per logical chunk, running rightmost argmin from the chunk start (prefix) + chunk minimum.
template <int NT>
SW_HD void phase2a_prefix(int tid, const SketchParams& P, const Tile& T, const TileSmem& S)
{
    const uint32_t c2 = P.c2;
    const uint32_t nc = (T.n_kmers + c2 - 1) / c2;
    for (uint32_t c = (uint32_t)tid; c < nc; c += NT) {
        const uint32_t base = c * c2;
        const uint32_t nend = T.n_kmers - base < c2 ? T.n_kmers - base : c2;
        uint64_t bh = S.h0[base];
        uint32_t bi = base;
        S.pidx[base] = (uint16_t)base;
        for (uint32_t n = 1; n < nend; ++n) {
            const uint64_t h = S.h0[base + n];
            if (h <= bh) { bh = h; bi = base + n; }
            S.pidx[base + n] = (uint16_t)bi;
        }
        S.cm_h[c] = bh;
        S.cm_i[c] = (uint16_t)bi;
    }
}

// 2b: evaluate every window that starts in chunk c.
template <int NT>
SW_HD void phase2b_windows(int tid, const SketchParams& P, const Tile& T, const TileSmem& S)
{
    const uint32_t c2 = P.c2, w = P.w;
    const uint32_t n_eval = T.n_kmers - w + 1;
    const uint32_t nce = (n_eval + c2 - 1) / c2;
    for (uint32_t c = (uint32_t)tid; c < nce; c += NT) {
        const uint32_t base = c * c2;
        // end chunk of the chunk's first window; higher windows end there or one chunk later
        const uint32_t te_lo = (base + w - 1) / c2;
        const uint32_t e_split = (te_lo + 1) * c2;  // window ends >= e_split lie in chunk te_lo + 1
        // whole chunks strictly between c and te_lo
        bool m_lo_valid = false;
        uint64_t m_lo_h = 0;
        uint32_t m_lo_i = 0;
        for (uint32_t cc = c + 1; cc < te_lo; ++cc) {
            const uint64_t h = S.cm_h[cc];
            if (!m_lo_valid || h <= m_lo_h) { m_lo_h = h; m_lo_i = S.cm_i[cc]; m_lo_valid = true; }
        }
        // ... and including te_lo (used by windows that end in te_lo + 1)
        uint64_t m_hi_h = m_lo_h;
        uint32_t m_hi_i = m_lo_i;
        {
            const uint64_t h = S.cm_h[te_lo];
            if (!m_lo_valid || h <= m_lo_h) { m_hi_h = h; m_hi_i = S.cm_i[te_lo]; }
        }
        // suffix pass, right to left; strict '<' keeps the right-hand element on ties.
        // The chunk is complete (c2 real k-mers) because its first window spans >= c2 + 1.
        uint64_t sh = 0;
        uint32_t si = 0;
        for (uint32_t n = c2; n-- > 0;) {
            const uint32_t a = base + n;
            const uint64_t h = S.h0[a];
            if (n == c2 - 1 || h < sh) { sh = h; si = a; }
            if (a >= n_eval) continue;
            const uint32_t e = a + w - 1;
            uint64_t rh = sh;
            uint32_t ri = si;
            if (e < e_split) {
                if (m_lo_valid && m_lo_h <= rh) { rh = m_lo_h; ri = m_lo_i; }
            } else {
                if (m_hi_h <= rh) { rh = m_hi_h; ri = m_hi_i; }
            }
            const uint32_t pe = S.pidx[e];
            const uint64_t ph = S.h0[pe];
            if (ph <= rh) { rh = ph; ri = pe; }
            S.amin[a] = (uint16_t)ri;
        }
    }
}

// Direct scan for small w.
template <int NT>
SW_HD void phase2_direct(int tid, const SketchParams& P, const Tile& T, const TileSmem& S)
{
    const uint32_t w = P.w;
    const uint32_t n_eval = T.n_kmers - w + 1;
    for (uint32_t a = (uint32_t)tid; a < n_eval; a += NT) {
        uint64_t bh = S.h0[a];
        uint32_t bi = a;
        for (uint32_t n = 1; n < w; ++n) {
            const uint64_t h = S.h0[a + n];
            if (h <= bh) { bh = h; bi = a + n; }
        }
        S.amin[a] = (uint16_t)bi;
    }
}

// ---- phase 3 (generic): emit --------------------------------------------------------------------

// 3a: thread owns windows [tid*c3, (tid+1)*c3); returns the flag mask of the emitting ones.
SW_HD uint64_t phase3a_flags(int tid, uint32_t c3, const Tile& T, uint32_t w, const TileSmem& S)
{
    const uint32_t n_eval = T.n_kmers - w + 1;
    const uint32_t a0 = (uint32_t)tid * c3;
    uint64_t mask = 0;
    if (a0 >= n_eval) return 0;
    const uint32_t a1 = a0 + c3 < n_eval ? a0 + c3 : n_eval;
    uint32_t prev = a0 ? S.amin[a0 - 1] : 0xFFFFFFFFu;
    for (uint32_t a = a0; a < a1; ++a) {
        const uint32_t cur = S.amin[a];
        bool f = (a == 0) ? (T.first != 0) : (cur != prev);
        if (f && S.h0[cur] == ~0ULL) f = false;   // minimizer.cpp:45
        if (f) mask |= 1ULL << (a - a0);
        prev = cur;
    }
    return mask;
}

// 3b: write the thread's emitting k-mer indices into the ordered staging list.
SW_HD void phase3b_stage(int tid, uint32_t c3, uint64_t mask, uint32_t offset, const TileSmem& S)
{
    const uint32_t a0 = (uint32_t)tid * c3;
    uint32_t o = offset;
    while (mask) {
#if defined(__CUDA_ARCH__)
        const int b = __ffsll((long long)mask) - 1;
#else
        const int b = __builtin_ctzll(mask);
#endif
        mask &= mask - 1;
        S.pidx[o++] = S.amin[a0 + b];
    }
}

// record position of tile-local k-mer idx
SW_HD uint32_t kmer_pos(const SketchParams& P, const Tile& T, uint32_t idx)
{
    const Piece* pcs = P.pieces + T.piece_lo;
    const uint64_t g = (uint64_t)T.e0 + idx;
    uint32_t pi = 0;
    while ((uint64_t)pcs[pi].kidx + pcs[pi].n <= g) ++pi;
    return pcs[pi].pos + (uint32_t)(g - pcs[pi].kidx);
}

// write tile-local k-mer idx as a minimizer into a global slot
SW_HD void write_minimizer(unsigned long long slot, uint32_t idx, const SketchParams& P, const Tile& T,
                           const TileSmem& S)
{
    if (slot >= P.capacity) return;  // counted, not stored: the host re-runs with more room
    P.out_key[slot] = h1_of(S.h0[idx], P.h1_mult);
    P.out_val[slot] = (uint64_t)kmer_pos(P, T, idx) | ((uint64_t)(P.rec_base + T.rec) << 32);
}

// 3c: write staged minimizer i of the tile to its global slot.
SW_HD void phase3c_write(uint32_t i, unsigned long long gbase, const SketchParams& P, const Tile& T,
                         const TileSmem& S)
{
    write_minimizer(gbase + i, S.pidx[i], P, T, S);
}

// ---- fast path: w - 1 >= C1, logical chunk == thread chunk, everything unrolled -----------------
//
//  A   hash own chunk + prefix argmin (fused when the tile is one run), chunk minimum
//  --barrier--
//  B1  A of the FIRST window of every chunk (whole chunk + whole chunks + one prefix lookup).
//      A[] is monotone, so a chunk whose first window and the next chunk's first window select the
//      same k-mer emits nothing: at density 2/(w+1) most chunks are skipped.
//  --barrier + compaction of the remaining ("active") chunks--
//  B2  one thread per active chunk: suffix pass right-to-left, A[a] for every window of the chunk;
//      window a+1 emits iff A[a+1] != A[a]; the 2^64-1 exclusion; count
//  --block scan + one atomicAdd on the global cursor--
//  D   write own minimizers
struct FastState {
    uint64_t mask;     // bit n: window c*C1 + n + 1 emits
    uint32_t chunk;    // chunk this thread evaluated in B2
    uint32_t a_first;  // A of the chunk's first window
    uint32_t f0;       // chunk 0 only: window 0 emits
};

template <int NT, int C1>
SW_HD void fastA_hash_prefix(int tid, const SketchParams& P, const Tile& T, const TileSmem& S)
{
    const uint32_t j0 = (uint32_t)tid * C1;
    if (j0 >= T.n_kmers) return;
    const uint32_t* W = P.words + P.rec_word_off[T.rec];
    const Piece* pcs = P.pieces + T.piece_lo;
    uint64_t bh;
    uint32_t bi;
    if (T.n_pieces == 1) {
        const uint64_t p0 = (uint64_t)pcs[0].pos + ((uint64_t)T.e0 + j0 - pcs[0].kidx);
        hash_chunk_fast<C1, true>(W, p0, P.k, S.tab, P.tetra, S.h0 + j0, S.pidx + j0, j0, &bh, &bi);
    } else {
        const uint32_t j1 = j0 + C1 < T.n_kmers ? j0 + C1 : T.n_kmers;
        hash_range_generic(W, pcs, T.e0, j0, j1, P.k, S.tab, S.h0);
        bh = S.h0[j0];
        bi = j0;
        S.pidx[j0] = (uint16_t)j0;
        for (uint32_t j = j0 + 1; j < j1; ++j) {
            const uint64_t h = S.h0[j];
            if (h <= bh) { bh = h; bi = j; }
            S.pidx[j] = (uint16_t)bi;
        }
    }
    S.cm_h[tid] = bh;
    S.cm_i[tid] = (uint16_t)bi;
}

// B1: A of window tid*C1 -> first_a[tid]
template <int NT, int C1>
SW_HD void fastB1_boundary(int tid, const SketchParams& P, const Tile& T, const TileSmem& S)
{
    const uint32_t w = P.w;
    const uint32_t n_eval = T.n_kmers - w + 1;
    const uint32_t j0 = (uint32_t)tid * C1;
    if (j0 >= n_eval) return;
    const uint32_t q = (w - 1) / C1;
    // the window covers chunk tid entirely (w - 1 >= C1), then chunks tid+1 .. tid+q-1, then a
    // prefix of chunk tid+q; everything to the right wins ties
    uint64_t rh = S.cm_h[tid];
    uint32_t ri = S.cm_i[tid];
    for (uint32_t cc = (uint32_t)tid + 1; cc < (uint32_t)tid + q; ++cc) {
        const uint64_t h = S.cm_h[cc];
        if (h <= rh) { rh = h; ri = S.cm_i[cc]; }
    }
    const uint32_t pe = S.pidx[j0 + w - 1];
    if (S.h0[pe] <= rh) ri = pe;
    S.first_a[tid] = (uint16_t)ri;
}

// does chunk c need a full evaluation?
template <int NT, int C1>
SW_HD bool fast_chunk_active(int c, const SketchParams& P, const Tile& T, const TileSmem& S)
{
    const uint32_t n_eval = T.n_kmers - P.w + 1;
    const uint32_t j0 = (uint32_t)c * C1;
    if (j0 >= n_eval) return false;
    if (j0 + C1 >= n_eval) return true;             // last chunk: no next boundary to compare with
    if (c == 0 && T.first != 0) return true;         // window 0 of the record always emits
    return S.first_a[c] != S.first_a[c + 1];
}

// B2: one thread per active chunk.  The selected position A(n) of window c*C1 + n is monotone in n,
// A(0) is known from B1 and so is A(C1) (the next chunk's first window), so the windows where the
// selection changes are found by bisection instead of evaluating all C1 windows:
//   1. one right-to-left pass stores the suffix argmin of the chunk from every offset (in amin[])
//   2. repeat: binary-search the smallest n > lo with A(n) != A(lo)  (about log2(C1) window
//      evaluations per emitted minimizer; each evaluation is three lookups, see fastB1_boundary)
// Emitted selections are appended to amin[c*C1 ...] (slots the search has already passed) and their
// window offsets are the set bits of st.mask.  Returns the number of minimizers emitted by windows
// c*C1+1 .. (c+1)*C1, plus window 0 for chunk 0 of a record's first tile.
template <int NT, int C1>
SW_HD uint32_t fastB2_windows(uint32_t c, const SketchParams& P, const Tile& T, const TileSmem& S, FastState& st)
{
    const uint32_t w = P.w;
    const uint32_t n_eval = T.n_kmers - w + 1;
    const uint32_t j0 = c * C1;
    const uint32_t q = (w - 1) / C1, r = (w - 1) % C1;  // window end = chunk c+q (+1), offset (n+r) % C1
    // whole chunks c+1 .. c+q-1 (windows ending in chunk c+q) and .. c+q (ending in c+q+1).
    // Chunks past the tile only matter to windows that are not evaluated.
    bool lo_valid = false;
    uint64_t m_lo_h = 0;
    uint32_t m_lo_i = 0;
    for (uint32_t cc = c + 1; cc < c + q && cc < NT; ++cc) {
        const uint64_t h = S.cm_h[cc];
        if (!lo_valid || h <= m_lo_h) { m_lo_h = h; m_lo_i = S.cm_i[cc]; lo_valid = true; }
    }
    uint64_t m_hi_h = m_lo_h;
    uint32_t m_hi_i = m_lo_i;
    if (c + q < NT) {
        const uint64_t h = S.cm_h[c + q];
        if (!lo_valid || h <= m_lo_h) { m_hi_h = h; m_hi_i = S.cm_i[c + q]; }
    }
    // 1. suffix argmin from every offset of the chunk; strict '<' keeps the right-hand element on ties.
    //    (the chunk is complete: its first window is evaluated and spans more than C1 k-mers)
    {
        uint64_t sh = S.h0[j0 + C1 - 1];
        uint32_t si = j0 + C1 - 1;
        S.amin[j0 + C1 - 1] = (uint16_t)si;
#pragma unroll 4
        for (int n = C1 - 2; n >= 0; --n) {
            const uint64_t h = S.h0[j0 + n];
            if (h < sh) { sh = h; si = j0 + n; }
            S.amin[j0 + n] = (uint16_t)si;
        }
    }
    // rightmost minimum of window j0 + n (0 <= n < C1, j0 + n < n_eval)
    auto eval = [&](uint32_t n, uint64_t* out_h) -> uint32_t {
        uint32_t ri = S.amin[j0 + n];
        uint64_t rh = S.h0[ri];
        if (n + r >= (uint32_t)C1) {
            if (m_hi_h <= rh) { rh = m_hi_h; ri = m_hi_i; }
        } else {
            if (lo_valid && m_lo_h <= rh) { rh = m_lo_h; ri = m_lo_i; }
        }
        const uint32_t pe = S.pidx[j0 + n + w - 1];
        const uint64_t ph = S.h0[pe];
        if (ph <= rh) { rh = ph; ri = pe; }
        *out_h = rh;
        return ri;
    };
    // right end of the search: the next chunk's first window if it is evaluated, else the last
    // evaluated window of this chunk
    uint32_t a_lo = S.first_a[c];
    uint32_t hi_end, a_end;
    uint64_t h_end = 0;
    bool end_is_next = false;
    if (j0 + C1 < n_eval) {
        hi_end = C1;
        a_end = S.first_a[c + 1];
        end_is_next = true;
    } else {
        hi_end = n_eval - 1 - j0;
        a_end = hi_end ? eval(hi_end, &h_end) : a_lo;
    }
    st.chunk = c;
    st.a_first = a_lo;
    st.f0 = (c == 0 && T.first != 0 && S.h0[a_lo] != ~0ULL) ? 1u : 0u;
    uint64_t mask = 0;
    uint32_t n_out = 0, lo = 0;
    // 2. next change point after lo, until the selection at the right end is reached
    while (a_lo != a_end) {
        uint32_t l = lo, rr = hi_end, a_r = a_end;
        uint64_t h_r = h_end;
        bool r_is_end = true;
        while (rr - l > 1) {
            const uint32_t m = (l + rr) >> 1;
            uint64_t hm;
            const uint32_t a_m = eval(m, &hm);
            if (a_m != a_lo) { rr = m; a_r = a_m; h_r = hm; r_is_end = false; } else { l = m; }
        }
        if (r_is_end && end_is_next) h_r = S.h0[a_r];
        // window j0 + rr selects a new k-mer; minimizer.cpp:45: never emit one whose h0 is 2^64-1
        if (h_r != ~0ULL) {
            mask |= 1ULL << (rr - 1);
            S.amin[j0 + n_out] = (uint16_t)a_r;   // slot n_out < rr: the search never looks there again
            ++n_out;
        }
        lo = rr;
        a_lo = a_r;
    }
    st.mask = mask;
    return n_out + st.f0;
}

template <int NT, int C1>
SW_HD void fastD_write(const SketchParams& P, const Tile& T, const TileSmem& S, const FastState& st,
                       unsigned long long slot)
{
    const uint32_t j0 = st.chunk * C1;
    if (st.f0) write_minimizer(slot++, st.a_first, P, T, S);
#if defined(__CUDA_ARCH__)
    const uint32_t n_out = (uint32_t)__popcll(st.mask);
#else
    const uint32_t n_out = (uint32_t)__builtin_popcountll(st.mask);
#endif
    for (uint32_t i = 0; i < n_out; ++i) write_minimizer(slot++, S.amin[j0 + i], P, T, S);
}

// ---- sparse path: large windows, one run of hashable bases per tile ---------------------------------
//
// A k-mer p of the tile is selected by some window (minimizer.cpp:69-87, rightmost minimum) iff a
// window of w k-mers fits between its nearest strictly smaller k-mer on the left (L) and its nearest
// smaller-or-equal k-mer on the right (R): R - L - 1 >= w.  It is then the selection of exactly the
// windows [max(L + 1, p - w + 1), min(p, R - w)], and selections are monotone in the window index,
// so "emit when the selection changes" (minimizer.cpp:41-47) = every such p once, in position order.
//
// Only small hashes can be selected: if every window holds at least one CANDIDATE, a k-mer with
// (h0 >> 32) < cand_hi, then every selection is a candidate and L / R need only be looked for among
// candidates.  cand_hi keeps about 11 candidates per window (sparse_cand_per_window), so each thread
//   A  hashes C1 consecutive k-mers and appends the few candidates to a private list  (no h0 array)
//   C  the lists are compacted into one position-ordered array per tile, plus a second array with
//      only the SMALL candidates ((h0 >> 32) < cand_hi_a, about 4 per window)
//   S  one thread per candidate walks outwards until it meets L and R (or has gone w positions).
//      A small candidate can only be stopped by small ones, so it walks the short array.  Any other
//      candidate first looks at the small ones next to it: they are all smaller, so unless w k-mers
//      fit between them it is done without walking.  The two kinds run in separate passes so that
//      the lanes of a warp do similar amounts of work.
//   D  selected candidates are written out
// Tiles the argument does not cover -- a window without candidate (low-complexity sequence), a
// list that overflows, a tile that crosses a gap -- are appended to fallback_tiles and recomputed
// by the dense kernels above.  Both paths are exact; the split only decides who does the work.
struct SparseSmem {
    RollEntry* tab;    // [20]
    uint64_t* list;    // [CAP][NT] slot-major private candidate lists (h0)
    uint64_t* key;     // [MC + 2]  all candidates: tile-local index << 32 | h0 >> 32; [0], [m + 1] sentinels
    uint64_t* akey;    // [MA + 2]  small candidates, same encoding and sentinels
    uint32_t* lo;      // [MC + 2]  low word of h0
    uint32_t* selbits; // [MC / 32 + 2] bit j: candidate j is selected (written by the small pass)
    uint16_t* aj;      // [MA + 2]  index of the small candidate in key[]
    uint16_t* na;      // [MC + 2]  number of small candidates among key[1 .. j]
};

// capacities of key[] / akey[] for NT threads: the means are 64 * NT * 11 / w and 64 * NT * 4 / w
// (3.5 and 1.3 per thread at w = 200, falling with w); shared memory per CTA decides the occupancy
SW_HD constexpr uint32_t sparse_mc(uint32_t nt) { return nt * 5; }
SW_HD constexpr uint32_t sparse_ma(uint32_t nt) { return nt * 7 / 4; }
constexpr uint32_t kSparseCheck = 8;       // a private list is checked for room every 8 steps
constexpr uint32_t kSparseMinW = 144;      // below this the dense kernels are used for every tile
// Candidates / small candidates per window.  At most one k-mer in 16 may be a candidate (the lists and
// arrays are sized for that), so narrow windows get fewer per window and hand over more tiles for
// lack of coverage: e^-9 per candidate at w = 144 (6 % of the tiles), e^-11 from w = 176 on (0.7 %).
inline double sparse_cand_per_window(uint32_t w) { return w >= 176 ? 11.0 : (double)w / 16.0; }
inline double sparse_small_per_window(uint32_t w) { return sparse_cand_per_window(w) * 4.0 / 11.0; }

// Windows without any candidate (low-complexity stretches) are settled in place: the k-mers between the
// two candidates around such a stretch are hashed again into a workspace that overlays the private lists
// (dead after the compaction) and their windows are evaluated directly.
constexpr uint32_t kGapMax = 6;         // stretches per tile handled in place
constexpr uint32_t kGapLenMax = 1024;   // k-mers per stretch
constexpr uint32_t kGapEmitMax = 192;   // minimizers they may add per tile
// (kept small: the sparse kernel's shared memory is sized to the byte for 7 CTAs per SM)
struct GapList {
    uint32_t n;                   // stretches registered by the selection pass; bit 31: more minimizers than kGapEmitMax
    int16_t a[kGapMax];           // tile-local index of the candidate before the stretch (-1: none)
    int16_t b[kGapMax];           // ... and after it (n_kmers: none)
    uint16_t owner[kGapMax];      // index in key[] of the candidate after it (m + 1: none)
    uint16_t e0[kGapMax + 1];     // first entry of the stretch in the emit arrays
};
constexpr uint32_t kGapOverflow = 0x80000000u;
SW_HD constexpr size_t sparse_gap_ws_words() { return kGapLenMax + kGapLenMax / 4 + kGapEmitMax + kGapEmitMax / 4; }
SW_HD constexpr size_t sparse_list_words(uint32_t nt, uint32_t cap)
{
    return (size_t)cap * nt > sparse_gap_ws_words() ? (size_t)cap * nt : sparse_gap_ws_words();
}

SW_HD size_t sparse_smem_bytes(uint32_t nt, uint32_t cap)
{
    const size_t mc = (size_t)sparse_mc(nt) + 2, ma = (size_t)sparse_ma(nt) + 2;
    size_t b = sizeof(RollEntry) * 20 + sizeof(uint64_t) * (sparse_list_words(nt, cap) + mc + ma);
    b += sizeof(uint32_t) * (mc + mc / 32 + 2) + sizeof(uint16_t) * (ma + mc);
    return (b + 15) & ~(size_t)15;
}

SW_HD SparseSmem carve_sparse_smem(unsigned char* base, uint32_t nt, uint32_t cap)
{
    const size_t mc = (size_t)sparse_mc(nt) + 2, ma = (size_t)sparse_ma(nt) + 2;
    SparseSmem s;
    s.tab = reinterpret_cast<RollEntry*>(base);
    s.list = reinterpret_cast<uint64_t*>(base + sizeof(RollEntry) * 20);
    s.key = s.list + sparse_list_words(nt, cap);
    s.akey = s.key + mc;
    s.lo = reinterpret_cast<uint32_t*>(s.akey + ma);
    s.selbits = s.lo + mc;
    s.aj = reinterpret_cast<uint16_t*>(s.selbits + mc / 32 + 2);
    s.na = s.aj + ma;
    return s;
}

// (h0 >> 32) threshold that leaves about `per_window` candidates in a window of w k-mers
inline uint32_t sparse_threshold(uint32_t w, double per_window)
{
    const double f = per_window / (double)w;
    return f >= 1.0 ? 0xFFFFFFFFu : (uint32_t)(f * 4294967296.0);
}

// A: hash the thread's C1 k-mers; candidates go to list[slot * NT + tid] in position order and set
// their bit in *mask.  Returns false when the private list would overflow (tile -> dense path).
template <int NT, int C1, int CAP>
SW_HD bool sparseA_hash(int tid, const SketchParams& P, const Tile& T, const SparseSmem& S, uint64_t* mask_out)
{
    static_assert(C1 <= 64 && CAP > (int)kSparseCheck, "mask is 64 bits; the list needs room for one check interval");
    *mask_out = 0;
    const uint32_t j0 = (uint32_t)tid * C1;
    if (j0 >= T.n_kmers) return true;
    const uint32_t* W = P.words + P.rec_word_off[T.rec];
    const Piece pc = P.pieces[T.piece_lo];
    // reads may run past the tile's last k-mer (kTailPadWords of slack); those steps are masked off below
    const uint64_t p0 = (uint64_t)pc.pos + ((uint64_t)T.e0 + j0 - pc.kidx);
    uint64_t fwd, rev;
    seed_kmer(W, p0, P.k, S.tab, P.tetra, &fwd, &rev);
    // private list cursor: a 32-bit shared-memory address on the device (one predicated add per
    // candidate instead of 64-bit pointer arithmetic), a plain pointer in the host emulator
#if defined(__CUDA_ARCH__)
    uint32_t slot = (uint32_t)__cvta_generic_to_shared(S.list + tid);
    const uint32_t slot_limit = slot + (uint32_t)((CAP - (int)kSparseCheck) * NT * sizeof(uint64_t));
#define SW_LIST_PUSH(h) do { asm volatile("st.shared.u64 [%0], %1;" :: "r"(slot), "l"(h) : "memory"); slot += NT * 8; } while (0)
#else
    uint64_t* slot = S.list + tid;
    uint64_t* const slot_limit = slot + (size_t)(CAP - (int)kSparseCheck) * NT;   // room for one more interval
#define SW_LIST_PUSH(h) do { *slot = (h); slot += NT; } while (0)
#endif
    uint32_t thr = P.cand_hi;
    uint64_t mask = 0;
    bool ok = true;
    {
        const uint64_t h = fwd + rev;
        if ((uint32_t)(h >> 32) < thr) { SW_LIST_PUSH(h); mask |= 1u; }
    }
    // 16 rolling steps (one word of bases in, one out) unrolled, the words in a loop: the unrolled body of all
    // C1 steps would be most of the kernel's instruction footprint
    constexpr int NB = (C1 - 1 + 15) / 16;
#pragma unroll 1
    for (int b = 0; b < NB; ++b) {
        const uint32_t in = fetch16(W, p0 + P.k + 16 * b);
        const uint32_t out = fetch16(W, p0 + 16 * b);
        const uint32_t x = (in & 0x33333333u) | ((out & 0x33333333u) << 2);
        const uint32_t y = ((in >> 2) & 0x33333333u) | (out & 0xCCCCCCCCu);
        uint32_t m16 = 0;
#pragma unroll
        for (int s = 0; s < 16; ++s) {
            // step n = 16 b + s + 1; the last word may be a partial one
            if ((C1 - 1) % 16 != 0 && s >= (C1 - 1) % 16 && b == NB - 1) break;
            if ((s + 1) % (int)kSparseCheck == 0 && 16 * b + s + 1 > CAP - (int)kSparseCheck) {
                if (slot > slot_limit) { ok = false; thr = 0; }   // stop recording: the tile is recomputed
            }
            const uint32_t idx = (((s & 1) ? y : x) >> (4 * (s >> 1))) & 15u;
            roll_step(fwd, rev, S.tab[idx]);
            const uint64_t h = fwd + rev;
            if ((uint32_t)(h >> 32) < thr) {
                SW_LIST_PUSH(h);
                m16 |= 1u << s;
            }
        }
        mask |= (uint64_t)m16 << (16 * b + 1);
    }
#undef SW_LIST_PUSH
    const uint32_t lim = T.n_kmers - j0;   // k-mers of this chunk that exist
    if (lim < (uint32_t)C1) mask &= (1ULL << lim) - 1;
    *mask_out = mask;
    return ok;
}

SW_HD uint32_t popcount64(uint64_t x)
{
#if defined(__CUDA_ARCH__)
    return (uint32_t)__popcll(x);
#else
    return (uint32_t)__builtin_popcountll(x);
#endif
}

SW_HD uint32_t ctz64(uint64_t x)
{
#if defined(__CUDA_ARCH__)
    return (uint32_t)(__ffsll((long long)x) - 1);
#else
    return (uint32_t)__builtin_ctzll(x);
#endif
}

// how many of the thread's `cnt` listed candidates are small
template <int NT>
SW_HD uint32_t sparse_count_small(int tid, uint32_t cnt, const SketchParams& P, const SparseSmem& S)
{
    uint32_t c = 0;
    for (uint32_t s = 0; s < cnt; ++s) c += (uint32_t)(S.list[(size_t)s * NT + tid] >> 32) < P.cand_hi_a ? 1u : 0u;
    return c;
}

// C: copy the thread's candidates to key[1 + off ...] (and the small ones to akey[1 + aoff ...]);
// the first threads write the sentinels (h0 word 0 stops every walk, position -1 / n_kmers is the
// bound it then reports) and clear the selection bits.
template <int NT, int C1>
SW_HD void sparseC_compact(int tid, uint64_t mask, uint32_t off, uint32_t aoff, uint32_t m, uint32_t ma,
                           const SketchParams& P, const Tile& T, const SparseSmem& S)
{
    const uint32_t j0 = (uint32_t)tid * C1;
    uint32_t s = 0, sa = 0;
    while (mask) {
        const uint32_t b = ctz64(mask);
        mask &= mask - 1;
        const uint64_t h = S.list[(size_t)s * NT + tid];
        const uint64_t key = ((uint64_t)(j0 + b) << 32) | (h >> 32);
        const uint32_t j = 1 + off + s;
        S.key[j] = key;
        S.lo[j] = (uint32_t)h;
        if ((uint32_t)(h >> 32) < P.cand_hi_a) {
            const uint32_t a = 1 + aoff + sa;
            S.akey[a] = key;
            S.aj[a] = (uint16_t)j;
            ++sa;
        }
        S.na[j] = (uint16_t)(aoff + sa);
        ++s;
    }
    if (tid == 0) {
        const uint64_t left = 0xFFFFFFFFull << 32, right = (uint64_t)T.n_kmers << 32;
        S.key[0] = left;
        S.key[m + 1] = right;
        S.akey[0] = left;
        S.akey[ma + 1] = right;
        S.lo[0] = S.lo[m + 1] = 0;
        S.aj[0] = 0;
        S.aj[ma + 1] = (uint16_t)(m + 1);
    }
    for (uint32_t i = (uint32_t)tid; i < sparse_mc(NT) / 32 + 2; i += NT) S.selbits[i] = 0;
}

// Walk left from entry j of a candidate array to the nearest strictly smaller h0, giving up once w
// positions away; returns its position (-1 from the sentinel; anything <= p - w means "none in reach").
// SMALL: the array is akey[], whose low words are found through aj[].
template <bool SMALL>
SW_HD int32_t sparse_walk_left(const SparseSmem& S, uint32_t j, int32_t p, uint32_t hh, uint32_t hl, int32_t w)
{
    const uint64_t* key = SMALL ? S.akey : S.key;
    const int32_t reach = p - w;
    uint32_t q = j - 1;
    for (;;) {
        const uint64_t kq = key[q];
        const int32_t pq = (int32_t)(kq >> 32);
        if ((uint32_t)kq <= hh || pq <= reach) {
            // equal high words are rare: only then the low words (and the sentinel's index) matter
            if ((uint32_t)kq < hh || pq <= reach || q == 0 || S.lo[SMALL ? S.aj[q] : q] < hl) return pq;
        }
        --q;
    }
}

// Walk right to the nearest smaller-or-equal h0 (ties: the right-hand k-mer wins, minimizer.cpp:75).
template <bool SMALL>
SW_HD int32_t sparse_walk_right(const SparseSmem& S, uint32_t j, uint32_t last, int32_t p, uint32_t hh, uint32_t hl,
                                int32_t w)
{
    const uint64_t* key = SMALL ? S.akey : S.key;
    const int32_t reach = p + w;
    uint32_t q = j + 1;
    for (;;) {
        const uint64_t kq = key[q];
        const int32_t pq = (int32_t)(kq >> 32);
        if ((uint32_t)kq <= hh || pq >= reach) {
            if ((uint32_t)kq < hh || pq >= reach || q == last || S.lo[SMALL ? S.aj[q] : q] <= hl) return pq;
        }
        ++q;
    }
}

// is the k-mer at p, boxed in by L and R, the selection of a window this tile emits for?
SW_HD bool sparse_selected(int32_t p, int32_t L, int32_t R, int32_t w, const Tile& T)
{
    const int32_t i_lo = L + 1 > p - w + 1 ? L + 1 : p - w + 1;
    const int32_t i_hi = p < R - w ? p : R - w;
    // window 0 of a tile that is not the record's first only provides the previous selection
    return i_lo <= i_hi && (T.first != 0 || i_lo >= 1);
}

// S (small pass): thread handles small candidates 1 + tid, 1 + tid + NT, ... of ma, walking akey[].
template <int NT>
SW_HD void sparseS_small(int tid, uint32_t ma, const SketchParams& P, const Tile& T, const SparseSmem& S)
{
    const int32_t w = (int32_t)P.w;
    for (uint32_t a = 1 + (uint32_t)tid; a <= ma; a += NT) {
        const uint64_t ka = S.akey[a];
        const int32_t p = (int32_t)(ka >> 32);
        const uint32_t j = S.aj[a];
        const uint32_t hh = (uint32_t)ka, hl = S.lo[j];
        const int32_t L = sparse_walk_left<true>(S, a, p, hh, hl, w);
        const int32_t R = sparse_walk_right<true>(S, a, ma + 1, p, hh, hl, w);
        if (sparse_selected(p, L, R, w, T)) {
#if defined(__CUDA_ARCH__)
            atomicOr(&S.selbits[j >> 5], 1u << (j & 31));
#else
            S.selbits[j >> 5] |= 1u << (j & 31);
#endif
        }
    }
}

// S (main pass): thread owns candidates [1 + tid * per, 1 + (tid + 1) * per) of m; the small ones
// are skipped.  Returns the selection flags of its other candidates (bit i: candidate
// 1 + tid * per + i); *bad is set if some window of the tile holds no candidate.
SW_COLD void gap_register(GapList* G, int32_t a, int32_t b, uint32_t owner, bool* bad)
{
    if ((uint32_t)(b - a - 1) > kGapLenMax) { *bad = true; return; }
#if defined(__CUDA_ARCH__)
    const uint32_t i = atomicAdd(&G->n, 1u);
#else
    const uint32_t i = G->n++;
#endif
    if (i >= kGapMax) { *bad = true; return; }
    G->a[i] = (int16_t)a;
    G->b[i] = (int16_t)b;
    G->owner[i] = (uint16_t)owner;
}

template <int NT, bool GAPS = true>
SW_HD uint32_t sparseS_main(int tid, uint32_t m, uint32_t per, const SketchParams& P, const Tile& T,
                            const SparseSmem& S, GapList* G, bool* bad)
{
    const int32_t w = (int32_t)P.w, n = (int32_t)T.n_kmers;
    uint32_t f = 0;
    *bad = false;
    for (uint32_t i = 0; i < per; ++i) {
        const uint32_t j = 1 + (uint32_t)tid * per + i;
        if (j > m) break;
        const uint64_t kj = S.key[j];
        const int32_t p = (int32_t)(kj >> 32);
        const uint32_t hh = (uint32_t)kj;
        // coverage: no stretch of w k-mers without candidate before this one / after the last one
        const int32_t prev = (int32_t)(S.key[j - 1] >> 32);
        if (GAPS) {
            if (p - prev > w) gap_register(G, prev, p, j, bad);
            if (j == m && n - p > w) gap_register(G, p, n, m + 1, bad);
        } else if (p - prev > w || (j == m && n - p > w)) {
            *bad = true;   // this variant leaves such tiles to the one that settles stretches in place
        }
        if (hh < P.cand_hi_a) continue;
        // every small candidate is smaller: unless w k-mers fit between the nearest ones on either
        // side (about one window in fifty has no small candidate), this one is never selected
        const uint32_t a = S.na[j];
        if ((int32_t)(S.akey[a + 1] >> 32) - (int32_t)(S.akey[a] >> 32) - 1 < w) continue;
        const uint32_t hl = S.lo[j];
        const int32_t L = sparse_walk_left<false>(S, j, p, hh, hl, w);
        const int32_t R = sparse_walk_right<false>(S, j, m + 1, p, hh, hl, w);
        if (sparse_selected(p, L, R, w, T)) f |= 1u << i;
    }
    return f;
}

// after both passes: add the small candidates' selection bits to the thread's flags
SW_HD uint32_t sparse_merge_flags(int tid, uint32_t m, uint32_t per, uint32_t flags, const SparseSmem& S)
{
    const uint32_t j0 = 1 + (uint32_t)tid * per;
    if (j0 > m || per == 0) return flags;
    const uint64_t two = (uint64_t)S.selbits[j0 >> 5] | ((uint64_t)S.selbits[(j0 >> 5) + 1] << 32);
    return flags | ((uint32_t)(two >> (j0 & 31)) & ((1u << per) - 1u));
}

// D: write the thread's selected candidates to consecutive global slots.
template <int NT>
SW_HD void sparseD_write(int tid, uint32_t per, uint32_t flags, unsigned long long slot, const SketchParams& P,
                         const Tile& T, const SparseSmem& S)
{
    const Piece pc = P.pieces[T.piece_lo];
    while (flags) {
        const uint32_t i = ctz64(flags);
        flags &= flags - 1;
        const uint32_t j = 1 + (uint32_t)tid * per + i;
        const uint64_t kj = S.key[j];
        const uint64_t h = (kj << 32) | S.lo[j];
        if (slot < P.capacity) {
            P.out_key[slot] = h1_of(h, P.h1_mult);
            const uint32_t pos = pc.pos + (uint32_t)((uint64_t)T.e0 + (uint32_t)(kj >> 32) - pc.kidx);
            P.out_val[slot] = (uint64_t)pos | ((uint64_t)(P.rec_base + T.rec) << 32);
        }
        ++slot;
    }
}

// ---- stretches without a candidate ---------------------------------------------------------------------
struct GapWs {
    uint64_t* h;      // [kGapLenMax] h0 of the stretch's k-mers
    uint16_t* sel;    // [kGapLenMax] selection of every window of the stretch (index into h)
    uint64_t* eh;     // [kGapEmitMax] h0 of the minimizers to add
    uint16_t* ep;     // [kGapEmitMax] their tile-local k-mer index
};
SW_HD GapWs gap_workspace(const SparseSmem& S)
{
    GapWs g;
    g.h = S.list;
    g.sel = reinterpret_cast<uint16_t*>(S.list + kGapLenMax);
    g.eh = S.list + kGapLenMax + kGapLenMax / 4;
    g.ep = reinterpret_cast<uint16_t*>(g.eh + kGapEmitMax);
    return g;
}

// one thread: order the stretches by position (they were registered in any order)
SW_COLD void sparseG_sort(GapList* G)
{
    for (uint32_t i = 1; i < G->n; ++i)
        for (uint32_t j = i; j > 0 && G->a[j] < G->a[j - 1]; --j) {
            const int16_t ta = G->a[j], tb = G->b[j];
            const uint16_t to = G->owner[j];
            G->a[j] = G->a[j - 1]; G->b[j] = G->b[j - 1]; G->owner[j] = G->owner[j - 1];
            G->a[j - 1] = ta; G->b[j - 1] = tb; G->owner[j - 1] = to;
        }
    G->e0[0] = 0;
}

// thread t hashes k-mers [16 t, 16 t + 16), [16 (t + NT), ...) ... of stretch gi (one seed, then rolling steps)
template <int NT>
SW_COLD void sparseG_hash(int tid, uint32_t gi, const GapList& G, const SketchParams& P, const Tile& T, const SparseSmem& S)
{
    const GapWs ws = gap_workspace(S);
    const uint32_t g = (uint32_t)(G.b[gi] - G.a[gi] - 1);
    const uint32_t* W = P.words + P.rec_word_off[T.rec];
    const Piece pc = P.pieces[T.piece_lo];
    for (uint32_t q0 = (uint32_t)tid * 16; q0 < g; q0 += NT * 16) {
        const uint32_t q1 = q0 + 16 < g ? q0 + 16 : g;
        uint64_t p = (uint64_t)pc.pos + ((uint64_t)T.e0 + (uint32_t)(G.a[gi] + 1) + q0 - pc.kidx);
        uint64_t fwd, rev;
        seed_kmer(W, p, P.k, S.tab, P.tetra, &fwd, &rev);
        ws.h[q0] = fwd + rev;
        for (uint32_t q = q0 + 1; q < q1; ++q, ++p) {
            roll_step(fwd, rev, S.tab[(base_at(W, p) << 2) | base_at(W, p + P.k)]);
            ws.h[q] = fwd + rev;
        }
    }
}

// thread t evaluates windows t, t + NT, ... of the stretch: rightmost minimum of w consecutive k-mers
template <int NT>
SW_COLD void sparseG_windows(int tid, uint32_t gi, const GapList& G, const SketchParams& P, const SparseSmem& S)
{
    const GapWs ws = gap_workspace(S);
    const uint32_t g = (uint32_t)(G.b[gi] - G.a[gi] - 1), nw = g - P.w + 1;
    for (uint32_t i = (uint32_t)tid; i < nw; i += NT) {
        uint64_t best = ws.h[i];
        uint32_t arg = i;
        for (uint32_t x = 1; x < P.w; ++x) {
            const uint64_t h = ws.h[i + x];
            if (h <= best) { best = h; arg = i + x; }
        }
        ws.sel[i] = (uint16_t)arg;
    }
}

// one thread: the windows' selections, each once, in position order (minimizer.cpp:41-47)
SW_COLD void sparseG_emit(uint32_t gi, GapList* G, const SketchParams& P, const Tile& T, const SparseSmem& S)
{
    const GapWs ws = gap_workspace(S);
    const int32_t a = G->a[gi];
    const uint32_t g = (uint32_t)(G->b[gi] - a - 1), nw = g - P.w + 1;
    uint32_t e = G->e0[gi];
    for (uint32_t i = 0; i < nw; ++i) {
        const uint32_t sel = ws.sel[i];
        // the window before the stretch's first one selected the candidate at a; window 0 of a tile that
        // is not the record's first only provides the previous selection
        const bool emit = i == 0 ? (a >= 0 || T.first != 0) : sel != ws.sel[i - 1];
        if (!emit || ws.h[sel] == ~0ull) continue;
        if (e >= kGapEmitMax) { G->n |= kGapOverflow; break; }
        ws.eh[e] = ws.h[sel];
        ws.ep[e] = (uint16_t)((uint32_t)(a + 1) + sel);
        ++e;
    }
    G->e0[gi + 1] = (uint16_t)e;
}

// which thread writes the minimizers of stretch gi: the owner of the candidate after it
SW_HD uint32_t gap_owner_thread(const GapList& G, uint32_t gi, uint32_t m, uint32_t per)
{
    const uint32_t j = G.owner[gi] <= m ? G.owner[gi] : m;
    return (j - 1) / per;
}

SW_COLD uint32_t sparse_gap_count(int tid, uint32_t m, uint32_t per, const GapList& G)
{
    uint32_t c = 0;
    for (uint32_t gi = 0; gi < G.n; ++gi)
        if (gap_owner_thread(G, gi, m, per) == (uint32_t)tid) c += (uint32_t)G.e0[gi + 1] - G.e0[gi];
    return c;
}

SW_HD void gap_write(uint32_t gi, unsigned long long& slot, const GapList& G, const Piece& pc, const SketchParams& P,
                     const Tile& T, const SparseSmem& S)
{
    const GapWs ws = gap_workspace(S);
    for (uint32_t e = G.e0[gi]; e < G.e0[gi + 1]; ++e, ++slot) {
        if (slot < P.capacity) {
            P.out_key[slot] = h1_of(ws.eh[e], P.h1_mult);
            const uint32_t pos = pc.pos + (uint32_t)((uint64_t)T.e0 + ws.ep[e] - pc.kidx);
            P.out_val[slot] = (uint64_t)pos | ((uint64_t)(P.rec_base + T.rec) << 32);
        }
    }
}

// D with stretches: the thread's candidates in order, each preceded by the minimizers of the stretch before it
template <int NT>
SW_COLD void sparseD_write_gaps(int tid, uint32_t m, uint32_t per, uint32_t flags, unsigned long long slot, const GapList& G,
                              const SketchParams& P, const Tile& T, const SparseSmem& S)
{
    const Piece pc = P.pieces[T.piece_lo];
    for (uint32_t i = 0; i < per; ++i) {
        const uint32_t j = 1 + (uint32_t)tid * per + i;
        if (j > m) break;
        for (uint32_t gi = 0; gi < G.n; ++gi)
            if (G.owner[gi] == j) gap_write(gi, slot, G, pc, P, T, S);
        if (flags & (1u << i)) {
            const uint64_t kj = S.key[j];
            const uint64_t h = (kj << 32) | S.lo[j];
            if (slot < P.capacity) {
                P.out_key[slot] = h1_of(h, P.h1_mult);
                const uint32_t pos = pc.pos + (uint32_t)((uint64_t)T.e0 + (uint32_t)(kj >> 32) - pc.kidx);
                P.out_val[slot] = (uint64_t)pos | ((uint64_t)(P.rec_base + T.rec) << 32);
            }
            ++slot;
        }
        if (j == m)
            for (uint32_t gi = 0; gi < G.n; ++gi)
                if (G.owner[gi] == m + 1) gap_write(gi, slot, G, pc, P, T, S);
    }
}

}  // namespace sw
