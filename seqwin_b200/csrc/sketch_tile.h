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
    uint32_t n_tiles;
    uint32_t k, w, c2;
    uint32_t rec_base;             // global index of batch record 0
    uint64_t h1_mult;
    uint64_t* out_key;             // [capacity] h1 of each minimizer, (record, pos) order
    uint64_t* out_val;             // [capacity] pos | record_idx << 32
    unsigned long long capacity;
    unsigned long long* tile_status;  // decoupled look-back words, one per tile
    unsigned int* tile_counter;       // ticket dispenser
    unsigned long long* total_out;    // number of minimizers produced
    RollTable table;
};

// Shared-memory view of one tile (TK = NT * C1 k-mers).
struct TileSmem {
    RollEntry* tab;   // [20]
    uint64_t* h0;     // [TK]
    uint64_t* cm_h;   // [TK / 9 + 2]  chunk minima (value)
    uint16_t* cm_i;   // [TK / 9 + 2]  chunk minima (tile-local index)
    uint16_t* amin;   // [TK] A[a]
    uint16_t* pidx;   // [TK] in-chunk offset of the prefix argmin; reused as the emit staging list
};

SW_HD size_t tile_smem_bytes(uint32_t tk)
{
    const size_t nc = tk / 9 + 2;
    size_t b = sizeof(RollEntry) * 20 + sizeof(uint64_t) * tk + sizeof(uint64_t) * nc;
    b += sizeof(uint16_t) * (nc + 2 * (size_t)tk);
    return (b + 15) & ~(size_t)15;
}

SW_HD TileSmem carve_tile_smem(unsigned char* base, uint32_t tk)
{
    const size_t nc = tk / 9 + 2;
    TileSmem s;
    s.tab = reinterpret_cast<RollEntry*>(base);
    s.h0 = reinterpret_cast<uint64_t*>(base + sizeof(RollEntry) * 20);
    s.cm_h = s.h0 + tk;
    s.cm_i = reinterpret_cast<uint16_t*>(s.cm_h + nc);
    s.amin = s.cm_i + nc;
    s.pidx = s.amin + tk;
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
// k-mers starting at record position p0: k warm-up steps, then C1-1 rolling steps.
template <int C1>
SW_HD void hash_chunk_fast(const uint32_t* W, uint64_t p0, uint32_t k, const RollEntry* tab,
                           uint64_t* h0_out)
{
    uint64_t fwd = 0, rev = 0;
    for (uint32_t i = 0; i < k; i += 16) {
        const uint32_t x = fetch16(W, p0 + i);
        const uint32_t m = k - i;
#pragma unroll
        for (int s = 0; s < 16; ++s)
            if ((uint32_t)s < m) roll_step(fwd, rev, tab[16 + ((x >> (2 * s)) & 3u)]);
    }
    h0_out[0] = fwd + rev;
#pragma unroll
    for (int b = 0; b < (C1 - 1 + 15) / 16; ++b) {
        const uint32_t in = fetch16(W, p0 + k + 16 * b);
        const uint32_t out = fetch16(W, p0 + 16 * b);
#pragma unroll
        for (int s = 0; s < 16; ++s) {
            const int n = 16 * b + s + 1;
            if (n < C1) {
                const uint32_t idx = ((in >> (2 * s)) & 3u) | (((out >> (2 * s)) & 3u) << 2);
                roll_step(fwd, rev, tab[idx]);
                h0_out[n] = fwd + rev;
            }
        }
    }
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
        hash_chunk_fast<C1>(W, p0, P.k, S.tab, S.h0 + j0);
    } else {
        const uint32_t j1 = j0 + C1 < T.n_kmers ? j0 + C1 : T.n_kmers;
        hash_range_generic(W, pcs, T.e0, j0, j1, P.k, S.tab, S.h0);
    }
}

// ---- phase 2: window minima -------------------------------------------------------------------

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
        uint32_t bi = 0;
        S.pidx[base] = 0;
        for (uint32_t n = 1; n < nend; ++n) {
            const uint64_t h = S.h0[base + n];
            if (h <= bh) { bh = h; bi = n; }
            S.pidx[base + n] = (uint16_t)bi;
        }
        S.cm_h[c] = bh;
        S.cm_i[c] = (uint16_t)(base + bi);
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
            const uint32_t te = e / c2;
            uint64_t rh = sh;
            uint32_t ri = si;
            if (te == te_lo) {
                if (m_lo_valid && m_lo_h <= rh) { rh = m_lo_h; ri = m_lo_i; }
            } else {
                if (m_hi_h <= rh) { rh = m_hi_h; ri = m_hi_i; }
            }
            const uint32_t pe = te * c2 + S.pidx[e];
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

// ---- phase 3: emit ----------------------------------------------------------------------------

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

// 3c: write staged minimizer i of the tile to its global slot.
SW_HD void phase3c_write(uint32_t i, unsigned long long gbase, const SketchParams& P, const Tile& T,
                         const TileSmem& S)
{
    const unsigned long long slot = gbase + i;
    if (slot >= P.capacity) return;  // counted, not stored: the host re-runs with more room
    const uint32_t idx = S.pidx[i];
    P.out_key[slot] = h1_of(S.h0[idx], P.h1_mult);
    P.out_val[slot] = (uint64_t)kmer_pos(P, T, idx) | ((uint64_t)(P.rec_base + T.rec) << 32);
}

}  // namespace sw
