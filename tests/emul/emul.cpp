// emul.cpp -- TEST-ONLY host emulator of the sketch kernel.
//
// Compiles the very same per-thread phase functions the CUDA kernel runs
// (seqwin_b200/csrc/sketch_tile.h) with g++ and executes them serially: one loop over tid per
// phase, a loop boundary standing in for each __syncthreads().  This lets the CPU test suite
// check tiling, halos, gap handling and tie-breaking against the oracle without a GPU.
// It is never linked into libseqwin_b200.so and nothing in seqwin_b200/ can reach it.
#include <cstdlib>
#include <algorithm>
#include <cstring>
#include <vector>

#include "../../seqwin_b200/csrc/ingest.cpp"
#include "../../seqwin_b200/csrc/sketch_tile.h"

namespace sw {
void* alloc_host(size_t bytes, bool* pinned)
{
    *pinned = false;
    return aligned_alloc(64, ((bytes ? bytes : 1) + 63) / 64 * 64);
}
void free_host(void* p, bool) { free(p); }

template <int NT, int C1>
static void emulate(const sw_batch& b, uint32_t k, uint32_t w, std::vector<uint64_t>& keys,
                    std::vector<uint64_t>& vals, uint32_t* n_tiles_out, int force_generic)
{
    constexpr uint32_t TK = NT * C1;
    Plan plan = plan_tiles(b, k, w, TK);
    *n_tiles_out = (uint32_t)plan.tiles.size();
    const bool fast = (w - 1 >= (uint32_t)C1) && !force_generic;
    const uint32_t nc = fast ? (uint32_t)NT : TK / 9 + 2;
    std::vector<unsigned char> smem(tile_smem_bytes(TK, nc) + 64);
    TileSmem S = carve_tile_smem(smem.data(), TK, nc);
    SketchParams P;
    memset(&P, 0, sizeof P);
    P.words = b.words;
    P.rec_word_off = b.rec_word_off.data();
    P.tiles = plan.tiles.data();
    P.pieces = plan.pieces.data();
    P.n_tiles = (uint32_t)plan.tiles.size();
    P.k = k;
    P.w = w;
    P.c2 = choose_c2(w, C1);
    P.rec_base = 0;
    P.h1_mult = h1_multiplier(k);
    P.table = make_roll_table(k);
    static TetraTable tetra;
    make_tetra_table(tetra);
    P.tetra = &tetra;
    std::vector<uint64_t> ukeys(plan.n_windows, 0), uvals(plan.n_windows, 0);
    std::vector<unsigned long long> tile_count(P.n_tiles, 0), tile_slot(P.n_tiles, 0);
    P.out_key = ukeys.data();
    P.out_val = uvals.data();
    P.capacity = plan.n_windows;
    unsigned long long cursor = 0;
    // tiles complete in arbitrary order on the device: emulate a scrambled completion order
    std::vector<uint32_t> order(P.n_tiles);
    for (uint32_t t = 0; t < P.n_tiles; ++t) order[t] = t;
    for (uint32_t t = 0; t + 1 < P.n_tiles; t += 2) std::swap(order[t], order[t + 1]);
    std::reverse(order.begin(), order.end());
    for (uint32_t oi = 0; oi < P.n_tiles; ++oi) {
        const uint32_t t = order[oi];
        // poison shared memory between tiles so that reads of stale data show up as mismatches
        memset(smem.data(), 0xA5, smem.size());
        for (int i = 0; i < 20; ++i) S.tab[i] = P.table.e[i];
        const Tile T = plan.tiles[t];
        std::vector<uint32_t> excl(NT);
        uint32_t total = 0;
        if (fast) {
            std::vector<FastState> st(NT);
            std::vector<uint32_t> cnt(NT, 0), active;
            for (int tid = 0; tid < NT; ++tid) fastA_hash_prefix<NT, C1>(tid, P, T, S);
            for (int tid = 0; tid < NT; ++tid) fastB1_boundary<NT, C1>(tid, P, T, S);
            for (int tid = 0; tid < NT; ++tid)
                if (fast_chunk_active<NT, C1>(tid, P, T, S)) active.push_back((uint32_t)tid);
            for (size_t i = 0; i < active.size(); ++i) cnt[i] = fastB2_windows<NT, C1>(active[i], P, T, S, st[i]);
            for (int tid = 0; tid < NT; ++tid) { excl[tid] = total; total += cnt[tid]; }
            tile_count[t] = total;
            tile_slot[t] = cursor;
            for (int tid = 0; tid < NT; ++tid)
                if (cnt[tid]) fastD_write<NT, C1>(P, T, S, st[tid], cursor + excl[tid]);
        } else {
            for (int tid = 0; tid < NT; ++tid) phase1_hash<NT, C1>(tid, P, T, S);
            if (P.c2) {
                for (int tid = 0; tid < NT; ++tid) phase2a_prefix<NT>(tid, P, T, S);
                for (int tid = 0; tid < NT; ++tid) phase2b_windows<NT>(tid, P, T, S);
            } else {
                for (int tid = 0; tid < NT; ++tid) phase2_direct<NT>(tid, P, T, S);
            }
            const uint32_t n_eval = T.n_kmers - w + 1;
            const uint32_t c3 = (n_eval + NT - 1) / NT;
            std::vector<uint64_t> masks(NT);
            for (int tid = 0; tid < NT; ++tid) masks[tid] = phase3a_flags(tid, c3, T, w, S);
            for (int tid = 0; tid < NT; ++tid) { excl[tid] = total; total += (uint32_t)__builtin_popcountll(masks[tid]); }
            for (int tid = 0; tid < NT; ++tid) phase3b_stage(tid, c3, masks[tid], excl[tid], S);
            tile_count[t] = total;
            tile_slot[t] = cursor;
            for (uint32_t i = 0; i < total; ++i) phase3c_write(i, cursor, P, T, S);
        }
        cursor += total;
    }
    // reorder_kernel: exclusive scan of the counts in tile order, then segment copy
    keys.assign(cursor, 0);
    vals.assign(cursor, 0);
    unsigned long long off = 0;
    for (uint32_t t = 0; t < P.n_tiles; ++t) {
        for (unsigned long long i = 0; i < tile_count[t]; ++i) {
            keys[off + i] = ukeys[tile_slot[t] + i];
            vals[off + i] = uvals[tile_slot[t] + i];
        }
        off += tile_count[t];
    }
}
}  // namespace sw

extern "C" {

// records given as ASCII; returns the number of minimizers, fills up to cap entries
long emul_sketch(const uint8_t* const* seqs, const uint32_t* lens, size_t n_records, uint32_t k, uint32_t w,
                 int nt, int c1, int force_generic, uint64_t* h1_out, uint32_t* pos_out, uint32_t* rec_out, size_t cap,
                 uint32_t* n_tiles_out)
{
    try {
        std::vector<uint32_t> asm_of(n_records, 0);
        sw_batch* b = sw::batch_from_memory(seqs, lens, asm_of.data(), nullptr, n_records, 1, 1);
        std::vector<uint64_t> keys, vals;
        if (nt == 128 && c1 == 45) sw::emulate<128, 45>(*b, k, w, keys, vals, n_tiles_out, force_generic);
        else if (nt == 256 && c1 == 21) sw::emulate<256, 21>(*b, k, w, keys, vals, n_tiles_out, force_generic);
        else if (nt == 256 && c1 == 61) sw::emulate<256, 61>(*b, k, w, keys, vals, n_tiles_out, force_generic);
        else if (nt == 8 && c1 == 11) sw::emulate<8, 11>(*b, k, w, keys, vals, n_tiles_out, force_generic);
        else if (nt == 4 && c1 == 45) sw::emulate<4, 45>(*b, k, w, keys, vals, n_tiles_out, force_generic);
        else if (nt == 128 && c1 == 33) sw::emulate<128, 33>(*b, k, w, keys, vals, n_tiles_out, force_generic);
        else if (nt == 256 && c1 == 17) sw::emulate<256, 17>(*b, k, w, keys, vals, n_tiles_out, force_generic);
        else if (nt == 128 && c1 == 27) sw::emulate<128, 27>(*b, k, w, keys, vals, n_tiles_out, force_generic);
        else if (nt == 32 && c1 == 9) sw::emulate<32, 9>(*b, k, w, keys, vals, n_tiles_out, force_generic);
        else { delete b; return -2; }
        delete b;
        for (size_t i = 0; i < keys.size() && i < cap; ++i) {
            h1_out[i] = keys[i];
            pos_out[i] = (uint32_t)vals[i];
            rec_out[i] = (uint32_t)(vals[i] >> 32);
        }
        return (long)keys.size();
    } catch (const std::exception& e) {
        fprintf(stderr, "emul_sketch: %s\n", e.what());
        return -1;
    }
}
}
