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

struct EmulOut {
    std::vector<uint64_t> ukeys, uvals;
    std::vector<unsigned long long> tile_count, tile_slot;
    unsigned long long cursor = 0;
};

static SketchParams make_params(const sw_batch& b, const Plan& plan, uint32_t k, uint32_t w, uint32_t c1, EmulOut& o)
{
    SketchParams P;
    memset(&P, 0, sizeof P);
    P.words = b.words;
    P.rec_word_off = b.rec_word_off.data();
    P.tiles = plan.tiles.data();
    P.pieces = plan.pieces.data();
    P.n_tiles = (uint32_t)plan.tiles.size();
    P.k = k;
    P.w = w;
    P.c2 = choose_c2(w, c1);
    P.rec_base = 0;
    P.h1_mult = h1_multiplier(k);
    P.table = make_roll_table(k);
    static TetraTable tetra;
    make_tetra_table(tetra);
    P.tetra = &tetra;
    o.ukeys.assign(plan.n_windows, 0);
    o.uvals.assign(plan.n_windows, 0);
    o.tile_count.assign(P.n_tiles, 0);
    o.tile_slot.assign(P.n_tiles, 0);
    P.out_key = o.ukeys.data();
    P.out_val = o.uvals.data();
    P.capacity = plan.n_windows;
    return P;
}

// tiles complete in arbitrary order on the device: emulate a scrambled completion order
static std::vector<uint32_t> scrambled_order(uint32_t n)
{
    std::vector<uint32_t> order(n);
    for (uint32_t t = 0; t < n; ++t) order[t] = t;
    for (uint32_t t = 0; t + 1 < n; t += 2) std::swap(order[t], order[t + 1]);
    std::reverse(order.begin(), order.end());
    return order;
}

// reorder_kernel: exclusive scan of the counts in tile order, then segment copy
static void reorder(const EmulOut& o, std::vector<uint64_t>& keys, std::vector<uint64_t>& vals)
{
    keys.assign(o.cursor, 0);
    vals.assign(o.cursor, 0);
    unsigned long long off = 0;
    for (size_t t = 0; t < o.tile_count.size(); ++t) {
        for (unsigned long long i = 0; i < o.tile_count[t]; ++i) {
            keys[off + i] = o.ukeys[o.tile_slot[t] + i];
            vals[off + i] = o.uvals[o.tile_slot[t] + i];
        }
        off += o.tile_count[t];
    }
}

// one tile through the dense kernels' phases (sketch_fast_kernel / sketch_generic_kernel)
template <int NT, int C1>
static void dense_tile(const SketchParams& P, uint32_t t, bool fast, EmulOut& o)
{
    constexpr uint32_t TK = NT * C1;
    const uint32_t w = P.w;
    const uint32_t nc = fast ? (uint32_t)NT : TK / 9 + 2;
    std::vector<unsigned char> smem(tile_smem_bytes(TK, nc) + 64);
    TileSmem S = carve_tile_smem(smem.data(), TK, nc);
#ifdef EMUL_SPLIT_SMEM
    std::vector<RollEntry> v_tab(20);
    std::vector<uint64_t> v_h0(TK, 0xA5A5A5A5A5A5A5A5ull), v_cmh(nc);
    std::vector<uint16_t> v_cmi(nc), v_amin(TK), v_pidx(TK), v_first(nc);
    S = TileSmem{v_tab.data(), v_h0.data(), v_cmh.data(), v_cmi.data(), v_amin.data(), v_pidx.data(), v_first.data()};
#endif
    // poison shared memory between tiles so that reads of stale data show up as mismatches
    memset(smem.data(), 0xA5, smem.size());
    for (int i = 0; i < 20; ++i) S.tab[i] = P.table.e[i];
    const Tile T = P.tiles[t];
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
        o.tile_count[t] = total;
        o.tile_slot[t] = o.cursor;
        for (int tid = 0; tid < NT; ++tid)
            if (cnt[tid]) fastD_write<NT, C1>(P, T, S, st[tid], o.cursor + excl[tid]);
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
        o.tile_count[t] = total;
        o.tile_slot[t] = o.cursor;
        for (uint32_t i = 0; i < total; ++i) phase3c_write(i, o.cursor, P, T, S);
    }
    o.cursor += total;
}

template <int NT, int C1>
static void emulate(const sw_batch& b, uint32_t k, uint32_t w, std::vector<uint64_t>& keys,
                    std::vector<uint64_t>& vals, uint32_t* n_tiles_out, int force_generic)
{
    constexpr uint32_t TK = NT * C1;
    Plan plan = plan_tiles(b, k, w, TK);
    *n_tiles_out = (uint32_t)plan.tiles.size();
    const bool fast = (w - 1 >= (uint32_t)C1) && !force_generic;
    EmulOut o;
    const SketchParams P = make_params(b, plan, k, w, C1, o);
    for (uint32_t t : scrambled_order(P.n_tiles)) dense_tile<NT, C1>(P, t, fast, o);
    reorder(o, keys, vals);
}

// sketch_sparse_kernel<NT, C1, CAP>, with sketch_fast/generic_kernel<FNT, FC1> for the tiles it hands over
template <int NT, int C1, int CAP, int FNT, int FC1>
static void emulate_sparse(const sw_batch& b, uint32_t k, uint32_t w, double cand_per_window, std::vector<uint64_t>& keys,
                           std::vector<uint64_t>& vals, uint32_t* n_tiles_out, uint32_t* n_fallback_out)
{
    constexpr uint32_t TK = NT * C1;
    static_assert(FNT * FC1 >= NT * C1, "the hand-over configuration must hold a sparse tile");
    Plan plan = plan_tiles(b, k, w, TK);
    *n_tiles_out = (uint32_t)plan.tiles.size();
    EmulOut o;
    SketchParams P = make_params(b, plan, k, w, FC1, o);
    // cand_per_window <= 0: the library's thresholds; else that many candidates, a third of them small
    P.cand_hi = sparse_threshold(w, cand_per_window <= 0 ? sparse_cand_per_window(w) : cand_per_window);
    P.cand_hi_a = sparse_threshold(w, cand_per_window <= 0 ? sparse_small_per_window(w) : cand_per_window / 3.0);
    constexpr uint32_t MC = sparse_mc(NT), MA = sparse_ma(NT);
    std::vector<unsigned char> smem(sparse_smem_bytes(NT, CAP) + 64);
    SparseSmem S = carve_sparse_smem(smem.data(), NT, CAP);
#ifdef EMUL_SPLIT_SMEM
    // manual check (g++ -DEMUL_SPLIT_SMEM -fsanitize=address): every shared array in its own heap
    // block of exactly its size, so that AddressSanitizer sees an overrun from one array into the next
    const size_t mc = sparse_mc(NT) + 2, ma = sparse_ma(NT) + 2;
    std::vector<RollEntry> v_tab(20);
    std::vector<uint64_t> v_list(sparse_list_words(NT, CAP)), v_key(mc), v_akey(ma);
    std::vector<uint32_t> v_lo(mc), v_sel(mc / 32 + 2);
    std::vector<uint16_t> v_aj(ma), v_na(mc);
    S = SparseSmem{v_tab.data(), v_list.data(), v_key.data(), v_akey.data(), v_lo.data(), v_sel.data(), v_aj.data(), v_na.data()};
#endif
    std::vector<uint32_t> fallback;
    uint32_t n_gap_tiles = 0;
    for (uint32_t t : scrambled_order(P.n_tiles)) {
        memset(smem.data(), 0xA5, smem.size());
        for (int i = 0; i < 20; ++i) S.tab[i] = P.table.e[i];
        const Tile T = P.tiles[t];
        std::vector<uint64_t> mask(NT, 0);
        GapList gaps;
        memset(&gaps, 0, sizeof(gaps));
        bool hand_over = T.n_pieces != 1;
        if (!hand_over)
            for (int tid = 0; tid < NT; ++tid)
                if (!sparseA_hash<NT, C1, CAP>(tid, P, T, S, &mask[tid])) hand_over = true;
        std::vector<uint32_t> off(NT, 0), aoff(NT, 0), flags(NT, 0), excl(NT, 0);
        uint32_t m = 0, ma = 0;
        if (!hand_over) {
            for (int tid = 0; tid < NT; ++tid) {
                const uint32_t c = (uint32_t)__builtin_popcountll(mask[tid]);
                off[tid] = m;
                aoff[tid] = ma;
                m += c;
                ma += sparse_count_small<NT>(tid, c, P, S);
            }
            hand_over = m == 0 || m > MC || ma > MA;
        }
        const uint32_t per = (m + NT - 1) / NT;
        if (!hand_over) {
            for (int tid = 0; tid < NT; ++tid) sparseC_compact<NT, C1>(tid, mask[tid], off[tid], aoff[tid], m, ma, P, T, S);
            for (int tid = 0; tid < NT; ++tid) {
                bool bad;
                flags[tid] = sparseS_main<NT>(tid, m, per, P, T, S, &gaps, &bad);
                sparseS_small<NT>(tid, ma, P, T, S);
                if (bad) hand_over = true;
            }
            if (!hand_over && gaps.n != 0) {
                sparseG_sort(&gaps);
                const uint32_t n_gaps = gaps.n;   // the emit phase may set the overflow bit in n
                for (uint32_t gi = 0; gi < n_gaps; ++gi) {
                    for (int tid = 0; tid < NT; ++tid) sparseG_hash<NT>(tid, gi, gaps, P, T, S);
                    for (int tid = 0; tid < NT; ++tid) sparseG_windows<NT>(tid, gi, gaps, P, S);
                    sparseG_emit(gi, &gaps, P, T, S);
                }
                hand_over = (gaps.n & kGapOverflow) != 0;
                if (!hand_over) ++n_gap_tiles;
            }
        }
        if (hand_over) { fallback.push_back(t); continue; }
        uint32_t total = 0;
        for (int tid = 0; tid < NT; ++tid) {
            flags[tid] = sparse_merge_flags(tid, m, per, flags[tid], S);
            excl[tid] = total;
            total += (uint32_t)__builtin_popcount(flags[tid]) + (gaps.n ? sparse_gap_count(tid, m, per, gaps) : 0u);
        }
        o.tile_count[t] = total;
        o.tile_slot[t] = o.cursor;
        for (int tid = 0; tid < NT; ++tid) {
            if (gaps.n) sparseD_write_gaps<NT>(tid, m, per, flags[tid], o.cursor + excl[tid], gaps, P, T, S);
            else if (flags[tid]) sparseD_write<NT>(tid, per, flags[tid], o.cursor + excl[tid], P, T, S);
        }
        o.cursor += total;
    }
    *n_fallback_out = (uint32_t)fallback.size();
    if (getenv("EMUL_GAP_STATS")) fprintf(stderr, "[emul] %u tiles, %zu handed over, %u settled stretches in place\n",
                                          P.n_tiles, fallback.size(), n_gap_tiles);
    const bool fast = w - 1 >= (uint32_t)FC1;
    for (uint32_t t : fallback) dense_tile<FNT, FC1>(P, t, fast, o);
    reorder(o, keys, vals);
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

// sparse kernel + dense hand-over; cand_per_window <= 0 uses the library's threshold
long emul_sketch_sparse(const uint8_t* const* seqs, const uint32_t* lens, size_t n_records, uint32_t k, uint32_t w,
                        int nt, int c1, double cand_per_window, uint64_t* h1_out, uint32_t* pos_out, uint32_t* rec_out,
                        size_t cap, uint32_t* n_tiles_out, uint32_t* n_fallback_out)
{
    try {
        std::vector<uint32_t> asm_of(n_records, 0);
        sw_batch* b = sw::batch_from_memory(seqs, lens, asm_of.data(), nullptr, n_records, 1, 1);
        std::vector<uint64_t> keys, vals;
        if (nt == 128 && c1 == 64)
            sw::emulate_sparse<128, 64, 20, 256, 33>(*b, k, w, cand_per_window, keys, vals, n_tiles_out, n_fallback_out);
        else if (nt == 8 && c1 == 64)
            sw::emulate_sparse<8, 64, 24, 16, 33>(*b, k, w, cand_per_window, keys, vals, n_tiles_out, n_fallback_out);
        else if (nt == 16 && c1 == 32)
            sw::emulate_sparse<16, 32, 12, 32, 17>(*b, k, w, cand_per_window, keys, vals, n_tiles_out, n_fallback_out);
        else if (nt == 4 && c1 == 48)
            sw::emulate_sparse<4, 48, 16, 8, 25>(*b, k, w, cand_per_window, keys, vals, n_tiles_out, n_fallback_out);
        else { delete b; return -2; }
        delete b;
        for (size_t i = 0; i < keys.size() && i < cap; ++i) {
            h1_out[i] = keys[i];
            pos_out[i] = (uint32_t)vals[i];
            rec_out[i] = (uint32_t)(vals[i] >> 32);
        }
        return (long)keys.size();
    } catch (const std::exception& e) {
        fprintf(stderr, "emul_sketch_sparse: %s\n", e.what());
        return -1;
    }
}

// the same from FASTA / .gz files through the host ingest (parse_fasta + packer), one assembly per file
long emul_sketch_files(const char* const* paths, size_t n_paths, uint32_t k, uint32_t w, int sparse, uint64_t* h1_out,
                       uint32_t* pos_out, uint32_t* rec_out, size_t cap)
{
    try {
        sw_batch* b = sw::batch_from_fasta(paths, n_paths, 2);
        std::vector<uint64_t> keys, vals;
        uint32_t n_tiles = 0, n_fb = 0;
        if (sparse) sw::emulate_sparse<8, 64, 24, 16, 33>(*b, k, w, 0.0, keys, vals, &n_tiles, &n_fb);
        else sw::emulate<32, 9>(*b, k, w, keys, vals, &n_tiles, 0);
        delete b;
        for (size_t i = 0; i < keys.size() && i < cap; ++i) {
            h1_out[i] = keys[i];
            pos_out[i] = (uint32_t)vals[i];
            rec_out[i] = (uint32_t)(vals[i] >> 32);
        }
        return (long)keys.size();
    } catch (const std::exception& e) {
        fprintf(stderr, "emul_sketch_files: %s\n", e.what());
        return -1;
    }
}
}
