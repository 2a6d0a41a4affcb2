// agg_emul.cpp -- the aggregation kernels of seqwin_b200/csrc/agg.cuh executed on the CPU through the
// CUDA execution-model emulator (cuda_emul.h).  TEST INFRASTRUCTURE: the stable partition that radix.cu
// performs on the device is a std::stable_sort here; everything between the partition and the graph
// arrays is the device code itself.
#include "cuda_emul.h"

#include <numeric>

#include "../../seqwin_b200/csrc/agg.cuh"

using namespace sw;
using namespace sw::agg;

namespace {

template <typename V>
void stable_partition_top_bits(std::vector<uint64_t>& keys, std::vector<V>& vals, int shift)
{
    std::vector<uint32_t> idx(keys.size());
    std::iota(idx.begin(), idx.end(), 0u);
    std::stable_sort(idx.begin(), idx.end(), [&](uint32_t a, uint32_t b) { return (keys[a] >> shift) < (keys[b] >> shift); });
    std::vector<uint64_t> k2(keys.size());
    std::vector<V> v2(vals.size());
    for (size_t i = 0; i < idx.size(); ++i) {
        k2[i] = keys[idx[i]];
        v2[i] = vals[idx[i]];
    }
    keys.swap(k2);
    vals.swap(v2);
}

std::vector<unsigned long long> exclusive_scan(const std::vector<unsigned long long>& v)
{
    std::vector<unsigned long long> out(v.size() + 1, 0);
    for (size_t i = 0; i < v.size(); ++i) out[i + 1] = out[i] + v[i];
    return out;
}

}  // namespace

// the hash-range sample behind the bucket sizing: returns items per distinct key as the device estimates it
extern "C" double agg_emul_estimate(const uint64_t* keys, uint64_t n, int sbits)
{
    std::vector<unsigned long long> set(1ull << kSampleSetBits, kEmptyKey), out(2, 0);
    cuemu::launch(dim3(4), dim3(256), [&] { distinct_sample_kernel(keys, n, sbits, set.data(), out.data()); });
    return out[1] ? (double)out[0] / (double)out[1] : 1.0;
}

extern "C" long agg_emul_run(const uint64_t* stream_keys, const uint64_t* stream_vals, uint64_t M, const uint32_t* rec_asm,
                             uint32_t rec_base, const uint8_t* is_target, int score, uint32_t n_targets, uint32_t n_non_targets,
                             uint32_t per_bucket_nodes, uint32_t per_bucket_edges, uint32_t max_distinct_edges, sw_kmer* kmers_out,
                             sw_node* nodes_out, sw_edge* edges_out, uint64_t* n_nodes_out, uint64_t* n_edges_out,
                             uint64_t* n_overflow_out)
{
    *n_nodes_out = *n_edges_out = *n_overflow_out = 0;
    if (M == 0) return 0;
    // ---- nodes ----
    const int P = partition_bits(M, per_bucket_nodes);
    const int key_bits = 64 - P;
    const uint64_t nb = 1ull << P;
    std::vector<uint64_t> keys(stream_keys, stream_keys + M);
    std::vector<unsigned long long> vals(stream_vals, stream_vals + M);
    stable_partition_top_bits(keys, vals, key_bits);
    std::vector<uint32_t> start(nb + 1, 0xDEADBEEFu), bucket_d(nb, 0xABABABABu);
    cuemu::launch(dim3(3), dim3(256), [&] { bucket_bounds_kernel(keys.data(), M, key_bits, nb, start.data()); });
    std::vector<uint64_t> grp_keys(M);
    std::vector<uint32_t> grp_cnt(M);
    std::vector<uint16_t> item_rank(M, 0xEEEE);
    cuemu::launch(dim3((unsigned)nb), dim3(kNT), [&] {
        group_count_kernel(keys.data(), start.data(), key_bits, (uint32_t)kMaxDistinct, grp_keys.data(), grp_cnt.data(), bucket_d.data(),
                           item_rank.data());
    });
    std::vector<unsigned long long> d64(nb), tot(2, 0);
    cuemu::launch(dim3(2), dim3(256), [&] { bucket_counts_kernel(bucket_d.data(), start.data(), nb, d64.data(), nullptr, tot.data()); });
    if (tot[0]) return -1;   // a node bucket overflowed: the product falls back to the sort-based path
    std::vector<unsigned long long> grp_base = exclusive_scan(d64);
    const uint64_t U = grp_base[nb];
    std::vector<uint64_t> node_hash(U);
    PlaceArgs pa{item_rank.data(), start.data(), key_bits, grp_keys.data(), grp_cnt.data(), bucket_d.data(), grp_base.data()};
    NodeOut no{};
    no.vals = vals.data();
    no.placed = reinterpret_cast<unsigned long long*>(kmers_out);
    std::vector<uint32_t> node_asm(M);
    no.placed_asm = node_asm.data();
    no.nodes = nodes_out;
    no.node_hash = node_hash.data();
    no.rec_asm = rec_asm;
    no.rec_base = rec_base;
    no.is_target = is_target;
    no.inv_t = n_targets ? 1.0 / (double)n_targets : 0.0;
    no.inv_n = n_non_targets ? 1.0 / (double)n_non_targets : 0.0;
    no.counts_only = 0;
    if (score) cuemu::launch(dim3((unsigned)nb), dim3(kNT), [&] { group_place_kernel<NodeOut, true>(pa, no); });
    else cuemu::launch(dim3((unsigned)nb), dim3(kNT), [&] { group_place_kernel<NodeOut, false>(pa, no); });
    *n_nodes_out = U;

    // ---- edges ----
    int fbits = 1;
    while (fbits < 28 && (1ull << fbits) < U) ++fbits;
    std::vector<uint32_t> ftable((1ull << fbits) + 1);
    cuemu::launch(dim3(2), dim3(256), [&] { bucket_bounds_kernel(node_hash.data(), U, 64 - fbits, 1ull << fbits, ftable.data()); });
    int rank_bits = 1;
    while (rank_bits < 32 && (1ull << rank_bits) < U) ++rank_bits;
    const uint32_t n_blocks = (uint32_t)((M + kEmitItems - 1) / kEmitItems);
    std::vector<unsigned long long> block_cnt(n_blocks, 0);
    for (uint64_t i = 0; i + 1 < M; ++i)
        if ((stream_vals[i] >> 32) == (stream_vals[i + 1] >> 32)) ++block_cnt[i / kEmitItems];
    std::vector<unsigned long long> block_off = exclusive_scan(block_cnt);
    const uint64_t E = block_off[n_blocks];
    if (E == 0) return 0;
    std::vector<uint64_t> ekey(E);
    std::vector<uint32_t> easm(E);
    cuemu::launch(dim3(n_blocks), dim3(kNT), [&] {
        edge_emit_kernel(stream_keys, stream_vals, M, node_hash.data(), ftable.data(), 64 - fbits, rec_asm, rec_base,
                         block_off.data(), rank_bits, ekey.data(), easm.data(), 0, nullptr, nullptr);
    });
    const int Pe = partition_bits(E, per_bucket_edges);
    const int ekey_bits = 64 - Pe;
    const uint64_t neb = 1ull << Pe;
    stable_partition_top_bits(ekey, easm, ekey_bits);
    std::vector<uint32_t> estart(neb + 1), ebucket_d(neb);
    cuemu::launch(dim3(3), dim3(256), [&] { bucket_bounds_kernel(ekey.data(), E, ekey_bits, neb, estart.data()); });
    {   // the binary-search variant of the bounds must agree with the scan
        std::vector<uint32_t> estart2(neb + 1, 0xDEADBEEFu);
        cuemu::launch(dim3(3), dim3(256), [&] { bucket_search_kernel(ekey.data(), E, ekey_bits, neb, estart2.data()); });
        if (estart2 != estart) return -2;
    }
    std::vector<uint64_t> egrp_keys(E);
    std::vector<uint32_t> egrp_cnt(E);
    cuemu::launch(dim3((unsigned)neb), dim3(kNT), [&] {
        edge_group_kernel(ekey.data(), easm.data(), estart.data(), ekey_bits, max_distinct_edges, egrp_keys.data(), egrp_cnt.data(),
                          ebucket_d.data());
    });
    std::vector<unsigned long long> ed64(neb), ovf_items(neb), etot(2, 0);
    cuemu::launch(dim3(2), dim3(256), [&] { bucket_counts_kernel(ebucket_d.data(), estart.data(), neb, ed64.data(), ovf_items.data(), etot.data()); });
    std::vector<sw_edge> side_edges;
    std::vector<unsigned long long> ovf_d64(neb, 0), ovf_base;
    if (etot[0]) {
        // buckets with too many distinct pairs: the device sorts their records (radix sort) and run-length
        // encodes them; here the sort and the encoding are host code, the index plumbing is the kernels'
        *n_overflow_out = etot[0];
        std::vector<unsigned long long> side_off = exclusive_scan(ovf_items);
        std::vector<uint64_t> skeys(etot[1]);
        std::vector<uint32_t> sasm(etot[1]);
        cuemu::launch(dim3((unsigned)neb), dim3(kNT), [&] {
            overflow_gather_kernel<uint32_t>(ekey.data(), easm.data(), estart.data(), ebucket_d.data(), side_off.data(), skeys.data(), sasm.data());
        });
        stable_partition_top_bits(skeys, sasm, 0);
        cuemu::launch(dim3((unsigned)neb), dim3(kNT), [&] {
            overflow_count_kernel(skeys.data(), estart.data(), ebucket_d.data(), side_off.data(), ed64.data(), ovf_d64.data());
        });
        ovf_base = exclusive_scan(ovf_d64);
        for (size_t i = 0; i < skeys.size(); ++i) {
            if (i == 0 || skeys[i] != skeys[i - 1]) {
                sw_edge e;
                e.first = node_hash[skeys[i] >> (64 - rank_bits)];
                e.second = node_hash[(skeys[i] >> (64 - 2 * rank_bits)) & ((1ull << rank_bits) - 1)];
                e.weight = 1;
                side_edges.push_back(e);
            } else if (sasm[i] != sasm[i - 1]) {
                ++side_edges.back().weight;
            }
        }
    }
    std::vector<unsigned long long> egrp_base = exclusive_scan(ed64);
    const uint64_t UE = egrp_base[neb];
    cuemu::launch(dim3(2), dim3(256), [&] {
        edge_out_kernel(egrp_keys.data(), egrp_cnt.data(), estart.data(), ebucket_d.data(), egrp_base.data(), neb, node_hash.data(),
                          rank_bits, edges_out);
    });
    if (etot[0])
        cuemu::launch(dim3((unsigned)neb), dim3(kNT), [&] {
            overflow_copy_kernel(side_edges.data(), ebucket_d.data(), egrp_base.data(), ovf_base.data(), edges_out);
        });
    *n_edges_out = UE;
    return 0;
}
