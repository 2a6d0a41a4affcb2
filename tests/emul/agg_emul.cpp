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
                             uint32_t per_bucket_nodes, uint32_t max_distinct_edges, sw_kmer* kmers_out,
                             sw_node* nodes_out, sw_edge* edges_out, uint64_t* n_nodes_out, uint64_t* n_edges_out,
                             uint64_t* n_overflow_out)
{
    *n_nodes_out = *n_edges_out = *n_overflow_out = 0;
    if (M == 0) return 0;
    // ---- nodes ----
    const int P = partition_bits(M, per_bucket_nodes);
    const int key_bits = 64 - P;
    const uint64_t nb = 1ull << P;
    // what the first partition pass generates on the device (radix.cu): the owned neighbour hashes of every item
    std::vector<uint64_t> prev(M), next(M);
    for (uint64_t g = 0; g < M; ++g) {
        if (stream_keys[g] == 0) return -3;   // the product takes the sort-based path
        const uint32_t rec = (uint32_t)(stream_vals[g] >> 32);
        prev[g] = owned_prev(stream_keys, stream_vals, g, stream_keys[g], rec);
        next[g] = owned_next(stream_keys, stream_vals, g, M, stream_keys[g], rec);
    }
    std::vector<uint64_t> keys(stream_keys, stream_keys + M);
    std::vector<unsigned long long> vals(stream_vals, stream_vals + M);
    {   // stable partition of the four arrays on the top P bits
        std::vector<uint32_t> idx(M);
        std::iota(idx.begin(), idx.end(), 0u);
        std::stable_sort(idx.begin(), idx.end(), [&](uint32_t x, uint32_t y) { return (keys[x] >> key_bits) < (keys[y] >> key_bits); });
        std::vector<uint64_t> k2(M), p2(M), n2(M);
        std::vector<unsigned long long> v2(M);
        for (uint64_t i = 0; i < M; ++i) {
            k2[i] = keys[idx[i]];
            v2[i] = vals[idx[i]];
            p2[i] = prev[idx[i]];
            n2[i] = next[idx[i]];
        }
        keys.swap(k2);
        vals.swap(v2);
        prev.swap(p2);
        next.swap(n2);
    }
    std::vector<uint32_t> start(nb + 1, 0xDEADBEEFu), bucket_d(nb, 0xABABABABu);
    cuemu::launch(dim3(3), dim3(256), [&] { bucket_search_kernel(keys.data(), M, key_bits, nb, start.data()); });
    {   // the scan variant of the bounds must agree with the binary search
        std::vector<uint32_t> start2(nb + 1, 0xDEADBEEFu);
        cuemu::launch(dim3(3), dim3(256), [&] { bucket_bounds_kernel(keys.data(), M, key_bits, nb, start2.data()); });
        if (start2 != start) return -2;
    }
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
    PlaceArgs pa{item_rank.data(), start.data(), key_bits, grp_keys.data(), grp_cnt.data(), bucket_d.data(), grp_base.data()};
    NodeOut no{};
    no.vals = vals.data();
    no.placed = reinterpret_cast<unsigned long long*>(kmers_out);
    std::vector<uint32_t> node_asm(M);
    no.placed_asm = node_asm.data();
    no.nodes = nodes_out;
    no.node_hash = nullptr;
    no.rec_asm = rec_asm;
    no.rec_base = rec_base;
    no.is_target = is_target;
    no.inv_t = n_targets ? 1.0 / (double)n_targets : 0.0;
    no.inv_n = n_non_targets ? 1.0 / (double)n_non_targets : 0.0;
    no.counts_only = 0;
    if (score) cuemu::launch(dim3((unsigned)nb), dim3(kNT), [&] { group_place_kernel<NodeOut, true>(pa, no); });
    else cuemu::launch(dim3((unsigned)nb), dim3(kNT), [&] { group_place_kernel<NodeOut, false>(pa, no); });
    *n_nodes_out = U;

    // ---- edges: grouped inside the node buckets ----
    constexpr int ESB = 11, EI = 2;
    using Smem = BucketEdgeSmem<ESB, EI>;
    std::vector<uint64_t> te_second(2 * M, 0xDDDDDDDDDDDDDDDDull);
    std::vector<uint32_t> te_w(2 * M, 0xDDDDDDDDu), bucket_e(nb, 0xABABABABu), bucket_rec(nb, 0xABABABABu);
    std::vector<uint16_t> te_r(2 * M, 0xDDDD);
    BucketEdgeArgs ea{};
    ea.item_rank = item_rank.data();
    ea.nb_prev = prev.data();
    ea.nb_next = next.data();
    ea.vals = vals.data();
    ea.start = start.data();
    ea.rec_asm = rec_asm;
    ea.rec_base = rec_base;
    ea.max_distinct = std::min<uint32_t>(max_distinct_edges, Smem::kEMax);
    ea.te_second = te_second.data();
    ea.te_w = te_w.data();
    ea.te_r = te_r.data();
    ea.bucket_e = bucket_e.data();
    ea.bucket_rec = bucket_rec.data();
    cuemu::launch(dim3((unsigned)nb), dim3(kNT), [&] { bucket_edges_kernel<ESB, EI>(ea); });
    std::vector<unsigned long long> e64(nb), side_rec(nb), etot(2, 0);
    cuemu::launch(dim3(2), dim3(256), [&] {
        bucket_edge_counts_kernel(bucket_e.data(), bucket_rec.data(), nb, e64.data(), side_rec.data(), etot.data());
    });
    std::vector<sw_edge> side_edges;
    std::vector<unsigned long long> ovf_e64(nb, 0), ovf_base;
    if (etot[0]) {
        // buckets left to the side path: the device sorts their records by (node, second hash) with two stable
        // radix sorts; here that order comes from a host sort, everything else is the device code
        *n_overflow_out = etot[0];
        std::vector<unsigned long long> side_off = exclusive_scan(side_rec);
        const uint64_t ns = etot[1];
        std::vector<uint32_t> side_node(ns, 0xDDDDDDDDu), side_asm(ns), perm(ns);
        std::vector<uint64_t> side_second(ns);
        cuemu::launch(dim3((unsigned)nb), dim3(kNT), [&] {
            side_emit_kernel(ea, bucket_e.data(), side_off.data(), grp_base.data(), side_node.data(), side_second.data(), side_asm.data());
        });
        std::iota(perm.begin(), perm.end(), 0u);
        std::stable_sort(perm.begin(), perm.end(), [&](uint32_t x, uint32_t y) {
            return side_node[x] != side_node[y] ? side_node[x] < side_node[y] : side_second[x] < side_second[y];
        });
        std::vector<unsigned long long> flags(ns + 1, 0);
        cuemu::launch(dim3(2), dim3(256), [&] {
            side_final_kernel(perm.data(), side_node.data(), side_second.data(), side_asm.data(), ns, flags.data(), nullptr, nullptr, nullptr);
        });
        std::vector<unsigned long long> scanned = exclusive_scan(std::vector<unsigned long long>(flags.begin(), flags.end() - 1));
        cuemu::launch(dim3(2), dim3(256), [&] {
            side_count_kernel(bucket_e.data(), side_off.data(), scanned.data(), nb, e64.data(), ovf_e64.data());
        });
        ovf_base = exclusive_scan(ovf_e64);
        side_edges.assign(scanned[ns] & 0xFFFFFFFFull, sw_edge{0, 0, 0});
        cuemu::launch(dim3(2), dim3(256), [&] {
            side_final_kernel(perm.data(), side_node.data(), side_second.data(), side_asm.data(), ns, nullptr, scanned.data(), nodes_out,
                              side_edges.data());
        });
    }
    std::vector<unsigned long long> ebase = exclusive_scan(e64);
    const uint64_t UE = ebase[nb];
    cuemu::launch(dim3(2), dim3(256), [&] {
        bucket_edges_out_kernel(te_second.data(), te_w.data(), te_r.data(), start.data(), bucket_e.data(), ebase.data(), nb,
                                grp_keys.data(), edges_out);
    });
    if (etot[0])
        cuemu::launch(dim3((unsigned)nb), dim3(kNT), [&] {
            overflow_copy_kernel(side_edges.data(), bucket_e.data(), ebase.data(), ovf_base.data(), edges_out);
        });
    *n_edges_out = UE;
    return 0;
}
