"""First consumers of the graph on the GPU: the edge-weight / isolated-node
filter against the reference's own function (golden vectors) and against the oracle restatement, and
filter_kmers on a device-resident graph against the oracle's filter_kmers."""
from __future__ import annotations

import ctypes as C
from pathlib import Path

import numpy as np
import pytest

from oracle import oracle as O
from seqwin_b200 import _lib
from seqwin_b200.dist import export_graph
from seqwin_b200.graph import KmerGraph, _filter_edges_and_nodes, _get_penalty

pytestmark = pytest.mark.gpu
GOLDEN_ARRAYS = Path(__file__).resolve().parent / "golden" / "arrays"
THRESHOLDS = (0.0, 0.9, 1.0, 2.0, 3.5, 1e9)


@pytest.mark.parametrize("name", ["fixtures_17_10", "edge_17_10", "edge_21_200", "fixtures_31_50"])
def test_filter_edges_and_nodes_golden(name):
    """kmers.py:132-162 -- vectors produced by the reference's _filter_edges_and_nodes."""
    a = np.load(GOLDEN_ARRAYS / f"{name}.npz", allow_pickle=False)
    want = np.load(GOLDEN_ARRAYS / f"filter_{name}.npz", allow_pickle=False)
    for th in THRESHOLDS:
        tag = str(th).replace(".", "p").replace("+", "")
        nodes, edges = _filter_edges_and_nodes(a["nodes_penalty"], a["edges"], th)
        assert np.array_equal(nodes, want[f"nodes_{tag}"]), (name, th)
        assert np.array_equal(edges, want[f"edges_{tag}"]), (name, th)


def test_filter_edges_and_nodes_rejects_foreign_endpoints():
    a = np.load(GOLDEN_ARRAYS / "fixtures_17_10.npz", allow_pickle=False)
    edges = a["edges"].copy()
    edges["second"][3] = np.uint64(12345)   # not a node hash
    with pytest.raises(ValueError):
        _filter_edges_and_nodes(a["nodes_penalty"], edges, 0.0)
    with pytest.raises(TypeError):
        _filter_edges_and_nodes(a["nodes_penalty"].view(np.uint8), edges, 0.0)
    n, e = _filter_edges_and_nodes(a["nodes_penalty"][:0], edges[:0], 0.0)
    assert len(n) == 0 and len(e) == 0


@pytest.mark.parametrize("case", ["synth_medium", "synth_skew"])
def test_device_resident_filters_match_oracle(case, synth_sets):
    """build -> score -> filter edges / nodes -> filter k-mers without leaving the device."""
    L = _lib.lib()
    paths, is_t = synth_sets[case]
    is_t = np.ascontiguousarray(is_t, dtype=np.bool_)
    arr = (C.c_char_p * len(paths))(*[str(p).encode() for p in paths])
    for k, w in ((21, 200), (15, 20)):
        ref = KmerGraph(paths, k, w, n_cpu=2)
        nodes_ref = ref.nodes.copy()
        _get_penalty(ref.kmers, nodes_ref, ref.record_offsets, is_t)
        b, d, g = C.c_void_p(), C.c_void_p(), C.c_void_p()
        _lib.check(L.sw_batch_from_fasta(arr, len(paths), 2, C.byref(b)))
        try:
            _lib.check(L.sw_dev_upload(b, C.byref(d)))
            _lib.check(L.sw_dev_build_scored(d, k, w, is_t.ctypes.data, len(is_t), C.byref(g), None))
            for th in (1.0, 2.0):
                _lib.check(L.sw_graph_filter_edges(g, C.c_uint64(int(th))))
                nodes_want, edges_want = O.filter_edges_and_nodes(nodes_ref, ref.edges, th)
                kmers, nodes, edges = export_graph(L, g)
                assert np.array_equal(kmers, ref.kmers)
                assert np.array_equal(nodes, nodes_want) and np.array_equal(edges, edges_want), (case, k, w, th)
                nodes_ref, ref.edges = nodes_want, edges_want   # filters compose
            # filter_kmers: every third surviving node plus hashes that are not in the graph, shuffled
            rng = np.random.default_rng(k * 31 + w)
            used = np.concatenate([nodes_ref["hash"][::3], rng.integers(0, 2**63, 50, dtype=np.uint64),
                                   nodes_ref["hash"][:5]])
            rng.shuffle(used)
            used = np.ascontiguousarray(used)
            _lib.check(L.sw_graph_filter_kmers(g, used.ctypes.data, len(used)))
            kmers_want, nodes_want = O._filter_kmers_native(ref.kmers, nodes_ref, frozenset(used.tolist()))
            kmers, nodes, edges = export_graph(L, g)
            assert np.array_equal(kmers, kmers_want) and np.array_equal(nodes, nodes_want), (case, k, w)
            assert np.array_equal(edges, ref.edges)
            # nothing left
            _lib.check(L.sw_graph_filter_kmers(g, None, 0))
            assert L.sw_graph_size(g, _lib.SW_KMERS) == 0 and L.sw_graph_size(g, _lib.SW_NODES) == 0
        finally:
            if g:
                L.sw_graph_free(g)
            if d:
                L.sw_dev_batch_free(d)
            L.sw_batch_free(b)


def test_filter_on_full_size_graph_properties():
    """A 100-genome graph: the filter must keep exactly the edges above the threshold, exactly their
    endpoints, in order, and be idempotent."""
    from seqwin_b200.synth import SynthSet, SynthSpec
    import bench
    L = _lib.lib()
    spec = SynthSpec(n_genomes=100, n_targets=20, genome_len=1_000_000, n_contigs=20, seed=7)
    batch = bench.build_batch(SynthSet(spec), range(spec.n_genomes), 4)
    is_t = np.ascontiguousarray(np.arange(spec.n_genomes) < spec.n_targets, dtype=np.bool_)
    d, g = C.c_void_p(), C.c_void_p()
    try:
        _lib.check(L.sw_dev_upload(batch, C.byref(d)))
        _lib.check(L.sw_dev_build_scored(d, 21, 200, is_t.ctypes.data, len(is_t), C.byref(g), None))
        kmers0, nodes0, edges0 = export_graph(L, g)
        th = 15
        _lib.check(L.sw_graph_filter_edges(g, C.c_uint64(th)))
        _, nodes1, edges1 = export_graph(L, g)
        nodes_want, edges_want = O.filter_edges_and_nodes(nodes0, edges0, float(th))
        assert np.array_equal(edges1, edges_want) and np.array_equal(nodes1, nodes_want)
        assert 0 < len(edges1) < len(edges0) and 0 < len(nodes1) < len(nodes0)
        _lib.check(L.sw_graph_filter_edges(g, C.c_uint64(th)))
        _, nodes2, edges2 = export_graph(L, g)
        assert np.array_equal(nodes2, nodes1) and np.array_equal(edges2, edges1)
    finally:
        if g:
            L.sw_graph_free(g)
        if d:
            L.sw_dev_batch_free(d)
        L.sw_batch_free(batch)


def test_penalty_threshold_statistics(synth_sets, tmp_path):
    """the threshold statistics of kmers.py:424-440 from the device-resident nodes
    (exact integer sums) vs the reference's float formula on the host arrays (rtol 1e-12: the
    reference sums float products, the sums here are exact), and the graph.npz writer round trip."""
    from seqwin_b200.graph import count_sums, penalty_threshold, save_graph
    L = _lib.lib()
    paths, is_t = synth_sets["synth_medium"]
    is_t = np.ascontiguousarray(is_t, dtype=np.bool_)
    n_tar, n_neg = int(is_t.sum()), int(len(is_t) - is_t.sum())
    arr = (C.c_char_p * len(paths))(*[str(p).encode() for p in paths])
    b, d, g = C.c_void_p(), C.c_void_p(), C.c_void_p()
    _lib.check(L.sw_batch_from_fasta(arr, len(paths), 2, C.byref(b)))
    try:
        _lib.check(L.sw_dev_upload(b, C.byref(d)))
        _lib.check(L.sw_dev_build_scored(d, 21, 50, is_t.ctypes.data, len(is_t), C.byref(g), None))
        sums = np.zeros(3, dtype=np.uint64)
        _lib.check(L.sw_graph_count_sums(g, sums.ctypes.data))
        kmers, nodes, edges = export_graph(L, g)
        assert tuple(int(x) for x in sums) == count_sums(nodes)
        # the reference's arithmetic (kmers.py:424-436)
        frac_tar = nodes["n_tar"] / n_tar
        e_abs = 1 - np.sum(frac_tar * nodes["n_tar"]) / np.sum(nodes["n_tar"])
        frac_neg = nodes["n_neg"] / n_neg
        e_pre = np.sum(frac_neg * nodes["n_tar"]) / np.sum(nodes["n_tar"])
        th_ref = min((1 - 5 / 10) * (e_abs * e_pre) ** 0.5, 0.2)
        th, a, p = penalty_threshold(sums, n_tar, n_neg)
        assert abs(a - e_abs) <= 1e-12 and abs(p - e_pre) <= 1e-12 and abs(th - th_ref) <= 1e-12
        offs = np.zeros(len(paths) + 1, dtype=np.uint32)
        save_graph(tmp_path / "graph.npz", kmers, nodes, edges, offs)
        back = np.load(tmp_path / "graph.npz", allow_pickle=False)
        assert np.array_equal(back["nodes"], nodes) and np.array_equal(back["kmers"], kmers)
        assert np.array_equal(back["edges"], edges) and back["nodes"].dtype == nodes.dtype
    finally:
        if g:
            L.sw_graph_free(g)
        if d:
            L.sw_dev_batch_free(d)
        L.sw_batch_free(b)
