"""Parity of the CUDA path (through the C ABI / the _core shim) with the reference's golden vectors
and with the oracle.  Bit-exact: every array is integer / byte data except nodes.penalty, which is
an IEEE double computed without FMA contraction and is compared for exact equality as well."""
from __future__ import annotations

import ctypes as C
from pathlib import Path

import numpy as np
import pytest

from oracle import oracle as O
from seqwin_b200 import _lib
from seqwin_b200._core import EDGE_DTYPE, KMER_DTYPE, NODE_DTYPE
from seqwin_b200.graph import KmerGraph, _filter_kmers, _get_penalty
from tests.cases import GOLDEN_KW
from tests.helpers import assert_graph_equal, assert_matches_digest, check_graph_invariants, digest

pytestmark = pytest.mark.gpu
GOLDEN_ARRAYS = Path(__file__).resolve().parent / "golden" / "arrays"


def _build(*args, **kwargs):
    g = KmerGraph(*args, **kwargs)
    return g.kmers, g.nodes, g.edges, g.record_offsets, g.record_ids


def test_reference_golden_graph(fixture_paths, expected_graph):
    """Reference tests/smoke/test_outputs.py:22-58 -- graph.npz at k=17, w=10, array-equal."""
    kmers, nodes, edges, offsets, ids = _build(fixture_paths, 17, 10, n_cpu=1)
    assert kmers.dtype == KMER_DTYPE and nodes.dtype == NODE_DTYPE and edges.dtype == EDGE_DTYPE
    assert offsets.dtype == np.dtype(np.uint32)
    np.testing.assert_array_equal(kmers, expected_graph["kmers"])
    np.testing.assert_array_equal(edges, expected_graph["edges"])
    np.testing.assert_array_equal(offsets, expected_graph["record_offsets"])
    for f in ("hash", "start", "stop"):
        np.testing.assert_array_equal(nodes[f], expected_graph["nodes"][f])
    assert np.all(nodes["n_tar"] == 0) and np.all(nodes["n_neg"] == 0) and np.all(nodes["penalty"] == 0.0)
    assert ids == [("NR_114042.1",), ("NR_024570.1",), ("NR_074910.1",), ("NR_119108.1",)]
    assert edges[0].tolist() == (30758054237625032, 2986744083141097459, 4)


@pytest.mark.parametrize("kw", GOLDEN_KW, ids=lambda kw: f"k{kw[0]}w{kw[1]}")
@pytest.mark.parametrize("case", ["fixtures", "edge"])
def test_build_and_penalty_match_reference_arrays(case, kw, fixture_paths, edge_paths):
    paths, is_t = (fixture_paths, [True, True, False, False]) if case == "fixtures" else edge_paths
    k, w = kw
    want = np.load(GOLDEN_ARRAYS / f"{case}_{k}_{w}.npz", allow_pickle=False)
    got = _build(paths, k, w, n_cpu=2)
    assert_graph_equal(got, (want["kmers"], want["nodes_build"], want["edges"], want["record_offsets"]), f"{case} {kw}")
    check_graph_invariants(*got[:4])
    nodes = got[1].copy()
    assert _get_penalty(got[0], nodes, got[3], is_t, n_cpu=3) is None
    assert np.array_equal(nodes, want["nodes_penalty"]), "n_tar / n_neg / penalty differ from the reference"


@pytest.mark.parametrize("case", ["synth_small", "synth_medium", "synth_skew"])
def test_build_matches_reference_digests(case, synth_sets, digests):
    paths, is_t = synth_sets[case]
    for kw in GOLDEN_KW:
        d = digests[case][f"{kw[0]},{kw[1]}"]
        got = _build(paths, *kw, n_cpu=4)
        assert_matches_digest(got, d, f"{case} {kw}")
        nodes = got[1].copy()
        _get_penalty(got[0], nodes, got[3], is_t)
        assert digest(nodes) == d["nodes_penalty"], f"{case} {kw} penalty"


def test_threading_and_low_memory_equivalence(fixture_paths):
    """Reference tests/smoke/test_graph.py:67-127, 222-245."""
    base = _build(fixture_paths, kmerlen=7, windowsize=10, n_cpu=1)
    assert np.array_equal(base[3], np.array([0, 1, 2, 3, 4], dtype=np.uint32))
    assert np.array_equal(np.unique(base[0]["record_idx"]), np.arange(4, dtype=np.uint32))
    for n_cpu in (2, 99):
        for low in (False, True):
            other = _build(fixture_paths, kmerlen=7, windowsize=10, n_cpu=n_cpu, low_memory=low)
            assert_graph_equal(other, base)


def test_record_offsets_and_global_record_indices(tmp_path):
    """Reference tests/smoke/test_graph.py:144-165."""
    paths = []
    for i, n in enumerate([2, 1, 3, 1]):
        p = tmp_path / f"a{i}.fasta"
        p.write_text("".join(f">r{j}\n{'ACGT' * 20}\n" for j in range(n)))
        paths.append(p)
    kmers, _, _, offsets, ids = _build(paths, kmerlen=7, windowsize=10, n_cpu=2)
    assert [len(x) for x in ids] == [2, 1, 3, 1]
    assert np.array_equal(offsets, np.array([0, 2, 3, 6, 7], dtype=np.uint32))
    assert np.array_equal(np.unique(kmers["record_idx"]), np.arange(7, dtype=np.uint32))
    assert_graph_equal(_build(paths, 7, 10), O._build_native(paths, 7, 10))


def test_empty_inputs(tmp_path):
    """Reference tests/smoke/test_graph.py:168-187."""
    empty = tmp_path / "empty.fasta"
    empty.write_text("")
    for paths, want in (([], [0]), ([empty], [0, 0])):
        for low in (False, True):
            kmers, nodes, edges, offsets, ids = _build(paths, kmerlen=7, windowsize=10, n_cpu=2, low_memory=low)
            assert len(kmers) == 0 and len(nodes) == 0 and len(edges) == 0
            assert offsets.dtype == np.dtype(np.uint32) and offsets.tolist() == want
            assert ids == [()] * len(paths)


def test_missing_file_raises_runtime_error(tmp_path):
    with pytest.raises(RuntimeError):
        _build([tmp_path / "nope.fasta"], 7, 10)


def _penalty_inputs():
    kmers = np.array([(0, 0), (1, 0), (2, 1), (3, 2), (4, 4), (5, 2), (6, 3), (7, 5), (8, 6), (9, 4)], dtype=KMER_DTYPE)
    nodes = np.array([(10, 0, 5, 0, 0, 0.0), (20, 5, 7, 0, 0, 0.0), (30, 7, 9, 0, 0, 0.0), (40, 9, 10, 0, 0, 0.0),
                      (50, 10, 10, 9, 9, 9.0), (60, 5, 9, 0, 0, 0.0)], dtype=NODE_DTYPE)
    return kmers, nodes, np.array([0, 2, 4, 5, 7], dtype=np.uint32), np.array([True, False, True, False])


def test_get_penalty_exact_scoring():
    """Reference tests/smoke/test_graph.py:268-304."""
    kmers, nodes, offsets, is_t = _penalty_inputs()
    assert _get_penalty(kmers, nodes, offsets, is_t, n_cpu=1) is None
    assert nodes["n_tar"].tolist() == [2, 0, 0, 1, 0, 0]
    assert nodes["n_neg"].tolist() == [1, 1, 1, 0, 0, 2]
    np.testing.assert_allclose(nodes["penalty"], [0.5, np.hypot(1.0, 0.5), np.hypot(1.0, 0.5), 0.5, 1.0, np.sqrt(2.0)])
    ref_nodes = _penalty_inputs()[1]
    O._get_penalty_native(kmers, ref_nodes, offsets, is_t)
    assert np.array_equal(nodes, ref_nodes)
    kmers = np.array([(0, 0), (1, 1)], dtype=KMER_DTYPE)
    nodes = np.array([(10, 0, 2, 0, 0, 0.0)], dtype=NODE_DTYPE)
    _get_penalty(kmers, nodes, np.array([0, 1, 1, 1, 2], dtype=np.uint32), [True, False, True, False], n_cpu=2)
    assert (nodes[0]["n_tar"], nodes[0]["n_neg"]) == (1, 1) and nodes[0]["penalty"] == np.sqrt(0.5)


def test_get_penalty_device_side_validation():
    """Reference tests/smoke/test_graph.py:328-339 (the checks that need the data)."""
    kmers, nodes, offsets, is_t = _penalty_inputs()
    bad_nodes = nodes.copy()
    bad_nodes[0]["stop"] = len(kmers) + 1
    with pytest.raises(ValueError):
        _get_penalty(kmers, bad_nodes, offsets, is_t)
    bad_kmers = kmers.copy()
    bad_kmers[0]["record_idx"] = 7
    with pytest.raises(ValueError):
        _get_penalty(bad_kmers, nodes.copy(), offsets, is_t)
    desc = kmers.copy()
    desc[3]["record_idx"] = 0
    with pytest.raises(ValueError):
        _get_penalty(desc, nodes.copy(), offsets, is_t)


def test_scored_build_entry_points(edge_paths):
    """Build + get_penalty in one call (device-resident scoring) == reference build then get_penalty."""
    L = _lib.lib()
    paths, is_t = edge_paths
    is_t = np.ascontiguousarray(is_t, dtype=np.bool_)
    arr = (C.c_char_p * len(paths))(*[str(p).encode() for p in paths])
    for k, w in [(17, 10), (21, 200), (31, 50)]:
        want = np.load(GOLDEN_ARRAYS / f"edge_{k}_{w}.npz", allow_pickle=False)
        b, d, g = C.c_void_p(), C.c_void_p(), C.c_void_p()
        _lib.check(L.sw_batch_from_fasta(arr, len(paths), 2, C.byref(b)))
        try:
            # end-to-end entry: pinned host batch -> scored graph in host memory
            _lib.check(L.sw_build_from_batch_scored(b, k, w, is_t.ctypes.data, len(is_t), C.byref(g), None))
            from seqwin_b200.dist import export_graph
            kmers, nodes, edges = export_graph(L, g)
            L.sw_graph_free(g)
            assert np.array_equal(kmers, want["kmers"]) and np.array_equal(edges, want["edges"])
            assert np.array_equal(nodes, want["nodes_penalty"])
            # device-resident entry: build, then score in place on the device
            _lib.check(L.sw_dev_upload(b, C.byref(d)))
            g = C.c_void_p()
            _lib.check(L.sw_dev_build(d, k, w, C.byref(g), None))
            _lib.check(L.sw_graph_penalty(g, None, 0, is_t.ctypes.data, len(is_t), None))
            kmers, nodes, edges = export_graph(L, g)
            L.sw_graph_free(g)
            assert np.array_equal(nodes, want["nodes_penalty"]) and np.array_equal(kmers, want["kmers"])
            # device-resident entry with the classes known up front: scoring fused into the node stage
            g = C.c_void_p()
            _lib.check(L.sw_dev_build_scored(d, k, w, is_t.ctypes.data, len(is_t), C.byref(g), None))
            kmers, nodes, edges = export_graph(L, g)
            L.sw_graph_free(g)
            L.sw_dev_batch_free(d)
            assert np.array_equal(nodes, want["nodes_penalty"]) and np.array_equal(kmers, want["kmers"])
            assert np.array_equal(edges, want["edges"])
        finally:
            L.sw_batch_free(b)
    with pytest.raises(ValueError):
        b = C.c_void_p()
        _lib.check(L.sw_batch_from_fasta(arr, len(paths), 2, C.byref(b)))
        try:
            all_t = np.ones(len(paths), dtype=np.bool_)
            _lib.check(L.sw_build_from_batch_scored(b, 17, 10, all_t.ctypes.data, len(all_t), C.byref(g), None))
        finally:
            L.sw_batch_free(b)
    with pytest.raises(ValueError):   # one flag per assembly of the batch
        b = C.c_void_p()
        _lib.check(L.sw_batch_from_fasta(arr, len(paths), 2, C.byref(b)))
        try:
            _lib.check(L.sw_build_from_batch_scored(b, 17, 10, is_t.ctypes.data, len(is_t) - 1, C.byref(g), None))
        finally:
            L.sw_batch_free(b)


def test_filter_kmers_on_built_graph(fixture_paths):
    kmers, nodes, _, offsets, _ = _build(fixture_paths, 17, 10)
    used = frozenset(nodes["hash"][::3])
    kn, nn = _filter_kmers(kmers, nodes, used)
    ko, no = O._filter_kmers_native(kmers, nodes, used)
    assert np.array_equal(kn, ko) and np.array_equal(nn, no)


def _dev_sketch(seqs, k, w):
    """Sketch stage alone through the C ABI: (h1, pos, record) stream of in-memory records."""
    L = _lib.lib()
    n = len(seqs)
    arrs = [np.frombuffer(s, dtype=np.uint8) for s in seqs]
    ptrs = (C.c_void_p * max(1, n))(*[a.ctypes.data for a in arrs])
    lens = np.array([len(s) for s in seqs], dtype=np.uint32)
    asm = np.zeros(n, dtype=np.uint32)
    b, d = C.c_void_p(), C.c_void_p()
    _lib.check(L.sw_batch_from_memory(ptrs, lens.ctypes.data, asm.ctypes.data, None, n, 1, 2, C.byref(b)))
    try:
        _lib.check(L.sw_dev_upload(b, C.byref(d)))
        cnt = C.c_size_t()
        _lib.check(L.sw_dev_sketch(d, k, w, None, None, None, 0, C.byref(cnt)))
        h1, pos, rec = np.empty(cnt.value, np.uint64), np.empty(cnt.value, np.uint32), np.empty(cnt.value, np.uint32)
        _lib.check(L.sw_dev_sketch(d, k, w, h1.ctypes.data, pos.ctypes.data, rec.ctypes.data, cnt.value, C.byref(cnt)))
        return h1, pos, rec
    finally:
        if d:
            L.sw_dev_batch_free(d)
        L.sw_batch_free(b)


@pytest.mark.parametrize("kw", [(21, 200), (21, 10), (5, 3), (4, 1), (31, 50), (21, 46), (9, 9), (15, 64), (33, 1000),
                                (21, 3000), (255, 40), (21, 144), (15, 150), (31, 175), (21, 2048), (17, 143), (64, 500),
                                (3, 200)], ids=lambda kw: f"k{kw[0]}w{kw[1]}")
def test_sketch_stream_matches_oracle(kw):
    """Ordered minimizer stream of the sketch kernel vs btllib semantics (minimizer.cpp:53-90)."""
    k, w = kw
    rng = np.random.default_rng(k * 100003 + w)

    def rand(n, p_n=0.0, alphabet=b"ACGT"):
        a = np.frombuffer(alphabet, dtype=np.uint8)[rng.integers(0, len(alphabet), n)].copy()
        if p_n:
            for s in rng.integers(0, max(1, n), max(1, int(n * p_n / 5))):
                a[s:s + rng.integers(1, 12)] = ord("N")
        return a.tobytes()
    seqs = [rand(int(rng.integers(1, 9000))) for _ in range(4)] + [
        rand(150_000), rand(90_000, 0.001), rand(30_000, 0.02), rand(20_000, alphabet=b"AC"), rand(9000, alphabet=b"A"),
        b"", b"ACGT", b"ACGTTGCA" * 3000, rand(70_000).lower()]
    h1, pos, rec = _dev_sketch(seqs, k, w)
    H, P, R = [], [], []
    for r, s in enumerate(seqs):
        h, p = O.minimize(s, k, w)
        H.append(h), P.append(p), R.append(np.full(len(h), r, np.uint32))
    oh, op, orr = np.concatenate(H), np.concatenate(P), np.concatenate(R)
    assert len(h1) == len(oh)
    assert np.array_equal(rec, orr) and np.array_equal(pos, op) and np.array_equal(h1, oh)


def test_low_complexity_overflows_the_density_estimate():
    """Homopolymer / short-period input emits a minimizer for (almost) every window, far above the
    2/(w+1) density the output buffer is sized for: the sketch must re-run with the exact size."""
    rng = np.random.default_rng(5)
    rand = np.frombuffer(b"ACGT", dtype=np.uint8)[rng.integers(0, 4, 50_000)].tobytes()
    seqs = [b"A" * 400_000 + rand + b"AC" * 150_000, rand + b"T" * 250_000]
    for k, w in [(21, 200), (15, 33)]:
        h1, pos, rec = _dev_sketch(seqs, k, w)
        H, P, R = [], [], []
        for r, s in enumerate(seqs):
            h, p = O.minimize(s, k, w)
            H.append(h), P.append(p), R.append(np.full(len(h), r, np.uint32))
        assert len(h1) == sum(len(x) for x in H) > 700_000
        assert np.array_equal(h1, np.concatenate(H)) and np.array_equal(pos, np.concatenate(P))
        assert np.array_equal(rec, np.concatenate(R))


def test_window_too_large_fails_loudly():
    with pytest.raises(RuntimeError, match="windowsize"):
        _dev_sketch([b"ACGT" * 100], 21, 40_000)


AGG_MODES = {
    "legacy": {"SEQWIN_AGG": "legacy"},                               # sort + run-length path (round 1)
    "edge-overflow": {"SEQWIN_AGG_EDGE_DISTINCT": "6"},               # most edge buckets take the side (sort) path
    "tiny-buckets": {"SEQWIN_AGG_NODE_BUCKET": "12"},                 # 2-3 partition passes, tiny groups
    "one-bucket-pair": {"SEQWIN_AGG_NODE_BUCKET": "700", "SEQWIN_AGG_EDGE_DISTINCT": "700"},
    "large-edge-table": {"SEQWIN_AGG_EDGE_GEOM": "2"},
    # the reduced-footprint plan: hash slices of the stream aggregated one after the other
    "slices-4": {"SEQWIN_AGG_SLICES": "4"},
    "slices-16-tiny-buckets": {"SEQWIN_AGG_SLICES": "16", "SEQWIN_AGG_NODE_BUCKET": "12"},
    "slices-edge-overflow": {"SEQWIN_AGG_SLICES": "4", "SEQWIN_AGG_EDGE_DISTINCT": "6"},
    # ... with output arrays sized too small at first (the estimate is retried with more head room)
    "slices-capacity-retry": {"SEQWIN_AGG_SLICES": "2", "SEQWIN_AGG_SLACK_PCT": "40", "SEQWIN_AGG_SLACK_ITEMS": "1"},
}


@pytest.mark.parametrize("mode", sorted(AGG_MODES))
def test_aggregation_paths_agree(mode, synth_sets, digests, fixture_paths, expected_graph, edge_paths, monkeypatch):
    """The bucketed aggregation (csrc/agg.cuh) under settings that force its less common branches -- edge
    buckets handed to the sort path, several partition passes with a few items per bucket, a node bucket that
    overflows and sends the whole build to the sort-based path -- and the sort-based path itself: the graphs
    (build and scored build) must not change."""
    for k_, v_ in AGG_MODES[mode].items():
        monkeypatch.setenv(k_, v_)
    kmers, nodes, edges, offsets, _ = _build(fixture_paths, 17, 10, n_cpu=1)
    np.testing.assert_array_equal(kmers, expected_graph["kmers"])
    np.testing.assert_array_equal(edges, expected_graph["edges"])
    for f in ("hash", "start", "stop"):
        np.testing.assert_array_equal(nodes[f], expected_graph["nodes"][f])
    for case in ("synth_small", "synth_medium", "synth_skew"):
        paths, is_t = synth_sets[case]
        for kw in GOLDEN_KW[:4]:
            d = digests[case][f"{kw[0]},{kw[1]}"]
            got = _build(paths, *kw, n_cpu=4)
            assert_matches_digest(got, d, f"{case} {kw} {mode}")
            nodes2 = got[1].copy()
            _get_penalty(got[0], nodes2, got[3], is_t)
            assert digest(nodes2) == d["nodes_penalty"], f"{case} {kw} {mode} penalty"
    # scored build entry point (scoring fused into the placement kernel)
    L = _lib.lib()
    paths, is_t = edge_paths
    arr = (C.c_char_p * len(paths))(*[str(p).encode() for p in paths])
    b, g = C.c_void_p(), C.c_void_p()
    _lib.check(L.sw_batch_from_fasta(arr, len(paths), 2, C.byref(b)))
    t8 = np.ascontiguousarray(is_t, dtype=np.bool_)
    _lib.check(L.sw_build_from_batch_scored(b, 21, 46, t8.ctypes.data, len(t8), C.byref(g), None))
    from seqwin_b200.dist import export_graph
    gk, gn, ge = export_graph(L, g)
    L.sw_graph_free(g)
    L.sw_batch_free(b)
    want = np.load(GOLDEN_ARRAYS / "edge_21_46.npz", allow_pickle=False)
    assert np.array_equal(gk, want["kmers"]) and np.array_equal(ge, want["edges"])
    assert np.array_equal(gn, want["nodes_penalty"])


@pytest.mark.parametrize("begin_bit", [0, 40, 48, 56])
def test_node_sort_tie_fixup(begin_bit, synth_sets, digests, fixture_paths, expected_graph, monkeypatch):
    monkeypatch.setenv("SEQWIN_AGG", "legacy")
    """The node sort looks at the high word of h1 only (4 radix passes) and re-sorts the rare groups of
    equal high words afterwards.  Starting the passes at bit 40 / 48 / 56 makes such groups common
    (and, on the larger set, too big for the fix-up, which then falls back to the 64-bit sort);
    bit 0 is the plain 64-bit sort.  The graphs must not change."""
    monkeypatch.setenv("SEQWIN_SORT_BEGIN_BIT", str(begin_bit))
    kmers, nodes, edges, offsets, _ = _build(fixture_paths, 17, 10, n_cpu=1)
    np.testing.assert_array_equal(kmers, expected_graph["kmers"])
    np.testing.assert_array_equal(edges, expected_graph["edges"])
    for f in ("hash", "start", "stop"):
        np.testing.assert_array_equal(nodes[f], expected_graph["nodes"][f])
    for case in ("synth_small", "synth_medium", "synth_skew"):
        paths, is_t = synth_sets[case]
        for kw in GOLDEN_KW[:3]:
            d = digests[case][f"{kw[0]},{kw[1]}"]
            got = _build(paths, *kw, n_cpu=4)
            assert_matches_digest(got, d, f"{case} {kw} begin_bit={begin_bit}")
