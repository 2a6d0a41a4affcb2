"""Pins the oracle (oracle/oracle.c) on the reference's golden vectors -- CPU only."""
from __future__ import annotations

import numpy as np
import pytest

from oracle import oracle as O
from tests.cases import GOLDEN_KW
from tests.helpers import assert_graph_equal, assert_matches_digest, check_graph_invariants

GOLDEN_ARRAYS = __import__("pathlib").Path(__file__).resolve().parent / "golden" / "arrays"


def test_kmer_hash_known_answers():
    # vectors derived from the reference build
    assert O.hash_kmer("ACGTACGTACGTACGTACGTA") == (0x263b9f5675c43346, 0xf658ab02446c5143)
    assert O.hash_kmer("AAAAAAAAAAAAAAAAAAAAA") == (0xb8f8c2d7c8478e21, 0x5d10d0a5c5f9362f)
    fwd = O.hash_kmer("GATTACAGATTACAGATTACA")
    rc = O.hash_kmer("TGTAATCTGTAATCTGTAATC")
    assert fwd == rc == (0x69fc4e3a6b04988a, 0x5b0eabcb4010109a)
    assert O.hash_kmer("acgtacgtacgtacgtacgtu") == O.hash_kmer("ACGTACGTACGTACGTACGTT")
    assert O.hash_kmer("ACGTNCGTACGTACGTACGTA") is None


def test_minimizer_known_answer_with_gap_and_self_loop():
    h1, pos = O.minimize("ACGTTGCATGNCATGCAACGTAGCTAGCTA", 5, 3)
    assert pos.tolist() == [1, 3, 11, 13, 15, 16, 19, 20, 21, 24, 25]
    assert h1[-1] == h1[-2] == 1993289553699856947


def test_reference_golden_graph(fixture_paths, expected_graph):
    """tests/smoke/test_outputs.py:22-58 of the reference: graph.npz at k=17, w=10."""
    kmers, nodes, edges, offsets, ids = O._build_native(fixture_paths, 17, 10)
    assert (len(kmers), len(nodes), len(edges)) == (1061, 380, 404)
    assert np.array_equal(kmers, expected_graph["kmers"])
    assert np.array_equal(edges, expected_graph["edges"])
    assert np.array_equal(offsets, expected_graph["record_offsets"])
    for f in ("hash", "start", "stop"):
        assert np.array_equal(nodes[f], expected_graph["nodes"][f])
    assert ids == [("NR_114042.1",), ("NR_024570.1",), ("NR_074910.1",), ("NR_119108.1",)]
    O._get_penalty_native(kmers, nodes, offsets, np.array([True, True, False, False]))
    exp = expected_graph["nodes"]
    if exp["n_tar"].any():  # the fixture stores post-penalty nodes in some reference versions
        assert np.array_equal(nodes["n_tar"], exp["n_tar"]) and np.array_equal(nodes["n_neg"], exp["n_neg"])
        np.testing.assert_allclose(nodes["penalty"], exp["penalty"], rtol=0, atol=1e-12)


@pytest.mark.parametrize("kw", GOLDEN_KW, ids=lambda kw: f"k{kw[0]}w{kw[1]}")
@pytest.mark.parametrize("case", ["fixtures", "edge"])
def test_oracle_matches_reference_arrays(case, kw, fixture_paths, edge_paths):
    paths, is_t = (fixture_paths, [True, True, False, False]) if case == "fixtures" else edge_paths
    k, w = kw
    want = np.load(GOLDEN_ARRAYS / f"{case}_{k}_{w}.npz", allow_pickle=False)
    got = O._build_native(paths, k, w)
    assert_graph_equal(got, (want["kmers"], want["nodes_build"], want["edges"], want["record_offsets"]), f"{case} {kw}")
    check_graph_invariants(*got[:4])
    nodes = got[1].copy()
    O._get_penalty_native(got[0], nodes, got[3], np.asarray(is_t))
    assert np.array_equal(nodes, want["nodes_penalty"])


@pytest.mark.parametrize("case", ["synth_small", "synth_medium", "synth_skew"])
def test_oracle_matches_reference_digests(case, synth_sets, digests):
    paths, is_t = synth_sets[case]
    for kw in [(21, 200), (17, 10), (31, 50), (4, 1)]:
        d = digests[case][f"{kw[0]},{kw[1]}"]
        got = O._build_native(paths, *kw)
        assert_matches_digest(got, d, f"{case} {kw}")
        nodes = got[1].copy()
        O._get_penalty_native(got[0], nodes, got[3], np.asarray(is_t))
        from tests.helpers import digest
        assert digest(nodes) == d["nodes_penalty"]


def test_oracle_against_live_reference(fixture_paths, edge_paths):
    """Where the reference tree is present (authoring container) compare live, incl. get_penalty/filter."""
    ref = O.load_reference()
    if ref is None:
        pytest.skip("oracle/_ref not built (no /root/reference)")
    for paths in (fixture_paths, edge_paths[0]):
        for kw in [(17, 10), (9, 9), (12, 33)]:
            a = O._build_native(paths, *kw)
            b = ref._build_native([str(p) for p in paths], kw[0], kw[1], 3, False)
            assert_graph_equal(a, b, str(kw))
            used = frozenset(a[1]["hash"][::3].tolist())
            fa = O._filter_kmers_native(a[0], a[1], used)
            fb = ref._filter_kmers_native(b[0], b[1], list(used))
            assert np.array_equal(fa[0], fb[0]) and np.array_equal(fa[1], fb[1])


def test_penalty_known_answers():
    """tests/smoke/test_graph.py:248-304 of the reference."""
    K, N = O.KMER_DTYPE, O.NODE_DTYPE
    kmers = np.array([(0, 0), (1, 0), (2, 1), (3, 2), (4, 4), (5, 2), (6, 3), (7, 5), (8, 6), (9, 4)], dtype=K)
    nodes = np.array([(10, 0, 5, 0, 0, 0.0), (20, 5, 7, 0, 0, 0.0), (30, 7, 9, 0, 0, 0.0), (40, 9, 10, 0, 0, 0.0),
                      (50, 10, 10, 9, 9, 9.0), (60, 5, 9, 0, 0, 0.0)], dtype=N)
    offsets = np.array([0, 2, 4, 5, 7], dtype=np.uint32)
    O._get_penalty_native(kmers, nodes, offsets, np.array([True, False, True, False]))
    assert nodes["n_tar"].tolist() == [2, 0, 0, 1, 0, 0]
    assert nodes["n_neg"].tolist() == [1, 1, 1, 0, 0, 2]
    np.testing.assert_allclose(nodes["penalty"], [0.5, np.hypot(1.0, 0.5), np.hypot(1.0, 0.5), 0.5, 1.0, np.sqrt(2.0)])
    kmers = np.array([(0, 0), (1, 1)], dtype=K)
    nodes = np.array([(10, 0, 2, 0, 0, 0.0)], dtype=N)
    O._get_penalty_native(kmers, nodes, np.array([0, 1, 1, 1, 2], dtype=np.uint32), np.array([True, False, True, False]))
    assert (nodes[0]["n_tar"], nodes[0]["n_neg"]) == (1, 1) and nodes[0]["penalty"] == np.sqrt(0.5)


@pytest.mark.parametrize("name", ["fixtures_17_10", "edge_17_10", "edge_21_200", "fixtures_31_50"])
def test_filter_edges_and_nodes_restatement_matches_reference(name):
    """oracle.filter_edges_and_nodes vs the reference's own _filter_edges_and_nodes
    (kmers.py:132-173; vectors written by tests/golden/make_filter_golden.py)."""
    arrays = GOLDEN_ARRAYS
    a = np.load(arrays / f"{name}.npz", allow_pickle=False)
    want = np.load(arrays / f"filter_{name}.npz", allow_pickle=False)
    for th in (0.0, 0.9, 1.0, 2.0, 3.5, 1e9):
        tag = str(th).replace(".", "p").replace("+", "")
        n2, e2 = O.filter_edges_and_nodes(a["nodes_penalty"], a["edges"], th)
        assert np.array_equal(n2, want[f"nodes_{tag}"]) and np.array_equal(e2, want[f"edges_{tag}"]), (name, th)
