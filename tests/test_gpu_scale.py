"""Larger inputs: oracle comparison at sizes the oracle finishes in seconds, and size-independent
properties (sortedness, tiling, idempotence, stream <-> graph consistency) beyond that."""
from __future__ import annotations

import numpy as np
import pytest

from oracle import oracle as O
from seqwin_b200.graph import KmerGraph, _get_penalty
from seqwin_b200.synth import SynthSpec, write_set
from tests.helpers import assert_graph_equal, check_graph_invariants

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def big_set(tmp_path_factory):
    spec = SynthSpec(n_genomes=16, n_targets=5, genome_len=1_000_000, n_contigs=9, seed=99, n_runs_every=4)
    return write_set(spec, tmp_path_factory.mktemp("big"))


@pytest.mark.parametrize("kw", [(21, 200), (21, 10), (31, 50)], ids=lambda kw: f"k{kw[0]}w{kw[1]}")
def test_16x1mbp_matches_oracle(big_set, kw):
    paths, is_t = big_set
    g = KmerGraph(paths, *kw, n_cpu=8)
    got = (g.kmers, g.nodes, g.edges, g.record_offsets, g.record_ids)
    want = O._build_native(paths, *kw)
    assert_graph_equal(got, want, str(kw))
    check_graph_invariants(*got[:4])
    nodes, ref_nodes = g.nodes.copy(), want[1].copy()
    _get_penalty(g.kmers, nodes, g.record_offsets, is_t)
    O._get_penalty_native(want[0], ref_nodes, want[3], is_t)
    assert np.array_equal(nodes, ref_nodes)


def test_properties_and_idempotence(big_set):
    paths, is_t = big_set
    a = KmerGraph(paths, 25, 100, n_cpu=8)
    b = KmerGraph(paths, 25, 100, n_cpu=3, low_memory=True)
    assert_graph_equal((a.kmers, a.nodes, a.edges, a.record_offsets), (b.kmers, b.nodes, b.edges, b.record_offsets))
    check_graph_invariants(a.kmers, a.nodes, a.edges, a.record_offsets)
    # every genome of a clade shares most minimizers: node counts must respect the class sizes
    nodes = a.nodes.copy()
    _get_penalty(a.kmers, nodes, a.record_offsets, is_t)
    n_t, n_n = int(np.sum(is_t)), int(np.sum(~np.asarray(is_t)))
    assert nodes["n_tar"].max() <= n_t and nodes["n_neg"].max() <= n_n
    assert np.all((nodes["n_tar"] + nodes["n_neg"]) >= 1)
    assert np.all((nodes["penalty"] >= 0) & (nodes["penalty"] <= np.sqrt(2.0)))
    # a reversed assembly order permutes record indices but not the node / edge hash sets
    c = KmerGraph(paths[::-1], 25, 100)
    assert np.array_equal(c.nodes["hash"], a.nodes["hash"])
    assert np.array_equal(c.edges, a.edges)
    assert np.array_equal(c.nodes["stop"] - c.nodes["start"], a.nodes["stop"] - a.nodes["start"])


def test_chunked_batch_equals_whole_batch():
    """bench.build_batch packs the synthetic genomes chunk by chunk and concatenates the packed pieces
    (sw_batch_concat); the graph must be the one of the batch packed in one go."""
    import ctypes as C

    import bench
    from seqwin_b200 import _lib
    from seqwin_b200.dist import export_graph
    from seqwin_b200.synth import SynthSet
    from tests.helpers import digest
    L = _lib.lib()
    spec = SynthSpec(n_genomes=11, n_targets=3, genome_len=200_000, n_contigs=6, seed=17, n_runs_every=4)
    ss = SynthSet(spec)
    out = []
    for chunk in (3, 100):
        b = bench.build_batch(ss, range(spec.n_genomes), 4, chunk=chunk)
        g = C.c_void_p()
        _lib.check(L.sw_build_from_batch(b, 21, 50, C.byref(g), None))
        out.append([digest(a) for a in export_graph(L, g)])
        L.sw_graph_free(g)
        L.sw_batch_free(b)
    whole, first = bench.build_batch(ss, range(spec.n_genomes), 4, chunk=4, keep_first=True)
    assert L.sw_batch_n_records(first) == 4 * 6 and L.sw_batch_n_records(whole) == 11 * 6
    L.sw_batch_free(whole)
    L.sw_batch_free(first)
    assert out[0] == out[1]
