"""The aggregation kernels (seqwin_b200/csrc/agg.cuh: bucket bounds, group count, group place, the per-bucket
edge grouping, side-path plumbing; csrc/nbr.cuh: owned neighbours) executed on the CPU through tests/emul/cuda_emul.h and compared with the oracle's
graph: the minimizer stream comes from the oracle, the stable partition is a host sort, everything else is
the device code.  CPU only."""
from __future__ import annotations

import ctypes as C
import subprocess
from pathlib import Path

import numpy as np
import pytest

from oracle import oracle as O
from seqwin_b200.synth import SynthSet, SynthSpec, write_set

EMUL_DIR = Path(__file__).resolve().parent / "emul"
CSRC = EMUL_DIR.parents[1] / "seqwin_b200" / "csrc"


@pytest.fixture(scope="module")
def agg():
    so = EMUL_DIR / "libagg_emul.so"
    srcs = [EMUL_DIR / "agg_emul.cpp", EMUL_DIR / "cuda_emul.h", CSRC / "agg.cuh", CSRC / "nbr.cuh", CSRC / "sample.cuh", CSRC / "common.h"]
    if not so.exists() or so.stat().st_mtime < max(s.stat().st_mtime for s in srcs):
        subprocess.check_call(["g++", "-O1", "-g", "-std=c++17", "-Wall", "-Wno-unknown-pragmas", "-Wno-unused-function", "-fPIC",
                               "-shared", "-o", str(so), str(EMUL_DIR / "agg_emul.cpp")])
    L = C.CDLL(str(so))
    L.agg_emul_run.restype = C.c_long
    L.agg_emul_estimate.restype = C.c_double
    return L


def _stream(spec: SynthSpec, k: int, w: int):
    ss = SynthSet(spec)
    keys, vals, rec_asm = [], [], []
    rec = 0
    for g in range(spec.n_genomes):
        for _, seq in ss.records(g):
            h1, pos = O.minimize(seq.tobytes(), k, w)
            keys.append(h1)
            vals.append(pos.astype(np.uint64) | (np.uint64(rec) << np.uint64(32)))
            rec_asm.append(g)
            rec += 1
    return (np.ascontiguousarray(np.concatenate(keys)), np.ascontiguousarray(np.concatenate(vals)),
            np.asarray(rec_asm, dtype=np.uint32), ss.is_targets)


def _run(L, keys, vals, rec_asm, is_t, score, per_nodes, max_distinct_edges=1024, expect_rc=0):
    M = len(keys)
    kmers = np.zeros(M, O.KMER_DTYPE)
    nodes = np.zeros(M, O.NODE_DTYPE)
    edges = np.zeros(max(1, M), O.EDGE_DTYPE)
    nn, ne, novf = C.c_uint64(), C.c_uint64(), C.c_uint64()
    is_t8 = np.ascontiguousarray(is_t, dtype=np.uint8)
    rc = L.agg_emul_run(C.c_void_p(keys.ctypes.data), C.c_void_p(vals.ctypes.data), C.c_uint64(M), C.c_void_p(rec_asm.ctypes.data),
                        C.c_uint32(0), C.c_void_p(is_t8.ctypes.data), C.c_int(1 if score else 0),
                        C.c_uint32(int(is_t8.sum())), C.c_uint32(int(len(is_t8) - is_t8.sum())), C.c_uint32(per_nodes),
                        C.c_uint32(max_distinct_edges), C.c_void_p(kmers.ctypes.data),
                        C.c_void_p(nodes.ctypes.data), C.c_void_p(edges.ctypes.data), C.byref(nn), C.byref(ne), C.byref(novf))
    assert rc == expect_rc, f"emulated aggregation returned {rc}"
    return kmers, nodes[:nn.value], edges[:ne.value], novf.value


CASES = {
    "small": (SynthSpec(n_genomes=6, n_targets=2, genome_len=20_000, n_contigs=3, seed=11, n_runs_every=2), 17, 10),
    "dups": (SynthSpec(n_genomes=9, n_targets=3, genome_len=30_000, n_contigs=4, seed=5, n_runs_every=3), 21, 50),
    "skew": (SynthSpec(n_genomes=8, n_targets=3, genome_len=40_000, n_contigs=4, seed=3, n_runs_every=5, skew=True), 15, 8),
}


@pytest.mark.parametrize("name", sorted(CASES))
@pytest.mark.parametrize("per_bucket", [512, 40, 3000], ids=["default", "tiny-buckets", "few-buckets"])
def test_emulated_aggregation_matches_oracle(agg, tmp_path, name, per_bucket):
    spec, k, w = CASES[name]
    paths, is_t = write_set(spec, tmp_path)
    want_k, want_n, want_e, offsets, _ = O._build_native(paths, k, w)
    O._get_penalty_native(want_k, want_n, offsets, is_t)
    keys, vals, rec_asm, is_t2 = _stream(spec, k, w)
    assert len(keys) == len(want_k)
    if per_bucket >= 3000 and len(np.unique(keys)) > 1024 * 2:
        pytest.skip("more distinct nodes than two buckets hold")
    kmers, nodes, edges, _ = _run(agg, keys, vals, rec_asm, is_t2, True, per_bucket)
    assert np.array_equal(kmers, want_k)
    assert np.array_equal(nodes, want_n)
    assert np.array_equal(edges, want_e)


def test_emulated_unscored_build(agg, tmp_path):
    spec, k, w = CASES["small"]
    paths, _ = write_set(spec, tmp_path)
    want_k, want_n, want_e, _, _ = O._build_native(paths, k, w)
    keys, vals, rec_asm, is_t = _stream(spec, k, w)
    kmers, nodes, edges, _ = _run(agg, keys, vals, rec_asm, is_t, False, 512)
    assert np.array_equal(kmers, want_k) and np.array_equal(nodes, want_n) and np.array_equal(edges, want_e)


@pytest.mark.parametrize("name", ["small", "skew"])
def test_emulated_edge_overflow_buckets(agg, tmp_path, name):
    """Edge buckets with more distinct pairs than the table is allowed to take go through the side path."""
    spec, k, w = CASES[name]
    paths, is_t = write_set(spec, tmp_path)
    want_k, want_n, want_e, offsets, _ = O._build_native(paths, k, w)
    O._get_penalty_native(want_k, want_n, offsets, is_t)
    keys, vals, rec_asm, is_t2 = _stream(spec, k, w)
    kmers, nodes, edges, n_ovf = _run(agg, keys, vals, rec_asm, is_t2, True, 512, max_distinct_edges=60)
    assert n_ovf > 0
    assert np.array_equal(nodes, want_n) and np.array_equal(edges, want_e)


def test_emulated_hot_key_spans_chunks(agg):
    """One hash owning thousands of k-mers: a group larger than a placement chunk, across many records."""
    rng = np.random.default_rng(1)
    n_rec, per = 40, 300
    keys, vals, rec_asm = [], [], []
    hot = np.uint64(0x1234_5678_9ABC_DEF0)
    for r in range(n_rec):
        h = rng.integers(0, 2**63, per, dtype=np.uint64) * np.uint64(2) + np.uint64(1)
        h[rng.random(per) < 0.6] = hot
        keys.append(h)
        vals.append(np.arange(per, dtype=np.uint64) * np.uint64(7) | (np.uint64(r) << np.uint64(32)))
        rec_asm.append(r // 4)
    keys, vals = np.concatenate(keys), np.concatenate(vals)
    rec_asm = np.asarray(rec_asm, dtype=np.uint32)
    is_t = np.arange(10) < 3
    kmers, nodes, edges, _ = _run(agg, keys, vals, rec_asm, is_t, True, 512)
    # reference by plain numpy: stable sort by key
    order = np.argsort(keys, kind="stable")
    assert np.array_equal(kmers.view(np.uint64), vals[order])
    uk, first, cnt = np.unique(keys[order], return_index=True, return_counts=True)
    assert np.array_equal(nodes["hash"], uk) and np.array_equal(nodes["start"], first) and np.array_equal(nodes["stop"], first + cnt)
    hot_i = int(np.searchsorted(uk, hot))
    assert cnt[hot_i] > 4096
    asm_of = rec_asm[(vals[order] >> np.uint64(32)).astype(np.int64)]
    for i in (hot_i, 0, len(uk) - 1):
        a = np.unique(asm_of[first[i]:first[i] + cnt[i]])
        assert nodes["n_tar"][i] == np.count_nonzero(is_t[a]) and nodes["n_neg"][i] == len(a) - np.count_nonzero(is_t[a])
    # edges: adjacent pairs inside a record, distinct assemblies
    rank = np.searchsorted(uk, keys)
    same = (vals[:-1] >> np.uint64(32)) == (vals[1:] >> np.uint64(32))
    u, v = np.minimum(rank[:-1], rank[1:])[same], np.maximum(rank[:-1], rank[1:])[same]
    a = rec_asm[(vals[:-1] >> np.uint64(32)).astype(np.int64)][same]
    trip = np.unique(np.stack([u, v, a], axis=1), axis=0)
    pairs, wgt = np.unique(trip[:, :2], axis=0, return_counts=True)
    assert np.array_equal(edges["first"], uk[pairs[:, 0]]) and np.array_equal(edges["second"], uk[pairs[:, 1]])
    assert np.array_equal(edges["weight"], wgt.astype(np.uint64))


def test_emulated_distinct_estimate(agg):
    """The hash-range sample that sizes the buckets: items per distinct key, sampled 1 key in 2^sbits."""
    rng = np.random.default_rng(3)
    uniq = rng.integers(0, 2**63, 40_000, dtype=np.uint64)
    keys = np.ascontiguousarray(np.concatenate([np.repeat(uniq[:10_000], 9), uniq[10_000:]]))   # 120k items, 40k keys: 3.0
    rng.shuffle(keys)
    exact = agg.agg_emul_estimate(C.c_void_p(keys.ctypes.data), C.c_uint64(len(keys)), C.c_int(0))
    assert abs(exact - 3.0) < 1e-9
    est = agg.agg_emul_estimate(C.c_void_p(keys.ctypes.data), C.c_uint64(len(keys)), C.c_int(3))
    assert abs(est - 3.0) < 0.3


@pytest.mark.parametrize("per_nodes", [700, 6000], ids=["two-chunks", "many-chunks"])
def test_emulated_edge_buckets_span_chunks(agg, per_nodes):
    """Few distinct hashes, many records: edge buckets hold several chunks of records and the same (pair,
    assembly) shows up in more than one of them -- the weight must still count every assembly once."""
    rng = np.random.default_rng(7)
    n_rec, per = 60, 250
    pool = rng.integers(0, 2**63, 40, dtype=np.uint64) * np.uint64(2) + np.uint64(1)
    keys, vals, rec_asm = [], [], []
    for r in range(n_rec):
        keys.append(pool[rng.integers(0, len(pool), per)])
        vals.append(np.arange(per, dtype=np.uint64) * np.uint64(3) | (np.uint64(r) << np.uint64(32)))
        rec_asm.append(r // 5)
    keys, vals = np.concatenate(keys), np.concatenate(vals)
    rec_asm = np.asarray(rec_asm, dtype=np.uint32)
    is_t = np.arange(12) < 4
    kmers, nodes, edges, n_ovf = _run(agg, keys, vals, rec_asm, is_t, True, per_nodes)
    assert n_ovf == 0
    uk = np.unique(keys)
    rank = np.searchsorted(uk, keys)
    same = (vals[:-1] >> np.uint64(32)) == (vals[1:] >> np.uint64(32))
    u, v = np.minimum(rank[:-1], rank[1:])[same], np.maximum(rank[:-1], rank[1:])[same]
    a = rec_asm[(vals[:-1] >> np.uint64(32)).astype(np.int64)][same]
    trip = np.unique(np.stack([u, v, a], axis=1), axis=0)
    pairs, wgt = np.unique(trip[:, :2], axis=0, return_counts=True)
    assert len(edges) == len(pairs)
    assert np.array_equal(edges["first"], uk[pairs[:, 0]]) and np.array_equal(edges["second"], uk[pairs[:, 1]])
    assert np.array_equal(edges["weight"], wgt.astype(np.uint64))


def _numpy_edges(keys, vals, rec_asm):
    uk = np.unique(keys)
    rank = np.searchsorted(uk, keys)
    same = (vals[:-1] >> np.uint64(32)) == (vals[1:] >> np.uint64(32))
    u, v = np.minimum(rank[:-1], rank[1:])[same], np.maximum(rank[:-1], rank[1:])[same]
    a = rec_asm[(vals[:-1] >> np.uint64(32)).astype(np.int64)][same]
    trip = np.unique(np.stack([u, v, a], axis=1), axis=0)
    pairs, wgt = np.unique(trip[:, :2], axis=0, return_counts=True)
    return uk[pairs[:, 0]], uk[pairs[:, 1]], wgt.astype(np.uint64)


def test_emulated_edge_key_bits_clash(agg):
    """Two neighbours of one node that differ only in the 10 hash bits the pair table's key leaves out: the
    check of those bits must send the bucket to the side path, and the edges must still be exact."""
    rng = np.random.default_rng(11)
    x = np.uint64(0x0000_1000_0000_0001)
    a = np.uint64(0x0040_2222_3333_4445)
    b = a ^ np.uint64(1 << 60)                 # same low 54 bits as a
    fill = rng.integers(1 << 40, 1 << 62, 400, dtype=np.uint64) | np.uint64(1)
    keys, vals, rec_asm = [], [], []
    for r, nb_ in enumerate([a, b, a, b, a]):
        rec = np.concatenate([fill[r * 50:(r + 1) * 50], [x, nb_], fill[250 + r * 10:260 + r * 10]])
        keys.append(rec)
        vals.append(np.arange(len(rec), dtype=np.uint64) * np.uint64(5) | (np.uint64(r) << np.uint64(32)))
        rec_asm.append(r // 2)
    keys, vals = np.concatenate(keys), np.concatenate(vals)
    rec_asm = np.asarray(rec_asm, dtype=np.uint32)
    is_t = np.array([True, False, False])
    kmers, nodes, edges, n_ovf = _run(agg, keys, vals, rec_asm, is_t, True, 512)
    assert n_ovf >= 1
    f, s2, w = _numpy_edges(keys, vals, rec_asm)
    assert np.array_equal(edges["first"], f) and np.array_equal(edges["second"], s2) and np.array_equal(edges["weight"], w)
    i_a = np.flatnonzero((edges["first"] == x) & (edges["second"] == a))
    i_b = np.flatnonzero((edges["first"] == x) & (edges["second"] == b))
    assert len(i_a) == 1 and len(i_b) == 1 and edges["weight"][i_a[0]] == 3 and edges["weight"][i_b[0]] == 2


def test_emulated_zero_hash_takes_the_other_path(agg):
    """0 marks "no owned neighbour": a stream that holds the hash 0 is left to the sort-based path."""
    keys = np.array([5, 0, 9, 7], dtype=np.uint64)
    vals = np.arange(4, dtype=np.uint64)
    _run(agg, keys, vals, np.zeros(1, np.uint32), np.array([True]), False, 512, expect_rc=-3)


def test_emulated_all_distinct_hashes(agg):
    """Unrelated genomes: every minimizer its own node, as many edges as records, self-loops and ties absent."""
    rng = np.random.default_rng(2)
    n_rec, per = 30, 400
    keys = rng.permutation(np.arange(1, n_rec * per + 1, dtype=np.uint64) * np.uint64(0x9E3779B97F4A7C15))
    vals = np.concatenate([np.arange(per, dtype=np.uint64) | (np.uint64(r) << np.uint64(32)) for r in range(n_rec)])
    rec_asm = (np.arange(n_rec) // 3).astype(np.uint32)
    is_t = np.arange(10) < 2
    kmers, nodes, edges, _ = _run(agg, keys, vals, rec_asm, is_t, True, 300)
    assert len(nodes) == len(keys) and len(edges) == n_rec * (per - 1)
    f, s2, w = _numpy_edges(keys, vals, rec_asm)
    assert np.array_equal(edges["first"], f) and np.array_equal(edges["second"], s2) and np.array_equal(edges["weight"], w)
