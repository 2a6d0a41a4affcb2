"""BASELINE.json configs[1] at full size (500 genomes x 5 Mbp, k=21, w=200) on one B200: the oracle
cannot run 2.5 Gbp in seconds, so parity is checked through size-independent properties:

  * ordering / tiling / uniqueness invariants of the three output arrays,
  * the device-resident path and the pipelined end-to-end path give byte-identical arrays,
  * the number of k-mers equals the length of the ordered minimizer stream,
  * for records sampled at random the oracle's minimizers are EXACTLY the graph's k-mers of that record,
  * class counts are bounded by the class sizes and penalties recomputed on the host match bit for bit.
"""
from __future__ import annotations

import ctypes as C

import numpy as np
import pytest

from oracle import oracle as O
from seqwin_b200 import _lib
from seqwin_b200.dist import export_graph
from seqwin_b200.synth import SynthSet, SynthSpec
from tests.helpers import check_graph_invariants, digest

pytestmark = pytest.mark.gpu
K, W = 21, 200


def test_full_size_properties():
    L = _lib.lib()
    spec = SynthSpec(n_genomes=500, n_targets=100, genome_len=5_000_000, n_contigs=50, seed=42)
    ss = SynthSet(spec)
    arrs, lens, asm_of = [], [], []
    for g in range(spec.n_genomes):
        for _, seq in ss.records(g):
            arrs.append(np.ascontiguousarray(seq))
            lens.append(len(seq))
            asm_of.append(g)
    n = len(arrs)
    ptrs = (C.c_void_p * n)(*[a.ctypes.data for a in arrs])
    lens_a, asm_a = np.asarray(lens, dtype=np.uint32), np.asarray(asm_of, dtype=np.uint32)
    b, d, g = C.c_void_p(), C.c_void_p(), C.c_void_p()
    _lib.check(L.sw_batch_from_memory(ptrs, lens_a.ctypes.data, asm_a.ctypes.data, None, n, spec.n_genomes, 16, C.byref(b)))
    is_t = np.ascontiguousarray(ss.is_targets, dtype=np.bool_)
    try:
        assert L.sw_batch_n_bases(b) == spec.n_genomes * spec.genome_len
        # device-resident path
        _lib.check(L.sw_dev_upload(b, C.byref(d)))
        _lib.check(L.sw_dev_build(d, K, W, C.byref(g), None))
        _lib.check(L.sw_graph_penalty(g, None, 0, is_t.ctypes.data, len(is_t), None))
        kmers, nodes, edges = export_graph(L, g)
        L.sw_graph_free(g)
        cnt = C.c_size_t()
        _lib.check(L.sw_dev_sketch(d, K, W, None, None, None, 0, C.byref(cnt)))
        L.sw_dev_batch_free(d)
        assert cnt.value == len(kmers)
        offsets = np.arange(0, n + 1, 50, dtype=np.uint32)
        check_graph_invariants(kmers, nodes, edges, offsets)
        # expected density 2/(w+1) within a few percent on random sequence
        assert abs(len(kmers) / (2.5e9 * 2 / (W + 1)) - 1) < 0.05
        # pipelined end-to-end path: byte-identical
        g = C.c_void_p()
        _lib.check(L.sw_build_from_batch_scored(b, K, W, is_t.ctypes.data, len(is_t), C.byref(g), None))
        k2, n2, e2 = export_graph(L, g)
        L.sw_graph_free(g)
        assert digest(k2) == digest(kmers) and digest(n2) == digest(nodes) and digest(e2) == digest(edges)
    finally:
        L.sw_batch_free(b)

    # sampled records: the graph's k-mers of that record == the oracle's minimizers of that record
    rng = np.random.default_rng(0)
    hash_of_kmer = np.repeat(nodes["hash"], (nodes["stop"] - nodes["start"]).astype(np.int64))
    order = np.argsort(kmers["record_idx"], kind="stable")
    rec_sorted = kmers["record_idx"][order]
    for r in rng.integers(0, n, 6).tolist() + [4 * 50 + 3, 9 * 50 + 7]:   # genome 9 carries N runs / soft masking
        lo, hi = np.searchsorted(rec_sorted, [r, r + 1])
        sel = order[lo:hi]
        sel = sel[np.argsort(kmers["pos"][sel], kind="stable")]
        h1, pos = O.minimize(arrs[r].tobytes(), K, W)
        assert np.array_equal(kmers["pos"][sel], pos), f"record {r}: positions differ"
        assert np.array_equal(hash_of_kmer[sel], h1), f"record {r}: hashes differ"

    # class counts and penalties
    assert nodes["n_tar"].max() <= 100 and nodes["n_neg"].max() <= 400
    assert np.all(nodes["n_tar"] + nodes["n_neg"] >= 1)
    ft = nodes["n_tar"].astype(np.float64) * (1.0 / 100.0)
    fn = nodes["n_neg"].astype(np.float64) * (1.0 / 400.0)
    assert np.array_equal(nodes["penalty"], np.sqrt((1.0 - ft) * (1.0 - ft) + fn * fn))
    # every edge endpoint is a node and weights count assemblies
    assert edges["weight"].max() <= 500


def test_full_size_matches_reference(tmp_path_factory):
    """The whole C2 workload (500 genomes, 2.5 Gbp) as FASTA files: the drop-in entry points
    (KmerGraph + _get_penalty on the GPU) against the UNMODIFIED reference extension (oracle/_ref, all host
    cores) -- kmers, scored nodes, edges and record offsets compared bit for bit, as the reference's own
    tests/smoke/test_outputs.py:22-58 compares its graph.npz."""
    import os
    import shutil

    import bench
    from seqwin_b200.graph import KmerGraph, _get_penalty

    ref = O.load_reference()
    if ref is None:
        pytest.skip("oracle/_ref (the compiled reference) is not present")
    spec = SynthSpec(n_genomes=500, n_targets=100, genome_len=5_000_000, n_contigs=50, seed=42)
    ss = SynthSet(spec)
    d = bench.shm_dir("seqwin_b200_fullsize_")
    try:
        n_bases = bench.write_genomes(ss, range(spec.n_genomes), d, 8)
        assert n_bases == spec.n_genomes * spec.genome_len
        paths = [str(bench.fasta_path(d, g)) for g in range(spec.n_genomes)]
        is_t = np.ascontiguousarray(ss.is_targets, dtype=np.bool_)
        cores = os.cpu_count() or 8
        g = KmerGraph(paths, K, W, n_cpu=cores)
        _get_penalty(g.kmers, g.nodes, g.record_offsets, is_t)
        rk, rn, re_, ro, rids = ref._build_native(paths, K, W, cores, False)
        ref._get_penalty_native(rk, rn, ro, is_t, cores)
    finally:
        shutil.rmtree(d, ignore_errors=True)
    assert len(rk) > 24_000_000
    assert np.array_equal(g.record_offsets, ro) and g.record_ids == rids
    for name, ours, theirs in (("kmers", g.kmers, rk), ("nodes", g.nodes, rn), ("edges", g.edges, re_)):
        assert ours.dtype == theirs.dtype and ours.shape == theirs.shape, name
        assert digest(ours) == digest(theirs), f"{name} differ from the reference at full size"
