"""CPU-only checks of the boundary: library loads, exports every declared symbol, host logic."""
from __future__ import annotations

import ctypes as C
import re
from pathlib import Path

import numpy as np
import pytest

from seqwin_b200 import _lib
from seqwin_b200._core import EDGE_DTYPE, KMER_DTYPE, NODE_DTYPE
from seqwin_b200.graph import KmerGraph, _filter_kmers, _get_penalty

ROOT = Path(__file__).resolve().parents[1]


def test_library_exports_every_declared_symbol():
    header = (ROOT / "include" / "seqwin_b200.h").read_text()
    declared = set(re.findall(r"\b(sw_[a-z_0-9]+)\s*\(", header))
    assert len(declared) >= 20
    L = _lib.lib()
    for name in sorted(declared):
        assert hasattr(L, name), f"libseqwin_b200.so does not export {name}"
    assert declared == set(_lib.PROTOTYPES), declared ^ set(_lib.PROTOTYPES)


def test_dtype_layouts():
    # reference tests/smoke/test_graph.py:45-64
    assert KMER_DTYPE.itemsize == 8 and KMER_DTYPE.names == ("pos", "record_idx")
    assert NODE_DTYPE.itemsize == 40 and NODE_DTYPE.names == ("hash", "start", "stop", "n_tar", "n_neg", "penalty")
    assert EDGE_DTYPE.itemsize == 24 and [EDGE_DTYPE.fields[n][1] for n in ("first", "second", "weight")] == [0, 8, 16]


def test_no_cpu_fallback_without_device(have_gpu, fixture_paths):
    if have_gpu:
        pytest.skip("GPU present")
    with pytest.raises(RuntimeError, match="CUDA"):
        KmerGraph(fixture_paths, 17, 10)


def test_build_rejects_is_targets_argument(fixture_paths):
    # reference tests/smoke/test_graph.py:130-141
    with pytest.raises(TypeError):
        KmerGraph(fixture_paths[:2], kmerlen=7, windowsize=10, is_targets=[True, False])


def _batch_from_fasta(paths, threads=2):
    L = _lib.lib()
    arr = (C.c_char_p * max(1, len(paths)))(*[str(p).encode() for p in paths])
    b = C.c_void_p()
    _lib.check(L.sw_batch_from_fasta(arr, len(paths), threads, C.byref(b)))
    return b


def test_host_ingest_counts(edge_paths, fixture_paths):
    L = _lib.lib()
    b = _batch_from_fasta(edge_paths[0])
    try:
        # a_plain 2, b_crlf 2, c_odd 4, d_gz 2, e_empty 0, f_shared 2
        assert L.sw_batch_n_records(b) == 12
        from tests.cases import EDGE_FILES
        n_bases = 0
        for content in EDGE_FILES.values():
            for line in content.replace(b"\r\n", b"\n").split(b"\n"):
                if line and not line.startswith(b">") and line.strip():
                    n_bases += len(b"".join(line.split()))
        assert L.sw_batch_n_bases(b) == n_bases
        assert L.sw_batch_packed_bytes(b) % 16 == 0
    finally:
        L.sw_batch_free(b)
    b = _batch_from_fasta(fixture_paths, threads=99)
    assert L.sw_batch_n_records(b) == 4
    L.sw_batch_free(b)


def test_host_ingest_errors(tmp_path):
    L = _lib.lib()
    with pytest.raises(RuntimeError, match="Unable to open"):
        _batch_from_fasta([tmp_path / "missing.fasta"])
    bad = tmp_path / "bad.fasta"
    bad.write_text("ACGT\n>late\nACGT\n")
    with pytest.raises(RuntimeError, match="sequence encountered before header"):
        _batch_from_fasta([bad])
    assert L.sw_last_error()


def test_filter_kmers_known_answer():
    # reference tests/smoke/test_graph.py:190-219 (host-side entry point, no device needed)
    kmers = np.array([(10, 0), (11, 0), (20, 1), (30, 2), (31, 2), (32, 2)], dtype=KMER_DTYPE)
    nodes = np.array([(10, 0, 2, 1, 0, 0.1), (20, 2, 3, 1, 0, 0.2), (30, 3, 6, 1, 1, 0.3)], dtype=NODE_DTYPE)
    kn, nn = _filter_kmers(kmers, nodes, {30, 10})
    assert nn["hash"].tolist() == [10, 30] and nn["start"].tolist() == [0, 2] and nn["stop"].tolist() == [2, 5]
    assert nn["n_neg"].tolist() == [0, 1] and nn["penalty"].tolist() == [0.1, 0.3]
    assert np.array_equal(kn, np.array([(10, 0), (11, 0), (30, 2), (31, 2), (32, 2)], dtype=KMER_DTYPE))
    kn, nn = _filter_kmers(kmers, nodes, frozenset(np.array([20, 99], dtype=np.uint64)))
    assert nn["hash"].tolist() == [20] and len(kn) == 1
    kn, nn = _filter_kmers(kmers, nodes, [])
    assert len(kn) == 0 and len(nn) == 0


def _penalty_inputs():
    kmers = np.array([(0, 0), (1, 0), (2, 1), (3, 2), (4, 4), (5, 2), (6, 3), (7, 5), (8, 6), (9, 4)], dtype=KMER_DTYPE)
    nodes = np.array([(10, 0, 5, 0, 0, 0.0), (20, 5, 7, 0, 0, 0.0), (30, 7, 9, 0, 0, 0.0), (40, 9, 10, 0, 0, 0.0),
                      (50, 10, 10, 9, 9, 9.0), (60, 5, 9, 0, 0, 0.0)], dtype=NODE_DTYPE)
    return kmers, nodes, np.array([0, 2, 4, 5, 7], dtype=np.uint32), np.array([True, False, True, False])


def test_get_penalty_argument_validation():
    """The host-side half of reference tests/smoke/test_graph.py:307-341 (raised before any device work)."""
    kmers, nodes, offsets, is_t = _penalty_inputs()
    ro = nodes.copy()
    ro.flags.writeable = False
    with pytest.raises(ValueError):
        _get_penalty(kmers, ro, offsets, is_t)
    with pytest.raises(ValueError):
        _get_penalty(kmers, nodes.copy(), offsets[:-1], is_t)
    with pytest.raises(ValueError):
        _get_penalty(kmers, nodes.copy(), np.array([1, 2, 4, 5, 7], dtype=np.uint32), is_t)
    with pytest.raises(ValueError):
        _get_penalty(kmers, nodes.copy(), np.array([0, 3, 2, 5, 7], dtype=np.uint32), is_t)
    with pytest.raises(TypeError):
        _get_penalty(kmers, nodes.copy(), offsets.astype(np.uintp), is_t)
    with pytest.raises(TypeError):
        _get_penalty(kmers, nodes.copy(), offsets.astype(np.uint64), is_t)
    with pytest.raises(ValueError):
        _get_penalty(kmers, nodes.copy(), offsets, [False] * 4)
    with pytest.raises(ValueError):
        _get_penalty(kmers, nodes.copy(), offsets, [True] * 4)
    with pytest.raises(ValueError):
        _get_penalty(kmers, nodes.copy(), offsets, np.array([[True, False], [True, False]]))


def test_kw_limits(have_gpu, fixture_paths):
    if not have_gpu:
        pytest.skip("needs the device path to reach the k/w checks")
    for k, w in [(2, 10), (70000, 10), (17, 0)]:
        with pytest.raises(RuntimeError):
            KmerGraph(fixture_paths, k, w)


def test_batch_concat_tables(edge_paths, fixture_paths):
    """sw_batch_concat: the concatenation of two packed batches has the tables of the batch parsed in one go."""
    L = _lib.lib()
    paths = [str(p) for p in edge_paths[0]] + [str(p) for p in fixture_paths]
    whole = _batch_from_fasta(paths)
    a, b = _batch_from_fasta(paths[:4]), _batch_from_fasta(paths[4:])
    cat = C.c_void_p()
    arr = (C.c_void_p * 2)(a.value, b.value)
    _lib.check(L.sw_batch_concat(arr, 2, C.byref(cat)))
    try:
        for fn in (L.sw_batch_n_bases, L.sw_batch_n_records, L.sw_batch_packed_bytes):
            assert fn(cat) == fn(whole)
        n = len(paths) + 1
        o1, o2 = np.zeros(n, np.int64), np.zeros(n, np.int64)
        _lib.check(L.sw_batch_record_offsets(cat, o1.ctypes.data, n))
        _lib.check(L.sw_batch_record_offsets(whole, o2.ctypes.data, n))
        assert np.array_equal(o1, o2)
    finally:
        for x in (whole, a, b, cat):
            L.sw_batch_free(x)
