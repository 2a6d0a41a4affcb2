"""ctypes face of oracle/oracle.c -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
import this module.  It mirrors the three entry points of the reference extension
(cpp/src/bindings/python_bindings.cpp:50-168) so parity tests can call the oracle, the compiled
reference (``load_reference()``) and the CUDA path with the same arguments.
"""
from __future__ import annotations

import ctypes as C
import importlib
import os
import subprocess
import sys
from pathlib import Path

import numpy as np

HERE = Path(__file__).resolve().parent

KMER_DTYPE = np.dtype([("pos", np.uint32), ("record_idx", np.uint32)])
NODE_DTYPE = np.dtype([("hash", np.uint64), ("start", np.uintp), ("stop", np.uintp),
                       ("n_tar", np.uint32), ("n_neg", np.uint32), ("penalty", np.float64)])
EDGE_DTYPE = np.dtype([("first", np.uint64), ("second", np.uint64), ("weight", np.uintp)])

_lib = None


def build(force: bool = False) -> None:
    """Compile liboracle.so (and oracle/_ref when /root/reference is present)."""
    if force or not (HERE / "liboracle.so").exists():
        subprocess.check_call(["make", "-s", "-C", str(HERE), str(HERE / "liboracle.so")])
    if os.path.isdir(os.environ.get("SEQWIN_REF", "/root/reference")):
        subprocess.check_call(["make", "-s", "-C", str(HERE), "ref"])


def lib() -> C.CDLL:
    global _lib
    if _lib is None:
        build()
        L = C.CDLL(str(HERE / "liboracle.so"))
        L.orc_build.restype = C.c_void_p
        L.orc_build.argtypes = [C.POINTER(C.c_char_p), C.c_size_t, C.c_uint, C.c_size_t]
        L.orc_graph_error.restype = C.c_char_p
        L.orc_graph_error.argtypes = [C.c_void_p]
        L.orc_graph_size.restype = C.c_size_t
        L.orc_graph_size.argtypes = [C.c_void_p, C.c_int]
        L.orc_graph_export.restype = None
        L.orc_graph_export.argtypes = [C.c_void_p] * 5
        L.orc_graph_record_id.restype = C.c_char_p
        L.orc_graph_record_id.argtypes = [C.c_void_p, C.c_size_t]
        L.orc_graph_free.restype = None
        L.orc_graph_free.argtypes = [C.c_void_p]
        L.orc_hash_kmer.restype = C.c_int
        L.orc_hash_kmer.argtypes = [C.c_char_p, C.c_uint, C.POINTER(C.c_uint64), C.POINTER(C.c_uint64)]
        L.orc_minimize_export.restype = C.c_size_t
        L.orc_minimize_export.argtypes = [C.c_char_p, C.c_size_t, C.c_uint, C.c_size_t,
                                          C.c_void_p, C.c_void_p, C.c_size_t]
        L.orc_get_penalty.restype = C.c_int
        L.orc_get_penalty.argtypes = [C.c_void_p, C.c_size_t, C.c_void_p, C.c_size_t, C.c_void_p,
                                      C.c_size_t, C.c_void_p, C.c_size_t, C.c_char_p, C.c_size_t]
        L.orc_filter_kmers.restype = None
        L.orc_filter_kmers.argtypes = [C.c_void_p, C.c_void_p, C.c_size_t, C.c_void_p, C.c_size_t,
                                       C.c_void_p, C.c_void_p, C.POINTER(C.c_size_t), C.POINTER(C.c_size_t)]
        _lib = L
    return _lib


def hash_kmer(kmer: str) -> tuple[int, int] | None:
    h0, h1 = C.c_uint64(), C.c_uint64()
    b = kmer.encode()
    ok = lib().orc_hash_kmer(b, len(b), C.byref(h0), C.byref(h1))
    return (h0.value, h1.value) if ok else None


def minimize(seq: bytes | str, k: int, w: int) -> tuple[np.ndarray, np.ndarray]:
    """(h1, pos) of the minimizers of one record, in emission order."""
    if isinstance(seq, str):
        seq = seq.encode()
    cap = max(16, len(seq))
    h1 = np.empty(cap, np.uint64)
    pos = np.empty(cap, np.uint32)
    n = lib().orc_minimize_export(seq, len(seq), k, w, h1.ctypes.data, pos.ctypes.data, cap)
    return h1[:n].copy(), pos[:n].copy()


def _build_native(assembly_paths, kmerlen, windowsize, n_cpu=1, low_memory=False):
    L = lib()
    paths = [os.fsencode(str(p)) for p in assembly_paths]
    arr = (C.c_char_p * max(1, len(paths)))(*paths)
    g = L.orc_build(arr, len(paths), int(kmerlen), int(windowsize))
    try:
        err = L.orc_graph_error(g)
        if err:
            raise RuntimeError(err.decode())
        kmers = np.empty(L.orc_graph_size(g, 0), KMER_DTYPE)
        nodes = np.empty(L.orc_graph_size(g, 1), NODE_DTYPE)
        edges = np.empty(L.orc_graph_size(g, 2), EDGE_DTYPE)
        offsets = np.empty(L.orc_graph_size(g, 3), np.uint32)
        L.orc_graph_export(g, kmers.ctypes.data, nodes.ctypes.data, edges.ctypes.data, offsets.ctypes.data)
        ids = []
        for a in range(len(paths)):
            ids.append(tuple(L.orc_graph_record_id(g, r).decode() for r in range(offsets[a], offsets[a + 1])))
        return kmers, nodes, edges, offsets, ids
    finally:
        L.orc_graph_free(g)


def _get_penalty_native(kmers, nodes, record_offsets, is_targets, n_cpu=1):
    is_t = np.ascontiguousarray(is_targets, dtype=np.bool_)
    err = C.create_string_buffer(256)
    rc = lib().orc_get_penalty(kmers.ctypes.data, len(kmers), nodes.ctypes.data, len(nodes),
                               record_offsets.ctypes.data, len(record_offsets),
                               is_t.ctypes.data, len(is_t), err, 256)
    if rc:
        raise ValueError(err.value.decode())


def _filter_kmers_native(kmers, nodes, used_hashes):
    used = np.fromiter((int(h) for h in used_hashes), dtype=np.uint64)
    nk, nn = C.c_size_t(), C.c_size_t()
    L = lib()
    L.orc_filter_kmers(kmers.ctypes.data, nodes.ctypes.data, len(nodes), used.ctypes.data, len(used),
                       None, None, C.byref(nk), C.byref(nn))
    ko = np.empty(nk.value, KMER_DTYPE)
    no = np.empty(nn.value, NODE_DTYPE)
    L.orc_filter_kmers(kmers.ctypes.data, nodes.ctypes.data, len(nodes), used.ctypes.data, len(used),
                       ko.ctypes.data, no.ctypes.data, C.byref(nk), C.byref(nn))
    return ko, no


def filter_edges_and_nodes(nodes, edges, edge_weight_th):
    """Array part of the reference's `_filter_edges_and_nodes` (src/seqwin/kmers.py:147-162), restated:
    edges with weight > uintp(th), then the nodes found by searchsorted for the unique endpoints.
    Pinned against the reference's own function by tests/golden/arrays/filter_*.npz."""
    th = np.uint64(edge_weight_th)   # kmers.py:151 -- np.uintp(edge_weight_th): truncation
    kept = edges[edges["weight"] > th]
    ends = np.unique(np.concatenate([kept["first"], kept["second"]]))
    return nodes[np.searchsorted(nodes["hash"], ends)], kept


def load_reference():
    """The UNMODIFIED reference extension compiled into oracle/_ref (None if not built)."""
    ref_dir = HERE / "_ref"
    if not any((ref_dir / "seqwin_ref").glob("_core*.so")):
        try:
            build()
        except Exception:
            return None
    if not any((ref_dir / "seqwin_ref").glob("_core*.so")):
        return None
    if str(ref_dir) not in sys.path:
        sys.path.insert(0, str(ref_dir))
    return importlib.import_module("seqwin_ref._core")
