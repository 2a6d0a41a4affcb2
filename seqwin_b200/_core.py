"""Drop-in replacement for the reference's pybind11 module ``seqwin.graph._core``.

Same three callables, argument names, defaults, return types and exception types as
``cpp/src/bindings/python_bindings.cpp:43-169``; the work is done by the CUDA library behind the
C ABI of ``include/seqwin_b200.h``.  Copy (or symlink) this file as ``seqwin/graph/_core.py`` next
to ``libseqwin_b200.so`` and the unmodified Seqwin package runs on the GPU.
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

try:  # inside the seqwin_b200 package
    from ._lib import (SW_EDGES, SW_KMERS, SW_NODES, SW_OFFSETS, check, lib)
except ImportError:  # copied into seqwin/graph/ next to _lib.py
    from _lib import (SW_EDGES, SW_KMERS, SW_NODES, SW_OFFSETS, check, lib)  # type: ignore

KMER_DTYPE = np.dtype([("pos", np.uint32), ("record_idx", np.uint32)])
NODE_DTYPE = np.dtype([("hash", np.uint64), ("start", np.uintp), ("stop", np.uintp),
                       ("n_tar", np.uint32), ("n_neg", np.uint32), ("penalty", np.float64)])
EDGE_DTYPE = np.dtype([("first", np.uint64), ("second", np.uint64), ("weight", np.uintp)])


def _graph_to_numpy(L, g):
    kmers = np.empty(L.sw_graph_size(g, SW_KMERS), dtype=KMER_DTYPE)
    nodes = np.empty(L.sw_graph_size(g, SW_NODES), dtype=NODE_DTYPE)
    edges = np.empty(L.sw_graph_size(g, SW_EDGES), dtype=EDGE_DTYPE)
    offsets = np.empty(L.sw_graph_size(g, SW_OFFSETS), dtype=np.uint32)
    check(L.sw_graph_export(g, kmers.ctypes.data, nodes.ctypes.data, edges.ctypes.data, offsets.ctypes.data))
    ids = []
    for a in range(len(offsets) - 1):
        n = L.sw_graph_n_records(g, a)
        ids.append(tuple(L.sw_graph_record_id(g, a, i).decode() for i in range(n)))
    return kmers, nodes, edges, offsets, ids


def _build_native(assembly_paths, kmerlen, windowsize, n_cpu=1, low_memory=False):
    """python_bindings.cpp:50-90 -> (kmers, nodes, edges, record_offsets, ids_by_assembly)."""
    if isinstance(assembly_paths, (str, bytes)) or not all(isinstance(p, str) for p in assembly_paths):
        raise TypeError("assembly_paths must be a list of str")
    for name, v in (("kmerlen", kmerlen), ("windowsize", windowsize), ("n_cpu", n_cpu)):
        if isinstance(v, bool) or not isinstance(v, (int, np.integer)) or v < 0:
            raise TypeError(f"{name} must be a non-negative int")
    L = lib()
    paths = [os.fsencode(p) for p in assembly_paths]
    arr = (C.c_char_p * max(1, len(paths)))(*paths)
    g = C.c_void_p()
    check(L.sw_build(arr, len(paths), int(kmerlen), int(windowsize), max(1, int(n_cpu)),
                     1 if low_memory else 0, C.byref(g)))
    try:
        return _graph_to_numpy(L, g)
    finally:
        L.sw_graph_free(g)


def _require(arr, dtype, name, writable=False):
    # pybind11 `.noconvert()` array_t<T, c_style>: a wrong dtype / layout is a TypeError
    if not isinstance(arr, np.ndarray) or arr.dtype != dtype or not arr.flags.c_contiguous:
        raise TypeError(f"{name}: expected a C-contiguous numpy array of dtype {dtype}")
    if writable and not arr.flags.writeable:
        raise ValueError(f"{name} must be writable")


def _get_penalty_native(kmers, nodes, record_offsets, is_targets, n_cpu=1):
    """python_bindings.cpp:92-135; in place on ``nodes``; returns None."""
    _require(kmers, KMER_DTYPE, "kmers")
    _require(nodes, NODE_DTYPE, "nodes", writable=True)
    _require(record_offsets, np.dtype(np.uint32), "record_offsets")
    _require(is_targets, np.dtype(np.bool_), "is_targets")
    if nodes.ndim != 1 or kmers.ndim != 1 or record_offsets.ndim != 1:
        raise ValueError("kmers, nodes and record_offsets must be 1-D")
    # the reference reads shape[0] of whatever it is given; a 2-D is_targets then fails the
    # len(record_offsets) == len(is_targets) + 1 check (ValueError)
    n_assemblies = is_targets.shape[0] if is_targets.ndim >= 1 else 0
    if is_targets.ndim != 1:
        raise ValueError("len(record_offsets) must equal len(is_targets) + 1")
    check(lib().sw_get_penalty(kmers.ctypes.data, kmers.shape[0], nodes.ctypes.data, nodes.shape[0],
                               record_offsets.ctypes.data, record_offsets.shape[0],
                               is_targets.ctypes.data, n_assemblies, max(1, int(n_cpu))))
    return None


def _filter_kmers_native(kmers, nodes, used_hashes):
    """python_bindings.cpp:137-168 -> (kmers_new, nodes_new)."""
    _require(kmers, KMER_DTYPE, "kmers")
    _require(nodes, NODE_DTYPE, "nodes")
    used = np.fromiter((int(h) for h in used_hashes), dtype=np.uint64)
    L = lib()
    nk, nn = C.c_size_t(), C.c_size_t()
    args = (kmers.ctypes.data, kmers.shape[0], nodes.ctypes.data, nodes.shape[0], used.ctypes.data, used.shape[0])
    check(L.sw_filter_kmers(*args, None, None, C.byref(nk), C.byref(nn)))
    kmers_new = np.empty(nk.value, dtype=KMER_DTYPE)
    nodes_new = np.empty(nn.value, dtype=NODE_DTYPE)
    check(L.sw_filter_kmers(*args, kmers_new.ctypes.data, nodes_new.ctypes.data, C.byref(nk), C.byref(nn)))
    return kmers_new, nodes_new
