"""Host-side mirror of the reference's graph API (``src/seqwin/graph/__init__.py:40-196``).

Same names, argument meaning and error behaviour, so parity tests read like the reference's
``tests/smoke/test_graph.py``; the native module is :mod:`seqwin_b200._core`.
"""
from __future__ import annotations

from collections.abc import Iterable
from pathlib import Path

import numpy as np
from numpy.typing import NDArray

from ._core import (EDGE_DTYPE, KMER_DTYPE, NODE_DTYPE, _build_native, _filter_kmers_native,
                    _get_penalty_native)

__all__ = ["KMER_DTYPE", "NODE_DTYPE", "EDGE_DTYPE", "KmerGraph", "_get_penalty", "_filter_kmers",
           "_filter_edges_and_nodes", "penalty_threshold", "save_graph"]


class KmerGraph:
    """The minimizer graph (``graph/__init__.py:61-146``): kmers grouped and sorted by hash, nodes
    and edges sorted by hash, ``record_offsets`` cumulative per assembly, ``record_ids`` per assembly."""
    __slots__ = ("kmers", "nodes", "edges", "record_offsets", "record_ids")

    def __init__(self, assembly_paths: Iterable[Path], kmerlen: int, windowsize: int,
                 low_memory: bool = False, n_cpu: int = 1) -> None:
        self.kmers, self.nodes, self.edges, self.record_offsets, self.record_ids = _build_native(
            list(str(p) for p in assembly_paths), int(kmerlen), int(windowsize), int(n_cpu), bool(low_memory))


def _get_penalty(kmers: NDArray[np.void], nodes: NDArray[np.void], record_offsets: NDArray[np.uint32],
                 is_targets: Iterable[bool], n_cpu: int = 1) -> None:
    """``graph/__init__.py:149-171``: populate n_tar / n_neg / penalty in place."""
    _get_penalty_native(kmers, nodes, record_offsets,
                        np.asarray(is_targets, dtype=np.bool_, order="C"), int(n_cpu))


def _filter_kmers(kmers: NDArray[np.void], nodes: NDArray[np.void], used_hashes):
    """``graph/__init__.py:174-196``: keep the nodes (and their k-mers) whose hash is in ``used_hashes``."""
    return _filter_kmers_native(kmers, nodes, used_hashes)


def _filter_edges_and_nodes(nodes: NDArray[np.void], edges: NDArray[np.void], edge_weight_th: float):
    """Array part of ``kmers.py:132-162``: remove edges with ``weight <= uintp(edge_weight_th)`` and the
    nodes left without an edge.  Returns ``(nodes, edges)``; the caller builds the networkx graph from
    them as the reference does (``kmers.py:164-171``).  Runs on the GPU (``sw_filter_edges_and_nodes``)."""
    import ctypes as C

    from . import _lib
    if nodes.dtype != NODE_DTYPE or edges.dtype != EDGE_DTYPE:
        raise TypeError("nodes / edges must have the graph dtypes")
    nodes = np.ascontiguousarray(nodes)
    edges = np.ascontiguousarray(edges)
    nodes_out, edges_out = np.empty_like(nodes), np.empty_like(edges)
    nn, ne = C.c_size_t(), C.c_size_t()
    _lib.check(_lib.lib().sw_filter_edges_and_nodes(
        nodes.ctypes.data, len(nodes), edges.ctypes.data, len(edges), C.c_uint64(int(np.uint64(edge_weight_th))),
        nodes_out.ctypes.data, edges_out.ctypes.data, C.byref(nn), C.byref(ne)))
    return nodes_out[:nn.value].copy(), edges_out[:ne.value].copy()


def penalty_threshold(count_sums, n_tar: int, n_neg: int, stringency: int = 5, penalty_th_cap: float = 0.2):
    """Automatic penalty threshold of ``kmers.py:424-440`` (the branch without Mash) from the three
    integer sums ``(sum n_tar, sum n_tar^2, sum n_tar * n_neg)`` over the scored nodes --
    ``sw_graph_count_sums`` on a device-resident graph, or ``count_sums(nodes)`` for a numpy array.
    Returns ``(penalty_th, e_absence_tar, e_presence_neg)``."""
    s_t, s_tt, s_tn = (int(x) for x in count_sums)
    # The reference evaluates np.sum(frac * n_tar) / np.sum(n_tar) in float64; the exact integer sums used
    # here agree with it to the last few ulps (not bit for bit).  Empty denominators give nan / inf like
    # numpy does there (with a warning), not an exception.
    with np.errstate(divide="ignore", invalid="ignore"):
        e_absence_tar = float(1.0 - np.float64(s_tt) / np.float64(n_tar * s_t))
        e_presence_neg = float(np.float64(s_tn) / np.float64(n_neg * s_t))
    penalty_th = (1 - stringency / 10) * (e_absence_tar * e_presence_neg) ** 0.5
    return min(penalty_th, penalty_th_cap), e_absence_tar, e_presence_neg


def count_sums(nodes: NDArray[np.void]):
    """The sums of :func:`penalty_threshold` for a host array (exact integer arithmetic)."""
    t = nodes["n_tar"].astype(np.uint64)
    return int(t.sum()), int((t * t).sum()), int((t * nodes["n_neg"].astype(np.uint64)).sum())


def save_graph(path, kmers, nodes, edges, record_offsets) -> None:
    """``graph.npz`` exactly as ``core.py:134-145`` writes it."""
    np.savez(path, allow_pickle=False, kmers=kmers, nodes=nodes, edges=edges, record_offsets=record_offsets)
