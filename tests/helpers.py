from __future__ import annotations

import hashlib

import numpy as np


def digest(a: np.ndarray) -> str:
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


def assert_graph_equal(got, want, what=""):
    """Bit-exact comparison of (kmers, nodes, edges, record_offsets[, ids])."""
    names = ["kmers", "nodes", "edges", "record_offsets"]
    for name, g, w in zip(names, got[:4], want[:4]):
        assert g.dtype == w.dtype, f"{what} {name} dtype {g.dtype} != {w.dtype}"
        assert g.shape == w.shape, f"{what} {name} shape {g.shape} != {w.shape}"
        if not np.array_equal(g, w):
            bad = np.nonzero(g != w)[0][:5]
            raise AssertionError(f"{what} {name} differs at {bad}: {g[bad]} != {w[bad]}")
    if len(got) > 4 and len(want) > 4:
        assert list(got[4]) == list(want[4]), f"{what} record ids differ"


def assert_matches_digest(got, d, what=""):
    kmers, nodes, edges, offsets = got[:4]
    assert (len(kmers), len(nodes), len(edges)) == (d["n_kmers"], d["n_nodes"], d["n_edges"]), \
        f"{what} sizes {(len(kmers), len(nodes), len(edges))} != {(d['n_kmers'], d['n_nodes'], d['n_edges'])}"
    assert offsets.tolist() == d["offsets"], what
    for name, arr in (("kmers", kmers), ("nodes", nodes), ("edges", edges), ("record_offsets", offsets)):
        assert digest(arr) == d[name], f"{what} {name} digest mismatch"
    if len(got) > 4:
        assert hashlib.sha256(repr(got[4]).encode()).hexdigest() == d["ids"], f"{what} ids digest mismatch"


def check_graph_invariants(kmers, nodes, edges, record_offsets):
    """Size-independent properties of a well-formed graph (SURVEY.md 8a ordering invariants)."""
    n = len(kmers)
    if len(nodes):
        h = nodes["hash"]
        assert np.all(h[1:] > h[:-1]), "nodes not strictly ascending by hash"
        assert nodes["start"][0] == 0 and nodes["stop"][-1] == n
        assert np.array_equal(nodes["start"][1:], nodes["stop"][:-1]), "node ranges do not tile kmers"
        assert np.all(nodes["stop"] > nodes["start"])
    else:
        assert n == 0
    if n:
        # inside a node: (record_idx, pos) strictly ascending
        key = kmers["record_idx"].astype(np.uint64) << np.uint64(32) | kmers["pos"].astype(np.uint64)
        brk = np.zeros(n, dtype=bool)
        brk[nodes["start"]] = True
        inc = key[1:] > key[:-1]
        assert np.all(inc | brk[1:]), "kmers not sorted by (record, pos) inside a node"
        assert kmers["record_idx"].max() < record_offsets[-1]
    if len(edges):
        f, s = edges["first"], edges["second"]
        assert np.all(f <= s)
        assert np.all((f[1:] > f[:-1]) | ((f[1:] == f[:-1]) & (s[1:] > s[:-1]))), "edges not sorted / unique"
        assert np.all(edges["weight"] >= 1) and np.all(edges["weight"] <= len(record_offsets) - 1)
        assert np.all(np.isin(f, nodes["hash"])) and np.all(np.isin(s, nodes["hash"]))
