#!/usr/bin/env python3
"""Time the graph consumers on the N=1 bench workload: GPU vs the numpy / C
restatements of the reference on the same arrays (GPU box only).

    python tools/bench_filter.py [--genomes 500]
"""
import argparse
import ctypes as C
import os
import sys
import time
from pathlib import Path

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
import numpy as np  # noqa: E402
import torch  # noqa: E402

import bench  # noqa: E402
from oracle import oracle as O  # noqa: E402
from seqwin_b200 import _lib  # noqa: E402
from seqwin_b200.dist import export_graph  # noqa: E402
from seqwin_b200.graph import _filter_edges_and_nodes  # noqa: E402
from seqwin_b200.synth import SynthSet, SynthSpec  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--genomes", type=int, default=500)
    ap.add_argument("--th", type=int, default=50)
    a = ap.parse_args()
    spec = SynthSpec(n_genomes=a.genomes, n_targets=max(1, a.genomes // 5), genome_len=5_000_000, n_contigs=50, seed=42)
    L = _lib.lib()
    batch = bench.build_batch(SynthSet(spec), range(a.genomes), os.cpu_count() or 1)
    is_t = np.ascontiguousarray(np.arange(a.genomes) < spec.n_targets, dtype=np.bool_)
    dev = C.c_void_p()
    _lib.check(L.sw_dev_upload(batch, C.byref(dev)))

    def build():
        g = C.c_void_p()
        _lib.check(L.sw_dev_build_scored(dev, 21, 200, is_t.ctypes.data, len(is_t), C.byref(g), None))
        return g

    g = build()
    kmers, nodes, edges = export_graph(L, g)
    L.sw_graph_free(g)
    print(f"graph: {len(kmers)} kmers, {len(nodes)} nodes, {len(edges)} edges; threshold weight > {a.th}")

    t0 = time.perf_counter()
    n_ref, e_ref = O.filter_edges_and_nodes(nodes, edges, float(a.th))
    t_np = time.perf_counter() - t0
    for _ in range(2):
        t0 = time.perf_counter()
        n_gpu, e_gpu = _filter_edges_and_nodes(nodes, edges, float(a.th))
        t_host = time.perf_counter() - t0
    assert np.array_equal(n_gpu, n_ref) and np.array_equal(e_gpu, e_ref)
    times = []
    for _ in range(3):
        g = build()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        _lib.check(L.sw_graph_filter_edges(g, C.c_uint64(a.th)))
        times.append(time.perf_counter() - t0)
        L.sw_graph_free(g)
    print(f"filter_edges_and_nodes: numpy restatement {t_np * 1e3:.1f} ms | GPU from/to host arrays {t_host * 1e3:.1f} ms | "
          f"device-resident {min(times) * 1e3:.2f} ms  -> {len(n_ref)} nodes, {len(e_ref)} edges")

    rng = np.random.default_rng(0)
    used = np.ascontiguousarray(rng.choice(n_ref["hash"], size=max(1, len(n_ref) // 10), replace=False))
    used_set = frozenset(used.tolist())
    t0 = time.perf_counter()
    k_ref, n2_ref = O._filter_kmers_native(kmers, n_ref, used_set)
    t_c = time.perf_counter() - t0
    times = []
    for _ in range(3):
        g = build()
        _lib.check(L.sw_graph_filter_edges(g, C.c_uint64(a.th)))
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        _lib.check(L.sw_graph_filter_kmers(g, used.ctypes.data, len(used)))
        times.append(time.perf_counter() - t0)
        k_gpu, n_gpu2, _ = export_graph(L, g)
        L.sw_graph_free(g)
    assert np.array_equal(k_gpu, k_ref) and np.array_equal(n_gpu2, n2_ref)
    print(f"filter_kmers ({len(used)} hashes): C restatement on host arrays {t_c * 1e3:.1f} ms | device-resident "
          f"{min(times) * 1e3:.2f} ms  -> {len(k_ref)} kmers, {len(n2_ref)} nodes")
    L.sw_dev_batch_free(dev)


if __name__ == "__main__":
    main()
