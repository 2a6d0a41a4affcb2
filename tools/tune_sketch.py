#!/usr/bin/env python3
"""Sweep sketch-kernel tuning knobs on one resident batch (GPU box only).

    python tools/tune_sketch.py [--genomes 200] "SEQWIN_SPARSE_CPW=12 SEQWIN_SPARSE_SMALL=4" "SEQWIN_SKETCH_DENSE=1" ...

Each argument is a space-separated list of environment settings applied for one measurement
(the library reads them at every build); prints the CUDA-event stage times of sw_dev_build.
"""
import argparse
import ctypes as C
import os
import sys
from pathlib import Path

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
import numpy as np  # noqa: E402

import bench  # noqa: E402
from seqwin_b200 import _lib  # noqa: E402
from seqwin_b200.synth import SynthSet, SynthSpec  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--genomes", type=int, default=200)
    ap.add_argument("--genome-len", type=int, default=5_000_000)
    ap.add_argument("--k", type=int, default=21)
    ap.add_argument("--w", type=int, default=200)
    ap.add_argument("--reps", type=int, default=4)
    ap.add_argument("settings", nargs="*", default=[""])
    a = ap.parse_args()
    spec = SynthSpec(n_genomes=a.genomes, n_targets=max(1, a.genomes // 5), genome_len=a.genome_len, n_contigs=50, seed=42)
    L = _lib.lib()
    batch = bench.build_batch(SynthSet(spec), range(a.genomes), os.cpu_count() or 1)
    if isinstance(batch, tuple):
        batch = batch[0]
    dev = C.c_void_p()
    _lib.check(L.sw_dev_upload(batch, C.byref(dev)))
    knobs = set()
    for s in a.settings:
        knobs.update(kv.split("=")[0] for kv in s.split())
    for s in a.settings:
        for kname in knobs:
            os.environ.pop(kname, None)
        for kv in s.split():
            key, val = kv.split("=")
            os.environ[key] = val
        rows = []
        for _ in range(a.reps + 1):
            g, t = C.c_void_p(), _lib.StageTimes()
            _lib.check(L.sw_dev_build(dev, a.k, a.w, C.byref(g), C.byref(t)))
            L.sw_graph_free(g)
            rows.append((t.sketch_kernel_ms, t.total_ms, t.n_kmers, t.sort_nodes_ms, t.nodes_ms, t.edges_ms))
        rows = rows[1:]
        m = [float(np.mean([r[i] for r in rows])) for i in range(6)]
        print(f"{s or '(default)':50s} sketch_kernel {m[0]:7.3f} ms  total {m[1]:7.3f} ms  sort_nodes {m[3]:6.3f}  nodes {m[4]:6.3f}"
              f"  edges {m[5]:6.3f}  minimizers {rows[0][2]}", flush=True)


if __name__ == "__main__":
    main()
