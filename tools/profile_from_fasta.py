#!/usr/bin/env python3
"""Where the time of the drop-in `_build_native(paths, ...)` goes when it starts from FASTA files
(GPU box only): parse + pack, pinned batch assembly, GPU build, export into numpy arrays.

    python tools/profile_from_fasta.py [--genomes 96]
"""
import argparse
import ctypes as C
import os
import shutil
import sys
import tempfile
import time
from pathlib import Path

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
import numpy as np  # noqa: E402

from seqwin_b200 import _lib  # noqa: E402
from seqwin_b200._core import _build_native  # noqa: E402
from seqwin_b200.dist import export_graph  # noqa: E402
from seqwin_b200.synth import SynthSet, SynthSpec, write_fasta  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--genomes", type=int, default=96)
    a = ap.parse_args()
    spec = SynthSpec(n_genomes=a.genomes, n_targets=max(1, a.genomes // 5), genome_len=5_000_000, n_contigs=50, seed=42)
    ss = SynthSet(spec)
    d = Path(tempfile.mkdtemp(prefix="sw_fasta_", dir="/dev/shm"))
    try:
        paths = []
        for g in range(a.genomes):
            p = d / f"g{g:04d}.fasta"
            write_fasta(p, ss.records(g))
            paths.append(str(p))
        n_bases = a.genomes * spec.genome_len
        L = _lib.lib()
        arr = (C.c_char_p * len(paths))(*[p.encode() for p in paths])
        threads = os.cpu_count() or 1
        for rep in range(3):
            t0 = time.perf_counter()
            b = C.c_void_p()
            _lib.check(L.sw_batch_from_fasta(arr, len(paths), threads, C.byref(b)))
            t1 = time.perf_counter()
            g = C.c_void_p()
            _lib.check(L.sw_build_from_batch(b, 21, 200, C.byref(g), None))
            t2 = time.perf_counter()
            arrays = export_graph(L, g)
            t3 = time.perf_counter()
            L.sw_graph_free(g)
            L.sw_batch_free(b)
            t4 = time.perf_counter()
            out = _build_native(paths, 21, 200, threads, False)
            t5 = time.perf_counter()
            print(f"rep {rep}: ingest {1e3 * (t1 - t0):7.1f} ms ({n_bases / (t1 - t0) / 1e9:5.2f} Gbp/s, {threads} threads) | "
                  f"GPU build from pinned batch {1e3 * (t2 - t1):6.1f} | export to numpy {1e3 * (t3 - t2):6.1f} | free {1e3 * (t4 - t3):5.1f} || "
                  f"_build_native one call {1e3 * (t5 - t4):7.1f} ms = {n_bases / (t5 - t4) / 1e9:5.2f} Gbp/s", flush=True)
            del arrays, out
    finally:
        shutil.rmtree(d, ignore_errors=True)


if __name__ == "__main__":
    main()
