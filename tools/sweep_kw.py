#!/usr/bin/env python3
"""BASELINE.json configs[2] / configs[4] on one GPU: one resident synthetic batch, a sweep over (k, w).

    python tools/sweep_kw.py [--genomes 2000] [--skew] [--kw 21:200,21:50,21:10,31:50,63:20,127:10]

Prints one JSON line per (k, w): CUDA-event stage times of the scored device-resident build, the
graph sizes, Gbp/s, and the HBM view of the aggregation stage (algorithmic 24*M + 40*U_n + 24*U_e
bytes over sort + nodes + edges time against the measured HBM peak) -- SURVEY.md 8(d)'s second regime.
"""
import argparse
import ctypes as C
import json
import os
import sys
import time
from pathlib import Path

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
import numpy as np  # noqa: E402

import bench  # noqa: E402
from seqwin_b200 import _lib  # noqa: E402
from seqwin_b200.synth import SynthSet, SynthSpec  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--genomes", type=int, default=2000)
    ap.add_argument("--genome-len", type=int, default=5_000_000)
    ap.add_argument("--skew", action="store_true")
    ap.add_argument("--kw", default="21:200,21:50,21:10,31:50,63:20,127:10")
    ap.add_argument("--reps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=1)
    a = ap.parse_args()
    spec = SynthSpec(n_genomes=a.genomes, n_targets=max(1, a.genomes // 5), genome_len=a.genome_len, n_contigs=50,
                     seed=42, skew=a.skew)
    peaks = {}
    pk = ROOT / "MEASURED_PEAKS.json"
    if pk.exists():
        peaks = json.loads(pk.read_text())
    hbm_peak = float(peaks.get("hbm_gbs", 6650.0))
    L = _lib.lib()
    ss = SynthSet(spec)
    t0 = time.perf_counter()
    batch = bench.build_batch(ss, range(a.genomes), os.cpu_count() or 1)
    gen_s = time.perf_counter() - t0
    dev = C.c_void_p()
    _lib.check(L.sw_dev_upload(batch, C.byref(dev)))
    is_t = np.ascontiguousarray(ss.is_targets, dtype=np.bool_)
    n_bases = L.sw_batch_n_bases(batch)
    st = _lib.StageTimes()
    for kw in a.kw.split(","):
        k, w = (int(x) for x in kw.split(":"))
        line = {"genomes": a.genomes, "skew": a.skew, "k": k, "w": w, "n_bases": int(n_bases), "gen_seconds": round(gen_s, 1)}
        try:
            runs = []
            for i in range(a.warmup + a.reps):
                g = C.c_void_p()
                _lib.check(L.sw_dev_build_scored(dev, k, w, is_t.ctypes.data, len(is_t), C.byref(g), C.byref(st)))
                L.sw_graph_free(g)
                if i >= a.warmup:
                    runs.append(st.as_dict())
            m = {n: float(np.mean([r[n] for r in runs])) for n in
                 ("total_ms", "plan_ms", "sketch_kernel_ms", "reorder_ms", "sketch_ms", "sort_nodes_ms", "nodes_ms", "edges_ms")}
            M, Un, Ue = (int(runs[-1][n]) for n in ("n_kmers", "n_nodes", "n_edges"))
            agg_ms = m["sort_nodes_ms"] + m["nodes_ms"] + m["edges_ms"]
            agg_bytes = 24 * M + 40 * Un + 24 * Ue
            path_bytes = n_bases / 4 + 40 * M + 40 * Un + 24 * Ue
            line.update({"gbp_s": n_bases / (m["total_ms"] * 1e-3) / 1e9, "stage_ms": m,
                         "n_kmers": M, "n_nodes": Un, "n_edges": Ue,
                         "aggregation": {"ms": agg_ms, "algorithmic_bytes": agg_bytes,
                                         "achieved_gbs": agg_bytes / (agg_ms * 1e-3) / 1e9,
                                         "hbm_frac": agg_bytes / (agg_ms * 1e-3) / 1e9 / hbm_peak},
                         "path": {"algorithmic_bytes": path_bytes, "achieved_gbs": path_bytes / (m["total_ms"] * 1e-3) / 1e9,
                                  "hbm_frac": path_bytes / (m["total_ms"] * 1e-3) / 1e9 / hbm_peak},
                         "hbm_peak_gbs": hbm_peak})
            ms = (C.c_uint64 * 4)()
            _lib.check(L.sw_mem_stats(ms))
            line["device_memory_gb"] = {"scratch_arena": ms[0] / 1e9, "pool_used_high": ms[1] / 1e9, "in_use_now": ms[3] / 1e9}
        except Exception as exc:  # noqa: BLE001 - a config that does not fit is a result, not a crash
            line["error"] = f"{type(exc).__name__}: {exc}"
        print(json.dumps(line), flush=True)
    L.sw_dev_batch_free(dev)
    L.sw_batch_free(batch)


if __name__ == "__main__":
    main()
