#!/usr/bin/env python3
"""DRAM traffic of the hot-path kernels from `ncu --set full` captures -> profiles/r2_traffic.json (read by bench.py).

    python tools/ncu_traffic.py --n-bases N --n-kmers M --n-nodes U --n-edges E  report1.ncu-rep [report2.ncu-rep ...]

Every captured launch of a repo kernel is listed with its dram__bytes_read.sum + dram__bytes_write.sum, duration
and the pipe / issue utilisation; the `sketch` entry sums the sketch kernels, the `aggregation` entry everything
between the ordered stream and the graph arrays (one build each: if a kernel was captured in several builds, the
per-launch mean times the launches of one build is used).  Algorithmic bytes are SURVEY.md 8(d)'s:
sketch N/4 + 16 M, aggregation 24 M + 40 U_n + 24 U_e.
"""
import argparse
import csv
import json
import re
import subprocess
from collections import OrderedDict

SKETCH = ("sketch_sparse_kernel", "sketch_fast_kernel", "sketch_generic_kernel")
AGG = ("radix_onesweep_kernel", "radix_hist_kernel", "group_count_kernel", "group_place_kernel", "bucket_edges_kernel",
       "bucket_edges_out_kernel", "bucket_search_kernel", "reorder_kernel")


def short(name: str) -> str:
    name = re.sub(r"\(.*", "", name)
    return name.replace("void ", "").replace("unnamed>::", "").replace("agg::", "").replace("sw::", "").strip()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("reports", nargs="+")
    ap.add_argument("--n-bases", type=float, required=True)
    ap.add_argument("--n-kmers", type=float, required=True)
    ap.add_argument("--n-nodes", type=float, required=True)
    ap.add_argument("--n-edges", type=float, required=True)
    ap.add_argument("--launches-per-build", default="",
                    help="kernels launched more than once per build, name=count[,name=count]")
    ap.add_argument("--out", default="profiles/r2_traffic.json")
    a = ap.parse_args()
    per_build = {k: int(v) for k, v in (kv.split("=") for kv in a.launches_per_build.split(",") if kv)}
    launches = []
    for rep in a.reports:
        out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
        rows = list(csv.reader(out.splitlines()))
        hdr = rows[0]
        ix = {h: i for i, h in enumerate(hdr)}
        units = dict(zip(rows[0], rows[1]))   # the unit row: every column has its own (Mbyte here, Gbyte there)
        scale = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "ns": 1e-6, "us": 1e-3, "ms": 1.0, "s": 1e3}

        def val(r, key):
            try:
                return float(r[ix[key]]) * scale.get(units.get(key, ""), 1.0)
            except (KeyError, ValueError):
                return None
        for r in rows[2:]:
            launches.append({
                "kernel": short(r[ix["Kernel Name"]]), "report": rep.split("/")[-1],
                "duration_ms_under_ncu": val(r, "gpu__time_duration.sum"),
                "dram_bytes_read": val(r, "dram__bytes_read.sum"), "dram_bytes_write": val(r, "dram__bytes_write.sum"),
                "dram_throughput_pct": val(r, "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed"),
                "alu_pipe_busy_pct": val(r, "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active"),
                "issue_active_pct": val(r, "smsp__issue_active.avg.pct_of_peak_sustained_active"),
                "warp_instructions": val(r, "smsp__inst_executed.sum"),
                "registers": val(r, "launch__registers_per_thread"),
                "warps_active_pct": val(r, "sm__warps_active.avg.pct_of_peak_sustained_active"),
                "lanes_per_instruction": val(r, "smsp__thread_inst_executed_per_inst_executed.ratio"),
            })
    # one build = the launches from a first-pass sketch kernel up to the edge kernel; only the LAST complete build of the
    # capture is summed (a capture window may start or end in the middle of one)
    builds, cur = [], []
    for l in launches:
        if l["kernel"].startswith(SKETCH) and cur and any(x["kernel"].startswith("bucket_edges_kernel") for x in cur):
            builds.append(cur)
            cur = []
        cur.append(l)
    if cur:
        builds.append(cur)
    complete = [b for b in builds if any(x["kernel"].startswith("reorder_kernel") for x in b)
                and any(x["kernel"].startswith("bucket_edges_kernel") for x in b)
                and sum(x["kernel"].startswith("sketch_sparse_kernel") for x in b) >= 2]
    build = complete[-1] if complete else launches

    def group(names):
        by = OrderedDict()
        for l in build:
            if any(l["kernel"].startswith(n) for n in names):
                by.setdefault(l["kernel"], []).append(l)
        tot_b = tot_ms = 0.0
        parts = {}
        for k, ls in by.items():
            b = sum((x["dram_bytes_read"] or 0) + (x["dram_bytes_write"] or 0) for x in ls)
            ms = sum(x["duration_ms_under_ncu"] or 0 for x in ls)
            parts[k] = {"launches_per_build": len(ls), "dram_bytes": b, "ms_under_ncu": ms}
            tot_b += b
            tot_ms += ms
        return tot_b, tot_ms, parts
    N, M, U, E = a.n_bases, a.n_kmers, a.n_nodes, a.n_edges
    sk_b, sk_ms, sk_parts = group(SKETCH)
    ag_b, ag_ms, ag_parts = group(AGG)
    sparse = max((l for l in build if l["kernel"].startswith("sketch_sparse_kernel")), key=lambda l: l["duration_ms_under_ncu"] or 0,
                 default=None)
    res = {
        "capture": "ncu --set full --clock-control none over tools/sweep_kw.py --genomes 500 --kw 21:200 (the bench.py N=1 workload)",
        "workload": {"n_bases": N, "n_kmers": M, "n_nodes": U, "n_edges": E},
        "sketch": {"dram_bytes": sk_b, "algorithmic_bytes": N / 4 + 16 * M, "ms_under_ncu": sk_ms, "kernels": sk_parts,
                   "alu_pipe_busy_pct": sparse and sparse["alu_pipe_busy_pct"], "issue_active_pct": sparse and sparse["issue_active_pct"],
                   "warp_instructions": sparse and sparse["warp_instructions"]},
        "aggregation": {"dram_bytes": ag_b, "algorithmic_bytes": 24 * M + 40 * U + 24 * E, "ms_under_ncu": ag_ms,
                        "kernels": ag_parts,
                        "note": "the main kernels of the stage (the scans, bound searches and count kernels move < 1 % of it)"},
        "launches": launches,
    }
    with open(a.out, "w") as f:
        json.dump(res, f, indent=1)
    print(json.dumps({k: (v if not isinstance(v, dict) else {kk: vv for kk, vv in v.items() if kk != "kernels"})
                      for k, v in res.items() if k != "launches"}, indent=1))


if __name__ == "__main__":
    main()
