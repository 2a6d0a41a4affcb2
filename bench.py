#!/usr/bin/env python3
"""bench.py -- Gbp/s sketched + graph-built on synthetic bacterial-genome sets (BASELINE.json).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--genomes G] ...

One "step" = one pass of the hot path (ntHash -> window minimizers -> minimizer graph) over the
whole synthetic batch.  Workload at N=1: BASELINE.json configs[1] -- 500 genomes x 5 Mbp (100
targets / 400 non-targets), k=21, w=200.  Under torchrun (N>1) every rank owns one such shard
(weak scaling) and minimizer / edge records are exchanged with NCCL all-to-all (seqwin_b200.dist).

  value      device-resident: packed input already in HBM, graph (build + get_penalty scoring) left in
             HBM, CUDA-event timed
  e2e        the same through the host-buffer C-ABI call (pinned 2-bit batch -> H2D -> build ->
             D2H of kmers/nodes/edges), wall-clock around the call
  roofline   dominant kernel (sketch): algorithmic bytes / CUDA-event kernel time vs measured HBM
             peak, plus the INT32-pipe view that actually bounds it
  cpu_baseline / --impl reference: the UNMODIFIED reference extension (oracle/_ref) on the box's
             host cores over a bounded sample of the same genomes (FASTA on /dev/shm)
"""
from __future__ import annotations

import argparse
import ctypes as C
import json
import os
import shutil
import subprocess
import sys
import tempfile
import threading
import time
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent
sys.path.insert(0, str(ROOT))

from seqwin_b200.synth import SynthSet, SynthSpec, write_fasta  # noqa: E402

K_DEFAULT, W_DEFAULT = 21, 200


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--genomes", type=int, default=500, help="genomes per GPU (weak scaling)")
    ap.add_argument("--genome-len", type=int, default=5_000_000)
    ap.add_argument("--k", type=int, default=K_DEFAULT)
    ap.add_argument("--w", type=int, default=W_DEFAULT)
    ap.add_argument("--sample-genomes", type=int, default=0, help="genomes in the CPU-baseline sample (0 = auto)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--skew", action="store_true", help="config C5: near-clonal + repeats")
    return ap.parse_args()


def spec_for(args, world: int) -> SynthSpec:
    n = args.genomes * world
    return SynthSpec(n_genomes=n, n_targets=max(1, n // 5), genome_len=args.genome_len, n_contigs=50,
                     seed=42, skew=args.skew)


# ---- clocks ----------------------------------------------------------------------------------
class ClockSampler:
    """SM clock / throttle-reason sampling during the timed region (NVML, 5 ms period; the
    nvidia-smi line of B200_PROFILING.md polls the same counters but cannot start fast enough
    for a sub-second region)."""

    def __init__(self, gpu_index: int):
        self.gpu = gpu_index
        self.samples: list[tuple[int, int, int]] = []
        self._stop = threading.Event()
        self._thread = None
        self._h = None
        try:
            import pynvml
            pynvml.nvmlInit()
            vis = os.environ.get("CUDA_VISIBLE_DEVICES")
            phys = int(vis.split(",")[gpu_index]) if vis and vis.split(",")[gpu_index].isdigit() else gpu_index
            self._h = pynvml.nvmlDeviceGetHandleByIndex(phys)
            self._nv = pynvml
        except Exception:
            self._h = None

    def _run(self):
        nv = self._nv
        while not self._stop.is_set():
            try:
                sm = nv.nvmlDeviceGetClockInfo(self._h, nv.NVML_CLOCK_SM)
                mx = nv.nvmlDeviceGetMaxClockInfo(self._h, nv.NVML_CLOCK_SM)
                rs = nv.nvmlDeviceGetCurrentClocksEventReasons(self._h) if hasattr(
                    nv, "nvmlDeviceGetCurrentClocksEventReasons") else nv.nvmlDeviceGetCurrentClocksThrottleReasons(self._h)
                self.samples.append((sm, mx, rs))
            except Exception:
                pass
            time.sleep(0.005)

    def start(self):
        if self._h is None:
            return
        self._thread = threading.Thread(target=self._run, daemon=True)
        self._thread.start()

    def stop(self) -> dict:
        if self._h is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvml unavailable"]}
        self._stop.set()
        self._thread.join()
        nv = self._nv
        names = {"hw_slowdown": getattr(nv, "nvmlClocksThrottleReasonHwSlowdown", 0x8),
                 "hw_thermal_slowdown": getattr(nv, "nvmlClocksThrottleReasonHwThermalSlowdown", 0x40),
                 "sw_thermal_slowdown": getattr(nv, "nvmlClocksThrottleReasonSwThermalSlowdown", 0x20),
                 "sw_power_cap": getattr(nv, "nvmlClocksThrottleReasonSwPowerCap", 0x4)}
        reasons = sorted(n for n, bit in names.items() if any(r & bit for _, _, r in self.samples))
        sm = [s for s, _, _ in self.samples]
        return {"sm_mhz": float(np.median(sm)) if sm else None,
                "sm_max_mhz": float(max(m for _, m, _ in self.samples)) if sm else None,
                "reasons": reasons, "samples": len(sm)}


# ---- reference / CPU baseline ------------------------------------------------------------------
def sample_indices(spec: SynthSpec, n_sample: int) -> list[int]:
    n_sample = max(2, min(n_sample, spec.n_genomes))
    n_t = max(1, min(spec.n_targets, n_sample // 5))
    n_n = min(spec.n_genomes - spec.n_targets, n_sample - n_t)
    return list(range(n_t)) + list(range(spec.n_targets, spec.n_targets + n_n))


def write_sample(ss: SynthSet, idx: list[int]) -> tuple[Path, list[Path], np.ndarray, int]:
    base = Path("/dev/shm") if Path("/dev/shm").is_dir() else Path(tempfile.gettempdir())
    d = Path(tempfile.mkdtemp(prefix="seqwin_b200_sample_", dir=base))
    paths, n_bases = [], 0
    for g in idx:
        p = d / f"g{g:05d}.fasta"
        recs = ss.records(g)
        n_bases += sum(len(s) for _, s in recs)
        write_fasta(p, recs)
        paths.append(p)
    is_t = np.array([g < ss.spec.n_targets for g in idx])
    return d, paths, is_t, n_bases


def time_reference(paths, is_t, k, w, n_bases, steps, warmup):
    """The reference's own CPU path: _build_native + _get_penalty_native, --threads = all cores."""
    from oracle import oracle as O
    ref = O.load_reference()
    kind = "reference"
    cores = os.cpu_count() or 1
    if ref is None:  # oracle/_ref not present: fall back to the scalar C port of the oracle
        ref, kind, cores = O, "port", 1
    spaths = [str(p) for p in paths]
    times = []
    for i in range(warmup + steps):
        t0 = time.perf_counter()
        kmers, nodes, edges, offsets, ids = ref._build_native(spaths, k, w, cores, False)
        ref._get_penalty_native(kmers, nodes, offsets, np.asarray(is_t, dtype=np.bool_), cores)
        dt = time.perf_counter() - t0
        if i >= warmup:
            times.append(dt)
    mean = float(np.mean(times))
    return {"value": n_bases / mean / 1e9, "unit": "Gbp/s", "cores": cores, "kind": kind,
            "sample": f"{len(paths)} of the workload's genomes ({n_bases / 1e6:.0f} Mbp), plain FASTA on tmpfs, "
                      f"build+get_penalty, mean of {steps} (warmup {warmup})",
            "seconds": mean, "graph": (kmers, nodes, edges, offsets)}


# ---- our arm -----------------------------------------------------------------------------------
def build_batch(ss: SynthSet, genomes: range, threads: int):
    """Generate genomes and pack them into one pinned 2-bit batch through the C ABI."""
    from seqwin_b200 import _lib
    L = _lib.lib()
    arrs, lens, asm_of, ids = [], [], [], []
    for a, g in enumerate(genomes):
        for rid, seq in ss.records(g):
            arrs.append(np.ascontiguousarray(seq))
            lens.append(len(seq))
            asm_of.append(a)
            ids.append(rid.encode())
    n = len(arrs)
    ptrs = (C.c_void_p * max(1, n))(*[a.ctypes.data for a in arrs])
    idp = (C.c_char_p * max(1, n))(*ids)
    lens_a = np.asarray(lens, dtype=np.uint32)
    asm_a = np.asarray(asm_of, dtype=np.uint32)
    b = C.c_void_p()
    _lib.check(L.sw_batch_from_memory(ptrs, lens_a.ctypes.data, asm_a.ctypes.data, idp, n, len(genomes), threads,
                                      C.byref(b)))
    return b


def main():
    args = parse_args()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    k, w = args.k, args.w
    spec = spec_for(args, world)
    workload = (f"synthetic {spec.n_genomes} genomes x {spec.genome_len / 1e6:g} Mbp "
                f"({spec.n_targets} targets / {spec.n_genomes - spec.n_targets} non-targets), 50 contigs each, "
                f"k={k} w={w}" + (", skew (C5)" if args.skew else ""))

    if args.impl == "reference":
        if rank != 0:
            return
        ss = SynthSet(spec)
        n_s = args.sample_genomes or int(np.clip(os.cpu_count() or 8, 32, 96))
        idx = sample_indices(spec, n_s)
        d, paths, is_t, n_bases = write_sample(ss, idx)
        try:
            r = time_reference(paths, is_t, k, w, n_bases, args.steps, args.warmup)
        finally:
            shutil.rmtree(d, ignore_errors=True)
        r.pop("graph")
        line = {"impl": "reference", "metric": "Gbp/s sketched+graph-built", "value": r["value"], "unit": "Gbp/s",
                "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": r["seconds"] * 1e3,
                "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u64", "data": "synthetic",
                "config": {"workload": workload},
                "cpu_baseline": {k2: r[k2] for k2 in ("value", "unit", "cores", "kind", "sample")},
                "e2e": {"value": r["value"], "unit": "Gbp/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
        print(json.dumps(line))
        return

    import torch
    torch.cuda.set_device(local_rank)
    from seqwin_b200 import _lib
    from seqwin_b200._lib import StageTimes
    L = _lib.lib()
    dist = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))

    threads = max(1, (os.cpu_count() or 8) // max(1, world))
    ss = SynthSet(spec)
    my_genomes = range(rank * args.genomes, (rank + 1) * args.genomes)
    t0 = time.perf_counter()
    batch = build_batch(ss, my_genomes, threads)
    gen_s = time.perf_counter() - t0
    n_bases_local = L.sw_batch_n_bases(batch)
    packed_bytes = L.sw_batch_packed_bytes(batch)
    n_records = L.sw_batch_n_records(batch)

    if world > 1:
        from seqwin_b200 import dist as swdist
        result = swdist.bench_loop(L, batch, spec, rank, world, k, w, args.steps, args.warmup, ClockSampler)
    else:
        dev = C.c_void_p()
        _lib.check(L.sw_dev_upload(batch, C.byref(dev)))
        st = StageTimes()
        is_t = np.ascontiguousarray(ss.is_targets[my_genomes.start:my_genomes.stop], dtype=np.bool_)
        def dev_step():
            # build + scoring, like the reference arm (_build_native + _get_penalty_native); the classes are
            # known up front, so the scoring is fused into the node stage (sw_dev_build_scored)
            g = C.c_void_p()
            _lib.check(L.sw_dev_build_scored(dev, k, w, is_t.ctypes.data, len(is_t), C.byref(g), C.byref(st)))
            L.sw_graph_free(g)
            return st.as_dict()

        for _ in range(args.warmup):
            dev_step()
        sampler = ClockSampler(local_rank)
        sampler.start()
        torch.cuda.synchronize()
        stages = [dev_step() for _ in range(args.steps)]
        torch.cuda.synchronize()
        clocks = sampler.stop()
        L.sw_dev_batch_free(dev)

        # end to end through the host-buffer call: pinned packed batch -> graph arrays on the host
        def e2e_step():
            g = C.c_void_p()
            t0 = time.perf_counter()
            _lib.check(L.sw_build_from_batch_scored(batch, k, w, is_t.ctypes.data, len(is_t), C.byref(g), C.byref(st)))
            dt = time.perf_counter() - t0
            L.sw_graph_free(g)
            return dt, st.as_dict()

        for _ in range(max(1, args.warmup // 2)):
            e2e_step()
        e2e_runs = [e2e_step() for _ in range(args.steps)]
        result = {"stages": stages, "clocks": clocks, "e2e_runs": e2e_runs}

    if rank != 0:
        if dist is not None:
            dist.destroy_process_group()
        return

    stages = result["stages"]
    total_ms = float(np.mean([s["total_ms"] for s in stages]))
    sketch_ms = float(np.mean([s["sketch_kernel_ms"] for s in stages]))
    s0 = stages[-1]
    n_bases_total = result.get("n_bases_total", n_bases_local)
    M, Un, Ue = s0["n_kmers"], s0["n_nodes"], s0["n_edges"]
    value = n_bases_total / (total_ms * 1e-3) / 1e9

    peaks = {}
    pk = ROOT / "MEASURED_PEAKS.json"
    if pk.exists():
        peaks = json.loads(pk.read_text())
    hbm_peak = float(peaks.get("hbm_gbs", 6650.0))
    peak_src = "measured (MEASURED_PEAKS.json)" if peaks else "fallback (B200_PROFILING.md)"
    # dominant kernel = sketch: reads the 2-bit input once, writes 16 B per minimizer
    sketch_bytes = n_bases_local / 4 + 16 * s0.get("n_kmers_local", M)
    ach = sketch_bytes / (sketch_ms * 1e-3) / 1e9
    clocks = result["clocks"]
    sm_mhz = clocks.get("sm_mhz") or float(peaks.get("sm_max_mhz", 1965.0))
    int_peak = 148 * 128 * sm_mhz * 1e6  # lane-ops/s at the clock seen under load
    int_alg = 45.0 * n_bases_local + 12.0 * s0.get("n_kmers_local", M)
    # DRAM traffic of the same kernel from the committed `ncu --set full` capture, scaled by the
    # algorithmic bytes when the workload differs from the captured one
    traffic, kernel_name, alu_pct = None, "sketch_sparse_kernel<128,64,20>", 73
    tp = ROOT / "profiles" / "r1_sketch_traffic.json"
    if tp.exists():
        tj = json.loads(tp.read_text())
        traffic = (tj["dram_bytes_read"] + tj["dram_bytes_write"]) * sketch_bytes / tj["algorithmic_bytes"]
        kernel_name, alu_pct = tj.get("kernel", kernel_name), tj.get("alu_pipe_pct", alu_pct)
    if k != K_DEFAULT or w != W_DEFAULT or os.environ.get("SEQWIN_SKETCH_DENSE"):
        kernel_name = "sketch kernel selected for this (k, w)"
    roofline = {"bound": "hbm", "kernel": kernel_name, "achieved": ach, "peak": hbm_peak, "unit": "GB/s",
                "frac": ach / hbm_peak, "traffic": traffic, "peak_source": peak_src,
                "algorithmic_bytes_per_launch": sketch_bytes, "kernel_ms": sketch_ms,
                "note": f"the sketch kernel is bound by the INT32 ALU pipe, not HBM (SURVEY 8d; ncu: ALU pipe {alu_pct:g} % busy, "
                        "DRAM < 3 %): see int_roofline and path",
                "int_roofline": {"achieved_Tops": int_alg / (sketch_ms * 1e-3) / 1e12, "peak_Tops": int_peak / 1e12,
                                 "frac": int_alg / (sketch_ms * 1e-3) / int_peak,
                                 "model": "I_alg = 45*N + 12*M lane-ops; peak = 148 SM x 128 lanes x sm clock under load"},
                "path": {"algorithmic_bytes": n_bases_local / 4 + 40 * M + 40 * Un + 24 * Ue, "ms": total_ms,
                         "achieved": (n_bases_local / 4 + 40 * M + 40 * Un + 24 * Ue) / (total_ms * 1e-3) / 1e9,
                         "dist_ms": {n: float(np.mean([s[n] for s in stages])) for n in
                                     ("phase_local_ms", "phase_exchange_merge_ms", "phase_merge_ms") if n in stages[0]},
                         "stage_ms": {n: float(np.mean([s[n] for s in stages]))
                                      for n in ("plan_ms", "sketch_kernel_ms", "reorder_ms", "sketch_ms", "sort_nodes_ms",
                                                "nodes_ms", "edges_ms", "penalty_ms") if n in stages[0]}}}

    e2e_runs = result["e2e_runs"]
    e2e_s = float(np.mean([r[0] for r in e2e_runs]))
    d2h = 8 * M + 40 * Un + 24 * Ue
    e2e = {"value": n_bases_total / e2e_s / 1e9, "unit": "Gbp/s",
           "h2d_bytes_per_step": int(packed_bytes + 12 * n_records) * world, "d2h_bytes_per_step": int(d2h),
           "ms_per_step": e2e_s * 1e3,
           "stage_ms": {n: float(np.mean([r[1][n] for r in e2e_runs])) for n in ("h2d_ms", "plan_ms", "sketch_kernel_ms", "reorder_ms", "sort_nodes_ms", "nodes_ms",
                                  "penalty_ms", "edges_ms", "total_ms", "d2h_ms")},
           "what": "sw_build_from_batch_scored: pinned 2-bit host batch -> H2D -> sketch + graph + get_penalty -> "
                   "D2H host arrays (copies overlapped with the kernels)"}

    cpu_baseline = None
    parity_sample = None
    if not args.no_cpu_baseline and world == 1:
        n_s = args.sample_genomes or int(np.clip(os.cpu_count() or 8, 32, 96))
        idx = sample_indices(spec, n_s)
        d, paths, is_t, nb = write_sample(ss, idx)
        try:
            r = time_reference(paths, is_t, k, w, nb, steps=1, warmup=1)
            # the same sample through our FASTA entry point: bit-exact check + FASTA-inclusive e2e
            from seqwin_b200.graph import KmerGraph, _get_penalty
            for _ in range(2):   # like the reference arm: one warm-up pass, one timed pass
                t0 = time.perf_counter()
                g = KmerGraph(paths, k, w, n_cpu=os.cpu_count() or 8)
                _get_penalty(g.kmers, g.nodes, g.record_offsets, is_t)   # in place, like the reference arm
                ours_s = time.perf_counter() - t0
            rk, rn, re_, ro = r.pop("graph")
            parity_sample = bool(np.array_equal(g.kmers, rk) and np.array_equal(g.nodes, rn)
                                 and np.array_equal(g.edges, re_) and np.array_equal(g.record_offsets, ro))
            cpu_baseline = {k2: r[k2] for k2 in ("value", "unit", "cores", "kind", "sample")}
            cpu_baseline["ours_same_sample_from_fasta_gbps"] = nb / ours_s / 1e9
        finally:
            shutil.rmtree(d, ignore_errors=True)

    line = {"metric": "Gbp/s sketched+graph-built", "value": value, "unit": "Gbp/s", "n_gpus": world,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": total_ms, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "u64", "data": "synthetic",
            "config": {"workload": workload, "l2": "inputs larger than L2 (packed batch %.0f MB > 126 MB)" % (packed_bytes / 1e6),
                       "timing": "CUDA events on the library stream around each step (host tile planning included)",
                       "genomes_per_gpu": args.genomes, "gen_seconds": round(gen_s, 1)},
            "graph": {"n_bases": int(n_bases_total), "n_kmers": int(M), "n_nodes": int(Un), "n_edges": int(Ue)},
            "clocks": clocks, "e2e": e2e, "gpu_launches": int(sum(s["total_launches"] for s in stages)),
            "roofline": roofline, "cpu_baseline": cpu_baseline, "parity_sample_bit_exact": parity_sample}
    print(json.dumps(line))
    L.sw_batch_free(batch)
    if dist is not None:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
