#!/usr/bin/env python3
"""bench.py -- Gbp/s sketched + graph-built on synthetic bacterial-genome sets (BASELINE.json).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--genomes G] ...

One "step" = one pass of the hot path (ntHash -> window minimizers -> minimizer graph + get_penalty
scoring) over the whole synthetic batch.  Workloads (genomes x 5 Mbp, 1 in 5 a target, k=21, w=200):

  N = 1   BASELINE.json configs[1]: 500 genomes (2.5 Gbp) on one B200
  N >= 2  1,875 genomes per GPU, so that N = 8 is configs[3]: 15,000 genomes (75 Gbp) on 8 B200 with the
          nodes sharded by hash range over an NCCL all-to-all (seqwin_b200.dist)

  value      device-resident: packed input already in HBM, graph (build + scoring) left in HBM, CUDA events
  e2e        the same through the host-buffer C-ABI call (pinned 2-bit batch -> H2D -> build -> D2H of
             kmers / nodes / edges), wall clock around the call
  roofline   the sketch kernel against the measured INT32-pipe peak (its binding roofline) and the
             measured HBM peak; the aggregation stage against the HBM peak
  parity     N = 1: the whole workload is written as FASTA and built by the UNMODIFIED reference
             (oracle/_ref) -- every array compared bit for bit.  N > 1: a subset (125 genomes per rank)
             goes through the same NCCL exchange + merge and is compared with the reference, piece by
             piece and by SHA-256; the full-size step is checked through multiset checksums
  cpu_baseline / --impl reference: the reference's own CPU path on the box's cores
"""
from __future__ import annotations

import argparse
import ctypes as C
import hashlib
import json
import os
import shutil
import sys
import tempfile
import threading
import time
from concurrent.futures import ThreadPoolExecutor
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent
sys.path.insert(0, str(ROOT))

from seqwin_b200.synth import SynthSet, SynthSpec, write_fasta  # noqa: E402

K_DEFAULT, W_DEFAULT = 21, 200
GENOMES_1GPU, GENOMES_PER_GPU_MULTI = 500, 1875
PARITY_GENOMES_PER_RANK = 125      # N > 1: the first genomes of every shard form the parity subset
REF_SAMPLE_GENOMES_MULTI = 500     # N > 1 reference arm: a 2.5 Gbp sample per step


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--genomes", type=int, default=0, help="genomes per GPU (0 = 500 at N=1, 1875 at N>1)")
    ap.add_argument("--genome-len", type=int, default=5_000_000)
    ap.add_argument("--k", type=int, default=K_DEFAULT)
    ap.add_argument("--w", type=int, default=W_DEFAULT)
    ap.add_argument("--sample-genomes", type=int, default=0, help="genomes in the CPU-baseline run (0 = auto)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-parity", action="store_true")
    ap.add_argument("--skew", action="store_true", help="config C5: near-clonal + repeats")
    return ap.parse_args()


def genomes_per_gpu(args, world: int) -> int:
    return args.genomes or (GENOMES_1GPU if world == 1 else GENOMES_PER_GPU_MULTI)


def spec_for(args, world: int) -> SynthSpec:
    n = genomes_per_gpu(args, world) * world
    return SynthSpec(n_genomes=n, n_targets=max(1, n // 5), genome_len=args.genome_len, n_contigs=50,
                     seed=42, skew=args.skew)


def workload_name(spec: SynthSpec, k: int, w: int) -> str:
    return (f"synthetic {spec.n_genomes} genomes x {spec.genome_len / 1e6:g} Mbp "
            f"({spec.n_targets} targets / {spec.n_genomes - spec.n_targets} non-targets), 50 contigs each, "
            f"k={k} w={w}" + (", skew (C5)" if spec.skew else ""))


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


# ---- FASTA for the reference --------------------------------------------------------------------
def shm_dir(prefix: str) -> Path:
    base = Path("/dev/shm") if Path("/dev/shm").is_dir() else Path(tempfile.gettempdir())
    return Path(tempfile.mkdtemp(prefix=prefix, dir=base))


def fasta_path(d: Path, g: int) -> Path:
    return d / f"g{g:05d}.fasta"


def write_genomes(ss: SynthSet, genomes, d: Path, threads: int) -> int:
    """80-column plain FASTA of the given genomes into d (numpy releases the GIL: a few threads help)."""
    def one(g):
        recs = ss.records(g)
        write_fasta(fasta_path(d, g), recs)
        return sum(len(s) for _, s in recs)
    with ThreadPoolExecutor(max(1, threads)) as ex:
        return int(sum(ex.map(one, list(genomes))))


def sample_indices(spec: SynthSpec, n_sample: int) -> list[int]:
    """n_sample genomes of the workload with its 1:4 class mix (targets first)."""
    n_sample = max(2, min(n_sample, spec.n_genomes))
    if n_sample == spec.n_genomes:
        return list(range(spec.n_genomes))
    n_t = max(1, min(spec.n_targets, n_sample // 5))
    n_n = min(spec.n_genomes - spec.n_targets, n_sample - n_t)
    return list(range(n_t)) + list(range(spec.n_targets, spec.n_targets + n_n))


def run_reference(paths, is_t, k, w, steps, warmup, keep_graph=False):
    """The reference's own CPU path: _build_native + _get_penalty_native, --threads = all cores."""
    from oracle import oracle as O
    ref = O.load_reference()
    kind = "reference"
    cores = os.cpu_count() or 1
    if ref is None:  # oracle/_ref not present: the scalar C port of the oracle
        ref, kind, cores = O, "port", 1
    spaths = [str(p) for p in paths]
    is_t = np.ascontiguousarray(is_t, dtype=np.bool_)
    times, graph = [], None
    for i in range(warmup + steps):
        graph = None
        t0 = time.perf_counter()
        kmers, nodes, edges, offsets, ids = ref._build_native(spaths, k, w, cores, False)
        ref._get_penalty_native(kmers, nodes, offsets, is_t, cores)
        dt = time.perf_counter() - t0
        if i >= warmup:
            times.append(dt)
        graph = (kmers, nodes, edges, offsets)
    return {"seconds": float(np.mean(times)), "cores": cores, "kind": kind, "graph": graph if keep_graph else None}


def sha(a: np.ndarray) -> str:
    return hashlib.sha256(np.ascontiguousarray(a).view(np.uint8).data).hexdigest()


# ---- our arm -----------------------------------------------------------------------------------
def build_batch(ss: SynthSet, genomes: range, threads: int, chunk: int = PARITY_GENOMES_PER_RANK, keep_first: bool = False):
    """Generate genomes and pack them into one pinned 2-bit batch through the C ABI, `chunk` genomes at
    a time (the unpacked ASCII of one chunk only is ever held).  keep_first also returns the batch of the
    first chunk on its own (the parity subset of a multi-GPU run)."""
    from seqwin_b200 import _lib
    L = _lib.lib()
    parts = []
    glist = list(genomes)
    pool = ThreadPoolExecutor(max(1, min(threads, 8)))
    for c0 in range(0, max(1, len(glist)), chunk):
        gs = glist[c0:c0 + chunk]
        recs_by_g = list(pool.map(ss.records, gs))
        arrs, lens, asm_of, ids = [], [], [], []
        for a, recs in enumerate(recs_by_g):
            for rid, seq in recs:
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
        _lib.check(L.sw_batch_from_memory(ptrs, lens_a.ctypes.data, asm_a.ctypes.data, idp, n, len(gs), threads, C.byref(b)))
        parts.append(b)
    pool.shutdown()
    if len(parts) == 1 and not keep_first:
        return parts[0]
    whole = C.c_void_p()
    arr = (C.c_void_p * len(parts))(*[p.value for p in parts])
    _lib.check(L.sw_batch_concat(arr, len(parts), C.byref(whole)))
    for p in parts[1 if keep_first else 0:]:
        L.sw_batch_free(p)
    return (whole, parts[0]) if keep_first else whole


def int_peak(L, _lib) -> dict:
    out = (C.c_double * 4)()
    _lib.check(L.sw_measure_int_peak(out))
    return {"lop3_Tops": out[0] / 1e12, "shf_Tops": out[1] / 1e12, "iadd_Tops": out[2] / 1e12, "mix_Tops": out[3] / 1e12}


def main():
    args = parse_args()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    k, w = args.k, args.w
    spec = spec_for(args, world)
    per_gpu = genomes_per_gpu(args, world)
    workload = workload_name(spec, k, w)
    cores = os.cpu_count() or 8

    if args.impl == "reference":
        if rank != 0:
            return
        ss = SynthSet(spec)
        # N = 1: the whole workload per step.  N > 1: a bounded 500-genome sample of it per step
        n_s = args.sample_genomes or (spec.n_genomes if world == 1 else REF_SAMPLE_GENOMES_MULTI)
        idx = sample_indices(spec, n_s)
        d = shm_dir("seqwin_b200_ref_")
        try:
            n_bases = write_genomes(ss, idx, d, min(cores, 8))
            r = run_reference([fasta_path(d, g) for g in idx], ss.is_targets[idx], k, w, args.steps, args.warmup)
        finally:
            shutil.rmtree(d, ignore_errors=True)
        value = n_bases / r["seconds"] / 1e9
        sample = ("the whole workload" if len(idx) == spec.n_genomes else f"{len(idx)} of the workload's genomes") + \
                 f" ({n_bases / 1e6:.0f} Mbp) per step, plain FASTA on tmpfs, _build_native + _get_penalty_native"
        line = {"impl": "reference", "metric": "Gbp/s sketched+graph-built", "value": value, "unit": "Gbp/s",
                "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": r["seconds"] * 1e3,
                "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u64", "data": "synthetic",
                "config": {"workload": workload, "genomes_per_step": len(idx)},
                "cpu_baseline": {"value": value, "unit": "Gbp/s", "cores": r["cores"], "kind": r["kind"], "sample": sample},
                "e2e": {"value": value, "unit": "Gbp/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
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
        # the record exchange is one large all-to-all per array: with NCCL's default P2P channel count it reaches
        # about 430 GB/s per GPU over NVLink 5, with 64 channels about 620 (tools/a2a_bench.py, profiles/)
        os.environ.setdefault("NCCL_MIN_P2P_NCHANNELS", "64")
        os.environ.setdefault("NCCL_MAX_P2P_NCHANNELS", "64")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))

    threads = max(1, cores // max(1, world))
    ss = SynthSet(spec)
    my_genomes = range(rank * per_gpu, (rank + 1) * per_gpu)
    t0 = time.perf_counter()
    parity_batch = None
    if world > 1 and not args.no_parity:
        batch, parity_batch = build_batch(ss, my_genomes, threads, keep_first=True)
    else:
        batch = build_batch(ss, my_genomes, threads)
    gen_s = time.perf_counter() - t0
    n_bases_local = L.sw_batch_n_bases(batch)
    packed_bytes = L.sw_batch_packed_bytes(batch)
    n_records = L.sw_batch_n_records(batch)
    parity = None

    if world > 1:
        from seqwin_b200 import dist as swdist
        result = swdist.bench_loop(L, batch, spec, rank, world, k, w, args.steps, args.warmup, ClockSampler)
        if parity_batch is not None:
            parity = swdist.bench_parity(L, parity_batch, ss, rank, world, per_gpu, PARITY_GENOMES_PER_RANK, k, w,
                                         sys.modules[__name__])
            L.sw_batch_free(parity_batch)
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
        L.sw_batch_free(batch)
        if dist is not None:
            dist.destroy_process_group()
        return

    stages = result["stages"]
    total_ms = float(np.mean([s["total_ms"] for s in stages]))
    sketch_ms = float(np.mean([s["sketch_kernel_ms"] for s in stages]))
    s0 = stages[-1]
    n_bases_total = result.get("n_bases_total", n_bases_local)
    M, Un, Ue = s0["n_kmers"], s0["n_nodes"], s0["n_edges"]
    M_loc = s0.get("n_kmers_local", M)
    Un_loc, Ue_loc = s0.get("n_nodes_local", Un), s0.get("n_edges_local", Ue)
    value = n_bases_total / (total_ms * 1e-3) / 1e9

    peaks = {}
    pk = ROOT / "MEASURED_PEAKS.json"
    if pk.exists():
        peaks = json.loads(pk.read_text())
    hbm_peak = float(peaks.get("hbm_gbs", 6650.0))
    peak_src = "measured copy bandwidth (MEASURED_PEAKS.json)" if peaks else "fallback (B200_PROFILING.md)"
    clocks = result["clocks"]
    stage_ms = {n: float(np.mean([s[n] for s in stages]))
                for n in ("plan_ms", "sketch_kernel_ms", "reorder_ms", "sketch_ms", "sort_nodes_ms", "nodes_ms", "edges_ms")
                if n in stages[0]}

    # sketch kernel (dominant): INT32 ALU pipe.  Algorithmic lane-ops per launch = 45 * N + 12 * M;
    # the peak is measured on this device, in this run (dependent LOP3 chains, csrc/diag.cu)
    ip = int_peak(L, _lib)
    int_alg = 45.0 * n_bases_local + 12.0 * M_loc
    int_ach = int_alg / (sketch_ms * 1e-3) / 1e12
    sketch_bytes = n_bases_local / 4 + 16 * M_loc
    traffic = {}
    tp = ROOT / "profiles" / "r2_traffic.json"
    if tp.exists():
        traffic = json.loads(tp.read_text())
    sk_t = traffic.get("sketch")
    sketch_traffic = (sk_t["dram_bytes"] * sketch_bytes / sk_t["algorithmic_bytes"]) if sk_t else None
    kernel_name = "sketch_sparse_kernel<128,64,20>" if (w >= 144 and w <= 2048 and not os.environ.get("SEQWIN_SKETCH_DENSE")) \
        else "sketch kernel selected for this (k, w)"
    # aggregation stage (sort + nodes + edges): HBM.  Algorithmic bytes = read the 16 B records once, write the
    # k-mers, nodes and edges once
    agg_ms = stage_ms.get("sort_nodes_ms", 0) + stage_ms.get("nodes_ms", 0) + stage_ms.get("edges_ms", 0)
    agg_bytes = 24 * M_loc + 40 * Un_loc + 24 * Ue_loc
    ag_t = traffic.get("aggregation")
    path_bytes = n_bases_local / 4 + 40 * M_loc + 40 * Un_loc + 24 * Ue_loc
    local_ms = float(np.mean([s.get("local_build_ms", s["total_ms"]) for s in stages]))
    # The model counts integer instructions of any kind, so the denominator is the measured integer ISSUE rate (four
    # schedulers x 32 lanes per clock: what independent IADD chains reach, since adds go down two pipes).  LOP3 / SHF /
    # ISETP -- most of a hash step -- share ONE of those pipes at half that rate: the ALU-pipe view (ncu) is the
    # tighter bound and is reported beside it.
    roofline = {"bound": "int", "kernel": kernel_name, "achieved": int_ach, "peak": ip["iadd_Tops"], "unit": "Tlane-op/s",
                "frac": int_ach / ip["iadd_Tops"], "traffic": sketch_traffic,
                "peak_source": "measured in this run: integer issue rate, independent IADD chains on every SM "
                               "(sw_measure_int_peak, csrc/diag.cu)",
                "algorithmic_ops_per_launch": int_alg, "kernel_ms": sketch_ms, "int_peaks_measured": ip,
                "model": "I_alg = 45*N + 12*M lane-ops (SURVEY 8d)",
                "alu_pipe_view": {"peak_lop3_Tops": ip["lop3_Tops"],
                                  "busy_pct_ncu": (sk_t or {}).get("alu_pipe_busy_pct"),
                                  "issue_slots_busy_pct_ncu": (sk_t or {}).get("issue_active_pct"),
                                  "what": "LOP3 / SHF / ISETP issue on one pipe at half the issue rate; ncu capture of this "
                                          "round (profiles/r2_traffic.json)"},
                "hbm_view": {"achieved": sketch_bytes / (sketch_ms * 1e-3) / 1e9, "peak": hbm_peak, "unit": "GB/s",
                             "frac": sketch_bytes / (sketch_ms * 1e-3) / 1e9 / hbm_peak,
                             "algorithmic_bytes_per_launch": sketch_bytes, "peak_source": peak_src},
                "stages": {"aggregation": {"bound": "hbm", "what": "node sort + nodes/k-mers + edge stage",
                                           "algorithmic_bytes": agg_bytes, "ms": agg_ms,
                                           "achieved": agg_bytes / (agg_ms * 1e-3) / 1e9 if agg_ms else None,
                                           "peak": hbm_peak, "unit": "GB/s",
                                           "frac": agg_bytes / (agg_ms * 1e-3) / 1e9 / hbm_peak if agg_ms else None,
                                           "traffic": (ag_t["dram_bytes"] * agg_bytes / ag_t["algorithmic_bytes"]) if ag_t else None},
                           "path": {"bound": "hbm", "algorithmic_bytes": path_bytes, "ms": local_ms,
                                    "achieved": path_bytes / (local_ms * 1e-3) / 1e9, "peak": hbm_peak, "unit": "GB/s",
                                    "frac": path_bytes / (local_ms * 1e-3) / 1e9 / hbm_peak}},
                "stage_ms": stage_ms,
                "dist_ms": {n: float(np.mean([s[n] for s in stages])) for n in
                            ("phase_local_ms", "phase_exchange_merge_ms", "phase_merge_ms") if n in stages[0]}}

    e2e_runs = result["e2e_runs"]
    e2e_s = float(np.mean([r[0] for r in e2e_runs]))
    d2h = 8 * M + 40 * Un + 24 * Ue
    e2e = {"value": n_bases_total / e2e_s / 1e9, "unit": "Gbp/s",
           "h2d_bytes_per_step": int(packed_bytes + 12 * n_records) * world, "d2h_bytes_per_step": int(d2h),
           "ms_per_step": e2e_s * 1e3,
           "stage_ms": {n: float(np.mean([r[1][n] for r in e2e_runs])) for n in
                        ("h2d_ms", "plan_ms", "sketch_kernel_ms", "reorder_ms", "sort_nodes_ms", "nodes_ms",
                         "edges_ms", "total_ms", "d2h_ms", "local_build_ms", "fetch_ms") if n in e2e_runs[0][1]},
           "what": "pinned 2-bit host batch -> H2D -> sketch + graph + get_penalty -> D2H host arrays (copies overlapped "
                   "with the kernels)" + ("; every rank fetches its hash range" if world > 1 else "")}

    cpu_baseline = None
    if world == 1 and not args.no_cpu_baseline:
        # the whole workload through the unmodified reference, timed once after a warm-up pass, and every
        # array of OUR graph from the same FASTA files compared with it
        n_s = args.sample_genomes or spec.n_genomes
        idx = sample_indices(spec, n_s)
        d = shm_dir("seqwin_b200_cpu_")
        try:
            nb = write_genomes(ss, idx, d, min(cores, 8))
            paths = [fasta_path(d, g) for g in idx]
            is_t_s = np.ascontiguousarray(ss.is_targets[idx], dtype=np.bool_)
            r = run_reference(paths, is_t_s, k, w, steps=1, warmup=1, keep_graph=not args.no_parity)
            cpu_baseline = {"value": nb / r["seconds"] / 1e9, "unit": "Gbp/s", "cores": r["cores"], "kind": r["kind"],
                            "sample": ("the whole workload" if len(idx) == spec.n_genomes else f"{len(idx)} genomes")
                                      + f" ({nb / 1e6:.0f} Mbp), plain FASTA on tmpfs, build + get_penalty, 1 timed pass after 1 warm-up"}
            if not args.no_parity:
                from seqwin_b200.graph import KmerGraph, _get_penalty
                for _ in range(2):   # like the reference: one warm-up pass, one timed pass
                    t0 = time.perf_counter()
                    g = KmerGraph(paths, k, w, n_cpu=cores)
                    _get_penalty(g.kmers, g.nodes, g.record_offsets, is_t_s)
                    ours_s = time.perf_counter() - t0
                rk, rn, re_, ro = r["graph"]
                ok = {"kmers": bool(np.array_equal(g.kmers, rk)), "nodes": bool(np.array_equal(g.nodes, rn)),
                      "edges": bool(np.array_equal(g.edges, re_)), "record_offsets": bool(np.array_equal(g.record_offsets, ro))}
                parity = {"bit_exact": all(ok.values()), "arrays": ok, "against": r["kind"],
                          "what": f"{len(idx)} genomes ({nb / 1e6:.0f} Mbp) from FASTA: KmerGraph + _get_penalty on the GPU vs "
                                  "the reference's _build_native + _get_penalty_native, every array compared",
                          "sha256": {"kmers": sha(g.kmers), "nodes": sha(g.nodes), "edges": sha(g.edges)},
                          "graph": {"n_kmers": len(rk), "n_nodes": len(rn), "n_edges": len(re_)}}
                cpu_baseline["ours_same_input_from_fasta_gbps"] = nb / ours_s / 1e9
        finally:
            shutil.rmtree(d, ignore_errors=True)
    elif world > 1 and parity is not None:
        cpu_baseline = parity.pop("cpu_baseline", None)

    line = {"metric": "Gbp/s sketched+graph-built", "value": value, "unit": "Gbp/s", "n_gpus": world,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": total_ms, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "u64", "data": "synthetic",
            "config": {"workload": workload, "l2": "inputs larger than L2 (packed batch %.0f MB per GPU > 126 MB)" % (packed_bytes / 1e6),
                       "timing": "CUDA events on the library stream around each step, max over ranks (host tile planning included)",
                       "genomes_per_gpu": per_gpu, "gen_seconds": round(gen_s, 1)},
            "graph": {"n_bases": int(n_bases_total), "n_kmers": int(M), "n_nodes": int(Un), "n_edges": int(Ue)},
            "clocks": clocks, "e2e": e2e, "gpu_launches": int(sum(s["total_launches"] for s in stages)),
            "roofline": roofline, "cpu_baseline": cpu_baseline,
            "parity_sample_bit_exact": None if parity is None else bool(parity["bit_exact"]), "parity": parity}
    ms = (C.c_uint64 * 4)()
    _lib.check(L.sw_mem_stats(ms))
    line["device_memory_gb"] = {"scratch_arena_high": ms[0] / 1e9, "pool_used_high": ms[1] / 1e9, "in_use_now": ms[3] / 1e9}
    if "single_gpu" in result:
        line["single_gpu_same_shard"] = result["single_gpu"]
    if "full_size_checks" in result:
        line["full_size_checks"] = result["full_size_checks"]
    if "mode" in result:
        line["config"]["multi_gpu_build"] = result["mode"]
        line["last_step_ms_by_rank"] = result["last_step_ms_by_rank"]
    print(json.dumps(line))
    L.sw_batch_free(batch)
    if dist is not None:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
