"""Multi-GPU graph build: one process per GPU, genomes sharded, hash ranges owned.

    rank r:  shard of assemblies --(single-GPU kernels)--> local graph (global record indices)
             cut the sorted node / k-mer / edge arrays at the hash boundaries i * 2^64 / P
             all-to-all the slices (NCCL over NVLink; torch.distributed is only the plumbing)
             merge what arrived for hash range r  (csrc/dist.cu, mirrors merge_thread_graphs,
             cpp/src/seqwin/build_internals.cpp:295-392)

An assembly lives on exactly one rank, so k-mer lists concatenate in rank order and edge weights
add; concatenating the ranks' outputs in rank order is the reference graph.

The exchange logic is backend-agnostic: :class:`CudaStages` runs the stages through the C ABI on
device pointers; tests drive the same :func:`exchange_and_merge` over ``gloo`` with a numpy stand-in.
"""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass

import numpy as np
import torch
import torch.distributed as dist

KMER_BYTES, NODE_BYTES, EDGE_BYTES = 8, 40, 24


@dataclass
class LocalGraph:
    """A shard's graph as three flat uint8 tensors plus the hash-range cut points."""
    kmers: torch.Tensor
    nodes: torch.Tensor
    edges: torch.Tensor
    node_split: np.ndarray   # [P+1]
    kmer_split: np.ndarray   # [P+1]
    edge_split: np.ndarray   # [P+1]
    handle: object = None
    early: object = None     # node / k-mer exchange already in flight (exchange_nodes result)


def all_to_all_bytes(chunks: list[torch.Tensor], group=None) -> list[torch.Tensor]:
    """chunks[d] (uint8, 1-D) goes to rank d; returns what every rank sent to this one."""
    world = dist.get_world_size(group)
    dev = chunks[0].device
    send_counts = torch.tensor([c.numel() for c in chunks], dtype=torch.int64, device=dev)
    recv_counts = torch.empty(world, dtype=torch.int64, device=dev)
    if dist.get_backend(group) == "nccl":
        dist.all_to_all_single(recv_counts, send_counts, group=group)
        rc = recv_counts.tolist()
        out = torch.empty(sum(rc), dtype=torch.uint8, device=dev)
        dist.all_to_all_single(out, torch.cat(chunks), output_split_sizes=rc,
                               input_split_sizes=send_counts.tolist(), group=group)
        return list(out.split(rc))
    # gloo has no all_to_all: pairwise exchange through host memory (CPU tests, 1-GPU debugging)
    rank = dist.get_rank(group)
    gathered = [torch.empty(world, dtype=torch.int64) for _ in range(world)]
    dist.all_gather(gathered, send_counts.cpu(), group=group)
    rc = [int(gathered[s][rank]) for s in range(world)]
    host_chunks = [c.cpu().contiguous() for c in chunks]
    outs = [torch.empty(n, dtype=torch.uint8) for n in rc]
    outs[rank].copy_(host_chunks[rank])
    reqs = []
    for peer in range(world):
        if peer == rank:
            continue
        if host_chunks[peer].numel():
            reqs.append(dist.isend(host_chunks[peer], peer, group=group))
        if rc[peer]:
            reqs.append(dist.irecv(outs[peer], peer, group=group))
    for r in reqs:
        r.wait()
    return [o.to(dev) for o in outs]


def all_to_all_slices(tensors: list[torch.Tensor], splits: list[np.ndarray], item_bytes: list[int], group=None,
                      async_op: bool = False, extra: np.ndarray | None = None):
    """Several byte tensors, each already laid out as P consecutive destination slices
    (splits[t][d] .. splits[t][d+1], in items).  One count exchange for all of them, then one
    all_to_all_single per tensor straight out of / into contiguous buffers (no concatenation).
    `extra` ([P] int64, one value per destination) rides on the count exchange.
    Returns (received tensors, received item counts per source [T, P], pending work handles,
    received extra values [P] or None); with async_op (NCCL only) the data transfers are left in
    flight on NCCL's stream."""
    world = dist.get_world_size(group)
    dev = tensors[0].device
    send_items = np.stack([np.diff(sp.astype(np.int64)) for sp in splits])          # [T, P]
    if dist.get_backend(group) != "nccl":
        outs, counts = [], []
        for t, sp, ib in zip(tensors, splits, item_bytes):
            chunks = [t[int(sp[d]) * ib:int(sp[d + 1]) * ib] for d in range(world)]
            rc = all_to_all_bytes(chunks, group)
            outs.append(torch.cat(rc))
            counts.append([c.numel() // ib for c in rc])
        recv_extra = None
        if extra is not None:
            ex = torch.from_numpy(np.ascontiguousarray(extra, dtype=np.int64)).to(dev).view(torch.uint8)
            rc = all_to_all_bytes([ex[8 * d:8 * d + 8] for d in range(world)], group)
            recv_extra = torch.cat(rc).view(torch.int64).cpu().numpy()
        return outs, np.array(counts, dtype=np.uint64), [], recv_extra
    rows = [send_items] if extra is None else [send_items, np.asarray(extra, dtype=np.int64)[None, :]]
    send_t = torch.from_numpy(np.ascontiguousarray(np.concatenate(rows).T)).to(dev)  # [P, T(+1)]: row d goes to rank d
    recv_t = torch.empty_like(send_t)
    dist.all_to_all_single(recv_t, send_t, group=group)
    recv_all = recv_t.cpu().numpy().T                                                # [T(+1), P]
    recv_items = recv_all[:len(tensors)]
    recv_extra = recv_all[len(tensors)].copy() if extra is not None else None
    outs, works = [], []
    for i, (t, sp, ib) in enumerate(zip(tensors, splits, item_bytes)):
        out = torch.empty(int(recv_items[i].sum()) * ib, dtype=torch.uint8, device=dev)
        src = t[int(sp[0]) * ib:int(sp[-1]) * ib]
        wk = dist.all_to_all_single(out, src, output_split_sizes=(recv_items[i] * ib).tolist(),
                                    input_split_sizes=(send_items[i] * ib).tolist(), group=group, async_op=async_op)
        if async_op:
            works.append(wk)
        outs.append(out)
    return outs, recv_items.astype(np.uint64), works, recv_extra


def exchange_nodes(nodes, kmers, node_split, kmer_split, group=None, async_op=False):
    """First half of the exchange: node and k-mer slices.  The owner also needs each sender's k-mer
    base to rebase node.start: one value per destination, carried by the count exchange.
    Returns ((recv_nodes, recv_kmers, recv_kmer_base), counts, works)."""
    world = dist.get_world_size(group)
    (recv_nodes, recv_kmers), counts, works, recv_base = all_to_all_slices(
        [nodes, kmers], [node_split, kmer_split], [NODE_BYTES, KMER_BYTES], group, async_op,
        extra=kmer_split[:world].astype(np.int64))
    return (recv_nodes, recv_kmers, recv_base.astype(np.uint64)), counts, works


def exchange_and_merge(stages, local: LocalGraph, group=None, early=None):
    """Send every hash range to its owner and merge what arrives; returns the merged graph.
    `early` is the result of an exchange_nodes() already started while the edge stage was running.
    Stages that can merge in two halves (CudaStages) merge nodes + k-mers while the edge slices are
    still on the wire."""
    if early is None:
        early = exchange_nodes(local.nodes, local.kmers, local.node_split, local.kmer_split, group)
    (recv_nodes, recv_kmers, recv_base), counts, works = early
    two_halves = hasattr(stages, "merge_nodes") and dist.get_backend(group) == "nccl"
    (recv_edges,), ecounts, eworks, _ = all_to_all_slices([local.edges], [local.edge_split], [EDGE_BYTES], group,
                                                          async_op=two_halves)
    for wk in works:
        wk.wait()
    kmer_base = np.ascontiguousarray(recv_base, dtype=np.uint64)
    node_args = (recv_nodes, np.ascontiguousarray(counts[0]), recv_kmers, np.ascontiguousarray(counts[1]), kmer_base)
    if not two_halves:
        return stages.merge(*node_args, recv_edges, np.ascontiguousarray(ecounts[0]))
    g = stages.merge_nodes(*node_args)
    for wk in eworks:
        wk.wait()
    stages.merge_edges(g, recv_edges, np.ascontiguousarray(ecounts[0]))
    return g


def record_base(n_records_local: int, device, group=None) -> tuple[int, int]:
    """(global index of this shard's first record, total records)."""
    world, rank = dist.get_world_size(group), dist.get_rank(group)
    if dist.get_backend(group) != "nccl":
        device = torch.device("cpu")
    counts = [torch.zeros(1, dtype=torch.int64, device=device) for _ in range(world)]
    dist.all_gather(counts, torch.tensor([n_records_local], dtype=torch.int64, device=device), group=group)
    counts = [int(c) for c in counts]
    return sum(counts[:rank]), sum(counts)


# ---- CUDA stages (C ABI) ------------------------------------------------------------------------

class _DevView:
    """Zero-copy uint8 view of device memory owned by libseqwin_b200 (CUDA array interface)."""

    def __init__(self, ptr: int, nbytes: int):
        self.__cuda_array_interface__ = {"shape": (nbytes,), "typestr": "|u1", "data": (ptr, False),
                                         "version": 3, "strides": None}


def _view(ptr, nbytes: int, device) -> torch.Tensor:
    if not nbytes or not ptr:
        return torch.empty(0, dtype=torch.uint8, device=device)
    return torch.as_tensor(_DevView(int(ptr), int(nbytes)), device=device)


class CudaStages:
    """The product stages: everything runs in libseqwin_b200.so on torch's current stream."""

    def __init__(self, lib, device):
        from . import _lib
        self.L, self._lib, self.device = lib, _lib, device
        self._lib.check(self.L.sw_set_stream(C.c_void_p(torch.cuda.current_stream(device).cuda_stream)))
        self.times = None
        self.merge_launches = 0
        self.phase_events = None
        self.merge_ms = 0.0

    def local_build(self, dev_batch, k: int, w: int, rec_base: int, world: int, host_batch=None,
                    on_nodes=None, is_targets=None) -> LocalGraph:
        """Single-GPU kernels on this rank's shard.  on_nodes(nodes, kmers, node_split, kmer_split), if
        given, is called from inside the build as soon as nodes + k-mers are final on the device (the
        edge stage follows); its return value lands in LocalGraph.early.  is_targets (classes of this
        shard's assemblies) makes the shard's nodes carry their n_tar / n_neg counts."""
        L, lb = self.L, self._lib
        g = C.c_void_p()
        self.times = lb.StageTimes()
        early, failure = [], []

        def cuts(gh, want_edges: bool):
            pk, pn, pe = C.c_void_p(), C.c_void_p(), C.c_void_p()
            lb.check(L.sw_graph_device_ptrs(gh, C.byref(pk), C.byref(pn), C.byref(pe)))
            n_k, n_n, n_e = (L.sw_graph_size(gh, i) for i in (lb.SW_KMERS, lb.SW_NODES, lb.SW_EDGES))
            ns, ks, es = (np.zeros(world + 1, dtype=np.uint64) for _ in range(3))
            lb.check(L.sw_graph_split(gh, world, ns.ctypes.data, ks.ctypes.data, es.ctypes.data if want_edges else None))
            return (_view(pk.value, n_k * KMER_BYTES, self.device), _view(pn.value, n_n * NODE_BYTES, self.device),
                    _view(pe.value, n_e * EDGE_BYTES, self.device), ns, ks, es)

        def hook(_user, gh):
            try:   # an exception must not unwind through the C frames
                kmers, nodes, _, ns, ks, _ = cuts(C.c_void_p(gh), want_edges=False)
                early.append(on_nodes(nodes, kmers, ns, ks))
            except BaseException as exc:   # noqa: BLE001 - re-raised below
                failure.append(exc)

        cb = lb.NODES_READY_FN(hook)
        if on_nodes is not None:
            lb.check(L.sw_set_nodes_ready(cb, None))
        try:
            t_ptr = is_targets.ctypes.data if is_targets is not None else None
            t_len = len(is_targets) if is_targets is not None else 0
            if host_batch is not None:   # end-to-end: the H2D copy is sliced and overlapped with the sketch
                lb.check(L.sw_build_from_batch_ex(host_batch, k, w, rec_base, 0, t_ptr, t_len, C.byref(g),
                                                  C.byref(self.times)))
            else:
                lb.check(L.sw_dev_build_ex(dev_batch, k, w, rec_base, t_ptr, t_len, C.byref(g), C.byref(self.times)))
        finally:
            if on_nodes is not None:
                L.sw_set_nodes_ready(lb.NODES_READY_FN(0), None)
        if failure:
            L.sw_graph_free(g)
            raise failure[0]
        kmers, nodes, edges, ns, ks, es = cuts(g, want_edges=True)
        local = LocalGraph(kmers, nodes, edges, ns, ks, es, handle=g)
        local.early = early[0] if early else None
        return local

    def free_local(self, local: LocalGraph) -> None:
        if local.handle:
            self.L.sw_graph_free(local.handle)
            local.handle = None

    def merge(self, nodes, node_counts, kmers, kmer_counts, kmer_base, edges, edge_counts):
        g = self.merge_nodes(nodes, node_counts, kmers, kmer_counts, kmer_base)
        self.merge_edges(g, edges, edge_counts)
        return g

    def merge_nodes(self, nodes, node_counts, kmers, kmer_counts, kmer_base):
        L, lb = self.L, self._lib
        torch.cuda.current_stream(self.device).synchronize()   # received slices visible to the library stream
        g = C.c_void_p()
        n_launch = C.c_uint32()
        m0, m1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        m0.record()
        self._merge_events = (m0, m1)
        lb.check(L.sw_dist_merge(C.c_void_p(nodes.data_ptr()), node_counts.ctypes.data, C.c_void_p(kmers.data_ptr()),
                                 kmer_counts.ctypes.data, kmer_base.ctypes.data, None, None, len(node_counts),
                                 C.byref(g), C.byref(n_launch)))
        self.merge_launches = n_launch.value
        m1.record()
        return g

    def finish_penalty(self, g, class_totals) -> None:
        """penalty of the merged nodes from their summed n_tar / n_neg and the global class sizes"""
        self._lib.check(self.L.sw_graph_finish_penalty(g, C.c_uint64(class_totals[0]), C.c_uint64(class_totals[1])))

    def merge_edges(self, g, edges, edge_counts):
        L, lb = self.L, self._lib
        torch.cuda.current_stream(self.device).synchronize()
        n_launch = C.c_uint32()
        lb.check(L.sw_dist_merge_edges(g, C.c_void_p(edges.data_ptr()), edge_counts.ctypes.data, len(edge_counts),
                                       C.byref(n_launch)))
        self.merge_launches += n_launch.value
        self._merge_events[1].record()


def dist_build(stages: CudaStages, dev_batch, n_records_local: int, k: int, w: int, group=None, rec_base=None,
               host_batch=None, overlap: bool = True, is_targets=None, class_totals=None):
    """Full multi-GPU build of this rank's hash range; returns the merged sw_graph handle.
    With is_targets (bool array, the classes of THIS rank's assemblies) the graph comes back scored:
    every shard counts its own assemblies, the merge adds the counts, and the penalty is finished with
    class_totals = (targets, non-targets) over all ranks (all-reduced here when not given)."""
    world = dist.get_world_size(group)
    if rec_base is None:
        rec_base, _ = record_base(n_records_local, stages.device, group)
    timed = torch.device(stages.device).type == "cuda"   # phase events for bench.py (the CPU stand-in has none)
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(3)] if timed else None
    if timed:
        ev[0].record()
    # the node / k-mer slices leave as soon as they are final; over NCCL the transfer stays in flight
    # while the edge stage runs (gloo exchanges synchronously inside the hook: same code path, no overlap)
    on_nodes = None
    if overlap:
        nccl = dist.get_backend(group) == "nccl"

        def on_nodes(nodes, kmers, node_split, kmer_split):
            return exchange_nodes(nodes, kmers, node_split, kmer_split, group, async_op=nccl)
    if is_targets is not None:
        is_targets = np.ascontiguousarray(is_targets, dtype=np.bool_)
        if class_totals is None:
            dev_t = stages.device if dist.get_backend(group) == "nccl" else torch.device("cpu")
            tot = torch.tensor([int(is_targets.sum()), int(len(is_targets) - is_targets.sum())], dtype=torch.int64, device=dev_t)
            dist.all_reduce(tot, group=group)
            class_totals = (int(tot[0]), int(tot[1]))
    local = stages.local_build(dev_batch, k, w, rec_base, world, host_batch=host_batch, on_nodes=on_nodes,
                               is_targets=is_targets)
    if timed:
        ev[1].record()
    try:
        g = exchange_and_merge(stages, local, group, early=local.early)
    finally:
        stages.free_local(local)
    if is_targets is not None:
        stages.finish_penalty(g, class_totals)
    if timed:
        ev[2].record()
    stages.phase_events = ev
    return g


def export_graph(lib, g):
    """Host numpy copies (kmers, nodes, edges) of a device-resident graph handle."""
    from . import _lib
    from ._core import EDGE_DTYPE, KMER_DTYPE, NODE_DTYPE
    kmers = np.empty(lib.sw_graph_size(g, _lib.SW_KMERS), dtype=KMER_DTYPE)
    nodes = np.empty(lib.sw_graph_size(g, _lib.SW_NODES), dtype=NODE_DTYPE)
    edges = np.empty(lib.sw_graph_size(g, _lib.SW_EDGES), dtype=EDGE_DTYPE)
    _lib.check(lib.sw_graph_export(g, kmers.ctypes.data, nodes.ctypes.data, edges.ctypes.data, None))
    return kmers, nodes, edges


def concat_rank_graphs(parts):
    """Concatenate per-rank (kmers, nodes, edges) in rank order into the global graph."""
    kmers = np.concatenate([p[0] for p in parts])
    nodes_l, off = [], 0
    for k_, n_, _ in parts:
        n_ = n_.copy()
        n_["start"] += off
        n_["stop"] += off
        off += len(k_)
        nodes_l.append(n_)
    return kmers, np.concatenate(nodes_l), np.concatenate([p[2] for p in parts])


def gather_graph(parts_local, group=None):
    """Gather every rank's (kmers, nodes, edges) on rank 0 and concatenate (tests / small graphs)."""
    world, rank = dist.get_world_size(group), dist.get_rank(group)
    gathered = [None] * world if rank == 0 else None
    dist.gather_object(parts_local, gathered, dst=0, group=group)
    return concat_rank_graphs(gathered) if rank == 0 else None


# ---- bench driver (bench.py --gpus N under torchrun) ----------------------------------------------

def bench_loop(L, batch, spec, rank: int, world: int, k: int, w: int, steps: int, warmup: int, sampler_cls) -> dict:
    from . import _lib
    device = torch.device("cuda", torch.cuda.current_device())
    stages = CudaStages(L, device)
    n_records = L.sw_batch_n_records(batch)
    n_bases_local = L.sw_batch_n_bases(batch)
    dev = C.c_void_p()
    _lib.check(L.sw_dev_upload(batch, C.byref(dev)))

    rec_base, _ = record_base(n_records, device)   # shard bookkeeping is input metadata, not per-step work
    # classes of the assemblies (input metadata): this rank's slice and the two class sizes
    n_asm_local = spec.n_genomes // world
    is_t = np.ascontiguousarray(np.arange(spec.n_genomes) < spec.n_targets, dtype=np.bool_)

    import os
    overlap = os.environ.get("SEQWIN_DIST_OVERLAP", "1") != "0"   # A/B switch for profiles/

    is_t_local = np.ascontiguousarray(is_t[rank * n_asm_local:(rank + 1) * n_asm_local])
    totals = (int(is_t.sum()), int(len(is_t) - is_t.sum()))   # class sizes are input metadata too

    def step():
        # build + scoring: every shard counts its assemblies, the merge adds, the penalty is finished last
        g = dist_build(stages, dev, n_records, k, w, rec_base=rec_base, overlap=overlap, is_targets=is_t_local,
                       class_totals=totals)
        sizes = [L.sw_graph_size(g, i) for i in (_lib.SW_KMERS, _lib.SW_NODES, _lib.SW_EDGES)]
        L.sw_graph_free(g)
        return sizes

    for _ in range(warmup):
        step()
    sampler = sampler_cls(torch.cuda.current_device())
    times, stage_dicts, sizes = [], [], None
    sampler.start()
    for _ in range(steps):
        dist.barrier()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        sizes = step()
        e1.record()
        torch.cuda.synchronize()
        t = torch.tensor([e0.elapsed_time(e1)], device=device)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        times.append(float(t))
        d = stages.times.as_dict()
        d["n_kmers_local"] = d["n_kmers"]
        d["local_build_ms"] = d["total_ms"]
        d["total_ms"] = float(t)
        d["total_launches"] = d["total_launches"] + stages.merge_launches
        pe = stages.phase_events
        d["phase_local_ms"] = pe[0].elapsed_time(pe[1])
        d["phase_exchange_merge_ms"] = pe[1].elapsed_time(pe[2])
        d["phase_merge_ms"] = stages._merge_events[0].elapsed_time(stages._merge_events[1])
        stage_dicts.append(d)
    clocks = sampler.stop()
    L.sw_dev_batch_free(dev)

    # end to end: pinned host batch -> H2D -> distributed build -> D2H of this rank's range
    import time
    e2e = []
    for i in range(max(1, warmup // 2) + steps):
        dist.barrier()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        g = dist_build(stages, None, n_records, k, w, rec_base=rec_base, host_batch=batch, overlap=overlap,
                       is_targets=is_t_local, class_totals=totals)
        _lib.check(L.sw_graph_fetch(g))     # this rank's hash range -> pinned host memory
        L.sw_graph_free(g)
        torch.cuda.synchronize()
        t = torch.tensor([time.perf_counter() - t0], device=device, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        if i >= max(1, warmup // 2):
            e2e.append((float(t), stage_dicts[-1]))

    tot = torch.tensor([n_bases_local] + sizes, dtype=torch.int64, device=device)
    dist.all_reduce(tot, op=dist.ReduceOp.SUM)
    n_bases_total, n_k, n_n, n_e = (int(x) for x in tot)
    for d in stage_dicts:
        d["n_kmers"], d["n_nodes"], d["n_edges"] = n_k, n_n, n_e
    L.sw_set_stream(None)
    return {"stages": stage_dicts, "clocks": clocks, "e2e_runs": e2e, "n_bases_total": n_bases_total}
