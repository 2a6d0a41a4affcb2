"""Multi-GPU graph build: one process per GPU, genomes sharded, hash ranges owned.

Fused build (the default):
    rank r:  shard of assemblies --sketch--> minimizer stream; records per hash bin counted, the counts all-gathered
             (the only collective: 64 numbers per rank at 8 ranks)
             ONE partition pass derives every record's owned neighbour hashes and scatters (h1, k-mer, neighbours)
             straight into the arrays of the GPU that owns the record's hash range -- peer memory mapped through
             CUDA IPC, written over NVLink -- in (bin, source rank) order (csrc/radix.cu route_scatter)
             after a barrier the owner aggregates its range with the single-GPU bucket kernels (csrc/agg.cuh): nodes,
             k-mers, the edges those nodes own, scoring -- nothing is left to merge
Routed build (SEQWIN_DIST=routed): the same with the records partitioned locally on the top byte of h1 and exchanged
    by four NCCL all-to-alls (csrc/graph.cu route_stream).
Merge-based build (SEQWIN_DIST=merge; round 1, and the fallback where peer memory cannot be mapped): every rank builds the graph of its shard, the sorted node / k-mer /
edge arrays are cut at the hash boundaries, exchanged and merged by the owner (csrc/dist.cu, mirrors
merge_thread_graphs, cpp/src/seqwin/build_internals.cpp:295-392).

An assembly lives on exactly one rank and ranks hold consecutive assemblies, so the records a range owner
receives, concatenated in rank order, are in global stream order; concatenating the ranks' outputs in rank order is
the reference graph.

The exchange logic is backend-agnostic: :class:`CudaStages` runs the stages through the C ABI on device
pointers; tests drive the same functions over ``gloo`` with a numpy stand-in.
"""
from __future__ import annotations

import ctypes as C
import os
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

# ---- routed build -----------------------------------------------------------------------------------------

PAIR_WEIGHT = 0.6    # cost of an owned adjacent pair relative to a record, at the owner (partition + node kernels vs edge kernel)


def range_bounds(world: int, bits: int = 8, pair_weight: float = PAIR_WEIGHT) -> np.ndarray:
    """Boundaries of the ranks' hash ranges in units of 2^-bits of the hash space ("bins"; bits = 8: top bytes),
    [world + 1] ascending from 0 to 2^bits.  A range owner handles the records of its range and the adjacent pairs
    they own; a pair belongs to the smaller hash, so low ranges own more pairs (density 2 (1 - x)): the boundaries
    are the quantiles of records + a * pairs, (x + a (2 x - x^2)) / (1 + a) = i / world."""
    n_bins = 1 << bits
    if world > n_bins:
        raise ValueError(f"at most {n_bins} hash ranges")
    a = float(pair_weight)
    b = [int(round(n_bins * ((1 + 2 * a) - np.sqrt((1 + 2 * a) ** 2 - 4 * a * (1 + a) * i / world)) / (2 * a)))
         for i in range(world + 1)]
    b[0], b[-1] = 0, n_bins
    for i in range(1, world):          # strictly increasing, whatever the rounding did
        b[i] = min(max(b[i], b[i - 1] + 1), n_bins - (world - i))
    return np.asarray(b, dtype=np.int64)


def fused_route_bits(world: int) -> int:
    """Bins of the fused routing pass: about eight per rank -- enough to balance the ranges, few enough that the
    owner's per-bin partition launches stay negligible."""
    bits = 3
    while bits < 8 and (1 << bits) < 8 * world:
        bits += 1
    return bits


@dataclass
class RoutedContext:
    """Input metadata every rank needs once per set of shards: where its records start, and the record offsets and
    classes of ALL assemblies (the owner of a hash range scores records of every shard)."""
    rec_base: int
    record_offsets: np.ndarray          # [A_total + 1] uint32, global record indices
    is_targets: np.ndarray | None       # [A_total] bool
    bounds: np.ndarray                  # [world + 1] top-byte boundaries
    peer: object = None                 # PeerBuffers of the fused build, allocated on first use
    peer_failed: bool = False           # peer memory could not be set up: the merge-based build is used instead
    prefer_merge: bool = False          # auto mode found the set repetitive: the merge-based build is used from then on


def routed_context(record_offsets_local: np.ndarray, is_targets_local, group=None) -> RoutedContext:
    world, rank = dist.get_world_size(group), dist.get_rank(group)
    counts = np.diff(np.asarray(record_offsets_local, dtype=np.int64))
    mine = (counts, None if is_targets_local is None else np.asarray(is_targets_local, dtype=np.bool_))
    every = [None] * world
    dist.all_gather_object(every, mine, group=group)
    all_counts = np.concatenate([c for c, _ in every])
    offsets = np.concatenate([[0], np.cumsum(all_counts)])
    if offsets[-1] > 0xFFFFFFFF:
        raise RuntimeError("more than 2^32-1 records")
    is_t = None if any(t is None for _, t in every) else np.ascontiguousarray(np.concatenate([t for _, t in every]), dtype=np.bool_)
    rec_base = int(sum(int(c.sum()) for c, _ in every[:rank]))
    return RoutedContext(rec_base, np.ascontiguousarray(offsets, dtype=np.uint32), is_t, range_bounds(world))


def dist_build_routed(stages, dev_batch, k: int, w: int, ctx: RoutedContext, group=None, host_batch=None, inspect=None):
    """The routed multi-GPU build of this rank's hash range; returns the graph handle (scored when the context holds
    the classes)."""
    world, rank = dist.get_world_size(group), dist.get_rank(group)
    timed = torch.device(stages.device).type == "cuda"
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(4)] if timed else None
    if timed:
        ev[0].record()
    routed = stages.sketch_route(dev_batch, k, w, ctx.rec_base, host_batch=host_batch)
    if timed:
        ev[1].record()
    try:
        split = np.asarray([routed.byte_off[int(b)] for b in ctx.bounds], dtype=np.uint64)
        recv, counts, _, _ = all_to_all_slices(routed.arrays, [split] * 4, [8] * 4, group)
    finally:
        stages.free_routed(routed)
    if timed:
        ev[2].record()
    n = int(counts[0].sum())
    g = stages.aggregate(recv, n, int(ctx.bounds[rank]), int(ctx.bounds[rank + 1]), ctx.record_offsets, ctx.is_targets,
                         routed.pairs_per_edge)
    if inspect is not None:
        inspect(None, g)
    if timed:
        ev[3].record()
        stages.phase_events = [ev[0], ev[1], ev[3]]
        stages._merge_events = (ev[2], ev[3])
    return g


@dataclass
class PeerBuffers:
    """Receive arrays of the fused routing: every rank owns 4 x capacity words (keys | vals | prev | next) of plain
    device memory that the other ranks have mapped through CUDA IPC (peer access over NVLink)."""
    capacity: int
    bases: list        # bases[r] = rank r's buffer as seen from this process (bases[rank] is this rank's own)


class PeerMemoryUnavailable(RuntimeError):
    """CUDA IPC / peer access could not be set up on some rank: the caller takes the merge-based build."""


def ensure_peer_buffers(stages, ctx: "RoutedContext", need_items: int, group=None) -> PeerBuffers:
    """Collective: (re)allocate the receive arrays when the largest hash range of this step does not fit.
    Raises PeerMemoryUnavailable on EVERY rank if any rank cannot allocate or map them."""
    world, rank = dist.get_world_size(group), dist.get_rank(group)
    pb = ctx.peer
    if pb is not None and pb.capacity >= need_items:
        return pb
    if pb is not None:
        for r, b in enumerate(pb.bases):
            if r != rank:
                stages.peer_close(b)
        dist.barrier(group)            # nobody has this rank's memory mapped any more
        stages.peer_free(pb.bases[rank])
        ctx.peer = None
    cap = int(need_items * 1.25) + 4096
    mine, handle = None, None
    try:
        mine, handle = stages.peer_alloc(4 * cap * 8)
    except RuntimeError:
        pass
    handles = [None] * world
    dist.all_gather_object(handles, handle, group=group)
    bases, ok = [None] * world, all(h is not None for h in handles)
    if ok:
        try:
            bases = [mine if r == rank else stages.peer_open(handles[r]) for r in range(world)]
        except RuntimeError:
            ok = False
    oks = [None] * world
    dist.all_gather_object(oks, ok, group=group)
    if not all(oks):
        for r, b in enumerate(bases):
            if b is not None and r != rank:
                stages.peer_close(b)
        dist.barrier(group)
        if mine is not None:
            stages.peer_free(mine)
        raise PeerMemoryUnavailable("CUDA IPC peer memory is not available on every rank")
    ctx.peer = PeerBuffers(cap, bases)
    return ctx.peer


def release_peer_buffers(stages, ctx: "RoutedContext", group=None) -> None:
    """Collective: unmap and free the receive arrays of a context."""
    pb, rank = ctx.peer, dist.get_rank(group)
    if pb is None:
        return
    for r, b in enumerate(pb.bases):
        if r != rank:
            stages.peer_close(b)
    dist.barrier(group)
    stages.peer_free(pb.bases[rank])
    ctx.peer = None


def dist_build_fused(stages, dev_batch, k: int, w: int, ctx: "RoutedContext", group=None, host_batch=None, inspect=None,
                     auto: bool = False):
    """The routed build with the exchange fused into the routing pass: no data-path collective at all.  The ranks
    all-gather 256 counts each, derive where every shard's records of every top byte go in the owner's arrays --
    (top byte, source rank) order: global stream order inside a byte --, and the pass that generates the owned
    neighbour hashes scatters its output straight into those arrays over NVLink (csrc/radix.cu route_scatter)."""
    world, rank = dist.get_world_size(group), dist.get_rank(group)
    timed = torch.device(stages.device).type == "cuda"
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(4)] if timed else None
    if timed:
        ev[0].record()
    rb = fused_route_bits(world)
    n_bins = 1 << rb
    routed, counts = stages.sketch_hist(dev_batch, k, w, ctx.rec_base, rb, host_batch=host_batch)
    if timed:
        ev[1].record()
    try:
        # the counts, and (last column) this shard's sampled pairs per distinct pair in thousandths
        mine_np = np.concatenate([counts.astype(np.int64), [int(round(1000.0 * routed.pairs_per_edge))]])
        if dist.get_backend(group) == "nccl":
            mine_t = torch.from_numpy(mine_np).to(stages.device)
            every_t = torch.empty(world * 257, dtype=torch.int64, device=stages.device)
            dist.all_gather_into_tensor(every_t, mine_t, group=group)
            mat = every_t.cpu().numpy().reshape(world, 257)
        else:
            every = [None] * world
            dist.all_gather_object(every, mine_np, group=group)
            mat = np.stack(every)
        nonempty = mat[:, :256].sum(axis=1) > 0
        if (auto and nonempty.any() and float(mat[nonempty, 256].mean()) / 1000.0 >= AUTO_MERGE_PAIRS_PER_EDGE
                and int(mat[:, :256].sum(axis=1).max()) <= MERGE_MAX_LOCAL_MINIMIZERS):
            raise _PreferMerge()      # every rank sees the same numbers
        mat = mat[:, :n_bins]
        bounds = range_bounds(world, rb)
        owner = np.repeat(np.arange(world), np.diff(bounds))                 # owner of every bin
        per_byte = mat.sum(axis=0)                                            # records of every bin, all shards
        n_owner = np.array([int(per_byte[bounds[o]:bounds[o + 1]].sum()) for o in range(world)])
        pb = ensure_peer_buffers(stages, ctx, int(n_owner.max()), group)
        # inside an owner's arrays: top bytes ascending, sources ascending inside a byte
        byte_start = np.zeros(n_bins, dtype=np.int64)
        for o in range(world):
            lo, hi = int(bounds[o]), int(bounds[o + 1])
            byte_start[lo:hi] = np.concatenate([[0], np.cumsum(per_byte[lo:hi])[:-1]])
        byte_base = np.zeros(256, dtype=np.uint64)
        byte_base[:n_bins] = byte_start + mat[:rank].sum(axis=0)
        route_ptrs = np.zeros(4 * 256, dtype=np.uint64)
        for a in range(4):
            route_ptrs[a * 256:a * 256 + n_bins] = [pb.bases[int(owner[b])] + a * pb.capacity * 8 for b in range(n_bins)]
        stages.routed_scatter(routed, route_ptrs, byte_base)     # returns when this shard's records have left
    finally:
        stages.free_routed(routed)
    dist.barrier(group)                                           # ... and now everybody's have arrived
    if timed:
        ev[2].record()
    lo, hi = int(bounds[rank]), int(bounds[rank + 1])
    seg = np.ascontiguousarray(np.concatenate([[0], np.cumsum(per_byte[lo:hi])]), dtype=np.uint64)
    mine = pb.bases[rank]
    g = stages.aggregate([mine + a * pb.capacity * 8 for a in range(4)], int(n_owner[rank]), lo, hi, ctx.record_offsets,
                         ctx.is_targets, routed.pairs_per_edge, byte_off=seg, range_bits=rb)
    if inspect is not None:
        inspect(None, g)
    if timed:
        ev[3].record()
        stages.phase_events = [ev[0], ev[1], ev[3]]
        stages._merge_events = (ev[2], ev[3])
    return g


@dataclass
class Routed:
    """A shard's routed records: four flat uint8 tensors (keys, vals, prev, next; 8 bytes per record)."""
    arrays: list
    byte_off: np.ndarray        # [257] first record of every top-byte value
    pairs_per_edge: float
    handle: object = None
    dev_batch: object = None    # a batch uploaded for this call only


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

    def sketch_route(self, dev_batch, k: int, w: int, rec_base: int, host_batch=None) -> Routed:
        """Sketch this rank's shard and partition its records on the top byte of h1 (sw_dev_sketch_route)."""
        L, lb = self.L, self._lib
        own = None
        if dev_batch is None:      # end-to-end: the packed batch starts in pinned host memory
            own = C.c_void_p()
            lb.check(L.sw_dev_upload(host_batch, C.byref(own)))
            dev_batch = own
        r = C.c_void_p()
        self.times = lb.StageTimes()
        try:
            lb.check(L.sw_dev_sketch_route(dev_batch, k, w, rec_base, C.byref(r), C.byref(self.times)))
        except BaseException:
            if own is not None:
                L.sw_dev_batch_free(own)
            raise
        ptrs = [C.c_void_p() for _ in range(4)]
        n = C.c_uint64()
        off = np.zeros(257, dtype=np.uint64)
        stats = np.zeros(2, dtype=np.float64)
        lb.check(L.sw_routed_info(r, C.byref(ptrs[0]), C.byref(ptrs[1]), C.byref(ptrs[2]), C.byref(ptrs[3]), C.byref(n),
                                  off.ctypes.data, stats.ctypes.data))
        arrays = [_view(p_.value, n.value * 8, self.device) for p_ in ptrs]
        self.merge_launches = 0
        return Routed(arrays, off, float(stats[1]), handle=r, dev_batch=own)

    def sketch_hist(self, dev_batch, k: int, w: int, rec_base: int, route_bits: int, host_batch=None):
        """Fused routing, first half: sketch the shard, count its records per top byte of h1 (sw_dev_sketch_hist).
        Returns (Routed without arrays, counts[256])."""
        L, lb = self.L, self._lib
        own = None
        if dev_batch is None:
            own = C.c_void_p()
            lb.check(L.sw_dev_upload(host_batch, C.byref(own)))
            dev_batch = own
        r = C.c_void_p()
        self.times = lb.StageTimes()
        counts = np.zeros(256, dtype=np.uint64)
        try:
            lb.check(L.sw_dev_sketch_hist(dev_batch, k, w, rec_base, route_bits, C.byref(r), counts.ctypes.data, C.byref(self.times)))
        except BaseException:
            if own is not None:
                L.sw_dev_batch_free(own)
            raise
        stats = np.zeros(2, dtype=np.float64)
        lb.check(L.sw_routed_info(r, None, None, None, None, None, None, stats.ctypes.data))
        self.merge_launches = 0
        return Routed([], counts, float(stats[1]), handle=r, dev_batch=own), counts

    def routed_scatter(self, routed: Routed, route_ptrs: np.ndarray, byte_base: np.ndarray) -> None:
        """Fused routing, second half: the generating partition pass whose scatter writes go to the owners' arrays."""
        ms = C.c_float()
        self._lib.check(self.L.sw_routed_scatter(routed.handle, route_ptrs.ctypes.data, byte_base.ctypes.data, C.byref(ms)))
        if self.times is not None:
            self.times.sort_nodes_ms += ms.value
            self.times.total_ms += ms.value
            self.times.total_launches += 1

    def peer_alloc(self, n_bytes: int):
        p_, h = C.c_void_p(), (C.c_ubyte * 64)()
        self._lib.check(self.L.sw_peer_alloc(n_bytes, C.byref(p_), h))
        return p_.value, bytes(h)

    def peer_open(self, handle: bytes) -> int:
        p_ = C.c_void_p()
        buf = (C.c_ubyte * 64).from_buffer_copy(handle)
        self._lib.check(self.L.sw_peer_open(buf, C.byref(p_)))
        return p_.value

    def peer_close(self, ptr: int) -> None:
        self._lib.check(self.L.sw_peer_close(C.c_void_p(ptr)))

    def peer_free(self, ptr: int) -> None:
        self._lib.check(self.L.sw_peer_free(C.c_void_p(ptr)))

    def free_routed(self, routed: Routed) -> None:
        if routed.handle:
            self.L.sw_routed_free(routed.handle)
            routed.handle = None
        if routed.dev_batch is not None:
            self.L.sw_dev_batch_free(routed.dev_batch)
            routed.dev_batch = None

    def aggregate(self, recv, n: int, byte_lo: int, byte_hi: int, record_offsets, is_targets, pairs_per_edge: float,
                  byte_off=None, range_bits: int = 8):
        """Graph of the records this rank owns (sw_dev_aggregate): the single-GPU bucket kernels on what arrived."""
        L, lb = self.L, self._lib
        torch.cuda.current_stream(self.device).synchronize()   # received records visible to the library stream
        g = C.c_void_p()
        agg = lb.StageTimes()
        t_ptr = is_targets.ctypes.data if is_targets is not None else None
        t_len = len(is_targets) if is_targets is not None else 0
        ptrs = [x if isinstance(x, int) else x.data_ptr() for x in recv]   # tensors, or raw device pointers (peer buffers)
        lb.check(L.sw_dev_aggregate(C.c_void_p(ptrs[0]), C.c_void_p(ptrs[1]), C.c_void_p(ptrs[2]),
                                    C.c_void_p(ptrs[3]), n, byte_lo, byte_hi, record_offsets.ctypes.data,
                                    len(record_offsets), t_ptr, t_len, float(pairs_per_edge),
                                    None if byte_off is None else byte_off.ctypes.data, range_bits, C.byref(g), C.byref(agg)))
        if self.times is not None:    # one stage record per step: the routing pass counts as part of the sort stage
            self.times.sort_nodes_ms += agg.sort_nodes_ms
            self.times.nodes_ms, self.times.edges_ms = agg.nodes_ms, agg.edges_ms
            self.times.n_nodes, self.times.n_edges = agg.n_nodes, agg.n_edges
            self.times.total_ms += agg.total_ms
            self.times.total_launches += agg.total_launches
        return g

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


AUTO_MERGE_PAIRS_PER_EDGE = 8.0
# The merge-based build has been run with up to 1.9e8 minimizers per rank; with 9.1e8 (1,000 genomes per GPU at
# w = 10) its owner-side merge fails with an illegal address (profiles/r2_c3_n2_notes.txt; not yet understood).
# auto never picks it beyond this size, and bench.py's cross-check of the two builds is skipped there.
MERGE_MAX_LOCAL_MINIMIZERS = 400_000_000


def dist_mode() -> str:
    """SEQWIN_DIST: 'auto' (default), 'fused' (the routing pass writes the records into the owners' memory; falls back
    to 'merge' where CUDA IPC peer memory is not available), 'routed' (records exchanged through NCCL before they are
    aggregated) or 'merge' (shard graphs merged by the range owners).
    'auto' starts as 'fused' and looks at the shards' sampled adjacent pairs per distinct pair: highly repetitive
    sets (>= 8, e.g. the near-clonal skew set: 12) collapse so much in a local build that shipping reduced graphs
    beats shipping records -- measured on 8 B200: 62 vs 96 ms per step -- so those take 'merge' from then on;
    ordinary sets (15,000 synthetic genomes: 5) stay 'fused' (44.4 vs 45.4 ms; 38.0 vs 42.4 ms on 2 GPUs)."""
    return os.environ.get("SEQWIN_DIST", "auto")


class _PreferMerge(Exception):
    """auto mode, raised on every rank alike: the set is repetitive enough for the merge-based build."""


def dist_build(stages: CudaStages, dev_batch, n_records_local: int, k: int, w: int, group=None, rec_base=None,
               host_batch=None, overlap: bool = True, is_targets=None, class_totals=None, inspect=None, ctx=None):
    """Full multi-GPU build of this rank's hash range; returns the sw_graph handle.
    With ctx (routed_context()) the build SEQWIN_DIST names runs (dist_mode(): fused by default, which falls back to
    the merge-based build below when peer memory cannot be set up); without ctx the merge-based build runs.
    With is_targets (bool array, the classes of THIS rank's assemblies) the graph comes back scored:
    every shard counts its own assemblies, the merge adds the counts, and the penalty is finished with
    class_totals = (targets, non-targets) over all ranks (all-reduced here when not given).
    inspect(local, merged_handle), if given, runs after the merge while the shard's own graph is still alive
    (verification hooks of bench.py)."""
    world = dist.get_world_size(group)
    mode = dist_mode()
    if ctx is not None and mode in ("fused", "auto") and not ctx.peer_failed and not (mode == "auto" and ctx.prefer_merge):
        try:
            return dist_build_fused(stages, dev_batch, k, w, ctx, group, host_batch=host_batch, inspect=inspect,
                                    auto=mode == "auto")
        except PeerMemoryUnavailable:      # raised on every rank alike: all of them go on with the merge-based build
            ctx.peer_failed = True
        except _PreferMerge:               # likewise
            ctx.prefer_merge = True
    elif ctx is not None and mode == "routed":
        return dist_build_routed(stages, dev_batch, k, w, ctx, group, host_batch=host_batch, inspect=inspect)
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
    except BaseException:
        # slices of the local graph may still be on the wire (async all_to_all started from the hook):
        # let them land before the buffers they read go back to the allocator
        if local.early is not None:
            for wk in local.early[2]:
                try:
                    wk.wait()
                except Exception:   # noqa: BLE001 - the original error is the one to report
                    pass
        if timed:
            torch.cuda.synchronize()
        stages.free_local(local)
        raise
    if inspect is not None:
        inspect(local, g)
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


# ---- verification helpers (bench.py, tests): torch is the checker here, not the product -------------

_C1, _C2 = -7046029254386353131, -4417276706812531889   # odd 64-bit constants (as signed)
_MIN64 = -(1 << 63)


def _mix(a: torch.Tensor, b: torch.Tensor) -> torch.Tensor:
    x = a * _C1 + b
    x = x ^ (x >> 29)
    return x * _C2


def _as_i64(t: torch.Tensor, cols: int) -> torch.Tensor:
    v = t.view(torch.int64)
    return v.view(-1, cols) if cols > 1 else v


def graph_checksums(kmers_u8: torch.Tensor, nodes_u8: torch.Tensor, edges_u8: torch.Tensor) -> dict:
    """Order-independent checksums and ordering properties of a (shard or merged) graph held as flat byte
    tensors on the device: the multiset of (node hash, record, pos) and of (first, second) x weight must
    survive the exchange + merge; the arrays must come out sorted and tiled."""
    out = {}
    k = _as_i64(kmers_u8, 1)
    n = _as_i64(nodes_u8, 5)
    e = _as_i64(edges_u8, 3)
    out["n_kmers"], out["n_nodes"], out["n_edges"] = int(k.numel()), int(n.shape[0]), int(e.shape[0])
    if n.shape[0]:
        h, start, stop = n[:, 0], n[:, 1], n[:, 2]
        cnt = stop - start
        per_kmer_hash = torch.repeat_interleave(h, cnt)
        out["kmer_sum"] = int(_mix(per_kmer_hash, k).sum()) if per_kmer_hash.numel() == k.numel() else None
        hs = h ^ _MIN64
        out["nodes_sorted"] = bool((hs[1:] > hs[:-1]).all())
        out["nodes_tile"] = bool(start[0] == 0 and stop[-1] == k.numel() and (start[1:] == stop[:-1]).all() and (cnt > 0).all())
        brk = torch.zeros(k.numel(), dtype=torch.bool, device=k.device)
        brk[start] = True
        out["kmers_sorted_in_node"] = bool(((k[1:] > k[:-1]) | brk[1:]).all())
        cls = n[:, 3]                       # n_tar (low 32 bits) | n_neg (high 32 bits)
        out["n_tar_sum"] = int((cls & 0xFFFFFFFF).sum())
        out["n_neg_sum"] = int(((cls >> 32) & 0xFFFFFFFF).sum())
        out["first_hash"], out["last_hash"] = int(h[0]) & (2**64 - 1), int(h[-1]) & (2**64 - 1)
    else:
        out.update({"kmer_sum": 0, "nodes_sorted": True, "nodes_tile": k.numel() == 0, "kmers_sorted_in_node": True,
                    "n_tar_sum": 0, "n_neg_sum": 0, "first_hash": None, "last_hash": None})
    if e.shape[0]:
        f, s2, wgt = e[:, 0], e[:, 1], e[:, 2]
        out["edge_sum"] = int((_mix(f, s2) * wgt).sum())
        fs, ss = f ^ _MIN64, s2 ^ _MIN64
        out["edges_sorted"] = bool(((fs[1:] > fs[:-1]) | ((fs[1:] == fs[:-1]) & (ss[1:] > ss[:-1]))).all() and (fs <= ss).all())
        out["weight_sum"] = int(wgt.sum())
    else:
        out.update({"edge_sum": 0, "edges_sorted": True, "weight_sum": 0})
    return out


def _wrap64(x: int) -> int:
    return x & (2**64 - 1)


def batch_record_offsets(L, batch, n_assemblies: int) -> np.ndarray:
    """[A + 1] record offsets of a packed host batch (records per assembly, cumulative)."""
    from . import _lib
    buf = np.zeros(n_assemblies + 1, dtype=np.int64)
    _lib.check(L.sw_batch_record_offsets(batch, buf.ctypes.data, len(buf)))
    return buf


def full_size_checks(stages: "CudaStages", dev, n_records: int, k: int, w: int, rec_base: int, is_t_local, totals, ctx=None) -> dict:
    """Two more full-size distributed steps.  The merge-based build gives the checksums of every shard's own graph
    (before any exchange) and of every rank's merged hash range; the routed build -- the one that is timed -- those
    of every rank's range as its owner aggregated it.  Summed over the ranks all three must agree: nothing lost,
    nothing counted twice, whichever way the records travelled."""
    from . import _lib
    L = stages.L
    world, rank = dist.get_world_size(), dist.get_rank()
    device = stages.device
    got = {}

    def sums_of(g):
        pk, pn, pe = C.c_void_p(), C.c_void_p(), C.c_void_p()
        _lib.check(L.sw_graph_device_ptrs(g, C.byref(pk), C.byref(pn), C.byref(pe)))
        n_k, n_n, n_e = (L.sw_graph_size(g, i) for i in (_lib.SW_KMERS, _lib.SW_NODES, _lib.SW_EDGES))
        return graph_checksums(_view(pk.value, n_k * KMER_BYTES, device), _view(pn.value, n_n * NODE_BYTES, device),
                               _view(pe.value, n_e * EDGE_BYTES, device))

    def inspect(local, g):
        got["local"] = graph_checksums(local.kmers, local.nodes, local.edges)
        got["merged"] = sums_of(g)

    # the merge-based build is the cross-check (and the source of the shards' own checksums) only at sizes it has
    # been validated at; the decision must be the same on every rank
    biggest = torch.tensor([int(stages.times.n_kmers) if stages.times is not None else 0], dtype=torch.int64, device=device)
    dist.all_reduce(biggest, op=dist.ReduceOp.MAX)
    if int(biggest) > MERGE_MAX_LOCAL_MINIMIZERS and dist_mode() != "merge":
        return {"skipped": f"{int(biggest)} minimizers on one rank: beyond the size the merge-based cross-check is run at"} if rank == 0 else {}
    prev_mode = os.environ.get("SEQWIN_DIST")
    os.environ["SEQWIN_DIST"] = "merge"
    try:
        g = dist_build(stages, dev, n_records, k, w, rec_base=rec_base, is_targets=is_t_local, class_totals=totals,
                       inspect=inspect)
    finally:
        if prev_mode is None:
            del os.environ["SEQWIN_DIST"]
        else:
            os.environ["SEQWIN_DIST"] = prev_mode
    L.sw_graph_free(g)
    if ctx is not None and dist_mode() != "merge" and not ctx.peer_failed and not ctx.prefer_merge:
        g = dist_build(stages, dev, n_records, k, w, ctx=ctx)
        got["routed"] = sums_of(g)
        L.sw_graph_free(g)
    every = [None] * world
    dist.all_gather_object(every, got)
    if rank != 0:
        return {}
    loc, mer = [x["local"] for x in every], [x["merged"] for x in every]
    tot = lambda rows, key: _wrap64(sum(r[key] for r in rows))   # noqa: E731

    def conserved(out):
        return {
            "kmers_conserved": sum(r["n_kmers"] for r in loc) == sum(r["n_kmers"] for r in out),
            "kmer_multiset_conserved": all(r["kmer_sum"] is not None for r in loc + out) and tot(loc, "kmer_sum") == tot(out, "kmer_sum"),
            "edge_weight_multiset_conserved": tot(loc, "edge_sum") == tot(out, "edge_sum") and
                                              sum(r["weight_sum"] for r in loc) == sum(r["weight_sum"] for r in out),
            "class_counts_conserved": sum(r["n_tar_sum"] for r in loc) == sum(r["n_tar_sum"] for r in out) and
                                      sum(r["n_neg_sum"] for r in loc) == sum(r["n_neg_sum"] for r in out),
            "sorted_and_tiled_on_every_rank": all(r["nodes_sorted"] and r["nodes_tile"] and r["kmers_sorted_in_node"] and r["edges_sorted"]
                                                  for r in out),
            "rank_ranges_ascending": all(a["last_hash"] is None or b["first_hash"] is None or a["last_hash"] < b["first_hash"]
                                         for a, b in zip(out[:-1], out[1:])),
        }
    res = conserved(mer)
    final = mer
    if "routed" in every[0]:
        final = [x["routed"] for x in every]
        routed = conserved(final)
        res = {k2: bool(v and routed[k2]) for k2, v in res.items()}
        res["routed_equals_merged_totals"] = all(sum(r[key] for r in final) == sum(r[key] for r in mer)
                                                 for key in ("n_kmers", "n_nodes", "n_edges", "weight_sum", "n_tar_sum", "n_neg_sum"))
    res.update({"n_kmers": sum(r["n_kmers"] for r in final), "n_nodes": sum(r["n_nodes"] for r in final),
                "n_edges": sum(r["n_edges"] for r in final)})
    res["all_ok"] = all(v for k2, v in res.items() if not k2.startswith("n_"))
    return res


# ---- bench driver (bench.py --gpus N under torchrun) ----------------------------------------------

def bench_loop(L, batch, spec, rank: int, world: int, k: int, w: int, steps: int, warmup: int, sampler_cls) -> dict:
    import os
    import time

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
    overlap = os.environ.get("SEQWIN_DIST_OVERLAP", "1") != "0"   # A/B switch for profiles/
    is_t_local = np.ascontiguousarray(is_t[rank * n_asm_local:(rank + 1) * n_asm_local])
    totals = (int(is_t.sum()), int(len(is_t) - is_t.sum()))   # class sizes are input metadata too
    # ... and so are the record offsets / classes of all shards, which the owner of a hash range scores with
    ctx = routed_context(batch_record_offsets(L, batch, n_asm_local), is_t_local)

    def step():
        # build + scoring: records routed to the owners of their hash ranges, aggregated and scored there
        g = dist_build(stages, dev, n_records, k, w, rec_base=rec_base, overlap=overlap, is_targets=is_t_local,
                       class_totals=totals, ctx=ctx)
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
        d["n_kmers_local"], d["n_nodes_local"], d["n_edges_local"] = d["n_kmers"], d["n_nodes"], d["n_edges"]
        d["local_build_ms"] = d["total_ms"]
        d["total_ms"] = float(t)
        d["total_launches"] = d["total_launches"] + stages.merge_launches
        pe = stages.phase_events
        d["phase_local_ms"] = pe[0].elapsed_time(pe[1])
        d["phase_exchange_merge_ms"] = pe[1].elapsed_time(pe[2])
        d["phase_merge_ms"] = stages._merge_events[0].elapsed_time(stages._merge_events[1])
        stage_dicts.append(d)
    clocks = sampler.stop()

    # the same shard without the exchange: what one GPU does on its own with this much input (the
    # equal-work reference point for the scaling figures; scored like the distributed step)
    st = _lib.StageTimes()
    single = []
    for i in range(1 + min(steps, 5)):
        g = C.c_void_p()
        _lib.check(L.sw_dev_build_ex(dev, k, w, rec_base, is_t_local.ctypes.data, len(is_t_local), C.byref(g), C.byref(st)))
        L.sw_graph_free(g)
        if i:
            single.append(st.total_ms)
    sg = torch.tensor([float(np.mean(single))], device=device)
    dist.all_reduce(sg, op=dist.ReduceOp.MAX)

    checks = full_size_checks(stages, dev, n_records, k, w, rec_base, is_t_local, totals, ctx=ctx)
    L.sw_dev_batch_free(dev)

    # end to end: pinned host batch -> H2D -> distributed build -> D2H of this rank's range
    e2e = []
    for i in range(max(1, warmup // 2) + steps):
        dist.barrier()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        g = dist_build(stages, None, n_records, k, w, rec_base=rec_base, host_batch=batch, overlap=overlap,
                       is_targets=is_t_local, class_totals=totals, ctx=ctx)
        t1 = time.perf_counter()
        _lib.check(L.sw_graph_fetch(g))     # this rank's hash range -> pinned host memory
        t2 = time.perf_counter()
        L.sw_graph_free(g)
        torch.cuda.synchronize()
        t = torch.tensor([time.perf_counter() - t0], device=device, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        if i >= max(1, warmup // 2):
            d = stages.times.as_dict()    # this step's own local-build stages (H2D included)
            d["local_build_ms"] = d["total_ms"]
            d["total_ms"] = float(t) * 1e3
            d["fetch_ms"] = (t2 - t1) * 1e3
            e2e.append((float(t), d))

    by_rank = [None] * world
    last = stage_dicts[-1]
    dist.all_gather_object(by_rank, [round(last[n], 2) for n in ("phase_local_ms", "phase_exchange_merge_ms", "phase_merge_ms",
                                                                  "sort_nodes_ms", "nodes_ms", "edges_ms")])
    tot = torch.tensor([n_bases_local] + sizes, dtype=torch.int64, device=device)
    dist.all_reduce(tot, op=dist.ReduceOp.SUM)
    n_bases_total, n_k, n_n, n_e = (int(x) for x in tot)
    for d in stage_dicts:
        d["n_kmers"], d["n_nodes"], d["n_edges"] = n_k, n_n, n_e
    release_peer_buffers(stages, ctx)
    L.sw_set_stream(None)
    return {"stages": stage_dicts, "clocks": clocks, "e2e_runs": e2e, "n_bases_total": n_bases_total,
            "single_gpu": {"ms_per_step": float(sg), "gbp_s_per_gpu": n_bases_local / (float(sg) * 1e-3) / 1e9,
                           "what": "the same shard built and scored by one GPU without the exchange (slowest rank)"},
            "full_size_checks": checks,
            "mode": dist_mode() + (" -> merge (peer memory unavailable)" if ctx.peer_failed else
                                   " -> merge (repetitive set)" if ctx.prefer_merge else
                                   " -> fused" if dist_mode() == "auto" else ""),
            "last_step_ms_by_rank": {"columns": ["local", "exchange_and_owner", "owner", "partition_passes", "nodes", "edges"],
                                     "rows": by_rank}}


def bench_parity(L, parity_batch, ss, rank: int, world: int, per_gpu: int, per_rank: int, k: int, w: int, bench) -> dict | None:
    """Bit-exactness of the NCCL path on hardware: the first `per_rank` genomes of every shard go through
    the same distributed build (local build, all-to-all of hash ranges, owner-side merge, scoring); rank 0
    runs the unmodified reference on the same genomes from FASTA and compares every rank's piece of every
    array.  Rank 0 also builds the subset on its own GPU from the FASTA files (single-GPU == multi-GPU)."""
    import hashlib
    import shutil
    import time

    from . import _lib
    from ._core import NODE_DTYPE
    device = torch.device("cuda", torch.cuda.current_device())
    stages = CudaStages(L, device)
    per_rank = min(per_rank, per_gpu)
    mine = list(range(rank * per_gpu, rank * per_gpu + per_rank))
    is_t_local = np.ascontiguousarray(ss.is_targets[mine], dtype=np.bool_)
    n_records = L.sw_batch_n_records(parity_batch)
    dev = C.c_void_p()
    _lib.check(L.sw_dev_upload(parity_batch, C.byref(dev)))
    ctx = routed_context(batch_record_offsets(L, parity_batch, len(mine)), is_t_local)
    g = dist_build(stages, dev, n_records, k, w, is_targets=is_t_local, ctx=ctx)
    kmers, nodes, edges = export_graph(L, g)
    L.sw_graph_free(g)
    release_peer_buffers(stages, ctx)
    L.sw_dev_batch_free(dev)
    L.sw_set_stream(None)
    counts = [None] * world
    dist.all_gather_object(counts, (len(kmers), len(nodes), len(edges)))
    kbase = sum(c[0] for c in counts[:rank])
    nodes["start"] += np.uintp(kbase)
    nodes["stop"] += np.uintp(kbase)
    sha = lambda a: hashlib.sha256(np.ascontiguousarray(a).view(np.uint8).data).hexdigest()   # noqa: E731
    digests = [None] * world
    dist.all_gather_object(digests, (sha(kmers), sha(nodes), sha(edges)))

    # FASTA of the subset: every rank writes its own genomes into one tmpfs directory
    box = [str(bench.shm_dir("seqwin_b200_parity_"))] if rank == 0 else [None]
    dist.broadcast_object_list(box, src=0)
    from pathlib import Path
    d = Path(box[0])
    nb_local = bench.write_genomes(ss, mine, d, 4)
    nb = torch.tensor([nb_local], dtype=torch.int64, device=device)
    dist.all_reduce(nb)
    dist.barrier()
    res = None
    if rank == 0:
        try:
            subset = [g2 for r in range(world) for g2 in range(r * per_gpu, r * per_gpu + per_rank)]
            paths = [bench.fasta_path(d, g2) for g2 in subset]
            is_t = np.ascontiguousarray(ss.is_targets[subset], dtype=np.bool_)
            ref = bench.run_reference(paths, is_t, k, w, steps=1, warmup=0, keep_graph=True)
            rk, rn, re_, ro = ref["graph"]
            ok, ko, no, eo = True, 0, 0, 0
            pieces = []
            for r in range(world):
                ck, cn, ce = counts[r]
                same = (sha(rk[ko:ko + ck]), sha(rn[no:no + cn]), sha(re_[eo:eo + ce])) == tuple(digests[r])
                pieces.append(bool(same))
                ok = ok and same
                ko, no, eo = ko + ck, no + cn, eo + ce
            ok = ok and (ko, no, eo) == (len(rk), len(rn), len(re_))
            # the same FASTA files through the drop-in entry point on one GPU
            from .graph import KmerGraph, _get_penalty
            t0 = time.perf_counter()
            og = KmerGraph(paths, k, w, n_cpu=os_cpu_count())
            _get_penalty(og.kmers, og.nodes, og.record_offsets, is_t)
            ours_s = time.perf_counter() - t0
            single_ok = bool(np.array_equal(og.kmers, rk) and np.array_equal(og.nodes, rn) and np.array_equal(og.edges, re_)
                             and np.array_equal(og.record_offsets, ro))
            res = {"bit_exact": bool(ok and single_ok), "nccl_pieces_match_reference": pieces,
                   "single_gpu_from_fasta_matches_reference": single_ok, "against": ref["kind"],
                   "what": f"{len(subset)} genomes ({int(nb) / 1e6:.0f} Mbp; the first {per_rank} of every shard) through the same "
                           f"{world}-rank NCCL exchange of minimizer records + aggregation + scoring by the range owners; every rank's slice of kmers / nodes / edges compared by "
                           "SHA-256 with the reference's arrays built from FASTA",
                   "graph": {"n_kmers": len(rk), "n_nodes": len(rn), "n_edges": len(re_)},
                   "sha256": {"kmers": sha(rk), "nodes": sha(rn), "edges": sha(re_)},
                   "cpu_baseline": {"value": int(nb) / ref["seconds"] / 1e9, "unit": "Gbp/s", "cores": ref["cores"], "kind": ref["kind"],
                                    "sample": f"{len(subset)} of the workload's genomes ({int(nb) / 1e6:.0f} Mbp), plain FASTA on tmpfs, "
                                              "build + get_penalty, one pass",
                                    "ours_same_input_from_fasta_gbps": int(nb) / ours_s / 1e9}}
        finally:
            shutil.rmtree(d, ignore_errors=True)
    dist.barrier()
    return res


def os_cpu_count() -> int:
    import os
    return os.cpu_count() or 8
