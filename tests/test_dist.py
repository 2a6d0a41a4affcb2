"""N>1 host-side logic over gloo (world_size 2, CPU): shard bookkeeping, hash-range cuts, the byte
all-to-all and the rank-order concatenation of seqwin_b200.dist, with a numpy stand-in for the CUDA
stages (the stand-in lives here, in tests/, and uses the oracle for the per-shard build)."""
from __future__ import annotations

import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from tests.helpers import assert_graph_equal


class NumpyStages:
    """Test double for seqwin_b200.dist.CudaStages."""
    device = torch.device("cpu")

    def local_build(self, paths, k, w, rec_base, world, host_batch=None, on_nodes=None, is_targets=None):
        from oracle import oracle as O
        from seqwin_b200.dist import LocalGraph
        kmers, nodes, edges, offsets, ids = O._build_native(paths, k, w)
        kmers = kmers.copy()
        if is_targets is not None:   # shard counts: distinct assemblies of this shard per node, by class
            asm_of = np.repeat(np.arange(len(offsets) - 1), np.diff(offsets.astype(np.int64)))
            nodes = nodes.copy()
            for i in range(len(nodes)):
                a = np.unique(asm_of[kmers["record_idx"][int(nodes["start"][i]):int(nodes["stop"][i])]])
                nodes["n_tar"][i] = int(np.count_nonzero(is_targets[a]))
                nodes["n_neg"][i] = len(a) - int(nodes["n_tar"][i])
        kmers["record_idx"] += np.uint32(rec_base)
        bounds = [(i << 64) // world for i in range(world)]
        ebounds = [int((1.0 - (1.0 - i / world) ** 0.5) * 2.0 ** 64) for i in range(world)]   # balanced for min(u, v)
        ns = np.array([int(np.searchsorted(nodes["hash"], np.uint64(b), "left")) if b else 0 for b in bounds] + [len(nodes)],
                      dtype=np.uint64)
        es = np.array([int(np.searchsorted(edges["first"], np.uint64(b), "left")) if b else 0 for b in ebounds] + [len(edges)],
                      dtype=np.uint64)
        ks = np.array([int(nodes["start"][int(i)]) if int(i) < len(nodes) else len(kmers) for i in ns], dtype=np.uint64)
        as_t = lambda a: torch.from_numpy(np.frombuffer(a.tobytes(), dtype=np.uint8).copy())  # noqa: E731
        local = LocalGraph(as_t(kmers), as_t(nodes), as_t(edges), ns, ks, es)
        if on_nodes is not None:   # the nodes-ready hook of the CUDA build
            local.early = on_nodes(local.nodes, local.kmers, ns, ks)
        return local

    def free_local(self, local):
        pass

    def merge(self, nodes_t, node_counts, kmers_t, kmer_counts, kmer_base, edges_t, edge_counts):
        from oracle.oracle import EDGE_DTYPE, KMER_DTYPE, NODE_DTYPE
        nodes = np.frombuffer(nodes_t.numpy().tobytes(), dtype=NODE_DTYPE)
        kmers = np.frombuffer(kmers_t.numpy().tobytes(), dtype=KMER_DTYPE)
        edges = np.frombuffer(edges_t.numpy().tobytes(), dtype=EDGE_DTYPE)
        node_off = np.concatenate([[0], np.cumsum(node_counts)]).astype(np.int64)
        kmer_off = np.concatenate([[0], np.cumsum(kmer_counts)]).astype(np.int64)
        abs_start = np.empty(len(nodes), dtype=np.int64)
        for s in range(len(node_counts)):
            sl = slice(node_off[s], node_off[s + 1])
            abs_start[sl] = kmer_off[s] + (nodes["start"][sl].astype(np.int64) - int(kmer_base[s]))
        order = np.argsort(nodes["hash"], kind="stable")
        out_k, out_n = [], []
        for j in order:
            cnt = int(nodes["stop"][j] - nodes["start"][j])
            seg = kmers[abs_start[j]:abs_start[j] + cnt]
            if out_n and out_n[-1][0] == nodes["hash"][j]:
                out_n[-1][2] += cnt
                out_n[-1][3] += int(nodes["n_tar"][j])   # an assembly lives on one rank: counts add
                out_n[-1][4] += int(nodes["n_neg"][j])
            else:
                pos = sum(len(x) for x in out_k)
                out_n.append([nodes["hash"][j], pos, pos + cnt, int(nodes["n_tar"][j]), int(nodes["n_neg"][j])])
            out_k.append(seg)
        mk = np.concatenate(out_k) if out_k else np.empty(0, KMER_DTYPE)
        mn = np.zeros(len(out_n), dtype=NODE_DTYPE)
        for i, (h, a, b, nt, nn) in enumerate(out_n):
            mn[i] = (h, a, b, nt, nn, 0.0)
        eo = np.lexsort((edges["second"], edges["first"]))
        me = []
        for j in eo:
            if me and me[-1][0] == edges["first"][j] and me[-1][1] == edges["second"][j]:
                me[-1][2] += int(edges["weight"][j])
            else:
                me.append([edges["first"][j], edges["second"][j], int(edges["weight"][j])])
        mee = np.zeros(len(me), dtype=EDGE_DTYPE)
        for i, (a, b, c) in enumerate(me):
            mee[i] = (a, b, c)
        return mk, mn, mee


    # ---- routed build: records to the owners of their hash ranges, aggregated there ----
    def sketch_route(self, paths, k, w, rec_base, host_batch=None):
        from oracle import oracle as O
        from seqwin_b200.dist import Routed
        keys, vals = [], []
        rec = rec_base
        for path in paths:
            for seq in _fasta_records(path):
                h1, pos = O.minimize(seq, k, w)
                keys.append(h1.astype(np.uint64))
                vals.append(pos.astype(np.uint64) | (np.uint64(rec) << np.uint64(32)))
                rec += 1
        keys = np.concatenate(keys) if keys else np.empty(0, np.uint64)
        vals = np.concatenate(vals) if vals else np.empty(0, np.uint64)
        n = len(keys)
        prev, nxt = np.zeros(n, np.uint64), np.zeros(n, np.uint64)
        if n > 1:   # csrc/nbr.cuh: the adjacent pair belongs to the smaller hash, the earlier item on a tie
            same = (vals[:-1] >> np.uint64(32)) == (vals[1:] >> np.uint64(32))
            own_next = same & (keys[1:] >= keys[:-1])
            own_prev = same & (keys[:-1] > keys[1:])
            nxt[:-1][own_next] = keys[1:][own_next]
            prev[1:][own_prev] = keys[:-1][own_prev]
        order = np.argsort(keys >> np.uint64(56), kind="stable")
        keys, vals, prev, nxt = keys[order], vals[order], prev[order], nxt[order]
        byte_off = np.searchsorted(keys >> np.uint64(56), np.arange(257, dtype=np.uint64), "left").astype(np.uint64)
        as_t = lambda a: torch.from_numpy(np.frombuffer(a.tobytes(), dtype=np.uint8).copy())  # noqa: E731
        return Routed([as_t(keys), as_t(vals), as_t(prev), as_t(nxt)], byte_off, 1.0)

    def free_routed(self, routed):
        pass

    def aggregate(self, recv, n, byte_lo, byte_hi, record_offsets, is_targets, pairs_per_edge):
        from oracle.oracle import EDGE_DTYPE, KMER_DTYPE, NODE_DTYPE
        keys, vals, prev, nxt = (np.frombuffer(t.numpy().tobytes(), dtype=np.uint64) for t in recv)
        assert len(keys) == n and (n == 0 or (int(keys.min() >> np.uint64(56)) >= byte_lo and int(keys.max() >> np.uint64(56)) < byte_hi))
        order = np.argsort(keys, kind="stable")
        kmers = np.zeros(n, dtype=KMER_DTYPE)
        kmers["pos"] = (vals[order] & np.uint64(0xFFFFFFFF)).astype(np.uint32)
        kmers["record_idx"] = (vals[order] >> np.uint64(32)).astype(np.uint32)
        uk, first, cnt = np.unique(keys[order], return_index=True, return_counts=True)
        nodes = np.zeros(len(uk), dtype=NODE_DTYPE)
        nodes["hash"], nodes["start"], nodes["stop"] = uk, first, first + cnt
        asm_of_rec = np.searchsorted(record_offsets.astype(np.int64), np.arange(int(record_offsets[-1])), "right") - 1
        if is_targets is not None:
            n_t, n_n = int(is_targets.sum()), int(len(is_targets) - is_targets.sum())
            for i in range(len(nodes)):
                a = np.unique(asm_of_rec[kmers["record_idx"][first[i]:first[i] + cnt[i]]])
                nodes["n_tar"][i] = int(np.count_nonzero(is_targets[a]))
                nodes["n_neg"][i] = len(a) - int(nodes["n_tar"][i])
            ft = nodes["n_tar"] * (1.0 / n_t)
            fn = nodes["n_neg"] * (1.0 / n_n)
            nodes["penalty"] = np.sqrt((1.0 - ft) * (1.0 - ft) + fn * fn)
        asm = asm_of_rec[(vals >> np.uint64(32)).astype(np.int64)] if n else np.empty(0, np.int64)
        trip = set()
        for k_, s_, a_ in zip(np.concatenate([keys[prev != 0], keys[nxt != 0]]), np.concatenate([prev[prev != 0], nxt[nxt != 0]]),
                              np.concatenate([asm[prev != 0], asm[nxt != 0]])):
            trip.add((int(k_), int(s_), int(a_)))
        weight = {}
        for f_, s_, _ in trip:
            weight[(f_, s_)] = weight.get((f_, s_), 0) + 1
        edges = np.zeros(len(weight), dtype=EDGE_DTYPE)
        for i, ((f_, s_), w_) in enumerate(sorted(weight.items())):
            edges[i] = (f_, s_, w_)
        return kmers, nodes, edges

    def finish_penalty(self, g, class_totals):
        nodes = g[1]
        ft = nodes["n_tar"] * (1.0 / class_totals[0])
        fn = nodes["n_neg"] * (1.0 / class_totals[1])
        nodes["penalty"] = np.sqrt((1.0 - ft) * (1.0 - ft) + fn * fn)


def _fasta_records(path):
    """Sequences of a plain FASTA file (test sets: no blank lines inside records)."""
    seq = None
    with open(path) as f:
        for line in f:
            line = line.strip()
            if line.startswith(">"):
                if seq is not None:
                    yield "".join(seq)
                seq = []
            elif seq is not None:
                seq.append(line)
    if seq is not None:
        yield "".join(seq)


def _routed_worker(rank, world, port, paths, is_t, k, w, out_path):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    os.environ["SEQWIN_DIST"] = "routed"
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from oracle import oracle as O
        from seqwin_b200 import dist as swd
        per = (len(paths) + world - 1) // world
        mine = paths[rank * per:(rank + 1) * per]
        mine_t = None if is_t is None else np.asarray(is_t[rank * per:(rank + 1) * per], dtype=np.bool_)
        offsets_local = O._build_native(mine, k, w)[3]
        ctx = swd.routed_context(offsets_local, mine_t)
        merged = swd.dist_build(NumpyStages(), mine, int(offsets_local[-1]), k, w, is_targets=mine_t, ctx=ctx)
        full = swd.gather_graph(merged)
        if rank == 0:
            np.savez(out_path, kmers=full[0], nodes=full[1], edges=full[2], bounds=ctx.bounds, rec_base=np.array([ctx.rec_base]))
    finally:
        dist.destroy_process_group()


def _scored_worker(rank, world, port, paths, is_t, k, w, out_path):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from oracle import oracle as O
        from seqwin_b200 import dist as swd
        per = (len(paths) + world - 1) // world
        mine = paths[rank * per:(rank + 1) * per]
        mine_t = np.asarray(is_t[rank * per:(rank + 1) * per], dtype=np.bool_)
        n_rec_local = int(O._build_native(mine, k, w)[3][-1])
        # dist_build end to end (hook, class totals by all-reduce, merge, finish) with the numpy stand-in
        merged = swd.dist_build(NumpyStages(), mine, n_rec_local, k, w, is_targets=mine_t)
        full = swd.gather_graph(merged)
        if rank == 0:
            np.savez(out_path, kmers=full[0], nodes=full[1], edges=full[2])
    finally:
        dist.destroy_process_group()


def _worker(rank, world, port, paths, k, w, out_path):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from oracle import oracle as O
        from seqwin_b200 import dist as swd
        per = (len(paths) + world - 1) // world
        mine = paths[rank * per:(rank + 1) * per]
        n_rec_local = int(O._build_native(mine, k, w)[3][-1])
        stages = NumpyStages()
        rec_base, total = swd.record_base(n_rec_local, torch.device("cpu"))
        local = stages.local_build(mine, k, w, rec_base, world)
        merged = swd.exchange_and_merge(stages, local)
        full = swd.gather_graph(merged)
        if rank == 0:
            np.savez(out_path, kmers=full[0], nodes=full[1], edges=full[2], total=np.array([total]))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("kw", [(17, 10), (21, 46)], ids=lambda kw: f"k{kw[0]}w{kw[1]}")
def test_two_rank_exchange_matches_single_graph(synth_sets, tmp_path, kw):
    from oracle import oracle as O
    paths, _ = synth_sets["synth_small"]
    paths = [str(p) for p in paths]
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    out = tmp_path / "merged.npz"
    mp.spawn(_worker, args=(2, port, paths, kw[0], kw[1], str(out)), nprocs=2, join=True)
    got = np.load(out)
    want = O._build_native(paths, *kw)
    assert int(got["total"][0]) == int(want[3][-1])
    assert_graph_equal((got["kmers"], got["nodes"], got["edges"], want[3]), want, f"2 ranks {kw}")


@pytest.mark.parametrize("world", [2, 3])
def test_dist_build_scored_flow_over_gloo(synth_sets, tmp_path, world):
    """seqwin_b200.dist.dist_build with scoring, host logic only: shards that hold one class only,
    class totals by all-reduce, counts summed in the merge, penalty finished last == the oracle's
    build + get_penalty on all assemblies."""
    from oracle import oracle as O
    k, w = 17, 10
    paths, is_t = synth_sets["synth_small"]
    paths = [str(p) for p in paths]
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    out = tmp_path / "merged.npz"
    mp.spawn(_scored_worker, args=(world, port, paths, list(map(bool, is_t)), k, w, str(out)), nprocs=world, join=True)
    got = np.load(out)
    kmers, nodes, edges, offsets, _ = O._build_native(paths, k, w)
    O._get_penalty_native(kmers, nodes, offsets, np.asarray(is_t, dtype=np.bool_))
    assert np.array_equal(got["kmers"], kmers) and np.array_equal(got["edges"], edges)
    assert np.array_equal(got["nodes"], nodes)


@pytest.mark.parametrize("world", [2, 3])
@pytest.mark.parametrize("scored", [False, True], ids=["build", "scored"])
def test_routed_build_over_gloo(synth_sets, tmp_path, world, scored):
    """seqwin_b200.dist routed flow, host logic only: global record offsets / classes gathered once, records cut at
    the top-byte boundaries of the ranks' hash ranges, exchanged, aggregated by the owners; the ranks' graphs
    concatenated in rank order == the oracle's graph (and its get_penalty) on all assemblies."""
    from oracle import oracle as O
    k, w = 17, 10
    paths, is_t = synth_sets["synth_small"]
    paths = [str(p) for p in paths]
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    out = tmp_path / "routed.npz"
    mp.spawn(_routed_worker, args=(world, port, paths, list(map(bool, is_t)) if scored else None, k, w, str(out)),
             nprocs=world, join=True)
    got = np.load(out)
    kmers, nodes, edges, offsets, _ = O._build_native(paths, k, w)
    if scored:
        O._get_penalty_native(kmers, nodes, offsets, np.asarray(is_t, dtype=np.bool_))
    assert np.array_equal(got["kmers"], kmers) and np.array_equal(got["edges"], edges)
    assert np.array_equal(got["nodes"], nodes)
    b = got["bounds"]
    assert b[0] == 0 and b[-1] == 256 and np.all(np.diff(b) > 0)


def test_range_bounds():
    from seqwin_b200.dist import range_bounds
    for world in (1, 2, 3, 4, 8, 64, 256):
        b = range_bounds(world)
        assert len(b) == world + 1 and b[0] == 0 and b[-1] == 256 and np.all(np.diff(b) >= 1)
    b8 = range_bounds(8)
    assert b8[1] < 32 and 256 - b8[-2] > 32      # low ranges own more pairs, so they are narrower
