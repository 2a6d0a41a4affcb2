"""Multi-rank build on real hardware against the oracle, all three ways seqwin_b200.dist knows: fused (the routing
pass scatters the records straight into the range owners' memory, mapped through CUDA IPC), routed (records
partitioned on the top byte of h1, sent to the owners of their hash ranges through NCCL, aggregated there with the
single-GPU bucket kernels) and merge-based (local build with a record base, hash-range split, merge kernels of
csrc/dist.cu).  Uses NCCL when one GPU per rank is visible; on a single GPU the ranks share cuda:0 and exchange
through gloo (the device kernels exercised are the same)."""
from __future__ import annotations

import ctypes as C
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from tests.helpers import assert_graph_equal, check_graph_invariants

pytestmark = pytest.mark.gpu


def _worker(rank, world, port, paths, k, w, out_path, use_nccl, overlap=True, is_targets=None, mode="routed"):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    os.environ["SEQWIN_DIST"] = mode
    dev_index = rank if use_nccl else 0
    torch.cuda.set_device(dev_index)
    if use_nccl:
        dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", dev_index))
    else:
        dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from seqwin_b200 import _lib
        from seqwin_b200 import dist as swd
        L = _lib.lib()
        device = torch.device("cuda", dev_index)
        stages = swd.CudaStages(L, device)
        per = (len(paths) + world - 1) // world
        mine = paths[rank * per:(rank + 1) * per]
        arr = (C.c_char_p * max(1, len(mine)))(*[p.encode() for p in mine])
        b, d = C.c_void_p(), C.c_void_p()
        _lib.check(L.sw_batch_from_fasta(arr, len(mine), 2, C.byref(b)))
        _lib.check(L.sw_dev_upload(b, C.byref(d)))
        mine_t = None if is_targets is None else np.asarray(is_targets[rank * per:(rank + 1) * per], dtype=np.bool_)
        ctx = None
        if mode != "merge":
            ctx = swd.routed_context(swd.batch_record_offsets(L, b, len(mine)), mine_t)
        g = swd.dist_build(stages, d, L.sw_batch_n_records(b), k, w, overlap=overlap, is_targets=mine_t, ctx=ctx)
        parts = swd.export_graph(L, g)
        L.sw_graph_free(g)
        if mode in ("fused", "auto"):   # a second build reuses the mapped receive arrays (auto: or the choice it made)
            g = swd.dist_build(stages, d, L.sw_batch_n_records(b), k, w, overlap=overlap, is_targets=mine_t, ctx=ctx)
            again = swd.export_graph(L, g)
            L.sw_graph_free(g)
            assert all(np.array_equal(x, y) for x, y in zip(parts, again))
            swd.release_peer_buffers(stages, ctx)
        L.sw_dev_batch_free(d)
        L.sw_batch_free(b)
        full = swd.gather_graph(parts)
        if rank == 0:
            np.savez(out_path, kmers=full[0], nodes=full[1], edges=full[2])
    finally:
        dist.destroy_process_group()


def _backend(world: int) -> bool:
    """NCCL over NVLink whenever the box shows one GPU per rank; the 4- and 8-rank variants only run then."""
    use_nccl = torch.cuda.device_count() >= world
    if world > 3 and not use_nccl:
        pytest.skip(f"{world} ranks need {world} GPUs (NCCL); {torch.cuda.device_count()} visible")
    print(f"[dist test] world {world}: backend {'nccl' if use_nccl else 'gloo (ranks share cuda:0)'}")
    return use_nccl


@pytest.mark.parametrize("mode", ["fused", "routed", "merge"])
@pytest.mark.parametrize("world", [2, 3, 4, 8])
@pytest.mark.parametrize("kw", [(21, 200), (17, 10)], ids=lambda kw: f"k{kw[0]}w{kw[1]}")
def test_multi_rank_build_matches_oracle(synth_sets, tmp_path, kw, world, mode):
    from oracle import oracle as O
    use_nccl = _backend(world)
    paths, _ = synth_sets["synth_medium"]
    paths = [str(p) for p in paths]
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    out = tmp_path / "merged.npz"
    mp.spawn(_worker, args=(world, port, paths, kw[0], kw[1], str(out), use_nccl, True, None, mode), nprocs=world, join=True)
    got = np.load(out)
    want = O._build_native(paths, *kw)
    assert_graph_equal((got["kmers"], got["nodes"], got["edges"], want[3]), want, f"{world} ranks {kw} {mode}")
    check_graph_invariants(got["kmers"], got["nodes"], got["edges"], want[3])


def test_multi_rank_build_without_overlap(synth_sets, tmp_path):
    """overlap=False: the whole exchange after the local build (the path the CPU gloo tests model)."""
    from oracle import oracle as O
    world, kw = 2, (21, 200)
    use_nccl = torch.cuda.device_count() >= world
    paths = [str(p) for p in synth_sets["synth_medium"][0]]
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    out = tmp_path / "merged.npz"
    mp.spawn(_worker, args=(world, port, paths, kw[0], kw[1], str(out), use_nccl, False, None, "merge"), nprocs=world, join=True)
    got = np.load(out)
    want = O._build_native(paths, *kw)
    assert_graph_equal((got["kmers"], got["nodes"], got["edges"], want[3]), want, "2 ranks, no overlap")


@pytest.mark.parametrize("mode", ["fused", "routed", "merge"])
def test_more_ranks_than_assemblies(synth_sets, tmp_path, mode):
    """A rank with an empty shard (no records to route; in the merge-based build it never reaches the nodes-ready
    hook) must still join every collective."""
    from oracle import oracle as O
    world, kw = 3, (21, 50)
    use_nccl = torch.cuda.device_count() >= world
    paths = [str(p) for p in synth_sets["synth_medium"][0]][:2]
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    out = tmp_path / "merged.npz"
    mp.spawn(_worker, args=(world, port, paths, kw[0], kw[1], str(out), use_nccl, True, None, mode), nprocs=world, join=True)
    got = np.load(out)
    want = O._build_native(paths, *kw)
    assert_graph_equal((got["kmers"], got["nodes"], got["edges"], want[3]), want, f"3 ranks, 2 assemblies, {mode}")


@pytest.mark.parametrize("mode", ["auto", "fused", "routed", "merge"])
@pytest.mark.parametrize("world", [2, 3, 4, 8])
def test_multi_rank_scored_build(synth_sets, tmp_path, world, mode):
    """Scoring across ranks.  Routed: the owner of a hash range scores the records of every shard with the global
    record offsets and classes.  Merge-based: every shard counts its own assemblies (targets first, so some shards
    hold one class only), the merge adds the counts, the penalty is finished with the global class sizes.
    Must equal build + get_penalty of the oracle on all assemblies."""
    from oracle import oracle as O
    kw = (21, 200)
    use_nccl = _backend(world)
    paths, is_t = synth_sets["synth_medium"]
    paths = [str(p) for p in paths]
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    out = tmp_path / "merged.npz"
    mp.spawn(_worker, args=(world, port, paths, kw[0], kw[1], str(out), use_nccl, True, list(map(bool, is_t)), mode),
             nprocs=world, join=True)
    got = np.load(out)
    kmers, nodes, edges, offsets, _ = O._build_native(paths, *kw)
    O._get_penalty_native(kmers, nodes, offsets, np.asarray(is_t, dtype=np.bool_))
    assert np.array_equal(got["kmers"], kmers) and np.array_equal(got["edges"], edges)
    assert np.array_equal(got["nodes"], nodes), "scored nodes differ"
