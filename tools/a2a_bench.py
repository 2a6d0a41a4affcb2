#!/usr/bin/env python3
"""all_to_all_single bandwidth between the GPUs of one box (torchrun): what the record exchange of the routed
multi-GPU build can expect.  Prints GB/s sent per GPU for a few message sizes."""
import os
import time

import torch
import torch.distributed as dist


def main():
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    torch.cuda.set_device(int(os.environ["LOCAL_RANK"]))
    dist.init_process_group("nccl", device_id=torch.device("cuda", torch.cuda.current_device()))
    for mb in (64, 512, 3072):
        n = mb * (1 << 20) // world * world
        src = torch.empty(n, dtype=torch.uint8, device="cuda")
        dst = torch.empty(n, dtype=torch.uint8, device="cuda")
        for _ in range(3):
            dist.all_to_all_single(dst, src)
        torch.cuda.synchronize()
        dist.barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(5):
            dist.all_to_all_single(dst, src)
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 5
        if rank == 0:
            print(f"world {world} {mb} MB per GPU: {ms:.2f} ms, {n * (world - 1) / world / ms / 1e6:.0f} GB/s sent per GPU "
                  f"(NCCL_MIN/MAX_P2P_NCHANNELS={os.environ.get('NCCL_MIN_P2P_NCHANNELS')}/{os.environ.get('NCCL_MAX_P2P_NCHANNELS')})", flush=True)
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
