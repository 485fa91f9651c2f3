#!/usr/bin/env python3
"""Host-link ceiling for the e2e arm at N GPUs: H2D alone, D2H alone and both directions at once, with
k = 1, 2, 4, ... N GPUs of the box active CONCURRENTLY (pinned host memory, one copy stream per direction per GPU).

    python -m torch.distributed.run --nnodes=1 --nproc-per-node=8 --master-addr 127.0.0.1 tools/pcie_probe_multi.py > profiles/rNN_pcie_n8.txt

Every rank allocates its buffers; for each k the ranks < k copy between barriers, the others idle.  Rank 0 prints one line per
(k, direction): per-GPU and aggregate GB/s (aggregate = bytes of all active ranks / slowest rank's time)."""
import json
import os
import time

import torch
import torch.distributed as dist


def main():
    rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    n = 1 << 28                                         # 1 GiB of float32 per direction
    h_in = torch.empty(n, dtype=torch.float32).pin_memory()
    h_out = torch.empty(n, dtype=torch.float32).pin_memory()
    h_in.fill_(1.0)
    d_a = torch.empty(n, dtype=torch.float32, device="cuda")
    d_b = torch.ones(n, dtype=torch.float32, device="cuda")
    s_up, s_down = torch.cuda.Stream(), torch.cuda.Stream()
    gb = n * 4 / 1e9
    flag = torch.zeros(1, device="cuda")

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.all_reduce(flag)
            torch.cuda.synchronize()

    def up():
        with torch.cuda.stream(s_up):
            d_a.copy_(h_in, non_blocking=True)

    def down():
        with torch.cuda.stream(s_down):
            h_out.copy_(d_b, non_blocking=True)

    def both():
        up()
        down()

    ks = [k for k in (1, 2, 4, 8, 16) if k <= world]
    if rank == 0:
        print(json.dumps({"world": world, "host_cpus": len(os.sched_getaffinity(0)), "bytes_per_copy": n * 4}), flush=True)
    for k in ks:
        for name, fn, nbytes in (("h2d", up, gb), ("d2h", down, gb), ("both", both, 2 * gb)):
            best = float("inf")
            for _ in range(3):
                barrier()
                t0 = time.perf_counter()
                if rank < k:
                    fn()
                torch.cuda.synchronize()
                dt = time.perf_counter() - t0
                t = torch.tensor([dt if rank < k else 0.0], device="cuda", dtype=torch.float64)
                if world > 1:
                    dist.all_reduce(t, op=dist.ReduceOp.MAX)
                best = min(best, float(t.item()))
            if rank == 0:
                print(json.dumps({"active_gpus": k, "direction": name, "slowest_rank_s": round(best, 5), "GBps_per_gpu": round(nbytes / best, 1),
                                  "GBps_aggregate": round(k * nbytes / best, 1)}), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
