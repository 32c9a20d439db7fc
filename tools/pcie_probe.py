#!/usr/bin/env python
"""Host<->device copy bandwidth per rank with ALL ranks copying at once (what ceiling does the host give
the end-to-end path when 8 ranks share it?).  Launch like bench.py:
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P tools/pcie_probe.py
Every rank pins `--mb` MiB, then (after a barrier) times H2D alone, D2H alone and both directions together on two streams;
rank 0 prints one JSON line with per-rank and aggregate GB/s."""
import argparse
import json
import os
import time

import torch
import torch.distributed as dist


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--mb", type=int, default=512)
    ap.add_argument("--reps", type=int, default=5)
    a = ap.parse_args()
    rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    n = a.mb << 20
    h_up = torch.empty(n, dtype=torch.uint8).pin_memory()
    h_dn = torch.empty(n, dtype=torch.uint8).pin_memory()
    d_up = torch.empty(n, dtype=torch.uint8, device="cuda")
    d_dn = torch.empty(n, dtype=torch.uint8, device="cuda")
    s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize()

    def timed(fn):
        fn(); barrier()
        t0 = time.perf_counter()
        for _ in range(a.reps):
            fn()
        torch.cuda.synchronize()
        dt = (time.perf_counter() - t0) / a.reps
        barrier()
        return dt

    def up():
        with torch.cuda.stream(s1):
            d_up.copy_(h_up, non_blocking=True)

    def down():
        with torch.cuda.stream(s2):
            h_dn.copy_(d_dn, non_blocking=True)

    def both():
        up(); down()

    res = {}
    for name, fn, nbytes in (("h2d", up, n), ("d2h", down, n), ("both", both, 2 * n)):
        dt = timed(fn)
        t = torch.tensor([nbytes / dt / 1e9], dtype=torch.float64, device="cuda")
        if world > 1:
            lst = [torch.zeros_like(t) for _ in range(world)]
            dist.all_gather(lst, t)
            vals = [float(x.item()) for x in lst]
        else:
            vals = [float(t.item())]
        res[name] = {"per_rank_gbs": [round(v, 2) for v in vals], "aggregate_gbs": round(sum(vals), 1), "min_gbs": round(min(vals), 2)}
    if rank == 0:
        try:
            aff = sorted(os.sched_getaffinity(0))
            aff = f"{aff[0]}-{aff[-1]} ({len(aff)} cpus)"
        except OSError:
            aff = None
        print(json.dumps({"ranks": world, "mib_per_copy": a.mb, "pinned": True, "cpu_affinity_rank0": aff, **res}), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
