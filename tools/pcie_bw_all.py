#!/usr/bin/env python
"""Host<->device copy bandwidth of ALL GPUs of the box at once (pinned buffers, one process per GPU under torchrun): the ceiling
of the end-to-end batch figure at N GPUs.  Prints one JSON line on rank 0: per-GPU min and the aggregate, for H2D alone, D2H
alone and both directions together (the e2e pattern)."""
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
    n = 256 << 20
    h, h2 = torch.empty(n, dtype=torch.uint8).pin_memory(), torch.empty(n, dtype=torch.uint8).pin_memory()
    d, d2 = torch.empty(n, dtype=torch.uint8, device="cuda"), torch.empty(n, dtype=torch.uint8, device="cuda")
    s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()

    def timed(fn, reps=8):
        fn()
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        t0 = time.perf_counter()
        for _ in range(reps):
            fn()
        torch.cuda.synchronize()
        dt = (time.perf_counter() - t0) / reps
        t = torch.tensor([dt], device="cuda")
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    def both():
        with torch.cuda.stream(s1):
            d.copy_(h, non_blocking=True)
        with torch.cuda.stream(s2):
            h2.copy_(d2, non_blocking=True)

    res = {"h2d": n / timed(lambda: d.copy_(h, non_blocking=True)) / 1e9,
           "d2h": n / timed(lambda: h.copy_(d, non_blocking=True)) / 1e9,
           "bidir_each": n / timed(both) / 1e9}
    if rank == 0:
        print(json.dumps({"n_gpus": world, "per_gpu_gbs_slowest_rank": {k: round(v, 1) for k, v in res.items()},
                          "aggregate_gbs": {"h2d": round(res["h2d"] * world, 1), "d2h": round(res["d2h"] * world, 1),
                                            "bidir_total": round(2 * res["bidir_each"] * world, 1)},
                          "cpus": os.cpu_count()}))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
