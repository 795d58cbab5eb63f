#!/usr/bin/env python
"""BASELINE.json configs 3 and 4: parameter sweep over one 4K frame, or a batch of 4K frames, sharded over the GPUs
of one box with NO collective on the data path (one process per GPU; torch.distributed only for the final gather).

  python tools/run_sweep.py --mode sweep                      # 9 (sp, sr) variants of one 3840x2160 frame
  python -m torch.distributed.run --nproc-per-node 8 ... tools/run_sweep.py --mode batch --frames 256

Prints one JSON line on rank 0 with per-unit and aggregate Mpix/s (device-resident, CUDA-event timed per unit)."""
import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

import msegment_b200 as mseg  # noqa: E402

dev = mseg.device
W, H = 3840, 2160


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--mode", choices=["sweep", "batch"], default="sweep")
    ap.add_argument("--frames", type=int, default=256)
    args = ap.parse_args()
    rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
    torch.cuda.set_device(local)
    dist = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    ctx = mseg.Context(local)
    ctx.set_stream(torch.cuda.current_stream().cuda_stream)
    if args.mode == "sweep":
        units = [dict(seed=3, sp=sp, sr=sr) for sp in (5, 10, 20) for sr in (10, 20, 40)]
    else:
        units = [dict(seed=1000 + i, sp=10, sr=10) for i in range(args.frames)]
    mine = units[rank::world]                                   # static round-robin, no communication
    src = torch.empty((H, W, 3), dtype=torch.uint8, device="cuda")
    filt = torch.empty_like(src)
    ren = torch.empty_like(src)
    lab = torch.empty((H, W), dtype=torch.int32, device="cuda")
    nreg = torch.zeros((1,), dtype=torch.int32, device="cuda")
    results = []
    # warm-up: workspace allocation and first-launch overheads are not part of any unit's time
    if mine:
        dev.synth(ctx, src.data_ptr(), 3 * W, W, H, 1)
        dev.segment(ctx, src.data_ptr(), 3 * W, W, H, dev.params(min_size=50, color_dist=10, render_depth=0), filt.data_ptr(),
                    3 * W, lab.data_ptr(), 4 * W, ren.data_ptr(), 3 * W, nreg.data_ptr())
    torch.cuda.synchronize()
    e_all0, e_all1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e_all0.record()
    for u in mine:
        dev.synth(ctx, src.data_ptr(), 3 * W, W, H, u["seed"])
        prm = dev.params(sp=u["sp"], sr=u["sr"], max_level=1, lo_diff=2, min_size=50, color_dist=10, render_depth=0)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        dev.segment(ctx, src.data_ptr(), 3 * W, W, H, prm, filt.data_ptr(), 3 * W, lab.data_ptr(), 4 * W, ren.data_ptr(), 3 * W,
                    nreg.data_ptr())
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1)
        results.append(dict(u, ms=round(ms, 3), mpix_per_s=round(W * H / 1e6 / (ms / 1e3), 1), regions=int(nreg.item())))
    e_all1.record()
    torch.cuda.synchronize()
    total_ms = e_all0.elapsed_time(e_all1)
    gathered = [results]
    tmax = total_ms
    if dist is not None:
        gathered = [None] * world
        dist.all_gather_object(gathered, results)
        t = torch.tensor([total_ms], device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        tmax = float(t.item())
    if rank == 0:
        flat = [r for g in gathered for r in g]
        line = {"config": "%s: %d units of %dx%d over %d GPUs, no collectives" % (args.mode, len(units), W, H, world),
                "n_gpus": world, "units": len(flat), "wall_ms_max_over_ranks": round(tmax, 2),
                "aggregate_mpix_per_s": round(len(flat) * W * H / 1e6 / (tmax / 1e3), 1)}
        if args.mode == "sweep":
            line["variants"] = flat
        else:
            ms = sorted(r["ms"] for r in flat)
            line["per_frame_ms"] = {"min": ms[0], "median": ms[len(ms) // 2], "max": ms[-1]}
        print(json.dumps(line))
    if dist is not None:
        dist.destroy_process_group()
    ctx.close()


if __name__ == "__main__":
    main()
