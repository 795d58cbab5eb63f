#!/usr/bin/env python
"""Small, single-stream run of the bench workload for ncu (launch list / --set full capture of the top kernel).
Usage: python tools/profile_step.py [frames] [W H]   -- numbers printed by a run under ncu are not bench values."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

import bench  # noqa: E402
import msegment_b200 as mseg  # noqa: E402


def main():
    frames = int(sys.argv[1]) if len(sys.argv) > 1 else 2
    w, h = (int(sys.argv[2]), int(sys.argv[3])) if len(sys.argv) > 3 else (bench.W, bench.H)
    dev = mseg.device
    ctx = mseg.Context(0)
    prm = dev.params(**bench.PARAMS, render_depth=-1)
    src = torch.empty((frames, h, w, 3), dtype=torch.uint8, device="cuda")
    filt = torch.empty((h, w, 3), dtype=torch.uint8, device="cuda")
    lab = torch.empty((h, w), dtype=torch.int32, device="cuda")
    for i in range(frames):
        dev.synth(ctx, src[i].data_ptr(), 3 * w, w, h, bench.SEED0 + i)
    for rep in range(2):       # first pass warms allocations; ncu skips it with -s
        for i in range(frames):
            dev.segment(ctx, src[i].data_ptr(), 3 * w, w, h, prm, filt.data_ptr(), 3 * w, lab.data_ptr(), 4 * w)
        ctx.synchronize()
        if rep == 0:
            print("launches per pass:", ctx.stats()["kernel_launches"] - frames)
    print("done", ctx.stats())


if __name__ == "__main__":
    main()
