#!/usr/bin/env python
"""Experiment driver: times the mean-shift level kernels (CUDA events via msg_set_profiling) for tile-width /
accumulate variants selected through MSG_TILE_W / MSG_ACC.  Prints one line per (size, variant)."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

import msegment_b200 as mseg  # noqa: E402


def run(w, h, frames, sp, sr):
    dev = mseg.device
    ctx = mseg.Context(0)
    src = torch.empty((frames, h, w, 3), dtype=torch.uint8, device="cuda")
    dst = torch.empty((h, w, 3), dtype=torch.uint8, device="cuda")
    for i in range(frames):
        dev.synth(ctx, src[i].data_ptr(), 3 * w, w, h, 2 + i)
    for rep in range(2):
        if rep == 1:
            ctx.set_profiling(True)
        for i in range(frames):
            dev.meanshift(ctx, src[i].data_ptr(), 3 * w, dst.data_ptr(), 3 * w, w, h, sp, sr, 1)
        ctx.synchronize()
    p = ctx.kernel_profile()
    n = max(1, p["launches"][0])
    ops = 9 * p["tile_tests"][0] + 5 * p["tile_hits"][0]
    out = "L0 %.4f ms (%.2f Tiop/s, %.2f Ttests/s)  L1 %.4f ms  ovf %.4f ms" % (
        p["tile_ms"][0] / n, ops / (p["tile_ms"][0] * 1e-3) / 1e12 if p["tile_ms"][0] else 0,
        p["tile_tests"][0] / (p["tile_ms"][0] * 1e-3) / 1e12 if p["tile_ms"][0] else 0,
        p["tile_ms"][1] / n, (p["overflow_ms"][0] + p["overflow_ms"][1]) / n)
    ctx.close()
    return out


if __name__ == "__main__":
    for (w, h, frames) in ((1920, 1080, 8), (3840, 2160, 4)):
        for tw in ("32", "64"):
            for acc in ("0", "1"):
                os.environ["MSG_TILE_W"] = tw
                os.environ["MSG_ACC"] = acc
                print("%dx%d sp10 sr10 tile_w=%s acc=%s : %s" % (w, h, tw, acc, run(w, h, frames, 10, 10)), flush=True)
    os.environ["MSG_TILE_W"] = "64"
    os.environ["MSG_ACC"] = "1"
    print("3840x2160 sp20 sr40 tile_w=64 acc=1 :", run(3840, 2160, 2, 20, 40), flush=True)
    os.environ["MSG_ACC"] = "0"
    print("3840x2160 sp20 sr40 tile_w=64 acc=0 :", run(3840, 2160, 2, 20, 40), flush=True)
