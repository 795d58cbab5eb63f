#!/usr/bin/env python
"""Small end-to-end exercise of every kernel family for compute-sanitizer (memcheck / racecheck): ragged sizes, all stages once.
Usage: compute-sanitizer --tool memcheck python tools/sanitize_small.py"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np  # noqa: E402

import msegment_b200 as mseg  # noqa: E402


def main():
    with mseg.Context(0) as ctx:
        gi = mseg.GpuImgproc(ctx)
        for (w, h, seed) in ((131, 97, 1), (257, 64, 2)):
            im = mseg.synth_bgr(w, h, seed)
            out = gi.segment(im, 6, 9, 1, loDiff=2, minSize=20, colorDist=8, want=("filtered", "labels", "rendered"))
            ctx.set_option("merge_small_max", 0)
            gi.segment(im, 6, 9, 1, loDiff=2, minSize=20, colorDist=8)
            ctx.set_option("merge_small_max", -1)
            gi.labelRegions(out["filtered"], 2, 2, 8)
            n, mk = gi.colorSeeds(im)
            gi.watershed(im, mk.copy())
            gi.shapeSeeds(im)
            gray = gi.cvtColorBGR2GRAY(im)
            gi.bilateralFilter(gray, 5, 10, 10)
            t, bw = gi.threshold(gray, 40, 255, 8)
            gi.distanceTransform(bw)
            ctx.set_option("dt_fixed", 1)
            gi.distanceTransform(bw)
            ctx.set_option("dt_fixed", 0)
            gi.connectedComponents(bw, 8)
        print("sanitize_small done", ctx.stats()["kernel_launches"], "launches")


if __name__ == "__main__":
    main()
