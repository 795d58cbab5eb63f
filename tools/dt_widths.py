#!/usr/bin/env python
"""Kernel time of the chamfer distance transform against the number of column bands (width / 32 at P = 1)."""
import json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import msegment_b200 as mseg
res = {}
h = 2000
with mseg.Context(0) as ctx:
    gi = mseg.GpuImgproc(ctx)
    for w in (32, 64, 128, 256, 512, 1024, 2048, 4096):
        m = (np.random.default_rng(1).random((h, w)) < .9).astype(np.uint8)
        t = []
        for _ in range(3):
            gi.distanceTransform(m)
            t.append(ctx.timings()["filter_ms"])
        res[w] = round(min(t), 3)
print(json.dumps({"h": h, "kernel_ms_by_width": res}))
