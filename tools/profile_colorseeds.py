#!/usr/bin/env python
"""Two passes of the colour-method marker generator on one 1080p frame for ncu (launch list / --set full capture of the
distance-transform kernel).  Numbers printed by a run under ncu are not bench values."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import msegment_b200 as mseg  # noqa: E402


def main():
    w, h = (int(sys.argv[1]), int(sys.argv[2])) if len(sys.argv) > 2 else (1920, 1080)
    im = mseg.synth_bgr(w, h, 2)
    with mseg.Context(0) as ctx:
        gi = mseg.GpuImgproc(ctx)
        for rep in range(2):
            before = ctx.stats()["kernel_launches"]
            n, _ = gi.colorSeeds(im)
            print("pass", rep, "contours", n, "launches", ctx.stats()["kernel_launches"] - before)


if __name__ == "__main__":
    main()
