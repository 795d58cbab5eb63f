#!/usr/bin/env python
"""One exact watershed of one frame (colour-method markers) for ncu: python tools/profile_watershed.py [W H]"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import msegment_b200 as mseg  # noqa: E402


def main():
    w, h = (int(sys.argv[1]), int(sys.argv[2])) if len(sys.argv) > 2 else (960, 540)
    im = mseg.synth_bgr(w, h, 100)
    with mseg.Context(0) as ctx:
        gi = mseg.GpuImgproc(ctx)
        n, mk = gi.colorSeeds(im)
        gi.watershed(im, mk)
        print("watershed done", n, int((mk == -1).sum()))


if __name__ == "__main__":
    main()
