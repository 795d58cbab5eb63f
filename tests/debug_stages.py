#!/usr/bin/env python
"""GPU debugging aid (lives under tests/ because it uses the oracle): runs the mean-shift pipeline stage by stage and reports where it first departs from the
oracle (pyramid planes, top-level result, final result).  Test infrastructure, not product."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import msegment_b200 as mseg  # noqa: E402
from oracle import oracle as orc  # noqa: E402


def unpack(plane):
    return np.stack([(plane >> (8 * c)) & 0xFF for c in range(3)], axis=-1).astype(np.uint8)


def report(name, got, want):
    bad = np.any(got != want, axis=-1)
    n = int(bad.sum())
    msg = "%-28s %s  (%d / %d differ)" % (name, "OK  " if n == 0 else "FAIL", n, bad.size)
    if n:
        ys, xs = np.nonzero(bad)
        msg += " first (y,x)=(%d,%d) got %s want %s; bbox y[%d,%d] x[%d,%d]" % (
            ys[0], xs[0], got[ys[0], xs[0]].tolist(), want[ys[0], xs[0]].tolist(), ys.min(), ys.max(), xs.min(), xs.max())
    print(msg)
    return n


def main():
    w, h, seed = (int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3])) if len(sys.argv) > 3 else (300, 200, 1)
    sp, sr = (float(sys.argv[4]), float(sys.argv[5])) if len(sys.argv) > 5 else (10.0, 10.0)
    im = orc.synth_bgr(w, h, seed)
    ctx = mseg.Context(0)
    gi = mseg.GpuImgproc(ctx)
    total = 0
    for ml in (0, 1, 2):
        got = gi.pyrMeanShiftFiltering(im, sp, sr, ml)
        S = [im]
        for l in range(ml):
            S.append(orc.pyr_down(S[-1]))
        for l in range(ml + 1):
            total += report("L%d: S[%d] plane" % (ml, l), unpack(ctx.debug_plane(0, l)), S[l])
        top = orc.meanshift_filter(S[ml], max(sp / (1 << ml), 1.0), sr, 0)
        total += report("L%d: D[%d] (top level)" % (ml, ml), unpack(ctx.debug_plane(1, ml)), top)
        total += report("L%d: final" % ml, got, orc.meanshift_filter(im, sp, sr, ml))
        print("   stats:", ctx.stats())
    f = orc.meanshift_filter(im, sp, sr, 1)
    n0, l0 = orc.label_regions(f, 2)
    n1, l1 = gi.labelRegions(f, 2, 2, 4)
    print("label_regions: n oracle %d gpu %d, differ %d" % (n0, n1, int((l0 != l1).sum())))
    total += int((l0 != l1).sum())
    m0, lm0 = orc.merge_regions(f, l0, 50, 10)
    m1, lm1 = gi.mergeRegions(f, l0, 50, 10)
    print("merge_regions: n oracle %d gpu %d, differ %d, rounds %d" % (m0, m1, int((lm0 != lm1).sum()), ctx.stats()["merge_rounds"]))
    total += int((lm0 != lm1).sum())
    print("TOTAL MISMATCHES", total)
    return 1 if total else 0


if __name__ == "__main__":
    sys.exit(main())
