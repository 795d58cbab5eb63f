"""CPU: the strip decomposition bench.py's reference arm uses to spread ONE whole frame over several host cores
(cv2.pyrMeanShiftFiltering is single-threaded).  Even strip starts + a halo >= the dependency bound: the oracle
port, which takes the strip's offset and keeps absolute coordinates, gives exactly the whole-frame rows; cv2 on a strip works
in strip-local coordinates and differs from its own whole-frame call on isolated pixels -- same work, so
fine for a timing baseline, and documented as such."""
import importlib
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import oracle as orc  # noqa: E402

bench = importlib.import_module("bench")


def test_strip_plan_covers_the_frame_with_even_starts():
    for h, parts in ((2160, 8), (1080, 8), (400, 4), (2160, 5), (97, 3)):
        plan = bench.strip_plan(h, parts, bench.REF_HALO)
        assert plan[0][0] == 0 and plan[-1][1] == h
        for k, (lo, hi, top, bot) in enumerate(plan):
            assert lo % 2 == 0 and top % 2 == 0 and top <= lo < hi <= bot <= h
            if k:
                assert lo == plan[k - 1][1]


def test_strips_equal_whole_frame_oracle_port():
    im = orc.synth_bgr(300, 420, 1001)
    full = orc.meanshift_filter(im, 10, 10, 1)
    parts = [bench._ref_strip_job(("port", np.ascontiguousarray(im[t:b]), lo, hi, t, 420))
             for (lo, hi, t, b) in bench.strip_plan(420, 4, bench.REF_HALO)]
    assert np.array_equal(np.concatenate(parts, axis=0), full)


def test_strips_close_to_whole_frame_cv2():
    cv2 = pytest.importorskip("cv2")
    for (w, h, seed, parts) in ((300, 420, 1001, 4), (640, 800, 1002, 8)):
        im = orc.synth_bgr(w, h, seed)
        full = cv2.pyrMeanShiftFiltering(im, 10, 10, maxLevel=1, termcrit=(3, 5, 1.0))
        got = [bench._ref_strip_job(("reference", np.ascontiguousarray(im[t:b]), lo, hi, t, h))
               for (lo, hi, t, b) in bench.strip_plan(h, parts, bench.REF_HALO)]
        got = np.concatenate(got, axis=0)
        differ = (got != full).any(axis=2).mean()
        assert differ < 1e-3, (w, h, differ)              # strip-local coordinates: isolated pixels only
        assert np.array_equal(full, orc.meanshift_filter(im, 10, 10, 1))     # the whole-frame call is what the oracle pins
