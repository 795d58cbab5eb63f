"""GPU (B200): BASELINE.json's full sizes.  1080p is compared with the oracle outright (about 10 s of CPU);
4K uses size-independent properties plus an oracle check of the label stage (cheap on CPU)."""
import numpy as np
import pytest

import msegment_b200 as mseg
from oracle import oracle as orc

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def gi():
    ctx = mseg.Context(0)
    yield mseg.GpuImgproc(ctx)
    ctx.close()


def test_1080p_config2_vs_oracle(gi):
    im = orc.synth_bgr(1920, 1080, 2)
    out = gi.segment(im, 10, 10, 1, loDiff=2, minSize=50, colorDist=10)
    f = orc.meanshift_filter(im, 10, 10, 1)
    assert np.array_equal(out["filtered"], f), int((out["filtered"] != f).any(axis=2).sum())
    n0, l0 = orc.label_regions(f, 2)
    n1, l1 = orc.merge_regions(f, l0, 50, 10)
    assert out["n_regions"] == n1 and np.array_equal(out["labels"], l1)


def test_4k_properties(gi):
    w, h = 3840, 2160
    im = orc.synth_bgr(w, h, 3)
    out = gi.segment(im, 10, 10, 1, loDiff=2, want=("filtered", "labels"))
    f, lab = out["filtered"], out["labels"]
    # (a) a crop far from the borders is independent of the rest of the image (window reach = maxCount*sp per level):
    #     oracle on a crop with margin must reproduce the interior of the full-frame GPU result
    y0, x0, ch, cw, m = 1000, 2000, 96, 128, 160      # margin 160 >= dependency cone of level 1 (even origin)
    crop = np.ascontiguousarray(im[y0 - m:y0 + ch + m, x0 - m:x0 + cw + m])
    fc = orc.meanshift_filter_roi(crop, x0 - m, y0 - m, w, h, 10, 10, 1)   # global coordinates (App. A.2)
    assert np.array_equal(fc[m:m + ch, m:m + cw], f[y0:y0 + ch, x0:x0 + cw])
    # (b) label stage against the oracle at full size (union-find on CPU is fast)
    n0, l0 = orc.label_regions(f, 2)
    assert out["n_regions"] == n0 and np.array_equal(lab, l0)
    # (c) determinism
    out2 = gi.segment(im, 10, 10, 1, loDiff=2, want=("filtered", "labels"))
    assert np.array_equal(out2["filtered"], f) and np.array_equal(out2["labels"], lab)
