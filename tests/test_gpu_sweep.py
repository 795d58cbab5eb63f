"""GPU (B200): BASELINE.json config 3 -- one 3840x2160 frame, sp x sr sweep, full output batch per variant.
Filtered images are checked against the oracle on a crop evaluated in GLOBAL coordinates (orc_meanshift_filter_roi,
margin >= the dependency cone), label / merge / render stages against the oracle at full size."""
import numpy as np
import pytest

import msegment_b200 as mseg
from oracle import oracle as orc

pytestmark = pytest.mark.gpu
W, H, SEED = 3840, 2160, 3
SWEEP = [(sp, sr) for sp in (5, 10, 20) for sr in (10, 20, 40)]     # mirrors CorrelationTestService.java:29-39 in shape


@pytest.fixture(scope="module")
def frame():
    return orc.synth_bgr(W, H, SEED)


@pytest.fixture(scope="module")
def gi():
    ctx = mseg.Context(0)
    yield mseg.GpuImgproc(ctx)
    ctx.close()


@pytest.mark.parametrize("sp,sr", SWEEP, ids=lambda v: str(v))
def test_sweep_variant(gi, frame, sp, sr):
    out = gi.segment(frame, sp, sr, 1, loDiff=2, minSize=50, colorDist=10)
    f, lab, ren = out["filtered"], out["labels"], out["rendered"]
    # (1) filter: oracle on a crop in global coordinates; margin covers maxCount*sp per level plus the mask/pyramid taps
    m = 16 * int(sp) + 32
    ch, cw = 96, 128
    y0, x0 = 1024, 1900
    crop = np.ascontiguousarray(frame[y0 - m:y0 + ch + m, x0 - m:x0 + cw + m])
    fc = orc.meanshift_filter_roi(crop, x0 - m, y0 - m, W, H, sp, sr, 1)
    assert np.array_equal(fc[m:m + ch, m:m + cw], f[y0:y0 + ch, x0:x0 + cw]), (sp, sr)
    # a crop touching the image corner exercises the clamped windows / reflect borders
    cs = 64 + m
    fc2 = orc.meanshift_filter_roi(np.ascontiguousarray(frame[:cs, :cs]), 0, 0, W, H, sp, sr, 1)
    assert np.array_equal(fc2[:64, :64], f[:64, :64]), (sp, sr)
    # (2) label + merge + render at full size
    n0, l0 = orc.label_regions(f, 2)
    n1, l1 = orc.merge_regions(f, l0, 50, 10)
    assert out["n_regions"] == n1 and np.array_equal(lab, l1), (sp, sr, out["n_regions"], n1)
    assert np.array_equal(ren, orc.render_labels(l1, n1))
    # bw_result (PictureService.java:376-379) of a white render is any of its channels
    assert set(np.unique(ren)) <= {0, 255}
