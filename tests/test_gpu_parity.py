"""GPU (B200): the CUDA path, called through the C ABI, against the committed OpenCV golden vectors and the
CPU oracle on the same seeded inputs.  Bar: bit-exact (integer / byte / index work throughout)."""
import os

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


def _diff(a, b):
    bad = np.argwhere(np.any(a != b, axis=-1) if a.ndim == 3 else a != b)
    return "%d/%d differ, first at %s" % (len(bad), a.shape[0] * a.shape[1], bad[:3].tolist())


# ------------------------------------------------------------------ golden vectors (cv2 4.13.0)

def test_meanshift_golden(gi, golden_dir):
    g = np.load(os.path.join(golden_dir, "meanshift.npz"))
    params = g["params"]
    names = sorted(k[3:] for k in g.files if k.startswith("in/"))
    for name in names:
        im = g["in/" + name]
        for k, (sp, sr, ml, tt, tc, te) in enumerate(params):
            want = g["out/%s/%d" % (name, k)]
            got = gi.pyrMeanShiftFiltering(im, sp, sr, int(ml), (int(tt), int(tc), float(te)))
            assert np.array_equal(got, want), (name, k, _diff(got, want))


def test_label_regions_golden(gi, golden_dir):
    g = np.load(os.path.join(golden_dir, "labels.npz"))
    for name in sorted(k[3:] for k in g.files if k.startswith("in/")):
        f = g["in/" + name]
        for d in (0, 2, 5):
            want = g["ff%d/%s" % (d, name)]
            n, lab = gi.labelRegions(f, d, d, 4)
            assert n == want.max() and np.array_equal(lab, want), (name, d, n, want.max(), _diff(lab, want))
        want8 = g["ff2c8/" + name]
        n, lab = gi.labelRegions(f, 2, 2, 8)
        assert n == want8.max() and np.array_equal(lab, want8), (name, "8-conn", _diff(lab, want8))


def test_connected_components_golden(gi, golden_dir):
    g = np.load(os.path.join(golden_dir, "labels.npz"))
    for k in sorted(int(k[5:]) for k in g.files if k.startswith("mask/")):
        m = g["mask/%d" % k]
        for conn in (4, 8):
            n, lab = gi.connectedComponents(m, conn)
            assert n == int(g["ccn%d/%d" % (conn, k)]), (k, conn)
            assert np.array_equal(lab, g["cc%d/%d" % (conn, k)]), (k, conn, _diff(lab, g["cc%d/%d" % (conn, k)]))


# ------------------------------------------------------------------ oracle on seeded inputs

MS_CASES = [
    # w, h, seed, sp, sr, maxLevel, termcrit
    (131, 97, 1, 10, 10, 1, (3, 5, 1.0)),
    (200, 333, 2, 10, 10, 0, (3, 5, 1.0)),
    (333, 200, 3, 5, 20, 2, (3, 5, 1.0)),
    (257, 129, 4, 7.5, 12, 1, (3, 5, 1.0)),        # non-integral sp: window parity rule
    (160, 120, 5, 10, 3, 2, (3, 5, 1.0)),          # sr < 4: isr2 != isr22
    (96, 64, 6, 3, 300, 1, (3, 4, 1.0)),           # isr2 >= 254^2: generic (no-sentinel) kernel
    (150, 150, 7, 20, 40, 1, (3, 5, 1.0)),
    (64, 48, 8, 2, 8, 3, (1, 100, 0.0)),           # maxCount = 100, eps ignored -> 1
    (100, 80, 9, 6, 15, 1, (2, 0, 50.0)),          # eps only, large
    (41, 47, 10, 10, 10, 1, (3, 5, 1.0)),          # smaller than one tile
    (1, 50, 11, 4, 30, 2, (3, 5, 1.0)),
    (50, 1, 12, 4, 30, 2, (3, 5, 1.0)),
    (512, 512, 1, 10, 10, 1, (3, 5, 1.0)),         # BASELINE config 1 input
]


@pytest.mark.parametrize("case", MS_CASES, ids=lambda c: "%dx%d_sp%g_sr%g_L%d" % (c[0], c[1], c[3], c[4], c[5]))
def test_meanshift_vs_oracle(gi, case):
    w, h, seed, sp, sr, ml, term = case
    im = orc.synth_bgr(w, h, seed)
    want = orc.meanshift_filter(im, sp, sr, ml, term)
    got = gi.pyrMeanShiftFiltering(im, sp, sr, ml, term)
    assert np.array_equal(got, want), _diff(got, want)


def test_meanshift_noise_and_flat(gi):
    rng = np.random.default_rng(5)
    for im in (rng.integers(0, 256, (90, 123, 3), dtype=np.uint8), np.full((70, 70, 3), 200, np.uint8),
               np.zeros((33, 65, 3), np.uint8), np.full((40, 40, 3), 255, np.uint8)):
        for sp, sr, ml in ((10, 10, 1), (5, 40, 2), (8, 25, 0)):
            want = orc.meanshift_filter(im, sp, sr, ml)
            got = gi.pyrMeanShiftFiltering(im, sp, sr, ml)
            assert np.array_equal(got, want), (im.shape, sp, sr, ml, _diff(got, want))


def test_meanshift_strided_input_and_errors(gi):
    big = orc.synth_bgr(200, 100, 3)
    view = big[10:90, 20:150]                      # non-continuous Mat (step > 3*w)
    want = orc.meanshift_filter(np.ascontiguousarray(view), 6, 14, 1)
    assert np.array_equal(gi.pyrMeanShiftFiltering(view, 6, 14, 1), want)
    with pytest.raises(mseg.CvException):          # OpenCV: "The number of pyramid levels is too large or negative"
        gi.pyrMeanShiftFiltering(big, 5, 5, 9)
    with pytest.raises(mseg.CvException):
        gi.pyrMeanShiftFiltering(big, 5, 5, -1)
    with pytest.raises(mseg.CvException):
        gi.labelRegions(big, 2, 3, 4)              # asymmetric range rejected
    with pytest.raises(mseg.CvException):
        gi.labelRegions(big, 2, 2, 6)              # connectivity must be 4 or 8
    with pytest.raises(mseg.CvException):
        gi.connectedComponents(np.zeros((4, 4), np.uint8), 6)


@pytest.mark.parametrize("w,h,seed", [(131, 97, 1), (640, 360, 2), (31, 200, 3), (1000, 40, 4)])
def test_label_and_merge_vs_oracle(gi, w, h, seed):
    im = orc.synth_bgr(w, h, seed)
    f = orc.meanshift_filter(im, 6, 12, 1)
    for d in (0, 2, 6):
        n0, l0 = orc.label_regions(f, d)
        n1, l1 = gi.labelRegions(f, d, d, 4)
        assert n0 == n1 and np.array_equal(l0, l1), (d, n0, n1, _diff(l1, l0))
    n8, l8 = orc.label_regions(f, 3, 8)
    m8, g8 = gi.labelRegions(f, 3, 3, 8)
    assert n8 == m8 and np.array_equal(l8, g8), ("8-conn", n8, m8)
    out8 = gi.segment(im, 6, 12, 1, loDiff=3, connectivity=8, want=("labels",))
    assert out8["n_regions"] == n8 and np.array_equal(out8["labels"], l8)
    n0, l0 = orc.label_regions(f, 2)
    for min_size, cd in ((20, 0), (0, 10), (50, 10), (10**9, 0)):
        m0, lm0 = orc.merge_regions(f, l0, min_size, cd)
        m1, lm1 = gi.mergeRegions(f, l0, min_size, cd)
        assert m0 == m1 and np.array_equal(lm0, lm1), (min_size, cd, m0, m1, _diff(lm1, lm0))


def test_label_uniform_and_checker(gi):
    flat = np.full((100, 300, 3), 9, np.uint8)
    n, lab = gi.labelRegions(flat, 0, 0, 4)
    assert n == 1 and (lab == 1).all()
    yy, xx = np.mgrid[0:64, 0:96]
    chk = (((yy + xx) & 1) * 255).astype(np.uint8)
    img = np.stack([chk, chk, chk], axis=-1)
    n, lab = gi.labelRegions(img, 2, 2, 4)
    n0, l0 = orc.label_regions(img, 2)
    assert n == n0 == 64 * 96 and np.array_equal(lab, l0)
    # serpentine: one long winding component
    s = np.zeros((65, 129), np.uint8)
    s[::2, :] = 255
    s[1::4, -1] = 255
    s[3::4, 0] = 255
    for conn in (4, 8):
        n, lab = gi.connectedComponents(s, conn)
        n0, l0 = orc.connected_components(s, conn)
        assert n == n0 and np.array_equal(lab, l0)


@pytest.mark.parametrize("w,h,p,seed", [(640, 480, .5, 1), (1001, 333, .6, 2), (257, 300, .35, 3), (64, 64, .02, 4)])
def test_connected_components_vs_oracle(gi, w, h, p, seed):
    rng = np.random.default_rng(seed)
    m = (rng.random((h, w)) < p).astype(np.uint8) * 255
    for conn in (4, 8):
        n0, l0 = orc.connected_components(m, conn)
        n1, l1 = gi.connectedComponents(m, conn)
        assert n0 == n1 and np.array_equal(l0, l1), (conn, n0, n1, _diff(l1, l0))


def test_render_vs_oracle(gi):
    rng = np.random.default_rng(2)
    lab = rng.integers(-3, 40, (77, 131)).astype(np.int32)
    assert np.array_equal(gi.colorByIndexes(lab, 25), orc.render_labels(lab, 25))
    cols = rng.integers(0, 256, (25, 3)).astype(np.uint8)
    assert np.array_equal(gi.colorByIndexes(lab, 25, cols), orc.render_labels(lab, 25, cols))


def test_segment_fused_vs_oracle(gi):
    im = orc.synth_bgr(512, 512, 1)                # BASELINE config 1 (ii)
    out = gi.segment(im, 10, 10, 1, loDiff=2, minSize=50, colorDist=10)
    f = orc.meanshift_filter(im, 10, 10, 1)
    assert np.array_equal(out["filtered"], f), _diff(out["filtered"], f)
    n0, l0 = orc.label_regions(f, 2)
    n1, l1 = orc.merge_regions(f, l0, 50, 10)
    assert out["n_regions"] == n1 and np.array_equal(out["labels"], l1), _diff(out["labels"], l1)
    assert np.array_equal(out["rendered"], orc.render_labels(l1, n1))
    st = gi.ctx.stats()
    assert st["kernel_launches"] > 0 and st["ms_active_items"] > 0
