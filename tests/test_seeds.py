"""Shape-method marker generator (PictureService.java:404-442): Canny, dilate, subtract and the whole
gray -> median -> Canny -> dilate/dilate/subtract -> median 3 -> connectedComponents(8) chain.
CPU: oracle vs cv2 golden vectors.  GPU: CUDA vs golden vectors and vs the oracle at larger sizes."""
import os

import numpy as np
import pytest

import msegment_b200 as mseg
from oracle import oracle as orc


def _golden(golden_dir):
    return np.load(os.path.join(golden_dir, "seeds.npz"))


def _names(g):
    return sorted(k[3:] for k in g.files if k.startswith("in/"))


def test_oracle_seed_stages_golden(golden_dir):
    g = _golden(golden_dir)
    names = _names(g)
    assert len(names) >= 6
    for n in names:
        gray = orc.bgr2gray(g["in/" + n])
        for k, (lo, hi) in enumerate(g["canny_params"]):
            assert np.array_equal(orc.canny(gray, lo, hi), g["canny%d/%s" % (k, n)]), (n, k)
        for kw, kh in ((3, 3), (5, 5), (7, 2)):
            assert np.array_equal(orc.dilate_rect(gray, kw, kh), g["dilate%dx%d/%s" % (kw, kh, n)]), (n, kw, kh)


def test_oracle_seed_chain_golden(golden_dir):
    g = _golden(golden_dir)
    for n in _names(g):
        cnt, markers, st = orc.shape_seeds(g["in/" + n])
        assert np.array_equal(st["edges"], g["chain_edges/" + n]), n
        assert np.array_equal(st["dde"], g["chain_dde/" + n]), n
        assert np.array_equal(st["dde3"], g["chain_dde3/" + n]), n
        assert cnt == int(g["chain_n/" + n]) and np.array_equal(markers, g["chain_markers/" + n]), n


def test_blur_mask_size_rule():
    # PictureService.java:877-899, the values SURVEY quotes: 7 at 512^2, 11 at 1080p (1080 * 0.01 = 10 -> 11)
    assert [orc.blur_mask_size(w, h) for w, h in ((2, 50), (96, 80), (225, 225), (400, 373), (512, 512), (1920, 1080), (3840, 2160))] \
        == [1, 5, 5, 7, 7, 11, 11]
    for w, h in ((2, 50), (96, 80), (225, 225), (400, 373), (512, 512), (1920, 1080), (3840, 2160), (1500, 1500)):
        assert mseg.GpuImgproc.calculateSizeOfSquareBlurMask(w, h) == orc.blur_mask_size(w, h)


@pytest.mark.gpu
def test_gpu_seed_stages_golden(golden_dir):
    g = _golden(golden_dir)
    with mseg.Context(0) as ctx:
        gi = mseg.GpuImgproc(ctx)
        for n in _names(g):
            im = g["in/" + n]
            gray = gi.cvtColorBGR2GRAY(im)
            for k, (lo, hi) in enumerate(g["canny_params"]):
                assert np.array_equal(gi.Canny(gray, lo, hi), g["canny%d/%s" % (k, n)]), (n, k)
            for kw, kh in ((3, 3), (5, 5), (7, 2)):
                assert np.array_equal(gi.dilate(gray, (kh, kw)), g["dilate%dx%d/%s" % (kw, kh, n)]), (n, kw, kh)
            cnt, markers, st = gi.shapeSeeds(im, stages=True)
            assert np.array_equal(st["edges"], g["chain_edges/" + n]), n
            assert np.array_equal(st["dde"], g["chain_dde/" + n]), n
            assert np.array_equal(st["dde3"], g["chain_dde3/" + n]), n
            assert cnt == int(g["chain_n/" + n]) and np.array_equal(markers, g["chain_markers/" + n]), n


@pytest.mark.gpu
@pytest.mark.parametrize("size", [(640, 360), (1920, 1080), (333, 517)])
def test_gpu_seed_chain_vs_oracle(size):
    w, h = size
    im = orc.synth_bgr(w, h, 21)
    want_n, want_m, want = orc.shape_seeds(im)
    rng = np.random.default_rng(3)
    with mseg.Context(0) as ctx:
        gi = mseg.GpuImgproc(ctx)
        n, m, st = gi.shapeSeeds(im, stages=True)
        assert st["k"] == want["k"]
        for key in ("blurred", "edges", "dde", "dde3"):
            assert np.array_equal(st[key], want[key]), key
        assert n == want_n and np.array_equal(m, want_m)
        n2, m2 = gi.shapeSeeds(im)
        assert n2 == n and np.array_equal(m2, m)
        # the single calls, on noise and with swapped / fractional / extreme thresholds
        noise = rng.integers(0, 256, (h, w), dtype=np.uint8)
        for lo, hi in ((5, 50), (50, 5), (0, 0), (100.9, 300.2), (1, 5000), (-3, 10)):
            assert np.array_equal(gi.Canny(noise, lo, hi), orc.canny(noise, lo, hi)), (lo, hi)
            assert np.array_equal(gi.Canny(want["blurred"], lo, hi), orc.canny(want["blurred"], lo, hi)), (lo, hi)
        for kh, kw in ((3, 3), (5, 5), (1, 9), (8, 2), (31, 31)):
            assert np.array_equal(gi.dilate(noise, (kh, kw)), orc.dilate_rect(noise, kw, kh)), (kh, kw)
        other = rng.integers(0, 256, (h, w), dtype=np.uint8)
        assert np.array_equal(gi.subtract(noise, other), orc.subtract_u8(noise, other))
        # strided views
        view = np.ascontiguousarray(np.pad(noise, ((0, 0), (3, 5))))[:, 3:3 + w]
        assert np.array_equal(gi.Canny(view, 5, 50), orc.canny(noise, 5, 50))


@pytest.mark.gpu
def test_gpu_seed_argument_checks():
    with mseg.Context(0) as ctx:
        gi = mseg.GpuImgproc(ctx)
        gray = np.zeros((8, 8), np.uint8)
        with pytest.raises(mseg.CvException):
            gi.Canny(np.zeros((8, 8, 3), np.uint8), 5, 50)
        with pytest.raises(mseg.CvException):
            gi.dilate(gray, (0, 3))
        with pytest.raises(mseg.CvException):
            gi.subtract(gray, np.zeros((8, 9), np.uint8))
        with pytest.raises(mseg.CvException):
            gi.shapeSeeds(np.zeros((8, 8, 3), np.uint8), medianKsize=4)
        n, m = gi.shapeSeeds(np.full((8, 8, 3), 7, np.uint8))      # flat image: no edges, no seeds
        assert n == 1 and not m.any()
