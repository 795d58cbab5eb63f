"""Colour-method marker generator (PictureService.java:309-366, :938-943, :1018-1023) and the
bilateral pre-filter (row a5, :490): Otsu threshold, chamfer distance transform, min-max normalisation, peak threshold +
dilate, findContours / drawContours labelling, filled circle.
CPU: oracle vs cv2 golden vectors (tests/golden/gen_color_seeds.py).  GPU: CUDA vs golden vectors and vs the oracle at
larger sizes.  Integer / label stages are bit-exact; the float planes are compared bit for bit too (same operation order);
the bilateral filter is bit-exact against the oracle and within 1 LSB of cv2 (cv2's vector path fuses multiply-adds)."""
import os

import numpy as np
import pytest

import msegment_b200 as mseg
from oracle import oracle as orc


def _golden(golden_dir):
    return np.load(os.path.join(golden_dir, "color_seeds.npz"))


def _names(g):
    return sorted(k[3:] for k in g.files if k.startswith("in/"))


def _masks(g):
    return sorted(int(k[5:]) for k in g.files if k.startswith("mask/"))


def test_oracle_color_chain_golden(golden_dir):
    g = _golden(golden_dir)
    names = _names(g)
    assert len(names) >= 7
    for n in names:
        cnt, markers, st = orc.color_seeds(g["in/" + n])
        assert np.array_equal(st["black_bg"], g["black/" + n]), n
        assert np.array_equal(st["sharp"], g["sharp/" + n]), n
        assert np.array_equal(st["gray"], g["gray/" + n]), n
        assert st["otsu"] == int(g["otsu/" + n]), n
        assert np.array_equal(st["bw"], g["bw/" + n]), n
        assert np.array_equal(st["dist"], g["dist/" + n]), n
        assert np.array_equal(st["norm"], g["norm/" + n]), n
        assert np.array_equal(st["peaks"], g["peaks/" + n]), n
        assert cnt == int(g["n/" + n]), n
        assert np.array_equal(markers, g["markers/" + n]), n
        assert np.array_equal(orc.contour_markers(g["peaks/" + n])[1], g["contours/" + n]), n


def test_oracle_contours_dist_circle_golden(golden_dir):
    g = _golden(golden_dir)
    ks = _masks(g)
    assert len(ks) >= 10
    for k in ks:
        m = g["mask/%d" % k]
        n, lab = orc.contour_markers(m)
        assert n == int(g["mask_n/%d" % k]) and np.array_equal(lab, g["mask_markers/%d" % k]), k
        assert np.array_equal(orc.distance_transform(m * 255), g["mask_dist/%d" % k]), k
        cx, cy, r, v = (int(t) for t in g["mask_circle_args/%d" % k])
        assert np.array_equal(orc.circle_filled(lab, cx, cy, r, v), g["mask_circle/%d" % k]), k


def test_oracle_distance_transform_tie_cases_golden(golden_dir):
    """cv2 vectors where the forward pass' unrounded running value inside aligned groups of four columns decides the result
    (a horizontal step crosses 32 / 64 on an exact rounding tie); plain float rounding everywhere fails these by 1 ulp."""
    g = _golden(golden_dir)
    ks = sorted(int(k[7:]) for k in g.files if k.startswith("dt_tie/"))
    assert len(ks) >= 6
    plain_differs = 0
    for k in ks:
        m = g["dt_tie_mask/%d" % k] * 255
        want = g["dt_tie/%d" % k]
        assert np.array_equal(orc.distance_transform(m), want), k
        # the plain two-pass float recurrence (numpy, row by row) is NOT what cv2 computes on these
        plain_differs += int(not np.array_equal(_plain_float_chamfer(m), want))
    assert plain_differs >= 3


def _dt_fixed_golden(golden_dir):
    return np.load(os.path.join(golden_dir, "dt_fixed.npz"))


def test_oracle_distance_transform_fixed_golden(golden_dir):
    """cv2.distanceTransform with IPP switched off (OpenCV's own 16.16 fixed-point chamfer: what a non-IPP build such as the
    openpnp 3.4.2 natives computes, tests/golden/gen_dt_fixed.py) == orc_distance_transform_l2_5_fixed, incl. the images
    without any zero pixel (65533.805 everywhere) and the cases where the two arithmetics differ."""
    g = _dt_fixed_golden(golden_dir)
    names = [str(n) for n in g["names"]]
    assert len(names) >= 12
    differs_from_float_mode = 0
    for n in names:
        m, want = g["mask/" + n], g["dist/" + n]
        assert np.array_equal(orc.distance_transform(m, fixed=True), want), n
        differs_from_float_mode += int(not np.array_equal(orc.distance_transform(m), want))
    assert differs_from_float_mode >= 5


@pytest.mark.gpu
def test_gpu_distance_transform_fixed_mode(golden_dir):
    """Option dt_fixed = 1: the whole-GPU closed-form kernels (k_dt_fixed.cu) == cv2 without IPP on the golden vectors and ==
    the oracle's two-pass fixed-point recurrence on larger blobs, noise, single zero pixels far away, ragged sizes; the option
    switches back; the whole colour-seed chain follows the mode."""
    g = _dt_fixed_golden(golden_dir)
    with mseg.Context(0) as ctx:
        gi = mseg.GpuImgproc(ctx)
        ctx.set_option("dt_fixed", 1)
        assert ctx.get_option("dt_fixed") == 1
        for n in (str(n) for n in g["names"]):
            assert np.array_equal(gi.distanceTransform(g["mask/" + n]), g["dist/" + n]), n
        for (w, h, sigma, level) in ((640, 360, 6, .5), (1920, 1080, 12, .45), (2500, 300, 9, .3), (333, 517, 3, .6), (1031, 77, 2, .5)):
            mask = _blobs(w, h, 5, sigma, level) * 255
            assert np.array_equal(gi.distanceTransform(mask), orc.distance_transform(mask, fixed=True)), (w, h)
            noise = ((np.random.default_rng(w).random((h, w)) < .97) * 255).astype(np.uint8)
            assert np.array_equal(gi.distanceTransform(noise), orc.distance_transform(noise, fixed=True)), (w, h)
        for (w, h, zy, zx) in ((3000, 40, 0, 0), (3000, 40, 39, 2999), (1500, 700, 350, 750), (70, 2000, 1999, 0), (5000, 9, 4, 17),
                               (33, 65, 64, 32), (129, 31, 0, 128)):
            m = np.full((h, w), 255, np.uint8)
            m[zy, zx] = 0
            assert np.array_equal(gi.distanceTransform(m), orc.distance_transform(m, fixed=True)), (w, h, zy, zx)
        for shape in ((1, 1), (1, 40), (40, 1), (2, 2), (5, 3), (33, 1), (1, 257)):
            for fill in (0, 255):
                m = np.full(shape, fill, np.uint8)
                assert np.array_equal(gi.distanceTransform(m), orc.distance_transform(m, fixed=True)), (shape, fill)
        # strided ROI views in and out
        big = ((np.random.default_rng(3).random((90, 140)) < .95) * 255).astype(np.uint8)
        view = big[7:71, 13:120]
        assert np.array_equal(gi.distanceTransform(view), orc.distance_transform(np.ascontiguousarray(view), fixed=True))
        # the chain follows the mode (and differs from the float mode's distances on this frame)
        im = orc.synth_bgr(640, 360, 31)
        want_n, want_m, want = orc.color_seeds(im, dt_fixed=True)
        n, m, st = gi.colorSeeds(im, stages=True)
        assert n == want_n and np.array_equal(m, want_m)
        for key in ("bw", "norm", "peaks"):
            assert np.array_equal(st[key], want[key]), key
        with pytest.raises(mseg.CvException):
            gi.distanceTransform(np.zeros((2, 16385), np.uint8))
        ctx.set_option("dt_fixed", 0)
        m = _blobs(300, 200, 5, 4, .5) * 255
        assert np.array_equal(gi.distanceTransform(m), orc.distance_transform(m))


def _plain_float_chamfer(mask):
    """Two-pass 5x5 chamfer with every addition rounded to float32 (the textbook order)."""
    f = np.float32
    h, w = mask.shape
    a, b, c, inf = f(1.0), f(1.4), f(2.1969), np.finfo(np.float32).max
    t = np.full((h + 4, w + 4), inf, np.float32)
    fw = ((-2, -1, c), (-2, 1, c), (-1, -2, c), (-1, -1, b), (-1, 0, a), (-1, 1, b), (-1, 2, c), (0, -1, a))
    with np.errstate(over="ignore"):
        for y in range(h):
            for x in range(w):
                if mask[y, x] == 0:
                    t[y + 2, x + 2] = 0
                    continue
                t[y + 2, x + 2] = min(min(f(t[y + 2 + dy, x + 2 + dx] + m), inf) for dy, dx, m in fw)
        for y in range(h - 1, -1, -1):
            for x in range(w - 1, -1, -1):
                v = t[y + 2, x + 2]
                t[y + 2, x + 2] = min(v, min(min(f(t[y + 2 - dy, x + 2 - dx] + m), inf) for dy, dx, m in fw))
    return t[2:-2, 2:-2].copy()


def test_oracle_bilateral_golden(golden_dir):
    g = _golden(golden_dir)
    keys = [k for k in g.files if k.startswith("bil_")]
    assert len(keys) >= 12
    for key in keys:
        kind, name = key.split("/")
        d = int(kind[-1])
        im = g["in/" + name]
        src = orc.bgr2gray(im) if "gray" in kind else im
        got = orc.bilateral_filter(src, d, 2 * d, 2 * d)
        assert np.abs(got.astype(int) - g[key].astype(int)).max() <= 1, key        # tolerance: 1 LSB (float, fused vs not)


@pytest.mark.gpu
def test_gpu_color_chain_golden(golden_dir):
    g = _golden(golden_dir)
    with mseg.Context(0) as ctx:
        gi = mseg.GpuImgproc(ctx)
        for n in _names(g):
            im = g["in/" + n]
            assert np.array_equal(gi.whiteToBlack(im), g["black/" + n]), n
            t, bw = gi.threshold(g["gray/" + n], 40, 255, gi.THRESH_BINARY | gi.THRESH_OTSU)
            assert t == float(g["otsu/" + n]) and np.array_equal(bw, g["bw/" + n]), n
            dist = gi.distanceTransform(g["bw/" + n], gi.CV_DIST_L2, 5)
            assert np.array_equal(dist, g["dist/" + n]), n
            nrm = gi.normalize(dist, 0, 1., gi.NORM_MINMAX)
            assert np.array_equal(nrm, g["norm/" + n]), n
            _, th = gi.threshold(nrm, .4, 1., gi.THRESH_BINARY)
            pk = gi.convertToU8(gi.dilateF32(th, (3, 3)))
            assert np.array_equal(pk, g["peaks/" + n]), n
            cnt, m = gi.contourMarkers(pk)
            assert cnt == int(g["n/" + n]) and np.array_equal(m, g["contours/" + n]), n
            assert np.array_equal(gi.circle(m, (5, 5), 3, 255), g["markers/" + n]), n
            # the whole chain in one call
            cnt2, m2, st = gi.colorSeeds(im, stages=True)
            assert np.array_equal(st["sharp"], g["sharp/" + n]), n
            assert np.array_equal(st["bw"], g["bw/" + n]), n
            assert np.array_equal(st["norm"], g["norm/" + n]), n
            assert np.array_equal(st["peaks"], g["peaks/" + n]), n
            assert cnt2 == int(g["n/" + n]) and np.array_equal(m2, g["markers/" + n]), n
            cnt3, m3 = gi.colorSeeds(im)
            assert cnt3 == cnt2 and np.array_equal(m3, m2), n


@pytest.mark.gpu
def test_gpu_contours_dist_circle_golden(golden_dir):
    g = _golden(golden_dir)
    with mseg.Context(0) as ctx:
        gi = mseg.GpuImgproc(ctx)
        for k in _masks(g):
            m = g["mask/%d" % k]
            n, lab = gi.contourMarkers(m)
            assert n == int(g["mask_n/%d" % k]) and np.array_equal(lab, g["mask_markers/%d" % k]), k
            assert np.array_equal(gi.distanceTransform(m * 255), g["mask_dist/%d" % k]), k
            cx, cy, r, v = (int(t) for t in g["mask_circle_args/%d" % k])
            assert np.array_equal(gi.circle(lab, (cx, cy), r, v), g["mask_circle/%d" % k]), k


@pytest.mark.gpu
def test_gpu_bilateral_golden_and_oracle(golden_dir):
    g = _golden(golden_dir)
    rng = np.random.default_rng(9)
    with mseg.Context(0) as ctx:
        gi = mseg.GpuImgproc(ctx)
        for key in [k for k in g.files if k.startswith("bil_")]:
            kind, name = key.split("/")
            d = int(kind[-1])
            im = g["in/" + name]
            src = orc.bgr2gray(im) if "gray" in kind else im
            got = gi.bilateralFilter(src, d, 2 * d, 2 * d)
            assert np.abs(got.astype(int) - g[key].astype(int)).max() <= 1, key    # vs cv2: 1 LSB
            assert np.array_equal(got, orc.bilateral_filter(src, d, 2 * d, 2 * d)), key   # vs the oracle: bit-exact
        for (w, h, d, sc, ss) in ((640, 360, 11, 22, 22), (333, 217, 9, 50, 3), (200, 100, 0, 10, 2.5), (64, 64, 3, -1, -1)):
            gray = rng.integers(0, 256, (h, w), dtype=np.uint8)
            assert np.array_equal(gi.bilateralFilter(gray, d, sc, ss), orc.bilateral_filter(gray, d, sc, ss)), (w, h, d)
            bgr = orc.synth_bgr(w, h, 3)
            assert np.array_equal(gi.bilateralFilter(bgr, d, sc, ss), orc.bilateral_filter(bgr, d, sc, ss)), (w, h, d)
        with pytest.raises(mseg.CvException):
            gi.bilateralFilter(np.zeros((8, 8, 2), np.uint8), 5, 10, 10)
        with pytest.raises(mseg.CvException):
            gi.bilateralFilter(np.zeros((8, 8), np.uint8), 99, 10, 10)


def _blobs(w, h, seed, sigma, level):
    rng = np.random.default_rng(seed)
    f = rng.random((h, w)).astype(np.float32)
    # separable box smoothing by cumulative sums (no cv2 on the GPU box path)
    k = int(sigma)
    for axis in (0, 1):
        c = np.cumsum(np.pad(f, [(k, k) if a == axis else (0, 0) for a in (0, 1)], mode="wrap"), axis=axis)
        f = (np.take(c, np.arange(2 * k, 2 * k + f.shape[axis]), axis=axis) - np.take(c, np.arange(0, f.shape[axis]), axis=axis)) / (2 * k)
    return (f > np.quantile(f, level)).astype(np.uint8)


@pytest.mark.gpu
@pytest.mark.parametrize("case", [(640, 360, 6, .5), (1920, 1080, 12, .45), (2500, 300, 9, .3), (333, 517, 3, .6), (1031, 77, 2, .5)])
def test_gpu_dist_and_contours_vs_oracle(case):
    w, h, sigma, level = case
    mask = _blobs(w, h, 5, sigma, level)
    with mseg.Context(0) as ctx:
        gi = mseg.GpuImgproc(ctx)
        d = gi.distanceTransform(mask * 255)
        assert np.array_equal(d, orc.distance_transform(mask * 255))
        n, lab = gi.contourMarkers(mask)
        wn, wlab = orc.contour_markers(mask)
        assert n == wn and np.array_equal(lab, wlab)
        noise = (np.random.default_rng(1).random((h, w)) < .62).astype(np.uint8)     # percolating noise: deep nesting, many holes
        n, lab = gi.contourMarkers(noise)
        wn, wlab = orc.contour_markers(noise)
        assert n == wn and np.array_equal(lab, wlab)
        assert np.array_equal(gi.distanceTransform(noise), orc.distance_transform(noise))


@pytest.mark.gpu
def test_gpu_dist_long_runs_and_edges():
    """Distances far beyond 32 px: the row scan crosses float binades (dt_advance), one zero pixel in a corner / centre."""
    with mseg.Context(0) as ctx:
        gi = mseg.GpuImgproc(ctx)
        for (w, h, zy, zx) in ((3000, 40, 0, 0), (3000, 40, 39, 2999), (1500, 700, 350, 750), (70, 2000, 1999, 0), (5000, 9, 4, 17)):
            m = np.full((h, w), 255, np.uint8)
            m[zy, zx] = 0
            assert np.array_equal(gi.distanceTransform(m), orc.distance_transform(m)), (w, h, zy, zx)
        for shape in ((1, 1), (1, 40), (40, 1), (2, 2), (5, 3)):
            for fill in (0, 255):
                m = np.full(shape, fill, np.uint8)
                assert np.array_equal(gi.distanceTransform(m), orc.distance_transform(m)), (shape, fill)
                assert gi.contourMarkers(m)[0] == orc.contour_markers(m)[0]
                assert np.array_equal(gi.contourMarkers(m)[1], orc.contour_markers(m)[1])
        full = np.full((20, 30), 255, np.uint8)                # no zero pixel: FLT_MAX everywhere, normalises to all zero
        d = gi.distanceTransform(full)
        assert (d == np.finfo(np.float32).max).all()
        assert not gi.normalize(d).any()
        with pytest.raises(mseg.CvException):
            gi.distanceTransform(full, 1, 3)
        with pytest.raises(mseg.CvException):
            gi.threshold(full, 40, 255, 3)


@pytest.mark.gpu
@pytest.mark.parametrize("size", [(512, 512), (1920, 1080), (333, 217)])
def test_gpu_color_seeds_vs_oracle(size):
    w, h = size
    im = orc.synth_bgr(w, h, 31)
    im[20:60, 30:90] = 255                                      # a pure white patch for the white->black loop
    want_n, want_m, want = orc.color_seeds(im)
    with mseg.Context(0) as ctx:
        gi = mseg.GpuImgproc(ctx)
        n, m, st = gi.colorSeeds(im, stages=True)
        for key in ("sharp", "bw", "norm", "peaks"):
            assert np.array_equal(st[key], want[key]), key
        assert n == want_n and np.array_equal(m, want_m)
        # the intended 3 x 3 reading of the sharpen kernel is one argument away
        k33 = np.array([1, 1, 1, 1, -8, 1, 1, 1, 1], np.int8).reshape(3, 3)
        n3, m3 = gi.colorSeeds(im, kernel=k33, peakThresh=0.3)
        w3n, w3m, _ = orc.color_seeds(im, taps=k33, peak_thresh=0.3)
        assert n3 == w3n and np.array_equal(m3, w3m)
        # Otsu on noise and on a two-level image
        rng = np.random.default_rng(2)
        for g in (rng.integers(0, 256, (h, w), dtype=np.uint8), (rng.integers(0, 2, (h, w)) * 200).astype(np.uint8),
                  np.full((h, w), 9, np.uint8)):
            t, bw = gi.threshold(g, 0, 255, gi.THRESH_BINARY | gi.THRESH_OTSU)
            assert t == orc.otsu_threshold(g) and np.array_equal(bw, orc.threshold_binary(g, int(t)))
        t, bw = gi.threshold(want["gray"], 99.7, 200, gi.THRESH_BINARY)
        assert t == 99 and np.array_equal(bw, orc.threshold_binary(want["gray"], 99, 200))
