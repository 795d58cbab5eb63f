"""The pre-filters the reference calls around its segmentation stage: oracle vs cv2 golden vectors on CPU,
CUDA vs golden vectors and oracle on the GPU."""
import os

import numpy as np
import pytest

import msegment_b200 as mseg
from oracle import oracle as orc

TAPS = np.array([1, 1, 1, 1, -8, 1, 1, 1, 1], np.int8)     # PictureService.java:323


def _golden(golden_dir):
    return np.load(os.path.join(golden_dir, "filters.npz"))


def test_oracle_filters_golden(golden_dir):
    g = _golden(golden_dir)
    names = sorted(k[3:] for k in g.files if k.startswith("in/"))
    assert len(names) >= 4
    for n in names:
        im = g["in/" + n]
        assert np.array_equal(orc.laplacian_sharpen(im, TAPS.reshape(9, 1)), g["sharp_k91/" + n]), n
        assert np.array_equal(orc.laplacian_sharpen(im, TAPS.reshape(3, 3)), g["sharp_k33/" + n]), n
        assert np.array_equal(orc.bgr2gray(im), g["gray/" + n]), n
        for k in (3, 5, 7, 11):
            assert np.array_equal(orc.median_blur(g["gray/" + n], k), g["median%d/%s" % (k, n)]), (n, k)


@pytest.mark.gpu
def test_gpu_filters_golden_and_oracle(golden_dir):
    g = _golden(golden_dir)
    with mseg.Context(0) as ctx:
        gi = mseg.GpuImgproc(ctx)
        for n in sorted(k[3:] for k in g.files if k.startswith("in/")):
            im = g["in/" + n]
            assert np.array_equal(gi.sharpenLaplacian(im, TAPS.reshape(9, 1)), g["sharp_k91/" + n]), n
            assert np.array_equal(gi.sharpenLaplacian(im, TAPS.reshape(3, 3)), g["sharp_k33/" + n]), n
            assert np.array_equal(gi.cvtColorBGR2GRAY(im), g["gray/" + n]), n
            for k in (3, 5, 7, 11):
                assert np.array_equal(gi.medianBlur(g["gray/" + n], k), g["median%d/%s" % (k, n)]), (n, k)
        im = orc.synth_bgr(640, 360, 3)
        gray = orc.bgr2gray(im)
        assert np.array_equal(gi.cvtColorBGR2GRAY(im), gray)
        for k in (1, 3, 7, 11, 25):                      # calculateSizeOfSquareBlurMask gives 7 @512^2, 11 @1080p
            assert np.array_equal(gi.medianBlur(gray, k), orc.median_blur(gray, k)), k
        for shape in ((9, 1), (3, 3), (1, 9)):
            assert np.array_equal(gi.sharpenLaplacian(im, TAPS.reshape(shape)), orc.laplacian_sharpen(im, TAPS.reshape(shape)))
        with pytest.raises(mseg.CvException):
            gi.medianBlur(gray, 4)
        with pytest.raises(mseg.CvException):
            gi.sharpenLaplacian(im, TAPS.reshape(9, 1)[:8])
