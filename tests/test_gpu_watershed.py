"""GPU: msg_watershed (exact cv::watershed, PictureService.java:908-911) against cv2 golden vectors and the oracle."""
import os

import numpy as np
import pytest

import msegment_b200 as mseg
from oracle import oracle as orc

pytestmark = pytest.mark.gpu
GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


@pytest.fixture(scope="module")
def gi():
    with mseg.Context(0) as c:
        yield mseg.GpuImgproc(c)


def test_watershed_cv2_goldens(gi):
    g = np.load(os.path.join(GOLDEN, "watershed.npz"))
    for k in range(3):
        mk = g["markers/%d" % k].copy()
        gi.watershed(g["img/%d" % k], mk)
        assert np.array_equal(mk, g["out/%d" % k]), k
    g = np.load(os.path.join(GOLDEN, "watershed2.npz"))
    for name in g["names"]:
        mk = g["markers/%s" % name].copy()
        gi.watershed(g["img/%s" % name], mk)
        want = g["out/%s" % name]
        assert np.array_equal(mk, want), (name, int((mk != want).sum()))


def test_watershed_strided_markers_in_place(gi):
    im = orc.synth_bgr(160, 120, 5)
    _, mk, _ = orc.shape_seeds(im)
    want = orc.watershed(im, mk.copy())
    canvas = np.full((140, 200), 12345, np.int32)
    roi = canvas[10:130, 20:180]
    roi[:] = mk
    gi.watershed(im, roi)
    assert np.array_equal(roi, want)
    roi[:] = 12345
    assert (canvas == 12345).all()


@pytest.mark.parametrize("w,h,seed", [(640, 360, 2), (1920, 1080, 2)])
def test_watershed_pipeline_markers_vs_oracle(gi, w, h, seed):
    """Both CLI pipelines end in watershed (PictureService.java:372, :457): markers from the GPU generators, flood on the GPU,
    compared with the oracle's flood of the same markers -- full 1080p included."""
    im = orc.synth_bgr(w, h, seed)
    n, mk = gi.shapeSeeds(im)
    want = orc.watershed(im, mk.copy())
    got = mk.copy()
    gi.watershed(im, got)
    assert np.array_equal(got, want), int((got != want).sum())
    n, mk, st = gi.colorSeeds(im, stages=True)
    want = orc.watershed(st["sharp"], mk.copy())
    got = mk.copy()
    gi.watershed(st["sharp"], got)
    assert np.array_equal(got, want), int((got != want).sum())
    assert set(np.unique(got)) <= set(range(-1, n + 1)) | {255}


def test_watershed_batch_device(gi):
    torch = pytest.importorskip("torch")
    dev = mseg.device
    ctx = gi.ctx
    ctx.set_stream(torch.cuda.current_stream().cuda_stream)
    try:
        w, h, count = 200, 150, 40
        ims = np.stack([orc.synth_bgr(w, h, 700 + i) for i in range(count)])
        rng = np.random.default_rng(1)
        mks = np.zeros((count, h, w), np.int32)
        for i in range(count):
            for s in range(1, 10):
                y, x = rng.integers(2, h - 4), rng.integers(2, w - 4)
                mks[i, y:y + 2, x:x + 3] = s
        d_im = torch.from_numpy(ims).cuda()
        d_mk = torch.from_numpy(mks).cuda()
        pops = torch.zeros(1, dtype=torch.int64, device="cuda")
        dev.watershed_batch(ctx, d_im.data_ptr(), 3 * w, 3 * w * h, d_mk.data_ptr(), 4 * w, 4 * w * h, w, h, count, pops.data_ptr())
        got = d_mk.cpu().numpy()
        for i in range(count):
            assert np.array_equal(got[i], orc.watershed(ims[i], mks[i].copy())), i
        assert int(pops.item()) > count * (w - 2) * (h - 2) * 0.9
    finally:
        ctx.set_stream(None)
