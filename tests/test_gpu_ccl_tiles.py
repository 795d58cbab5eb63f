"""GPU: the tile-local union-find labelling (128 x 16 tiles, shared-memory unions + border pass + bitmap ranking) against the
oracle on sizes around the tile boundaries, both predicates, both connectivities, and against the round-1 row-run path."""
import numpy as np
import pytest

import msegment_b200 as mseg
from oracle import oracle as orc

pytestmark = pytest.mark.gpu

SIZES = [(1, 1), (1, 40), (40, 1), (127, 15), (128, 16), (129, 17), (255, 33), (256, 32), (257, 31), (300, 200), (640, 97),
         (1000, 40), (31, 300)]


@pytest.fixture(scope="module")
def gi():
    with mseg.Context(0) as c:
        yield mseg.GpuImgproc(c)


def _images(w, h, rng):
    """colour images with long thin structures, plateaus and noise, so that components cross many tiles"""
    yy, xx = np.mgrid[0:h, 0:w]
    stripes = (((xx + 2 * yy) // 7) % 3 * 40).astype(np.uint8)
    spiral = ((np.hypot(xx - w / 2, yy - h / 2) + 6 * np.arctan2(yy - h / 2, xx - w / 2)) % 12 < 6).astype(np.uint8) * 90
    out = []
    for base in (stripes, spiral, np.zeros((h, w), np.uint8)):
        im = np.repeat(base[..., None], 3, axis=2) + rng.integers(0, 4, (h, w, 3), dtype=np.uint8)
        out.append(np.ascontiguousarray(im.astype(np.uint8)))
    out.append(rng.integers(0, 256, (h, w, 3), dtype=np.uint8))
    return out


@pytest.mark.parametrize("w,h", SIZES)
def test_colour_labels_vs_oracle(gi, w, h):
    rng = np.random.default_rng(w * 1000 + h)
    for im in _images(w, h, rng):
        for d, conn in ((0, 4), (2, 4), (3, 8), (40, 4)):
            n0, l0 = orc.label_regions(im, d, conn)
            n1, l1 = gi.labelRegions(im, d, d, conn)
            assert n0 == n1 and np.array_equal(l0, l1), (w, h, d, conn, n0, n1)


@pytest.mark.parametrize("w,h", SIZES)
def test_binary_components_vs_oracle(gi, w, h):
    rng = np.random.default_rng(w * 77 + h)
    yy, xx = np.mgrid[0:h, 0:w]
    masks = [(rng.random((h, w)) < p).astype(np.uint8) * 255 for p in (0.3, 0.55, 0.75, 0.95)]
    masks.append((((xx // 3 + yy // 5) % 2) * 255).astype(np.uint8))
    masks.append(((np.hypot(xx - w / 2, yy - h / 2) + 5 * np.arctan2(yy - h / 2, xx - w / 2)) % 9 < 3).astype(np.uint8))
    masks.append(np.full((h, w), 7, np.uint8))
    masks.append(np.zeros((h, w), np.uint8))
    for m in masks:
        for conn in (4, 8):
            n0, l0 = orc.connected_components(m, conn)
            n1, l1 = gi.connectedComponents(m, conn)
            assert n0 == n1 and np.array_equal(l0, l1), (w, h, conn, n0, n1)


def test_tile_path_equals_legacy_path_large(gi):
    im = orc.synth_bgr(1920, 1080, 2)
    f = gi.pyrMeanShiftFiltering(im, 10, 10, 1)
    got = gi.labelRegions(f, 2, 2, 4)
    gi.ctx.set_option("ccl_legacy", 1)
    try:
        want = gi.labelRegions(f, 2, 2, 4)
        m_want = gi.connectedComponents((f[..., 1] > 128).astype(np.uint8), 8)
    finally:
        gi.ctx.set_option("ccl_legacy", 0)
    assert got[0] == want[0] and np.array_equal(got[1], want[1])
    m_got = gi.connectedComponents((f[..., 1] > 128).astype(np.uint8), 8)
    assert m_got[0] == m_want[0] and np.array_equal(m_got[1], m_want[1])
    n0, l0 = orc.label_regions(f, 2)
    assert got[0] == n0 and np.array_equal(got[1], l0)


def test_worst_case_single_component_snake(gi):
    """one serpentine component crossing every tile many times: long union chains across the border pass"""
    w, h = 1030, 260
    m = np.zeros((h, w), np.uint8)
    m[::2, :] = 1
    m[1::4, -1] = 1
    m[3::4, 0] = 1
    n1, l1 = gi.connectedComponents(m, 4)
    n0, l0 = orc.connected_components(m, 4)
    assert n1 == n0 == 2 and np.array_equal(l1, l0)
    im = np.repeat((m * 200)[..., None], 3, axis=2).astype(np.uint8)
    n1, l1 = gi.labelRegions(im, 1, 1, 4)
    n0, l0 = orc.label_regions(im, 1, 4)
    assert n1 == n0 and np.array_equal(l1, l0)


QUAD_SIZES = [(4, 1), (8, 3), (124, 15), (128, 16), (132, 17), (256, 32), (260, 33), (640, 97), (1000, 40), (2052, 19)]


@pytest.mark.parametrize("w,h", QUAD_SIZES)
def test_quad_tile_kernel_vs_oracle_and_one_pixel_per_lane(gi, w, h):
    """widths that are multiples of 4 take ccl_tile4_kernel (four pixels per lane); option ccl_quad = 0 keeps the
    one-pixel-per-lane kernel: both must give the oracle's labels, on BGR input (PRED 2) and through the fused pipeline
    (PRED 0, the packed plane of the mean-shift output)"""
    rng = np.random.default_rng(w * 31 + h)
    for im in _images(w, h, rng):
        for d in (0, 2, 40):
            n0, l0 = orc.label_regions(im, d, 4)
            n1, l1 = gi.labelRegions(im, d, d, 4)
            gi.ctx.set_option("ccl_quad", 0)
            try:
                n2, l2 = gi.labelRegions(im, d, d, 4)
            finally:
                gi.ctx.set_option("ccl_quad", 1)
            assert n0 == n1 == n2 and np.array_equal(l0, l1) and np.array_equal(l0, l2), (w, h, d, n0, n1, n2)
    if w >= 8 and h >= 8:
        im = _images(w, h, rng)[0]
        a = gi.segment(im, 4, 6, 1, loDiff=2, minSize=0, colorDist=0)
        gi.ctx.set_option("ccl_quad", 0)
        try:
            b = gi.segment(im, 4, 6, 1, loDiff=2, minSize=0, colorDist=0)
        finally:
            gi.ctx.set_option("ccl_quad", 1)
        assert a["n_regions"] == b["n_regions"] and np.array_equal(a["labels"], b["labels"])
        n0, l0 = orc.label_regions(a["filtered"], 2, 4)
        assert a["n_regions"] == n0 and np.array_equal(a["labels"], l0)


def test_quad_tile_kernel_snake_and_roi_view(gi):
    """one serpentine component crossing every tile (long union chains inside the tiles and across the border pass), and a
    device-resident ROI views with and without 4-byte aligned rows"""
    w, h = 1032, 260
    m = np.zeros((h, w), np.uint8)
    m[::2, :] = 1
    m[1::4, -1] = 1
    m[3::4, 0] = 1
    im = np.repeat((m * 200)[..., None], 3, axis=2).astype(np.uint8)
    n1, l1 = gi.labelRegions(im, 1, 1, 4)
    n0, l0 = orc.label_regions(im, 1, 4)
    assert n1 == n0 and np.array_equal(l1, l0)
    # device-resident ROI views: (row step, byte offset) decide between the two kernels -- 4-byte aligned rows take the quad
    # kernel, anything else the one-pixel-per-lane kernel; the labels must not depend on it
    import torch
    dev = mseg.device
    rng = np.random.default_rng(5)
    for wfull, x0 in ((300, 0), (300, 4), (300, 1), (301, 0), (301, 3)):
        big = (rng.integers(0, 3, (70, wfull, 3), dtype=np.uint8) * 3).astype(np.uint8)
        w, h = 292, 64
        view = np.ascontiguousarray(big[3:3 + h, x0:x0 + w])
        n0, l0 = orc.label_regions(view, 2, 4)
        d_big = torch.from_numpy(big).cuda()
        d_lab = torch.empty((h, w), dtype=torch.int32, device="cuda")
        d_n = torch.zeros(4, dtype=torch.int32, device="cuda")
        ptr = d_big.data_ptr() + 3 * wfull * 3 + 3 * x0
        dev.label_regions(gi.ctx, ptr, 3 * wfull, d_lab.data_ptr(), 4 * w, w, h, 2, d_n.data_ptr())
        gi.ctx.synchronize()
        assert int(d_n[0].item()) == n0 and np.array_equal(d_lab.cpu().numpy(), l0), (wfull, x0)


@pytest.mark.parametrize("w,h", [(8, 8), (128, 16), (132, 70), (512, 130), (1920, 1080)])
def test_stats_column_strips_equal_row_chunks(gi, w, h):
    """the statistics pass of the merge as a walk down column strips (register-resident sums, pairs staged per warp) against
    the row-chunk pass (option merge_strips = 0) and the oracle, single-CTA rounds and the cooperative large path"""
    im = orc.synth_bgr(w, h, 11) if w >= 64 else np.random.default_rng(3).integers(0, 40, (h, w, 3), dtype=np.uint8)
    f = gi.pyrMeanShiftFiltering(im, 5, 8, 1)
    n0, l0 = gi.labelRegions(f, 2, 2, 4)
    want = orc.merge_regions(f, l0, 20, 12) if w * h <= 512 * 130 else None
    got = {}
    for strips in (1, 0):
        for small_max in (-1, 0):
            gi.ctx.set_option("merge_strips", strips)
            gi.ctx.set_option("merge_small_max", small_max)
            try:
                got[(strips, small_max)] = gi.mergeRegions(f, l0, 20, 12)
            finally:
                gi.ctx.set_option("merge_strips", 1)
                gi.ctx.set_option("merge_small_max", -1)
    ref = got[(0, -1)]
    for k, (n, lab) in got.items():
        assert n == ref[0] and np.array_equal(lab, ref[1]), k
    if want is not None:
        assert ref[0] == want[0] and np.array_equal(ref[1], want[1])
