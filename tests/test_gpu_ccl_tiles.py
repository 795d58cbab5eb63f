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
