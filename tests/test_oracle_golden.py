"""CPU: the oracle (oracle/msg_oracle.c) against the committed OpenCV golden vectors.

The vectors were produced by tests/golden/gen_golden.py from cv2 4.13.0 -- the stand-in for the
OpenCV natives the reference binds (pom.xml:39-43); the reference ships no tests of its own.
"""
import os

import numpy as np
import pytest

from oracle import oracle as orc


def _load(golden_dir, name):
    return np.load(os.path.join(golden_dir, name))


def test_meanshift_golden(golden_dir):
    g = _load(golden_dir, "meanshift.npz")
    params = g["params"]
    names = sorted(k[3:] for k in g.files if k.startswith("in/"))
    assert len(names) >= 6
    for name in names:
        im = g["in/" + name]
        for k, (sp, sr, ml, tt, tc, te) in enumerate(params):
            want = g["out/%s/%d" % (name, k)]
            got = orc.meanshift_filter(im, sp, sr, int(ml), (int(tt), int(tc), float(te)))
            assert np.array_equal(got, want), (name, k)


def test_pyramid_golden(golden_dir):
    g = _load(golden_dir, "pyramid.npz")
    for name in sorted(k[3:] for k in g.files if k.startswith("in/")):
        im = g["in/" + name]
        h, w = im.shape[:2]
        assert np.array_equal(orc.pyr_down(im), g["down/" + name]), name
        assert np.array_equal(orc.pyr_up(im, (2 * w, 2 * h)), g["up_even/" + name]), name
        assert np.array_equal(orc.pyr_up(im, (2 * w - 1, 2 * h - 1)), g["up_odd/" + name]), name


def test_label_regions_golden(golden_dir):
    g = _load(golden_dir, "labels.npz")
    for name in sorted(k[3:] for k in g.files if k.startswith("in/")):
        f = g["in/" + name]
        for d in (0, 2, 5):
            want = g["ff%d/%s" % (d, name)]
            n, lab = orc.label_regions(f, d)
            assert n == want.max() and np.array_equal(lab, want), (name, d)
        want8 = g["ff2c8/" + name]                      # floodFill with the 8-connectivity flag
        n, lab = orc.label_regions(f, 2, 8)
        assert n == want8.max() and np.array_equal(lab, want8), name


def test_connected_components_golden(golden_dir):
    g = _load(golden_dir, "labels.npz")
    ks = sorted(int(k[5:]) for k in g.files if k.startswith("mask/"))
    assert ks
    for k in ks:
        m = g["mask/%d" % k]
        for conn in (4, 8):
            n, lab = orc.connected_components(m, conn)
            assert n == int(g["ccn%d/%d" % (conn, k)]), (k, conn)
            assert np.array_equal(lab, g["cc%d/%d" % (conn, k)]), (k, conn)


def test_watershed_golden(golden_dir):
    g = _load(golden_dir, "watershed.npz")
    for k in range(3):
        got = orc.watershed(g["img/%d" % k], g["markers/%d" % k])
        assert np.array_equal(got, g["out/%d" % k]), k


def test_watershed_golden_extended(golden_dir):
    """27 more cv2.watershed cases (tests/golden/gen_watershed.py): degenerate sizes, noise, flat images, border seeds, negative
    input markers, and the markers of the reference's two pipelines on its sample images and on synthetic frames."""
    g = _load(golden_dir, "watershed2.npz")
    for name in g["names"]:
        got = orc.watershed(g["img/%s" % name], g["markers/%s" % name])
        assert np.array_equal(got, g["out/%s" % name]), name


def test_oracle_reproduces_cv2_whole_frame_digest_1080p(golden_dir):
    """The 1080p unit of tests/golden/fullsize_digests.json (cv2 on the whole frame) against the oracle; the 4K units are
    checked by tests/golden/campaign_digests.py (25 CPU-minutes, result in PROVENANCE.txt) and, on the GPU, by
    tests/test_gpu_fullsize_cv2.py."""
    import hashlib
    import json
    u = json.load(open(os.path.join(golden_dir, "fullsize_digests.json")))["units"]["c2_seed2"]
    im = orc.synth_bgr(u["w"], u["h"], u["seed"])
    f = orc.meanshift_filter(im, u["sp"], u["sr"], 1)
    assert hashlib.sha256(np.ascontiguousarray(f).tobytes()).hexdigest() == u["filtered_sha256"]
    n, lab = orc.label_regions(f, 2)
    assert n == u["n_regions"] and hashlib.sha256(np.ascontiguousarray(lab).tobytes()).hexdigest() == u["labels_sha256"]


def test_render_rule():
    # PictureService.java:928: 0 < index <= depth -> colour, else background
    rng = np.random.default_rng(1)
    lab = rng.integers(-2, 9, (20, 30)).astype(np.int32)
    r = orc.render_labels(lab, 5)
    on = (lab > 0) & (lab <= 5)
    assert np.array_equal(r, np.repeat(np.where(on, 255, 0).astype(np.uint8)[..., None], 3, axis=2))
    cols = rng.integers(0, 256, (5, 3)).astype(np.uint8)
    r2 = orc.render_labels(lab, 5, cols)
    want = np.zeros((20, 30, 3), np.uint8)
    want[on] = cols[lab[on] - 1]
    assert np.array_equal(r2, want)


def test_merge_properties():
    im = orc.synth_bgr(160, 120, 4)
    f = orc.meanshift_filter(im, 6, 12, 1)
    n, lab = orc.label_regions(f, 2)
    n0, l0 = orc.merge_regions(f, lab, 0, 0)
    assert n0 == n and np.array_equal(l0, lab)           # minSize=0, colorDist=0 -> identity
    n1, l1 = orc.merge_regions(f, lab, 30, 0)
    areas = np.bincount(l1.ravel())[1:]
    assert n1 == len(areas) and (areas.min() >= 30 or n1 == 1)
    n2, l2 = orc.merge_regions(f, l1, 30, 0)            # idempotent
    assert n2 == n1 and np.array_equal(l2, l1)
    # merged partition is a coarsening of the input partition
    pairs = np.unique(np.stack([lab.ravel(), l1.ravel()]), axis=1)
    assert len(np.unique(pairs[0])) == pairs.shape[1]
    n3, l3 = orc.merge_regions(f, lab, 30, 12)
    assert n3 <= n


@pytest.mark.parametrize("seed", [1, 2])
def test_oracle_vs_cv2_live(seed):
    cv2 = pytest.importorskip("cv2")
    rng = np.random.default_rng(seed)
    w, h = int(rng.integers(20, 90)), int(rng.integers(20, 90))
    im = orc.synth_bgr(w, h, seed)
    for sp, sr, ml in [(4, 9, 1), (6.5, 3, 2), (9, 30, 0)]:
        want = cv2.pyrMeanShiftFiltering(im, sp, sr, maxLevel=ml)
        assert np.array_equal(orc.meanshift_filter(im, sp, sr, ml), want)
