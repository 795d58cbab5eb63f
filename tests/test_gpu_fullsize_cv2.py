"""GPU (B200): every full-size configuration of BASELINE.json against the REAL OpenCV result on the WHOLE frame.

tests/golden/fullsize_digests.json (generated in the build container by tests/golden/gen_fullsize_digests.py with cv2, whole
frames, no strips) holds, per unit, the SHA-256 of cv2.pyrMeanShiftFiltering's output, 512 sampled pixels of it, and the
region count + SHA-256 of the label map of OpenCV's floodFill labelling loop.  The GPU must reproduce all of them bit for bit:
  config 2  1920x1080 seed 2                       config 3  3840x2160 seed 3, 9 (sp, sr) variants
  config 4  3840x2160 seeds 1000..1015 (the 16-frame subsample of the 256-frame batch), through the asynchronous batch path
"""
import hashlib
import json
import os

import numpy as np
import pytest

import msegment_b200 as mseg

pytestmark = pytest.mark.gpu
DIGESTS = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "fullsize_digests.json")


@pytest.fixture(scope="module")
def digests():
    if not os.path.exists(DIGESTS):
        pytest.fail("tests/golden/fullsize_digests.json is missing (run tests/golden/gen_fullsize_digests.py)")
    return json.load(open(DIGESTS))


@pytest.fixture(scope="module")
def ctx():
    with mseg.Context(0) as c:
        yield c


def _sha(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


def _check_unit(name, u, filtered, labels, n_regions, digests):
    rng = np.random.default_rng(digests["sample_seed"])
    ys, xs = rng.integers(0, u["h"], digests["n_samples"]), rng.integers(0, u["w"], digests["n_samples"])
    want = np.array(u["filtered_samples"], np.uint8).reshape(-1, 3)
    got = filtered[ys, xs]
    assert np.array_equal(got, want), (name, "sampled pixels differ", int((got != want).any(axis=1).sum()))
    assert _sha(filtered) == u["filtered_sha256"], (name, "filtered frame differs from cv2.pyrMeanShiftFiltering")
    assert n_regions == u["n_regions"], (name, n_regions, u["n_regions"])
    assert _sha(labels.astype(np.int32)) == u["labels_sha256"], (name, "labels differ from the cv2 floodFill loop")


def _units(digests, prefix):
    return sorted((k, v) for k, v in digests["units"].items() if k.startswith(prefix))


def test_config2_1080p_whole_frame_vs_cv2(ctx, digests):
    gi = mseg.GpuImgproc(ctx)
    for name, u in _units(digests, "c2_"):
        im = mseg.synth_bgr(u["w"], u["h"], u["seed"])
        out = gi.segment(im, u["sp"], u["sr"], 1, loDiff=2, want=("filtered", "labels"))
        _check_unit(name, u, out["filtered"], out["labels"], out["n_regions"], digests)


def test_config3_4k_sweep_whole_frames_vs_cv2(ctx, digests):
    gi = mseg.GpuImgproc(ctx)
    units = _units(digests, "c3_")
    assert len(units) == 9
    im = mseg.synth_bgr(3840, 2160, 3)
    for name, u in units:
        assert (u["w"], u["h"], u["seed"]) == (3840, 2160, 3)
        out = gi.segment(im, u["sp"], u["sr"], 1, loDiff=2, want=("filtered", "labels"))
        _check_unit(name, u, out["filtered"], out["labels"], out["n_regions"], digests)


def test_config4_batch_subsample_vs_cv2(ctx, digests):
    """The 16-frame subsample of config 4 through msg_submit_segment / msg_wait (pageable numpy buffers, 16-bit labels): the
    path the batch bench times."""
    dev = mseg.device
    units = _units(digests, "c4_")
    assert len(units) == 16
    w, h = 3840, 2160
    frames = [mseg.synth_bgr(w, h, u["seed"]) for _, u in units]
    filt = [np.empty((h, w, 3), np.uint8) for _ in units]
    labs = [np.empty((h, w), np.uint16) for _ in units]
    prm = dev.params(sp=10, sr=10, lo_diff=2, render_depth=-1, labels_type=1)
    tickets, counts = [], []
    for i in range(len(units)):
        if len(tickets) == 2:
            counts.append(dev.wait(ctx, tickets.pop(0)))
        tickets.append(dev.submit_segment(ctx, frames[i].ctypes.data, 3 * w, w, h, prm, filt[i].ctypes.data, 3 * w,
                                          labs[i].ctypes.data, 2 * w))
    counts += [dev.wait(ctx, t) for t in tickets]
    for i, (name, u) in enumerate(units):
        _check_unit(name, u, filt[i], labs[i], counts[i], digests)
