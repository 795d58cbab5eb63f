"""GPU: the round-2 additions of the C ABI -- 16-bit labels, download mask, pageable buffers through the pinned staging
ring (synchronous and asynchronous path), host registration, per-context options, per-frame statistics."""
import ctypes as C

import numpy as np
import pytest

import msegment_b200 as mseg
from oracle import oracle as orc

pytestmark = pytest.mark.gpu
L = mseg.lib
dev = mseg.device


@pytest.fixture(scope="module")
def ctx():
    with mseg.Context(0) as c:
        yield c


def _oracle_segment(im, sp, sr, lo, min_size, cd):
    f = orc.meanshift_filter(im, sp, sr, 1)
    n0, l0 = orc.label_regions(f, lo)
    n1, l1 = orc.merge_regions(f, l0, min_size, cd)
    return f, n1, l1


def test_labels_u16_and_download_mask(ctx):
    gi = mseg.GpuImgproc(ctx)
    im = orc.synth_bgr(333, 211, 41)
    f, n, lab = _oracle_segment(im, 6, 12, 2, 20, 6)
    out = gi.segment(im, 6, 12, 1, loDiff=2, minSize=20, colorDist=6, want=("labels",), labelsType=mseg.imgproc.CV_16U)
    assert set(out) == {"n_regions", "labels"} and out["labels"].dtype == np.uint16
    assert out["n_regions"] == n and np.array_equal(out["labels"].astype(np.int32), lab)
    before = ctx.stats()["d2h_bytes"]
    gi.segment(im, 6, 12, 1, loDiff=2, minSize=20, colorDist=6, want=("labels",), labelsType=mseg.imgproc.CV_16U)
    assert ctx.stats()["d2h_bytes"] - before == im.shape[0] * im.shape[1] * 2      # nothing else came back


def test_labels_u16_overflow_is_reported(ctx):
    gi = mseg.GpuImgproc(ctx)
    rng = np.random.default_rng(5)
    noise = rng.integers(0, 256, (300, 400, 3), dtype=np.uint8)       # lo_diff 0 on noise: ~one region per pixel
    with pytest.raises(mseg.CvException) as ei:
        gi.segment(noise, 1, 1, 0, loDiff=0, want=("labels",), labelsType=mseg.imgproc.CV_16U)
    assert ei.value.status == L.MSG_ERANGE
    out = gi.segment(noise, 1, 1, 0, loDiff=0, want=("labels",))       # the context is still usable, 32-bit labels fine
    assert out["n_regions"] > 65535


def test_pageable_strided_buffers_sync_path(ctx):
    """numpy arrays are pageable: uploads and downloads go through the pinned ring, also for ROI views with row steps, and
    for images larger than the ring (4 x 4 MiB)."""
    gi = mseg.GpuImgproc(ctx)
    big = orc.synth_bgr(2600, 2300, 9)                                  # 17.9 MB > 16 MiB ring
    base = ctx.stats()["staged_bytes"]
    got = gi.pyrMeanShiftFiltering(big, 3, 8, 0)
    assert np.array_equal(got, orc.meanshift_filter(big, 3, 8, 0))
    assert ctx.stats()["staged_bytes"] - base == 2 * big.nbytes
    canvas = np.full((300, 500, 3), 77, np.uint8)
    roi = canvas[17:217, 40:371]                                       # strided view: 200 x 331
    roi[:] = orc.synth_bgr(331, 200, 3)
    want = orc.meanshift_filter(np.ascontiguousarray(roi), 5, 10, 1)
    dst_canvas = np.full((260, 420, 3), 5, np.uint8)
    dst = dst_canvas[30:230, 50:381]
    gi.pyrMeanShiftFiltering(roi, 5, 10, 1, dst=dst)
    assert np.array_equal(dst, want)
    dst[:] = 5
    assert (dst_canvas == 5).all()                                      # nothing outside the ROI was written


def test_dst_with_bad_strides_is_rejected(ctx):
    gi = mseg.GpuImgproc(ctx)
    im = orc.synth_bgr(64, 48, 1)
    dst = np.empty((48, 64, 3), np.uint8)
    with pytest.raises(mseg.CvException):
        gi.pyrMeanShiftFiltering(im, 3, 8, 0, dst=dst[::-1])           # negative row stride
    with pytest.raises(mseg.CvException):
        gi.pyrMeanShiftFiltering(im, 3, 8, 0, dst=np.empty((48, 64, 6), np.uint8)[:, :, ::2])   # non-contiguous pixels


@pytest.mark.parametrize("labels16", [False, True])
def test_async_pipeline_with_pageable_buffers(ctx, labels16):
    """msg_submit_segment / msg_wait with plain numpy (pageable) sources and destinations: the upload is staged, the
    outputs land in the frame's pinned staging and reach the caller's arrays inside msg_wait."""
    w, h, count = 320, 240, 11
    frames = [orc.synth_bgr(w, h, 500 + i) for i in range(count)]
    filt = [np.zeros((h, w, 3), np.uint8) for _ in range(count)]
    labs = [np.zeros((h, w), np.uint16 if labels16 else np.int32) for _ in range(count)]
    prm = dev.params(sp=6, sr=12, lo_diff=2, min_size=20, color_dist=6, render_depth=-1, labels_type=1 if labels16 else 0)
    tickets, out_n = [], []
    for i in range(count):
        if len(tickets) == 3:
            out_n.append(dev.wait(ctx, tickets.pop(0)))
        tickets.append(dev.submit_segment(ctx, frames[i].ctypes.data, 3 * w, w, h, prm, filt[i].ctypes.data, 3 * w,
                                          labs[i].ctypes.data, labs[i].strides[0]))
    out_n += [dev.wait(ctx, t) for t in tickets]
    for i in range(count):
        f, n, lab = _oracle_segment(frames[i], 6, 12, 2, 20, 6)
        assert out_n[i] == n and np.array_equal(filt[i], f), i
        assert np.array_equal(labs[i].astype(np.int32), lab), i
    assert ctx.stats()["ms_active_items"] > 0


def test_registered_host_range_is_used_in_place(ctx):
    gi = mseg.GpuImgproc(ctx)
    im = orc.synth_bgr(400, 300, 12)
    ctx.register_host(im)
    try:
        base = ctx.stats()["staged_bytes"]
        out = gi.segment(im, 6, 12, 1, loDiff=-1, want=("filtered",))
        staged = ctx.stats()["staged_bytes"] - base
        assert staged == im.nbytes                                     # only the (pageable) output was staged, not the input
        assert np.array_equal(out["filtered"], orc.meanshift_filter(im, 6, 12, 1))
    finally:
        ctx.unregister_host(im)
    with pytest.raises(mseg.CvException):
        ctx.unregister_host(im)


def test_options_roundtrip_and_gray_compat(ctx):
    gi = mseg.GpuImgproc(ctx)
    assert ctx.get_option("gray_compat") == 0 and ctx.get_option("staging") == 1
    with pytest.raises(mseg.CvException):
        ctx.set_option("no_such_option", 1)
    rng = np.random.default_rng(3)
    im = rng.integers(0, 256, (97, 131, 3), dtype=np.uint8)
    assert np.array_equal(gi.cvtColorBGR2GRAY(im), orc.bgr2gray(im))
    ctx.set_option("gray_compat", 1)
    try:
        got = gi.cvtColorBGR2GRAY(im)
        b, g, r = (im[..., k].astype(np.int64) for k in range(3))
        want = ((1868 * b + 9617 * g + 4899 * r + 8192) >> 14).astype(np.uint8)      # OpenCV 3.4.2 scalar path
        assert np.array_equal(got, want)
        assert np.array_equal(got, orc.bgr2gray(im, compat342=True))
        assert np.abs(got.astype(int) - orc.bgr2gray(im).astype(int)).max() <= 1
    finally:
        ctx.set_option("gray_compat", 0)


def test_merge_large_path_forced_on_small_image(ctx):
    """merge_small_max = 0 sends a small image through the cooperative-grid merge rounds; both paths must give the oracle's
    labels."""
    gi = mseg.GpuImgproc(ctx)
    im = orc.synth_bgr(400, 260, 23)
    f = orc.meanshift_filter(im, 6, 12, 1)
    n0, l0 = orc.label_regions(f, 2)
    want = orc.merge_regions(f, l0, 30, 8)
    ctx.set_option("merge_small_max", 0)
    try:
        got = gi.mergeRegions(f, l0, 30, 8)
    finally:
        ctx.set_option("merge_small_max", -1)
    assert got[0] == want[0] and np.array_equal(got[1], want[1])
    got = gi.mergeRegions(f, l0, 30, 8)
    assert got[0] == want[0] and np.array_equal(got[1], want[1])


@pytest.mark.parametrize("case", [(400, 260, 23, 6, 12, 30, 8), (517, 333, 4, 4, 6, 12, 5), (300, 200, 9, 3, 4, 0, 6), (640, 360, 5, 5, 8, 25, 0)])
def test_merge_medium_path_forced(ctx, case):
    """merge_medium_only = 1 sends images of <= 8191 labels through the single-CTA rounds kernel whose pair set lives in the global
    hash table (the regime of a 4K bench frame: 5.5 k regions); all three rounds kernels must give the oracle's labels, also on
    images with thousands of small regions and with one phase switched off."""
    w, h, seed, sp, sr, min_size, cd = case
    gi = mseg.GpuImgproc(ctx)
    im = orc.synth_bgr(w, h, seed)
    f = orc.meanshift_filter(im, sp, sr, 1)
    n0, l0 = orc.label_regions(f, 1)
    want = orc.merge_regions(f, l0, min_size, cd)
    for opt, val in (("merge_medium_only", 1), ("merge_small_max", 0), (None, None)):
        if opt:
            ctx.set_option(opt, val)
        try:
            got = gi.mergeRegions(f, l0, min_size, cd)
        finally:
            if opt:
                ctx.set_option(opt, 0 if opt == "merge_medium_only" else -1)
        assert got[0] == want[0] and np.array_equal(got[1], want[1]), (opt, n0)
    assert n0 <= 8191 or True


def test_tall_image_is_rejected_not_fatal(ctx):
    gi = mseg.GpuImgproc(ctx)
    tall = np.zeros((70000, 2, 3), np.uint8)
    with pytest.raises(mseg.CvException) as ei:
        gi.pyrMeanShiftFiltering(tall, 2, 5, 0)
    assert ei.value.status == L.MSG_EINVAL
    im = orc.synth_bgr(50, 40, 2)                                      # the context survives
    assert np.array_equal(gi.pyrMeanShiftFiltering(im, 3, 8, 0), orc.meanshift_filter(im, 3, 8, 0))


def test_two_contexts_with_different_sharpen_kernels():
    """The sharpen taps travel as a kernel argument: two contexts running different kernels concurrently cannot see each
    other's taps (they shared one __constant__ symbol in round 1)."""
    import threading
    im = orc.synth_bgr(640, 480, 8)
    k1 = np.array([[1, 1, 1], [1, -8, 1], [1, 1, 1]], np.int8)
    k2 = np.array([[1], [1], [1], [1], [-8], [1], [1], [1], [1]], np.int8)
    want = {0: orc.laplacian_sharpen(im, k1), 1: orc.laplacian_sharpen(im, k2)}
    bad = []

    def worker(which):
        with mseg.Context(0) as c:
            gi = mseg.GpuImgproc(c)
            for _ in range(40):
                if not np.array_equal(gi.sharpenLaplacian(im, k1 if which == 0 else k2), want[which]):
                    bad.append(which)
                    return
    ts = [threading.Thread(target=worker, args=(k,)) for k in (0, 1)]
    [t.start() for t in ts]
    [t.join() for t in ts]
    assert not bad


def test_merge_labels_canonical_option(ctx):
    """Option labels_canonical: msg_merge_regions_dev skips the validation / renumbering passes when the caller vouches for
    canonical labels (the output of msg_label_regions_dev) and passes their count; same result as the validating call."""
    torch = pytest.importorskip("torch")
    dev = mseg.device
    for (w, h, seed) in ((640, 360, 11), (517, 233, 12)):
        im = orc.synth_bgr(w, h, seed)
        filt = orc.meanshift_filter(im, 8, 10, 1)
        d_f = torch.from_numpy(filt).cuda()
        lab = torch.empty((h, w), dtype=torch.int32, device="cuda")
        cnt = torch.zeros(4, dtype=torch.int32, device="cuda")
        torch.cuda.synchronize()
        dev.label_regions(ctx, d_f.data_ptr(), 3 * w, lab.data_ptr(), 4 * w, w, h, 2, cnt.data_ptr())
        ctx.synchronize()
        n0 = int(cnt[0].item())
        want_n, want = orc.merge_regions(filt, lab.cpu().numpy(), 40, 9)
        a, b = lab.clone(), lab.clone()
        torch.cuda.synchronize()                     # torch ops run on torch's stream, the C-ABI calls on the context's
        dev.merge_regions(ctx, d_f.data_ptr(), 3 * w, a.data_ptr(), 4 * w, w, h, 40, 9, cnt.data_ptr())
        ctx.synchronize()
        assert int(cnt[0].item()) == want_n and np.array_equal(a.cpu().numpy(), want)
        ctx.set_option("labels_canonical", 1)
        try:
            base = ctx.stats()["kernel_launches"]
            cnt[0] = n0
            torch.cuda.synchronize()
            dev.merge_regions(ctx, d_f.data_ptr(), 3 * w, b.data_ptr(), 4 * w, w, h, 40, 9, cnt.data_ptr())
            ctx.synchronize()
            trusted_launches = ctx.stats()["kernel_launches"] - base
            assert int(cnt[0].item()) == want_n and np.array_equal(b.cpu().numpy(), want)
            with pytest.raises(mseg.CvException):
                dev.merge_regions(ctx, d_f.data_ptr(), 3 * w, b.data_ptr(), 4 * w, w, h, 40, 9, 0)
        finally:
            ctx.set_option("labels_canonical", 0)
        assert trusted_launches <= 8
