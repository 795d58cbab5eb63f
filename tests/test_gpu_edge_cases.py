"""GPU (B200): edge cases of the C ABI -- strided (non-continuous) Mats, background labels in the merge stage,
degenerate images, randomised parameter draws, asynchronous submit/wait, device-resident entry points."""
import ctypes as C

import numpy as np
import pytest

import msegment_b200 as mseg
from oracle import oracle as orc

pytestmark = pytest.mark.gpu
L = mseg.lib


@pytest.fixture(scope="module")
def ctx():
    c = mseg.Context(0)
    yield c
    c.close()


def test_strided_mats_through_the_abi(ctx):
    lib = ctx._lib
    w, h = 150, 90
    big = np.zeros((h + 7, w + 13, 3), np.uint8)
    big[3:3 + h, 5:5 + w] = orc.synth_bgr(w, h, 21)
    src = big[3:3 + h, 5:5 + w]                                   # ROI view: step > 3*w
    dst_big = np.full((h, w + 9, 3), 7, np.uint8)
    dst = dst_big[:, 2:2 + w]
    ctx.check(lib.msg_meanshift_filter(ctx._h, src.ctypes.data, src.strides[0], dst.ctypes.data, dst.strides[0], w, h,
                                       8.0, 14.0, 1, 3, 5, 1.0))
    want = orc.meanshift_filter(np.ascontiguousarray(src), 8, 14, 1)
    assert np.array_equal(dst, want)
    assert (dst_big[:, :2] == 7).all() and (dst_big[:, 2 + w:] == 7).all()      # nothing written outside the ROI
    lab_big = np.full((h, w + 5), -9, np.int32)
    lab = lab_big[:, 1:1 + w]
    n = C.c_int32()
    ctx.check(lib.msg_label_regions(ctx._h, dst.ctypes.data, dst.strides[0], lab.ctypes.data, lab.strides[0], w, h, 2, 2, 4,
                                    C.byref(n)))
    n0, l0 = orc.label_regions(want, 2)
    assert n.value == n0 and np.array_equal(lab, l0) and (lab_big[:, 0] == -9).all() and (lab_big[:, 1 + w:] == -9).all()
    ctx.check(lib.msg_merge_regions(ctx._h, dst.ctypes.data, dst.strides[0], lab.ctypes.data, lab.strides[0], w, h, 25, 6,
                                    C.byref(n)))
    n1, l1 = orc.merge_regions(want, l0, 25, 6)
    assert n.value == n1 and np.array_equal(lab, l1) and (lab_big[:, 0] == -9).all()


def test_merge_keeps_background_and_handles_unordered_labels(ctx):
    gi = mseg.GpuImgproc(ctx)
    rng = np.random.default_rng(4)
    im = orc.synth_bgr(180, 120, 8)
    m = (rng.random((120, 180)) < 0.8).astype(np.uint8) * 255
    n, cc = orc.connected_components(m, 8)                       # label 0 = background
    perm = rng.permutation(n - 1) + 1                            # scramble the numbering: not canonical any more
    scr = np.where(cc > 0, perm[np.maximum(cc, 1) - 1], 0).astype(np.int32)
    for min_size, cd in ((0, 0), (30, 0), (10, 12)):
        n0, l0 = orc.merge_regions(im, scr, min_size, cd)
        n1, l1 = gi.mergeRegions(im, scr, min_size, cd)
        assert n0 == n1 and np.array_equal(l0, l1), (min_size, cd)
        assert ((l1 == 0) == (cc == 0)).all()
    with pytest.raises(mseg.CvException):                         # labels must be <= width*height
        gi.mergeRegions(im, scr + np.where(scr > 0, 10**7, 0).astype(np.int32), 10, 0)


def test_every_pixel_its_own_region(ctx):
    gi = mseg.GpuImgproc(ctx)
    rng = np.random.default_rng(1)
    noise = rng.integers(0, 256, (64, 96, 3), dtype=np.uint8)
    n, lab = gi.labelRegions(noise, 0, 0, 4)
    n0, l0 = orc.label_regions(noise, 0)
    assert n == n0 and np.array_equal(lab, l0)
    n1, l1 = gi.mergeRegions(noise, lab, 4, 0)
    n2, l2 = orc.merge_regions(noise, l0, 4, 0)
    assert n1 == n2 and np.array_equal(l1, l2)


def test_randomised_parameters_vs_oracle(ctx):
    gi = mseg.GpuImgproc(ctx)
    rng = np.random.default_rng(20261018)
    for k in range(14):
        w, h = int(rng.integers(1, 260)), int(rng.integers(1, 200))
        sp = float(rng.choice([0.5, 1, 2.5, 3, 4.7, 6, 9.5, 10, 13]))
        sr = float(rng.choice([0.0, 1.5, 4, 8, 17.3, 30, 60, 260]))
        ml = int(rng.integers(0, 4))
        term = [(3, 5, 1.0), (1, int(rng.integers(1, 9)), 0.0), (2, 0, float(rng.integers(0, 30))), (3, 2, 0.0)][k % 4]
        im = orc.synth_bgr(w, h, 100 + k) if k % 3 else rng.integers(0, 256, (h, w, 3), dtype=np.uint8)
        want = orc.meanshift_filter(im, sp, sr, ml, term)
        got = gi.pyrMeanShiftFiltering(im, sp, sr, ml, term)
        assert np.array_equal(got, want), (k, w, h, sp, sr, ml, term, int((got != want).any(axis=2).sum()))


def test_async_submit_wait_pinned(ctx):
    dev = mseg.device
    w, h = 320, 240
    frames = [orc.synth_bgr(w, h, 50 + i) for i in range(5)]
    nb, nl = w * h * 3, w * h * 4
    hs, hf, hl = dev.alloc_pinned(5 * nb), dev.alloc_pinned(5 * nb), dev.alloc_pinned(5 * nl)
    try:
        for i, f in enumerate(frames):
            C.memmove(hs + i * nb, f.ctypes.data, nb)
        prm = dev.params(sp=6, sr=12, lo_diff=2, min_size=20, color_dist=6, render_depth=-1)
        tickets = []
        out_n = []
        for i in range(5):
            if len(tickets) == L.MAX_INFLIGHT:
                out_n.append(dev.wait(ctx, tickets.pop(0)))
            tickets.append(dev.submit_segment(ctx, hs + i * nb, 3 * w, w, h, prm, hf + i * nb, 3 * w, hl + i * nl, 4 * w))
        out_n += [dev.wait(ctx, t) for t in tickets]
        with pytest.raises(mseg.CvException):
            dev.wait(ctx, 0)                                       # ticket already consumed
        for i, f in enumerate(frames):
            ff = orc.meanshift_filter(f, 6, 12, 1)
            n0, l0 = orc.label_regions(ff, 2)
            n1, l1 = orc.merge_regions(ff, l0, 20, 6)
            got_f = np.ctypeslib.as_array((C.c_uint8 * nb).from_address(hf + i * nb)).reshape(h, w, 3)
            got_l = np.ctypeslib.as_array((C.c_int32 * (w * h)).from_address(hl + i * nl)).reshape(h, w)
            assert np.array_equal(got_f, ff) and np.array_equal(got_l, l1) and out_n[i] == n1, i
    finally:
        for p in (hs, hf, hl):
            dev.free_pinned(p)


def test_async_pipeline_graph_replay_matches_blocking_calls(ctx):
    """msg_submit_segment replays the frame's kernel sequence as a CUDA graph from the third use of a slot on: many
    different frames through the same slots, with a geometry change and a parameter change in between (re-capture),
    must give exactly what the blocking msg_segment call gives (itself checked against the oracle elsewhere)."""
    dev = mseg.device
    rng = np.random.default_rng(11)
    plan = [((320, 240), dict(sp=6, sr=12, lo_diff=2, min_size=20, color_dist=6), 14),
            ((200, 333), dict(sp=6, sr=12, lo_diff=2, min_size=20, color_dist=6), 9),     # geometry change
            ((200, 333), dict(sp=4, sr=20, lo_diff=3, min_size=0, color_dist=9), 9),      # parameter change
            ((320, 240), dict(sp=6, sr=12, lo_diff=2, min_size=20, color_dist=6), 6)]     # back to the first one
    with mseg.Context(0) as ref_ctx:
        ref = mseg.GpuImgproc(ref_ctx)
        for (w, h), kw, count in plan:
            nb, nl = w * h * 3, w * h * 4
            hs, hf, hl = dev.alloc_pinned(count * nb), dev.alloc_pinned(count * nb), dev.alloc_pinned(count * nl)
            try:
                frames = [orc.synth_bgr(w, h, 300 + i) if i % 3 else rng.integers(0, 256, (h, w, 3), dtype=np.uint8)
                          for i in range(count)]
                for i, f in enumerate(frames):
                    C.memmove(hs + i * nb, f.ctypes.data, nb)
                prm = dev.params(render_depth=-1, **kw)
                tickets, out_n = [], []
                for i in range(count):
                    if len(tickets) == 2:
                        out_n.append(dev.wait(ctx, tickets.pop(0)))
                    tickets.append(dev.submit_segment(ctx, hs + i * nb, 3 * w, w, h, prm, hf + i * nb, 3 * w, hl + i * nl, 4 * w))
                out_n += [dev.wait(ctx, t) for t in tickets]
                for i, f in enumerate(frames):
                    want = ref.segment(f, sp=kw["sp"], sr=kw["sr"], loDiff=kw["lo_diff"], minSize=kw["min_size"],
                                       colorDist=kw["color_dist"], renderDepth=-1, want=("filtered", "labels"))
                    got_f = np.ctypeslib.as_array((C.c_uint8 * nb).from_address(hf + i * nb)).reshape(h, w, 3)
                    got_l = np.ctypeslib.as_array((C.c_int32 * (w * h)).from_address(hl + i * nl)).reshape(h, w)
                    assert np.array_equal(got_f, want["filtered"]), (w, h, i)
                    assert np.array_equal(got_l, want["labels"]) and out_n[i] == want["n_regions"], (w, h, i)
            finally:
                for ptr in (hs, hf, hl):
                    dev.free_pinned(ptr)


def test_device_entry_points_with_torch():
    torch = pytest.importorskip("torch")
    dev = mseg.device
    c = mseg.Context(0)
    c.set_stream(torch.cuda.current_stream().cuda_stream)
    w, h = 211, 157
    im = orc.synth_bgr(w, h, 77)
    src = torch.from_numpy(im).cuda()
    filt = torch.empty_like(src)
    lab = torch.empty((h, w), dtype=torch.int32, device="cuda")
    ren = torch.empty_like(src)
    n = torch.zeros(1, dtype=torch.int32, device="cuda")
    dev.meanshift(c, src.data_ptr(), 3 * w, filt.data_ptr(), 3 * w, w, h, 7, 11, 2)
    dev.label_regions(c, filt.data_ptr(), 3 * w, lab.data_ptr(), 4 * w, w, h, 3, n.data_ptr())
    f = orc.meanshift_filter(im, 7, 11, 2)
    n0, l0 = orc.label_regions(f, 3)
    assert np.array_equal(filt.cpu().numpy(), f) and int(n.item()) == n0 and np.array_equal(lab.cpu().numpy(), l0)
    dev.merge_regions(c, filt.data_ptr(), 3 * w, lab.data_ptr(), 4 * w, w, h, 15, 5, n.data_ptr())
    n1, l1 = orc.merge_regions(f, l0, 15, 5)
    assert int(n.item()) == n1 and np.array_equal(lab.cpu().numpy(), l1)
    dev.render_labels(c, lab.data_ptr(), 4 * w, ren.data_ptr(), 3 * w, w, h, 10)
    assert np.array_equal(ren.cpu().numpy(), orc.render_labels(l1, 10))
    mask = torch.from_numpy((np.random.default_rng(2).random((h, w)) < 0.55).astype(np.uint8) * 255).cuda()
    dev.connected_components(c, mask.data_ptr(), w, lab.data_ptr(), 4 * w, w, h, 8, n.data_ptr())
    n2, l2 = orc.connected_components(mask.cpu().numpy(), 8)
    assert int(n.item()) == n2 and np.array_equal(lab.cpu().numpy(), l2)
    c.close()


@pytest.mark.parametrize("w,h,sp,sr,ml", [(260, 180, 40, 20, 0),      # wide windows: runtime-width tile kernel, drift shrunk to fit smem
                                          (150, 110, 64, 15, 1),      # radius 64 at level 0, 32 at level 1
                                          (120, 90, 130, 25, 0),      # radius > 120: generic warp-per-pixel kernel for every pixel
                                          (140, 100, 12, 255, 1)])    # sr >= 254: sentinel cannot be used -> generic kernel
def test_extreme_parameters_vs_oracle(ctx, w, h, sp, sr, ml):
    gi = mseg.GpuImgproc(ctx)
    im = orc.synth_bgr(w, h, 31)
    want = orc.meanshift_filter(im, sp, sr, ml)
    got = gi.pyrMeanShiftFiltering(im, sp, sr, ml)
    assert np.array_equal(got, want), int((got != want).any(axis=2).sum())
