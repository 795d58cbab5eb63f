"""CPU: host-side strip-sharding logic (opencv-msegment_b200/sharded.py), including a world_size-2 gloo run in which
two ranks label their strips (with the oracle standing in for the device kernels), exchange seam pairs with
all_gather and must reproduce the unsharded labelling exactly."""
import os
import socket
import sys

import numpy as np
import pytest

import msegment_b200 as mseg
from oracle import oracle as orc



def _sharded():
    import importlib
    return importlib.import_module("opencv_msegment_b200.sharded")


def test_plan_strips_alignment():
    s = _sharded()
    for h, n, ml in ((1080, 4, 1), (16384, 8, 1), (517, 3, 2), (64, 8, 3)):
        strips = s.plan_strips(h, n, ml)
        assert strips[0][0] == 0 and strips[-1][1] == h and len(strips) == n
        for (a0, a1), (b0, b1) in zip(strips, strips[1:]):
            assert a1 == b0 and a0 < a1 and b0 % (1 << ml) == 0
    with pytest.raises(ValueError):
        s.plan_strips(5, 8, 1)
    h0, h1 = s.halo_range(512, 1024, 2000, 134, 1)
    assert h0 % 2 == 0 and h0 <= 512 - 134 and h1 == 1024 + 134


def test_resolve_pairs():
    s = _sharded()
    f, t = s.resolve_pairs(np.array([[10, 7], [7, 3], [20, 21], [3, 10], [40, 40]]))
    m = dict(zip(f.tolist(), t.tolist()))
    assert m == {7: 3, 10: 3, 21: 20}
    assert list(f) == sorted(f)
    f, t = s.resolve_pairs(np.zeros((0, 2), np.int32))
    assert len(f) == 0 and len(t) == 0


def test_first_pixel_roundtrip():
    s = _sharded()
    im = orc.synth_bgr(90, 70, 3)
    f = orc.meanshift_filter(im, 5, 10, 1)
    n, lab = orc.label_regions(f, 2)
    fp = s.first_pixel_labels(lab)
    n2, back = s.dense_from_first_pixel(fp)
    assert n2 == n and np.array_equal(back, lab)
    ys, xs = np.nonzero(fp == np.arange(1, fp.size + 1).reshape(fp.shape))
    assert len(ys) == n                                    # exactly one self-pointing pixel per region


def _seam_pairs_np(up_bgr, up_lab, lo_bgr, lo_lab, d):
    close = (np.abs(up_bgr.astype(int) - lo_bgr.astype(int)) <= d).all(axis=1)
    a, b = up_lab[close], lo_lab[close]
    keep = a != b
    return np.stack([a[keep], b[keep]], axis=1).astype(np.int32)


def _rank_main(rank, world, port, q):
    import torch.distributed as dist
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    s = _sharded()
    w, h, d = 160, 120, 2
    f = orc.meanshift_filter(orc.synth_bgr(w, h, 11), 5, 10, 1)     # every rank can build the same filtered image
    strips = s.plan_strips(h, world, 1)
    r0, r1 = strips[rank]
    # strip labelling (device stand-in): dense labels of the strip -> 1 + global index of first pixel
    n, lab = orc.label_regions(np.ascontiguousarray(f[r0:r1]), d)
    fp = s.first_pixel_labels(lab)
    local = fp - 1
    fp = (r0 * w + (local // w) * w + local % w + 1).astype(np.int32)
    # seam with the strip above: needs its last row (colour + labels) -> exchange through all_gather of rows
    last_row = np.concatenate([f[r1 - 1].reshape(-1), fp[-1].astype(np.int32).view(np.uint8)])
    import torch
    rows = [torch.zeros(len(last_row), dtype=torch.uint8) for _ in range(world)]
    dist.all_gather(rows, torch.from_numpy(last_row.copy()))
    pairs = np.zeros((0, 2), np.int32)
    if rank > 0:
        up = rows[rank - 1].numpy()
        up_bgr = up[:3 * w].reshape(w, 3)
        up_lab = up[3 * w:].view(np.int32)
        pairs = _seam_pairs_np(up_bgr, up_lab, f[r0], fp[0], d)
    allp = s.allgather_pairs(dist, pairs)
    allp1 = s.allgather_pairs(dist, pairs, cap=w)              # single-collective form must agree
    assert np.array_equal(allp, allp1)
    frm, to = s.resolve_pairs(allp)
    if len(frm):
        idx = np.searchsorted(frm, fp)
        idx[idx >= len(frm)] = 0
        hit = frm[idx] == fp
        fp = np.where(hit, to[idx], fp).astype(np.int32)
    gathered = [None] * world
    dist.all_gather_object(gathered, fp)
    if rank == 0:
        full = np.concatenate(gathered, axis=0)
        n0, want = orc.label_regions(f, d)
        q.put(bool(np.array_equal(full, s.first_pixel_labels(want))))
    dist.destroy_process_group()


def test_two_rank_gloo_seam_resolution():
    import torch.multiprocessing as mp
    sock = socket.socket()
    sock.bind(("127.0.0.1", 0))
    port = sock.getsockname()[1]
    sock.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_rank_main, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(timeout=180)
        assert p.exitcode == 0
    assert q.get(timeout=10) is True


def _strip_local(colors, r0, r1, d):
    """Provisional strip labels as msg_label_strip_dev writes them: 1 + GLOBAL index of the strip-local region's first pixel."""
    h, w = colors.shape[:2]
    n, lab = orc.label_regions(np.ascontiguousarray(colors[r0:r1]), d)        # canonical 1..n inside the strip
    flat = lab.ravel()
    first = np.full(n + 1, -1, np.int64)
    idx = np.arange(flat.size)
    first[flat[::-1]] = idx[::-1]                                               # smallest index wins
    return n, (first[flat] + r0 * w + 1).reshape(lab.shape).astype(np.int32)


@pytest.mark.parametrize("w,h,n_strips,seed", [(64, 96, 3, 1), (37, 120, 5, 2), (90, 40, 2, 3), (50, 50, 4, 4)])
def test_resolve_dense_single_exchange(w, h, n_strips, seed):
    """Host logic of the single-exchange strip finalisation (sharded.resolve_dense) against the unsharded labelling."""
    sh = _sharded()
    rng = np.random.default_rng(seed)
    # a few flat colours in blobs: many regions cross the seams, some span several strips
    base = rng.integers(0, 3, (h // 4 + 1, w // 4 + 1))
    colors = (np.kron(base, np.ones((4, 4), np.int64))[:h, :w, None] * np.array([70, 40, 90])).astype(np.uint8)
    colors[rng.random((h, w)) < 0.05] = 255                                     # speckles: small regions
    d = 2
    want_n, want = orc.label_regions(np.ascontiguousarray(colors), d)
    strips = sh.plan_strips(h, n_strips, 0)
    prov, n_roots, ranks = [], [], []
    for (r0, r1) in strips:
        n, lab = _strip_local(colors, r0, r1, d)
        roots = np.unique(lab)                                                  # ascending label = ascending first pixel
        prov.append(lab); n_roots.append(n); ranks.append(roots)
    quads = []
    for s in range(1, len(strips)):
        up_c, lo_c = colors[strips[s][0] - 1].astype(int), colors[strips[s][0]].astype(int)
        close = (np.abs(up_c - lo_c) <= d).all(axis=1)
        a, b = prov[s - 1][-1][close], prov[s][0][close]
        ra = np.searchsorted(ranks[s - 1], a) + 1
        rb = np.searchsorted(ranks[s], b) + 1
        quads.append(np.stack([a, b, ra, rb], axis=1))
    frm, dense, offsets, frm_lo, total = sh.resolve_dense(np.concatenate(quads), n_roots, strips, w)
    assert total == want_n
    got = np.zeros((h, w), np.int32)
    for s, (r0, r1) in enumerate(strips):
        v = prov[s].astype(np.int64)
        j = np.searchsorted(frm, v)
        hit = (j < len(frm)) & (frm[np.minimum(j, max(len(frm) - 1, 0))] == v) if len(frm) else np.zeros(v.shape, bool)
        own = offsets[s] + np.searchsorted(ranks[s], v) - (j - frm_lo[s]) + 1
        got[r0:r1] = np.where(hit, dense[np.minimum(j, max(len(dense) - 1, 0))] if len(dense) else 0, own)
    assert np.array_equal(got, want)


def test_c_shard_plan_matches_python_planner():
    """msg_shard_plan_make (pure host C, what a Java / C host calls) gives the strips and halo ranges of the Python planner."""
    import msegment_b200 as mseg
    dev = mseg.device
    sh = mseg.pkg.sharded
    for (w, h, n, sp, ml) in [(16384, 16384, 8, 10, 1), (8192, 8192, 8, 10, 1), (4096, 4096, 2, 10, 1), (600, 518, 3, 10, 1),
                              (333, 400, 4, 6, 2), (257, 300, 2, 8, 0), (512, 96, 6, 4, 0), (1000, 1001, 7, 20, 3), (50, 64, 64, 2, 0)]:
        halo, strips, halos = dev.shard_plan(w, h, n, sp, ml)
        assert halo == dev.halo_rows(sp, ml)
        assert strips == sh.plan_strips(h, n, ml), (w, h, n, ml)
        assert halos == [sh.halo_range(r0, r1, h, halo, ml) for r0, r1 in strips]
        assert strips[0][0] == 0 and strips[-1][1] == h and all(a < b for a, b in strips)
    import pytest
    with pytest.raises(ValueError):
        dev.shard_plan(100, 10, 8, 10, 2)          # too small for 8 strips at alignment 4
    with pytest.raises(ValueError):
        dev.shard_plan(100, 1000, 65, 10, 0)       # more than MSG_MAX_STRIPS


def test_halo_rows_is_the_derived_bound():
    """msg_meanshift_halo_rows (pure host arithmetic, no GPU): need(l) = max(maxCount * ceil(sp_l), 2 * (need(l+1) + 3) + 2),
    rounded up to the pyramid phase + one unit (58 at defaults, not the 134 of a chained cone)."""
    import msegment_b200 as mseg
    hr = mseg.device.halo_rows
    assert hr(10, 1) == 60                       # max(50, 2 * (25 + 3) + 2 = 58) -> 58 + 2
    assert hr(10, 0) == 51                       # 5 * 10, + 1
    assert hr(10, 0, (1, 9, 0.0)) == 91          # COUNT only, 9 iterations
    assert hr(10, 1, (2, 0, 1.0)) == 60          # EPS only: OpenCV runs 5 iterations
    assert hr(20, 1) == 110                      # max(100, 2 * (50 + 3) + 2 = 108) -> 108 + 2
    assert hr(4, 3) == 104                       # levels 3..0: 5 -> 18 -> 44 -> 96, rounded to a multiple of 8, + 8
    assert hr(10, 9) == -1
