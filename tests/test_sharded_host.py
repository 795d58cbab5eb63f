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
