"""GPU (B200): strip sharding must be bit-identical to the unsharded result.  N strips are evaluated one after the
other on one GPU (each call sees only its rows + halo), which is how the multi-GPU path is checked without a cluster;
tools/shard_large_image.py runs the same calls with one process per GPU."""
import numpy as np
import pytest

import msegment_b200 as mseg
from oracle import oracle as orc

torch = pytest.importorskip("torch")
pytestmark = pytest.mark.gpu
dev = mseg.device
sh = mseg.pkg.sharded


def _sharded_run(ctx, src, w, h, n_strips, sp, sr, ml, lo):
    """src: torch uint8 [h,w,3] on the GPU.  Returns (filtered [h,w,3], labels [h,w] in first-pixel form)."""
    halo = dev.halo_rows(sp, ml)
    strips = sh.plan_strips(h, n_strips, ml)
    filt = torch.zeros_like(src)
    lab = torch.zeros((h, w), dtype=torch.int32, device="cuda")
    for (r0, r1) in strips:
        h0, h1 = sh.halo_range(r0, r1, h, halo, ml)
        rows = src[h0:h1].contiguous()                 # what a rank would hold: its strip + halo rows only
        out = torch.empty((r1 - r0, w, 3), dtype=torch.uint8, device="cuda")
        dev.meanshift_strip(ctx, rows.data_ptr(), 3 * w, h0, h1, out.data_ptr(), 3 * w, w, h, r0, r1, sp, sr, ml)
        filt[r0:r1] = out
    for (r0, r1) in strips:
        rows = filt[r0:r1].contiguous()
        l = torch.empty((r1 - r0, w), dtype=torch.int32, device="cuda")
        dev.label_strip(ctx, rows.data_ptr(), 3 * w, l.data_ptr(), 4 * w, w, r1 - r0, r0, w, lo)
        lab[r0:r1] = l
    pairs_all = []
    pairs = torch.zeros((w, 2), dtype=torch.int32, device="cuda")
    cnt = torch.zeros((1,), dtype=torch.int32, device="cuda")
    for (r0, r1) in strips[1:]:
        dev.seam_pairs(ctx, filt[r0 - 1].data_ptr(), lab[r0 - 1].data_ptr(), filt[r0].data_ptr(), lab[r0].data_ptr(), w, lo,
                       pairs.data_ptr(), cnt.data_ptr())
        ctx.synchronize()
        pairs_all.append(pairs[:int(cnt.item())].cpu().numpy().copy())
    frm, to = sh.resolve_pairs(np.concatenate(pairs_all) if pairs_all else np.zeros((0, 2), np.int32))
    if len(frm):
        d_from, d_to = torch.from_numpy(frm).cuda(), torch.from_numpy(to).cuda()
        dev.apply_label_map(ctx, lab.data_ptr(), 4 * w, w, h, d_from.data_ptr(), d_to.data_ptr(), len(frm))
    ctx.synchronize()
    return filt, lab, strips, to


@pytest.mark.parametrize("w,h,n,sp,sr,ml", [(600, 518, 3, 10, 10, 1), (333, 400, 4, 6, 15, 2), (257, 300, 2, 8, 12, 0),
                                           (420, 300, 5, 10, 10, 1)])
def test_strips_bit_identical(w, h, n, sp, sr, ml):
    ctx = mseg.Context(0)
    ctx.set_stream(torch.cuda.current_stream().cuda_stream)    # torch copies and C-ABI kernels on one stream
    im = orc.synth_bgr(w, h, 17)
    src = torch.from_numpy(im).cuda()
    full = torch.empty_like(src)
    dev.meanshift(ctx, src.data_ptr(), 3 * w, full.data_ptr(), 3 * w, w, h, sp, sr, ml)
    lab_full = torch.empty((h, w), dtype=torch.int32, device="cuda")
    dev.label_regions(ctx, full.data_ptr(), 3 * w, lab_full.data_ptr(), 4 * w, w, h, 2)
    ctx.synchronize()
    filt, lab, strips, to = _sharded_run(ctx, src, w, h, n, sp, sr, ml, 2)
    bad = (filt != full).any(dim=2)
    assert not bad.any(), "filtered differs at %d pixels, rows %s" % (int(bad.sum()), torch.nonzero(bad.any(dim=1)).flatten()[:8].tolist())
    want = sh.first_pixel_labels(lab_full.cpu().numpy())
    got = lab.cpu().numpy()
    assert np.array_equal(got, want), int((got != want).sum())
    # and the unsharded GPU result is the oracle's
    assert np.array_equal(full.cpu().numpy(), orc.meanshift_filter(im, sp, sr, ml))
    # dense global numbering: one context per simulated rank (each keeps its ranks between the three steps)
    ctxs = [mseg.Context(0) for _ in strips]
    for c in ctxs:
        c.set_stream(torch.cuda.current_stream().cuda_stream)
    cnt = torch.zeros(len(strips), dtype=torch.int32, device="cuda")
    for k, (r0, r1) in enumerate(strips):
        dev.strip_rank(ctxs[k], lab[r0:r1].data_ptr(), 4 * w, w, r1 - r0, r0, w, cnt[k:].data_ptr())
    counts = cnt.cpu().numpy().astype(np.int64)
    offsets = np.concatenate([[0], np.cumsum(counts)[:-1]])
    uniq_to = np.unique(to).astype(np.int32)
    dense_of_to = np.zeros(len(uniq_to), np.int32)
    for k, (r0, r1) in enumerate(strips):
        own = np.flatnonzero((uniq_to > r0 * w) & (uniq_to <= r1 * w))
        if len(own):
            q = torch.from_numpy(uniq_to[own]).cuda()
            o = torch.zeros(len(own), dtype=torch.int32, device="cuda")
            dev.strip_query_dense(ctxs[k], q.data_ptr(), len(own), w, r1 - r0, r0, w, int(offsets[k]), o.data_ptr())
            dense_of_to[own] = o.cpu().numpy()
    assert (dense_of_to > 0).all()
    d_lab, d_dense = torch.from_numpy(uniq_to).cuda(), torch.from_numpy(dense_of_to).cuda()
    for k, (r0, r1) in enumerate(strips):
        dev.strip_apply_dense(ctxs[k], lab[r0:r1].data_ptr(), 4 * w, w, r1 - r0, r0, w, int(offsets[k]), d_lab.data_ptr(),
                              d_dense.data_ptr(), len(uniq_to))
    torch.cuda.synchronize()
    assert int(counts.sum()) == int(lab_full.max().item())
    assert torch.equal(lab, lab_full), int((lab != lab_full).sum().item())      # identical to the unsharded dense numbering
    for c in ctxs:
        c.close()
    ctx.close()


@pytest.mark.parametrize("w,h,n,sp,sr,ml", [(600, 518, 3, 10, 10, 1), (333, 400, 4, 6, 15, 2), (420, 300, 5, 10, 10, 1),
                                           (512, 96, 6, 4, 30, 0)])
def test_strips_single_exchange_dense(w, h, n, sp, sr, ml):
    """The single-exchange form (msg_strip_rank_dev on provisional labels, msg_seam_quads_dev, one gather,
    sharded.resolve_dense, msg_strip_finalize_dense_dev): labels identical to the unsharded call's, integer for integer."""
    im = orc.synth_bgr(w, h, 23)
    src = torch.from_numpy(im).cuda()
    ctx = mseg.Context(0)
    ctx.set_stream(torch.cuda.current_stream().cuda_stream)
    full = torch.empty_like(src)
    dev.meanshift(ctx, src.data_ptr(), 3 * w, full.data_ptr(), 3 * w, w, h, sp, sr, ml)
    lab_full = torch.empty((h, w), dtype=torch.int32, device="cuda")
    dev.label_regions(ctx, full.data_ptr(), 3 * w, lab_full.data_ptr(), 4 * w, w, h, 2)
    ctx.synchronize()
    strips = sh.plan_strips(h, n, ml)
    ctxs = [mseg.Context(0) for _ in strips]                  # one context per simulated rank (keeps the rank tables)
    for c in ctxs:
        c.set_stream(torch.cuda.current_stream().cuda_stream)
    labs, cnt = [], torch.zeros(len(strips), dtype=torch.int32, device="cuda")
    for k, (r0, r1) in enumerate(strips):
        rows = full[r0:r1].contiguous()
        l = torch.empty((r1 - r0, w), dtype=torch.int32, device="cuda")
        dev.label_strip(ctxs[k], rows.data_ptr(), 3 * w, l.data_ptr(), 4 * w, w, r1 - r0, r0, w, 2)
        dev.strip_rank(ctxs[k], l.data_ptr(), 4 * w, w, r1 - r0, r0, w, cnt[k:].data_ptr())
        labs.append(l)
    quads_all = []
    for k in range(1, len(strips)):
        (u0, u1), (r0, r1) = strips[k - 1], strips[k]
        up_lab = labs[k - 1][-1].contiguous()
        up_rank1 = torch.zeros(w, dtype=torch.int32, device="cuda")           # what the rank above sends with its last row
        dev.strip_query_dense(ctxs[k - 1], up_lab.data_ptr(), w, w, u1 - u0, u0, w, 0, up_rank1.data_ptr())
        quads = torch.zeros((w, 4), dtype=torch.int32, device="cuda")
        qn = torch.zeros(1, dtype=torch.int32, device="cuda")
        dev.seam_quads(ctxs[k], full[r0 - 1].data_ptr(), up_lab.data_ptr(), up_rank1.data_ptr(), full[r0].data_ptr(),
                       labs[k][0].data_ptr(), w, 2, r1 - r0, r0, w, quads.data_ptr(), qn.data_ptr())
        torch.cuda.synchronize()
        quads_all.append(quads[:int(qn.item())].cpu().numpy().copy())
    frm, dense, offsets, frm_lo, total = sh.resolve_dense(np.concatenate(quads_all), cnt.cpu().numpy(), strips, w)
    d_frm = torch.from_numpy(np.ascontiguousarray(frm)).cuda() if len(frm) else torch.zeros(1, dtype=torch.int32, device="cuda")
    d_dense = torch.from_numpy(np.ascontiguousarray(dense)).cuda() if len(frm) else torch.zeros(1, dtype=torch.int32, device="cuda")
    for k, (r0, r1) in enumerate(strips):
        dev.strip_finalize_dense(ctxs[k], labs[k].data_ptr(), 4 * w, w, r1 - r0, r0, w, int(offsets[k]), d_frm.data_ptr(),
                                 d_dense.data_ptr(), len(frm), int(frm_lo[k]))
    torch.cuda.synchronize()
    got = torch.cat(labs)
    assert total == int(lab_full.max().item())
    assert torch.equal(got, lab_full), int((got != lab_full).sum().item())
    for c in ctxs:
        c.close()
    ctx.close()


@pytest.mark.parametrize("w,h,n,sp,sr,ml,min_size,cd", [(600, 518, 3, 10, 10, 1, 30, 8), (333, 400, 4, 6, 15, 2, 20, 0),
                                                        (420, 300, 5, 10, 10, 1, 0, 12), (512, 96, 6, 4, 30, 0, 25, 6),
                                                        (1024, 700, 8, 10, 10, 1, 50, 10)])
def test_strips_device_resolve_and_sharded_merge(w, h, n, sp, sr, ml, min_size, cd):
    """Round 2: the whole strip pipeline without host arithmetic.  msg_shard_plan_make; per strip label + rank + seam quads; the
    "all-gathered" payload buffer is resolved ON THE DEVICE (msg_strip_resolve_dense_dev) and consumed by
    msg_strip_finalize_tables_dev -> labels identical to the unsharded call; then the strip-sharded merge (per-strip tables,
    summed = the all-reduce, pair lists concatenated = the all-gather, identical rounds per strip) -> identical to
    msg_merge_regions_dev on the whole image and to the oracle."""
    im = orc.synth_bgr(w, h, 29)
    src = torch.from_numpy(im).cuda()
    ctx = mseg.Context(0)
    ctx.set_stream(torch.cuda.current_stream().cuda_stream)
    full = torch.empty_like(src)
    dev.meanshift(ctx, src.data_ptr(), 3 * w, full.data_ptr(), 3 * w, w, h, sp, sr, ml)
    lab_full = torch.empty((h, w), dtype=torch.int32, device="cuda")
    nfull = torch.zeros(1, dtype=torch.int32, device="cuda")
    dev.label_regions(ctx, full.data_ptr(), 3 * w, lab_full.data_ptr(), 4 * w, w, h, 2, nfull.data_ptr())
    ctx.synchronize()
    halo, strips, halos = dev.shard_plan(w, h, n, sp, ml)
    ctxs = [mseg.Context(0) for _ in strips]
    for c in ctxs:
        c.set_stream(torch.cuda.current_stream().cuda_stream)
    labs = []
    gathered = torch.zeros((n, w + 1, 4), dtype=torch.int32, device="cuda")       # what ONE all-gather delivers to every rank
    for k, (r0, r1) in enumerate(strips):
        rows = full[r0:r1].contiguous()
        l = torch.empty((r1 - r0, w), dtype=torch.int32, device="cuda")
        dev.label_strip(ctxs[k], rows.data_ptr(), 3 * w, l.data_ptr(), 4 * w, w, r1 - r0, r0, w, 2)
        dev.strip_rank(ctxs[k], l.data_ptr(), 4 * w, w, r1 - r0, r0, w, gathered[k, 0, 1:].data_ptr())
        labs.append(l)
    for k in range(1, n):
        (u0, u1), (r0, r1) = strips[k - 1], strips[k]
        up_lab = labs[k - 1][-1].contiguous()
        up_rank1 = torch.zeros(w, dtype=torch.int32, device="cuda")
        dev.strip_query_dense(ctxs[k - 1], up_lab.data_ptr(), w, w, u1 - u0, u0, w, 0, up_rank1.data_ptr())
        dev.seam_quads(ctxs[k], full[r0 - 1].data_ptr(), up_lab.data_ptr(), up_rank1.data_ptr(), full[r0].data_ptr(),
                       labs[k][0].data_ptr(), w, 2, r1 - r0, r0, w, gathered[k, 1:].data_ptr(), gathered[k, 0, 0:].data_ptr())
    tables_ints = mseg.lib.SHARD_TABLE_HEADER + 2 * n * w
    tables = torch.zeros(tables_ints, dtype=torch.int32, device="cuda")
    for k, (r0, r1) in enumerate(strips):            # every rank resolves the same payload and rewrites its strip
        dev.strip_resolve_dense(ctxs[k], gathered.data_ptr(), n, w, [s[0] for s in strips], tables.data_ptr(), tables_ints)
        dev.strip_finalize_tables(ctxs[k], labs[k].data_ptr(), 4 * w, w, r1 - r0, r0, w, k, n, tables.data_ptr())
    torch.cuda.synchronize()
    got = torch.cat(labs)
    total = int(tables[1].item())
    assert total == int(nfull.item()) == int(lab_full.max().item())
    assert torch.equal(got, lab_full), int((got != lab_full).sum().item())
    # the host tables of round 1 agree with the device tables
    host = gathered.cpu().numpy()
    quads = np.concatenate([host[r, 1:1 + int(host[r, 0, 0])] for r in range(n)], axis=0)
    frm, dense, offsets, frm_lo, total_h = sh.resolve_dense(quads, host[:, 0, 1], strips, w)
    t = tables.cpu().numpy()
    assert t[0] == len(frm) and t[1] == total_h
    assert np.array_equal(t[2:2 + n], offsets) and np.array_equal(t[2 + 64:2 + 64 + n], frm_lo)
    hdr = mseg.lib.SHARD_TABLE_HEADER
    assert np.array_equal(t[hdr:hdr + len(frm)], frm) and np.array_equal(t[hdr + n * w:hdr + n * w + len(frm)], dense)
    # ---- sharded merge
    want = lab_full.clone()
    nwant = torch.zeros(1, dtype=torch.int32, device="cuda")
    dev.merge_regions(ctx, full.data_ptr(), 3 * w, want.data_ptr(), 4 * w, w, h, min_size, cd, nwant.data_ptr())
    area = torch.zeros((n, total + 1), dtype=torch.int32, device="cuda")
    sums = torch.zeros((n, 3 * (total + 1)), dtype=torch.int64, device="cuda")
    cap = 2 * w * max(r1 - r0 for r0, r1 in strips)
    pairs = torch.zeros((n, cap, 2), dtype=torch.int32, device="cuda")
    npairs = torch.zeros(n, dtype=torch.int32, device="cuda")
    for k, (r0, r1) in enumerate(strips):
        up = labs[k - 1][-1].contiguous() if k else None
        dev.strip_merge_stats(ctxs[k], full[r0:r1].contiguous().data_ptr(), 3 * w, labs[k].data_ptr(), 4 * w, w, r1 - r0,
                              up.data_ptr() if up is not None else 0, total, area[k].data_ptr(), sums[k].data_ptr(),
                              pairs[k].data_ptr(), cap, npairs[k:].data_ptr())
    torch.cuda.synchronize()
    assert int(area.sum().item()) == w * h                                           # every pixel counted exactly once
    area_all, sums_all = area.sum(dim=0, dtype=torch.int32), sums.sum(dim=0)          # = all-reduce(sum)
    cnts = npairs.cpu().tolist()
    assert max(cnts) <= cap
    all_pairs = torch.cat([pairs[k, :cnts[k]] for k in range(n)]).contiguous()        # = all-gather
    nout = torch.zeros(n, dtype=torch.int32, device="cuda")
    for k, (r0, r1) in enumerate(strips):
        a_k, s_k = area_all.clone(), sums_all.clone()                                 # every rank owns its copy of the tables
        dev.strip_merge_finish(ctxs[k], labs[k].data_ptr(), 4 * w, w, r1 - r0, w * h, total, a_k.data_ptr(), s_k.data_ptr(),
                               all_pairs.data_ptr(), all_pairs.shape[0], min_size, cd, nout[k:].data_ptr())
    torch.cuda.synchronize()
    got = torch.cat(labs)
    assert len(set(nout.cpu().tolist())) == 1 and int(nout[0].item()) == int(nwant.item())
    assert torch.equal(got, want), int((got != want).sum().item())
    f = full.cpu().numpy()
    n0, l0 = orc.label_regions(f, 2)
    n1, l1 = orc.merge_regions(f, l0, min_size, cd)
    assert n1 == int(nwant.item()) and np.array_equal(got.cpu().numpy(), l1)
    for c in ctxs:
        c.close()
    ctx.close()


def test_synth_rows_matches_full():
    ctx = mseg.Context(0)
    ctx.set_stream(torch.cuda.current_stream().cuda_stream)
    w, h = 300, 200
    a = torch.empty((h, w, 3), dtype=torch.uint8, device="cuda")
    dev.synth(ctx, a.data_ptr(), 3 * w, w, h, 5)
    b = torch.empty((60, w, 3), dtype=torch.uint8, device="cuda")
    dev.synth_rows(ctx, b.data_ptr(), 3 * w, w, h, 70, 60, 5)
    ctx.synchronize()
    assert torch.equal(a[70:130], b)
    assert np.array_equal(a.cpu().numpy(), orc.synth_bgr(w, h, 5))
    ctx.close()


def _smooth_field(w, h, seed, k):
    """Low-frequency colour field (box-filtered noise, stretched to 0..255): mean-shift windows drift far on it."""
    rng = np.random.default_rng(seed)
    f = rng.random((h + 2 * k, w + 2 * k, 3))
    for axis in (0, 1):
        c = np.cumsum(f, axis=axis)
        n = f.shape[axis] - 2 * k
        f = np.take(c, np.arange(2 * k, 2 * k + n), axis=axis) - np.take(c, np.arange(0, n), axis=axis)
    f = (f - f.min()) / (f.max() - f.min())
    return np.ascontiguousarray((f * 255).astype(np.uint8))


@pytest.mark.parametrize("w,h,n,sp,sr,ml,term", [(200, 1200, 3, 10, 30, 1, (3, 5, 1.0)), (160, 1400, 4, 10, 40, 1, (1, 9, 0.0)),
                                                (180, 900, 2, 7.5, 25, 2, (3, 5, 1.0)), (150, 1000, 3, 12, 50, 0, (1, 7, 0.0)),
                                                (140, 1100, 3, 4, 60, 3, (3, 6, 1.0))])
def test_halo_bound_on_far_drifting_windows(w, h, n, sp, sr, ml, term):
    """The halo of msg_meanshift_halo_rows is the derived dependency bound max(own windows, via the level above), not a chain
    of the two.  Tall strips (far more rows than the halo), smooth fields and large sr make the windows travel
    as far as they can; every strip filtered from its rows + halo only must equal the unsharded call bit for bit -- and the
    drift statistics show that the windows did travel (the test would be vacuous otherwise)."""
    ctx = mseg.Context(0)
    ctx.set_stream(torch.cuda.current_stream().cuda_stream)
    im = _smooth_field(w, h, 100 + ml, 9)
    src = torch.from_numpy(im).cuda()
    full = torch.empty_like(src)
    dev.meanshift(ctx, src.data_ptr(), 3 * w, full.data_ptr(), 3 * w, w, h, sp, sr, ml, term)
    ctx.synchronize()
    halo = dev.halo_rows(sp, ml, term)
    mc = term[1] if term[0] & 1 else 5
    own = mc * int(np.ceil(max(sp, 1.0)))
    assert own <= halo <= 2 * own + 16 * (1 << ml), (halo, own)            # not the old chained bound
    strips = sh.plan_strips(h, n, ml)
    assert min(r1 - r0 for r0, r1 in strips) > 2 * halo
    for (r0, r1) in strips:
        h0, h1 = sh.halo_range(r0, r1, h, halo, ml)
        rows = src[h0:h1].contiguous()
        out = torch.empty((r1 - r0, w, 3), dtype=torch.uint8, device="cuda")
        dev.meanshift_strip(ctx, rows.data_ptr(), 3 * w, h0, h1, out.data_ptr(), 3 * w, w, h, r0, r1, sp, sr, ml, term)
        bad = (out != full[r0:r1]).any(dim=2)
        assert not bad.any(), "strip %d..%d differs at %d pixels, rows %s" % (
            r0, r1, int(bad.sum()), (torch.nonzero(bad.any(dim=1)).flatten()[:8] + r0).tolist())
    # the unsharded GPU result is the oracle's on this image too
    assert np.array_equal(full.cpu().numpy(), orc.meanshift_filter(im, sp, sr, ml, term))
    # the windows did move: a large share of the pixels changed colour by more than the noise level
    moved = (np.abs(full.cpu().numpy().astype(int) - im.astype(int)).max(axis=2) > 3).mean()
    assert moved > 0.05, moved
