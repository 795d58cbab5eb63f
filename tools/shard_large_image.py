#!/usr/bin/env python
"""BASELINE.json config 5: one very large synthetic image, row strips over the GPUs of one box.

  python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 tools/shard_large_image.py \
        --size 16384 [--min-size 50 --color-dist 10] [--verify]

This script is only the HOST of the strip pipeline: it owns the NCCL calls (torch.distributed) and nothing else.  Everything it
calls between two collectives is a C-ABI entry point of include/msegment.h, in this order:

  msg_shard_plan_make                                   strips + halo ranges (pure host C)
  [P2P]  halo rows from the neighbours                  input rows only
  msg_meanshift_filter_strip_dev, msg_label_strip_dev, msg_strip_rank_dev, msg_strip_query_dense_dev
  [P2P]  last filtered row, its labels and root ranks to the rank below
  msg_seam_quads_dev
  [ALL-GATHER]  (quad count, root count, quads) of every strip -- the ONE collective of the label stage
  msg_strip_resolve_dense_dev, msg_strip_finalize_tables_dev      seam union-find + dense numbering ON THE DEVICE
  -- labels are now identical to the unsharded call's; with --min-size / --color-dist the merge follows: --
  [4-byte read]  total number of regions (sizes the tables)
  [P2P]  last row of dense labels to the rank below
  msg_strip_merge_stats_dev
  [ALL-REDUCE] area + colour sums (one buffer)   [ALL-GATHER] adjacent-pair lists
  msg_strip_merge_finish_dev                            identical rounds on every rank, strip rewritten

--verify (sizes that fit one GPU) gathers the result on rank 0 and compares it bit for bit with the unsharded single-GPU
calls.  The check against the CPU ORACLE at any size lives in tests/shard_verify_oracle.py (only tests/ may touch the oracle):
it calls run() below with a hook, every rank then runs the strip-wise oracle (global coordinates, halo) on the first and last
rows of its strip -- both sides of all seams -- and rank 0 the oracle's union-find labelling (and merge) on the whole gathered
filtered image.  Prints one JSON line on rank 0.
"""
import argparse
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np  # noqa: E402
import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402

import msegment_b200 as mseg  # noqa: E402

dev = mseg.device


def parse_args(argv=None):
    ap = argparse.ArgumentParser()
    ap.add_argument("--size", type=int, default=16384)
    ap.add_argument("--width", type=int, default=0)
    ap.add_argument("--seed", type=int, default=5)
    ap.add_argument("--sp", type=float, default=10.0)
    ap.add_argument("--sr", type=float, default=10.0)
    ap.add_argument("--max-level", type=int, default=1)
    ap.add_argument("--lo", type=int, default=2)
    ap.add_argument("--min-size", type=int, default=50)
    ap.add_argument("--color-dist", type=int, default=10)
    ap.add_argument("--verify", action="store_true")
    return ap.parse_args(argv)


def run(args, oracle_hook=None):
    """oracle_hook(state) -> dict, called on every rank after the timed passes (tests/shard_verify_oracle.py)."""
    rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    h = args.size
    w = args.width or args.size
    ml = args.max_level
    do_merge = args.min_size > 0 or args.color_dist > 0
    ctx = mseg.Context(local)
    ctx.set_stream(torch.cuda.current_stream().cuda_stream)
    halo, strips, halos = dev.shard_plan(w, h, world, args.sp, ml)
    r0, r1 = strips[rank]
    h0, h1 = halos[rank]
    rows = r1 - r0

    # own rows, generated in place inside the halo buffer
    buf = torch.empty((h1 - h0, w, 3), dtype=torch.uint8, device="cuda")
    dev.synth_rows(ctx, buf[r0 - h0:].data_ptr(), 3 * w, w, h, r0, rows, args.seed)
    # warm up NCCL (communicator + P2P channels to both neighbours are created lazily on first use)
    warm = torch.zeros(1024, device="cuda")
    dist.all_reduce(warm)
    wops = []
    for peer in (rank - 1, rank + 1):
        if 0 <= peer < world:
            wops += [dist.P2POp(dist.isend, warm, peer), dist.P2POp(dist.irecv, torch.empty_like(warm), peer)]
    if wops:
        for req in dist.batch_isend_irecv(wops):
            req.wait()
    tables_ints = mseg.lib.SHARD_TABLE_HEADER + 2 * world * w
    row0s = [s[0] for s in strips]

    def p2p(ops):
        if ops:
            for req in dist.batch_isend_irecv(ops):
                req.wait()

    def process():
        torch.cuda.synchronize()
        dist.barrier()
        t_start = time.perf_counter()
        ev = [torch.cuda.Event(enable_timing=True) for _ in range(6)]
        ev[0].record()

        # ---- halo exchange (input rows only): rows [h0,r0) come from ranks above, rows [r1,h1) from ranks below
        ops = []
        for peer, (p0, p1) in enumerate(strips):
            if peer == rank:
                continue
            ph0, ph1 = halos[peer]
            for (a, b) in ((max(ph0, r0), min(p0, r1)), (max(p1, r0), min(ph1, r1))):      # what the peer needs from me
                if a < b:
                    ops.append(dist.P2POp(dist.isend, buf[a - h0:b - h0], peer))
            for (a, b) in ((max(h0, p0), min(r0, p1)), (max(r1, p0), min(h1, p1))):        # what I need from the peer
                if a < b:
                    ops.append(dist.P2POp(dist.irecv, buf[a - h0:b - h0], peer))
        halo_bytes = sum(op.tensor.numel() for op in ops if op.op == dist.irecv)
        p2p(ops)
        ev[1].record()

        # ---- filter + label the strip
        filt = torch.empty((rows, w, 3), dtype=torch.uint8, device="cuda")
        dev.meanshift_strip(ctx, buf.data_ptr(), 3 * w, h0, h1, filt.data_ptr(), 3 * w, w, h, r0, r1, args.sp, args.sr, ml)
        ev[2].record()
        lab = torch.empty((rows, w), dtype=torch.int32, device="cuda")
        dev.label_strip(ctx, filt.data_ptr(), 3 * w, lab.data_ptr(), 4 * w, w, rows, r0, w, args.lo)
        ev[3].record()

        # ---- seams + dense global numbering: one boundary-row transfer, ONE all-gather, tables resolved on the device
        payload = torch.zeros((w + 1, 4), dtype=torch.int32, device="cuda")      # row 0: (quad count, root count, 0, 0)
        dev.strip_rank(ctx, lab.data_ptr(), 4 * w, w, rows, r0, w, payload[0, 1:].data_ptr())
        last = torch.zeros((2, w), dtype=torch.int32, device="cuda")             # my last row: labels, rank + 1 of their roots
        last[0].copy_(lab[-1])
        if rank + 1 < world:
            dev.strip_query_dense(ctx, last[0].data_ptr(), w, w, rows, r0, w, 0, last[1].data_ptr())
        up_bgr = torch.empty((w, 3), dtype=torch.uint8, device="cuda")
        up_lr = torch.empty((2, w), dtype=torch.int32, device="cuda")
        ops = []
        if rank + 1 < world:
            ops += [dist.P2POp(dist.isend, filt[-1], rank + 1), dist.P2POp(dist.isend, last, rank + 1)]
        if rank > 0:
            ops += [dist.P2POp(dist.irecv, up_bgr, rank - 1), dist.P2POp(dist.irecv, up_lr, rank - 1)]
        p2p(ops)
        if rank > 0:
            dev.seam_quads(ctx, up_bgr.data_ptr(), up_lr[0].data_ptr(), up_lr[1].data_ptr(), filt[0].data_ptr(), lab[0].data_ptr(),
                           w, args.lo, rows, r0, w, payload[1:].data_ptr(), payload[0, 0:].data_ptr())
        gathered = torch.empty((world, w + 1, 4), dtype=torch.int32, device="cuda")
        dist.all_gather_into_tensor(gathered, payload)
        tables = torch.empty(tables_ints, dtype=torch.int32, device="cuda")
        dev.strip_resolve_dense(ctx, gathered.data_ptr(), world, w, row0s, tables.data_ptr(), tables_ints)
        dev.strip_finalize_tables(ctx, lab.data_ptr(), 4 * w, w, rows, r0, w, rank, world, tables.data_ptr())
        ev[4].record()

        # ---- region merge across the strips
        n_total = n_after = None
        lab_unmerged = lab.clone() if (args.verify or oracle_hook is not None) else None
        if do_merge:
            n_total = int(tables[1].item())                                      # sizes the tables: the one host read
            ops = []
            up_dense = torch.empty((w,), dtype=torch.int32, device="cuda")
            if rank + 1 < world:
                ops.append(dist.P2POp(dist.isend, lab[-1], rank + 1))
            if rank > 0:
                ops.append(dist.P2POp(dist.irecv, up_dense, rank - 1))
            p2p(ops)
            # one buffer, ONE all-reduce: 3 x int64 colour sums per label, then the int32 areas packed two per int64 (the areas
            # of a label summed over the ranks stay below 2^32 -- they are pixel counts -- so no carry crosses the halves)
            nl = n_total + 1
            stats = torch.zeros(3 * nl + (nl + 1) // 2, dtype=torch.int64, device="cuda")
            sums = stats[:3 * nl]
            area = stats[3 * nl:].view(torch.int32)[:nl]
            cap = 2 * w * rows
            pairs = torch.empty((cap, 2), dtype=torch.int32, device="cuda")
            npairs = torch.zeros(1, dtype=torch.int32, device="cuda")
            dev.strip_merge_stats(ctx, filt.data_ptr(), 3 * w, lab.data_ptr(), 4 * w, w, rows, up_dense.data_ptr() if rank > 0 else 0,
                                  n_total, area.data_ptr(), sums.data_ptr(), pairs.data_ptr(), cap, npairs.data_ptr())
            dist.all_reduce(stats)
            counts = torch.empty(world, dtype=torch.int32, device="cuda")
            dist.all_gather_into_tensor(counts, npairs)
            cl = counts.cpu().tolist()
            if max(cl) > cap:
                raise RuntimeError("adjacent-pair list overflow")
            mx = max(1, max(cl))
            allp = torch.empty((world, mx, 2), dtype=torch.int32, device="cuda")
            dist.all_gather_into_tensor(allp, pairs[:mx].contiguous())
            all_pairs = torch.cat([allp[r, :cl[r]] for r in range(world)]).contiguous()
            nout = torch.zeros(1, dtype=torch.int32, device="cuda")
            dev.strip_merge_finish(ctx, lab.data_ptr(), 4 * w, w, rows, w * h, n_total, area.data_ptr(), sums.data_ptr(),
                                   all_pairs.data_ptr(), all_pairs.shape[0], args.min_size, args.color_dist, nout.data_ptr())
            n_after = nout
        ev[5].record()
        torch.cuda.synchronize()
        dist.barrier()
        wall = time.perf_counter() - t_start
        ms = [ev[i].elapsed_time(ev[i + 1]) for i in range(5)]
        tmax = torch.tensor(ms + [wall * 1e3], device="cuda")
        dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
        if n_total is None:
            n_total = int(tables[1].item())
        quads = int(gathered[:, 0, 0].sum().item())
        return dict(filt=filt, lab=lab, lab_unmerged=lab_unmerged, halo_bytes=halo_bytes, quads=quads, n_total=n_total,
                    n_after=int(n_after.item()) if n_after is not None else None, tmax=tmax)

    process()                                    # warm-up pass: workspace allocation, first-launch overheads
    res = process()
    filt, lab = res["filt"], res["lab"]

    equal_strips = len({b - a for a, b in strips}) == 1
    ok = None
    if args.verify:
        if not equal_strips:
            ok = {"skipped": "unequal strips"}
        else:
            parts_f = [torch.empty_like(filt) for _ in strips]
            parts_l = [torch.empty_like(lab) for _ in strips]
            parts_u = [torch.empty_like(lab) for _ in strips]
            dist.all_gather(parts_f, filt)
            dist.all_gather(parts_l, lab)
            dist.all_gather(parts_u, res["lab_unmerged"])
            if rank == 0:
                full_f, full_l, full_u = torch.cat(parts_f), torch.cat(parts_l), torch.cat(parts_u)
                src = torch.empty((h, w, 3), dtype=torch.uint8, device="cuda")
                dev.synth(ctx, src.data_ptr(), 3 * w, w, h, args.seed)
                ref_f = torch.empty_like(src)
                dev.meanshift(ctx, src.data_ptr(), 3 * w, ref_f.data_ptr(), 3 * w, w, h, args.sp, args.sr, ml)
                ref_l = torch.empty((h, w), dtype=torch.int32, device="cuda")
                dev.label_regions(ctx, ref_f.data_ptr(), 3 * w, ref_l.data_ptr(), 4 * w, w, h, args.lo)
                ctx.synchronize()
                ok = {"filtered_bit_identical": bool(torch.equal(full_f, ref_f)), "labels_bit_identical": bool(torch.equal(full_u, ref_l))}
                if do_merge:
                    nref = torch.zeros(1, dtype=torch.int32, device="cuda")
                    dev.merge_regions(ctx, ref_f.data_ptr(), 3 * w, ref_l.data_ptr(), 4 * w, w, h, args.min_size, args.color_dist, nref.data_ptr())
                    ctx.synchronize()
                    ok["merged_labels_bit_identical"] = bool(torch.equal(full_l, ref_l)) and int(nref.item()) == res["n_after"]
    oracle = None
    if oracle_hook is not None:
        oracle = oracle_hook(dict(args=args, ctx=ctx, rank=rank, world=world, w=w, h=h, ml=ml, halo=halo, strips=strips, r0=r0, r1=r1,
                                  filt=filt, lab=lab, lab_unmerged=res["lab_unmerged"], n_total=res["n_total"], n_after=res["n_after"],
                                  do_merge=do_merge, equal_strips=equal_strips))
    if rank == 0:
        t = [float(x) for x in res["tmax"].tolist()]
        print(json.dumps({"config": "strip-sharded %dx%d over %d GPUs, sp=%g sr=%g maxLevel=%d lo=%d minSize=%d colorDist=%d"
                                    % (w, h, world, args.sp, args.sr, ml, args.lo, args.min_size, args.color_dist),
                          "n_gpus": world, "halo_rows": halo, "halo_bytes_received_rank0": int(res["halo_bytes"]),
                          "seam_quads_total": res["quads"], "regions": res["n_total"], "regions_after_merge": res["n_after"],
                          "ms_max_over_ranks": {"halo_exchange": round(t[0], 3), "meanshift": round(t[1], 3), "label": round(t[2], 3),
                                                "seams": round(t[3], 3), "merge": round(t[4], 3), "wall": round(t[5], 3)},
                          "mpix_per_s": round(w * h / 1e6 / (t[5] / 1e3), 1), "host_reads": "4 bytes (region count) + pair counts, merge only",
                          "verify": ok, "verify_oracle": oracle}))
    dist.barrier()
    dist.destroy_process_group()
    ctx.close()


if __name__ == "__main__":
    run(parse_args())
