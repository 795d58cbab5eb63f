#!/usr/bin/env python
"""Multi-GPU check of BASELINE.json config 5 against the CPU ORACLE, at any size incl. 16384 x 16384:

  python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 tests/shard_verify_oracle.py --size 16384

Runs the strip pipeline of tools/shard_large_image.py (the product path: C ABI + NCCL) and then, as the checker:
  (1) every rank: the strip-wise oracle (orc_meanshift_filter_roi: global coordinates, dependency-cone halo) on the first and last
      VERIFY_ROWS rows of its strip -- i.e. both sides of every seam -- against the GPU's filtered rows;
  (2) rank 0: the oracle's union-find labelling (and the oracle's merge) on the WHOLE gathered filtered image against all labels.
Lives under tests/ because only tests may use the oracle; not collected by pytest (no test_ prefix: it needs torchrun)."""
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tools"))
import numpy as np  # noqa: E402
import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402

import msegment_b200 as mseg  # noqa: E402
import shard_large_image as tool  # noqa: E402
from oracle import oracle as orc  # noqa: E402

dev = mseg.device
VERIFY_ROWS = 48


def hook(st):
    args, ctx, rank, w, h, ml, halo = st["args"], st["ctx"], st["rank"], st["w"], st["h"], st["ml"], st["halo"]
    r0, r1, filt, lab = st["r0"], st["r1"], st["filt"], st["lab"]
    t0 = time.perf_counter()
    bad_rows = checked = 0
    for (a, b) in ((r0, min(r1, r0 + VERIFY_ROWS)), (max(r0, r1 - VERIFY_ROWS), r1)):
        c0, c1 = max(0, a - halo), min(h, b + halo)
        c0 -= c0 % (1 << ml)
        crop = torch.empty((c1 - c0, w, 3), dtype=torch.uint8, device="cuda")
        dev.synth_rows(ctx, crop.data_ptr(), 3 * w, w, h, c0, c1 - c0, args.seed)
        ctx.synchronize()
        want = orc.meanshift_filter_roi(crop.cpu().numpy(), 0, c0, w, h, args.sp, args.sr, ml)[a - c0:b - c0]
        got = filt[a - r0:b - r0].cpu().numpy()
        bad_rows += int((got != want).any(axis=(1, 2)).sum())
        checked += b - a
    flt = torch.tensor([bad_rows, checked], device="cuda", dtype=torch.int64)
    dist.all_reduce(flt)
    out = {"seam_rows_checked_vs_oracle": int(flt[1].item()), "seam_rows_differing": int(flt[0].item()),
           "rows_per_strip_end": VERIFY_ROWS, "filter_oracle_seconds": round(time.perf_counter() - t0, 1)}
    if st["equal_strips"]:
        strips = st["strips"]
        gf = [torch.empty_like(filt) for _ in strips] if rank == 0 else None
        dist.gather(filt, gf, dst=0)
        gu = [torch.empty_like(lab) for _ in strips] if rank == 0 else None
        dist.gather(st["lab_unmerged"], gu, dst=0)
        gl = [torch.empty_like(lab) for _ in strips] if rank == 0 else None
        dist.gather(lab, gl, dst=0)
        if rank == 0:
            t0 = time.perf_counter()
            f_host = torch.cat(gf).cpu().numpy()
            del gf
            n0, l0 = orc.label_regions(f_host, args.lo)
            got_u = torch.cat(gu).cpu().numpy()
            out["labels_equal_oracle_whole_image"] = bool(n0 == st["n_total"] and np.array_equal(got_u, l0))
            del got_u, gu
            if st["do_merge"]:
                n1, l1 = orc.merge_regions(f_host, l0, args.min_size, args.color_dist)
                got_l = torch.cat(gl).cpu().numpy()
                out["merged_labels_equal_oracle_whole_image"] = bool(n1 == st["n_after"] and np.array_equal(got_l, l1))
            out["label_oracle_seconds"] = round(time.perf_counter() - t0, 1)
    return out


if __name__ == "__main__":
    tool.run(tool.parse_args(), oracle_hook=hook)
