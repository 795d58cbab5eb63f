#!/usr/bin/env python
"""Label + merge + render of one filtered frame, device-resident, for ncu (launch list / --set full of the HBM-bound stages).
Usage: python tools/profile_stages.py W H [reps]   -- numbers printed by a run under ncu are not bench values."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

import msegment_b200 as mseg  # noqa: E402


def main():
    w, h = int(sys.argv[1]), int(sys.argv[2])
    reps = int(sys.argv[3]) if len(sys.argv) > 3 else 2
    dev = mseg.device
    ctx = mseg.Context(0)
    ctx.set_stream(torch.cuda.current_stream().cuda_stream)
    ctx.set_option("labels_canonical", 1)     # the labels come straight from msg_label_regions_dev: no validation passes
    src = torch.empty((h, w, 3), dtype=torch.uint8, device="cuda")
    filt = torch.empty_like(src)
    ren = torch.empty_like(src)
    lab = torch.empty((h, w), dtype=torch.int32, device="cuda")
    cnt = torch.zeros(4, dtype=torch.int32, device="cuda")
    dev.synth(ctx, src.data_ptr(), 3 * w, w, h, 2)
    dev.meanshift(ctx, src.data_ptr(), 3 * w, filt.data_ptr(), 3 * w, w, h, 10, 10)
    torch.cuda.synchronize()
    print("MARK stages begin", ctx.stats()["kernel_launches"])
    for _ in range(reps):
        dev.label_regions(ctx, filt.data_ptr(), 3 * w, lab.data_ptr(), 4 * w, w, h, 2, cnt.data_ptr())
        dev.merge_regions(ctx, filt.data_ptr(), 3 * w, lab.data_ptr(), 4 * w, w, h, 50, 10, cnt.data_ptr())
        dev.render_labels(ctx, lab.data_ptr(), 4 * w, ren.data_ptr(), 3 * w, w, h, int(cnt[0].item()))
    torch.cuda.synchronize()
    print("done", int(cnt[0].item()), ctx.stats()["kernel_launches"])


if __name__ == "__main__":
    main()
