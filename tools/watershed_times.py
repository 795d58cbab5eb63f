#!/usr/bin/env python
"""Times the exact watershed (msg_watershed_batch_dev: one warp per image, SURVEY 8(f1), PictureService.java:908-911) on the
markers the colour-method generator produces for synthetic frames: one image alone (the honest per-image figure: the flood is
one dependent chain) and batches of independent images (where the GPU's throughput comes from), device-resident, CUDA events
on the launching stream.  cv2.watershed on one core is timed beside it when cv2 is importable (build container only).
Usage: python tools/watershed_times.py [W H] [max_batch]"""
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np  # noqa: E402
import torch  # noqa: E402

import msegment_b200 as mseg  # noqa: E402

dev = mseg.device


def main():
    w, h = (int(sys.argv[1]), int(sys.argv[2])) if len(sys.argv) > 2 else (1920, 1080)
    max_batch = int(sys.argv[3]) if len(sys.argv) > 3 else 296
    torch.cuda.set_device(0)
    out = {"size": "%dx%d" % (w, h), "what": "Imgproc.watershed on the colour-method markers of synthetic frames (seeds 100..)",
           "timing": "CUDA events on the launching stream, device-resident, best of 3"}
    with mseg.Context(0) as ctx:
        ctx.set_stream(torch.cuda.current_stream().cuda_stream)
        gi = mseg.GpuImgproc(ctx)
        nseed = 4
        ims, mks = [], []
        for s in range(nseed):
            im = mseg.synth_bgr(w, h, 100 + s)
            n, markers = gi.colorSeeds(im)
            ims.append(im)
            mks.append(markers)
        out["contours_frame0"] = int(mks[0].max())
        pops = torch.zeros(1, dtype=torch.int64, device="cuda")
        rows = []
        b = 1
        batches = []
        while b <= max_batch:
            batches.append(b)
            b *= 4
        if batches[-1] != max_batch:
            batches.append(max_batch)
        for count in batches:
            d_im = torch.empty((count, h, w, 3), dtype=torch.uint8, device="cuda")
            d_mk0 = torch.empty((count, h, w), dtype=torch.int32, device="cuda")
            for k in range(count):
                d_im[k].copy_(torch.from_numpy(ims[k % nseed]))
                d_mk0[k].copy_(torch.from_numpy(mks[k % nseed]))
            d_mk = torch.empty_like(d_mk0)
            best = 1e9
            for _ in range(3):
                d_mk.copy_(d_mk0)
                pops.zero_()
                a, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                a.record()
                dev.watershed_batch(ctx, d_im.data_ptr(), 3 * w, 3 * w * h, d_mk.data_ptr(), 4 * w, 4 * w * h, w, h, count, pops.data_ptr())
                e.record()
                torch.cuda.synchronize()
                best = min(best, a.elapsed_time(e))
            npops = int(pops.item())
            rows.append({"images": count, "ms": round(best, 3), "mpix_per_s": round(count * w * h / best / 1e3, 1),
                         "pops": npops, "ns_per_pop_per_image_chain": round(best * 1e6 / (npops / count), 1)})
            print(json.dumps(rows[-1]), flush=True)
            if count == 1:
                first = d_mk[0].cpu().numpy()
            del d_im, d_mk0, d_mk
            torch.cuda.empty_cache()
        out["gpu"] = rows
    try:
        import cv2
        cv2.setNumThreads(1)
        m = mks[0].copy()
        t0 = time.perf_counter()
        cv2.watershed(ims[0], m)
        dt = time.perf_counter() - t0
        out["cv2_one_core"] = {"ms": round(dt * 1e3, 2), "mpix_per_s": round(w * h / dt / 1e6, 2),
                               "equal_gpu": bool(np.array_equal(m, first))}
    except ImportError:
        out["cv2_one_core"] = None      # GPU box: no cv2 (tests/test_gpu_watershed.py holds the parity checks against the oracle)
    print(json.dumps(out))


if __name__ == "__main__":
    main()
