#!/usr/bin/env python
"""HBM-roofline figures of the label / merge / render stages (7, 11, 7 algorithmic bytes per pixel) over image
sizes, device-resident, CUDA events on the stream the kernels run on (the context is put on torch's current stream).
The input of the label stage is the mean-shift output of the synthetic image (computed once per size, not timed)."""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

import msegment_b200 as mseg  # noqa: E402

dev = mseg.device


def timed(fn, reps):
    fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    best = 1e9
    for _ in range(reps):
        a.record()
        fn()
        b.record()
        torch.cuda.synchronize()
        best = min(best, a.elapsed_time(b))
    return best


def main():
    try:
        hbm = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"]
    except Exception:
        hbm = 6650.0
    sizes = [(1920, 1080), (3840, 2160), (8192, 8192), (16384, 8192)]
    if os.environ.get("SIZES"):
        sizes = [tuple(int(v) for v in s.split("x")) for s in os.environ["SIZES"].split(",")]
    reps = int(os.environ.get("REPS", "5"))
    grid = int(os.environ.get("MERGE_GRID", "0"))
    out = {"peak_gbs": hbm, "merge_grid": grid or "auto", "rows": []}
    torch.cuda.set_device(0)
    with mseg.Context(0) as ctx:
        ctx.set_stream(torch.cuda.current_stream().cuda_stream)
        if grid:
            ctx.set_option("merge_grid", grid)
        ctx.set_option("labels_canonical", 1)     # the labels come straight from msg_label_regions_dev: no validation passes
        for w, h in sizes:
            n = w * h
            src = torch.empty((h, w, 3), dtype=torch.uint8, device="cuda")
            filt = torch.empty_like(src)
            lab = torch.empty((h, w), dtype=torch.int32, device="cuda")
            lab0 = torch.empty_like(lab)
            ren = torch.empty_like(src)
            cnt = torch.zeros(4, dtype=torch.int32, device="cuda")
            dev.synth(ctx, src.data_ptr(), 3 * w, w, h, 2)
            dev.meanshift(ctx, src.data_ptr(), 3 * w, filt.data_ptr(), 3 * w, w, h, 10, 10)
            t_label = timed(lambda: dev.label_regions(ctx, filt.data_ptr(), 3 * w, lab0.data_ptr(), 4 * w, w, h, 2, cnt.data_ptr()), reps)
            n0 = int(cnt[0].item())
            cnt0 = cnt.clone()

            def merge():
                lab.copy_(lab0)
                cnt.copy_(cnt0)                     # option labels_canonical: the label count goes in through *d_n_regions
                dev.merge_regions(ctx, filt.data_ptr(), 3 * w, lab.data_ptr(), 4 * w, w, h, 50, 10, cnt.data_ptr())
            t_copy = timed(lambda: (lab.copy_(lab0), cnt.copy_(cnt0)), reps)
            t_merge = timed(merge, reps) - t_copy
            n1 = int(cnt[0].item())
            ctx.synchronize()
            rounds = int(ctx.stats()["merge_rounds"])
            t_render = timed(lambda: dev.render_labels(ctx, lab.data_ptr(), 4 * w, ren.data_ptr(), 3 * w, w, h, n1), reps)
            row = {"size": "%dx%d" % (w, h), "mpix": round(n / 1e6, 2), "regions": n0, "regions_after_merge": n1, "merge_rounds": rounds}
            for name, t, bpp in (("label", t_label, 7), ("merge", t_merge, 11), ("render", t_render, 7)):
                gbs = bpp * n / (t * 1e-3) / 1e9
                row[name] = {"ms": round(t, 4), "gbs": round(gbs, 1), "frac": round(gbs / hbm, 4)}
            out["rows"].append(row)
            print(json.dumps(row), flush=True)
            del src, filt, lab, lab0, ren
            torch.cuda.empty_cache()
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    json.dump(out, open(os.path.join(ROOT, "gpurun_out", "stage_rooflines.json"), "w"), indent=1)


if __name__ == "__main__":
    main()
