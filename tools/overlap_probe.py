#!/usr/bin/env python
"""How much of the non-mean-shift stages does the multi-stream step hide?  Device-resident throughput of the 4K bench workload
with the stages switched on one by one (same frames, same contexts / streams as bench.py):
  filter        msg_segment_dev with lo_diff < 0 (mean shift only)
  +label        labelling, no merge
  +label+merge  the bench configuration
Usage: python tools/overlap_probe.py [streams] [frames]"""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

import bench  # noqa: E402
import msegment_b200 as mseg  # noqa: E402


def main():
    ns = int(sys.argv[1]) if len(sys.argv) > 1 else bench.N_STREAMS
    frames = int(sys.argv[2]) if len(sys.argv) > 2 else bench.FRAMES_PER_STEP
    torch.cuda.set_device(0)
    wl = bench.Workload(torch, mseg, 0, 0, bench.W, bench.H, frames, ns)
    dev = mseg.device
    base = dict(bench.PARAMS)
    variants = {"filter": dict(base, lo_diff=-1, min_size=0, color_dist=0),
                "+label": dict(base, min_size=0, color_dist=0),
                "+label+merge": base}
    out = {"workload": bench.workload_name(), "streams": ns, "frames_per_step": frames, "rows": []}
    for name, p in variants.items():
        wl.prm = dev.params(**p, render_depth=-1)
        for _ in range(3):
            wl.step_device()
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        steps = 5
        a.record(wl.streams[0])                      # events on the streams the kernels run on (as bench.py's timed())
        for s in wl.streams[1:]:
            s.wait_event(a)
        for _ in range(steps):
            wl.step_device()
        for s in wl.streams[1:]:
            done = torch.cuda.Event()
            done.record(s)
            wl.streams[0].wait_event(done)
        b.record(wl.streams[0])
        torch.cuda.synchronize()
        ms = a.elapsed_time(b) / steps
        row = {"stages": name, "ms_per_step": round(ms, 3), "ms_per_frame": round(ms / frames, 4),
               "mpix_per_s": round(frames * bench.W * bench.H / ms / 1e3, 1)}
        out["rows"].append(row)
        print(json.dumps(row), flush=True)
    wl.close()
    print(json.dumps(out))


if __name__ == "__main__":
    main()
