#!/usr/bin/env python
"""Generates tests/golden/fullsize_digests.json: digests of the REAL OpenCV results (cv2, whole frames, no strips) for the
full-size configurations of BASELINE.json, so that the GPU path is checked against OpenCV itself at size, on the whole frame:

  config 3: 3840x2160, seed 3, sp in {5,10,20} x sr in {10,20,40}, maxLevel 1      (9 variants)
  config 4: 3840x2160, seeds 1000..1015 (the 16-frame subsample), sp = sr = 10       (16 frames)
  config 2: 1920x1080, seed 2

Per unit: SHA-256 of the filtered frame (cv2.pyrMeanShiftFiltering), 512 sampled pixels of it (positions from a fixed
generator, stored), the number of regions and the SHA-256 of the int32 label map of the floodFill labelling loop
(lo = up = 2, 4-connectivity, OpenCV's meanshift_segmentation sample) on that frame.  Needs cv2 (build container); about 25
CPU-minutes, spread over the cores.  The GPU box only reads the JSON."""
import hashlib
import json
import multiprocessing as mp
import os
import sys
import time

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
import msegment_b200 as mseg  # noqa: E402  (numpy synthetic generator: identical to the oracle's and the device's)

N_SAMPLES = 512


def sample_positions(w, h):
    rng = np.random.default_rng(20261019)
    return rng.integers(0, h, N_SAMPLES), rng.integers(0, w, N_SAMPLES)


def floodfill_labels(f, d=2):
    import cv2
    h, w = f.shape[:2]
    mask = np.zeros((h + 2, w + 2), np.uint8)
    lab = np.zeros((h, w), np.int32)
    n = 0
    for y in range(h):
        row = mask[y + 1, 1:-1]
        x = 0
        while True:
            nz = np.flatnonzero(row[x:] == 0)
            if len(nz) == 0:
                break
            x += int(nz[0])
            n += 1
            _, _, _, rect = cv2.floodFill(f, mask, (x, y), (0, 0, 0), (d,) * 3, (d,) * 3, 4 | cv2.FLOODFILL_MASK_ONLY | (2 << 8))
            rx, ry, rw, rh = rect
            sub = mask[ry + 1:ry + 1 + rh, rx + 1:rx + 1 + rw]
            sel = sub == 2
            lab[ry:ry + rh, rx:rx + rw][sel] = n
            sub[sel] = 1
            x += 1
    return n, lab


def unit(args):
    import cv2
    cv2.setNumThreads(1)
    name, w, h, seed, sp, sr = args
    t0 = time.time()
    im = mseg.synth_bgr(w, h, seed)
    f = cv2.pyrMeanShiftFiltering(im, sp, sr, maxLevel=1, termcrit=(3, 5, 1.0))
    n, lab = floodfill_labels(f)
    ys, xs = sample_positions(w, h)
    return name, {"w": w, "h": h, "seed": seed, "sp": sp, "sr": sr,
                  "filtered_sha256": hashlib.sha256(np.ascontiguousarray(f).tobytes()).hexdigest(),
                  "filtered_samples": f[ys, xs].reshape(-1).tolist(),
                  "n_regions": int(n), "labels_sha256": hashlib.sha256(np.ascontiguousarray(lab).tobytes()).hexdigest(),
                  "seconds": round(time.time() - t0, 1)}


def main():
    import cv2
    units = [("c3_sp%d_sr%d" % (sp, sr), 3840, 2160, 3, sp, sr) for sp in (5, 10, 20) for sr in (10, 20, 40)]
    units += [("c4_seed%d" % s, 3840, 2160, s, 10, 10) for s in range(1000, 1016)]
    units += [("c2_seed2", 1920, 1080, 2, 10, 10)]
    units.sort(key=lambda u: -u[4] * u[1])            # longest first
    with mp.get_context("spawn").Pool(max(1, (os.cpu_count() or 2) - 1)) as pool:
        res = dict(pool.map(unit, units, chunksize=1))
    out = {"generator": "tests/golden/gen_fullsize_digests.py", "opencv": cv2.__version__, "n_samples": N_SAMPLES,
           "sample_seed": 20261019, "units": res}
    path = os.path.join(HERE, "fullsize_digests.json")
    json.dump(out, open(path, "w"), indent=0, separators=(",", ":"))
    print("wrote", path, os.path.getsize(path), "bytes")


if __name__ == "__main__":
    main()
