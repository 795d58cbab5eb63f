#!/usr/bin/env python
"""Times the pre-filter kernels through the host C ABI (msg_get_timings is not wired for them, so CUDA
events are not available here: wall time of the synchronous call, pinned-free) and cv2 on one core beside them."""
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np  # noqa: E402

import msegment_b200 as mseg  # noqa: E402


def best(fn, reps=5):
    fn()
    t = []
    for _ in range(reps):
        t0 = time.perf_counter()
        fn()
        t.append(time.perf_counter() - t0)
    return min(t) * 1e3


def main():
    w, h = 1920, 1080
    im = mseg.synth_bgr(w, h, 2)
    taps = np.array([1, 1, 1, 1, -8, 1, 1, 1, 1], np.int8)
    out = {"size": "%dx%d" % (w, h), "note": "host-buffer C-ABI calls incl. pageable H2D/D2H copies; cv2 on 1 thread"}
    try:
        import cv2
        cv2.setNumThreads(1)
    except Exception:
        cv2 = None
    with mseg.Context(0) as ctx:
        gi = mseg.GpuImgproc(ctx)
        gray = gi.cvtColorBGR2GRAY(im)
        out["gpu_ms"] = {"sharpen_9x1": round(best(lambda: gi.sharpenLaplacian(im, taps.reshape(9, 1))), 3),
                         "sharpen_3x3": round(best(lambda: gi.sharpenLaplacian(im, taps.reshape(3, 3))), 3),
                         "bgr2gray": round(best(lambda: gi.cvtColorBGR2GRAY(im)), 3),
                         "median_3": round(best(lambda: gi.medianBlur(gray, 3)), 3),
                         "median_11": round(best(lambda: gi.medianBlur(gray, 11)), 3),
                         "canny_5_50": round(best(lambda: gi.Canny(gray, 5, 50)), 3),
                         "shape_seeds_chain": round(best(lambda: gi.shapeSeeds(im)), 3)}
    if cv2 is not None:
        k91 = taps.reshape(9, 1).astype(np.float32)

        def sharp():
            lap = cv2.filter2D(im, cv2.CV_32F, k91)
            return np.clip(np.rint(im.astype(np.float32) - lap), 0, 255).astype(np.uint8)
        out["cv2_ms"] = {"sharpen_9x1": round(best(sharp, 3), 3),
                         "bgr2gray": round(best(lambda: cv2.cvtColor(im, cv2.COLOR_BGR2GRAY)), 3),
                         "median_3": round(best(lambda: cv2.medianBlur(gray, 3)), 3),
                         "median_11": round(best(lambda: cv2.medianBlur(gray, 11)), 3),
                         "canny_5_50": round(best(lambda: cv2.Canny(gray, 5, 50)), 3)}

        def chain():
            g = cv2.medianBlur(cv2.cvtColor(im, cv2.COLOR_BGR2GRAY), 11)
            e = cv2.Canny(g, 5, 50)
            d3 = cv2.dilate(e, np.ones((3, 3), np.uint8))
            d5 = cv2.dilate(d3, np.ones((5, 5), np.uint8))
            m = cv2.medianBlur(cv2.subtract(d5, d3), 3)
            return cv2.connectedComponents(m, connectivity=8, ltype=cv2.CV_32S)
        out["cv2_ms"]["shape_seeds_chain"] = round(best(chain, 3), 3)
    print(json.dumps(out))


if __name__ == "__main__":
    main()
