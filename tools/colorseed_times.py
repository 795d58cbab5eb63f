#!/usr/bin/env python
"""Times the colour-method marker generator and the bilateral filter through the host C ABI
(wall time of the synchronous calls incl. pageable copies) with cv2 on one core beside them."""
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
    return round(min(t) * 1e3, 3)


def main():
    w, h = (int(sys.argv[1]), int(sys.argv[2])) if len(sys.argv) > 2 else (1920, 1080)
    im = mseg.synth_bgr(w, h, 2)
    out = {"size": "%dx%d" % (w, h), "note": "host-buffer C-ABI calls incl. pageable H2D/D2H copies; cv2 on 1 thread"}
    try:
        import cv2
        cv2.setNumThreads(1)
    except Exception:
        cv2 = None
    with mseg.Context(0) as ctx:
        gi = mseg.GpuImgproc(ctx)
        n, markers, st = gi.colorSeeds(im, stages=True)
        gray = gi.cvtColorBGR2GRAY(st["sharp"])
        bw, pk = st["bw"], st["peaks"]
        noise = (np.random.default_rng(1).random((h, w)) < .62).astype(np.uint8)
        out["contours"] = n
        out["gpu_ms"] = {"otsu_threshold": best(lambda: gi.threshold(gray, 40, 255, 8)),
                         "distance_transform": best(lambda: gi.distanceTransform(bw)),
                         "contour_markers_peaks": best(lambda: gi.contourMarkers(pk)),
                         "contour_markers_noise": best(lambda: gi.contourMarkers(noise)),   # no cv2 twin: its drawContours loop is O(n^2)
                         "color_seeds_chain": best(lambda: gi.colorSeeds(im)),
                         "bilateral_gray_d11": best(lambda: gi.bilateralFilter(gray, 11, 22, 22)),
                         "bilateral_bgr_d11": best(lambda: gi.bilateralFilter(im, 11, 22, 22))}
        ctx.set_option("dt_fixed", 1)       # OpenCV's own 16.16 fixed-point chamfer: order independent, whole-GPU kernels
        out["gpu_ms"]["distance_transform_fixed_mode"] = best(lambda: gi.distanceTransform(bw))
        out["gpu_ms"]["color_seeds_chain_fixed_mode"] = best(lambda: gi.colorSeeds(im))
        gi.distanceTransform(bw)
        out["gpu_kernel_ms"] = {"distance_transform_fixed_mode": round(ctx.timings()["filter_ms"], 4)}
        ctx.set_option("dt_fixed", 0)
        gi.distanceTransform(bw)
        out["gpu_kernel_ms"]["distance_transform"] = round(ctx.timings()["filter_ms"], 4)
        ev = ctx.stats()
        out["kernel_launches_total"] = ev["kernel_launches"]
    if cv2 is not None:
        def contours(mask):
            cs, hier = cv2.findContours(mask, cv2.RETR_CCOMP, cv2.CHAIN_APPROX_NONE)
            m = np.zeros(mask.shape, np.int32)
            for i in range(len(cs)):
                cv2.drawContours(m, cs, i, (i + 1,) * 4, -1, 8, hier, 2 ** 31 - 1, (0, 0))
            return m

        def chain():
            black = im.copy()
            black[(im == 255).all(axis=2)] = 0
            lap = cv2.filter2D(black, cv2.CV_32F, np.array([1, 1, 1, 1, -8, 1, 1, 1, 1], np.float32).reshape(9, 1))
            sharp = np.clip(np.rint(black.astype(np.float32) - lap), 0, 255).astype(np.uint8)
            g = cv2.cvtColor(sharp, cv2.COLOR_BGR2GRAY)
            _, b = cv2.threshold(g, 40, 255, cv2.THRESH_BINARY | cv2.THRESH_OTSU)
            d = cv2.normalize(cv2.distanceTransform(b, cv2.DIST_L2, 5), None, 0, 1., cv2.NORM_MINMAX)
            _, t = cv2.threshold(d, .4, 1., cv2.THRESH_BINARY)
            p = cv2.dilate(t, np.ones((3, 3), np.uint8)).astype(np.uint8)
            m = contours(p)
            cv2.circle(m, (5, 5), 3, (255, 255, 255), -1)
            return m
        out["cv2_ms"] = {"otsu_threshold": best(lambda: cv2.threshold(gray, 40, 255, cv2.THRESH_BINARY | cv2.THRESH_OTSU)),
                         "distance_transform": best(lambda: cv2.distanceTransform(bw, cv2.DIST_L2, 5)),
                         "contour_markers_peaks": best(lambda: contours(pk), 3),
                         "color_seeds_chain": best(chain, 3),
                         "bilateral_gray_d11": best(lambda: cv2.bilateralFilter(gray, 11, 22, 22), 3),
                         "bilateral_bgr_d11": best(lambda: cv2.bilateralFilter(im, 11, 22, 22), 3)}
        out["chain_equal_cv2"] = bool(np.array_equal(chain(), markers))
    print(json.dumps(out))


if __name__ == "__main__":
    main()
