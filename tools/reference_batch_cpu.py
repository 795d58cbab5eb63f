#!/usr/bin/env python
"""BASELINE.json config 1 (i): the reference's own CLI batch, timed on the CPU through cv2 -- REPORTED BASELINE ONLY.

A Python transliteration (cv2 calls, parameters from BASELINE.md section 2) of what `App.main` runs
(App.java:28-29): PictureService.colorAutoMarkerWatershed (PictureService.java:301-382) and
shapeAutoMarkerWatershed (:396-467), each producing its 8-Result batch.  The Java program itself cannot
run here (no JDK), and its per-pixel Java loops (:309-318, :925-934) are not reproduced -- the OpenCV calls are.
Quirks kept: the white->black loop is a no-op (App. C#1); the Laplacian kernel is the literal 9x1 column (App. C#2,
--kernel 3x3 for the intended reading); depth = number of contours incl. holes (App. C#3); colored = false.

  python tools/reference_batch_cpu.py [--size 512] [--seed 1] [--image file]      -> one JSON line
"""
import argparse
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import cv2  # noqa: E402
import numpy as np  # noqa: E402


def blur_mask_size(gray):                      # PictureService.calculateSizeOfSquareBlurMask, :877-899
    m = min(gray.shape[:2])
    if m < 3:
        return 1
    if m <= 100:
        return 5
    scale = 0.025 if m <= 360 else 0.02 if m <= 480 else 0.015 if m <= 720 else 0.01 if m <= 1080 else 0.005
    r = int(m * scale)
    return r + 1 if r % 2 == 0 else r


def color_by_indexes(markers, depth):           # :913-936, colored = false
    dst = np.zeros(markers.shape + (3,), np.uint8)
    dst[(markers > 0) & (markers <= depth)] = 255
    return dst


class Timer:
    def __init__(self):
        self.t = {}

    def run(self, name, fn):
        t0 = time.perf_counter()
        r = fn()
        self.t[name] = self.t.get(name, 0.0) + (time.perf_counter() - t0) * 1e3
        return r


def color_pipeline(src, kernel_shape, tm):
    results = [("black_bg", src.copy())]                                                  # :309-321 (no-op loop)
    kern = np.array([1, 1, 1, 1, -8, 1, 1, 1, 1], np.float32).reshape(kernel_shape)
    lap = tm.run("filter2D", lambda: cv2.filter2D(src, cv2.CV_32F, kern))                 # :323-327
    res = tm.run("sharpen_arith", lambda: np.clip(np.rint(src.astype(np.float32) - lap), 0, 255).astype(np.uint8))  # :328-333
    src = res
    results.append(("laplassian_sharp", res.copy()))
    gray = tm.run("cvtColor", lambda: cv2.cvtColor(src, cv2.COLOR_BGR2GRAY))              # :940
    _, bw = tm.run("threshold_otsu", lambda: cv2.threshold(gray, 40, 255, cv2.THRESH_BINARY | cv2.THRESH_OTSU))   # :941
    results.append(("bw", bw.copy()))
    dist = tm.run("distanceTransform", lambda: cv2.distanceTransform(bw, cv2.DIST_L2, 5))  # :1020
    dist = tm.run("normalize", lambda: cv2.normalize(dist, None, 0, 1.0, cv2.NORM_MINMAX))  # :1021
    results.append(("distance_transform", dist.copy()))
    _, peaks = tm.run("threshold", lambda: cv2.threshold(dist, .4, 1., cv2.THRESH_BINARY))  # :348
    peaks = tm.run("dilate", lambda: cv2.dilate(peaks, np.ones((3, 3), np.uint8)))        # :349-350
    results.append(("distance_peaks", peaks.copy()))
    d8 = peaks.astype(np.uint8)                                                           # :355-356
    contours, hier = tm.run("findContours", lambda: cv2.findContours(d8, cv2.RETR_CCOMP, cv2.CHAIN_APPROX_NONE))  # :360
    markers = np.zeros(d8.shape, np.int32)

    def draw():
        for i in range(len(contours)):                                                    # :361-364
            cv2.drawContours(markers, contours, i, (i + 1,) * 4, -1, 8, hier, 2 ** 31 - 1)
    tm.run("drawContours", draw)
    depth = len(contours)                                                                 # :365
    cv2.circle(markers, (5, 5), 3, (255, 255, 255), -1)                                   # :366
    results.append(("markers", markers.copy()))
    tm.run("watershed", lambda: cv2.watershed(src, markers))                              # :909
    dst = tm.run("colorByIndexes", lambda: color_by_indexes(markers, depth))
    results.append(("result", dst))
    results.append(("bw_result", cv2.cvtColor(dst, cv2.COLOR_BGR2GRAY)))                  # :376-378
    return results


def shape_pipeline(src, tm):
    gray = tm.run("cvtColor", lambda: cv2.cvtColor(src, cv2.COLOR_BGR2GRAY))              # :405
    k = blur_mask_size(gray)
    gray = tm.run("medianBlur_k", lambda: cv2.medianBlur(gray, k))                        # :408
    results = [("blured_by_%dx%d" % (k, k), gray.copy())]
    edges = tm.run("Canny", lambda: cv2.Canny(gray, 5, 50))                               # :416
    brd = np.zeros_like(src)
    brd[edges != 0] = src[edges != 0]                                                     # :418 copyTo(mask)
    results += [("borders", brd), ("gray_borders", edges.copy())]
    d3 = tm.run("dilate", lambda: cv2.dilate(edges, np.ones((3, 3), np.uint8)))           # :428
    d5 = tm.run("dilate", lambda: cv2.dilate(d3, np.ones((5, 5), np.uint8)))              # :429
    mask = cv2.subtract(d5, d3)                                                           # :430
    results.append(("dde_step", mask.copy()))
    mask = tm.run("medianBlur_3", lambda: cv2.medianBlur(mask, 3))                        # :436
    results.append(("dde_step_blurred_3x3", mask.copy()))
    _, markers = tm.run("connectedComponents", lambda: cv2.connectedComponents(mask, connectivity=8, ltype=cv2.CV_32S))  # :442
    results.append(("markers", markers.copy()))
    contours, _ = tm.run("findContours", lambda: cv2.findContours(mask, cv2.RETR_CCOMP, cv2.CHAIN_APPROX_NONE))          # :450
    if not contours:
        return results
    depth = len(contours)
    tm.run("watershed", lambda: cv2.watershed(src, markers))                              # :457 -> :909
    dst = tm.run("colorByIndexes", lambda: color_by_indexes(markers, depth))
    results += [("result", dst), ("bw_result", cv2.cvtColor(dst, cv2.COLOR_BGR2GRAY))]
    return results


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--size", type=int, default=512)
    ap.add_argument("--seed", type=int, default=1)
    ap.add_argument("--image", default=None)
    ap.add_argument("--kernel", default="9x1", choices=["9x1", "3x3"])
    ap.add_argument("--reps", type=int, default=3)
    args = ap.parse_args()
    cv2.setNumThreads(1)
    if args.image:
        src = cv2.imread(args.image)
    else:
        import msegment_b200                    # host-side numpy generator (no CUDA needed)
        src = msegment_b200.synth_bgr(args.size, args.size, args.seed)
    kshape = (9, 1) if args.kernel == "9x1" else (3, 3)
    best = None
    for _ in range(args.reps):
        tc, ts = Timer(), Timer()
        t0 = time.perf_counter()
        rc = color_pipeline(src.copy(), kshape, tc)
        t1 = time.perf_counter()
        rs = shape_pipeline(src.copy(), ts)
        t2 = time.perf_counter()
        cur = dict(color_ms=(t1 - t0) * 1e3, shape_ms=(t2 - t1) * 1e3, color=tc.t, shape=ts.t, n_color=len(rc), n_shape=len(rs))
        if best is None or cur["color_ms"] + cur["shape_ms"] < best["color_ms"] + best["shape_ms"]:
            best = cur
    h, w = src.shape[:2]
    tot = best["color_ms"] + best["shape_ms"]
    print(json.dumps({"config": "reference CLI batch (color + shape pipelines, OpenCV calls only) on %dx%d, cv2 %s, 1 thread" % (w, h, cv2.__version__),
                      "results_in_batch": best["n_color"] + best["n_shape"], "total_ms": round(tot, 2),
                      "mpix_per_s": round(w * h / 1e6 / (tot / 1e3), 2),
                      "color_pipeline_ms": round(best["color_ms"], 2), "shape_pipeline_ms": round(best["shape_ms"], 2),
                      "color_stages_ms": {k: round(v, 2) for k, v in best["color"].items()},
                      "shape_stages_ms": {k: round(v, 2) for k, v in best["shape"].items()},
                      "note": "reported baseline only; the Java per-pixel loops (2-3 JNI calls per pixel) are not included"}))


if __name__ == "__main__":
    main()
