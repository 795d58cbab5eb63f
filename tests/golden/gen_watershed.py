#!/usr/bin/env python
"""Generates tests/golden/watershed2.npz: inputs + cv2.watershed outputs (Imgproc.watershed, PictureService.java:909) for
the exact GPU flood (msg_watershed) and the oracle (orc_watershed).  Run in the build container (needs cv2 and
/root/reference for the two real-image cases); the GPU box only reads the .npz.

Cases: tiny / degenerate sizes, noise (many level changes), seeds on the border and negative values in the input markers
(cv::watershed resets them), markers produced by transliterations of the reference's own pipelines
(PictureService.java:301-372 colour method, :396-457 shape method) on crops of its sample images and on synthetic frames.
"""
import os
import sys

import cv2
import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
import msegment_b200 as mseg  # noqa: E402  (numpy synthetic generator only)

REF_IMAGES = "/root/reference/src/main/resources/images"


def shape_markers(im):
    """PictureService.java:396-442 (k = 7 for these sizes)."""
    gray = cv2.cvtColor(im, cv2.COLOR_BGR2GRAY)
    gray = cv2.medianBlur(gray, 7 if min(im.shape[:2]) >= 200 else 5)
    edges = cv2.Canny(gray, 5, 50)
    d3 = cv2.dilate(edges, np.ones((3, 3), np.uint8))
    d5 = cv2.dilate(d3, np.ones((5, 5), np.uint8))
    band = cv2.medianBlur(cv2.subtract(d5, d3), 3)
    _, markers = cv2.connectedComponents(band, connectivity=8, ltype=cv2.CV_32S)
    return markers


def color_markers(im):
    """PictureService.java:301-366 (literal 9x1 sharpen kernel)."""
    k = np.array([1, 1, 1, 1, -8, 1, 1, 1, 1], np.float32).reshape(9, 1)
    lap = cv2.filter2D(im, cv2.CV_32F, k)
    sharp = np.clip(np.rint(im.astype(np.float32) - lap), 0, 255).astype(np.uint8)
    gray = cv2.cvtColor(sharp, cv2.COLOR_BGR2GRAY)
    _, bw = cv2.threshold(gray, 40, 255, cv2.THRESH_BINARY | cv2.THRESH_OTSU)
    dist = cv2.distanceTransform(bw, cv2.DIST_L2, 5)
    dist = cv2.normalize(dist, None, 0, 1.0, cv2.NORM_MINMAX)
    _, peaks = cv2.threshold(dist, 0.4, 1.0, cv2.THRESH_BINARY)
    peaks = cv2.dilate(peaks, np.ones((3, 3), np.uint8)).astype(np.uint8)
    contours, hier = cv2.findContours(peaks, cv2.RETR_CCOMP, cv2.CHAIN_APPROX_NONE)
    markers = np.zeros(im.shape[:2], np.int32)
    for i in range(len(contours)):
        cv2.drawContours(markers, contours, i, (i + 1,), -1, 8, hier, 2 ** 31 - 1)
    cv2.circle(markers, (5, 5), 3, (255, 255, 255), -1)
    return sharp, markers


def main():
    rng = np.random.default_rng(20261019)
    cases = []

    def add(name, im, mk):
        out = mk.copy()
        cv2.watershed(np.ascontiguousarray(im), out)
        cases.append((name, im, mk, out))

    # degenerate sizes: everything is border
    for (h, w) in [(1, 1), (1, 7), (5, 1), (2, 2), (2, 9), (3, 3), (3, 8), (4, 4)]:
        im = rng.integers(0, 256, (h, w, 3), dtype=np.uint8)
        mk = rng.integers(-2, 4, (h, w)).astype(np.int32)
        add("tiny_%dx%d" % (h, w), im, mk)
    # pure noise, sparse point seeds; negative junk in the markers and seeds on the border
    for k, (h, w) in enumerate([(40, 53), (64, 64), (97, 131)]):
        im = rng.integers(0, 256, (h, w, 3), dtype=np.uint8)
        mk = np.zeros((h, w), np.int32)
        for s in range(1, 9):
            mk[rng.integers(0, h), rng.integers(0, w)] = s
        mk[rng.random((h, w)) < 0.02] = -1
        mk[rng.random((h, w)) < 0.01] = -7
        mk[0, :5] = 3
        mk[h - 1, w - 4:] = 2
        add("noise_%d" % k, im, mk)
    # flat image: every difference is 0 (pure FIFO order), two and five seeds
    for k, ns in enumerate((2, 5)):
        im = np.full((50, 70, 3), 90, np.uint8)
        mk = np.zeros((50, 70), np.int32)
        for s in range(1, ns + 1):
            mk[rng.integers(1, 49), rng.integers(1, 69)] = s
        add("flat_%d" % k, im, mk)
    # low-amplitude noise (levels 0..6: frequent level changes), blob seeds
    im = (120 + rng.integers(-3, 4, (120, 160, 3))).astype(np.uint8)
    mk = np.zeros((120, 160), np.int32)
    for s in range(1, 13):
        y, x = rng.integers(5, 110), rng.integers(5, 150)
        mk[y:y + 4, x:x + 5] = s
    add("lowamp", im, mk)
    # unreachable pocket: a zero region fenced by WSHED cannot happen in the input (negatives are reset), but a region
    # fenced by another basin's pixels can stay 0 only if no seed reaches it -> an image with no seeds at all
    add("noseeds", rng.integers(0, 256, (30, 30, 3), dtype=np.uint8), np.zeros((30, 30), np.int32))
    # synthetic frames with the reference pipelines' markers
    for k, (w, h, seed) in enumerate([(200, 150, 1), (256, 192, 7), (131, 97, 3)]):
        im = mseg.synth_bgr(w, h, seed)
        add("synth_shape_%d" % k, im, shape_markers(im))
        sharp, mk = color_markers(im)
        add("synth_color_%d" % k, sharp, mk)
    # the reference's own sample images (crops)
    for name, (y0, x0, hh, ww) in (("hkp.jpg", (10, 20, 150, 150)), ("guide.png", (0, 0, 200, 220)), ("haha.jpg", (60, 80, 180, 240))):
        path = os.path.join(REF_IMAGES, name)
        if not os.path.exists(path):
            continue
        im = np.ascontiguousarray(cv2.imread(path)[y0:y0 + hh, x0:x0 + ww])
        add("ref_shape_" + name.split(".")[0], im, shape_markers(im))
        sharp, mk = color_markers(im)
        add("ref_color_" + name.split(".")[0], sharp, mk)
    out = {"names": np.array([c[0] for c in cases])}
    for name, im, mk, res in cases:
        out["img/" + name] = im
        out["markers/" + name] = mk
        out["out/" + name] = res
    path = os.path.join(HERE, "watershed2.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, os.path.getsize(path), "bytes,", len(cases), "cases")


if __name__ == "__main__":
    main()
