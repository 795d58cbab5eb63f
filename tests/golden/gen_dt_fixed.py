#!/usr/bin/env python
"""Generates tests/golden/dt_fixed.npz: cv2.distanceTransform(mask, DIST_L2, 5) with IPP switched OFF
(cv2.ipp.setUseIPP(False)), i.e. OpenCV's own 16.16 fixed-point chamfer -- what a non-IPP build such as the openpnp 3.4.2
natives the reference binds computes (PictureService.java:1020).  Pins the "dt_fixed" mode of msg_distance_transform and
orc_distance_transform_l2_5_fixed.  Build container only (needs cv2)."""
import os
import sys

import cv2
import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
import msegment_b200 as mseg  # noqa: E402


def main():
    assert hasattr(cv2, "ipp")
    cv2.ipp.setUseIPP(False)
    rng = np.random.default_rng(20261019)
    masks = {}
    for k, (h, w, p) in enumerate([(1, 1, 0.5), (1, 9, 0.7), (7, 1, 0.7), (2, 2, 0.5), (5, 5, 0.9), (33, 47, 0.97), (64, 64, 0.995),
                                   (97, 131, 0.9), (120, 300, 0.999), (200, 150, 0.98)]):
        masks["rand_%d" % k] = ((rng.random((h, w)) < p) * 255).astype(np.uint8)
    masks["all_zero"] = np.zeros((9, 11), np.uint8)
    masks["no_zero"] = np.full((9, 11), 255, np.uint8)
    one = np.full((150, 180), 255, np.uint8)
    one[70, 90] = 0
    masks["single_zero"] = one
    for k, (w, h, seed) in enumerate([(256, 192, 7), (640, 360, 2)]):       # Otsu masks of synthetic frames (the pipeline's input)
        g = cv2.cvtColor(mseg.synth_bgr(w, h, seed), cv2.COLOR_BGR2GRAY)
        _, bw = cv2.threshold(g, 40, 255, cv2.THRESH_BINARY | cv2.THRESH_OTSU)
        masks["otsu_%d" % k] = bw
    out = {"names": np.array(sorted(masks))}
    n_diff_ipp = 0
    for name, m in masks.items():
        out["mask/" + name] = m
        out["dist/" + name] = cv2.distanceTransform(m, cv2.DIST_L2, 5)
    cv2.ipp.setUseIPP(True)
    for name, m in masks.items():
        n_diff_ipp += int((cv2.distanceTransform(m, cv2.DIST_L2, 5) != out["dist/" + name]).sum())
    path = os.path.join(HERE, "dt_fixed.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, os.path.getsize(path), "bytes;", len(masks), "masks; pixels where the IPP build differs:", n_diff_ipp)


if __name__ == "__main__":
    main()
