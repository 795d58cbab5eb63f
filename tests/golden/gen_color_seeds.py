#!/usr/bin/env python
"""Generates tests/golden/color_seeds.npz from the real OpenCV (cv2 4.13.0 in this image): the colour-method marker
generator of the reference (PictureService.java:309-366, :938-943, :1018-1023) stage by stage, the
contour labelling on random masks (nested holes, islands), filled circles and the bilateral filter (:490).

Run from the repo root IN THE BUILD CONTAINER (needs cv2 and, for the two real-image crops, /root/reference).  Inputs are
stored with the outputs so the tests never need cv2 or /root/reference.

Distance transform: this cv2 build dispatches distanceTransform(DIST_L2, 5) to IPP, which accumulates the float metrics
1 / 1.4 / 2.1969 in float32 with an unrounded running value inside aligned groups of four columns of the forward pass; the
oracle restates exactly that (oracle/msg_oracle.c: orc_distance_transform_l2_5) and equals cv2 bit for bit on every image tried
(asserted below for the stored vectors)."""
import os
import sys

import cv2
import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
import gen_golden  # noqa: E402  (shared input images; its module-level rng makes inputs() reproducible when called first)

cv2.setNumThreads(1)


def contour_markers(mask):
    cs, hier = cv2.findContours(mask, cv2.RETR_CCOMP, cv2.CHAIN_APPROX_NONE)      # PictureService.java:360
    m = np.zeros(mask.shape, np.int32)
    for i in range(len(cs)):                                                       # :361-364
        cv2.drawContours(m, cs, i, (i + 1,) * 4, -1, 8, hier, 2 ** 31 - 1, (0, 0))
    return len(cs), m


def color_chain(im):
    black = im.copy()
    black[(im == 255).all(axis=2)] = 0                                             # :309-318
    k91 = np.array([1, 1, 1, 1, -8, 1, 1, 1, 1], np.float32).reshape(9, 1)         # :320 (literal 9 x 1 reading)
    lap = cv2.filter2D(black, cv2.CV_32F, k91)
    sharp = np.clip(np.rint(black.astype(np.float32) - lap), 0, 255).astype(np.uint8)   # :323-329
    gray = cv2.cvtColor(sharp, cv2.COLOR_BGR2GRAY)                                 # :940
    t, bw = cv2.threshold(gray, 40, 255, cv2.THRESH_BINARY | cv2.THRESH_OTSU)      # :941
    dist = cv2.distanceTransform(bw, cv2.DIST_L2, 5)                               # :1020
    nrm = cv2.normalize(dist, None, 0, 1., cv2.NORM_MINMAX)                        # :1021
    _, th = cv2.threshold(nrm, .4, 1., cv2.THRESH_BINARY)                          # :348
    pk = cv2.dilate(th, np.ones((3, 3), np.uint8)).astype(np.uint8)                # :349-356
    n, m = contour_markers(pk)
    final = m.copy()
    cv2.circle(final, (5, 5), 3, (255, 255, 255), -1)                              # :366
    return dict(black=black, sharp=sharp, gray=gray, otsu=np.int32(t), bw=bw, dist=dist, norm=nrm, peaks=pk,
                n=np.int32(n), contours=m, markers=final)


def main():
    ins = gen_golden.inputs()
    rng = np.random.default_rng(20261019)
    out = {}
    names = [n for n in ("synth96x80", "noise41x47", "smooth131x97", "hkp_crop96", "guide_crop90x75", "row50", "col50",
                         "flat20x33") if n in ins]
    # a blob image whose Otsu foreground has holes and islands
    blob = cv2.GaussianBlur(rng.integers(0, 256, (120, 150, 3), dtype=np.uint8), (0, 0), 5)
    blob = np.clip((blob.astype(np.int32) - 128) * 6 + 128, 0, 255).astype(np.uint8)
    blob[10:20, 10:30] = 255                                                       # pure white patch: the white->black loop
    ins = dict(ins, blob150x120=blob)
    names.append("blob150x120")
    for name in names:
        im = ins[name]
        out["in/" + name] = im
        for k, v in color_chain(im).items():
            out["%s/%s" % (k, name)] = v
    # contour labelling, distance transform and circles on masks
    masks = []
    for (w, h, p) in [(64, 48, .5), (101, 37, .62), (40, 40, .8), (33, 57, .9), (1, 9, .5), (9, 1, .5), (50, 50, .35), (80, 60, .97)]:
        masks.append((rng.random((h, w)) < p).astype(np.uint8))
    smooth = (cv2.GaussianBlur(rng.random((90, 120)).astype(np.float32), (0, 0), 3) > .5).astype(np.uint8)
    ring = np.zeros((60, 70), np.uint8)
    cv2.circle(ring, (35, 30), 28, 1, 5); cv2.circle(ring, (35, 30), 18, 1, 4); cv2.circle(ring, (35, 30), 8, 1, -1)
    ring[30, 33:38] = 0
    masks += [smooth, ring, np.ones((12, 17), np.uint8), np.zeros((7, 5), np.uint8)]
    for k, m in enumerate(masks):
        out["mask/%d" % k] = m
        n, lab = contour_markers(m)
        out["mask_n/%d" % k] = np.int32(n)
        out["mask_markers/%d" % k] = lab
        out["mask_dist/%d" % k] = cv2.distanceTransform(m * 255, cv2.DIST_L2, 5)
        c = lab.copy()
        cx, cy, r = int(rng.integers(-3, m.shape[1] + 3)), int(rng.integers(-3, m.shape[0] + 3)), int(rng.integers(0, 9))
        cv2.circle(c, (cx, cy), r, (77, 77, 77), -1)
        out["mask_circle/%d" % k] = c
        out["mask_circle_args/%d" % k] = np.array([cx, cy, r, 77], np.int32)
    # distance transform where the unrounded running value of the forward pass matters: a horizontal step crosses 32 / 64 on an
    # exact rounding tie (single zero pixels far away), at several group phases (x % 4) and row tails (w % 4)
    for k, (w, h, zy, zx) in enumerate([(140, 40, 0, 0), (141, 40, 0, 1), (139, 36, 0, 2), (123, 120, 119, 122), (75, 70, 0, 3),
                                        (130, 33, 2, 5)]):
        m = np.full((h, w), 1, np.uint8)
        m[zy, zx] = 0
        out["dt_tie_mask/%d" % k] = m
        out["dt_tie/%d" % k] = cv2.distanceTransform(m * 255, cv2.DIST_L2, 5)
    # bilateral filter (d = mask, sigmas = 2 * mask as PictureService.java:490 calls it)
    for name in ("synth96x80", "noise41x47", "hkp_crop96", "row50", "col50"):
        if name not in ins:
            continue
        im = ins[name]
        gray = cv2.cvtColor(im, cv2.COLOR_BGR2GRAY)
        for d in (5, 7):
            out["bil_gray%d/%s" % (d, name)] = cv2.bilateralFilter(gray, d, 2 * d, 2 * d)
            out["bil_bgr%d/%s" % (d, name)] = cv2.bilateralFilter(im, d, 2 * d, 2 * d)
    # the stored distance transforms must be free of the IPP rounding corner described in the module docstring
    from oracle import oracle as orc
    for key in list(out):
        if key.startswith("dist/"):
            assert np.array_equal(orc.distance_transform(out["bw/" + key[5:]]), out[key]), key
        if key.startswith("mask_dist/"):
            assert np.array_equal(orc.distance_transform(out["mask/" + key[10:]] * 255), out[key]), key
        if key.startswith("dt_tie/"):
            assert np.array_equal(orc.distance_transform(out["dt_tie_mask/" + key[7:]] * 255), out[key]), key
    np.savez_compressed(os.path.join(HERE, "color_seeds.npz"), **out)
    with open(os.path.join(HERE, "PROVENANCE.txt"), "a") as f:
        f.write("color_seeds.npz generated by tests/golden/gen_color_seeds.py with cv2 %s, numpy %s\n" % (cv2.__version__, np.__version__))
    print("color_seeds.npz", os.path.getsize(os.path.join(HERE, "color_seeds.npz")), "bytes,", len(out), "arrays")


if __name__ == "__main__":
    main()
