#!/usr/bin/env python
"""Generates tests/golden/*.npz from the real OpenCV (cv2 4.13.0 in this image).

Run from the repo root IN THE BUILD CONTAINER (needs cv2 and, for the two real-image crops,
/root/reference/src/main/resources/images).  The reference itself (Java + OpenCV 3.4.2
natives, pom.xml:39-43) cannot run offline, and ships no tests or golden vectors
(SURVEY.md section 4), so these vectors pin the oracle -- and through it the CUDA path --
against the OpenCV functions the reference's `Imgproc` class binds:
  pyrMeanShiftFiltering / pyrDown / pyrUp           (north_star subsystem 1)
  floodFill loop (meanshift_segmentation sample)    (north_star subsystem 2)
  connectedComponents(mask, 8, CV_32S)              (PictureService.java:441-442)
  watershed                                         (PictureService.java:909)
Inputs are stored with the outputs so the tests never need cv2 or /root/reference.
"""
import os
import sys

import cv2
import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
from oracle import oracle as orc  # noqa: E402  (only for the synthetic-image generator)

REF_IMAGES = "/root/reference/src/main/resources/images"
cv2.setNumThreads(1)
rng = np.random.default_rng(20261018)


def inputs():
    out = {
        "synth96x80": orc.synth_bgr(96, 80, 11),
        "noise41x47": rng.integers(0, 256, (47, 41, 3), dtype=np.uint8),
        "smooth131x97": cv2.GaussianBlur(rng.integers(0, 256, (97, 131, 3), dtype=np.uint8), (0, 0), 3),
        "row50": rng.integers(0, 256, (1, 50, 3), dtype=np.uint8),
        "col50": rng.integers(0, 256, (50, 1, 3), dtype=np.uint8),
        "flat20x33": np.full((33, 20, 3), 77, np.uint8),
    }
    hkp = cv2.imread(os.path.join(REF_IMAGES, "hkp.jpg"))
    guide = cv2.imread(os.path.join(REF_IMAGES, "guide.png"))
    if hkp is not None:
        out["hkp_crop96"] = np.ascontiguousarray(hkp[40:136, 40:136])
    if guide is not None:
        out["guide_crop90x75"] = np.ascontiguousarray(guide[60:135, 70:160])
    return out


MS_PARAMS = [  # sp, sr, maxLevel, (type, maxCount, eps)
    (10.0, 10.0, 1, (3, 5, 1.0)),   # Java 4-arg overload defaults
    (5.5, 20.0, 2, (3, 5, 1.0)),
    (3.0, 3.0, 0, (1, 3, 0.0)),
    (20.0, 40.0, 1, (3, 5, 1.0)),
    (0.3, 1.0, 3, (2, 0, 3.0)),
    (10.0, 2.5, 1, (3, 5, 1.0)),    # sr < 4: isr2 != isr22 in the pyramid mask rule
    (7.0, 25.0, 0, (3, 100, 0.0)),
]


def floodfill_labels(img, d, conn=4):
    """OpenCV samples/cpp/meanshift_segmentation.cpp floodFillPostprocess, recording region ids."""
    h, w = img.shape[:2]
    mask = np.zeros((h + 2, w + 2), np.uint8)
    lab = np.zeros((h, w), np.int32)
    n = 0
    for y in range(h):
        for x in range(w):
            if mask[y + 1, x + 1] == 0:
                n += 1
                before = mask[1:-1, 1:-1] != 0
                cv2.floodFill(img, mask, (x, y), (0, 0, 0), (d, d, d), (d, d, d),
                              conn | cv2.FLOODFILL_MASK_ONLY | (1 << 8))
                lab[(mask[1:-1, 1:-1] != 0) & ~before] = n
    return n, lab


def canonical(lab):
    out = np.zeros_like(lab)
    seen = {}
    flat, o = lab.ravel(), out.ravel()
    for i, v in enumerate(flat):
        if v > 0:
            if v not in seen:
                seen[v] = len(seen) + 1
            o[i] = seen[v]
        else:
            o[i] = v
    return out


def main():
    ins = inputs()
    ms = {}
    for name, im in ins.items():
        ms["in/" + name] = im
        for k, (sp, sr, ml, term) in enumerate(MS_PARAMS):
            ms["out/%s/%d" % (name, k)] = cv2.pyrMeanShiftFiltering(im, sp, sr, maxLevel=ml, termcrit=term)
    ms["params"] = np.array([(sp, sr, ml, t[0], t[1], t[2]) for sp, sr, ml, t in MS_PARAMS], np.float64)
    np.savez_compressed(os.path.join(HERE, "meanshift.npz"), **ms)

    pyr = {}
    for name, im in ins.items():
        if min(im.shape[:2]) < 2:
            continue
        pyr["in/" + name] = im
        pyr["down/" + name] = cv2.pyrDown(im)
        h, w = im.shape[:2]
        pyr["up_even/" + name] = cv2.pyrUp(im, dstsize=(2 * w, 2 * h))
        pyr["up_odd/" + name] = cv2.pyrUp(im, dstsize=(2 * w - 1, 2 * h - 1))
    np.savez_compressed(os.path.join(HERE, "pyramid.npz"), **pyr)

    lab = {}
    for name in ("synth96x80", "smooth131x97", "hkp_crop96", "guide_crop90x75", "row50", "flat20x33"):
        if name not in ins:
            continue
        f = cv2.pyrMeanShiftFiltering(ins[name], 6, 12, maxLevel=1)
        lab["in/" + name] = f
        for d in (0, 2, 5):
            n, l = floodfill_labels(f.copy(), d)
            lab["ff%d/%s" % (d, name)] = l
        n, l = floodfill_labels(f.copy(), 2, 8)
        lab["ff2c8/%s" % name] = l
    for k, (w, h, p) in enumerate([(64, 48, .5), (101, 37, .3), (17, 90, .7), (1, 9, .5), (9, 1, .5), (40, 40, .95)]):
        m = (rng.random((h, w)) < p).astype(np.uint8) * 255
        lab["mask/%d" % k] = m
        for conn in (4, 8):
            n, l = cv2.connectedComponents(m, connectivity=conn, ltype=cv2.CV_32S)
            lab["cc%d/%d" % (conn, k)] = canonical(l)
            lab["ccn%d/%d" % (conn, k)] = np.int32(n)
    np.savez_compressed(os.path.join(HERE, "labels.npz"), **lab)

    ws = {}
    for k, (w, h, seed) in enumerate([(64, 64, 1), (120, 80, 2), (75, 133, 3)]):
        im = orc.synth_bgr(w, h, seed)
        mk = np.zeros((h, w), np.int32)
        for lbl in range(1, 12):
            y, x = rng.integers(1, h - 1), rng.integers(1, w - 1)
            mk[max(y - 2, 0):y + 2, max(x - 2, 0):x + 2] = lbl
        ws["img/%d" % k] = im
        ws["markers/%d" % k] = mk
        out = mk.copy()
        cv2.watershed(im, out)
        ws["out/%d" % k] = out
    np.savez_compressed(os.path.join(HERE, "watershed.npz"), **ws)
    # f2 pre-filters the reference really calls: sharpen chain (PictureService.java:323-333), medianBlur (:408, :436), gray
    fl = {}
    k91 = np.array([1, 1, 1, 1, -8, 1, 1, 1, 1], np.float32).reshape(9, 1)   # literal MatOfFloat reading (9x1 column)
    k33 = k91.reshape(3, 3)                                                 # intended 3x3 Laplacian
    for name in ("synth96x80", "noise41x47", "guide_crop90x75", "row50", "col50"):
        if name not in ins:
            continue
        im = ins[name]
        fl["in/" + name] = im
        for tag, kern in (("k91", k91), ("k33", k33)):
            lap = cv2.filter2D(im, cv2.CV_32F, kern)
            res = im.astype(np.float32) - lap
            fl["sharp_%s/%s" % (tag, name)] = np.clip(np.rint(res), 0, 255).astype(np.uint8)   # == convertTo(CV_8U)
        gray = cv2.cvtColor(im, cv2.COLOR_BGR2GRAY)
        fl["gray/" + name] = gray
        for k in (3, 5, 7, 11):
            fl["median%d/%s" % (k, name)] = cv2.medianBlur(gray, k)
    np.savez_compressed(os.path.join(HERE, "filters.npz"), **fl)
    # f3 shape-method seed generator (PictureService.java:404-442): Canny, dilate, subtract and the whole chain
    sd = {}
    for name in ("synth96x80", "noise41x47", "smooth131x97", "hkp_crop96", "guide_crop90x75", "row50", "col50", "flat20x33"):
        if name not in ins:
            continue
        im = ins[name]
        gray = cv2.cvtColor(im, cv2.COLOR_BGR2GRAY)
        sd["in/" + name] = im
        for k, (lo, hi) in enumerate([(5, 50), (100, 200), (0, 0), (20.7, 60.2)]):
            sd["canny%d/%s" % (k, name)] = cv2.Canny(gray, lo, hi)
        for kw, kh in ((3, 3), (5, 5), (7, 2)):
            sd["dilate%dx%d/%s" % (kw, kh, name)] = cv2.dilate(gray, np.ones((kh, kw), np.uint8))
        h, w = gray.shape
        m = min(w, h)                                    # calculateSizeOfSquareBlurMask (PictureService.java:877-899)
        ksz = 1 if m < 3 else 5                          # all golden inputs are <= 100 px on the short side
        blurred = cv2.medianBlur(gray, ksz)
        edges = cv2.Canny(blurred, 5, 50)                # :416
        d3 = cv2.dilate(edges, np.ones((3, 3), np.uint8))   # :428
        d5 = cv2.dilate(d3, np.ones((5, 5), np.uint8))      # :429
        dde = cv2.subtract(d5, d3)                          # :430
        dde3 = cv2.medianBlur(dde, 3)                       # :436
        n, l = cv2.connectedComponents(dde3, connectivity=8, ltype=cv2.CV_32S)   # :441-442
        sd["chain_edges/" + name] = edges
        sd["chain_dde/" + name] = dde
        sd["chain_dde3/" + name] = dde3
        sd["chain_markers/" + name] = canonical(l)
        sd["chain_n/" + name] = np.int32(n)
    sd["canny_params"] = np.array([(5, 50), (100, 200), (0, 0), (20.7, 60.2)], np.float64)
    np.savez_compressed(os.path.join(HERE, "seeds.npz"), **sd)
    with open(os.path.join(HERE, "PROVENANCE.txt"), "w") as f:
        f.write("generated by tests/golden/gen_golden.py with cv2 %s, numpy %s\n" % (cv2.__version__, np.__version__))
    for fn in sorted(os.listdir(HERE)):
        print(fn, os.path.getsize(os.path.join(HERE, fn)))


if __name__ == "__main__":
    main()
