#!/usr/bin/env python
"""One-off pinning campaign (build container, needs cv2): the oracle's contour-labelling rule (orc_contour_markers) against
findContours(RETR_CCOMP, CHAIN_APPROX_NONE) + the drawContours loop of PictureService.java:360-364 on 1500 masks of six kinds
(noise at several densities, smooth blobs, circles / rings with punched holes, rectangles + lines, 3x3 block noise, diagonal
stripes + speckle), sizes 1x1 .. 160x160.  Result when run for round 1: 0 mismatches of 1500 masks (up to 9713 contours)."""
import numpy as np, cv2, sys, time
sys.path.insert(0, __import__('os').path.dirname(__import__('os').path.dirname(__import__('os').path.dirname(__import__('os').path.abspath(__file__)))))
from oracle import oracle as orc
rng = np.random.default_rng(77)
def ref(mask):
    cs, hier = cv2.findContours(mask, cv2.RETR_CCOMP, cv2.CHAIN_APPROX_NONE)
    m = np.zeros(mask.shape, np.int32)
    for i in range(len(cs)):
        cv2.drawContours(m, cs, i, (i + 1,) * 4, -1, 8, hier, 2**31 - 1, (0, 0))
    return len(cs), m
bad = 0; tot = 0; t0 = time.time(); maxn = 0
for t in range(1500):
    h, w = int(rng.integers(1, 160)), int(rng.integers(1, 160))
    k = t % 6
    if k == 0: a = (rng.random((h, w)) < rng.choice([.2, .4, .55, .62, .7, .85, .95])).astype(np.uint8)
    elif k == 1: a = (cv2.GaussianBlur(rng.random((h, w)).astype(np.float32), (0, 0), float(rng.choice([1, 2, 4]))) > .5).astype(np.uint8)
    elif k == 2:
        a = np.zeros((h, w), np.uint8)
        for _ in range(int(rng.integers(1, 12))):
            c = (int(rng.integers(0, w)), int(rng.integers(0, h))); r = int(rng.integers(1, 40))
            cv2.circle(a, c, r, 1, int(rng.choice([-1, 1, 2, 3])))
            if rng.random() < .5: cv2.circle(a, c, max(r // 2, 1), 0, -1)
    elif k == 3:
        a = np.zeros((h, w), np.uint8)
        for _ in range(int(rng.integers(1, 10))):
            p1 = (int(rng.integers(0, w)), int(rng.integers(0, h))); p2 = (int(rng.integers(0, w)), int(rng.integers(0, h)))
            cv2.rectangle(a, p1, p2, int(rng.integers(0, 2)), int(rng.choice([-1, 1, 2])))
            cv2.line(a, p1, p2, 1, 1)
    elif k == 4: a = np.kron((rng.random((h // 3 + 1, w // 3 + 1)) < .6), np.ones((3, 3))).astype(np.uint8)[:h, :w]
    else: a = ((np.indices((h, w)).sum(0) % int(rng.integers(2, 5))) == 0).astype(np.uint8) | (rng.random((h, w)) < .1).astype(np.uint8)
    a = np.ascontiguousarray(a)
    n, m = ref(a); n2, m2 = orc.contour_markers(a)
    tot += 1; maxn = max(maxn, n)
    if n != n2 or not np.array_equal(m, m2):
        bad += 1
        if bad < 3: print('MISMATCH', t, k, h, w, n, n2)
print('contour rule: %d mismatches of %d masks (max %d contours), %.1f s' % (bad, tot, maxn, time.time() - t0))
