"""CPU: the bit arithmetic of the four-pixels-per-lane CCL tile kernel (ccl_tile4_kernel, k_ccl.cu) and of the register-resident
sums of the column-strip statistics pass (stats_pass_strips, k_merge.cu), restated in Python and checked against brute force.

These are the checks that were run before the kernels first went to a GPU: the run structure of a 128-pixel tile row from two
ballot words (C = "my first pixel is linked to the lane on my left", T = "my four pixels are linked to each other"), the
vertical-union rule on 4-bit masks, and the two-entry accumulator that keeps a lane's sums in registers."""
import collections
import random

import numpy as np

W, H = 128, 16


def _clz(x):
    return 32 - x.bit_length()


def _row_parents(link, row_base):
    """parents of the 128 pixels of one tile row exactly as the kernel computes them (lane = 4 pixels)"""
    hb = [sum(link[4 * ln + j] << j for j in range(4)) for ln in range(32)]
    c_word = sum((hb[ln] & 1) << ln for ln in range(32))
    t_word = sum((1 if (hb[ln] & 0xE) == 0xE else 0) << ln for ln in range(32))
    lb = [3 if not hb[ln] & 8 else (2 if not hb[ln] & 4 else (1 if not hb[ln] & 2 else 0)) for ln in range(32)]
    par = []
    for ln in range(32):
        stop = (~(c_word & t_word)) & ((1 << ln) - 1) & 0xFFFFFFFF
        if hb[ln] & 1:
            assert stop != 0                       # lane 0 never links left
            sl = 31 - _clz(stop)
            s0 = row_base + 4 * sl + lb[sl]
        else:
            s0 = row_base + 4 * ln
        p = [s0, 0, 0, 0]
        for j in range(1, 4):
            p[j] = p[j - 1] if hb[ln] & (1 << j) else row_base + 4 * ln + j
        par += p
    return hb, par


def test_run_starts_from_two_ballot_words():
    rng = random.Random(1)
    for _ in range(3000):
        pden = rng.choice([0.1, 0.5, 0.8, 0.95, 0.99])
        link = [1 if rng.random() < pden else 0 for _ in range(W)]
        link[0] = 0
        want = [0] * W
        for x in range(W):
            want[x] = want[x - 1] if link[x] else x
        _, par = _row_parents(link, 0)
        assert par == want


def _conn(a, b, d):
    return all(abs(int(a[k]) - int(b[k])) <= d for k in range(3))


def test_tile_labelling_equals_brute_force():
    rng = random.Random(3)
    nprng = np.random.default_rng(3)
    for _ in range(40):
        d = 2
        w, h = rng.choice([128, 124, 64, 8]), rng.choice([16, 15, 3, 1])
        img = (nprng.integers(0, 3, size=(H, W, 3)) * rng.choice([1, 2, 3])).astype(np.int32)
        fg = np.zeros((H, W), bool)
        fg[:h, :w] = True
        lab = -np.ones(H * W, int)
        hbr = {}

        def find(a):
            while lab[a] != a:
                a = lab[a]
            return a

        def union(a, b):
            a, b = find(a), find(b)
            if a != b:
                lab[max(a, b)] = min(a, b)

        for r in range(H):
            link = [0] * W
            for x in range(1, W):
                link[x] = 1 if fg[r, x] and _conn(img[r, x], img[r, x - 1], d) else 0
            hb, par = _row_parents(link, r * W)
            hbr[r] = hb
            for x in range(W):
                lab[r * W + x] = par[x] if fg[r, x] else -1
        for r in range(1, H):
            vbs = [sum((1 << j) for j in range(4) if fg[r, 4 * ln + j] and _conn(img[r, 4 * ln + j], img[r - 1, 4 * ln + j], d))
                   for ln in range(32)]
            for ln in range(32):
                carry = (vbs[ln - 1] >> 3) & 1 if ln > 0 else 0
                need = vbs[ln] & ~(hbr[r][ln] & hbr[r - 1][ln] & ((vbs[ln] << 1) | carry)) & 0xF
                for j in range(4):
                    if need >> j & 1:
                        union(r * W + 4 * ln + j, r * W + 4 * ln + j - W)
        got = np.array([find(p) if lab[p] >= 0 else -1 for p in range(H * W)]).reshape(H, W)
        want = -np.ones((H, W), int)
        for y in range(H):
            for x in range(W):
                if fg[y, x] and want[y, x] < 0:
                    stack = [(y, x)]
                    want[y, x] = y * W + x
                    while stack:
                        cy, cx = stack.pop()
                        for ny, nx in ((cy - 1, cx), (cy + 1, cx), (cy, cx - 1), (cy, cx + 1)):
                            if 0 <= ny < H and 0 <= nx < W and fg[ny, nx] and want[ny, nx] < 0 and _conn(img[cy, cx], img[ny, nx], d):
                                want[ny, nx] = y * W + x
                                stack.append((ny, nx))
        assert (got == want).all()


def test_two_entry_accumulator_loses_nothing():
    rng = random.Random(5)
    for _ in range(2000):
        rows, nlab = rng.randint(1, 64), rng.choice([1, 2, 3, 5])
        total, flushed = collections.Counter(), collections.Counter()
        e0, e1 = [0, 0], [0, 0]

        def flush(e):
            if e[0] > 0 and e[1]:
                flushed[e[0]] += e[1]

        cur = [rng.randint(0, nlab) for _ in range(4)]
        for _y in range(rows):
            if rng.random() < 0.2:
                cur = [rng.randint(0, nlab) for _ in range(4)]
            if rng.random() < 0.3:
                cur = [cur[0]] * 4
            for v in cur:
                if v > 0:
                    total[v] += 1
            uni = cur[0] == cur[1] == cur[2] == cur[3]
            if uni and cur[0] == e1[0] and cur[0] != e0[0]:
                e0, e1 = e1, e0
            if uni and cur[0] == e0[0]:
                e0[1] += 4
            else:
                for v in cur:
                    if v == e0[0]:
                        e0[1] += 1
                    elif v == e1[0]:
                        e1[1] += 1
                    else:
                        flush(e1)
                        e1 = [v, 1]
        flush(e0)
        flush(e1)
        assert flushed == total
