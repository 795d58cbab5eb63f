"""Deterministic synthetic test image in vectorised numpy -- the host-side twin of msg_synth_bgr_dev.

Jittered-grid Voronoi patches (one site per 64x64 cell) plus triangular noise, all from a counter-based hash
(splitmix64), so the same (width, height, seed) gives the same bytes here, in the CUDA generator and in the oracle's C
generator (tests check all three against each other)."""
import numpy as np

_M = np.uint64(0xFFFFFFFFFFFFFFFF)


def _splitmix64(z):
    z = (z + np.uint64(0x9E3779B97F4A7C15)) & _M
    z = ((z ^ (z >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)) & _M
    z = ((z ^ (z >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)) & _M
    return z ^ (z >> np.uint64(31))


def _hash(sm, a, b, c):
    a = np.asarray(a, dtype=np.uint64)
    b = np.asarray(b, dtype=np.uint64)
    return _splitmix64(sm ^ ((a << np.uint64(40)) | (b << np.uint64(16)) | np.uint64(c)))


def synth_bgr(width, height, seed):
    with np.errstate(over="ignore"):
        sm = _splitmix64(np.uint64(seed))
        ncx, ncy = (width + 63) // 64, (height + 63) // 64
        cx, cy = np.meshgrid(np.arange(ncx), np.arange(ncy))
        sx = 64 * cx + (_hash(sm, cx, cy, 0) % np.uint64(64)).astype(np.int64)
        sy = 64 * cy + (_hash(sm, cx, cy, 1) % np.uint64(64)).astype(np.int64)
        col = (_hash(sm, cx, cy, 2) & np.uint64(0xFFFFFF)).astype(np.int64)
        x, y = np.meshgrid(np.arange(width, dtype=np.int64), np.arange(height, dtype=np.int64))
        cx0, cy0 = x // 64, y // 64
        best = np.full((height, width), -1, np.int64)
        bc = np.zeros((height, width), np.int64)
        for dy in (-1, 0, 1):                     # scan order (cy, cx) ascending, strict '<' keeps the first minimum
            for dx in (-1, 0, 1):
                ccx, ccy = cx0 + dx, cy0 + dy
                ok = (ccx >= 0) & (ccy >= 0) & (ccx < ncx) & (ccy < ncy)
                ix, iy = np.clip(ccx, 0, ncx - 1), np.clip(ccy, 0, ncy - 1)
                d = (x - sx[iy, ix]) ** 2 + (y - sy[iy, ix]) ** 2
                take = ok & ((best < 0) | (d < best))
                best = np.where(take, d, best)
                bc = np.where(take, col[iy, ix], bc)
        out = np.empty((height, width, 3), np.uint8)
        for c in range(3):
            base = (bc >> (8 * c)) & 0xFF
            nz = ((_hash(sm, x, y, 16 + c) % np.uint64(13)).astype(np.int64) +
                  (_hash(sm, x, y, 32 + c) % np.uint64(13)).astype(np.int64) - 12)
            out[..., c] = np.clip(base + nz, 0, 255).astype(np.uint8)
    return out
