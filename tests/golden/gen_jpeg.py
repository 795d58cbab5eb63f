#!/usr/bin/env python
"""Generates the JPEG fixtures of the CLI's reader (opencv-msegment_b200/host/jpeg_io.hpp): small baseline files written by
cv2.imwrite (libjpeg-turbo) over sizes that are not MCU multiples, qualities, chroma sampling factors, gray, optimised Huffman
tables and restart intervals, with the pixels cv2.imread decodes from them (tests/golden/jpeg/expected.npz).  Also records
the SHA-256 of cv2.imread's pixels for the reference's own three sample images (read from /root/reference when present;
the images themselves are not copied).  Run from the repo root in the build container (needs cv2)."""
import hashlib
import json
import os
import sys

import cv2
import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
OUT = os.path.join(HERE, "jpeg")
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
from oracle import oracle as orc  # noqa: E402  (synthetic image generator only)


def main():
    rng = np.random.default_rng(7)
    os.makedirs(OUT, exist_ok=True)
    smooth = cv2.GaussianBlur(rng.integers(0, 256, (97, 131, 3), dtype=np.uint8), (0, 0), 2)
    cases = [
        ("synth_61x47_q90_420", orc.synth_bgr(61, 47, 3), [cv2.IMWRITE_JPEG_QUALITY, 90]),
        ("synth_128x96_q75_420", orc.synth_bgr(128, 96, 4), [cv2.IMWRITE_JPEG_QUALITY, 75]),
        ("smooth_131x97_q95_444", smooth, [cv2.IMWRITE_JPEG_QUALITY, 95, cv2.IMWRITE_JPEG_SAMPLING_FACTOR, cv2.IMWRITE_JPEG_SAMPLING_FACTOR_444]),
        ("smooth_131x97_q60_422", smooth, [cv2.IMWRITE_JPEG_QUALITY, 60, cv2.IMWRITE_JPEG_SAMPLING_FACTOR, cv2.IMWRITE_JPEG_SAMPLING_FACTOR_422]),
        ("smooth_131x97_q85_440", smooth, [cv2.IMWRITE_JPEG_QUALITY, 85, cv2.IMWRITE_JPEG_SAMPLING_FACTOR, cv2.IMWRITE_JPEG_SAMPLING_FACTOR_440]),
        ("noise_33x17_q50_420", rng.integers(0, 256, (17, 33, 3), dtype=np.uint8), [cv2.IMWRITE_JPEG_QUALITY, 50]),
        ("synth_75x90_q80_rst", orc.synth_bgr(75, 90, 5), [cv2.IMWRITE_JPEG_QUALITY, 80, cv2.IMWRITE_JPEG_RST_INTERVAL, 3]),
        ("synth_75x90_q80_opt", orc.synth_bgr(75, 90, 5), [cv2.IMWRITE_JPEG_QUALITY, 80, cv2.IMWRITE_JPEG_OPTIMIZE, 1]),
        ("gray_50x41_q85", cv2.cvtColor(orc.synth_bgr(50, 41, 6), cv2.COLOR_BGR2GRAY), [cv2.IMWRITE_JPEG_QUALITY, 85]),
        ("flat_16x16_q90", np.full((16, 16, 3), (10, 200, 90), np.uint8), [cv2.IMWRITE_JPEG_QUALITY, 90]),
        ("synth_3x5_q90", orc.synth_bgr(3, 5, 8), [cv2.IMWRITE_JPEG_QUALITY, 90]),
    ]
    exp = {}
    for name, im, params in cases:
        path = os.path.join(OUT, name + ".jpg")
        assert cv2.imwrite(path, im, params)
        exp[name] = cv2.imread(path, cv2.IMREAD_COLOR)
    np.savez_compressed(os.path.join(OUT, "expected.npz"), **exp)
    ref = {}
    d = "/root/reference/src/main/resources/images"
    for f in ("album.jpg", "haha.jpg", "hkp.jpg"):
        p = os.path.join(d, f)
        if os.path.exists(p):
            im = cv2.imread(p, cv2.IMREAD_COLOR)
            ref[f] = {"shape": list(im.shape), "sha256": hashlib.sha256(np.ascontiguousarray(im).tobytes()).hexdigest()}
    json.dump(ref, open(os.path.join(OUT, "reference_images.json"), "w"), indent=1)
    print(sorted(os.listdir(OUT)), sum(os.path.getsize(os.path.join(OUT, f)) for f in os.listdir(OUT)), "bytes")


if __name__ == "__main__":
    main()
