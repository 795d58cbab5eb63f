#!/usr/bin/env python
"""One-off campaign (not part of the test suite: ~25 CPU-minutes): the ORACLE reproduces every unit of
tests/golden/fullsize_digests.json (cv2 on whole 4K / 1080p frames) -- filtered frame SHA-256 and label map SHA-256.
Result of the last run: tests/golden/PROVENANCE.txt."""
import hashlib
import json
import multiprocessing as mp
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))


def unit(args):
    from oracle import oracle as orc
    name, u = args
    im = orc.synth_bgr(u["w"], u["h"], u["seed"])
    f = orc.meanshift_filter(im, u["sp"], u["sr"], 1)
    n, lab = orc.label_regions(f, 2)
    ok_f = hashlib.sha256(np.ascontiguousarray(f).tobytes()).hexdigest() == u["filtered_sha256"]
    ok_l = n == u["n_regions"] and hashlib.sha256(np.ascontiguousarray(lab).tobytes()).hexdigest() == u["labels_sha256"]
    return name, ok_f, ok_l


def main():
    d = json.load(open(os.path.join(HERE, "fullsize_digests.json")))
    units = sorted(d["units"].items(), key=lambda kv: -kv[1]["sp"] * kv[1]["w"])
    with mp.get_context("spawn").Pool(max(1, (os.cpu_count() or 2) - 1)) as pool:
        res = pool.map(unit, units, chunksize=1)
    bad = [r for r in res if not (r[1] and r[2])]
    print("oracle == cv2 digests on %d / %d units" % (len(res) - len(bad), len(res)), bad)


if __name__ == "__main__":
    main()
