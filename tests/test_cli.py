"""The C++ host program keeps the reference CLI contract (App.java:14-31) and, on a GPU, writes the output batch."""
import os
import subprocess

import numpy as np
import pytest

import msegment_b200 as mseg
from oracle import oracle as orc

CLI = os.path.join(mseg.PKG_DIR, "host", "msegment_cli")


def _write_ppm(path, bgr):
    h, w = bgr.shape[:2]
    with open(path, "wb") as f:
        f.write(b"P6\n%d %d\n255\n" % (w, h))
        f.write(np.ascontiguousarray(bgr[..., ::-1]).tobytes())


def _read_pnm(path):
    with open(path, "rb") as f:
        magic = f.readline().strip()
        w, h = map(int, f.readline().split())
        f.readline()
        data = np.frombuffer(f.read(), np.uint8)
    return data.reshape(h, w, 3)[..., ::-1] if magic == b"P6" else data.reshape(h, w)


def test_cli_argument_contract():
    assert os.path.exists(CLI), "run `python __graft_entry__.py` first"
    r = subprocess.run([CLI, "only", "two"], capture_output=True, text=True)
    assert r.returncode == 0 and r.stdout.strip() == "error parsing args"          # App.java:17-20
    r = subprocess.run([CLI], capture_output=True, text=True)
    assert r.returncode == 0 and "error parsing args" in r.stdout
    r = subprocess.run([CLI, "/nonexistent", "out", "x.ppm"], capture_output=True, text=True)
    assert r.stdout.splitlines()[:3] == ["arg 0: /nonexistent", "arg 1: out", "arg 2: x.ppm"]   # App.java:22-24
    assert "error with file stream processing" in r.stderr and r.returncode == 0   # IOException path: logged, no crash


def test_cli_fails_loudly_without_gpu(tmp_path):
    if mseg.lib.load().msg_device_count() > 0:
        pytest.skip("a GPU is present")
    _write_ppm(str(tmp_path / "in.ppm"), orc.synth_bgr(64, 48, 1))
    r = subprocess.run([CLI, str(tmp_path), "out", "in.ppm"], capture_output=True, text=True)
    assert r.returncode != 0 and "no CPU fallback" in r.stderr


@pytest.mark.gpu
def test_cli_batch_matches_oracle(tmp_path):
    im = orc.synth_bgr(200, 150, 9)
    _write_ppm(str(tmp_path / "input.ppm"), im)
    r = subprocess.run([CLI, str(tmp_path), "unused_out_root", "input.ppm"], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    outdir = tmp_path / "input_output"
    stamp = sorted(os.listdir(outdir))[0]
    files = sorted(os.listdir(outdir / stamp))
    assert files == ["MEANSHIFT_METHOD_input_00001_meanshift_filtered.ppm", "MEANSHIFT_METHOD_input_00002_markers.pgm",
                     "MEANSHIFT_METHOD_input_00003_merged_markers.pgm", "MEANSHIFT_METHOD_input_00004_result.ppm"]
    f = orc.meanshift_filter(im, 10, 10, 1)
    assert np.array_equal(_read_pnm(str(outdir / stamp / files[0])), f)
    n0, l0 = orc.label_regions(f, 2)
    n1, l1 = orc.merge_regions(f, l0, 50, 10)
    assert np.array_equal(_read_pnm(str(outdir / stamp / files[2])), np.clip(l1, 0, 255).astype(np.uint8))
    assert np.array_equal(_read_pnm(str(outdir / stamp / files[3])), orc.render_labels(l1, n1))
    assert "regions after merge: %d" % n1 in r.stdout
