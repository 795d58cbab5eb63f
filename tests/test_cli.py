"""The C++ host program keeps the reference CLI contract (App.java:14-31) and, on a GPU, writes the output batch."""
import os
import struct
import subprocess
import zlib

import numpy as np
import pytest

import msegment_b200 as mseg
from oracle import oracle as orc

CLI = os.path.join(mseg.PKG_DIR, "host", "msegment_cli")
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


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


def _read_png(path):
    """Independent PNG decoder (python zlib) for 8-bit gray / RGB, any filter type."""
    d = open(path, "rb").read()
    assert d[:8] == b"\x89PNG\r\n\x1a\n"
    o, idat = 8, b""
    while o < len(d):
        n, tag = struct.unpack(">I4s", d[o:o + 8])
        body = d[o + 8:o + 8 + n]
        assert zlib.crc32(d[o + 4:o + 8 + n]) == struct.unpack(">I", d[o + 8 + n:o + 12 + n])[0], "bad chunk CRC"
        if tag == b"IHDR":
            w, h, depth, ctype, _, _, interlace = struct.unpack(">IIBBBBB", body)
            assert depth == 8 and ctype in (0, 2) and interlace == 0
        elif tag == b"IDAT":
            idat += body
        o += 12 + n
    ch = 3 if ctype == 2 else 1
    raw = np.frombuffer(zlib.decompress(idat), np.uint8).reshape(h, w * ch + 1)
    assert not raw[:, 0].any()                       # our writer uses filter 0 only
    px = raw[:, 1:]
    return px.reshape(h, w, 3)[..., ::-1] if ch == 3 else px.reshape(h, w)


def _write_png_filtered(path, rgb, palette=None):
    """PNG encoder that cycles through all five filter types and compresses (dynamic Huffman blocks)."""
    h, w = rgb.shape[:2]
    ch = 1 if rgb.ndim == 2 else rgb.shape[2]
    rows = rgb.reshape(h, w * ch).astype(np.int32)
    out = bytearray()
    prev = np.zeros(w * ch, np.int32)
    for y in range(h):
        cur = rows[y]
        a = np.concatenate([np.zeros(ch, np.int32), cur[:-ch]])
        c = np.concatenate([np.zeros(ch, np.int32), prev[:-ch]])
        ft = y % 5
        if ft == 0:
            f = cur
        elif ft == 1:
            f = cur - a
        elif ft == 2:
            f = cur - prev
        elif ft == 3:
            f = cur - ((a + prev) >> 1)
        else:
            p = a + prev - c
            pa, pb, pc = abs(p - a), abs(p - prev), abs(p - c)
            pred = np.where((pa <= pb) & (pa <= pc), a, np.where(pb <= pc, prev, c))
            f = cur - pred
        out.append(ft)
        out += (f & 255).astype(np.uint8).tobytes()
        prev = cur

    def chunk(tag, body):
        return struct.pack(">I", len(body)) + tag + body + struct.pack(">I", zlib.crc32(tag + body))
    ctype = 3 if palette is not None else {1: 0, 2: 4, 3: 2, 4: 6}[ch]
    data = b"\x89PNG\r\n\x1a\n" + chunk(b"IHDR", struct.pack(">IIBBBBB", w, h, 8, ctype, 0, 0, 0))
    if palette is not None:
        data += chunk(b"PLTE", palette.astype(np.uint8).tobytes())
    data += chunk(b"IDAT", zlib.compress(bytes(out), 9)) + chunk(b"IEND", b"")
    open(path, "wb").write(data)


@pytest.mark.parametrize("kind", ["rgb", "rgba", "gray", "palette", "noise_stored"])
def test_png_codec_roundtrip(tmp_path, kind):
    """host/png_io.hpp: inflate + unfilter + palette on read, stored-deflate on write, checked against python zlib."""
    rt = os.path.join(mseg.PKG_DIR, "host", "png_roundtrip")
    assert os.path.exists(rt), "run `python __graft_entry__.py` first"
    rng = np.random.default_rng(5)
    bgr = orc.synth_bgr(300, 231, 4)                 # > 65535 raw bytes: several stored blocks on write
    rgb = bgr[..., ::-1]
    src = str(tmp_path / "in.png")
    if kind == "rgb":
        _write_png_filtered(src, rgb)
        want = bgr
    elif kind == "rgba":
        _write_png_filtered(src, np.dstack([rgb, rng.integers(0, 256, rgb.shape[:2], dtype=np.uint8)]))
        want = bgr
    elif kind == "gray":
        _write_png_filtered(src, rgb[..., 0])
        want = np.repeat(rgb[..., :1], 3, axis=2)
    elif kind == "palette":
        pal = rng.integers(0, 256, (200, 3), dtype=np.uint8)
        idx = (rgb[..., 0].astype(np.int32) * 199 // 255).astype(np.uint8)
        _write_png_filtered(src, idx, palette=pal)
        want = pal[idx][..., ::-1]
    else:
        noise = rng.integers(0, 256, (97, 61, 3), dtype=np.uint8)     # incompressible: zlib emits stored blocks
        _write_png_filtered(src, noise)
        want = noise[..., ::-1]
    dst = str(tmp_path / "out.png")
    r = subprocess.run([rt, src, dst])
    assert r.returncode == 0
    assert np.array_equal(_read_png(dst), want)
    assert subprocess.run([rt, str(tmp_path / "missing.png"), dst]).returncode == 2


def _decode_with_cli_reader(path, tmp_path):
    rt = os.path.join(mseg.PKG_DIR, "host", "png_roundtrip")
    dst = str(tmp_path / "decoded.png")
    r = subprocess.run([rt, path, dst])
    return r.returncode, (_read_png(dst) if r.returncode == 0 else None)


def test_jpeg_reader_matches_imread_fixtures(tmp_path):
    """host/jpeg_io.hpp against cv2.imread (libjpeg-turbo) on the committed baseline fixtures: sizes that are not MCU multiples,
    4:2:0 / 4:2:2 / 4:4:0 / 4:4:4, gray, optimised Huffman tables, restart intervals (tests/golden/gen_jpeg.py)."""
    d = os.path.join(ROOT, "tests", "golden", "jpeg")
    exp = np.load(os.path.join(d, "expected.npz"))
    assert len(exp.files) >= 10
    for name in exp.files:
        rc, got = _decode_with_cli_reader(os.path.join(d, name + ".jpg"), tmp_path)
        assert rc == 0, name
        assert np.array_equal(got, exp[name]), name
    rc, _ = _decode_with_cli_reader(os.path.join(d, "expected.npz"), tmp_path)      # not an image: unreadable, no crash
    assert rc == 2


def test_jpeg_reader_on_the_reference_images(tmp_path):
    """The reference's own sample inputs (baseline 4:2:0 JPEGs) decode to the pixels cv2.imread gives (hashes recorded by
    tests/golden/gen_jpeg.py).  Only where /root/reference exists (the build container)."""
    import hashlib
    import json
    d = "/root/reference/src/main/resources/images"
    if not os.path.isdir(d):
        pytest.skip("/root/reference is not available here")
    want = json.load(open(os.path.join(ROOT, "tests", "golden", "jpeg", "reference_images.json")))
    assert set(want) == {"album.jpg", "haha.jpg", "hkp.jpg"}
    for f, meta in want.items():
        rc, got = _decode_with_cli_reader(os.path.join(d, f), tmp_path)
        assert rc == 0 and list(got.shape) == meta["shape"], f
        assert hashlib.sha256(np.ascontiguousarray(got).tobytes()).hexdigest() == meta["sha256"], f


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
    # same naming scheme and file type as PictureService.saveResultsToFS (PictureService.java:209-216)
    ms = [f for f in files if f.startswith("MEANSHIFT_METHOD_")]
    assert ms == ["MEANSHIFT_METHOD_input_00001_meanshift_filtered.png", "MEANSHIFT_METHOD_input_00002_markers.png",
                  "MEANSHIFT_METHOD_input_00003_merged_markers.png", "MEANSHIFT_METHOD_input_00004_result.png"]
    # the reference's two pipelines, all 8 + 8 Results under its own step names (PictureService.java:301-382, :396-467;
    # SURVEY App. C#4)
    assert [f for f in files if f.startswith("COLOR_METHOD_")] == [
        "COLOR_METHOD_input_00001_black_bg.png", "COLOR_METHOD_input_00002_laplassian_sharp.png", "COLOR_METHOD_input_00003_bw.png",
        "COLOR_METHOD_input_00004_distance_transform.png", "COLOR_METHOD_input_00005_distance_peaks.png",
        "COLOR_METHOD_input_00006_markers.png", "COLOR_METHOD_input_00007_result.png", "COLOR_METHOD_input_00008_bw_result.png"]
    assert [f for f in files if f.startswith("SHAPE_METHOD_")] == [
        "SHAPE_METHOD_input_00001_blured_by_%dx%d.png" % ((orc.blur_mask_size(200, 150),) * 2), "SHAPE_METHOD_input_00002_borders.png",
        "SHAPE_METHOD_input_00003_gray_borders.png",
        "SHAPE_METHOD_input_00004_dde_step.png", "SHAPE_METHOD_input_00005_dde_step_blurred_3x3.png",
        "SHAPE_METHOD_input_00006_markers.png", "SHAPE_METHOD_input_00007_result.png", "SHAPE_METHOD_input_00008_bw_result.png"]
    cn, cm, cst = orc.color_seeds(im)
    rd = lambda name: _read_png(str(outdir / stamp / name))
    assert np.array_equal(rd("COLOR_METHOD_input_00002_laplassian_sharp.png"), cst["sharp"])
    assert np.array_equal(rd("COLOR_METHOD_input_00003_bw.png"), cst["bw"])
    assert np.array_equal(rd("COLOR_METHOD_input_00004_distance_transform.png"),
                          np.clip(np.rint(cst["norm"].astype(np.float64) * 1000), 0, 255).astype(np.uint8))
    assert np.array_equal(rd("COLOR_METHOD_input_00005_distance_peaks.png"), cst["peaks"] * 255)
    assert np.array_equal(rd("COLOR_METHOD_input_00006_markers.png"), np.clip(cm.astype(np.int64) * 10000, 0, 255).astype(np.uint8))
    assert "colour-method contours: %d" % cn in r.stdout
    # result / bw_result: Imgproc.watershed on the SHARPENED image (the reference copies it into src, :333) + colorByIndexes
    ws = orc.watershed(cst["sharp"], cm.copy())
    res = orc.render_labels(ws, cn)
    assert np.array_equal(rd("COLOR_METHOD_input_00007_result.png"), res)
    assert np.array_equal(rd("COLOR_METHOD_input_00008_bw_result.png"), orc.bgr2gray(res))
    sn, sm, sst = orc.shape_seeds(im)
    want_borders = np.where((sst["edges"] != 0)[..., None], im, 0).astype(np.uint8)
    assert np.array_equal(rd("SHAPE_METHOD_input_00002_borders.png"), want_borders)
    sdepth, _ = orc.contour_markers(sst["dde3"])                     # contours.size() incl. holes (:450-455)
    ws = orc.watershed(im, sm.copy())
    res = orc.render_labels(ws, sdepth)
    assert np.array_equal(rd("SHAPE_METHOD_input_00007_result.png"), res)
    assert np.array_equal(rd("SHAPE_METHOD_input_00008_bw_result.png"), orc.bgr2gray(res))
    assert np.array_equal(rd("SHAPE_METHOD_input_00003_gray_borders.png"), sst["edges"])
    assert np.array_equal(rd("SHAPE_METHOD_input_00005_dde_step_blurred_3x3.png"), sst["dde3"])
    assert np.array_equal(rd("SHAPE_METHOD_input_00006_markers.png"), np.clip(sm.astype(np.int64) * 10000, 0, 255).astype(np.uint8))
    assert "shape-method labels: %d" % sn in r.stdout
    files = ms
    f = orc.meanshift_filter(im, 10, 10, 1)
    assert np.array_equal(_read_png(str(outdir / stamp / files[0])), f)
    n0, l0 = orc.label_regions(f, 2)
    n1, l1 = orc.merge_regions(f, l0, 50, 10)
    assert np.array_equal(_read_png(str(outdir / stamp / files[2])), np.clip(l1, 0, 255).astype(np.uint8))
    assert np.array_equal(_read_png(str(outdir / stamp / files[3])), orc.render_labels(l1, n1))
    assert "regions after merge: %d" % n1 in r.stdout
    # PNG input + PNM output switch
    _write_png_filtered(str(tmp_path / "second.png"), im[..., ::-1])
    r = subprocess.run([CLI, str(tmp_path), "unused_out_root", "second.png"], capture_output=True, text=True,
                       env=dict(os.environ, MSG_OUT_FORMAT="pnm"))
    assert r.returncode == 0, r.stderr
    outdir = tmp_path / "second_output"
    stamp = sorted(os.listdir(outdir))[0]
    assert np.array_equal(_read_pnm(str(outdir / stamp / "MEANSHIFT_METHOD_second_00001_meanshift_filtered.ppm")), f)
    # JPEG input (the reference's sample images are baseline JPEGs): decoded as imread decodes it, then the same pipeline
    import shutil
    jdir = os.path.join(ROOT, "tests", "golden", "jpeg")
    shutil.copy(os.path.join(jdir, "synth_128x96_q75_420.jpg"), str(tmp_path / "third.jpg"))
    r = subprocess.run([CLI, str(tmp_path), "unused_out_root", "third.jpg"], capture_output=True, text=True,
                       env=dict(os.environ, MSG_OUT_FORMAT="pnm", MSG_REFERENCE_MARKERS="0"))
    assert r.returncode == 0, r.stderr
    decoded = np.load(os.path.join(jdir, "expected.npz"))["synth_128x96_q75_420"]
    outdir = tmp_path / "third_output"
    stamp = sorted(os.listdir(outdir))[0]
    assert np.array_equal(_read_pnm(str(outdir / stamp / "MEANSHIFT_METHOD_third_00001_meanshift_filtered.ppm")),
                          orc.meanshift_filter(decoded, 10, 10, 1))
