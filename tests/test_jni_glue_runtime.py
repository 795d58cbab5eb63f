"""The JNI glue (opencv-msegment_b200/java/msegment_jni.c) EXECUTED, not only type-checked: it is compiled against
tests/stubs/jni.h into a shared library and driven from Python through a mock JNIEnv whose function table has exactly the
layout of that stub header (NewStringUTF, Get/ReleaseStringUTFChars, Get/ReleaseByteArrayElements, SetIntArrayRegion,
SetDoubleArrayRegion).  Java arrays / strings are plain buffers whose address plays the role of the jobject.  Every native
method of GpuImgproc.java is called once and its result is compared with the same operator called through the C ABI directly
(the ctypes mirror, itself parity-tested against the oracle): an argument forwarded in the wrong order or with the wrong type
shows up here.  (No JDK in this image: the Java class itself still cannot be compiled)"""
import ctypes as C
import os
import re
import subprocess

import numpy as np
import pytest

import msegment_b200 as mseg

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GLUE = os.path.join(ROOT, "opencv-msegment_b200", "java", "msegment_jni.c")
PFX = "Java_ru_shayhulud_opencvcmsegment_gpu_GpuImgproc_"


class _Table(C.Structure):          # tests/stubs/jni.h: struct JNINativeInterface_, in that order
    _fields_ = [("NewStringUTF", C.c_void_p), ("GetStringUTFChars", C.c_void_p), ("ReleaseStringUTFChars", C.c_void_p),
                ("GetByteArrayElements", C.c_void_p), ("ReleaseByteArrayElements", C.c_void_p),
                ("SetIntArrayRegion", C.c_void_p), ("SetDoubleArrayRegion", C.c_void_p)]


class MockEnv:
    """JNIEnv* = pointer to a pointer to the function table."""

    def __init__(self):
        self.strings = []
        self.released = []
        V, I = C.c_void_p, C.c_int32
        self._cbs = [
            C.CFUNCTYPE(V, V, C.c_char_p)(self._new_string),
            C.CFUNCTYPE(V, V, V, V)(lambda env, s, is_copy: s),                       # a jstring IS its char buffer here
            C.CFUNCTYPE(None, V, V, V)(lambda env, s, chars: self.released.append(("str", s, chars))),
            C.CFUNCTYPE(V, V, V, V)(lambda env, arr, is_copy: arr),                   # a jbyteArray IS its element buffer
            C.CFUNCTYPE(None, V, V, V, I)(lambda env, arr, elems, mode: self.released.append(("bytes", arr, elems, mode))),
            C.CFUNCTYPE(None, V, V, I, I, V)(lambda env, arr, start, n, buf: C.memmove(arr + 4 * start, buf, 4 * n)),
            C.CFUNCTYPE(None, V, V, I, I, V)(lambda env, arr, start, n, buf: C.memmove(arr + 8 * start, buf, 8 * n)),
        ]
        self.table = _Table(*[C.cast(cb, C.c_void_p) for cb in self._cbs])
        self.inner = C.pointer(self.table)
        self.env = C.pointer(self.inner)

    def _new_string(self, env, utf):
        buf = C.create_string_buffer(utf if utf is not None else b"")
        self.strings.append(buf)
        return C.addressof(buf)


@pytest.fixture(scope="module")
def glue(tmp_path_factory):
    out = str(tmp_path_factory.mktemp("jni") / "libmsegment_jni_stub.so")
    libdir = os.path.dirname(mseg.lib.LIB_PATH)
    r = subprocess.run(["gcc", "-shared", "-fPIC", "-O1", "-std=c99", "-Wall", "-Wextra", "-Wno-unused-parameter", "-Werror",
                        "-I", os.path.join(ROOT, "tests", "stubs"), "-I", os.path.join(ROOT, "include"), GLUE,
                        "-L", libdir, "-lmsegment_b200", "-Wl,-rpath," + libdir, "-o", out], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    return C.CDLL(out)


def _natives():
    java = open(os.path.join(ROOT, "opencv-msegment_b200", "java", "GpuImgproc.java")).read()
    return sorted(set(re.findall(r"private static native \w+(?:\[\])? (\w+)\(", java)))


def test_glue_builds_and_exports_every_native(glue):
    """CPU: the glue links against the real library and exports one JNI function per native method; without a GPU nCreate
    returns 0 and nLastError hands the library's message to NewStringUTF."""
    names = _natives()
    assert len(names) >= 31
    for n in names:
        assert hasattr(glue, PFX + n), n
    env = MockEnv()
    import torch
    if not torch.cuda.is_available():
        f = getattr(glue, PFX + "nCreate")
        f.restype, f.argtypes = C.c_int64, [C.c_void_p, C.c_void_p, C.c_int32]
        assert f(env.env, None, 0) == 0
        g = getattr(glue, PFX + "nLastError")
        g.restype, g.argtypes = C.c_void_p, [C.c_void_p, C.c_void_p, C.c_int64]
        s = g(env.env, None, 0)
        assert s and b"CUDA" in C.string_at(s)


def _fn(glue, name, restype, *argtypes):
    f = getattr(glue, PFX + name)
    f.restype = restype
    f.argtypes = [C.c_void_p, C.c_void_p] + list(argtypes)
    return f


@pytest.mark.gpu
def test_every_native_forwards_its_arguments(glue):
    from oracle import oracle as orc
    J, I, D, V = C.c_int64, C.c_int32, C.c_double, C.c_void_p
    env = MockEnv()
    e = env.env
    called = set()

    def call(name, restype, argtypes, *args):
        called.add(name)
        return _fn(glue, name, restype, *argtypes)(e, None, *args)

    ctx = call("nCreate", J, [I], 0)
    assert ctx != 0
    with mseg.Context(0) as pctx:
        gi = mseg.GpuImgproc(pctx)
        w, h = 157, 93                                           # odd sizes, and steps larger than the rows
        im = orc.synth_bgr(w, h, 5)
        pad = np.zeros((h, w + 7, 3), np.uint8)
        pad[:, :w] = im
        src = pad[:, :w]                                         # step = 3 * (w + 7)
        A = lambda a: a.ctypes.data                              # noqa: E731
        S = lambda a: a.strides[0]                               # noqa: E731
        IMG = [J, J, J, J, J, I, I]                              # ctx, src, sstep, dst, dstep, w, h

        # pyrMeanShiftFiltering
        dst = np.zeros((h, w + 3, 3), np.uint8)[:, :w]
        assert call("nMeanshift", I, IMG + [D, D, I, I, I, D], ctx, A(src), S(src), A(dst), S(dst), w, h, 7.0, 11.0, 1, 3, 4, 1.0) == 0
        filt = gi.pyrMeanShiftFiltering(im, 7, 11, 1, (3, 4, 1.0))
        assert np.array_equal(dst, filt)
        # labelRegions / mergeRegions / connectedComponents / render
        n = np.zeros(1, np.int32)
        lab = np.zeros((h, w + 5), np.int32)[:, :w]
        fc = np.ascontiguousarray(filt)
        assert call("nLabelRegions", I, [J, J, J, J, J, I, I, I, I, I, V], ctx, A(fc), S(fc), A(lab), S(lab), w, h, 3, 3, 8, A(n)) == 0
        wn, wlab = gi.labelRegions(fc, 3, 3, 8)
        assert int(n[0]) == wn and np.array_equal(lab, wlab)
        lab2 = np.ascontiguousarray(lab)
        assert call("nMergeRegions", I, [J, J, J, J, J, I, I, I, I, V], ctx, A(fc), S(fc), A(lab2), S(lab2), w, h, 30, 9, A(n)) == 0
        mn, mlab = gi.mergeRegions(fc, wlab, 30, 9)
        assert int(n[0]) == mn and np.array_equal(lab2, mlab)
        mask = ((im[..., 1] > 128) * 255).astype(np.uint8)
        cc = np.zeros((h, w), np.int32)
        assert call("nConnectedComponents", I, [J, J, J, J, J, I, I, I, V], ctx, A(mask), S(mask), A(cc), S(cc), w, h, 4, A(n)) == 0
        cn, clab = gi.connectedComponents(mask, 4)
        assert int(n[0]) == cn and np.array_equal(cc, clab)
        colors = np.random.default_rng(1).integers(0, 256, (mn, 3), dtype=np.uint8)
        ren = np.zeros((h, w, 3), np.uint8)
        assert call("nRender", I, [J, J, J, J, J, I, I, I, V], ctx, A(lab2), S(lab2), A(ren), S(ren), w, h, mn, A(colors)) == 0
        assert np.array_equal(ren, gi.colorByIndexes(mlab, mn, colors))
        assert call("nRender", I, [J, J, J, J, J, I, I, I, V], ctx, A(lab2), S(lab2), A(ren), S(ren), w, h, mn, None) == 0
        assert np.array_equal(ren, gi.colorByIndexes(mlab, mn, None))
        assert any(r[0] == "bytes" and r[3] == 2 for r in env.released)             # ReleaseByteArrayElements(JNI_ABORT)
        # pre-filters
        taps = np.array([1, 1, 1, 1, -8, 1, 1, 1, 1], np.int8)
        o3 = np.zeros((h, w, 3), np.uint8)
        assert call("nSharpen", I, IMG + [V, I, I], ctx, A(src), S(src), A(o3), S(o3), w, h, A(taps), 9, 1) == 0
        assert np.array_equal(o3, gi.sharpenLaplacian(im, taps.reshape(9, 1)))
        gray = np.zeros((h, w), np.uint8)
        assert call("nGray", I, IMG, ctx, A(src), S(src), A(gray), S(gray), w, h) == 0
        assert np.array_equal(gray, gi.cvtColorBGR2GRAY(im))
        o1 = np.zeros((h, w), np.uint8)
        assert call("nMedian", I, IMG + [I], ctx, A(gray), S(gray), A(o1), S(o1), w, h, 5) == 0
        assert np.array_equal(o1, gi.medianBlur(gray, 5))
        edges = np.zeros((h, w), np.uint8)
        assert call("nCanny", I, IMG + [D, D], ctx, A(gray), S(gray), A(edges), S(edges), w, h, 5.0, 50.0) == 0
        assert np.array_equal(edges, gi.Canny(gray, 5, 50))
        dil = np.zeros((h, w), np.uint8)
        assert call("nDilate", I, IMG + [I, I], ctx, A(edges), S(edges), A(dil), S(dil), w, h, 5, 3) == 0
        assert np.array_equal(dil, gi.dilate(edges, (3, 5)))                          # (rows, cols) = (kh, kw)
        sub = np.zeros((h, w), np.uint8)
        assert call("nSubtract", I, [J, J, J, J, J, J, J, I, I], ctx, A(dil), S(dil), A(edges), S(edges), A(sub), S(sub), w, h) == 0
        assert np.array_equal(sub, gi.subtract(dil, edges))
        mk = np.zeros((h, w), np.int32)
        assert call("nShapeSeeds", I, [J, J, J, I, I, I, D, D, J, J, V], ctx, A(src), S(src), w, h, 3, 5.0, 50.0, A(mk), S(mk), A(n)) == 0
        sn, smk = gi.shapeSeeds(im, 5, 10, medianKsize=3)
        assert int(n[0]) == sn and np.array_equal(mk, smk)
        cm = np.zeros((h, w, 3), np.uint8)
        assert call("nCopyMasked", I, [J, J, J, J, J, J, J, I, I], ctx, A(src), S(src), A(edges), S(edges), A(cm), S(cm), w, h) == 0
        assert np.array_equal(cm, np.where(edges[..., None] != 0, im, 0).astype(np.uint8))     # src.copyTo(zeros, mask)
        # colour-method stages
        white = im.copy()
        white[10:20, 30:60] = 255
        wb = np.zeros((h, w, 3), np.uint8)
        assert call("nWhiteToBlack", I, IMG, ctx, A(white), S(white), A(wb), S(wb), w, h) == 0
        assert np.array_equal(wb, gi.whiteToBlack(white))
        used = np.zeros(1, np.float64)
        bw = np.zeros((h, w), np.uint8)
        assert call("nThreshold", I, IMG + [D, D, I, V], ctx, A(gray), S(gray), A(bw), S(bw), w, h, 40.0, 255.0, 8, A(used)) == 0
        t, wbw = gi.threshold(gray, 40, 255, 8)
        assert used[0] == t and np.array_equal(bw, wbw)
        dist = np.zeros((h, w), np.float32)
        assert call("nDistanceTransform", I, IMG + [I, I], ctx, A(bw), S(bw), A(dist), S(dist), w, h, 2, 5) == 0
        wdist = gi.distanceTransform(bw)
        assert np.array_equal(dist, wdist)
        # options travel through GetStringUTFChars / ReleaseStringUTFChars
        name = C.create_string_buffer(b"dt_fixed")
        assert call("nSetOption", I, [J, V, I], ctx, C.addressof(name), 1) == 0
        assert call("nDistanceTransform", I, IMG + [I, I], ctx, A(bw), S(bw), A(dist), S(dist), w, h, 2, 5) == 0
        assert np.array_equal(dist, orc.distance_transform(bw, fixed=True)) and not np.array_equal(dist, wdist)
        assert call("nSetOption", I, [J, V, I], ctx, C.addressof(name), 0) == 0
        bad = C.create_string_buffer(b"no_such_option")
        assert call("nSetOption", I, [J, V, I], ctx, C.addressof(bad), 1) != 0
        assert any(r[0] == "str" for r in env.released)
        s = call("nLastError", V, [J], ctx)
        assert b"no_such_option" in C.string_at(s) or b"option" in C.string_at(s)
        nrm = np.zeros((h, w), np.float32)
        assert call("nNormalize", I, IMG + [D, D], ctx, A(wdist), S(wdist), A(nrm), S(nrm), w, h, 0.0, 1.0) == 0
        assert np.array_equal(nrm, gi.normalize(wdist, 0, 1))
        thr = np.zeros((h, w), np.float32)
        assert call("nThresholdF32", I, IMG + [D, D], ctx, A(nrm), S(nrm), A(thr), S(thr), w, h, 0.4, 1.0) == 0
        assert np.array_equal(thr, gi.threshold(nrm, 0.4, 1.0, 0)[1])
        dlf = np.zeros((h, w), np.float32)
        assert call("nDilateF32", I, IMG + [I, I], ctx, A(thr), S(thr), A(dlf), S(dlf), w, h, 3, 3) == 0
        assert np.array_equal(dlf, gi.dilateF32(thr, (3, 3)))
        pk = np.zeros((h, w), np.uint8)
        assert call("nConvertU8", I, IMG, ctx, A(dlf), S(dlf), A(pk), S(pk), w, h) == 0
        assert np.array_equal(pk, gi.convertToU8(dlf))
        cmk = np.zeros((h, w), np.int32)
        assert call("nContourMarkers", I, [J, J, J, J, J, I, I, V], ctx, A(pk), S(pk), A(cmk), S(cmk), w, h, A(n)) == 0
        kn, kmk = gi.contourMarkers(pk)
        assert int(n[0]) == kn and np.array_equal(cmk, kmk)
        assert call("nCircle", I, [J, J, J, I, I, I, I, I, I], ctx, A(cmk), S(cmk), w, h, 5, 5, 3, 255) == 0
        assert np.array_equal(cmk, gi.circle(kmk.copy(), (5, 5), 3, 255))
        fmk = np.zeros((h, w), np.int32)
        assert call("nColorSeeds", I, [J, J, J, I, I, V, I, I, D, J, J, V], ctx, A(src), S(src), w, h, A(taps), 9, 1, 0.4, A(fmk), S(fmk),
                    A(n)) == 0
        fn_, fm = gi.colorSeeds(im)
        assert int(n[0]) == fn_ and np.array_equal(fmk, fm)
        bl = np.zeros((h, w), np.uint8)
        assert call("nBilateral", I, IMG + [I, I, D, D], ctx, A(gray), S(gray), A(bl), S(bl), w, h, 1, 5, 10.0, 10.0) == 0
        assert np.array_equal(bl, gi.bilateralFilter(gray, 5, 10, 10))
        # watershed on the colour-method markers
        wsm = fm.copy()
        assert call("nWatershed", I, [J, J, J, J, J, I, I], ctx, A(src), S(src), A(wsm), S(wsm), w, h) == 0
        assert np.array_equal(wsm, gi.watershed(im, fm.copy()))
        # fused segment: 32-bit and 16-bit labels, filtered on and off
        f2 = np.zeros((h, w, 3), np.uint8)
        l32 = np.zeros((h, w), np.int32)
        assert call("nSegment", I, [J, J, J, I, I, D, D, I, I, I, I, I, J, J, J, J, V], ctx, A(src), S(src), w, h, 7.0, 11.0, 1, 3, 30, 9, 0,
                    A(f2), S(f2), A(l32), S(l32), A(n)) == 0
        want = gi.segment(im, 7, 11, 1, loDiff=3, minSize=30, colorDist=9)
        assert np.array_equal(l32, want["labels"]) and int(n[0]) == want["n_regions"]
        assert np.array_equal(f2, gi.pyrMeanShiftFiltering(im, 7, 11, 1))
        l16 = np.zeros((h, w), np.uint16)
        assert call("nSegment", I, [J, J, J, I, I, D, D, I, I, I, I, I, J, J, J, J, V], ctx, A(src), S(src), w, h, 7.0, 11.0, 1, 3, 30, 9, 1,
                    0, 0, A(l16), S(l16), A(n)) == 0
        assert np.array_equal(l16.astype(np.int32), want["labels"])
        # host registration
        big = np.zeros(1 << 20, np.uint8)
        assert call("nRegisterHost", I, [J, J, J], ctx, A(big), big.nbytes) == 0
        assert call("nUnregisterHost", I, [J, J], ctx, A(big)) == 0
        assert call("nUnregisterHost", I, [J, J], ctx, A(big)) != 0
    assert called >= set(_natives()) - {"nDestroy"}, sorted(set(_natives()) - called)
    if hasattr(glue, PFX + "nDestroy"):
        _fn(glue, "nDestroy", None, J)(e, None, ctx)
