"""ctypes loader for the CPU oracle (TEST INFRASTRUCTURE ONLY -- see oracle/msg_oracle.h).

Importable only from tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
``--impl reference`` legs.  Nothing under ``opencv-msegment_b200/`` may import this module.
"""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "libmsg_oracle.so")

TERM_COUNT = 1
TERM_EPS = 2


class MsCounters(C.Structure):
    _fields_ = [("window_tests", C.c_uint64), ("hits", C.c_uint64), ("iterations", C.c_uint64),
                ("pixels", C.c_uint64), ("max_drift", C.c_uint64)]


def build(force=False):
    src = os.path.join(_HERE, "msg_oracle.c")
    if (force or not os.path.exists(_LIB_PATH)
            or os.path.getmtime(_LIB_PATH) < max(os.path.getmtime(src), os.path.getmtime(src[:-2] + ".h"))):
        subprocess.check_call(["make", "-C", _HERE, "-s", "-B", "libmsg_oracle.so"])
    return _LIB_PATH


_lib = None


def lib():
    global _lib
    if _lib is None:
        build()
        L = C.CDLL(_LIB_PATH)
        u8p, i32p, sz, i, d = C.c_void_p, C.c_void_p, C.c_size_t, C.c_int, C.c_double
        L.orc_pyr_down_8uc3.argtypes = [u8p, sz, i, i, u8p, sz]
        L.orc_pyr_up_8uc3.argtypes = [u8p, sz, i, i, u8p, sz, i, i]
        L.orc_meanshift_filter.argtypes = [u8p, sz, u8p, sz, i, i, d, d, i, i, i, d, C.POINTER(MsCounters)]
        L.orc_meanshift_filter.restype = i
        L.orc_meanshift_filter_roi.argtypes = [u8p, sz, u8p, sz, i, i, i, i, i, i, d, d, i, i, i, d, C.POINTER(MsCounters)]
        L.orc_meanshift_filter_roi.restype = i
        L.orc_label_regions.argtypes = [u8p, sz, i32p, sz, i, i, i]
        L.orc_label_regions.restype = C.c_int32
        L.orc_label_regions_conn.argtypes = [u8p, sz, i32p, sz, i, i, i, i]
        L.orc_label_regions_conn.restype = C.c_int32
        L.orc_connected_components.argtypes = [u8p, sz, i32p, sz, i, i, i]
        L.orc_connected_components.restype = C.c_int32
        L.orc_relabel_canonical.argtypes = [i32p, sz, i, i]
        L.orc_relabel_canonical.restype = C.c_int32
        L.orc_merge_regions.argtypes = [u8p, sz, i32p, sz, i, i, i, i]
        L.orc_merge_regions.restype = C.c_int32
        L.orc_render_labels.argtypes = [i32p, sz, u8p, sz, i, i, i, u8p]
        L.orc_watershed.argtypes = [u8p, sz, i32p, sz, i, i]
        L.orc_synth_bgr.argtypes = [u8p, sz, i, i, C.c_uint64]
        L.orc_laplacian_sharpen.argtypes = [u8p, sz, u8p, sz, i, i, C.c_void_p, i, i]
        L.orc_median_blur_8uc1.argtypes = [u8p, sz, u8p, sz, i, i, i]
        L.orc_bgr2gray.argtypes = [u8p, sz, u8p, sz, i, i]
        L.orc_bgr2gray_342.argtypes = [u8p, sz, u8p, sz, i, i]
        L.orc_canny.argtypes = [u8p, sz, u8p, sz, i, i, d, d]
        L.orc_dilate_rect.argtypes = [u8p, sz, u8p, sz, i, i, i, i]
        L.orc_subtract_u8.argtypes = [u8p, sz, u8p, sz, u8p, sz, i, i]
        L.orc_otsu_threshold.argtypes = [u8p, sz, i, i]
        L.orc_otsu_threshold.restype = i
        L.orc_threshold_binary_u8.argtypes = [u8p, sz, u8p, sz, i, i, i, i]
        L.orc_distance_transform_l2_5.argtypes = [u8p, sz, C.c_void_p, sz, i, i]
        L.orc_distance_transform_l2_5_fixed.argtypes = [u8p, sz, C.c_void_p, sz, i, i]
        L.orc_normalize_minmax01_f32.argtypes = [C.c_void_p, sz, C.c_void_p, sz, i, i]
        L.orc_peaks_u8.argtypes = [C.c_void_p, sz, u8p, sz, i, i, d]
        L.orc_contour_markers.argtypes = [u8p, sz, i32p, sz, i, i]
        L.orc_contour_markers.restype = C.c_int32
        L.orc_circle_filled_i32.argtypes = [i32p, sz, i, i, i, i, i, C.c_int32]
        L.orc_white_to_black.argtypes = [u8p, sz, u8p, sz, i, i]
        L.orc_bilateral_filter.argtypes = [u8p, sz, u8p, sz, i, i, i, i, d, d]
        _lib = L
    return _lib


def _img(a):
    a = np.ascontiguousarray(a, dtype=np.uint8)
    assert a.ndim == 3 and a.shape[2] == 3
    return a


def pyr_down(src):
    src = _img(src)
    h, w = src.shape[:2]
    dst = np.empty(((h + 1) // 2, (w + 1) // 2, 3), np.uint8)
    lib().orc_pyr_down_8uc3(src.ctypes.data, src.strides[0], w, h, dst.ctypes.data, dst.strides[0])
    return dst


def pyr_up(src, dsize=None):
    src = _img(src)
    h, w = src.shape[:2]
    dw, dh = dsize if dsize is not None else (2 * w, 2 * h)
    dst = np.empty((dh, dw, 3), np.uint8)
    lib().orc_pyr_up_8uc3(src.ctypes.data, src.strides[0], w, h, dst.ctypes.data, dst.strides[0], dw, dh)
    return dst


def meanshift_filter(src, sp, sr, max_level=1, term=(TERM_COUNT | TERM_EPS, 5, 1.0), counters=False):
    src = _img(src)
    h, w = src.shape[:2]
    dst = np.empty_like(src)
    ct = MsCounters()
    rc = lib().orc_meanshift_filter(src.ctypes.data, src.strides[0], dst.ctypes.data, dst.strides[0], w, h,
                                    float(sp), float(sr), int(max_level), int(term[0]), int(term[1]),
                                    float(term[2]), C.byref(ct))
    if rc != 0:
        raise ValueError("orc_meanshift_filter: invalid arguments")
    if counters:
        return dst, {k: getattr(ct, k) for k, _ in MsCounters._fields_}
    return dst


def meanshift_filter_roi(crop, xoff, yoff, full_w, full_h, sp, sr, max_level=1, term=(TERM_COUNT | TERM_EPS, 5, 1.0)):
    crop = _img(crop)
    h, w = crop.shape[:2]
    dst = np.empty_like(crop)
    rc = lib().orc_meanshift_filter_roi(crop.ctypes.data, crop.strides[0], dst.ctypes.data, dst.strides[0], w, h,
                                        int(xoff), int(yoff), int(full_w), int(full_h), float(sp), float(sr),
                                        int(max_level), int(term[0]), int(term[1]), float(term[2]), None)
    if rc != 0:
        raise ValueError("orc_meanshift_filter_roi: invalid arguments")
    return dst


def label_regions(bgr, d=2, connectivity=4):
    bgr = _img(bgr)
    h, w = bgr.shape[:2]
    lab = np.empty((h, w), np.int32)
    n = lib().orc_label_regions_conn(bgr.ctypes.data, bgr.strides[0], lab.ctypes.data, lab.strides[0], w, h, int(d),
                                     int(connectivity))
    return n, lab


def connected_components(mask, connectivity=8):
    mask = np.ascontiguousarray(mask, dtype=np.uint8)
    h, w = mask.shape
    lab = np.empty((h, w), np.int32)
    n = lib().orc_connected_components(mask.ctypes.data, mask.strides[0], lab.ctypes.data, lab.strides[0], w, h,
                                       int(connectivity))
    return n, lab


def relabel_canonical(labels):
    lab = np.ascontiguousarray(labels, dtype=np.int32).copy()
    h, w = lab.shape
    n = lib().orc_relabel_canonical(lab.ctypes.data, lab.strides[0], w, h)
    return n, lab


def merge_regions(bgr, labels, min_size, color_dist):
    bgr = _img(bgr)
    lab = np.ascontiguousarray(labels, dtype=np.int32).copy()
    h, w = lab.shape
    n = lib().orc_merge_regions(bgr.ctypes.data, bgr.strides[0], lab.ctypes.data, lab.strides[0], w, h,
                                int(min_size), int(color_dist))
    return n, lab


def render_labels(labels, depth, colors=None):
    lab = np.ascontiguousarray(labels, dtype=np.int32)
    h, w = lab.shape
    dst = np.empty((h, w, 3), np.uint8)
    cptr = None
    if colors is not None:
        colors = np.ascontiguousarray(colors, dtype=np.uint8)
        assert colors.shape == (depth, 3)
        cptr = colors.ctypes.data
    lib().orc_render_labels(lab.ctypes.data, lab.strides[0], dst.ctypes.data, dst.strides[0], w, h, int(depth), cptr)
    return dst


def watershed(bgr, markers):
    bgr = _img(bgr)
    m = np.ascontiguousarray(markers, dtype=np.int32).copy()
    h, w = m.shape
    lib().orc_watershed(bgr.ctypes.data, bgr.strides[0], m.ctypes.data, m.strides[0], w, h)
    return m


def synth_bgr(w, h, seed):
    dst = np.empty((h, w, 3), np.uint8)
    lib().orc_synth_bgr(dst.ctypes.data, dst.strides[0], w, h, int(seed))
    return dst


def laplacian_sharpen(src, taps):
    src = _img(src)
    taps = np.ascontiguousarray(taps, dtype=np.int8)
    assert taps.ndim == 2
    h, w = src.shape[:2]
    dst = np.empty_like(src)
    lib().orc_laplacian_sharpen(src.ctypes.data, src.strides[0], dst.ctypes.data, dst.strides[0], w, h, taps.ctypes.data,
                                taps.shape[0], taps.shape[1])
    return dst


def median_blur(gray, k):
    gray = np.ascontiguousarray(gray, dtype=np.uint8)
    h, w = gray.shape
    dst = np.empty_like(gray)
    lib().orc_median_blur_8uc1(gray.ctypes.data, gray.strides[0], dst.ctypes.data, dst.strides[0], w, h, int(k))
    return dst


def bgr2gray(src, compat342=False):
    src = _img(src)
    h, w = src.shape[:2]
    dst = np.empty((h, w), np.uint8)
    fn = lib().orc_bgr2gray_342 if compat342 else lib().orc_bgr2gray
    fn(src.ctypes.data, src.strides[0], dst.ctypes.data, dst.strides[0], w, h)
    return dst


def canny(gray, low, high):
    gray = np.ascontiguousarray(gray, dtype=np.uint8)
    h, w = gray.shape
    dst = np.empty_like(gray)
    lib().orc_canny(gray.ctypes.data, gray.strides[0], dst.ctypes.data, dst.strides[0], w, h, float(low), float(high))
    return dst


def dilate_rect(img, kw, kh):
    img = np.ascontiguousarray(img, dtype=np.uint8)
    h, w = img.shape
    dst = np.empty_like(img)
    lib().orc_dilate_rect(img.ctypes.data, img.strides[0], dst.ctypes.data, dst.strides[0], w, h, int(kw), int(kh))
    return dst


def subtract_u8(a, b):
    a = np.ascontiguousarray(a, dtype=np.uint8)
    b = np.ascontiguousarray(b, dtype=np.uint8)
    h, w = a.shape
    dst = np.empty_like(a)
    lib().orc_subtract_u8(a.ctypes.data, a.strides[0], b.ctypes.data, b.strides[0], dst.ctypes.data, dst.strides[0], w, h)
    return dst


def blur_mask_size(w, h):
    """PictureService.calculateSizeOfSquareBlurMask (PictureService.java:877-899)."""
    m = min(w, h)
    if m < 3:
        return 1
    if m <= 100:
        return 5
    scale = 0.025 if m <= 360 else 0.02 if m <= 480 else 0.015 if m <= 720 else 0.01 if m <= 1080 else 0.005
    r = int(m * scale)
    return r + 1 if r % 2 == 0 else r


def shape_seeds(bgr, low=5, high=50):
    """Shape-method marker generator (PictureService.java:404-442): returns (n_labels, markers int32, stages dict)."""
    g = bgr2gray(bgr)
    k = blur_mask_size(g.shape[1], g.shape[0])
    blurred = median_blur(g, k)
    edges = canny(blurred, low, high)
    d3 = dilate_rect(edges, 3, 3)
    d5 = dilate_rect(d3, 5, 5)
    dde = subtract_u8(d5, d3)
    dde3 = median_blur(dde, 3)
    n, markers = connected_components(dde3, 8)
    return n, markers, {"gray": g, "blurred": blurred, "edges": edges, "dde": dde, "dde3": dde3, "k": k}


# ---------------------------------------------------------------- colour-method seeds (rows a6 / a4) and bilateral filter
def _u8(a):
    return np.ascontiguousarray(a, dtype=np.uint8)


def otsu_threshold(gray):
    gray = _u8(gray)
    h, w = gray.shape
    return int(lib().orc_otsu_threshold(gray.ctypes.data, gray.strides[0], w, h))


def threshold_binary(gray, thresh, maxval=255):
    gray = _u8(gray)
    h, w = gray.shape
    dst = np.empty_like(gray)
    lib().orc_threshold_binary_u8(gray.ctypes.data, gray.strides[0], dst.ctypes.data, dst.strides[0], w, h, int(thresh), int(maxval))
    return dst


def distance_transform(mask, fixed=False):
    """fixed=False: the float arithmetic of the IPP-backed cv2 build; fixed=True: OpenCV's own 16.16 fixed-point chamfer."""
    mask = _u8(mask)
    h, w = mask.shape
    dst = np.empty((h, w), np.float32)
    fn = lib().orc_distance_transform_l2_5_fixed if fixed else lib().orc_distance_transform_l2_5
    fn(mask.ctypes.data, mask.strides[0], dst.ctypes.data, dst.strides[0], w, h)
    return dst


def normalize_minmax01(dist):
    dist = np.ascontiguousarray(dist, dtype=np.float32)
    h, w = dist.shape
    dst = np.empty_like(dist)
    lib().orc_normalize_minmax01_f32(dist.ctypes.data, dist.strides[0], dst.ctypes.data, dst.strides[0], w, h)
    return dst


def peaks(nrm, thresh=0.4):
    nrm = np.ascontiguousarray(nrm, dtype=np.float32)
    h, w = nrm.shape
    dst = np.empty((h, w), np.uint8)
    lib().orc_peaks_u8(nrm.ctypes.data, nrm.strides[0], dst.ctypes.data, dst.strides[0], w, h, float(thresh))
    return dst


def contour_markers(mask):
    mask = _u8(mask)
    h, w = mask.shape
    out = np.empty((h, w), np.int32)
    n = lib().orc_contour_markers(mask.ctypes.data, mask.strides[0], out.ctypes.data, out.strides[0], w, h)
    return int(n), out


def circle_filled(markers, cx, cy, radius, value):
    markers = np.ascontiguousarray(markers, dtype=np.int32).copy()
    h, w = markers.shape
    lib().orc_circle_filled_i32(markers.ctypes.data, markers.strides[0], w, h, int(cx), int(cy), int(radius), int(value))
    return markers


def white_to_black(bgr):
    bgr = _img(bgr)
    h, w = bgr.shape[:2]
    dst = np.empty_like(bgr)
    lib().orc_white_to_black(bgr.ctypes.data, bgr.strides[0], dst.ctypes.data, dst.strides[0], w, h)
    return dst


def bilateral_filter(img, d, sigma_color, sigma_space):
    img = _u8(img)
    h, w = img.shape[:2]
    cn = 1 if img.ndim == 2 else img.shape[2]
    dst = np.empty_like(img)
    lib().orc_bilateral_filter(img.ctypes.data, img.strides[0], dst.ctypes.data, dst.strides[0], w, h, cn, int(d),
                               float(sigma_color), float(sigma_space))
    return dst


SHARPEN_TAPS_9x1 = np.array([1, 1, 1, 1, -8, 1, 1, 1, 1], np.int8).reshape(9, 1)   # literal MatOfFloat reading (App. C#2)


def color_seeds(bgr, taps=SHARPEN_TAPS_9x1, peak_thresh=0.4, dt_fixed=False, gray_compat=False):
    """Colour-method marker generator (PictureService.java:309-366): returns (n_contours, markers int32, stages dict).
    dt_fixed / gray_compat: the arithmetic of a non-IPP OpenCV 3.4.2 build (the library's options of the same names)."""
    black = white_to_black(bgr)
    sharp = laplacian_sharpen(black, taps)
    gray = bgr2gray(sharp, compat342=gray_compat)
    t = otsu_threshold(gray)
    bw = threshold_binary(gray, t, 255)
    dist = distance_transform(bw, fixed=dt_fixed)
    nrm = normalize_minmax01(dist)
    pk = peaks(nrm, peak_thresh)
    n, markers = contour_markers(pk)
    markers = circle_filled(markers, 5, 5, 3, 255)
    return n, markers, {"black_bg": black, "sharp": sharp, "gray": gray, "otsu": t, "bw": bw, "dist": dist, "norm": nrm,
                        "peaks": pk}
