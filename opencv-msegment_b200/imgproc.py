"""Host-side mirror of the reference's operator interface for the segmentation hot path.

The reference (Java) calls static methods of ``org.opencv.imgproc.Imgproc`` on ``Mat`` objects
(PictureService.java:441-442 connectedComponents, :909 watershed, :913-936 colorByIndexes) and
BASELINE.json's north_star adds pyrMeanShiftFiltering / floodFill-style labelling / region merge.
This module exposes the same names with the argument meaning of the OpenCV Java/Python API on
numpy arrays (the Python spelling of ``Mat``): ``GpuImgproc.pyrMeanShiftFiltering(src, sp, sr,
maxLevel, termcrit)`` etc.  Every call goes through the C ABI (include/msegment.h) into CUDA; a
non-zero status raises ``CvException`` exactly where the Java shim throws
``org.opencv.core.CvException``.  No CPU implementation exists here.
"""
import ctypes as C

import numpy as np

from . import _lib as L

CV_32S = 4                     # org.opencv.core.CvType.CV_32S
CV_16U = 2                     # org.opencv.core.CvType.CV_16U
TERM_COUNT, TERM_EPS = L.TERM_COUNT, L.TERM_EPS
DEFAULT_TERMCRIT = (TERM_COUNT | TERM_EPS, 5, 1.0)   # Imgproc.pyrMeanShiftFiltering 4-arg overload


class CvException(RuntimeError):
    """Mirror of org.opencv.core.CvException (unchecked)."""

    def __init__(self, status, message):
        super().__init__("msegment status %d: %s" % (status, message))
        self.status = status


class Context:
    """One msg_ctx: one CUDA device, one stream, a grow-only HBM workspace.  Not thread-safe."""

    def __init__(self, device=0):
        self._lib = L.load()
        h = C.c_void_p()
        rc = self._lib.msg_create(int(device), C.byref(h))
        if rc != L.MSG_OK:
            raise CvException(rc, (self._lib.msg_last_error(None) or b"").decode())
        self._h = h
        self.device = device

    def close(self):
        if getattr(self, "_h", None):
            self._lib.msg_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()

    def check(self, rc):
        if rc != L.MSG_OK:
            raise CvException(rc, (self._lib.msg_last_error(self._h) or b"").decode())

    # -- introspection
    def timings(self):
        t = L.Timings()
        self.check(self._lib.msg_get_timings(self._h, C.byref(t)))
        return {n: getattr(t, n) for n, _ in L.Timings._fields_}

    def stats(self):
        s = L.Stats()
        self.check(self._lib.msg_get_stats(self._h, C.byref(s)))
        return {n: getattr(s, n) for n, _ in L.Stats._fields_}

    def set_option(self, name, value):
        self.check(self._lib.msg_set_option(self._h, name.encode(), int(value)))

    def get_option(self, name):
        v = C.c_int()
        self.check(self._lib.msg_get_option(self._h, name.encode(), C.byref(v)))
        return v.value

    def register_host(self, array):
        """Page-lock a long-lived numpy array (msg_register_host); unregister_host before it is freed."""
        self.check(self._lib.msg_register_host(self._h, array.ctypes.data, array.nbytes))

    def unregister_host(self, array):
        self.check(self._lib.msg_unregister_host(self._h, array.ctypes.data))

    def set_profiling(self, enable):
        self.check(self._lib.msg_set_profiling(self._h, 1 if enable else 0))

    def kernel_profile(self):
        p = L.KernelProfile()
        self.check(self._lib.msg_get_kernel_profile(self._h, C.byref(p)))
        return {n: list(getattr(p, n)) for n, _ in L.KernelProfile._fields_}

    def set_stream(self, cuda_stream):
        """cuda_stream: a cudaStream_t handle as int; None = the context's own stream; 0 = the legacy default stream
        (what torch.cuda.current_stream().cuda_stream returns outside a stream context), passed as cudaStreamLegacy."""
        if cuda_stream is None:
            handle = None
        else:
            handle = int(cuda_stream) or 1          # cudaStreamLegacy == (cudaStream_t)0x1
        self.check(self._lib.msg_set_stream(self._h, C.c_void_p(handle)))

    def synchronize(self):
        self.check(self._lib.msg_synchronize(self._h))

    def debug_plane(self, kind, level):
        w, h = C.c_int(), C.c_int()
        self.check(self._lib.msg_debug_get_plane(self._h, kind, level, None, 0, C.byref(w), C.byref(h)))
        out = np.empty((h.value, w.value), np.uint32)
        self.check(self._lib.msg_debug_get_plane(self._h, kind, level, out.ctypes.data, out.size, C.byref(w), C.byref(h)))
        return out


def _mat8uc3(a, what):
    a = np.asarray(a)
    if a.dtype != np.uint8 or a.ndim != 3 or a.shape[2] != 3:
        raise CvException(L.MSG_EINVAL, "%s must be CV_8UC3 (HxWx3 uint8), got %s %s" % (what, a.dtype, a.shape))
    if a.strides[2] != 1 or a.strides[1] != 3:
        a = np.ascontiguousarray(a)
    return a


def _mat32s(a, what):
    a = np.asarray(a)
    if a.dtype != np.int32 or a.ndim != 2:
        raise CvException(L.MSG_EINVAL, "%s must be CV_32SC1 (HxW int32), got %s %s" % (what, a.dtype, a.shape))
    if a.strides[1] != 4:
        a = np.ascontiguousarray(a)
    return a


class GpuImgproc:
    """Static-method style mirror of ``Imgproc`` bound to a Context (``GpuImgproc(ctx).name(...)``)."""

    def __init__(self, ctx=None):
        self.ctx = ctx if ctx is not None else Context(0)
        self._lib = self.ctx._lib

    # Imgproc.pyrMeanShiftFiltering(Mat src, Mat dst, double sp, double sr, int maxLevel, TermCriteria termcrit)
    def pyrMeanShiftFiltering(self, src, sp, sr, maxLevel=1, termcrit=DEFAULT_TERMCRIT, dst=None):
        src = _mat8uc3(src, "src")
        h, w = src.shape[:2]
        if dst is None:
            dst = np.empty((h, w, 3), np.uint8)
        else:
            if not isinstance(dst, np.ndarray) or dst.shape != src.shape or dst.dtype != np.uint8:
                raise CvException(L.MSG_EINVAL, "dst must have the size and type of src")
            # a Mat's pixels are contiguous and its rows ascend: a sliced / reversed view would be written to the wrong place
            if dst.strides[2] != 1 or dst.strides[1] != 3 or dst.strides[0] < 3 * w:
                raise CvException(L.MSG_EINVAL, "dst must have contiguous pixels and a positive row stride (got strides %s)" % (dst.strides,))
        self.ctx.check(self._lib.msg_meanshift_filter(self.ctx._h, src.ctypes.data, src.strides[0], dst.ctypes.data,
                                                      dst.strides[0], w, h, float(sp), float(sr), int(maxLevel),
                                                      int(termcrit[0]), int(termcrit[1]), float(termcrit[2])))
        return dst

    # floodFill region-growing loop of OpenCV's meanshift_segmentation sample, as one call
    def labelRegions(self, image, loDiff=2, upDiff=2, connectivity=4):
        image = _mat8uc3(image, "image")
        h, w = image.shape[:2]
        labels = np.empty((h, w), np.int32)
        n = C.c_int32()
        self.ctx.check(self._lib.msg_label_regions(self.ctx._h, image.ctypes.data, image.strides[0], labels.ctypes.data,
                                                   labels.strides[0], w, h, int(loDiff), int(upDiff), int(connectivity),
                                                   C.byref(n)))
        return n.value, labels

    def mergeRegions(self, image, labels, minSize, colorDist):
        image = _mat8uc3(image, "image")
        labels = _mat32s(labels, "labels").copy()
        h, w = image.shape[:2]
        if labels.shape != (h, w):
            raise CvException(L.MSG_EINVAL, "labels must have the size of image")
        n = C.c_int32()
        self.ctx.check(self._lib.msg_merge_regions(self.ctx._h, image.ctypes.data, image.strides[0], labels.ctypes.data,
                                                   labels.strides[0], w, h, int(minSize), int(colorDist), C.byref(n)))
        return n.value, labels

    # Imgproc.connectedComponents(Mat image, Mat labels, int connectivity, int ltype) -- PictureService.java:441-442
    def connectedComponents(self, image, connectivity=8, ltype=CV_32S):
        image = np.asarray(image)
        if image.dtype != np.uint8 or image.ndim != 2:
            raise CvException(L.MSG_EINVAL, "image must be CV_8UC1")
        if ltype != CV_32S:
            raise CvException(L.MSG_EINVAL, "only ltype CV_32S is supported (the reference passes CvType.CV_32S)")
        if image.strides[1] != 1:
            image = np.ascontiguousarray(image)
        h, w = image.shape
        labels = np.empty((h, w), np.int32)
        n = C.c_int32()
        self.ctx.check(self._lib.msg_connected_components(self.ctx._h, image.ctypes.data, image.strides[0],
                                                          labels.ctypes.data, labels.strides[0], w, h,
                                                          int(connectivity), C.byref(n)))
        return n.value, labels

    # PictureService.colorByIndexes(Mat markers, Integer depth, boolean colored) -- PictureService.java:913-936
    def colorByIndexes(self, markers, depth, colors=None):
        markers = _mat32s(markers, "markers")
        h, w = markers.shape
        dst = np.empty((h, w, 3), np.uint8)
        cptr = None
        if colors is not None:
            colors = np.ascontiguousarray(colors, dtype=np.uint8)
            if colors.shape != (depth, 3):
                raise CvException(L.MSG_EINVAL, "colors must be depth x 3 bytes")
            cptr = colors.ctypes.data
        self.ctx.check(self._lib.msg_render_labels(self.ctx._h, markers.ctypes.data, markers.strides[0], dst.ctypes.data,
                                                   dst.strides[0], w, h, int(depth), cptr))
        return dst

    # Imgproc.watershed(Mat image, Mat markers) -- PictureService.java:908-911; markers is modified IN PLACE like the Mat
    def watershed(self, image, markers):
        image = _mat8uc3(image, "image")
        if not (isinstance(markers, np.ndarray) and markers.dtype == np.int32 and markers.ndim == 2 and markers.strides[1] == 4
                and markers.strides[0] >= 4 * markers.shape[1]):
            raise CvException(L.MSG_EINVAL, "markers must be a CV_32SC1 ndarray with contiguous pixels")
        if markers.shape != image.shape[:2]:
            raise CvException(L.MSG_EINVAL, "markers must have the size of image")
        h, w = markers.shape
        self.ctx.check(self._lib.msg_watershed(self.ctx._h, image.ctypes.data, image.strides[0], markers.ctypes.data,
                                               markers.strides[0], w, h))
        return markers

    # ---- pre-filters the reference calls around the segmentation stage
    # filter2D + convertTo + subtract + convertTo chain of PictureService.java:323-333 as one call; kernel = integer taps
    def sharpenLaplacian(self, src, kernel):
        src = _mat8uc3(src, "src")
        taps = np.ascontiguousarray(kernel, dtype=np.int8)
        if taps.ndim != 2:
            raise CvException(L.MSG_EINVAL, "kernel must be 2-D (MatOfFloat(...) is N x 1)")
        h, w = src.shape[:2]
        dst = np.empty_like(src)
        self.ctx.check(self._lib.msg_laplacian_sharpen(self.ctx._h, src.ctypes.data, src.strides[0], dst.ctypes.data,
                                                       dst.strides[0], w, h, taps.ctypes.data, taps.shape[0], taps.shape[1]))
        return dst

    # Imgproc.cvtColor(src, dst, Imgproc.COLOR_BGR2GRAY)
    def cvtColorBGR2GRAY(self, src):
        src = _mat8uc3(src, "src")
        h, w = src.shape[:2]
        dst = np.empty((h, w), np.uint8)
        self.ctx.check(self._lib.msg_bgr2gray(self.ctx._h, src.ctypes.data, src.strides[0], dst.ctypes.data, dst.strides[0], w, h))
        return dst

    # Imgproc.medianBlur(src, dst, ksize)
    def medianBlur(self, src, ksize):
        src = np.asarray(src)
        if src.dtype != np.uint8 or src.ndim != 2:
            raise CvException(L.MSG_EINVAL, "medianBlur: CV_8UC1 only (the reference blurs gray images)")
        if src.strides[1] != 1:
            src = np.ascontiguousarray(src)
        h, w = src.shape
        dst = np.empty_like(src)
        self.ctx.check(self._lib.msg_median_blur(self.ctx._h, src.ctypes.data, src.strides[0], dst.ctypes.data, dst.strides[0],
                                                 w, h, int(ksize)))
        return dst

    # ---- shape-method marker generator (PictureService.java:404-442)
    def _gray(self, a, what):
        a = np.asarray(a)
        if a.dtype != np.uint8 or a.ndim != 2:
            raise CvException(L.MSG_EINVAL, "%s: CV_8UC1 only" % what)
        return a if a.strides[1] == 1 else np.ascontiguousarray(a)

    # Imgproc.Canny(image, edges, threshold1, threshold2)
    def Canny(self, image, threshold1, threshold2):
        image = self._gray(image, "Canny")
        h, w = image.shape
        dst = np.empty((h, w), np.uint8)
        self.ctx.check(self._lib.msg_canny(self.ctx._h, image.ctypes.data, image.strides[0], dst.ctypes.data, dst.strides[0], w, h,
                                           float(threshold1), float(threshold2)))
        return dst

    # Imgproc.dilate(src, dst, Mat.ones(kh, kw, CV_8U))
    def dilate(self, src, kernel_shape):
        src = self._gray(src, "dilate")
        kh, kw = kernel_shape
        h, w = src.shape
        dst = np.empty((h, w), np.uint8)
        self.ctx.check(self._lib.msg_dilate(self.ctx._h, src.ctypes.data, src.strides[0], dst.ctypes.data, dst.strides[0], w, h,
                                            int(kw), int(kh)))
        return dst

    # Core.subtract(src1, src2, dst)
    def subtract(self, src1, src2):
        a, b = self._gray(src1, "subtract"), self._gray(src2, "subtract")
        if a.shape != b.shape:
            raise CvException(L.MSG_EINVAL, "subtract: sizes differ")
        h, w = a.shape
        dst = np.empty((h, w), np.uint8)
        self.ctx.check(self._lib.msg_subtract(self.ctx._h, a.ctypes.data, a.strides[0], b.ctypes.data, b.strides[0],
                                              dst.ctypes.data, dst.strides[0], w, h))
        return dst

    @staticmethod
    def calculateSizeOfSquareBlurMask(cols, rows):
        """PictureService.calculateSizeOfSquareBlurMask (PictureService.java:877-899)."""
        m = min(cols, rows)
        if m < 3:
            return 1
        if m <= 100:
            return 5
        scale = 0.025 if m <= 360 else 0.02 if m <= 480 else 0.015 if m <= 720 else 0.01 if m <= 1080 else 0.005
        r = int(m * scale)
        return r + 1 if r % 2 == 0 else r

    # the marker half of PictureService.shapeAutoMarkerWatershed (:404-442) as one call, intermediates on the device
    def shapeSeeds(self, src, lowThreshold=5, ratio=10, medianKsize=None, stages=False):
        src = _mat8uc3(src, "src")
        h, w = src.shape[:2]
        k = self.calculateSizeOfSquareBlurMask(w, h) if medianKsize is None else int(medianKsize)
        markers = np.empty((h, w), np.int32)
        st = np.empty((4, h, w), np.uint8) if stages else None
        n = C.c_int32(0)
        self.ctx.check(self._lib.msg_shape_seeds(self.ctx._h, src.ctypes.data, src.strides[0], w, h, k, float(lowThreshold),
                                                 float(lowThreshold * ratio), markers.ctypes.data, markers.strides[0],
                                                 C.byref(n), st.ctypes.data if stages else None, w))
        if stages:
            return n.value, markers, {"blurred": st[0], "edges": st[1], "dde": st[2], "dde3": st[3], "k": k}
        return n.value, markers

    # ---- colour-method marker generator (PictureService.java:309-366)
    THRESH_BINARY, THRESH_OTSU, CV_DIST_L2, NORM_MINMAX = 0, 8, 2, 32
    SHARPEN_KERNEL = ((1,), (1,), (1,), (1,), (-8,), (1,), (1,), (1,), (1,))   # new MatOfFloat(1,1,1,1,-8,1,1,1,1): 9 x 1

    def _f32(self, a, what):
        a = np.asarray(a)
        if a.dtype != np.float32 or a.ndim != 2:
            raise CvException(L.MSG_EINVAL, "%s: CV_32FC1 only" % what)
        return a if a.strides[1] == 4 else np.ascontiguousarray(a)

    # the Java loop of PictureService.java:309-318
    def whiteToBlack(self, src):
        src = _mat8uc3(src, "src")
        h, w = src.shape[:2]
        dst = np.empty_like(src)
        self.ctx.check(self._lib.msg_white_to_black(self.ctx._h, src.ctypes.data, src.strides[0], dst.ctypes.data, dst.strides[0], w, h))
        return dst

    # Imgproc.threshold(src, dst, thresh, maxval, type) -> (computed threshold, dst); 8UC1 (BINARY [| OTSU]) or 32FC1 (BINARY)
    def threshold(self, src, thresh, maxval, type=0):
        src = np.asarray(src)
        if src.dtype == np.float32:
            src = self._f32(src, "threshold")
            if type != self.THRESH_BINARY:
                raise CvException(L.MSG_EINVAL, "threshold: CV_32F supports THRESH_BINARY only")
            h, w = src.shape
            dst = np.empty((h, w), np.float32)
            self.ctx.check(self._lib.msg_threshold_f32(self.ctx._h, src.ctypes.data, src.strides[0], dst.ctypes.data, dst.strides[0],
                                                       w, h, float(thresh), float(maxval)))
            return float(thresh), dst
        src = self._gray(src, "threshold")
        h, w = src.shape
        dst = np.empty((h, w), np.uint8)
        used = C.c_double(0)
        self.ctx.check(self._lib.msg_threshold(self.ctx._h, src.ctypes.data, src.strides[0], dst.ctypes.data, dst.strides[0], w, h,
                                               float(thresh), float(maxval), int(type), C.byref(used)))
        return used.value, dst

    # Imgproc.distanceTransform(src, dst, Imgproc.CV_DIST_L2, 5)
    def distanceTransform(self, src, distanceType=2, maskSize=5):
        src = self._gray(src, "distanceTransform")
        h, w = src.shape
        dst = np.empty((h, w), np.float32)
        self.ctx.check(self._lib.msg_distance_transform(self.ctx._h, src.ctypes.data, src.strides[0], dst.ctypes.data, dst.strides[0],
                                                        w, h, int(distanceType), int(maskSize)))
        return dst

    # Core.normalize(src, dst, alpha, beta, Core.NORM_MINMAX)
    def normalize(self, src, alpha=0.0, beta=1.0, norm_type=32):
        if norm_type != self.NORM_MINMAX:
            raise CvException(L.MSG_EINVAL, "normalize: NORM_MINMAX only")
        src = self._f32(src, "normalize")
        h, w = src.shape
        dst = np.empty((h, w), np.float32)
        self.ctx.check(self._lib.msg_normalize_minmax(self.ctx._h, src.ctypes.data, src.strides[0], dst.ctypes.data, dst.strides[0],
                                                      w, h, float(alpha), float(beta)))
        return dst

    # Imgproc.dilate on CV_32FC1 (PictureService.java:349-350)
    def dilateF32(self, src, kernel_shape):
        src = self._f32(src, "dilate")
        kh, kw = kernel_shape
        h, w = src.shape
        dst = np.empty((h, w), np.float32)
        self.ctx.check(self._lib.msg_dilate_f32(self.ctx._h, src.ctypes.data, src.strides[0], dst.ctypes.data, dst.strides[0], w, h,
                                                int(kw), int(kh)))
        return dst

    # Mat.convertTo(dst, CvType.CV_8U) from CV_32FC1
    def convertToU8(self, src):
        src = self._f32(src, "convertTo")
        h, w = src.shape
        dst = np.empty((h, w), np.uint8)
        self.ctx.check(self._lib.msg_convert_f32_to_u8(self.ctx._h, src.ctypes.data, src.strides[0], dst.ctypes.data, dst.strides[0], w, h))
        return dst

    # findContours(RETR_CCOMP, CHAIN_APPROX_NONE) + the drawContours loop of PictureService.java:360-364 -> (contours.size(), markers)
    def contourMarkers(self, image):
        image = self._gray(image, "findContours")
        h, w = image.shape
        markers = np.empty((h, w), np.int32)
        n = C.c_int32(0)
        self.ctx.check(self._lib.msg_contour_markers(self.ctx._h, image.ctypes.data, image.strides[0], markers.ctypes.data,
                                                     markers.strides[0], w, h, C.byref(n)))
        return n.value, markers

    # Imgproc.circle(img, center, radius, color, -1) on CV_32SC1, in place like OpenCV
    def circle(self, img, center, radius, value):
        if not (isinstance(img, np.ndarray) and img.dtype == np.int32 and img.ndim == 2 and img.strides[1] == 4):
            raise CvException(L.MSG_EINVAL, "circle: CV_32SC1 ndarray only")
        h, w = img.shape
        self.ctx.check(self._lib.msg_circle_filled(self.ctx._h, img.ctypes.data, img.strides[0], w, h, int(center[0]), int(center[1]),
                                                   int(radius), int(value)))
        return img

    # the marker half of PictureService.colorAutoMarkerWatershed (:309-366) as one call, intermediates on the device
    def colorSeeds(self, src, kernel=None, peakThresh=0.4, stages=False):
        src = _mat8uc3(src, "src")
        taps = np.ascontiguousarray(self.SHARPEN_KERNEL if kernel is None else kernel, dtype=np.int8)
        if taps.ndim != 2:
            raise CvException(L.MSG_EINVAL, "kernel must be 2-D (MatOfFloat(...) is N x 1)")
        h, w = src.shape[:2]
        markers = np.empty((h, w), np.int32)
        n = C.c_int32(0)
        if stages:
            sharp, bw = np.empty((h, w, 3), np.uint8), np.empty((h, w), np.uint8)
            dist, peaks = np.empty((h, w), np.float32), np.empty((h, w), np.uint8)
            args = (sharp.ctypes.data, sharp.strides[0], bw.ctypes.data, bw.strides[0], dist.ctypes.data, dist.strides[0],
                    peaks.ctypes.data, peaks.strides[0])
        else:
            args = (None, 0, None, 0, None, 0, None, 0)
        self.ctx.check(self._lib.msg_color_seeds(self.ctx._h, src.ctypes.data, src.strides[0], w, h, taps.ctypes.data, taps.shape[0],
                                                 taps.shape[1], float(peakThresh), markers.ctypes.data, markers.strides[0],
                                                 C.byref(n), *args))
        if stages:
            return n.value, markers, {"sharp": sharp, "bw": bw, "norm": dist, "peaks": peaks}
        return n.value, markers

    # Imgproc.bilateralFilter(src, dst, d, sigmaColor, sigmaSpace) on CV_8UC1 / CV_8UC3 (PictureService.java:490)
    def bilateralFilter(self, src, d, sigmaColor, sigmaSpace):
        src = np.asarray(src)
        if src.dtype != np.uint8 or not (src.ndim == 2 or (src.ndim == 3 and src.shape[2] == 3)):
            raise CvException(L.MSG_EINVAL, "bilateralFilter: CV_8UC1 or CV_8UC3 only")
        cn = 1 if src.ndim == 2 else 3
        if src.strides[-1] != 1 or (cn == 3 and src.strides[1] != 3):
            src = np.ascontiguousarray(src)
        h, w = src.shape[:2]
        dst = np.empty_like(src)
        self.ctx.check(self._lib.msg_bilateral_filter(self.ctx._h, src.ctypes.data, src.strides[0], dst.ctypes.data, dst.strides[0],
                                                      w, h, cn, int(d), float(sigmaColor), float(sigmaSpace)))
        return dst

    # fused pipeline
    def segment(self, src, sp=10.0, sr=10.0, maxLevel=1, termcrit=DEFAULT_TERMCRIT, loDiff=2, minSize=0, colorDist=0,
                renderDepth=0, want=("filtered", "labels", "rendered"), connectivity=4, labelsType=CV_32S):
        """want = the download mask: products not named are neither converted nor copied back.  labelsType = CV_32S (the
        reference's markers Mat) or CV_16U (2 bytes per pixel; CvException MSG_ERANGE beyond 65535 regions)."""
        src = _mat8uc3(src, "src")
        h, w = src.shape[:2]
        if labelsType not in (CV_32S, CV_16U):
            raise CvException(L.MSG_EINVAL, "labelsType must be CV_32S or CV_16U")
        p = L.SegmentParams(float(sp), float(sr), int(maxLevel), int(termcrit[0]), int(termcrit[1]), float(termcrit[2]),
                            int(loDiff), int(minSize), int(colorDist), int(renderDepth), int(connectivity),
                            L.LABELS_16U if labelsType == CV_16U else L.LABELS_32S)
        out = {}
        f = np.empty((h, w, 3), np.uint8) if "filtered" in want else None
        lab = np.empty((h, w), np.uint16 if labelsType == CV_16U else np.int32) if "labels" in want else None
        r = np.empty((h, w, 3), np.uint8) if "rendered" in want else None
        n = C.c_int32()
        self.ctx.check(self._lib.msg_segment(
            self.ctx._h, src.ctypes.data, src.strides[0], w, h, C.byref(p),
            f.ctypes.data if f is not None else None, f.strides[0] if f is not None else 0,
            lab.ctypes.data if lab is not None else None, lab.strides[0] if lab is not None else 0,
            r.ctypes.data if r is not None else None, r.strides[0] if r is not None else 0, C.byref(n)))
        out["n_regions"] = n.value
        if f is not None:
            out["filtered"] = f
        if lab is not None:
            out["labels"] = lab
        if r is not None:
            out["rendered"] = r
        return out
