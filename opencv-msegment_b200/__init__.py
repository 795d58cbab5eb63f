"""opencv-msegment_b200: B200-native (sm_100a CUDA) segmentation hot path behind the OpenCV Imgproc
calls of ShayHulud/opencv-msegment.  The product is csrc/ -> libmsegment_b200.so (C ABI in
include/msegment.h); this package is the thin host-side mirror used by tests, bench and the CLI."""
from . import _lib, device, imgproc, sharded, synth
from .imgproc import CV_16U, CV_32S, DEFAULT_TERMCRIT, TERM_COUNT, TERM_EPS, Context, CvException, GpuImgproc

__all__ = ["_lib", "device", "sharded", "synth", "Context", "CvException", "GpuImgproc", "CV_32S", "CV_16U", "imgproc", "TERM_COUNT", "TERM_EPS", "DEFAULT_TERMCRIT"]
