"""Import shim: the package directory is `opencv-msegment_b200/` (the name the project brief fixes),
which is not a valid Python identifier; this module loads it as `opencv_msegment_b200`."""
import importlib.util
import os
import sys

_ROOT = os.path.dirname(os.path.abspath(__file__))
_PKG_DIR = os.path.join(_ROOT, "opencv-msegment_b200")
_NAME = "opencv_msegment_b200"

if _NAME not in sys.modules:
    _spec = importlib.util.spec_from_file_location(_NAME, os.path.join(_PKG_DIR, "__init__.py"),
                                                   submodule_search_locations=[_PKG_DIR])
    _mod = importlib.util.module_from_spec(_spec)
    sys.modules[_NAME] = _mod
    _spec.loader.exec_module(_mod)

pkg = sys.modules[_NAME]
Context, CvException, GpuImgproc = pkg.Context, pkg.CvException, pkg.GpuImgproc
lib = pkg._lib
device = pkg.device
imgproc = pkg.imgproc
CV_32S, CV_16U = pkg.imgproc.CV_32S, pkg.imgproc.CV_16U
synth_bgr = pkg.synth.synth_bgr
PKG_DIR = _PKG_DIR
