"""CPU: the C-ABI library builds, loads without a GPU and exports every symbol include/msegment.h declares."""
import ctypes
import os
import re

import pytest

import msegment_b200 as mseg

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    src = open(os.path.join(ROOT, "include", "msegment.h")).read()
    return sorted(set(re.findall(r"^MSG_API\s+[\w\s\*]+?\b(msg_\w+)\s*\(", src, flags=re.M)))


def test_header_declares_expected_surface():
    syms = declared_symbols()
    for must in ("msg_create", "msg_meanshift_filter", "msg_label_regions", "msg_merge_regions",
                 "msg_connected_components", "msg_render_labels", "msg_segment", "msg_submit_segment", "msg_wait",
                 "msg_alloc_pinned", "msg_meanshift_filter_dev", "msg_meanshift_filter_strip_dev"):
        assert must in syms
    assert len(syms) >= 30


def test_library_exports_every_declared_symbol():
    assert os.path.exists(mseg.lib.LIB_PATH), "run `python __graft_entry__.py build` first"
    lib = ctypes.CDLL(mseg.lib.LIB_PATH)
    missing = [s for s in declared_symbols() if not hasattr(lib, s)]
    assert not missing, missing


def test_python_binding_covers_the_header():
    assert sorted(mseg.lib.SIGNATURES) == declared_symbols()


def test_version_and_no_cpu_fallback():
    lib = mseg.lib.load()
    assert lib.msg_version() == 200
    if lib.msg_device_count() == 0:
        with pytest.raises(mseg.CvException) as ei:
            mseg.Context(0)
        assert ei.value.status == mseg.lib.MSG_ECUDA
        assert "no CPU fallback" in str(ei.value)


def test_halo_rows_is_pure_host_function():
    lib = mseg.lib.load()
    h1 = lib.msg_meanshift_halo_rows(10.0, 1, 3, 5)
    h0 = lib.msg_meanshift_halo_rows(10.0, 0, 3, 5)
    assert h0 >= 50 and h1 >= 58 and h1 % 2 == 0
    assert lib.msg_meanshift_halo_rows(10.0, 9, 3, 5) < 0


def test_product_does_not_use_oracle():
    """The product path must never import, link or call the CPU oracle."""
    pkg = os.path.join(ROOT, "opencv-msegment_b200")
    bad = re.compile(r"libmsg_oracle|from\s+oracle|import\s+oracle|msg_oracle\.h|\borc_\w+\s*\(")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".h", ".cpp", ".hpp", ".java", ".c")):
                txt = open(os.path.join(dirpath, f), errors="ignore").read()
                assert not bad.search(txt), (dirpath, f)
    out = os.popen("ldd %s 2>/dev/null" % mseg.lib.LIB_PATH).read()
    assert "oracle" not in out


def test_numpy_synth_matches_oracle_generator():
    """Three generators (numpy in the package, C in the oracle, CUDA kernel -- checked on the GPU) must agree byte for byte."""
    import numpy as np
    from oracle import oracle as orc
    for (w, h, s) in [(200, 150, 1), (64, 64, 7), (333, 129, 5), (1, 1, 2), (70, 200, 1000)]:
        assert np.array_equal(mseg.synth_bgr(w, h, s), orc.synth_bgr(w, h, s)), (w, h, s)


def test_only_allowed_files_touch_the_oracle():
    """tools/ and the package must not import the oracle; tests/, __graft_entry__.smoke() and bench.py's CPU legs may."""
    bad = re.compile(r"^\s*(from\s+oracle|import\s+oracle)", re.M)
    for d in ("tools", "opencv-msegment_b200"):
        for dirpath, _, files in os.walk(os.path.join(ROOT, d)):
            for f in files:
                if f.endswith(".py"):
                    assert not bad.search(open(os.path.join(dirpath, f)).read()), (dirpath, f)


def test_bench_cpu_binding_is_best_effort():
    """bench.bind_to_gpu_cpus never raises (no NVML / no GPU here) and leaves the allowed CPU set non-empty."""
    import importlib
    import os
    import sys
    sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    bench = importlib.import_module("bench")
    before = os.sched_getaffinity(0)
    info = bench.bind_to_gpu_cpus(0)
    assert info["cpus_allowed"] == len(before) and "affinity" in info
    assert os.sched_getaffinity(0) and os.sched_getaffinity(0) <= before
    os.sched_setaffinity(0, before)


def test_jni_glue_type_checks_against_stub_jni_h():
    """No JDK in this image: compile-check java/msegment_jni.c against tests/stubs/jni.h (the JNI types and the JNIEnv entries
    the glue uses, with the specification's signatures) and the real include/msegment.h, warnings as errors -- every call
    into the C ABI is checked for argument count and types."""
    import subprocess
    r = subprocess.run(["gcc", "-fsyntax-only", "-std=c99", "-Wall", "-Wextra", "-Wno-unused-parameter", "-Werror",
                        "-I", os.path.join(ROOT, "tests", "stubs"), "-I", os.path.join(ROOT, "include"),
                        os.path.join(ROOT, "opencv-msegment_b200", "java", "msegment_jni.c")], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr


def test_java_natives_match_jni_glue():
    """No JDK here, so the Java shim cannot be compiled: check at least that every `native` method of GpuImgproc.java has a JNI
    function of the same name and arity in msegment_jni.c (JNIEnv*, jclass + the Java parameters), and vice versa, and that
    every msg_* function the glue calls is declared in include/msegment.h."""
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    java = open(os.path.join(root, "opencv-msegment_b200", "java", "GpuImgproc.java")).read()
    glue = open(os.path.join(root, "opencv-msegment_b200", "java", "msegment_jni.c")).read()
    header = open(os.path.join(root, "include", "msegment.h")).read()
    natives = {m.group(1): len([a for a in m.group(2).split(",") if a.strip()])
               for m in re.finditer(r"private static native \w+(?:\[\])? (\w+)\(([^)]*)\)", java, re.S)}
    jni = {m.group(1): len([a for a in m.group(2).split(",") if a.strip()]) - 2
           for m in re.finditer(r"JNICALL J\((\w+)\)\(([^)]*)\)", glue, re.S)}
    assert len(natives) >= 31
    assert natives == jni, (sorted(set(natives.items()) ^ set(jni.items())))
    declared = set(re.findall(r"\b(msg_\w+)\s*\(", header))
    called = set(re.findall(r"\b(msg_\w+)\s*\(", glue))
    assert called <= declared, sorted(called - declared)
    # every public method of the shim that forwards to a native uses one that exists
    used = set(re.findall(r"\b(n[A-Z]\w+)\(", java)) - {"new"}
    assert used <= set(natives), sorted(used - set(natives))
