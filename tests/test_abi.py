"""CPU tier: the C-ABI library loads and exports every symbol include/klt_b200.h declares; the host
logic that needs no GPU (planning, validation, sharding) behaves like the reference boundary."""
import ctypes
import os
import re
import subprocess

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def header_symbols():
    src = open(os.path.join(ROOT, "include", "klt_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(klt_[a-z0-9_]+)\s*\(", src)))


def test_header_declares_expected_entry_points():
    syms = header_symbols()
    for s in ["klt_create", "klt_destroy", "klt_pyr_plan", "klt_pyr_build", "klt_pyr_down", "klt_lk_track",
              "klt_calc_optical_flow_pyr_lk_host", "klt_build_optical_flow_pyramid_host", "klt_host_alloc", "klt_host_free"]:
        assert s in syms


def test_library_exports_every_declared_symbol(native_lib):
    lib = ctypes.CDLL(native_lib)
    for s in header_symbols():
        assert hasattr(lib, s), "libklt_b200.so does not export %s" % s
    out = subprocess.run(["nm", "-D", "--defined-only", native_lib], capture_output=True, text=True).stdout
    exported = set(re.findall(r" T (klt_\w+)", out))
    assert set(header_symbols()) <= exported


def test_python_binding_covers_header(klt):
    from visual_odom_pipeline_b200 import _lib
    assert sorted(_lib.SYMBOLS) == header_symbols()
    assert _lib.load().klt_version() == 120


def test_library_holds_sm100a_sass_only(native_lib):
    out = subprocess.run(["cuobjdump", "-lelf", native_lib], capture_output=True, text=True).stdout
    archs = set(re.findall(r"sm_(\d+a?)", out))
    assert archs == {"100a"}, archs


def test_pyr_plan_follows_cv2_level_rule(klt, oracle):
    from visual_odom_pipeline_b200 import _lib
    L = _lib.load()
    for (w, h, win, ml) in [(1241, 376, (21, 21), 3), (1241, 376, (21, 21), 8), (1241, 376, (31, 31), 8), (640, 480, (21, 21), 3),
                            (3840, 2160, (31, 31), 5), (70, 50, (21, 21), 3), (1024, 768, (5, 40), 10)]:
        lay = _lib.klt_pyr_layout()
        assert L.klt_pyr_plan(w, h, win[0], win[1], ml, 4, ctypes.byref(lay)) == 0
        assert lay.top == oracle.pyr_max_level(w, h, win, ml)
        lw, lh = w, h
        for l in range(1, lay.top + 1):
            lw, lh = (lw + 1) // 2, (lh + 1) // 2
            lv = lay.level[l]
            assert (lv.w, lv.h) == (lw, lh) and lv.pitch >= lw and lv.pitch % 32 == 0 and lv.batch_stride >= lv.pitch * lh
        assert lay.bytes == sum(lay.level[l].batch_stride * 4 for l in range(1, lay.top + 1))
    assert L.klt_pyr_plan(100, 100, 2, 21, 3, 1, ctypes.byref(lay)) == _lib.KLT_ERR_INVALID_ARG
    assert L.klt_pyr_plan(100, 100, 21, 21, -1, 1, ctypes.byref(lay)) == _lib.KLT_ERR_INVALID_ARG


def test_argument_validation_matches_cv2_error_class(klt):
    """Same inputs cv2 rejects raise cv2.error here, before any GPU work."""
    import cv2
    img = np.zeros((48, 64), np.uint8)
    pts = np.zeros((4, 1, 2), np.float32)
    bad = [
        dict(prevPts=pts.astype(np.float64)),
        dict(prevPts=np.zeros((4, 3), np.float32)),
        dict(prevImg=img.astype(np.float32)),
        dict(nextImg=np.zeros((48, 65), np.uint8)),
        dict(winSize=(2, 21)),
        dict(maxLevel=-1),
    ]
    for kw in bad:
        args = dict(prevImg=img, nextImg=img, prevPts=pts, nextPts=None, winSize=(21, 21), maxLevel=3)
        args.update(kw)
        with pytest.raises(cv2.error):
            klt.calcOpticalFlowPyrLK(**args)
        with pytest.raises(cv2.error):  # and cv2 itself rejects the same call
            cv2.calcOpticalFlowPyrLK(args["prevImg"], args["nextImg"], args["prevPts"], None, winSize=args["winSize"], maxLevel=args["maxLevel"])
    assert issubclass(klt.error, cv2.error)


def test_empty_point_set_returns_nones(klt):
    img = np.zeros((48, 64), np.uint8)
    assert klt.calcOpticalFlowPyrLK(img, img, np.zeros((0, 1, 2), np.float32), None) == (None, None, None)


def test_no_cpu_fallback_without_gpu(klt):
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    img = np.zeros((48, 64), np.uint8)
    with pytest.raises(klt.KLTLibraryError):
        klt.calcOpticalFlowPyrLK(img, img, np.zeros((4, 1, 2), np.float32), None)


def test_product_package_never_imports_oracle():
    pkg = os.path.join(ROOT, "visual-odom-pipeline_b200")
    for dp, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                txt = open(os.path.join(dp, f)).read()
                assert "klt_oracle" not in txt and "from oracle" not in txt and "import oracle" not in txt, f


def test_shard_range_is_a_partition():
    from visual_odom_pipeline_b200.sharding import shard_range, shard_sizes
    for n in [0, 1, 7, 256, 1000]:
        for w in [1, 2, 3, 4, 8]:
            spans = [shard_range(n, w, r) for r in range(w)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(spans[i][1] == spans[i + 1][0] for i in range(w - 1))
            assert max(shard_sizes(n, w)) - min(shard_sizes(n, w)) <= 1
