"""ctypes front-end of oracle/klt_oracle.c -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.

Restates what the reference delegates to ``cv2.calcOpticalFlowPyrLK`` at
``src/extractor/extractor.py:44,45,65,66``.  Pinned bit-exactly against the
live ``cv2`` module (the reference's own implementation of the path) in tests/test_oracle_vs_cv2.py
and against tests/golden/*.npz.
"""
import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "_build", "libklt_oracle.so")
_lib = None

TERM_COUNT, TERM_EPS = 1, 2
USE_INITIAL_FLOW, GET_MIN_EIGENVALS = 4, 8


def build(force=False):
    """Compile the oracle with gcc (seconds)."""
    src = os.path.join(_HERE, "klt_oracle.c")
    if force or not os.path.exists(_SO) or os.path.getmtime(_SO) < os.path.getmtime(src):
        subprocess.check_call(["make", "-s", "-C", _HERE, "-B", "all"])
    return _SO


def lib():
    global _lib
    if _lib is None:
        build()
        L = ctypes.CDLL(_SO)
        c = ctypes
        L.klt_oracle_pyr_down.argtypes = [c.c_void_p, c.c_int, c.c_int, c.c_int64, c.c_void_p, c.c_int64]
        L.klt_oracle_pyr_down.restype = c.c_int
        L.klt_oracle_pyr_max_level.argtypes = [c.c_int] * 5
        L.klt_oracle_pyr_max_level.restype = c.c_int
        L.klt_oracle_scharr.argtypes = [c.c_void_p, c.c_int, c.c_int, c.c_int64, c.c_void_p]
        L.klt_oracle_scharr.restype = c.c_int
        L.klt_oracle_calc_optical_flow_pyr_lk.argtypes = [
            c.c_void_p, c.c_int64, c.c_void_p, c.c_int64, c.c_int, c.c_int,
            c.c_void_p, c.c_void_p, c.c_void_p, c.c_void_p, c.c_int,
            c.c_int, c.c_int, c.c_int, c.c_int, c.c_int, c.c_double, c.c_int, c.c_double, c.c_void_p]
        L.klt_oracle_calc_optical_flow_pyr_lk.restype = c.c_int
        _lib = L
    return _lib


def _u8_image(img):
    img = np.asarray(img)
    if img.dtype != np.uint8 or img.ndim != 2:
        raise ValueError("oracle: image must be 2-D uint8")
    if img.strides[1] != 1:
        img = np.ascontiguousarray(img)
    return img


def pyr_down(img):
    img = _u8_image(img)
    h, w = img.shape
    dst = np.empty(((h + 1) // 2, (w + 1) // 2), np.uint8)
    rc = lib().klt_oracle_pyr_down(img.ctypes.data, w, h, img.strides[0], dst.ctypes.data, dst.strides[0])
    assert rc == 0
    return dst


def pyr_max_level(w, h, win_size, max_level):
    return lib().klt_oracle_pyr_max_level(w, h, win_size[0], win_size[1], max_level)


def build_pyramid(img, win_size=(21, 21), max_level=3):
    """-> (last level index, [level0, level1, ...]) like cv2.buildOpticalFlowPyramid (no derivs)."""
    img = _u8_image(img)
    h, w = img.shape
    top = pyr_max_level(w, h, win_size, max_level)
    levels = [np.ascontiguousarray(img)]
    for _ in range(top):
        levels.append(pyr_down(levels[-1]))
    return top, levels


def scharr(img):
    """-> int16 (h, w, 2) with [...,0]=dI/dx, [...,1]=dI/dy (== cv2.Scharr(..., CV_16S))."""
    img = _u8_image(img)
    h, w = img.shape
    out = np.empty((h, w, 2), np.int16)
    rc = lib().klt_oracle_scharr(img.ctypes.data, w, h, img.strides[0], out.ctypes.data)
    assert rc == 0
    return out


def calc_optical_flow_pyr_lk(prev_img, next_img, prev_pts, next_pts=None, win_size=(21, 21), max_level=3,
                             criteria=(3, 30, 0.01), flags=0, min_eig_threshold=1e-4, return_iters=False):
    """Same contract as cv2.calcOpticalFlowPyrLK -> (nextPts, status, err) [+ iters]."""
    prev_img, next_img = _u8_image(prev_img), _u8_image(next_img)
    if prev_img.shape != next_img.shape:
        raise ValueError("oracle: image size mismatch")
    p = np.asarray(prev_pts)
    if p.dtype != np.float32 or p.size % 2:
        raise ValueError("oracle: prevPts must be float32 Nx2")
    shape = p.shape
    p = np.ascontiguousarray(p.reshape(-1, 2))
    n = p.shape[0]
    if n == 0:
        return (None, None, None) + ((None,) if return_iters else ())
    if flags & USE_INITIAL_FLOW:
        q = np.ascontiguousarray(np.asarray(next_pts, np.float32).reshape(-1, 2)).copy()
    else:
        q = np.zeros_like(p)
    status = np.empty((n, 1), np.uint8)
    err = np.empty((n, 1), np.float32)
    iters = np.zeros(n, np.int32)
    h, w = prev_img.shape
    rc = lib().klt_oracle_calc_optical_flow_pyr_lk(
        prev_img.ctypes.data, prev_img.strides[0], next_img.ctypes.data, next_img.strides[0], w, h,
        p.ctypes.data, q.ctypes.data, status.ctypes.data, err.ctypes.data, n,
        int(win_size[0]), int(win_size[1]), int(max_level), int(criteria[0]), int(criteria[1]), float(criteria[2]),
        int(flags), float(min_eig_threshold), iters.ctypes.data)
    if rc < 0:
        raise ValueError("oracle: invalid arguments (rc=%d)" % rc)
    out = (q.reshape(shape), status, err)
    return out + (iters,) if return_iters else out
