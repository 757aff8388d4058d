"""ctypes front-end of oracle/klt_oracle.c -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.

Restates what the reference delegates to ``cv2.calcOpticalFlowPyrLK`` at
``src/extractor/extractor.py:44,45,65,66``.  Pinned bit-exactly against the
live ``cv2`` module (the reference's own implementation of the path) in tests/test_oracle.py
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
    srcs = [os.path.join(_HERE, f) for f in ("klt_oracle.c", "gftt_oracle.c", "bilateral_oracle.c", "Makefile")]
    if force or not os.path.exists(_SO) or os.path.getmtime(_SO) < max(os.path.getmtime(f) for f in srcs):
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
        L.klt_oracle_corner_min_eigen_val.argtypes = [c.c_void_p, c.c_int, c.c_int, c.c_int64, c.c_int, c.c_void_p]
        L.klt_oracle_corner_min_eigen_val.restype = c.c_int
        L.klt_oracle_select_corners.argtypes = [c.c_void_p, c.c_int, c.c_int, c.c_void_p, c.c_int64, c.c_int, c.c_double,
                                                c.c_double, c.c_void_p, c.c_int]
        L.klt_oracle_select_corners.restype = c.c_int
        L.klt_oracle_good_features_to_track.argtypes = [c.c_void_p, c.c_int, c.c_int, c.c_int64, c.c_void_p, c.c_int64, c.c_int,
                                                        c.c_double, c.c_double, c.c_int, c.c_void_p, c.c_int]
        L.klt_oracle_good_features_to_track.restype = c.c_int
        L.klt_oracle_circle_half_widths.argtypes = [c.c_int, c.c_void_p]
        L.klt_oracle_circle_half_widths.restype = c.c_int
        L.klt_oracle_mask_from_points.argtypes = [c.c_void_p, c.c_int, c.c_int, c.c_int, c.c_int, c.c_void_p, c.c_int64]
        L.klt_oracle_mask_from_points.restype = c.c_int
        L.klt_oracle_bilateral_filter.argtypes = [c.c_void_p, c.c_int, c.c_int, c.c_int64, c.c_int, c.c_double, c.c_double,
                                                  c.c_int, c.c_int, c.c_void_p, c.c_int64]
        L.klt_oracle_bilateral_filter.restype = c.c_int
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


def corner_min_eigen_val(img, block_size, ksize=3):
    """cv2.cornerMinEigenVal(img, blockSize, ksize=3) -> float32 (h, w)   (oracle/gftt_oracle.c G.1-G.6)."""
    if ksize != 3:
        raise ValueError("oracle: only ksize 3")
    img = _u8_image(img)
    h, w = img.shape
    out = np.empty((h, w), np.float32)
    rc = lib().klt_oracle_corner_min_eigen_val(img.ctypes.data, w, h, img.strides[0], int(block_size), out.ctypes.data)
    if rc < 0:
        raise ValueError("oracle: corner_min_eigen_val failed (%d)" % rc)
    return out


def _mask(mask, shape):
    if mask is None:
        return None, 0, 0
    mask = np.asarray(mask)
    if mask.dtype != np.uint8 or mask.shape != shape:
        raise ValueError("oracle: mask must be uint8 of the image's shape")
    if mask.strides[1] != 1:
        mask = np.ascontiguousarray(mask)
    return mask, mask.ctypes.data, mask.strides[0]


def select_corners(eig, max_corners, quality_level, min_distance, mask=None):
    """G.7 + G.8 on a given eigenvalue map -> float32 (n, 1, 2) or None."""
    eig = np.ascontiguousarray(eig, np.float32)
    h, w = eig.shape
    mask, mp, ms = _mask(mask, (h, w))
    cap = w * h
    out = np.empty((cap, 2), np.float32)
    n = lib().klt_oracle_select_corners(eig.ctypes.data, w, h, mp, ms, int(max_corners), float(quality_level),
                                        float(min_distance), out.ctypes.data, cap)
    if n < 0:
        raise ValueError("oracle: select_corners failed (%d)" % n)
    return out[:n].reshape(-1, 1, 2).copy() if n else None


def good_features_to_track(img, max_corners, quality_level, min_distance, mask=None, block_size=3):
    """cv2.goodFeaturesToTrack(img, maxCorners, qualityLevel, minDistance, mask=mask, blockSize=...) as called at
    reference src/extractor/extractor.py:110-111 -> float32 (n, 1, 2) or None."""
    img = _u8_image(img)
    h, w = img.shape
    mask, mp, ms = _mask(mask, (h, w))
    cap = w * h
    out = np.empty((cap, 2), np.float32)
    n = lib().klt_oracle_good_features_to_track(img.ctypes.data, w, h, img.strides[0], mp, ms, int(max_corners),
                                                float(quality_level), float(min_distance), int(block_size),
                                                out.ctypes.data, cap)
    if n < 0:
        raise ValueError("oracle: good_features_to_track failed (%d)" % n)
    return out[:n].reshape(-1, 1, 2).copy() if n else None


def mask_from_points(points, radius, shape):
    """The detection mask of reference src/extractor/extractor.py:102-107: 255 everywhere, filled cv2.circle of `radius`
    with value 0 around np.int32 of every point -> uint8 (h, w)."""
    pts = np.ascontiguousarray(np.asarray(points, np.float32).reshape(-1, 2))
    h, w = int(shape[0]), int(shape[1])
    out = np.empty((h, w), np.uint8)
    rc = lib().klt_oracle_mask_from_points(pts.ctypes.data if len(pts) else None, len(pts), int(radius), w, h, out.ctypes.data, w)
    if rc < 0:
        raise ValueError("oracle: mask_from_points failed (%d)" % rc)
    return out


def bilateral_filter(img, d, sigma_color, sigma_space, simd_lanes=8, tail_fma=False):
    """cv2.bilateralFilter(img, d, sigmaColor, sigmaSpace) on a uint8 (h, w) image, OpenCV's own code path (B.1-B.6 in
    oracle/bilateral_oracle.c; reference src/loader/loader.py:16-20,86).  simd_lanes = vector width of the OpenCV build
    (8: AVX2 dispatch of the wheel), columns of the scalar tail take OpenCV's 4-at-a-time path."""
    img = np.ascontiguousarray(img, np.uint8)
    h, w = img.shape
    out = np.empty((h, w), np.uint8)
    rc = lib().klt_oracle_bilateral_filter(img.ctypes.data, w, h, img.strides[0], int(d), float(sigma_color), float(sigma_space),
                                           int(simd_lanes), int(bool(tail_fma)), out.ctypes.data, w)
    if rc < 0:
        raise ValueError("oracle: bilateral_filter failed (%d)" % rc)
    return out
