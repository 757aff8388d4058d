"""CPU tier of the detection step: the C oracle of cv2.cornerMinEigenVal /
cv2.goodFeaturesToTrack (oracle/gftt_oracle.c) is pinned bit-for-bit against the live cv2 module (the
reference's own implementation, called at src/extractor/extractor.py:110-111) and against the committed
golden vectors; the sequential host tail of the product (klt_select_corners_host: sort + greedy
minimum-distance selection, no GPU needed) is checked against cv2's selection."""
import ctypes

import numpy as np
import pytest

from conftest import load_golden

GOLDEN = ["gftt_reference", "gftt_default", "gftt_even_block", "gftt_block5"]


def _mask_with_discs(shape, n, seed):
    import cv2
    rng = np.random.default_rng(seed)
    m = np.full(shape, 255, np.uint8)
    for _ in range(n):
        cv2.circle(m, (int(rng.integers(0, shape[1])), int(rng.integers(0, shape[0]))), 10, 0, -1)
    return m


def _same_corners(a, b):
    if a is None or b is None:
        return a is None and b is None
    return a.shape == b.shape and np.array_equal(a, b)


@pytest.mark.parametrize("name", GOLDEN)
def test_oracle_matches_golden(oracle, name):
    g = load_golden(name)
    img, bs = g["img"], int(g["blockSize"])
    mask = g["mask"] if g["mask"].size else None
    eig = oracle.corner_min_eigen_val(img, bs)
    assert np.array_equal(eig.view(np.uint32), g["eig"].view(np.uint32)), "eigenvalue map differs from cv2's"
    c = oracle.good_features_to_track(img, int(g["maxCorners"]), float(g["qualityLevel"]), float(g["minDistance"]), mask=mask, block_size=bs)
    want = g["corners"] if len(g["corners"]) else None
    assert _same_corners(c, want)


@pytest.mark.parametrize("hw", [(376, 1241), (480, 640), (100, 101), (57, 43), (64, 96)])
@pytest.mark.parametrize("bs", [31, 3, 5, 7, 4])
def test_oracle_bit_exact_vs_live_cv2(oracle, hw, bs):
    import cv2
    from visual_odom_pipeline_b200 import synth as S
    if bs // 2 >= min(hw):
        pytest.skip("block larger than the image")
    img = S.frame_pair(hw[0], hw[1], seed=hw[1] + bs)[0]
    ref = cv2.cornerMinEigenVal(img, bs, ksize=3)
    got = oracle.corner_min_eigen_val(img, bs)
    bad = np.argwhere(got.view(np.uint32) != ref.view(np.uint32))
    assert bad.size == 0, "%d eigenvalues differ, first at %s: %r vs %r" % (len(bad), bad[0], got[tuple(bad[0])], ref[tuple(bad[0])])
    mask = _mask_with_discs(img.shape, 30, bs)
    for (mc, ql, md, m) in [(1000, 0.03, 10, mask), (1000, 0.03, 7, None), (0, 0.01, 3.5, mask), (50, 0.2, 0, None), (200, 0.001, 25, mask)]:
        c = oracle.good_features_to_track(img, mc, ql, md, mask=m, block_size=bs)
        k = cv2.goodFeaturesToTrack(img, mc, ql, md, mask=m, blockSize=bs)
        assert _same_corners(c, k), (mc, ql, md, m is not None)


def test_oracle_edge_cases_vs_live_cv2(oracle):
    import cv2
    flat = np.full((60, 80), 77, np.uint8)
    assert oracle.good_features_to_track(flat, 100, 0.01, 5, block_size=7) is None
    assert cv2.goodFeaturesToTrack(flat, 100, 0.01, 5, blockSize=7) is None
    from visual_odom_pipeline_b200 import synth as S
    img = S.frame_pair(60, 80, seed=4)[0]
    empty = np.zeros(img.shape, np.uint8)
    assert oracle.good_features_to_track(img, 100, 0.01, 5, mask=empty, block_size=7) is None
    assert cv2.goodFeaturesToTrack(img, 100, 0.01, 5, mask=empty, blockSize=7) is None
    # periodic pattern: plateaus of exactly equal eigenvalues exercise the tie order (later pixel first)
    per = np.tile(np.array([[0, 255], [255, 0]], np.uint8).repeat(8, 0).repeat(8, 1), (6, 8))
    for md in (0, 4, 9.5):
        assert _same_corners(oracle.good_features_to_track(per, 0, 0.01, md, block_size=3), cv2.goodFeaturesToTrack(per, 0, 0.01, md, blockSize=3))


def _keys_from_eig(eig, quality, mask=None):
    """The candidate list the device kernels hand to the host tail (include/klt_b200.h: klt_corner_candidates)."""
    import cv2
    mx = eig[mask != 0].max() if mask is not None and (mask != 0).any() else (eig.max() if mask is None else np.float32(0))
    thr = np.float32(float(mx) * quality)
    e = np.where(eig > thr, eig, np.float32(0)).astype(np.float32)
    d = cv2.dilate(e, np.ones((3, 3), np.uint8))
    cand = (e != 0) & (e == d)
    if mask is not None:
        cand &= mask != 0
    cand[0, :] = cand[-1, :] = False
    cand[:, 0] = cand[:, -1] = False
    ys, xs = np.nonzero(cand)
    bits = e[ys, xs].view(np.uint32).astype(np.uint64) | np.uint64(0x80000000)   # positive floats
    return (bits << np.uint64(32)) | (ys.astype(np.uint64) << np.uint64(16)) | xs.astype(np.uint64)


@pytest.mark.parametrize("params", [(1000, 0.03, 10.0), (0, 0.01, 3.5), (50, 0.2, 0.0), (300, 0.001, 25.0), (0, 0.05, 1.0)])
def test_host_selection_tail_equals_cv2(klt, params):
    """klt_select_corners_host (product code, runs on the host) on cv2's own eigenvalue map == cv2's corners."""
    import cv2
    from visual_odom_pipeline_b200 import _lib, synth as S
    mc, ql, md = params
    L = _lib.load()
    for hw, seed in [((376, 1241), 3), ((120, 161), 5)]:
        img = S.frame_pair(hw[0], hw[1], seed=seed)[0]
        mask = _mask_with_discs(img.shape, 25, seed)
        eig = cv2.cornerMinEigenVal(img, 31 if hw[0] > 200 else 7, ksize=3)
        keys = _keys_from_eig(eig, ql, mask)
        keys = keys[np.random.default_rng(seed).permutation(len(keys))].copy()   # device order is arbitrary
        cap = mc if mc > 0 else len(keys)
        out = np.empty((max(cap, 1), 2), np.float32)
        n = ctypes.c_int()
        rc = L.klt_select_corners_host(keys.ctypes.data, len(keys), hw[1], hw[0], mc, md, out.ctypes.data, cap, ctypes.byref(n))
        assert rc == 0
        ref = cv2.goodFeaturesToTrack(img, mc, ql, md, mask=mask, blockSize=31 if hw[0] > 200 else 7)
        ref = np.zeros((0, 2), np.float32) if ref is None else ref.reshape(-1, 2)
        assert n.value == len(ref) and np.array_equal(out[:n.value], ref)


def test_host_selection_rejects_bad_arguments(klt):
    from visual_odom_pipeline_b200 import _lib
    L = _lib.load()
    n = ctypes.c_int()
    keys = np.array([(0xC0000000 << 32) | (5 << 16) | 7], np.uint64)
    out = np.empty((4, 2), np.float32)
    assert L.klt_select_corners_host(keys.ctypes.data, 1, 64, 48, 10, 5.0, out.ctypes.data, 4, ctypes.byref(n)) == 0
    assert n.value == 1 and tuple(out[0]) == (7.0, 5.0)
    assert L.klt_select_corners_host(keys.ctypes.data, 1, 64, 48, -1, 5.0, out.ctypes.data, 4, ctypes.byref(n)) == _lib.KLT_ERR_INVALID_ARG
    assert L.klt_select_corners_host(keys.ctypes.data, 1, 64, 48, 10, -1.0, out.ctypes.data, 4, ctypes.byref(n)) == _lib.KLT_ERR_INVALID_ARG
    assert L.klt_select_corners_host(keys.ctypes.data, 1, 6, 48, 10, 5.0, out.ctypes.data, 4, ctypes.byref(n)) == _lib.KLT_ERR_INVALID_ARG   # x >= w


def test_detection_argument_validation_matches_cv2_error_class(klt):
    import cv2
    img = np.zeros((48, 64), np.uint8)
    for kw in [dict(qualityLevel=0.0), dict(minDistance=-1.0), dict(maxCorners=-5), dict(mask=np.zeros((48, 65), np.uint8)),
               dict(mask=np.zeros((48, 64), np.float32))]:
        args = dict(image=img, maxCorners=10, qualityLevel=0.01, minDistance=3.0, mask=None, blockSize=3)
        args.update(kw)
        with pytest.raises(cv2.error):
            klt.goodFeaturesToTrack(**args)
        with pytest.raises(cv2.error):
            cv2.goodFeaturesToTrack(args["image"], args["maxCorners"], args["qualityLevel"], args["minDistance"], mask=args["mask"], blockSize=args["blockSize"])
    with pytest.raises(cv2.error):
        klt.goodFeaturesToTrack(img.astype(np.float64), 10, 0.01, 3.0)
    with pytest.raises(cv2.error):
        klt.goodFeaturesToTrack(img, 10, 0.01, 3.0, useHarrisDetector=True)   # not implemented here: loud, never silent


def test_detection_has_no_cpu_fallback(klt):
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    img = np.zeros((48, 64), np.uint8)
    with pytest.raises(klt.KLTLibraryError):
        klt.goodFeaturesToTrack(img, 10, 0.01, 3.0)
    with pytest.raises(klt.KLTLibraryError):
        klt.cornerMinEigenVal(img, 3)


@pytest.mark.parametrize("radius", [0, 1, 2, 3, 5, 7, 10, 16, 31, 50])
def test_oracle_circle_mask_equals_cv2_circle(oracle, radius):
    """Detection mask of extractor.py:102-107: np.int32 centres (inside, on and outside the border), filled cv2.circle."""
    import cv2
    rng = np.random.default_rng(radius)
    for shape in [(60, 80), (37, 23), (200, 300)]:
        h, w = shape
        pts = np.stack([rng.uniform(-radius - 5, w + radius + 5, 60), rng.uniform(-radius - 5, h + radius + 5, 60)], -1).astype(np.float32)
        want = np.zeros(shape, np.uint8)
        want[:] = 255
        for x, y in [np.int32(p) for p in pts.astype(np.float64)]:
            cv2.circle(want, (int(x), int(y)), radius, 0, -1)
        assert np.array_equal(oracle.mask_from_points(pts, radius, shape), want)
