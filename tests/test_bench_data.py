"""The synthetic pool of bench.py: distinct pairs without wrap-around seams (CPU)."""
import importlib.util
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _bench():
    spec = importlib.util.spec_from_file_location("klt_bench", os.path.join(ROOT, "bench.py"))
    mod = importlib.util.module_from_spec(spec)
    argv = sys.argv
    sys.argv = ["bench.py"]
    try:
        spec.loader.exec_module(mod)
    finally:
        sys.argv = argv
    return mod


def _column_jumps(img):
    """mean |difference| between neighbouring columns, per column boundary"""
    return np.abs(np.diff(img.astype(np.int32), axis=1)).mean(axis=0)


def test_pool_entries_are_seam_free_crops():
    B = _bench()
    wl = dict(B.WORKLOADS["kitti"])
    wl.update(h=94, w=310, n=50)                        # a quarter-size KITTI frame keeps this fast
    pool = B.host_pool(wl, 8, seed0=7, max_extra_off=2 * 53)
    assert len(pool) == 8
    for a, b, p in pool:
        assert a.shape == b.shape == (94, 310) and a.dtype == np.uint8 and a.flags.c_contiguous
        assert p.shape == (50, 1, 2) and p.dtype == np.float32
    # distinct bytes in every entry, and the device pool's further offsets give new crops of the same canvas
    assert len({a.tobytes() for a, _, _ in pool}) == 8
    a0, b0 = pool.crop(0)
    assert np.array_equal(a0, pool[0][0]) and np.array_equal(b0, pool[0][1])
    a1, _ = pool.crop(0, 53)
    assert np.array_equal(a1[:, :-53], a0[:, 53:])      # the same canvas, 53 columns further
    # no seam: no column boundary stands out (a cyclic shift leaves one where unrelated columns meet)
    for i in (4, 5, 7):                                 # entries with a non-zero offset
        j = _column_jumps(pool[i][0])
        assert j.max() < 4 * np.median(j), (i, j.max(), np.median(j))


def test_legacy_roll_pool_has_the_seam():
    B = _bench()
    wl = dict(B.WORKLOADS["kitti"])
    wl.update(h=94, w=310, n=50)
    B.LEGACY_ROLL_POOL = True
    try:
        pool = B.host_pool(wl, 8, seed0=7)
    finally:
        B.LEGACY_ROLL_POOL = False
    j = _column_jumps(pool[4][0])                       # shifted by 37 columns: the seam sits at boundary 36 | 37
    assert int(np.argmax(j)) == 36 and j.max() > 4 * np.median(j)
