"""Generates tests/golden/bilateral.npz from the live cv2 wheel: small frames with the output of OpenCV's own code path
(cv2.ipp.setUseIPP(False)) and of the wheel's default (IPP) path for cv2.bilateralFilter as the reference calls it
(src/loader/loader.py:16-20,86).   usage: python tests/golden/make_bilateral_golden.py"""
import os
import sys

import cv2
import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from visual_odom_pipeline_b200 import synth as S  # noqa: E402

out = {}
cases = [((96, 161), (5, 1.5, 1.5)), ((61, 64), (5, 1.5, 1.5)), ((50, 77), (7, 25.0, 3.0)), ((40, 33), (3, 10.0, 2.0))]
rng = np.random.default_rng(11)
for k, ((h, w), p) in enumerate(cases):
    img = np.clip(S.texture(h, w, seed=20 + k) + rng.integers(-25, 25, (h, w)), 0, 255).astype(np.uint8)
    cv2.ipp.setUseIPP(False)
    out["generic%d" % k] = cv2.bilateralFilter(img, int(p[0]), p[1], p[2])
    cv2.ipp.setUseIPP(True)
    out["ipp%d" % k] = cv2.bilateralFilter(img, int(p[0]), p[1], p[2])
    out["img%d" % k] = img
    out["p%d" % k] = np.array(p, np.float64)
out["n"] = np.array(len(cases))
out["cv2_version"] = np.array(cv2.__version__)
np.savez_compressed(os.path.join(ROOT, "tests", "golden", "bilateral.npz"), **out)
print("wrote bilateral.npz with", len(cases), "cases")
