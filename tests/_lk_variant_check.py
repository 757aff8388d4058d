"""Run in a fresh process with KLT_LK_WPP set (the library reads it once): every team size of the specialised LK kernel
(1, 2 or 4 warps per keypoint; the default picks by window and point count) must stay
bit-identical to live cv2.  Used by tests/test_gpu_parity.py::test_lk_team_sizes_bit_exact."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import cv2  # noqa: E402
import visual_odom_pipeline_b200 as K  # noqa: E402
from visual_odom_pipeline_b200 import synth as S  # noqa: E402

CASES = [
    (376, 1241, 2000, (21, 21), 3, (3, 30, 0.01), S.BENIGN, 0, {}),
    (376, 1241, 1500, (31, 31), 3, (3, 30, 0.03), S.HARD, 60, dict(noise_sigma=3.0, flat_cols=(400, 700))),
    (480, 640, 500, (21, 21), 3, (3, 10, 0.01), S.BENIGN, 10, {}),
    (120, 160, 96, (21, 21), 3, (3, 30, 0.01), S.BENIGN, 5, {}),
]
# windows larger than a white-noise image: the gradient energy of a window sits in one or two threads (clamp regression)
NOISE_CASES = [(12, 12, (21, 21)), (12, 12, (31, 31)), (25, 25, (31, 31))]
bad = 0
for rep in range(2):
    for h, w, n, win, lvl, crit, motion, margin, kw in CASES:
        a, b = S.frame_pair(h, w, seed=7 + rep, motion=motion, **kw)
        p = S.uniform_points(n, h, w, seed=3 + rep, margin=margin)
        q, st, er = K.calcOpticalFlowPyrLK(a, b, p, None, winSize=win, maxLevel=lvl, criteria=crit)
        rq, rs, re_ = cv2.calcOpticalFlowPyrLK(a, b, p, None, winSize=win, maxLevel=lvl, criteria=crit)
        m = rs.ravel() == 1
        ok = (np.array_equal(q.view(np.uint32), rq.view(np.uint32)) and np.array_equal(st, rs)
              and np.array_equal(er.ravel()[m].view(np.uint32), re_.ravel()[m].view(np.uint32)))
        if not ok:
            bad += 1
            print("MISMATCH", (h, w, n, win, crit), "points differing:", int((q.view(np.uint32) != rq.view(np.uint32)).any(-1).sum()))
rng = np.random.default_rng(0)
for h, w, win in NOISE_CASES:
    a = rng.integers(0, 256, (h, w), dtype=np.uint8)
    b = np.roll(a, (0, 1), axis=(0, 1))
    p = np.stack([rng.uniform(-8, w + 8, 200), rng.uniform(-8, h + 8, 200)], -1).astype(np.float32).reshape(-1, 1, 2)
    q, st, er = K.calcOpticalFlowPyrLK(a, b, p, None, winSize=win, maxLevel=0, criteria=(3, 30, 0.01))
    rq, rs, re_ = cv2.calcOpticalFlowPyrLK(a, b, p, None, winSize=win, maxLevel=0, criteria=(3, 30, 0.01))
    if not (np.array_equal(q.view(np.uint32), rq.view(np.uint32)) and np.array_equal(st, rs)):
        bad += 1
        print("MISMATCH noise", (h, w, win), "points differing:", int((q.view(np.uint32) != rq.view(np.uint32)).any(-1).sum()))
print("variant", {k: v for k, v in os.environ.items() if k.startswith("KLT_LK")}, "bad cases:", bad)
sys.exit(1 if bad else 0)
