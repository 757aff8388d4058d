"""Generates the committed golden vectors from the reference's own implementation of the path:
the `cv2` module that src/extractor/extractor.py imports (cv2 4.13.0 here).  Run once, in this
container:  python tests/golden/make_golden.py
Each .npz holds the inputs and cv2's outputs; err is zeroed where status == 0 (cv2 leaves it
uninitialised there).
"""
import os
import sys

import cv2
import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
from visual_odom_pipeline_b200 import synth as S  # noqa: E402

LK_CASES = {
    # name: (h, w, n, win, maxLevel, criteria, flags, motion, margin, noise, flat)
    "lk_default":     (120, 160, 96, (21, 21), 3, (3, 30, 0.01), 0, S.BENIGN, 0, 0.0, None),
    "lk_reference":   (144, 192, 96, (31, 31), 3, (3, 30, 0.03), 0, S.BENIGN, 25, 0.0, None),  # extractor.py:16-19
    "lk_hard":        (120, 200, 128, (13, 9), 2, (3, 30, 0.01), 0, S.HARD, 30, 3.0, (60, 110)),
    "lk_count_only":  (96, 128, 64, (5, 5), 4, (1, 10, 0.0), 0, S.BENIGN, 10, 0.0, None),
    "lk_eps_only":    (96, 128, 64, (8, 24), 1, (2, 0, 0.05), 0, S.HARD, 10, 1.0, None),
    "lk_mineig_flag": (96, 128, 64, (21, 21), 2, (3, 30, 0.01), 8, S.BENIGN, 10, 0.0, (30, 60)),
    "lk_level0":      (96, 128, 64, (15, 15), 0, (3, 30, 0.01), 0, S.BENIGN, 10, 0.0, None),
}


def main():
    for name, (h, w, n, win, lvl, crit, flags, motion, margin, noise, flat) in LK_CASES.items():
        a, b = S.frame_pair(h, w, seed=11, motion=motion, noise_sigma=noise, flat_cols=flat)
        p = S.uniform_points(n, h, w, seed=5, margin=margin)
        q, st, er = cv2.calcOpticalFlowPyrLK(a, b, p, None, winSize=win, maxLevel=lvl, criteria=crit, flags=flags)
        er = np.where(st == 1, er, 0).astype(np.float32)
        np.savez_compressed(os.path.join(HERE, name + ".npz"), prev=a, next=b, prevPts=p, nextPts=q, status=st, err=er,
                            winSize=np.array(win), maxLevel=lvl, criteria=np.array(crit, np.float64), flags=flags,
                            cv2_version=cv2.__version__)
        print(name, "status mean %.3f" % st.mean())
    img = S.texture(135, 241, seed=21).astype(np.uint8)
    levels = [img]
    for _ in range(4):
        levels.append(cv2.pyrDown(levels[-1]))
    top, pyr = cv2.buildOpticalFlowPyramid(img, (21, 21), 8, None, False)
    np.savez_compressed(os.path.join(HERE, "pyramid.npz"), img=img, l1=levels[1], l2=levels[2], l3=levels[3], l4=levels[4],
                        top_win21_max8=top, scharr_x=cv2.Scharr(img, cv2.CV_16S, 1, 0), scharr_y=cv2.Scharr(img, cv2.CV_16S, 0, 1),
                        cv2_version=cv2.__version__)
    print("pyramid top", top)


if __name__ == "__main__":
    main()
