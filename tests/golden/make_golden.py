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
    make_gftt()


GFTT_CASES = {
    # name: (h, w, seed, blockSize, maxCorners, qualityLevel, minDistance, n_mask_discs)
    "gftt_reference": (144, 192, 31, 31, 1000, 0.03, 10, 12),   # extractor.py:21-24 parameters, mask as in :102-107
    "gftt_default":   (120, 161, 32, 3, 0, 0.01, 3.5, 0),       # cv2 defaults: blockSize 3, no mask, no corner limit
    "gftt_even_block": (96, 128, 33, 4, 40, 0.05, 0, 6),        # even block, no minimum distance, maxCorners binding
    "gftt_block5":    (90, 70, 34, 5, 200, 0.001, 25, 3),       # fresh-sum row filter of OpenCV (block 3 / 5), large distance
}


def make_gftt():
    """Detection step: cv2.cornerMinEigenVal / cv2.goodFeaturesToTrack outputs."""
    for name, (h, w, seed, bs, mc, ql, md, discs) in GFTT_CASES.items():
        img = S.frame_pair(h, w, seed=seed)[0]
        rng = np.random.default_rng(seed)
        mask = None
        if discs:
            mask = np.full(img.shape, 255, np.uint8)
            for _ in range(discs):
                cv2.circle(mask, (int(rng.integers(0, w)), int(rng.integers(0, h))), 10, 0, -1)
        eig = cv2.cornerMinEigenVal(img, bs, ksize=3)
        c = cv2.goodFeaturesToTrack(img, mc, ql, md, mask=mask, blockSize=bs)
        c = np.zeros((0, 1, 2), np.float32) if c is None else c
        np.savez_compressed(os.path.join(HERE, name + ".npz"), img=img, mask=np.zeros((0, 0), np.uint8) if mask is None else mask,
                            eig=eig, corners=c, blockSize=bs, maxCorners=mc, qualityLevel=ql, minDistance=md,
                            cv2_version=cv2.__version__)
        print(name, "corners", len(c))


if __name__ == "__main__":
    main()
