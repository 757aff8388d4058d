"""Quick GPU check of the corner-detection path against live cv2 and the oracle, plus timings."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, cv2
import visual_odom_pipeline_b200 as K
from visual_odom_pipeline_b200 import synth as S
from oracle import klt_oracle as O

rng = np.random.default_rng(0)
allok = True
for hw in [(376, 1241), (480, 640), (768, 1024), (100, 101), (57, 43), (64, 96), (2160, 3840)]:
    a = S.frame_pair(hw[0], hw[1], seed=3)[0]
    mask = np.full(a.shape, 255, np.uint8)
    for _ in range(60):
        cv2.circle(mask, (int(rng.integers(0, hw[1])), int(rng.integers(0, hw[0]))), 10, 0, -1)
    for bs in (31, 3, 5, 7, 4):
        if bs // 2 >= min(hw) or (hw[0] > 2000 and bs != 31):
            continue
        e = K.cornerMinEigenVal(a, bs)
        r = cv2.cornerMinEigenVal(a, bs, ksize=3)
        ok_e = np.array_equal(e.view(np.uint32), r.view(np.uint32))
        res = []
        for (mc, ql, md, m) in [(1000, 0.03, 10, mask), (1000, 0.03, 7, None), (0, 0.01, 3.5, mask), (50, 0.2, 0, None)]:
            c = K.goodFeaturesToTrack(a, mc, ql, md, mask=m, blockSize=bs)
            k = cv2.goodFeaturesToTrack(a, mc, ql, md, mask=m, blockSize=bs)
            same = (c is None and k is None) or (c is not None and k is not None and c.shape == k.shape and np.array_equal(c, k))
            res.append((bool(same), 0 if c is None else len(c)))
        allok = allok and ok_e and all(x[0] for x in res)
        if not ok_e:
            bad = np.argwhere(e != r)
            print("   eig mismatches:", len(bad), bad[:5], e[tuple(bad[0])], r[tuple(bad[0])])
        print(hw, "bs", bs, "eig bit-exact", ok_e, res, flush=True)
print("ALL OK" if allok else "MISMATCH")
a = S.frame_pair(376, 1241, seed=3)[0]
mask = np.full(a.shape, 255, np.uint8)
for f, name in ((lambda: K.goodFeaturesToTrack(a, 1000, 0.03, 10, mask=mask, blockSize=31), "b200 gftt"),
                (lambda: K.cornerMinEigenVal(a, 31), "b200 eig(+D2H)"),
                (lambda: cv2.goodFeaturesToTrack(a, 1000, 0.03, 10, mask=mask, blockSize=31), "cv2 gftt"),
                (lambda: cv2.cornerMinEigenVal(a, 31, ksize=3), "cv2 eig")):
    for _ in range(5): f()
    t = time.perf_counter()
    for _ in range(50): f()
    print(name, "%.1f us" % ((time.perf_counter() - t) / 50 * 1e6))
