"""Quick GPU parity sweep (debug aid): CUDA path vs live cv2 and vs the C oracle."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import cv2
import visual_odom_pipeline_b200 as K
from visual_odom_pipeline_b200 import synth as S
from oracle import klt_oracle as O

bad = 0

def cmp(name, h, w, n, win, L, crit, motion=S.BENIGN, margin=0, flags=0, **kw):
    global bad
    a, b = S.frame_pair(h, w, motion=motion, **kw)
    p = S.uniform_points(n, h, w, margin=margin)
    q1, s1, e1 = cv2.calcOpticalFlowPyrLK(a, b, p, None, winSize=win, maxLevel=L, criteria=crit, flags=flags)
    t = time.time(); q2, s2, e2 = K.calcOpticalFlowPyrLK(a, b, p, None, winSize=win, maxLevel=L, criteria=crit, flags=flags); dt = time.time() - t
    dp = (q1.view(np.uint32) != q2.view(np.uint32)).any(-1).ravel()
    ds = (s1 != s2).ravel()
    m = ((s1 == 1) & (s2 == 1)).ravel()
    de = (e1.view(np.uint32) != e2.view(np.uint32)).ravel() & m
    ok = not (dp.any() or ds.any() or de.any())
    bad += not ok
    print(f"{'OK ' if ok else 'BAD'} {name}: pts differ {dp.sum()} (max {np.nanmax(np.abs(q1 - q2)):.4g}) status differ {ds.sum()} err differ {de.sum()} st1={s1.mean():.3f} t={dt*1e3:.2f}ms", flush=True)
    if not ok:
        idx = np.nonzero(dp | ds | de)[0][:5]
        for i in idx:
            print("   pt", i, p.reshape(-1, 2)[i], "cv2", q1.reshape(-1, 2)[i], s1[i], e1[i], "klt", q2.reshape(-1, 2)[i], s2[i], e2[i])

ctx = K.default_context(0)
print("device", ctx.name, ctx.sm_count, ctx.cc)
# pyramid
for (h, w) in [(376, 1241), (480, 640), (768, 1024), (135, 241), (50, 70), (33, 17), (1080, 1920)]:
    a = S.texture(h, w, seed=h).astype(np.uint8)
    top, lv = K.buildOpticalFlowPyramid(a, (5, 5), 6)
    ref = [a]
    for i in range(top): ref.append(cv2.pyrDown(ref[-1]))
    oks = [np.array_equal(x, y) for x, y in zip(lv, ref)]
    bad += not all(oks)
    print("pyramid", (h, w), "top", top, oks, flush=True)
cmp("tiny", 50, 70, 100, (21, 21), 3, (3, 30, 0.01), margin=30)
cmp("parking", 480, 640, 500, (21, 21), 3, (3, 30, 0.01))
cmp("kitti21", 376, 1241, 2000, (21, 21), 3, (3, 30, 0.01))
cmp("kitti21 again", 376, 1241, 2000, (21, 21), 3, (3, 30, 0.01))
cmp("kitti31", 376, 1241, 2000, (31, 31), 3, (3, 30, 0.03))
cmp("kitti hard", 376, 1241, 2000, (21, 21), 3, (3, 30, 0.01), motion=S.HARD, margin=60, noise_sigma=3.0, flat_cols=(400, 700))
cmp("kitti hard31", 376, 1241, 2000, (31, 31), 3, (3, 30, 0.03), motion=S.HARD, margin=60, noise_sigma=3.0, flat_cols=(400, 700))
cmp("malaga", 768, 1024, 3000, (21, 21), 3, (3, 30, 0.01))
for win in [(5, 5), (7, 7), (24, 24), (20, 12), (13, 29), (3, 3), (8, 8), (9, 16), (40, 40), (64, 64)]:
    cmp(f"win{win}", 240, 320, 400, win, 3, (3, 30, 0.01), motion=S.HARD, margin=30)
for crit in [(1, 10, 0.01), (2, 30, 0.05), (3, 200, 20.0), (3, 0, 0.01), (0, 5, 0.5), (3, -3, -1.0)]:
    cmp(f"crit{crit}", 240, 320, 400, (21, 21), 3, crit, margin=30)
for L in [0, 1, 5, 8]:
    cmp(f"maxLevel{L}", 376, 1241, 500, (21, 21), L, (3, 30, 0.01), margin=30)
cmp("mineig flag", 240, 320, 400, (21, 21), 3, (3, 30, 0.01), flags=8, margin=30)
cmp("4k", 2160, 3840, 20000, (31, 31), 5, (3, 30, 0.01))
print("TOTAL BAD", bad)
