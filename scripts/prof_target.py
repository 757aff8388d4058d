"""Small deterministic workload for ncu: a few single-pair KITTI LK calls and a batched pyramid build."""
import ctypes, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import visual_odom_pipeline_b200 as K
from visual_odom_pipeline_b200 import synth as S, tracker as T

mode = sys.argv[1] if len(sys.argv) > 1 else "all"
win = (31, 31) if "win31" in sys.argv else (21, 21)
if mode in ("all", "lk"):
    a, b = S.frame_pair(376, 1241, seed=7)
    p = S.uniform_points(2000, 376, 1241, seed=3)
    for _ in range(3):
        K.calcOpticalFlowPyrLK(a, b, p, None, winSize=win, maxLevel=3, criteria=(3, 30, 0.01))
if mode in ("all", "pyr", "batch"):
    # 155 KITTI-sized pairs = 310 images, as bench.py's pool: crops of wide synthetic canvases at their own offsets
    P = 155
    canv = [S.frame_pair(376, 1241 + 31 * (P // 4), seed=7 + s) for s in range(4)]
    imgs = T.alloc_image_batch(2 * P, 376, 1241)
    prev, nxt_imgs = T.alloc_image_batch(P, 376, 1241), T.alloc_image_batch(P, 376, 1241)
    for i in range(P):
        off = 31 * (i // 4)
        a = torch.from_numpy(np.ascontiguousarray(canv[i % 4][0][:, off:off + 1241]))
        b = torch.from_numpy(np.ascontiguousarray(canv[i % 4][1][:, off:off + 1241]))
        imgs[2 * i].copy_(a); imgs[2 * i + 1].copy_(b)
        prev[i].copy_(a); nxt_imgs[i].copy_(b)
    torch.cuda.synchronize()
    pyr = T.DevicePyramid(imgs, win, 3)
    for _ in range(2):
        pyr.build()
    torch.cuda.synchronize()
    if mode in ("all", "batch"):
        p0, p1 = T.DevicePyramid(prev, win, 3), T.DevicePyramid(nxt_imgs, win, 3)
        pts = torch.from_numpy(np.stack([S.uniform_points(2000, 376, 1241, seed=1007 + i).reshape(-1, 2) for i in range(16)])).cuda()
        pts = pts.repeat(P // 16 + 1, 1, 1)[:P].contiguous()
        torch.cuda.synchronize()
        T.lk_track(p0, p1, pts, criteria=(3, 30, 0.03 if win[0] == 31 else 0.01))
        torch.cuda.synchronize()
