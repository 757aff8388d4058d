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
    B = 310
    base = [S.texture(376, 1241, seed=s).astype(np.uint8) for s in range(4)]
    imgs = T.alloc_image_batch(B, 376, 1241)
    for i in range(B):
        imgs[i].copy_(torch.from_numpy(np.roll(base[i % 4], 31 * (i // 4), axis=1)))
    torch.cuda.synchronize()
    pyr = T.DevicePyramid(imgs, win, 3)
    for _ in range(2):
        pyr.build()
    torch.cuda.synchronize()
    if mode in ("all", "batch"):
        nxt = T.DevicePyramid(imgs.roll(1, 0).contiguous() if False else imgs, win, 3)
        pts = torch.from_numpy(np.stack([S.uniform_points(2000, 376, 1241, seed=i).reshape(-1, 2) for i in range(8)])).cuda()
        pts = pts.repeat(B // 8 + 1, 1, 1)[:B].contiguous()
        T.lk_track(pyr, nxt, pts)
        torch.cuda.synchronize()
