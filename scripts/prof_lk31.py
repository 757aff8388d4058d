"""ncu target: batched LK launches (B=8 pairs, 16000 points, win 31 unless 'win21' is given) for old-vs-new comparisons."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from visual_odom_pipeline_b200 import synth as S, tracker as T
win = (21, 21) if "win21" in sys.argv else (31, 31)
crit = (3, 30, 0.01) if "win21" in sys.argv else (3, 30, 0.03)
B, n, h, w = 8, 2000, 376, 1241
prs = [S.frame_pair(h, w, seed=7 + i) for i in range(4)]
a = T.alloc_image_batch(B, h, w); b = T.alloc_image_batch(B, h, w)
for i in range(B):
    a[i].copy_(torch.from_numpy(prs[i % 4][0])); b[i].copy_(torch.from_numpy(prs[i % 4][1]))
pts = torch.from_numpy(np.stack([S.uniform_points(n, h, w, seed=3 + i).reshape(n, 2) for i in range(B)])).cuda()
P0 = T.DevicePyramid(a, win, 3); P1 = T.DevicePyramid(b, win, 3)
for _ in range(4):
    T.lk_track(P0, P1, pts, criteria=crit)
torch.cuda.synchronize()
