"""Cycles per phase of the resume-team iteration (debug flag 0x200, phase selector in flag bits 12..14) for the longest points."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from visual_odom_pipeline_b200 import synth as S, tracker as T
win = (31, 31) if "win31" in sys.argv else (21, 21)
crit = (3, 30, 0.03) if "win31" in sys.argv else (3, 30, 0.01)
h, w, n = 376, 1241, 2000
pa, pb = S.frame_pair(h, w, seed=7)
a = T.alloc_image_batch(1, h, w); b = T.alloc_image_batch(1, h, w)
a[0].copy_(torch.from_numpy(pa)); b[0].copy_(torch.from_numpy(pb))
pts = torch.from_numpy(S.uniform_points(n, h, w, seed=3).reshape(1, n, 2)).cuda()
P0 = T.DevicePyramid(a, win, 3); P1 = T.DevicePyramid(b, win, 3)
_, st, _, it = T.lk_track(P0, P1, pts, criteria=crit, return_iters=True)
it = it[0].cpu().numpy()
top = np.argsort(-it)[:5]
names = ["top->pixels", "pixels+stores", "barrier 1", "replay", "barrier 2", "solve+tests"]
rows = []
for ph in range(6):
    _, _, _, dbg = T.lk_track(P0, P1, pts, criteria=crit, flags=0x200 | (ph << 12), return_iters=True)
    rows.append((dbg[0].cpu().numpy() & 0xfffffff) * 4)
_, _, _, dbg = T.lk_track(P0, P1, pts, criteria=crit, flags=0x100, return_iters=True)
dbg = dbg[0].cpu().numpy(); cyc = (dbg & 0xfffff) * 64; t2 = (dbg >> 26) & 63
for i in top:
    print("pt %d iters %d, resume-team iterations %d, resume cycles %d" % (i, it[i], t2[i], cyc[i]))
    for ph in range(6):
        print("   %-14s %7d cycles total, %6.0f per iteration" % (names[ph], rows[ph][i], rows[ph][i] / max(t2[i], 1)))
