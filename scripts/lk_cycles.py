"""Per-point latency model of the fast LK kernel (debug flag 0x100): cycles vs iterations / fallbacks."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from visual_odom_pipeline_b200 import synth as S, tracker as T
win = (31, 31) if "win31" in sys.argv else (21, 21)
h, w, n = 376, 1241, 2000
pa, pb = S.frame_pair(h, w, seed=7)
a = T.alloc_image_batch(1, h, w); b = T.alloc_image_batch(1, h, w)
a[0].copy_(torch.from_numpy(pa)); b[0].copy_(torch.from_numpy(pb))
pts = torch.from_numpy(S.uniform_points(n, h, w, seed=3).reshape(1, n, 2)).cuda()
P0 = T.DevicePyramid(a, win, 3); P1 = T.DevicePyramid(b, win, 3)
crit = (3, 30, 0.01)
_, st, _, it = T.lk_track(P0, P1, pts, criteria=crit, return_iters=True)
for _ in range(2):
    _, _, _, dbg = T.lk_track(P0, P1, pts, criteria=crit, flags=0x100, return_iters=True)
it = it[0].cpu().numpy(); dbg = dbg[0].cpu().numpy()
cyc = (dbg & 0xfffff) * 64; t1 = (dbg >> 20) & 63; t2 = (dbg >> 26) & 63
print("WPP", os.environ.get("KLT_LK_WPP"), "win", win)
print("cycles: mean %.0f median %.0f p99 %.0f max %.0f" % (cyc.mean(), np.median(cyc), np.percentile(cyc, 99), cyc.max()))
A = np.stack([np.ones_like(it), it, t1, t2], 1).astype(np.float64)
coef, *_ = np.linalg.lstsq(A, cyc.astype(np.float64), rcond=None)
print("fit cycles = %.0f + %.0f*iters + %.0f*tier1 + %.0f*tier2" % tuple(coef))
print("iters mean %.2f max %d; tier1 total %d (%.1f%% of iters) tier2 total %d (%.1f%%)" % (it.mean(), it.max(), t1.sum(), 100 * t1.sum() / it.sum(), t2.sum(), 100 * t2.sum() / it.sum()))
top = np.argsort(-cyc)[:8]
for i in top: print("  pt %d cycles %d iters %d tier1 %d tier2 %d status %d" % (i, cyc[i], it[i], t1[i], t2[i], st[0, i].item()))
