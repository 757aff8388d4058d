"""Timeline of one single-pair LK launch (debug flags 0x200 / 0x400): when every point starts and ends, how many points
are resident over time, and which points end last.  usage: python scripts/lk_timeline.py [win31]"""
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
for _ in range(3):
    _, st, _, tl = T.lk_track(P0, P1, pts, criteria=crit, flags=0x200, return_iters=True)
_, _, _, lv = T.lk_track(P0, P1, pts, criteria=crit, flags=0x400, return_iters=True)
tl = tl[0].cpu().numpy().astype(np.uint32); lv = lv[0].cpu().numpy().astype(np.uint32); st = st[0].cpu().numpy()
start = (tl & 0xffff).astype(np.int64); dur = (tl >> 16).astype(np.int64)
start = (start - start.min()) & 0xffff
s_us, d_us = start * 0.032, dur * 0.032
e_us = s_us + d_us
itl = np.stack([(lv >> (8 * l)) & 0xff for l in range(4)], 1)   # per level 0..3
tot = itl.sum(1)
print("WPP", os.environ.get("KLT_LK_WPP"), "QUEUE", os.environ.get("KLT_LK_QUEUE"), "win", win)
print("kernel span %.1f us; last start %.1f us; point duration mean %.1f median %.1f p99 %.1f max %.1f us" %
      (e_us.max(), s_us.max(), d_us.mean(), np.median(d_us), np.percentile(d_us, 99), d_us.max()))
print("iters/pt mean %.2f, per level L0..L3 %s; points with a 30-iteration level: %d; max total %d" %
      (tot.mean(), np.round(itl.mean(0), 2), int((itl >= 30).any(1).sum()), tot.max()))
A = np.stack([np.ones(n), tot], 1).astype(np.float64)
coef, *_ = np.linalg.lstsq(A, d_us, rcond=None)
print("fit duration = %.2f us + %.3f us * iters" % tuple(coef))
edges = np.arange(0, e_us.max() + 5, 5.0)
print("resident points every 5 us:", [int(((s_us <= t) & (e_us > t)).sum()) for t in edges])
print("starts per 5 us:", np.histogram(s_us, edges)[0].tolist())
for i in np.argsort(-e_us)[:10]:
    print("  pt %4d start %6.1f dur %6.1f end %6.1f iters L0..L3 %s status %d" % (i, s_us[i], d_us[i], e_us[i], itl[i].tolist(), st[i]))
