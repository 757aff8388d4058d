"""Timeline of one single-pair LK launch (debug flag 0x400: per-point start / end stamps from %globaltimer)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from visual_odom_pipeline_b200 import synth as S, tracker as T
win = (31, 31) if "win31" in sys.argv else (21, 21)
crit = (3, 30, 0.03) if "win31" in sys.argv else (3, 30, 0.01)
h, w, n = 376, 1241, 2000
for k in range(2):
    pa, pb = S.frame_pair(h, w, seed=7 + k)
    a = T.alloc_image_batch(1, h, w); b = T.alloc_image_batch(1, h, w)
    a[0].copy_(torch.from_numpy(pa)); b[0].copy_(torch.from_numpy(pb))
    pts = torch.from_numpy(S.uniform_points(n, h, w, seed=3 + k).reshape(1, n, 2)).cuda()
    P0 = T.DevicePyramid(a, win, 3); P1 = T.DevicePyramid(b, win, 3)
    _, st, _, it = T.lk_track(P0, P1, pts, criteria=crit, return_iters=True)
    for _ in range(3):
        _, _, _, dbg = T.lk_track(P0, P1, pts, criteria=crit, flags=0x400, return_iters=True)
    torch.cuda.synchronize()
    ts = []
    for _ in range(10):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); T.lk_track(P0, P1, pts, criteria=crit); e1.record(); torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1) * 1e3)
    d = dbg[0].cpu().numpy(); it = it[0].cpu().numpy()
    t0 = (d & 0x7fff).astype(np.int64); t1 = ((d >> 15) & 0x7fff).astype(np.int64); lg = (d >> 30) & 1
    base = t0.min()
    t0 = ((t0 - base) % 32768) * 0.128; t1 = ((t1 - base) % 32768) * 0.128
    bulk = lg == 0
    print("pair %d: event time median %.1f us; stamps: span %.1f us; bulk points %d: last start %.1f, last end %.1f; long points %d: first start %.1f, last start %.1f, last end %.1f"
          % (k, np.median(ts), t1.max(), bulk.sum(), t0[bulk].max(), t1[bulk].max(), (~bulk).sum(), t0[~bulk].min() if (~bulk).any() else -1,
             t0[~bulk].max() if (~bulk).any() else -1, t1[~bulk].max() if (~bulk).any() else -1))
    hist, edges = np.histogram(t1[bulk], bins=np.arange(0, t1.max() + 10, 10))
    print("   bulk end-time histogram (10 us bins):", hist.tolist())
    order = np.argsort(-t1)[:8]
    for i in order:
        print("   pt %4d long=%d start %.1f end %.1f (%.1f us) iters %d" % (i, lg[i], t0[i], t1[i], t1[i] - t0[i], it[i]))
