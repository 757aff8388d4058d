"""Per-point diagnostics of the fast LK kernel (debug flag 0x100): dumps cycles / iterations / tier counts per point for
several KITTI pairs so that the launch tail can be analysed offline.  Output: gpurun_out/<tag>/lk_diag_win<W>.npz"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from visual_odom_pipeline_b200 import synth as S, tracker as T
tag = sys.argv[1] if len(sys.argv) > 1 else "diag"
out = os.path.join("gpurun_out", tag)
os.makedirs(out, exist_ok=True)
h, w, n = 376, 1241, 2000
for win, crit in (((21, 21), (3, 30, 0.01)), ((31, 31), (3, 30, 0.03))):
    rec = {}
    for k in range(6):
        pa, pb = S.frame_pair(h, w, seed=7 + k)
        a = T.alloc_image_batch(1, h, w); b = T.alloc_image_batch(1, h, w)
        a[0].copy_(torch.from_numpy(pa)); b[0].copy_(torch.from_numpy(pb))
        ph = S.uniform_points(n, h, w, seed=3 + k).reshape(1, n, 2)
        pts = torch.from_numpy(ph).cuda()
        P0 = T.DevicePyramid(a, win, 3); P1 = T.DevicePyramid(b, win, 3)
        _, st, _, it = T.lk_track(P0, P1, pts, criteria=crit, return_iters=True)
        for _ in range(2):
            _, _, _, dbg = T.lk_track(P0, P1, pts, criteria=crit, flags=0x100, return_iters=True)
        torch.cuda.synchronize()
        ts = []
        for _ in range(20):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(); T.lk_track(P0, P1, pts, criteria=crit); e1.record(); torch.cuda.synchronize()
            ts.append(e0.elapsed_time(e1))
        it = it[0].cpu().numpy(); dbg = dbg[0].cpu().numpy()
        cyc = (dbg & 0xfffff) * 64; t1 = (dbg >> 20) & 63; t2 = (dbg >> 26) & 63
        rec["pts%d" % k] = ph[0]; rec["it%d" % k] = it; rec["cyc%d" % k] = cyc; rec["t1_%d" % k] = t1; rec["t2_%d" % k] = t2
        rec["st%d" % k] = st[0].cpu().numpy(); rec["us%d" % k] = np.array(ts) * 1e3
        order = np.argsort(-cyc)[:6]
        print("win %d pair %d: launch median %.1f us; cycles mean %.0f p50 %.0f p90 %.0f p99 %.0f max %.0f; iters mean %.2f; t1 %d t2 %d of %d iters"
              % (win[0], k, np.median(ts) * 1e3, cyc.mean(), np.median(cyc), np.percentile(cyc, 90), np.percentile(cyc, 99), cyc.max(),
                 it.mean(), t1.sum(), t2.sum(), it.sum()))
        for i in order:
            print("    pt %4d (%.1f, %.1f) cycles %6d iters %3d t1 %2d t2 %2d st %d" % (i, ph[0, i, 0], ph[0, i, 1], cyc[i], it[i], t1[i], t2[i], st[0, i].item()))
    np.savez_compressed(os.path.join(out, "lk_diag_win%d.npz" % win[0]), **rec)
