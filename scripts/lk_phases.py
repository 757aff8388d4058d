"""Cycles per phase of the iteration loop of the longest points of one single-pair LK launch.

Needs the diagnostics build:  python visual-odom-pipeline_b200/build.py --variant phases --extra -DKLT_LK_PHASES
and  KLT_LIB_PATH=visual-odom-pipeline_b200/lib/libklt_b200_phases.so  (scripts/gpu_lk_phases.sh does both)."""
import ctypes, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from visual_odom_pipeline_b200 import _lib, synth as S, tracker as T
NAMES = ["loop top (range, region, weights)", "per-pixel pass", "tier 0 sums", "tier 1 sums + tests", "replay stores", "barrier 1",
         "replay, warp 0 chains", "barrier 2 (tail chains)", "result loads + combine", "solve + tests"]
win = (31, 31) if "win31" in sys.argv else (21, 21)
crit = (3, 30, 0.03) if "win31" in sys.argv else (3, 30, 0.01)
h, w, n = 376, 1241, 2000
L = _lib.load()
L.klt_debug_lk_phase_select.argtypes = [ctypes.c_longlong]
L.klt_debug_lk_phase_read.argtypes = [ctypes.c_void_p]
for k in range(2):
    pa, pb = S.frame_pair(h, w, seed=7 + k)
    a = T.alloc_image_batch(1, h, w); b = T.alloc_image_batch(1, h, w)
    a[0].copy_(torch.from_numpy(pa)); b[0].copy_(torch.from_numpy(pb))
    pts = torch.from_numpy(S.uniform_points(n, h, w, seed=3 + k).reshape(1, n, 2)).cuda()
    P0 = T.DevicePyramid(a, win, 3); P1 = T.DevicePyramid(b, win, 3)
    _, st, _, it = T.lk_track(P0, P1, pts, criteria=crit, return_iters=True)
    for _ in range(2):
        _, _, _, dbg = T.lk_track(P0, P1, pts, criteria=crit, flags=0x100, return_iters=True)
    torch.cuda.synchronize()
    dbg = dbg[0].cpu().numpy(); it = it[0].cpu().numpy()
    cyc = (dbg & 0xfffff) * 64
    for i in np.argsort(-cyc)[:3].tolist() + [int(np.argsort(cyc)[n // 2])]:
        assert L.klt_debug_lk_phase_select(i) == 0
        T.lk_track(P0, P1, pts, criteria=crit)
        torch.cuda.synchronize()
        out = (ctypes.c_ulonglong * 16)()
        assert L.klt_debug_lk_phase_read(out) == 0
        ph = np.array(out[:10], dtype=np.float64); n_it = out[15]; n_rep = out[14]
        print("pair %d pt %4d: %d cycles in the launch under flag 0x100, %d iterations, %d replays; stamped %.0f cycles"
              % (k, i, cyc[i], n_it, n_rep, ph.sum()))
        for name, c in zip(NAMES, ph):
            print("    %-36s %8.0f total  %6.0f per iteration" % (name, c, c / max(n_it, 1)))
    L.klt_debug_lk_phase_select(-1)
