import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import visual_odom_pipeline_b200 as K
from visual_odom_pipeline_b200 import synth as S
a, b = S.frame_pair(376, 1241, seed=7); p = S.uniform_points(2000, 376, 1241, seed=3)
pa, pb, pp = K.pinned_empty(a.shape), K.pinned_empty(b.shape), K.pinned_empty(p.shape, np.float32)
pa[...] = a; pb[...] = b; pp[...] = p
for name, (x, y, z) in {"pinned": (pa, pb, pp), "pageable": (a, b, p)}.items():
    for _ in range(5): K.calcOpticalFlowPyrLK(x, y, z, None)
    ts = []
    for _ in range(50):
        t = time.perf_counter(); K.calcOpticalFlowPyrLK(x, y, z, None); ts.append(time.perf_counter() - t)
    print(name, "median %.1f us min %.1f us" % (np.median(ts) * 1e6, min(ts) * 1e6), file=sys.stderr)
