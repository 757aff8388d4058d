"""Times the corner-detection call (host API) for ncu launch lists / wall-clock breakdowns."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import visual_odom_pipeline_b200 as K
from visual_odom_pipeline_b200 import synth as S
h, w = (2160, 3840) if "4k" in sys.argv else (376, 1241)
a = S.frame_pair(h, w, seed=3)[0]
mask = np.full(a.shape, 255, np.uint8)
reps = int(os.environ.get("REPS", "20"))
for _ in range(3): K.goodFeaturesToTrack(a, 1000, 0.03, 10, mask=mask, blockSize=31)
t = time.perf_counter()
for _ in range(reps): c = K.goodFeaturesToTrack(a, 1000, 0.03, 10, mask=mask, blockSize=31)
print("gftt %dx%d: %.1f us per call, %d corners" % (w, h, (time.perf_counter() - t) / reps * 1e6, len(c)))
