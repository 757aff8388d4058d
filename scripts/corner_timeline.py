"""Chain-warp timeline of row_scan_kernel (profiling aid): where one chunk's cycles go."""
import ctypes, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import visual_odom_pipeline_b200 as K
from visual_odom_pipeline_b200 import _lib, synth as S
L = _lib.load()
L.klt_debug_corner_timeline.argtypes = [ctypes.c_int, ctypes.c_void_p, ctypes.c_int]
a = S.frame_pair(376, 1241, seed=3)[0]
for _ in range(3): K.cornerMinEigenVal(a, 31)
mode = 2 if 'col' in sys.argv else 1   # 1 = row_scan_kernel, 2 = col_scan_kernel
assert L.klt_debug_corner_timeline(mode, None, 0) == 0
K.cornerMinEigenVal(a, 31)
buf = np.zeros(2560, np.int64)
assert L.klt_debug_corner_timeline(0, buf.ctypes.data, 2560) == 0
hb = buf[1024:].reshape(-1, 6)
t = buf[:1024].reshape(-1, 4)
t0 = t[0, 0]
nh = int((hb[:, 0] != 0).sum())
print("helper warp 0, per chunk: start, convert, arrive+fetch issue, wait s-full, copy out, arrive (cycles)")
for k in range(nh):
    h = hb[k]
    print("%3d: %7d  %5d %5d %5d %5d %5d" % (k, h[0] - t0, h[1] - h[0], h[2] - h[1], (h[3] - h[2]) if h[3] else 0, (h[4] - h[3]) if h[3] else 0, h[5] - (h[4] if h[4] else h[2])))
n = int((t[:, 0] != 0).sum())
t = t[:n] - t[0, 0]
print("chunk: start, wait d-full, wait s-empty, compute+arrive (cycles)")
for k in range(n):
    print("%3d: %7d  %6d %6d %6d" % (k, t[k, 0], t[k, 1] - t[k, 0], t[k, 2] - t[k, 1], t[k, 3] - t[k, 2]))
print("total", t[n - 1, 3], "cycles for", n, "chunks", flush=True)
