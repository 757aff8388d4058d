"""ncu target: a few batched pyrDown launches (level 0 -> 1) on inputs larger than L2.  argv: h w B"""
import ctypes, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from visual_odom_pipeline_b200 import synth as S, tracker as T, _lib
h, w, B = (int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3])) if len(sys.argv) > 3 else (376, 1241, 310)
base = [S.texture(h, w, seed=s).astype(np.uint8) for s in range(2)]
imgs = T.alloc_image_batch(B, h, w)
for i in range(B):
    imgs[i].copy_(torch.from_numpy(np.roll(base[i % 2], 31 * (i // 2), axis=1)))
pyr = T.DevicePyramid(imgs, (21, 21), 1)
for _ in range(3):
    pyr.build()
torch.cuda.synchronize()
