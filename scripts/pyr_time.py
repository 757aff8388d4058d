"""Times the batched pyrDown kernel (level 0->1 and the whole pyramid) on inputs larger than L2; checks vs the oracle."""
import ctypes, os, sys, statistics, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import visual_odom_pipeline_b200 as K
from visual_odom_pipeline_b200 import synth as S, tracker as T, _lib
from oracle import klt_oracle as O

PEAK = json.load(open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "MEASURED_PEAKS.json")))["hbm_gbs"]

def run(h, w, B, win=(21, 21), max_level=3, reps=20):
    base = [S.texture(h, w, seed=s).astype(np.uint8) for s in range(3)]
    imgs = T.alloc_image_batch(B, h, w)
    for i in range(B):
        imgs[i].copy_(torch.from_numpy(np.roll(base[i % 3], 31 * (i // 3), axis=1)))
    pyr = T.DevicePyramid(imgs, win, max_level)
    torch.cuda.synchronize()
    L = _lib.load(); lay = pyr.layout; st = ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)
    l1 = lay.level[1]
    def down():
        assert L.klt_pyr_down(pyr.ctx.handle, imgs.data_ptr(), w, h, imgs.stride(1), imgs.stride(0), pyr.buffer.data_ptr() + l1.offset,
                              l1.pitch, l1.batch_stride, B, st) == 0
    for _ in range(3): down(); pyr.build()
    # all repetitions are queued back to back and synchronised once: an event recorded on an idle stream would also time
    # the CPU's launch path (~5 us through ctypes), which is not kernel time
    evs = []
    for _ in range(reps):
        e = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
        e[0].record(); down(); e[1].record(); pyr.build(); e[2].record()
        evs.append(e)
    torch.cuda.synchronize()
    t1 = [e[0].elapsed_time(e[1]) for e in evs]
    t2 = [e[1].elapsed_time(e[2]) for e in evs]
    b01 = B * (w * h + ((w + 1) // 2) * ((h + 1) // 2)); ball = pyr.algorithmic_bytes()
    m1, m2 = statistics.median(t1), statistics.median(t2)
    ok = True
    for b in (0, B // 2, B - 1):
        _, ol = O.build_pyramid(np.roll(base[b % 3], 31 * (b // 3), axis=1), win, max_level)
        for l in range(1, pyr.top + 1):
            ok = ok and np.array_equal(pyr.level(l)[b].cpu().numpy(), ol[l])
    print(f"{w}x{h} B={B}: level0->1 {m1*1e3:.1f} us = {b01/m1/1e6:.0f} GB/s ({b01/m1/1e6/PEAK*100:.1f}% of measured {PEAK:.0f}); "
          f"whole pyramid ({pyr.top} levels) {m2*1e3:.1f} us = {ball/m2/1e6:.0f} GB/s ({ball/m2/1e6/PEAK*100:.1f}%); bit-exact={ok}", flush=True)

if "single" in sys.argv:      # the single-pair shape only (latency of the one-launch build; env switches: DESIGN s10)
    run(376, 1241, 2, reps=100)
    sys.exit(0)
run(376, 1241, 310)
run(376, 1241, 2)
run(480, 640, 600)
run(768, 1024, 256)
run(2160, 3840, 24, win=(31, 31), max_level=5)
