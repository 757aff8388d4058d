#!/usr/bin/env python
"""bench.py -- KLT tracking hot path (pyramid build + pyramidal Lucas-Kanade) on B200.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference] [--workload kitti]

One "step" = one pass of the hot path over one synthetic frame pair of the workload: build both
Gaussian pyramids and track all keypoints coarse-to-fine (what the reference does inside ONE
cv2.calcOpticalFlowPyrLK call, src/extractor/extractor.py:44).  Default workload = BASELINE.json
configs[1]: KITTI-shape 1241x376, 2000 keypoints, winSize 21, maxLevel 3.

Printed JSON line (rank 0): `value` = tracked keypoints/s with inputs resident in HBM (device timed,
CUDA events, max over ranks); `e2e` = the same metric through the public drop-in
`calcOpticalFlowPyrLK(numpy, ...)` with pinned HOST buffers, H2D + D2H inside the timed region;
`roofline` = the dominant kernel of the step (LK) on its algorithmic bytes; `pyramid_roofline` = the
HBM-bound pyrDown kernel on a batch larger than L2; `cpu_baseline` = cv2 (the reference's own
implementation of the path) on this box's host cores.  `--impl reference` times only that cv2 path.
"""
import argparse
import ctypes
import json
import os
import statistics
import subprocess
import sys
import tempfile
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

WORKLOADS = {
    # BASELINE.json configs; criteria = cv2 defaults unless the config is the reference's own parameters
    "parking": dict(idx=0, h=480, w=640, n=500, win=(21, 21), max_level=3, criteria=(3, 30, 0.01)),
    "kitti": dict(idx=1, h=376, w=1241, n=2000, win=(21, 21), max_level=3, criteria=(3, 30, 0.01)),
    "kitti_ref_params": dict(idx=1, h=376, w=1241, n=2000, win=(31, 31), max_level=3, criteria=(3, 30, 0.03)),
    "malaga": dict(idx=2, h=768, w=1024, n=3000, win=(21, 21), max_level=3, criteria=(3, 30, 0.01)),
    "stress4k": dict(idx=4, h=2160, w=3840, n=100014, win=(31, 31), max_level=5, criteria=(3, 30, 0.01)),
}
L2_BYTES = 126 * 1024 * 1024


def describe(wl_name, wl):
    return ("configs[%d] %s: %dx%d frame pair, %d keypoints, winSize %d, maxLevel %d, criteria %s"
            % (wl["idx"], wl_name, wl["w"], wl["h"], wl["n"], wl["win"][0], wl["max_level"], tuple(wl["criteria"])))


def host_pool(wl, count, seed0):
    """`count` distinct (prev, next, pts) host triples.  A few base pairs are synthesised with the
    SURVEY s8d generator and the rest are derived by cyclic shifts (distinct bytes at distinct
    addresses is what matters for the cache behaviour)."""
    from visual_odom_pipeline_b200 import synth as S
    h, w, n = wl["h"], wl["w"], wl["n"]
    n_base = min(count, 4)
    bases = [S.frame_pair(h, w, seed=seed0 + i) for i in range(n_base)]
    pool = []
    for i in range(count):
        a, b = bases[i % n_base]
        sh = (i // n_base) * 37
        if sh:
            a, b = np.roll(a, sh, axis=1), np.roll(b, sh, axis=1)
        if wl["n"] == 100014:
            p = S.grid_points(422, 237, h, w)
        else:
            p = S.uniform_points(n, h, w, seed=seed0 + 1000 + i)
        pool.append((np.ascontiguousarray(a), np.ascontiguousarray(b), p))
    return pool


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed regions (B200_PROFILING.md)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, device_index):
        self.f = tempfile.NamedTemporaryFile("w+", suffix=".csv", delete=False)
        self.p = None
        try:
            self.p = subprocess.Popen(["nvidia-smi", "-i", str(device_index), "--query-gpu=" + self.Q, "--format=csv,noheader,nounits",
                                       "-lms", "100"], stdout=self.f, stderr=subprocess.DEVNULL)
        except Exception:
            self.p = None

    def stop(self):
        out = {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        if self.p is None:
            return out
        self.p.terminate()
        try:
            self.p.wait(timeout=5)
        except Exception:
            self.p.kill()
        self.f.flush()
        self.f.seek(0)
        sm, mx, reasons = [], [], set()
        for line in self.f.read().splitlines():
            c = [x.strip() for x in line.split(",")]
            if len(c) < 9:
                continue
            try:
                sm.append(float(c[1])); mx.append(float(c[2]))
            except ValueError:
                continue
            for name, v in zip(["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"], c[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        try:
            os.unlink(self.f.name)
        except OSError:
            pass
        if sm:
            busy = [s for s in sm if s >= 0.5 * max(sm)]
            out.update(sm_mhz=statistics.median(busy), sm_max_mhz=max(mx), reasons=sorted(reasons), samples=len(sm))
        return out


def measured_peak_gbs():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    try:
        with open(path) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md)"


def ncu_counters(report, pick_largest_grid=False):
    """Counters of the committed ncu capture of a kernel (profiles/r01/counters.json, written by
    scripts/summarize_profiles.py from `ncu --set full` runs of scripts/prof_target.py): DRAM traffic per launch and the
    utilisation figures the north star asks for.  Profiler figures, labelled as such -- never timings of this run."""
    try:
        with open(os.path.join(ROOT, "profiles", "r01", "counters.json")) as f:
            rows = json.load(f).get(report) or []
        if not rows:
            return None
        def grid_size(r):
            try:
                return int(str(r.get("grid", "0")).strip("() ").split(",")[0])
            except ValueError:
                return 0
        row = max(rows, key=grid_size) if pick_largest_grid else rows[-1]
        return row
    except Exception:
        return None


def time_cv2(wl, pool, reps, warm, threads=None):
    import cv2
    if threads is not None:
        cv2.setNumThreads(threads)
    lk = dict(winSize=wl["win"], maxLevel=wl["max_level"], criteria=wl["criteria"])
    for i in range(warm):
        a, b, p = pool[i % len(pool)]
        cv2.calcOpticalFlowPyrLK(a, b, p, None, **lk)
    ts = []
    for i in range(reps):
        a, b, p = pool[i % len(pool)]
        t = time.perf_counter()
        cv2.calcOpticalFlowPyrLK(a, b, p, None, **lk)
        ts.append(time.perf_counter() - t)
    return ts


def run_reference(args, wl_name, wl):
    """--impl reference: the reference's own CPU implementation of the path (cv2.calcOpticalFlowPyrLK as
    called at src/extractor/extractor.py:44) on this box's host cores, all threads OpenCV will use."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    import cv2
    pool = host_pool(wl, 8, seed0=7)
    per_call = time_cv2(wl, pool, 3, 2)
    est = statistics.median(per_call)
    steps = max(1, min(args.steps, int(60.0 / max(est, 1e-6))))   # bounded: at most ~1 min of CPU work
    ts = time_cv2(wl, pool, steps, args.warmup)
    total = sum(ts)
    value = wl["n"] * steps / total
    line = {
        "impl": "reference", "metric": "tracked_keypoints_per_sec", "value": value, "unit": "keypoints/s",
        "n_gpus": args.gpus, "steps": steps, "warmup": args.warmup, "ms_per_step": 1e3 * total / steps,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u8", "data": "synthetic",
        "config": {"workload": describe(wl_name, wl)},
        "pairs_per_sec": steps / total,
        "cpu_baseline": {"value": value, "unit": "keypoints/s", "cores": cv2.getNumThreads(), "kind": "reference",
                         "sample": "%d cv2.calcOpticalFlowPyrLK calls (cv2 %s, %d OpenCV threads, os.cpu_count()=%s), median %.3f ms, min %.3f ms"
                                   % (steps, cv2.__version__, cv2.getNumThreads(), os.cpu_count(), 1e3 * statistics.median(ts), 1e3 * min(ts))},
        "e2e": {"value": value, "unit": "keypoints/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line), flush=True)
    return 0


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=400)
    ap.add_argument("--warmup", type=int, default=20)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="kitti", choices=sorted(WORKLOADS))
    ap.add_argument("--pool", type=int, default=0, help="distinct device-resident pairs (0 = enough to exceed L2)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-detection", action="store_true", help="skip the Shi-Tomasi detection section (SURVEY s8f rank 2)")
    args = ap.parse_args()
    wl_name, wl = args.workload, WORKLOADS[args.workload]
    if args.warmup < 3:
        args.warmup = 3
    if args.impl == "reference":
        return run_reference(args, wl_name, wl)

    import torch
    import visual_odom_pipeline_b200 as K
    from visual_odom_pipeline_b200 import _lib, sharding
    from visual_odom_pipeline_b200.lk import make_params

    rank, local_rank, world = sharding.init_process_group()
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device -- the hot path has no CPU fallback")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    ctx = K.default_context(local_rank)
    L = _lib.load()
    stream = torch.cuda.current_stream(dev)
    sptr = ctypes.c_void_p(stream.cuda_stream)
    h, w, n = wl["h"], wl["w"], wl["n"]
    win, max_level = wl["win"], wl["max_level"]
    params = make_params(win, wl["criteria"], 0, 1e-4)

    # ---- device-resident pool of distinct pairs, larger than L2 (no flush needed between steps) --------
    # ONE batched pyramid: item 2i = prev frame of pair i, item 2i+1 = its next frame.
    pitch = (w + 127) // 128 * 128
    lay1 = _lib.klt_pyr_layout()
    assert L.klt_pyr_plan(w, h, win[0], win[1], max_level, 1, ctypes.byref(lay1)) == 0
    pair_bytes = 2 * (pitch * h + int(lay1.bytes))
    P = args.pool or max(8, -(-int(1.5 * L2_BYTES) // pair_bytes))
    P = min(P, 1024)
    layB = _lib.klt_pyr_layout()
    assert L.klt_pyr_plan(w, h, win[0], win[1], max_level, 2 * P, ctypes.byref(layB)) == 0
    layB.level[0].pitch = pitch
    layB.level[0].batch_stride = pitch * h
    n_host = min(P, 16)
    hp = host_pool(wl, n_host, seed0=7 + 100 * rank)
    imgs = torch.empty((2 * P, h, pitch), dtype=torch.uint8, device=dev)
    pyrs = torch.empty(max(int(layB.bytes), 1), dtype=torch.uint8, device=dev)
    pts = torch.empty((P, n, 2), dtype=torch.float32, device=dev)
    for i in range(P):
        a, b, p = hp[i % n_host]
        sh = (i // n_host) * 53
        imgs[2 * i, :, :w].copy_(torch.from_numpy(np.roll(a, sh, axis=1) if sh else a))
        imgs[2 * i + 1, :, :w].copy_(torch.from_numpy(np.roll(b, sh, axis=1) if sh else b))
        pts[i].copy_(torch.from_numpy(p.reshape(n, 2)))
    out_q = torch.empty((P, n, 2), dtype=torch.float32, device=dev)
    out_s = torch.empty((P, n), dtype=torch.uint8, device=dev)
    out_e = torch.empty((P, n), dtype=torch.float32, device=dev)
    iters = torch.zeros((P, n), dtype=torch.int32, device=dev)
    img0, pyr0, pts0 = imgs.data_ptr(), pyrs.data_ptr(), pts.data_ptr()
    q0, s0, e0, it0 = out_q.data_ptr(), out_s.data_ptr(), out_e.data_ptr(), iters.data_ptr()
    h_ctx = ctx.handle
    layB_ref, params_ref = ctypes.byref(layB), ctypes.byref(params)

    def pyr_step(i):
        rc = L.klt_pyr_build(h_ctx, img0, layB_ref, pyr0, 2 * i, 2, sptr)
        assert rc == 0, _lib.status_string(rc)

    def lk_step(i, want_iters=False):
        rc = L.klt_lk_track(h_ctx, img0, pyr0, img0, pyr0, layB_ref, 2 * i, 2 * i + 1, 2, 1, pts0 + i * n * 8, q0 + i * n * 8,
                            s0 + i * n, e0 + i * n * 4, (it0 + i * n * 4) if want_iters else None, n, params_ref, sptr)
        assert rc == 0, _lib.status_string(rc)

    def step(i):
        pyr_step(i)
        lk_step(i)

    launches_per_step = int(layB.top) + 1   # one pyrDown launch per level (both images batched) + one LK launch

    # ---- headline: device-resident steps ------------------------------------------------------------------
    for s in range(args.warmup):
        step(s % P)
    torch.cuda.synchronize()
    sharding.barrier()
    sampler = ClockSampler(local_rank) if rank == 0 else None
    time.sleep(0.25)
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    ev0.record(stream)
    for s in range(args.steps):
        step((args.warmup + s) % P)
    ev1.record(stream)
    torch.cuda.synchronize()
    sharding.barrier()
    dev_ms = sharding.max_over_ranks(ev0.elapsed_time(ev1))
    value = world * n * args.steps / (dev_ms * 1e-3)

    # ---- the same steps pipelined over 4 streams: independent pairs overlap, the idle tail of one pair's LK launch is
    # filled by the next pair (throughput of a job of independent pairs; per-pair latency is the number above) -------
    n_str = 4
    side = [torch.cuda.Stream(device=dev) for _ in range(n_str)]
    sptrs = [ctypes.c_void_p(st_.cuda_stream) for st_ in side]
    pe0, pe1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)

    def step_on(i, sp):
        rc = L.klt_pyr_build(h_ctx, img0, layB_ref, pyr0, 2 * i, 2, sp)
        assert rc == 0, _lib.status_string(rc)
        rc = L.klt_lk_track(h_ctx, img0, pyr0, img0, pyr0, layB_ref, 2 * i, 2 * i + 1, 2, 1, pts0 + i * n * 8, q0 + i * n * 8,
                            s0 + i * n, e0 + i * n * 4, None, n, params_ref, sp)
        assert rc == 0, _lib.status_string(rc)

    for rep in range(2):      # first pass = warm-up
        torch.cuda.synchronize()
        sharding.barrier()
        pe0.record(stream)
        for st_ in side:
            st_.wait_event(pe0)
        for s_ in range(args.steps):
            step_on((args.warmup + s_) % P, sptrs[s_ % n_str])
        for st_ in side:
            ev = torch.cuda.Event()
            ev.record(st_)
            stream.wait_event(ev)
        pe1.record(stream)
        torch.cuda.synchronize()
    pipe_ms = sharding.max_over_ranks(pe0.elapsed_time(pe1))
    pipelined = {"streams": n_str, "keypoints_per_sec": world * n * args.steps / (pipe_ms * 1e-3), "pairs_per_sec": world * args.steps / (pipe_ms * 1e-3),
                 "ms_per_step": pipe_ms / args.steps, "note": "same steps, independent pairs issued round-robin on 4 streams"}

    # ---- per-kernel durations over the same steps (events on the launching stream) ---------------------------
    ksteps = min(args.steps, 200)
    evs = [[torch.cuda.Event(enable_timing=True) for _ in range(3)] for _ in range(ksteps)]
    for s in range(ksteps):
        i = (args.warmup + s) % P
        evs[s][0].record(stream)
        pyr_step(i)
        evs[s][1].record(stream)
        lk_step(i, want_iters=True)
        evs[s][2].record(stream)
    torch.cuda.synchronize()
    pyr_ms = statistics.mean(e[0].elapsed_time(e[1]) for e in evs)
    lk_ms = statistics.mean(e[1].elapsed_time(e[2]) for e in evs)
    used = sorted(set((args.warmup + s) % P for s in range(ksteps)))
    it_mean = float(iters[used].float().mean().item())
    st_mean = float(out_s[used].float().mean().item())

    # ---- end to end through the public drop-in: pinned host buffers in, numpy out ------------------------------
    hpin = []
    for (a, b, p) in hp:
        pa, pb, pp = K.pinned_empty(a.shape, np.uint8), K.pinned_empty(b.shape, np.uint8), K.pinned_empty(p.shape, np.float32)
        pa[...] = a; pb[...] = b; pp[...] = p
        hpin.append((pa, pb, pp))
    lk_kw = dict(winSize=win, maxLevel=max_level, criteria=wl["criteria"], device=local_rank)
    e2e_steps = args.steps
    for s in range(args.warmup):
        a, b, p = hpin[s % n_host]
        K.calcOpticalFlowPyrLK(a, b, p, None, **lk_kw)
    sharding.barrier()
    t0 = time.perf_counter()
    for s in range(e2e_steps):
        a, b, p = hpin[s % n_host]
        res = K.calcOpticalFlowPyrLK(a, b, p, None, **lk_kw)
    e2e_s = sharding.max_over_ranks(time.perf_counter() - t0)
    sharding.barrier()
    e2e_value = world * n * e2e_steps / e2e_s
    clocks = sampler.stop() if sampler else None

    line = None
    if rank == 0:
        peak, peak_src = measured_peak_gbs()
        # LK algorithmic bytes per point: every window fetched once, u8:
        #   sum over levels (win+3)^2  +  iterations * (win+1)^2
        levels = int(layB.top) + 1
        lk_bytes_pt = levels * (win[0] + 3) * (win[1] + 3) + it_mean * (win[0] + 1) * (win[1] + 1)
        lk_achieved = lk_bytes_pt * n / (lk_ms * 1e-3) / 1e9
        mac_pt = win[0] * win[1] * (15 * levels + 7 * it_mean + 6)
        c_lk = ncu_counters("prof_lk") if wl_name == "kitti" else None
        roofline = {"kernel": "lk_fast_kernel (fused Scharr + pyramidal LK, all levels, 1 launch)", "bound": "hbm",
                    "achieved": lk_achieved, "peak": peak, "unit": "GB/s", "frac": lk_achieved / peak,
                    "traffic": (c_lk["dram_bytes_read"] + c_lk["dram_bytes_write"]) if c_lk else None,
                    "ncu": ({"source": "profiles/r01/counters.json (ncu --set full of the same launch shape; profiler figures)",
                             "issue_slots_active_pct": c_lk["issue_active_pct"], "shared_mem_wavefronts_pct_of_peak": c_lk["smem_wavefronts_pct_of_peak"],
                             "sm_active_fraction_of_elapsed": (c_lk["sm_active_cycles_avg"] / c_lk["sm_elapsed_cycles_max"]) if c_lk.get("sm_elapsed_cycles_max") else None}
                            if c_lk else None),
                    "peak_source": peak_src, "launch_ms": lk_ms, "algorithmic_bytes_per_launch": lk_bytes_pt * n,
                    "iters_per_point": it_mean, "int_mac_per_point": mac_pt, "gmac_per_s": mac_pt * n / (lk_ms * 1e-3) / 1e9,
                    "note": "working set is L2/L1-resident: this kernel is issue/latency-bound, not HBM-bound (see DESIGN.md, profiles/)"}

        # ---- the HBM-bound kernel: pyrDown over the whole pool (input larger than L2) ---------------------------
        nb = 2 * P
        l1 = layB.level[1] if layB.top >= 1 else None
        reps = 20
        pe = [[torch.cuda.Event(enable_timing=True) for _ in range(3)] for _ in range(reps)]
        for r in range(3 + reps):
            k = r - 3
            if k >= 0:
                pe[k][0].record(stream)
            if l1 is not None:
                rc = L.klt_pyr_down(h_ctx, img0, w, h, pitch, pitch * h, pyr0 + l1.offset, l1.pitch, l1.batch_stride, nb, sptr)
                assert rc == 0
            if k >= 0:
                pe[k][1].record(stream)
            rc = L.klt_pyr_build(h_ctx, img0, layB_ref, pyr0, 0, 0, sptr)
            assert rc == 0
            if k >= 0:
                pe[k][2].record(stream)
        torch.cuda.synchronize()
        d01_ms = statistics.mean(e[0].elapsed_time(e[1]) for e in pe)
        full_ms = statistics.mean(e[1].elapsed_time(e[2]) for e in pe)
        b01 = nb * (w * h + ((w + 1) // 2) * ((h + 1) // 2))
        ball = 0
        for l in range(int(layB.top)):
            ball += nb * (layB.level[l].w * layB.level[l].h + layB.level[l + 1].w * layB.level[l + 1].h)
        pyr_roof = {"kernel": "pyr_down_kernel level 0->1, batch of %d images (%.0f MB read, > L2)" % (nb, nb * w * h / 1e6),
                    "bound": "hbm", "achieved": b01 / (d01_ms * 1e-3) / 1e9, "peak": peak, "unit": "GB/s",
                    "frac": b01 / (d01_ms * 1e-3) / 1e9 / peak,
                    "traffic": ((lambda c: (c["dram_bytes_read"] + c["dram_bytes_write"]) if c else None)(ncu_counters("prof_pyr", True)) if (w, h) == (1241, 376) else None),
                    "traffic_note": "ncu dram__bytes_read+write of the same launch (310 KITTI images), profiles/r01/counters.json; part of the output is still in L2 at kernel end",
                    "peak_source": peak_src, "launch_ms": d01_ms,
                    "algorithmic_bytes_per_launch": b01,
                    "whole_pyramid": {"levels_built": int(layB.top), "ms": full_ms, "algorithmic_bytes": ball,
                                      "achieved": ball / (full_ms * 1e-3) / 1e9, "frac": ball / (full_ms * 1e-3) / 1e9 / peak}}

        # ---- batched LK throughput (configs[3] style: all P independent pairs in ONE launch) -----------------------
        be = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
        for r in range(3):
            if r == 2:
                be[0].record(stream)
            rc = L.klt_lk_track(h_ctx, img0, pyr0, img0, pyr0, layB_ref, 0, 1, 2, P, pts0, q0, s0, e0, None, n, params_ref, sptr)
            assert rc == 0, _lib.status_string(rc)
        be[1].record(stream)
        torch.cuda.synchronize()
        bms = be[0].elapsed_time(be[1])
        batched = {"pairs": P, "points": P * n, "ms": bms, "keypoints_per_sec": P * n / (bms * 1e-3), "pairs_per_sec": P / (bms * 1e-3)}

        # ---- next row of the scope table: Shi-Tomasi detection with the reference's parameters
        # (src/extractor/extractor.py:21-24), once per frame.  Reported next to the headline, not part of `value`. -----
        detection = None
        if not args.no_detection and h > 31 and w > 31:
            from visual_odom_pipeline_b200 import detector as D
            det_kw = dict(maxCorners=1000, qualityLevel=0.03, minDistance=10, blockSize=31)
            dimgs = [x[0] for x in hpin]
            dmask = K.pinned_empty(dimgs[0].shape, np.uint8)
            dmask[...] = 255
            for k_ in range(0, 400):   # discs around "tracked" keypoints, as extractor.py:102-107 builds the mask
                cy, cx = int(hp[0][2].reshape(-1, 2)[k_ % n][1]), int(hp[0][2].reshape(-1, 2)[k_ % n][0])
                dmask[max(cy - 7, 0):cy + 8, max(cx - 7, 0):cx + 8] = 0
            for r_ in range(5):
                got = K.goodFeaturesToTrack(dimgs[r_ % n_host], mask=dmask, device=local_rank, **det_kw)
            dreps = 200
            t0 = time.perf_counter()
            for r_ in range(dreps):
                K.goodFeaturesToTrack(dimgs[r_ % n_host], mask=dmask, device=local_rank, **det_kw)
            det_ms = 1e3 * (time.perf_counter() - t0) / dreps
            # the same step with the mask rasterised on the device from the tracked keypoints (opt-in fused call)
            tracked = K.pinned_empty((n, 2), np.float32)
            tracked[...] = hp[0][2].reshape(-1, 2)
            for r_ in range(5):
                K.detectNewFeatures(dimgs[r_ % n_host], tracked, 10, device=local_rank, **det_kw)
            t0 = time.perf_counter()
            for r_ in range(dreps):
                K.detectNewFeatures(dimgs[r_ % n_host], tracked, 10, device=local_rank, **det_kw)
            det_pts_ms = 1e3 * (time.perf_counter() - t0) / dreps
            # the whole data-parallel part of a frame chained on the device (KLTTracker.step: forward + second LK pass with
            # the reference's parameters, filter, mask, detection) against the same sequence of cv2 calls
            frame_step = None
            try:
                import cv2
                from visual_odom_pipeline_b200 import synth as S2, tracker as T2
                lkp = dict(winSize=(31, 31), maxLevel=3, criteria=(3, 30, 0.03))      # extractor.py:16-19
                seq = S2.sequence(h, w, 6, seed=11)
                dseq = [torch.from_numpy(f).to(dev) for f in seq]
                p_init = hp[0][2].reshape(-1, 2).copy()
                def run_b200(reps):
                    trk = T2.KLTTracker(**lkp).reset(dseq[0])
                    pts_d = torch.from_numpy(p_init).to(dev)
                    torch.cuda.synchronize()
                    t0 = time.perf_counter()
                    for r_ in range(reps):
                        surv, keep_, new_ = trk.step(dseq[1 + r_ % 5], pts_d)
                    torch.cuda.synchronize()
                    return (time.perf_counter() - t0) / reps, int(surv.shape[0]), 0 if new_ is None else len(new_)
                run_b200(5)
                fs_ms, n_surv, n_new = run_b200(50)
                def run_cv2(reps):
                    t0 = time.perf_counter()
                    for r_ in range(reps):
                        im0, im1 = seq[0], seq[1 + r_ % 5]
                        p0 = p_init.reshape(-1, 1, 2)
                        p1, _a, _b = cv2.calcOpticalFlowPyrLK(im0, im1, p0, None, **lkp)
                        p0r, _a, _b = cv2.calcOpticalFlowPyrLK(im0, im1, p1, None, **lkp)
                        good = abs(p0 - p0r).reshape(-1, 2).max(-1) < 30
                        q = p1.reshape(-1, 2)
                        kp_ = q[good & (0 <= q[:, 0]) & (q[:, 0] <= w) & (0 <= q[:, 1]) & (q[:, 1] <= h)]
                        m_ = np.zeros_like(im1)
                        m_[:] = 255
                        for x_, y_ in [np.int32(p_) for p_ in kp_]:
                            cv2.circle(m_, (int(x_), int(y_)), 10, 0, -1)
                        cv2.goodFeaturesToTrack(im1, mask=m_, **det_kw)
                    return (time.perf_counter() - t0) / reps
                frame_step = {"api": "KLTTracker.step(frame, points): 2 LK passes (win 31, eps 0.03), filter, mask, goodFeaturesToTrack; frames and points device-resident",
                              "ms_per_frame": 1e3 * fs_ms, "tracked": int(len(p_init)), "survivors": n_surv, "new_corners": n_new}
                if world == 1 and not args.no_cpu_baseline:
                    run_cv2(2)
                    frame_step["cv2_ms_per_frame"] = 1e3 * run_cv2(10)
                    frame_step["cv2_note"] = "same steps with cv2 on the host (%d threads), incl. the reference's Python loop over cv2.circle" % cv2.getNumThreads()
            except Exception as ex:   # pragma: no cover
                frame_step = {"error": repr(ex)}
            # device-resident, batched eigenvalue maps (the kernels only): 64 frames per launch sequence
            nbd = min(64, 2 * P)
            de = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
            for r_ in range(3):
                if r_ == 2:
                    de[0].record(stream)
                eigs = D.corner_min_eigen_val(imgs[:nbd, :, :w], 31)
            de[1].record(stream)
            torch.cuda.synchronize()
            eig_ms = de[0].elapsed_time(de[1])
            detection = {"api": "visual_odom_pipeline_b200.goodFeaturesToTrack(numpy pinned image, mask, %s) -> numpy" % det_kw,
                         "e2e_ms_per_frame": det_ms, "frames_per_sec": 1e3 / det_ms, "corners": 0 if got is None else int(len(got)),
                         "h2d_bytes_per_frame": 2 * w * h,
                         "frame_step": frame_step,
                         "fused_from_tracked_points": {"api": "visual_odom_pipeline_b200.detectNewFeatures(image, %d tracked keypoints, mask_radius=10)" % n,
                                                       "e2e_ms_per_frame": det_pts_ms, "h2d_bytes_per_frame": w * h + 8 * n,
                                                       "note": "mask of extractor.py:102-107 rasterised on the device; the reference additionally spends "
                                                               "a Python loop over cv2.circle per tracked keypoint building it on the host"},
                         "batched_min_eig": {"frames": nbd, "ms": eig_ms, "us_per_frame": 1e3 * eig_ms / nbd,
                                             "mpixels_per_sec": nbd * w * h / (eig_ms * 1e-3) / 1e6,
                                             "algorithmic_bytes_per_frame": 5 * w * h,
                                             "note": "u8 frame in, float32 eigenvalue map out; 4 launches per batch (products, running row sums, "
                                                     "running column sums + eigenvalue, all in OpenCV's summation order)"}}
            try:
                import cv2
                got = K.goodFeaturesToTrack(dimgs[0], mask=dmask, device=local_rank, **det_kw)
                ref = cv2.goodFeaturesToTrack(np.array(dimgs[0]), mask=np.array(dmask), **det_kw)
                detection["parity"] = {"corners_identical_to_cv2": bool((got is None and ref is None) or (got is not None and ref is not None
                                                                                                      and got.shape == ref.shape and np.array_equal(got, ref)))}
                if world == 1 and not args.no_cpu_baseline:
                    ci, cm = np.array(dimgs[0]), np.array(dmask)
                    for r_ in range(3):
                        cv2.goodFeaturesToTrack(ci, mask=cm, **det_kw)
                    t0 = time.perf_counter()
                    for r_ in range(40):
                        cv2.goodFeaturesToTrack(ci, mask=cm, **det_kw)
                    detection["cv2_ms_per_frame"] = 1e3 * (time.perf_counter() - t0) / 40
                    detection["cv2_threads"] = cv2.getNumThreads()
            except Exception as ex:   # pragma: no cover
                detection["parity"] = {"error": repr(ex)}

        # ---- parity spot check of pool entry used first, against live cv2 ------------------------------------
        parity = None
        try:
            import cv2
            i = args.warmup % P
            a, b, p = hp[i % n_host]
            sh = (i // n_host) * 53
            a2, b2 = (np.roll(a, sh, axis=1), np.roll(b, sh, axis=1)) if sh else (a, b)
            rq, rs, re_ = cv2.calcOpticalFlowPyrLK(np.ascontiguousarray(a2), np.ascontiguousarray(b2), p, None, winSize=win, maxLevel=max_level,
                                                   criteria=wl["criteria"])
            gq, gs = out_q[i].cpu().numpy(), out_s[i].cpu().numpy()
            both = (rs.ravel() == 1) & (gs == 1)
            parity = {"points": int(n), "status_agree": float((rs.ravel() == gs).mean()),
                      "max_abs_diff_px": float(np.abs(rq.reshape(-1, 2)[both] - gq[both]).max()) if both.any() else 0.0,
                      "bit_exact": bool(np.array_equal(rq.reshape(-1, 2).view(np.uint32), gq.view(np.uint32)))}
        except Exception as ex:   # pragma: no cover
            parity = {"error": repr(ex)}

        cpu = None
        if world == 1 and not args.no_cpu_baseline:
            import cv2
            hp_cpu = [(np.array(a), np.array(b), np.array(p)) for (a, b, p) in hp[:8]]
            est = statistics.median(time_cv2(wl, hp_cpu, 3, 2))
            reps_cpu = max(5, min(2000, int(12.0 / max(est, 1e-6))))
            ts = time_cv2(wl, hp_cpu, reps_cpu, 3)
            ts1 = time_cv2(wl, hp_cpu, max(3, min(200, int(4.0 / max(est * cv2.getNumThreads(), 1e-6)))), 1, threads=1)
            cv2.setNumThreads(-1)
            cpu = {"value": n * len(ts) / sum(ts), "unit": "keypoints/s", "cores": cv2.getNumThreads(), "kind": "reference",
                   "sample": "%d cv2.calcOpticalFlowPyrLK calls on the same workload (cv2 %s, %d OpenCV threads of os.cpu_count()=%s): "
                             "median %.3f ms, min %.3f ms per call; single thread: median %.3f ms"
                             % (len(ts), cv2.__version__, cv2.getNumThreads(), os.cpu_count(), 1e3 * statistics.median(ts), 1e3 * min(ts),
                                1e3 * statistics.median(ts1)),
                   "ms_per_call_median": 1e3 * statistics.median(ts), "ms_per_call_single_thread": 1e3 * statistics.median(ts1)}

        line = {
            "metric": "tracked_keypoints_per_sec", "value": value, "unit": "keypoints/s", "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": dev_ms / args.steps, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "u8", "data": "synthetic",
            "config": {"workload": describe(wl_name, wl),
                       "l2": "rotating pool of %d distinct device-resident pairs (%.0f MB) > 126 MB L2; no flush needed" % (P, P * pair_bytes / 1e6),
                       "sharding": "each rank tracks its own independent sequences; no collective on the data path"},
            "pairs_per_sec": world * args.steps / (dev_ms * 1e-3),
            "status1_fraction": st_mean,
            "kernel_ms": {"pyramid_build_both_images": pyr_ms, "lk": lk_ms},
            "e2e": {"value": e2e_value, "unit": "keypoints/s", "h2d_bytes_per_step": 2 * w * h + n * 8, "d2h_bytes_per_step": n * 13,
                    "ms_per_step": 1e3 * e2e_s / e2e_steps, "pairs_per_sec": world * e2e_steps / e2e_s,
                    "api": "visual_odom_pipeline_b200.calcOpticalFlowPyrLK(numpy pinned host arrays) -> numpy"},
            "gpu_launches": launches_per_step * args.steps,
            "clocks": clocks,
            "roofline": roofline,
            "pyramid_roofline": pyr_roof,
            "pipelined": pipelined,
            "batched_lk": batched,
            "parity": parity,
            "detection": detection,
            "cpu_baseline": cpu,
            "device": ctx.name,
        }
        print(json.dumps(line), flush=True)
    sharding.barrier()
    if world > 1:
        import torch.distributed as dist
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
