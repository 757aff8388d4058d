"""Builds libklt_b200.so in-tree with nvcc for sm_100a (no JIT cache: the .so travels with the repo).

    python visual-odom-pipeline_b200/build.py [--force] [--verbose]
"""
import os
import subprocess
import sys

PKG = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(PKG)
CSRC = os.path.join(PKG, "csrc")
LIB_DIR = os.path.join(PKG, "lib")
LIB = os.path.join(LIB_DIR, "libklt_b200.so")
SOURCES = ["klt_pyramid.cu", "klt_lk.cu", "klt_lk_fast.cu", "klt_lk_warp.cu", "klt_filter.cu", "klt_corners.cu", "klt_capi.cu"]
HEADERS = [os.path.join(CSRC, "klt_common.cuh"), os.path.join(ROOT, "include", "klt_b200.h")]

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",   # B200 only; no PTX for other archs, no fallback
    "-O3", "-std=c++17", "-lineinfo",
    "-fmad=false",            # OpenCV's x86 build rounds every float op once
    "-prec-div=true", "-prec-sqrt=true",
    "-Xcompiler", "-fPIC,-O2,-Wall",
    "-cudart", "static",
]


def nvcc():
    for cand in (os.environ.get("NVCC"), "/usr/local/cuda/bin/nvcc", "nvcc"):
        if cand and (os.path.sep not in cand or os.path.exists(cand)):
            return cand
    raise RuntimeError("nvcc not found")


def needs_build():
    if not os.path.exists(LIB):
        return True
    t = os.path.getmtime(LIB)
    deps = [os.path.join(CSRC, s) for s in SOURCES] + HEADERS + [os.path.abspath(__file__)]
    return any(os.path.getmtime(d) > t for d in deps)


def build(force=False, verbose=False):
    if not force and not needs_build():
        return LIB
    os.makedirs(LIB_DIR, exist_ok=True)
    cmd = [nvcc()] + NVCC_FLAGS + ["-I", os.path.join(ROOT, "include"), "-I", CSRC, "-shared", "-o", LIB]
    if verbose:
        cmd += ["-Xptxas", "-v"]
    cmd += os.environ.get("KLT_NVCC_EXTRA", "").split()   # e.g. -DKLT_LK_TIMELINE for scripts/lk_timeline.py / lk_phases.py
    cmd += [os.path.join(CSRC, s) for s in SOURCES]
    res = subprocess.run(cmd, capture_output=True, text=True)
    if verbose or res.returncode != 0:
        sys.stderr.write(res.stdout + res.stderr)
    if res.returncode != 0:
        raise RuntimeError("nvcc failed building libklt_b200.so")
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="--verbose" in sys.argv))
