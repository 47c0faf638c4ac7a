"""Builds the native pieces in-tree:

  roborts_edu_slam_b200/librsm.so    CUDA kernels + C ABI (nvcc, sm_100a only)
  roborts_edu_slam_b200/libsynth.so  synthetic-workload ray-caster (gcc; not on the hot path)

Run as `python -m roborts_edu_slam_b200.build` or through __graft_entry__.build().
nvcc cross-compiles sm_100a without a GPU.
"""
import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-O3", "-std=c++17", "-lineinfo",
    "-fmad=false",  # the reference is built without FMA; every a*b+c must round twice
    "-Xcompiler", "-fPIC,-ffp-contract=off,-fno-fast-math,-Wall,-Wno-unused-function",
    "-shared",
]


def _newer(target, sources):
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(s) > t for s in sources)


def _nvcc():
    return shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"


def build_rsm(force=False, verbose=False):
    out = os.path.join(HERE, "librsm.so")
    srcs = [os.path.join(CSRC, f) for f in ("rsm_kernels.cu", "rsm_score.cu", "rsm_api.cu")]
    deps = srcs + [os.path.join(CSRC, f) for f in ("rsm_device.h", "rsm_kernels.h", "rsm_host.h")] + \
        [os.path.join(os.path.dirname(HERE), "include", "rsm.h")]
    if not force and not _newer(out, deps):
        return out
    cmd = [_nvcc(), "-t", "0"] + NVCC_FLAGS + (["-Xptxas", "-v"] if verbose else []) + ["-o", out] + srcs
    print("[build]", " ".join(cmd), flush=True)
    subprocess.check_call(cmd)
    return out


def build_synth(force=False):
    out = os.path.join(HERE, "libsynth.so")
    src = os.path.join(CSRC, "synth_raycast.c")
    if not force and not _newer(out, [src]):
        return out
    cmd = ["gcc", "-O2", "-fPIC", "-shared", "-ffp-contract=off", "-o", out, src, "-lm"]
    print("[build]", " ".join(cmd), flush=True)
    subprocess.check_call(cmd)
    return out


def main():
    force = "--force" in sys.argv
    build_synth(force)
    build_rsm(force, verbose="-v" in sys.argv)


if __name__ == "__main__":
    main()
