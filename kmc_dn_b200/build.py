"""In-tree build of libkmcb200.so for sm_100a with plain nvcc (no torch involved).

    python -m kmc_dn_b200.build [--force] [--verbose]

The .so lands next to this file (git-ignored, but it travels to the GPU box with gpurun).
"""
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
OBJ = os.path.join(HERE, "build")
SO = os.path.join(HERE, "libkmcb200.so")

ARCH = ["-gencode", "arch=compute_100a,code=sm_100a"]
COMMON = ["-O3", "-std=c++17", "-lineinfo", "-Xcompiler", "-fPIC"]
# per-file extra flags: the replay kernels must not contract a*b+c (Go on amd64 and numba never do)
UNITS = {
    "hop_fast.cu": [],
    "hop_memo.cu": [],
    "hop_lanes.cu": [],
    "hop_wide.cu": [],
    "hop_reforder.cu": [],
    "hop_exact.cu": ["-fmad=false", "-prec-div=true", "-prec-sqrt=true"],
    "hop_prob.cu": ["-fmad=false", "-prec-div=true", "-prec-sqrt=true"],
    "kmc_api.cu": [],
    "peaks.cu": [],
    "reduce.cu": [],
}
DEPS = ["kmc_internal.cuh", "kmc_device.cuh", "memo_common.cuh", os.path.join("..", "..", "include", "kmc_b200.h")]


def _nvcc():
    for c in (os.environ.get("NVCC"), "/usr/local/cuda/bin/nvcc", "nvcc"):
        if c and (os.path.isabs(c) and os.path.exists(c) or not os.path.isabs(c)):
            return c
    raise RuntimeError("nvcc not found")


def _stale(target, sources):
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(s) > t for s in sources)


def build(force=False, verbose=False):
    os.makedirs(OBJ, exist_ok=True)
    nvcc = _nvcc()
    deps = [os.path.join(CSRC, d) for d in DEPS] + [os.path.abspath(__file__)]
    tune = os.environ.get("KMCB200_NVCC_FLAGS", "").split()  # tuning experiments only, e.g. -DLANES_MIN_CTAS=7
    objs, jobs = [], []
    for src, extra in UNITS.items():
        s = os.path.join(CSRC, src)
        o = os.path.join(OBJ, src.replace(".cu", ".o"))
        objs.append(o)
        if force or _stale(o, [s] + deps):
            jobs.append([nvcc] + ARCH + COMMON + extra + tune + (["-Xptxas", "-v"] if verbose else []) + ["-c", s, "-o", o])
    if jobs:  # the translation units are independent: compile them side by side
        from concurrent.futures import ThreadPoolExecutor

        def run(cmd):
            if verbose:
                print(" ".join(cmd), flush=True)
            r = subprocess.run(cmd, capture_output=True, text=True)
            return cmd, r
        with ThreadPoolExecutor(max_workers=min(len(jobs), os.cpu_count() or 1)) as ex:
            results = list(ex.map(run, jobs))
        for cmd, r in results:
            if verbose or r.returncode:
                sys.stdout.write(r.stdout)
                sys.stderr.write(r.stderr)
            if r.returncode:
                raise subprocess.CalledProcessError(r.returncode, cmd)
    if force or _stale(SO, objs):
        cmd = [nvcc] + ARCH + ["-shared", "-o", SO] + objs + ["-Xlinker", "--exclude-libs=ALL"]
        if verbose:
            print(" ".join(cmd))
        subprocess.check_call(cmd)
    return SO


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="--verbose" in sys.argv))
