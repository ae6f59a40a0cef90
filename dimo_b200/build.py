"""Builds dimo_b200/lib/libdimo_b200.so from dimo_b200/csrc/*.cu with nvcc for sm_100a.

In-tree, no torch headers, plain C ABI (include/dimo_b200.h).  `python -m dimo_b200.build`
or `__graft_entry__.build()`.  Objects are rebuilt only when a source or header is newer.
"""
import os
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
CSRC = os.path.join(HERE, "csrc")
LIBDIR = os.path.join(HERE, "lib")
OBJDIR = os.path.join(HERE, "build")
LIB = os.path.join(LIBDIR, "libdimo_b200.so")

ARCH = ["-gencode", "arch=compute_100a,code=sm_100a"]
COMMON = ["-O3", "-std=c++17", "-lineinfo", "-Xcompiler", "-fPIC",
          "-I", os.path.join(ROOT, "include"), "--expt-relaxed-constexpr"]

# translation unit -> extra flags.  The two "index path" units are compiled without FMA
# contraction so their fp32 arithmetic is reproducible op-for-op by the oracle (bit-exact
# radii / tile rectangles / sort keys / neighbour indices).
SOURCES = {
    "raster_preprocess.cu": ["-fmad=false"],
    "raster_preprocess_bwd.cu": [],
    "raster_bin.cu": [],
    "raster_blend.cu": [],
    "knn.cu": ["-fmad=false"],
    "points.cu": ["-fmad=false"],
    "arap.cu": ["-fmad=false"],
    "deform.cu": [],
    "mlp.cu": [],
    "mlp_tc.cu": [],
    "timenet_tc.cu": [],
    "ssim.cu": [],
    "optim.cu": [],
    "smooth.cu": [],
    "gtcache.cu": [],
}


def _nvcc():
    for c in (os.environ.get("NVCC"), "/usr/local/cuda/bin/nvcc", "nvcc"):
        if c and (os.path.isabs(c) and os.path.exists(c) or not os.path.isabs(c)):
            return c
    raise RuntimeError("nvcc not found")


def _newer(src_list, target):
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(s) > t for s in src_list)


def build(verbose=False, force=False):
    os.makedirs(LIBDIR, exist_ok=True)
    os.makedirs(OBJDIR, exist_ok=True)
    nvcc = _nvcc()
    headers = [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith((".cuh", ".h"))]
    headers.append(os.path.join(ROOT, "include", "dimo_b200.h"))
    headers.append(os.path.abspath(__file__))
    jobs = []
    objs = []
    for src, extra in SOURCES.items():
        sp = os.path.join(CSRC, src)
        if not os.path.exists(sp):
            if os.environ.get("DIMO_ALLOW_MISSING"):
                continue
            raise RuntimeError(f"missing source {sp}")
        obj = os.path.join(OBJDIR, src.replace(".cu", ".o"))
        objs.append(obj)
        if force or _newer([sp] + headers, obj):
            cmd = [nvcc] + ARCH + COMMON + extra + (["-Xptxas", "-v"] if verbose else []) + ["-c", sp, "-o", obj]
            jobs.append(cmd)

    def run(cmd):
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError("nvcc failed:\n" + " ".join(cmd) + "\n" + r.stdout + r.stderr)
        return r.stderr

    if jobs:
        with ThreadPoolExecutor(max_workers=min(8, len(jobs))) as ex:
            for log in ex.map(run, jobs):
                if verbose and log:
                    print(log)
    if jobs or not os.path.exists(LIB):
        cmd = [nvcc] + ARCH + ["-shared", "-o", LIB] + objs + ["-lcudart"]
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError("link failed:\n" + r.stdout + r.stderr)
    return LIB


if __name__ == "__main__":
    print(build(verbose="-v" in sys.argv, force="-f" in sys.argv))
