"""Builds visco_b200/libvisco_b200.so in-tree with nvcc for sm_100a (cross-compiles without a GPU).

Each source is compiled to its own object under csrc/_obj/ (only when it or a header changed), then linked."""
import os
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
OBJ = os.path.join(CSRC, "_obj")
LIB = os.path.join(HERE, "libvisco_b200.so")
SOURCES = ["api.cu", "jacobi.cu", "stages.cu", "gram_tc.cu", "layout.cu", "cgemm_tc.cu", "topk.cu", "tridiag.cu", "recon_tc.cu",
           "tridiag_sym.cu", "tridiag_small.cu"]
CFLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17", "-Xcompiler", "-fPIC"]
CFLAGS += os.environ.get("VISCO_EXTRA_NVCC_FLAGS", "").split()   # development (e.g. -DVK_TRIDIAG_CLOCKS)
LFLAGS = ["--shared", "-cudart", "static", "-gencode", "arch=compute_100a,code=sm_100a"]


def _headers_mtime():
    hs = [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith((".cuh", ".h"))]
    hs.append(os.path.join(HERE, "..", "include", "visco_b200.h"))
    return max(os.path.getmtime(h) for h in hs)


def _stale(src, obj, hm):
    return not os.path.exists(obj) or os.path.getmtime(obj) < max(os.path.getmtime(src), hm)


def build(force=False, verbose=False):
    nvcc = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
    os.makedirs(OBJ, exist_ok=True)
    hm = _headers_mtime()
    jobs = []
    for s in SOURCES:
        src, obj = os.path.join(CSRC, s), os.path.join(OBJ, s[:-3] + ".o")
        if force or _stale(src, obj, hm):
            cmd = [nvcc] + CFLAGS + (["-Xptxas=-v"] if verbose else []) + ["-c", src, "-o", obj]
            jobs.append(cmd)
    if jobs:
        def run(cmd):
            r = subprocess.run(cmd, cwd=CSRC, capture_output=True, text=True)
            return cmd, r
        with ThreadPoolExecutor(max_workers=min(len(jobs), os.cpu_count() or 4)) as ex:
            for cmd, r in ex.map(run, jobs):
                if verbose or r.returncode:
                    sys.stderr.write(" ".join(cmd) + "\n" + r.stdout + r.stderr)
                if r.returncode:
                    raise subprocess.CalledProcessError(r.returncode, cmd)
    objs = [os.path.join(OBJ, s[:-3] + ".o") for s in SOURCES]
    if jobs or not os.path.exists(LIB) or any(os.path.getmtime(o) > os.path.getmtime(LIB) for o in objs):
        subprocess.run([nvcc] + LFLAGS + objs + ["-o", LIB], check=True, cwd=CSRC)
    return LIB


if __name__ == "__main__":
    build(force="--force" in sys.argv, verbose="-v" in sys.argv)
    print(LIB)
