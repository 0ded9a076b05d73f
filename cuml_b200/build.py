"""Build the C-ABI shared library in-tree with nvcc for sm_100a (no CMake, no network).

    python -m cuml_b200.build            # builds cuml_b200/lib/libcuml_b200.so if stale

The .so is git-ignored but travels to the GPU box with the gpurun snapshot.
"""
from __future__ import annotations

import os
import shutil
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
CSRC = os.path.join(HERE, "csrc")
LIBDIR = os.path.join(HERE, "lib")
OBJDIR = os.path.join(HERE, "build")
LIBNAME = "libcuml_b200.so"

SOURCES = ["handle.cu", "peer_comm.cu", "distance_simt.cu", "centroid_update.cu", "fused_l2_argmin_sm100.cu", "centroid_update_tma.cu", "seeding.cu",
           "kmeans_api.cu"]

NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17",
              "-Xcompiler", "-fPIC,-fvisibility=hidden", "--expt-relaxed-constexpr",
              "-Xptxas", "-v"]


def _nvcc():
    return shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"


def lib_path():
    # CUML_B200_LIB: load another build of the library (A/B measurements of compile-time variants)
    return os.environ.get("CUML_B200_LIB") or os.path.join(LIBDIR, LIBNAME)


def _deps():
    out = [os.path.join(ROOT, "include", "cuml_b200", "kmeans_c.h")]
    for f in os.listdir(CSRC):
        out.append(os.path.join(CSRC, f))
    return out


def is_stale():
    lib = lib_path()
    if not os.path.exists(lib):
        return True
    t = os.path.getmtime(lib)
    return any(os.path.getmtime(p) > t for p in _deps())


def build(force=False, verbose=False):
    if not force and not is_stale():
        return lib_path()
    os.makedirs(LIBDIR, exist_ok=True)
    os.makedirs(OBJDIR, exist_ok=True)
    nvcc = _nvcc()
    headers_mtime = max(os.path.getmtime(p) for p in _deps() if p.endswith((".cuh", ".h")))

    def compile_one(src):
        obj = os.path.join(OBJDIR, src.replace(".cu", ".o"))
        srcp = os.path.join(CSRC, src)
        if (not force and os.path.exists(obj) and os.path.getmtime(obj) > os.path.getmtime(srcp)
                and os.path.getmtime(obj) > headers_mtime):
            return obj, ""
        cmd = [nvcc, *NVCC_FLAGS, "-c", srcp, "-o", obj]
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError("nvcc failed for %s:\n%s\n%s" % (src, r.stdout, r.stderr))
        return obj, r.stderr

    with ThreadPoolExecutor(max_workers=min(8, len(SOURCES))) as ex:
        results = list(ex.map(compile_one, SOURCES))
    objs = [o for o, _ in results]
    if verbose:
        for _, log in results:
            if log:
                print(log)
    with open(os.path.join(OBJDIR, "ptxas.log"), "w") as f:
        for _, log in results:
            f.write(log)
    cmd = [nvcc, "-shared", "-o", lib_path(), *objs, "-gencode", "arch=compute_100a,code=sm_100a",
           "-Xlinker", "--no-undefined", "-ldl"]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError("link failed:\n%s\n%s" % (r.stdout, r.stderr))
    return lib_path()


if __name__ == "__main__":
    p = build(force="--force" in sys.argv, verbose=True)
    print("built", p)
