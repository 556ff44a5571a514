"""Compile csrc/*.cu into libtkbo.so (sm_100a only), in-tree, with plain nvcc.

No torch / pybind involvement: the library's surface is the C ABI of include/tkbo.h and Python binds it
with ctypes (``_native.py``).  nvcc cross-compiles without a GPU, so this runs in the build container;
the resulting .so travels to the GPU box with the repo snapshot.
"""
from __future__ import annotations

import concurrent.futures
import hashlib
import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
BUILD = os.path.join(HERE, "build")
LIB = os.path.join(HERE, "libtkbo.so")

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-O3", "-lineinfo", "-std=c++17",
    "-Xcompiler", "-fPIC",
    "--expt-relaxed-constexpr",
    "-Xptxas", "-v",
]


def nvcc_path():
    for cand in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found: libtkbo.so cannot be built (there is no CPU fallback)")


def sources():
    return sorted(os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith(".cu"))


def _digest():
    h = hashlib.sha256()
    for root in (CSRC, os.path.join(os.path.dirname(HERE), "include")):
        for f in sorted(os.listdir(root)):
            if f.endswith((".cu", ".cuh", ".h")):
                with open(os.path.join(root, f), "rb") as fh:
                    h.update(f.encode())
                    h.update(fh.read())
    h.update(" ".join(NVCC_FLAGS).encode())
    return h.hexdigest()


def _compile_one(nvcc, src, extra=(), suffix=""):
    obj = os.path.join(BUILD, os.path.basename(src)[:-3] + suffix + ".o")
    cmd = [nvcc, *NVCC_FLAGS, *extra, "-c", src, "-o", obj]
    r = subprocess.run(cmd, capture_output=True, text=True)
    log = r.stdout + r.stderr
    with open(obj + ".log", "w") as fh:
        fh.write(" ".join(cmd) + "\n" + log)
    if r.returncode != 0:
        raise RuntimeError(f"nvcc failed on {src}:\n{log}")
    return obj


def build(force=False, verbose=True):
    """Build libtkbo.so if missing or stale. Returns its path."""
    os.makedirs(BUILD, exist_ok=True)
    stamp = os.path.join(BUILD, "stamp")
    digest = _digest()
    if not force and os.path.exists(LIB) and os.path.exists(stamp) and open(stamp).read() == digest:
        return LIB
    nvcc = nvcc_path()
    srcs = sources()
    if verbose:
        print(f"[tkbo] compiling {len(srcs)} sources for sm_100a with {nvcc}", file=sys.stderr)
    with concurrent.futures.ThreadPoolExecutor(max_workers=min(8, len(srcs))) as ex:
        objs = list(ex.map(lambda s: _compile_one(nvcc, s), srcs))
    cmd = [nvcc, "-shared", "-o", LIB, *objs, "-lcudart"]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError("link failed:\n" + r.stdout + r.stderr)
    with open(stamp, "w") as fh:
        fh.write(digest)
    return LIB


def build_trace():
    """libtkbo_trace.so: the same library with the K1 pipeline trace compiled in (-DTKBO_TRACE); diagnostics only."""
    os.makedirs(BUILD, exist_ok=True)
    nvcc = nvcc_path()
    with concurrent.futures.ThreadPoolExecutor(max_workers=8) as ex:
        objs = list(ex.map(lambda s: _compile_one(nvcc, s, ("-DTKBO_TRACE",), "_trace"), sources()))
    out = os.path.join(HERE, "libtkbo_trace.so")
    r = subprocess.run([nvcc, "-shared", "-o", out, *objs, "-lcudart"], capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError("link failed:\n" + r.stdout + r.stderr)
    return out


def build_variant(name, defines):
    """libtkbo_<name>.so: the same library compiled with extra -D defines (tuning sweeps, e.g. TKBO_POLY_MASK=0x1111u);
    select it at run time with TKBO_LIB=<path>."""
    os.makedirs(BUILD, exist_ok=True)
    nvcc = nvcc_path()
    extra = tuple("-D" + d for d in defines)
    with concurrent.futures.ThreadPoolExecutor(max_workers=8) as ex:
        objs = list(ex.map(lambda s: _compile_one(nvcc, s, extra, "_" + name), sources()))
    out = os.path.join(HERE, f"libtkbo_{name}.so")
    r = subprocess.run([nvcc, "-shared", "-o", out, *objs, "-lcudart"], capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError("link failed:\n" + r.stdout + r.stderr)
    return out


if __name__ == "__main__":
    print(build(force="--force" in sys.argv))
    if "--trace" in sys.argv:
        print(build_trace())
    for a in sys.argv[1:]:
        if a.startswith("--variant="):          # --variant=name:DEF1=x,DEF2=y
            vname, _, defs = a[len("--variant="):].partition(":")
            print(build_variant(vname, [d for d in defs.split(",") if d]))
