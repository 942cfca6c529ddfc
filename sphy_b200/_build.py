"""In-tree native builds: the sm_100a CUDA library behind the C-ABI and the host-side generator.

Everything is built next to the sources (the .so files travel to the GPU box with the snapshot;
they are git-ignored).  No JIT cache, no multi-arch: sm_100a only.
"""
import os
import shutil
import subprocess

PKG = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(PKG)
CSRC = os.path.join(PKG, "csrc")
INCLUDE = os.path.join(ROOT, "include")

CUDA_LIB = os.path.join(PKG, "libsphy_b200.so")
HOSTGEN_LIB = os.path.join(PKG, "_hostgen.so")

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-O3", "-lineinfo", "-std=c++17",
    "-fmad=false",            # parity: the reference rounds every REAL4 operation (no FMA contraction)
    "-prec-div=true", "-prec-sqrt=true", "-ftz=false",
    "-Xcompiler", "-fPIC", "-shared",
]


def _newer(target, sources):
    if not os.path.exists(target):
        return False
    t = os.path.getmtime(target)
    return all(os.path.getmtime(s) <= t for s in sources)


def cuda_sources():
    srcs = sorted(os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith(".cu"))
    deps = sorted(os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith((".cuh", ".h")))
    deps += sorted(os.path.join(INCLUDE, f) for f in os.listdir(INCLUDE) if f.endswith(".h"))
    return srcs, deps


def build_cuda(force=False, verbose=False):
    """Compile every csrc/*.cu to an object (in parallel) and link libsphy_b200.so."""
    from concurrent.futures import ThreadPoolExecutor

    srcs, deps = cuda_sources()
    if not force and _newer(CUDA_LIB, srcs + deps):
        return CUDA_LIB
    nvcc = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not os.path.exists(nvcc):
        raise RuntimeError("nvcc not found: cannot build libsphy_b200.so (no CPU fallback exists)")
    objdir = os.path.join(PKG, "build")
    os.makedirs(objdir, exist_ok=True)
    cflags = [f for f in NVCC_FLAGS if f != "-shared"]

    def compile_one(src):
        obj = os.path.join(objdir, os.path.basename(src)[:-3] + ".o")
        if not force and _newer(obj, [src] + deps):
            return obj, ""
        cmd = [nvcc] + cflags + (["-Xptxas=-v"] if verbose else []) + ["-I", INCLUDE, "-I", CSRC, "-c", "-o", obj, src]
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError("nvcc failed for %s:\n%s" % (src, r.stderr))
        return obj, r.stderr

    with ThreadPoolExecutor(max_workers=min(8, len(srcs))) as ex:
        results = list(ex.map(compile_one, srcs))
    if verbose:
        for _, log in results:
            if log:
                print(log)
    objs = [o for o, _ in results]
    subprocess.check_call([nvcc, "-gencode", "arch=compute_100a,code=sm_100a", "-shared", "-o", CUDA_LIB] + objs)
    return CUDA_LIB


def build_hostgen(force=False):
    src = os.path.join(CSRC, "hostgen.c")
    if not force and _newer(HOSTGEN_LIB, [src]):
        return HOSTGEN_LIB
    subprocess.check_call(["gcc", "-O2", "-fPIC", "-shared", "-o", HOSTGEN_LIB, src, "-lm"])
    return HOSTGEN_LIB


def build_all(force=False):
    build_hostgen(force)
    build_cuda(force)
