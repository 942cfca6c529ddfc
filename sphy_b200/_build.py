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


def _digest(sources, flags=()):
    import hashlib
    h = hashlib.sha256(" ".join(flags).encode())
    for s in sorted(sources):
        h.update(os.path.basename(s).encode())
        with open(s, "rb") as fh:
            h.update(fh.read())
    return h.hexdigest()


def _newer(target, sources, flags=()):
    """Is `target` the build of exactly these sources?  Decided by a content digest kept next to it -- file times do not
    survive the copy onto a GPU box, and a spurious rebuild there costs half a minute per process."""
    stamp = target + ".stamp"
    if not os.path.exists(target) or not os.path.exists(stamp):
        return False
    with open(stamp) as fh:
        return fh.read().strip() == _digest(sources, flags)


def _stamp(target, sources, flags=()):
    with open(target + ".stamp", "w") as fh:
        fh.write(_digest(sources, flags))


def cuda_sources():
    srcs = sorted(os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith(".cu"))
    deps = sorted(os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith((".cuh", ".h")))
    deps += sorted(os.path.join(INCLUDE, f) for f in os.listdir(INCLUDE) if f.endswith(".h"))
    return srcs, deps


class _BuildLock:
    """Inter-process lock around the in-tree builds: the ranks of a torchrun launch all import the package at once, and a
    rank must never dlopen a library another rank is still linking."""

    def __init__(self, name):
        self.path = os.path.join(PKG, name)

    def __enter__(self):
        import fcntl
        self.fh = open(self.path, "w")
        fcntl.flock(self.fh, fcntl.LOCK_EX)
        return self

    def __exit__(self, *exc):
        import fcntl
        fcntl.flock(self.fh, fcntl.LOCK_UN)
        self.fh.close()


def build_cuda(force=False, verbose=False):
    """Compile every csrc/*.cu to an object (in parallel) and link libsphy_b200.so (atomically, under a file lock)."""
    srcs, deps = cuda_sources()
    if not force and _newer(CUDA_LIB, srcs + deps, NVCC_FLAGS):
        return CUDA_LIB
    with _BuildLock(".build.lock"):
        if not force and _newer(CUDA_LIB, srcs + deps, NVCC_FLAGS):    # another process built it while this one waited
            return CUDA_LIB
        return _build_cuda_locked(srcs, deps, force, verbose)


def _build_cuda_locked(srcs, deps, force, verbose):
    from concurrent.futures import ThreadPoolExecutor

    nvcc = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not os.path.exists(nvcc):
        raise RuntimeError("nvcc not found: cannot build libsphy_b200.so (no CPU fallback exists)")
    objdir = os.path.join(PKG, "build")
    os.makedirs(objdir, exist_ok=True)
    cflags = [f for f in NVCC_FLAGS if f != "-shared"]

    def compile_one(src):
        obj = os.path.join(objdir, os.path.basename(src)[:-3] + ".o")
        if not force and _newer(obj, [src] + deps, cflags):
            return obj, ""
        cmd = [nvcc] + cflags + (["-Xptxas=-v"] if verbose else []) + ["-I", INCLUDE, "-I", CSRC, "-c", "-o", obj, src]
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError("nvcc failed for %s:\n%s" % (src, r.stderr))
        _stamp(obj, [src] + deps, cflags)
        return obj, r.stderr

    with ThreadPoolExecutor(max_workers=min(8, len(srcs))) as ex:
        results = list(ex.map(compile_one, srcs))
    if verbose:
        for _, log in results:
            if log:
                print(log)
    objs = [o for o, _ in results]
    tmp = CUDA_LIB + f".tmp{os.getpid()}"
    subprocess.check_call([nvcc, "-gencode", "arch=compute_100a,code=sm_100a", "-shared", "-o", tmp] + objs)
    os.replace(tmp, CUDA_LIB)                                  # readers see the old library or the complete new one
    _stamp(CUDA_LIB, srcs + deps, NVCC_FLAGS)
    return CUDA_LIB


def build_hostgen(force=False):
    src = os.path.join(CSRC, "hostgen.c")
    if not force and _newer(HOSTGEN_LIB, [src]):
        return HOSTGEN_LIB
    with _BuildLock(".build.lock"):
        if not force and _newer(HOSTGEN_LIB, [src]):
            return HOSTGEN_LIB
        tmp = HOSTGEN_LIB + f".tmp{os.getpid()}"
        subprocess.check_call(["gcc", "-O2", "-fPIC", "-shared", "-o", tmp, src, "-lm"])
        os.replace(tmp, HOSTGEN_LIB)
        _stamp(HOSTGEN_LIB, [src])
    return HOSTGEN_LIB


def build_all(force=False):
    build_hostgen(force)
    build_cuda(force)
