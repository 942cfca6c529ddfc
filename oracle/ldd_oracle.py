"""oracle/ldd_oracle.py -- TEST INFRASTRUCTURE ONLY: ctypes binding of oracle/ldd_oracle.c.

CPU oracle for the LDD integer structures and the accuflux family
(/root/reference/modules/routing.py:27, /root/reference/modules/advanced_routing.py:29,35,159).
Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may use it.
"""
import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))

_LIB = None


def ldd_lib():
    global _LIB
    if _LIB is not None:
        return _LIB
    src = os.path.join(_HERE, "ldd_oracle.c")
    out_dir = os.path.join(_HERE, "_build")
    so = os.path.join(out_dir, "libldd_oracle.so")
    if not os.path.exists(so) or os.path.getmtime(so) < os.path.getmtime(src):
        os.makedirs(out_dir, exist_ok=True)
        tmp = so + f".tmp{os.getpid()}"                        # several processes may build at once: publish atomically
        subprocess.check_call(["gcc", "-O2", "-fPIC", "-shared", "-ffp-contract=off", "-o", tmp, src])
        os.replace(tmp, so)
    lib = ctypes.CDLL(so)
    P = ctypes.c_void_p
    lib.ldd_compact.restype = ctypes.c_int64
    lib.ldd_compact.argtypes = [P, P, ctypes.c_int, ctypes.c_int, P]
    lib.ldd_build.restype = ctypes.c_int32
    lib.ldd_build.argtypes = [P, P, ctypes.c_int, ctypes.c_int, ctypes.c_int64] + [P] * 8
    lib.ldd_accuflux.restype = None
    lib.ldd_accuflux.argtypes = [P, P, ctypes.c_int64, P, P, P, P]
    lib.ldd_accuflux_f64.restype = None
    lib.ldd_accuflux_f64.argtypes = [P, P, ctypes.c_int64, P, P, P]
    lib.ldd_accuflux_f32inc.restype = None
    lib.ldd_accuflux_f32inc.argtypes = [P, P, ctypes.c_int64, P, P]
    lib.ldd_accucapacity.restype = None
    lib.ldd_accucapacity.argtypes = [P, P, ctypes.c_int64, P, P, P, P]
    lib.ldd_upstream.restype = None
    lib.ldd_upstream.argtypes = [P, P, ctypes.c_int64, P, P]
    _LIB = lib
    return lib


def _p(a):
    return None if a is None else a.ctypes.data_as(ctypes.c_void_p)


class LddStruct:
    """Appendix-C integer structures of one LDD raster (CPU oracle)."""

    def __init__(self, ldd_u8, active=None):
        lib = ldd_lib()
        ldd_u8 = np.ascontiguousarray(ldd_u8, dtype=np.uint8)
        self.nrows, self.ncols = ldd_u8.shape
        act = None if active is None else np.ascontiguousarray(active, dtype=np.uint8)
        self.cidx = np.empty(ldd_u8.shape, dtype=np.int32)
        n = lib.ldd_compact(_p(ldd_u8), _p(act), self.nrows, self.ncols, _p(self.cidx))
        self.n = int(n)
        self.down = np.empty(n, np.int32)
        self.indeg = np.empty(n, np.uint8)
        self.level = np.empty(n, np.int32)
        self.order = np.empty(n, np.int32)
        level_ptr = np.empty(n + 2, np.int32)
        self.don_ptr = np.empty(n + 1, np.int32)
        self.don_idx = np.empty(max(n, 1), np.int32)
        self.nup = np.empty(n, np.int64)
        nlev = lib.ldd_build(_p(ldd_u8), _p(self.cidx), self.nrows, self.ncols, n, _p(self.down),
                             _p(self.indeg), _p(self.level), _p(self.order), _p(level_ptr),
                             _p(self.don_ptr), _p(self.don_idx), _p(self.nup))
        if nlev == -1:
            raise ValueError("LDD is unsound: cycle detected")
        if nlev < 0:
            raise MemoryError("ldd_build failed")
        self.nlevels = int(nlev)
        self.level_ptr = level_ptr[: nlev + 1].copy()
        self.don_idx = self.don_idx[: int(self.don_ptr[-1])] if n else self.don_idx[:0]
        self.flat_active = np.flatnonzero(self.cidx.ravel() >= 0)

    # compact <-> raster helpers
    def gather(self, raster, dtype=np.float32):
        return np.ascontiguousarray(np.asarray(raster).ravel()[self.flat_active], dtype=dtype)

    def scatter(self, compact, fill=np.nan, dtype=np.float32):
        out = np.full(self.nrows * self.ncols, fill, dtype=dtype)
        out[self.flat_active] = compact
        return out.reshape(self.nrows, self.ncols)

    def accuflux(self, m, frac=None, want_state=False):
        lib = ldd_lib()
        m = np.ascontiguousarray(m, np.float32)
        fr = None if frac is None else np.ascontiguousarray(frac, np.float32)
        flux = np.empty(self.n, np.float32)
        state = np.empty(self.n, np.float32) if want_state else None
        lib.ldd_accuflux(_p(self.order), _p(self.down), self.n, _p(m), _p(fr), _p(flux), _p(state))
        return (flux, state) if want_state else flux

    def accucapacity(self, m, cap):
        """accucapacityflux / accucapacitystate -> (flux, state), REAL4."""
        lib = ldd_lib()
        m = np.ascontiguousarray(m, np.float32)
        cap = np.ascontiguousarray(cap, np.float32)
        flux = np.empty(self.n, np.float32)
        state = np.empty(self.n, np.float32)
        lib.ldd_accucapacity(_p(self.order), _p(self.down), self.n, _p(m), _p(cap), _p(flux), _p(state))
        return flux, state

    def subcatchment(self, points):
        """subcatchment(ldd, points) [PCR-doc]: every cell takes the id of the first non-zero point met on its way
        downstream (itself included), 0 when it drains out without meeting one."""
        pts = np.asarray(points, np.int64)
        out = np.zeros(self.n, np.int64)
        for k in range(self.n - 1, -1, -1):                  # outlets first
            i = self.order[k]
            d = self.down[i]
            out[i] = pts[i] if pts[i] != 0 else (out[d] if d >= 0 else 0)
        return out

    def streamorder(self):
        """streamorder(ldd) [PCR-doc]: Strahler order (headwater cells 1; two or more donors of the maximum order -> +1)."""
        out = np.ones(self.n, np.int64)
        for k in range(self.n):
            i = self.order[k]
            a, b = self.don_ptr[i], self.don_ptr[i + 1]
            if b > a:
                o = out[self.don_idx[a:b]]
                mx = o.max()
                out[i] = mx + 1 if (o == mx).sum() >= 2 else mx
        return out

    def accuflux_f64(self, m, frac=None):
        lib = ldd_lib()
        m = np.ascontiguousarray(m, np.float32)
        fr = None if frac is None else np.ascontiguousarray(frac, np.float32)
        flux = np.empty(self.n, np.float64)
        lib.ldd_accuflux_f64(_p(self.order), _p(self.down), self.n, _p(m), _p(fr), _p(flux))
        return flux

    def accuflux_f32inc(self, m):
        """accuflux with REAL4 adds donor by donor in routing order (the second candidate reference model)."""
        lib = ldd_lib()
        m = np.ascontiguousarray(m, np.float32)
        flux = np.empty(self.n, np.float32)
        lib.ldd_accuflux_f32inc(_p(self.order), _p(self.down), self.n, _p(m), _p(flux))
        return flux

    def upstream(self, x):
        lib = ldd_lib()
        x = np.ascontiguousarray(x, np.float32)
        out = np.empty(self.n, np.float32)
        lib.ldd_upstream(_p(self.don_ptr), _p(self.don_idx), self.n, _p(x), _p(out))
        return out
