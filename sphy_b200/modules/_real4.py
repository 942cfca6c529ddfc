"""REAL4 map arithmetic for the init-time parameter derivations of the process modules.

The reference evaluates its init() expressions with PCRaster: Python numbers become non-spatial REAL4 values,
every map operation rounds to REAL4, transcendentals are evaluated in double on the REAL4 operand (assumptions A2/A3,
DESIGN.md).  These helpers restate that for float32 NumPy arrays; they run once per model, never per time step."""
import numpy as np

f32 = np.float32


def r4(x):
    if isinstance(x, np.ndarray):
        return x if x.dtype == np.float32 else x.astype(np.float32)
    return f32(x)


def _dbl(fn, x):
    with np.errstate(all="ignore"):
        return np.asarray(fn(np.asarray(r4(x), np.float64))).astype(np.float32)[()]


def r4_exp(x): return _dbl(np.exp, x)
def r4_sin(x): return _dbl(np.sin, x)
def r4_cos(x): return _dbl(np.cos, x)


def r4_pow(x, y):
    with np.errstate(all="ignore"):
        return np.asarray(np.power(np.asarray(r4(x), np.float64), np.asarray(r4(y), np.float64))).astype(np.float32)[()]


def r4_div(a, b):
    """x / 0 is a missing value in PCRaster."""
    a, b = r4(a), r4(b)
    with np.errstate(all="ignore"):
        q = np.asarray(a / b, np.float32)
        return np.where(np.broadcast_to(b, q.shape) == 0, f32(np.nan), q)[()]


def r4_if(cond, a, b):
    return np.where(cond, r4(a), r4(b)).astype(np.float32)
