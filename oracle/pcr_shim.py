"""oracle/pcr_shim.py -- TEST INFRASTRUCTURE ONLY (never imported by the product path).

A float32 NumPy stand-in for the `pcraster` / `pcraster.framework` modules, exposing exactly the
operations the in-scope reference files call, so that the UNMODIFIED reference
(`/root/reference/sphy.py` and the modules it imports) can be executed in this container, where
PCRaster itself is absent (third-party C++ dependency, version unpinned by the reference, no wheel,
no network).  It is used only by `oracle/run_reference.py` to generate the committed golden vectors
under `tests/golden/` and to validate the travelling restatement `oracle/sphy_port.py`.

PARITY UNPINNED at the PCRaster boundary: the reference ships no tests, fixtures or golden outputs
(SURVEY.md section 4), so the semantics below are restated from PCRaster's public documentation
 and could not be checked against a real PCRaster build:
  * Scalar fields are REAL4; every operator rounds its result to float32; MV is NaN here.
  * Python numbers mixed with fields become REAL4 non-spatial fields (so `pcr.cos(2*pi*J/365)` first
    rounds its argument to float32).
  * x/0 -> MV.  Any MV operand -> MV (ifthenelse: only condition + selected branch).
  * Assumption A2: transcendental functions (exp, sin, cos, tan, acos, **) are evaluated in double
    on the REAL4 operand(s) and the result is stored as REAL4.
  * Assumption A1: accuflux-family per-cell sums are formed in double, stored REAL4 (ldd_oracle.c).
  * `generateNameT("dir/prec", 1)` -> "dir/prec0000.001".
"""
from __future__ import annotations

import os
import types

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))

F32 = np.float32
NAN = np.float32(np.nan)
MV_U1 = np.uint8(255)
MV_I4 = np.int32(-2147483648)

Scalar, Nominal, Boolean, Ldd, Ordinal, Directional = "scalar", "nominal", "boolean", "ldd", "ordinal", "directional"


# ----------------------------------------------------------------------------------------------
# global state (PCRaster is full of it: clone, global options)
# ----------------------------------------------------------------------------------------------
class _Globals:
    def __init__(self):
        self.reset()

    def reset(self):
        self.nrows = 0
        self.ncols = 0
        self.cellsize = 1.0
        self.radians = False
        self.matrixtable = False
        self.maps = {}          # normalised path -> ndarray | Field | callable() -> ndarray
        self.reports = {}       # name -> {timestep -> ndarray}
        self.tss = {}           # name -> TimeoutputTimeseries
        self.step_hook = None
        self.dtype = np.float32  # float64 twin for error attribution: set to np.float64
        self._ldd_cache = {}


G = _Globals()


def _norm(path):
    return os.path.normpath(str(path))


def register_map(path, value):
    """value: ndarray (dtype decides the value scale), Field, or a zero-arg callable returning one."""
    G.maps[_norm(path)] = value


from .ldd_oracle import LddStruct  # noqa: E402  (C helper shared with the travelling port)


# ----------------------------------------------------------------------------------------------
# Field
# ----------------------------------------------------------------------------------------------
class Field:
    """A PCRaster map value: spatial (nrows, ncols) or non-spatial (0-d)."""

    __slots__ = ("a", "vs")
    __array_ufunc__ = None       # make numpy scalars defer to our reflected operators

    def __init__(self, a, vs=Scalar):
        self.a = a
        self.vs = vs

    def isSpatial(self):
        return self.a.ndim > 0

    def dataType(self):
        return self.vs

    def __repr__(self):
        return f"Field<{self.vs},{'spatial' if self.a.ndim else 'nonspatial'}>({self.a!r})"

    # arithmetic -----------------------------------------------------------------------------
    def __add__(self, o): return _arith(np.add, self, o)
    def __radd__(self, o): return _arith(np.add, o, self)
    def __sub__(self, o): return _arith(np.subtract, self, o)
    def __rsub__(self, o): return _arith(np.subtract, o, self)
    def __mul__(self, o): return _arith(np.multiply, self, o)
    def __rmul__(self, o): return _arith(np.multiply, o, self)
    def __truediv__(self, o): return _divide(self, o)
    def __rtruediv__(self, o): return _divide(o, self)
    def __pow__(self, o): return _power(self, o)
    def __rpow__(self, o): return _power(o, self)
    def __neg__(self): return Field((-_sc(self)).astype(G.dtype), Scalar)
    def __pos__(self): return self
    def __abs__(self): return Field(np.abs(_sc(self)), Scalar)

    # comparisons ----------------------------------------------------------------------------
    def __gt__(self, o): return _compare(np.greater, self, o)
    def __ge__(self, o): return _compare(np.greater_equal, self, o)
    def __lt__(self, o): return _compare(np.less, self, o)
    def __le__(self, o): return _compare(np.less_equal, self, o)
    def __eq__(self, o): return _compare(np.equal, self, o)
    def __ne__(self, o): return _compare(np.not_equal, self, o)
    __hash__ = None

    # boolean --------------------------------------------------------------------------------
    def __and__(self, o): return pcrand(self, o)
    def __rand__(self, o): return pcrand(o, self)
    def __or__(self, o): return pcror(self, o)
    def __ror__(self, o): return pcror(o, self)
    def __invert__(self): return pcrnot(self)

    def __bool__(self):
        if self.a.ndim:
            raise ValueError("truth value of a spatial field is ambiguous")
        return bool(_sc(self) != 0)

    def __float__(self):
        if self.a.ndim:
            raise TypeError("cannot convert a spatial field to float")
        return float(_sc(self))


def _is_field(x):
    return isinstance(x, Field)


def _sc(x):
    """Operand -> REAL4 array (0-d for non-spatials and Python numbers), NaN = MV."""
    if isinstance(x, Field):
        if x.vs in (Scalar, Directional):
            return x.a
        if x.vs in (Boolean, Ldd):
            r = x.a.astype(G.dtype)
            return np.where(x.a == MV_U1, G.dtype(np.nan), r)
        r = x.a.astype(G.dtype)
        return np.where(x.a == MV_I4, G.dtype(np.nan), r)
    if isinstance(x, (bool, np.bool_)):
        return np.asarray(G.dtype(1.0 if x else 0.0))
    return np.asarray(G.dtype(x))


def _arith(op, x, y):
    with np.errstate(all="ignore"):
        r = op(_sc(x), _sc(y))
    return Field(np.asarray(r, dtype=G.dtype), Scalar)


def _divide(x, y):
    a, b = _sc(x), _sc(y)
    with np.errstate(all="ignore"):
        r = np.asarray(np.divide(a, b), dtype=G.dtype)
        r = np.where(np.broadcast_to(b, r.shape) == 0, G.dtype(np.nan), r)
    return Field(np.asarray(r, dtype=G.dtype), Scalar)


def _power(x, y):
    a, b = _sc(x).astype(np.float64), _sc(y).astype(np.float64)
    with np.errstate(all="ignore"):
        r = np.power(a, b)
    r = np.where(np.isfinite(r) | np.isnan(r), r, np.nan)
    return Field(np.asarray(r, dtype=G.dtype), Scalar)


def _intlike(x):
    return isinstance(x, Field) and x.vs in (Nominal, Ordinal)


def _compare(op, x, y):
    if _intlike(x) or _intlike(y):
        a, b = _to_int(x), _to_int(y)
        mv = (a == MV_I4) | (b == MV_I4)
        r = op(a, b)
    else:
        a, b = _sc(x), _sc(y)
        mv = np.isnan(a) | np.isnan(b)
        with np.errstate(all="ignore"):
            r = op(a, b)
    out = np.where(mv, MV_U1, np.asarray(r, dtype=np.uint8)).astype(np.uint8)
    return Field(out, Boolean)


def _b(x):
    """Operand -> (truth ndarray[bool], mv ndarray[bool])."""
    if isinstance(x, Field):
        if x.vs == Boolean or x.vs == Ldd:
            return x.a == 1, x.a == MV_U1
        s = _sc(x)
        return s != 0, np.isnan(s)
    return np.asarray(bool(x)), np.asarray(False)


def _mkbool(t, mv):
    return Field(np.where(mv, MV_U1, t.astype(np.uint8)).astype(np.uint8), Boolean)


def pcrand(x, y):
    (a, ma), (b, mb) = _b(x), _b(y)
    return _mkbool(a & b, ma | mb)


def pcror(x, y):
    (a, ma), (b, mb) = _b(x), _b(y)
    return _mkbool(a | b, ma | mb)


def pcrnot(x):
    a, ma = _b(x)
    return _mkbool(~a, ma)


# ----------------------------------------------------------------------------------------------
# conversions
# ----------------------------------------------------------------------------------------------
def scalar(x):
    return Field(np.asarray(_sc(x), dtype=G.dtype), Scalar)


def boolean(x):
    t, mv = _b(x)
    return _mkbool(np.asarray(t), np.asarray(mv))


def _to_int(x):
    if isinstance(x, Field) and x.vs in (Nominal, Ordinal):
        return x.a
    s = _sc(x)
    with np.errstate(all="ignore"):
        r = np.where(np.isnan(s), MV_I4, np.nan_to_num(s).astype(np.int32)).astype(np.int32)
    return r


def nominal(x):
    return Field(_to_int(x), Nominal)


def ordinal(x):
    return Field(_to_int(x), Ordinal)


def ldd(x):
    s = _sc(x)
    return Field(np.where(np.isnan(s), MV_U1, np.nan_to_num(s).astype(np.uint8)).astype(np.uint8), Ldd)


def _from_array(arr, vs=None):
    arr = np.asarray(arr)
    if vs is None:
        if arr.dtype == np.bool_:
            vs = Boolean
        elif arr.dtype == np.uint8:
            vs = Ldd
        elif np.issubdtype(arr.dtype, np.integer):
            vs = Nominal
        else:
            vs = Scalar
    if vs in (Scalar, Directional):
        return Field(np.ascontiguousarray(arr, dtype=G.dtype), vs)
    if vs == Boolean:
        return Field(np.ascontiguousarray(arr, dtype=np.uint8), vs)
    if vs == Ldd:
        return Field(np.ascontiguousarray(arr, dtype=np.uint8), vs)
    return Field(np.ascontiguousarray(arr, dtype=np.int32), vs)


def numpy2pcr(vs, arr, mv):
    arr = np.asarray(arr)
    if vs in (Scalar, Directional):
        a = arr.astype(G.dtype)
        a[arr == mv] = np.nan
        return Field(a, vs)
    if vs in (Boolean, Ldd):
        a = arr.astype(np.uint8)
        a[arr == mv] = MV_U1
        return Field(a, vs)
    a = arr.astype(np.int32)
    a[arr == mv] = MV_I4
    return Field(a, vs)


def pcr2numpy(f, mv):
    f = _spatial(f)
    if f.vs in (Scalar, Directional):
        a = f.a.copy()
        a[np.isnan(a)] = mv
        return a
    if f.vs in (Boolean, Ldd):
        a = f.a.astype(np.int16 if mv < 0 else np.uint8)
        a[f.a == MV_U1] = mv
        return a
    a = f.a.copy()
    a[f.a == MV_I4] = mv
    return a


def _spatial(f):
    if not isinstance(f, Field):
        f = scalar(f)
    if f.a.ndim == 0:
        return Field(np.full((G.nrows, G.ncols), f.a[()], dtype=f.a.dtype), f.vs)
    return f


def spatial(f):
    return _spatial(f)


# ----------------------------------------------------------------------------------------------
# elementwise functions
# ----------------------------------------------------------------------------------------------
def _select_parts(x):
    """Field/number -> (array in its own storage dtype, mv mask, value scale)."""
    if isinstance(x, Field):
        if x.vs in (Scalar, Directional):
            return x.a, np.isnan(x.a), x.vs
        if x.vs in (Boolean, Ldd):
            return x.a, x.a == MV_U1, x.vs
        return x.a, x.a == MV_I4, x.vs
    a = np.asarray(G.dtype(x))
    return a, np.isnan(a), None


def _mv_of(vs):
    return G.dtype(np.nan) if vs in (Scalar, Directional) else (MV_U1 if vs in (Boolean, Ldd) else MV_I4)


def _cast_to(a, vs_from, vs_to):
    if vs_from == vs_to:
        return a
    if vs_to in (Scalar, Directional):
        return a.astype(G.dtype)
    if vs_to in (Boolean, Ldd):
        return (np.asarray(a) != 0).astype(np.uint8) if vs_to == Boolean else np.asarray(a).astype(np.uint8)
    return np.asarray(a).astype(np.int32)


def ifthenelse(cond, x, y):
    c, cm = _b(cond)
    xa, xm, xvs = _select_parts(x)
    ya, ym, yvs = _select_parts(y)
    vs = xvs or yvs or Scalar
    xa = _cast_to(xa, xvs or Scalar, vs)
    ya = _cast_to(ya, yvs or Scalar, vs)
    r = np.where(c, xa, ya)
    mv = cm | np.where(c, xm, ym)
    r = np.where(mv, _mv_of(vs), r).astype(xa.dtype if xa.dtype == ya.dtype else np.result_type(xa, ya))
    if vs in (Scalar, Directional):
        r = r.astype(G.dtype)
    return Field(np.asarray(r), vs)


def ifthen(cond, x):
    c, cm = _b(cond)
    xa, xm, xvs = _select_parts(x)
    vs = xvs or Scalar
    mv = cm | ~c | xm
    r = np.where(mv, _mv_of(vs), xa).astype(xa.dtype)
    return Field(np.asarray(r), vs)


def cover(*args):
    xa, xm, vs = _select_parts(args[0])
    vs = vs or Scalar
    r, m = xa, xm
    for nxt in args[1:]:
        na, nm, nvs = _select_parts(nxt)
        na = _cast_to(na, nvs or Scalar, vs)
        r = np.where(m, na, r)
        m = m & nm
    r = np.where(m, _mv_of(vs), r)
    if vs in (Scalar, Directional):
        r = r.astype(G.dtype)
    elif vs in (Boolean, Ldd):
        r = r.astype(np.uint8)
    else:
        r = r.astype(np.int32)
    return Field(np.asarray(r), vs)


def defined(x):
    _, m, _ = _select_parts(x)
    return Field((~np.asarray(m)).astype(np.uint8), Boolean)


def _reduce(fn, args):
    r = _sc(args[0])
    for nxt in args[1:]:
        r = fn(r, _sc(nxt))          # np.maximum / np.minimum propagate NaN (= MV)
    return Field(np.asarray(r, dtype=G.dtype), Scalar)


def max(*args):  # noqa: A001  (PCRaster's own name)
    return _reduce(np.maximum, args)


def min(*args):  # noqa: A001
    return _reduce(np.minimum, args)


def _unary_double(fn, x, domain=None):
    """Assumption A2: evaluate in double on the REAL4 operand, store REAL4."""
    a = _sc(x).astype(np.float64)
    with np.errstate(all="ignore"):
        r = fn(a)
        if domain is not None:
            r = np.where(domain(a), r, np.nan)
    return Field(np.asarray(r, dtype=G.dtype), Scalar)


def _angle(a):
    return a if G.radians else np.deg2rad(a)


def exp(x): return _unary_double(np.exp, x)
def ln(x): return _unary_double(np.log, x, lambda a: a > 0)
def log10(x): return _unary_double(np.log10, x, lambda a: a > 0)
def sqrt(x): return _unary_double(np.sqrt, x, lambda a: a >= 0)
def sqr(x): return _arith(np.multiply, x, x)
def sin(x): return _unary_double(lambda a: np.sin(_angle(a)), x)
def cos(x): return _unary_double(lambda a: np.cos(_angle(a)), x)
def tan(x): return _unary_double(lambda a: np.tan(_angle(a)), x)
def abs(x): return Field(np.abs(_sc(x)), Scalar)  # noqa: A001


def acos(x):
    f = _unary_double(np.arccos, x, lambda a: np.abs(a) <= 1)
    return f if G.radians else Field(np.rad2deg(f.a.astype(np.float64)).astype(G.dtype), Scalar)


def asin(x):
    f = _unary_double(np.arcsin, x, lambda a: np.abs(a) <= 1)
    return f if G.radians else Field(np.rad2deg(f.a.astype(np.float64)).astype(G.dtype), Scalar)


def atan(x):
    f = _unary_double(np.arctan, x)
    return f if G.radians else Field(np.rad2deg(f.a.astype(np.float64)).astype(G.dtype), Scalar)


# ----------------------------------------------------------------------------------------------
# clone / geometry / IO
# ----------------------------------------------------------------------------------------------
def setclone(*args):
    if len(args) == 1:
        m = G.maps.get(_norm(args[0]))
        if m is None:
            raise RuntimeError(f"setclone: cannot open '{args[0]}'")
        arr = m.a if isinstance(m, Field) else (m() if callable(m) else m)
        G.nrows, G.ncols = np.asarray(arr).shape
    else:
        G.nrows, G.ncols, G.cellsize = int(args[0]), int(args[1]), float(args[2])


def setglobaloption(opt):
    if opt == "radians":
        G.radians = True
    elif opt == "degrees":
        G.radians = False
    elif opt == "matrixtable":
        G.matrixtable = True
    elif opt == "columntable":
        G.matrixtable = False


def cellarea():
    return Field(np.full((G.nrows, G.ncols), G.dtype(G.cellsize) * G.dtype(G.cellsize), dtype=G.dtype), Scalar)


def celllength():
    return Field(np.full((G.nrows, G.ncols), G.dtype(G.cellsize), dtype=G.dtype), Scalar)


def cellvalue(f, *idx):
    f = _spatial(f)
    if len(idx) == 1:
        r, c = divmod(int(idx[0]) - 1, G.ncols)
    else:
        r, c = int(idx[0]) - 1, int(idx[1]) - 1
    a, m, vs = _select_parts(f)
    v = a[r, c]
    valid = not bool(m[r, c])
    if vs in (Scalar, Directional):
        return float(v), valid
    return int(v), valid


def readmap(path):
    key = _norm(path)
    if key not in G.maps:
        raise RuntimeError(f"readmap: cannot open '{path}'")
    v = G.maps[key]
    if callable(v):
        v = v()
    if isinstance(v, Field):
        return v
    return _from_array(v)


def report(f, path):
    f = _spatial(f if isinstance(f, Field) else scalar(f))
    G.reports.setdefault(_norm(path), {})[None] = f.a.copy()


# ----------------------------------------------------------------------------------------------
# lookup tables (files on disk, whitespace separated)
# ----------------------------------------------------------------------------------------------
def _read_table(path):
    rows = []
    with open(path) as fh:
        for line in fh:
            line = line.strip()
            if not line or line.startswith("#"):
                continue
            rows.append([float(t) for t in line.replace(",", " ").split()])
    return rows


def _lookup(table, args):
    rows = _read_table(table)
    if G.matrixtable or len(args) == 2:
        # matrix table: first row = column keys (first entry ignored), first column = row keys.
        # reservoirs.initial / lakes.initial (reservoirs.py:153, lakes.py:221) issue the same 3-argument
        # lookups after the option was switched back to "columntable"; the intent is plainly the matrix
        # form, so 3-argument lookups are always read that way here (init-time only, not on the hot path).
        col, key = args
        colkeys = rows[0][1:]
        j = colkeys.index(float(col)) + 1
        mapping = {r[0]: r[j] for r in rows[1:]}
        keyf = key
    else:
        (keyf,) = args
        mapping = {r[0]: r[-1] for r in rows}
    ka, km, _ = _select_parts(keyf)
    ka = np.asarray(ka)
    out = np.full(ka.shape, np.nan, dtype=np.float64)
    for k, v in mapping.items():
        out[(ka == k) & ~km] = v
    return out


def lookupscalar(table, *args):
    return Field(_lookup(table, args).astype(G.dtype), Scalar)


def lookupnominal(table, *args):
    r = _lookup(table, args)
    return Field(np.where(np.isnan(r), MV_I4, np.nan_to_num(r).astype(np.int32)).astype(np.int32), Nominal)


# ----------------------------------------------------------------------------------------------
# area operations
# ----------------------------------------------------------------------------------------------
def _area(fn, x, classes, init):
    xa = _sc(_spatial(x if isinstance(x, Field) else scalar(x)))
    ca = _to_int(_spatial(classes))
    out = np.full(xa.shape, np.nan, dtype=np.float64)
    valid = (ca != MV_I4)
    ids = np.unique(ca[valid])
    for k in ids:
        sel = ca == k
        vals = xa[sel].astype(np.float64)
        out[sel] = np.nan if np.isnan(vals).any() else fn(vals)
    return Field(out.astype(G.dtype), Scalar)


def areatotal(x, classes): return _area(np.sum, x, classes, 0.0)
def areamaximum(x, classes): return _area(np.max, x, classes, -np.inf)
def areaminimum(x, classes): return _area(np.min, x, classes, np.inf)
def areaaverage(x, classes): return _area(np.mean, x, classes, 0.0)


def clump(classes):
    from scipy import ndimage
    ca = _to_int(_spatial(classes))
    out = np.full(ca.shape, MV_I4, dtype=np.int32)
    nxt = 1
    for k in np.unique(ca[ca != MV_I4]):
        lab, n = ndimage.label(ca == k, structure=np.ones((3, 3), dtype=int))
        out[lab > 0] = lab[lab > 0] + (nxt - 1)
        nxt += n
    return Field(out, Nominal)


# ----------------------------------------------------------------------------------------------
# LDD operations
# ----------------------------------------------------------------------------------------------
def _ldd_struct(lddf):
    key = (lddf.a.ctypes.data, lddf.a.shape)
    hit = G._ldd_cache.get(key)
    if hit is not None and hit[0] is lddf.a:
        return hit[1]
    code = np.where(lddf.a == MV_U1, 0, lddf.a).astype(np.uint8)
    st = LddStruct(code)
    G._ldd_cache[key] = (lddf.a, st)
    return st


def _ldd_apply(lddf, fn, *fields):
    st = _ldd_struct(lddf)
    comp = [st.gather(_sc(_spatial(f if isinstance(f, Field) else scalar(f)))) for f in fields]
    # MV in -> MV for the cell and everything downstream: NaN propagates through the double sums
    out = fn(st, *comp)
    return Field(st.scatter(out).astype(G.dtype), Scalar)


def accuflux(lddf, material):
    return _ldd_apply(lddf, lambda st, m: st.accuflux(m), material)


def accufractionflux(lddf, material, fraction):
    return _ldd_apply(lddf, lambda st, m, f: st.accuflux(m, f), material, fraction)


def accufractionstate(lddf, material, fraction):
    return _ldd_apply(lddf, lambda st, m, f: st.accuflux(m, f, want_state=True)[1], material, fraction)


def accucapacityflux(lddf, material, capacity):
    return _ldd_apply(lddf, lambda st, m, c: st.accucapacity(m, c)[0], material, capacity)


def accucapacitystate(lddf, material, capacity):
    return _ldd_apply(lddf, lambda st, m, c: st.accucapacity(m, c)[1], material, capacity)


def subcatchment(lddf, points):
    st = _ldd_struct(lddf)
    pts = _to_int(_spatial(points))
    ids = st.subcatchment(np.where(pts.ravel()[st.flat_active] == MV_I4, 0, pts.ravel()[st.flat_active]))
    out = np.full(G.nrows * G.ncols, MV_I4, np.int32)
    out[st.flat_active] = ids
    return Field(out.reshape(G.nrows, G.ncols), Nominal)


def streamorder(lddf):
    st = _ldd_struct(lddf)
    out = np.full(G.nrows * G.ncols, MV_I4, np.int32)
    out[st.flat_active] = st.streamorder()
    return Field(out.reshape(G.nrows, G.ncols), Ordinal)


def upstream(lddf, x):
    return _ldd_apply(lddf, lambda st, v: st.upstream(v), x)


def catchmenttotal(x, lddf):
    return accuflux(lddf, x)


# ----------------------------------------------------------------------------------------------
# pcraster.framework
# ----------------------------------------------------------------------------------------------
def generateNameT(name, time):
    head, tail = os.path.split(name)
    digits = "%d" % time
    pad = 11 - (len(tail) + len(digits))
    if pad < 0:
        raise ValueError("generateNameT: name too long")
    s = tail + "0" * pad + digits
    return os.path.join(head, s[:8] + "." + s[8:])


class DynamicModel:
    _t = 0

    def currentTimeStep(self):
        return self._t

    def report(self, var, name):
        f = _spatial(var if isinstance(var, Field) else scalar(var))
        G.reports.setdefault(_norm(name), {})[self._t] = f.a.copy()

    def readmap(self, name):
        return readmap(generateNameT(name, self._t))


class DynamicFramework:
    def __init__(self, model, lastTimeStep=0, firstTimestep=1):
        self.model, self.last, self.first = model, lastTimeStep, firstTimestep

    def run(self):
        self.model._t = 0
        self.model.initial()
        if G.step_hook:
            G.step_hook(self.model, 0)
        for t in range(self.first, self.last + 1):
            self.model._t = t
            self.model.dynamic()
            if G.step_hook:
                G.step_hook(self.model, t)


class TimeoutputTimeseries:
    def __init__(self, name, model, idMap, noHeader=False):
        self.name, self.model = name, model
        ids = _to_int(_spatial(idMap if isinstance(idMap, Field) else readmap(idMap)))
        sel = (ids != MV_I4) & (ids > 0)
        flat = np.flatnonzero(sel.ravel())
        order = np.argsort(ids.ravel()[flat], kind="stable")
        self.cells = flat[order]
        self.ids = ids.ravel()[self.cells]
        self.rows = []
        G.tss[name] = self

    def sample(self, expr):
        f = _spatial(expr if isinstance(expr, Field) else scalar(expr))
        self.rows.append((self.model._t, _sc(f).ravel()[self.cells].copy()))


def make_modules():
    """Return (pcraster, pcraster.framework) module objects suitable for sys.modules."""
    pcr = types.ModuleType("pcraster")
    this = globals()
    for k in ("Scalar", "Nominal", "Boolean", "Ldd", "Ordinal", "Directional", "Field", "scalar", "boolean",
              "nominal", "ordinal", "ldd", "spatial", "numpy2pcr", "pcr2numpy", "ifthenelse", "ifthen", "cover",
              "defined", "max", "min", "exp", "ln", "log10", "sqrt", "sqr", "sin", "cos", "tan", "acos", "asin",
              "atan", "abs", "pcrand", "pcror", "pcrnot", "setclone", "setglobaloption", "cellarea", "celllength",
              "cellvalue", "readmap", "report", "lookupscalar", "lookupnominal", "areatotal", "areamaximum",
              "areaminimum", "areaaverage", "clump", "accuflux", "accufractionflux", "accufractionstate",
              "upstream", "catchmenttotal", "accucapacityflux", "accucapacitystate", "subcatchment", "streamorder"):
        setattr(pcr, k, this[k])
    fw = types.ModuleType("pcraster.framework")
    for k in ("DynamicModel", "DynamicFramework", "TimeoutputTimeseries", "generateNameT"):
        setattr(fw, k, this[k])
    pcr.framework = fw
    return pcr, fw
