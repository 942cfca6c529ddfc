"""oracle/run_reference.py -- TEST INFRASTRUCTURE ONLY.

Runs the UNMODIFIED reference (`/root/reference/sphy.py`, executed with runpy exactly as
`python sphy.py sphy_config.cfg` would) on a synthetic catchment, with `oracle/pcr_shim.py` registered
as `pcraster` / `pcraster.framework` (PCRaster itself is absent from this container).  Works only where
/root/reference exists (this container, not the GPU box); its outputs are committed as golden vectors
under tests/golden/ by `oracle/make_golden.py`.
"""
from __future__ import annotations

import contextlib
import glob
import io
import os
import runpy
import sys
import tempfile
import types

import numpy as np

from . import pcr_shim as shim

REFERENCE_DIR = os.environ.get("SPHY_REFERENCE_DIR", "/root/reference")

STATE_ATTRS = (
    "RootWater", "SubWater", "CapRise", "RootDrain", "GwRecharge", "Gw", "BaseR", "H_gw", "SubDrain",
    "SnowStore", "SnowWatStore", "TotalSnowStore", "SoilWater",
    "QRAold", "RootRRAold", "RootDRAold", "RainRAold", "SnowRAold", "GlacRAold", "BaseRAold",
    "StorRES", "RootRR", "RootDR", "RainR", "SnowR", "GlacR", "SnowSoil", "waterbalanceTot", "GlacFrac", "Scanopy", "Kc",
)

_REF_MODULE_PREFIXES = ("ET", "rootzone", "subzone", "hargreaves", "modules", "utilities")


def reference_available():
    return os.path.isfile(os.path.join(REFERENCE_DIR, "sphy.py"))


def _purge_reference_modules():
    for name in list(sys.modules):
        root = name.split(".")[0]
        if root in _REF_MODULE_PREFIXES:
            mod = sys.modules[name]
            f = getattr(mod, "__file__", "") or ""
            if f.startswith(REFERENCE_DIR) or root in ("modules", "utilities"):
                del sys.modules[name]


def _as_raster(v, nrows, ncols):
    if isinstance(v, shim.Field):
        a = shim._sc(v)
        if a.ndim == 0:
            return np.full((nrows, ncols), a[()], dtype=np.float32)
        return np.asarray(a, dtype=np.float32).copy()
    if isinstance(v, (int, float, np.floating, np.integer)):
        return np.full((nrows, ncols), np.float32(v), dtype=np.float32)
    return None


def run_reference(cat, nsteps=None, dtype=np.float32, keep_table=False, quiet=True):
    """Run the reference on Catchment `cat`; returns dict with
       states[attr] -> float32 [nsteps+1, nrows, ncols] (index 0 = after initial()),
       reports[name] -> float32 [nsteps, nrows, ncols] for every daily-reported variable,
       glac_table -> list of per-step DataFrame copies (if keep_table), seconds -> wall time of the step loop."""
    if not reference_available():
        raise RuntimeError("reference sources not found at " + REFERENCE_DIR)
    import time

    nsteps = nsteps or cat.nsteps
    cat.nsteps = nsteps
    shim.G.reset()
    shim.G.dtype = dtype
    shim.G.cellsize = cat.cellsize
    tmp = tempfile.mkdtemp(prefix="sphy_ref_")
    inp = os.path.join(tmp, "in")
    out = os.path.join(tmp, "out")
    cfg = cat.write_inputs(inp, out)
    mask = getattr(cat, "mask", None)
    for name, arr in cat.static_maps().items():
        arr = np.asarray(arr)
        if mask is not None and name != "clone.map":          # PCRaster maps carry MV outside the catchment
            if arr.dtype == np.float32:
                arr = np.where(mask, arr, np.float32(np.nan)).astype(np.float32)
            elif arr.dtype == np.uint8:
                arr = np.where(mask, arr, 255).astype(np.uint8)
            elif arr.dtype == np.int32:
                arr = np.where(mask, arr, -2147483648).astype(np.int32)
        shim.register_map(os.path.join(inp, name), arr)
    cache = {}

    def provider(day, key):
        def get():
            if cache.get("day") != day:
                cache.clear()
                cache["day"] = day
                cache["f"] = cat.forcing(day)
            return cache["f"][key]
        return get

    for day in range(1, nsteps + 1):
        for key, name in cat.forcing_map_names(day).items():
            shim.register_map(os.path.join(inp, name), provider(day, key))

    states = {k: [] for k in STATE_ATTRS}
    tables = []
    t_marks = []

    def hook(model, t):
        t_marks.append(time.perf_counter())
        for k in STATE_ATTRS:
            if hasattr(model, k):
                r = _as_raster(getattr(model, k), cat.nrows, cat.ncols)
                if r is not None:
                    states[k].append(r)
        if keep_table and hasattr(model, "GlacTable"):
            tables.append(model.GlacTable.copy())

    shim.G.step_hook = hook
    pcr, fw = shim.make_modules()
    saved = {k: sys.modules.get(k) for k in ("pcraster", "pcraster.framework", "netCDF4", "pyproj")}
    sys.modules["pcraster"] = pcr
    sys.modules["pcraster.framework"] = fw
    nc = types.ModuleType("netCDF4")
    pj = types.ModuleType("pyproj")
    pj.Proj = pj.transform = None
    sys.modules["netCDF4"] = nc
    sys.modules["pyproj"] = pj
    _purge_reference_modules()
    old_argv, old_cwd, old_path = sys.argv, os.getcwd(), list(sys.path)
    sys.argv = ["sphy.py", cfg]
    sys.path.insert(0, REFERENCE_DIR)
    os.chdir(out)
    sink = io.StringIO()
    # pandas >= 3 refuses the silent int64 -> float64 column upcast the reference relies on
    # (`GlacTable["GlacMelt"] = 0` at glacier.py:355 followed by float assignments at :366-378).  Restore
    # the pandas-2 outcome without touching the reference: whole-column Python-int scalars are stored as float.
    import pandas as _pd
    _orig_setitem = _pd.DataFrame.__setitem__

    def _setitem(self, key, value):
        if type(value) is int:
            value = float(value)
        return _orig_setitem(self, key, value)

    _pd.DataFrame.__setitem__ = _setitem
    # modules/sediment_transport.py:164-166 still uses the `np.int` alias NumPy removed in 1.24
    _had_np_int = hasattr(np, "int")
    if not _had_np_int:
        np.int = int
    try:
        with (contextlib.redirect_stdout(sink) if quiet else contextlib.nullcontext()):
            runpy.run_path(os.path.join(REFERENCE_DIR, "sphy.py"), run_name="__main__")
    finally:
        _pd.DataFrame.__setitem__ = _orig_setitem
        if not _had_np_int:
            del np.int
        os.chdir(old_cwd)
        sys.argv = old_argv
        sys.path[:] = old_path
        for k, v in saved.items():
            if v is None:
                sys.modules.pop(k, None)
            else:
                sys.modules[k] = v
        _purge_reference_modules()
    result = {"states": {k: np.stack(v) for k, v in states.items() if len(v) == nsteps + 1},
              "reports": {}, "seconds": t_marks[-1] - t_marks[0] if len(t_marks) > 1 else 0.0,
              "tss": {k: v.rows for k, v in shim.G.tss.items()}}
    for name, per_t in shim.G.reports.items():
        steps = sorted(t for t in per_t if t is not None)
        if len(steps) == nsteps:
            result["reports"][os.path.basename(name)] = np.stack([per_t[t] for t in steps]).astype(np.float32)
    # everything self.report / pcr.report wrote, keyed like the files PCRaster would create
    from sphy_b200.csf import stack_name
    result["report_files"] = {}
    for name, per_t in shim.G.reports.items():
        base = os.path.basename(name)
        for t, arr in per_t.items():
            result["report_files"][base if t is None else stack_name(base, t)] = np.asarray(arr, np.float32)
    # tables the reference wrote with pandas (per-glacier reporting, glacier table after the yearly update)
    result["csv_files"] = {}
    for path in sorted(glob.glob(os.path.join(out, "*.csv"))):
        with open(path) as fh:
            result["csv_files"][os.path.basename(path)] = fh.read()
    result["tss_tables"] = {k: np.array([[t] + list(v) for t, v in rows], np.float64) for k, rows in result["tss"].items()}
    if keep_table:
        result["glac_table"] = tables
    return result
