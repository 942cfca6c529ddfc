"""SphyModel -- host side of the B200 hot path, mirroring the reference's model object.

The reference is one class `sphy(pcrm.DynamicModel)` with `__init__` (cfg + map loading,
/root/reference/sphy.py:38-562), `initial()` (sphy.py:565-717) and `dynamic()` (sphy.py:719-1263), driven by
`DynamicFramework.run()`.  This class keeps that contract -- same `sphy_config.cfg` switches, same state
attribute names (`RootWater`, `SubWater`, `CapRise`, `RootDrain`, `GwRecharge`, `Gw`, `BaseR`, `H_gw`,
`SnowStore`, `SnowWatStore`, `TotalSnowStore`, `QRAold`, `*RAold`, `StorRES`, `QFRAC`, `GlacFrac`) -- but the
~250 PCRaster raster passes of one `dynamic()` collapse into 2-4 CUDA kernel launches through the C ABI.
State lives in torch CUDA tensors over the N active cells (compact row-major order); torch is used for
buffers, streams and copies only.  There is no CPU fallback.
"""
from __future__ import annotations

import configparser
import ctypes as C
import datetime
import math
import os

import numpy as np
import torch

from . import _lib
from ._lib import SphyParam

f32 = np.float32
COMPONENTS = ("RootR", "RootD", "Rain", "Snow", "Glac", "Base")       # modules/routing.py:53


# ----------------------------------------------------------------------------------------------
# map sources (the `pcr.readmap` side of the boundary)
# ----------------------------------------------------------------------------------------------
class MapSource:
    """Where `readmap(inputdir + name)` gets its rasters from.  read() returns a 2-D ndarray or raises
    FileNotFoundError (the reference's scalar-or-map idiom relies on readmap failing)."""

    def read(self, name):
        raise FileNotFoundError(name)


class DictMapSource(MapSource):
    def __init__(self, static_maps, forcing_fn=None, forcing_names=None):
        self.maps = dict(static_maps)
        self.forcing_fn = forcing_fn
        self.forcing_names = forcing_names
        self._cache = (None, None)

    def read(self, name):
        if name in self.maps:
            return self.maps[name]
        # map stacks `forcings/<key>0000.001` (pcrm.generateNameT, sphy.py:756) are served by the forcing generator
        base = os.path.basename(name)
        if self.forcing_fn is not None and len(base) == 12 and base[8] == ".":
            key = base[:8].rstrip("0123456789")
            try:
                t = int(base[len(key):8] + base[9:])
            except ValueError:
                raise FileNotFoundError(name) from None
            if self._cache[0] != t:
                self._cache = (t, self.forcing_fn(t))
            if key in self._cache[1]:
                return self._cache[1][key]
        raise FileNotFoundError(name)


class CatchmentSource(DictMapSource):
    """Feeds a synthetic `synth.Catchment` (its static maps and seeded daily forcing)."""

    def __init__(self, cat):
        super().__init__(cat.static_maps(), forcing_fn=cat.forcing)
        self.cellsize = cat.cellsize          # the reference takes it from the clone map's CSF header


class DirMapSource(MapSource):
    """The reference's input directory: PCRaster CSF maps (`dem.map`, `forcings/prec0000.001`, ...) read with
    sphy_b200.csf, or `<name>.npy` rasters.  The cell size comes from the clone map's CSF header like in PCRaster."""

    def __init__(self, root, cellsize=None):
        self.root = root
        self.cellsize = cellsize

    def read(self, name):
        from . import csf
        path = os.path.join(self.root, name)
        if os.path.exists(path):
            arr, info = csf.read_map(path)
            if self.cellsize is None:
                self.cellsize = info["cellsize"]
            return arr
        if os.path.exists(path + ".npy"):
            return np.load(path + ".npy")
        raise FileNotFoundError(path)


def name_t(t):
    """PCRaster stack-name suffix for a 4-letter prefix: 1 -> '0000.001', 1000 -> '0001.000' (generateNameT)."""
    s = "xxxx" + "0" * (11 - 4 - len(str(t))) + str(t)
    return (s[:8] + "." + s[8:])[4:]


def _clone_mask(clone):
    """Active cells of the clone map: non-zero and not a missing value.  PCRaster's MV is NaN for REAL4, the most negative
    value for INT4 (-2147483648), 255 for UINT1 and -128 for INT1 -- all of which `astype(bool)` would read as active."""
    clone = np.asarray(clone)
    if clone.dtype == bool:
        return clone
    if np.issubdtype(clone.dtype, np.floating):
        return np.isfinite(clone) & (clone != 0)
    mv = {np.dtype(np.int32): -2147483648, np.dtype(np.uint8): 255, np.dtype(np.int8): -128, np.dtype(np.int64): -2147483648}
    mask = clone != 0
    if clone.dtype in mv:
        mask &= clone != mv[clone.dtype]
    return mask


def julian(date):
    """utilities/timecalc.py:24-29"""
    return date.toordinal() - datetime.date(date.year, 1, 1).toordinal() + 1


def _read_matrix_table(path):
    rows = []
    with open(path) as fh:
        for line in fh:
            line = line.strip()
            if line and not line.startswith("#"):
                rows.append([float(t) for t in line.replace(",", " ").split()])
    return rows


def _table_col(rows, col):
    j = [float(x) for x in rows[0][1:]].index(float(col)) + 1
    return {int(r[0]): float(r[j]) for r in rows[1:]}


# ----------------------------------------------------------------------------------------------
class StepGraph:
    """CUDA graphs of the device step of one model or of a whole list of models sharing a context (ensemble members).

    Nothing in a captured step may depend on the day through a kernel argument.  The day-of-year scalars of the Ra table
    therefore come from a device table with one row per remaining day of the run (evaluated on the host like the reference
    does, sphy_day_scalars); a one-thread kernel at the head of every device step copies the row of the current day into
    the block the Ra kernel reads and advances a device-side day counter -- the host never touches memory a queued step
    still has to read."""

    @staticmethod
    def state(model):
        ndays = max((model.enddate.date() - model.curdate.date()).days + 1, 1) + 400      # + slack for runs past `enddate`
        k0 = float(getattr(model._wbP, "k0", 0.0))
        rows = np.zeros((ndays, 5), np.float32)
        out = (C.c_float * 5)()
        cache = {}
        for i in range(ndays):
            doy = julian(model.curdate + datetime.timedelta(days=i))
            if doy not in cache:
                model._ck(model.lib.sphy_day_scalars(doy, C.c_float(k0), out), "sphy_day_scalars")
                cache[doy] = list(out)
            rows[i] = cache[doy]
        table = torch.from_numpy(rows).to(model.device)
        block = torch.zeros(8, dtype=torch.float32, device=model.device)
        counter = torch.zeros(1, dtype=torch.int32, device=model.device)
        model._ck(model.lib.sphy_set_day_block(model._ctx, block.data_ptr()), "sphy_set_day_block")
        return {"graphs": {}, "seen": set(), "table": table, "block": block, "counter": counter, "ndays": ndays, "model": model,
                "steps": 0}

    @staticmethod
    def run(st, models, forcings, doy, capture=True, step_fn=None):
        """One device step of every model: eagerly the first time a set of forcing buffers is seen (or when `capture` is
        off), from a graph after that."""
        lead = st["model"]
        st["steps"] += 1
        if st["steps"] > st["ndays"]:
            raise RuntimeError("the day table of the CUDA-graph mode is exhausted: call enable_graphs() again")
        key = tuple(v.data_ptr() for f in forcings for _, v in sorted(f.items()))

        def body():
            lead._ck(lead.lib.sphy_day_table_step(lead._ctx, st["table"].data_ptr(), st["ndays"], st["counter"].data_ptr(),
                                                  st["block"].data_ptr(), lead._stream()), "sphy_day_table_step")
            if step_fn is not None:
                step_fn(models, forcings, doy)
            else:
                for m, f in zip(models, forcings):
                    m._device_step(f, doy)

        g = st["graphs"].get(key) if capture else None
        if g is not None:
            g.replay()
        elif not capture or key not in st["seen"]:
            st["seen"].add(key)
            body()                                   # first sight: run eagerly (lazy allocations, attribute opt-ins happen here)
        else:
            torch.cuda.synchronize(lead.device)
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g):
                body()
            st["graphs"][key] = g
            g.replay()                               # capturing executes nothing


# ----------------------------------------------------------------------------------------------
class SphyModel:
    def __init__(self, cfg, maps: MapSource, device="cuda:0", route_method="scan", diagnostics=False,
                 collect_tss=True, cell_order="dfs", rank=0, world=1, share_from=None, reporting=False, write_outputs=False,
                 ra_mode="auto", scan_drift=True, overlap_routing=False):
        if not torch.cuda.is_available():
            raise RuntimeError("SphyModel needs a CUDA device (B200); there is no CPU fallback")
        self.lib = _lib.load()
        self.device = torch.device(device)
        torch.cuda.set_device(self.device)
        self.maps = maps
        self.diagnostics = diagnostics or reporting          # reported fluxes come from the diagnostic outputs of K1
        self.enable_reporting = reporting
        self.write_outputs = write_outputs
        self.collect_tss = collect_tss
        self.route_method = {"levels": _lib.ROUTE_LEVELS, "scan": _lib.ROUTE_SCAN}[route_method]
        self.cell_order_mode = {"rowmajor": _lib.ORDER_ROWMAJOR, "dfs": _lib.ORDER_DFS}[cell_order]
        if self.route_method == _lib.ROUTE_SCAN and self.cell_order_mode != _lib.ORDER_DFS:
            raise ValueError("route_method='scan' needs cell_order='dfs'")
        # overlap_routing: route day t on a second stream while the water-balance kernel of day t + 1 runs (nothing in the
        # vertical balance reads the routed discharge).  Routed maps (QRAold, the .tss samples) are then only valid after
        # sync(); everything else behaves as before.  Plain total-flow scan routing only; otherwise silently sequential.
        self.overlap_routing = bool(overlap_routing)
        self.scan_drift = bool(scan_drift)   # scan routing: carry the reference's REAL4 rounding drift on the trunk cells
        self.ra_mode = ra_mode           # auto | table (latitude classes) | cell (per-cell sin/cos/tan maps)
        self.rank, self.world = int(rank), int(world)
        if self.world > 1 and (self.route_method != _lib.ROUTE_SCAN):
            raise ValueError("multi-GPU partitioning needs route_method='scan'")
        if isinstance(cfg, configparser.RawConfigParser):
            self.config = cfg
        else:
            self.config = configparser.RawConfigParser()
            if not self.config.read(cfg):
                raise FileNotFoundError(cfg)
        config = self.config
        self.MV = -9999
        # -- module switches, sphy.py:49-57,81-83
        g = lambda k: config.getint("MODULES", k)  # noqa: E731
        self.GlacFLAG, self.SnowFLAG, self.RoutFLAG = g("GlacFLAG"), g("SnowFLAG"), g("RoutFLAG")
        self.ResFLAG, self.LakeFLAG, self.DynVegFLAG = g("ResFLAG"), g("LakeFLAG"), g("DynVegFLAG")
        self.GroundFLAG, self.SedFLAG = g("GroundFLAG"), g("SedFLAG")
        if self.GlacFLAG == 1:
            self.SnowFLAG = 1
            self.GroundFLAG = 1
        self.SedTransFLAG = g("SedTransFLAG") if config.has_option("MODULES", "SedTransFLAG") else 0
        self.inpath = config.get("DIRS", "inputdir")
        self.outpath = config.get("DIRS", "outputdir")
        # -- timing, sphy.py:90-110
        ti = lambda k: config.getint("TIMING", k)  # noqa: E731
        self.startdate = datetime.datetime(ti("startyear"), ti("startmonth"), ti("startday"))
        self.enddate = datetime.datetime(ti("endyear"), ti("endmonth"), ti("endday"))
        self.ts1date = datetime.datetime(ti("startyear_timestep1"), ti("startmonth_timestep1"), ti("startday_timestep1"))
        self.startYear, self.endYear = ti("startyear"), ti("endyear")                          # sphy.py:107-110
        self.spinUpYears = ti("spinupyears")
        self.simYears = self.endYear - self.startYear - self.spinUpYears + 1
        # -- clone, sphy.py:141-145
        clone = np.asarray(self._readmap(config.get("GENERAL", "mask")))
        self.nrows, self.ncols = clone.shape
        self.clone_mask = _clone_mask(clone)
        self.cellsize = float(config.get("GENERAL", "cellsize")) if config.has_option("GENERAL", "cellsize") else None
        if self.cellsize is None:
            self.cellsize = float(getattr(maps, "cellsize", 0) or 0) or None
        if self.cellsize is None:
            raise ValueError("cell size unknown: give [GENERAL] cellsize in the cfg or a `cellsize` attribute on the map source")
        self.cellArea = f32(f32(self.cellsize) * f32(self.cellsize))
        # ensemble members (BASELINE config 5) share one context: the LDD structures are built and held once per GPU
        self._owns_ctx = share_from is None
        if share_from is not None:
            if (share_from.nrows, share_from.ncols, share_from.device) != (self.nrows, self.ncols, self.device):
                raise ValueError("share_from: the members of an ensemble must live on the same grid and GPU")
            self._ctx = share_from._ctx
        else:
            self._ctx = C.c_void_p()
            _lib.check(self.lib, None, self.lib.sphy_ctx_create(C.byref(self._ctx), self.device.index or 0, self.nrows,
                                                                self.ncols, C.c_float(self.cellsize)), "sphy_ctx_create")
        self._share_from = share_from
        self._keep = []               # tensors referenced by raw pointers
        self._stream = lambda: C.c_void_p(torch.cuda.current_stream(self.device).cuda_stream)  # noqa: E731
        # -- routing structures first: they define the active cells / compact order (routing.py:33-38)
        import time
        t0 = time.time()
        self._build_cells()
        torch.cuda.synchronize(self.device)
        t1 = time.time()
        self._read_parameters()
        self.setup_times = {"ldd_build_and_trunk_s": round(t1 - t0, 2), "read_parameters_s": round(time.time() - t1, 2)}
        # the process modules under the reference's attribute names (sphy.py:60-78, 210-213, 297-300, 374-377 ...)
        from . import modules
        modules.bind(self)
        self.ET, self.rootzone, self.subzone, self.Hargreaves = modules.ET, modules.rootzone, modules.subzone, modules.hargreaves
        self.groundwater, self.snow, self.glacier = modules.groundwater, modules.snow, modules.glacier
        self.routing, self.advanced_routing = modules.routing, modules.advanced_routing
        self.lakes, self.reservoirs = modules.lakes, modules.reservoirs
        self.mmf, self.musle, self.sediment_transport = modules.mmf, modules.musle, modules.sediment_transport

    # ------------------------------------------------------------------------------------------
    def close(self):
        if getattr(self, "_ctx", None):
            self.sync()
            torch.cuda.synchronize(self.device)
            self.check_exchange()
            if self._owns_ctx:
                self.lib.sphy_ctx_destroy(self._ctx)
            self._ctx = None

    def __del__(self):
        try:
            self.close()
        except RuntimeError:
            raise                          # a timed-out confluence exchange must not pass silently (Python reports it)
        except Exception:
            pass

    def check_exchange(self):
        """Raise if a peer-memory confluence exchange gave up waiting for a peer since the last check (synchronises)."""
        self.sync()
        if getattr(self, "_mg_p2p", False) and self.lib.sphy_p2p_timed_out(self._ctx, self._stream()):
            raise RuntimeError("a peer-memory confluence exchange timed out: discharges of this run are not valid "
                               "(the trunk cells of the affected steps hold NaN)")

    def _ck(self, rc, what):
        _lib.check(self.lib, self._ctx, rc, what)

    def _readmap(self, name):
        return self.maps.read(name)

    def _scalar_or_map(self, section, key):
        """`try getfloat / except readmap` (e.g. groundwater.py:51-57): float or compact device-ready ndarray."""
        raw = self.config.get(section, key)
        try:
            return float(raw)
        except ValueError:
            return self._compact_np(self._readmap(raw))

    # -- compact layout ------------------------------------------------------------------------
    def _build_cells(self):
        cfg = self.config
        if self._share_from is not None:
            o = self._share_from
            for k in ("n", "n_global", "nlevels", "flat_active", "cell_order", "full_raster", "_cidx_host"):
                setattr(self, k, getattr(o, k))
            for k in ("part_range", "_mg_stride", "trunk_cells", "_rl_part", "_rl_global", "_rl_plan_host"):
                if hasattr(o, k):
                    setattr(self, k, getattr(o, k))
            if self.route_method != _lib.ROUTE_SCAN:
                self._rl_part = False              # the level sweep does not use the trunk list
            return
        need_ldd = self.RoutFLAG == 1 or self.ResFLAG == 1 or self.LakeFLAG == 1
        if need_ldd:
            ldd = np.ascontiguousarray(self._readmap(cfg.get("ROUTING", "flowdir")), dtype=np.uint8)
            self._ldd_host = ldd
            ldd_d = torch.from_numpy(ldd).to(self.device)
            act_d = torch.from_numpy(self.clone_mask.astype(np.uint8)).to(self.device)
            self._ck(self.lib.sphy_set_cell_order(self._ctx, self.cell_order_mode), "sphy_set_cell_order")
            self._ck(self.lib.sphy_ldd_build(self._ctx, ldd_d.data_ptr(), act_d.data_ptr(), self._stream()), "sphy_ldd_build")
            self.n = int(self.lib.sphy_ncells(self._ctx))
            self.nlevels = int(self.lib.sphy_nlevels(self._ctx))
            flat = torch.empty(self.n, dtype=torch.int32, device=self.device)
            self._ck(self.lib.sphy_ldd_export(self._ctx, b"flat_active", flat.data_ptr(), flat.numel() * 4, self._stream()),
                     "sphy_ldd_export")
            self.flat_active = flat.cpu().numpy().astype(np.int64)      # device cell index -> raster index
            self._ck(self.lib.sphy_ldd_export(self._ctx, b"cell_order", flat.data_ptr(), flat.numel() * 4, self._stream()),
                     "sphy_ldd_export")
            self.cell_order = flat.cpu().numpy().astype(np.int64)       # device cell index -> canonical compact index
            self.n_global = self.n
            if self.route_method == _lib.ROUTE_SCAN:
                self._install_trunk()
        else:
            self.flat_active = np.flatnonzero(self.clone_mask.ravel())
            self.cell_order = np.arange(self.flat_active.size)
            self.n = int(self.flat_active.size)
            self.n_global = self.n
            if self.world > 1:
                raise ValueError("multi-GPU partitioning needs the routing structures (RoutFLAG = 1)")
            self.nlevels = 0
            self._ck(self.lib.sphy_set_ncells(self._ctx, self.n), "sphy_set_ncells")
        self.full_raster = (self.n == self.nrows * self.ncols and
                            bool(np.array_equal(self.flat_active, np.arange(self.n))))   # compact order == raster order
        self._cidx_host = np.full(self.nrows * self.ncols, -1, np.int64)
        self._cidx_host[self.flat_active] = np.arange(self.n)

    def _install_trunk(self):
        """Scan routing: list the trunk cells (sphy_trunk_select), restrict a multi-GPU rank to its contiguous DFS-order
        range, and install the exchange plan (sphy_b200/multigpu.py: plan_trunk).  On one GPU the trunk only carries the
        reference's REAL4 rounding drift; on a partitioned basin it is also what crosses the rank boundaries."""
        from . import multigpu
        adv = self.ResFLAG == 1 or self.LakeFLAG == 1
        # lakes / reservoirs: a partitioned basin puts the cells downstream of the structures on the trunk list
        # (sphy_route_reslake_begin / _finish).  One GPU does the same from 1e6 cells on (measured on BASELINE config 3, 4e6
        # cells: 0.207 instead of 0.234 ms per step, and no pass over a structure's whole upstream range), otherwise it routes
        # with float64 cut corrections inside the main pass (sphy_route_reslake_step), which also serves open-water evaporation
        # and measured lake levels.  SPHY_RESLAKE_PLAN=0/1 forces either path.
        cfg = self.config
        extras = (cfg.getint("OPENWATER", "ETOpenWaterFLAG") == 1 or
                  (self.LakeFLAG == 1 and cfg.has_option("LAKE", "updatelakelevel") and bool(cfg.get("LAKE", "updatelakelevel").strip())))
        force = os.environ.get("SPHY_RESLAKE_PLAN", "")
        self._rl_part = adv and (self.world > 1 or force == "1" or (force != "0" and not extras and self.n >= 1_000_000))
        self.trunk_cells = 0
        if (adv and not self._rl_part) or (self.world == 1 and not self.scan_drift and not adv):
            return                                   # accufractionflux cuts / plain float64 range sums
        big = 2 ** 31 - 1
        drift = self.scan_drift and not adv          # the structure path rounds like the one-GPU cut path: once per cell
        min_level = int(os.environ.get("SPHY_TRUNK_MIN_LEVEL", multigpu.TRUNK_MIN_LEVEL)) if drift else big
        min_cells = int(os.environ.get("SPHY_TRUNK_MIN_CELLS", multigpu.TRUNK_MIN_CELLS)) if drift else 2 ** 40
        bounds = multigpu.partition_bounds(self.n, self.world) if self.world > 1 else np.array([0, self.n], np.int64)
        barr = (C.c_int64 * len(bounds))(*[int(b) for b in bounds])
        m = C.c_int64(0)
        cut_dev = donors = None
        if adv:
            ids = self._structure_ids()                        # whole-basin id maps in device order (before the partition)
            cut = np.flatnonzero((ids["lake"] > 0) | (ids["res"] > 0)).astype(np.int32)
            self._rl_global = dict(ids, cut=cut)
            cut_dev = self._dev(cut if cut.size else np.zeros(1, np.int32))
            self._ck(self.lib.sphy_trunk_select_cuts(self._ctx, int(cut.size), cut_dev.data_ptr() if cut.size else None, min_level,
                                                     min_cells, self.world, barr, C.byref(m), self._stream()), "sphy_trunk_select_cuts")
            don = torch.full((max(cut.size, 1) * 8,), -1, dtype=torch.int32, device=self.device)
            self._ck(self.lib.sphy_cell_donors(self._ctx, int(cut.size), cut_dev.data_ptr() if cut.size else None,
                                               don.data_ptr() if cut.size else None, self._stream()), "sphy_cell_donors")
            donors = don[:cut.size * 8].cpu().numpy().reshape(-1, 8)
        else:
            self._ck(self.lib.sphy_trunk_select(self._ctx, min_level, min_cells, self.world, barr, C.byref(m), self._stream()),
                     "sphy_trunk_select")
        m = int(m.value)
        self.trunk_cells = m

        def export(name, dtype, per=1):
            t = torch.empty(max(m * per, 1), dtype=dtype, device=self.device)
            self._ck(self.lib.sphy_trunk_export(self._ctx, name.encode(), t.data_ptr(), t.numel() * t.element_size(), self._stream()),
                     "sphy_trunk_export")
            return t[:m * per].cpu().numpy()

        if self.world > 1:
            lo, hi = int(bounds[self.rank]), int(bounds[self.rank + 1])
            self._ck(self.lib.sphy_set_partition(self._ctx, lo, hi, self._stream()), "sphy_set_partition")
            self.part_range = (lo, hi)
            self._bounds = np.asarray(bounds, np.int64)
            if self.rank == 0:                     # rank 0 assembles reported rasters from the ranks' slices (gather_raster)
                self._flat_active_global = self.flat_active.astype(np.int32 if self.nrows * self.ncols < 2 ** 31 else np.int64)
            self.flat_active = self.flat_active[lo:hi]
            self.cell_order = self.cell_order[lo:hi]
            self.n = hi - lo
        elif m == 0 and not adv:
            return
        cells_l, sub_l = export("cell", torch.int32), export("sub_size", torch.int32)
        plan = multigpu.plan_trunk(cells_l, sub_l, bounds)
        stride = int(plan["stride"])
        if adv:                                                # the per-structure extras ride at the end of every block
            self._rl_plan_host = multigpu.plan_reslake(cells_l, sub_l, self._rl_global["cut"], donors, bounds)
            self._rl_plan_host["listed"] = cells_l.astype(np.int64)
            self._rl_plan_host["bounds"] = np.asarray(bounds, np.int64)
            stride += _lib.RL_EXTRA * int(self._rl_global["cut"].size)
        d = {k: self._dev(np.ascontiguousarray(plan[k], np.int32)) for k in ("before_rank", "before_slot", "end_rank", "end_slot")}
        pub = self._dev(np.ascontiguousarray(plan["pub"][self.rank], np.int32))
        ptr = lambda t: t.data_ptr() if t.numel() else None  # noqa: E731
        self._ck(self.lib.sphy_trunk_set_plan(self._ctx, int(pub.numel()), ptr(pub), ptr(d["before_rank"]), ptr(d["before_slot"]),
                                              ptr(d["end_rank"]), ptr(d["end_slot"]), stride, self.rank, self.world,
                                              self._stream()), "sphy_trunk_set_plan")
        self._mg_stride = stride

    def _structure_ids(self):
        """LakeId / ResId over the active cells of the WHOLE basin, in device order (lakes.py:166, reservoirs.py:94)."""
        cfg = self.config
        pick = lambda name: np.maximum(np.asarray(self._readmap(name)).ravel()[self.flat_active].astype(np.int32), 0)  # noqa: E731
        zero = np.zeros(self.flat_active.size, np.int32)
        return {"lake": pick(cfg.get("LAKE", "LakeId")) if self.LakeFLAG == 1 else zero,
                "res": pick(cfg.get("RESERVOIR", "ResId")) if self.ResFLAG == 1 else zero}

    def ldd_structure(self, name):
        """Device copy of one Appendix-C structure (bit-exact parity checks)."""
        n, L = self.n, self.nlevels
        spec = {"cidx": (torch.int32, self.nrows * self.ncols), "down": (torch.int32, n), "indeg": (torch.uint8, n),
                "level": (torch.int32, n), "order": (torch.int32, n), "level_ptr": (torch.int32, L + 1),
                "don_ptr": (torch.int32, n + 1), "nup": (torch.int64, n), "sub_size": (torch.int32, n),
                "cell_order": (torch.int32, n), "flat_active": (torch.int32, n)}
        if name == "don_idx":
            ptr = self.ldd_structure("don_ptr")
            spec["don_idx"] = (torch.int32, int(ptr[-1].item()))
        dt, cnt = spec[name]
        out = torch.empty(max(cnt, 1), dtype=dt, device=self.device)
        self._ck(self.lib.sphy_ldd_export(self._ctx, name.encode(), out.data_ptr(), out.numel() * out.element_size(),
                                          self._stream()), "sphy_ldd_export")
        return out[:cnt]

    def _compact_np(self, raster, dtype=np.float32):
        raster = np.asarray(raster)
        if raster.ndim == 2 and raster.strides == (0, 0):      # spatially uniform map delivered as a zero-stride view
            return np.asarray(raster.flat[0], dtype=dtype)     # 0-d array: keeps "map" (REAL4) arithmetic semantics
        flat = raster.ravel()
        if not self.full_raster:
            flat = flat[self.flat_active]
        return np.ascontiguousarray(flat, dtype=dtype)

    def _dev(self, arr, dtype=None):
        a = np.ascontiguousarray(arr if dtype is None else np.asarray(arr, dtype=dtype))
        if not a.flags.writeable:
            a = a.copy()
        return torch.from_numpy(a).to(self.device)

    def _full(self, value, dtype=torch.float32):
        return torch.full((self.n,), float(value), dtype=dtype, device=self.device)

    def to_raster(self, compact, fill=np.nan):
        """compact device map -> (nrows, ncols) ndarray (the `report` side)."""
        out = np.full(self.nrows * self.ncols, fill, np.float32)
        out[self.flat_active] = compact.detach().cpu().numpy()
        return out.reshape(self.nrows, self.ncols)

    def gather_raster(self, compact, fill=np.nan):
        """to_raster for a basin split across GPUs: COLLECTIVE; rank 0 gets the whole (nrows, ncols) raster assembled from
        every rank's slice of the compact map (one all-gather of the padded slices), the other ranks get None."""
        if self.world == 1:
            return self.to_raster(compact, fill)
        import torch.distributed as dist
        sizes = np.diff(self._bounds)
        maxn = int(sizes.max())
        send = torch.full((maxn,), float("nan"), dtype=torch.float32, device=self.device)
        send[:self.n] = compact.detach().to(torch.float32)
        recv = torch.empty(self.world * maxn, dtype=torch.float32, device=self.device)
        dist.all_gather_into_tensor(recv, send)
        if self.rank != 0:
            return None
        vals = recv.view(self.world, maxn).cpu().numpy()
        out = np.full(self.nrows * self.ncols, fill, np.float32)
        for r in range(self.world):
            lo, hi = int(self._bounds[r]), int(self._bounds[r + 1])
            out[self._flat_active_global[lo:hi]] = vals[r, :hi - lo]
        return out.reshape(self.nrows, self.ncols)

    def _param(self, x):
        """scalar-or-map -> (SphyParam, keepalive)"""
        if isinstance(x, torch.Tensor):
            self._keep.append(x)
            return SphyParam(x.data_ptr(), 0.0)
        if isinstance(x, np.ndarray):
            if x.ndim == 0 or (x.size and np.all(x == x.flat[0])):     # spatially uniform map: behaves as the scalar
                return SphyParam(None, float(x.flat[0]))
            t = self._dev(x, np.float32)
            self._keep.append(t)
            return SphyParam(t.data_ptr(), 0.0)
        return SphyParam(None, float(f32(x)))

    # -- parameters, sphy.py:147-458 -------------------------------------------------------------
    def _read_parameters(self):
        cfg = self.config
        cn = self._compact_np
        rd = lambda sec, key: cn(self._readmap(cfg.get(sec, key)))  # noqa: E731
        self.PedotransferFLAG = cfg.getint("PEDOTRANSFER", "PedotransferFLAG")
        sc = lambda k: f32(cfg.getfloat("SOIL_CAL", k))  # noqa: E731
        self.Slope = rd("GENERAL", "Slope")
        self.Locations = cn(self._readmap(cfg.get("GENERAL", "locations")), np.int32)
        if self.PedotransferFLAG == 1:                                                         # sphy.py:163-171
            from .modules import pedotransfer
            self.pedotransfer = pedotransfer
            pedotransfer.init(self, None, cfg)
            RootFieldMap, RootSatMap, RootDryMap, RootWiltMap = self.RootFieldMap, self.RootSatMap, self.RootDryMap, self.RootWiltMap
            SubSatMap, SubFieldMap = self.SubSatMap, self.SubFieldMap
        else:
            RootFieldMap = rd("SOIL", "RootFieldMap") * sc("RootFieldFrac")                 # sphy.py:173-192
            RootSatMap = rd("SOIL", "RootSatMap") * sc("RootSatFrac")
            RootDryMap = rd("SOIL", "RootDryMap") * sc("RootDryFrac")
            RootWiltMap = rd("SOIL", "RootWiltMap") * sc("RootWiltFrac")
            self.RootKsat = rd("SOIL", "RootKsat") * sc("RootKsatFrac")
            SubSatMap, SubFieldMap, self.SubKsat = rd("SOIL", "SubSatMap"), rd("SOIL", "SubFieldMap"), rd("SOIL", "SubKsat")
        self.RootDrainVel = self.RootKsat * self.Slope                                        # sphy.py:198
        for k in ("CapRiseMax", "RootDepthFlat", "SubDepthFlat"):                              # sphy.py:201-206
            setattr(self, k, self._scalar_or_map("SOILPARS", k))
        fl = lambda x: x if isinstance(x, np.ndarray) else f32(x)  # noqa: E731
        self.RootField = RootFieldMap * fl(self.RootDepthFlat)                                # sphy.py:240-247
        self.RootSat = RootSatMap * fl(self.RootDepthFlat)
        self.RootDry = RootDryMap * fl(self.RootDepthFlat)
        self.RootWilt = RootWiltMap * fl(self.RootDepthFlat)
        self.SubSat = SubSatMap * fl(self.SubDepthFlat)
        self.SubField = SubFieldMap * fl(self.SubDepthFlat)
        with np.errstate(all="ignore"):
            self.RootTT = np.maximum((self.RootSat - self.RootField) / self.RootKsat, f32(0.0001))
            self.SubTT = np.maximum((self.SubSat - self.SubField) / self.SubKsat, f32(0.0001))
            # static transcendentals, evaluated like the reference (double on the REAL4 operand, REAL4 result)
            self._e_root = np.exp((f32(-1) / self.RootTT).astype(np.float64)).astype(np.float32)
            self._e_sub = np.exp((f32(-1) / self.SubTT).astype(np.float64)).astype(np.float32)
        if self.GroundFLAG == 1:                                                               # groundwater.py:51-57
            for k in ("GwDepth", "GwSat", "deltaGw", "BaseThresh", "alphaGw", "YieldGw"):
                setattr(self, k, self._scalar_or_map("GROUNDW_PARS", k))
        else:
            self.SeepStatFLAG = cfg.getint("SOILPARS", "SeepStatic")                           # sphy.py:220-229
            if self.SeepStatFLAG == 0:
                self.Seepmaps = cfg.get("SOILPARS", "SeePage")
                self.SeePage = None                # daily map series, sphy.py:992-1000
            else:
                self.SeePage = self._scalar_or_map("SOILPARS", "SeePage")
            self.GWL_base = self._scalar_or_map("SOILPARS", "GWL_base")
            self.SubDrainVel = self.SubKsat * self.Slope                                      # sphy.py:237
        # land use / crop coefficient, sphy.py:254-281
        self.LandUse = cn(self._readmap(cfg.get("LANDUSE", "LandUse")), np.int32)
        if self.DynVegFLAG == 1:                                                               # dynamic_veg.py:52-63
            self.ndvi = cfg.get("DYNVEG", "NDVI")
            self.LAImax = self._lookup_column_table(os.path.join(self.inpath, cfg.get("DYNVEG", "LAImax")), self.LandUse)
            for k in ("NDVImax", "NDVImin", "NDVIbase", "KCmax", "KCmin", "FPARmax", "FPARmin"):
                setattr(self, k, self._scalar_or_map("DYNVEG", k))            # map first, cfg float otherwise (dynamic_veg.py:57-63)
            self.Kc = None                         # per-cell map produced by sphy_veg_step every day
        else:
            self.KcStatFLAG = cfg.getint("LANDUSE", "KCstatic")                                # sphy.py:268-275
            if self.KcStatFLAG == 1:
                self.Kc = self._lookup_column_table(os.path.join(self.inpath, cfg.get("LANDUSE", "CropFac")), self.LandUse)
            else:
                self.Kcmaps = cfg.get("LANDUSE", "KC")
                self.Kc = None                     # daily map series; the last map read is kept over gaps (sphy.py:824-830)
        self.PlantWaterStressFLAG = cfg.getint("PWS", "PWS_FLAG")
        self.PMap = (self._lookup_column_table(os.path.join(self.inpath, cfg.get("PWS", "PFactor")), self.LandUse)
                     if self.PlantWaterStressFLAG == 1 else None)
        # climate, sphy.py:305-370
        flag = lambda sec, k: int(cfg.has_option(sec, k) and cfg.get(sec, k).strip() != "" and cfg.getint(sec, k) == 1)  # noqa: E731
        # forcing key -> (reference forcing name, cfg section) of the variables fed from NetCDF (sphy.py:309-363)
        self._nc_cfg = {}
        if flag("CLIMATE", "precNetcdfFLAG"):
            self._nc_cfg["prec"] = ("Prec", "CLIMATE")
        if flag("CLIMATE", "tempNetcdfFLAG"):
            self._nc_cfg["tavg"] = ("Temp", "CLIMATE")
        if flag("ETREF", "TminNetcdfFLAG"):
            self._nc_cfg["tmin"] = ("Tmin", "ETREF")
        if flag("ETREF", "TmaxNetcdfFLAG"):
            self._nc_cfg["tmax"] = ("Tmax", "ETREF")
        self._nc = {}
        self.Prec = cfg.get("CLIMATE", "Prec") if "prec" not in self._nc_cfg else None         # map-stack prefixes, sphy.py:318,334
        self.Tair = cfg.get("CLIMATE", "Tair") if "tavg" not in self._nc_cfg else None
        self.Pcorr = cfg.getfloat("CLIMATE", "Pcorr_fact")
        self.Tcorr = cfg.getfloat("CLIMATE", "Tcorr_fact")
        self.ETREF_FLAG = cfg.getint("ETREF", "ETREF_FLAG")
        if self.ETREF_FLAG == 0:
            self.Lat = rd("ETREF", "Lat")
            self.Gsc = cfg.getfloat("ETREF", "Gsc")
            self.Tmin = cfg.get("ETREF", "Tmin") if "tmin" not in self._nc_cfg else None
            self.Tmax = cfg.get("ETREF", "Tmax") if "tmax" not in self._nc_cfg else None
        else:
            self.ETref = cfg.get("ETREF", "ETref")
        if self.SnowFLAG == 1:                                                                 # snow.py:84-90
            for k in ("Tcrit", "SnowSC", "DDFS", "SnowF", "SnowCth"):                          # map first, float otherwise
                v = self._scalar_or_map("SNOW", k)
                if isinstance(v, np.ndarray) and v.ndim > 0 and self.GlacFLAG == 1:
                    # glacier.py:285-372 mixes these with pandas table columns: a map cannot work there in the reference either
                    raise ValueError(f"[SNOW] {k} as a map cannot be combined with GlacFLAG = 1 (modules/glacier.py:285)")
                setattr(self, k, v)
            if self.ETREF_FLAG == 1:
                raise ValueError("SnowFLAG = 1 needs ETREF_FLAG = 0: TempMax is only defined on that branch "
                                 "(reference quirk Q3, sphy.py:788-813,875)")
        if self.RoutFLAG == 1 or self.ResFLAG == 1 or self.LakeFLAG == 1:                      # routing.py:33-38
            self.kx = self._scalar_or_map("ROUTING", "kx")
        # open water, sphy.py:416-436
        self.ETOpenWaterFLAG = cfg.getint("OPENWATER", "ETOpenWaterFLAG")
        self.openWaterFrac = None
        if self.ETOpenWaterFLAG == 1:
            if self.ResFLAG != 1:
                raise ValueError("ETOpenWaterFLAG = 1 needs ResFLAG = 1: the open-water classes are labelled with ResID "
                                 "(reference quirk Q21, sphy.py:430)")
            self.kcOpenWater = cfg.getfloat("OPENWATER", "kcOpenWater")
            self._ow_raster = np.asarray(self._readmap(cfg.get("OPENWATER", "openWaterFrac")), np.float32)
            self.openWaterFrac = cn(self._ow_raster)
        # infiltration excess, sphy.py:439-458
        self.InfilFLAG = cfg.getfloat("INFILTRATION", "Infil_excess")
        if self.InfilFLAG == 1:
            self.K_eff = cfg.getfloat("INFILTRATION", "K_eff")
            self.Alpha = cfg.getfloat("INFILTRATION", "Alpha")
            self.Labda_Infil = cfg.getfloat("INFILTRATION", "Labda_infil")
            try:
                self.pavedFrac = self._lookup_column_table(os.path.join(self.inpath, cfg.get("INFILTRATION", "PavedFrac")),
                                                           self.LandUse)
            except (OSError, configparser.Error):
                self.pavedFrac = 0.0

    def _lookup_column_table(self, path, keys):
        """pcr.lookupscalar(table, nominal map) in column-table mode: rows `key value`."""
        table = {}
        with open(path) as fh:
            for line in fh:
                t = line.replace(",", " ").split()
                if t and not t[0].startswith("#"):
                    table[int(float(t[0]))] = float(t[-1])
        if keys.ndim == 0:
            return np.asarray(table.get(int(keys), np.nan), np.float32)
        out = np.full(keys.shape, np.nan, np.float32)
        for k, v in table.items():
            out[keys == k] = f32(v)
        return out

    # -- initial(), sphy.py:565-717 -----------------------------------------------------------------
    def _init_value(self, section, key, default=None):
        raw = self.config.get(section, key) if self.config.has_option(section, key) else ""
        if not raw.strip():
            return default
        try:
            return float(raw)
        except ValueError:
            return self._compact_np(self._readmap(raw))

    def _state(self, value):
        if isinstance(value, np.ndarray) and value.ndim > 0:
            return self._dev(value, np.float32).clone()
        return self._full(f32(value))

    def initial(self):
        cfg = self.config
        self.counter = (self.startdate - self.ts1date).days
        self.curdate = self.startdate
        self.RootWater = self._state(self._init_value("SOIL_INIT", "RootWater", self.RootField))
        self.SubWater = self._state(self._init_value("SOIL_INIT", "SubWater", self.SubField))
        self.CapRise = self._state(self._init_value("SOIL_INIT", "CapRise"))
        self.RootDrain = self._state(self._init_value("SOIL_INIT", "RootDrain"))
        if self.GroundFLAG == 1:
            self.groundwater.initial(self, None, cfg)                                          # groundwater.py:61-86
        else:
            self.SubDrain = self._state(self._init_value("SOIL_INIT", "SubDrain"))
            self.BaseR = self._full(0.0)
            self.GwRecharge = self.Gw = self.H_gw = None
        if self.SnowFLAG == 1:                                                                 # snow.py:94-106
            self.SnowStore = self._state(self._init_value("SNOW_INIT", "SnowIni"))
            self.SnowWatStore = self._state(self._init_value("SNOW_INIT", "SnowWatStore"))
            self.TotalSnowStore = self.SnowStore + self.SnowWatStore
        else:
            self.SnowStore = self.SnowWatStore = self.TotalSnowStore = None
        self.Scanopy = self._full(0.0) if self.DynVegFLAG == 1 else None                       # dynamic_veg.py:66-72
        self.GlacFrac = None
        if self.GlacFLAG == 1:
            self._glacier_initial()
        self._zeros = self._full(0.0)
        if self.LakeFLAG == 1 or self.ResFLAG == 1:
            self.advanced_routing.initial(self, None, cfg)                                     # advanced_routing.py:51-130
            self._reslake_initial()                                                            # lakes / reservoirs .init + .initial
        elif self.RoutFLAG == 1:
            self.routing.initial(self, None, cfg)                                              # routing.py:42-66
        self.waterbalanceTot = self._full(0.0) if self.diagnostics else None
        # per-step outputs (device maps)
        self.TotR = self._full(0.0)
        comp_needed = self.diagnostics or (self.RoutFLAG == 1 and any(self.RA_FLAG.values()))
        self.RootRR = self._full(0.0) if comp_needed else None
        self.RootDR = self._full(0.0) if comp_needed else None
        self.RainR = self._full(0.0) if comp_needed else None
        self.SnowR = self._full(0.0) if (comp_needed and self.SnowFLAG == 1) else None
        self.out = {}
        if self.diagnostics:
            for k in ("ETref", "ETact", "ActETact", "Infil", "Rain", "RootPerc", "SubPerc", "wbal", "Precip"):
                self.out[k] = self._full(0.0)
        if self.ETOpenWaterFLAG == 1:
            self.out["ETOpenWater"] = self._full(0.0)                   # ET.ETpot(ETref, kcOpenWater), sphy.py:898
        self._prepare_wb_structs()
        self._ovl = None
        if (self.overlap_routing and self.RoutFLAG == 1 and self.ResFLAG != 1 and self.LakeFLAG != 1 and self.SedFLAG != 1
                and self.GlacFLAG != 1 and self.route_method == _lib.ROUTE_SCAN and not any(self.RA_FLAG.values())
                and not self.diagnostics and self._share_from is None):
            lo, hi = torch.cuda.Stream.priority_range()
            ev = torch.cuda.Event()
            ev.record()                                        # materialises the cudaEvent_t the library records behind the main pass
            self._ovl = {"stream": torch.cuda.Stream(device=self.device, priority=hi), "main_pass": ev}
            self._ck(self.lib.sphy_set_overlap_tail(self._ctx, 1, C.c_void_p(ev.cuda_event)), "sphy_set_overlap_tail")
        if self.SedFLAG == 1:
            self._sediment_initial()
        # time series at `Locations` (TimeoutputTimeseries, sphy.py:683-691): cells with id > 0, ordered by id
        loc_cells = np.flatnonzero(self.Locations > 0)
        loc_cells = loc_cells[np.argsort(self.Locations[loc_cells], kind="stable")]
        self.station_ids = self.Locations[loc_cells]
        self._station_cells = self._dev(loc_cells.astype(np.int32))
        self._station_buf = torch.empty(max(len(loc_cells), 1), dtype=torch.float32, device=self.device)
        self.tss = {"QallRAtot": []}
        self.timestep = 0
        self.report = None
        rep_csv = os.path.join(self.inpath, self.config.get("REPORTING", "RepTab"))
        if self.enable_reporting and os.path.exists(rep_csv):                                  # reporting.initial, sphy.py:708
            from .reporting import Reporting
            self.report = Reporting(self, rep_csv, write_files=self.write_outputs)
        self._forcing_dev = None
        self._forcing_host = None
        self._forcing_host_compact = False
        self._fdev = None
        self._copy_stream = None
        self._staging = None
        self._prof = None
        self._graphs = None
        if os.environ.get("SPHY_GRAPHS", "0") == "1":
            self.enable_graphs(True)

    # -- glacier.init + glacier.initial, modules/glacier.py:52-252 ------------------------------------
    def _glacier_initial(self):
        import pandas as pd
        cfg = self.config
        tab = pd.read_csv(os.path.join(self.inpath, cfg.get("GLACIER", "GlacTable")))
        tab = tab.sort_values(by="MOD_ID", kind="stable").reset_index(drop=True)
        if self.world > 1:
            # a partitioned basin: every rank keeps the fragments of its own cells (the fragment physics and the per-cell
            # aggregation are local); the per-glacier tables and the yearly retreat work on whole glaciers and are not split
            if cfg.getint("GLACIER", "GlacID_flag") or cfg.getint("GLACIER", "GlacRetreat"):
                raise NotImplementedError("per-glacier tables / glacier retreat are not available on a partitioned basin")
            model_id0 = np.asarray(self._readmap(cfg.get("GLACIER", "ModelID"))).ravel()
            lookup0 = {int(v): i for i, v in enumerate(model_id0)}
            mine = np.array([self._cidx_host[lookup0[int(m)]] >= 0 if int(m) in lookup0 else False for m in tab["MOD_ID"]], bool)
            tab = tab[mine].reset_index(drop=True)
        model_id = np.asarray(self._readmap(cfg.get("GLACIER", "ModelID"))).ravel()
        lookup = {int(v): i for i, v in enumerate(model_id)}                                   # ModelID_1d == ID
        mod = tab["MOD_ID"].to_numpy().astype(np.int64)
        try:
            flat = np.array([lookup[int(m)] for m in mod], np.int64)
        except KeyError as e:
            raise ValueError(f"glacier table MOD_ID {e} is not present in the ModelID map") from None
        cell = self._cidx_host[flat]
        if (cell < 0).any():
            raise ValueError("glacier fragment on an inactive cell")
        uniq, first = np.unique(mod, return_index=True)
        seg_ptr = np.concatenate([first, [len(mod)]]).astype(np.int32)
        seg_cell = cell[first].astype(np.int32)
        for k in ("DDFG", "DDFDG", "GlacF"):
            v = self._scalar_or_map("GLACIER", k)
            if isinstance(v, np.ndarray) and v.ndim > 0:
                # glacier.dynamic multiplies these with pandas table columns (modules/glacier.py:28-48,374-420): a PCRaster
                # map cannot be used there in the reference either
                raise ValueError(f"[GLACIER] {k} must be a number: the reference applies it to the fragment table, not to a map")
            v = float(v)
            setattr(self, k, v)
        lapse = pd.read_csv(os.path.join(self.inpath, cfg.get("GLACIER", "TLapse")), header=None, index_col=0, sep=" ",
                            skipinitialspace=True)
        self.TLapse = lapse.iloc[:, 0].astype(float)
        d64 = lambda a: self._dev(np.asarray(a, np.float64))  # noqa: E731
        n_rows = len(mod)
        self._glac_host = {k: tab[k].to_numpy() for k in ("U_ID", "MOD_ID", "GLAC_ID", "MOD_H", "GLAC_H", "DEBRIS")}
        self._glac_host["ICE_DEPTH"] = (tab["ICE_DEPTH"].to_numpy().astype(np.float64) if "ICE_DEPTH" in tab.columns
                                        else np.zeros(n_rows))
        self._glac_seg = np.repeat(np.arange(len(uniq)), np.diff(seg_ptr))                    # model-cell group of each row
        self._glac_seg_cell = seg_cell
        z64 = lambda: torch.zeros(n_rows, dtype=torch.float64, device=self.device)  # noqa: E731
        G = {"cell": self._dev(cell.astype(np.int32)), "seg_ptr": self._dev(seg_ptr), "seg_cell": self._dev(seg_cell),
             "MOD_H": d64(tab["MOD_H"]), "GLAC_H": d64(tab["GLAC_H"]), "FRAC_GLAC": d64(tab["FRAC_GLAC"]),
             "DEBRIS": self._dev((tab["DEBRIS"].to_numpy() != 0).astype(np.uint8))}
        for k in ("Rain_GLAC", "Snow_GLAC", "AccuGlacMelt", "GlacMelt", "GlacR", "GlacPerc", "SnowR_GLAC",
                  "ActSnowMelt_GLAC", "GLAC_T"):
            G[k] = z64()
        snow0 = self.SnowStore.cpu().numpy()[cell].astype(np.float64)                          # glacier.py:161-184
        wat0 = self.SnowWatStore.cpu().numpy()[cell].astype(np.float64)
        G["SnowStore_GLAC"], G["SnowWatStore_GLAC"] = d64(snow0), d64(wat0)
        G["TotalSnowStore_GLAC"] = d64(snow0 + wat0)
        self.GlacTable = G
        self._glac_n = (n_rows, len(uniq))
        gf = np.zeros(self.n, np.float64)                                                      # glacier.py:188-196
        np.add.at(gf, cell, tab["FRAC_GLAC"].to_numpy().astype(np.float64))
        self.GlacFrac = self._dev(gf.astype(np.float32))
        self._glac_agg = {k: self._full(0.0) for k in ("Rain_GLAC", "Snow_GLAC", "ActSnowMelt_GLAC",
                                                       "TotalSnowStore_GLAC", "SnowR_GLAC", "GlacMelt", "GlacR", "GlacPerc")}
        self.GlacR = self._glac_agg["GlacR"]
        self.TotalSnowStore_GLAC = self._glac_agg["TotalSnowStore_GLAC"]
        self._glacier_reporting_initial(tab)

    def _glacier_reporting_initial(self, tab):
        """glacier.init / initial, the parts behind dynamic_reporting (modules/glacier.py:84-110,146-154,208-252)."""
        cfg, G = self.config, self.GlacTable
        self.glac_files = {}                       # name -> array / DataFrame of everything written around an update
        self.GlacID_flag = cfg.getint("GLACIER", "GlacID_flag")
        self.GlacVars = [v.strip() for v in cfg.get("GLACIER", "GlacVars").split(",") if v.strip()]
        self.GlacRetreat = cfg.getint("GLACIER", "GlacRetreat")
        self.dateAfterUpdate = None
        self._glac_wbal = None
        G["ICE_DEPTH"] = self._dev(self._glac_host["ICE_DEPTH"])
        self._glac_emit({"GlacFrac_" + self.startdate.strftime("%Y%m%d") + ".map": self.GlacFrac.cpu().numpy()})   # glacier.py:196-199
        if self.GlacID_flag:
            gid = self._glac_host["GLAC_ID"]
            self.glacid = sorted(np.unique(gid).tolist())
            rows = np.argsort(gid, kind="stable").astype(np.int32)
            ptr = np.searchsorted(gid[rows], self.glacid + [self.glacid[-1] + 1]).astype(np.int32)
            known = {k: v for k, v in G.items() if v.dtype == torch.float64}
            missing = [v for v in self.GlacVars if v not in known]
            if missing or len(self.GlacVars) > 16:
                raise NotImplementedError(f"GlacVars {missing} are not kept on the device; available: {sorted(known)} (at most 16)")
            ndays = (self.enddate - self.startdate).days + 1
            nv, ng = len(self.GlacVars), len(self.glacid)
            self._glac_rep = dict(rows=self._dev(rows), ptr=self._dev(ptr), nv=nv, ng=ng,
                                  cols=(C.c_void_p * nv)(*[known[v].data_ptr() for v in self.GlacVars]),
                                  is_frac=(C.c_uint8 * nv)(*[int(v == "FRAC_GLAC") for v in self.GlacVars]),
                                  out=torch.zeros((ndays, nv, ng), dtype=torch.float32, device=self.device))
        if self.GlacRetreat == 1:
            day, month = (int(x) for x in cfg.get("GLACIER", "GlacUpdate").split(","))
            self.GlacUpdate = {"month": month, "day": day}
            frac, ice = tab["FRAC_GLAC"].to_numpy().astype(np.float64), self._glac_host["ICE_DEPTH"]
            ice_map = self._by_model_cell(ice * frac)
            area_map = self._by_model_cell(frac * float(self.cellArea))
            tag = self.startdate.strftime("%Y%m%d")
            self._glac_emit({"GlacArea_" + tag + ".map": area_map, "iceDepth_" + tag + ".map": ice_map,
                             "vIce_" + tag + ".map": ice_map * area_map})
            self._old_ice_mm = ice_map * f32(1000)                                             # glacier.py:241
            self._ice_delta = self._full(0.0)

    def _by_model_cell(self, col):
        """groupby(MOD_ID).sum() of a float64 table column (pandas: NaN skipped) scattered to the cells as REAL4."""
        import pandas as pd
        sums = pd.Series(np.asarray(col, np.float64)).groupby(self._glac_seg).sum().to_numpy()
        out = np.zeros(self.n, np.float64)
        out[self._glac_seg_cell] = sums
        return out.astype(np.float32)

    def _glac_emit(self, files):
        """Keep (and, with write_outputs, write) the maps / tables glacier.py reports outside reporting.csv."""
        from . import csf
        self.glac_files.update(files)
        if not self.write_outputs:
            return
        os.makedirs(self.outpath, exist_ok=True)
        for name, v in files.items():
            path = os.path.join(self.outpath, name)
            if name.endswith(".csv"):
                v.to_csv(path, index=False)
            else:
                raster = self.gather_raster(torch.from_numpy(np.ascontiguousarray(v)).to(self.device))   # collective when partitioned
                if raster is not None:
                    csf.write_map(path, raster, cellsize=self.cellsize or 1.0)

    def _glacier_reporting(self, F):
        """glacier.dynamic_reporting (modules/glacier.py:501-810).  It only depends on the fragment table as sphy_glac_step
        left it, so it runs BEFORE the water-balance kernel; what the reference applies to cell maps afterwards
        (TotalSnowStore on the update day, GlacFrac the day after) is returned as a list of deferred actions."""
        import pandas as pd
        G, after = self.GlacTable, []
        if self.GlacID_flag:                                                                   # glacier.py:503-533
            r = self._glac_rep
            day = (self.curdate - self.startdate).days
            self._ck(self.lib.sphy_glac_report(self._ctx, r["ng"], r["ptr"].data_ptr(), r["rows"].data_ptr(),
                                               G["FRAC_GLAC"].data_ptr(), r["nv"], r["cols"], r["is_frac"],
                                               r["out"][day].data_ptr(), self._stream()), "sphy_glac_report")
            if self.curdate == self.enddate and self.write_outputs:                           # glacier.py:534-538
                os.makedirs(self.outpath, exist_ok=True)
                for v, tbl in self.glacier_tables().items():
                    tbl.to_csv(os.path.join(self.outpath, v + ".csv"))
        if self.GlacRetreat != 1:
            return after
        if self._glac_wbal is None:                                                            # sphy.py:1114-1146: oldGw never advances
            self._glac_wbal = self.Gw.clone()
        F.wbal_old_gw = self._glac_wbal.data_ptr()
        if self.curdate.month == self.GlacUpdate["month"] and self.curdate.day == self.GlacUpdate["day"]:
            self.dateAfterUpdate = self.curdate + datetime.timedelta(days=1)                   # glacier.py:561-736
            gid = self._glac_host["GLAC_ID"]
            frac = G["FRAC_GLAC"].cpu().numpy()
            ice = self._glac_host["ICE_DEPTH"]
            snow = (G["TotalSnowStore_GLAC"] - G["AccuGlacMelt"]).cpu().numpy()
            area = float(self.cellArea)
            gsum = lambda x: pd.Series(x).groupby(gid).transform("sum").to_numpy()  # noqa: E731
            with np.errstate(invalid="ignore", divide="ignore"):
                v0 = frac * ice * area                                                         # ice volume before the update, m3
                dmelt = snow / 1000 * frac * area                                              # snow left minus ice melted, m3
                abl = dmelt < 0.0
                acc, vabl = np.where(abl, 0.0, dmelt), np.where(abl, v0, 0.0)
                dmelt_g, acc_g, vabl_g = gsum(dmelt), gsum(acc), gsum(vabl)
                # shrinking glaciers move their accumulation down to the ablation fragments, by ice volume
                redist = np.where((dmelt_g < 0.0) & ~abl, -acc, 0.0)
                redist = np.where((dmelt_g < 0.0) & abl, vabl / vabl_g * acc_g, redist)
                v1 = v0 + acc + np.where(abl, dmelt, 0.0) + redist
                vneg, vpos = np.minimum(0.0, v1), np.maximum(0.0, v1)
                spread = vpos / gsum(vpos) * gsum(vneg)                                        # negative volumes are shared out
                v2 = np.maximum(0.0, v1 + np.where(np.isnan(spread), 0.0, spread))
                ice_new = v2 / (frac * area)
            frac_new = np.where(ice_new <= 0.0, 0.0, frac)                                     # fragments that melted away
            self._glac_host["ICE_DEPTH"] = ice_new
            G["FRAC_GLAC"].copy_(torch.from_numpy(frac_new))
            G["ICE_DEPTH"].copy_(torch.from_numpy(ice_new))
            for k in ("SnowStore_GLAC", "SnowWatStore_GLAC", "TotalSnowStore_GLAC", "AccuGlacMelt"):
                G[k].zero_()
            ice_mm = self._by_model_cell(ice_new * frac_new) * f32(1000)                       # sphy.py:1115-1131
            self._ice_delta.copy_(torch.from_numpy(ice_mm - self._old_ice_mm))
            self._old_ice_mm = ice_mm
            F.wbal_snow_adjust = self.TotalSnowStore_GLAC.data_ptr()
            F.wbal_ice_delta = self._ice_delta.data_ptr()
            after.append(lambda: self.TotalSnowStore.sub_(self.TotalSnowStore_GLAC))           # glacier.py:729
        if self.curdate == self.dateAfterUpdate:                                               # glacier.py:739-810
            tag = self.curdate.strftime("%Y%m%d")
            h = self._glac_host
            frac = G["FRAC_GLAC"].cpu().numpy()
            table = pd.DataFrame({"U_ID": h["U_ID"], "MOD_ID": h["MOD_ID"], "GLAC_ID": h["GLAC_ID"], "MOD_H": h["MOD_H"],
                                  "GLAC_H": h["GLAC_H"], "DEBRIS": h["DEBRIS"], "FRAC_GLAC": frac, "ICE_DEPTH": h["ICE_DEPTH"]})
            gf = self._by_model_cell(frac)
            area_map = self._by_model_cell(frac * float(self.cellArea))
            ice_map = self._by_model_cell(h["ICE_DEPTH"] * frac)
            self._glac_emit({"glacTable_" + tag + ".csv": table, "GlacFrac_" + tag + ".map": gf, "GlacArea_" + tag + ".map": area_map,
                             "iceDepth_" + tag + ".map": ice_map, "vIce_" + tag + ".map": ice_map * area_map})
            self._glac_frac_new = self._dev(gf)
            F.wbal_glac_frac = self._glac_frac_new.data_ptr()                                  # sphy.py:1137 already sees the new map
            after.append(lambda: self.GlacFrac.copy_(self._glac_frac_new))                     # the physics of this day saw the old one
        return after

    def glacier_tables(self):
        """GlacVars name -> DataFrame (dates x glacier ids, float32) like the reference's `<var>_Table` (glacier.py:98-110)."""
        import pandas as pd
        out = self._glac_rep["out"].cpu().numpy()
        idx = pd.date_range(self.startdate, self.enddate, freq="D")
        return {v: pd.DataFrame(out[:, k, :], index=idx, columns=self.glacid, dtype=np.float32) for k, v in enumerate(self.GlacVars)}

    # -- soil erosion and sediment transport, sphy.py:460-559 --------------------------------------------------
    def _sediment_initial(self):
        cfg = self.config
        if self.world > 1:
            raise NotImplementedError("sediment transport is not available on a partitioned basin")
        self.SedModel = cfg.getfloat("SEDIMENT", "SedModel")
        if self.SedModel not in (1, 2):
            raise NotImplementedError("erosion models MUSLE (1) and MMF (2) are supported; INCA/SHETRAN/DHSVM/HSPF are not")
        self.RockFrac = self._compact_np(self._readmap(cfg.get("SEDIMENT", "RockFrac")))
        if self.SedModel == 1:                                                                 # sphy.py:485-493
            if self.SedTransFLAG == 1:
                raise NotImplementedError("sediment transport after MUSLE: sediment_transport.dynamic_musle reads undefined names "
                                          "(modules/sediment_transport.py:227-241) and cannot run in the reference either")
            self.musle.init(self, None, cfg)
            self.sed_out = {}
            return
        self.exclChannelsFLAG = cfg.getint("SEDIMENT", "exclChannelsFLAG")
        if self.exclChannelsFLAG == 1:                                                         # sphy.py:474-482
            ones = torch.ones(self.n, dtype=torch.float32, device=self.device)
            acc = torch.empty_like(ones)
            self._ck(self.lib.sphy_accuflux(self._ctx, ones.data_ptr(), None, acc.data_ptr(), _lib.ROUTE_LEVELS, self._stream()),
                     "sphy_accuflux")
            up = acc.cpu().numpy() * self.cellArea / f32(10 ** 6)
            self.Hillslope = (up <= f32(cfg.getfloat("SEDIMENT", "upstream_km2"))).astype(np.float32)
        self.mmf.init(self, None, cfg)
        if self.SedTransFLAG == 1:
            self.sediment_transport.init(self, None, cfg)
        self.mmf.prepare(self)

    # -- lakes.init/initial, reservoirs.init/initial, modules/lakes.py:162-224, reservoirs.py:90-162 ---
    def _reslake_initial(self):
        cfg = self.config
        part = getattr(self, "_rl_part", False)    # structures on the trunk list: the table covers the WHOLE basin on every rank
        if part and (self.ETOpenWaterFLAG == 1 or (self.LakeFLAG == 1 and cfg.has_option("LAKE", "updatelakelevel")
                                                   and cfg.get("LAKE", "updatelakelevel").strip())):
            raise NotImplementedError("open-water evaporation and measured lake levels are not available with the structures on "
                                      "the trunk list (a partitioned basin)")
        n = self.n_global if part else self.n
        res_id = np.zeros(n, np.int32)
        lake_id = np.zeros(n, np.int32)
        stor = np.zeros(n, np.float32)
        million = f32(10 ** 6)
        rows = {}                                   # cell -> parameter column
        if self.LakeFLAG == 1:
            lake_id = (self._rl_global["lake"] if part else
                       np.maximum(self._compact_np(self._readmap(cfg.get("LAKE", "LakeId")), np.int32), 0))
            T = {k: _read_matrix_table(os.path.join(self.inpath, cfg.get("LAKE", k)))
                 for k in ("LakeStor", "LakeFunc", "LakeQH", "LakeSH", "LakeHS")}
            st = _table_col(T["LakeStor"], 1)
            for c in np.flatnonzero(lake_id > 0):
                lid = int(lake_id[c])
                stor[c] = f32(st.get(lid, 0.0)) * million                                     # lakes.py:220-222
                col = np.full(_lib.RL_NPAR, np.nan, np.float32)
                col[9] = _table_col(T["LakeFunc"], 3).get(lid, np.nan)                         # HS func
                for j in range(6):
                    col[10 + j] = _table_col(T["LakeHS"], j + 1).get(lid, np.nan)              # exp_a exp_b pol_b a1 a2 a3
                col[16] = _table_col(T["LakeFunc"], 1).get(lid, np.nan)                        # QH func
                qh = [_table_col(T["LakeQH"], j + 1).get(lid, np.nan) for j in range(6)]
                col[17], col[18], col[19], col[20], col[21], col[22] = qh
                col[23] = _table_col(T["LakeFunc"], 2).get(lid, np.nan)                        # SH func (measured levels)
                for j in range(6):
                    col[24 + j] = _table_col(T["LakeSH"], j + 1).get(lid, np.nan)              # exp_a exp_b pol_b a1 a2 a3
                rows[int(c)] = (3, col)
            self._lake_update = None                                                           # lakes.py:205-213
            if cfg.has_option("LAKE", "updatelakelevel") and cfg.get("LAKE", "updatelakelevel").strip():
                try:
                    self._lake_update = self._compact_np(self._readmap(cfg.get("LAKE", "updatelakelevel")), np.float32) != 0
                    self.LLevel = cfg.get("LAKE", "LakeFile")
                except (FileNotFoundError, KeyError, configparser.Error):
                    self._lake_update = None
        if self.ResFLAG == 1:
            res_id = (self._rl_global["res"] if part else
                      np.maximum(self._compact_np(self._readmap(cfg.get("RESERVOIR", "ResId")), np.int32), 0))
            fs = _read_matrix_table(os.path.join(self.inpath, cfg.get("RESERVOIR", "ResFuncStor")))
            func, st = _table_col(fs, 1), _table_col(fs, 2)
            simple = adv = None
            if cfg.has_option("RESERVOIR", "ResSimple") and cfg.get("RESERVOIR", "ResSimple").strip():
                simple = _read_matrix_table(os.path.join(self.inpath, cfg.get("RESERVOIR", "ResSimple")))
            if cfg.has_option("RESERVOIR", "ResAdv") and cfg.get("RESERVOIR", "ResAdv").strip():
                adv = _read_matrix_table(os.path.join(self.inpath, cfg.get("RESERVOIR", "ResAdv")))
            for c in np.flatnonzero(res_id > 0):
                rid = int(res_id[c])
                stor[c] = stor[c] + f32(st.get(rid, 0.0)) * million                           # reservoirs.py:152-156
                col = np.full(_lib.RL_NPAR, np.nan, np.float32)
                fn = int(func.get(rid, 0))
                kind = 0
                if fn == 1 and simple is not None:
                    kind = 1
                    col[0] = _table_col(simple, 1).get(rid, np.nan)
                    col[1] = _table_col(simple, 2).get(rid, np.nan)
                    col[2] = f32(_table_col(simple, 3).get(rid, np.nan)) * million
                elif fn == 2 and adv is not None:
                    kind = 2
                    for j in range(4):
                        col[3 + j] = f32(_table_col(adv, j + 1).get(rid, np.nan)) * million
                    col[7] = _table_col(adv, 5).get(rid, np.nan)
                    col[8] = _table_col(adv, 6).get(rid, np.nan)
                rows[int(c)] = (kind, col)                                                    # ResID != 0 wins over LakeID
        cells = np.array(sorted(rows), np.int32)
        ns = len(cells)
        par = np.full((_lib.RL_NPAR, max(ns, 1)), np.nan, np.float32)
        kind = np.zeros(max(ns, 1), np.int32)
        for j, c in enumerate(cells):
            kind[j], par[:, j] = rows[int(c)]
        qfrac = np.ones(n, np.float32)
        qfrac[cells] = 0.0
        lo, hi = getattr(self, "part_range", (0, n))
        self.QFRAC = self._dev(qfrac[lo:hi])
        self._rl_cells_host = cells                # GLOBAL device indices when the structures sit on the trunk list
        self._rl = {"cell": self._dev(cells if ns else np.zeros(1, np.int32)), "kind": self._dev(kind), "par": self._dev(par),
                    "StorRES": self._dev(stor[cells] if ns else np.zeros(1, np.float32)),
                    "Qout": torch.zeros(max(ns, 1), dtype=torch.float32, device=self.device),
                    "Qin": torch.zeros(max(ns, 1), dtype=torch.float32, device=self.device)}
        if self.LakeFLAG == 1 and self._lake_update is not None:
            upd = np.broadcast_to(self._lake_update, (n,))[cells].astype(np.int32) if ns else np.zeros(1, np.int32)
            self._rl["update_level"] = self._dev(upd)
            self._rl["lake_level"] = torch.zeros(max(ns, 1), dtype=torch.float32, device=self.device)
            self._rl["cells64"] = self._dev(cells.astype(np.int64) if ns else np.zeros(1, np.int64))
        if self.ETOpenWaterFLAG == 1:
            self._openwater_initial(cells)
        r = self._rl
        self._rl_struct = _lib.ResLake(ns, r["cell"].data_ptr(), r["kind"].data_ptr(), r["par"].data_ptr(),
                                       self.QFRAC.data_ptr(), r["StorRES"].data_ptr(), r["Qout"].data_ptr(), r["Qin"].data_ptr(),
                                       None, r["update_level"].data_ptr() if "update_level" in r else None)
        if part:
            self._reslake_plan_initial(cells, qfrac[lo:hi], lo, hi)

    def _reslake_plan_initial(self, cells, qfrac_local, lo, hi):
        """Device tables of multigpu.plan_reslake (sphy_reslake_plan) and the local workspaces of the structure path."""
        ph = self._rl_plan_host
        if not np.array_equal(cells, self._rl_global["cut"]):
            raise RuntimeError("structure cells changed between the trunk selection and the table build")
        listed = ph["listed"]
        mine = listed[(listed >= lo) & (listed < hi)] - lo
        qmask = qfrac_local.copy()
        qmask[mine] = 2.0                          # the main pass leaves listed cells to the trunk kernels
        ns, m = len(cells), len(listed)
        i32 = lambda a: self._dev(np.ascontiguousarray(a if len(a) else np.zeros(1), np.int32))  # noqa: E731
        p = {k: i32(ph[k]) for k in ("cut_slot", "entry_cut", "entry_ptr", "entry_idx", "x_rank", "x_local")}
        p["qmask"] = self._dev(qmask)
        p["qday"] = self._full(0.0)
        p["extra"] = torch.zeros(max(ns * _lib.RL_EXTRA, 1), dtype=torch.float64, device=self.device)
        p["top"] = torch.zeros(max(m, 1), dtype=torch.float64, device=self.device)
        self._rl_plan_dev = p
        self._rl_plan = _lib.ResLakePlan(*[p[k].data_ptr() if k in p else None for k in
                                           ("cut_slot", "entry_cut", "entry_ptr", "entry_idx", "x_rank", "x_local", "qmask", "material", "qday",
                                            "extra", "top")])
        self._rl_pending = False
        # the slots of the exchange extras this rank owns (flush_structures)
        xr, xl = ph["x_rank"].astype(np.int64), ph["x_local"].astype(np.int64)
        donor_slot = (np.arange(xr.size) % _lib.RL_EXTRA) < _lib.RL_EXTRA - 1
        own = np.flatnonzero((xr == self.rank) & donor_slot)
        self._rl_flush = (self._dev(own), self._dev(xl[own]))

    def _openwater_initial(self, struct_cells):
        """sphy.py:425-431: 8-connected clumps of the open-water mask, each labelled with the largest ResID inside;
        per structure the list of cells of its class with openWaterFrac > 0 (what areatotal sums over)."""
        from scipy import ndimage
        mask = self._ow_raster > 0
        clumps = np.zeros(mask.shape, np.int64)
        nxt = 0
        for val in (False, True):
            lab, cnt = ndimage.label(mask == val, structure=np.ones((3, 3), dtype=int))
            clumps[lab > 0] = lab[lab > 0] + nxt
            nxt += cnt
        res_raster = np.maximum(np.asarray(self._readmap(self.config.get("RESERVOIR", "ResId"))), 0).astype(np.int64)
        best = np.zeros(nxt + 1, np.int64)
        np.maximum.at(best, clumps.ravel(), res_raster.ravel())
        cls = self._compact_np(best[clumps], np.int64)                 # class of every active cell (device order)
        wet = np.flatnonzero(self.openWaterFrac > 0)
        by_class = {}
        for c in wet:
            by_class.setdefault(int(cls[c]), []).append(int(c))
        ptr, lst = [0], []
        for c in struct_cells:
            lst.extend(by_class.get(int(cls[int(c)]), []))
            ptr.append(len(lst))
        self._ow_ptr = self._dev(np.asarray(ptr, np.int32))
        self._ow_cell = self._dev(np.asarray(lst if lst else [0], np.int32))
        self._ow_frac_dev = self._dev(self.openWaterFrac)

    @property
    def StorRES(self):
        """Dense view of the lake/reservoir storage map (reference attribute name), m3."""
        out = torch.zeros(self.n, dtype=torch.float32, device=self.device)
        cells = self._rl_cells_host.astype(np.int64)
        if getattr(self, "_rl_part", False):       # structures on the trunk list: global indices, storage possibly one update behind
            self.flush_structures()
            lo, hi = getattr(self, "part_range", (0, self.n))
            sel = np.flatnonzero((cells >= lo) & (cells < hi))
            if len(sel):
                out[torch.from_numpy(cells[sel] - lo).to(self.device)] = self._rl["StorRES"][torch.from_numpy(sel).to(self.device)]
            return out
        if len(cells):
            out[torch.from_numpy(cells).to(self.device)] = self._rl["StorRES"]
        return out

    # -- kernel argument structs ---------------------------------------------------------------------
    def _prepare_wb_structs(self):
        P = _lib.WbParams()
        flags = 0
        if self.SnowFLAG == 1:
            flags |= _lib.WB_SNOW
        if self.GroundFLAG == 1:
            flags |= _lib.WB_GROUND
        if self.InfilFLAG == 1:
            flags |= _lib.WB_INFIL_EXCESS
        if self.PlantWaterStressFLAG == 1:
            flags |= _lib.WB_PWS
        if self.ETREF_FLAG == 1 or self.DynVegFLAG == 1:   # with dynamic vegetation ETref comes from sphy_veg_step
            flags |= _lib.WB_ETREF_INPUT
        P.flags = flags
        if self.ETREF_FLAG == 0:
            try:                                   # hash-based, O(N): np.unique sorts 1e8 values for ~10 s
                import pandas as pd
                inv, uniq = pd.factorize(np.ascontiguousarray(self.Lat), sort=False)
            except ImportError:
                uniq, inv = np.unique(self.Lat, return_inverse=True)
            if self.ra_mode == "table" and len(uniq) > 65536:
                raise ValueError("ra_mode='table' needs at most 65536 distinct latitudes")
            if (len(uniq) <= 65536 and self.ra_mode != "cell"):
                P.ra_mode = _lib.RA_TABLE
                self._lat_class = self._dev(inv.astype(np.uint16))
                self._lat_unique = self._dev(uniq.astype(np.float32))
                P.lat_class, P.lat_unique_deg, P.n_lat_unique = self._lat_class.data_ptr(), self._lat_unique.data_ptr(), len(uniq)
            else:
                P.ra_mode = _lib.RA_CELL
                lat_d = self._dev(self.Lat)
                self._lat_trig = [torch.empty_like(lat_d) for _ in range(3)]
                for op, t in zip((_lib.UNARY_SIN_DEG, _lib.UNARY_COS_DEG, _lib.UNARY_TAN_DEG), self._lat_trig):
                    self._ck(self.lib.sphy_map_unary(self._ctx, op, lat_d.data_ptr(), t.data_ptr(), self.n, self._stream()),
                             "sphy_map_unary")
                P.sin_lat, P.cos_lat, P.tan_lat = (t.data_ptr() for t in self._lat_trig)
            P.k0 = float(f32(((24 * 60) / math.pi) * self.Gsc))                              # hargreaves.py:32-33
        P.pcorr, P.tcorr = float(f32(self.Pcorr)), float(f32(self.Tcorr))
        self._pcorr = P.pcorr
        pr = self._param
        if self.DynVegFLAG == 1:
            self._prepare_veg_struct(P)
        P.root_field, P.root_sat, P.root_dry, P.root_wilt = pr(self.RootField), pr(self.RootSat), pr(self.RootDry), pr(self.RootWilt)
        P.root_ksat, P.root_drainvel, P.e_root = pr(self.RootKsat), pr(self.RootDrainVel), pr(self._e_root)
        P.sub_field, P.sub_sat, P.e_sub = pr(self.SubField), pr(self.SubSat), pr(self._e_sub)
        P.caprise_max = pr(self.CapRiseMax)
        self._series = {}                          # optional daily map series: key -> (stack prefix, device map holding the last one read)
        if self.DynVegFLAG == 1:
            P.kc = SphyParam(self._veg_maps["Kc"].data_ptr(), 0.0)
        elif self.Kc is None:
            self._series["kc"] = (self.Kcmaps, self._full(1.0))                                # KcOld = 1, sphy.py:611-613
            P.kc = SphyParam(self._series["kc"][1].data_ptr(), 0.0)
        else:
            P.kc = pr(self.Kc)
        if self.PMap is not None:
            P.pmap = pr(self.PMap)
        if self.GlacFrac is not None:
            P.glac_frac = SphyParam(self.GlacFrac.data_ptr(), 0.0)
        if self.openWaterFrac is not None:
            P.open_water_frac = SphyParam(self._ow_frac_dev.data_ptr(), 0.0)
            P.kc_open_water = float(f32(self.kcOpenWater))
        if self.InfilFLAG == 1:
            P.k_eff, P.alpha = float(f32(self.K_eff)), float(f32(self.Alpha))
            P.alpha_sq, P.labda_infil = float(f32(self.Alpha ** 2)), float(f32(self.Labda_Infil))
            P.paved_frac = pr(self.pavedFrac)
        dbl_exp = lambda x: np.exp(np.asarray(x, np.float32).astype(np.float64)).astype(np.float32)  # noqa: E731
        if self.GroundFLAG == 1:
            P.gw_sat, P.base_thresh = pr(self.GwSat), pr(self.BaseThresh)
            dg, ag, yg = self.deltaGw, self.alphaGw, self.YieldGw
            # exp(-1/deltaGw): `-1/deltaGw` is a Python double when deltaGw is a cfg float (groundwater.py:27)
            e_delta = dbl_exp(f32(-1) / dg) if isinstance(dg, np.ndarray) else dbl_exp(-1 / dg)
            e_alpha = dbl_exp(-ag)
            if isinstance(yg, np.ndarray) or isinstance(ag, np.ndarray):
                fl = lambda x: x if isinstance(x, np.ndarray) else f32(x)  # noqa: E731
                h_den = f32(800) * fl(yg) * fl(ag)
            else:
                h_den = f32(800 * yg * ag)                                                     # groundwater.py:45
            P.e_delta, P.e_alpha, P.h_den = pr(e_delta), pr(e_alpha), pr(h_den)
        else:
            if self.SeePage is None:
                self._series["seep"] = (self.Seepmaps, self._full(0.0))                        # SeepOld = 0, sphy.py:628-629
                P.seepage = SphyParam(self._series["seep"][1].data_ptr(), 0.0)
            else:
                P.seepage = pr(self.SeePage)
            P.sub_drainvel = pr(self.SubDrainVel)
        if self.SnowFLAG == 1:
            P.tcrit, P.snow_sc, P.ddfs, P.snow_f = pr(self.Tcrit), pr(self.SnowSC), pr(self.DDFS), pr(self.SnowF)
            if not P.snow_f.map:                               # a cfg float: Python-double subtraction, then REAL4 (snow.py:182)
                sf_ = self.SnowF
                P.one_minus_snow_f = (float(f32(1) - f32(P.snow_f.value)) if isinstance(sf_, np.ndarray) else float(f32(1 - sf_)))
            for h in range(1, 25):                                                             # snow.py:30 (3.1415)
                P.melt_cos[h - 1] = float(f32(math.cos(float(f32(3.1415 * h / 12)))))
        if self.LakeFLAG == 1 and getattr(self, "_lake_update", None) is not None:
            self._series["llev"] = (self.LLevel, None)         # measured lake levels: used on the day they exist only
        self._wbP = P
        ptr = lambda t: t.data_ptr() if t is not None else None  # noqa: E731
        S = _lib.WbState()
        for k in ("RootWater", "SubWater", "CapRise", "RootDrain", "GwRecharge", "Gw", "BaseR", "H_gw", "SubDrain",
                  "SnowStore", "SnowWatStore", "TotalSnowStore", "waterbalanceTot"):
            setattr(S, k, ptr(getattr(self, k)))
        if self.GroundFLAG != 1:
            S.BaseR = None
        self._wbS = S
        O = _lib.WbOut()
        O.TotR = self.TotR.data_ptr()
        for k in ("RootRR", "RootDR", "RainR", "SnowR"):
            setattr(O, k, ptr(getattr(self, k)))
        for k, t in self.out.items():
            setattr(O, k, t.data_ptr())
        self._wbO = O
        if self.GlacFLAG == 1:
            G = self.GlacTable
            t = _lib.GlacTable()
            t.n_rows, t.n_seg = self._glac_n
            for k in ("cell", "seg_ptr", "seg_cell", "MOD_H", "GLAC_H", "FRAC_GLAC", "DEBRIS", "Rain_GLAC", "Snow_GLAC",
                      "SnowStore_GLAC", "SnowWatStore_GLAC", "TotalSnowStore_GLAC", "AccuGlacMelt", "GlacMelt", "GlacR",
                      "GlacPerc", "SnowR_GLAC", "ActSnowMelt_GLAC", "GLAC_T"):
                setattr(t, k, G[k].data_ptr())
            t.tcrit, t.ddfs, t.snow_sc = self.Tcrit, self.DDFS, self.SnowSC
            t.ddfg, t.ddfdg, t.glac_f = self.DDFG, self.DDFDG, self.GlacF
            self._glT = t
            o = _lib.GlacOut()
            for k, v in self._glac_agg.items():
                setattr(o, k, v.data_ptr())
            self._glO = o
        if self.RoutFLAG == 1 or self.ResFLAG == 1 or self.LakeFLAG == 1:
            self._kx = pr(self.kx)
            if self._kx.map:
                self._one_m_kx = 0.0                       # per-cell 1 - kx is formed in the kernel (REAL4)
            elif isinstance(self.kx, np.ndarray):          # spatially uniform map: `1 - kx` was a REAL4 map operation
                self._one_m_kx = float(f32(1) - f32(self._kx.value))
            else:                                          # cfg float: Python-double subtraction, then REAL4 (routing.py:28)
                self._one_m_kx = float(f32(1 - self.kx))

    def _prepare_veg_struct(self, P):
        """sphy_veg: cfg floats folded like the reference's Python expressions (dynamic_veg.py:30-40).  The veg kernel
        keeps the raw Pcorr/Tcorr and latitude fields; K1 and the glacier kernel then see pcorr = 1 because they consume
        the effective precipitation."""
        Pv = _lib.WbParams()
        C.memmove(C.byref(Pv), C.byref(P), C.sizeof(P))
        self._vegP = Pv
        P.pcorr = 1.0
        z = lambda: torch.zeros(self.n, dtype=torch.float32, device=self.device)  # noqa: E731
        self._veg_maps = {k: z() for k in ("ETref", "Kc", "PrecipEff", "LAI", "Int")}
        # cfg floats combine as Python doubles and become REAL4 where they meet a map; parameters read as maps are float32
        # arrays (0-d for a uniform map) and combine in REAL4 map arithmetic, exactly like the reference's expressions
        ndvi0 = (self.NDVImax + self.NDVImin) / 2                                             # dynamic_veg.py:70
        self._ndvi_old = (self._dev(np.broadcast_to(np.asarray(ndvi0, np.float32), (self.n,)).copy()) if isinstance(ndvi0, np.ndarray)
                          else torch.full((self.n,), float(f32(ndvi0)), dtype=torch.float32, device=self.device))
        V = _lib.Veg()
        V.lai_max = self._param(self.LAImax)
        if self.openWaterFrac is not None:
            V.open_water_frac = SphyParam(self._ow_frac_dev.data_ptr(), 0.0)
        with np.errstate(all="ignore"):
            sr_max = (1 + self.NDVImax) / (1 - self.NDVImax)
            sr_min = (1 + self.NDVImin) / (1 - self.NDVImin)
            one_m_f = 1 - self.FPARmax
            log_f = (np.log10(one_m_f.astype(np.float64)).astype(np.float32) if isinstance(one_m_f, np.ndarray)
                     else f32(np.log10(np.float64(f32(one_m_f)))))                           # double on the REAL4 operand
            pr = self._param
            V.sr_min, V.sr_range = pr(sr_min), pr(sr_max - sr_min)
            V.fpar_range = pr(self.FPARmax - self.FPARmin)
            V.log_one_m_fparmax = pr(log_f)
            V.kc_min, V.kc_range = pr(self.KCmin), pr(self.KCmax - self.KCmin)
            V.ndvi_min, V.ndvi_range = pr(self.NDVImin), pr(self.NDVImax - self.NDVImin)
        V.Scanopy = self.Scanopy.data_ptr()
        for k, t in self._veg_maps.items():
            setattr(V, k, t.data_ptr())
        self._veg = V

    def _veg_step(self, f, doy):
        """dynamic_veg.dynamic (modules/dynamic_veg.py:76-128) -> ETref, Kc and effective precipitation maps."""
        V = self._veg
        ndvi = f.get("ndvi")
        if ndvi is None:                                                                       # dynamic_veg.py:78-82
            ndvi = self._ndvi_old
        elif ndvi is not self._ndvi_old:
            self._ndvi_old.copy_(ndvi, non_blocking=True)
        V.ndvi = self._ndvi_old.data_ptr()
        V.prec, V.tavg = f["prec"].data_ptr(), f["tavg"].data_ptr()
        if self.ETREF_FLAG == 1:
            V.etref_in = f["etref"].data_ptr()
        else:
            V.tmin, V.tmax = f["tmin"].data_ptr(), f["tmax"].data_ptr()
        self._ck(self.lib.sphy_veg_step(self._ctx, C.byref(self._vegP), C.byref(V), doy, self._stream()), "sphy_veg_step")
        return self._veg_maps

    # -- forcing ---------------------------------------------------------------------------------------
    def set_device_forcing(self, fn):
        """fn(counter) -> dict of compact float32 CUDA tensors (prec, tavg, tmin, tmax[, etref]) already resident in HBM."""
        self._forcing_dev = fn

    def set_host_forcing(self, fn, compact=False):
        """fn(counter) -> dict of float32 PINNED host tensors holding the day's forcing RASTERS (nrows*ncols, the
        layout readmap delivers); each step copies them host->device and compacts them into device cell order."""
        self._forcing_host = fn
        self._forcing_host_compact = compact      # True: fn already returns this rank's cells in device order
        self._copy_stream = None

    def enable_profiling(self, on=True):
        """Record CUDA events around the water-balance and routing launches of every dynamic() call."""
        self._prof = {"wb": [], "route": []} if on else None

    def _forcing_names(self, counter):
        from . import csf
        if self.ETREF_FLAG == 1:
            pre = {"prec": self.Prec, "tavg": self.Tair, "etref": self.ETref}
        else:
            pre = {"prec": self.Prec, "tavg": self.Tair, "tmin": self.Tmin, "tmax": self.Tmax}
        if self.DynVegFLAG == 1:
            pre["ndvi"] = self.ndvi
        for k, (prefix, _) in self._series.items():
            pre[k] = prefix
        return {k: csf.stack_name(v, counter) for k, v in pre.items() if v is not None}

    def _ingest(self, counter):
        """The day's forcing as compact device maps.  Three sources: maps already resident in HBM
        (set_device_forcing), pinned host rasters (set_host_forcing: asynchronous H2D on a copy stream, double
        buffered so that day t+1 uploads while day t computes, then compaction into device cell order), or the
        map source (readmap of `generateNameT(prefix, counter)` like sphy.py:755-809)."""
        if self._forcing_dev is not None:
            return self._forcing_dev(counter)
        keys = ("prec", "tavg", "etref") if self.ETREF_FLAG == 1 else ("prec", "tavg", "tmin", "tmax")
        optional = (("ndvi",) if self.DynVegFLAG == 1 else ()) + tuple(self._series)           # series that may have gaps
        self._optional_keys = optional
        keys = keys + optional
        if self._fdev is None:
            self._fdev = {k: torch.empty(self.n, dtype=torch.float32, device=self.device) for k in keys}
        if self._forcing_host is not None:
            return self._ingest_host(counter, keys)
        if self._staging is None:            # two pinned sets: the host may refill one while the other's upload is still queued
            self._staging = [{k: torch.empty(self.n, dtype=torch.float32).pin_memory() for k in keys} for _ in range(2)]
            self._staging_ev = [None, None]
        slot = counter % 2
        if self._staging_ev[slot] is not None:
            self._staging_ev[slot].synchronize()               # the upload that last used this set has completed
        staging = self._staging[slot]
        names = self._forcing_names(counter)
        out = dict(self._fdev)
        for k in keys:
            if k in self._nc_cfg:                  # NetCDF forcing: interpolated on the device (netcdf2PCraster.netcdf2pcrDynamic)
                if k not in self._nc:
                    from .netcdf_forcing import NetcdfForcing
                    self._nc[k] = NetcdfForcing(self, *self._nc_cfg[k])
                out[k] = self._nc[k].day(self.curdate)
                continue
            try:
                src = np.asarray(self.maps.read(names[k]), np.float32).ravel()
            except (KeyError, FileNotFoundError):
                if k not in optional:
                    raise
                out.pop(k)                         # no map for this day: the previous one is kept (dynamic_veg.py:78-82, sphy.py:824-830)
                continue
            staging[k].numpy()[:] = src if self.full_raster else src[self.flat_active]
            self._fdev[k].copy_(staging[k], non_blocking=True)
        ev = torch.cuda.Event()
        ev.record(torch.cuda.current_stream(self.device))
        self._staging_ev[slot] = ev
        return out

    def _ingest_host(self, counter, keys):
        """Pinned host rasters -> compact device maps.  The device pulls exactly its own cells out of the pinned raster
        (sphy_gather_indexed_f32 over unified addressing, cells visited in raster order so that the PCIe reads coalesce)
        on a copy stream, double buffered: day t+1 is fetched while day t computes.  A partitioned rank therefore moves
        1 / world of every raster, with the host holding plain rasters as `readmap` delivers them."""
        compact = self._forcing_host_compact
        main = torch.cuda.current_stream(self.device)
        if self._copy_stream is None:
            self._copy_stream = torch.cuda.Stream(device=self.device)
            self._hbuf = [{k: torch.empty(self.n, dtype=torch.float32, device=self.device) for k in keys} for _ in range(2)]
            self._hready = [None, None]          # (counter, event) of the upload sitting in each buffer
            self.h2d_bytes_per_map = self.n * 4
            if not compact:
                order = np.argsort(self.flat_active, kind="stable")
                self._pull_src = self._dev(self.flat_active[order].astype(np.int32))     # raster index, ascending
                self._pull_dst = self._dev(order.astype(np.int32))                       # device cell of each
        cs = self._copy_stream

        def upload(t, slot):
            try:
                host = self._forcing_host(t)
            except (IndexError, KeyError, FileNotFoundError):
                return
            cs.wait_stream(main)                 # the buffer's previous contents have been consumed
            present = []
            with torch.cuda.stream(cs):
                for k in keys:
                    if k in self._optional_keys and host.get(k) is None:
                        continue                 # a map series without a map for this day
                    if compact:
                        self._hbuf[slot][k].copy_(host[k], non_blocking=True)
                    else:
                        if not host[k].is_pinned():
                            raise ValueError("set_host_forcing: the host rasters must be pinned (torch.Tensor.pin_memory())")
                        self._ck(self.lib.sphy_gather_indexed_f32(self._ctx, host[k].data_ptr(), self._pull_src.data_ptr(),
                                                                  self._pull_dst.data_ptr(), self.n, self._hbuf[slot][k].data_ptr(),
                                                                  C.c_void_p(cs.cuda_stream)), "sphy_gather_indexed_f32")
                    present.append(k)
                ev = torch.cuda.Event()
                ev.record(cs)
            self._hready[slot] = (t, ev, present, host)     # `host` stays referenced until the slot is reused: the pull reads it asynchronously

        slot = counter % 2
        if self._hready[slot] is None or self._hready[slot][0] != counter:
            upload(counter, slot)
        if self._hready[slot] is None or self._hready[slot][0] != counter:
            raise FileNotFoundError(f"no forcing for time step {counter}: the host forcing source has no entry for it")
        main.wait_event(self._hready[slot][1])
        buf, present = self._hbuf[slot], self._hready[slot][2]
        out = {k: buf[k] for k in present}
        upload(counter + 1, 1 - slot)            # prefetch tomorrow while today computes
        return out

    # -- dynamic(), sphy.py:719-1263 ------------------------------------------------------------------------
    def dynamic(self):
        f, doy = self._pre_step()
        if self._graphs is not None:
            StepGraph.run(self._graphs, [self], [f], doy, capture=self._graph_ok())
        else:
            self._device_step(f, doy)
        return self._post_step()

    def _pre_step(self):
        """Host bookkeeping of one step and the day's forcing as compact device maps."""
        self.counter += 1
        f = self._ingest(self.counter)
        doy = julian(self.curdate)
        self._doy = doy
        return f, doy

    def _device_step(self, f, doy, route=True, ra_table=True):
        """Everything of one daily step that runs on the device (sphy.py:732-1240): [vegetation] -> [glaciers] -> fused
        vertical water balance -> routing -> [erosion / sediment transport].  No host synchronisation.  `route=False`
        leaves plain routing to the caller (an Ensemble routes its members together); `ra_table=False` reuses the context's
        per-day Ra table (computed by another member sharing the latitude map)."""
        lib, ctx, st = self.lib, self._ctx, self._stream()
        F = _lib.WbForcing()
        F.prec, F.tavg = f["prec"].data_ptr(), f["tavg"].data_ptr()
        self._prec_ptr = F.prec
        if self.ETREF_FLAG == 1:
            F.etref = f["etref"].data_ptr()
        else:
            F.tmin, F.tmax = f["tmin"].data_ptr(), f["tmax"].data_ptr()
        self._prec_raw = f["prec"]
        for k, (_, last) in self._series.items():                                              # sphy.py:824-830, 992-1000
            if last is not None and f.get(k) is not None and f[k] is not last:
                last.copy_(f[k], non_blocking=True)
        if "llev" in self._series:                                                             # lakes.py:31-36
            lev = f.get("llev")
            if lev is not None:
                torch.index_select(lev, 0, self._rl["cells64"], out=self._rl["lake_level"])   # plumbing: levels at the structures
                self._rl_struct.lake_level = self._rl["lake_level"].data_ptr()
            else:
                self._rl_struct.lake_level = None
        if self.DynVegFLAG == 1:                                                               # sphy.py:820-822
            v = self._veg_step(f, doy)
            f = dict(f, prec=v["PrecipEff"], etref=v["ETref"])
            F.prec, F.etref = f["prec"].data_ptr(), f["etref"].data_ptr()
            if self.ETREF_FLAG != 1 and self.SnowFLAG != 1:
                F.tmin = F.tmax = None
        if self.GlacFLAG == 1:                                                                 # sphy.py:843-853
            a = self._glacier_step(f["tavg"], f["prec"])
            F.TotalSnowStore_GLAC, F.SnowR_GLAC = a["TotalSnowStore_GLAC"].data_ptr(), a["SnowR_GLAC"].data_ptr()
            F.GlacPerc, F.GlacR = a["GlacPerc"].data_ptr(), a["GlacR"].data_ptr()
        deferred = []
        if self.GlacFLAG == 1 and self.RoutFLAG == 1 and not (self.LakeFLAG == 1 or self.ResFLAG == 1):
            deferred = self._glacier_reporting(F)                                              # sphy.py:1106-1112
        prof = self._prof
        if self._ovl is not None and route:
            return self._device_step_overlapped(F, doy, ra_table)
        if prof is not None:
            ev = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
            ev[0].record()
        self._ck(lib.sphy_wb_step(ctx, C.byref(self._wbP), C.byref(self._wbS), C.byref(F), C.byref(self._wbO),
                                  doy if ra_table else -1, st), "sphy_wb_step")
        if prof is not None:
            ev[1].record()
        for fn in deferred:
            fn()
        self.Q = None
        if self.LakeFLAG == 1 or self.ResFLAG == 1:                                            # sphy.py:1100-1104
            self.Q = self.advanced_routing.dynamic(self, None, None, self.config, self.TotR, None, None)
        elif self.RoutFLAG == 1 and route:                                                     # sphy.py:1107-1108
            self.Q = self.routing.dynamic(self, None, self.TotR)
        elif self.RoutFLAG == 1:
            self.Q = self.QRAold
        if prof is not None:
            ev[2].record()
            prof["wb"].append((ev[0], ev[1]))
            prof["route"].append((ev[1], ev[2]))
        if self.SedFLAG == 1 and self.SedModel == 1:                                           # sphy.py:1222-1225
            self.musle.dynamic(self, None, self.Q if self.Q is not None else self.TotR)
        elif self.SedFLAG == 1:                                                                # sphy.py:1214-1240
            G = self.mmf.dynamic(self, None, f["prec"], self.Q if self.Q is not None else self.TotR)
            if self.SedTransFLAG == 1:
                self.sediment_transport.dynamic_mmf(self, None, None, None, G)

    def _device_step_overlapped(self, F, doy, ra_table):
        """The water-balance kernel of this day on the current stream, its routing on the routing stream.  Only the main
        pass of the routing is bandwidth-bound; everything behind it (tile totals, crossing cells, publish / peer push, the
        wait for the other ranks, the trunk kernels) is a latency-bound tail of small kernels.  The next day's water-balance
        kernel starts as soon as the main pass is done (an event the library records right behind it -- the main pass is
        also the only reader of TotR) and the tail runs underneath it, in CTAs small enough to be co-resident."""
        o = self._ovl
        main, rt = torch.cuda.current_stream(self.device), o["stream"]
        main.wait_event(o["main_pass"])                        # the main pass of the previous day's routing has finished
        prof = self._prof
        if prof is not None:
            ev = [torch.cuda.Event(enable_timing=True) for _ in range(4)]
            ev[0].record(main)
        self._ck(self.lib.sphy_wb_step(self._ctx, C.byref(self._wbP), C.byref(self._wbS), C.byref(F), C.byref(self._wbO),
                                       doy if ra_table else -1, self._stream()), "sphy_wb_step")
        done = torch.cuda.Event()
        done.record(main)
        if prof is not None:
            ev[1].record(main)
        with torch.cuda.stream(rt):
            rt.wait_event(done)
            if prof is not None:
                ev[2].record(rt)
            self.Q = self.routing.dynamic(self, None, self.TotR)
            if prof is not None:
                ev[3].record(rt)
                prof["wb"].append((ev[0], ev[1]))
                prof["route"].append((ev[2], ev[3]))

    def sync(self):
        """Make the routed maps of the last step visible to the current stream (only needed with overlap_routing)."""
        if getattr(self, "_ovl", None) is not None:
            torch.cuda.current_stream(self.device).wait_stream(self._ovl["stream"])

    def _post_step(self):
        """Time series at the stations, reporting, the date."""
        if self.collect_tss and self.Q is not None and len(self.station_ids):
            with torch.cuda.stream(self._ovl["stream"] if self._ovl is not None else torch.cuda.current_stream(self.device)):
                self._ck(self.lib.sphy_sample(self._ctx, self.Q.data_ptr(), self._station_cells.data_ptr(), len(self.station_ids),
                                              self._station_buf.data_ptr(), self._stream()), "sphy_sample")
                self.tss["QallRAtot"].append(self._station_buf[: len(self.station_ids)].clone())
        self.timestep += 1
        if self.world > 1 and (self.timestep % 32 == 0 or self.report is not None):
            self.check_exchange()          # before anything of this step leaves the device (reports, every 32 steps otherwise)
        if self.report is not None:
            self.report.step()
        self.curdate = self.curdate + datetime.timedelta(days=1)
        return self.Q if self.Q is not None else self.TotR

    # -- CUDA-graph replay of the device step ------------------------------------------------------------------
    def enable_graphs(self, on=True):
        """Replay the device part of dynamic() from a CUDA graph (one graph per set of forcing buffers, captured the second
        time the set is seen).  Launch-bound runs -- small catchments, ensembles, a basin split over many GPUs -- then cost
        one graph launch per step instead of a chain of Python -> ctypes -> kernel launches.  The only per-day kernel
        arguments, the day-of-year scalars of the Ra table, come from a device table indexed by a device-side day counter;
        the peer-exchange step counter already lives on the device.  Steps the graph cannot express (glaciers,
        vegetation, lakes/reservoirs, sediment, map series, profiling, the NCCL exchange) keep running eagerly."""
        if on and not (self.ETREF_FLAG == 1 or self._wbP.ra_mode == _lib.RA_TABLE) or self.DynVegFLAG == 1:
            return                                   # per-cell latitude maps / the vegetation kernel take the day as an argument
        if not on and self._graphs is not None:
            self._ck(self.lib.sphy_set_day_block(self._ctx, None), "sphy_set_day_block")
        self._graphs = StepGraph.state(self) if on else None

    def _graph_ok(self):
        return (self._ovl is None and self.DynVegFLAG != 1 and self.GlacFLAG != 1 and self.ResFLAG != 1 and self.LakeFLAG != 1 and self.SedFLAG != 1
                and not self._series and self._prof is None and self.ETOpenWaterFLAG != 1
                and (self.ETREF_FLAG == 1 or self._wbP.ra_mode == _lib.RA_TABLE)
                and (self.world == 1 or getattr(self, "_mg_p2p", None) is True))

    def _glacier_step(self, tavg, prec):
        """glacier.dynamic (modules/glacier.py:256-498) on the raw forcing maps of the day -> per-cell aggregates."""
        self._glT.lapse = float(self.TLapse.at[self.curdate.month])
        self._ck(self.lib.sphy_glac_step(self._ctx, C.byref(self._glT), tavg.data_ptr(), prec.data_ptr(), C.c_float(self._wbP.tcorr),
                                         C.c_float(self._wbP.pcorr), C.byref(self._glO), self._stream()), "sphy_glac_step")
        return self._glac_agg

    def _route_simple(self, TotR):
        """routing.dynamic (modules/routing.py:70-116): total flow + every flagged component share one pass."""
        src = [TotR]
        dst = [self.QRAold]
        comp_src = {"RootR": self.RootRR, "RootD": self.RootDR, "Rain": self.RainR,
                    "Snow": self.SnowR if self.SnowR is not None else self._zeros,
                    "Glac": self.GlacR if self.GlacFLAG == 1 else self._zeros,
                    "Base": self.BaseR if self.GroundFLAG == 1 else self.SubDrain}              # sphy.py:1085
        for c in COMPONENTS:
            if self.RA_FLAG[c]:
                src.append(comp_src[c])
                dst.append(getattr(self, c + "RAold"))
        nc = len(src)
        srcp = (C.c_void_p * nc)(*[t.data_ptr() for t in src])
        dstp = (C.c_void_p * nc)(*[t.data_ptr() for t in dst])
        st = self._stream()
        if self.world > 1:
            self._route_partitioned(nc, srcp, dstp, st)
        else:
            self._ck(self.lib.sphy_route_step(self._ctx, nc, srcp, dstp, self._kx, C.c_float(self._one_m_kx), self.route_method, st),
                     "sphy_route_step")
        return self.QRAold

    def _route_reslake(self, TotR):
        """advanced_routing.dynamic (modules/advanced_routing.py:134-217) incl. lakes.QLake / reservoirs.QRes."""
        if getattr(self, "_rl_part", False):
            return self._route_reslake_listed(TotR)
        self._ck(self.lib.sphy_route_reslake_step(self._ctx, C.byref(self._rl_struct), TotR.data_ptr(), self.QRAold.data_ptr(),
                                                  self._kx, C.c_float(self._one_m_kx), self._doy, self.route_method, self._stream()),
                 "sphy_route_reslake_step")
        if self.ETOpenWaterFLAG == 1:                                                          # advanced_routing.py:167-196
            ow = _lib.OpenWater(self._ow_ptr.data_ptr(), self._ow_cell.data_ptr(), self._ow_frac_dev.data_ptr(),
                                self.out["ETOpenWater"].data_ptr(), self._prec_ptr, self._pcorr)
            self._ck(self.lib.sphy_reslake_openwater(self._ctx, C.byref(self._rl_struct), C.byref(ow), self._stream()),
                     "sphy_reslake_openwater")
        return self.QRAold

    def _route_reslake_listed(self, TotR):
        """The same with the structures on the trunk list (sphy_route_reslake_begin / _finish): a basin split across GPUs, one
        exchange per step.  StorRES takes `- Qout + Qin` of day t at the head of day t + 1 (see flush_structures)."""
        import torch.distributed as dist
        st = self._stream()
        gathered = send = None
        if self.world > 1:
            if getattr(self, "_mg_p2p", None) is None:
                self._mg_p2p = self._p2p_connect()
            if not self._mg_p2p:
                if getattr(self, "_mg_send", None) is None:
                    self._mg_send = torch.zeros(self._mg_stride, dtype=torch.float64, device=self.device)
                    self._mg_recv = torch.zeros((self.world, self._mg_stride), dtype=torch.float64, device=self.device)
                send = self._mg_send.data_ptr()
        self._ck(self.lib.sphy_route_reslake_begin(self._ctx, C.byref(self._rl_struct), C.byref(self._rl_plan), TotR.data_ptr(),
                                                   self.QRAold.data_ptr(), self._kx, C.c_float(self._one_m_kx), send, self._mg_stride, st),
                 "sphy_route_reslake_begin")
        if send is not None:
            dist.all_gather_into_tensor(self._mg_recv.view(-1), self._mg_send)
            gathered = self._mg_recv.data_ptr()
        self._ck(self.lib.sphy_route_reslake_finish(self._ctx, C.byref(self._rl_struct), C.byref(self._rl_plan), self.QRAold.data_ptr(),
                                                    self._kx, C.c_float(self._one_m_kx), gathered, self._mg_stride, self.rank, self.world,
                                                    self._doy, int(self._rl_pending), st), "sphy_route_reslake_finish")
        self._rl_pending = True
        return self.QRAold

    def flush_structures(self):
        """Structures on the trunk list: bring StorRES / Qin up to date with the last routed day.  `upstream(ldd, Q)` at a
        structure (advanced_routing.py:35) needs the discharge at its donors, which may live on other ranks and only exists
        once the listed cells are finished -- so the storage update of day t normally rides in the exchange of day t + 1.
        This applies it right away: every rank contributes the donors it owns (COLLECTIVE on a partitioned basin)."""
        if not getattr(self, "_rl_part", False) or not self._rl_pending:
            return
        ns = len(self._rl_cells_host)
        x = torch.zeros(max(ns, 1) * _lib.RL_EXTRA, dtype=torch.float64, device=self.device)
        slots, local = self._rl_flush
        if slots.numel():
            x[slots] = self._rl_plan_dev["qday"][local].double()
        if self.world > 1:
            import torch.distributed as dist
            dist.all_reduce(x)                     # every slot has exactly one owner: the sum is exact
        v = x.view(-1, _lib.RL_EXTRA)
        s = v[:, 0].clone()
        for q in range(1, _lib.RL_EXTRA - 1):      # REAL8 sum in donor order, like the kernels
            s = s + v[:, q]
        r = self._rl
        qin = s[:max(ns, 1)].float()
        r["Qin"].copy_(qin)
        r["StorRES"].copy_((r["StorRES"] - r["Qout"]) + qin)                               # advanced_routing.py:36
        self._rl_pending = False

    def structure_outflow(self, kind=None):
        """Dense map of the last step's lake/reservoir outflow Qout [m3/day] (0 elsewhere); kind 1 simple reservoir,
        2 advanced reservoir, 3 lake, or a tuple of kinds."""
        out = torch.zeros(self.n, dtype=torch.float32, device=self.device)
        if len(self._rl_cells_host):
            kinds = self._rl["kind"]
            q = self._rl["Qout"]
            if kind is not None:
                ks = kind if isinstance(kind, (tuple, list)) else (kind,)
                sel = torch.zeros_like(kinds, dtype=torch.bool)
                for k in ks:
                    sel |= kinds == k
                q = torch.where(sel, q, torch.zeros_like(q))
            cells = self._rl_cells_host.astype(np.int64)
            if getattr(self, "_rl_part", False):               # global indices: keep this rank's structures
                lo, hi = getattr(self, "part_range", (0, self.n))
                own = np.flatnonzero((cells >= lo) & (cells < hi))
                cells, q = cells[own] - lo, q[torch.from_numpy(own).to(self.device)]
            if len(cells):
                out[torch.from_numpy(cells).to(self.device)] = q
        return out

    def _route_partitioned(self, nc, srcp, dstp, st):
        """Per-step confluence exchange: local scan -> all-gather of range totals / prefixes over NCCL -> trunk fix-up."""
        import torch.distributed as dist
        if getattr(self, "_mg_send", None) is None or self._mg_send.shape[0] != nc:
            self._mg_send = torch.zeros((nc, self._mg_stride), dtype=torch.float64, device=self.device)
            self._mg_recv = torch.zeros((self.world, nc, self._mg_stride), dtype=torch.float64, device=self.device)
        if getattr(self, "_mg_p2p", None) is None:
            self._mg_p2p = self._p2p_connect()
        self._ck(self.lib.sphy_route_scan_begin(self._ctx, nc, srcp, dstp, self._kx, C.c_float(self._one_m_kx),
                                                None if self._mg_p2p else self._mg_send.data_ptr(), self._mg_stride, st),
                 "sphy_route_scan_begin")
        if self._mg_p2p:                           # the publishing blocks stored into the peers' buffers over NVLink and raised
            gathered = None                        # their flags; the trunk kernel waits for the peers' flags: no collective call
        else:
            dist.all_gather_into_tensor(self._mg_recv.view(-1), self._mg_send.view(-1))
            gathered = self._mg_recv.data_ptr()
        self._ck(self.lib.sphy_route_scan_finish(self._ctx, nc, dstp, self._kx, C.c_float(self._one_m_kx),
                                                 gathered, self._mg_stride, self.rank, self.world, st),
                 "sphy_route_scan_finish")

    def _p2p_connect(self):
        """Set up the peer-memory confluence exchange (CUDA IPC handles exchanged once through torch.distributed).  Falls
        back to the NCCL all-gather when SPHY_MG_EXCHANGE=nccl, on a non-NCCL backend or if any rank cannot open its peers."""
        import torch.distributed as dist
        if os.environ.get("SPHY_MG_EXCHANGE", "p2p").lower() != "p2p" or dist.get_backend() != "nccl":
            return False
        h = (C.c_ubyte * 128)()
        rc = self.lib.sphy_p2p_setup_local(self._ctx, self.world, _lib.ROUTE_MAX_COMP * self._mg_stride,
                                           C.cast(h, C.c_void_p), C.c_void_p(C.addressof(h) + 64))
        mine = torch.tensor(list(bytes(h)) + [int(rc == 0)], dtype=torch.uint8, device=self.device)
        every = torch.empty(self.world * 129, dtype=torch.uint8, device=self.device)
        dist.all_gather_into_tensor(every, mine)
        every = every.cpu().numpy().reshape(self.world, 129)
        ok = bool(every[:, 128].all())
        if ok:
            recv_h, flag_h = every[:, :64].tobytes(), every[:, 64:128].tobytes()
            ok = self.lib.sphy_p2p_connect(self._ctx, self.rank, self.world, recv_h, flag_h) == 0
        agree = torch.tensor([int(ok)], dtype=torch.int32, device=self.device)
        dist.all_reduce(agree, op=dist.ReduceOp.MIN)
        return bool(agree.item())

    def reportable(self):
        """reporting.csv variable name -> callable giving the device map of the day (utilities/reporting.py call sites in
        sphy.py:762-1097, groundwater.py:96-125, routing.py:73-101).  Names whose expression carries a (1-GlacFrac) or
        (1-openWaterFrac) factor are only offered while that factor is exactly 1."""
        o, plain = self.out, (self.GlacFLAG != 1 and self.openWaterFrac is None)
        d = {}
        if o:
            d.update(TotPrec=lambda: o["Precip"], TotETref=lambda: o["ETref"], TotRain=lambda: o["Rain"],
                     TotETactF=lambda: o["ActETact"], Infil=lambda: o["Infil"], wbal=lambda: o["wbal"], TotSubPF=lambda: o["SubPerc"])
            if plain:
                d.update(TotPrecF=lambda: o["Precip"], TotETrefF=lambda: o["ETref"], TotRainF=lambda: o["Rain"],
                         TotETact=lambda: o["ETact"], TotRootPF=lambda: o["RootPerc"])
        if self.DynVegFLAG == 1:                                   # dynamic_veg.py:100-126; TotPrec is the RAW precipitation
            vm = self._veg_maps
            raw = lambda: self._prec_raw / f32(self._pcorr)  # noqa: E731  (sphy.py:757-763)
            d["TotPrec"] = raw
            d["LAI"] = lambda: vm["LAI"]
            if o and plain:
                d["TotPrecF"] = raw
            if self.GlacFLAG != 1:
                d.update(TotIntF=lambda: vm["Int"], TotPrecEF=lambda: vm["PrecipEff"])
            if self.openWaterFrac is None:
                d["StorCanop"] = lambda: self.Scanopy
        if self.SedFLAG == 1:                                      # mmf.py:259-306, sediment_transport.py:261-275
            so = self.sed_out
            d["SedTrans"] = lambda: so["SedTrans"]
            if self.SedModel == 2:
                d.update(DetRn=lambda: so["DetRn"], DetRun=lambda: so["DetRun"], SDepFld=lambda: so["SDepFld"])
            if self.SedTransFLAG == 1:
                d.update(TC=lambda: so["TC"], SedDep=lambda: so["SedDep"], SedYld=lambda: so["SedYld"], SedFlux=lambda: so["SedFlux"])
        d["TotRF"] = lambda: self.TotR
        if self.RootRR is not None:
            d.update(TotRootRF=lambda: self.RootRR, TotRootDF=lambda: self.RootDR, TotRainRF=lambda: self.RainR)
        if self.SnowR is not None:
            d["TotSnowRF"] = lambda: self.SnowR
        if self.GlacFLAG == 1:
            d["TotGlacRF"] = lambda: self.GlacR
        if self.GroundFLAG == 1:
            d.update(TotBaseRF=lambda: self.BaseR, TotGwRechargeF=lambda: self.GwRecharge)
            if self.openWaterFrac is None:
                d["StorGroundW"] = lambda: self.Gw
        else:
            d["TotSubDF"] = lambda: self.SubDrain
        if plain:
            d["TotCapRF"] = lambda: self.CapRise
        if self.openWaterFrac is None:
            d.update(StorRootW=lambda: self.RootWater, StorSubW=lambda: self.SubWater)
        if self.SnowFLAG == 1:
            d.update(StorSnow=lambda: self.TotalSnowStore, StorSnowW=lambda: self.SnowWatStore)
        if getattr(self, "QRAold", None) is not None:
            d["QallRAtot"] = lambda: self.QRAold
            for c in COMPONENTS:
                if self.RA_FLAG.get(c):
                    d[c + "RAtot"] = (lambda c=c: getattr(self, c + "RAold"))
        return d

    def timesteps(self):
        """utilities/timecalc.py:32-36"""
        return (self.enddate - self.startdate).days + 1

    def run(self, nsteps=None):
        """DynamicFramework(model, lastTimeStep, firstTimestep=1).run(): initial() once, then dynamic() per step."""
        self.initial()
        for _ in range(nsteps or self.timesteps()):
            self.dynamic()
        torch.cuda.synchronize(self.device)
        return self
