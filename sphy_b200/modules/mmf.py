"""modules/mmf.py of the reference (Morgan-Morgan-Finney soil erosion, /root/reference/modules/mmf.py) on device maps.

init() derives the static parameter maps once on the host (REAL4 NumPy, table-sized work exactly like mmf.init,
mmf.py:107-190); dynamic() is one launch of the fused per-cell kernel `sphy_mmf_step` (csrc/sediment.cu), which also
evaluates the transport capacity the sediment-transport module needs."""
import ctypes as C
import os

import numpy as np

from .. import _lib
from ._real4 import f32, r4_cos, r4_div, r4_exp, r4_if, r4_pow, r4_sin

TABLE_COLS = ("PlantHeight", "NoElements", "Diameter", "CC_table", "GC_table", "NoErosion", "Tillage", "n_table", "NoVegetation")
HARVEST_COLS = ("Sowing", "Harvest", "PlantHeight_harvest", "NoElements_harvest", "Diameter_harvest", "CC_harvest", "GC_harvest",
                "Tillage_harvest")
FLOATS = ("K_c", "K_z", "K_s", "DR_c", "DR_z", "DR_s", "deltaClay", "deltaSilt", "deltaSand", "RFR", "rho_s", "rho", "eta")


def _matrix_lookup(path, col, keys):
    """pcr.lookupscalar(table, col, nominal map) in matrix-table mode (mmf.py:127-137)."""
    from ..model import _read_matrix_table, _table_col
    tab = _table_col(_read_matrix_table(path), col)
    keys = np.asarray(keys)
    out = np.full(keys.shape, np.nan, np.float32)
    for k, v in tab.items():
        out[keys == k] = f32(v)
    return out


def ManningTillage(self, pcr=None):
    """mmf.py:70-72"""
    return r4_exp(-2.1132 + 0.0349 * self.RFR)


def manningVegetation(self, pcr, waterDepth, diameter, noElements):
    """mmf.py:75-77"""
    return r4_div(f32(waterDepth ** 0.67), r4_pow(r4_div(f32(2 * 9.81), diameter * noElements), 0.5))


def FlowVelocity(self, pcr, manning, waterDepth):
    """mmf.py:80-82; `manning` is a map or a cfg float (then the left factor is a Python double)."""
    slope = r4_if(self.Slope == 0, 1e-5, self.Slope)
    if isinstance(manning, np.ndarray):
        return r4_div(f32(1), manning) * f32(waterDepth ** 0.67) * r4_pow(slope, 0.5)
    return f32(1 / manning * waterDepth ** 0.67) * r4_pow(slope, 0.5)


def init(self, pcr, config):
    """mmf.init (mmf.py:107-190)."""
    if self.PedotransferFLAG == 0:                                                            # mmf.py:109-112
        rd = lambda k: self._compact_np(self._readmap(config.get("PEDOTRANSFER", k)))  # noqa: E731
        self.RootSandMap = rd("RootSandMap") / f32(100)
        self.RootClayMap = rd("RootClayMap") / f32(100)
        self.RootSiltMap = f32(1) - self.RootSandMap - self.RootClayMap
    try:
        self.PrecInt = config.getfloat("MMF", "PrecInt")
    except ValueError:
        raise NotImplementedError("[MMF] PrecInt as a map is not supported yet") from None
    self.CanopyCoverLAIFlag = config.getfloat("MMF", "CanopyCoverLAIFlag")
    path = os.path.join(self.inpath, config.get("MMF", "MMF_table"))
    for j, k in enumerate(TABLE_COLS):
        setattr(self, k, _matrix_lookup(path, j + 1, self.LandUse))
    if self.ResFLAG == 1:                                                                     # mmf.py:133-140
        if self.ETOpenWaterFLAG == 1:
            raise NotImplementedError("soil erosion with open-water classes as reservoir extent (mmf.py:135-137) is not supported yet")
        self.Reservoirs = self._compact_np(self._readmap(config.get("RESERVOIR", "reservoirs")), np.int32)
        self.NoErosion = np.minimum(self.NoErosion + self.Reservoirs.astype(np.float32), f32(1))
    self.harvest_FLAG = config.getfloat("MMF", "harvestFLAG")
    if self.harvest_FLAG:
        path = os.path.join(self.inpath, config.get("MMF", "MMF_harvest"))
        for j, k in enumerate(HARVEST_COLS):
            setattr(self, k, _matrix_lookup(path, j + 1, self.LandUse))
    for k in FLOATS:
        setattr(self, k, config.getfloat("MMF", k))
    self.n_bare = config.getfloat("MMF", "manning")
    self.d_bare = config.getfloat("MMF", "depthBare")
    self.d_field = config.getfloat("MMF", "depthInField")
    self.d_TC = config.getfloat("MMF", "depthTC")
    self.n_tilled = ManningTillage(self)
    self.n_soil = r4_if(self.Tillage == 1, self.n_tilled, self.n_bare)
    self.n_veg_field = _veg_manning(self, self.d_field)
    self.n_field = r4_pow(r4_pow(self.n_soil, 2) + r4_pow(self.n_veg_field, 2), 0.5)
    self.v_field = FlowVelocity(self, None, self.n_field, self.d_field)
    if self.harvest_FLAG:
        nv = manningVegetation(self, None, self.d_field, self.Diameter_harvest, self.NoElements_harvest)
        self.n_veg_field_harvest = r4_if(self.Tillage_harvest == 1, 0, nv)
        self.n_field_harvest = r4_pow(r4_pow(self.n_soil, 2) + r4_pow(self.n_veg_field_harvest, 2), 0.5)
        self.v_field_harvest = FlowVelocity(self, None, self.n_field_harvest, self.d_field)


def _veg_manning(self, depth):
    """Manning's n of the vegetation at a flow depth (mmf.py:175-179, sediment_transport.py:175-178)."""
    nv = manningVegetation(self, None, depth, self.Diameter, self.NoElements)
    nv = r4_if(self.NoVegetation == 1, 0, nv)
    nv = r4_if(self.NoErosion == 1, 0, nv)
    return r4_if(self.n_table > 0, self.n_table, nv)


def prepare(self):
    """Fill the sphy_mmf argument block once all init() parts ran (called from SphyModel.initial)."""
    P = _lib.Mmf()
    pr = self._param
    P.pcorr = 1.0 if self.DynVegFLAG == 1 else float(f32(self.Pcorr))
    P.q_is_discharge = int(self.RoutFLAG == 1 or self.ResFLAG == 1 or self.LakeFLAG == 1)
    P.cc_mode = 1 if (self.CanopyCoverLAIFlag == 1 and self.DynVegFLAG == 1) else (0 if self.CanopyCoverLAIFlag == 0 else 2)
    P.harvest_on = int(bool(self.harvest_FLAG))
    P.excl_channels = int(self.exclChannelsFLAG == 1)
    P.cos_slope, P.sin_slope_pow = pr(r4_cos(self.Slope)), pr(r4_pow(r4_sin(self.Slope), 0.3))
    P.cc_table, P.gc_table, P.plant_height = pr(self.CC_table), pr(self.GC_table), pr(self.PlantHeight)
    P.no_erosion, P.rock_frac = pr(self.NoErosion), pr(self.RockFrac)
    if self.harvest_FLAG:
        P.sowing, P.harvest = pr(self.Sowing), pr(self.Harvest)
        P.cc_harvest, P.gc_harvest, P.plant_height_harvest = pr(self.CC_harvest), pr(self.GC_harvest), pr(self.PlantHeight_harvest)
        P.v_field_harvest = pr(self.v_field_harvest)
    P.clay, P.silt, P.sand = pr(self.RootClayMap), pr(self.RootSiltMap), pr(self.RootSandMap)
    P.hillslope = pr(self.Hillslope) if self.exclChannelsFLAG == 1 else _lib.SphyParam(None, 1.0)
    P.v_field = pr(self.v_field)
    for j, (k, d, delta) in enumerate((("K_c", "DR_c", self.deltaClay), ("K_z", "DR_z", self.deltaSilt), ("K_s", "DR_s", self.deltaSand))):
        P.K[j], P.DR[j] = float(f32(getattr(self, k))), float(f32(getattr(self, d)))
        P.v_s[j] = float(f32((float(1) / 18 * (delta ** 2) * (self.rho_s - self.rho) * 9.81) / self.eta))   # mmf.py:86
    if self.InfilFLAG == 1:                                                                   # mmf.py:236-237
        P.infil_alpha = float(f32(self.Alpha))
    else:
        P.ke_const = float(f32(0.29) * (f32(1) - f32(0.72) * r4_exp(-0.05 * self.PrecInt)))   # mmf.py:42
    P.d_field, P.cell_length = float(f32(self.d_field)), float(f32(self.cellsize))
    self.sed_out = {k: self._full(0.0) for k in ("DetRn", "DetRun", "SDepFld", "SedTrans")}
    for k, t in self.sed_out.items():
        setattr(P, k, t.data_ptr())
    if self.SedTransFLAG == 1:
        P.roughness = pr(self.roughnessFactor)
        if self.harvest_FLAG:
            P.roughness_harvest = pr(r4_div(self.v_TC_harvest, self.v_b))                     # sediment_transport.py:247
        P.slope_streams_pow = pr(r4_pow(self.SlopeStreams, self.TC_gamma))
        P.tc_beta = float(f32(self.TC_beta))
        for k in ("TC", "SedDep", "SedFlux"):
            self.sed_out[k] = self._full(0.0)
        self.sed_out["SedYld"] = self._full(0.0) if self.ResFLAG == 1 else self.sed_out["SedDep"]   # sediment_transport.py:43
        P.TC = self.sed_out["TC"].data_ptr()
    self._mmfP = P


def dynamic(self, pcr, Precip, Runoff):
    """mmf.dynamic (mmf.py:193-309).  Precip: the day's precipitation map as the kernels see it (device tensor, divided
    by `pcorr` inside); Runoff: the routed discharge Q in m3/s (or TotR in mm without routing, sphy.py:1217-1220).
    Returns the sediment taken into transport in ton per cell (the reference's G * cellarea / 1000)."""
    P = self._mmfP
    P.prec, P.q = Precip.data_ptr(), Runoff.data_ptr()
    P.total_snow_store = self.TotalSnowStore.data_ptr() if self.SnowFLAG == 1 else None
    P.lai = self._veg_maps["LAI"].data_ptr() if P.cc_mode == 1 else None
    yday = self.curdate.timetuple().tm_yday
    self._ck(self.lib.sphy_mmf_step(self._ctx, C.byref(P), yday, self._stream()), "sphy_mmf_step")
    return self.sed_out["SedTrans"]
