"""modules/musle.py of the reference (MUSLE soil loss, Williams 1995; /root/reference/modules/musle.py) on device maps.

init() derives the static factors once on the host in REAL4 NumPy, in the reference's operation order.  The reference runs
musle.init inside sphy.__init__, i.e. BEFORE `pcr.setglobaloption("radians")` (sphy.py:562): PCRaster's angle unit is still
degrees there, so `atan` returns degrees and `sin`/`cos` take degrees -- restated as such.  dynamic() is one
`sphy_module_op(SPHY_OP_MUSLE)` launch."""
import os

import numpy as np

from ._real4 import _dbl, f32, r4_div, r4_exp, r4_if, r4_pow


def _atan_deg(x): return np.rad2deg(np.asarray(_dbl(np.arctan, x), np.float64)).astype(np.float32)
def _sin_deg(x): return _dbl(lambda a: np.sin(np.deg2rad(a)), x)
def _cos_deg(x): return _dbl(lambda a: np.cos(np.deg2rad(a)), x)
def _ln(x): return _dbl(np.log, x)


def LS_ULSE(self, pcr=None):
    """LS topographic factor (musle.py:49-57)"""
    m = f32(0.6) * (f32(1) - r4_exp(f32(-35.835) * self.Slope))
    alpha_hill = _atan_deg(self.Slope)
    L_hill = r4_div(f32(self.cellsize), _cos_deg(alpha_hill))
    return r4_pow(r4_div(L_hill, 22.1), m) * (f32(65.41) * r4_pow(_sin_deg(alpha_hill), 2) + f32(4.56) * _sin_deg(alpha_hill) + f32(0.065))


def K_USLE(self, pcr=None):
    """K factor from texture, organic matter and permeability class (musle.py:60-87; structure code s = 2)"""
    ksat_hourly = self.RootKsat / f32(24)
    M_textural = (self.RootSiltMap * f32(100) + f32(0)) * (f32(100) - self.RootClayMap * f32(100))
    perm = (ksat_hourly > f32(150)).astype(np.float32) * f32(1)
    for lo, hi, w in ((50, 150, 2), (15, 50, 3), (5, 15, 4), (1, 5, 5)):
        perm = perm + ((ksat_hourly > f32(lo)) & (ksat_hourly < f32(hi))).astype(np.float32) * f32(w)
    perm = perm + (ksat_hourly < f32(1)).astype(np.float32) * f32(6)
    s = 2
    return r4_div(f32(2.1 * 10 ** -4) * r4_pow(M_textural, 1.14) * (f32(12) - self.RootOMMap) + f32(3.25 * (s - 2))
                  + f32(2.5) * (perm - f32(3)), 100) * f32(0.1317)


def Tc(self, pcr=None):
    """time of concentration, hours: Kirpich channel flow + Kerby overland flow (musle.py:90-101)"""
    slope_adjusted = r4_if(self.Slope < f32(0.0025), self.Slope + f32(0.0005), self.Slope)
    alpha_hill = _atan_deg(slope_adjusted)
    L = r4_div(f32(self.cellsize), _cos_deg(alpha_hill))
    T_ch = f32(0.0195) * r4_pow(L, 0.77) * r4_pow(slope_adjusted, -0.385)
    T_ov = f32(1.44) * r4_pow(L * self.N, 0.467) * r4_pow(slope_adjusted, -0.235)
    return r4_div(T_ch + T_ov, 60)


def init(self, pcr, config):
    """musle.init (musle.py:113-139)"""
    from .mmf import _matrix_lookup
    if self.InfilFLAG != 1:
        raise ValueError("MUSLE needs Infil_excess = 1: musle.init reads self.Alpha (modules/musle.py:134), which only the "
                         "infiltration-excess branch defines (sphy.py:448-454)")
    table = os.path.join(self.inpath, config.get("MUSLE", "musle_table"))
    self.C_USLE = _matrix_lookup(table, 1, self.LandUse)
    self.N = _matrix_lookup(table, 2, self.LandUse)
    try:
        self.P_USLE = config.getfloat("MUSLE", "P_USLE")
    except ValueError:
        self.P_USLE = self._compact_np(self._readmap(config.get("MUSLE", "P_USLE")))
    if self.PedotransferFLAG == 1:
        self.K_USLE = K_USLE(self)
    else:
        self.K_USLE = self._compact_np(self._readmap(config.get("MUSLE", "K_USLE")))
    self.LS_USLE = LS_ULSE(self)
    self.CFRG = r4_exp(f32(-0.053) * (self.RockFrac * f32(100)))
    self.Tc = Tc(self)
    self.Alpha_tc = f32(1) - r4_exp(f32(2) * self.Tc * _ln(1 - (self.Alpha / 2)))
    self.ha_area = self.cellArea / f32(10000)
    self._musle_in = [self._dev(np.ascontiguousarray(np.broadcast_to(x, (self.n,)), np.float32)) if isinstance(x, np.ndarray) and x.ndim
                      else float(x) for x in (self.Alpha_tc, self.Tc, self.K_USLE, self.C_USLE, self.P_USLE, self.LS_USLE, self.CFRG)]


def dynamic(self, pcr, Runoff):
    """musle.dynamic (musle.py:143-150).  Runoff: the routed discharge Q in m3/s, or TotR in mm without routing
    (sphy.py:1217-1220); returns the sediment yield in ton per cell (reporting name SedTrans)."""
    import ctypes as C

    import torch

    from .. import _lib
    routed = float(self.RoutFLAG == 1 or self.ResFLAG == 1 or self.LakeFLAG == 1)
    inputs = [Runoff] + self._musle_in + [float(self.cellArea), float(self.ha_area), routed]
    params = (_lib.SphyParam * len(inputs))(*[_lib.SphyParam(x.data_ptr(), 0.0) if isinstance(x, torch.Tensor)
                                              else _lib.SphyParam(None, float(f32(x))) for x in inputs])
    if "SedTrans" not in self.sed_out:
        self.sed_out["SedTrans"] = torch.empty(self.n, dtype=torch.float32, device=self.device)
    outp = (C.c_void_p * 1)(self.sed_out["SedTrans"].data_ptr())
    self._ck(self.lib.sphy_module_op(self._ctx, _lib.OP["MUSLE"], params, len(inputs), outp, 1, self.n, C.c_float(0.0), 1,
                                     self._stream()), "sphy_module_op(MUSLE)")          # on the model's own context
    return self.sed_out["SedTrans"]
