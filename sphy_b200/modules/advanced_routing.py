"""modules/advanced_routing.py of the reference (/root/reference/modules/advanced_routing.py:27-217) on device maps."""
import ctypes as C

import numpy as np
import torch

from .. import _lib
from . import _stream, op


def ROUT(self, pcr, rvolume, qold, qout, sres):
    """advanced_routing.ROUT (:27-38) with externally supplied maps: Q = accufractionflux(ldd, rvolume, QFRAC) (bit-exact
    level sweep), the kx blend with the structure outflow at QFRAC == 0, Qin = upstream(ldd, Q) at the structures, the
    storage update; -> (sres, Q [m3/s], Qin).  SphyModel.dynamic() runs the same arithmetic fused (sphy_route_reslake_step)."""
    lib, ctx = self.lib, self._ctx
    acc = torch.empty(self.n, dtype=torch.float32, device=self.device)
    rv = rvolume.contiguous()
    _lib.check(lib, ctx, lib.sphy_accuflux(ctx, rv.data_ptr(), self.QFRAC.data_ptr(), acc.data_ptr(), _lib.ROUTE_LEVELS, _stream()),
               "sphy_accuflux")
    kx = self.kx
    if isinstance(kx, np.ndarray) and kx.ndim > 0 and not np.all(kx == kx.flat[0]):
        kx_in, one_m = torch.from_numpy(np.ascontiguousarray(kx, np.float32)).to(self.device), float("nan")
    elif isinstance(kx, np.ndarray):
        kx_in = float(kx.flat[0])
        one_m = float(np.float32(1) - np.float32(kx_in))       # `1 - kx` was a REAL4 map operation
    else:
        kx_in, one_m = float(kx), float(np.float32(1 - kx))    # cfg float: Python-double subtraction, then REAL4
    q_day = op("ADV_BLEND", [acc, qold, qout, self.QFRAC, kx_in, one_m])
    up = torch.empty_like(q_day)
    _lib.check(lib, ctx, lib.sphy_upstream(ctx, q_day.data_ptr(), up.data_ptr(), _stream()), "sphy_upstream")
    sres_new, q, qin = op("ADV_FINISH", [q_day, up, self.QFRAC, sres, qout], n_out=3)
    return sres_new, q, qin


def init(self, pcr, config):
    """advanced_routing.py:42-47: flow direction (built into the context by SphyModel) and kx."""
    self.kx = self._scalar_or_map("ROUTING", "kx")


def initial(self, pcr, config):
    """advanced_routing.py:51-130: initial routed discharge; every component flag ends up False (quirk Q9)."""
    from ..model import COMPONENTS
    self.QRAold = self._state(self._init_value("ROUT_INIT", "QRA_init", 0.0))
    self.RA_FLAG = {c: False for c in COMPONENTS}


def dynamic(self, pcr, pcrm, config, TotR, ETOpenWater, PrecipTot):
    """advanced_routing.dynamic (:134-217): storage update, structure outflow, accufractionflux, kx blend, inflow; -> Q [m3/s]."""
    return self._route_reslake(TotR)
