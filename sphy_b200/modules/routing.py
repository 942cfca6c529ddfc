"""modules/routing.py of the reference (/root/reference/modules/routing.py:25-116) on device maps."""
import ctypes as C

import numpy as np
import torch

from .. import _lib
from . import RT, _stream, as_param


def ROUT(pcr, q, oldq, flowdir, kx, method=None):
    """Q = (1-kx)*accuflux(ldd, q*0.001*cellarea/86400) + kx*oldq (routing.py:25-29).  `flowdir` is accepted for
    signature compatibility: the LDD structures live in the bound context.  Returns a new map; oldq is untouched."""
    if RT.lib is None:
        raise RuntimeError("sphy_b200.modules is not bound to a model")
    keep = []
    out = oldq.clone() if isinstance(oldq, torch.Tensor) else torch.full((RT.n,), float(oldq), dtype=torch.float32, device=RT.device)
    kxp = as_param(kx, keep)
    if kxp.map:
        one_m = 0.0
    elif isinstance(kx, (np.ndarray, torch.Tensor)):
        one_m = float(np.float32(1) - np.float32(kxp.value))
    else:
        one_m = float(np.float32(1 - kx))
    src = (C.c_void_p * 1)(q.data_ptr())
    dst = (C.c_void_p * 1)(out.data_ptr())
    m = _lib.ROUTE_SCAN if method in (None, "scan") else _lib.ROUTE_LEVELS
    _lib.check(RT.lib, RT.ctx, RT.lib.sphy_route_step(RT.ctx, 1, src, dst, kxp, C.c_float(one_m), m, _stream()), "sphy_route_step")
    return out


def init(self, pcr, config):
    """routing.py:33-38: the flow-direction map and kx (SphyModel builds the LDD structures in its constructor)."""
    self.kx = self._scalar_or_map("ROUTING", "kx")


def initial(self, pcr, config):
    """routing.py:42-66: initial routed discharge and, for every `<comp>RA_init` given in the cfg, the component state."""
    from ..model import COMPONENTS
    self.QRAold = self._state(self._init_value("ROUT_INIT", "QRA_init", 0.0))
    self.RA_FLAG = {}
    for c in COMPONENTS:
        v = self._init_value("ROUT_INIT", c + "RA_init")
        self.RA_FLAG[c] = v is not None
        if v is not None:
            setattr(self, c + "RAold", self._state(v))


def dynamic(self, pcr, TotR):
    """routing.dynamic (routing.py:70-116): routes the total and every flagged component in one shared sweep; -> Q."""
    return self._route_simple(TotR)
