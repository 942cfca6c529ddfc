"""modules/lakes.py of the reference (/root/reference/modules/lakes.py:28-262) on device maps."""
import ctypes as C

import numpy as np
import torch

from .. import _lib
from . import _stream


def _struct_op(self, op, table_in=None):
    """One lake/reservoir function over the structure table (sphy_reslake_op) -> per-structure values."""
    ns = len(self._rl_cells_host)
    out = torch.full((max(ns, 1),), float("nan"), dtype=torch.float32, device=self.device)
    if ns:
        _lib.check(self.lib, self._ctx, self.lib.sphy_reslake_op(self._ctx, C.byref(self._rl_struct), op, getattr(self, "_doy", 1) or 1,
                                                                 table_in.data_ptr() if table_in is not None else None, out.data_ptr(),
                                                                 _stream()), "sphy_reslake_op")
    return out[:ns]


def _dense(self, table, fill):
    """per-structure values -> map over the active cells (`fill` elsewhere)."""
    out = torch.full((self.n,), float(fill), dtype=torch.float32, device=self.device)
    if len(self._rl_cells_host):
        out[_cells64(self)] = table
    return out


def _cells64(self):
    if "cells64" not in self._rl:
        self._rl["cells64"] = torch.from_numpy(self._rl_cells_host.astype(np.int64)).to(self.device)
    return self._rl["cells64"]


def UpdateLakeHStore(self, pcr, pcrm):
    """lakes.py:28-132 -> (LakeLevel, StorRES).  On a day with a measured-level map (`LakeFile` stack, lakes.py:31-36) the
    gauged lakes take their storage from S(h) and their level from the map; every other lake gets h(S).  The level map is
    MV (NaN) outside the lakes, like the reference's lookups on LakeID."""
    from .. import csf
    rl = self._rl_struct
    rl.lake_level = None
    if getattr(self, "_lake_update", None) is not None:
        try:
            lev = self._compact_np(self.maps.read(csf.stack_name(self.LLevel, self.counter)))
            lev_d = self._dev(np.broadcast_to(lev, (self.n,)).astype(np.float32))
            torch.index_select(lev_d, 0, _cells64(self), out=self._rl["lake_level"])
            rl.lake_level = self._rl["lake_level"].data_ptr()
        except (KeyError, FileNotFoundError):
            pass                                               # no map for this day: levels follow from the storage
    level = _struct_op(self, _lib.RL_OP_UPDATE_LAKE)
    return _dense(self, level, float("nan")), self.StorRES


def QLake(self, pcr, LakeLevel):
    """lakes.py:136-158: outflow [m3/day] of every lake from its level map; 0 outside the lakes (pcr.cover)."""
    lev = LakeLevel.to(device=self.device, dtype=torch.float32)
    table = torch.index_select(lev, 0, _cells64(self)) if len(self._rl_cells_host) else lev[:0]
    return _dense(self, _struct_op(self, _lib.RL_OP_QLAKE, table.contiguous()), 0.0)


def init(self, pcr, config):
    """lakes.py:162-201 (lake ids, S(h) / h(S) / Q(h) tables): read by SphyModel._reslake_initial into the structure table."""
    if not hasattr(self, "_rl"):
        self._reslake_initial()


def initial(self, pcr, config):
    """lakes.py:218-224: initial lake storage (performed with init: both fill the same structure table)."""
    init(self, pcr, config)
