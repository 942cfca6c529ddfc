"""modules/reservoirs.py of the reference (/root/reference/modules/reservoirs.py:26-206) on device maps."""
from .. import _lib
from .lakes import _dense, _struct_op


def QAdv(self, pcr=None):
    """reservoirs.py:26-58: outflow [m3/day] of the advanced reservoirs from the current storage and the day of the year
    (flood or dry season); MV (NaN) where the reservoir parameters are undefined, like the reference's table lookups."""
    return _dense(self, _struct_op(self, _lib.RL_OP_QADV), float("nan"))


def QSimple(self, pcr=None):
    """reservoirs.py:62-70"""
    return _dense(self, _struct_op(self, _lib.RL_OP_QSIMPLE), float("nan"))


def QRes(self, pcr=None):
    """reservoirs.py:74-86: simple or advanced by ResFunc; 0 at structures that are neither, MV (NaN) outside."""
    return _dense(self, _struct_op(self, _lib.RL_OP_QRES), float("nan"))


def init(self, pcr, config):
    """reservoirs.py:90-146 (reservoir ids, function / storage / parameter tables): read by SphyModel._reslake_initial."""
    if not hasattr(self, "_rl"):
        self._reslake_initial()


def initial(self, pcr, config):
    """reservoirs.py:150-162: initial reservoir storage (performed with init: both fill the same structure table)."""
    init(self, pcr, config)
