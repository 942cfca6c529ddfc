"""modules/groundwater.py of the reference (/root/reference/modules/groundwater.py:26-125) on device maps."""
import numpy as np

from . import exp_neg, exp_neg_inv, op


def GroundWaterRecharge(pcr, deltagw, gwrecharge, subperc, glacperc):
    return op("GWRECHARGE", [exp_neg_inv(deltagw), gwrecharge, subperc, glacperc])


def BaseFlow(pcr, gw, baser, gwrecharge, basethresh, alphagw):
    return op("BASEFLOW", [gw, baser, gwrecharge, basethresh, exp_neg(alphagw)])


def _h_den(yield_gw, alphagw):
    if hasattr(yield_gw, "shape") or hasattr(alphagw, "shape"):
        return np.float32(800) * yield_gw * alphagw
    return float(np.float32(800 * yield_gw * alphagw))            # Python-double product, then REAL4 (groundwater.py:45)


def HLevel(pcr, Hgw, alphagw, gwrecharge, yield_gw):
    return op("HLEVEL", [Hgw, exp_neg(alphagw), gwrecharge, _h_den(yield_gw, alphagw)])


def init(self, pcr, config):
    """groundwater.py:51-57"""
    for k in ("GwDepth", "GwSat", "deltaGw", "BaseThresh", "alphaGw", "YieldGw"):
        setattr(self, k, self._scalar_or_map("GROUNDW_PARS", k))


def initial(self, pcr, config):
    """groundwater.py:61-86 (SphyModel.initial performs it together with the other states)."""
    raise NotImplementedError("use SphyModel.initial(): the groundwater states are created with the other state maps")


def dynamic(self, pcr, ActSubPerc, GlacPerc):
    """groundwater.dynamic (groundwater.py:90-125) as ONE pass; updates self.GwRecharge, Gw, BaseR, H_gw in place."""
    ow = getattr(self, "openWaterFrac", None)
    one_m_ow = 1.0 if ow is None else (1.0 - ow)
    gwr, gw, baser, hgw = op("GROUNDWATER_DYNAMIC",
                             [exp_neg_inv(self.deltaGw), self.GwRecharge, ActSubPerc, GlacPerc, self.Gw, self.BaseR, self.BaseThresh,
                              exp_neg(self.alphaGw), self.H_gw, _h_den(self.YieldGw, self.alphaGw), one_m_ow], n_out=4)
    self.GwRecharge.copy_(gwr)
    self.Gw.copy_(gw)
    self.BaseR.copy_(baser)
    self.H_gw.copy_(hgw)
