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
    """groundwater.py:61-86: GwRecharge, BaseR, Gw and the groundwater table height H_gw from the cfg values or maps."""
    f32 = np.float32
    self.GwRecharge = self._state(self._init_value("GROUNDW_INIT", "GwRecharge"))
    self.BaseR = self._state(self._init_value("GROUNDW_INIT", "BaseR"))
    self.Gw = self._state(self._init_value("GROUNDW_INIT", "Gw"))
    h0 = self._init_value("GROUNDW_INIT", "H_gw")
    parts = (self.RootDepthFlat, self.SubDepthFlat, self.GwDepth, h0)
    if any(isinstance(x, np.ndarray) for x in parts):
        a = [x if isinstance(x, np.ndarray) else f32(x) for x in parts]
        hg = np.maximum((a[0] + a[1] + a[2]) / f32(1000) - a[3], f32(0)).astype(np.float32)
    else:
        hg = f32(max((parts[0] + parts[1] + parts[2]) / 1000 - parts[3], 0))
    self.H_gw = self._state(hg)
    self.SubDrain = None


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
