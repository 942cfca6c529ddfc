"""utilities/pedotransfer.py of the reference (/root/reference/utilities/pedotransfer.py, Saxton & Rawls 2006): soil
hydraulic properties from sand / clay fractions, organic matter and a bulk-density factor.  Init-time only: REAL4 NumPy
on the host in the reference's operation order, same function names and argument meaning (`pcr` is accepted and ignored)."""
import numpy as np

from ._real4 import _dbl, f32, r4, r4_div, r4_exp, r4_pow


def _log10(x): return _dbl(np.log10, x)
def _ln(x): return _dbl(np.log, x)


def Dry(pcr, self, sand, clay, OM, bulk):
    """wilting point, 1500 kPa (pedotransfer.py:21-26)"""
    first = (f32(-0.024) * sand + f32(0.487) * clay + f32(0.006) * OM + f32(0.005) * sand * OM - f32(0.013) * clay * OM
             + f32(0.068) * sand * clay + f32(0.031))
    return first + (f32(0.14) * first - f32(0.02))


def Field(pcr, self, sand, clay, OM, bulk):
    """field capacity, 33 kPa (pedotransfer.py:30-35)"""
    first = (f32(-0.251) * sand + f32(0.195) * clay + f32(0.011) * OM + f32(0.006) * sand * OM - f32(0.027) * clay * OM
             + f32(0.452) * sand * clay + f32(0.299))
    return first + (f32(1.283) * r4_pow(first, 2) - f32(0.374) * first - f32(0.015))


def Sat(pcr, self, sand, clay, OM, bulk):
    """saturated water content, 0 kPa (pedotransfer.py:39-46)"""
    first = (f32(0.278) * sand + f32(0.034) * clay + f32(0.022) * OM - f32(0.018) * sand * OM - f32(0.027) * clay * OM
             - f32(0.584) * sand * clay + f32(0.078))
    poros = first + (f32(0.636) * first - f32(0.107))
    return poros + Field(pcr, self, sand, clay, OM, bulk) - f32(0.097) * sand + f32(0.043)


def FieldAdj(pcr, self, sand, clay, OM, bulk):
    """field capacity, porosity and saturation adjusted for density (pedotransfer.py:50-61)"""
    density = (f32(1) - Sat(pcr, self, sand, clay, OM, bulk)) * f32(2.65)
    sat_adj_dens = f32(1) - (density * r4(bulk)) / f32(2.65)
    field_adj_dens = Field(pcr, self, sand, clay, OM, bulk) + f32(0.2) * (Sat(pcr, self, sand, clay, OM, bulk) - sat_adj_dens)
    return field_adj_dens, sat_adj_dens - field_adj_dens, sat_adj_dens


def KSat(pcr, self, sand, clay, OM, bulk):
    """saturated hydraulic conductivity, mm/day (pedotransfer.py:65-75)"""
    dry = Dry(pcr, self, sand, clay, OM, bulk)
    field_adj_dens, poros_adj_dens, _ = FieldAdj(pcr, self, sand, clay, OM, bulk)
    lamda = r4_div(_log10(field_adj_dens) - _log10(dry), _log10(1500) - _log10(33))
    return (f32(1930) * r4_pow(poros_adj_dens, f32(3) - lamda)) * f32(24)


def Wilt(pcr, self, np_=None):
    """wilting point from the logarithmic retention curve through field capacity and the 1500 kPa point (pedotransfer.py:78-86)"""
    B = r4_div(np.log(1500) - np.log(33), _ln(self.RootFieldMap) - _ln(self.RootDryMap))
    A = r4_exp(f32(np.log(33)) + B * _ln(self.RootFieldMap))
    return r4_pow(r4_div(100, A), r4_div(-1, B))


def init(self, pcr, config, np_=None):
    """pedotransfer.init (pedotransfer.py:89-119); the *Frac calibration factors are applied like sphy.py:173-192."""
    rd = lambda k: self._compact_np(self._readmap(config.get("PEDOTRANSFER", k)))  # noqa: E731
    sc = lambda k: f32(config.getfloat("SOIL_CAL", k))  # noqa: E731

    def bulk(k):
        try:
            return config.getfloat("PEDOTRANSFER", k)
        except ValueError:
            return rd(k)
    self.RootSandMap, self.RootClayMap = rd("RootSandMap") / f32(100), rd("RootClayMap") / f32(100)
    self.RootSiltMap = f32(1) - self.RootSandMap - self.RootClayMap
    self.RootOMMap, self.RootBulkMap = rd("RootOMMap"), bulk("RootBulkMap")
    self.SubSandMap, self.SubClayMap = rd("SubSandMap") / f32(100), rd("SubClayMap") / f32(100)
    self.SubOMMap, self.SubBulkMap = rd("SubOMMap"), bulk("SubBulkMap")
    root = (self.RootSandMap, self.RootClayMap, self.RootOMMap, self.RootBulkMap)
    sub = (self.SubSandMap, self.SubClayMap, self.SubOMMap, self.SubBulkMap)
    self.RootDryMap = Dry(None, self, *root)
    fld, _, sat = FieldAdj(None, self, *root)
    self.RootFieldMap = fld * sc("RootFieldFrac")
    self.RootSatMap = sat * sc("RootSatFrac")
    self.RootFieldMap = np.minimum(self.RootSatMap - f32(0.0001), self.RootFieldMap)
    self.RootWiltMap = Wilt(None, self) * sc("RootWiltFrac")
    self.RootDryMap = self.RootDryMap * sc("RootDryFrac")
    self.RootKsat = KSat(None, self, *root) * sc("RootKsatFrac")
    fld, _, sat = FieldAdj(None, self, *sub)
    self.SubFieldMap, self.SubSatMap = fld, sat
    self.SubKsat = KSat(None, self, *sub)
