"""modules/glacier.py of the reference (/root/reference/modules/glacier.py:28-810) on the device fragment table."""
import torch


def GlacCDMelt(pcr, temp, ddf, glacfrac):
    """glacier.py:28-30: melt from clean-ice or debris-covered fragments (float64 table columns, like the pandas Series)."""
    return torch.clamp(torch.as_tensor(temp, dtype=torch.float64), min=0) * ddf * glacfrac


def GMelt(glaccimelt, glacdcmelt):
    """glacier.py:34-36"""
    return glaccimelt + glacdcmelt


def GlacR(glacf, gmelt, glacfrac):
    """glacier.py:40-42"""
    return glacf * gmelt * glacfrac


def GPerc(glacf, gmelt, glacfrac):
    """glacier.py:46-48"""
    return (1 - glacf) * gmelt * glacfrac


def init(self, pcr, config, pd=None, np=None, os=None):
    """glacier.py:52-154: the fragment table sorted by MOD_ID, DDFG / DDFDG / GlacF, the lapse rates, the reporting and retreat
    options.  SphyModel keeps init and initial together (both fill the same device table)."""
    if getattr(self, "GlacTable", None) is None:
        self._glacier_initial()


def initial(self, pcr, pd=None):
    """glacier.py:158-252: per-fragment snow stores from the cell stores, GlacFrac map, initial ice maps."""
    init(self, pcr, self.config)


def dynamic(self, pcr, pd, Temp, Precip):
    """glacier.dynamic: fragment-table physics + per-cell aggregation in one kernel (sphy_glac_step); returns
    (Rain_GLAC, Snow_GLAC, ActSnowMelt_GLAC, SnowR_GLAC, GlacMelt, GlacPerc, GlacR) like glacier.py:498.
    Temp / Precip are the raw forcing maps (the kernel applies Tcorr / Pcorr like sphy.py:755-786)."""
    a = self._glacier_step(Temp, Precip)
    return a["Rain_GLAC"], a["Snow_GLAC"], a["ActSnowMelt_GLAC"], a["SnowR_GLAC"], a["GlacMelt"], a["GlacPerc"], a["GlacR"]


def dynamic_reporting(self, pcr=None, pd=None, np=None):
    """glacier.dynamic_reporting (glacier.py:501-810): per-glacier daily tables, the yearly ice redistribution and the maps
    written the day after.  Called on its own (outside SphyModel.dynamic) the cell-map updates are applied right away."""
    from .. import _lib
    for fn in self._glacier_reporting(_lib.WbForcing()):
        fn()
