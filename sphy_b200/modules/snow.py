"""modules/snow.py of the reference (/root/reference/modules/snow.py:27-187) on device maps."""
import numpy as np

from . import op


def PotSnowMelt(pcr, temp, tempmax, ddfs):
    return op("POTSNOWMELT", [temp, tempmax, ddfs])


def ActSnowMelt(pcr, snowstore, potmelt):
    return op("ACTSNOWMELT", [snowstore, potmelt])


def SnowStoreUpdate(pcr, snowstore, snow, actmelt, temp, snowwatstore):
    return op("SNOWSTOREUPDATE", [snowstore, snow, actmelt, temp, snowwatstore])


def MaxSnowWatStorage(snowsc, snowstore):
    return op("MAXSNOWWATSTORAGE", [snowsc, snowstore])


def SnowWatStorage(pcr, temp, maxsnowwatstore, snowwatstore, actmelt, rain):
    return op("SNOWWATSTORAGE", [temp, maxsnowwatstore, snowwatstore, actmelt, rain])


def TotSnowStorage(snowstore, snowwatstore, snowfrac, rainfrac):
    return op("TOTSNOWSTORAGE", [snowstore, snowwatstore, snowfrac, rainfrac])


def SnowR(pcr, snowwatstore, maxsnowwatstore, actmelt, rain, oldsnowwatstore, snowfrac):
    return op("SNOWR", [snowwatstore, maxsnowwatstore, actmelt, rain, oldsnowwatstore, snowfrac])


def init(self, pcr, config):
    """snow.py:84-90: Tcrit, SnowSC, DDFS, SnowF, SnowCth as scalar-or-map."""
    for k in ("Tcrit", "SnowSC", "DDFS", "SnowF", "SnowCth"):
        setattr(self, k, self._scalar_or_map("SNOW", k))


def initial(self, pcr, config):
    """snow.py:94-106"""
    self.SnowStore = self._state(self._init_value("SNOW_INIT", "SnowIni"))
    self.SnowWatStore = self._state(self._init_value("SNOW_INIT", "SnowWatStore"))
    self.TotalSnowStore = self.SnowStore + self.SnowWatStore


def dynamic(self, pcr, Temp, TempMax, Precip, Snow_GLAC, ActSnowMelt_GLAC, SnowFrac, RainFrac, SnowR_GLAC):
    """snow.dynamic (snow.py:110-187) as ONE pass; updates self.SnowStore / SnowWatStore / TotalSnowStore and returns
    (Rain, SnowR, SnowSoil, OldTotalSnowStore) like the reference."""
    old_total = self.TotalSnowStore
    tss_glac = getattr(self, "TotalSnowStore_GLAC", 0.0)
    ow = getattr(self, "openWaterFrac", None)
    one_m_ow = 1.0 if ow is None else (1.0 - ow)
    Rain, SnowR_, SnowSoil, ss, sws, tss = op(
        "SNOW_DYNAMIC", [Temp, TempMax, Precip, self.Tcrit, self.DDFS, self.SnowSC, self.SnowStore, self.SnowWatStore, SnowFrac,
                         RainFrac, tss_glac, SnowR_GLAC, self.SnowF, float(np.float32(1 - self.SnowF)), one_m_ow], n_out=6)
    self.SnowStore, self.SnowWatStore, self.TotalSnowStore = ss, sws, tss
    return Rain, SnowR_, SnowSoil, old_total
