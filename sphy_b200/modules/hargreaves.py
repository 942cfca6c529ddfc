"""hargreaves.py of the reference (/root/reference/hargreaves.py:24-47) on device maps."""
import math

import numpy as np

from . import op


def julian(self):
    """utilities/timecalc.py:24-29 -> (day of year, 1)"""
    d = self.curdate
    return d.toordinal() - type(d)(d.year, 1, 1).toordinal() + 1, 1


def extrarad(self, pcr=None):
    """Extraterrestrial radiation Ra from the latitude map and the day of year (hargreaves.py:24-39)."""
    k0 = float(np.float32(((24 * 60) / math.pi) * self.Gsc))
    return op("EXTRARAD", [self.Lat], k0=k0, doy=julian(self)[0])


def Hargreaves(pcr, ra, temp, tempmax, tempmin):
    """ETref = max(0.0023*0.408*Ra*(T+17.8)*max(Tmax-Tmin,0)**0.5, 0) (hargreaves.py:43-47)."""
    return op("HARGREAVES", [ra, temp, tempmax, tempmin])
