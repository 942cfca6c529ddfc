"""ET.py of the reference (/root/reference/ET.py:24-47) on device maps."""
from . import op


def ETpot(etr, kc):
    return op("ETPOT", [etr, kc])


def ETact(pcr, etpot, rootwater, rootsat, etreddry, rainfrac):
    return op("ETACT", [etpot, rootwater, rootsat, etreddry, rainfrac])


def ks(self, pcr, etpot):
    """Plant-water-stress reduction (ET.py:39-47) from the model's RootField, RootDry, PMap and RootWater."""
    return op("KS", [self.RootField, self.RootDry, self.PMap, self.RootWater, etpot])
