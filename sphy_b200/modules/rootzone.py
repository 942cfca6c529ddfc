"""rootzone.py of the reference (/root/reference/rootzone.py:24-94) on device maps."""
import numpy as np

from . import exp_neg_inv, op


def RootRunoff(self, pcr, rainfrac, rain):
    """-> (rootrunoff, Infil); branch on self.InfilFLAG like rootzone.py:26."""
    if self.InfilFLAG == 1:
        alpha_sq = float(np.float32(self.Alpha ** 2)) if not hasattr(self.Alpha, "shape") else self.Alpha * self.Alpha
        return op("ROOTRUNOFF_INFIL", [self.K_eff, self.RootKsat, self.RootSat, self.RootWater, self.Labda_Infil, self.Alpha, alpha_sq,
                                       rain, self.pavedFrac], n_out=2)
    return op("ROOTRUNOFF_SAT", [rainfrac, rain, self.RootWater, self.RootSat, self.RootKsat], n_out=2)


def RootDrainage(pcr, rootwater, rootdrain, rootfield, rootsat, drainvel, rootTT):
    return op("ROOTDRAINAGE", [rootwater, rootdrain, rootfield, rootsat, drainvel, exp_neg_inv(rootTT)])


def RootPercolation(pcr, rootwater, subwater, rootfield, rootTT, subsat):
    return op("ROOTPERCOLATION", [rootwater, subwater, rootfield, exp_neg_inv(rootTT), subsat])


def CalcFrac(pcr, rootwater, rootfield, rootdrain, rootperc):
    """-> (rootdrain, rootperc) rescaled so that their sum equals the root excess (rootzone.py:89-94)."""
    return op("CALCFRAC", [rootwater, rootfield, rootdrain, rootperc], n_out=2)
