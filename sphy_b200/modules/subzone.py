"""subzone.py of the reference (/root/reference/subzone.py:24-51) on device maps."""
from . import exp_neg_inv, op


def CapilRise(pcr, subfield, subwater, capmax, rootwater, rootsat, rootfield):
    return op("CAPILRISE", [subfield, subwater, capmax, rootwater, rootsat, rootfield])


def SubPercolation(pcr, subwater, subfield, subTT, gw, gwsat):
    return op("SUBPERCOLATION", [subwater, subfield, exp_neg_inv(subTT), gw, gwsat])


def SubDrainage(pcr, subwater, subfield, subsat, drainvel, subdrainage, subTT):
    return op("SUBDRAINAGE", [subwater, subfield, subsat, drainvel, subdrainage, exp_neg_inv(subTT)])
