"""modules/advanced_routing.py of the reference (/root/reference/modules/advanced_routing.py:27-217) on device maps."""


def ROUT(self, pcr, rvolume, qold, qout, sres):
    """advanced_routing.ROUT (:27-38) is fused into sphy_route_reslake_step together with lakes.QLake / reservoirs.QRes;
    it cannot be called with externally computed qout/sres maps."""
    raise NotImplementedError("fused into advanced_routing.dynamic (sphy_route_reslake_step)")


def init(self, pcr, config):
    self.kx = self._scalar_or_map("ROUTING", "kx")


def dynamic(self, pcr, pcrm, config, TotR, ETOpenWater, PrecipTot):
    """advanced_routing.dynamic (:134-217): storage update, structure outflow, accufractionflux, kx blend, inflow; -> Q [m3/s]."""
    return self._route_reslake(TotR)
