"""modules/lakes.py of the reference (/root/reference/modules/lakes.py:28-262)."""


def UpdateLakeHStore(self, pcr, pcrm):
    """lakes.py:28-132 without measured levels: the level follows from the storage inside sphy_route_reslake_step."""
    raise NotImplementedError("fused into advanced_routing.dynamic (sphy_route_reslake_step)")


def QLake(self, pcr, LakeLevel=None):
    """Outflow of every lake [m3/day] from the current storage (lakes.py:136-158): the Qout of the last step."""
    return self.structure_outflow(kind=3)
