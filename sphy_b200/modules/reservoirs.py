"""modules/reservoirs.py of the reference (/root/reference/modules/reservoirs.py:26-206)."""


def QRes(self, pcr=None):
    """Outflow of every reservoir [m3/day] (reservoirs.py:74-86): the Qout of the last step."""
    return self.structure_outflow(kind=(1, 2))


def QSimple(self, pcr=None):
    return self.structure_outflow(kind=1)


def QAdv(self, pcr=None):
    return self.structure_outflow(kind=2)
