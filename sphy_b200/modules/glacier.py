"""modules/glacier.py of the reference (/root/reference/modules/glacier.py:256-498)."""


def dynamic(self, pcr, pd, Temp, Precip):
    """glacier.dynamic: fragment-table physics + per-cell aggregation in one kernel (sphy_glac_step); returns
    (Rain_GLAC, Snow_GLAC, ActSnowMelt_GLAC, SnowR_GLAC, GlacMelt, GlacPerc, GlacR) like glacier.py:498.
    Temp / Precip are the raw forcing maps (the kernel applies Tcorr / Pcorr like sphy.py:755-786)."""
    a = self._glacier_step(Temp, Precip)
    return a["Rain_GLAC"], a["Snow_GLAC"], a["ActSnowMelt_GLAC"], a["SnowR_GLAC"], a["GlacMelt"], a["GlacPerc"], a["GlacR"]
