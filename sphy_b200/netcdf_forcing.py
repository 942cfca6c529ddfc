"""NetCDF climate forcing (reference: /root/reference/utilities/netcdf2PCraster.py, cfg flags `precNetcdfFLAG`,
`tempNetcdfFLAG`, `TminNetcdfFLAG`, `TmaxNetcdfFLAG`, sphy.py:309-363, 744-808).

The reference interpolates the coarse forcing grid onto the model raster every day with
`scipy.interpolate.griddata((x, y), z, (xi, yi), method="linear")`, i.e. a Delaunay triangulation of the forcing nodes
and a barycentric combination of three node values per model cell, on the CPU, ~N log N per day.  The triangulation and
the weights depend only on the node positions, not on the day: here they are computed ONCE on the host (same Qhull
triangulation, same `find_simplex` / barycentric transform scipy uses) for the active cells in device order, and the daily
work is one gather kernel (`sphy_interp_linear`) over a slab of a few thousand float64 node values.

What this build cannot do, and says so: the image has neither netCDF4 nor pyproj.  Files are read with
`scipy.io.netcdf_file` (NetCDF-3 classic / 64-bit offset); NetCDF-4/HDF5 files, the `rotated` pole grids and any
`InProj != OutProj` reprojection raise NotImplementedError.  Days with missing values in the slab (the reference drops those
nodes and re-triangulates, netcdf2PCraster.py:184-190) rebuild the weights for that day's node set (cached by mask).
"""
from __future__ import annotations

import ctypes as C
import datetime

import numpy as np
import torch

from . import _lib


def time_index(time_var, units, date):
    """nc.date2index(date, time, select="exact") for `<days|hours|minutes|seconds> since YYYY-MM-DD[ hh:mm:ss]` axes."""
    parts = units.strip().split()
    if len(parts) < 3 or parts[1].lower() != "since":
        raise ValueError(f"unsupported time units {units!r}")
    per_day = {"days": 1.0, "day": 1.0, "hours": 24.0, "hour": 24.0, "minutes": 1440.0, "seconds": 86400.0}[parts[0].lower()]
    origin = datetime.datetime.strptime(parts[2][:10], "%Y-%m-%d")
    if len(parts) > 3:
        h, m, s = (parts[3].split(":") + ["0", "0"])[:3]
        origin += datetime.timedelta(hours=int(h), minutes=int(m), seconds=float(s))
    want = (datetime.datetime(date.year, date.month, date.day) - origin).total_seconds() / 86400.0 * per_day
    t = np.asarray(time_var, np.float64)
    hit = np.flatnonzero(np.abs(t - want) < 1e-6 * max(per_day, 1.0))
    if hit.size == 0:
        raise ValueError(f"{date:%Y-%m-%d} is not on the time axis of the NetCDF file (select='exact')")
    return int(hit[0])


def linear_weights(x, y, xi, yi):
    """Triangle vertices and barycentric weights of every target point, exactly as scipy's LinearNDInterpolator finds them
    (Delaunay over (x, y), find_simplex, affine transform).  -> idx int32 [n, 3] (-1 outside the hull), w float64 [n, 3]."""
    from scipy.spatial import Delaunay
    pts = np.column_stack([np.asarray(x, np.float64), np.asarray(y, np.float64)])
    tri = Delaunay(pts)
    tgt = np.column_stack([np.asarray(xi, np.float64).ravel(), np.asarray(yi, np.float64).ravel()])
    simplex = tri.find_simplex(tgt)
    ok = simplex >= 0
    T = tri.transform[simplex[ok]]                              # [m, 3, 2]: inverse of the edge matrix, and the offset r
    b = np.einsum("mij,mj->mi", T[:, :2, :], tgt[ok] - T[:, 2, :])
    w = np.zeros((tgt.shape[0], 3))
    w[ok, :2] = b
    w[ok, 2] = 1.0 - b.sum(axis=1)
    idx = np.full((tgt.shape[0], 3), -1, np.int32)
    idx[ok] = tri.simplices[simplex[ok]].astype(np.int32)
    return idx, w


class NetcdfForcing:
    """One forcing variable read from a NetCDF file and interpolated onto the model's active cells on the GPU."""

    def __init__(self, model, forcing, section):
        from scipy.io import netcdf_file
        cfg = model.config
        self.model = model
        self.path = cfg.get(section, forcing + "Netcdf")
        inp = [t.strip() for t in cfg.get(section, forcing + "NetcdfInput").split(",")]      # netcdf2PCraster.py:216-227
        self.var, self.varx, self.vary, self.method = inp[0], inp[1], inp[2], inp[3]
        self.factor = float(inp[4])
        in_proj, out_proj = inp[5], inp[6]
        if in_proj == "rotated" or in_proj.lower() != out_proj.lower():
            raise NotImplementedError("NetCDF forcing: reprojection needs pyproj, which this image does not have "
                                      f"(InProj {in_proj!r}, OutProj {out_proj!r}); only InProj == OutProj is supported")
        if self.method != "linear":
            raise NotImplementedError("NetCDF forcing: only griddata(method='linear') is built (cubic = Clough-Tocher is not)")
        try:
            self.f = netcdf_file(self.path, "r", mmap=False)
        except Exception as e:                                   # noqa: BLE001
            raise NotImplementedError(f"NetCDF forcing: {self.path} is not a NetCDF-3 file scipy can read (netCDF4/HDF5 is not "
                                      f"available in this image): {e}") from None
        geo = getattr(model.maps, "georef", None)                # (xUL, yUL) of the clone map (mapattr -p in the reference)
        if geo is None:
            raise ValueError("NetCDF forcing needs the clone map's upper-left corner: give the map source a `georef = (xUL, yUL)`")
        xul, yul = float(geo[0]), float(geo[1])
        cs, rows, cols = float(model.cellsize), model.nrows, model.ncols
        X = np.asarray(self.f.variables[self.varx][:], np.float64)
        Y = np.asarray(self.f.variables[self.vary][:], np.float64)
        dy = float(Y[1] - Y[0])                                  # "cellsizeInput", netcdf2PCraster.py:120-121
        # subset of the forcing grid covering the model domain plus a buffer of two forcing cells (:124-131)
        self.x0 = int(np.argmin(np.abs(X - (xul - 2 * dy))))
        self.x1 = int(np.argmin(np.abs(X - (xul + cs * cols + 2 * dy))))
        self.y1 = int(np.argmin(np.abs(Y - (yul + 2 * dy))))
        self.y0 = int(np.argmin(np.abs(Y - (yul - cs * rows - 2 * dy))))
        if self.x1 < self.x0 or self.y1 < self.y0:
            raise ValueError("NetCDF forcing: the coordinate axes must ascend over the model domain (netcdf2PCraster.py:124-131)")
        gx, gy = np.meshgrid(X[self.x0:self.x1 + 1], Y[self.y0:self.y1 + 1])
        self.x, self.y = gx.ravel(), gy.ravel()
        # centres of the model's ACTIVE cells in device order (:144-148 builds them for the whole raster)
        r, c = np.divmod(model.flat_active, cols)
        self.xi = xul + cs * 0.5 + c * cs
        self.yi = (yul + cs * 0.5 - rows * cs) + (rows - 1 - r) * cs
        self._weights = {}
        self._out = torch.empty(model.n, dtype=torch.float32, device=model.device)
        self._time = np.asarray(self.f.variables["time"][:], np.float64)
        units = self.f.variables["time"].units
        self._units = units.decode() if isinstance(units, bytes) else str(units)

    def _tables(self, valid):
        key = valid.tobytes() if not valid.all() else b"all"
        if key not in self._weights:
            idx, w = linear_weights(self.x[valid], self.y[valid], self.xi, self.yi)
            m = self.model
            self._weights[key] = (m._dev(np.ascontiguousarray(idx.ravel())), m._dev(np.ascontiguousarray(w.ravel())))
            if len(self._weights) > 8:                           # keep the all-valid tables and the most recent masks
                for k in list(self._weights)[:-4]:
                    if k != b"all":
                        self._weights.pop(k)
        return self._weights[key]

    def day(self, date):
        """netcdf2pcrDynamic (:167-197): the day's slab, missing values dropped, interpolated -> compact REAL4 device map."""
        m = self.model
        t = time_index(self._time, self._units, date)
        z = np.asarray(self.f.variables[self.var][t, self.y0:self.y1 + 1, self.x0:self.x1 + 1], np.float64).ravel()
        with np.errstate(invalid="ignore"):
            z = np.where(z <= -9999, np.nan, z) * self.factor
        valid = ~np.isnan(z)
        idx, w = self._tables(valid)
        zd = m._dev(np.ascontiguousarray(z[valid]))
        m._keep_day = (zd,)                                      # alive until the kernel has run
        m._ck(m.lib.sphy_interp_linear(m._ctx, zd.data_ptr(), idx.data_ptr(), w.data_ptr(), m.n, self._out.data_ptr(), m._stream()),
              "sphy_interp_linear")
        return self._out
