"""PCRaster CSF (version 2) raster files: the `.map` inputs SPHY reads with pcr.readmap (sphy.py:141-198) and the
maps it writes with self.report / pcr.report (utilities/reporting.py:24-95).

Layout restated from PCRaster's published libcsf description (unverifiable here -- PCRaster is absent):
  main header   offset 0   char[32] signature "RUU CROSS SYSTEM MAP FORMAT", UINT2 version (2), UINT4 gisFileId,
                           UINT2 projection, UINT4 attrTable, UINT2 mapType (1), UINT4 byteOrder (1)
  raster header offset 64  UINT2 valueScale, UINT2 cellRepr, 8 bytes minVal, 8 bytes maxVal, REAL8 xUL, REAL8 yUL,
                           UINT4 nrRows, UINT4 nrCols, REAL8 cellSizeX, REAL8 cellSizeY, REAL8 angle
  cells         offset 256 row-major, little endian
Missing values: REAL4 all bits set (a NaN), INT4 -2147483648, UINT1 255.
"""
from __future__ import annotations

import struct

import numpy as np

SIGNATURE = b"RUU CROSS SYSTEM MAP FORMAT"
VS_BOOLEAN, VS_NOMINAL, VS_ORDINAL, VS_SCALAR, VS_DIRECTION, VS_LDD = 0xE0, 0xE2, 0xF2, 0xEB, 0xFB, 0xF0
CR_UINT1, CR_INT4, CR_REAL4, CR_REAL8 = 0x00, 0x26, 0x5A, 0xDB
_DTYPE = {CR_UINT1: np.uint8, CR_INT4: np.int32, CR_REAL4: np.float32, CR_REAL8: np.float64}
_MV_BITS_R4 = np.uint32(0xFFFFFFFF)


def write_map(path, array, cellsize=1.0, x_ul=0.0, y_ul=0.0, value_scale=None):
    """Write a 2-D array as a CSF map.  dtype decides the representation: bool/uint8 -> UINT1 (boolean / ldd),
    integer -> INT4 nominal, float -> REAL4 scalar (NaN = MV)."""
    a = np.asarray(array)
    if a.ndim != 2:
        raise ValueError("CSF maps are 2-D")
    if a.dtype == np.bool_:
        vs, cr, data = value_scale or VS_BOOLEAN, CR_UINT1, a.astype(np.uint8)
    elif a.dtype == np.uint8:
        vs, cr, data = value_scale or VS_LDD, CR_UINT1, a
    elif np.issubdtype(a.dtype, np.integer):
        vs, cr, data = value_scale or VS_NOMINAL, CR_INT4, a.astype("<i4")
    else:
        vs, cr = value_scale or VS_SCALAR, CR_REAL4
        data = a.astype("<f4")
        bits = data.view(np.uint32).copy()
        bits[np.isnan(data)] = _MV_BITS_R4
        data = bits.view("<f4")
    valid = data[~np.isnan(data)] if cr == CR_REAL4 else data[data != (255 if cr == CR_UINT1 else -2147483648)]
    lo, hi = (valid.min(), valid.max()) if valid.size else (0, 0)
    fmt = {CR_UINT1: "<B7x", CR_INT4: "<i4x", CR_REAL4: "<f4x"}[cr]
    hdr = bytearray(256)
    hdr[0:len(SIGNATURE)] = SIGNATURE
    struct.pack_into("<HIHIHI", hdr, 32, 2, 0, 1, 0, 1, 1)
    struct.pack_into("<HH", hdr, 64, vs, cr)
    struct.pack_into(fmt, hdr, 68, lo)
    struct.pack_into(fmt, hdr, 76, hi)
    struct.pack_into("<ddIIddd", hdr, 84, float(x_ul), float(y_ul), a.shape[0], a.shape[1], float(cellsize), float(cellsize), 0.0)
    with open(path, "wb") as fh:
        fh.write(hdr)
        fh.write(np.ascontiguousarray(data).tobytes())


def read_map(path):
    """-> (array, info).  UINT1 boolean -> bool (MV -> False), ldd -> uint8 (MV 255), INT4 -> int32, REAL4 -> float32 (MV NaN)."""
    with open(path, "rb") as fh:
        hdr = fh.read(256)
        if len(hdr) < 256 or not hdr.startswith(SIGNATURE):
            raise ValueError(f"{path} is not a CSF map")
        vs, cr = struct.unpack_from("<HH", hdr, 64)
        x_ul, y_ul, nrows, ncols, csx, csy, angle = struct.unpack_from("<ddIIddd", hdr, 84)
        dt = np.dtype(_DTYPE[cr]).newbyteorder("<")
        data = np.frombuffer(fh.read(nrows * ncols * dt.itemsize), dtype=dt).reshape(nrows, ncols)
    info = dict(value_scale=vs, cell_repr=cr, cellsize=csx, x_ul=x_ul, y_ul=y_ul, nrows=nrows, ncols=ncols)
    if cr == CR_REAL4:
        out = data.astype(np.float32)
        out[data.view(np.uint32) == _MV_BITS_R4] = np.nan
        return out, info
    if cr == CR_UINT1 and vs == VS_BOOLEAN:
        return (data == 1), info
    return data.astype(_DTYPE[cr]), info


def stack_name(name, t):
    """pcraster.framework.generateNameT: 'dir/prec', 1 -> 'dir/prec0000.001' (name padded to 8 characters + 3 digits)."""
    import os
    head, tail = os.path.split(name)
    digits = str(int(t))
    pad = 11 - (len(tail) + len(digits))
    if pad < 0:
        raise ValueError("stack name too long for the 8.3 scheme")
    s = tail + "0" * pad + digits
    return os.path.join(head, s[:8] + "." + s[8:])


def write_tss(path, ids, rows, header=True):
    """PCRaster time-series file: one line per sampled step, one column per station id (TimeoutputTimeseries)."""
    with open(path, "w") as fh:
        if header:
            fh.write("timeseries scalar\n%d\ntimestep\n" % (len(ids) + 1))
            for i in ids:
                fh.write("%d\n" % int(i))
        for t, vals in rows:
            fh.write("%8d" % t + "".join(" %11.6g" % float(v) for v in vals) + "\n")
