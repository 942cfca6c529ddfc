"""CPU: the host side of the NetCDF forcing path (sphy_b200/netcdf_forcing.py) against what the reference calls,
scipy.interpolate.griddata(method="linear") (/root/reference/utilities/netcdf2PCraster.py:192), and its time-axis lookup."""
import datetime

import numpy as np
import pytest

from sphy_b200.netcdf_forcing import linear_weights, time_index


def test_linear_weights_reproduce_griddata():
    from scipy.interpolate import griddata
    rng = np.random.default_rng(0)
    gx, gy = np.meshgrid(np.arange(0.0, 9000.0, 1000.0), np.arange(0.0, 7000.0, 1000.0))
    x, y = gx.ravel(), gy.ravel()
    z = rng.normal(10.0, 3.0, x.size)
    xi, yi = np.meshgrid(np.arange(-300.0, 8600.0, 250.0), np.arange(6500.0, -400.0, -250.0))
    idx, w = linear_weights(x, y, xi, yi)
    ok = idx[:, 0] >= 0
    got = np.full(xi.size, np.nan)
    got[ok] = (w[ok] * z[idx[ok]]).sum(axis=1)
    want = griddata((x, y), z, (xi, yi), method="linear").ravel()
    assert np.array_equal(np.isnan(got), np.isnan(want))                   # the same cells fall outside the convex hull
    assert np.allclose(got[ok], want[ok], rtol=1e-13, atol=1e-13)
    assert np.array_equal(got[ok].astype(np.float32), want[ok].astype(np.float32))    # identical once stored as REAL4


def test_time_index_exact_match():
    t = np.arange(0.0, 40.0)
    assert time_index(t, "days since 2001-01-01 00:00:00", datetime.datetime(2001, 1, 11)) == 10
    assert time_index(t * 24, "hours since 2001-01-01", datetime.datetime(2001, 2, 1)) == 31
    with pytest.raises(ValueError):
        time_index(t, "days since 2001-01-01", datetime.datetime(2002, 1, 1))
