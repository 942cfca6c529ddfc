"""CPU: CSF map / TSS writers and PCRaster stack names (sphy_b200/csf.py) round-trip and follow the documented layout."""
import os
import struct

import numpy as np

from sphy_b200 import csf


def test_stack_names_like_generateNameT():
    assert csf.stack_name("forcings/prec", 1) == "forcings/prec0000.001"
    assert csf.stack_name("prec", 365) == "prec0000.365"
    assert csf.stack_name("prec", 1000) == "prec0001.000"
    assert csf.stack_name("out/TotRM", 42) == "out/TotRM000.042"


def test_map_round_trip(tmp_path):
    rng = np.random.default_rng(0)
    a = rng.normal(size=(7, 5)).astype(np.float32)
    a[2, 3] = np.nan
    for arr in (a, rng.integers(1, 10, (7, 5)).astype(np.uint8), rng.integers(-5, 500, (7, 5)).astype(np.int32), a > 0):
        p = str(tmp_path / "m.map")
        csf.write_map(p, arr, cellsize=250.0, x_ul=10.0, y_ul=99.0)
        back, info = csf.read_map(p)
        assert np.array_equal(back, arr, equal_nan=arr.dtype.kind == "f")
        assert (info["nrows"], info["ncols"], info["cellsize"]) == (7, 5, 250.0)
        assert os.path.getsize(p) == 256 + arr.size * (1 if arr.dtype in (np.uint8, np.bool_) else 4)
        raw = open(p, "rb").read(256)
        assert raw.startswith(b"RUU CROSS SYSTEM MAP FORMAT") and struct.unpack_from("<H", raw, 32)[0] == 2


def test_tss_file(tmp_path):
    p = str(tmp_path / "q.tss")
    csf.write_tss(p, [1, 2], [(1, [0.5, 1.5]), (2, [0.25, 2.0])])
    lines = open(p).read().splitlines()
    assert lines[:5] == ["timeseries scalar", "3", "timestep", "1", "2"] and len(lines) == 7


def test_clone_mask_excludes_missing_values():
    """PCRaster treats MV in the clone as outside the catchment, whatever the clone's value scale (sphy.py:141-145)."""
    import numpy as np
    from sphy_b200.model import _clone_mask
    assert _clone_mask(np.array([[1.0, np.nan], [0.0, 2.5]], np.float32)).tolist() == [[True, False], [False, True]]
    assert _clone_mask(np.array([[1, -2147483648], [0, 7]], np.int32)).tolist() == [[True, False], [False, True]]
    assert _clone_mask(np.array([[1, 255], [0, 1]], np.uint8)).tolist() == [[True, False], [False, True]]
    assert _clone_mask(np.array([[True, False]])).tolist() == [[True, False]]


def test_run_script_usage_errors(capsys):
    """`python -m sphy_b200`: wrong invocations are refused with the usage line before anything touches a GPU."""
    from sphy_b200.__main__ import main
    assert main([]) == 2
    assert main(["--route", "sideways", "x.cfg"]) == 2
    assert main(["a.cfg", "b.cfg"]) == 2
    assert main(["/nonexistent/sphy_config.cfg"]) == 2
    err = capsys.readouterr().err
    assert "usage: python -m sphy_b200" in err and "cannot read" in err
