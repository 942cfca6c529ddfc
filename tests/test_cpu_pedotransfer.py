"""CPU: the host-side pedotransfer module (sphy_b200/modules/pedotransfer.py, what SphyModel runs at init) against the
oracle port's restatement, which is pinned to the reference by the golden cases c1_pedotransfer / c1_pedo_sed."""
import types

import numpy as np


def _import_without_cuda():
    # sphy_b200.modules imports torch-backed siblings; only the NumPy helpers are exercised here
    from sphy_b200.modules import pedotransfer
    return pedotransfer


def test_pedotransfer_module_matches_oracle_port():
    from oracle.sphy_port import _pedotransfer
    ped = _import_without_cuda()
    rng = np.random.default_rng(3)
    n = 5000
    m = dict(root_sand=rng.uniform(10, 80, n).astype(np.float32), root_clay=rng.uniform(5, 40, n).astype(np.float32),
             root_OM=rng.uniform(0.3, 6, n).astype(np.float32), sub_sand=rng.uniform(10, 85, n).astype(np.float32),
             sub_clay=rng.uniform(5, 40, n).astype(np.float32), sub_OM=rng.uniform(0.1, 3, n).astype(np.float32),
             sub_bulk=rng.uniform(0.9, 1.15, n).astype(np.float32))
    want = _pedotransfer(m)
    f32 = np.float32
    rs, rc = m["root_sand"] / f32(100), m["root_clay"] / f32(100)
    ss, sc = m["sub_sand"] / f32(100), m["sub_clay"] / f32(100)
    me = types.SimpleNamespace()
    me.RootDryMap = ped.Dry(None, me, rs, rc, m["root_OM"], 1.0)
    fld, _, sat = ped.FieldAdj(None, me, rs, rc, m["root_OM"], 1.0)
    me.RootFieldMap = np.minimum(sat - f32(0.0001), fld)
    got = dict(RootFieldMap=me.RootFieldMap, RootSatMap=sat, RootDryMap=me.RootDryMap, RootWiltMap=ped.Wilt(None, me),
               RootKsat=ped.KSat(None, me, rs, rc, m["root_OM"], 1.0))
    fld, _, sat = ped.FieldAdj(None, me, ss, sc, m["sub_OM"], m["sub_bulk"])
    got.update(SubFieldMap=fld, SubSatMap=sat, SubKsat=ped.KSat(None, me, ss, sc, m["sub_OM"], m["sub_bulk"]))
    for k, w in want.items():
        a, b = np.asarray(got[k], np.float32), np.asarray(w, np.float32)
        assert np.array_equal(a.view(np.int32), b.view(np.int32)) or np.array_equal(a, b, equal_nan=True), k
