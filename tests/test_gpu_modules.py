"""GPU: the reference's module call signatures (sphy_b200/modules/*.py -> sphy_module_op) function by function against
the oracle port's restatement of the same reference function, on random maps that hit every branch.  Bit-exact."""
import datetime
import os
import tempfile

import numpy as np
import pytest

from tests.golden_util import bit_equal

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def env():
    import torch
    from sphy_b200 import modules, synth
    from sphy_b200.model import CatchmentSource, SphyModel
    cat = synth.make_catchment(60, 50, seed=4, nsteps=5, snow=True, infil_excess=1, pws=True, start=datetime.date(2001, 6, 10),
                               zrange=(1500.0, 5200.0), params=dict(SubWater=470.0, RootDepthFlat=150.0, Pscale=2.5, Toffset=9.0))
    tmp = tempfile.mkdtemp(prefix="sphy_mod_")
    cfg = cat.write_inputs(os.path.join(tmp, "in"), os.path.join(tmp, "out"))
    model = SphyModel(cfg, CatchmentSource(cat), route_method="levels", diagnostics=True)
    model.initial()
    modules.bind(model)
    rng = np.random.default_rng(0)
    n = model.n

    def rnd(lo, hi, zero_frac=0.0):
        a = rng.uniform(lo, hi, n).astype(np.float32)
        if zero_frac:
            a[rng.random(n) < zero_frac] = 0
        return a

    def dev(a):
        return torch.from_numpy(np.ascontiguousarray(a, np.float32)).cuda()

    yield dict(model=model, cat=cat, rnd=rnd, dev=dev, torch=torch, n=n)
    model.close()


def same(t, want):
    got = t.cpu().numpy()
    ok = bit_equal(got, np.asarray(want, np.float32))
    assert ok.all(), f"{int((~ok).sum())} of {got.size} cells differ"


def test_rootzone_subzone_functions(env):
    from oracle.sphy_port import PortModel as P
    from sphy_b200.modules import rootzone, subzone
    r, d = env["rnd"], env["dev"]
    rw, rd, rf, rs = r(20, 80), r(0, 5, 0.3), r(40, 50), r(60, 70)
    dv, tt, sw, ss = r(0.01, 5), r(0.5, 8), r(100, 740), r(700, 730)
    same(rootzone.RootDrainage(None, d(rw), d(rd), d(rf), d(rs), d(dv), d(tt)), P.RootDrainage(rw, rd, rf, rs, dv, tt))
    same(rootzone.RootPercolation(None, d(rw), d(sw), d(rf), d(tt), d(ss)), P.RootPercolation(rw, sw, rf, tt, ss))
    dr, pc = r(0, 6, 0.2), r(0, 6, 0.2)
    nd, npc = rootzone.CalcFrac(None, d(rw), d(rf), d(dr), d(pc))
    wd, wp = P.CalcFrac(rw, rf, dr, pc)
    fin = np.isfinite(wd)                                # 0/0 -> MV in the reference; the GPU yields NaN there too
    same(nd[env["torch"].from_numpy(fin).cuda()], wd[fin])
    same(npc[env["torch"].from_numpy(fin).cuda()], wp[fin])
    sf_, gw = r(400, 500), r(150, 250)
    same(subzone.CapilRise(None, d(sf_), d(sw), 0.571, d(rw), d(rs), d(rf)), P.CapilRise(sf_, sw, 0.571, rw, rs, rf))
    same(subzone.SubPercolation(None, d(sw), d(sf_), d(tt), d(gw), 200.0), P.SubPercolation(sw, sf_, tt, gw, 200.0))
    sd = r(0, 3)
    same(subzone.SubDrainage(None, d(sw), d(sf_), d(ss + 50), d(dv), d(sd), d(tt)), P.SubDrainage(sw, sf_, ss + 50, dv, sd, tt))
    # scalar travel time: `-1/TT` is a Python-double division in the reference
    same(rootzone.RootDrainage(None, d(rw), d(rd), d(rf), d(rs), d(dv), 1.37), P.RootDrainage(rw, rd, rf, rs, dv, np.float32(1.37)))


def test_hargreaves_and_et(env):
    from oracle.sphy_port import PortModel as P
    from sphy_b200.modules import ET, hargreaves
    m, r, d = env["model"], env["rnd"], env["dev"]
    from oracle.sphy_port import PortInputs, PortModel
    port = PortModel(PortInputs(env["cat"]))
    port.curdate = m.curdate
    same(hargreaves.extrarad(m), port.extrarad()[m.cell_order])
    ra, t, tmax, tmin = r(20, 45), r(-10, 30), r(0, 35), r(-15, 20)
    same(hargreaves.Hargreaves(None, d(ra), d(t), d(tmax), d(tmin)), P.Hargreaves(ra, t, tmax, tmin))
    etp, rw, rs, red, rfr = r(0, 8), r(10, 70), r(60, 70), r(0, 1), r(0, 1, 0.4)
    rw[::7] = rs[::7]                                     # the ET cliff at RootWater >= RootSat
    from oracle.sphy_port import p_if, p_min
    want = p_if(rfr > 0, p_min(etp * red * p_if(rw >= rs, 0, 1), rw), 0)
    same(ET.ETact(None, d(etp), d(rw), d(rs), d(red), d(rfr)), want)
    same(ET.ETpot(d(etp), 1.2), etp * np.float32(1.2))


def test_snow_groundwater_dynamic(env):
    """snow.dynamic / groundwater.dynamic called with the reference's signature reproduce the fused kernel's state."""
    import torch
    from sphy_b200 import modules
    m = env["model"]
    f = m._ingest(1)
    snow_before = (m.SnowStore.clone(), m.SnowWatStore.clone(), m.TotalSnowStore.clone())
    gw_before = (m.GwRecharge.clone(), m.Gw.clone(), m.BaseR.clone(), m.H_gw.clone())
    m.dynamic()
    torch.cuda.synchronize()

    class Shadow:                                         # a `self` in the state before the step
        pass
    s = Shadow()
    s.SnowStore, s.SnowWatStore, s.TotalSnowStore = snow_before
    for k in ("Tcrit", "DDFS", "SnowSC", "SnowF"):
        setattr(s, k, getattr(m, k))
    ss0 = snow_before[0]
    SnowFrac = torch.where(ss0 > 0, torch.ones_like(ss0), torch.zeros_like(ss0))
    RainFrac = torch.where(ss0 == 0, torch.ones_like(ss0), torch.zeros_like(ss0))
    Temp, TempMax, Precip = f["tavg"] + np.float32(m.Tcorr), f["tmax"] + np.float32(m.Tcorr), f["prec"] / np.float32(m.Pcorr)
    Rain, SnowR, SnowSoil, old = modules.snow.dynamic(s, None, Temp, TempMax, Precip, 0.0, 0.0, SnowFrac, RainFrac, 0.0)
    assert torch.equal(s.SnowStore, m.SnowStore) and torch.equal(s.SnowWatStore, m.SnowWatStore)
    assert torch.equal(s.TotalSnowStore, m.TotalSnowStore)
    assert torch.equal(Rain, m.out["Rain"]) and torch.equal(SnowR, m.SnowR)
    g = Shadow()
    g.GwRecharge, g.Gw, g.BaseR, g.H_gw = gw_before
    for k in ("deltaGw", "alphaGw", "YieldGw", "BaseThresh"):
        setattr(g, k, getattr(m, k))
    modules.groundwater.dynamic(g, None, m.out["SubPerc"], 0.0)
    for k in ("GwRecharge", "Gw", "BaseR", "H_gw"):
        assert torch.equal(getattr(g, k), getattr(m, k)), k


def test_routing_rout_signature(env):
    import torch
    from oracle.sphy_port import PortInputs, PortModel
    from sphy_b200.modules import routing
    m = env["model"]
    port = PortModel(PortInputs(env["cat"]))
    q = env["rnd"](0, 12, 0.3)
    old = env["rnd"](0, 2)
    want = port.ROUT(q, old)[m.cell_order]
    got = routing.ROUT(None, env["dev"](q[m.cell_order]), env["dev"](old[m.cell_order]), None, 0.971, method="levels")
    same(got, want)
    got_scan = routing.ROUT(None, env["dev"](q[m.cell_order]), env["dev"](old[m.cell_order]), None, 0.971).cpu().numpy()
    assert np.allclose(got_scan, want, rtol=1e-5, atol=1e-12)
