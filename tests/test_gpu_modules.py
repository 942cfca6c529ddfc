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


@pytest.fixture(scope="module")
def lake_env():
    """A catchment with simple + advanced reservoirs and lakes (BASELINE config 3 physics in miniature) next to its oracle port."""
    import torch
    from oracle.sphy_port import PortInputs, PortModel
    from sphy_b200 import synth
    from sphy_b200.model import CatchmentSource, SphyModel
    cat = synth.make_catchment(nrows=200, ncols=240, seed=7, cellsize=1000.0, reservoirs=True, lakes=True, n_res_simple=6, n_res_adv=4,
                               n_lakes=5, nsteps=6, start=datetime.date(2001, 5, 1),
                               params=dict(SubWater=470.0, RootDepthFlat=150.0, Pscale=2.5))
    tmp = tempfile.mkdtemp(prefix="sphy_lake_")
    cfg = cat.write_inputs(os.path.join(tmp, "in"), os.path.join(tmp, "out"))
    model = SphyModel(cfg, CatchmentSource(cat), route_method="levels")
    model.initial()
    inp = PortInputs(cat)
    port = PortModel(inp)
    g = inp.lddst.gather
    for t in range(1, 4):                                     # a few days so that storages and discharges are not the initial ones
        port.dynamic({k: g(v) for k, v in cat.forcing(t).items()})
        model.dynamic()
    torch.cuda.synchronize()
    yield dict(model=model, port=port, inp=inp, cat=cat, torch=torch)
    model.close()


def test_lake_and_reservoir_function_signatures(lake_env):
    """lakes.UpdateLakeHStore / QLake(self, pcr, LakeLevel) and reservoirs.QSimple / QAdv / QRes(self, pcr) as callable
    functions with the reference's signatures (modules/lakes.py:28-158, modules/reservoirs.py:26-86), bit-exact against the
    port's restatement at the structure cells."""
    m, port, inp = lake_env["model"], lake_env["port"], lake_env["inp"]
    order = m.cell_order
    pos = np.empty(m.n, np.int64)
    pos[order] = np.arange(m.n)                                # canonical cell -> device index
    same(m.StorRES, port.StorRES[order])
    lake_cells, res_cells = pos[inp.lake["cells"]], pos[inp.res["cells"]]
    level, stor = m.lakes.UpdateLakeHStore(m, None, None)
    same(stor, port.StorRES[order])
    want_level = port._lakefun(inp.lake, "HS", port.StorRES[inp.lake["cells"]])
    same(level[lake_env["torch"].from_numpy(lake_cells).cuda()], want_level)
    outside = np.ones(m.n, bool)
    outside[lake_cells] = False
    assert np.isnan(level.cpu().numpy()[outside]).all()        # MV outside the lakes, like the reference's lookups on LakeID
    q = m.lakes.QLake(m, None, level).cpu().numpy()
    assert bit_equal(q[lake_cells], port.QLake()).all() and (q[outside] == 0).all()
    port.curdate = m.curdate                                   # reservoirs.QAdv reads the day of the year
    m._doy = int(m.curdate.timetuple().tm_yday)
    qres = m.reservoirs.QRes(m, None).cpu().numpy()
    assert bit_equal(qres[res_cells], port.QRes()).all()
    func = inp.res["func"]
    qs, qa = m.reservoirs.QSimple(m, None).cpu().numpy()[res_cells], m.reservoirs.QAdv(m, None).cpu().numpy()[res_cells]
    assert bit_equal(qs[func == 1], port.QRes()[func == 1]).all() and np.isnan(qs[func != 1]).all()
    assert bit_equal(qa[func == 2], port.QRes()[func == 2]).all() and np.isnan(qa[func != 2]).all()


def test_advanced_routing_rout_signature(lake_env):
    """advanced_routing.ROUT(self, pcr, rvolume, qold, qout, sres) with externally supplied maps (modules/advanced_routing.py:27-38)
    against the same arithmetic restated with the oracle's accufractionflux / upstream."""
    m, inp, torch = lake_env["model"], lake_env["inp"], lake_env["torch"]
    st = inp.lddst
    order = m.cell_order
    rng = np.random.default_rng(5)
    f32 = np.float32
    cut = m.QFRAC.cpu().numpy() == 0                           # device order
    rv = (rng.gamma(0.8, 5e4, m.n)).astype(f32)
    qold = rng.uniform(0, 30, m.n).astype(f32)
    qout = np.where(cut, rng.uniform(0, 5e5, m.n), 0).astype(f32)
    sres = np.where(cut, rng.uniform(1e6, 5e7, m.n), 0).astype(f32)
    d = lambda a: torch.from_numpy(a).cuda()  # noqa: E731
    sres_n, q, qin = m.advanced_routing.ROUT(m, None, d(rv), d(qold), d(qout), d(sres))
    canon = lambda a: a_canon(a, order)  # noqa: E731
    kx = f32(m._kx.value)
    acc = st.accuflux(canon(rv), canon(m.QFRAC.cpu().numpy()))
    qd = np.where(canon(cut), canon(qout), f32(m._one_m_kx) * acc + (canon(qold) * f32(24) * f32(3600)) * kx).astype(f32)
    want_qin = np.where(canon(cut), st.upstream(qd), f32(0)).astype(f32)
    same(q, (qd / f32(24 * 3600))[order])
    same(qin, want_qin[order])
    same(sres_n, (canon(sres) - canon(qout) + want_qin)[order])


def a_canon(a, order):
    out = np.empty_like(a)
    out[order] = a
    return out
