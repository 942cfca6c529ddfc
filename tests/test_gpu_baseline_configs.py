"""GPU: BASELINE.json configurations 1-3 (and 5 in miniature) at their NAMED sizes and lengths against the oracle port
(oracle/sphy_port.py, the restatement pinned bit-identical to the unmodified reference by tests/test_oracle_golden.py):

  C1  500 x 500, 365 daily steps, rootzone + subzone + groundwater + routing
  C2  the same catchment with the snow and glacier modules on
  C3  2000 x 2000 at 1 km with 34 reservoirs and 6 lakes, 30 steps (routing through accufractionflux)
  C5  two members of the forcing/parameter ensemble on the 2000 x 2000 basin, sharing one context

Each runs once on the CPU port and, in the same loop, on two CUDA models sharing one context: the level sweep (must be
bit-identical: at most 0.01 % of the compared values may differ in the last bit) and the scan path that bench.py times
(states bit-identical, discharge within the north star's 1e-5 relative)."""
import datetime

import numpy as np
import pytest

from tests.golden_util import bit_equal
from tests.test_gpu_golden import close, make_model, model_attr

pytestmark = pytest.mark.gpu

WET = dict(SubWater=470.0)
STATES = ["RootWater", "SubWater", "CapRise", "RootDrain", "GwRecharge", "Gw", "BaseR", "H_gw"]


def _run(kw, nsteps, every, start=datetime.date(2001, 1, 1)):
    import torch
    from oracle.sphy_port import PortInputs, PortModel
    from sphy_b200 import synth
    cat = synth.make_catchment(nsteps=nsteps, start=start, **kw)
    inp = PortInputs(cat)
    port = PortModel(inp)
    scan = make_model(cat, route_method="scan", diagnostics=False)
    lev = make_model(cat, route_method="levels", diagnostics=False, share_from=scan)
    assert scan.n == inp.n
    g = inp.lddst.gather
    keys = list(STATES)
    if cat.flags["SnowFLAG"]:
        keys += ["SnowStore", "SnowWatStore", "TotalSnowStore"]
    if cat.flags["ResFLAG"] or cat.flags["LakeFLAG"]:
        keys += ["StorRES"]
    nb = tot = 0
    worst_q = 0.0
    order = scan.cell_order
    for t in range(1, nsteps + 1):
        port.dynamic({k: g(v) for k, v in cat.forcing(t).items()})
        scan.dynamic()
        lev.dynamic()
        if t % every and t != nsteps:
            continue
        torch.cuda.synchronize()
        for k in keys + ["QRAold"]:
            want = getattr(port, k)[order]                     # the oracle is in canonical order, the models in device order
            got = model_attr(lev, k).cpu().numpy()
            ok = close(got, want)
            assert ok.all(), f"levels t={t} {k}: {int((~ok).sum())} cells beyond 1e-5"
            nb += int((~bit_equal(got, want)).sum())
            tot += got.size
            got = model_attr(scan, k).cpu().numpy()
            if k in ("QRAold", "StorRES"):                     # routed quantities: float64 range sums, 1e-5 relative
                ok = close(got, want)
                assert ok.all(), f"scan t={t} {k}: {int((~ok).sum())} cells beyond 1e-5"
                if k == "QRAold":
                    worst_q = max(worst_q, float(np.max(np.abs(got.astype(np.float64) - want) / np.maximum(np.abs(want), 1e-9))))
            else:
                assert bit_equal(got, want).all(), f"scan t={t} {k}: the water balance must not depend on the routing path"
    assert nb <= 1e-4 * tot, f"{nb} of {tot} values are not bit-identical on the level sweep"
    print(f"{kw.get('nrows')}x{kw.get('ncols')} {nsteps} steps: level sweep {nb} of {tot} values not bit-identical; "
          f"scan discharge worst {worst_q:.2e} relative")
    lev.close()
    scan.close()


def test_config1_500x500_365_steps():
    _run(dict(nrows=500, ncols=500, params=WET), 365, every=5)


def test_config2_snow_and_glaciers_365_steps():
    _run(dict(nrows=500, ncols=500, snow=True, glacier=True, params=WET), 365, every=5)


def test_config3_2000x2000_reservoirs_and_lakes_30_steps():
    _run(dict(nrows=2000, ncols=2000, cellsize=1000.0, reservoirs=True, lakes=True, n_res_simple=24, n_res_adv=10, n_lakes=6,
              params=WET), 30, every=3)


def test_config5_two_ensemble_members_2000x2000():
    """Members with perturbed kx / alphaGw / deltaGw and their own forcing share the LDD structures of one context."""
    import torch
    from oracle.sphy_port import PortInputs, PortModel
    from sphy_b200 import synth
    base = dict(nrows=2000, ncols=2000, cellsize=1000.0, seed=1, nsteps=6, start=datetime.date(2001, 6, 1), outlets="single")
    members, ports, cats = [], [], []
    for m in range(2):
        p = dict(SubWater=470.0, kx=0.971 * (1 - 0.03 * m), alphaGw=0.508 * (1 + 0.05 * m), deltaGw=300.0 * (1 - 0.04 * m))
        cat = synth.make_catchment(params=p, **base)
        cat.seed = 100 + m                                     # member-specific forcing, same terrain
        cats.append(cat)
        members.append(make_model(cat, route_method="scan", share_from=members[0] if members else None))
        ports.append(PortModel(PortInputs(cat)))
    assert members[1]._ctx is members[0]._ctx
    from sphy_b200.multigpu import Ensemble
    ens = Ensemble(members, graphs=True)          # members routed together (sphy_route_step_members), step replayed from a graph
    assert ens.grouped
    g = PortInputs(cats[0]).lddst.gather
    for t in range(1, 7):
        ens.dynamic()
        for port, cat in zip(ports, cats):
            port.dynamic({k: g(v) for k, v in cat.forcing(t).items()})
    assert len(ens.state["graphs"]) == 1          # steps 3..6 were graph replays
    torch.cuda.synchronize()
    for model, port in zip(members, ports):
        for k in STATES:
            assert bit_equal(getattr(model, k).cpu().numpy(), getattr(port, k)[model.cell_order]).all(), k
        assert close(model.QRAold.cpu().numpy(), port.QRAold[model.cell_order]).all()
    for model in reversed(members):
        model.close()
