"""GPU: the partitioned scan-routing kernels (sphy_set_partition / sphy_route_scan_begin / _finish) with the ranks
emulated one after the other on a single GPU (no kernel ever waits on another: the all-gather is a torch.cat),
against the single-context scan path and the oracle.  The real NCCL exchange is exercised by bench.py --gpus N."""
import ctypes as C

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _ctx(lib, torch, ldd, cellsize):
    from sphy_b200 import _lib
    ctx = C.c_void_p()
    _lib.check(lib, None, lib.sphy_ctx_create(C.byref(ctx), 0, ldd.shape[0], ldd.shape[1], C.c_float(cellsize)), "create")
    ldd_d = torch.from_numpy(np.ascontiguousarray(ldd, np.uint8)).cuda()
    _lib.check(lib, ctx, lib.sphy_ldd_build(ctx, ldd_d.data_ptr(), None, None), "ldd_build")
    return ctx


def _export(lib, torch, ctx, name, n):
    from sphy_b200 import _lib
    t = torch.empty(n, dtype=torch.int32, device="cuda")
    _lib.check(lib, ctx, lib.sphy_ldd_export(ctx, name.encode(), t.data_ptr(), n * 4, None), name)
    torch.cuda.synchronize()
    return t.cpu().numpy().astype(np.int64)


def _trunk_arrays(lib, torch, ctx, m):
    from sphy_b200 import _lib
    out = {}
    for name, dt, per in (("cell", torch.int32, 1), ("sub_size", torch.int32, 1), ("is_trunk", torch.uint8, 1), ("last", torch.int32, 1),
                          ("donors", torch.int32, 8)):
        t = torch.empty(max(m * per, 1), dtype=dt, device="cuda")
        _lib.check(lib, ctx, lib.sphy_trunk_export(ctx, name.encode(), t.data_ptr(), t.numel() * t.element_size(), None), name)
        torch.cuda.synchronize()
        out[name] = t[:m * per].cpu().numpy()
    return out


def install_plan(lib, torch, ctx, bounds, rank, world, min_level, min_cells):
    """sphy_trunk_select -> [sphy_set_partition] -> plan_trunk -> sphy_trunk_set_plan, as SphyModel._install_trunk does."""
    from sphy_b200 import _lib, multigpu
    barr = (C.c_int64 * len(bounds))(*[int(b) for b in bounds])
    m = C.c_int64(0)
    _lib.check(lib, ctx, lib.sphy_trunk_select(ctx, min_level, min_cells, world, barr, C.byref(m), None), "trunk_select")
    m = int(m.value)
    arr = _trunk_arrays(lib, torch, ctx, m)
    if world > 1:
        _lib.check(lib, ctx, lib.sphy_set_partition(ctx, int(bounds[rank]), int(bounds[rank + 1]), None), "set_partition")
    plan = multigpu.plan_trunk(arr["cell"], arr["sub_size"], bounds)
    d = {k: torch.from_numpy(np.ascontiguousarray(plan[k], np.int32)).cuda() for k in ("before_rank", "before_slot", "end_rank", "end_slot")}
    pub = torch.from_numpy(np.ascontiguousarray(plan["pub"][rank], np.int32)).cuda()
    ptr = lambda t: t.data_ptr() if t.numel() else None  # noqa: E731
    _lib.check(lib, ctx, lib.sphy_trunk_set_plan(ctx, int(pub.numel()), ptr(pub), ptr(d["before_rank"]), ptr(d["before_slot"]),
                                                 ptr(d["end_rank"]), ptr(d["end_slot"]), int(plan["stride"]), rank, world, None),
               "trunk_set_plan")
    return plan, arr


@pytest.mark.parametrize("world,shape,outlets,ncomp,trunk", [(2, (300, 260), "single", 1, False), (3, (420, 380), "single", 2, True),
                                                             (4, (500, 610), "border", 2, False), (8, (640, 700), "single", 1, True)])
def test_partitioned_scan_matches_single_context(world, shape, outlets, ncomp, trunk):
    import torch
    from oracle.ldd_oracle import LddStruct
    from sphy_b200 import _lib, multigpu, synth
    lib = _lib.load()
    cellsize = 250.0
    _, _, ldd, _ = synth.make_dem_ldd(shape[0], shape[1], cellsize, seed=world, outlets=outlets)
    st = LddStruct(ldd)
    n = st.n
    min_level, min_cells = (1024, 1100) if trunk else (2 ** 31 - 1, 2 ** 40)      # with / without rounding-drift cells
    ref = _ctx(lib, torch, ldd, cellsize)
    canon = _export(lib, torch, ref, "cell_order", n)
    _, ref_arr = install_plan(lib, torch, ref, np.array([0, n]), 0, 1, min_level, min_cells)
    assert (len(ref_arr["cell"]) > 0) == trunk
    bounds = multigpu.partition_bounds(n, world)
    parts = []
    stride = None
    for p in range(world):
        ctx = _ctx(lib, torch, ldd, cellsize)
        plan, arr = install_plan(lib, torch, ctx, bounds, p, world, min_level, min_cells)
        stride = int(plan["stride"])
        assert lib.sphy_ncells(ctx) == bounds[p + 1] - bounds[p]
        assert set(ref_arr["cell"][ref_arr["is_trunk"] != 0]) == set(arr["cell"][arr["is_trunk"] != 0])   # the same trunk at any split
        parts.append((ctx, None))
    kx = _lib.SphyParam(None, 0.971)
    one_m = float(np.float32(1 - 0.971))
    rng = np.random.default_rng(world)
    q_ref = [torch.zeros(n, device="cuda") for _ in range(ncomp)]
    q_par = [[torch.zeros(int(bounds[p + 1] - bounds[p]), device="cuda") for _ in range(ncomp)] for p in range(world)]
    q64 = [np.zeros(n) for _ in range(ncomp)]
    area = np.float32(cellsize) * np.float32(cellsize)
    for step in range(3):
        mm = [(rng.gamma(0.8, 4.0, n) * (rng.random(n) < 0.7)).astype(np.float32) for _ in range(ncomp)]     # canonical order
        mm_dev = [torch.from_numpy(m[canon]).cuda() for m in mm]                                            # device order
        src = (C.c_void_p * ncomp)(*[t.data_ptr() for t in mm_dev])
        dst = (C.c_void_p * ncomp)(*[t.data_ptr() for t in q_ref])
        _lib.check(lib, ref, lib.sphy_route_step(ref, ncomp, src, dst, kx, C.c_float(one_m), _lib.ROUTE_SCAN, None), "route")
        sends = []
        locals_ = []
        for p, (ctx, _) in enumerate(parts):
            lo, hi = int(bounds[p]), int(bounds[p + 1])
            loc = [t[lo:hi].clone() for t in mm_dev]                     # this rank's cells (16-byte aligned copies)
            locals_.append(loc)
            send = torch.zeros((ncomp, stride), dtype=torch.float64, device="cuda")
            srcp = (C.c_void_p * ncomp)(*[t.data_ptr() for t in loc])
            dstp = (C.c_void_p * ncomp)(*[t.data_ptr() for t in q_par[p]])
            _lib.check(lib, ctx, lib.sphy_route_scan_begin(ctx, ncomp, srcp, dstp, kx, C.c_float(one_m), send.data_ptr(), stride, None),
                       "scan_begin")
            sends.append(send)
        gathered = torch.stack(sends).contiguous()                       # == all_gather_into_tensor
        for p, (ctx, _) in enumerate(parts):
            dstp = (C.c_void_p * ncomp)(*[t.data_ptr() for t in q_par[p]])
            _lib.check(lib, ctx, lib.sphy_route_scan_finish(ctx, ncomp, dstp, kx, C.c_float(one_m), gathered.data_ptr(), stride, p, world,
                                                            None), "scan_finish")
        torch.cuda.synchronize()
        for c in range(ncomp):
            got = torch.cat([q_par[p][c] for p in range(world)]).cpu().numpy().astype(np.float64)
            single = q_ref[c].cpu().numpy().astype(np.float64)
            rr = (mm[c] * np.float32(0.001) * area) / np.float32(86400)
            q64[c] = float(np.float32(one_m)) * st.accuflux_f64(rr) + float(np.float32(0.971)) * q64[c]
            want = q64[c][canon]
            assert np.all(np.abs(got - want) <= 2e-6 * np.abs(want) + 1e-12), f"step {step} comp {c}: partitioned vs float64 oracle"
            # the split only shows through the float64 prefixes: a rare REAL4 ulp, never more (the drift sums are exact)
            assert np.all(np.abs(got - single) <= 2.5e-7 * np.abs(single) + 1e-12), f"step {step} comp {c}: partitioned vs single context"
    # a partitioned context refuses the whole-basin entry points
    assert lib.sphy_upstream(parts[0][0], mm_dev[0].data_ptr(), q_ref[0].data_ptr(), None) == -4
    for ctx, _ in parts:
        lib.sphy_ctx_destroy(ctx)
    lib.sphy_ctx_destroy(ref)


def _install_reslake_plan(lib, torch, ctx, bounds, rank, world, cut, qfrac):
    """sphy_trunk_select_cuts -> export / sphy_cell_donors -> [sphy_set_partition] -> plan_trunk + plan_reslake ->
    sphy_trunk_set_plan, as SphyModel._install_trunk / _reslake_plan_initial do; returns the plan struct and its tensors."""
    from sphy_b200 import _lib, multigpu
    barr = (C.c_int64 * len(bounds))(*[int(b) for b in bounds])
    m = C.c_int64(0)
    cut_d = torch.from_numpy(cut.astype(np.int32)).cuda()
    _lib.check(lib, ctx, lib.sphy_trunk_select_cuts(ctx, len(cut), cut_d.data_ptr(), 2 ** 31 - 1, 2 ** 40, world, barr, C.byref(m), None),
               "trunk_select_cuts")
    m = int(m.value)
    arr = _trunk_arrays(lib, torch, ctx, m)
    don = torch.empty(len(cut) * 8, dtype=torch.int32, device="cuda")
    _lib.check(lib, ctx, lib.sphy_cell_donors(ctx, len(cut), cut_d.data_ptr(), don.data_ptr(), None), "cell_donors")
    donors = don.cpu().numpy().reshape(-1, 8)
    lo, hi = int(bounds[rank]), int(bounds[rank + 1])
    if world > 1:
        _lib.check(lib, ctx, lib.sphy_set_partition(ctx, lo, hi, None), "set_partition")
    plan = multigpu.plan_trunk(arr["cell"], arr["sub_size"], bounds)
    rl = multigpu.plan_reslake(arr["cell"], arr["sub_size"], cut, donors, bounds)
    stride = int(plan["stride"]) + _lib.RL_EXTRA * len(cut)
    d = {k: torch.from_numpy(np.ascontiguousarray(plan[k], np.int32)).cuda() for k in ("before_rank", "before_slot", "end_rank", "end_slot")}
    pub = torch.from_numpy(np.ascontiguousarray(plan["pub"][rank], np.int32)).cuda()
    ptr = lambda t: t.data_ptr() if t.numel() else None  # noqa: E731
    _lib.check(lib, ctx, lib.sphy_trunk_set_plan(ctx, int(pub.numel()), ptr(pub), ptr(d["before_rank"]), ptr(d["before_slot"]),
                                                 ptr(d["end_rank"]), ptr(d["end_slot"]), stride, rank, world, None), "trunk_set_plan")
    listed = arr["cell"].astype(np.int64)
    qmask = qfrac[lo:hi].copy()
    qmask[listed[(listed >= lo) & (listed < hi)] - lo] = 2.0
    t = {k: torch.from_numpy(np.ascontiguousarray(rl[k] if len(rl[k]) else np.zeros(1), np.int32)).cuda()
         for k in ("cut_slot", "entry_cut", "entry_ptr", "entry_idx", "x_rank", "x_local")}
    t["qmask"] = torch.from_numpy(qmask).cuda()
    t["qday"] = torch.zeros(hi - lo, device="cuda")
    t["extra"] = torch.zeros(len(cut) * _lib.RL_EXTRA, dtype=torch.float64, device="cuda")
    t["top"] = torch.zeros(max(m, 1), dtype=torch.float64, device="cuda")
    struct = _lib.ResLakePlan(*[t[k].data_ptr() if k in t else None for k in ("cut_slot", "entry_cut", "entry_ptr", "entry_idx", "x_rank",
                                                                              "x_local", "qmask", "material", "qday", "extra", "top")])
    x_cell = np.where(rl["x_rank"] >= 0, np.asarray(bounds)[np.maximum(rl["x_rank"], 0)] + rl["x_local"], -1)
    return struct, t, stride, listed, x_cell


def _structure_table(torch, rng, cut, S0):
    """Simple reservoirs, advanced reservoirs and lakes in turn, with parameters that keep every branch active."""
    from sphy_b200 import _lib
    ns = len(cut)
    par = np.full((_lib.RL_NPAR, ns), np.nan, np.float32)
    kind = (np.arange(ns) % 3 + 1).astype(np.int32)
    for j in range(ns):
        s0 = float(S0[j])
        if kind[j] == 1:
            par[0, j], par[1, j], par[2, j] = 0.04 + 0.02 * rng.random(), 1.5, 2.0 * s0
        elif kind[j] == 2:
            par[3, j], par[4, j], par[5, j], par[6, j], par[7, j], par[8, j] = 3.0 * s0, 0.2 * s0, 0.1 * s0, 0.02 * s0, 100, 250
        else:
            par[9, j], par[13, j], par[12, j] = 2, 1e-6, 0.5           # h = a1 * S + b
            par[16, j], par[20, j], par[19, j] = 2, 0.4, 0.0           # Q = a1 * h + b  [m3/s]
    return kind, par


def _rl_struct(torch, cut, kind, par, qfrac_local, S0):
    from sphy_b200 import _lib
    ns = len(cut)
    t = {"cell": torch.from_numpy(cut.astype(np.int32)).cuda(), "kind": torch.from_numpy(kind).cuda(), "par": torch.from_numpy(par).cuda(),
         "qfrac": torch.from_numpy(qfrac_local).cuda(), "StorRES": torch.from_numpy(S0.copy()).cuda(),
         "Qout": torch.zeros(ns, device="cuda"), "Qin": torch.zeros(ns, device="cuda")}
    st = _lib.ResLake(ns, t["cell"].data_ptr(), t["kind"].data_ptr(), t["par"].data_ptr(), t["qfrac"].data_ptr(), t["StorRES"].data_ptr(),
                      t["Qout"].data_ptr(), t["Qin"].data_ptr(), None, None)
    return st, t


@pytest.mark.parametrize("world,shape", [(1, (330, 300)), (2, (420, 380)), (4, (500, 610))])
def test_partitioned_lakes_and_reservoirs_match_the_level_sweep(world, shape):
    """sphy_route_reslake_begin / _finish (structures on the trunk list, the ranks emulated one after the other, the
    all-gather a torch.stack) against the bit-exact level sweep of sphy_route_reslake_step on the whole basin: discharge
    within the north star's 1e-5, and the replicated structure tables (storage one update behind until flushed)."""
    import torch
    from oracle.ldd_oracle import LddStruct
    from sphy_b200 import _lib, multigpu, synth
    lib = _lib.load()
    cellsize = 250.0
    _, _, ldd, _ = synth.make_dem_ldd(shape[0], shape[1], cellsize, seed=20 + world, outlets="single")
    n = LddStruct(ldd).n
    ref = _ctx(lib, torch, ldd, cellsize)
    sub_dev = _export(lib, torch, ref, "sub_size", n)          # upstream cell counts, by device (DFS) index
    rng = np.random.default_rng(world)
    # structures on streams of very different size; two of them directly behind each other on the main stem
    big = np.flatnonzero(sub_dev > 2000)
    small = np.flatnonzero((sub_dev > 30) & (sub_dev < 600))
    stem = np.flatnonzero(sub_dev > n // 3)
    cut = np.unique(np.concatenate([rng.choice(big, 10, replace=False), rng.choice(small, 6, replace=False),
                                    stem[[len(stem) // 2, len(stem) // 2 + 1]]]))
    cut = cut[cut > 0]
    ns = len(cut)
    qfrac = np.ones(n, np.float32)
    qfrac[cut] = 0.0
    S0 = rng.uniform(2e6, 4e7, ns).astype(np.float32)
    kind, par = _structure_table(torch, rng, cut, S0)
    rl_ref, t_ref = _rl_struct(torch, cut, kind, par, qfrac, S0)
    bounds = multigpu.partition_bounds(n, world) if world > 1 else np.array([0, n], np.int64)
    ranks = []
    for p in range(world):
        ctx = _ctx(lib, torch, ldd, cellsize)
        plan, pt, stride, listed, x_cell = _install_reslake_plan(lib, torch, ctx, bounds, p, world, cut, qfrac)
        lo, hi = int(bounds[p]), int(bounds[p + 1])
        rl, rt = _rl_struct(torch, cut, kind, par, qfrac[lo:hi], S0)
        ranks.append(dict(ctx=ctx, plan=plan, pt=pt, rl=rl, rt=rt, q=torch.zeros(hi - lo, device="cuda"), lo=lo, hi=hi))
    assert set(cut) <= set(listed)
    kx = _lib.SphyParam(None, 0.93)
    one_m = float(np.float32(1 - 0.93))
    q_ref = torch.zeros(n, device="cuda")
    for step in range(5):
        doy = 150 + step
        totr = torch.from_numpy((rng.gamma(0.8, 4.0, n) * (rng.random(n) < 0.6)).astype(np.float32)).cuda()      # device order
        _lib.check(lib, ref, lib.sphy_route_reslake_step(ref, C.byref(rl_ref), totr.data_ptr(), q_ref.data_ptr(), kx, C.c_float(one_m), doy,
                                                         _lib.ROUTE_LEVELS, None), "reslake_step")
        sends, locs = [], []
        for r in ranks:
            loc = totr[r["lo"]:r["hi"]].clone()
            locs.append(loc)
            send = torch.zeros(stride, dtype=torch.float64, device="cuda") if world > 1 else None
            _lib.check(lib, r["ctx"], lib.sphy_route_reslake_begin(r["ctx"], C.byref(r["rl"]), C.byref(r["plan"]), loc.data_ptr(),
                                                                   r["q"].data_ptr(), kx, C.c_float(one_m),
                                                                   send.data_ptr() if send is not None else None, stride, None), "reslake_begin")
            sends.append(send)
        gathered = torch.stack(sends).contiguous() if world > 1 else None                    # == all_gather_into_tensor
        for p, r in enumerate(ranks):
            _lib.check(lib, r["ctx"], lib.sphy_route_reslake_finish(r["ctx"], C.byref(r["rl"]), C.byref(r["plan"]), r["q"].data_ptr(), kx,
                                                                    C.c_float(one_m), gathered.data_ptr() if gathered is not None else None,
                                                                    stride, p, world, doy, int(step > 0), None), "reslake_finish")
        torch.cuda.synchronize()
        got = torch.cat([r["q"] for r in ranks]).cpu().numpy().astype(np.float64)
        want = q_ref.cpu().numpy().astype(np.float64)
        err = np.abs(got - want) / np.maximum(np.abs(want), 1e-9)
        bad = np.flatnonzero(np.abs(got - want) > 1e-5 * np.abs(want) + 1e-9)
        assert bad.size == 0, (f"step {step}: {bad.size} cells off, worst {err.max():.3e} at {int(err.argmax())}; listed among the bad: "
                               f"{int(np.isin(bad, listed).sum())}; structure cells among the bad: {int(np.isin(bad, cut).sum())}; "
                               f"first {bad[:8]} got {got[bad[:4]]} want {want[bad[:4]]}")
        # the structure tables are the same on every rank, and one storage update behind the reference until flushed
        for r in ranks[1:]:
            for k in ("StorRES", "Qout", "Qin"):
                assert torch.equal(r["rt"][k], ranks[0]["rt"][k]), (step, k)
        qday = torch.cat([r["pt"]["qday"] for r in ranks]).cpu().numpy().astype(np.float64)
        v = np.where(x_cell >= 0, qday[np.maximum(x_cell, 0)], 0.0).reshape(ns, _lib.RL_EXTRA)
        qin = np.zeros(ns)
        for q in range(_lib.RL_EXTRA - 1):
            qin = qin + v[:, q]
        qin = qin.astype(np.float32)
        t0 = ranks[0]["rt"]
        flushed = (t0["StorRES"].cpu().numpy() - t0["Qout"].cpu().numpy()) + qin
        assert np.allclose(t0["Qout"].cpu().numpy(), t_ref["Qout"].cpu().numpy(), rtol=2e-5, atol=1e-3), step
        assert np.allclose(qin, t_ref["Qin"].cpu().numpy(), rtol=2e-5, atol=1e-3), step
        assert np.allclose(flushed, t_ref["StorRES"].cpu().numpy(), rtol=2e-5), step
    for r in ranks:
        lib.sphy_ctx_destroy(r["ctx"])
    lib.sphy_ctx_destroy(ref)


@pytest.mark.parametrize("exchange,glacier,overlap,reslake", [("p2p", False, False, False), ("nccl", False, False, False),
                                                              ("p2p", True, False, False), ("p2p", False, True, False),
                                                              ("nccl", False, True, False), ("p2p", False, False, True),
                                                              ("nccl", False, False, True)])
def test_real_two_gpu_exchange(exchange, glacier, overlap, reslake):
    """Two real ranks (one per GPU) under torch.distributed.run: both confluence exchanges -- stores into peer memory over
    NVLink with flags, and the NCCL all-gather -- against the unpartitioned basin and the level sweep
    (tools/check_multigpu.py).  Skipped where only one GPU is visible."""
    import json
    import os
    import subprocess
    import sys

    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    port = 29600 + (os.getpid() % 300) + (0 if exchange == "p2p" else 301) + (602 if glacier else 0) + (1203 if overlap else 0) + (2406 if reslake else 0)
    env = dict(os.environ, SPHY_MG_EXCHANGE=exchange)
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
                        "--master-port", str(port), os.path.join(root, "tools", "check_multigpu.py"), "--rows", "1500", "--cols", "1300",
                        "--steps", "4"] + (["--glacier"] if glacier else []) + (["--overlap"] if overlap else []) + (["--reslake"] if reslake else []),
                       env=env, capture_output=True, text=True, timeout=900, cwd=root)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-3000:]
    line = json.loads([ln for ln in r.stdout.splitlines() if ln.startswith('{"tool"')][-1])
    assert line["exchange"] == exchange and line["world"] == 2 and line["overlap_routing"] == overlap
    assert line["worst_rel_partitioned_vs_one_gpu"] <= (3e-6 if reslake else 3e-7) and line["worst_rel_partitioned_vs_level_sweep"] <= 1e-5


def test_model_with_structures_on_the_trunk_list_matches_the_oracle_port(monkeypatch):
    """SphyModel with lakes and reservoirs routed through sphy_route_reslake_begin / _finish (SPHY_RESLAKE_PLAN=1: the path a
    partitioned basin runs, here on one GPU) against the oracle port over 12 days: discharge and lake / reservoir storage
    within 1e-5, the water-balance states bit-identical."""
    import datetime
    import os
    import tempfile

    import torch
    from oracle.sphy_port import PortInputs, PortModel
    from sphy_b200 import synth
    from sphy_b200.model import CatchmentSource, SphyModel
    monkeypatch.setenv("SPHY_RESLAKE_PLAN", "1")
    cat = synth.make_catchment(nrows=260, ncols=300, seed=23, cellsize=1000.0, reservoirs=True, lakes=True, n_res_simple=7, n_res_adv=5,
                               n_lakes=5, nsteps=12, start=datetime.date(2001, 5, 1), params=dict(SubWater=470.0, RootDepthFlat=150.0, Pscale=2.5))
    tmp = tempfile.mkdtemp(prefix="sphy_rlplan_")
    cfg = cat.write_inputs(os.path.join(tmp, "in"), os.path.join(tmp, "out"))
    model = SphyModel(cfg, CatchmentSource(cat), route_method="scan")
    model.initial()
    assert model._rl_part and model.trunk_cells >= len(model._rl_cells_host) > 0
    inp = PortInputs(cat)
    port = PortModel(inp)
    g = inp.lddst.gather
    order = model.cell_order
    for t in range(1, 13):
        port.dynamic({k: g(v) for k, v in cat.forcing(t).items()})
        model.dynamic()
        got, want = model.QRAold.cpu().numpy().astype(np.float64), port.QRAold[order].astype(np.float64)
        bad = np.flatnonzero(np.abs(got - want) > 1e-5 * np.abs(want) + 1e-12)
        assert bad.size == 0, (t, bad[:8], got[bad[:4]], want[bad[:4]])
        if t % 3 == 0:                                         # StorRES applies the pending storage update first
            s_got, s_want = model.StorRES.cpu().numpy().astype(np.float64), port.StorRES[order].astype(np.float64)
            assert np.all(np.abs(s_got - s_want) <= 1e-5 * np.abs(s_want) + 1e-3), t
    for k in ("RootWater", "SubWater", "Gw"):
        assert np.array_equal(getattr(model, k).cpu().numpy(), getattr(port, k)[order]), k
    torch.cuda.synchronize()
    model.close()


def test_run_script_on_two_gpus_writes_the_same_outputs():
    """`python -m torch.distributed.run --nproc-per-node 2 -m sphy_b200 sphy_config.cfg`: the basin split across two GPUs, rank 0
    assembling the reported rasters and merging the .tss columns -- against the same run on one GPU, file by file."""
    import datetime
    import os
    import subprocess
    import sys
    import tempfile

    import torch
    from sphy_b200 import csf, synth
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    outs = []
    for world in (1, 2):
        cat = synth.make_catchment(70, 64, seed=29, nsteps=6, start=datetime.date(2001, 6, 1), report_daily=("TotRF", "StorRootW"),
                                   report_opts={"QallRAtot": ("D", "NONE", "D"), "TotRF": ("D", "NONE", "D")},
                                   params=dict(SubWater=470.0, RootDepthFlat=150.0, Pscale=2.5))
        tmp = tempfile.mkdtemp(prefix=f"sphy_cli{world}_")
        cfg = cat.write_inputs(os.path.join(tmp, "in"), os.path.join(tmp, "out"))
        cat.write_csf_inputs(os.path.join(tmp, "in"), 6)
        launch = ([sys.executable, "-m", "sphy_b200", cfg] if world == 1 else
                  [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
                   "--master-port", str(29850 + os.getpid() % 100), "-m", "sphy_b200", cfg])
        r = subprocess.run(launch, cwd=root, capture_output=True, text=True, timeout=900)
        assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-3000:]
        assert r.stdout.count("Simulation succesfully completed") == 1
        outs.append(os.path.join(tmp, "out"))
    one, two = (sorted(os.listdir(o)) for o in outs)
    assert one == two and len(one) >= 14, (one, two)
    for f in one:
        if f.endswith(".tss"):
            a = np.loadtxt(os.path.join(outs[0], f), skiprows=3 + int(open(os.path.join(outs[0], f)).readlines()[1]) - 1)
            b = np.loadtxt(os.path.join(outs[1], f), skiprows=3 + int(open(os.path.join(outs[1], f)).readlines()[1]) - 1)
            assert a.shape == b.shape and a.shape[1] >= 2 and np.allclose(a, b, rtol=2e-5, atol=1e-9), f
        else:
            a, _ = csf.read_map(os.path.join(outs[0], f))
            b, _ = csf.read_map(os.path.join(outs[1], f))
            assert a.shape == b.shape == (70, 64)
            if f.lower().startswith("qall"):
                assert np.allclose(a, b, rtol=1e-5, atol=1e-12, equal_nan=True), f
            else:
                assert np.array_equal(a, b, equal_nan=True), f          # the water balance does not depend on the split
