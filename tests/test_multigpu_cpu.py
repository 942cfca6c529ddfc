"""CPU: the host-side logic of the scan routing's trunk list and of the multi-GPU path -- partition bounds, the exchange
plan (sphy_b200/multigpu.py: plan_trunk) and the per-step confluence exchange -- with world_size 2 over gloo (and 1..5
ranks emulated in-process), against the oracle: float64 accuflux for the ordinary cells, and the REAL4-per-cell
accuflux (oracle/ldd_oracle.c) for what the rounding-drift accumulation on the trunk cells has to reproduce.
The kernels themselves are covered by tests/test_gpu_multigpu.py and tests/test_gpu_scan.py."""
import os

import numpy as np
import pytest

from oracle.ldd_oracle import LddStruct
from sphy_b200 import multigpu, synth
from tests.mgpu_util import device_order_structures, dfs_order, rank_begin, trunk_finish, trunk_select


def _case(nrows, ncols, seed, outlets="single", uniform=False):
    _, _, ldd, _ = synth.make_dem_ldd(nrows, ncols, 100.0, seed=seed, outlets=outlets)
    st = LddStruct(ldd)
    pre, sub_size = dfs_order(st)
    down, level = device_order_structures(st, pre)
    rng = np.random.default_rng(seed)
    m = (np.full(st.n, 3.3) if uniform else rng.gamma(0.8, 3.0, st.n) * (rng.random(st.n) < 0.7)).astype(np.float32)
    exact, real4 = np.empty(st.n), np.empty(st.n, np.float32)
    exact[pre] = st.accuflux_f64(m)                 # oracle, reordered to device (DFS) order
    real4[pre] = st.accuflux(m)                     # the reference's REAL4-per-cell arithmetic
    vals = np.empty(st.n, np.float32)
    vals[pre] = m
    return dict(sub_size=sub_size, down=down, level=level, vals=vals, exact=exact, real4=real4, n=st.n)


def _route(case, world, min_level, min_cells, exchange=None):
    """Every rank's begin -> exchange -> finish; returns the routed REAL4 map of the whole basin and the listed cells."""
    n, sub_size = case["n"], case["sub_size"]
    b = multigpu.partition_bounds(n, world) if world > 1 else np.array([0, n], np.int64)
    cell, sub, is_trunk, last, parent_slot = trunk_select(sub_size, case["level"], case["down"], min_level, min_cells, b)
    plan = multigpu.plan_trunk(cell, sub, b)
    out = np.empty(n, np.float32)
    sends = []
    for p in range(world):
        lo, hi = int(b[p]), int(b[p + 1])
        mine = cell[(cell >= lo) & (cell < hi)] - lo
        acc, send = rank_begin(case["vals"][lo:hi], sub_size[lo:hi], plan["pub"][p], mine)
        out[lo:hi] = acc.astype(np.float32)
        sends.append(send)
    gathered = exchange(sends, plan["stride"]) if exchange else sends
    out[cell] = trunk_finish(plan, gathered, is_trunk, last, parent_slot)
    return out, cell, plan


def test_bounds_are_tile_aligned_and_balanced():
    b = multigpu.partition_bounds(10_000_000, 8)
    assert b[0] == 0 and b[-1] == 10_000_000 and (np.diff(b) > 0).all()
    assert (b[:-1] % multigpu.TILE == 0).all()
    assert np.diff(b).max() - np.diff(b).min() <= multigpu.TILE
    with pytest.raises(ValueError):
        multigpu.partition_bounds(1000, 8)


@pytest.mark.parametrize("world", [1, 2, 3, 5])
@pytest.mark.parametrize("shape,outlets", [((150, 130), "single"), ((170, 140), "border")])
def test_plan_and_exchange_emulated(world, shape, outlets):
    """No trunk (thresholds out of reach): the listed cells are exactly those whose range leaves their rank, and the
    result is the float64 accumulation rounded once -- whatever the number of ranks."""
    case = _case(shape[0], shape[1], seed=world + 3, outlets=outlets)
    got, cell, plan = _route(case, world, 10 ** 9, 10 ** 12)
    assert not np.isnan(got).any()
    assert np.allclose(got, case["exact"], rtol=2e-7, atol=1e-9)
    assert (cell.size == 0) == (world == 1)
    for p in range(world):                           # plan invariants
        pub = plan["pub"][p]
        assert (np.diff(pub) > 0).all() if len(pub) > 1 else True
        assert plan["stride"] >= 1 + len(pub)
    assert (plan["end_rank"] >= plan["before_rank"]).all()           # ranges only ever extend to higher ranks (pre-order)


@pytest.mark.parametrize("uniform", [False, True])
@pytest.mark.parametrize("world", [1, 3])
def test_rounding_drift_follows_the_real4_reference(world, uniform):
    """The reference rounds every cell's sum to REAL4.  With (nearly) equal runoff everywhere those roundings are biased
    and the reference drifts away from exact accumulation; the trunk accumulation must follow it, independent of the
    number of ranks."""
    case = _case(500, 450, seed=3, uniform=uniform)
    den = np.maximum(np.abs(case["real4"]), 1e-30)
    plain = np.max(np.abs(case["exact"].astype(np.float32).astype(np.float64) - case["real4"]) / den)
    got, cell, _ = _route(case, world, 16, 64)
    err = np.max(np.abs(got.astype(np.float64) - case["real4"]) / den)
    assert cell.size > 30000
    assert err <= 5e-7, err
    if uniform:
        assert plain > 2e-6 and err < 0.25 * plain
    one, _, _ = _route(case, 1, 16, 64)
    # the drift sums are exact, so the split only shows through the float64 prefixes: at most one REAL4 ulp, rarely
    # (the absolute term covers the cancellation noise of this emulation's rank-wide np.cumsum on tiny fluxes; the kernels
    # use tile-local prefixes for the ordinary cells)
    diff = np.abs(got.astype(np.float64) - one)
    assert (diff <= 1.3e-7 * den + 1e-10).all() and np.mean(diff > 0) < 1e-3


def _worker(rank, world, port, ret):
    import torch
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    case = _case(200, 160, seed=7)

    def exchange(sends, stride):
        mine = np.zeros(stride)
        mine[:sends[rank].size] = sends[rank]                          # this process only owns its own rank's block
        recv = torch.zeros(world * stride, dtype=torch.float64)
        dist.all_gather_into_tensor(recv, torch.from_numpy(mine))      # the per-step confluence exchange
        return list(recv.numpy().reshape(world, stride))

    got, cell, _ = _route(case, world, 32, 256, exchange=exchange)
    b = multigpu.partition_bounds(case["n"], world)
    lo, hi = int(b[rank]), int(b[rank + 1])
    den = np.maximum(np.abs(case["real4"]), 1e-30)
    ok = bool(np.max(np.abs(got[lo:hi].astype(np.float64) - case["real4"][lo:hi]) / den[lo:hi]) <= 5e-7) and cell.size > 0
    flag = torch.tensor([1 if ok else 0])
    dist.all_reduce(flag, op=dist.ReduceOp.MIN)
    if rank == 0:
        ret.put(int(flag.item()))
    dist.barrier()
    dist.destroy_process_group()


def test_exchange_world2_gloo():
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    ret = ctx.Queue()
    port = 29500 + (os.getpid() % 2000)
    procs = [ctx.Process(target=_worker, args=(r, 2, port, ret)) for r in range(2)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(120)
        assert p.exitcode == 0
    assert ret.get(timeout=5) == 1


def test_plan_reslake_edge_cases():
    """No structures at all; a structure on the outlet cell (nothing before it in the DFS order); unsorted input is refused."""
    _, _, ldd, _ = synth.make_dem_ldd(90, 80, 100.0, seed=4, outlets="single")
    st = LddStruct(ldd)
    pre, sub_size = dfs_order(st)
    down, level = device_order_structures(st, pre)
    b = multigpu.partition_bounds(st.n, 3)
    cell, sub, *_ = trunk_select(sub_size, level, down, 2 ** 31 - 1, 2 ** 62, b, np.zeros(0, np.int64))
    rl = multigpu.plan_reslake(cell, sub, np.zeros(0, np.int64), np.zeros((0, 8), np.int64), b)
    assert rl["n_struct"] == 0 and rl["entry_idx"].size == 0 and (rl["entry_cut"] == -1).all() and rl["entry_ptr"][-1] == 0
    cut = np.array([0, int(np.flatnonzero(sub_size > 500)[-1])], np.int64)          # the outlet itself and a tributary
    cell, sub, *_ = trunk_select(sub_size, level, down, 2 ** 31 - 1, 2 ** 62, b, cut)
    rl = multigpu.plan_reslake(cell, sub, cut, np.full((2, 8), -1, np.int64), b)
    assert rl["cut_slot"][0] == 0 and rl["entry_cut"][0] == 0
    ptr, idx = rl["entry_ptr"], rl["entry_idx"]
    assert list(idx[ptr[0]:ptr[1]]) == [1]                     # the tributary structure is top-level inside the outlet's range
    inner = rl["cut_slot"][1]
    assert ptr[inner + 1] == ptr[inner]                        # nothing upstream of it
    assert (rl["x_rank"].reshape(2, 9)[:, :8] == -1).all() and (rl["x_rank"].reshape(2, 9)[:, 8] >= 0).all()
    with pytest.raises(ValueError):
        multigpu.plan_reslake(cell, sub, cut[::-1], np.full((2, 8), -1, np.int64), b)
    with pytest.raises(ValueError):
        multigpu.plan_reslake(cell[1:], sub[1:], cut, np.full((2, 8), -1, np.int64), b)   # the outlet structure is not listed


@pytest.mark.parametrize("world", [1, 2, 4])
def test_lakes_and_reservoirs_on_the_trunk_list(world):
    """Lakes / reservoirs on a (partitioned) basin: the plain scan plus a correction of the cells downstream of the
    structures (multigpu.plan_reslake, what sphy_route_reslake_begin / _finish run on the device) against
    accufractionflux + upstream of the oracle, several days, nested structures, any number of ranks."""
    from tests.mgpu_util import ReslakeEmulation, reslake_rule
    _, _, ldd, _ = synth.make_dem_ldd(190, 170, 100.0, seed=11, outlets="single")
    st = LddStruct(ldd)
    n = st.n
    pre, sub_size = dfs_order(st)
    down, level = device_order_structures(st, pre)
    rng = np.random.default_rng(5)
    # structures on streams of very different size, some nested on the main stem, one right below another
    big = np.flatnonzero(sub_size > 400)
    cut = np.unique(np.concatenate([rng.choice(big, 9, replace=False), rng.choice(np.flatnonzero((sub_size > 20) & (sub_size < 200)), 5, replace=False)]))
    stem = np.flatnonzero(sub_size > n // 3)
    cut = np.unique(np.concatenate([cut, stem[[len(stem) // 2, len(stem) // 2 + 1]]]))       # receiver of one is a structure itself
    if world != 2:
        cut = cut[cut > 0]
    else:
        cut = np.unique(np.concatenate([[0], cut]))             # a reservoir on the outlet cell itself
    ns = cut.size
    canon = np.empty(n, np.int64)
    canon[pre] = np.arange(n)                                   # device index -> canonical cell
    donors = np.full((ns, 8), -1, np.int64)
    for j, c in enumerate(cut):
        d = st.don_idx[st.don_ptr[canon[c]]:st.don_ptr[canon[c] + 1]]
        donors[j, :len(d)] = pre[d]
    qfrac = np.ones(n, np.float32)
    qfrac[cut] = 0
    kx, vol = 0.93, np.float32(0.001) * np.float32(100.0 * 100.0)
    S0 = rng.uniform(1e4, 1e6, ns).astype(np.float32)
    b = multigpu.partition_bounds(n, world) if world > 1 else np.array([0, n], np.int64)
    em = ReslakeEmulation(sub_size, level, down, b, cut, donors, qfrac, kx, vol, S0)
    assert set(cut) <= set(em.cell)
    # reference: the reference's own sequence (advanced_routing.py:134-217) with the oracle's accufractionflux / upstream
    f32 = np.float32
    fr_c = qfrac[pre]                                           # canonical order
    S, q_c = S0.copy(), np.zeros(n, np.float32)
    cut_c = canon[cut]
    for day in range(5):
        totr = (rng.gamma(0.8, 4.0, n) * (rng.random(n) < 0.6)).astype(np.float32)          # device order
        got = em.step(totr)
        totr_c = totr[pre]
        S = (S + vol * totr_c[cut_c]).astype(np.float32)
        Qout = reslake_rule(S)
        qo = np.zeros(n, np.float32)
        qo[cut_c] = Qout
        m = (st.upstream(qo) + np.where(fr_c == 0, f32(0), vol * totr_c)).astype(np.float32)
        F = st.accuflux_f64(m, fr_c).astype(np.float32)
        qd = np.where(fr_c == 0, qo, f32(1 - kx) * F + ((q_c * f32(24)) * f32(3600)) * f32(kx)).astype(np.float32)
        Qin = st.upstream(qd)[cut_c]
        S = ((S - Qout) + Qin).astype(np.float32)
        q_c = (qd / f32(86400)).astype(np.float32)
        want = q_c[canon]
        assert np.allclose(got, want, rtol=3e-6, atol=1e-9), (day, np.max(np.abs(got - want) / np.maximum(np.abs(want), 1e-9)))
        em.flush() if day % 2 else None                        # the storage update may arrive late or right away
        if not em.pending:
            assert np.allclose(em.S, S, rtol=2e-6), day
    em.flush()
    assert np.allclose(em.S, S, rtol=2e-6) and np.allclose(em.Qin, Qin, rtol=2e-6, atol=1e-6)


def test_tss_tables_of_a_split_basin_merge_by_station_id():
    """Every rank samples the stations inside its own range; rank 0 merges the columns by ascending station id
    (sphy_b200/reporting.py: merge_tss), also when a rank owns no station or misses a step."""
    from sphy_b200.reporting import merge_tss
    every = [{"ids": [4, 1], "tss": {"QAllDTS": [(1, [40.0, 10.0]), (2, [41.0, 11.0])], "TotRDTS": [(1, [4.0, 1.0])]}},
             {"ids": [], "tss": {}},
             {"ids": [3], "tss": {"QAllDTS": [(1, [30.0]), (2, [31.0])]}}]
    out = merge_tss(every)
    ids, rows = out["QAllDTS"]
    assert ids.tolist() == [1, 3, 4]
    assert [t for t, _ in rows] == [1, 2] and rows[0][1].tolist() == [10.0, 30.0, 40.0] and rows[1][1].tolist() == [11.0, 31.0, 41.0]
    ids, rows = out["TotRDTS"]
    assert rows[0][1][0] == 1.0 and np.isnan(rows[0][1][1]) and rows[0][1][2] == 4.0
    with pytest.raises(ValueError):
        merge_tss([{"ids": [1], "tss": {}}, {"ids": [1], "tss": {}}])


def test_write_tss_single_process_path(tmp_path):
    """Reporting.write_tss on one GPU: the model's own station ids and rows go straight to the PCRaster .tss layout."""
    import types
    from sphy_b200.reporting import Reporting
    m = types.SimpleNamespace(world=1, rank=0, station_ids=np.array([1, 2, 3]), outpath=str(tmp_path))
    rep = types.SimpleNamespace(m=m, tss={"QAllDTS": [(1, np.array([1., 2., 3.], np.float32)), (2, np.array([4., 5., 6.], np.float32))]})
    rep.tss_tables = lambda: Reporting.tss_tables(rep)
    Reporting.write_tss(rep)
    lines = open(os.path.join(str(tmp_path), "QAllDTS.tss")).read().splitlines()
    assert lines[:6] == ["timeseries scalar", "4", "timestep", "1", "2", "3"]
    assert [float(x) for x in lines[6].split()] == [1, 1, 2, 3] and [float(x) for x in lines[7].split()] == [2, 4, 5, 6]


@pytest.mark.parametrize("name", ["c3_res_lake", "c3_res_only", "c3_lake_only"])
def test_structure_path_reproduces_the_reference_golden(name):
    """The algorithm behind sphy_route_reslake_begin / _finish (plain range sums, E(k) - Qout(k) taken out downstream of every
    structure, storage fed by upstream(ldd, Q) one step late) driven by the daily TotR maps the UNMODIFIED reference reported, with
    the port's lakes.QLake / reservoirs.QRes as the outflow functions: discharge and storage must follow the reference's own
    QRAold / StorRES states of the golden run (within the scan path's 1e-5; the reference rounds every cell's sum to REAL4)."""
    from oracle.sphy_port import PortInputs, PortModel
    from tests.golden_util import catchment_from_golden
    from tests.mgpu_util import ReslakeEmulation
    cat, z, meta = catchment_from_golden(name)
    inp = PortInputs(cat)
    port = PortModel(inp)                                       # only its structure functions and tables are used
    st = inp.lddst
    n = st.n
    pre, sub_size = dfs_order(st)
    down, level = device_order_structures(st, pre)
    canon = np.empty(n, np.int64)
    canon[pre] = np.arange(n)
    cut_c = np.flatnonzero(port.QFRAC == 0)                      # canonical structure cells
    cut = np.sort(pre[cut_c])
    cut_c = canon[cut]                                          # reordered like the device list
    donors = np.full((cut.size, 8), -1, np.int64)
    for j, c in enumerate(cut_c):
        d = st.don_idx[st.don_ptr[c]:st.don_ptr[c + 1]]
        donors[j, :len(d)] = pre[d]
    qfrac = np.ones(n, np.float32)
    qfrac[cut] = 0
    g = st.gather

    def rule(S):                                                # the reference's outflow functions on the emulation's storages
        port.StorRES = np.zeros(n, np.float32)
        port.StorRES[cut_c] = S
        q = np.zeros(n, np.float32)
        if inp.lake is not None:
            q[inp.lake["cells"]] = port.QLake(None)
        if inp.res is not None:
            q[inp.res["cells"]] = port.QRes()
        return q[cut_c].astype(np.float32)

    S0 = g(z["state/StorRES"][0])[cut_c]
    em = ReslakeEmulation(sub_size, level, down, np.array([0, n], np.int64), cut, donors, qfrac, float(inp.kx),
                          np.float32(0.001) * np.float32(inp.cellarea), S0, rule=rule)
    em.q = g(z["state/QRAold"][0])[canon].astype(np.float32)
    import datetime
    for t in range(1, meta["nsteps"] + 1):
        port.curdate = cat.start + datetime.timedelta(days=t - 1)
        totr = g(z["report/TotR"][t - 1])[canon]                # the reference's own TotR of the day, device order
        got = em.step(totr)
        want = g(z["state/QRAold"][t])[canon]
        assert np.allclose(got, want, rtol=1e-5, atol=1e-12), (name, t, float(np.max(np.abs(got - want) / np.maximum(np.abs(want), 1e-12))))
        em.flush()
        assert np.allclose(em.S, g(z["state/StorRES"][t])[cut_c], rtol=1e-5, atol=1e-3), (name, t)


def _reslake_worker(rank, world, port, ret):
    """One process per rank: the blocks (prefixes + per-structure extras) really travel through an all-gather."""
    import torch
    import torch.distributed as dist
    from tests.mgpu_util import ReslakeEmulation, reslake_rule
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    _, _, ldd, _ = synth.make_dem_ldd(120, 110, 100.0, seed=9, outlets="single")
    st = LddStruct(ldd)
    n = st.n
    pre, sub_size = dfs_order(st)
    down, level = device_order_structures(st, pre)
    rng = np.random.default_rng(2)
    cut = np.unique(np.concatenate([rng.choice(np.flatnonzero(sub_size > 300), 6, replace=False), rng.choice(np.flatnonzero(sub_size > n // 4), 2)]))
    cut = cut[cut > 0]
    canon = np.empty(n, np.int64)
    canon[pre] = np.arange(n)
    donors = np.full((cut.size, 8), -1, np.int64)
    for j, c in enumerate(cut):
        d = st.don_idx[st.don_ptr[canon[c]]:st.don_ptr[canon[c] + 1]]
        donors[j, :len(d)] = pre[d]
    qfrac = np.ones(n, np.float32)
    qfrac[cut] = 0
    kx, vol = 0.9, np.float32(10.0)
    S0 = rng.uniform(1e4, 1e6, cut.size).astype(np.float32)
    b = multigpu.partition_bounds(n, world)
    em = ReslakeEmulation(sub_size, level, down, b, cut, donors, qfrac, kx, vol, S0)

    def exchange(blocks, stride):
        recv = torch.zeros(world * stride, dtype=torch.float64)
        dist.all_gather_into_tensor(recv, torch.from_numpy(np.ascontiguousarray(blocks[rank])))    # only this rank's own block
        return list(recv.numpy().reshape(world, stride))

    f32 = np.float32
    fr_c, cut_c = qfrac[pre], canon[cut]
    S, q_c, ok = S0.copy(), np.zeros(n, np.float32), True
    lo, hi = int(b[rank]), int(b[rank + 1])
    for day in range(4):
        totr = (rng.gamma(0.8, 4.0, n) * (rng.random(n) < 0.6)).astype(np.float32)
        got = em.step(totr, exchange=exchange)
        totr_c = totr[pre]
        S = (S + vol * totr_c[cut_c]).astype(np.float32)
        Qout = reslake_rule(S)
        qo = np.zeros(n, np.float32)
        qo[cut_c] = Qout
        m = (st.upstream(qo) + np.where(fr_c == 0, f32(0), vol * totr_c)).astype(np.float32)
        F = st.accuflux_f64(m, fr_c).astype(np.float32)
        qd = np.where(fr_c == 0, qo, f32(1 - kx) * F + ((q_c * f32(24)) * f32(3600)) * f32(kx)).astype(np.float32)
        S = ((S - Qout) + st.upstream(qd)[cut_c]).astype(np.float32)
        q_c = (qd / f32(86400)).astype(np.float32)
        ok = ok and bool(np.allclose(got[lo:hi], q_c[canon][lo:hi], rtol=3e-6, atol=1e-9))
    em.flush()
    ok = ok and bool(np.allclose(em.S, S, rtol=2e-6))
    flag = torch.tensor([1 if ok else 0])
    dist.all_reduce(flag, op=dist.ReduceOp.MIN)
    if rank == 0:
        ret.put(int(flag.item()))
    dist.barrier()
    dist.destroy_process_group()


def test_structure_exchange_world2_gloo():
    """World size 2 over gloo: every process contributes only its own rank's block (confluence prefixes + the per-structure
    extras it owns); lakes / reservoirs downstream of the split see the other rank's water and structures."""
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    ret = ctx.Queue()
    port = 31500 + (os.getpid() % 2000)
    procs = [ctx.Process(target=_reslake_worker, args=(r, 2, port, ret)) for r in range(2)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(180)
        assert p.exitcode == 0
    assert ret.get(timeout=5) == 1


def _random_forest(rng, n, p_root=0.002):
    """A random drainage forest with at most 8 donors per cell, numbered in depth-first pre-order (donors by ascending
    original index): -> (down, sub_size) by device index."""
    parent = np.full(n, -1, np.int64)
    nchild = np.zeros(n, np.int64)
    for i in range(1, n):
        if rng.random() < p_root:
            continue
        for _ in range(20):
            # mostly recent cells (long chains), sometimes anywhere (bushy confluences)
            p = int(i - 1 - min(rng.geometric(0.3) - 1, i - 1)) if rng.random() < 0.7 else int(rng.integers(0, i))
            if nchild[p] < 8:
                parent[i] = p
                nchild[p] += 1
                break
    kids = [[] for _ in range(n)]
    for i in range(n):
        if parent[i] >= 0:
            kids[parent[i]].append(i)
    pre = np.empty(n, np.int64)
    size = np.ones(n, np.int64)
    k = 0
    for r in np.flatnonzero(parent < 0):
        stack = [(int(r), 0)]
        while stack:
            v, j = stack.pop()
            if j == 0:
                pre[v] = k
                k += 1
            if j < len(kids[v]):
                stack.append((v, j + 1))
                stack.append((kids[v][j], 0))
            elif parent[v] >= 0:
                size[parent[v]] += size[v]
    down = np.full(n, -1, np.int64)
    has = parent >= 0
    down[pre[has]] = pre[parent[has]]
    sub = np.empty(n, np.int64)
    sub[pre] = size
    return down, sub


@pytest.mark.parametrize("seed,world", [(1, 1), (2, 2), (3, 3), (4, 2)])
def test_structure_plan_on_random_forests(seed, world):
    """plan_reslake + the listed-cell correction on adversarial topologies: several outlets, bushy confluences, structures on
    outlets, structures directly behind each other, structures whose donors live on another rank -- against a brute-force
    accufractionflux walk."""
    from tests.mgpu_util import ReslakeEmulation, reslake_rule
    rng = np.random.default_rng(seed)
    n = 3300
    down, sub_size = _random_forest(rng, n)
    assert np.all(down[1:][down[1:] >= 0] < np.arange(1, n)[down[1:] >= 0])            # pre-order: receivers come first
    roots = np.flatnonzero(down < 0)
    chain = [int(np.argmax(sub_size * (down >= 0)))]                                   # a cell and its receiver: adjacent structures
    chain.append(int(down[chain[0]]))
    cut = np.unique(np.concatenate([rng.choice(n, 25, replace=False), roots[:2], chain]))
    ns = cut.size
    kids = [[] for _ in range(n)]
    for i in range(n):
        if down[i] >= 0:
            kids[down[i]].append(i)
    donors = np.full((ns, 8), -1, np.int64)
    for j, c in enumerate(cut):
        donors[j, :len(kids[c])] = kids[c]
    qfrac = np.ones(n, np.float32)
    qfrac[cut] = 0
    b = multigpu.partition_bounds(n, world) if world > 1 else np.array([0, n], np.int64)
    kx, vol = 0.8, np.float32(7.5)
    S0 = rng.uniform(1e3, 1e5, ns).astype(np.float32)
    em = ReslakeEmulation(sub_size, np.zeros(n, np.int64), down, b, cut, donors, qfrac, kx, vol, S0)
    f32 = np.float32
    S, q = S0.copy(), np.zeros(n, np.float32)
    for day in range(3):
        totr = (rng.gamma(0.8, 4.0, n) * (rng.random(n) < 0.7)).astype(np.float32)
        got = em.step(totr)
        S = (S + vol * totr[cut]).astype(np.float32)
        Qout = reslake_rule(S)
        m = np.where(qfrac == 0, f32(0), vol * totr).astype(np.float64)
        for j, c in enumerate(cut):
            if down[c] >= 0:
                m[down[c]] += float(Qout[j])                    # upstream(ldd, Qout) joins the receiver's material
        F = m.copy()
        for i in range(n - 1, -1, -1):                          # pre-order: every donor has a larger index
            F[i] *= qfrac[i]
            if down[i] >= 0:
                F[down[i]] += F[i]
        qd = (f32(1 - kx) * F.astype(np.float32) + ((q * f32(24)) * f32(3600)) * f32(kx)).astype(np.float32)
        qd[cut] = Qout
        Qin = np.array([np.sum(qd[kids[c]].astype(np.float64)) for c in cut]).astype(np.float32)
        S = ((S - Qout) + Qin).astype(np.float32)
        q = (qd / f32(86400)).astype(np.float32)
        assert np.allclose(got, q, rtol=3e-6, atol=1e-9), (day, float(np.max(np.abs(got - q) / np.maximum(np.abs(q), 1e-9))))
    em.flush()
    assert np.allclose(em.S, S, rtol=3e-6)


@pytest.mark.parametrize("uniform", [False, True])
def test_rounding_drift_on_a_random_chain_heavy_forest(uniform):
    """The drift accumulation on a topology no DEM produces: long chains (hundreds of levels) joined by random confluences of
    very unequal size.  It must still follow the REAL4-per-cell reference, where plain float64 range sums drift away."""
    rng = np.random.default_rng(11)
    n = 30000
    parent = np.full(n, -1, np.int64)
    nchild = np.zeros(n, np.int64)
    for i in range(1, n):
        for _ in range(20):
            p = i - 1 if rng.random() < 0.97 else int(rng.integers(0, i))
            if nchild[p] < 8:
                parent[i] = p
                nchild[p] += 1
                break
    # the construction is already a pre-order only along the sticks: renumber depth-first
    kids = [[] for _ in range(n)]
    for i in range(1, n):
        kids[parent[i]].append(i)
    pre, size, k, stack = np.empty(n, np.int64), np.ones(n, np.int64), 0, [(0, 0)]
    while stack:
        v, j = stack.pop()
        if j == 0:
            pre[v] = k
            k += 1
        if j < len(kids[v]):
            stack.append((v, j + 1))
            stack.append((kids[v][j], 0))
        elif parent[v] >= 0:
            size[parent[v]] += size[v]
    down = np.full(n, -1, np.int64)
    down[pre[1:]] = pre[parent[1:]]
    sub = np.empty(n, np.int64)
    sub[pre] = size
    level = np.zeros(n, np.int64)
    for i in range(n - 1, -1, -1):
        if down[i] >= 0:
            level[down[i]] = max(level[down[i]], level[i] + 1)
    assert level.max() > 300
    m = (np.full(n, 3.3) if uniform else rng.gamma(0.8, 3.0, n) * (rng.random(n) < 0.7)).astype(np.float32)
    exact, acc, real4 = m.astype(np.float64), m.astype(np.float64), np.zeros(n, np.float32)
    for i in range(n - 1, -1, -1):                              # donors have larger indices: REAL4 rounding at every cell
        real4[i] = np.float32(acc[i])
        if down[i] >= 0:
            acc[down[i]] += float(real4[i])
            exact[down[i]] += exact[i]
    case = dict(sub_size=sub, down=down, level=level, vals=m, exact=exact, real4=real4, n=n)
    den = np.maximum(np.abs(real4), 1e-30)
    plain = np.max(np.abs(exact.astype(np.float32).astype(np.float64) - real4) / den)
    for world in (1, 3):
        got, cell, _ = _route(case, world, 16, 64)
        err = np.max(np.abs(got.astype(np.float64) - real4) / den)
        assert cell.size > 10000 and err <= 5e-7, (world, err)
        if uniform:
            assert plain > 1.5e-6 and err < 0.25 * plain
