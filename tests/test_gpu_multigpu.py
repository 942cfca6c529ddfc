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


@pytest.mark.parametrize("exchange,glacier,overlap", [("p2p", False, False), ("nccl", False, False), ("p2p", True, False),
                                                      ("p2p", False, True), ("nccl", False, True)])
def test_real_two_gpu_exchange(exchange, glacier, overlap):
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
    port = 29600 + (os.getpid() % 300) + (0 if exchange == "p2p" else 301) + (602 if glacier else 0) + (1203 if overlap else 0)
    env = dict(os.environ, SPHY_MG_EXCHANGE=exchange)
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
                        "--master-port", str(port), os.path.join(root, "tools", "check_multigpu.py"), "--rows", "1500", "--cols", "1300",
                        "--steps", "4"] + (["--glacier"] if glacier else []) + (["--overlap"] if overlap else []), env=env, capture_output=True, text=True, timeout=900, cwd=root)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-3000:]
    line = json.loads([ln for ln in r.stdout.splitlines() if ln.startswith('{"tool"')][-1])
    assert line["exchange"] == exchange and line["world"] == 2 and line["overlap_routing"] == overlap
    assert line["worst_rel_partitioned_vs_one_gpu"] <= 3e-7 and line["worst_rel_partitioned_vs_level_sweep"] <= 1e-5
