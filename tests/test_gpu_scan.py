"""GPU: the SPHY_ROUTE_SCAN path (DFS pre-order + float64 prefix sums) vs the oracle and vs the level sweep.
Tolerance 1e-5 relative (north_star): float64 range sums have no per-cell REAL4 rounding along the flow path,
so this path is close to, not bit-identical with, the REAL4-per-cell oracle; the float64 oracle brackets it."""
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


@pytest.mark.parametrize("shape,outlets,seed", [((9, 7), "single", 1), ((120, 90), "border", 2), ((700, 900), "single", 3)])
def test_device_order_is_a_valid_dfs_numbering(shape, outlets, seed):
    import torch
    from oracle.ldd_oracle import LddStruct
    from sphy_b200 import _lib, synth
    lib = _lib.load()
    _, _, ldd, _ = synth.make_dem_ldd(shape[0], shape[1], 100.0, seed=seed, outlets=outlets)
    st = LddStruct(ldd)
    ctx = _ctx(lib, torch, ldd, 100.0)
    n = st.n
    got = {}
    for name in ("cell_order", "sub_size", "flat_active"):
        t = torch.empty(n, dtype=torch.int32, device="cuda")
        _lib.check(lib, ctx, lib.sphy_ldd_export(ctx, name.encode(), t.data_ptr(), n * 4, None), name)
        torch.cuda.synchronize()
        got[name] = t.cpu().numpy().astype(np.int64)
    canon, size = got["cell_order"], got["sub_size"]          # device index -> canonical cell; subtree size by device index
    assert np.array_equal(np.sort(canon), np.arange(n))                     # a permutation
    pre = np.empty(n, np.int64)
    pre[canon] = np.arange(n)                                               # canonical cell -> device index
    assert np.array_equal(size, st.nup[canon])                              # subtree sizes = upstream counts
    assert np.array_equal(got["flat_active"], st.flat_active[canon])        # raster index of every device cell
    has_down = st.down >= 0
    d = st.down[has_down]
    i = np.flatnonzero(has_down)
    assert (pre[i] > pre[d]).all()                                          # receiver first
    assert (pre[i] + st.nup[i] <= pre[d] + st.nup[d]).all()                 # donor range nested in the receiver's
    lib.sphy_ctx_destroy(ctx)


@pytest.mark.parametrize("shape,ncomp,kxmap,seed", [((60, 50), 1, False, 1), ((400, 300), 3, False, 2), ((900, 1100), 2, True, 3)])
def test_scan_routing_matches_oracle(shape, ncomp, kxmap, seed):
    import torch
    from oracle.ldd_oracle import LddStruct
    from sphy_b200 import _lib, synth
    lib = _lib.load()
    cellsize = 250.0
    _, _, ldd, _ = synth.make_dem_ldd(shape[0], shape[1], cellsize, seed=seed, outlets="single")
    st = LddStruct(ldd)
    n = st.n
    ctx = _ctx(lib, torch, ldd, cellsize)
    rng = np.random.default_rng(seed)
    t = torch.empty(n, dtype=torch.int32, device="cuda")
    _lib.check(lib, ctx, lib.sphy_ldd_export(ctx, b"cell_order", t.data_ptr(), n * 4, None), "cell_order")
    torch.cuda.synchronize()
    canon = t.cpu().numpy().astype(np.int64)                  # device index -> canonical cell
    kx_np = (0.9 + 0.09 * rng.random(n)).astype(np.float32) if kxmap else None      # canonical order
    kx_d = torch.from_numpy(kx_np[canon]).cuda() if kxmap else None
    kx = _lib.SphyParam(kx_d.data_ptr() if kxmap else None, 0.971)
    one_m = float(np.float32(1 - 0.971))
    area = np.float32(cellsize) * np.float32(cellsize)
    q_lv = [torch.zeros(n, device="cuda") for _ in range(ncomp)]
    q_sc = [torch.zeros(n, device="cuda") for _ in range(ncomp)]
    q_or = [np.zeros(n, np.float32) for _ in range(ncomp)]
    q_64 = [np.zeros(n, np.float64) for _ in range(ncomp)]
    for step in range(4):
        mm = [(rng.gamma(0.8, 4.0, n) * (rng.random(n) < 0.7)).astype(np.float32) for _ in range(ncomp)]
        mm_d = [torch.from_numpy(m[canon]).cuda() for m in mm]
        src = (C.c_void_p * ncomp)(*[t.data_ptr() for t in mm_d])
        for q, method in ((q_lv, _lib.ROUTE_LEVELS), (q_sc, _lib.ROUTE_SCAN)):
            dst = (C.c_void_p * ncomp)(*[t.data_ptr() for t in q])
            _lib.check(lib, ctx, lib.sphy_route_step(ctx, ncomp, src, dst, kx, C.c_float(one_m), method, None), "route")
        torch.cuda.synchronize()
        for c in range(ncomp):
            rr = (mm[c] * np.float32(0.001) * area) / np.float32(86400)
            ra = st.accuflux(rr)
            ra64 = st.accuflux_f64(rr)
            if kxmap:
                q_or[c] = (np.float32(1) - kx_np) * ra + kx_np * q_or[c]
                q_64[c] = (1.0 - kx_np.astype(np.float64)) * ra64 + kx_np.astype(np.float64) * q_64[c]
            else:
                q_or[c] = np.float32(one_m) * ra + np.float32(0.971) * q_or[c]
                q_64[c] = float(np.float32(one_m)) * ra64 + float(np.float32(0.971)) * q_64[c]
            lv = np.empty(n, np.float32)
            lv[canon] = q_lv[c].cpu().numpy()                 # back to canonical order
            sc = np.empty(n, np.float64)
            sc[canon] = q_sc[c].cpu().numpy().astype(np.float64)
            assert np.array_equal(lv.view(np.int32), q_or[c].view(np.int32)), "level sweep must stay bit-exact"
            tol = 1e-5 * np.abs(q_or[c]) + 1e-12
            assert (np.abs(sc - q_or[c]) <= tol).all(), f"step {step} comp {c}: scan vs REAL4 oracle {np.max(np.abs(sc - q_or[c]) / (np.abs(q_or[c]) + 1e-30)):.2e}"
            # the scan path must be at least as close to the float64 oracle as 1e-6 relative
            assert (np.abs(sc - q_64[c]) <= 2e-6 * np.abs(q_64[c]) + 1e-12).all()
    lib.sphy_ctx_destroy(ctx)


def test_scan_routing_dry_regions_are_exactly_zero():
    """Range sums are differences of float64 prefixes; a sub-catchment without runoff must still route exactly 0 (never a
    +-1e-16 residue, never a negative discharge -- modules/mmf.py:56 raises the routed runoff to the power 1.5)."""
    import torch
    from oracle.ldd_oracle import LddStruct
    from sphy_b200 import _lib, synth
    lib = _lib.load()
    cellsize = 250.0
    _, _, ldd, _ = synth.make_dem_ldd(500, 420, cellsize, seed=9, outlets="single")
    st = LddStruct(ldd)
    n = st.n
    ctx = _ctx(lib, torch, ldd, cellsize)
    t = torch.empty(n, dtype=torch.int32, device="cuda")
    _lib.check(lib, ctx, lib.sphy_ldd_export(ctx, b"cell_order", t.data_ptr(), n * 4, None), "cell_order")
    torch.cuda.synchronize()
    canon = t.cpu().numpy().astype(np.int64)
    rng = np.random.default_rng(9)
    rows = st.flat_active // 420
    mm = (rng.gamma(0.8, 40.0, n) * (rng.random(n) < 0.6)).astype(np.float32)
    mm[(rows < 230) | (rng.random(n) < 0.3)] = 0.0                          # a dry half plus scattered dry cells
    mm_d = torch.from_numpy(mm[canon]).cuda()
    q = torch.zeros(n, device="cuda")
    src, dst = (C.c_void_p * 1)(mm_d.data_ptr()), (C.c_void_p * 1)(q.data_ptr())
    kx = _lib.SphyParam(None, 0.0)
    _lib.check(lib, ctx, lib.sphy_route_step(ctx, 1, src, dst, kx, C.c_float(1.0), _lib.ROUTE_SCAN, None), "route")
    torch.cuda.synchronize()
    got = np.empty(n, np.float32)
    got[canon] = q.cpu().numpy()
    want = st.accuflux((mm * np.float32(0.001) * (np.float32(cellsize) * np.float32(cellsize))) / np.float32(86400))
    assert (want == 0).sum() > n // 4
    assert (got >= 0).all()
    assert (got[want == 0] == 0).all()
    assert (np.abs(got - want) <= 1e-5 * want).all()
    lib.sphy_ctx_destroy(ctx)


@pytest.mark.parametrize("uniform", [False, True])
def test_trunk_drift_follows_the_real4_reference(uniform):
    """With a trunk plan installed the scan path carries the reference's per-cell REAL4 rounding drift on the listed cells
    (csrc/route_scan.cu: k_trunk_drift).  Spatially uniform runoff is the hard case: the reference's roundings are biased
    and its accumulated flux drifts away from exact arithmetic, which plain float64 range sums cannot follow."""
    import torch
    from oracle.ldd_oracle import LddStruct
    from sphy_b200 import _lib, synth
    from tests.test_gpu_multigpu import install_plan
    lib = _lib.load()
    cellsize = 250.0
    _, _, ldd, _ = synth.make_dem_ldd(1500, 1400, cellsize, seed=4, outlets="single")
    st = LddStruct(ldd)
    n = st.n
    plain, trunk = _ctx(lib, torch, ldd, cellsize), _ctx(lib, torch, ldd, cellsize)
    _, arr = install_plan(lib, torch, trunk, np.array([0, n]), 0, 1, 1024, 1100)
    listed = arr["cell"][arr["is_trunk"] != 0].astype(np.int64)
    assert listed.size > 1000
    t = torch.empty(n, dtype=torch.int32, device="cuda")
    _lib.check(lib, plain, lib.sphy_ldd_export(plain, b"cell_order", t.data_ptr(), n * 4, None), "cell_order")
    torch.cuda.synchronize()
    canon = t.cpu().numpy().astype(np.int64)
    rng = np.random.default_rng(4)
    mm = np.full(n, 3.3, np.float32) if uniform else (rng.gamma(0.8, 4.0, n) * (rng.random(n) < 0.7)).astype(np.float32)
    mm_d = torch.from_numpy(mm[canon]).cuda()
    want = st.accuflux((mm * np.float32(0.001) * (np.float32(cellsize) * np.float32(cellsize))) / np.float32(86400))[canon].astype(np.float64)
    kx = _lib.SphyParam(None, 0.0)
    err = {}
    for name, ctx in (("plain", plain), ("trunk", trunk)):
        q = torch.zeros(n, device="cuda")
        src, dst = (C.c_void_p * 1)(mm_d.data_ptr()), (C.c_void_p * 1)(q.data_ptr())
        _lib.check(lib, ctx, lib.sphy_route_step(ctx, 1, src, dst, kx, C.c_float(1.0), _lib.ROUTE_SCAN, None), "route")
        torch.cuda.synchronize()
        got = q.cpu().numpy().astype(np.float64)
        err[name] = np.abs(got - want) / np.maximum(want, 1e-30)
    print(f"uniform={uniform}: listed {listed.size}; on the listed cells plain {err['plain'][listed].max():.2e} -> trunk "
          f"{err['trunk'][listed].max():.2e}; elsewhere {np.delete(err['trunk'], listed).max():.2e}")
    rest = np.ones(n, bool)
    rest[listed] = False
    assert np.array_equal(err["plain"][rest], err["trunk"][rest])           # the ordinary cells are untouched
    assert err["trunk"].max() <= 1e-5
    # the listed cells inherit what their (uncorrected) tributaries deliver, but add nothing of their own
    assert err["trunk"][listed].max() <= max(err["trunk"][rest].max(), 3e-7) * 1.05
    if uniform:
        assert err["trunk"][listed].max() < 0.7 * err["plain"][listed].max()
    lib.sphy_ctx_destroy(plain)
    lib.sphy_ctx_destroy(trunk)
