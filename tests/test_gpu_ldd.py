"""GPU: K3 (LDD -> integer structures) and the accuflux family vs the CPU oracle, through the C ABI.
Integer structures must be bit-exact; accuflux/upstream are float32 and must also be bit-exact
(per-cell REAL8 sum stored REAL4 makes the result independent of donor order)."""
import ctypes as C

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _build(lib, torch, ldd, active=None, cellsize=100.0):
    from sphy_b200 import _lib
    ctx = C.c_void_p()
    _lib.check(lib, None, lib.sphy_ctx_create(C.byref(ctx), 0, ldd.shape[0], ldd.shape[1], C.c_float(cellsize)), "create")
    ldd_d = torch.from_numpy(np.ascontiguousarray(ldd, np.uint8)).cuda()
    act_d = None if active is None else torch.from_numpy(np.ascontiguousarray(active, np.uint8)).cuda()
    rc = lib.sphy_ldd_build(ctx, ldd_d.data_ptr(), act_d.data_ptr() if act_d is not None else None, None)
    return ctx, rc


def _export(lib, torch, ctx, name, dtype, count):
    out = torch.empty(max(count, 1), dtype=dtype, device="cuda")
    from sphy_b200 import _lib
    _lib.check(lib, ctx, lib.sphy_ldd_export(ctx, name.encode(), out.data_ptr(), out.numel() * out.element_size(), None), name)
    torch.cuda.synchronize()
    return out[:count].cpu().numpy()


def _check_structures(lib, torch, ldd, active=None):
    from oracle.ldd_oracle import LddStruct
    st = LddStruct(ldd, active)
    ctx, rc = _build(lib, torch, ldd, active)
    assert rc == 0, lib.sphy_last_error(ctx)
    n = int(lib.sphy_ncells(ctx))
    assert n == st.n
    assert int(lib.sphy_nlevels(ctx)) == st.nlevels
    assert np.array_equal(_export(lib, torch, ctx, "cidx", torch.int32, ldd.size), st.cidx.ravel())
    assert np.array_equal(_export(lib, torch, ctx, "down", torch.int32, n), st.down)
    assert np.array_equal(_export(lib, torch, ctx, "indeg", torch.uint8, n), st.indeg)
    assert np.array_equal(_export(lib, torch, ctx, "level", torch.int32, n), st.level)
    assert np.array_equal(_export(lib, torch, ctx, "order", torch.int32, n), st.order)
    assert np.array_equal(_export(lib, torch, ctx, "level_ptr", torch.int32, st.nlevels + 1), st.level_ptr)
    assert np.array_equal(_export(lib, torch, ctx, "don_ptr", torch.int32, n + 1), st.don_ptr)
    assert np.array_equal(_export(lib, torch, ctx, "don_idx", torch.int32, len(st.don_idx)), st.don_idx)
    assert np.array_equal(_export(lib, torch, ctx, "nup", torch.int64, n), st.nup)
    return ctx, st


@pytest.fixture(scope="module")
def env():
    import torch
    from sphy_b200 import _lib
    assert torch.cuda.is_available()
    return _lib.load(), torch


@pytest.mark.parametrize("shape,outlets,seed", [((7, 5), "single", 1), ((64, 48), "single", 2), ((131, 257), "border", 3),
                                                ((500, 500), "single", 4), ((1200, 900), "border", 5)])
def test_structures_bit_exact(env, shape, outlets, seed):
    lib, torch = env
    from sphy_b200 import synth
    _, _, ldd, _ = synth.make_dem_ldd(shape[0], shape[1], 100.0, seed=seed, outlets=outlets)
    ctx, _ = _check_structures(lib, torch, ldd)
    lib.sphy_ctx_destroy(ctx)


def test_masked_catchment_and_mv(env):
    lib, torch = env
    from sphy_b200 import synth
    _, _, ldd, _ = synth.make_dem_ldd(90, 70, 100.0, seed=9, outlets="single")
    rng = np.random.default_rng(0)
    active = np.ones(ldd.shape, np.uint8)
    rr, cc = np.ogrid[:90, :70]
    active[(rr - 30) ** 2 + (cc - 20) ** 2 < 150] = 0          # a hole: cells draining into it become outlets
    ldd = ldd.copy()
    ldd[rng.random(ldd.shape) < 0.01] = 255                    # scattered MV codes
    ctx, _ = _check_structures(lib, torch, ldd, active)
    lib.sphy_ctx_destroy(ctx)


def test_degenerate_rasters(env):
    lib, torch = env
    for ldd in (np.full((1, 1), 5, np.uint8), np.full((3, 4), 5, np.uint8), np.array([[6, 6, 6, 6, 5]], np.uint8),
                np.array([[2], [2], [2], [5]], np.uint8)):
        ctx, _ = _check_structures(lib, torch, ldd)
        lib.sphy_ctx_destroy(ctx)
    # no active cell at all -> clean error, not a crash
    ctx, rc = _build(lib, torch, np.full((4, 4), 255, np.uint8))
    assert rc == -1
    lib.sphy_ctx_destroy(ctx)


def test_cycle_is_detected(env):
    lib, torch = env
    ldd = np.array([[6, 2], [8, 4]], np.uint8)                 # 4-cycle
    ctx, rc = _build(lib, torch, ldd)
    assert rc == -3 and b"cycle" in lib.sphy_last_error(ctx)
    lib.sphy_ctx_destroy(ctx)
    # calls before a successful build are refused
    assert lib.sphy_upstream(ctx if False else None, None, None, None) == -1


@pytest.mark.parametrize("shape,seed", [((40, 30), 1), ((300, 400), 2), ((1000, 1000), 3)])
def test_accuflux_upstream_bit_exact(env, shape, seed):
    lib, torch = env
    from sphy_b200 import _lib, synth
    _, _, ldd, _ = synth.make_dem_ldd(shape[0], shape[1], 100.0, seed=seed, outlets="single")
    ctx, st = _check_structures(lib, torch, ldd)
    rng = np.random.default_rng(seed)
    m = (rng.gamma(0.7, 3.0, st.n) * (rng.random(st.n) < 0.6)).astype(np.float32)
    canon = _export(lib, torch, ctx, "cell_order", torch.int32, st.n).astype(np.int64)   # device index -> canonical
    m_d = torch.from_numpy(m[canon]).cuda()
    out = torch.empty_like(m_d)
    _lib.check(lib, ctx, lib.sphy_accuflux(ctx, m_d.data_ptr(), None, out.data_ptr(), 0, None), "accuflux")
    torch.cuda.synchronize()
    got = out.cpu().numpy()
    want = st.accuflux(m)[canon]
    assert np.array_equal(got.view(np.int32), want.view(np.int32))
    # accuflux(ldd, 1) == upstream cell count (exact below 2**24), a pit holds its catchment total
    ones = torch.ones(st.n, device="cuda")
    _lib.check(lib, ctx, lib.sphy_accuflux(ctx, ones.data_ptr(), None, out.data_ptr(), 0, None), "accuflux")
    torch.cuda.synchronize()
    if st.n < 2 ** 24:
        assert np.array_equal(out.cpu().numpy().astype(np.int64), st.nup[canon])
    # fraction cut (QFRAC in {0,1}): nothing crosses a cut cell
    frac = np.ones(st.n, np.float32)
    frac[rng.choice(st.n, size=max(1, st.n // 200), replace=False)] = 0.0
    f_d = torch.from_numpy(frac[canon]).cuda()
    _lib.check(lib, ctx, lib.sphy_accuflux(ctx, m_d.data_ptr(), f_d.data_ptr(), out.data_ptr(), 0, None), "accufractionflux")
    torch.cuda.synchronize()
    want = st.accuflux(m, frac)[canon]
    got = out.cpu().numpy()
    assert np.array_equal(got.view(np.int32), want.view(np.int32))
    assert (got[frac[canon] == 0] == 0).all()
    # upstream
    _lib.check(lib, ctx, lib.sphy_upstream(ctx, m_d.data_ptr(), out.data_ptr(), None), "upstream")
    torch.cuda.synchronize()
    assert np.array_equal(out.cpu().numpy().view(np.int32), st.upstream(m)[canon].view(np.int32))
    # linearity (size-independent property): accuflux(a*m) == a*accuflux(m) for a power of two
    m2 = torch.from_numpy((m * np.float32(4.0))[canon]).cuda()
    out2 = torch.empty_like(m_d)
    _lib.check(lib, ctx, lib.sphy_accuflux(ctx, m2.data_ptr(), None, out2.data_ptr(), 0, None), "accuflux")
    _lib.check(lib, ctx, lib.sphy_accuflux(ctx, m_d.data_ptr(), None, out.data_ptr(), 0, None), "accuflux")
    torch.cuda.synchronize()
    assert torch.equal(out2, out * 4.0)
    lib.sphy_ctx_destroy(ctx)
