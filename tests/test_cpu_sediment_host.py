"""CPU: the host-side network derivations of the sediment module (sphy_b200/modules/sediment_transport.py: sub-catchments of
the stations / reservoirs, Strahler stream order; level-vectorised NumPy) against the oracle's cell-by-cell walks."""
import numpy as np
import pytest


@pytest.mark.parametrize("shape,outlets,seed", [((40, 33), "single", 1), ((150, 120), "border", 2), ((260, 300), "single", 3)])
def test_subcatchment_and_streamorder_match_oracle(shape, outlets, seed):
    from oracle.ldd_oracle import LddStruct
    from sphy_b200 import synth
    from sphy_b200.modules.sediment_transport import streamorder_levels, subcatchment_levels
    _, _, ldd, _ = synth.make_dem_ldd(shape[0], shape[1], 100.0, seed=seed, outlets=outlets)
    st = LddStruct(ldd)
    rng = np.random.default_rng(seed)
    pts = np.zeros(st.n, np.int64)
    where = rng.choice(st.n, size=max(3, st.n // 400), replace=False)
    pts[where] = rng.integers(1, 40, size=where.size)
    down, order, lp = st.down.astype(np.int64), st.order.astype(np.int64), st.level_ptr.astype(np.int64)
    assert np.array_equal(subcatchment_levels(down, order, lp, pts), st.subcatchment(pts))
    assert np.array_equal(streamorder_levels(down, order, lp), st.streamorder())
