"""CPU: the host-side logic of the multi-GPU path -- partition planning (sphy_b200/multigpu.py) and the per-step
confluence exchange -- with world_size 2 over gloo (and 1..5 ranks emulated in-process), against the oracle's
float64 accuflux.  The kernels themselves are covered by tests/test_gpu_multigpu.py."""
import os

import numpy as np
import pytest

from oracle.ldd_oracle import LddStruct
from sphy_b200 import multigpu, synth
from tests.mgpu_util import dfs_order, rank_begin, rank_finish


def _case(nrows, ncols, seed, outlets="single"):
    _, _, ldd, _ = synth.make_dem_ldd(nrows, ncols, 100.0, seed=seed, outlets=outlets)
    st = LddStruct(ldd)
    pre, sub_size = dfs_order(st)
    rng = np.random.default_rng(seed)
    m = (rng.gamma(0.8, 3.0, st.n) * (rng.random(st.n) < 0.7)).astype(np.float32)
    want = np.empty(st.n)
    want[pre] = st.accuflux_f64(m)                  # oracle, reordered to device (DFS) order
    vals = np.empty(st.n, np.float32)
    vals[pre] = m
    return sub_size, vals, want


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
    sub_size, vals, want = _case(shape[0], shape[1], seed=world + 3, outlets=outlets)
    b, ranks, stride = multigpu.plan(sub_size, world)
    assert len(ranks) == world and stride >= 1
    begun = [rank_begin(vals[b[p]:b[p + 1]], sub_size[b[p]:b[p + 1]], ranks[p], stride) for p in range(world)]
    gathered = [x[1] for x in begun]
    got = np.concatenate([rank_finish(begun[p][0], begun[p][2], ranks[p], p, gathered) for p in range(world)])
    assert not np.isnan(got).any()
    assert np.allclose(got, want, rtol=1e-9, atol=1e-9)
    for p in range(world):                           # plan invariants
        me = ranks[p]
        assert (np.diff(me["pub_e"]) > 0).all() if len(me["pub_e"]) > 1 else True
        assert (me["remote_rank"] > p).all()         # ranges only ever extend to higher ranks (pre-order)
        assert len(me["remote_k"]) == len(me["remote_rank"]) == len(me["remote_slot"])


def _worker(rank, world, port, ret):
    import torch
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    sub_size, vals, want = _case(200, 160, seed=7)
    b, ranks, stride = multigpu.plan(sub_size, world)
    lo, hi = b[rank], b[rank + 1]
    acc, send, partial = rank_begin(vals[lo:hi], sub_size[lo:hi], ranks[rank], stride)
    recv = torch.zeros(world * stride, dtype=torch.float64)
    dist.all_gather_into_tensor(recv, torch.from_numpy(send))          # the per-step confluence exchange
    got = rank_finish(acc, partial, ranks[rank], rank, recv.numpy().reshape(world, stride))
    ok = bool(np.allclose(got, want[lo:hi], rtol=1e-9, atol=1e-9)) and not np.isnan(got).any()
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
