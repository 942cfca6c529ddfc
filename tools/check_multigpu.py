"""Real multi-GPU check (run under torchrun, one rank per GPU): the basin split across the ranks must give the same
discharge as the unpartitioned basin on one GPU, with either confluence exchange (NCCL all-gather / peer memory), and
must stay within the north star's 1e-5 of the bit-exact level sweep (the REAL4-per-cell reference arithmetic).

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 \
        tools/check_multigpu.py --rows 1500 --cols 1300 --steps 6
    ... tools/check_multigpu.py --bench-size 10000 --steps 4        # BASELINE config 4 at its full size

Every rank also runs the whole basin by itself (scan path and level sweep), so each comparison is local; rank 0 prints
one JSON line with the worst differences over all ranks and steps.
"""
import argparse
import datetime
import json
import os
import sys
import tempfile

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def main():
    import torch
    import torch.distributed as dist
    import bench
    from sphy_b200 import synth
    from sphy_b200.model import CatchmentSource, SphyModel
    ap = argparse.ArgumentParser()
    ap.add_argument("--rows", type=int, default=1500)
    ap.add_argument("--cols", type=int, default=1300)
    ap.add_argument("--steps", type=int, default=6)
    ap.add_argument("--bench-size", type=int, default=0, help="use bench.py's basin of this size and its device forcing instead")
    ap.add_argument("--glacier", action="store_true", help="snow and glacier modules on (BASELINE config 2 physics): the fragment "
                                                           "table is split by rank")
    ap.add_argument("--overlap", action="store_true", help="partitioned model with overlap_routing=True (routing on a second stream)")
    ap.add_argument("--reslake", action="store_true", help="reservoirs and lakes on (BASELINE config 3 physics): the structures sit on "
                                                           "the trunk list, their table is replicated on every rank")
    args = ap.parse_args()
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    local = int(os.environ.get("LOCAL_RANK", rank))
    torch.cuda.set_device(local)
    dev = torch.device(f"cuda:{local}")
    dist.init_process_group("nccl", device_id=dev)
    if args.bench_size:
        cat = bench.make_case(args.bench_size)
        comps = ("QRAold",)
    else:
        extra = dict(snow=True, glacier=True, zrange=(1500.0, 5200.0)) if args.glacier else {}
        if args.reslake:
            extra.update(reservoirs=True, lakes=True, n_res_simple=7, n_res_adv=5, n_lakes=4, cellsize=1000.0)
        par = dict(SubWater=470.0, RootDepthFlat=150.0, Pscale=2.5)
        if args.glacier:
            par.update(Toffset=15.0, RootKsat=12.0)
        cat = synth.make_catchment(nrows=args.rows, ncols=args.cols, seed=77, nsteps=args.steps, start=datetime.date(2001, 5, 1),
                                   rout_components=not (args.overlap or args.reslake), params=par, **extra)
        comps = ("QRAold",) if (args.overlap or args.reslake) else ("QRAold", "RootRRAold", "BaseRAold") + (("GlacRAold", "SnowRAold") if args.glacier else ())
    tmp = tempfile.mkdtemp(prefix=f"sphy_mgcheck_r{rank}_")
    cfg = cat.write_inputs(os.path.join(tmp, "in"), os.path.join(tmp, "out"))
    kw = dict(device=str(dev), collect_tss=False)
    whole = SphyModel(cfg, CatchmentSource(cat), route_method="scan", **kw)
    whole.initial()
    lev = SphyModel(cfg, CatchmentSource(cat), route_method="levels", share_from=whole, **kw)
    lev.initial()
    part = SphyModel(cfg, CatchmentSource(cat), route_method="scan", rank=rank, world=world, overlap_routing=args.overlap, **kw)
    part.initial()
    lo, hi = part.part_range
    if args.bench_size:
        dem = torch.from_numpy(whole._compact_np(cat.dem)).to(dev)
        ring = bench.device_forcing_ring(torch, dem, 2, seed=1234)
        del dem
        for m in (whole, lev):
            m.set_device_forcing(lambda t: ring[t % 2])
        ring_p = [{k: v[lo:hi].clone() for k, v in day.items()} for day in ring]
        part.set_device_forcing(lambda t: ring_p[t % 2])
    vs_whole = vs_levels = whole_vs_levels = stor_vs_levels = 0.0
    for t in range(args.steps):
        whole.dynamic()
        lev.dynamic()
        part.dynamic()
        part.sync()
        torch.cuda.synchronize()
        for k in comps:
            a = getattr(part, k).double()
            b = getattr(whole, k).double()[lo:hi]
            c = getattr(lev, k).double()[lo:hi]
            vs_whole = max(vs_whole, float(((a - b).abs() / b.abs().clamp_min(1e-12)).max().item()))
            vs_levels = max(vs_levels, float(((a - c).abs() / c.abs().clamp_min(1e-12)).max().item()))
            whole_vs_levels = max(whole_vs_levels, float(((b - c).abs() / c.abs().clamp_min(1e-12)).max().item()))
        assert torch.equal(part.RootWater, whole.RootWater[lo:hi]), "water balance must not depend on the split"
        if args.reslake and t % 2 == 1:                        # StorRES flushes the pending storage update (a collective)
            a, c = part.StorRES.double(), lev.StorRES.double()[lo:hi]
            stor_vs_levels = max(stor_vs_levels, float(((a - c).abs() / c.abs().clamp_min(1.0)).max().item()))
        if args.glacier:
            assert torch.equal(part.TotalSnowStore, whole.TotalSnowStore[lo:hi]) and torch.equal(part.GlacR, whole.GlacR[lo:hi])
    mode = "p2p" if part._mg_p2p else "nccl"
    w = torch.tensor([vs_whole, vs_levels, whole_vs_levels, stor_vs_levels], device=dev, dtype=torch.float64)
    dist.all_reduce(w, op=dist.ReduceOp.MAX)
    check = bench.discharge_check(torch, part, dist)
    check1 = bench.discharge_check(torch, whole)
    part.check_exchange()
    if rank == 0:
        print(json.dumps({"tool": "check_multigpu", "world": world, "exchange": mode, "cells": whole.n, "steps": args.steps, "glacier": bool(args.glacier), "overlap_routing": part._ovl is not None,
                          "reslake": bool(args.reslake), "structures": len(getattr(part, "_rl_cells_host", [])),
                          "worst_rel_storage_partitioned_vs_level_sweep": float(w[3]),
                          "trunk_cells_partitioned": part.trunk_cells, "trunk_cells_one_gpu": whole.trunk_cells,
                          "worst_rel_partitioned_vs_one_gpu": float(w[0]), "worst_rel_partitioned_vs_level_sweep": float(w[1]),
                          "worst_rel_one_gpu_scan_vs_level_sweep": float(w[2]), "check_partitioned": check, "check_one_gpu": check1}),
              flush=True)
    # plain routing: the split shows through the float64 prefixes only, a REAL4 ulp at most.  Lakes / reservoirs: the one-GPU
    # scan path places float64 corrections inside the main pass, the partitioned one corrects the listed cells afterwards
    assert float(w[0]) <= (3e-6 if args.reslake else 3e-7), float(w[0])
    assert float(w[3]) <= 1e-5, float(w[3])
    assert float(w[1]) <= 1e-5, float(w[1])          # north star
    part.close()
    lev.close()
    whole.close()
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
