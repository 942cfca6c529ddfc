"""Real multi-GPU check (run under torchrun, one rank per GPU): the basin split across the ranks must give the same
discharge as the unpartitioned basin on one GPU, with either confluence exchange (NCCL all-gather / peer memory).

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 \
        tools/check_multigpu.py --rows 1500 --cols 1300 --steps 6
"""
import argparse
import datetime
import os
import sys
import tempfile

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def main():
    import torch
    import torch.distributed as dist
    from sphy_b200 import synth
    from sphy_b200.model import CatchmentSource, SphyModel
    ap = argparse.ArgumentParser()
    ap.add_argument("--rows", type=int, default=1500)
    ap.add_argument("--cols", type=int, default=1300)
    ap.add_argument("--steps", type=int, default=6)
    args = ap.parse_args()
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    local = int(os.environ.get("LOCAL_RANK", rank))
    torch.cuda.set_device(local)
    dev = torch.device(f"cuda:{local}")
    dist.init_process_group("nccl", device_id=dev)
    cat = synth.make_catchment(nrows=args.rows, ncols=args.cols, seed=77, nsteps=args.steps, start=datetime.date(2001, 5, 1),
                               rout_components=True, params=dict(SubWater=470.0, RootDepthFlat=150.0, Pscale=2.5))
    tmp = tempfile.mkdtemp(prefix=f"sphy_mgcheck_r{rank}_")
    cfg = cat.write_inputs(os.path.join(tmp, "in"), os.path.join(tmp, "out"))
    whole = SphyModel(cfg, CatchmentSource(cat), device=str(dev), route_method="scan")
    whole.initial()
    part = SphyModel(cfg, CatchmentSource(cat), device=str(dev), route_method="scan", rank=rank, world=world)
    part.initial()
    lo, hi = part.part_range
    worst = 0.0
    for t in range(args.steps):
        whole.dynamic()
        part.dynamic()
        torch.cuda.synchronize()
        for k in ("QRAold", "RootRRAold", "BaseRAold"):
            a = getattr(part, k).cpu().numpy().astype(np.float64)
            b = getattr(whole, k).cpu().numpy().astype(np.float64)[lo:hi]
            err = np.max(np.abs(a - b) / np.maximum(np.abs(b), 1e-12))
            worst = max(worst, float(err))
        assert np.array_equal(part.RootWater.cpu().numpy(), whole.RootWater.cpu().numpy()[lo:hi]), "water balance must not depend on the split"
    mode = "p2p" if part._mg_p2p else "nccl"
    w = torch.tensor([worst], device=dev)
    dist.all_reduce(w, op=dist.ReduceOp.MAX)
    part.close()
    whole.close()
    if rank == 0:
        print(f"multi-GPU check: world={world} exchange={mode} cells={whole.n} worst relative difference {float(w.item()):.3e}", flush=True)
    assert float(w.item()) < 1e-6, float(w.item())
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
