"""`python -m sphy_b200 sphy_config.cfg` -- the run script of the reference (`python sphy.py sphy_config.cfg`,
/root/reference/sphy.py:1268-1282) on the B200 path: the same cfg file, the input directory it names (PCRaster CSF maps
and map stacks, or .npy rasters), outputs (reported maps, .tss files, glacier tables) into its output directory.

One process per GPU splits the basin across the GPUs of the box; rank 0 assembles and writes the outputs:

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 -m sphy_b200 sphy_config.cfg
"""
import os
import sys
import time


def main(argv=None):
    argv = list(sys.argv[1:] if argv is None else argv)
    # --route levels: the bit-exact level sweep; --route scan (default): range sums + the reference's REAL4 rounding drift on
    # the trunk cells (within 1e-5 of the level sweep at any size, 5-20x faster)
    route = "scan"
    if "--route" in argv:
        k = argv.index("--route")
        if k + 1 >= len(argv) or argv[k + 1] not in ("levels", "scan"):
            print("usage: python -m sphy_b200 [--route levels|scan] <sphy_config.cfg>", file=sys.stderr)
            return 2
        route = argv[k + 1]
        del argv[k:k + 2]
    if len(argv) != 1:
        print("usage: python -m sphy_b200 [--route levels|scan] <sphy_config.cfg>", file=sys.stderr)
        return 2
    import configparser

    from .model import DirMapSource, SphyModel
    tic = time.perf_counter()
    cfg = configparser.RawConfigParser()
    if not cfg.read(argv[0]):
        print(f"cannot read {argv[0]}", file=sys.stderr)
        return 2
    rank, world = int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1"))
    device = "cuda:0"
    if world > 1:                                # launched by torch.distributed.run: one rank per GPU
        import torch
        import torch.distributed as dist
        if route != "scan":
            print("a basin split across GPUs is routed on the scan path (--route scan)", file=sys.stderr)
            return 2
        local = int(os.environ.get("LOCAL_RANK", rank))
        torch.cuda.set_device(local)
        device = f"cuda:{local}"
        dist.init_process_group("nccl", device_id=torch.device(device))
    model = SphyModel(cfg, DirMapSource(cfg.get("DIRS", "inputdir")), reporting=True, write_outputs=True, route_method=route,
                      device=device, rank=rank, world=world)
    model.run()
    if model.report is not None:
        model.report.write_tss()                 # the reference moves its *.tss into outputdir at the end (sphy.py:1273-1278)
    steps = model.timestep
    model.close()
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    if rank == 0:
        print(f"Simulation succesfully completed in {time.perf_counter() - tic} seconds! ({steps} daily steps)")
    return 0


if __name__ == "__main__":
    sys.exit(main())
