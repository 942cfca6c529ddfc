"""`python -m sphy_b200 sphy_config.cfg` -- the run script of the reference (`python sphy.py sphy_config.cfg`,
/root/reference/sphy.py:1268-1282) on the B200 path: the same cfg file, the input directory it names (PCRaster CSF maps
and map stacks, or .npy rasters), outputs (reported maps, .tss files, glacier tables) into its output directory."""
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
    model = SphyModel(cfg, DirMapSource(cfg.get("DIRS", "inputdir")), reporting=True, write_outputs=True, route_method=route)
    model.run()
    if model.report is not None:
        model.report.write_tss()                 # the reference moves its *.tss into outputdir at the end (sphy.py:1273-1278)
    steps = model.timestep
    model.close()
    print(f"Simulation succesfully completed in {time.perf_counter() - tic} seconds! ({steps} daily steps)")
    return 0


if __name__ == "__main__":
    sys.exit(main())
