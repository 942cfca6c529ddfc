#!/usr/bin/env python
"""Throughput of the five BASELINE.json configurations on one GPU (configs 1-3, and the per-GPU share of 4-5 are in
bench.py).  Prints one JSON line per configuration: Mcell-timesteps/s with the forcing resident in HBM, CUDA events.

    python tools/bench_configs.py [--steps 365]
"""
import argparse
import datetime
import json
import os
import sys
import tempfile

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def run(name, kw, steps, route):
    import torch
    import bench
    from sphy_b200 import synth
    from sphy_b200.model import CatchmentSource, SphyModel
    cat = synth.make_catchment(nsteps=steps, start=datetime.date(2001, 1, 1), **kw)
    tmp = tempfile.mkdtemp(prefix="sphy_cfg_")
    cfg = cat.write_inputs(os.path.join(tmp, "in"), os.path.join(tmp, "out"))
    model = SphyModel(cfg, CatchmentSource(cat), route_method=route, diagnostics=False)
    model.initial()
    if getattr(model, "_rl_part", False):
        name += " (structures on the trunk list)"
    dem = torch.from_numpy(np.ascontiguousarray(model._compact_np(cat.dem))).cuda()
    ring = bench.device_forcing_ring(torch, dem, 8, seed=7)
    model.set_device_forcing(lambda t: ring[t % len(ring)])
    for _ in range(2 * len(ring) + 2):           # with SPHY_GRAPHS=1 every ring slot is seen twice before its graph exists
        model.dynamic()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    profiled = os.environ.get("SPHY_BENCH_PROFILE") == "1"     # ncu --profile-from-start off: capture the timed steps only
    if profiled:
        torch.cuda.cudart().cudaProfilerStart()
    e0.record()
    for _ in range(steps):
        model.dynamic()
    e1.record()
    torch.cuda.synchronize()
    if profiled:
        torch.cuda.cudart().cudaProfilerStop()
    ms = e0.elapsed_time(e1)
    print(json.dumps({"config": name, "cells": model.n, "levels": model.nlevels, "steps": steps, "routing": route,
                      "ms_per_step": ms / steps, "Mcell_timesteps_per_s": model.n * steps / (ms * 1e-3) / 1e6}), flush=True)
    model.close()


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--steps", type=int, default=365)
    ap.add_argument("--configs", default="C1,C2,C3", help="comma-separated subset of C1,C2,C3,C3x")
    ap.add_argument("--routes", default="scan,levels")
    a = ap.parse_args()
    wet = dict(SubWater=470.0)
    want = set(a.configs.split(","))
    for route in a.routes.split(","):
        if "C1" in want:
            run("C1 500x500 ground+routing", dict(nrows=500, ncols=500, params=wet), a.steps, route)
        if "C2" in want:
            run("C2 500x500 + snow + glacier", dict(nrows=500, ncols=500, snow=True, glacier=True, params=wet), a.steps, route)
        if "C3" in want:
            run("C3 2000x2000 1 km, 34 reservoirs + 6 lakes", dict(nrows=2000, ncols=2000, cellsize=1000.0, reservoirs=True, lakes=True, n_res_simple=24, n_res_adv=10, n_lakes=6, params=wet),
                min(a.steps, 100), route)
        if "C3x" in want:      # config-3 physics on the basin of config 4: what the structure path costs at 1e8 cells
            run("C3x 10000x10000 250 m, 34 reservoirs + 6 lakes", dict(nrows=10000, ncols=10000, cellsize=250.0, reservoirs=True, lakes=True,
                                                                      n_res_simple=24, n_res_adv=10, n_lakes=6, params=wet),
                min(a.steps, 20), route)
