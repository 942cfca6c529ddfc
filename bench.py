#!/usr/bin/env python
"""bench.py -- Mcell-timesteps/s of SPHY's per-timestep raster hot path on B200 (BASELINE.json metric).

    python bench.py --gpus 1 --steps K --warmup W            # this repo's CUDA path
    python bench.py --impl reference --steps K --warmup W    # the reference's CPU algorithm (oracle port), host cores

A "step" is one daily time step of the hot path (fused vertical water balance + flow routing) over one synthetic
catchment.  `value` has the forcing already resident in HBM; `e2e` copies each day's four forcing maps from pinned
host memory inside the timed region and reads the station discharges back.  One JSON line is printed by rank 0.
"""
import argparse
import datetime
import json
import math
import os
import subprocess
import sys
import tempfile
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

PARAMS = dict(SubWater=470.0)          # cfg defaults except a wet subsoil, so percolation/baseflow are active from day 1
ALG_BYTES_WB = 88                      # 4 * (n_forcing 4 + n_param_maps 1 + 2 * n_state 8 + n_out 1) -- what `achieved` is quoted on
STREAM_BYTES_WB = 90                   # what this instance streams: 4 forcing, 8 state in + 8 out, TotR, drain velocity, 2-byte latitude class
ALG_BYTES_SCAN = 16                    # k_scan_main per routed component: runoff, sub_size, Qold in, Q out


def kernel_traffic_per_cell(kernel="k_wb_step_pf"):
    """dram__bytes_read.sum + dram__bytes_write.sum per cell of one launch, from the committed ncu --set full capture of the
    benched kernels at the benched size (profiles/r2_kernel_traffic.json)."""
    try:
        return float(json.load(open(os.path.join(ROOT, "profiles", "r2_kernel_traffic.json")))[kernel]["dram_bytes_per_cell"])
    except Exception:
        return None


def k1_traffic_per_cell():
    return kernel_traffic_per_cell("k_wb_step_pf")


def measured_peak():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        try:
            return float(json.load(open(path))["hbm_gbs"]), "measured"
        except Exception:
            pass
    return 6650.0, "fallback"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled during the timed region (B200_PROFILING.md recipe)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index=0):
        self.index, self.rows, self.proc = index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "100",
                                          "-i", str(self.index)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._pump, daemon=True)
            self.thread.start()
        except OSError:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.rows.append([t.strip() for t in line.split(",")])

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        sm = [float(r[1]) for r in self.rows if len(r) >= 9 and r[1].replace(".", "").isdigit()]
        mx = [float(r[2]) for r in self.rows if len(r) >= 9 and r[2].replace(".", "").isdigit()]
        reasons = set()
        for r in self.rows:
            if len(r) >= 9:
                for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[5:9]):
                    if v.lower().startswith("active"):
                        reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


def make_case(size, seed=1, cellsize=250.0, nsteps=365):
    """The synthetic basin of BASELINE config 4 (and, at 2000 x 2000 / 1 km, the grid of configs 3 and 5)."""
    from sphy_b200 import synth
    return synth.make_catchment(size, size, cellsize=cellsize, seed=seed, nsteps=nsteps, outlets="single", params=PARAMS,
                                start=datetime.date(2001, 6, 1))


def _hash_uniform(torch, idx, key):
    """Counter-based uniform numbers in [0, 1): a splitmix64-style integer hash of (global cell index, key).  Every value
    depends only on the cell's GLOBAL device index, so 1, 2, 4 or 8 ranks generate bit-identical forcing for the same cell."""
    m = (1 << 63) - 1
    x = idx * -7046029254386353131 + (key * 0x632BE59BD9B4E019 & m)          # int64 arithmetic wraps
    x = (x ^ ((x >> 30) & ((1 << 34) - 1))) * -4658895280553007687
    x = (x ^ ((x >> 27) & ((1 << 37) - 1))) * -7723592293110705685
    x = x ^ ((x >> 31) & ((1 << 33) - 1))
    return ((x >> 40) & 0xFFFFFF).to(torch.float32) * (1.0 / 16777216.0)


def _hash_normal(torch, idx, key):
    """Approximately N(0, 1): sum of four hashed uniforms (Irwin-Hall), exact integer hashing + a few float32 adds."""
    s = _hash_uniform(torch, idx, 4 * key) + _hash_uniform(torch, idx, 4 * key + 1)
    s = s + _hash_uniform(torch, idx, 4 * key + 2) + _hash_uniform(torch, idx, 4 * key + 3)
    return (s - 2.0) * 1.7320508


def device_forcing_ring(torch, dem_dev, ring, seed, cell_begin=0):
    """Synthetic daily forcing generated directly in HBM (SURVEY 8d formulas; wet days every other day).  `cell_begin` is
    the global device index of this rank's first cell: the forcing of a cell does not depend on the partitioning."""
    days = []
    n = dem_dev.numel()
    idx = torch.arange(cell_begin, cell_begin + n, dtype=torch.int64, device=dem_dev.device)
    for d in range(ring):
        doy = 152 + d
        key = seed * 64 + d * 4
        season = 10.0 * math.sin(2.0 * math.pi * (doy - 105) / 365.0)
        tavg = (12.0 - 6.5e-3 * (dem_dev - 1000.0) + season + _hash_normal(torch, idx, key)).float()
        tmax = tavg + 5.0 + 2.0 * _hash_uniform(torch, idx, 16 * key + 1001)
        tmin = tavg - 5.0 - 2.0 * _hash_uniform(torch, idx, 16 * key + 1002)
        if d % 2 == 0:
            prec = (8.0 * torch.clamp(1.0 + 0.3 * _hash_normal(torch, idx, key + 3), min=0.0)).float()
        else:
            prec = torch.zeros(n, device=dem_dev.device)
        days.append(dict(prec=prec.contiguous(), tavg=tavg.contiguous(), tmin=tmin.contiguous(), tmax=tmax.contiguous()))
    del idx
    return days


def discharge_check(torch, model, dist=None):
    """What makes bench lines at different GPU counts comparable: discharge at the outlet (global cell 0 of the DFS order),
    the sum and the maximum of the discharge map, and the sum over the first eighth of the cells (rank 0's range on 8
    GPUs, inside rank 0's range for any smaller count), all in float64."""
    from sphy_b200 import multigpu
    q = model.QRAold.double()
    n_glob = model.n_global
    lo = model.part_range[0] if getattr(model, "part_range", None) else 0
    first = int(multigpu.partition_bounds(n_glob, 8)[1]) if n_glob >= 8 * multigpu.TILE else n_glob
    hi_local = max(0, min(first - lo, q.numel()))
    vals = torch.stack([q.sum(), q[:hi_local].sum(), q[0] if lo == 0 else q.new_zeros(())])
    mx = q.max().reshape(1)
    if dist is not None:
        dist.all_reduce(vals, op=dist.ReduceOp.SUM)
        dist.all_reduce(mx, op=dist.ReduceOp.MAX)
    v = vals.cpu().tolist()
    return {"outlet_q": v[2], "sum_q": v[0], "sum_q_first_eighth": v[1], "max_q": float(mx.item()), "timestep": int(model.timestep)}


def h2d_ceiling(torch, dev, dist=None, nbytes=1 << 30, reps=3):
    """Aggregate host->device bandwidth of this box for the e2e path's pattern: one pinned cudaMemcpyAsync per GPU, all
    GPUs at once (GB/s summed over the ranks)."""
    src = torch.empty(nbytes // 4, dtype=torch.float32).pin_memory()
    dst = torch.empty(nbytes // 4, dtype=torch.float32, device=dev)
    dst.copy_(src, non_blocking=True)
    torch.cuda.synchronize()
    if dist is not None:
        dist.barrier()
    t0 = time.perf_counter()
    for _ in range(reps):
        dst.copy_(src, non_blocking=True)
    torch.cuda.synchronize()
    dt = torch.tensor([time.perf_counter() - t0], device=dev, dtype=torch.float64)
    world = 1
    if dist is not None:
        dist.all_reduce(dt, op=dist.ReduceOp.MAX)
        world = dist.get_world_size()
    del src, dst
    return world * reps * nbytes / float(dt.item()) / 1e9


def cpu_port_rate(size, steps, seed=1):
    """Oracle port (NumPy restatement of the reference algorithm, 1 thread) on a bounded sample: Mcell-steps/s."""
    from oracle.sphy_port import PortInputs, PortModel
    cat = make_case(size, seed=seed)
    inp = PortInputs(cat)
    port = PortModel(inp)
    g = inp.lddst.gather
    forc = [{k: g(v) for k, v in cat.forcing(t).items()} for t in range(1, steps + 2)]
    port.dynamic(forc[0])                     # warm-up (first-touch, LDD build excluded like on the GPU side)
    t0 = time.perf_counter()
    for t in range(1, steps + 1):
        port.dynamic(forc[t])
    dt = time.perf_counter() - t0
    return inp.n * steps / dt / 1e6, dt


def _ref_worker(size, steps, seed, barrier, queue, warmup=1):
    """One host core: its own sub-catchment through the oracle port; the timed loops of all workers start together."""
    from oracle.sphy_port import PortInputs, PortModel
    cat = make_case(size, seed=seed)
    inp = PortInputs(cat)
    port = PortModel(inp)
    g = inp.lddst.gather
    forc = [{k: g(v) for k, v in cat.forcing(t).items()} for t in range(1, 5)]     # a ring of four days, like the GPU arm
    for t in range(warmup):
        port.dynamic(forc[t % 4])
    barrier.wait(timeout=900)
    t0 = time.time()
    for t in range(warmup, warmup + steps):
        port.dynamic(forc[t % 4])
    queue.put((inp.n * steps, t0, time.time()))


def cpu_port_rate_parallel(size, steps, procs, warmup=1):
    """The reference is single-threaded (Python + PCRaster); the way to use a whole host with it is one independent
    sub-catchment per core.  `procs` workers, each an oracle-port run of its own size x size basin; aggregate Mcell-steps/s
    over the window from the first loop start to the last loop end."""
    import multiprocessing as mp
    if procs <= 1:
        rate, dt = cpu_port_rate(size, steps)
        return rate, dt, 1
    ctx = mp.get_context("spawn")
    barrier, queue = ctx.Barrier(procs), ctx.Queue()
    ws = [ctx.Process(target=_ref_worker, args=(size, steps, 1 + i, barrier, queue, warmup)) for i in range(procs)]
    for w in ws:
        w.start()
    got, deadline = [], time.time() + 1800
    import queue as _queue
    while len(got) < procs:
        try:
            got.append(queue.get(timeout=2))
        except _queue.Empty:
            if time.time() > deadline or any(w.exitcode not in (None, 0) for w in ws):
                break                            # a worker died (out of memory?) or the box is far too slow: fall back below
    for w in ws:
        if len(got) < procs and w.is_alive():
            w.terminate()                        # our own children, by handle
        w.join()
    if len(got) < procs:
        rate, dt = cpu_port_rate(size, steps)
        return rate, dt, 1
    work = sum(g[0] for g in got)
    dt = max(g[2] for g in got) - min(g[1] for g in got)
    return work / dt / 1e6, dt, procs


def default_ref_procs():
    try:
        avail = len(os.sched_getaffinity(0))
    except AttributeError:
        avail = os.cpu_count() or 1
    return max(1, min(avail, 32))


def pcraster_probe():
    """SURVEY 8c, last row: if a real PCRaster build happens to be importable on this box, say so -- it would pin the
    accumulation arithmetic the oracle can only assume (A1).  It never is in this image; the answer is recorded either way."""
    try:
        import pcraster  # noqa: F401
        return {"importable": True, "version": getattr(pcraster, "__version__", "unknown")}
    except Exception as e:                                    # noqa: BLE001
        return {"importable": False, "why": type(e).__name__}


def run_reference(args):
    size = args.ref_size
    steps = max(1, args.steps)
    procs = args.ref_procs or default_ref_procs()
    warmup = max(1, args.warmup)
    rate, dt, procs = cpu_port_rate_parallel(size, steps, procs, warmup)
    sample = (f"{procs} independent {size}x{size} sub-catchments (one per host core), {warmup} warm-up + {steps} timed daily steps each, same physics switches "
              f"(oracle port, NumPy float32, one pass per operation)")
    line = {
        "impl": "reference", "metric": "Mcell-timesteps/s", "value": rate, "unit": "Mcell-timesteps/s", "n_gpus": args.gpus,
        "steps": steps, "warmup": warmup, "ms_per_step": dt / steps * 1e3, "higher_is_better": True, "scaling": "strong",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": workload_name(args.size), "sample": sample},
        "cpu_baseline": {"value": rate, "unit": "Mcell-timesteps/s", "cores": procs, "kind": "port", "sample": sample,
                         "host_cores_available": os.cpu_count(), "pcraster": pcraster_probe(),
                         "note": "PCRaster and SPHY are single-threaded, so every core runs its own sub-catchment; PCRaster is not "
                                 "installable here, so this is the NumPy-shim restatement of the reference algorithm, not PCRaster"},
        "e2e": {"value": rate, "unit": "Mcell-timesteps/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line))


def workload_name(size):
    return (f"synthetic {size}x{size} 250 m basin (single outlet), rootzone+subzone+groundwater+routing, snow/glacier off, "
            f"Hargreaves ET, uniform soils, total-flow routing (BASELINE config 4 physics)")


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--size", type=int, default=int(os.environ.get("SPHY_BENCH_SIZE", "10000")))
    ap.add_argument("--ref-size", type=int, default=2000)
    ap.add_argument("--ref-procs", type=int, default=0, help="host processes of the CPU arm (0 = one per available core, at most 32)")
    ap.add_argument("--route", default=os.environ.get("SPHY_BENCH_ROUTE", "scan"), choices=["levels", "scan"])
    ap.add_argument("--ring", type=int, default=4)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--overlap", action="store_true", help="run the routing tail of day t on a second stream underneath the water "
                                                           "balance of day t + 1 (default on a partitioned basin, where it hides the "
                                                           "exchange latency; on one GPU it gains 1-2 %% and blurs the per-kernel timing)")
    ap.add_argument("--no-overlap", action="store_true", help="route day t strictly before the water balance of day t + 1")
    ap.add_argument("--workload", default="basin", choices=["basin", "ensemble"],
                    help="basin: BASELINE config 4 (default, the headline); ensemble: BASELINE config 5")
    ap.add_argument("--members", type=int, default=64)
    ap.add_argument("--ens-size", type=int, default=2000)
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if args.impl == "reference":
        if rank == 0:
            run_reference(args)
        return
    if args.workload == "ensemble":
        from sphy_b200 import multigpu
        return multigpu.ensemble_main(args, rank, world)
    if world > 1:
        from sphy_b200 import multigpu
        return multigpu.bench_main(args, rank, world)

    import torch
    from sphy_b200.model import CatchmentSource, SphyModel
    assert torch.cuda.is_available(), "bench.py needs a GPU (no CPU fallback); use --impl reference for the CPU arm"
    dev = torch.device("cuda:0")
    torch.cuda.set_device(dev)
    t_setup = time.time()
    cat = make_case(args.size)
    t_synth = time.time() - t_setup
    tmp = tempfile.mkdtemp(prefix="sphy_bench_")
    cfg = cat.write_inputs(os.path.join(tmp, "in"), os.path.join(tmp, "out"))
    model = SphyModel(cfg, CatchmentSource(cat), device="cuda:0", route_method=args.route, diagnostics=False,
                      overlap_routing=args.overlap and not args.no_overlap)
    t_init = time.time()
    model.initial()
    torch.cuda.synchronize()
    setup_parts = dict(model.setup_times, synthetic_inputs_on_host_s=round(t_synth, 2), initial_s=round(time.time() - t_init, 2))
    n = model.n
    dem_dev = torch.from_numpy(model._compact_np(cat.dem)).to(dev)
    ring = device_forcing_ring(torch, dem_dev, args.ring, seed=1234)
    del dem_dev
    torch.cuda.synchronize()
    setup_s = time.time() - t_setup

    # ---- value: forcing resident in HBM ---------------------------------------------------------------
    model.set_device_forcing(lambda t: ring[t % len(ring)])
    for _ in range(max(args.warmup, 3)):
        model.dynamic()
    torch.cuda.synchronize()
    model.enable_profiling(True)
    model.lib.sphy_route_profile(model._ctx, 1)
    sampler = ClockSampler(0)
    sampler.start()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    profiled = os.environ.get("SPHY_BENCH_PROFILE") == "1"     # ncu --profile-from-start off: capture the timed steps only
    if profiled:
        torch.cuda.cudart().cudaProfilerStart()
    e0.record()
    for _ in range(args.steps):
        model.dynamic()
    model.sync()                                               # the routing stream's last step is inside the timed region
    e1.record()
    torch.cuda.synchronize()
    if profiled:
        torch.cuda.cudart().cudaProfilerStop()
    clocks = sampler.stop()
    ms_total = e0.elapsed_time(e1)
    ms_step = ms_total / args.steps
    wb_ms = float(np.mean([a.elapsed_time(b) for a, b in model._prof["wb"]]))
    rt_ms = float(np.mean([a.elapsed_time(b) for a, b in model._prof["route"]]))
    model.enable_profiling(False)
    value = n * args.steps / (ms_total * 1e-3) / 1e6
    peak, peak_kind = measured_peak()
    ach = ALG_BYTES_WB * n / (wb_ms * 1e-3) / 1e9
    check = discharge_check(torch, model)
    rprof = route_profile(model)

    # ---- e2e: host buffers, H2D of the day's forcing + D2H of station discharges inside the timed region ----
    # host side holds RASTERS (what readmap delivers); the model copies them H2D and compacts them into device order
    fa = torch.from_numpy(model.flat_active).to(dev)
    host_ring = []
    for day in ring[:2]:
        hd = {}
        for k, v in day.items():
            raster = torch.zeros(model.nrows * model.ncols, dtype=torch.float32, device=dev)
            raster[fa] = v
            hd[k] = raster.cpu().pin_memory()
            del raster
        host_ring.append(hd)
    del fa
    h2d_ceiling_gbs = h2d_ceiling(torch, dev)
    model.set_device_forcing(None)
    model.set_host_forcing(lambda t: host_ring[t % len(host_ring)])
    e2e_steps = max(3, min(args.steps, 10))
    for _ in range(2):
        model.dynamic()
    torch.cuda.synchronize()
    model.tss["QallRAtot"].clear()
    t0 = time.perf_counter()
    d2h = 0
    for _ in range(e2e_steps):
        model.dynamic()
        model.sync()
        q = model.tss["QallRAtot"][-1].cpu()          # device -> host read of the step's result (.tss sample)
        d2h += q.numel() * 4
    torch.cuda.synchronize()
    e2e_s = time.perf_counter() - t0
    e2e_val = n * e2e_steps / e2e_s / 1e6
    model.set_host_forcing(None)

    # ---- CPU baseline on a bounded sample -------------------------------------------------------------------
    cpu = None
    if not args.no_cpu_baseline:
        procs = args.ref_procs or default_ref_procs()
        rate, dt, procs = cpu_port_rate_parallel(args.ref_size, 3, procs)
        cpu = {"value": rate, "unit": "Mcell-timesteps/s", "cores": procs, "kind": "port",
               "sample": f"{procs} independent {args.ref_size}x{args.ref_size} sub-catchments (one per host core), 3 daily steps each, "
                         f"oracle port (NumPy float32, one pass per operation)",
               "host_cores_available": os.cpu_count(), "pcraster": pcraster_probe()}

    # my kernels per step: k_ra_table, k_wb_step, k_sample + routing (3 scan kernels, or one launch per plan segment)
    # k_ra_table, k_wb_step_pf, k_sample + routing.  scan: k_scan_main, tile-total scan (one kernel of this repo when the tail is
    # overlapped, else the library's two), k_scan_cross_pub, and k_trunk_drift + k_trunk_apply when there is a trunk list
    tile_scan = 1 if (model._ovl is not None or model.n < 4096 * 1024) else 0      # cub's kernels are not this repo's
    nlaunch_step = 3 + ((2 + tile_scan + (2 if model.trunk_cells else 0)) if args.route == "scan" else len(model_plan(model)))
    line = {
        "metric": "Mcell-timesteps/s", "value": value, "unit": "Mcell-timesteps/s", "n_gpus": 1, "steps": args.steps,
        "warmup": max(args.warmup, 3), "ms_per_step": ms_step, "higher_is_better": True, "scaling": "strong",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": workload_name(args.size), "cells": n, "levels": model.nlevels, "routing": args.route,
                   "trunk_cells": getattr(model, "trunk_cells", 0), "overlap_routing": model._ovl is not None,
                   "l2": "per-step working set %.1f GB >> 126 MB L2 (inputs larger than L2, no flush needed)" % (ALG_BYTES_WB * n / 1e9)
                   if ALG_BYTES_WB * n > 4 * 126e6 else "working set fits L2: throughput only, no roofline claim",
                   "forcing_ring_days": args.ring, "setup_s": round(setup_s, 1), "setup_parts": setup_parts},
        "roofline": {"kernel": "k_wb_step (fused vertical water balance)", "bound": "hbm", "achieved": ach, "peak": peak,
                     "peak_kind": peak_kind, "unit": "GB/s", "frac": ach / peak,
                     "traffic": (k1_traffic_per_cell() * n if k1_traffic_per_cell() else None),
                     "traffic_note": "bytes per launch = ncu dram bytes per cell (profiles/r2_kernel_traffic.json, 1e8-cell capture) x cells",
                     "alg_bytes_per_cell": ALG_BYTES_WB, "alg_bytes_per_launch": ALG_BYTES_WB * n, "kernel_ms": wb_ms, "route_ms": rt_ms,
                     "streamed_bytes_per_cell": STREAM_BYTES_WB, "achieved_on_streamed_bytes": STREAM_BYTES_WB * n / (wb_ms * 1e-3) / 1e9},
        "roofline_route": scan_roofline(rprof, n, 1, peak),
        "check": check,
        "cpu_baseline": cpu,
        "e2e": {"value": e2e_val, "unit": "Mcell-timesteps/s", "h2d_bytes_per_step": 4 * model.h2d_bytes_per_map,
                "d2h_bytes_per_step": d2h // e2e_steps, "steps": e2e_steps,
                "h2d_achieved_gbs": 4 * model.h2d_bytes_per_map * e2e_steps / e2e_s / 1e9, "h2d_ceiling_gbs": h2d_ceiling_gbs,
                "note": "ceiling = one pinned cudaMemcpyAsync of 1 GiB, measured in this run"},
        "gpu_launches": nlaunch_step * args.steps,
        "clocks": clocks,
    }
    print(json.dumps(line))
    model.close()


def route_profile(model):
    """Average per-step device time of the scan routing's phases (CUDA events recorded by the library on the launching
    stream): main pass, tile-crossing cells, trunk cells."""
    import ctypes as C
    if model.route_method != 1:
        return None
    out = (C.c_float * 4)()
    cnt = C.c_int(0)
    model.lib.sphy_route_profile_read(model._ctx, out, C.byref(cnt))
    return {"scan_main_ms": out[0], "totals_cross_ms": out[1], "trunk_ms": out[2], "steps": cnt.value} if cnt.value else None


def scan_roofline(rprof, n, ncomp, peak):
    if not rprof:
        return None
    ach = ALG_BYTES_SCAN * ncomp * n / (rprof["scan_main_ms"] * 1e-3) / 1e9
    t = kernel_traffic_per_cell("k_scan_main")
    return dict(rprof, kernel="k_scan_main (float64 tile prefix + range sums + recession blend)", bound="hbm", achieved=ach, peak=peak,
                unit="GB/s", frac=ach / peak, alg_bytes_per_cell=ALG_BYTES_SCAN * ncomp, traffic=t * n * ncomp if t else None)


def model_plan(model):
    """Number of level-sweep launches per step, from the level histogram: one per level wider than 65 536 cells, the
    cooperative launch of the middle levels, the short-chain and the long-chain kernel."""
    import torch
    lp = model.ldd_structure("level_ptr").cpu().numpy()
    sizes = np.diff(lp)
    launches = int((sizes > 65536).sum()) + 3
    torch.cuda.synchronize()
    return [0] * launches


if __name__ == "__main__":
    main()
