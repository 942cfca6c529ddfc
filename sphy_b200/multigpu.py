"""Multi-GPU: one large basin split along its drainage tree across the GPUs of one box.

In the DFS device order the upstream area of every cell is a contiguous index range, so the basin is cut into P
contiguous, tile-aligned index ranges (whole sub-basins plus pieces of the trunk), one per rank.  The vertical water
balance is per-cell: no communication.  Routing: every rank scans its own range; the only cells that need remote
data are the trunk cells whose upstream range ends on another rank (a few thousand out of 1e8).  Per step each rank
publishes (i) the total of its range and (ii) the range-local prefix at the few positions other ranks' ranges end
at; one small all-gather over NCCL/NVLink delivers these "confluence inflows" to everybody, and a sparse kernel
finishes the trunk cells.  The reference has no counterpart (it is a single process).

`plan()` is pure host logic (NumPy) and is what the CPU/gloo tests exercise.
"""
from __future__ import annotations

import json
import os
import tempfile
import time

import numpy as np

from . import _lib

TILE = _lib.SCAN_TILE


def partition_bounds(n, world, tile=TILE):
    """P+1 cell boundaries: contiguous ranges of whole tiles, balanced by cell count."""
    ntiles = -(-n // tile)
    tb = np.round(np.arange(world + 1) * ntiles / world).astype(np.int64)
    b = np.minimum(tb * tile, n)
    b[-1] = n
    if np.any(np.diff(b) <= 0):
        raise ValueError(f"{n} cells are too few for {world} ranks of {tile}-cell tiles")
    return b


# default thresholds of the trunk list (sphy_trunk_select): every cell with at least this many upstream levels / cells
# carries the reference's REAL4 rounding drift.  Both must exceed the scan tile.
TRUNK_MIN_LEVEL = 1024
TRUNK_MIN_CELLS = 1 << 18


def plan_trunk(cell, sub_size, bounds):
    """Exchange plan of the trunk list (pure host logic over the small exported arrays).

    cell      int  global device indices of the listed cells, ascending (sphy_trunk_export "cell")
    sub_size  int  their upstream cell counts                          (sphy_trunk_export "sub_size")
    bounds    int  world + 1 partition bounds (partition_bounds), [0, n] for one GPU

    A listed cell j is finished from two global prefixes of the routed material: the one just before it (position
    j - 1, none for j = 0) and the one at the end of its upstream range (position j + sub_size - 1).  Every rank
    publishes the range-local prefix at the positions that fall into its range.  Returns a dict with
      pub          list of int32 arrays: per rank the published positions as LOCAL indices, ascending and unique
      before_rank  int32 [m]  rank publishing the prefix before entry s (-1: none)
      before_slot  int32 [m]  its slot in that rank's `pub`
      end_rank / end_slot     the same for the end of the upstream range
      stride       1 + the longest `pub` (slot 0 of a rank's block carries the range total)
    """
    cell = np.asarray(cell, np.int64)
    sub_size = np.asarray(sub_size, np.int64)
    bounds = np.asarray(bounds, np.int64)
    world = bounds.size - 1
    before, end = cell - 1, cell + sub_size - 1
    has_b = before >= 0
    pos = np.unique(np.concatenate([before[has_b], end]))
    owner = np.searchsorted(bounds, pos, side="right") - 1
    first = np.searchsorted(pos, bounds[:-1])                  # start of each rank's segment of `pos`
    seg_end = np.append(first[1:], pos.size)
    pub = [(pos[first[q]:seg_end[q]] - bounds[q]).astype(np.int32) for q in range(world)]

    def locate(p):
        idx = np.searchsorted(pos, p)
        rk = owner[idx]
        return rk.astype(np.int32), (idx - first[rk]).astype(np.int32)

    er, es = locate(end) if cell.size else (np.zeros(0, np.int32), np.zeros(0, np.int32))
    br = np.full(cell.size, -1, np.int32)
    bs = np.zeros(cell.size, np.int32)
    if has_b.any():
        br[has_b], bs[has_b] = locate(before[has_b])
    return dict(pub=pub, before_rank=br, before_slot=bs, end_rank=er, end_slot=es,
                stride=1 + max(len(p) for p in pub))


RL_EXTRA = _lib.RL_EXTRA  # per structure: the day's discharge [m3/day] at its <= 8 donors, then TotR at the structure cell


def plan_reslake(cell, sub_size, cut_cell, donors, bounds):
    """Lakes / reservoirs on the listed cells (pure host logic; the device side is sphy_route_reslake_begin / _finish).

    A lake or reservoir cell has QFRAC = 0 (modules/lakes.py:224): `accufractionflux` passes nothing of what arrives there
    downstream, and `advanced_routing.dynamic` puts the structure's outflow on the cell below it
    (modules/advanced_routing.py:159-161).  With E(c) the plain range sum of the per-cell material over the upstream range
    of c (structure cells carry none), the flux of any cell is
        F(c) = E(c) - sum of (E(k) - Qout(k)) over the TOP-LEVEL structures k strictly inside the range of c
    (top-level: no other structure between k and c -- whatever lies above k, nested structures included, stays behind k).
    Only cells downstream of a structure have a non-empty sum, so they are put on the trunk list (sphy_trunk_select_cuts)
    and finished from published prefixes; every other cell keeps the plain scan.  This function builds the static tables
    of that correction.

    cell, sub_size  the listed cells (global device indices, ascending) and their upstream counts; the list must contain
                    every structure cell and every cell downstream of one
    cut_cell        [ns] structure cells, global device indices, ascending
    donors          [ns, 8] the cells draining directly into each structure cell (global device indices, -1 padded)
    bounds          world + 1 partition bounds ([0, n] on one GPU)
    Returns int32 arrays:
      cut_slot   [ns]     list slot of each structure cell
      entry_cut  [m]      structure index of a listed structure cell, -1 elsewhere
      entry_ptr  [m + 1], entry_idx   CSR: the top-level structures strictly inside each listed cell's upstream range
      x_rank, x_local  [ns * RL_EXTRA]   owner rank / local index of donor q (slots 0..7) and of the structure cell itself
                                         (slot 8) of every structure; rank -1 = no such donor
    """
    cell = np.asarray(cell, np.int64)
    end = cell + np.asarray(sub_size, np.int64) - 1
    cut = np.asarray(cut_cell, np.int64)
    donors = np.asarray(donors, np.int64).reshape(-1, 8) if len(cut) else np.zeros((0, 8), np.int64)
    bounds = np.asarray(bounds, np.int64)
    ns, m = cut.size, cell.size
    if ns and np.any(np.diff(cut) <= 0):
        raise ValueError("structure cells must be listed by ascending device index")
    cut_slot = np.searchsorted(cell, cut)
    if ns and (np.any(cut_slot >= m) or np.any(cell[np.minimum(cut_slot, m - 1)] != cut)):
        raise ValueError("every structure cell must be on the trunk list")
    entry_cut = np.full(m, -1, np.int32)
    entry_cut[cut_slot] = np.arange(ns, dtype=np.int32)
    cut_end = end[cut_slot] if ns else np.zeros(0, np.int64)
    # nearest structure downstream of each structure (pre-order: an ancestor has a smaller index and a range that contains it)
    parent = np.full(ns, -1, np.int64)
    stack = []
    for k in range(ns):
        while stack and cut_end[stack[-1]] < cut[k]:
            stack.pop()
        if stack:
            parent[k] = stack[-1]
        stack.append(k)
    # structure k is top-level inside the range of entry s  <=>  cell_s < cut_k <= end_s  and  its parent structure (if
    # any) is not strictly inside that range, i.e. cut_parent <= cell_s
    pairs_s, pairs_k = [], []
    for k in range(ns):
        lo = np.searchsorted(cell, cut[parent[k]]) if parent[k] >= 0 else 0
        hi = cut_slot[k]                                   # entries with cell < cut_k
        s = np.arange(lo, hi)
        s = s[end[s] >= cut[k]]
        pairs_s.append(s)
        pairs_k.append(np.full(s.size, k, np.int64))
    ps = np.concatenate(pairs_s) if pairs_s else np.zeros(0, np.int64)
    pk = np.concatenate(pairs_k) if pairs_k else np.zeros(0, np.int64)
    o = np.lexsort((pk, ps))
    ps, pk = ps[o], pk[o]
    entry_ptr = np.zeros(m + 1, np.int64)
    np.add.at(entry_ptr, ps + 1, 1)
    entry_ptr = np.cumsum(entry_ptr)
    # where the per-structure extras of the exchange block come from
    x_cell = np.concatenate([donors, cut[:, None]], axis=1).reshape(-1) if ns else np.zeros(0, np.int64)
    has = x_cell >= 0
    x_rank = np.full(x_cell.size, -1, np.int64)
    x_rank[has] = np.searchsorted(bounds, x_cell[has], side="right") - 1
    x_local = np.zeros(x_cell.size, np.int64)
    x_local[has] = x_cell[has] - bounds[x_rank[has]]
    i32 = lambda a: np.ascontiguousarray(a, np.int32)  # noqa: E731
    return dict(cut_slot=i32(cut_slot), entry_cut=entry_cut, entry_ptr=i32(entry_ptr), entry_idx=i32(pk),
                x_rank=i32(x_rank), x_local=i32(x_local), n_struct=int(ns))


# ----------------------------------------------------------------------------------------------
def bind_to_gpu_numa_node(local):
    """Pin this process to the CPUs next to its GPU (NVML's CPU affinity of the device) before any pinned host buffer is
    allocated: first-touch then places the staging memory on the GPU's own NUMA node, which is what the per-rank
    host->device uploads of the end-to-end path run from.  Best effort: any failure leaves the affinity untouched."""
    try:
        import pynvml
        import torch
        pynvml.nvmlInit()
        try:
            uuid = str(torch.cuda.get_device_properties(local).uuid)
            handle = pynvml.nvmlDeviceGetHandleByUUID(("GPU-" + uuid) if not uuid.startswith("GPU-") else uuid)
        except Exception:
            handle = pynvml.nvmlDeviceGetHandleByIndex(local)
        words = (os.cpu_count() + 63) // 64
        mask = pynvml.nvmlDeviceGetCpuAffinity(handle, words)
        cpus = [w * 64 + b for w, m in enumerate(mask) for b in range(64) if (int(m) >> b) & 1]
        allowed = os.sched_getaffinity(0)
        cpus = [c for c in cpus if c in allowed]
        if cpus:
            os.sched_setaffinity(0, cpus)
        return len(cpus)
    except Exception:
        return 0


def bench_main(args, rank, world):
    """bench.py for N > 1: the BASELINE config-4 basin split across `world` GPUs (strong scaling)."""
    import torch
    import torch.distributed as dist

    import bench
    from .model import CatchmentSource, SphyModel

    local = int(os.environ.get("LOCAL_RANK", rank))
    torch.cuda.set_device(local)
    dev = torch.device(f"cuda:{local}")
    numa_cpus = bind_to_gpu_numa_node(local)
    dist.init_process_group("nccl", device_id=dev)
    t_setup = time.time()
    cat = bench.make_case(args.size)
    t_synth = time.time() - t_setup
    tmp = tempfile.mkdtemp(prefix=f"sphy_bench_r{rank}_")
    cfg = cat.write_inputs(os.path.join(tmp, "in"), os.path.join(tmp, "out"))
    model = SphyModel(cfg, CatchmentSource(cat), device=str(dev), route_method="scan", diagnostics=False, rank=rank, world=world,
                      overlap_routing=not args.no_overlap)
    t_init = time.time()
    model.initial()
    torch.cuda.synchronize()
    setup_parts = dict(model.setup_times, synthetic_inputs_on_host_s=round(t_synth, 2), initial_s=round(time.time() - t_init, 2))
    n_local, n_global = model.n, model.n_global
    dem_dev = torch.from_numpy(model._compact_np(cat.dem)).to(dev)
    ring = bench.device_forcing_ring(torch, dem_dev, args.ring, seed=1234, cell_begin=model.part_range[0])
    del dem_dev
    torch.cuda.synchronize()
    dist.barrier()
    setup_s = time.time() - t_setup

    def timed(nsteps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        dist.barrier()
        torch.cuda.synchronize()
        e0.record()
        for _ in range(nsteps):
            model.dynamic()
        model.sync()                                           # the routing stream's last step is inside the timed region
        e1.record()
        torch.cuda.synchronize()
        ms = torch.tensor([e0.elapsed_time(e1)], device=dev)
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)          # max over ranks
        dist.barrier()
        return float(ms.item())

    model.set_device_forcing(lambda t: ring[t % len(ring)])
    for _ in range(max(args.warmup, 3)):
        model.dynamic()
    model.enable_profiling(True)
    model.lib.sphy_route_profile(model._ctx, 1)
    sampler = bench.ClockSampler(local) if rank == 0 else None
    if sampler:
        sampler.start()
    ms_total = timed(args.steps)
    clocks = sampler.stop() if sampler else None
    wb_ms = float(np.mean([a.elapsed_time(b) for a, b in model._prof["wb"]]))
    rt_ms = float(np.mean([a.elapsed_time(b) for a, b in model._prof["route"]]))
    model.enable_profiling(False)
    value = n_global * args.steps / (ms_total * 1e-3) / 1e6
    check = bench.discharge_check(torch, model, dist)
    rprof = bench.route_profile(model)

    # e2e: the host holds forcing RASTERS (what readmap delivers) in pinned memory; every rank pulls its own cells out of
    # them each step (sphy_gather_indexed_f32) and reads its stations back -- the same raster -> device path as on one GPU
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
    h2d_ceiling = bench.h2d_ceiling(torch, dev, dist)
    model.set_device_forcing(None)
    model.set_host_forcing(lambda t: host_ring[t % len(host_ring)])
    for _ in range(2):
        model.dynamic()
    e2e_steps = max(3, min(args.steps, 10))
    dist.barrier()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    d2h = 0
    for _ in range(e2e_steps):
        model.dynamic()
        if model.tss["QallRAtot"]:
            model.sync()
            q = model.tss["QallRAtot"][-1].cpu()
            d2h += q.numel() * 4
    torch.cuda.synchronize()
    dt = torch.tensor([time.perf_counter() - t0], device=dev)
    dist.all_reduce(dt, op=dist.ReduceOp.MAX)
    e2e_val = n_global * e2e_steps / float(dt.item()) / 1e6
    h2d = torch.tensor([4.0 * model.h2d_bytes_per_map], device=dev, dtype=torch.float64)
    dist.all_reduce(h2d, op=dist.ReduceOp.SUM)
    h2d_bytes = float(h2d.item())
    peak, peak_kind = bench.measured_peak()
    ach = bench.ALG_BYTES_WB * n_local / (wb_ms * 1e-3) / 1e9
    if rank == 0:
        line = {
            "metric": "Mcell-timesteps/s", "value": value, "unit": "Mcell-timesteps/s", "n_gpus": world, "steps": args.steps,
            "warmup": max(args.warmup, 3), "ms_per_step": ms_total / args.steps, "higher_is_better": True, "scaling": "strong",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": bench.workload_name(args.size), "cells": n_global, "cells_per_gpu": n_local,
                       "partition": f"{world} contiguous DFS-order ranges (sub-basins + trunk pieces), per-step exchange of "
                                    f"{model._mg_stride} float64 per rank (confluence inflows; "
                                    + ("stores into peer memory over NVLink + flags" if getattr(model, "_mg_p2p", False)
                                       else "NCCL all-gather") + ")",
                       "routing": "scan", "trunk_cells": model.trunk_cells, "overlap_routing": model._ovl is not None, "forcing_ring_days": args.ring, "setup_s": round(setup_s, 1),
                       "setup_parts": setup_parts,
                       "host_affinity": f"rank 0 bound to the {numa_cpus} CPUs NVML lists next to its GPU" if numa_cpus else "unchanged",
                       "l2": "per-GPU per-step working set %.2f GB (larger than the 126 MB L2)" % (bench.ALG_BYTES_WB * n_local / 1e9)},
            "roofline": {"kernel": "k_wb_step (fused vertical water balance), rank 0", "bound": "hbm", "achieved": ach, "peak": peak,
                         "peak_kind": peak_kind, "unit": "GB/s", "frac": ach / peak,
                         "traffic": (bench.k1_traffic_per_cell() * n_local if bench.k1_traffic_per_cell() else None),
                         "alg_bytes_per_cell": bench.ALG_BYTES_WB, "kernel_ms": wb_ms, "route_ms": rt_ms,
                         "streamed_bytes_per_cell": bench.STREAM_BYTES_WB},
            "roofline_route": bench.scan_roofline(rprof, n_local, 1, peak),
            "check": check,
            "cpu_baseline": None,
            "e2e": {"value": e2e_val, "unit": "Mcell-timesteps/s", "h2d_bytes_per_step": int(h2d_bytes),
                    "d2h_bytes_per_step": d2h // e2e_steps, "steps": e2e_steps,
                    "h2d_achieved_gbs": h2d_bytes * e2e_steps / float(dt.item()) / 1e9, "h2d_ceiling_gbs": h2d_ceiling,
                    "note": "every rank pulls its own cells out of pinned host rasters (all ranks together: "
                            f"{h2d_bytes / (16.0 * model.nrows * model.ncols):.2f} x the raster, algorithmic bytes); ceiling = "
                            "concurrent pinned cudaMemcpyAsync of 1 GiB per GPU"},
            "gpu_launches": 8 * args.steps * world,     # k_ra_table, K1, k_scan_main, tile scan, cross+publish, drift, apply, k_sample
            "clocks": clocks,
        }
        print(json.dumps(line), flush=True)
    model.close()
    dist.barrier()
    dist.destroy_process_group()


# ----------------------------------------------------------------------------------------------
class Ensemble:
    """The members of a forcing/parameter ensemble that live on one GPU (BASELINE config 5): independent SphyModels sharing
    one context, i.e. one set of LDD structures.  One ensemble step runs every member's fused water-balance kernel and
    then routes the members in groups of up to 8 maps per pass (sphy_route_step_members): sub_size, the crossing and the
    trunk lists are read once per group instead of once per member.  With `graphs` the whole ensemble step -- a few hundred
    launches -- is replayed from one CUDA graph, which removes the per-member Python / ctypes / launch chain that
    otherwise bounds this workload."""

    GROUP = _lib.ROUTE_MAX_COMP

    def __init__(self, members, graphs=True):
        from .model import StepGraph
        if not members:
            raise ValueError("an ensemble needs at least one member")
        lead = members[0]
        if any(m._ctx is not lead._ctx for m in members):
            raise ValueError("ensemble members must share one context (share_from=)")
        self.members = list(members)
        # members route together when they use plain total-flow routing on the scan path with scalar or per-cell kx alike
        self.grouped = all(m.RoutFLAG == 1 and m.ResFLAG != 1 and m.LakeFLAG != 1 and m.route_method == _lib.ROUTE_SCAN
                           and m.world == 1 and not any(m.RA_FLAG.values()) and bool(m._kx.map) == bool(lead._kx.map)
                           for m in members)
        # members that share the latitude classes (same ETref mode) take the Ra table of the lead member
        self.shared_ra = all(m.ETREF_FLAG == lead.ETREF_FLAG and m.DynVegFLAG != 1 and
                             (m.ETREF_FLAG == 1 or (m._wbP.ra_mode == _lib.RA_TABLE and np.array_equal(m.Lat, lead.Lat)))
                             for m in members)
        self.streams = None
        self.state = StepGraph.state(lead) if graphs and lead._graphs is None else lead._graphs
        if graphs and self.state is None:
            raise ValueError("this configuration cannot be replayed from a graph (per-cell latitude maps or dynamic vegetation)")

    def dynamic(self):
        from .model import StepGraph
        pre = [m._pre_step() for m in self.members]
        doy = pre[0][1]
        if any(p[1] != doy for p in pre):
            raise ValueError("ensemble members must be on the same date")
        forcings = [p[0] for p in pre]
        if self.state is not None:
            StepGraph.run(self.state, self.members, forcings, doy, capture=all(m._graph_ok() for m in self.members),
                          step_fn=self._device_step)
        else:
            self._device_step(self.members, forcings, doy)
        return [m._post_step() for m in self.members]

    def _device_step(self, models, forcings, doy):
        import ctypes as C

        import torch
        lead = models[0]
        # the members' water-balance kernels are independent: spread them over a few streams so that the drain of one
        # overlaps the ramp-up of the next (inside a captured graph these become parallel branches)
        if self.streams is None:
            self.streams = [torch.cuda.Stream(device=lead.device) for _ in range(3)]
        main = torch.cuda.current_stream(lead.device)
        lead._device_step(forcings[0], doy, route=not self.grouped)
        ready = torch.cuda.Event()
        ready.record(main)                                     # the Ra table of the day is in place
        for st in self.streams:
            st.wait_event(ready)
        for j, (m, f) in enumerate(zip(models[1:], forcings[1:])):
            with torch.cuda.stream(self.streams[j % len(self.streams)] if self.grouped else main):
                m._device_step(f, doy, route=not self.grouped, ra_table=not self.shared_ra)
        for st in self.streams:
            main.wait_stream(st)
        if not self.grouped:
            return
        for g0 in range(0, len(models), self.GROUP):
            grp = models[g0:g0 + self.GROUP]
            k = len(grp)
            src = (C.c_void_p * k)(*[m.TotR.data_ptr() for m in grp])
            dst = (C.c_void_p * k)(*[m.QRAold.data_ptr() for m in grp])
            kx = (_lib.SphyParam * k)(*[m._kx for m in grp])
            omk = (C.c_float * k)(*[m._one_m_kx for m in grp])
            lead._ck(lead.lib.sphy_route_step_members(lead._ctx, k, src, dst, kx, omk, lead._stream()), "sphy_route_step_members")


# ----------------------------------------------------------------------------------------------
def ensemble_main(args, rank, world):
    """BASELINE config 5: a forcing/parameter ensemble on the 2000 x 2000 basin, members sharded across the GPUs.
    Members are independent: no data-path collective (only the timing barrier / max-over-ranks)."""
    import copy

    import torch

    import bench
    from .model import CatchmentSource, SphyModel

    dist = None
    local = int(os.environ.get("LOCAL_RANK", rank))
    torch.cuda.set_device(local)
    dev = torch.device(f"cuda:{local}")
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=dev)
    total = args.members
    mine = [m for m in range(total) if m % world == rank]
    size = args.ens_size
    cat = bench.make_case(size, cellsize=1000.0)
    members, rings = [], []
    t_setup = time.time()
    for m in mine:
        c = copy.copy(cat)
        rng = np.random.default_rng(1000 + m)
        c.params = dict(cat.params)
        for k in ("kx", "alphaGw", "deltaGw"):                # +-10 % parameter perturbation
            c.params[k] = cat.params[k] * (1.0 + 0.1 * (2.0 * rng.random() - 1.0))
        c.params["kx"] = min(c.params["kx"], 0.999)
        tmp = tempfile.mkdtemp(prefix=f"sphy_ens_r{rank}_m{m}_")
        cfg = c.write_inputs(os.path.join(tmp, "in"), os.path.join(tmp, "out"))
        model = SphyModel(cfg, CatchmentSource(c), device=str(dev), route_method="scan", diagnostics=False, collect_tss=False,
                          share_from=members[0] if members else None)
        model.initial()
        members.append(model)
        dem_dev = torch.from_numpy(np.ascontiguousarray(model._compact_np(cat.dem))).to(dev)
        ring = bench.device_forcing_ring(torch, dem_dev, 2, seed=5000 + m)
        rings.append(ring)
        model.set_device_forcing(lambda t, ring=ring: ring[t % len(ring)])
    n = members[0].n
    torch.cuda.synchronize()
    setup_s = time.time() - t_setup

    ens = Ensemble(members, graphs=os.environ.get("SPHY_ENSEMBLE_GRAPHS", "1") == "1")

    def step_all():
        ens.dynamic()

    for _ in range(max(args.warmup, 3)):
        step_all()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    if dist:
        dist.barrier()
    torch.cuda.synchronize()
    e0.record()
    for _ in range(args.steps):
        step_all()
    e1.record()
    torch.cuda.synchronize()
    ms = torch.tensor([e0.elapsed_time(e1)], device=dev)
    if dist:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        dist.barrier()
    ms_total = float(ms.item())
    if rank == 0:
        value = total * n * args.steps / (ms_total * 1e-3) / 1e6
        line = {"metric": "Mcell-timesteps/s", "value": value, "unit": "Mcell-timesteps/s", "n_gpus": world, "steps": args.steps,
                "warmup": max(args.warmup, 3), "ms_per_step": ms_total / args.steps, "higher_is_better": True, "scaling": "strong",
                "vs_baseline": None, "dtype": "f32", "data": "synthetic",
                "config": {"workload": f"{total}-member forcing/parameter ensemble on the {size}x{size} 1 km basin (BASELINE config 5), "
                                       f"members sharded across {world} GPU(s), LDD structures shared per GPU, no collective",
                           "cells_per_member": n, "members_per_gpu": len(mine), "setup_s": round(setup_s, 1),
                           "routing": f"members routed {Ensemble.GROUP} maps per pass" if ens.grouped else "one pass per member",
                           "cuda_graph": ens.state is not None},
                "gpu_launches": (2 * total + 5 * -(-total // Ensemble.GROUP)) * args.steps}
        print(json.dumps(line), flush=True)
    for model in reversed(members):
        model.close()
    if dist:
        dist.barrier()
        dist.destroy_process_group()
