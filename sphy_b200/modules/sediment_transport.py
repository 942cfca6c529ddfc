"""modules/sediment_transport.py of the reference (/root/reference/modules/sediment_transport.py) on device maps:
init() derives the channel network, the stream slopes and the roughness factor once (host NumPy over the exported
LDD structures, sediment_transport.py:113-214); dynamic_mmf() routes the day's sediment with the capacity-limited level
sweep `sphy_accucapacity` (accucapacityflux / accucapacitystate, sediment_transport.py:38-44)."""
import numpy as np
import torch

from . import mmf
from ._real4 import f32, r4_div, r4_if, r4_pow


def _canonical(self):
    """device index -> canonical (row-major) index and its inverse"""
    to_canon = self.cell_order
    to_dev = np.empty(self.n, np.int64)
    to_dev[to_canon] = np.arange(self.n)
    return to_canon, to_dev


def subcatchment_levels(down, order, level_ptr, points):
    """pcr.subcatchment(ldd, points) over the canonical LDD structures (`down` receiver or -1, `order` cells by level,
    `level_ptr`): every cell takes the id of the first non-zero point on its way downstream.  Level by level from the
    outlets, vectorised per level."""
    n = len(down)
    pts = np.asarray(points, np.int64)
    out = np.zeros(n + 1, np.int64)                              # slot n: "drains out of the map"
    for lv in range(len(level_ptr) - 2, -1, -1):
        c = order[level_ptr[lv]:level_ptr[lv + 1]]
        d = np.where(down[c] >= 0, down[c], n)
        out[c] = np.where(pts[c] != 0, pts[c], out[d])
    return out[:n]


def streamorder_levels(down, order, level_ptr):
    """pcr.streamorder(ldd): Strahler order (headwater cells 1; two or more donors of the maximum order -> +1), level by
    level from the headwaters."""
    n = len(down)
    mx = np.zeros(n, np.int64)      # largest donor order seen so far
    cnt = np.zeros(n, np.int64)     # how many donors have it
    so = np.ones(n, np.int64)
    for lv in range(len(level_ptr) - 1):
        c = order[level_ptr[lv]:level_ptr[lv + 1]]
        so[c] = np.where(mx[c] == 0, 1, np.where(cnt[c] >= 2, mx[c] + 1, mx[c]))
        c = c[down[c] >= 0]
        d, o = down[c], so[c]
        # donors of this level raise the maximum of their receivers first, then the ties are counted
        before = mx[d].copy()
        np.maximum.at(mx, d, o)
        raised = np.unique(d[mx[d] > before])
        cnt[raised] = 0
        np.add.at(cnt, d, (o == mx[d]).astype(np.int64))
    return so


def _structures(self):
    return (self.ldd_structure("down").cpu().numpy().astype(np.int64), self.ldd_structure("order").cpu().numpy().astype(np.int64),
            self.ldd_structure("level_ptr").cpu().numpy().astype(np.int64))


def subcatchment(self, points_dev):
    """pcr.subcatchment for maps in device cell order (the LDD structures are exported in canonical numbering)."""
    to_canon, to_dev = _canonical(self)
    return subcatchment_levels(*_structures(self), np.asarray(points_dev, np.int64)[to_dev])[to_canon]


def streamorder(self):
    to_canon, _ = _canonical(self)
    return streamorder_levels(*_structures(self))[to_canon]


def init(self, pcr, config, csv=None, np_=None):
    """sediment_transport.init for ResFLAG = 0 (sediment_transport.py:113-136,175-214)."""
    n = self.n
    if self.ResFLAG == 1:                                                                     # sediment_transport.py:115-118
        self.ResSedID = np.maximum(self._compact_np(self._readmap(config.get("RESERVOIR", "ResId")), np.int32), 0)
    else:
        self.ResSedID = np.broadcast_to(self.Locations, (n,))
    ones = torch.ones(n, dtype=torch.float32, device=self.device)
    acc = torch.empty_like(ones)
    self._ck(self.lib.sphy_accuflux(self._ctx, ones.data_ptr(), None, acc.data_ptr(), 0, self._stream()), "sphy_accuflux")
    self.UpstreamArea = acc.cpu().numpy() * self.cellArea / f32(10 ** 6)                      # sediment_transport.py:122
    self.Upstream_km2 = config.getfloat("SEDIMENT_TRANS", "upstream_km2")
    self.Channel = self.UpstreamArea > f32(self.Upstream_km2)
    self.Basin = subcatchment(self, self.ResSedID)
    self.StreamOrder = streamorder(self)
    ch = self.Channel.astype(np.float32)
    self.Streams = ch * self.Basin.astype(np.float32) * f32(10) + ch * self.StreamOrder.astype(np.float32)
    # areaaverage(Slope, nominal(Streams)): float64 mean over the cells of a class in raster order, REAL4 result
    to_canon, to_dev = _canonical(self)
    slope = np.ascontiguousarray(np.broadcast_to(self.Slope, (n,)), np.float32)
    cls = self.Streams.astype(np.int64)
    slope_c, cls_c = slope[to_dev], cls[to_dev]
    avg_c = np.empty(n, np.float32)
    for k in np.unique(cls_c):
        sel = cls_c == k
        avg_c[sel] = f32(np.mean(slope_c[sel].astype(np.float64)))
    avg = avg_c[to_canon]
    self.SlopeStreams = r4_if(self.Channel, avg, slope)                                       # sediment_transport.py:135-136
    if self.ResFLAG == 1:
        _reservoir_init(self, config)
    self.TC_beta = config.getfloat("SEDIMENT_TRANS", "TC_beta")
    self.TC_gamma = config.getfloat("SEDIMENT_TRANS", "TC_gamma")
    if self.SedModel != 2:
        raise NotImplementedError("only the MMF erosion model (SedModel = 2) is supported")
    self.manningChannelsFLAG = config.getint("SEDIMENT_TRANS", "manningChannelFLAG")
    if self.manningChannelsFLAG == 1:
        self.manningChannel = config.getfloat("SEDIMENT_TRANS", "manningChannel")
    self.n_veg_TC = mmf._veg_manning(self, self.d_TC)
    self.n_TC = r4_pow(r4_pow(self.n_soil, 2) + r4_pow(self.n_veg_TC, 2), 0.5)
    if self.manningChannelsFLAG == 1:
        self.n_TC = r4_if(self.Channel, self.manningChannel, self.n_TC)
    self.v_TC = mmf.FlowVelocity(self, None, self.n_TC, self.d_TC)
    if self.harvest_FLAG:                          # sediment_transport.py:188-196: the in-field harvest Manning is reused
        self.n_veg_TC_harvest = r4_if(self.Tillage_harvest == 1, 0, self.n_veg_field_harvest)
        self.n_TC_harvest = r4_pow(r4_pow(self.n_soil, 2) + r4_pow(self.n_veg_TC_harvest, 2), 0.5)
        if self.manningChannelsFLAG == 1:
            self.n_TC_harvest = r4_if(self.Channel & (self.n_TC_harvest > 0), self.manningChannel, self.n_TC_harvest)
        self.v_TC_harvest = mmf.FlowVelocity(self, None, self.n_TC_harvest, self.d_TC)
    self.v_b = mmf.FlowVelocity(self, None, self.n_bare, self.d_bare)
    self.roughnessFactor = r4_div(self.v_TC, self.v_b)


def _reservoir_init(self, config):
    """sediment_transport.init, reservoir part (sediment_transport.py:138-173): extent, trapping efficiencies, the
    sub-catchment of every reservoir and the routing steps of resorder.txt, folded into the sphy_sedres block."""
    import ctypes as C
    import os

    from .. import _lib
    if self.ETOpenWaterFLAG == 1:
        raise NotImplementedError("open-water classes as reservoir extent (sediment_transport.py:144-145) are not supported yet")
    n = self.n
    extent = np.maximum(self._compact_np(self._readmap(config.get("RESERVOIR", "reservoirs")), np.int32), 0)
    trap = {}
    with open(os.path.join(self.inpath, config.get("SEDIMENT_TRANS", "TrapEffTab"))) as fh:
        for line in fh:
            t = line.replace(",", " ").split()
            if t and not t[0].startswith("#"):
                trap[int(float(t[0]))] = f32(float(t[-1]))
    order, steps = [], []
    with open(os.path.join(self.inpath, config.get("SEDIMENT_TRANS", "ResOrder"))) as fh:
        next(fh)                                                 # heading
        for line in fh:
            t = line.split()
            if len(t) >= 2:
                order.append(int(t[0]))
                steps.append(int(t[1]))
    order, steps = np.asarray(order, np.int64), np.asarray(steps, np.int64)
    self.subcatchmentOrder, self.subcatchmentSteps = order, steps
    self.subcatchmentRes = self.Basin                            # subcatchment(FlowDir, ResSedID), computed by init()
    n_steps = int(steps.max()) + 1 if len(steps) else 1
    # the reference's iteration order: step by step, within a step in file order
    seq = [k for st in range(n_steps) for k in np.flatnonzero(steps == st)]
    step_of, seq_of = {}, {}
    for g, k in enumerate(seq):
        step_of.setdefault(int(order[k]), int(steps[k]))
        seq_of.setdefault(int(order[k]), g)
    never = np.iinfo(np.int32).max
    finish_step = np.array([step_of.get(int(r), never) for r in self.subcatchmentRes], np.int32)
    down = self.ldd_structure("down").cpu().numpy().astype(np.int64)           # canonical numbering
    to_canon, to_dev = _canonical(self)
    res_cell, res_down, res_step, dfs, te, oe = [], [], [], [], [], []
    step_last = np.full(n_steps, -1, np.int32)
    for g, k in enumerate(seq):
        rid = int(order[k])
        cells = np.flatnonzero(self.ResSedID == rid)
        if len(cells) != 1:
            raise NotImplementedError(f"reservoir {rid} of resorder.txt must mark exactly one cell of ResId")
        c = int(cells[0])
        d = int(down[to_canon[c]])
        x = int(to_dev[d]) if d >= 0 else -1
        res_cell.append(c)
        res_down.append(x)
        res_step.append(int(steps[k]))
        dfs.append(seq_of.get(int(self.subcatchmentRes[x]), never) if x >= 0 else never)
        t = trap.get(rid)
        te.append(f32(0) if t is None else t)                    # cover(lookupscalar(...), 0) / cover(1 - lookupscalar(...), 1)
        oe.append(f32(1) if t is None else f32(1) - t)
        step_last[int(steps[k]):] = g
    i32 = lambda a: self._dev(np.asarray(a, np.int32).reshape(-1) if len(a) else np.zeros(1, np.int32))  # noqa: E731
    R = _lib.SedRes()
    R.n_res, R.n_steps = len(seq), n_steps
    keep = dict(res_cell=i32(res_cell), res_down=i32(res_down), res_step=i32(res_step), down_finish_seq=i32(dfs),
                trap_eff=self._dev(np.asarray(te if te else [0], np.float32)), outflow_eff=self._dev(np.asarray(oe if oe else [1], np.float32)),
                finish_step=self._dev(finish_step), extent=self._dev(extent.astype(np.int32)),
                outflow=self._full(0.0)[: max(len(seq), 1)].clone())
    for k in ("sed_trans", "tc_route", "cap_flux", "cap_state"):
        keep[k] = self._full(0.0)
    for k, t in keep.items():
        setattr(R, k, t.data_ptr())
    self._sedres_last = (C.c_int32 * n_steps)(*[int(v) for v in step_last])
    R.step_last_seq_host = C.cast(self._sedres_last, C.c_void_p)
    self._sedres, self._sedres_keep = R, keep


def dynamic_mmf(self, pcr, Runoff, np_, G):
    """sediment_transport.dynamic_mmf / SedTrans without reservoirs (sediment_transport.py:38-44,244-275): the transport
    capacity was evaluated by sphy_mmf_step; G is the sediment in transport in ton per cell.  -> (yield, deposition, flux)."""
    import ctypes as C
    o = self.sed_out
    if self.ResFLAG == 1:                                        # sediment_transport.py:46-107
        self._ck(self.lib.sphy_sed_reservoir_route(self._ctx, C.byref(self._sedres), G.data_ptr(), o["TC"].data_ptr(),
                                                   o["SedYld"].data_ptr(), o["SedDep"].data_ptr(), o["SedFlux"].data_ptr(),
                                                   self._stream()), "sphy_sed_reservoir_route")
        return o["SedYld"], o["SedDep"], o["SedFlux"]
    self._ck(self.lib.sphy_accucapacity(self._ctx, G.data_ptr(), o["TC"].data_ptr(), o["SedFlux"].data_ptr(),
                                        o["SedDep"].data_ptr(), self._stream()), "sphy_accucapacity")
    return o["SedDep"], o["SedDep"], o["SedFlux"]
