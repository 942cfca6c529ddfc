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


def subcatchment(self, points_dev):
    """pcr.subcatchment(ldd, points): every cell takes the id of the first non-zero point on its way downstream.
    Level by level from the outlets, over the canonical LDD structures; returns the ids by device index."""
    to_canon, to_dev = _canonical(self)
    down = self.ldd_structure("down").cpu().numpy().astype(np.int64)
    order = self.ldd_structure("order").cpu().numpy().astype(np.int64)
    lptr = self.ldd_structure("level_ptr").cpu().numpy().astype(np.int64)
    pts = np.asarray(points_dev, np.int64)[to_dev]
    out = np.zeros(self.n + 1, np.int64)                         # slot n: "drains out of the map"
    for lv in range(len(lptr) - 2, -1, -1):
        c = order[lptr[lv]:lptr[lv + 1]]
        d = np.where(down[c] >= 0, down[c], self.n)
        out[c] = np.where(pts[c] != 0, pts[c], out[d])
    return out[:self.n][to_canon]


def streamorder(self):
    """pcr.streamorder(ldd): Strahler order, level by level from the headwaters; by device index."""
    to_canon, _ = _canonical(self)
    down = self.ldd_structure("down").cpu().numpy().astype(np.int64)
    order = self.ldd_structure("order").cpu().numpy().astype(np.int64)
    lptr = self.ldd_structure("level_ptr").cpu().numpy().astype(np.int64)
    n = self.n
    mx = np.zeros(n, np.int64)      # largest donor order seen so far
    cnt = np.zeros(n, np.int64)     # how many donors have it
    so = np.ones(n, np.int64)
    for lv in range(len(lptr) - 1):
        c = order[lptr[lv]:lptr[lv + 1]]
        so[c] = np.where(mx[c] == 0, 1, np.where(cnt[c] >= 2, mx[c] + 1, mx[c]))
        c = c[down[c] >= 0]
        d, o = down[c], so[c]
        # donors of this level raise the maximum of their receivers first, then the ties are counted
        before = mx[d].copy()
        np.maximum.at(mx, d, o)
        raised = np.unique(d[mx[d] > before])
        cnt[raised] = 0
        np.add.at(cnt, d, (o == mx[d]).astype(np.int64))
    return so[to_canon]


def init(self, pcr, config, csv=None, np_=None):
    """sediment_transport.init for ResFLAG = 0 (sediment_transport.py:113-136,175-214)."""
    if self.ResFLAG == 1:
        raise NotImplementedError("sediment transport through reservoirs (sediment_transport.py:46-107,138-173) is not supported yet")
    n = self.n
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


def dynamic_mmf(self, pcr, Runoff, np_, G):
    """sediment_transport.dynamic_mmf / SedTrans without reservoirs (sediment_transport.py:38-44,244-275): the transport
    capacity was evaluated by sphy_mmf_step; G is the sediment in transport in ton per cell.  -> (yield, deposition, flux)."""
    o = self.sed_out
    self._ck(self.lib.sphy_accucapacity(self._ctx, G.data_ptr(), o["TC"].data_ptr(), o["SedFlux"].data_ptr(),
                                        o["SedDep"].data_ptr(), self._stream()), "sphy_accucapacity")
    return o["SedDep"], o["SedDep"], o["SedFlux"]
