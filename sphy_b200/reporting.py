"""Reporting accumulators and outputs -- utilities/reporting.py of the reference (/root/reference/utilities/
reporting.py:24-181, driven by reporting.csv) on device maps.

Per variable the csv selects map outputs (D, M, Y, F(inal), MS, YS), average maps (M, Y, MA, YA) and time series at
`Locations` (D, M, Y).  The accumulations (`tot + var`, `var / dim`, `tot / ydays`) are single REAL4 map operations in
the reference; here each is one launch of sphy_map_accumulate / sphy_map_divide on device tensors, and only the
requested outputs leave the device (as CSF maps / .tss files, or kept in memory for tests).
The reference evaluates ~45 reporting argument expressions every step whether or not they are reported (SURVEY.md
section 5); here nothing is computed for variables the csv does not select.
"""
from __future__ import annotations

import calendar
import csv
import os

import numpy as np
import torch

from . import _lib
from . import csf

PERIODS = ("Day", "Month", "Year", "Final", "MonthSum", "YearSum", "MonthAvg", "YearAvg")


def merge_tss(every):
    """Time-series tables of a basin split across GPUs (pure host logic).  `every[r]` = {"ids": station ids rank r owns,
    "tss": {name: [(timestep, values at those stations)]}}; -> {name: (all station ids ascending, [(timestep, values)])}
    with the columns ordered by station id like the reference's TimeoutputTimeseries (NaN where nobody sampled)."""
    ids = np.array(sorted(int(i) for e in every for i in e["ids"]), np.int64)
    if len(set(ids.tolist())) != len(ids):
        raise ValueError("a station id is owned by more than one rank")
    col = {int(i): k for k, i in enumerate(ids)}
    out = {}
    for name in sorted({k for e in every for k in e["tss"]}):
        steps = sorted({t for e in every for t, _ in e["tss"].get(name, [])})
        table = {t: np.full(len(ids), np.nan, np.float32) for t in steps}
        for e in every:
            cols = [col[int(i)] for i in e["ids"]]
            for t, v in e["tss"].get(name, []):
                table[t][cols] = np.asarray(v, np.float32)
        out[name] = (ids, [(t, table[t]) for t in steps])
    return out


class Reporting:
    def __init__(self, model, csv_path, write_files=False):
        self.m = model
        self.write_files = write_files
        self.vars = {}             # name -> dict(fname, map=set, avg=set, ts=set)
        self.acc = {}              # (name, period) -> tensor | dict month -> tensor
        self.maps = {}             # output path (relative to outputdir) -> ndarray raster   (what self.report wrote)
        self.tss = {}              # tss name -> list of (timestep, values at the stations)
        with open(csv_path) as fh:
            next(fh)
            for row in csv.reader(fh):
                if not row or row[0][:1] == "#":
                    continue
                name, mp, av, ts, fname = row[0], row[1], row[2], row[3], row[4]
                if mp == av == ts == "NONE":
                    continue
                self.vars[name] = dict(fname=fname, map=set() if mp == "NONE" else set(mp.split("+")),
                                       avg=set() if av == "NONE" else set(av.split("+")),
                                       ts=set() if ts == "NONE" else set(ts.split("+")))
        missing = [n for n in self.vars if n not in model.reportable()]
        if missing:
            raise NotImplementedError("reporting.csv selects variables this build cannot report with the current switches: "
                                      + ", ".join(missing))
        z = lambda: torch.zeros(model.n, dtype=torch.float32, device=model.device)  # noqa: E731
        for name, v in self.vars.items():
            if "M" in v["map"] or "M" in v["avg"] or "M" in v["ts"]:
                self.acc[(name, "Month")] = z()
            if "Y" in v["map"] or "Y" in v["avg"] or "Y" in v["ts"]:
                self.acc[(name, "Year")] = z()
            if v["map"] - {"D", "M", "Y", "MS", "YS"}:                     # anything else is the Final option
                self.acc[(name, "Final")] = z()
            if "MS" in v["map"]:
                self.acc[(name, "MonthSum")] = {mth: z() for mth in range(12)}
            if "YS" in v["map"]:
                self.acc[(name, "YearSum")] = z()
            if "MA" in v["avg"]:
                self.acc[(name, "MonthAvg")] = {mth: z() for mth in range(12)}
            if "YA" in v["avg"]:
                self.acc[(name, "YearAvg")] = z()
        self._tmp = z()
        self._sbuf = torch.empty(max(len(model.station_ids), 1), dtype=torch.float32, device=model.device)

    # -- helpers -----------------------------------------------------------------------------------
    def _accumulate(self, acc, x, divisor=0.0):
        m = self.m
        _lib.check(m.lib, m._ctx, m.lib.sphy_map_accumulate(m._ctx, acc.data_ptr(), x.data_ptr(), float(divisor), m.n, m._stream()),
                   "sphy_map_accumulate")

    def _divided(self, x, divisor):
        m = self.m
        _lib.check(m.lib, m._ctx, m.lib.sphy_map_divide(m._ctx, x.data_ptr(), float(divisor), self._tmp.data_ptr(), m.n, m._stream()),
                   "sphy_map_divide")
        return self._tmp

    def _report(self, tensor, rel_path):
        raster = self.m.gather_raster(tensor)       # a basin split across GPUs: collective, rank 0 holds and writes the raster
        if raster is None:
            return
        self.maps[rel_path] = raster
        if self.write_files:
            path = os.path.join(self.m.outpath, rel_path)
            os.makedirs(os.path.dirname(path) or ".", exist_ok=True)
            csf.write_map(path, raster, cellsize=self.m.cellsize)

    def _sample(self, tss_name, tensor):
        m = self.m
        n = len(m.station_ids)
        if n == 0:
            return
        _lib.check(m.lib, m._ctx, m.lib.sphy_sample(m._ctx, tensor.data_ptr(), m._station_cells.data_ptr(), n, self._sbuf.data_ptr(),
                                                    m._stream()), "sphy_sample")
        self.tss.setdefault(tss_name, []).append((m.timestep, self._sbuf[:n].cpu().numpy().copy()))

    # -- one day, after dynamic() has produced the day's maps (REPM, reporting.py:24-95) ------------------
    def step(self):
        m = self.m
        date = m.curdate                                # the day just simulated
        dim = calendar.monthrange(date.year, date.month)[1]
        ydays = 366 if calendar.isleap(date.year) else 365
        doy = date.timetuple().tm_yday
        last_dom = date.day == dim
        after_spinup = date.year >= m.startYear + m.spinUpYears
        at_end = date == m.enddate
        t = m.timestep
        src = m.reportable()
        for name, v in self.vars.items():
            var = src[name]()
            f = v["fname"]
            if "D" in v["ts"]:
                self._sample(f + "DTS", var)
            if "D" in v["map"]:
                self._report(var, csf.stack_name(f, t))
            if (name, "Month") in self.acc:
                tot = self.acc[(name, "Month")]
                self._accumulate(tot, var)
                if last_dom:
                    if "M" in v["ts"]:
                        self._sample(f + "MTS", tot)
                    if "M" in v["map"]:
                        self._report(tot, csf.stack_name(f + "M", t))
                    if "M" in v["avg"]:
                        self._report(self._divided(tot, dim), csf.stack_name(f + "M", t))
                    tot.zero_()
            if (name, "Year") in self.acc:
                tot = self.acc[(name, "Year")]
                self._accumulate(tot, var)
                if doy == ydays:
                    if "Y" in v["ts"]:
                        self._sample(f + "YTS", tot)
                    if "Y" in v["map"]:
                        self._report(tot, csf.stack_name(f + "Y", t))
                    if "Y" in v["avg"]:
                        self._report(self._divided(tot, ydays), csf.stack_name(f + "Y", t))
                    tot.zero_()
            if (name, "Final") in self.acc:
                tot = self.acc[(name, "Final")]
                if not at_end:
                    self._accumulate(tot, var)
                else:
                    self._report(tot, f + ".map")           # the end day itself is not added (reporting.py:89-94)
                    tot.zero_()
            if (name, "MonthSum") in self.acc:
                tot = self.acc[(name, "MonthSum")][date.month - 1]
                if not last_dom and after_spinup:
                    self._accumulate(tot, var)
                if last_dom and date.year == m.endYear:
                    self._report(self._divided(tot, m.simYears), f + "SumM" + str(date.month).zfill(2) + ".map")
            if (name, "YearSum") in self.acc:
                tot = self.acc[(name, "YearSum")]
                if after_spinup:
                    self._accumulate(tot, var)
                if at_end:
                    self._report(self._divided(tot, m.simYears), f + "SumY.map")
            if (name, "MonthAvg") in self.acc:
                tot = self.acc[(name, "MonthAvg")][date.month - 1]
                if not last_dom and after_spinup:
                    self._accumulate(tot, var, dim)
                elif last_dom and date.year == m.endYear:
                    self._report(self._divided(tot, m.simYears), f + "AvgM" + str(date.month).zfill(2) + ".map")
            if (name, "YearAvg") in self.acc:
                tot = self.acc[(name, "YearAvg")]
                if after_spinup:
                    self._accumulate(tot, var, ydays)
                if at_end:
                    self._report(self._divided(tot, m.simYears), f + "AvgY.map")

    def tss_tables(self):
        """tss name -> (station ids, [(timestep, values)]).  On a basin split across GPUs every rank sampled the stations
        inside its own range: COLLECTIVE, rank 0 gets the merged tables (stations by ascending id, like the reference's
        TimeoutputTimeseries), the other ranks an empty dict."""
        m = self.m
        if m.world == 1:
            return {name: (m.station_ids, rows) for name, rows in self.tss.items()}
        import torch.distributed as dist
        mine = {"ids": [int(i) for i in m.station_ids], "tss": {k: [(t, v.tolist()) for t, v in rows] for k, rows in self.tss.items()}}
        every = [None] * m.world
        dist.all_gather_object(every, mine)
        return merge_tss(every) if m.rank == 0 else {}

    def write_tss(self):
        """The reference moves *.tss to the output directory at the end of the run (sphy.py:1273-1278)."""
        for name, (ids, rows) in self.tss_tables().items():
            os.makedirs(self.m.outpath, exist_ok=True)
            csf.write_tss(os.path.join(self.m.outpath, name + ".tss"), ids, rows)
