"""Synthetic catchments for tests and bench: DEM -> priority-flood LDD, soils, forcing,
glacier table, reservoirs/lakes, plus a complete `sphy_config.cfg` + lookup tables + `reporting.csv`
written in the reference's own formats (/root/reference/sphy_config.cfg), so the same case can be fed to
the reference (through the oracle harness), to the oracle port and to the B200 SphyModel.

Map names are the cfg values relative to `inputdir`, exactly as the reference concatenates them
(`self.inpath + config.get(...)`, /root/reference/sphy.py:141-198).
"""
from __future__ import annotations

import ctypes
import datetime
import os
from dataclasses import dataclass, field

import numpy as np

from . import _build

_HG = None


def _hostgen():
    global _HG
    if _HG is None:
        lib = ctypes.CDLL(_build.build_hostgen())
        P = ctypes.c_void_p
        lib.hostgen_flood_ldd.restype = ctypes.c_int
        lib.hostgen_flood_ldd.argtypes = [P, ctypes.c_int, ctypes.c_int, ctypes.c_float, ctypes.c_int, ctypes.c_int,
                                          ctypes.c_int, P, P]
        lib.hostgen_dem.restype = None
        lib.hostgen_dem.argtypes = [P, ctypes.c_int, ctypes.c_int, ctypes.c_uint64, ctypes.c_float, ctypes.c_float]
        lib.hostgen_nup.restype = ctypes.c_int
        lib.hostgen_nup.argtypes = [P, ctypes.c_int, ctypes.c_int, P]
        _HG = lib
    return _HG


def _ptr(a):
    return a.ctypes.data_as(ctypes.c_void_p)


def make_dem_ldd(nrows, ncols, cellsize, seed=1, outlets="single", zrange=None):
    """Returns (dem_raw, dem_filled, ldd uint8, slope float32)."""
    hg = _hostgen()
    dem = np.empty((nrows, ncols), np.float32)
    zmin, zmax = (zrange if zrange else (0.0, 0.0))
    hg.hostgen_dem(_ptr(dem), nrows, ncols, seed, zmin, zmax)
    raw = dem.copy()
    ldd = np.empty((nrows, ncols), np.uint8)
    slope = np.empty((nrows, ncols), np.float32)
    mode = 1 if outlets == "single" else 0
    rc = hg.hostgen_flood_ldd(_ptr(dem), nrows, ncols, float(cellsize), mode, nrows - 1, (ncols - 1) // 2,
                              _ptr(ldd), _ptr(slope))
    if rc != 0:
        raise MemoryError("hostgen_flood_ldd failed")
    return raw, dem, ldd, slope


def upstream_counts(ldd):
    hg = _hostgen()
    ldd = np.ascontiguousarray(ldd, np.uint8)
    nup = np.empty(ldd.shape, np.int32)
    rc = hg.hostgen_nup(_ptr(ldd), ldd.shape[0], ldd.shape[1], _ptr(nup))
    if rc != 0:
        raise ValueError("LDD has a cycle")
    return nup


def _smooth_noise(rng, nrows, ncols, k=8):
    """Bilinear up-sampling of a (k+1)x(k+1) N(0,1) lattice -> smooth field, float32."""
    lat = rng.standard_normal((k + 1, k + 1))
    r = np.linspace(0, k, nrows)
    c = np.linspace(0, k, ncols)
    r0 = np.minimum(r.astype(int), k - 1)
    c0 = np.minimum(c.astype(int), k - 1)
    tr = (r - r0)[:, None]
    tc = (c - c0)[None, :]
    a = lat[r0][:, c0]
    b = lat[r0][:, c0 + 1]
    d = lat[r0 + 1][:, c0]
    e = lat[r0 + 1][:, c0 + 1]
    return ((a * (1 - tc) + b * tc) * (1 - tr) + (d * (1 - tc) + e * tc) * tr).astype(np.float32)


# cfg defaults of the reference (/root/reference/sphy_config.cfg:161-176,197-202,241-267,391-407,491)
DEFAULT_PARAMS = dict(
    RootFieldMap=0.30, RootSatMap=0.45, RootDryMap=0.08, RootWiltMap=0.15, RootKsat=50.0,
    SubFieldMap=0.30, SubSatMap=0.45, SubKsat=20.0,
    RootDepthFlat=400.0, SubDepthFlat=1600.0, CapRiseMax=0.571,
    GwDepth=2000.0, GwSat=200.0, deltaGw=300.0, BaseThresh=0.0, alphaGw=0.508, YieldGw=0.05,
    GwRecharge=0.0, BaseR=0.0, Gw=100.0, H_gw=3.0,
    SubWater=10.0, CapRise=0.0, RootDrain=0.0, SubDrain=0.0,
    kx=0.971, Gsc=0.0820, Kc=1.0,
    Tcrit=-0.52, SnowSC=0.196, DDFS=1.401, SnowF=0.274, SnowCth=5.0, SnowIni=0.0, SnowWatStore=0.0,
    DDFG=4.0, DDFDG=4.0, GlacF=0.555,
    Alpha=0.32, Labda_infil=1.65, K_eff=0.5,
    Pcorr_fact=1.0, Tcorr_fact=0.0,
    SeePage=0.5, GWL_base=2.0,
    kcOpenWater=1.2,
)


@dataclass
class Catchment:
    nrows: int
    ncols: int
    cellsize: float
    seed: int
    dem: np.ndarray
    ldd: np.ndarray
    slope: np.ndarray
    lat: np.ndarray
    flags: dict
    params: dict
    start: datetime.date
    nsteps: int
    locations: np.ndarray
    soil_maps: dict
    rout_components: bool = False
    infil_excess: int = 0
    glac_table: dict | None = None       # column name -> ndarray (U_ID, MOD_ID, GLAC_ID, MOD_H, GLAC_H, DEBRIS, FRAC_GLAC, ICE_DEPTH)
    model_id: np.ndarray | None = None
    glac_id: np.ndarray | None = None
    tlapse: np.ndarray | None = None     # 12 monthly lapse rates
    res_id: np.ndarray | None = None
    lake_id: np.ndarray | None = None
    res_tables: dict = field(default_factory=dict)   # file name -> 2-D list (matrix table incl. header row)
    lake_tables: dict = field(default_factory=dict)
    open_water_frac: np.ndarray | None = None
    report_daily: tuple = ()
    pws: bool = False                    # plant water stress (ET.ks, PWS_FLAG)
    report_opts: dict = field(default_factory=dict)   # reporting.csv: name -> (map, avg, timeseries) option strings
    kc_series: bool = False              # KCstatic = 0: crop-coefficient map series with gaps (sphy.py:824-830)
    seep_series: bool = False            # SeepStatic = 0 (GroundFLAG = 0): seepage map series with gaps (sphy.py:992-1000)
    pedotransfer: dict | None = None     # PedotransferFLAG = 1: soil hydraulic properties from sand/clay/organic matter (utilities/pedotransfer.py)
    lake_levels: bool = False            # measured lake levels update the lake storage on some days (lakes.py:28-97)
    sediment: dict | None = None         # SedFLAG = 1, MMF erosion (SedModel = 2) + capacity-limited transport: maps and tables
    glac_retreat: tuple | None = None    # (day, month) of the yearly glacier update (GlacRetreat = 1, glacier.py:562-736)
    glac_vars: tuple = ()                # per-glacier reporting (GlacID_flag = 1, glacier.py:501-559): GlacTable columns
    dynveg: bool = False                 # DynVegFLAG = 1: NDVI map series -> Kc, LAI, canopy interception (modules/dynamic_veg.py)
    dynveg_maps: bool = False            # NDVImax, KCmax and FPARmax as maps instead of cfg floats (dynamic_veg.py:57-63 reads maps first)
    etref_input: bool = False            # ETREF_FLAG = 1: ETref map series instead of Hargreaves (sphy.py:814)
    mask: np.ndarray | None = None       # clone mask (True = inside the catchment); None = the whole rectangle

    # ------------------------------------------------------------------------------------------
    @property
    def ncell(self):
        return self.nrows * self.ncols

    def forcing(self, day):
        """Day `day` (1-based model time step) -> dict of float32 rasters prec, tavg, tmin, tmax."""
        rng = np.random.default_rng([self.seed + 2, int(day)])
        doy = (self.start + datetime.timedelta(days=int(day) - 1)).timetuple().tm_yday
        z = self.dem
        season = 10.0 * np.sin(2.0 * np.pi * (doy - 105) / 365.0)
        tavg = (12.0 + self.params.get("Toffset", 0.0) - 6.5e-3 * (z - 1000.0) + season + 2.0 * rng.standard_normal()
                + 1.0 * _smooth_noise(rng, self.nrows, self.ncols)).astype(np.float32)
        tmax = (tavg + np.float32(5.0) + (2.0 * rng.random((self.nrows, self.ncols))).astype(np.float32)).astype(np.float32)
        tmin = (tavg - np.float32(5.0) - (2.0 * rng.random((self.nrows, self.ncols))).astype(np.float32)).astype(np.float32)
        wet = rng.random() < 0.35
        if wet:
            base = rng.gamma(0.8, 12.0)
            mult = np.maximum(0.0, 1.0 + 0.3 * rng.standard_normal((self.nrows, self.ncols)))
            prec = (self.params.get("Pscale", 1.0) * base * mult).astype(np.float32)
        else:
            rng.gamma(0.8, 12.0)
            prec = np.zeros((self.nrows, self.ncols), np.float32)
        out = dict(prec=prec, tavg=tavg, tmin=tmin, tmax=tmax)
        if self.etref_input:
            out["etref"] = np.clip(np.float32(0.12) * (tavg + np.float32(12.0)), 0, None).astype(np.float32)
        if self.kc_series and day % 3 == 2:                      # gaps: the previous map (1 before the first) is kept
            out["kc"] = np.clip(0.9 + 0.3 * _smooth_noise(rng, self.nrows, self.ncols, 4), 0.3, 1.4).astype(np.float32)
        if self.seep_series and day % 4 == 0:
            out["seep"] = np.clip(0.4 + 0.4 * _smooth_noise(rng, self.nrows, self.ncols, 4), 0.0, 1.5).astype(np.float32)
        if self.lake_levels and self.lake_id is not None and day % 5 == 3:
            lev = np.full((self.nrows, self.ncols), np.nan, np.float32)        # MV where nothing was measured
            for r_, c_ in zip(*np.nonzero(self.lake_id > 0)):
                lid = int(self.lake_id[r_, c_])
                if lid % 2 == 1 and not (lid == 3 and day % 10 == 3):          # lake 3 has gaps in its record
                    lev[r_, c_] = np.float32(2.0 * rng.uniform(0.85, 1.15))
            out["llev"] = lev
        if self.dynveg:
            season_v = 0.35 + 0.25 * np.sin(2.0 * np.pi * (doy - 120) / 365.0)
            out["ndvi"] = np.clip(season_v + 0.15 * _smooth_noise(rng, self.nrows, self.ncols, 6)
                                  + 0.05 * rng.standard_normal((self.nrows, self.ncols)), -0.05, 0.95).astype(np.float32)
        return out

    # ------------------------------------------------------------------------------------------
    def static_maps(self):
        """name (relative to inputdir) -> ndarray; dtype decides the PCRaster value scale
        (bool -> boolean, uint8 -> ldd, int32 -> nominal, float32 -> scalar)."""
        m = {
            "clone.map": (np.broadcast_to(np.bool_(True), (self.nrows, self.ncols)) if self.mask is None
                          else np.asarray(self.mask, np.bool_)),
            "dem.map": np.asarray(self.dem, np.float32),
            "slope.map": np.asarray(self.slope, np.float32),
            "stations.map": np.asarray(self.locations, np.int32),
            "ldd.map": np.asarray(self.ldd, np.uint8),
            "latitude.map": np.asarray(self.lat, np.float32),
            "landuse.map": (np.broadcast_to(np.int32(1), (self.nrows, self.ncols)) if self.sediment is None
                            else self.sediment["landuse"]),
        }
        if self.pedotransfer is not None:
            for k, v in self.pedotransfer.items():
                if isinstance(v, np.ndarray):
                    m[k + ".map"] = v
        if self.sediment is not None and "reservoirs" in self.sediment:
            m["reservoirs_nominal.map"] = self.sediment["reservoirs"]
        if self.sediment is not None and int(self.sediment.get("model", 2)) == 1 and self.pedotransfer is None:
            m["k_usle.map"] = self.sediment["k_usle"]
        if self.sediment is not None:
            m["rock_fraction.map"] = self.sediment["rock"]
            m["root_sand.map"] = self.sediment["sand"]
            m["root_clay.map"] = self.sediment["clay"]
        for k, v in self.soil_maps.items():
            m[_SOIL_FILES[k]] = np.asarray(v, np.float32)
        if self.model_id is not None:
            m["MOD_ID.map"] = self.model_id.astype(np.int32)
            m["GLAC_ID.map"] = self.glac_id.astype(np.int32)
        if self.res_id is not None:
            m["res_id.map"] = self.res_id.astype(np.int32)
        if self.lake_id is not None:
            m["lake_id.map"] = self.lake_id.astype(np.int32)
        if self.open_water_frac is not None:
            m["openwater_frac.map"] = self.open_water_frac.astype(np.float32)
        if self.lake_levels and self.lake_id is not None:
            m["update_lakelevel.map"] = (self.lake_id % 2 == 1)               # boolean: lakes with a gauge
        if self.dynveg_maps:
            rng = np.random.default_rng(self.seed + 4711)
            z = _smooth_noise(rng, self.nrows, self.ncols, 4)
            m["ndvimax.map"] = np.clip(0.65 + 0.08 * z, 0.5, 0.8).astype(np.float32)
            m["kcmax.map"] = np.clip(1.4 + 0.2 * _smooth_noise(rng, self.nrows, self.ncols, 4), 1.0, 1.8).astype(np.float32)
            m["fparmax.map"] = np.where(z > 0, np.float32(0.95), np.float32(0.9)).astype(np.float32)
        return m

    def forcing_map_names(self, day):
        from .csf import stack_name
        names = {k: stack_name("forcings/" + k, day) for k in ("prec", "tavg", "tmin", "tmax")}
        if self.etref_input:
            names["etref"] = stack_name("forcings/etref", day)
        if self.dynveg:
            names["ndvi"] = stack_name("forcings/ndvi", day)
        if self.lake_levels and self.lake_id is not None and day % 5 == 3:
            names["llev"] = stack_name("forcings/llev", day)
        if self.kc_series and day % 3 == 2:
            names["kc"] = stack_name("forcings/kc", day)
        if self.seep_series and day % 4 == 0:
            names["seep"] = stack_name("forcings/seep", day)
        return names

    def write_csf_inputs(self, inputdir, nsteps=None):
        """Write every static map and the daily forcing stacks as PCRaster CSF files under `inputdir`, laid out exactly
        as the reference expects them (clone.map ... forcings/prec0000.001)."""
        from . import csf
        os.makedirs(os.path.join(inputdir, "forcings"), exist_ok=True)
        for name, arr in self.static_maps().items():
            csf.write_map(os.path.join(inputdir, name), np.ascontiguousarray(arr), cellsize=self.cellsize)
        for t in range(1, (nsteps or self.nsteps) + 1):
            f = self.forcing(t)
            for key, name in self.forcing_map_names(t).items():
                csf.write_map(os.path.join(inputdir, name), f[key], cellsize=self.cellsize)

    # ------------------------------------------------------------------------------------------
    def write_inputs(self, inputdir, outputdir=None):
        """Write sphy_config.cfg, reporting.csv and lookup tables under `inputdir`; return cfg path."""
        os.makedirs(inputdir, exist_ok=True)
        outputdir = outputdir or os.path.join(inputdir, "out")
        os.makedirs(outputdir, exist_ok=True)
        p, f = self.params, self.flags
        end = self.start + datetime.timedelta(days=self.nsteps - 1)
        classes = (1,) if self.sediment is None else (1, 2, 3)
        for fname, val in (("kc.tbl", p["Kc"]), ("paved.tbl", 0.0), ("laimax.tbl", p.get("LAImax", 5.5)),
                           ("depletion.tbl", p.get("PFactor", 0.5))):
            with open(os.path.join(inputdir, fname), "w") as fh:
                for c in classes:
                    fh.write(f"{c} {val}\n")
        if self.sediment is not None and "reservoirs" in self.sediment:
            with open(os.path.join(inputdir, "trapefftab.tbl"), "w") as fh:
                for rid, eff in self.sediment["trap_eff"]:
                    fh.write(f"{rid} {eff}\n")
            with open(os.path.join(inputdir, "resorder.txt"), "w") as fh:
                fh.write("order\tsteps\n")
                for rid, step in self.sediment["res_order"]:
                    fh.write(f"{rid}\t{step}\n")
        if self.sediment is not None:
            for fname in ("mmf.tbl", "mmf_harvest.tbl", "musle.tbl"):
                with open(os.path.join(inputdir, fname), "w") as fh:
                    for row in self.sediment[fname]:
                        fh.write(" ".join(repr(float(v)) for v in row) + "\n")
        with open(os.path.join(inputdir, "reporting.csv"), "w") as fh:
            fh.write("name,map,avg,timeseries,filename,comment\n")
            for name, fname in _REPORT_VARS:
                mode = "D" if name in self.report_daily else "NONE"
                mp, av, ts = self.report_opts.get(name, (mode, "NONE", "NONE"))
                fh.write(f"{name},{mp},{av},{ts},{fname},synthetic\n")
        if self.glac_table is not None:
            cols = ["U_ID", "MOD_ID", "GLAC_ID", "MOD_H", "GLAC_H", "DEBRIS", "FRAC_GLAC", "ICE_DEPTH"]
            with open(os.path.join(inputdir, "glaciers.csv"), "w") as fh:
                fh.write(",".join(cols) + "\n")
                n = len(self.glac_table["U_ID"])
                for i in range(n):
                    fh.write(",".join(repr(float(self.glac_table[c][i])) if c in ("MOD_H", "GLAC_H", "FRAC_GLAC", "ICE_DEPTH")
                                      else str(int(self.glac_table[c][i])) for c in cols) + "\n")
            with open(os.path.join(inputdir, "tlapse.tbl"), "w") as fh:
                for mth in range(12):
                    fh.write(f"{mth + 1} {float(self.tlapse[mth])!r}\n")
        for name, rows in list(self.res_tables.items()) + list(self.lake_tables.items()):
            with open(os.path.join(inputdir, name), "w") as fh:
                for row in rows:
                    fh.write(" ".join(repr(float(v)) for v in row) + "\n")

        rout_init = "QRA_init = 0\n"
        if self.rout_components:
            rout_init += "".join(f"{c}RA_init = 0\n" for c in ("RootR", "RootD", "Rain", "Snow", "Glac", "Base"))
        ground = f["GroundFLAG"] == 1 or f["GlacFLAG"] == 1
        cfg = f"""[MODULES]
GlacFLAG = {f['GlacFLAG']}
SnowFLAG = {f['SnowFLAG']}
RoutFLAG = {f['RoutFLAG']}
LakeFLAG = {f['LakeFLAG']}
ResFLAG = {f['ResFLAG']}
DynVegFLAG = {1 if self.dynveg else 0}
GroundFLAG = {f['GroundFLAG']}
SedFLAG = {1 if self.sediment else 0}
SedTransFLAG = {1 if (self.sediment and int(self.sediment.get("model", 2)) == 2) else 0}
[DIRS]
inputdir = {inputdir.rstrip('/')}/
outputdir = {outputdir.rstrip('/')}/
[TIMING]
startyear_timestep1 = {self.start.year}
startmonth_timestep1 = {self.start.month}
startday_timestep1 = {self.start.day}
startyear = {self.start.year}
startmonth = {self.start.month}
startday = {self.start.day}
endyear = {end.year}
endmonth = {end.month}
endday = {end.day}
spinupyears = 0
[GENERAL]
mask = clone.map
dem = dem.map
Slope = slope.map
locations = stations.map
[SOIL]
RootFieldMap = root_field.map
RootSatMap = root_sat.map
RootDryMap = root_dry.map
RootWiltMap = root_wilt.map
RootKsat = root_ksat.map
SubFieldMap = sub_field.map
SubSatMap = sub_sat.map
SubKsat = sub_ksat.map
[PEDOTRANSFER]
PedotransferFLAG = {1 if self.pedotransfer else 0}
RootSandMap = root_sand.map
RootClayMap = root_clay.map
RootOMMap = root_OM.map
RootBulkMap = 1
SubSandMap = sub_sand.map
SubClayMap = sub_clay.map
SubOMMap = sub_OM.map
SubBulkMap = sub_bulk.map
[SEDIMENT]
SedModel = {int((self.sediment or {}).get("model", 2))}
RockFrac = rock_fraction.map
exclChannelsFLAG = {int((self.sediment or {}).get("excl_channels", 1))}
upstream_km2 = {(self.sediment or {}).get("upstream_km2", 1.0)}
[MUSLE]
K_USLE = {"" if self.pedotransfer else "k_usle.map"}
P_USLE = {(self.sediment or {}).get("P_USLE", 1)}
musle_table = musle.tbl
[MMF]
MMF_table = mmf.tbl
harvestFLAG = {int((self.sediment or {}).get("harvest", 1))}
MMF_harvest = mmf_harvest.tbl
CanopyCoverLAIFlag = {int((self.sediment or {}).get("cc_lai", 0))}
{chr(10).join(f"{k} = {v!r}" for k, v in MMF_PARS.items())}
[SEDIMENT_TRANS]
Sed_init = 0
upstream_km2 = {(self.sediment or {}).get("upstream_km2", 1.0)}
{chr(10).join(f"{k} = {v!r}" for k, v in SEDTRANS_PARS.items())}
TrapEffTab = trapefftab.tbl
ResOrder = resorder.txt
[SOIL_INIT]
RootWater =
SubWater = {p['SubWater']}
CapRise = {p['CapRise']}
RootDrain = {p['RootDrain']}
SubDrain = {p['SubDrain']}
[SOIL_CAL]
RootFieldFrac = 1
RootSatFrac = 1
RootDryFrac = 1
RootWiltFrac = 1
RootKsatFrac = 1
[SOILPARS]
RootDepthFlat = {p['RootDepthFlat']}
SubDepthFlat = {p['SubDepthFlat']}
CapRiseMax = {p['CapRiseMax']}
SeepStatic = {0 if self.seep_series else 1}
SeePage = {"forcings/seep" if self.seep_series else p['SeePage']}
GWL_base = {p['GWL_base']}
[INFILTRATION]
Infil_excess = {self.infil_excess}
Alpha = {p['Alpha']}
Labda_infil = {p['Labda_infil']}
K_eff = {p['K_eff']}
PavedFrac = paved.tbl
[GROUNDW_PARS]
GwDepth = {p['GwDepth']}
GwSat = {p['GwSat']}
deltaGw = {p['deltaGw']}
BaseThresh = {p['BaseThresh']}
alphaGw = {p['alphaGw']}
YieldGw = {p['YieldGw']}
[GROUNDW_INIT]
GwRecharge = {p['GwRecharge']}
BaseR = {p['BaseR']}
Gw = {p['Gw']}
H_gw = {p['H_gw']}
[LANDUSE]
LandUse = landuse.map
KCstatic = {0 if self.kc_series else 1}
CropFac = kc.tbl
KC = {"forcings/kc" if self.kc_series else ""}
[DYNVEG]
NDVI = forcings/ndvi
NDVImax = {"ndvimax.map" if self.dynveg_maps else 0.65}
NDVImin = 0.1
NDVIbase = 0.15
KCmax = {"kcmax.map" if self.dynveg_maps else 1.5}
KCmin = 0.5
LAImax = laimax.tbl
FPARmax = {"fparmax.map" if self.dynveg_maps else 0.95}
FPARmin = 0.001
[PWS]
PWS_FLAG = {1 if self.pws else 0}
PFactor = depletion.tbl
[OPENWATER]
ETOpenWaterFLAG = {1 if self.open_water_frac is not None else 0}
kcOpenWater = {p['kcOpenWater']}
openWaterFrac = openwater_frac.map
[GLACIER]
GlacTable = glaciers.csv
ModelID = MOD_ID.map
GlacID = GLAC_ID.map
DDFG = {p['DDFG']}
DDFDG = {p['DDFDG']}
GlacF = {p['GlacF']}
GlacID_flag = {1 if self.glac_vars else 0}
GlacVars = {",".join(self.glac_vars) if self.glac_vars else "FRAC_GLAC,GlacMelt"}
GlacRetreat = {1 if self.glac_retreat else 0}
GlacID_memerror = {0 if self.glac_vars else 1}
GlacUpdate = {"%d,%d" % tuple(self.glac_retreat) if self.glac_retreat else "30,9"}
TLapse = tlapse.tbl
[SNOW]
TCrit = {p['Tcrit']}
SnowSC = {p['SnowSC']}
DDFS = {p['DDFS']}
SnowF = {p['SnowF']}
SnowCth = {p['SnowCth']}
[SNOW_INIT]
SnowIni = {p['SnowIni']}
SnowWatStore = {p['SnowWatStore']}
[CLIMATE]
Prec = forcings/prec
Pcorr_fact = {p['Pcorr_fact']}
precNetcdfFLAG = 0
Tair = forcings/tavg
Tcorr_fact = {p['Tcorr_fact']}
tempNetcdfFLAG = 0
[ETREF]
ETREF_FLAG = {1 if self.etref_input else 0}
ETref = forcings/etref
Lat = latitude.map
Gsc = {p['Gsc']}
Tmin = forcings/tmin
TminNetcdfFLAG = 0
Tmax = forcings/tmax
TmaxNetcdfFLAG = 0
[ROUTING]
flowdir = ldd.map
kx = {p['kx']}
Rout_components = 1
[ROUT_INIT]
{rout_init}[LAKE]
LakeId = lake_id.map
{"updatelakelevel = update_lakelevel.map" if self.lake_levels else ""}
{"LakeFile = forcings/llev" if self.lake_levels else ""}
LakeStor = lake_stor.tbl
LakeFunc = lake_function.tbl
LakeQH = lake_qh.tbl
LakeSH = lake_sh.tbl
LakeHS = lake_hs.tbl
[RESERVOIR]
ResId = res_id.map
reservoirs = reservoirs_nominal.map
ResFuncStor = res_funcstor.tbl
{'ResSimple = res_simple.tbl' if 'res_simple.tbl' in self.res_tables else ''}
{'ResAdv = res_adv.tbl' if 'res_adv.tbl' in self.res_tables else ''}
[REPORTING]
RepTab = reporting.csv
mm_rep_FLAG = 0
QTOT_mm_FLAG = 0
QSNOW_mm_FLAG = 0
QROOTR_mm_FLAG = 0
QROOTD_mm_FLAG = 0
QRAIN_mm_FLAG = 0
GMelt_mm_FLAG = 0
QGLAC_mm_FLAG = 0
QBASE_mm_FLAG = 0
Prec_mm_FLAG = 0
ETa_mm_FLAG = 0
Seep_mm_FLAG = 0
wbal_TSS_FLAG = 0
Lake_wbal = 0
Res_wbal = 0
RepSnow_FLAG = 0
RepRootR_FLAG = 0
RepRootD_FLAG = 0
RepRain_FLAG = 0
RepGlac_FLAG = 0
RepBase_FLAG = 0
"""
        del ground
        path = os.path.join(inputdir, "sphy_config.cfg")
        with open(path, "w") as fh:
            fh.write(cfg)
        return path


_SOIL_FILES = dict(RootFieldMap="root_field.map", RootSatMap="root_sat.map", RootDryMap="root_dry.map",
                   RootWiltMap="root_wilt.map", RootKsat="root_ksat.map", SubFieldMap="sub_field.map",
                   SubSatMap="sub_sat.map", SubKsat="sub_ksat.map")

# (reporting name, file prefix) pairs -- the names `reporting.reporting()` is called with on the hot path
# [MMF] / [SEDIMENT_TRANS] constants of the shipped sphy_config.cfg (sphy_config.cfg:694-753,889-911)
MMF_PARS = dict(PrecInt=30.0, K_c=0.1, K_z=0.5, K_s=0.3, DR_c=1.0, DR_z=1.6, DR_s=1.5, deltaClay=2e-6, deltaSilt=60e-6,
                deltaSand=200e-6, manning=0.015, depthBare=0.005, depthInField=0.1, depthTC=0.25, RFR=6.0, rho_s=2650.0,
                rho=1100.0, eta=0.0015)
SEDTRANS_PARS = dict(TC_beta=1.4, TC_gamma=1.4, manningChannelFLAG=1, manningChannel=0.045)

_REPORT_VARS = [
    ("LAI", "LAI"), ("TotIntF", "IntF"), ("TotPrecEF", "PrecEF"), ("StorCanop", "Canop"),
    ("wbal", "wbal"), ("TotPrec", "Prec"), ("TotPrecF", "PrecF"), ("TotRain", "Rain"), ("TotRainF", "RainF"),
    ("TotETref", "ETr"), ("TotETrefF", "ETrF"), ("TotETpot", "ETp"), ("TotETpotF", "ETpF"), ("TotETact", "ETa"),
    ("TotETactF", "ETaF"), ("PlantStress", "PStr"), ("TotSnow", "Snow"), ("TotSnowF", "SnowF"),
    ("TotSnowMelt", "SMel"), ("TotSnowMeltF", "SMelF"), ("TotGlacMelt", "GMel"), ("TotGlacR", "GlacR"),
    ("TotGlacPerc", "GlacP"), ("TotGlacRF", "GlacRF"), ("TotRootRF", "RootR"), ("TotRootDF", "RootD"),
    ("TotRootPF", "RootP"), ("TotSubDF", "SubD"), ("TotSubPF", "SubP"), ("TotCapRF", "CapR"),
    ("TotSeepF", "SeepF"), ("TotGwRechargeF", "Rchg"), ("TotRainRF", "RainR"), ("TotSnowRF", "SnowR"),
    ("TotBaseRF", "BaseR"), ("TotRF", "TotR"), ("Infil", "Infil"), ("StorRootW", "RootW"), ("StorSubW", "SubW"),
    ("StorGroundW", "GrndW"), ("StorSnow", "SnowS"), ("StorSnowW", "SnowW"), ("SCover", "SCov"), ("GWL", "GWL"),
    ("QallRAtot", "QAll"), ("RootRRAtot", "QRoR"), ("RootDRAtot", "QRoD"), ("RainRAtot", "QRain"),
    ("SnowRAtot", "QSnow"), ("GlacRAtot", "QGlac"), ("BaseRAtot", "QBase"),
    ("DetRn", "DetRn"), ("DetRun", "DetRun"), ("SDepFld", "SDpFld"), ("SedTrans", "STrans"), ("SedDep", "Sdep"),
    ("SedYld", "SdYld"), ("SedFlux", "SdFlux"), ("TC", "TC"),
]


def _name_t(t):
    """Suffix of PCRaster's stack naming for an <=4-letter prefix such as 'prec': prec0000.001."""
    s = "prec" + "0" * (11 - 4 - len(str(t))) + str(t)
    return (s[:8] + "." + s[8:])[4:]


# ----------------------------------------------------------------------------------------------
def make_catchment(nrows, ncols, cellsize=250.0, seed=1, nsteps=365, snow=False, glacier=False,
                   reservoirs=False, lakes=False, ground=True, routing=True, infil_excess=0,
                   rout_components=False, outlets="single", start=datetime.date(2001, 1, 1),
                   n_res_simple=0, n_res_adv=0, n_lakes=0, open_water=False, params=None,
                   report_daily=(), soil_classes=1, zrange=None, pws=False, report_opts=None, etref_input=False, mask=None, dynveg=False,
                   glac_retreat=None, glac_vars=(), sediment=False, kc_series=False,
                   seep_series=False, lake_levels=False, pedotransfer=False):
    """Build one of the BASELINE.json synthetic configurations."""
    p = dict(DEFAULT_PARAMS)
    if params:
        p.update(params)
    if zrange is None:
        zrange = (1500.0, 6500.0) if (snow or glacier) else None
    raw, filled, ldd, slope = make_dem_ldd(nrows, ncols, cellsize, seed=seed, outlets=outlets, zrange=zrange)
    lat = np.broadcast_to(np.linspace(27.0, 31.0, nrows, dtype=np.float32)[::-1, None], (nrows, ncols))
    flags = dict(GlacFLAG=int(glacier), SnowFLAG=int(snow or glacier), RoutFLAG=int(routing),
                 LakeFLAG=int(lakes), ResFLAG=int(reservoirs), GroundFLAG=int(ground or glacier))
    rng = np.random.default_rng([seed, 77])
    soil_maps = {}
    cls = (rng.integers(0, soil_classes, size=(nrows, ncols)) if soil_classes > 1 else np.zeros((nrows, ncols), int))
    scale = 1.0 + 0.05 * np.arange(soil_classes)
    for k in _SOIL_FILES:
        if soil_classes > 1:
            soil_maps[k] = (np.float32(p[k]) * scale[cls].astype(np.float32)).astype(np.float32)
        else:       # uniform soils: a zero-stride view, no memory (a 10000 x 10000 map would be 400 MB each)
            soil_maps[k] = np.broadcast_to(np.float32(np.float32(p[k]) * np.float32(scale[0])), (nrows, ncols))
    nup = upstream_counts(ldd)
    # stations: the outlet(s) with the largest upstream area + a few interior cells
    locations = np.zeros((nrows, ncols), np.int32)
    nflat = nup.ravel()
    picks = [int(np.argmax(nflat))]
    for frac in (0.1, 0.01, 0.001):               # first cell (row-major) draining at least that share of the raster
        hit = np.flatnonzero(nflat >= max(2, int(frac * nflat.size)))
        if hit.size:
            picks.append(int(hit[0]))
    for k, g in enumerate(dict.fromkeys(picks)):
        locations.ravel()[g] = k + 1
    cat = Catchment(nrows=nrows, ncols=ncols, cellsize=float(cellsize), seed=seed, dem=raw, ldd=ldd, slope=slope,
                    lat=lat, flags=flags, params=p, start=start, nsteps=nsteps, locations=locations,
                    soil_maps=soil_maps, rout_components=rout_components, infil_excess=infil_excess,
                    report_daily=tuple(report_daily), pws=bool(pws), report_opts=dict(report_opts or {}),
                    etref_input=bool(etref_input), dynveg=bool(dynveg), dynveg_maps=(dynveg == "maps"), kc_series=bool(kc_series),
                    seep_series=bool(seep_series), lake_levels=bool(lake_levels))
    if mask == "disk":                    # a non-rectangular catchment: everything outside the disk is MV
        rr, cc = np.ogrid[:nrows, :ncols]
        inside = ((rr - (nrows - 1) / 2.0) / (nrows / 2.0)) ** 2 + ((cc - (ncols - 1) / 2.0) / (ncols / 2.0)) ** 2 <= 1.0
        inside[nrows - 1, (ncols - 1) // 2] = True
        cat.mask = inside
        cat.ldd = np.where(inside, cat.ldd, 255).astype(np.uint8)
        cat.locations = np.where(inside, cat.locations, 0).astype(np.int32)
    if pedotransfer:                      # texture in percent, organic matter in % weight, bulk-density factor
        r3 = np.random.default_rng([seed, 606])
        tex = lambda mean, amp, lo, hi: np.clip(mean + amp * _smooth_noise(r3, nrows, ncols, 5), lo, hi).astype(np.float32)  # noqa: E731
        cat.pedotransfer = dict(root_sand=tex(45.0, 20.0, 15.0, 75.0), root_clay=tex(22.0, 8.0, 8.0, 20.0 + 15.0),
                                root_OM=tex(2.5, 1.0, 0.5, 5.0), sub_sand=tex(50.0, 20.0, 15.0, 80.0),
                                sub_clay=tex(20.0, 8.0, 5.0, 35.0), sub_OM=tex(1.0, 0.5, 0.1, 3.0),
                                sub_bulk=tex(1.02, 0.05, 0.9, 1.15))
    if sediment:
        _add_sediment(cat, seed, sediment if isinstance(sediment, dict) else {})
        if cat.pedotransfer is not None:  # one texture for both modules
            cat.sediment["sand"], cat.sediment["clay"] = cat.pedotransfer["root_sand"], cat.pedotransfer["root_clay"]
    if glacier:
        _add_glaciers(cat, rng)
        cat.glac_vars = tuple(glac_vars)
        if glac_retreat:
            cat.glac_retreat = tuple(glac_retreat)
            # thin ice on every fifth fragment so that some of them melt away at the update (glacier.py:713-719)
            r2 = np.random.default_rng([seed, 991])
            depth = cat.glac_table["ICE_DEPTH"].astype(np.float64)
            thin = r2.random(depth.size) < 0.2
            depth[thin] = r2.uniform(0.005, 0.12, size=int(thin.sum()))
            cat.glac_table["ICE_DEPTH"] = depth
    if reservoirs or lakes:
        _add_reservoirs_lakes(cat, rng, nup, n_res_simple if reservoirs else 0, n_res_adv if reservoirs else 0,
                              n_lakes if lakes else 0)
        if open_water:
            # water bodies of a few cells around every reservoir (fractions 0.3-0.9): clumps, some merging
            ow = np.zeros((nrows, ncols), np.float32)
            if cat.res_id is not None:
                for r, c in zip(*np.nonzero(cat.res_id > 0)):
                    r0, r1, c0, c1 = max(r - 1, 0), min(r + 2, nrows), max(c - 1, 0), min(c + 2, ncols)
                    blk = rng.uniform(0.3, 0.9, size=(r1 - r0, c1 - c0)).astype(np.float32)
                    blk[rng.random(blk.shape) < 0.4] = 0.0
                    ow[r0:r1, c0:c1] = np.maximum(ow[r0:r1, c0:c1], blk)
                    ow[r, c] = np.float32(rng.uniform(0.5, 0.9))
            cat.open_water_frac = ow
    if sediment and reservoirs:
        _add_sediment_reservoirs(cat, seed)
    return cat


def _add_glaciers(cat, rng):
    """Glacier fragments on cells above 5000 m: 1-3 fragments per cell."""
    z = cat.dem
    thresh = 5000.0 if z.max() > 5200.0 else np.quantile(z, 0.85)
    cells = np.flatnonzero(z.ravel() > thresh)
    cat.model_id = (np.arange(cat.ncell, dtype=np.int32) + 1).reshape(cat.nrows, cat.ncols)
    glac_id = np.zeros(cat.ncell, np.int32)
    cols = {k: [] for k in ("U_ID", "MOD_ID", "GLAC_ID", "MOD_H", "GLAC_H", "DEBRIS", "FRAC_GLAC", "ICE_DEPTH")}
    uid = 1
    for g in cells:
        nfrag = int(rng.integers(1, 4))
        gid = int(g // max(1, cat.ncols // 4) % 7 + 1)
        glac_id[g] = gid
        for _ in range(nfrag):
            cols["U_ID"].append(uid)
            uid += 1
            cols["MOD_ID"].append(int(g) + 1)
            cols["GLAC_ID"].append(gid)
            cols["MOD_H"].append(float(z.ravel()[g]))
            cols["GLAC_H"].append(float(z.ravel()[g] + rng.uniform(-200.0, 200.0)))
            cols["DEBRIS"].append(int(rng.random() < 0.3))
            cols["FRAC_GLAC"].append(float(rng.uniform(0.1, 0.5) / nfrag * 1.5))
            cols["ICE_DEPTH"].append(float(rng.uniform(20.0, 150.0)))
    # shuffle rows so the reference's sort_values("MOD_ID") (glacier.py:82) has work to do
    perm = rng.permutation(len(cols["U_ID"]))
    cat.glac_table = {k: np.asarray(v)[perm] for k, v in cols.items()}
    cat.glac_id = glac_id.reshape(cat.nrows, cat.ncols)
    cat.tlapse = np.full(12, -0.0065)


def _add_sediment(cat, seed, opts):
    """Inputs of the MMF erosion model and the sediment transport module (modules/mmf.py:107-190,
    modules/sediment_transport.py:113-214): three land-use classes (forest, cropland with a harvest window, bare /
    no-erosion patches), soil texture in percent, rock fraction."""
    r = np.random.default_rng([seed, 4242])
    nr, nc = cat.nrows, cat.ncols
    blobs = _smooth_noise(r, nr, nc, 4)
    lu = np.where(blobs > 0.35, 3, np.where(blobs > -0.2, 2, 1)).astype(np.int32)
    sand = np.clip(45.0 + 25.0 * _smooth_noise(r, nr, nc, 5), 10.0, 80.0).astype(np.float32)
    clay = np.clip(22.0 + 12.0 * _smooth_noise(r, nr, nc, 5), 5.0, np.minimum(45.0, 95.0 - sand)).astype(np.float32)
    rock = np.clip(0.12 + 0.15 * _smooth_noise(r, nr, nc, 3), 0.0, 0.6).astype(np.float32)
    #          class PlantH NoElem Diam   CC   GC  NoEros Till n_tab NoVeg
    mmf = [[0, 1, 2, 3, 4, 5, 6, 7, 8, 9],
           [1, 12.0, 0.8, 0.4, 0.85, 0.6, 0, 0, 0.0, 0],
           [2, 0.9, 150.0, 0.01, 0.5, 0.3, 0, 1, 0.0, 0],
           [3, 0.1, 0.0, 0.0, 0.0, 0.05, 0, 0, 0.03, 1]]
    #          class Sowing Harvest PlantH NoElem Diam  CC   GC  Till
    day0 = cat.start.timetuple().tm_yday
    harv = [[0, 1, 2, 3, 4, 5, 6, 7, 8],
            [1, 0, 0, 0.0, 0.0, 0.0, 0.0, 0.0, 0],
            [2, (day0 + 300) % 365 + 1, (day0 + 12) % 365 + 1, 0.05, 20.0, 0.005, 0.05, 0.1, 1],
            [3, 0, 0, 0.0, 0.0, 0.0, 0.0, 0.0, 0]]
    cat.sediment = dict(opts, landuse=lu, sand=sand, clay=clay, rock=rock)
    cat.sediment["mmf.tbl"] = mmf
    cat.sediment["mmf_harvest.tbl"] = harv
    #            class C-factor retardance N (modules/musle.py:115-119)
    cat.sediment["musle.tbl"] = [[0, 1, 2], [1, 0.003, 0.6], [2, 0.2, 0.2], [3, 0.45, 0.05]]
    cat.sediment["k_usle"] = np.clip(0.03 + 0.015 * _smooth_noise(r, nr, nc, 4), 0.005, 0.07).astype(np.float32)


def _add_sediment_reservoirs(cat, seed):
    """Reservoir inputs of the sediment transport module (modules/sediment_transport.py:138-188): the extent map, the
    trapping efficiencies and resorder.txt (reservoir id -> routing step; upstream reservoirs get earlier steps)."""
    r = np.random.default_rng([seed, 5151])
    rid = cat.res_id.ravel()
    cells = np.flatnonzero(rid > 0)
    dr = {1: (1, -1), 2: (1, 0), 3: (1, 1), 4: (0, -1), 6: (0, 1), 7: (-1, -1), 8: (-1, 0), 9: (-1, 1)}
    ldd = cat.ldd
    down = {}
    for g in range(cat.ncell):
        code = int(ldd.ravel()[g])
        if code in dr:
            rr, cc = divmod(g, cat.ncols)
            r2, c2 = rr + dr[code][0], cc + dr[code][1]
            if 0 <= r2 < cat.nrows and 0 <= c2 < cat.ncols:
                down[g] = r2 * cat.ncols + c2
    extent = np.zeros(cat.ncell, np.int32)
    for g in cells:                                   # the dam cell plus its direct donors
        extent[g] = rid[g]
        for u, d in down.items():
            if d == g and rid[u] == 0:
                extent[u] = rid[g]
    step = {}
    for g in cells:                                   # number of reservoirs further upstream decides the step
        step[int(rid[g])] = 0
    for g in cells:
        c = down.get(int(g))
        depth = 0
        while c is not None:
            if rid[c] > 0:
                depth += 1
                step[int(rid[c])] = max(step[int(rid[c])], depth)
            c = down.get(c)
    ids = sorted(step)
    cat.sediment["reservoirs"] = extent.reshape(cat.nrows, cat.ncols)
    cat.sediment["trap_eff"] = [(i, round(float(r.uniform(0.5, 0.95)), 3)) for i in ids]
    cat.sediment["res_order"] = [(i, step[i]) for i in ids]


def _add_reservoirs_lakes(cat, rng, nup, n_simple, n_adv, n_lakes):
    ncell = cat.ncell
    lo, hi = max(20, ncell // 4000), max(200, ncell // 40)
    cand = np.flatnonzero((nup.ravel() >= lo) & (nup.ravel() <= hi) & (cat.ldd.ravel() != 5))
    total = n_simple + n_adv + n_lakes
    if cand.size < total:
        raise ValueError("not enough candidate cells for reservoirs/lakes")
    chosen = rng.choice(cand, size=total, replace=False)
    area = cat.cellsize * cat.cellsize
    res_id = np.zeros(ncell, np.int32)
    lake_id = np.zeros(ncell, np.int32)
    funcstor = [[0, 1, 2]]
    simple = [[0, 1, 2, 3]]
    adv = [[0, 1, 2, 3, 4, 5, 6]]
    for k in range(n_simple + n_adv):
        g = chosen[k]
        rid = k + 1
        res_id[g] = rid
        inflow_mcm = nup.ravel()[g] * area * 0.002 / 1e6            # ~2 mm/day over the upstream area
        smax = 40.0 * inflow_mcm
        if k < n_simple:
            funcstor.append([rid, 1, 0.5 * smax])
            simple.append([rid, float(rng.uniform(0.02, 0.08)), 1.5, smax])
        else:
            funcstor.append([rid, 2, 0.6 * smax])
            adv.append([rid, smax, 0.1 * smax, 3.0 * inflow_mcm, 0.8 * inflow_mcm, int(rng.integers(100, 200)),
                        int(rng.integers(220, 300))])
    if n_simple + n_adv:
        cat.res_id = res_id.reshape(cat.nrows, cat.ncols)
        cat.res_tables["res_funcstor.tbl"] = funcstor
        if n_simple:
            cat.res_tables["res_simple.tbl"] = simple
        if n_adv:
            cat.res_tables["res_adv.tbl"] = adv
    if n_lakes:
        stor = [[0, 1]]
        func = [[0, 1, 2, 3]]
        qh = [[0, 1, 2, 3, 4, 5, 6]]
        sh = [[0, 1, 2, 3, 4, 5, 6]]
        hs = [[0, 1, 2, 3, 4, 5, 6]]
        for k in range(n_lakes):
            g = chosen[n_simple + n_adv + k]
            lid = k + 1
            lake_id[g] = lid
            inflow = nup.ravel()[g] * area * 0.002                   # m3/day
            s0 = 60.0 * inflow                                       # m3, reference storage
            ftype = k % 4 + 1
            stor.append([lid, s0 / 1e6])
            func.append([lid, ftype, ftype, ftype])
            # h(S): level ~ 2 m at S0; Q(h) in m3/s ~ inflow at h = 2
            qref = inflow / 86400.0
            qh.append([lid, qref / np.exp(1.6), 0.8, 0.0, qref / 2.0, qref / 8.0, qref / 32.0])
            sh.append([lid, s0 / np.exp(1.0), 0.5, 0.0, s0 / 2.0, s0 / 8.0, s0 / 32.0])
            hs.append([lid, 2.0 / np.exp(0.5), 0.5 / s0, 1.0, 1.0 / s0, 0.2 / s0 ** 2, 0.05 / s0 ** 3])
        cat.lake_id = lake_id.reshape(cat.nrows, cat.ncols)
        cat.lake_tables = {"lake_stor.tbl": stor, "lake_function.tbl": func, "lake_qh.tbl": qh, "lake_sh.tbl": sh,
                           "lake_hs.tbl": hs}
