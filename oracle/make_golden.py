"""oracle/make_golden.py -- TEST INFRASTRUCTURE ONLY: (re)generate tests/golden/*.npz.

Runs the UNMODIFIED reference (/root/reference/sphy.py, via oracle/run_reference.py + oracle/pcr_shim.py)
on small seeded synthetic catchments and stores its per-step state maps, daily reported flux maps and
the exact inputs it saw.  Must be run in the build container (needs /root/reference); the resulting
fixtures travel to the GPU box.

    python -m oracle.make_golden            # all cases
    python -m oracle.make_golden c2_glac    # one case
"""
from __future__ import annotations

import json
import os
import sys

import numpy as np

from sphy_b200 import synth

from .run_reference import run_reference

GOLDEN_DIR = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")

REPORT = ("LAI", "TotIntF", "TotPrecEF", "StorCanop", "TotRF", "TotETactF", "TotETref", "Infil", "wbal", "TotRootRF", "TotRootDF", "TotRootPF", "TotSubPF",
          "TotCapRF", "TotBaseRF", "TotSnowRF", "TotGlacRF", "TotGlacMelt", "TotSnowMelt", "StorSnow", "QallRAtot",
          "DetRn", "DetRun", "SDepFld", "SedTrans", "SedDep", "SedYld", "SedFlux", "TC")

# wetter / shallower than the cfg defaults so that saturation, runoff, drainage, percolation, baseflow and the
# ET cliff (ET.py:31) all occur inside a short run
WET = dict(SubWater=470.0, RootDepthFlat=150.0, Pscale=2.5)
WETK = dict(WET, RootKsat=12.0)          # rain often exceeds Ksat: the min(rain, Ksat, deficit) branches
COLD = dict(WETK, Toffset=9.0)
GLAC = dict(WETK, Toffset=15.0)

# name -> (make_catchment kwargs, nsteps)
CASES = {
    # BASELINE config 1: rootzone+subzone+groundwater+routing, snow/glacier off
    "c1_base": (dict(nrows=12, ncols=10, seed=11, params=WET), 60),
    # same with infiltration excess (the shipped cfg default) and all 7 routed components
    "c1_infil_comp": (dict(nrows=12, ncols=10, seed=12, infil_excess=1, rout_components=True, params=WET), 60),
    # groundwater module off: seepage + sub-drainage branch
    "c1_noground": (dict(nrows=12, ncols=10, seed=13, ground=False, params=WET), 40),
    # three soil classes: distributed soil parameters
    "c1_soilcls": (dict(nrows=12, ncols=10, seed=14, soil_classes=3, params=WETK), 40),
    # ETREF_FLAG = 1: ETref read from a map series instead of Hargreaves
    "c1_etref": (dict(nrows=12, ncols=10, seed=17, etref_input=True, params=WET), 40),
    # a non-rectangular catchment: MV outside the clone mask, cells at the rim drain out of the map
    "c1_masked": (dict(nrows=14, ncols=12, seed=18, mask="disk", rout_components=True, params=WET), 40),
    # dynamic vegetation: NDVI map series -> Kc, LAI, canopy interception (modules/dynamic_veg.py)
    "c1_dynveg": (dict(nrows=12, ncols=10, seed=19, dynveg=True, params=WET, start=(2001, 4, 1)), 50),
    # the same with NDVImax, KCmax and FPARmax read as maps (dynamic_veg.py:57-63 tries readmap first)
    "c1_dynveg_maps": (dict(nrows=12, ncols=10, seed=20, dynveg="maps", params=WET, start=(2001, 4, 1)), 40),
    "c2_dynveg_snow": (dict(nrows=12, ncols=10, seed=23, dynveg=True, snow=True, glacier=True, params=GLAC, start=(2001, 4, 15),
                            zrange=(1500.0, 5200.0)), 50),
    # crop-coefficient and seepage map series with missing days (sphy.py:824-830, 992-1000)
    "c1_kc_seep_series": (dict(nrows=12, ncols=10, seed=26, ground=False, kc_series=True, seep_series=True, params=WET), 30),
    "c1_kc_series": (dict(nrows=12, ncols=10, seed=27, kc_series=True, params=WET), 30),
    # soil hydraulic properties from texture (PedotransferFLAG = 1, utilities/pedotransfer.py), also feeding the erosion model
    "c1_pedotransfer": (dict(nrows=12, ncols=10, seed=28, pedotransfer=True, params=WET), 30),
    "c1_pedo_sed": (dict(nrows=12, ncols=10, seed=29, pedotransfer=True, sediment=True, params=WET, start=(2001, 6, 1)), 25),
    # plant water stress (ET.ks) instead of the wilting-point reduction
    "c1_pws": (dict(nrows=12, ncols=10, seed=15, pws=True, params=WET), 40),
    # reporting.csv driven outputs (utilities/reporting.py): month / year / final maps, averages, long-term sums, TSS
    "c1_reporting": (dict(nrows=10, ncols=9, seed=16, rout_components=True, params=WET, start=(2001, 11, 20),
                          report_opts={"TotRF": ("M+F", "Y", "D"), "TotETactF": ("Y", "M", "M"), "QallRAtot": ("D", "NONE", "D+Y"),
                                       "StorRootW": ("YS", "YA", "NONE"), "TotPrec": ("MS", "MA", "NONE"),
                                       "BaseRAtot": ("M", "NONE", "M")}), 75),
    # BASELINE config 2: snow on (spring melt season, snow line inside the catchment)
    "c2_snow": (dict(nrows=12, ncols=10, seed=21, snow=True, params=COLD, start=(2001, 3, 15), zrange=(1500.0, 5200.0)), 90),
    # BASELINE config 2: snow + glacier, all components routed
    "c2_glac": (dict(nrows=12, ncols=10, seed=22, snow=True, glacier=True, rout_components=True, params=GLAC,
                     start=(2001, 4, 15), zrange=(1500.0, 5200.0)), 90),
    # yearly glacier update (ice redistribution, fragments melting away) + per-glacier tables, modules/glacier.py:501-810
    "c2_glac_retreat": (dict(nrows=12, ncols=10, seed=24, snow=True, glacier=True, params=dict(WETK, Toffset=12.0), start=(2001, 9, 20),
                             zrange=(1500.0, 5200.0), glac_retreat=(30, 9),
                             glac_vars=("FRAC_GLAC", "GlacMelt", "GlacR", "TotalSnowStore_GLAC", "GLAC_T")), 25),
    # soil erosion (MMF, modules/mmf.py) + capacity-limited sediment transport (modules/sediment_transport.py), no reservoirs
    "c1_sed_mmf": (dict(nrows=14, ncols=12, seed=41, sediment=True, infil_excess=1, params=WET, start=(2001, 6, 1)), 40),
    "c2_sed_snow_veg": (dict(nrows=14, ncols=12, seed=42, sediment=dict(cc_lai=1, excl_channels=0, harvest=0), snow=True, dynveg=True,
                             params=COLD, start=(2001, 3, 15), zrange=(1500.0, 5200.0)), 40),
    # sediment trapped in reservoirs, sub-catchments routed in the steps of resorder.txt (sediment_transport.py:46-107)
    "c3_sed_reservoirs": (dict(nrows=18, ncols=16, seed=47, cellsize=500.0, sediment=dict(upstream_km2=2.0), infil_excess=1,
                               reservoirs=True, n_res_simple=3, n_res_adv=2, params=WET, start=(2001, 6, 1)), 25),
    # MUSLE soil loss (SedModel = 1, modules/musle.py), K-factor from a map / from texture (pedotransfer)
    "c1_sed_musle": (dict(nrows=12, ncols=10, seed=44, sediment=dict(model=1), infil_excess=1, params=WET, start=(2001, 6, 1)), 25),
    "c1_sed_musle_pedo": (dict(nrows=12, ncols=10, seed=45, sediment=dict(model=1, P_USLE=0.8), pedotransfer=True, infil_excess=1,
                               params=WET, start=(2001, 6, 1)), 25),
    # BASELINE config 3: reservoirs (simple + advanced) and lakes
    "c3_res_lake": (dict(nrows=16, ncols=14, seed=31, cellsize=1000.0, reservoirs=True, lakes=True, n_res_simple=2,
                         n_res_adv=2, n_lakes=4, params=WET, start=(2001, 5, 1)), 70),
    # open-water evaporation / precipitation on the water bodies (advanced_routing.py:167-196)
    "c3_openwater": (dict(nrows=16, ncols=14, seed=34, cellsize=1000.0, reservoirs=True, lakes=True, n_res_simple=2, n_res_adv=1,
                          n_lakes=2, open_water=True, params=WET, start=(2001, 6, 1)), 50),
    "c3_res_only": (dict(nrows=14, ncols=12, seed=32, cellsize=1000.0, reservoirs=True, n_res_simple=2, n_res_adv=1,
                         params=WETK), 40),
    # measured lake levels replace the storage of gauged lakes on the days a level map exists (lakes.py:28-97)
    "c3_lake_levels": (dict(nrows=16, ncols=14, seed=35, cellsize=1000.0, reservoirs=True, lakes=True, n_res_simple=1, n_res_adv=1,
                            n_lakes=4, lake_levels=True, params=WET, start=(2001, 5, 1)), 30),
    "c3_lake_only": (dict(nrows=14, ncols=12, seed=33, cellsize=1000.0, lakes=True, n_lakes=4, params=WET), 40),
}


def build_case(name):
    import datetime
    kw, nsteps = CASES[name]
    kw = dict(kw)
    if "start" in kw:
        kw["start"] = datetime.date(*kw["start"])
    daily = () if "report_opts" in kw else REPORT
    return synth.make_catchment(nsteps=nsteps, report_daily=daily, **kw), nsteps


def make(name):
    cat, nsteps = build_case(name)
    ref = run_reference(cat, nsteps, keep_table=bool(cat.glac_retreat or cat.glac_vars))
    fkeys = ("prec", "tavg", "tmin", "tmax") + (("etref",) if cat.etref_input else ()) + (("ndvi",) if cat.dynveg else ())
    forcing = {k: np.stack([cat.forcing(t)[k] for t in range(1, nsteps + 1)]) for k in fkeys}
    nan_map = np.full((cat.nrows, cat.ncols), np.nan, np.float32)           # series with gaps: NaN map = no file that day
    for k, on in (("kc", cat.kc_series), ("seep", cat.seep_series), ("llev", cat.lake_levels)):
        if on:
            forcing[k] = np.stack([cat.forcing(t).get(k, nan_map) for t in range(1, nsteps + 1)])
    payload = {"meta": np.array(json.dumps(dict(case=name, nsteps=nsteps, kwargs=CASES[name][0])))}
    for k, v in cat.static_maps().items():
        payload["map/" + k] = v
    for k, v in forcing.items():
        payload["forcing/" + k] = v
    for k, v in ref["states"].items():
        payload["state/" + k] = v
    for k, v in ref["reports"].items():
        payload["report/" + k] = v
    if CASES[name][0].get("report_opts"):
        for k, v in ref["report_files"].items():
            payload["repfile/" + k] = v
        for k, v in ref["tss_tables"].items():
            payload["tss/" + k] = v
    if cat.glac_retreat or cat.glac_vars:
        import io
        import pandas as pd
        for k, v in ref["report_files"].items():                  # GlacFrac_/GlacArea_/iceDepth_/vIce_<date>.map
            if k.endswith(".map"):
                payload["repfile/" + k] = v
        for k, text in ref["csv_files"].items():
            df = pd.read_csv(io.StringIO(text), index_col=None if k.startswith("glacTable") else 0)
            payload["csv/" + k] = df.to_numpy(np.float64)
            payload["csvcols/" + k] = np.asarray([str(c) for c in df.columns])
        tabs = ref["glac_table"][1:]                               # after every dynamic() step
        payload["glac/U_ID"] = tabs[0]["U_ID"].to_numpy(np.int64)  # the reference's row order (sort_values is not stable)
        payload["glac/FRAC_GLAC"] = np.stack([t["FRAC_GLAC"].to_numpy(np.float64) for t in tabs])
        payload["glac/ICE_DEPTH"] = np.stack([t["ICE_DEPTH"].to_numpy(np.float64) for t in tabs])
    os.makedirs(GOLDEN_DIR, exist_ok=True)
    path = os.path.join(GOLDEN_DIR, name + ".npz")
    np.savez_compressed(path, **payload)
    # coverage summary: make sure the case exercises the branches it is meant to pin
    st, rp = ref["states"], ref["reports"]
    cov = {
        "RootDrain>0": float((st["RootDrain"][1:] > 0).mean()),
        "report files": len(ref["report_files"]) if CASES[name][0].get("report_opts") else None,
        "tss": sorted(ref["tss_tables"]) if CASES[name][0].get("report_opts") else None,
        "runoff>0": float((rp["RootR"] > 0).mean()) if "RootR" in rp else None,
        "perc>0": float((rp["RootP"] > 0).mean()) if "RootP" in rp else None,
        "subperc>0": float((rp["SubP"] > 0).mean()) if "SubP" in rp else None,
        "BaseR>0": float((st["BaseR"][1:] > 0).mean()) if "BaseR" in st else None,
        "ETa==0": float((rp["ETaF"] == 0).mean()) if "ETaF" in rp else None,
        "Snow>0": float((st["SnowStore"][1:] > 0).mean()) if "SnowStore" in st else None,
        "GlacMelt>0": float((rp["GMel"] > 0).mean()) if "GMel" in rp else None,
        "StorRES range": [float(st["StorRES"][1:].max()), float(st["StorRES"][-1].max())] if "StorRES" in st else None,
    }
    print(name, os.path.getsize(path) // 1024, "KiB", {k: (round(v, 3) if isinstance(v, float) else v)
                                                       for k, v in cov.items() if v is not None})


if __name__ == "__main__":
    for nm in (sys.argv[1:] or list(CASES)):
        make(nm)
