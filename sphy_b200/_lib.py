"""ctypes binding of libsphy_b200.so (the C ABI in include/sphy_b200.h).

There is no fallback: if the CUDA library cannot be loaded, or a call fails, a RuntimeError is raised.
"""
from __future__ import annotations

import ctypes as C
import os

from . import _build

P = C.c_void_p


class SphyParam(C.Structure):
    _fields_ = [("map", P), ("value", C.c_float)]


class WbParams(C.Structure):
    _fields_ = [
        ("flags", C.c_uint32), ("ra_mode", C.c_int32),
        ("lat_class", P), ("lat_unique_deg", P), ("n_lat_unique", C.c_int32),
        ("sin_lat", P), ("cos_lat", P), ("tan_lat", P),
        ("k0", C.c_float), ("pcorr", C.c_float), ("tcorr", C.c_float),
        ("root_field", SphyParam), ("root_sat", SphyParam), ("root_dry", SphyParam), ("root_wilt", SphyParam),
        ("root_ksat", SphyParam), ("root_drainvel", SphyParam), ("e_root", SphyParam),
        ("sub_field", SphyParam), ("sub_sat", SphyParam), ("sub_drainvel", SphyParam), ("e_sub", SphyParam),
        ("caprise_max", SphyParam), ("kc", SphyParam), ("pmap", SphyParam), ("glac_frac", SphyParam),
        ("open_water_frac", SphyParam), ("kc_open_water", C.c_float),
        ("k_eff", C.c_float), ("alpha", C.c_float), ("alpha_sq", C.c_float), ("labda_infil", C.c_float),
        ("paved_frac", SphyParam),
        ("gw_sat", SphyParam), ("base_thresh", SphyParam), ("e_delta", SphyParam), ("e_alpha", SphyParam),
        ("h_den", SphyParam), ("seepage", SphyParam),
        ("tcrit", SphyParam), ("snow_sc", SphyParam), ("ddfs", SphyParam), ("snow_f", SphyParam),
        ("one_minus_snow_f", C.c_float), ("melt_cos", C.c_float * 24),
    ]


class WbState(C.Structure):
    _fields_ = [(k, P) for k in ("RootWater", "SubWater", "CapRise", "RootDrain", "GwRecharge", "Gw", "BaseR", "H_gw",
                                 "SubDrain", "SnowStore", "SnowWatStore", "TotalSnowStore", "waterbalanceTot")]


class WbForcing(C.Structure):
    _fields_ = [(k, P) for k in ("prec", "tavg", "tmin", "tmax", "etref", "TotalSnowStore_GLAC", "SnowR_GLAC",
                                 "GlacPerc", "GlacR", "wbal_old_gw", "wbal_snow_adjust", "wbal_ice_delta", "wbal_glac_frac")]


class WbOut(C.Structure):
    _fields_ = [(k, P) for k in ("TotR", "RootRR", "RootDR", "RainR", "SnowR", "ETref", "ETact", "ActETact", "Infil",
                                 "Rain", "RootPerc", "SubPerc", "wbal", "ETOpenWater", "Precip")]


class GlacTable(C.Structure):
    _fields_ = ([("n_rows", C.c_int32), ("n_seg", C.c_int32), ("cell", P), ("seg_ptr", P), ("seg_cell", P),
                 ("MOD_H", P), ("GLAC_H", P), ("FRAC_GLAC", P), ("DEBRIS", P), ("Rain_GLAC", P), ("Snow_GLAC", P),
                 ("SnowStore_GLAC", P), ("SnowWatStore_GLAC", P), ("TotalSnowStore_GLAC", P), ("AccuGlacMelt", P),
                 ("GlacMelt", P), ("GlacR", P), ("GlacPerc", P), ("SnowR_GLAC", P), ("ActSnowMelt_GLAC", P),
                 ("GLAC_T", P)]
                + [(k, C.c_double) for k in ("tcrit", "ddfs", "snow_sc", "ddfg", "ddfdg", "glac_f", "lapse")])


class GlacOut(C.Structure):
    _fields_ = [(k, P) for k in ("Rain_GLAC", "Snow_GLAC", "ActSnowMelt_GLAC", "TotalSnowStore_GLAC", "SnowR_GLAC",
                                 "GlacMelt", "GlacR", "GlacPerc")]


class ResLake(C.Structure):
    _fields_ = [("n_struct", C.c_int32), ("cell", P), ("kind", P), ("par", P), ("qfrac", P), ("StorRES", P),
                ("Qout", P), ("Qin", P), ("lake_level", P), ("update_level", P)]


class ResLakePlan(C.Structure):
    _fields_ = [(k, P) for k in ("cut_slot", "entry_cut", "entry_ptr", "entry_idx", "x_rank", "x_local", "qmask", "material", "qday",
                                 "extra", "top")]


class OpenWater(C.Structure):
    _fields_ = [("ow_ptr", P), ("ow_cell", P), ("open_water_frac", P), ("et_open_water", P), ("prec", P), ("pcorr", C.c_float)]


class Veg(C.Structure):
    _fields_ = ([(k, P) for k in ("ndvi", "prec", "tavg", "tmin", "tmax", "etref_in")]
                + [("lai_max", SphyParam), ("open_water_frac", SphyParam)]
                + [(k, SphyParam) for k in ("sr_min", "sr_range", "fpar_range", "log_one_m_fparmax", "kc_min", "kc_range", "ndvi_min",
                                             "ndvi_range")]
                + [(k, P) for k in ("Scanopy", "ETref", "Kc", "PrecipEff", "LAI", "Int")])


class Mmf(C.Structure):
    _fields_ = ([("prec", P), ("pcorr", C.c_float), ("q", P), ("q_is_discharge", C.c_int32), ("total_snow_store", P), ("lai", P),
                 ("cc_mode", C.c_int32), ("harvest_on", C.c_int32), ("excl_channels", C.c_int32)]
                + [(k, SphyParam) for k in ("cos_slope", "sin_slope_pow", "cc_table", "gc_table", "plant_height", "no_erosion",
                                            "rock_frac", "sowing", "harvest", "cc_harvest", "gc_harvest", "plant_height_harvest",
                                            "clay", "silt", "sand", "hillslope", "v_field", "v_field_harvest", "roughness",
                                            "roughness_harvest", "slope_streams_pow")]
                + [("K", C.c_float * 3), ("DR", C.c_float * 3), ("v_s", C.c_float * 3)]
                + [(k, C.c_float) for k in ("infil_alpha", "ke_const", "d_field", "cell_length", "tc_beta")]
                + [(k, P) for k in ("DetRn", "DetRun", "SDepFld", "SedTrans", "TC")])


class SedRes(C.Structure):
    _fields_ = ([("n_res", C.c_int32), ("n_steps", C.c_int32)]
                + [(k, P) for k in ("res_cell", "res_down", "res_step", "down_finish_seq", "trap_eff", "outflow_eff", "finish_step",
                                    "extent", "step_last_seq_host", "sed_trans", "tc_route", "cap_flux", "cap_state", "outflow")])


RL_NPAR = 30
RL_EXTRA = 9                          # float64 per structure at the end of an exchange block (include/sphy_b200.h)
RL_OP_UPDATE_LAKE, RL_OP_QLAKE, RL_OP_QSIMPLE, RL_OP_QADV, RL_OP_QRES = range(5)
WB_SNOW, WB_GROUND, WB_INFIL_EXCESS, WB_PWS, WB_ETREF_INPUT = 1, 2, 4, 8, 16
RA_TABLE, RA_CELL = 0, 1
ROUTE_LEVELS, ROUTE_SCAN = 0, 1
ROUTE_MAX_COMP = 8                    # routed components per call (csrc/common.cuh)
ORDER_ROWMAJOR, ORDER_DFS = 0, 1
SCAN_TILE = 1024                      # cells per scan-routing tile (csrc/route_scan.cu)
OPS = ("EXTRARAD", "HARGREAVES", "ETPOT", "ETACT", "KS", "POTSNOWMELT", "ACTSNOWMELT", "SNOWSTOREUPDATE", "MAXSNOWWATSTORAGE",
       "SNOWWATSTORAGE", "TOTSNOWSTORAGE", "SNOWR", "ROOTRUNOFF_SAT", "ROOTRUNOFF_INFIL", "ROOTDRAINAGE", "ROOTPERCOLATION",
       "CALCFRAC", "CAPILRISE", "SUBPERCOLATION", "SUBDRAINAGE", "GWRECHARGE", "BASEFLOW", "HLEVEL", "SNOW_DYNAMIC",
       "GROUNDWATER_DYNAMIC", "MUSLE", "ADV_BLEND", "ADV_FINISH")
OP = {name: i for i, name in enumerate(OPS)}
UNARY_EXP_NEG_INV, UNARY_SIN_DEG, UNARY_COS_DEG, UNARY_TAN_DEG, UNARY_EXP_NEG = 0, 1, 2, 3, 4

# every symbol include/sphy_b200.h declares: name -> (restype, argtypes)
SYMBOLS = {
    "sphy_ctx_create": (C.c_int, [C.POINTER(P), C.c_int, C.c_int, C.c_int, C.c_float]),
    "sphy_ctx_destroy": (None, [P]),
    "sphy_last_error": (C.c_char_p, [P]),
    "sphy_sm_count": (C.c_int, [P]),
    "sphy_abi_sizeof": (C.c_size_t, [C.c_int]),
    "sphy_set_cell_order": (C.c_int, [P, C.c_int]),
    "sphy_ldd_build": (C.c_int, [P, P, P, P]),
    "sphy_ncells": (C.c_int64, [P]),
    "sphy_nlevels": (C.c_int32, [P]),
    "sphy_ldd_export": (C.c_int, [P, C.c_char_p, P, C.c_size_t, P]),
    "sphy_gather_f32": (C.c_int, [P, P, P, P]),
    "sphy_gather_indexed_f32": (C.c_int, [P, P, P, P, C.c_int64, P, P]),
    "sphy_interp_linear": (C.c_int, [P, P, P, P, C.c_int64, P, P]),
    "sphy_scatter_f32": (C.c_int, [P, P, P, C.c_float, P]),
    "sphy_sample": (C.c_int, [P, P, P, C.c_int, P, P]),
    "sphy_map_accumulate": (C.c_int, [P, P, P, C.c_float, C.c_int64, P]),
    "sphy_map_divide": (C.c_int, [P, P, C.c_float, P, C.c_int64, P]),
    "sphy_map_unary": (C.c_int, [P, C.c_int, P, P, C.c_int64, P]),
    "sphy_set_ncells": (C.c_int, [P, C.c_int64]),
    "sphy_module_op": (C.c_int, [P, C.c_int, C.POINTER(SphyParam), C.c_int, C.POINTER(P), C.c_int, C.c_int64, C.c_float, C.c_int, P]),
    "sphy_scan_tile_cells": (C.c_int, []),
    "sphy_set_overlap_tail": (C.c_int, [P, C.c_int, P]),
    "sphy_route_profile": (C.c_int, [P, C.c_int]),
    "sphy_route_profile_read": (C.c_int, [P, C.POINTER(C.c_float), C.POINTER(C.c_int)]),
    "sphy_p2p_setup_local": (C.c_int, [P, C.c_int, C.c_int64, P, P]),
    "sphy_p2p_connect": (C.c_int, [P, C.c_int, C.c_int, P, P]),
    "sphy_p2p_timed_out": (C.c_int, [P, P]),
    "sphy_mmf_step": (C.c_int, [P, C.POINTER(Mmf), C.c_int, P]),
    "sphy_accucapacity": (C.c_int, [P, P, P, P, P, P]),
    "sphy_sed_reservoir_route": (C.c_int, [P, C.POINTER(SedRes), P, P, P, P, P, P]),
    "sphy_glac_report": (C.c_int, [P, C.c_int32, P, P, P, C.c_int32, P, P, P, P]),
    "sphy_veg_step": (C.c_int, [P, C.POINTER(WbParams), C.POINTER(Veg), C.c_int, P]),
    "sphy_day_scalars": (C.c_int, [C.c_int, C.c_float, C.POINTER(C.c_float)]),
    "sphy_set_day_block": (C.c_int, [P, P]),
    "sphy_day_table_step": (C.c_int, [P, P, C.c_int, P, P, P]),
    "sphy_wb_step": (C.c_int, [P, C.POINTER(WbParams), C.POINTER(WbState), C.POINTER(WbForcing), C.POINTER(WbOut),
                               C.c_int, P]),
    "sphy_glac_step": (C.c_int, [P, C.POINTER(GlacTable), P, P, C.c_float, C.c_float, C.POINTER(GlacOut), P]),
    "sphy_route_step": (C.c_int, [P, C.c_int, C.POINTER(P), C.POINTER(P), SphyParam, C.c_float, C.c_int, P]),
    "sphy_route_step_members": (C.c_int, [P, C.c_int, C.POINTER(P), C.POINTER(P), C.POINTER(SphyParam), C.POINTER(C.c_float), P]),
    "sphy_set_partition": (C.c_int, [P, C.c_int64, C.c_int64, P]),
    "sphy_trunk_select": (C.c_int, [P, C.c_int32, C.c_int64, C.c_int, P, C.POINTER(C.c_int64), P]),
    "sphy_trunk_export": (C.c_int, [P, C.c_char_p, P, C.c_size_t, P]),
    "sphy_trunk_set_plan": (C.c_int, [P, C.c_int32, P, P, P, P, P, C.c_int32, C.c_int, C.c_int, P]),
    "sphy_route_scan_begin": (C.c_int, [P, C.c_int, C.POINTER(P), C.POINTER(P), SphyParam, C.c_float, P, C.c_int32, P]),
    "sphy_route_scan_finish": (C.c_int, [P, C.c_int, C.POINTER(P), SphyParam, C.c_float, P, C.c_int32, C.c_int, C.c_int, P]),
    "sphy_accuflux": (C.c_int, [P, P, P, P, C.c_int, P]),
    "sphy_upstream": (C.c_int, [P, P, P, P]),
    "sphy_reslake_openwater": (C.c_int, [P, C.POINTER(ResLake), C.POINTER(OpenWater), P]),
    "sphy_reslake_op": (C.c_int, [P, C.POINTER(ResLake), C.c_int, C.c_int, P, P, P]),
    "sphy_route_reslake_step": (C.c_int, [P, C.POINTER(ResLake), P, P, SphyParam, C.c_float, C.c_int, C.c_int, P]),
    "sphy_trunk_select_cuts": (C.c_int, [P, C.c_int32, P, C.c_int32, C.c_int64, C.c_int, P, C.POINTER(C.c_int64), P]),
    "sphy_cell_donors": (C.c_int, [P, C.c_int32, P, P, P]),
    "sphy_route_reslake_begin": (C.c_int, [P, C.POINTER(ResLake), C.POINTER(ResLakePlan), P, P, SphyParam, C.c_float, P, C.c_int32, P]),
    "sphy_route_reslake_finish": (C.c_int, [P, C.POINTER(ResLake), C.POINTER(ResLakePlan), P, SphyParam, C.c_float, P, C.c_int32,
                                            C.c_int, C.c_int, C.c_int, C.c_int, P]),
}

ABI_STRUCTS = (SphyParam, WbParams, WbState, WbForcing, WbOut, GlacTable, GlacOut, ResLake, OpenWater, Veg, Mmf, SedRes, ResLakePlan)

_LIB = None


def library_path():
    return _build.CUDA_LIB


def load(build=True):
    """Load libsphy_b200.so, building it in-tree first if needed.  Raises if that is impossible."""
    global _LIB
    if _LIB is not None:
        return _LIB
    path = library_path()
    if build:
        path = _build.build_cuda()
    if not os.path.exists(path):
        raise RuntimeError(f"{path} is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                           "(sphy_b200 has no CPU fallback)")
    lib = C.CDLL(path)
    for name, (res, args) in SYMBOLS.items():
        fn = getattr(lib, name)          # AttributeError here = header and library disagree
        fn.restype = res
        fn.argtypes = args
    _LIB = lib
    return lib


class SphyError(RuntimeError):
    pass


def check(lib, ctx, rc, what):
    if rc != 0:
        msg = lib.sphy_last_error(ctx).decode() if ctx else ""
        raise SphyError(f"{what} failed with code {rc}: {msg}")
