/*
 * sphy_b200.h -- C ABI of the B200-native SPHY hot path (libsphy_b200.so, sm_100a only).
 *
 * The reference (FutureWater/SPHY v3.1) has no FFI: its operator boundary is the `pcr` namespace that
 * every process module receives as an argument plus the initial()/dynamic() contract.  Each entry point
 * below names the reference interface it replaces (file:line relative to /root/reference).
 *
 * Conventions
 *   - returns 0 on success, a negative SPHY_E_* code otherwise; never throws, never exits;
 *     sphy_last_error() gives the message of the last failure on that context.
 *   - every data buffer is DEVICE memory owned by the caller (torch tensors on the Python side);
 *     the context owns only its LDD structures and workspaces.
 *   - calls are asynchronous on the passed cudaStream_t (void*); no hidden synchronisation and no
 *     allocation after sphy_ldd_build() except where stated.
 *   - one context per GPU, not thread-safe per context (the reference is single-threaded).
 *   - all maps are float32 arrays over the N active cells in compact row-major order
 *     (SURVEY.md Appendix C); "scalar-or-map" parameters mirror the reference's
 *     `try readmap / except getfloat` idiom (e.g. modules/groundwater.py:51-57).
 */
#ifndef SPHY_B200_H
#define SPHY_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define SPHY_OK 0
#define SPHY_E_INVALID (-1)   /* bad argument */
#define SPHY_E_CUDA (-2)      /* CUDA runtime error */
#define SPHY_E_CYCLE (-3)     /* unsound LDD (cycle) */
#define SPHY_E_STATE (-4)     /* call order (e.g. routing before sphy_ldd_build) */
#define SPHY_E_NOMEM (-5)

typedef struct sphy_ctx sphy_ctx;

/* scalar-or-map parameter: map != NULL -> per-cell values, else the broadcast scalar */
typedef struct { const float *map; float value; } sphy_param;

/* ---- context ------------------------------------------------------------------------------ */
/* replaces pcr.setclone / the clone map (sphy.py:141-145); cellsize feeds pcr.cellarea() */
int sphy_ctx_create(sphy_ctx **out, int device, int nrows, int ncols, float cellsize);
void sphy_ctx_destroy(sphy_ctx *ctx);
const char *sphy_last_error(const sphy_ctx *ctx);
int sphy_sm_count(const sphy_ctx *ctx);
/* sizeof() of the ABI structs, so foreign-language bindings can verify their layout without a GPU:
 * 0 sphy_param, 1 sphy_wb_params, 2 sphy_wb_state, 3 sphy_wb_forcing, 4 sphy_wb_out, 5 sphy_glac_table,
 * 6 sphy_glac_out, 7 sphy_reslake, 8 sphy_openwater, 9 sphy_veg, 10 sphy_mmf, 11 sphy_sedres, 12 sphy_reslake_plan */
size_t sphy_abi_sizeof(int which);

/* ---- K3: LDD -> integer structures (once per run) ------------------------------------------
 * replaces what pcr.accuflux/accufractionflux/upstream derive implicitly from the UINT1 LDD map read at
 * modules/routing.py:34.  ldd_dev: nrows*ncols keypad codes (1-9, 5 = pit, anything else = MV);
 * active_dev: optional nrows*ncols mask (clone, sphy.py:143).  Synchronises the stream (one-time build). */
int sphy_ldd_build(sphy_ctx *ctx, const uint8_t *ldd_dev, const uint8_t *active_dev, void *stream);
/* Order of the N cells in every compact map (call before sphy_ldd_build).  ROWMAJOR = the canonical compact
 * numbering; DFS (default) = depth-first pre-order of the drainage forest, so that a cell's upstream area is the
 * contiguous index range [j, j + sub_size[j]) -- what SPHY_ROUTE_SCAN streams over.  The water balance is per-cell
 * and indifferent to the order; raster <-> compact conversion goes through sphy_gather_f32 / sphy_scatter_f32 or
 * the "flat_active" export (device index -> raster index). */
#define SPHY_ORDER_ROWMAJOR 0
#define SPHY_ORDER_DFS 1
int sphy_set_cell_order(sphy_ctx *ctx, int mode);
int64_t sphy_ncells(const sphy_ctx *ctx);
int32_t sphy_nlevels(const sphy_ctx *ctx);
/* copy one structure to a caller device buffer for bit-exact checks.  name / element type / count
 * (canonical compact row-major numbering, SURVEY.md Appendix C, whatever the device cell order is):
 *   "cidx" i32 nrows*ncols | "down" i32 N | "indeg" u8 N | "level" i32 N | "order" i32 N |
 *   "level_ptr" i32 L+1 | "don_ptr" i32 N+1 | "don_idx" i32 don_ptr[N] | "nup" i64 N
 * and the device-order maps: "flat_active" i32 N (device index -> raster index), "cell_order" i32 N (device
 * index -> canonical index), "sub_size" i32 N (upstream cell count by device index) */
int sphy_ldd_export(sphy_ctx *ctx, const char *name, void *dst_dev, size_t dst_bytes, void *stream);

/* raster <-> compact (readmap / report side of the boundary): out[k] = raster[cell k] and back */
int sphy_gather_f32(sphy_ctx *ctx, const float *raster_dev, float *compact_dev, void *stream);
int sphy_scatter_f32(sphy_ctx *ctx, const float *compact_dev, float *raster_dev, float fill, void *stream);
/* NetCDF forcing (utilities/netcdf2PCraster.py:167-197): the daily scipy.interpolate.griddata(method="linear") of the coarse
 * forcing grid onto the model cells as one gather -- out[i] = REAL4(sum_k w3[3i+k] * z[idx3[3i+k]]) in float64, with the
 * triangle and barycentric weights of every cell computed once on the host (they do not change from day to day);
 * idx3[3i] < 0 = outside the convex hull -> NaN (the reference's -9999 / MV). */
int sphy_interp_linear(sphy_ctx *ctx, const double *z_dev, const int32_t *idx3_dev, const double *w3_dev, int64_t count,
                       float *out_dev, void *stream);
/* out[dst_idx[e]] = src[src_idx[e]], e < count, src_idx ascending.  `src` may be a PINNED HOST raster (unified addressing):
 * the device then pulls only the cells it owns out of the host raster (pcr.readmap's result, sphy.py:756) over PCIe, which
 * is what lets a partitioned basin ingest rasters at the aggregate host->device bandwidth of the box. */
int sphy_gather_indexed_f32(sphy_ctx *ctx, const float *src, const int32_t *src_idx_dev, const int32_t *dst_idx_dev,
                            int64_t count, float *out_dev, void *stream);
/* reporting accumulators (utilities/reporting.py:24-95): acc = acc + x (divisor == 0) or acc + x/divisor, and
 * out = in/divisor, each one REAL4 operation per cell like the reference's `tot + var`, `var / dim`, `tot / ydays` */
int sphy_map_accumulate(sphy_ctx *ctx, float *acc, const float *x, float divisor, int64_t count, void *stream);
int sphy_map_divide(sphy_ctx *ctx, const float *in, float divisor, float *out, int64_t count, void *stream);
/* TimeoutputTimeseries.sample (sphy.py:683-691): out[j] = map[cells[j]] */
int sphy_sample(sphy_ctx *ctx, const float *map_dev, const int32_t *cells_dev, int n, float *out_dev, void *stream);

/* ---- K1: fused vertical water balance of one daily step -------------------------------------
 * replaces sphy.dynamic() sphy.py:732-1097 + 1114-1212, i.e. hargreaves.extrarad/Hargreaves
 * (hargreaves.py:24-47), ET.ETpot/ETact/ks (ET.py:24-47), snow.dynamic (modules/snow.py:110-187),
 * rootzone.RootRunoff/RootDrainage/RootPercolation/CalcFrac (rootzone.py:24-94), subzone.CapilRise/
 * SubPercolation/SubDrainage (subzone.py:24-51) and groundwater.dynamic (modules/groundwater.py:90-125). */
#define SPHY_WB_SNOW 1u          /* SnowFLAG (sphy.py:869) */
#define SPHY_WB_GROUND 2u        /* GroundFLAG (sphy.py:1033,1080) */
#define SPHY_WB_INFIL_EXCESS 4u  /* InfilFLAG == 1 (rootzone.py:26) */
#define SPHY_WB_PWS 8u           /* PlantWaterStressFLAG (sphy.py:919) */
#define SPHY_WB_ETREF_INPUT 16u  /* ETREF_FLAG == 1: forcing.etref is read instead of Hargreaves (sphy.py:814) */

#define SPHY_RA_TABLE 0          /* Ra looked up by latitude class (lat_class u16 + per-day table) */
#define SPHY_RA_CELL 1           /* Ra from per-cell sin/cos/tan(latitude) maps */

typedef struct {
    uint32_t flags;
    int32_t ra_mode;
    /* latitude: TABLE mode: class map + unique latitudes in degrees (the kernel forms Lat*REAL4(pi/180)) */
    const uint16_t *lat_class;      /* N */
    const float *lat_unique_deg;    /* n_lat_unique */
    int32_t n_lat_unique;
    /* CELL mode: REAL4(sin), REAL4(cos), REAL4(tan) of LatRad, each evaluated in double (sphy_map_unary) */
    const float *sin_lat, *cos_lat, *tan_lat;
    float k0;                       /* REAL4(((24*60)/pi)*Gsc), the Python-double prefix of hargreaves.py:32-33 */
    float pcorr, tcorr;             /* Pcorr_fact (divides, sphy.py:757), Tcorr_fact */
    /* soil (sphy.py:173-247); e_root = REAL4(exp(-1/RootTT)), e_sub = REAL4(exp(-1/SubTT)) */
    sphy_param root_field, root_sat, root_dry, root_wilt, root_ksat, root_drainvel, e_root;
    sphy_param sub_field, sub_sat, sub_drainvel, e_sub;
    sphy_param caprise_max, kc, pmap /* PWS p-factor */, glac_frac /* NULL+0 when no glaciers */;
    sphy_param open_water_frac;
    float kc_open_water;            /* [OPENWATER] kcOpenWater (sphy.py:898) */
    /* infiltration excess (sphy.py:439-458) */
    float k_eff, alpha, alpha_sq /* REAL4(Alpha**2) */, labda_infil;
    sphy_param paved_frac;
    /* groundwater (modules/groundwater.py:51-57); e_delta = REAL4(exp(-1/deltaGw)), e_alpha = REAL4(exp(-alphaGw)),
     * h_den = REAL4(800*YieldGw*alphaGw) */
    sphy_param gw_sat, base_thresh, e_delta, e_alpha, h_den;
    sphy_param seepage;             /* GroundFLAG == 0 (sphy.py:1004) */
    /* snow (modules/snow.py:84-90): scalar-or-map like the reference reads them (readmap first, getfloat otherwise);
     * one_minus_snow_f = REAL4(1 - SnowF) for a cfg float (Python-double subtraction), formed per cell for a map */
    sphy_param tcrit, snow_sc, ddfs, snow_f;
    float one_minus_snow_f;
    float melt_cos[24];             /* REAL4(cos(3.1415*h/12)), h = 1..24 (snow.py:30) */
} sphy_wb_params;

typedef struct {                     /* in/out, updated in place (attribute names of the reference) */
    float *RootWater, *SubWater, *CapRise, *RootDrain;
    float *GwRecharge, *Gw, *BaseR, *H_gw;            /* GROUND */
    float *SubDrain;                                  /* !GROUND */
    float *SnowStore, *SnowWatStore, *TotalSnowStore; /* SNOW */
    float *waterbalanceTot;                           /* optional accumulator (sphy.py:1192) */
} sphy_wb_state;

typedef struct {                     /* forcing maps of the day (sphy.py:755-809) */
    const float *prec, *tavg, *tmin, *tmax, *etref;
    /* glacier aggregates of the day from sphy_glac_step (NULL when GlacFLAG = 0) */
    const float *TotalSnowStore_GLAC, *SnowR_GLAC, *GlacPerc, *GlacR;
    /* water balance diagnostics of a GlacRetreat = 1 run (sphy.py:1114-1146), all optional: the frozen oldGw map, the
     * glacier snow store that leaves TotalSnowStore on the update day, (iceDepth - oldIceDepth) in mm on that day, the new
     * GlacFrac map on the day after (the soil-storage term of dS uses it while the physics of that day used the old one) */
    const float *wbal_old_gw, *wbal_snow_adjust, *wbal_ice_delta, *wbal_glac_frac;
} sphy_wb_forcing;

typedef struct {                     /* outputs; every pointer except TotR may be NULL */
    float *TotR;                     /* sphy.py:1096 */
    float *RootRR, *RootDR, *RainR, *SnowR;           /* component runoff (sphy.py:1067-1075, snow.py:183) */
    float *ETref, *ETact, *ActETact, *Infil, *Rain, *RootPerc, *SubPerc, *wbal, *ETOpenWater, *Precip;
} sphy_wb_out;

/* init-time derived maps, evaluated like the oracle (double on the REAL4 operand, REAL4 result):
 *   EXP_NEG_INV: exp(-1/x) with -1/x a REAL4 division (rootzone.py:70, subzone.py:38, groundwater.py:27)
 *   SIN/COS/TAN_DEG2RAD: f(x*REAL4(pi/180)) for the latitude map (hargreaves.py:26-37); EXP_NEG: exp(-x) */
#define SPHY_UNARY_EXP_NEG_INV 0
#define SPHY_UNARY_SIN_DEG 1
#define SPHY_UNARY_COS_DEG 2
#define SPHY_UNARY_TAN_DEG 3
#define SPHY_UNARY_EXP_NEG 4
int sphy_map_unary(sphy_ctx *ctx, int op, const float *in, float *out, int64_t count, void *stream);

/* The reference's per-cell process functions one at a time (whole-map pass per call, like PCRaster), so that the
 * module call signatures stay usable: the same device functions the fused kernel composes (csrc/sphy_functions.cuh).
 * Inputs are scalar-or-map in the order of the reference's argument list (see csrc/wb_step.cu: k_module_op);
 * time-constant transcendentals are passed evaluated (e_root = REAL4(exp(-1/rootTT)) ...). */
enum {
    SPHY_OP_EXTRARAD = 0,        /* hargreaves.extrarad        hargreaves.py:24-39   in: Lat[deg] (+ k0, day_of_year) */
    SPHY_OP_HARGREAVES,          /* hargreaves.Hargreaves      hargreaves.py:43-47 */
    SPHY_OP_ETPOT,               /* ET.ETpot                   ET.py:24-26 */
    SPHY_OP_ETACT,               /* ET.ETact                   ET.py:30-35 */
    SPHY_OP_KS,                  /* ET.ks                      ET.py:39-47 */
    SPHY_OP_POTSNOWMELT,         /* snow.PotSnowMelt           modules/snow.py:27-33 */
    SPHY_OP_ACTSNOWMELT,         /* snow.ActSnowMelt           :37-39 */
    SPHY_OP_SNOWSTOREUPDATE,     /* snow.SnowStoreUpdate       :43-50 */
    SPHY_OP_MAXSNOWWATSTORAGE,   /* snow.MaxSnowWatStorage     :54-56 */
    SPHY_OP_SNOWWATSTORAGE,      /* snow.SnowWatStorage        :60-64 */
    SPHY_OP_TOTSNOWSTORAGE,      /* snow.TotSnowStorage        :68-70 */
    SPHY_OP_SNOWR,               /* snow.SnowR                 :74-80 */
    SPHY_OP_ROOTRUNOFF_SAT,      /* rootzone.RootRunoff, InfilFLAG = 0   rootzone.py:50-57  (2 outputs) */
    SPHY_OP_ROOTRUNOFF_INFIL,    /* rootzone.RootRunoff, InfilFLAG = 1   rootzone.py:26-47  (2 outputs) */
    SPHY_OP_ROOTDRAINAGE,        /* rootzone.RootDrainage      rootzone.py:63-74 */
    SPHY_OP_ROOTPERCOLATION,     /* rootzone.RootPercolation   :78-85 */
    SPHY_OP_CALCFRAC,            /* rootzone.CalcFrac          :89-94  (2 outputs) */
    SPHY_OP_CAPILRISE,           /* subzone.CapilRise          subzone.py:24-31 */
    SPHY_OP_SUBPERCOLATION,      /* subzone.SubPercolation     :35-41 */
    SPHY_OP_SUBDRAINAGE,         /* subzone.SubDrainage        :45-51 */
    SPHY_OP_GWRECHARGE,          /* groundwater.GroundWaterRecharge  modules/groundwater.py:26-29 */
    SPHY_OP_BASEFLOW,            /* groundwater.BaseFlow       :33-39 */
    SPHY_OP_HLEVEL,              /* groundwater.HLevel         :43-47 */
    SPHY_OP_SNOW_DYNAMIC,        /* snow.dynamic               modules/snow.py:110-187   (6 outputs) */
    SPHY_OP_GROUNDWATER_DYNAMIC, /* groundwater.dynamic        modules/groundwater.py:90-125 (4 outputs) */
    SPHY_OP_MUSLE,               /* musle.q_peak + musle.MUSLE modules/musle.py:29-46,143-150: in Q|TotR, Alpha_tc, Tc, K, C, P, LS,
                                    CFRG, cellarea, ha_area, q_is_discharge -> sediment yield, ton per cell */
    SPHY_OP_ADV_BLEND,           /* advanced_routing.ROUT, the blend   modules/advanced_routing.py:31-33: in accufractionflux, qold, qout,
                                    QFRAC, kx, REAL4(1-kx) or NaN -> Q [m3/day] */
    SPHY_OP_ADV_FINISH,          /* advanced_routing.ROUT, the rest    :34-38: in Q, upstream(Q), QFRAC, sres, qout -> sres, Q [m3/s], Qin */
    SPHY_OP_COUNT
};
int sphy_module_op(sphy_ctx *ctx, int op, const sphy_param *in, int n_in, float *const *out, int n_out, int64_t count,
                   float k0, int day_of_year, void *stream);

/* Dynamic vegetation (modules/dynamic_veg.py:27-128, DynVegFLAG = 1): NDVI -> Kc, LAI, canopy storage Smax; canopy
 * interception turns the day's precipitation into effective precipitation.  Runs BEFORE the glacier and water-balance
 * kernels (sphy.py:820-822) and produces the three maps they then consume: ETref (Hargreaves, or the input series),
 * the per-cell crop coefficient and the effective precipitation -- sphy_wb_step is called with SPHY_WB_ETREF_INPUT,
 * kc = Kc map, prec = PrecipEff and pcorr = 1.  The latitude / k0 / pcorr / tcorr fields are taken from `p`. */
typedef struct {
    const float *ndvi;               /* NDVI of the day (the caller keeps the previous map when a day is missing) */
    const float *prec, *tavg, *tmin, *tmax, *etref_in;   /* raw forcing; etref_in replaces Hargreaves when non-NULL */
    sphy_param lai_max, open_water_frac;
    /* the [DYNVEG] parameters folded on the host exactly like the reference's Python expressions (dynamic_veg.py:30-40):
     * cfg floats combine as Python doubles and are rounded to REAL4 where they meet a map; parameters given as maps
     * (dynamic_veg.py:57-63 tries readmap first) combine in REAL4 map arithmetic -- hence scalar-or-map, like every parameter */
    sphy_param sr_min;               /* (1+NDVImin)/(1-NDVImin) */
    sphy_param sr_range;             /* SR_max - SR_min */
    sphy_param fpar_range;           /* FPARmax - FPARmin */
    sphy_param log_one_m_fparmax;    /* log10(REAL4(1 - FPARmax)) */
    sphy_param kc_min, kc_range;     /* KCmin, KCmax - KCmin */
    sphy_param ndvi_min, ndvi_range; /* NDVImin, NDVImax - NDVImin */
    float *Scanopy;                  /* canopy storage, in/out */
    float *ETref, *Kc, *PrecipEff;   /* outputs consumed by sphy_glac_step / sphy_wb_step */
    float *LAI, *Int;                /* optional outputs (reporting: LAI, TotIntF) */
} sphy_veg;
int sphy_veg_step(sphy_ctx *ctx, const sphy_wb_params *p, const sphy_veg *v, int day_of_year, void *stream);

/* K6: Morgan-Morgan-Finney soil erosion + transport capacity (SedFLAG = 1, SedModel = 2; modules/mmf.py:193-309,
 * modules/sediment_transport.py:244-258).  Parameter maps are what mmf.init / sediment_transport.init derive once
 * (modules/mmf.py:107-190, modules/sediment_transport.py:113-214); scalar-or-map like every other parameter.  Order of
 * the texture classes in K / DR / v_s: clay, silt, sand. */
typedef struct {
    const float *prec;               /* precipitation reaching the ground (raw / pcorr, or the effective one of sphy_veg_step) */
    float pcorr;
    const float *q;                  /* routed discharge m3/s (q_is_discharge = 1) or TotR in mm (0), sphy.py:1217-1220 */
    int32_t q_is_discharge;
    const float *total_snow_store;   /* SnowFLAG = 1: snow cover protects the soil (mmf.py:220-224); else NULL */
    const float *lai;                /* cc_mode 1 */
    int32_t cc_mode;                 /* 0 table (+ harvest window), 1 min(1, LAI), 2 table only */
    int32_t harvest_on, excl_channels;
    sphy_param cos_slope, sin_slope_pow;             /* REAL4 cos(Slope), sin(Slope)**0.3 */
    sphy_param cc_table, gc_table, plant_height, no_erosion, rock_frac;
    sphy_param sowing, harvest, cc_harvest, gc_harvest, plant_height_harvest;
    sphy_param clay, silt, sand, hillslope;
    sphy_param v_field, v_field_harvest;             /* in-field flow velocity (mmf.py:175-189) */
    sphy_param roughness, roughness_harvest;         /* v_TC / v_b (sediment_transport.py:175-205) */
    sphy_param slope_streams_pow;                    /* SlopeStreams ** TC_gamma */
    float K[3], DR[3], v_s[3];                       /* detachability, REAL4 of the Stokes settling velocities (mmf.py:86) */
    float infil_alpha;               /* InfilFLAG = 1: PrecInt = DT * Alpha (mmf.py:236-237); 0 = fixed intensity */
    float ke_const;                  /* REAL4 0.29 * (1 - 0.72 * exp(-0.05 * PrecInt)) for the fixed intensity */
    float d_field, cell_length, tc_beta;
    float *DetRn, *DetRun, *SDepFld; /* optional reporting outputs, ton per cell */
    float *SedTrans;                 /* sediment taken into transport, ton per cell: the material of sphy_accucapacity */
    float *TC;                       /* transport capacity (NULL when SedTransFLAG = 0) */
} sphy_mmf;
int sphy_mmf_step(sphy_ctx *ctx, const sphy_mmf *p, int day_of_year, void *stream);

/* accucapacityflux / accucapacitystate (modules/sediment_transport.py:40-44): level sweep where every cell passes on at
 * most its capacity; flux and state are optional outputs (at least one). */
int sphy_accucapacity(sphy_ctx *ctx, const float *material, const float *capacity, float *flux, float *state, void *stream);

/* Sediment transport through reservoirs (modules/sediment_transport.py:46-107): sub-catchments are routed in the steps
 * of resorder.txt.  Reservoirs are listed in the reference's iteration order (step, then position in the file). */
typedef struct {
    int32_t n_res, n_steps;                 /* n_steps = max(step) + 1 */
    const int32_t *res_cell, *res_down;     /* [n_res] dam cell and the cell below it (-1 = none), device indices */
    const int32_t *res_step;                /* [n_res] */
    const int32_t *down_finish_seq;         /* [n_res] iteration at which the sub-catchment of res_down is finished (INT32_MAX = never) */
    const float *trap_eff, *outflow_eff;    /* [n_res] */
    const int32_t *finish_step;             /* [N] step at which the cell's sub-catchment is finished (INT32_MAX = never) */
    const int32_t *extent;                  /* [N] > 0 inside a reservoir: unlimited transport capacity */
    const int32_t *step_last_seq_host;      /* HOST array [n_steps]: last iteration index of steps <= k (-1 = none yet) */
    float *sed_trans, *tc_route, *cap_flux, *cap_state;   /* [N] work maps */
    float *outflow;                         /* [n_res] work */
} sphy_sedres;
int sphy_sed_reservoir_route(sphy_ctx *ctx, const sphy_sedres *r, const float *sed, const float *tc, float *yield, float *dep,
                             float *flux, void *stream);

/* K1 alone does not need the LDD: declare the number of active cells directly */
int sphy_set_ncells(sphy_ctx *ctx, int64_t n);

/* CUDA-graph replay of a step: the only per-day kernel arguments of sphy_wb_step are the day-of-year scalars of
 * hargreaves.extrarad (hargreaves.py:25-34).  sphy_day_scalars evaluates them on the host exactly as the reference does
 * (Python double -> REAL4 -> libm double -> REAL4) into out5 = {k0*dr, sin/cos/tan(delta), 0.0023*0.408}; with a device
 * block installed (5 floats, caller-owned; NULL uninstalls) the Ra table kernel reads them from there and ignores
 * `day_of_year`, so a captured graph stays valid for any day (SPHY_RA_TABLE or SPHY_WB_ETREF_INPUT only). */
int sphy_day_scalars(int day_of_year, float k0, float *out5);
int sphy_set_day_block(sphy_ctx *ctx, const float *dev5);
/* block_dev[0..5) = table_dev[*counter_dev][0..5); ++*counter_dev -- the head of a replayable step: the day advances on the
 * device, so the host never rewrites memory that a queued step still has to read */
int sphy_day_table_step(sphy_ctx *ctx, const float *table_dev, int n_rows, int *counter_dev, float *block_dev, void *stream);
/* day_of_year < 0: the context's per-day Ra table (SPHY_RA_TABLE) is already up to date -- ensemble members sharing a
 * context and a latitude map compute it once per day, not once per member. */
int sphy_wb_step(sphy_ctx *ctx, const sphy_wb_params *p, const sphy_wb_state *s, const sphy_wb_forcing *f,
                 const sphy_wb_out *o, int day_of_year, void *stream);

/* ---- K2: glacier fragment table ------------------------------------------------------------
 * replaces glacier.dynamic (modules/glacier.py:256-498): float64 per-fragment physics + groupby(MOD_ID).sum().
 * Table columns are device float64 arrays of n_rows (rows sorted by MOD_ID); `cell` = compact cell of each row;
 * `seg_ptr` (n_seg+1) delimits the rows of one model cell, `seg_cell` (n_seg) is that cell. */
typedef struct {
    int32_t n_rows, n_seg;
    const int32_t *cell, *seg_ptr, *seg_cell;
    const double *MOD_H, *GLAC_H, *FRAC_GLAC;
    const uint8_t *DEBRIS;
    double *Rain_GLAC, *Snow_GLAC;                     /* stale-value semantics (quirk Q7) */
    double *SnowStore_GLAC, *SnowWatStore_GLAC, *TotalSnowStore_GLAC, *AccuGlacMelt;
    double *GlacMelt, *GlacR, *GlacPerc, *SnowR_GLAC, *ActSnowMelt_GLAC, *GLAC_T;   /* per-row diagnostics */
    double tcrit, ddfs, snow_sc, ddfg, ddfdg, glac_f, lapse;
} sphy_glac_table;

typedef struct {                     /* per-cell aggregates (REAL4, written only at glacier cells) */
    float *Rain_GLAC, *Snow_GLAC, *ActSnowMelt_GLAC, *TotalSnowStore_GLAC, *SnowR_GLAC, *GlacMelt, *GlacR, *GlacPerc;
} sphy_glac_out;

int sphy_glac_step(sphy_ctx *ctx, const sphy_glac_table *t, const float *tavg, const float *prec, float tcorr,
                   float pcorr, const sphy_glac_out *o, void *stream);

/* Per-glacier daily tables (glacier.dynamic_reporting, modules/glacier.py:501-533): FRAC_GLAC-weighted average of up to 16
 * float64 table columns over the fragments of each glacier id (`row[row_ptr[g] .. row_ptr[g+1])`), written as REAL4
 * to out[v * n_glac + g]; columns flagged in `is_frac` report the summed fraction instead (the FRAC_GLAC column). */
int sphy_glac_report(sphy_ctx *ctx, int32_t n_glac, const int32_t *row_ptr, const int32_t *row, const double *frac,
                     int32_t n_vars, const double *const *cols, const uint8_t *is_frac, float *out, void *stream);

/* ---- K4: flow routing ------------------------------------------------------------------------
 * replaces routing.ROUT / routing.dynamic (modules/routing.py:25-29,70-116):
 *   Q = (1-kx) * accuflux(ldd, q*0.001*cellarea/86400) + kx*Qold   for ncomp runoff maps sharing one sweep.
 * runoff_mm[c]: N-cell map [mm/day]; q_state[c]: Qold in, Q out (N cells, compact order).
 * method: SPHY_ROUTE_LEVELS = level-by-level sweep, per-cell REAL8 sum stored REAL4 (bit-exact with the
 * oracle); SPHY_ROUTE_SCAN = subtree index ranges + float64 prefix sums (needs SPHY_ORDER_DFS). */
#define SPHY_ROUTE_LEVELS 0
#define SPHY_ROUTE_SCAN 1
int sphy_route_step(sphy_ctx *ctx, int ncomp, const float *const *runoff_mm, float *const *q_state, sphy_param kx,
                    float one_minus_kx /* REAL4(1-kx) when kx is a scalar: Python-double subtraction, routing.py:28 */,
                    int method, void *stream);
/* The same routing (SPHY_ROUTE_SCAN) for up to 8 runoff maps with a recession coefficient each -- the members of a
 * forcing/parameter ensemble on one basin (BASELINE config 5): one pass reads the index structures once for all of them. */
int sphy_route_step_members(sphy_ctx *ctx, int nmaps, const float *const *runoff_mm, float *const *q_state,
                            const sphy_param *kx, const float *one_minus_kx, void *stream);
/* The caller runs the routing of step t on a second stream while the water-balance kernel of step t + 1 occupies the GPU
 * (nothing in the vertical balance reads the routed discharge): every routing kernel behind the main pass then uses CTAs
 * small enough to be resident next to it (the tile totals are scanned by 128-thread blocks in two levels instead of by the
 * library), and `main_pass_event` (a cudaEvent_t of the caller, may be NULL) is recorded on the routing stream right
 * behind the bandwidth-bound main pass, so that the caller can start the next water-balance kernel exactly then. */
int sphy_set_overlap_tail(sphy_ctx *ctx, int on, void *main_pass_event);
/* Optional per-phase device timing of SPHY_ROUTE_SCAN (what bench.py's second roofline entry is measured with): while
 * enabled every routing step records four CUDA events on its stream; sphy_route_profile_read synchronises on the last
 * one and returns the mean milliseconds per step of {main pass, tile totals + crossing cells, trunk cells (incl. the
 * exchange on a partitioned context), whole routing} and the number of steps seen, then clears the record. */
int sphy_route_profile(sphy_ctx *ctx, int enable);
int sphy_route_profile_read(sphy_ctx *ctx, float *ms4, int *steps);
/* Cells per scan-routing tile (a compile-time constant of the library); partition ranges must start on a tile. */
int sphy_scan_tile_cells(void);

/* Trunk list of the scan path.  The reference rounds every cell's accuflux sum to REAL4 (PCRaster, routing.py:27); along
 * the ~10^4-cell main stem of a 10^8-cell basin those roundings add up to ~1e-5 of the flux, which exact float64 range
 * sums do not contain.  sphy_trunk_select lists the downstream-closed set {level >= min_level or sub_size >= min_cells}
 * (both thresholds must exceed the 1024-cell scan tile) -- plus, with world > 1, every cell whose upstream range crosses
 * one of the rank bounds (world + 1 host values, tile-aligned) -- by ascending device index (= DFS pre-order of the trunk
 * tree; global indices, identical on every rank).  Listed cells are finished from PUBLISHED prefixes: the context
 * publishes its range total and the range-local inclusive prefix at the positions in `pub_pos` (local indices, ascending);
 * entry s of the list reads the global prefix just before its cell at (before_rank[s], before_slot[s]) (rank -1 = none)
 * and at the end of its upstream range at (end_rank[s], end_slot[s]).  Trunk entries additionally carry the reference's
 * rounding drift as a second linear accumulation over the list (csrc/route_scan.cu: k_trunk_drift), which keeps
 * SPHY_ROUTE_SCAN within 1e-5 of the REAL4-per-cell arithmetic at any basin size.  The plan is host logic
 * (sphy_b200/multigpu.py: plan_trunk) over the small exported arrays "cell", "sub_size" (int32), "is_trunk" (uint8),
 * "last" (int32) and "donors" (8 x int32 per entry).  With a plan installed sphy_route_step(SPHY_ROUTE_SCAN) uses it. */
int sphy_trunk_select(sphy_ctx *ctx, int32_t min_level, int64_t min_cells, int world, const int64_t *bounds_host,
                      int64_t *count_out, void *stream);
int sphy_trunk_export(sphy_ctx *ctx, const char *name, void *dst_dev, size_t dst_bytes, void *stream);
int sphy_trunk_set_plan(sphy_ctx *ctx, int32_t n_pub, const int32_t *pub_pos_dev, const int32_t *before_rank_dev,
                        const int32_t *before_slot_dev, const int32_t *end_rank_dev, const int32_t *end_slot_dev,
                        int32_t stride /* >= 1 + max n_pub over the ranks */, int rank, int world, void *stream);

/* Multi-GPU (one context per GPU, each built on the whole LDD): restrict this context to the contiguous device-index
 * range [cell_begin, cell_end) of the DFS order (cell_begin on a 1024-cell tile).  After the call every compact map
 * and sphy_ncells() refer to the local cells only, and every whole-basin structure and workspace of the context is
 * released (memory per rank scales with N / world; sphy_ldd_export and the level sweep are no longer available).
 * Call order: sphy_trunk_select(world, bounds) -> sphy_set_partition -> sphy_trunk_set_plan.  The per-step confluence exchange is
 *   sphy_route_scan_begin -> all-gather of `send` (stride doubles per component and rank) -> sphy_route_scan_finish
 * and replaces the single-GPU sphy_route_step (modules/routing.py:25-29) on a partitioned context. */
int sphy_set_partition(sphy_ctx *ctx, int64_t cell_begin, int64_t cell_end, void *stream);
int sphy_route_scan_begin(sphy_ctx *ctx, int ncomp, const float *const *runoff_mm, float *const *q_state, sphy_param kx,
                          float one_minus_kx, double *send_dev, int32_t stride, void *stream);
int sphy_route_scan_finish(sphy_ctx *ctx, int ncomp, float *const *q_state, sphy_param kx, float one_minus_kx,
                           const double *gathered_dev, int32_t stride, int rank, int world, void *stream);
/* plain pcr.accuflux(ldd, m) / pcr.catchmenttotal(m, ldd) (sphy.py:838, routing.py:27) and pcr.upstream */
int sphy_accuflux(sphy_ctx *ctx, const float *material, const float *fraction /*nullable*/, float *flux, int method,
                  void *stream);
int sphy_upstream(sphy_ctx *ctx, const float *x, float *out, void *stream);

/* The same exchange over peer memory (NVLink P2P through CUDA IPC) instead of an NCCL all-gather: the publishing blocks of
 * sphy_route_scan_begin store this rank's block straight into the peers' receive buffers and release a per-peer flag; the
 * trunk kernel of sphy_route_scan_finish acquires the flags of all peers (bounded wait, SPHY_P2P_TIMEOUT_S seconds,
 * default 30: a dead peer yields NaN discharges on the trunk cells and sphy_p2p_timed_out() != 0, not a hang).  The step
 * counter lives in device memory.  sphy_p2p_setup_local allocates this rank's buffers (`pitch` >= ncomp * stride float64
 * per rank) and returns two 64-byte CUDA IPC handles; the host exchanges them (any transport) and calls sphy_p2p_connect
 * with all of them, ordered by rank.  Once connected, per step: sphy_route_scan_begin(send_dev = NULL) ->
 * sphy_route_scan_finish(gathered_dev = NULL); no host call in between. */
int sphy_p2p_setup_local(sphy_ctx *ctx, int world, int64_t pitch, unsigned char *handle_recv64, unsigned char *handle_flags64);
int sphy_p2p_connect(sphy_ctx *ctx, int rank, int world, const unsigned char *handles_recv, const unsigned char *handles_flags);
int sphy_p2p_timed_out(sphy_ctx *ctx, void *stream);

/* ---- K5 + routing with lakes/reservoirs ------------------------------------------------------
 * replaces advanced_routing.dynamic / ROUT (modules/advanced_routing.py:27-38,134-217), lakes.UpdateLakeHStore /
 * QLake (modules/lakes.py:28-158, incl. the assimilation of measured lake levels) and reservoirs.QRes/QSimple/QAdv (modules/reservoirs.py:26-86).
 * Structures are listed once (n_struct cells with QFRAC == 0); kind 1 = simple reservoir, 2 = advanced reservoir,
 * 3 = lake.  par[k][j]: parameter k of structure j (see reslake.cu for the column layout). */
#define SPHY_RL_NPAR 30
typedef struct {
    int32_t n_struct;
    const int32_t *cell;             /* compact cell of each structure */
    const int32_t *kind;
    const float *par;                /* [SPHY_RL_NPAR][n_struct] */
    const float *qfrac;              /* dense N-cell QFRAC map (0 at structures, 1 elsewhere; lakes.py:224) */
    float *StorRES;                  /* [n_struct] in/out, m3 */
    float *Qout, *Qin;               /* [n_struct] outputs of the step, m3/day */
    /* measured lake levels (lakes.py:28-97): NULL on days without a level map; else the level at every structure (NaN =
     * not measured) and which lakes are gauged (the `updatelakelevel` map) */
    const float *lake_level;
    const int32_t *update_level;
} sphy_reslake;

int sphy_route_reslake_step(sphy_ctx *ctx, const sphy_reslake *rl, const float *totr_mm, float *q_state,
                            sphy_param kx, float one_minus_kx, int day_of_year,
                            int method /* SPHY_ROUTE_LEVELS (bit-exact) or SPHY_ROUTE_SCAN; structures listed by ascending cell */,
                            void *stream);

/* Lakes / reservoirs on the trunk list: what a PARTITIONED basin runs instead of sphy_route_reslake_step (and what one GPU
 * may run as well).  A structure cell (QFRAC = 0, modules/lakes.py:224) passes nothing downstream and its outflow enters
 * the cell below it (modules/advanced_routing.py:29-38,159-161), so with E(c) the plain range sum of the per-cell material
 *     flux(c) = E(c) - sum over the top-level structures k strictly inside the upstream range of c of (E(k) - Qout(k)).
 * Only cells downstream of a structure see a correction: sphy_trunk_select_cuts puts them (and the structure cells) on the
 * trunk list next to the cells whose range leaves their rank, every other cell keeps the plain scan, and the listed cells
 * are finished from the published prefixes.  The structure table is replicated on every rank; what its update needs from
 * cells of other ranks -- TotR at the structure cell, and the day's discharge at the cells draining into it for
 * `upstream(ldd, Q)` (advanced_routing.py:35) -- travels in the same exchange block (SPHY_RL_EXTRA float64 per structure
 * behind the prefixes: stride >= 1 + max n_pub + SPHY_RL_EXTRA * n_struct).  The discharge of the donors only exists once
 * the listed cells are finished, so StorRES takes `- Qout + Qin` of day t at the head of day t + 1 (apply_pending != 0),
 * from the values in that day's block; the host flushes the last one itself (sphy_b200/model.py: flush_structures).
 * The tables are host logic over the exported trunk list (sphy_b200/multigpu.py: plan_reslake).  Call order:
 *   sphy_trunk_select_cuts -> sphy_trunk_export / sphy_cell_donors -> [sphy_set_partition] -> sphy_trunk_set_plan, then per
 *   step sphy_route_reslake_begin -> exchange (all-gather of `send`, or peer memory) -> sphy_route_reslake_finish. */
#define SPHY_RL_EXTRA 9
typedef struct {
    const int32_t *cut_slot;         /* [n_struct] trunk-list slot of every structure cell */
    const int32_t *entry_cut;        /* [m] structure index of a listed structure cell, -1 elsewhere */
    const int32_t *entry_ptr, *entry_idx;   /* CSR over the m list entries: top-level structures strictly inside the range */
    const int32_t *x_rank, *x_local; /* [n_struct * SPHY_RL_EXTRA] owner rank (-1: none) / local cell of donor 0..7 and of the structure cell */
    const float *qmask;              /* [n local] 2 at listed cells, QFRAC elsewhere (the main pass leaves listed cells alone) */
    float *material;                 /* reserved (may be NULL): the main pass forms the material 0.001 * cellarea * TotR itself */
    float *qday;                     /* [n local] in/out: the day's blended discharge, m3/day (zero before the first step) */
    double *extra;                   /* [n_struct * SPHY_RL_EXTRA] workspace */
    double *top;                     /* [m] workspace: global prefix at the end of each listed cell's range (its noise scale) */
} sphy_reslake_plan;
int sphy_trunk_select_cuts(sphy_ctx *ctx, int32_t n_cut, const int32_t *cut_cell_dev /* ascending device indices */, int32_t min_level,
                           int64_t min_cells, int world, const int64_t *bounds_host, int64_t *count_out, void *stream);
/* the cells draining directly into each of n given cells: 8 device indices per cell, -1 padded (before sphy_set_partition) */
int sphy_cell_donors(sphy_ctx *ctx, int32_t n_cells, const int32_t *cells_dev, int32_t *donors_out_dev, void *stream);
int sphy_route_reslake_begin(sphy_ctx *ctx, const sphy_reslake *rl, const sphy_reslake_plan *plan, const float *totr_mm, float *q_state,
                             sphy_param kx, float one_minus_kx, double *send_dev /* NULL: peer memory */, int32_t stride, void *stream);
int sphy_route_reslake_finish(sphy_ctx *ctx, const sphy_reslake *rl, const sphy_reslake_plan *plan, float *q_state, sphy_param kx,
                              float one_minus_kx, const double *gathered_dev /* NULL: peer memory */, int32_t stride, int rank, int world,
                              int day_of_year, int apply_pending, void *stream);

/* Open-water evaporation and precipitation on lakes/reservoirs (modules/advanced_routing.py:167-196, ETOpenWaterFLAG):
 * StorRES -= min(areatotal(ETOpenWater*cellarea*openWaterFrac, class)*0.001, StorRES); += areatotal(Precip*...)*0.001.
 * ow_ptr/ow_cell: for each structure, the cells of its open-water class (sphy.py:425-431: 8-connected clumps of
 * openWaterFrac > 0 labelled with the largest reservoir id inside) that have openWaterFrac > 0.  Call after
 * sphy_route_reslake_step. */
typedef struct {
    const int32_t *ow_ptr;           /* [n_struct + 1] */
    const int32_t *ow_cell;          /* [ow_ptr[n_struct]] */
    const float *open_water_frac;    /* dense N-cell map */
    const float *et_open_water;      /* dense N-cell map: sphy_wb_out.ETOpenWater of the step */
    const float *prec;               /* dense N-cell map: the day's raw precipitation forcing */
    float pcorr;
} sphy_openwater;
int sphy_reslake_openwater(sphy_ctx *ctx, const sphy_reslake *rl, const sphy_openwater *ow, void *stream);

/* The lake / reservoir functions one at a time over the structure table (n_struct values in, n_struct values out; NaN
 * where a function does not apply to a structure's kind) -- what the module-signature shims call:
 *   UPDATE_LAKE  lakes.UpdateLakeHStore  modules/lakes.py:28-132  -> lake level (rl->StorRES of gauged lakes is updated when
 *                                                                   rl->lake_level holds the day's measured levels)
 *   QLAKE        lakes.QLake             modules/lakes.py:136-158   in: lake level -> outflow [m3/day], 0 for non-lakes
 *   QSIMPLE / QADV / QRES  reservoirs.QSimple / QAdv / QRes  modules/reservoirs.py:26-86 -> outflow [m3/day] from rl->StorRES */
enum { SPHY_RL_OP_UPDATE_LAKE = 0, SPHY_RL_OP_QLAKE, SPHY_RL_OP_QSIMPLE, SPHY_RL_OP_QADV, SPHY_RL_OP_QRES, SPHY_RL_OP_COUNT };
int sphy_reslake_op(sphy_ctx *ctx, const sphy_reslake *rl, int op, int day_of_year, const float *in, float *out, void *stream);

#ifdef __cplusplus
}
#endif
#endif /* SPHY_B200_H */
