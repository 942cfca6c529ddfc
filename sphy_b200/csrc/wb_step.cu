// wb_step.cu -- K1 host side: day scalars, the per-day Ra table, init-time derived maps, launch.
// Device code lives in wb_kernel.cuh (instantiated in wb_inst_*.cu).
#include "wb_kernel.cuh"

using namespace wb;

namespace {

__global__ void k_ra_table(const float *__restrict__ lat_deg, int n, DayScalars d, float *__restrict__ ra)
{
    int u = blockIdx.x * blockDim.x + threadIdx.x;
    if (u >= n) return;
    float latrad = lat_deg[u] * 0.017453292519943295f;     // Lat * REAL4(pi/180)
    float sinl = (float)sin((double)latrad), cosl = (float)cos((double)latrad), tanl = (float)tan((double)latrad);
    ra[u] = extrarad(sinl, cosl, tanl, d);
}

// the same with the day scalars read from a device block (sphy_set_day_block): nothing in the launch depends on the day,
// so a CUDA graph of the step can be replayed for any day
__global__ void k_ra_table_dev(const float *__restrict__ lat_deg, int n, const DayScalars *__restrict__ dp, float *__restrict__ ra)
{
    int u = blockIdx.x * blockDim.x + threadIdx.x;
    if (u >= n) return;
    const DayScalars d = *dp;
    float latrad = lat_deg[u] * 0.017453292519943295f;
    float sinl = (float)sin((double)latrad), cosl = (float)cos((double)latrad), tanl = (float)tan((double)latrad);
    ra[u] = extrarad(sinl, cosl, tanl, d);
}

__global__ void k_day_select(const float *__restrict__ table, int n_rows, int *counter, float *__restrict__ block)
{
    if (blockIdx.x != 0 || threadIdx.x != 0) return;
    int t = *counter;
    if (t >= n_rows) t = n_rows - 1;
    for (int k = 0; k < 5; ++k) block[k] = table[(size_t)t * 5 + k];
    *counter = t + 1;
}

__global__ void k_map_unary(int op, const float *__restrict__ in, float *__restrict__ out, int64_t n)
{
    int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (i >= n) return;
    float x = in[i];
    float r;
    switch (op) {
        case SPHY_UNARY_EXP_NEG_INV: r = (float)exp((double)(-1.0f / x)); break;
        case SPHY_UNARY_SIN_DEG: r = (float)sin((double)(x * 0.017453292519943295f)); break;
        case SPHY_UNARY_COS_DEG: r = (float)cos((double)(x * 0.017453292519943295f)); break;
        case SPHY_UNARY_TAN_DEG: r = (float)tan((double)(x * 0.017453292519943295f)); break;
        default: r = (float)exp((double)(-x)); break;
    }
    out[i] = r;
}


inline bool aligned16(const void *p) { return (reinterpret_cast<uintptr_t>(p) & 15u) == 0; }

}  // namespace

// Day-of-year scalars, evaluated on the host exactly as the reference does (hargreaves.py:25-29):
// Python double angle -> REAL4 non-spatial -> libm double -> REAL4.
static DayScalars day_scalars(int doy, float k0)
{
    const double pi = 3.141592653589793;
    const double ang = (2.0 * pi * (double)doy) / 365.0;
    DayScalars d;
    const float dr = 1.0f + 0.033f * (float)cos((double)(float)ang);
    const float delta = 0.409f * (float)sin((double)(float)(ang - 1.39));
    d.k0dr = k0 * dr;
    d.sin_delta = (float)sin((double)delta);
    d.cos_delta = (float)cos((double)delta);
    d.tan_delta = (float)tan((double)delta);
    d.hg_c = (float)(0.0023 * 0.408);
    return d;
}

extern "C" int sphy_day_scalars(int day_of_year, float k0, float *out5)
{
    if (!out5) return SPHY_E_INVALID;
    const DayScalars d = day_scalars(day_of_year, k0);
    out5[0] = d.k0dr; out5[1] = d.sin_delta; out5[2] = d.cos_delta; out5[3] = d.tan_delta; out5[4] = d.hg_c;
    return SPHY_OK;
}

extern "C" int sphy_day_table_step(sphy_ctx *ctx, const float *table_dev, int n_rows, int *counter_dev, float *block_dev, void *stream)
{
    if (!ctx || !table_dev || n_rows < 1 || !counter_dev || !block_dev) return SPHY_E_INVALID;
    k_day_select<<<1, 32, 0, (cudaStream_t)stream>>>(table_dev, n_rows, counter_dev, block_dev);
    SPHY_LAUNCH_CHECK(ctx);
    return SPHY_OK;
}

extern "C" int sphy_set_day_block(sphy_ctx *ctx, const float *dev5)
{
    if (!ctx) return SPHY_E_INVALID;
    static_assert(sizeof(DayScalars) == 5 * sizeof(float), "day block layout");
    ctx->day_block = dev5;
    return SPHY_OK;
}

extern "C" int sphy_map_unary(sphy_ctx *ctx, int op, const float *in, float *out, int64_t count, void *stream)
{
    if (!ctx || !in || !out || count < 0 || op < 0 || op > 4) return SPHY_E_INVALID;
    if (count == 0) return SPHY_OK;
    k_map_unary<<<div_up(count, 256), 256, 0, (cudaStream_t)stream>>>(op, in, out, count);
    SPHY_LAUNCH_CHECK(ctx);
    return SPHY_OK;
}

extern "C" int sphy_wb_step(sphy_ctx *ctx, const sphy_wb_params *p, const sphy_wb_state *s, const sphy_wb_forcing *f,
                            const sphy_wb_out *o, int doy, void *stream_)
{
    if (!ctx || !p || !s || !f || !o) return SPHY_E_INVALID;
    if (!ctx->ldd_ready && ctx->n <= 0) SPHY_FAIL(ctx, SPHY_E_STATE, "sphy_wb_step: no active cells (call sphy_ldd_build or sphy_set_ncells)");
    cudaStream_t st = (cudaStream_t)stream_;
    const bool snow = p->flags & SPHY_WB_SNOW, ground = p->flags & SPHY_WB_GROUND, infil = p->flags & SPHY_WB_INFIL_EXCESS;
    const bool etin = p->flags & SPHY_WB_ETREF_INPUT;
    const int ra = etin ? 2 : p->ra_mode;
    // argument validation: the reference would raise a Python exception on a missing map
    if (!f->prec || !f->tavg || !o->TotR || !s->RootWater || !s->SubWater || !s->CapRise || !s->RootDrain)
        SPHY_FAIL(ctx, SPHY_E_INVALID, "sphy_wb_step: missing mandatory forcing/state/output map");
    if (etin ? !f->etref : (!f->tmin || !f->tmax)) SPHY_FAIL(ctx, SPHY_E_INVALID, "sphy_wb_step: missing ETref / Tmin / Tmax forcing");
    if (snow && (!f->tmax || !s->SnowStore || !s->SnowWatStore || !s->TotalSnowStore))
        SPHY_FAIL(ctx, SPHY_E_INVALID, "sphy_wb_step: SnowFLAG needs tmax and the three snow state maps (quirk Q3)");
    if (ground ? (!s->GwRecharge || !s->Gw || !s->BaseR || !s->H_gw) : !s->SubDrain)
        SPHY_FAIL(ctx, SPHY_E_INVALID, "sphy_wb_step: missing groundwater / SubDrain state map");
    if (!etin && ra == SPHY_RA_TABLE && (!p->lat_class || !p->lat_unique_deg || p->n_lat_unique <= 0 || p->n_lat_unique > 65536))
        SPHY_FAIL(ctx, SPHY_E_INVALID, "sphy_wb_step: RA_TABLE needs lat_class and 1..65536 unique latitudes");
    if (!etin && ra == SPHY_RA_CELL && (!p->sin_lat || !p->cos_lat || !p->tan_lat))
        SPHY_FAIL(ctx, SPHY_E_INVALID, "sphy_wb_step: RA_CELL needs sin/cos/tan latitude maps");
    const bool g_any = f->GlacR || f->GlacPerc || f->SnowR_GLAC || f->TotalSnowStore_GLAC;
    if (g_any && !(f->GlacR && f->GlacPerc && f->SnowR_GLAC && f->TotalSnowStore_GLAC))
        SPHY_FAIL(ctx, SPHY_E_INVALID, "sphy_wb_step: glacier aggregates must be given together");
    const void *ptrs[] = {f->prec, f->tavg, f->tmin, f->tmax, f->etref, f->TotalSnowStore_GLAC, f->SnowR_GLAC, f->GlacPerc,
                          f->GlacR, s->RootWater, s->SubWater, s->CapRise, s->RootDrain, s->GwRecharge, s->Gw, s->BaseR,
                          s->H_gw, s->SubDrain, s->SnowStore, s->SnowWatStore, s->TotalSnowStore, s->waterbalanceTot,
                          o->TotR, o->RootRR, o->RootDR, o->RainR, o->SnowR, o->ETref, o->ETact, o->ActETact, o->Infil,
                          o->Rain, o->RootPerc, o->SubPerc, o->wbal, o->ETOpenWater, o->Precip, p->sin_lat, p->cos_lat, p->tan_lat,
                          p->root_field.map, p->root_sat.map, p->root_dry.map, p->root_wilt.map, p->root_ksat.map,
                          p->root_drainvel.map, p->e_root.map, p->sub_field.map, p->sub_sat.map, p->sub_drainvel.map,
                          p->e_sub.map, p->caprise_max.map, p->kc.map, p->pmap.map, p->glac_frac.map,
                          p->open_water_frac.map, p->paved_frac.map, p->gw_sat.map, p->base_thresh.map, p->e_delta.map,
                          p->e_alpha.map, p->h_den.map, p->seepage.map};
    for (const void *q : ptrs)
        if (q && !aligned16(q)) SPHY_FAIL(ctx, SPHY_E_INVALID, "sphy_wb_step: every map must be 16-byte aligned");
    if (p->lat_class && (reinterpret_cast<uintptr_t>(p->lat_class) & 7u))
        SPHY_FAIL(ctx, SPHY_E_INVALID, "sphy_wb_step: lat_class must be 8-byte aligned");

    WbArgs a;
    a.p = *p;
    a.s = *s;
    a.f = *f;
    a.o = *o;
    a.n = ctx->n;
    a.d = day_scalars(doy, p->k0);
    a.ra_table = nullptr;
    if (ra == SPHY_RA_TABLE) {
        if (ctx->ra_table_cap < p->n_lat_unique) {
            // grows only when a larger latitude table is first seen (normally once, at the first step)
            int rc = dev_alloc(ctx, &ctx->ra_table, (size_t)p->n_lat_unique);
            if (rc != SPHY_OK) return rc;
            ctx->ra_table_cap = p->n_lat_unique;
        }
        if (doy < 0)
            ;                                                   // another member of the ensemble refreshed this context's table today
        else if (ctx->day_block)
            k_ra_table_dev<<<div_up(p->n_lat_unique, 128), 128, 0, st>>>(p->lat_unique_deg, p->n_lat_unique,
                                                                        reinterpret_cast<const DayScalars *>(ctx->day_block), ctx->ra_table);
        else
            k_ra_table<<<div_up(p->n_lat_unique, 128), 128, 0, st>>>(p->lat_unique_deg, p->n_lat_unique, a.d, ctx->ra_table);
        SPHY_LAUNCH_CHECK(ctx);
        a.ra_table = ctx->ra_table;
    } else if (ra == SPHY_RA_CELL && ctx->day_block) {
        SPHY_FAIL(ctx, SPHY_E_STATE, "sphy_wb_step: a device day block (sphy_set_day_block) needs SPHY_RA_TABLE or ETref input");
    }
    const bool diag = s->waterbalanceTot || o->ETref || o->ETact || o->ActETact || o->Infil || o->Rain || o->RootPerc ||
                      o->SubPerc || o->wbal || o->ETOpenWater || o->Precip;
    const bool usoil = !(p->root_field.map || p->root_sat.map || p->root_dry.map || p->root_wilt.map || p->root_ksat.map ||
                         p->e_root.map || p->sub_field.map || p->sub_sat.map || p->e_sub.map || p->caprise_max.map ||
                         p->kc.map || p->gw_sat.map || p->base_thresh.map || p->e_delta.map || p->e_alpha.map ||
                         p->h_den.map || p->seepage.map);
    const int cpt = 4;                                          // cells per thread (128-bit accesses; see wb_kernel.cuh)
    wb_kernel_t kern = snow ? (ground ? wb_pick_11(infil, ra, diag, usoil, cpt) : wb_pick_10(infil, ra, diag, usoil, cpt))
                            : (ground ? wb_pick_01(infil, ra, diag, usoil, cpt) : wb_pick_00(infil, ra, diag, usoil, cpt));
    if (!kern) SPHY_FAIL(ctx, SPHY_E_INVALID, "sphy_wb_step: no kernel for this switch combination");
    const int64_t n4 = (a.n + cpt - 1) / cpt;
    int64_t want = (n4 + WB_THREADS - 1) / WB_THREADS;
    // large grids without diagnostics: the variant that prefetches every thread's next tile with cp.async
    // worthwhile once every CTA walks several tiles
    const char *pf_env = getenv("SPHY_WB_PREFETCH");             // 0 = off, 2 = also on small grids (tests)
    const bool pf_enabled = !(pf_env && pf_env[0] == '0');
    const bool pf_force = pf_env && pf_env[0] == '2';
    size_t smem = 0;
    if (pf_enabled && !diag && (pf_force || want >= (int64_t)ctx->sm_count * 16)) {
        wb_kernel_t pf = snow ? (ground ? wb_pick_11(infil, ra, diag, usoil, -1) : wb_pick_10(infil, ra, diag, usoil, -1))
                              : (ground ? wb_pick_01(infil, ra, diag, usoil, -1) : wb_pick_00(infil, ra, diag, usoil, -1));
        if (pf) {
            SPHY_CUDA(ctx, cudaFuncSetAttribute(pf, cudaFuncAttributeMaxDynamicSharedMemorySize, WB_PF_SMEM));
            kern = pf;
            smem = WB_PF_SMEM;
        }
    }
    int occ = 0;
    SPHY_CUDA(ctx, cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kern, WB_THREADS, smem));
    if (occ < 1) occ = 1;
    int64_t cap = (int64_t)ctx->sm_count * occ;
    int grid = (int)(want < cap ? (want > 0 ? want : 1) : cap);
    kern<<<grid, WB_THREADS, smem, st>>>(a);
    SPHY_LAUNCH_CHECK(ctx);
    return SPHY_OK;
}

// K1 does not need the LDD: let callers that only run the water balance declare N directly
extern "C" int sphy_set_ncells(sphy_ctx *ctx, int64_t n)
{
    if (!ctx || n <= 0) return SPHY_E_INVALID;
    if (ctx->ldd_ready) SPHY_FAIL(ctx, SPHY_E_STATE, "sphy_set_ncells after sphy_ldd_build");
    ctx->n = n;
    return SPHY_OK;
}


// ---- dynamic vegetation (modules/dynamic_veg.py) --------------------------------------------------------------------
namespace {

struct VegArgs {
    sphy_veg v;
    DayScalars d;
    float pcorr, tcorr;
    int ra;                      // 0 table, 1 cell, 2 etref input
    const uint16_t *lat_class;
    const float *ra_table, *sin_lat, *cos_lat, *tan_lat;
    int64_t n;
};

__global__ void k_veg_step(const __grid_constant__ VegArgs a)
{
    const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (i >= a.n) return;
    const sphy_veg &v = a.v;
    const float Precip = __ldg(v.prec + i) / a.pcorr;                        // sphy.py:757
    float ETref;
    if (a.ra == 2) {
        ETref = __ldg(v.etref_in + i);
    } else {
        const float Temp = __ldg(v.tavg + i) + a.tcorr, TempMin = __ldg(v.tmin + i) + a.tcorr, TempMax = __ldg(v.tmax + i) + a.tcorr;
        const float Ra = a.ra == 0 ? __ldg(a.ra_table + __ldg(a.lat_class + i))
                                   : extrarad(__ldg(a.sin_lat + i), __ldg(a.cos_lat + i), __ldg(a.tan_lat + i), a.d);
        ETref = sf::Hargreaves(a.d.hg_c, Ra, Temp, TempMax, TempMin);
    }
    const float ndvi = fminf(__ldg(v.ndvi + i), 0.999f);                     // dynamic_veg.py:84
    const float lai_max = v.lai_max.map ? __ldg(v.lai_max.map + i) : v.lai_max.value;
    float Kc, Smax, LAI, Int, PreT, S;
    auto P = [&](const sphy_param &q) { return q.map ? __ldg(q.map + i) : q.value; };
    sf::Veg_function(ndvi, lai_max, P(v.sr_min), P(v.sr_range), P(v.fpar_range), P(v.log_one_m_fparmax), P(v.kc_min), P(v.kc_range),
                     P(v.ndvi_min), P(v.ndvi_range), Kc, Smax, LAI);
    sf::Inter_function(v.Scanopy[i] + Precip, Smax, ETref, Int, PreT, S);    // dynamic_veg.py:108-110
    v.Scanopy[i] = S;
    v.ETref[i] = ETref;
    v.Kc[i] = Kc;
    v.PrecipEff[i] = PreT;
    if (v.LAI) v.LAI[i] = LAI;
    if (v.Int) v.Int[i] = v.open_water_frac.map ? Int * (1.0f - __ldg(v.open_water_frac.map + i)) : Int;
}

}  // namespace

extern "C" int sphy_veg_step(sphy_ctx *ctx, const sphy_wb_params *p, const sphy_veg *v, int doy, void *stream_)
{
    if (!ctx || !p || !v) return SPHY_E_INVALID;
    if (ctx->n <= 0) SPHY_FAIL(ctx, SPHY_E_STATE, "sphy_veg_step: no active cells");
    if (!v->ndvi || !v->prec || !v->Scanopy || !v->ETref || !v->Kc || !v->PrecipEff)
        SPHY_FAIL(ctx, SPHY_E_INVALID, "sphy_veg_step: missing mandatory map");
    if (!v->etref_in && (!v->tavg || !v->tmin || !v->tmax)) SPHY_FAIL(ctx, SPHY_E_INVALID, "sphy_veg_step: missing temperature forcing");
    cudaStream_t st = (cudaStream_t)stream_;
    VegArgs a = {};
    a.v = *v;
    a.pcorr = p->pcorr;
    a.tcorr = p->tcorr;
    a.d = day_scalars(doy, p->k0);
    a.n = ctx->n;
    a.ra = v->etref_in ? 2 : p->ra_mode;
    if (a.ra == SPHY_RA_TABLE) {
        if (!p->lat_class || !p->lat_unique_deg || p->n_lat_unique <= 0) SPHY_FAIL(ctx, SPHY_E_INVALID, "sphy_veg_step: latitude classes missing");
        if (ctx->ra_table_cap < p->n_lat_unique) {
            int rc = dev_alloc(ctx, &ctx->ra_table, (size_t)p->n_lat_unique);
            if (rc != SPHY_OK) return rc;
            ctx->ra_table_cap = p->n_lat_unique;
        }
        k_ra_table<<<div_up(p->n_lat_unique, 128), 128, 0, st>>>(p->lat_unique_deg, p->n_lat_unique, a.d, ctx->ra_table);
        a.ra_table = ctx->ra_table;
        a.lat_class = p->lat_class;
    } else if (a.ra == SPHY_RA_CELL) {
        if (!p->sin_lat || !p->cos_lat || !p->tan_lat) SPHY_FAIL(ctx, SPHY_E_INVALID, "sphy_veg_step: latitude maps missing");
        a.sin_lat = p->sin_lat; a.cos_lat = p->cos_lat; a.tan_lat = p->tan_lat;
    }
    k_veg_step<<<div_up(ctx->n, 256), 256, 0, st>>>(a);
    SPHY_LAUNCH_CHECK(ctx);
    return SPHY_OK;
}

// ---- the reference's per-cell process functions, one at a time -----------------------------------------------------
// sphy_module_op keeps the module call signatures of the reference usable on device maps (sphy_b200/modules/*.py):
// the same __device__ functions the fused kernel composes, evaluated one whole-map pass per call like PCRaster does.
namespace {

struct OpArgs {
    int op;
    sphy_param in[16];
    float *out[6];
    DayScalars d;
    float melt_cos[24];
    int64_t n;
};

__global__ void k_module_op(const __grid_constant__ OpArgs a)
{
    const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (i >= a.n) return;
    auto I = [&](int k) { return a.in[k].map ? __ldg(a.in[k].map + i) : a.in[k].value; };
    float r0 = 0.0f, r1 = 0.0f, r2 = 0.0f, r3 = 0.0f, r4 = 0.0f, r5 = 0.0f;
    switch (a.op) {
        case SPHY_OP_SNOW_DYNAMIC: {                   // snow.dynamic, modules/snow.py:110-187
            // in: Temp TempMax Precip Tcrit DDFS SnowSC SnowStore SnowWatStore SnowFrac RainFrac TotalSnowStore_GLAC
            //     SnowR_GLAC SnowF (1-SnowF) (1-openWaterFrac)     out: Rain SnowR SnowSoil SnowStore SnowWatStore TotalSnowStore
            const float Temp = I(0), TempMax = I(1), Precip = I(2);
            const float Snow = Temp >= I(3) ? 0.0f : Precip;
            const float Rain = Temp < I(3) ? 0.0f : Precip;
            const float Pot = TempMax < 0.0f ? 0.0f : sf::PotSnowMelt(Temp, TempMax, I(4), a.melt_cos);
            const float Act = sf::ActSnowMelt(I(6), Pot);
            const float ss = sf::SnowStoreUpdate(I(6), Snow, Act, Temp, I(7));
            const float MaxSWS = sf::MaxSnowWatStorage(I(5), ss);
            const float sws = sf::SnowWatStorage(TempMax, MaxSWS, I(7), Act, Rain);
            const float tss = sf::TotSnowStorage(ss, sws, I(8), I(9)) + I(10);
            float SnowR = sf::SnowR(sws, MaxSWS, Act, Rain, I(7), I(8)) + I(11);
            const float SnowSoil = SnowR * I(12);
            SnowR = (SnowR * I(13)) * I(14);
            r0 = Rain; r1 = SnowR; r2 = SnowSoil; r3 = ss; r4 = sws; r5 = tss;
        } break;
        case SPHY_OP_MUSLE: {                          // musle.dynamic, modules/musle.py:29-46,143-150
            // in: Q (m3/s) or TotR (mm), Alpha_tc, Tc, K_USLE, C_USLE, P_USLE, LS_USLE, CFRG, cellarea, ha_area, q_is_discharge
            const float area = I(8);
            const float Runoff = I(10) != 0.0f ? (((I(0) * 3600.0f) * 24.0f) / area) * 1000.0f : I(0);       // sphy.py:1217-1220
            const float q_peak = ((I(1) * Runoff) * area) / (3600000.0f * I(2));
            const float sed = ((((((11.8f * sf::d_pow((Runoff * q_peak) * I(9), 0.56f)) * I(3)) * I(4)) * I(5)) * I(6)) * I(7));
            r0 = (sed / 10000.0f) * area;
        } break;
        case SPHY_OP_GROUNDWATER_DYNAMIC: {            // groundwater.dynamic, modules/groundwater.py:90-125
            // in: e_delta GwRecharge ActSubPerc GlacPerc Gw BaseR BaseThresh e_alpha H_gw h_den (1-openWaterFrac)
            // out: GwRecharge Gw BaseR H_gw
            const float gwr = sf::GroundWaterRecharge(I(0), I(1), I(2) + I(3));
            float gw = I(4) + gwr;
            float baser = sf::BaseFlow(gw, I(5), gwr, I(6), I(7));
            gw = gw - baser;
            baser = baser * I(10);
            r0 = gwr; r1 = gw; r2 = baser; r3 = sf::HLevel(I(8), I(7), gwr, I(9));
        } break;
        case SPHY_OP_EXTRARAD: {                       // in: Lat [deg]
            const float latrad = I(0) * 0.017453292519943295f;
            r0 = extrarad(sf::d_sin(latrad), sf::d_cos(latrad), sf::d_tan(latrad), a.d);
        } break;
        case SPHY_OP_HARGREAVES: r0 = sf::Hargreaves(a.d.hg_c, I(0), I(1), I(2), I(3)); break;        // ra temp tempmax tempmin
        case SPHY_OP_ETPOT: r0 = sf::ETpot(I(0), I(1)); break;                                         // etr kc
        case SPHY_OP_ETACT: r0 = sf::ETact(I(0), I(1), I(2), I(3), I(4)); break;                       // etpot rootwater rootsat etreddry rainfrac
        case SPHY_OP_KS: r0 = sf::ks(I(0), I(1), I(2), I(3), I(4)); break;                             // rootfield rootdry pmap rootwater etpot
        case SPHY_OP_POTSNOWMELT: r0 = sf::PotSnowMelt(I(0), I(1), I(2), a.melt_cos); break;           // temp tempmax ddfs
        case SPHY_OP_ACTSNOWMELT: r0 = sf::ActSnowMelt(I(0), I(1)); break;
        case SPHY_OP_SNOWSTOREUPDATE: r0 = sf::SnowStoreUpdate(I(0), I(1), I(2), I(3), I(4)); break;
        case SPHY_OP_MAXSNOWWATSTORAGE: r0 = sf::MaxSnowWatStorage(I(0), I(1)); break;
        case SPHY_OP_SNOWWATSTORAGE: r0 = sf::SnowWatStorage(I(0), I(1), I(2), I(3), I(4)); break;
        case SPHY_OP_TOTSNOWSTORAGE: r0 = sf::TotSnowStorage(I(0), I(1), I(2), I(3)); break;
        case SPHY_OP_SNOWR: r0 = sf::SnowR(I(0), I(1), I(2), I(3), I(4), I(5)); break;
        case SPHY_OP_ROOTRUNOFF_SAT: sf::RootRunoff_sat(I(0), I(1), I(2), I(3), I(4), r0, r1); break;  // rainfrac rain rootwater rootsat rootksat
        case SPHY_OP_ROOTRUNOFF_INFIL:                 // k_eff rootksat rootsat rootwater labda alpha alpha_sq rain paved
            sf::RootRunoff_infil(I(0), I(1), I(2), I(3), I(4), I(5), I(6), I(7), I(8), r0, r1);
            break;
        case SPHY_OP_ROOTDRAINAGE: r0 = sf::RootDrainage(I(0), I(1), I(2), I(3), I(4), I(5)); break;   // rootwater rootdrain rootfield rootsat drainvel e_root
        case SPHY_OP_ROOTPERCOLATION: r0 = sf::RootPercolation(I(0), I(1), I(2), I(3), I(4)); break;   // rootwater subwater rootfield e_root subsat
        case SPHY_OP_CALCFRAC: sf::CalcFrac(I(0), I(1), I(2), I(3), r0, r1); break;                    // rootwater rootfield rootdrain rootperc
        case SPHY_OP_CAPILRISE: r0 = sf::CapilRise(I(0), I(1), I(2), I(3), I(4), I(5)); break;         // subfield subwater capmax rootwater rootsat rootfield
        case SPHY_OP_SUBPERCOLATION: r0 = sf::SubPercolation(I(0), I(1), I(2), I(3), I(4)); break;     // subwater subfield e_sub gw gwsat
        case SPHY_OP_SUBDRAINAGE: r0 = sf::SubDrainage(I(0), I(1), I(2), I(3), I(4), I(5)); break;     // subwater subfield subsat drainvel subdrainage e_sub
        case SPHY_OP_GWRECHARGE: r0 = sf::GroundWaterRecharge(I(0), I(1), I(2) + I(3)); break;         // e_delta gwrecharge subperc glacperc
        case SPHY_OP_ADV_BLEND: {                      // advanced_routing.ROUT, advanced_routing.py:31-33: the recession blend [m3/day]
            // in: accufractionflux  qold[m3/s]  qout[m3/day]  QFRAC  kx  (1-kx) as REAL4 (NaN: kx is a map, form it per cell)
            const float omk = I(5) == I(5) ? I(5) : 1.0f - I(4);
            r0 = I(3) == 0.0f ? I(2) : omk * I(0) + ((I(1) * 24.0f) * 3600.0f) * I(4);
        } break;
        case SPHY_OP_ADV_FINISH: {                     // advanced_routing.py:34-38
            // in: Q[m3/day]  upstream(ldd, Q)  QFRAC  sres  qout        out: sres  Q[m3/s]  Qin
            const float qin = I(2) == 0.0f ? I(1) : 0.0f;
            r0 = (I(3) - I(4)) + qin;
            r1 = I(0) / 86400.0f;
            r2 = qin;
        } break;
        case SPHY_OP_BASEFLOW: r0 = sf::BaseFlow(I(0), I(1), I(2), I(3), I(4)); break;                 // gw baser gwrecharge basethresh e_alpha
        case SPHY_OP_HLEVEL: r0 = sf::HLevel(I(0), I(1), I(2), I(3)); break;                           // Hgw e_alpha gwrecharge h_den
        default: break;
    }
    const float r[6] = {r0, r1, r2, r3, r4, r5};
#pragma unroll
    for (int k = 0; k < 6; ++k)
        if (a.out[k]) a.out[k][i] = r[k];
}

}  // namespace

extern "C" int sphy_module_op(sphy_ctx *ctx, int op, const sphy_param *in, int n_in, float *const *out, int n_out, int64_t count,
                              float k0, int day_of_year, void *stream)
{
    if (!ctx || !in || !out || n_in < 1 || n_in > 16 || n_out < 1 || n_out > 6 || count < 0 || op < 0 || op >= SPHY_OP_COUNT)
        return SPHY_E_INVALID;
    if (count == 0) return SPHY_OK;
    OpArgs a = {};
    a.op = op;
    for (int k = 0; k < n_in; ++k) a.in[k] = in[k];
    for (int k = 0; k < n_out; ++k) a.out[k] = out[k];
    if (!a.out[0]) SPHY_FAIL(ctx, SPHY_E_INVALID, "sphy_module_op: first output is NULL");
    a.d = day_scalars(day_of_year, k0);
    for (int h = 1; h <= 24; ++h) a.melt_cos[h - 1] = (float)cos((double)(float)(3.1415 * h / 12));     // snow.py:30
    a.n = count;
    k_module_op<<<div_up(count, 256), 256, 0, (cudaStream_t)stream>>>(a);
    SPHY_LAUNCH_CHECK(ctx);
    return SPHY_OK;
}
