// sediment.cu -- K6: Morgan-Morgan-Finney soil erosion per cell (modules/mmf.py:193-309) fused with the transport
// capacity of the runoff (modules/mmf.py:99-103, modules/sediment_transport.py:244-258).  One thread per cell, REAL4
// arithmetic in the reference's operation order (-fmad=false), transcendentals evaluated in double on the REAL4
// operand (assumption A2).  Everything that does not change from day to day (Manning velocities, roughness factor,
// cos/sin of the slope, SlopeStreams**gamma) is folded into parameter maps at init time by
// sphy_b200/modules/mmf.py / sediment_transport.py, exactly as the reference's init() does.
// The capacity-limited routing that follows (accucapacityflux/state) is sphy_accucapacity in route.cu.
#include "common.cuh"
#include "sphy_functions.cuh"

namespace {

__device__ __forceinline__ float par(const sphy_param &p, int64_t i) { return p.map ? __ldg(p.map + i) : p.value; }

__global__ void __launch_bounds__(256) k_mmf_step(const __grid_constant__ sphy_mmf a, int64_t n, float cellarea, float yday)
{
    const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (i >= n) return;
    const float Precip = __ldg(a.prec + i) / a.pcorr;
    const float qv = __ldg(a.q + i);
    const float Runoff = a.q_is_discharge ? (((qv * 3600.0f) * 24.0f) / cellarea) * 1000.0f : qv;     // sphy.py:1217-1220
    // harvested areas, mmf.py:201-206
    bool hv = false;
    if (a.harvest_on) {
        const float H = par(a.harvest, i), S = par(a.sowing, i);
        float h = 0.0f;
        if (H < S) h = (H < yday && S > yday) ? 1.0f : h;
        if (H > S) h = (yday > H || yday < S) ? 1.0f : h;
        if (H == 0.0f) h = 0.0f;
        hv = h == 1.0f;
    }
    float CC = par(a.cc_table, i);
    if (a.cc_mode == 1) CC = fminf(1.0f, __ldg(a.lai + i));                                           // mmf.py:195-196
    else if (a.cc_mode == 0 && a.harvest_on) CC = hv ? par(a.cc_harvest, i) : CC;                     // mmf.py:209-210
    const float GC = (a.harvest_on && hv) ? par(a.gc_harvest, i) : par(a.gc_table, i);
    const float rock = par(a.rock_frac, i);
    float Cover;
    if (a.total_snow_store) Cover = fminf(((__ldg(a.total_snow_store + i) > 0.0f ? 1.0f : 0.0f) + GC) + rock, 1.0f);
    else Cover = fminf(GC + rock, 1.0f);
    const float cosS = par(a.cos_slope, i);
    const float Rf = Precip * cosS;                                                                    // mmf.py:21-33
    const float LD = Rf * CC;
    const float DT = Rf - LD;
    float KE_DT;
    if (a.infil_alpha > 0.0f) KE_DT = (DT * (0.29f * (1.0f - 0.72f * sf::d_exp(-0.05f * (DT * a.infil_alpha))))) * 100.0f;
    else KE_DT = (DT * a.ke_const) * 100.0f;                                                           // mmf.py:36-42
    if (DT == 0.0f) KE_DT = 0.0f;
    const float PH = (a.harvest_on && hv) ? par(a.plant_height_harvest, i) : par(a.plant_height, i);
    const float KE_LD = PH < 0.15f ? 0.0f : LD * (15.8f * sf::d_pow(PH, 0.5f) - 5.87f);               // mmf.py:45-47
    const float KE = KE_DT + KE_LD;
    const float bare = fmaxf(0.0f, 1.0f - (par(a.no_erosion, i) + Cover));
    const float q15 = sf::d_pow(Runoff, 1.5f);
    const float sin03 = par(a.sin_slope_pow, i);
    const float hill = par(a.hillslope, i);
    const float v = (a.harvest_on && hv) ? par(a.v_field_harvest, i) : par(a.v_field, i);
    const float tex[3] = {par(a.clay, i), par(a.silt, i), par(a.sand, i)};
    float F[3], Hh[3], G[3], D[3];
#pragma unroll
    for (int k = 0; k < 3; ++k) {
        F[k] = (((a.K[k] * tex[k]) * bare) * KE) * 1e-3f;                                              // mmf.py:50-52
        float h = ((((a.DR[k] * tex[k]) * q15) * bare) * sin03) * 1e-3f;                               // mmf.py:55-62
        if (a.excl_channels) h = h * hill;
        Hh[k] = h;
        const float den = v * a.d_field;
        const float Nf = den == 0.0f ? nanf("") : ((a.cell_length / cosS) * a.v_s[k]) / den;           // mmf.py:85-88
        const float DEP = fminf(100.0f, 44.1f * sf::d_pow(Nf, 0.29f));                                 // mmf.py:91-93
        G[k] = (F[k] + h) * (1.0f - DEP / 100.0f);                                                     // mmf.py:96-99
        D[k] = ((F[k] + h) * DEP) / 100.0f;
    }
    const float Gt = (G[0] + G[1]) + G[2];
    const float sed = (Gt * cellarea) / 1000.0f;                                                       // ton per cell
    if (a.DetRn) a.DetRn[i] = (((F[0] + F[1]) + F[2]) * cellarea) / 1000.0f;
    if (a.DetRun) a.DetRun[i] = (((Hh[0] + Hh[1]) + Hh[2]) * cellarea) / 1000.0f;
    if (a.SDepFld) a.SDepFld[i] = (((D[0] + D[1]) + D[2]) * cellarea) / 1000.0f;
    a.SedTrans[i] = sed;
    if (a.TC) {                                                                                        // mmf.py:99-103
        const float rf = (a.harvest_on && hv) ? par(a.roughness_harvest, i) : par(a.roughness, i);
        const float q = (Runoff / 1000.0f) * a.cell_length;
        a.TC[i] = (rf * sf::d_pow(q, a.tc_beta)) * par(a.slope_streams_pow, i);
    }
}

}  // namespace

extern "C" int sphy_mmf_step(sphy_ctx *ctx, const sphy_mmf *p, int day_of_year, void *stream)
{
    if (!ctx || !p) return SPHY_E_INVALID;
    if (ctx->n <= 0) SPHY_FAIL(ctx, SPHY_E_STATE, "sphy_mmf_step: no active cells");
    if (!p->prec || !p->q || !p->SedTrans) SPHY_FAIL(ctx, SPHY_E_INVALID, "sphy_mmf_step: prec, q and SedTrans are mandatory");
    if (p->cc_mode == 1 && !p->lai) SPHY_FAIL(ctx, SPHY_E_INVALID, "sphy_mmf_step: canopy cover from LAI needs the LAI map");
    k_mmf_step<<<div_up(ctx->n, 256), 256, 0, (cudaStream_t)stream>>>(*p, ctx->n, ctx->cellarea, (float)day_of_year);
    SPHY_LAUNCH_CHECK(ctx);
    return SPHY_OK;
}

// ---- sediment transport through reservoirs (modules/sediment_transport.py:46-107) -----------------------------------
// The reference routes the sub-catchments of the reservoirs step by step (resorder.txt): capacity sweep, then per
// reservoir of the step: trapped yield, the sub-catchment is marked finished and emptied, the untrapped part is
// handed to the cell below the dam.  Here every step is: sphy_accucapacity sweep -> k_sedres_collect (one thread per
// reservoir) -> k_sedres_update (per cell: accumulate flux / deposition, rebuild the material of the next sweep) ->
// k_sedres_outflows (the few dam outflows, added in the reference's order by ONE thread so that sums are identical).
// An outflow added at iteration g survives in the material as long as the sub-catchment of its target cell has not been
// finished by a LATER iteration (the reference zeroes every finished sub-catchment again at each iteration).
namespace {

__global__ void k_sedres_prepare(const sphy_sedres r, const float *sed, const float *tc, int64_t n, float *yield, float *dep, float *flux)
{
    const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (i >= n) return;
    r.sed_trans[i] = sed[i];
    r.tc_route[i] = r.extent[i] > 0 ? 1e10f : tc[i];            // sediment_transport.py:57
    yield[i] = 0.0f; dep[i] = 0.0f; flux[i] = 0.0f;
}

__global__ void k_sedres_collect(const sphy_sedres r, int step, float *yield)
{
    const int s = blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= r.n_res || r.res_step[s] != step) return;
    const float f = r.cap_flux[r.res_cell[s]];
    yield[r.res_cell[s]] = f * r.trap_eff[s];                   // sediment_transport.py:79
    r.outflow[s] = f * r.outflow_eff[s];                        // sediment_transport.py:87
}

__global__ void k_sedres_update(const sphy_sedres r, const float *sed, int step, int64_t n, float *dep, float *flux)
{
    const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (i >= n) return;
    const int fs = r.finish_step[i];
    const float finished = fs <= step ? 1.0f : 0.0f, now = fs == step ? 1.0f : 0.0f;
    flux[i] = r.cap_flux[i] * finished + flux[i];               // sediment_transport.py:93
    dep[i] = dep[i] + r.cap_state[i] * now;                     // sediment_transport.py:96
    r.sed_trans[i] = fs <= step ? 0.0f : sed[i];                // sediment_transport.py:85 (outflows are added next)
}

__global__ void k_sedres_outflows(const sphy_sedres r, int last_seq)
{
    if (blockIdx.x != 0 || threadIdx.x != 0) return;
    for (int s = 0; s <= last_seq; ++s) {                       // reservoirs are stored in iteration order
        const int x = r.res_down[s];
        if (x < 0) continue;
        if (s == last_seq || r.down_finish_seq[s] > last_seq) r.sed_trans[x] = r.sed_trans[x] + r.outflow[s];
    }
}

__global__ void k_sedres_final(const sphy_sedres r, int last_step, int64_t n, float *flux)
{
    const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (i >= n) return;
    const float finished = r.finish_step[i] <= last_step ? 1.0f : 0.0f;
    flux[i] = r.cap_flux[i] * (1.0f - finished) + flux[i];      // sediment_transport.py:103
}

}  // namespace

extern "C" int sphy_sed_reservoir_route(sphy_ctx *ctx, const sphy_sedres *r, const float *sed, const float *tc, float *yield,
                                        float *dep, float *flux, void *stream)
{
    if (!ctx || !r || !sed || !tc || !yield || !dep || !flux) return SPHY_E_INVALID;
    if (!ctx->ldd_ready) SPHY_FAIL(ctx, SPHY_E_STATE, "sphy_sed_reservoir_route before sphy_ldd_build");
    if (r->n_res < 0 || r->n_steps < 1 || !r->finish_step || !r->extent || !r->sed_trans || !r->tc_route || !r->cap_flux ||
        !r->cap_state || !r->step_last_seq_host)
        SPHY_FAIL(ctx, SPHY_E_INVALID, "sphy_sed_reservoir_route: incomplete sphy_sedres");
    cudaStream_t st = (cudaStream_t)stream;
    const int64_t n = ctx->n;
    const int grid = (int)div_up(n, 256);
    k_sedres_prepare<<<grid, 256, 0, st>>>(*r, sed, tc, n, yield, dep, flux);
    for (int step = 0; step < r->n_steps; ++step) {
        int rc = sphy_accucapacity(ctx, r->sed_trans, r->tc_route, r->cap_flux, r->cap_state, stream);
        if (rc != SPHY_OK) return rc;
        if (r->n_res > 0) k_sedres_collect<<<(int)div_up(r->n_res, 128), 128, 0, st>>>(*r, step, yield);
        k_sedres_update<<<grid, 256, 0, st>>>(*r, sed, step, n, dep, flux);
        if (r->step_last_seq_host[step] >= 0) k_sedres_outflows<<<1, 32, 0, st>>>(*r, r->step_last_seq_host[step]);
    }
    int rc = sphy_accucapacity(ctx, r->sed_trans, r->tc_route, r->cap_flux, nullptr, stream);
    if (rc != SPHY_OK) return rc;
    k_sedres_final<<<grid, 256, 0, st>>>(*r, r->n_steps - 1, n, flux);
    SPHY_LAUNCH_CHECK(ctx);
    return SPHY_OK;
}
