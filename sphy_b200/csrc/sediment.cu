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
