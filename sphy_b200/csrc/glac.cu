// glac.cu -- K2: glacier fragment table physics + per-model-cell aggregation.
//
// Replaces glacier.dynamic (/root/reference/modules/glacier.py:256-498): the reference gathers T and P at the
// glacier cells into a pandas float64 table, updates ~20 columns per fragment, multiplies by FRAC_GLAC and
// groupby(MOD_ID).sum()s eight of them back onto rasters (cast to REAL4 by numpy2pcr).  Here one thread owns
// one model cell (segment): it walks that cell's fragments (rows are sorted by MOD_ID, so they are contiguous)
// in float64, accumulates the eight aggregates sequentially in row order and writes them as REAL4.
// Quirk Q7 is kept: Rain_GLAC / Snow_GLAC are only assigned under their mask, so stale values persist.
#include "common.cuh"

namespace {

__global__ void k_glac_step(sphy_glac_table t, const float *__restrict__ tavg, const float *__restrict__ prec, float tcorr,
                            float pcorr, sphy_glac_out o)
{
    int sgm = blockIdx.x * blockDim.x + threadIdx.x;
    if (sgm >= t.n_seg) return;
    double a_rain = 0, a_snow = 0, a_asm = 0, a_tss = 0, a_snowr = 0, a_melt = 0, a_r = 0, a_perc = 0;
    for (int32_t r = t.seg_ptr[sgm]; r < t.seg_ptr[sgm + 1]; ++r) {
        const int32_t c = t.cell[r];
        // Temp / Precip maps as the reference forms them (sphy.py:755-786) before pcr2numpy -> float64
        const double MOD_T = (double)(tavg[c] + tcorr);
        const double Prec = (double)(prec[c] / pcorr);
        const double GLAC_T = MOD_T - (t.MOD_H[r] - t.GLAC_H[r]) * t.lapse;
        double Rain = t.Rain_GLAC[r], Snow = t.Snow_GLAC[r];
        if (GLAC_T >= t.tcrit) Rain = Prec; else Snow = Prec;                    // glacier.py:284-289
        const double Tmelt = fmax(GLAC_T, 0.0);
        const double Pot = Tmelt * t.ddfs;
        double ss = t.SnowStore_GLAC[r], sws = t.SnowWatStore_GLAC[r];
        const double Act = fmin(ss, Pot);
        const double OldSS = ss;
        ss = ss + Snow - Act;
        const bool cold = GLAC_T < 0.0;
        if (cold) ss = ss + sws;
        const double MaxSWS = t.snow_sc * ss;
        const double OldSWS = sws;
        sws = fmin(MaxSWS, sws + Act + Rain);
        if (cold) sws = 0.0;
        const double tss = ss + sws;
        const bool mask = (sws == MaxSWS) && (OldSS > 0.0);
        const double SnowR = mask ? Act + Rain - (sws - OldSWS) : 0.0;
        const bool partial = (OldSS > 0.0) && (ss == 0.0), full = (OldSS == 0.0) && (ss == 0.0);
        const double ddf = t.DEBRIS[r] ? t.ddfdg : t.ddfg;
        double melt = 0.0;
        if (partial) melt = fmax(GLAC_T - (OldSS / t.ddfs), 0.0) * ddf;
        if (full) melt = Tmelt * ddf;
        const bool nosnow = OldSS == 0.0;
        const double GlacR = nosnow ? t.glac_f * (melt + Rain) : t.glac_f * melt;
        const double GlacPerc = nosnow ? (1.0 - t.glac_f) * (melt + Rain) : (1.0 - t.glac_f) * melt;
        t.Rain_GLAC[r] = Rain;
        t.Snow_GLAC[r] = Snow;
        t.SnowStore_GLAC[r] = ss;
        t.SnowWatStore_GLAC[r] = sws;
        t.TotalSnowStore_GLAC[r] = tss;
        t.AccuGlacMelt[r] = t.AccuGlacMelt[r] + melt;
        if (t.GlacMelt) t.GlacMelt[r] = melt;
        if (t.GlacR) t.GlacR[r] = GlacR;
        if (t.GlacPerc) t.GlacPerc[r] = GlacPerc;
        if (t.SnowR_GLAC) t.SnowR_GLAC[r] = SnowR;
        if (t.ActSnowMelt_GLAC) t.ActSnowMelt_GLAC[r] = Act;
        if (t.GLAC_T) t.GLAC_T[r] = GLAC_T;
        const double fr = t.FRAC_GLAC[r];                                        // glacier.py:421-427
        a_rain += Rain * fr;
        a_snow += Snow * fr;
        a_asm += Act * fr;
        a_tss += tss * fr;
        a_snowr += SnowR * fr;
        a_melt += melt * fr;
        a_r += GlacR * fr;
        a_perc += GlacPerc * fr;
    }
    const int32_t cell = t.seg_cell[sgm];
    if (o.Rain_GLAC) o.Rain_GLAC[cell] = (float)a_rain;
    if (o.Snow_GLAC) o.Snow_GLAC[cell] = (float)a_snow;
    if (o.ActSnowMelt_GLAC) o.ActSnowMelt_GLAC[cell] = (float)a_asm;
    o.TotalSnowStore_GLAC[cell] = (float)a_tss;
    o.SnowR_GLAC[cell] = (float)a_snowr;
    if (o.GlacMelt) o.GlacMelt[cell] = (float)a_melt;
    o.GlacR[cell] = (float)a_r;
    o.GlacPerc[cell] = (float)a_perc;
}

}  // namespace

extern "C" int sphy_glac_step(sphy_ctx *ctx, const sphy_glac_table *t, const float *tavg, const float *prec, float tcorr,
                              float pcorr, const sphy_glac_out *o, void *stream)
{
    if (!ctx || !t || !tavg || !prec || !o) return SPHY_E_INVALID;
    if (t->n_seg < 0 || t->n_rows < 0) return SPHY_E_INVALID;
    if (t->n_seg == 0) return SPHY_OK;
    if (!t->cell || !t->seg_ptr || !t->seg_cell || !t->MOD_H || !t->GLAC_H || !t->FRAC_GLAC || !t->DEBRIS || !t->Rain_GLAC ||
        !t->Snow_GLAC || !t->SnowStore_GLAC || !t->SnowWatStore_GLAC || !t->TotalSnowStore_GLAC || !t->AccuGlacMelt)
        SPHY_FAIL(ctx, SPHY_E_INVALID, "sphy_glac_step: incomplete glacier table");
    if (!o->TotalSnowStore_GLAC || !o->SnowR_GLAC || !o->GlacR || !o->GlacPerc)
        SPHY_FAIL(ctx, SPHY_E_INVALID, "sphy_glac_step: the four aggregates consumed by sphy_wb_step are mandatory");
    k_glac_step<<<div_up(t->n_seg, 128), 128, 0, (cudaStream_t)stream>>>(*t, tavg, prec, tcorr, pcorr, *o);
    SPHY_LAUNCH_CHECK(ctx);
    return SPHY_OK;
}

// ---- per-glacier daily tables (glacier.dynamic_reporting, modules/glacier.py:501-533) ------------------------------
// FRAC_GLAC-weighted average of table columns over the fragments of every glacier id: one warp per glacier, float64
// sums, REAL4 results (the reference casts the day's row with astype(float32)); 0 where the summed fraction is 0.
namespace {

constexpr int GLAC_REPORT_MAX_VARS = 16;
struct GlacReportArgs {
    const double *col[GLAC_REPORT_MAX_VARS];
    uint8_t is_frac[GLAC_REPORT_MAX_VARS];   // the FRAC_GLAC column reports the summed fraction itself
    int n_vars, n_glac;
    const int32_t *row_ptr, *row;            // fragments of glacier g: row[row_ptr[g] .. row_ptr[g+1])
    const double *frac;
    float *out;                              // [n_vars][n_glac]
};

__global__ void k_glac_report(const __grid_constant__ GlacReportArgs a)
{
    const int g = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    if (g >= a.n_glac) return;
    const int b = a.row_ptr[g], e = a.row_ptr[g + 1];
    double fs = 0.0;
    for (int k = b + lane; k < e; k += 32) fs += a.frac[a.row[k]];
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) fs += __shfl_xor_sync(0xffffffffu, fs, d);
    for (int v = 0; v < a.n_vars; ++v) {
        double s = 0.0;
        for (int k = b + lane; k < e; k += 32) {
            const int r = a.row[k];
            s += a.col[v][r] * a.frac[r];
        }
#pragma unroll
        for (int d = 16; d > 0; d >>= 1) s += __shfl_xor_sync(0xffffffffu, s, d);
        if (lane == 0) {
            const double q = a.is_frac[v] ? fs : s / fs;
            a.out[(size_t)v * a.n_glac + g] = (q != q) ? 0.0f : (float)q;          // fillna(0.0), glacier.py:524
        }
    }
}

}  // namespace

extern "C" int sphy_glac_report(sphy_ctx *ctx, int32_t n_glac, const int32_t *row_ptr, const int32_t *row, const double *frac,
                                int32_t n_vars, const double *const *cols, const uint8_t *is_frac, float *out, void *stream)
{
    if (!ctx || n_glac < 0 || n_vars < 0 || n_vars > GLAC_REPORT_MAX_VARS) return SPHY_E_INVALID;
    if (n_glac == 0 || n_vars == 0) return SPHY_OK;
    if (!row_ptr || !row || !frac || !cols || !is_frac || !out) SPHY_FAIL(ctx, SPHY_E_INVALID, "sphy_glac_report: null argument");
    GlacReportArgs a = {};
    for (int v = 0; v < n_vars; ++v) {
        if (!cols[v]) SPHY_FAIL(ctx, SPHY_E_INVALID, "sphy_glac_report: null column %d", v);
        a.col[v] = cols[v];
        a.is_frac[v] = is_frac[v];
    }
    a.n_vars = n_vars; a.n_glac = n_glac; a.row_ptr = row_ptr; a.row = row; a.frac = frac; a.out = out;
    k_glac_report<<<div_up((int64_t)n_glac * 32, 128), 128, 0, (cudaStream_t)stream>>>(a);
    SPHY_LAUNCH_CHECK(ctx);
    return SPHY_OK;
}
