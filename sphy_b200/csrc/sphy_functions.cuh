// sphy_functions.cuh -- the reference's per-cell process functions as __device__ functions, one per reference
// function, same argument meaning, REAL4 arithmetic in the reference's operation order.
// The fused kernel K1 (wb_kernel.cuh) composes them; module_ops.cu exposes each one on its own through
// sphy_module_op so that sphy_b200/modules/*.py keeps the reference's module call signatures.
// Compiled with -fmad=false: every operation rounds like a PCRaster REAL4 operation.
#pragma once
#include <math.h>

namespace sf {

// x / d for x >= 0 that is very often exactly 0 (no excess water, a dry day): the IEEE division takes its slow path for a
// zero numerator, and 0 / d is +0 for every d > 0 -- same bits, no division
__device__ __forceinline__ float div_often_zero(float x, float d) { return (x == 0.0f && d > 0.0f) ? 0.0f : x / d; }

// double-evaluated transcendentals stored as REAL4 (assumption A2, see oracle/pcr_shim.py)
__device__ __forceinline__ float d_acos(float x) { return (float)acos((double)x); }
__device__ __forceinline__ float d_sin(float x) { return (float)sin((double)x); }
__device__ __forceinline__ float d_cos(float x) { return (float)cos((double)x); }
__device__ __forceinline__ float d_tan(float x) { return (float)tan((double)x); }
__device__ __forceinline__ float d_exp(float x) { return (float)exp((double)x); }
__device__ __forceinline__ float d_pow(float x, float y) { return (float)pow((double)x, (double)y); }

// ---- hargreaves.py ------------------------------------------------------------------------------------
// extrarad (hargreaves.py:24-39) for one latitude, given REAL4 sin/cos/tan of LatRad and the day's solar terms
__device__ __forceinline__ float extrarad(float sinl, float cosl, float tanl, float k0dr, float sin_delta, float cos_delta,
                                          float tan_delta)
{
    const float x = (-1.0f * tanl) * tan_delta;
    const float omegas = d_acos(x);
    return k0dr * ((omegas * sinl) * sin_delta + (cosl * cos_delta) * d_sin(omegas));
}
// Hargreaves (hargreaves.py:43-47); hg_c = REAL4(0.0023*0.408); x**0.5 == sqrtf(x) (double-rounding safe)
__device__ __forceinline__ float Hargreaves(float hg_c, float ra, float temp, float tempmax, float tempmin)
{
    return fmaxf(((hg_c * ra) * (temp + 17.8f)) * sqrtf(fmaxf(tempmax - tempmin, 0.0f)), 0.0f);
}

// ---- ET.py ----------------------------------------------------------------------------------------------
__device__ __forceinline__ float ETpot(float etr, float kc) { return etr * kc; }                       // ET.py:24-26
__device__ __forceinline__ float ETact(float etpot, float rootwater, float rootsat, float etreddry, float rainfrac)
{                                                                                                      // ET.py:30-35
    const float etredwet = rootwater >= rootsat ? 0.0f : 1.0f;
    return rainfrac > 0.0f ? fminf((etpot * etreddry) * etredwet, rootwater) : 0.0f;
}
__device__ __forceinline__ float ks(float rootfield, float rootdry, float pmap, float rootwater, float etpot)
{                                                                                                      // ET.py:39-47
    const float TAW = rootfield - rootdry;
    const float p = fmaxf(fminf(pmap + 0.04f * (5.0f - etpot), 0.8f), 0.1f);
    const float RootPWS = rootfield - TAW * p;
    return fmaxf(fminf((rootwater - rootdry) / (RootPWS - rootdry), 1.0f), 0.0f);
}
__device__ __forceinline__ float etreddry(float rootwater, float rootdry, float rootwilt)               // sphy.py:922-927
{
    return fmaxf(fminf((rootwater - rootdry) / (rootwilt - rootdry), 1.0f), 0.0f);
}

// ---- modules/dynamic_veg.py ---------------------------------------------------------------------------------
// Veg_function (dynamic_veg.py:27-40) with the cfg floats folded on the host; ndvi already limited to 0.999
__device__ __forceinline__ void Veg_function(float ndvi, float lai_max, float sr_min, float sr_range, float fpar_range,
                                             float log_one_m_fparmax, float kc_min, float kc_range, float ndvi_min, float ndvi_range,
                                             float &Kc, float &Smax, float &LAI)
{
    const float SR = (1.0f + ndvi) / (1.0f - ndvi);
    const float FPAR = fminf(((SR - sr_min) / sr_range) * fpar_range, 0.95f);
    LAI = (lai_max * (float)log10((double)(1.0f - FPAR))) / log_one_m_fparmax;
    Smax = (0.935f + 0.498f * LAI) - 0.00575f * (LAI * LAI);
    Kc = kc_min + kc_range * fmaxf(fminf((ndvi - ndvi_min) / ndvi_range, 1.0f), 0.0f);
}
// Inter_function (dynamic_veg.py:44-49)
__device__ __forceinline__ void Inter_function(float S, float Smax, float Etr, float &Int, float &PreT, float &Sout)
{
    PreT = fmaxf(0.0f, S - Smax);
    S = S - PreT;
    Int = fminf(1.5f * Etr, S);
    Sout = S - Int;
}

// ---- modules/snow.py -------------------------------------------------------------------------------------
// PotSnowMelt (snow.py:27-33); melt_cos[h-1] = REAL4(cos(3.1415*h/12))
__device__ __forceinline__ float PotSnowMelt(float temp, float tempmax, float ddfs, const float *melt_cos)
{
    float thour = 0.0f;
    const float dT = tempmax - temp;
#pragma unroll
    for (int h = 0; h < 24; ++h) thour = thour + fmaxf(0.0f, temp + dT * melt_cos[h]);
    return div_often_zero(thour * ddfs, 24.0f);
}
__device__ __forceinline__ float ActSnowMelt(float snowstore, float potmelt) { return fminf(snowstore, potmelt); }   // :37-39
__device__ __forceinline__ float SnowStoreUpdate(float snowstore, float snow, float actmelt, float temp, float snowwatstore)
{                                                                                                      // snow.py:43-50
    return ((snowstore + snow) - actmelt) + (temp < 0.0f ? snowwatstore : 0.0f);
}
__device__ __forceinline__ float MaxSnowWatStorage(float snowsc, float snowstore) { return snowsc * snowstore; }     // :54-56
__device__ __forceinline__ float SnowWatStorage(float temp, float maxsnowwatstore, float snowwatstore, float actmelt, float rain)
{                                                                                                      // snow.py:60-64
    return temp < 0.0f ? 0.0f : fminf(maxsnowwatstore, (snowwatstore + actmelt) + rain);
}
__device__ __forceinline__ float TotSnowStorage(float snowstore, float snowwatstore, float snowfrac, float rainfrac)
{                                                                                                      // snow.py:68-70
    return (snowstore + snowwatstore) * (snowfrac + rainfrac);
}
__device__ __forceinline__ float SnowR(float snowwatstore, float maxsnowwatstore, float actmelt, float rain, float oldsnowwatstore,
                                       float snowfrac)
{                                                                                                      // snow.py:74-80
    return snowwatstore == maxsnowwatstore ? ((actmelt + rain) - (snowwatstore - oldsnowwatstore)) * snowfrac : 0.0f;
}

// ---- rootzone.py --------------------------------------------------------------------------------------------
// RootRunoff, saturation-excess branch (rootzone.py:50-57)
__device__ __forceinline__ void RootRunoff_sat(float rainfrac, float rain, float rootwater, float rootsat, float rootksat,
                                               float &rootrunoff, float &infil)
{
    infil = fmaxf(0.0f, fminf(fminf(rain, rootksat), rootsat - rootwater));
    rootrunoff = rainfrac > 0.0f ? rain - infil : 0.0f;
}
// RootRunoff, infiltration-excess branch (rootzone.py:26-47); alpha_sq = REAL4(Alpha**2)
__device__ __forceinline__ void RootRunoff_infil(float k_eff, float rootksat, float rootsat, float rootwater, float labda,
                                                 float alpha, float alpha_sq, float rain, float paved, float &rootrunoff, float &infil)
{
    const float cap = ((k_eff * rootksat) / 24.0f) * d_pow(1.0f + ((rootsat - rootwater) / rootsat), labda);
    const float ar = alpha * rain;
    const float dd = ar - cap;
    const float iex = ar > cap ? rain - (dd * dd) / (alpha_sq * rain) : rain;
    infil = fmaxf(0.0f, fminf(iex, rootsat - rootwater)) * (1.0f - paved);
    rootrunoff = rain - infil;
}
// RootDrainage (rootzone.py:63-74); e_root = REAL4(exp(-1/rootTT))
__device__ __forceinline__ float RootDrainage(float rootwater, float rootdrain, float rootfield, float rootsat, float drainvel,
                                              float e_root)
{
    const float rootexcess = fmaxf(rootwater - rootfield, 0.0f);
    const float rootlat = div_often_zero(rootexcess, rootsat - rootfield) * drainvel;
    return fmaxf(fminf(rootexcess, rootlat * (1.0f - e_root) + rootdrain * e_root), 0.0f);
}
// RootPercolation (rootzone.py:78-85)
__device__ __forceinline__ float RootPercolation(float rootwater, float subwater, float rootfield, float e_root, float subsat)
{
    const float rootexcess = fmaxf(rootwater - rootfield, 0.0f);
    float rootperc = rootexcess * (1.0f - e_root);
    rootperc = subwater >= subsat ? 0.0f : fminf(subsat - subwater, rootperc);
    return fmaxf(fminf(rootperc, rootexcess), 0.0f);
}
// CalcFrac (rootzone.py:89-94)
__device__ __forceinline__ void CalcFrac(float rootwater, float rootfield, float rootdrain, float rootperc, float &newdrain,
                                         float &newperc)
{
    const float rootexcess = fmaxf(rootwater - rootfield, 0.0f);
    const float frac = ((rootdrain + rootperc) - rootexcess) / (rootdrain + rootperc);
    newdrain = rootdrain - rootdrain * frac;
    newperc = rootperc - rootperc * frac;
}

// ---- subzone.py -----------------------------------------------------------------------------------------------
__device__ __forceinline__ float CapilRise(float subfield, float subwater, float capmax, float rootwater, float rootsat,
                                           float rootfield)
{                                                                                                      // subzone.py:24-31
    const float subrelwat = fmaxf(fminf(subwater / subfield, 1.0f), 0.0f);
    const float rootrelwat = fmaxf(fminf(rootwater / rootfield, 1.0f), 0.0f);
    const float caprise = fminf(subwater, (capmax * (1.0f - rootrelwat)) * subrelwat);
    return fminf(caprise, rootsat - rootwater);
}
__device__ __forceinline__ float SubPercolation(float subwater, float subfield, float e_sub, float gw, float gwsat)
{                                                                                                      // subzone.py:35-41
    const float d = subwater - subfield;
    return (gw < gwsat && d > 0.0f) ? d * (1.0f - e_sub) : 0.0f;
}
__device__ __forceinline__ float SubDrainage(float subwater, float subfield, float subsat, float drainvel, float subdrainage,
                                             float e_sub)
{                                                                                                      // subzone.py:45-51
    const float subexcess = fmaxf(subwater - subfield, 0.0f);
    const float sublateral = div_often_zero(subexcess, subsat - subfield) * drainvel;
    const float s = (sublateral + subdrainage) * (1.0f - e_sub);
    return fmaxf(fminf(s, subwater), 0.0f);
}

// ---- modules/groundwater.py --------------------------------------------------------------------------------------
// e_delta = REAL4(exp(-1/deltaGw)), e_alpha = REAL4(exp(-alphaGw)), h_den = REAL4(800*YieldGw*alphaGw)
__device__ __forceinline__ float GroundWaterRecharge(float e_delta, float gwrecharge, float subperc_plus_glacperc)
{                                                                                                      // groundwater.py:26-29
    const float gwseep = (1.0f - e_delta) * subperc_plus_glacperc;
    return (e_delta * gwrecharge) + gwseep;
}
__device__ __forceinline__ float BaseFlow(float gw, float baser, float gwrecharge, float basethresh, float e_alpha)
{                                                                                                      // groundwater.py:33-39
    return gw <= basethresh ? 0.0f : (baser * e_alpha + gwrecharge * (1.0f - e_alpha));
}
__device__ __forceinline__ float HLevel(float Hgw, float e_alpha, float gwrecharge, float h_den)
{                                                                                                      // groundwater.py:43-47
    return (Hgw * e_alpha) + div_often_zero(gwrecharge * (1.0f - e_alpha), h_den);
}

// ---- modules/routing.py -----------------------------------------------------------------------------------------
__device__ __forceinline__ float runoff_to_m3s(float q_mm, float cellarea) { return ((q_mm * 0.001f) * cellarea) / 86400.0f; }  // :26

}  // namespace sf
