// wb_kernel.cuh -- K1 device code: the fused vertical water balance, templated on the module switches.
// Included by the wb_inst_*.cu translation units (one per SnowFLAG x GroundFLAG combination, so that the
// 48 instantiations compile in parallel) and by wb_step.cu (host launcher, Ra table, init-time maps).
#pragma once
#include <math.h>

#include "common.cuh"
#include "sphy_functions.cuh"

namespace wb {

struct DayScalars {
    float k0dr;        // REAL4(k0) * dr                                   hargreaves.py:31-34
    float sin_delta, cos_delta, tan_delta;
    float hg_c;        // REAL4(0.0023 * 0.408)                            hargreaves.py:45
};

struct WbArgs {
    sphy_wb_params p;
    sphy_wb_state s;
    sphy_wb_forcing f;
    sphy_wb_out o;
    DayScalars d;
    const float *ra_table;
    int64_t n;
};

// ---- 128-bit streaming accessors ------------------------------------------------------------------
struct V4 { float v[4]; };

__device__ __forceinline__ V4 ld4(const float *p, int64_t i)
{
    float4 t = __ldg(reinterpret_cast<const float4 *>(p + i));
    return V4{{t.x, t.y, t.z, t.w}};
}
__device__ __forceinline__ V4 ld4_rw(const float *p, int64_t i)
{
    float4 t = *reinterpret_cast<const float4 *>(p + i);
    return V4{{t.x, t.y, t.z, t.w}};
}
__device__ __forceinline__ void st4(float *p, int64_t i, const V4 &a)
{
    *reinterpret_cast<float4 *>(p + i) = make_float4(a.v[0], a.v[1], a.v[2], a.v[3]);
}
__device__ __forceinline__ V4 splat(float x) { return V4{{x, x, x, x}}; }

template <int W>
struct Lanes;   // W cells per thread: 4 (vector path) or 1 (tail)

template <>
struct Lanes<4> {
    static __device__ __forceinline__ V4 in(const float *p, int64_t i) { return ld4(p, i); }
    static __device__ __forceinline__ V4 inrw(const float *p, int64_t i) { return ld4_rw(p, i); }
    static __device__ __forceinline__ void out(float *p, int64_t i, const V4 &a) { st4(p, i, a); }
    static __device__ __forceinline__ V4 par(const sphy_param &q, int64_t i) { return q.map ? ld4(q.map, i) : splat(q.value); }
};
template <>
struct Lanes<2> {
    static __device__ __forceinline__ V4 in(const float *p, int64_t i)
    {
        float2 t = __ldg(reinterpret_cast<const float2 *>(p + i));
        return V4{{t.x, t.y, 0.f, 0.f}};
    }
    static __device__ __forceinline__ V4 inrw(const float *p, int64_t i)
    {
        float2 t = *reinterpret_cast<const float2 *>(p + i);
        return V4{{t.x, t.y, 0.f, 0.f}};
    }
    static __device__ __forceinline__ void out(float *p, int64_t i, const V4 &a)
    {
        *reinterpret_cast<float2 *>(p + i) = make_float2(a.v[0], a.v[1]);
    }
    static __device__ __forceinline__ V4 par(const sphy_param &q, int64_t i) { return q.map ? in(q.map, i) : splat(q.value); }
};
template <>
struct Lanes<1> {
    static __device__ __forceinline__ V4 in(const float *p, int64_t i) { return splat(__ldg(p + i)); }
    static __device__ __forceinline__ V4 inrw(const float *p, int64_t i) { return splat(p[i]); }
    static __device__ __forceinline__ void out(float *p, int64_t i, const V4 &a) { p[i] = a.v[0]; }
    static __device__ __forceinline__ V4 par(const sphy_param &q, int64_t i) { return splat(q.map ? __ldg(q.map + i) : q.value); }
};

// Ra for one latitude (hargreaves.py:24-39)
__device__ __forceinline__ float extrarad(float sinl, float cosl, float tanl, const DayScalars &d)
{
    return sf::extrarad(sinl, cosl, tanl, d.k0dr, d.sin_delta, d.cos_delta, d.tan_delta);
}

constexpr int WB_THREADS = 128;

// ---- asynchronous prefetch of the next tile (cp.async, one 16-byte slot per stream and thread) ------------------
// The kernel is a grid-stride loop; without prefetch a thread's ~14 loads are only in flight while it waits for them,
// not while it computes and stores.  With PF every thread copies the inputs of its NEXT tile into its own shared-
// memory slots right after it has moved the current ones into registers, so HBM requests are outstanding during the
// whole compute/store phase.  A thread only ever reads its own slots: no barrier, just cp.async.wait_group.
enum { S_PREC, S_TAVG, S_T1 /* tmin | etref */, S_T2 /* tmax */, S_LAT, S_RW, S_SW, S_CAPR, S_RDRAIN,
       S_G0 /* GwRecharge | SubDrain */, S_G1, S_G2, S_G3, S_SN0, S_SN1, S_SN2, S_DVEL, S_COUNT };

__device__ __forceinline__ void cp_async16(float4 *dst, const void *src)
{
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"((unsigned)__cvta_generic_to_shared(dst)), "l"(src) : "memory");
}
__device__ __forceinline__ void cp_async8(float4 *dst, const void *src)
{
    asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"((unsigned)__cvta_generic_to_shared(dst)), "l"(src) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_group 0;" ::: "memory"); }
__device__ __forceinline__ V4 pf_get(const float4 *pf, int slot)
{
    const float4 t = pf[slot * WB_THREADS];
    return V4{{t.x, t.y, t.z, t.w}};
}

// issue the copies of the 4 cells starting at cell i (same stream set as the load phase of wb_cells below)
template <bool SNOW, bool GROUND, int RA>
__device__ __forceinline__ void pf_issue(const WbArgs &a, int64_t i, float4 *pf)
{
#define PF_CP(slot, ptr) cp_async16(pf + (slot) * WB_THREADS, (ptr) + i)
    PF_CP(S_PREC, a.f.prec);
    PF_CP(S_TAVG, a.f.tavg);
    if (RA == 2) {
        PF_CP(S_T1, a.f.etref);
        if (SNOW) PF_CP(S_T2, a.f.tmax);
    } else {
        PF_CP(S_T1, a.f.tmin);
        PF_CP(S_T2, a.f.tmax);
    }
    if (RA == 0) cp_async8(pf + S_LAT * WB_THREADS, a.p.lat_class + i);
    PF_CP(S_RW, a.s.RootWater);
    PF_CP(S_SW, a.s.SubWater);
    PF_CP(S_CAPR, a.s.CapRise);
    PF_CP(S_RDRAIN, a.s.RootDrain);
    if (GROUND) {
        PF_CP(S_G0, a.s.GwRecharge);
        PF_CP(S_G1, a.s.Gw);
        PF_CP(S_G2, a.s.BaseR);
        PF_CP(S_G3, a.s.H_gw);
    } else {
        PF_CP(S_G0, a.s.SubDrain);
    }
    if (SNOW) {
        PF_CP(S_SN0, a.s.SnowStore);
        PF_CP(S_SN1, a.s.SnowWatStore);
        PF_CP(S_SN2, a.s.TotalSnowStore);
    }
    if (a.p.root_drainvel.map) PF_CP(S_DVEL, a.p.root_drainvel.map);
#undef PF_CP
    cp_async_commit();
}

// ---- the per-cell physics, W cells per call --------------------------------------------------------
template <bool SNOW, bool GROUND, bool INFIL, int RA /*0 table, 1 cell, 2 etref input*/, bool DIAG, bool USOIL, int W, bool PF = false>
__device__ __forceinline__ void wb_cells(const WbArgs &a, int64_t i, float4 *pf = nullptr, int64_t i_next = -1)
{
    using L = Lanes<W>;
    const sphy_wb_params &p = a.p;
    static_assert(!PF || W == 4, "the prefetch path moves 4 cells per thread");
    if (PF) cp_async_wait_all();                       // this tile's inputs have landed in the thread's slots
#define IN(slot, ptr) (PF ? pf_get(pf, slot) : L::in(ptr, i))
#define INRW(slot, ptr) (PF ? pf_get(pf, slot) : L::inrw(ptr, i))
    // ---- loads (all issued up front: ~14 independent 128-bit requests per thread) -------------------
    V4 prec = IN(S_PREC, a.f.prec), tavg = IN(S_TAVG, a.f.tavg);
    V4 tmin, tmax, etref_in, ra;
    if (RA == 2) {
        etref_in = IN(S_T1, a.f.etref);
    } else {
        tmin = IN(S_T1, a.f.tmin);
        tmax = IN(S_T2, a.f.tmax);
    }
    if (SNOW && RA == 2) tmax = IN(S_T2, a.f.tmax);
    V4 sinl, cosl, tanl;
    if (RA == 0) {
        if (W == 4) {
            uint2 c;
            if (PF) c = *reinterpret_cast<const uint2 *>(pf + S_LAT * WB_THREADS);
            else c = __ldg(reinterpret_cast<const uint2 *>(p.lat_class + i));
            ra.v[0] = __ldg(a.ra_table + (c.x & 0xffffu));
            ra.v[1] = __ldg(a.ra_table + (c.x >> 16));
            ra.v[2] = __ldg(a.ra_table + (c.y & 0xffffu));
            ra.v[3] = __ldg(a.ra_table + (c.y >> 16));
        } else if (W == 2) {
            unsigned c = __ldg(reinterpret_cast<const unsigned *>(p.lat_class + i));
            ra.v[0] = __ldg(a.ra_table + (c & 0xffffu));
            ra.v[1] = __ldg(a.ra_table + (c >> 16));
        } else {
            ra = splat(__ldg(a.ra_table + __ldg(p.lat_class + i)));
        }
    } else if (RA == 1) {
        sinl = L::in(p.sin_lat, i);
        cosl = L::in(p.cos_lat, i);
        tanl = L::in(p.tan_lat, i);
    }
    V4 RW = INRW(S_RW, a.s.RootWater), SW = INRW(S_SW, a.s.SubWater), CapRise = INRW(S_CAPR, a.s.CapRise),
       RootDrain = INRW(S_RDRAIN, a.s.RootDrain);
    V4 GwRecharge, Gw, BaseR, H_gw, SubDrain, SnowStore, SnowWatStore, TotalSnowStore;
    if (GROUND) {
        GwRecharge = INRW(S_G0, a.s.GwRecharge);
        Gw = INRW(S_G1, a.s.Gw);
        BaseR = INRW(S_G2, a.s.BaseR);
        H_gw = INRW(S_G3, a.s.H_gw);
    } else {
        SubDrain = INRW(S_G0, a.s.SubDrain);
    }
    if (SNOW) {
        SnowStore = INRW(S_SN0, a.s.SnowStore);
        SnowWatStore = INRW(S_SN1, a.s.SnowWatStore);
        TotalSnowStore = INRW(S_SN2, a.s.TotalSnowStore);
    }
    // soil / groundwater parameters: USOIL = every one of them is a scalar (uniform soils, the cfg defaults):
    // they are then read straight from the kernel-parameter constant bank and cost no registers and no bytes
    V4 root_field, root_sat, root_dry, root_wilt, root_ksat, e_root, sub_field, sub_sat, e_sub, capmax, kc;
    if (!USOIL) {
        root_field = L::par(p.root_field, i); root_sat = L::par(p.root_sat, i); root_dry = L::par(p.root_dry, i);
        root_wilt = L::par(p.root_wilt, i); root_ksat = L::par(p.root_ksat, i); e_root = L::par(p.e_root, i);
        sub_field = L::par(p.sub_field, i); sub_sat = L::par(p.sub_sat, i); e_sub = L::par(p.e_sub, i);
        capmax = L::par(p.caprise_max, i); kc = L::par(p.kc, i);
    }
    V4 drainvel = (PF && p.root_drainvel.map) ? pf_get(pf, S_DVEL) : L::par(p.root_drainvel, i);
#undef IN
#undef INRW
    // the slots are free again: start the copies of this thread's next tile (in flight while this one is computed)
    if (PF && i_next >= 0) pf_issue<SNOW, GROUND, RA>(a, i_next, pf);
#define PV(vec, par) (USOIL ? p.par.value : vec.v[j])
    const bool has_glac = p.glac_frac.map != nullptr;
    const bool has_ow = p.open_water_frac.map != nullptr;
    V4 glac_frac = L::par(p.glac_frac, i), ow = L::par(p.open_water_frac, i);
    V4 g_tss, g_snowr, g_perc, g_r;
    const bool glac_in = a.f.GlacR != nullptr;
    if (glac_in) {
        g_tss = L::in(a.f.TotalSnowStore_GLAC, i);
        g_snowr = L::in(a.f.SnowR_GLAC, i);
        g_perc = L::in(a.f.GlacPerc, i);
        g_r = L::in(a.f.GlacR, i);
    }
    V4 gw_sat, base_thresh, e_delta, e_alpha, h_den, seep, sub_drainvel, pmap, paved;
    if (GROUND) {
        if (!USOIL) {
            gw_sat = L::par(p.gw_sat, i); base_thresh = L::par(p.base_thresh, i); e_delta = L::par(p.e_delta, i);
            e_alpha = L::par(p.e_alpha, i); h_den = L::par(p.h_den, i);
        }
    } else {
        if (!USOIL) seep = L::par(p.seepage, i);
        sub_drainvel = L::par(p.sub_drainvel, i);
    }
    // snow parameters: scalars in every cfg the reference ships; maps are honoured (modules/snow.py:84-90)
    V4 tcrit_m, snow_sc_m, ddfs_m, snow_f_m;
    if (SNOW) {
        if (p.tcrit.map) tcrit_m = L::in(p.tcrit.map, i);
        if (p.snow_sc.map) snow_sc_m = L::in(p.snow_sc.map, i);
        if (p.ddfs.map) ddfs_m = L::in(p.ddfs.map, i);
        if (p.snow_f.map) snow_f_m = L::in(p.snow_f.map, i);
    }
#define PS(par) (p.par.map ? par##_m.v[j] : p.par.value)
    const bool pws = (p.flags & SPHY_WB_PWS) != 0;
    if (pws) pmap = L::par(p.pmap, i);
    if (INFIL) paved = L::par(p.paved_frac, i);
    V4 wtot;
    if (DIAG && a.s.waterbalanceTot) wtot = L::inrw(a.s.waterbalanceTot, i);

    V4 o_totr, o_rootrr, o_rootdr, o_rainr, o_snowr, o_etref, o_etact, o_actet, o_infil, o_rain, o_rperc, o_sperc, o_wbal, o_etow, o_prec;

#pragma unroll
    for (int j = 0; j < W; ++j) {
        const float one = 1.0f;
        const float gf = glac_frac.v[j];
        const float omg = has_glac ? one - gf : one;                      // 1 - GlacFrac
        const float omow = has_ow ? one - ow.v[j] : one;                   // 1 - openWaterFrac
        const float ss0 = SNOW ? SnowStore.v[j] : 0.0f;
        const float SnowFrac = ss0 > 0.0f ? omg : 0.0f;                    // sphy.py:732
        const float RainFrac = ss0 == 0.0f ? omg : 0.0f;                   // sphy.py:733
        const float Precip = sf::div_often_zero(prec.v[j], p.pcorr);       // sphy.py:757 (divides); dry days
        const float Temp = tavg.v[j] + p.tcorr;
        float TempMax = 0.0f, ETref;
        if (RA == 2) {
            ETref = etref_in.v[j];
            if (SNOW) TempMax = tmax.v[j] + p.tcorr;
        } else {
            const float TempMin = tmin.v[j] + p.tcorr;
            TempMax = tmax.v[j] + p.tcorr;
            const float Ra = (RA == 0) ? ra.v[j] : extrarad(sinl.v[j], cosl.v[j], tanl.v[j], a.d);
            ETref = sf::Hargreaves(a.d.hg_c, Ra, Temp, TempMax, TempMin);                   // hargreaves.py:43-47
        }
        // ---- snow, modules/snow.py:110-187 ----------------------------------------------------------
        float Rain = Precip, SnowR = 0.0f, SnowSoil = 0.0f, OldTotalSnowStore = 0.0f, tss = 0.0f;
        float ss = ss0, sws = 0.0f;
        if (SNOW) {
            sws = SnowWatStore.v[j];
            const float tcrit = PS(tcrit);
            const float Snow = Temp >= tcrit ? 0.0f : Precip;
            Rain = Temp < tcrit ? 0.0f : Precip;
            const float PotSnowMelt = TempMax < 0.0f ? 0.0f : sf::PotSnowMelt(Temp, TempMax, PS(ddfs), p.melt_cos);
            const float ActSnowMelt = sf::ActSnowMelt(ss, PotSnowMelt);
            ss = sf::SnowStoreUpdate(ss, Snow, ActSnowMelt, Temp, sws);
            const float MaxSWS = sf::MaxSnowWatStorage(PS(snow_sc), ss);
            const float OldSWS = sws;
            sws = sf::SnowWatStorage(TempMax, MaxSWS, sws, ActSnowMelt, Rain);          // tested on Tmax (quirk Q5)
            OldTotalSnowStore = TotalSnowStore.v[j];
            tss = sf::TotSnowStorage(ss, sws, SnowFrac, RainFrac);
            if (glac_in) tss = tss + g_tss.v[j];
            SnowR = sf::SnowR(sws, MaxSWS, ActSnowMelt, Rain, OldSWS, SnowFrac);
            if (glac_in) SnowR = SnowR + g_snowr.v[j];
            const float snow_f = PS(snow_f);
            SnowSoil = SnowR * snow_f;
            SnowR = SnowR * (p.snow_f.map ? 1.0f - snow_f : p.one_minus_snow_f);
            if (has_ow) SnowR = SnowR * omow;
        }
        const float ETpot = sf::ETpot(ETref, PV(kc, kc));                  // ET.py:24-26
        // ---- rootzone, sphy.py:904-981 + rootzone.py ---------------------------------------------------
        float rw = RW.v[j] + CapRise.v[j];
        if (SNOW) rw = rw + SnowSoil;
        const float rsat = PV(root_sat, root_sat), rfield = PV(root_field, root_field);
        float Infil, RootRunoff;
        if (INFIL)
            sf::RootRunoff_infil(p.k_eff, PV(root_ksat, root_ksat), rsat, rw, p.labda_infil, p.alpha, p.alpha_sq, Rain, paved.v[j],
                                 RootRunoff, Infil);
        else
            sf::RootRunoff_sat(RainFrac, Rain, rw, rsat, PV(root_ksat, root_ksat), RootRunoff, Infil);
        rw = RainFrac > 0.0f ? rw + Infil : rw;
        const float etreddry = pws ? sf::ks(rfield, PV(root_dry, root_dry), pmap.v[j], rw, ETpot)
                                   : sf::etreddry(rw, PV(root_dry, root_dry), PV(root_wilt, root_wilt));
        const float ETact = sf::ETact(ETpot, rw, rsat, etreddry, RainFrac);
        const float ActETact = ETact * RainFrac;
        rw = fmaxf(rw - ETact, 0.0f);
        const float er = PV(e_root, e_root);
        const float ex = fmaxf(rw - rfield, 0.0f);
        const float tD = sf::RootDrainage(rw, RootDrain.v[j], rfield, rsat, drainvel.v[j], er);
        float sw = SW.v[j];
        const float ssat = PV(sub_sat, sub_sat);
        const float tP = sf::RootPercolation(rw, sw, rfield, er, ssat);
        const float RootOut = tD + tP;
        const bool over = RootOut > ex;                                    // sphy.py:976-978
        float rdrain = tD, rootperc = tP;
        if (over) sf::CalcFrac(rw, rfield, tD, tP, rdrain, rootperc);      // only then used; elsewhere it is 0/0 (IEEE slow path)
        rw = rw - (rdrain + rootperc);
        // ---- subzone, sphy.py:991-1056 + subzone.py ----------------------------------------------------
        sw = sw + rootperc;
        if (!GROUND) sw = fminf(fmaxf(sw - PV(seep, seepage), 0.0f), ssat);
        const float sfield = PV(sub_field, sub_field);
        const float cr = sf::CapilRise(sfield, sw, PV(capmax, caprise_max), rw, rsat, rfield);
        sw = sw - cr;
        float subperc = 0.0f, ActSubPerc = 0.0f, sdrain = 0.0f;
        const float es = PV(e_sub, e_sub);
        if (GROUND) {
            subperc = sf::SubPercolation(sw, sfield, es, Gw.v[j], PV(gw_sat, gw_sat));
            ActSubPerc = has_glac ? subperc * omg : subperc;
            sw = sw - subperc;
        } else {
            sdrain = sf::SubDrainage(sw, sfield, ssat, sub_drainvel.v[j], SubDrain.v[j], es);
            sw = sw - sdrain;
        }
        // ---- runoff assembly, sphy.py:1063-1077 --------------------------------------------------------
        float RootRR = RootRunoff * RainFrac;
        if (has_ow) RootRR = RootRR * omow;
        float RootDR = has_glac ? rdrain * omg : rdrain;
        if (has_ow) RootDR = RootDR * omow;
        const float RainR = RootRR + RootDR;
        // ---- groundwater, modules/groundwater.py:90-125 -------------------------------------------------
        float baser, gw = 0.0f, gwr = 0.0f, hgw = 0.0f;
        if (GROUND) {
            const float ed = PV(e_delta, e_delta), ea = PV(e_alpha, e_alpha);
            float inflow = ActSubPerc;
            if (glac_in) inflow = inflow + g_perc.v[j];
            gwr = sf::GroundWaterRecharge(ed, GwRecharge.v[j], inflow);
            gw = Gw.v[j] + gwr;
            baser = sf::BaseFlow(gw, BaseR.v[j], gwr, PV(base_thresh, base_thresh), ea);
            gw = gw - baser;
            if (has_ow) baser = baser * omow;
            hgw = sf::HLevel(H_gw.v[j], ea, gwr, PV(h_den, h_den));
        } else {
            baser = sdrain;                                                // sphy.py:1085
        }
        float TotR = (baser + RainR) + SnowR;                              // sphy.py:1096
        if (glac_in) TotR = TotR + g_r.v[j];
        // ---- water balance, sphy.py:1114-1212 (GlacRetreat = 0 branches) ----------------------------------
        if (DIAG && (a.o.wbal || a.s.waterbalanceTot)) {
            float dS = (rw - RW.v[j]) + (sw - SW.v[j]);
            // the day after a glacier update the reference forms dS with the NEW GlacFrac map (glacier.py:766, sphy.py:1137)
            if (a.f.wbal_glac_frac && i + j < a.n) dS = dS * (one - __ldg(a.f.wbal_glac_frac + i + j));
            else if (has_glac) dS = dS * omg;
            float wb;
            if (GROUND) {
                const bool live = i + j < a.n;
                // GlacRetreat = 1 (sphy.py:1114-1146): oldGw is never advanced, the glacier snow store leaves
                // TotalSnowStore on the update day before dS is formed, and the ice-depth change joins dS
                const float old_gw = (a.f.wbal_old_gw && live) ? __ldg(a.f.wbal_old_gw + i + j) : Gw.v[j];
                dS = dS + (gw - old_gw);
                if (SNOW) {
                    const float tss_now = (a.f.wbal_snow_adjust && live) ? tss - __ldg(a.f.wbal_snow_adjust + i + j) : tss;
                    dS = dS + (tss_now - OldTotalSnowStore);
                }
                if (a.f.wbal_ice_delta && live) dS = dS + __ldg(a.f.wbal_ice_delta + i + j);
                wb = (((Precip - ActETact) - baser) - RainR) - SnowR;
                if (glac_in) wb = wb - g_r.v[j];
                wb = wb - dS;
            } else {
                if (SNOW) dS = dS + (tss - OldTotalSnowStore);
                wb = (((((Precip - ActETact) - baser) - RainR) - SnowR) - dS) - PV(seep, seepage);
            }
            o_wbal.v[j] = wb;
            if (a.s.waterbalanceTot) wtot.v[j] = wtot.v[j] + wb;
        }
        // ---- write back to the register tile ------------------------------------------------------------
        RW.v[j] = rw;
        SW.v[j] = sw;
        CapRise.v[j] = cr;
        RootDrain.v[j] = rdrain;
        if (GROUND) {
            GwRecharge.v[j] = gwr;
            Gw.v[j] = gw;
            BaseR.v[j] = baser;
            H_gw.v[j] = hgw;
        } else {
            SubDrain.v[j] = sdrain;
        }
        if (SNOW) {
            SnowStore.v[j] = ss;
            SnowWatStore.v[j] = sws;
            TotalSnowStore.v[j] = tss;
        }
        o_totr.v[j] = TotR;
        o_rootrr.v[j] = RootRR;
        o_rootdr.v[j] = RootDR;
        o_rainr.v[j] = RainR;
        o_snowr.v[j] = SnowR;
        if (DIAG) {
            o_etref.v[j] = ETref;
            o_etact.v[j] = ETact;
            o_actet.v[j] = ActETact;
            o_infil.v[j] = Infil;
            o_rain.v[j] = Rain;
            o_rperc.v[j] = rootperc;
            o_sperc.v[j] = ActSubPerc;
            o_etow.v[j] = ETref * p.kc_open_water;
            o_prec.v[j] = Precip;
        }
    }
#undef PV
#undef PS
    // ---- stores ----------------------------------------------------------------------------------------
    L::out(a.s.RootWater, i, RW);
    L::out(a.s.SubWater, i, SW);
    L::out(a.s.CapRise, i, CapRise);
    L::out(a.s.RootDrain, i, RootDrain);
    if (GROUND) {
        L::out(a.s.GwRecharge, i, GwRecharge);
        L::out(a.s.Gw, i, Gw);
        L::out(a.s.BaseR, i, BaseR);
        L::out(a.s.H_gw, i, H_gw);
    } else {
        L::out(a.s.SubDrain, i, SubDrain);
    }
    if (SNOW) {
        L::out(a.s.SnowStore, i, SnowStore);
        L::out(a.s.SnowWatStore, i, SnowWatStore);
        L::out(a.s.TotalSnowStore, i, TotalSnowStore);
    }
    if (DIAG && a.s.waterbalanceTot) L::out(a.s.waterbalanceTot, i, wtot);
    L::out(a.o.TotR, i, o_totr);
    if (a.o.RootRR) L::out(a.o.RootRR, i, o_rootrr);
    if (a.o.RootDR) L::out(a.o.RootDR, i, o_rootdr);
    if (a.o.RainR) L::out(a.o.RainR, i, o_rainr);
    if (a.o.SnowR) L::out(a.o.SnowR, i, o_snowr);
    if (DIAG) {
        if (a.o.ETref) L::out(a.o.ETref, i, o_etref);
        if (a.o.ETact) L::out(a.o.ETact, i, o_etact);
        if (a.o.ActETact) L::out(a.o.ActETact, i, o_actet);
        if (a.o.Infil) L::out(a.o.Infil, i, o_infil);
        if (a.o.Rain) L::out(a.o.Rain, i, o_rain);
        if (a.o.RootPerc) L::out(a.o.RootPerc, i, o_rperc);
        if (a.o.SubPerc) L::out(a.o.SubPerc, i, o_sperc);
        if (a.o.wbal) L::out(a.o.wbal, i, o_wbal);
        if (a.o.ETOpenWater) L::out(a.o.ETOpenWater, i, o_etow);
        if (a.o.Precip) L::out(a.o.Precip, i, o_prec);
    }
}

template <bool SNOW, bool GROUND, bool INFIL, int RA, bool DIAG, bool USOIL, int CPT>
__global__ void __launch_bounds__(WB_THREADS) k_wb_step(const __grid_constant__ WbArgs a)
{
    const int64_t ng = a.n / CPT;                       // groups of CPT consecutive cells
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t q = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; q < ng; q += stride)
        wb_cells<SNOW, GROUND, INFIL, RA, DIAG, USOIL, CPT>(a, q * CPT);
    // ragged tail (< CPT cells)
    const int64_t tail0 = ng * CPT;
    if (blockIdx.x == 0 && threadIdx.x < (a.n - tail0)) wb_cells<SNOW, GROUND, INFIL, RA, DIAG, USOIL, 1>(a, tail0 + threadIdx.x);
}

// the same kernel with the next tile prefetched through shared memory (S_COUNT 16-byte slots per thread)
constexpr int WB_PF_SMEM = S_COUNT * WB_THREADS * (int)sizeof(float4);

template <bool SNOW, bool GROUND, bool INFIL, int RA, bool USOIL>
__global__ void __launch_bounds__(WB_THREADS) k_wb_step_pf(const __grid_constant__ WbArgs a)
{
    extern __shared__ float4 pf_smem[];
    float4 *pf = pf_smem + threadIdx.x;
    const int64_t ng = a.n / 4;
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    int64_t q = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (q < ng) pf_issue<SNOW, GROUND, RA>(a, q * 4, pf);
    for (; q < ng; q += stride) {
        const int64_t qn = q + stride;
        wb_cells<SNOW, GROUND, INFIL, RA, false, USOIL, 4, true>(a, q * 4, pf, qn < ng ? qn * 4 : -1);
    }
    const int64_t tail0 = ng * 4;
    if (blockIdx.x == 0 && threadIdx.x < (a.n - tail0)) wb_cells<SNOW, GROUND, INFIL, RA, false, USOIL, 1>(a, tail0 + threadIdx.x);
}

typedef void (*wb_kernel_t)(const WbArgs);

template <bool S, bool G, bool U>
wb_kernel_t pick_sgu_pf(bool infil, int ra)
{
#define SPHY_WB_CASE(I, R) if (infil == I && ra == R) return k_wb_step_pf<S, G, I, R, U>;
    SPHY_WB_CASE(false, 0) SPHY_WB_CASE(false, 1) SPHY_WB_CASE(false, 2)
    SPHY_WB_CASE(true, 0) SPHY_WB_CASE(true, 1) SPHY_WB_CASE(true, 2)
#undef SPHY_WB_CASE
    return nullptr;
}

template <bool S, bool G, bool U, int CPT>
wb_kernel_t pick_sgu(bool infil, int ra, bool diag)
{
#define SPHY_WB_CASE(I, R, D) if (infil == I && ra == R && diag == D) return k_wb_step<S, G, I, R, D, U, CPT>;
    SPHY_WB_CASE(false, 0, false) SPHY_WB_CASE(false, 1, false) SPHY_WB_CASE(false, 2, false)
    SPHY_WB_CASE(true, 0, false) SPHY_WB_CASE(true, 1, false) SPHY_WB_CASE(true, 2, false)
    SPHY_WB_CASE(false, 0, true) SPHY_WB_CASE(false, 1, true) SPHY_WB_CASE(false, 2, true)
    SPHY_WB_CASE(true, 0, true) SPHY_WB_CASE(true, 1, true) SPHY_WB_CASE(true, 2, true)
#undef SPHY_WB_CASE
    return nullptr;
}

// cpt < 0 selects the prefetching variant (no diagnostics)
template <bool S, bool G>
wb_kernel_t pick_sg(bool infil, int ra, bool diag, bool usoil, int cpt)
{
    if (cpt < 0) return diag ? nullptr : (usoil ? pick_sgu_pf<S, G, true>(infil, ra) : pick_sgu_pf<S, G, false>(infil, ra));
    // measured on B200 (profiles/r1_k1_variants.md): 4 cells/thread at 96 registers (5 CTAs/SM) beats both
    // 2 cells/thread (80 registers) and a 6-CTA register cap (spills), so only CPT = 4 is instantiated
    (void)cpt;
    return usoil ? pick_sgu<S, G, true, 4>(infil, ra, diag) : pick_sgu<S, G, false, 4>(infil, ra, diag);
}

}  // namespace wb

// one per translation unit wb_inst_{s}{g}.cu
wb::wb_kernel_t wb_pick_00(bool infil, int ra, bool diag, bool usoil, int cpt);
wb::wb_kernel_t wb_pick_01(bool infil, int ra, bool diag, bool usoil, int cpt);
wb::wb_kernel_t wb_pick_10(bool infil, int ra, bool diag, bool usoil, int cpt);
wb::wb_kernel_t wb_pick_11(bool infil, int ra, bool diag, bool usoil, int cpt);
