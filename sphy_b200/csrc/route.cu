// route.cu -- K4: flow accumulation over the LDD as a topological level sweep, fused with the kx recession.
//
// Replaces pcr.accuflux / pcr.accufractionflux / pcr.upstream as called by routing.ROUT
// (/root/reference/modules/routing.py:25-29) and advanced_routing.ROUT (modules/advanced_routing.py:27-38).
// Cells are visited in routing order (position k in `order`, i.e. sorted by (level, cell)); a cell's flux is
// its own material plus the fluxes of its <=8 donors, summed in REAL8 and stored REAL4 (assumption A1), which
// makes the result independent of the donor order and bit-identical with oracle/ldd_oracle.c.
//   * wide levels (>= 2048 cells): one grid launch per level, one thread per cell;
//   * runs of narrow levels: ONE CTA walks the run level by level; 8 lanes per cell fetch the donors in
//     parallel and reduce with warp shuffles, the previous level's fluxes (the frontier) are staged in
//     shared memory so the dependent read does not go back to L2.
// All routed components (total + up to 6 fractions, routing.py:81-101) share one sweep: the index
// structures are read once.
#include <cstdio>

#include <cooperative_groups.h>
#include <cub/cub.cuh>
#include <cstdlib>

#include "common.cuh"

namespace cg = cooperative_groups;

namespace {

enum { MODE_PLAIN = 0, MODE_ROUT = 1, MODE_ADV = 2, MODE_CAPACITY = 3 };

struct SweepArgs {
    int ncomp;
    int mode;
    const float *m[ROUTE_MAX_COMP];      // material per component, cell order
    float *qstate[ROUTE_MAX_COMP];       // MODE_ROUT/ADV: Qold in, Q out (cell order)
    float *flux_out[ROUTE_MAX_COMP];     // MODE_PLAIN: flux in cell order
    const float *frac;                   // accufractionflux fraction (cell order) or NULL
    const float *qout_dense;             // MODE_ADV: structure outflow, m3/day, cell order
    float *acc;                          // [ncomp][n] flux in routing order
    float *qday;                         // MODE_ADV: [n] blended Q in m3/day, routing order
    const int32_t *order, *dpos_ptr, *dpos, *level_ptr;
    const int32_t *rec;                  // static donor records of the tail (k_sweep_records), positions >= rec_k0
    int32_t rec_k0;
    sphy_param kx;
    float one_m_kx;
    float cellarea;
    int64_t n;
};

__device__ __forceinline__ float material(const SweepArgs &a, int c, int32_t i)
{
    float q = __ldg(a.m[c] + i);
    if (a.mode == MODE_ROUT) return ((q * 0.001f) * a.cellarea) / 86400.0f;   // routing.py:26
    return q;
}

// what `finish` needs from global memory for one cell (component 0 can have it prefetched a level ahead)
struct CellIn {
    float fr, kx, qold, qout;
};

__device__ __forceinline__ CellIn load_cell_in(const SweepArgs &a, int c, int32_t i)
{
    CellIn ci;
    ci.fr = a.frac ? __ldg(a.frac + i) : 1.0f;
    ci.kx = a.kx.map ? __ldg(a.kx.map + i) : a.kx.value;
    ci.qold = (a.mode == MODE_PLAIN || a.mode == MODE_CAPACITY) ? 0.0f : a.qstate[c][i];
    ci.qout = (a.mode == MODE_ADV && ci.fr == 0.0f) ? __ldg(a.qout_dense + i) : 0.0f;
    return ci;
}

// what leaves a cell given what arrived (REAL8 sum): the REAL4 flux handed to the receiver
__device__ __forceinline__ float flux_of(int mode, bool has_frac, double tot, float fr)
{
    if (mode == MODE_CAPACITY) return tot < (double)fr ? (float)tot : fr;
    return has_frac ? (float)((double)fr * tot) : (float)tot;
}

// finish a cell: fraction cut, store flux, recession blend
__device__ __forceinline__ float finish(const SweepArgs &a, int c, int64_t k, int32_t i, double tot, const CellIn &ci)
{
    if (a.mode == MODE_CAPACITY) {
        // accucapacityflux / accucapacitystate: what arrives leaves up to the transport capacity (held in `frac`),
        // the rest stays behind (modules/sediment_transport.py:40-44)
        const float f = tot < (double)ci.fr ? (float)tot : ci.fr;
        a.acc[(int64_t)c * a.n + k] = f;
        if (a.flux_out[c]) a.flux_out[c][i] = f;
        if (a.qstate[c]) a.qstate[c][i] = (float)(tot - (double)f);
        return f;
    }
    const float f = a.frac ? (float)((double)ci.fr * tot) : (float)tot;
    a.acc[(int64_t)c * a.n + k] = f;
    if (a.mode == MODE_PLAIN) {
        if (a.flux_out[c]) a.flux_out[c][i] = f;
    } else {
        const float omk = a.kx.map ? 1.0f - ci.kx : a.one_m_kx;
        float *qs = a.qstate[c];
        if (a.mode == MODE_ROUT) {
            qs[i] = omk * f + ci.kx * ci.qold;                                 // routing.py:28
        } else {
            // advanced_routing.py:31-37: Q = QFRAC==0 ? Qout : (1-kx)*Q + (Qold*24*3600)*kx ; Q/86400 is the state
            const float qd = ci.fr == 0.0f ? ci.qout : omk * f + ((ci.qold * 24.0f) * 3600.0f) * ci.kx;
            a.qday[k] = qd;
            qs[i] = qd / 86400.0f;
        }
    }
    return f;
}

// one wide level: positions [k0, k1)
__global__ void __launch_bounds__(256) k_sweep_wide(const __grid_constant__ SweepArgs a, int32_t k0, int32_t k1)
{
    int64_t k = k0 + blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (k >= k1) return;
    const int32_t i = __ldg(a.order + k);
    const int32_t b = __ldg(a.dpos_ptr + k), e = __ldg(a.dpos_ptr + k + 1);
    for (int c = 0; c < a.ncomp; ++c) {
        double s = (double)material(a, c, i);
        const float *acc = a.acc + (int64_t)c * a.n;
        for (int32_t q = b; q < e; ++q) s += (double)acc[__ldg(a.dpos + q)];
        finish(a, c, k, i, s, load_cell_in(a, c, i));
    }
}

constexpr int SMALL_LEVEL_MAX = 1024 / 8;   // levels up to 128 cells: 8 donor lanes per cell, software-pipelined

// static donor records of the tail: rec[(k - k0) * 8 + s] = routing position of donor s of the cell at routing position
// k (or -1), for every position k >= k0 (the first level with <= SMALL_LEVEL cells).  One coalesced load replaces the
// dpos_ptr -> dpos chain on the latency-critical path of the narrow sweep.
__global__ void k_sweep_records(const int32_t *__restrict__ dpos_ptr, const int32_t *__restrict__ dpos, int64_t k0, int64_t n,
                                int32_t *__restrict__ rec)
{
    const int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (t >= (n - k0) * 8) return;
    const int64_t k = k0 + (t >> 3);
    const int s = (int)(t & 7);
    const int32_t b = __ldg(dpos_ptr + k), e = __ldg(dpos_ptr + k + 1);
    rec[t] = b + s < e ? __ldg(dpos + b + s) : -1;
}

// The tail levels are latency-bound (one memory round trip per level), and what they read -- materials, old discharges,
// donor records, the fluxes of tributary outlets -- is scattered and was evicted from L2 by the streaming kernels before.
// One massively parallel pass pulls those lines into the 126 MB L2 right before the serial walk starts, so that the
// walk's round trips are L2 hits instead of HBM accesses.
__device__ __forceinline__ void prefetch_l2(const void *p) { asm volatile("prefetch.global.L2 [%0];" ::"l"(p)); }

__global__ void k_sweep_warm(const __grid_constant__ SweepArgs a)
{
    const int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (t >= (a.n - a.rec_k0) * 8) return;
    const int64_t k = a.rec_k0 + (t >> 3);
    const int sub = (int)(t & 7);
    const int32_t dp = __ldg(a.rec + t);
    if (dp >= 0 && dp < a.rec_k0)                               // donors inside the tail are produced by the walk itself
        for (int c = 0; c < a.ncomp; ++c) prefetch_l2(a.acc + (int64_t)c * a.n + dp);
    if (sub == 0) {
        const int32_t i = __ldg(a.order + k);
        for (int c = 0; c < a.ncomp; ++c) {
            prefetch_l2(a.m[c] + i);
            if (a.qstate[c]) prefetch_l2(a.qstate[c] + i);
        }
        if (a.kx.map) prefetch_l2(a.kx.map + i);
        if (a.frac) prefetch_l2(a.frac + i);
        if (a.qout_dense) prefetch_l2(a.qout_dense + i);
    }
}

// A run of narrow levels [l0, l0+nl) walked by ONE CTA (the latency chain of the sweep: ~10^4 dependent levels).
//  * small levels (<= 128 cells, the long tail): 8 lanes per cell fetch the donors in parallel and reduce with warp
//    shuffles.  Everything that does not depend on the previous level is prefetched in a two-level software pipeline:
//      level l+2: cell index and donor record (static arrays, one hop);
//      level l+1: material, old discharge, kx/fraction, and the fluxes of donors older than level l (final already);
//      level l  : shared-memory read of the previous level's frontier -> 3 shuffles -> REAL8 add -> store -> barrier.
//    so a level costs one memory latency instead of a chain of four dependent loads.
//  * medium levels (129 .. 2047 cells): one thread per cell, sequential donor loop.
// The fluxes of the previous level (the frontier) are staged in shared memory; older ones come from global memory.
struct LevelRegs {            // what one lane holds for one small level in flight
    int32_t i, dp;            // cell (lane sub == 0 uses it), donor position of this lane
    float old, mat;
    bool valid, old_valid;    // valid: hop 1 done; old_valid: the donor was final when hop 2 ran
    CellIn ci;
};

// THREADS = 1024 (any narrow level), 256 or 32 (runs whose levels all have <= THREADS/8 cells: the long tail of a river
// network has 1-4 cells per level, and a 32-warp CTA would spend its time issuing the idle warps' loop bodies).
template <int THREADS>
__global__ void __launch_bounds__(THREADS) k_sweep_narrow(const __grid_constant__ SweepArgs a, int32_t l0, int32_t nl)
{
    constexpr int NARROW_THREADS = THREADS;
    constexpr int SMALL_LEVEL = THREADS / 8;
    constexpr int FRONT_CAP = THREADS == 1024 ? ROUTE_WIDE_LEVEL : THREADS / 8;
    extern __shared__ float front[];     // [2][ncomp][FRONT_CAP] fluxes of the previous / current level
    const int tid = threadIdx.x;
    const int sub = tid & 7;             // donor lane
    const int slot = tid >> 3;           // cell slot (128 cells)
    const int ncomp = a.ncomp;
    const int32_t lend = l0 + nl;
    int cur = 0;
    int32_t prev_begin = -1, prev_end = -1;
    // level boundaries k[l], k[l+1], k[l+2], k[l+3] (loaded ahead: they sit on the prefetch path)
    auto lp = [&](int32_t l) { return __ldg(a.level_ptr + (l < lend ? l : lend)); };
    int32_t kb = lp(l0), ke = lp(l0 + 1), k2 = lp(l0 + 2), k3 = lp(l0 + 3);
    LevelRegs A = {}, B = {};            // A: level l+2 (hop 1 done), B: level l+1 (hops 1 and 2 done) -- after the shifts below
    // hop 1 of level `lev` = positions [b, e): cell index + donor record
    auto hop1 = [&](int32_t b, int32_t e, LevelRegs &r) {
        r.valid = false; r.old_valid = false; r.i = 0; r.dp = -1;
        if (e - b > SMALL_LEVEL || e <= b || b < a.rec_k0) return;
        r.valid = true;
        if (slot < e - b) {
            const int32_t k = b + slot;
            r.i = __ldg(a.order + k);
            r.dp = __ldg(a.rec + ((int64_t)(k - a.rec_k0) << 3) + sub);
        }
    };
    // hop 2 of the level whose hop 1 is in r; `final_below`: donors at positions < final_below are final and in global memory
    auto hop2 = [&](int32_t b, int32_t e, int32_t final_below, LevelRegs &r) {
        if (!r.valid || slot >= e - b) return;
        if (r.dp >= 0 && r.dp < final_below) { r.old = a.acc[r.dp]; r.old_valid = true; }
        if (sub == 0) { r.mat = material(a, 0, r.i); r.ci = load_cell_in(a, 0, r.i); }
    };
    // fill the pipeline: level l0 completely (nothing is in flight yet), level l0+1 hop 1
    LevelRegs Cr = {};
    hop1(kb, ke, Cr);
    hop2(kb, ke, kb, Cr);                // every donor of the run's first level is older than the run
    hop1(ke, k2, B);
    for (int32_t l = l0; l < lend; ++l) {
        const int32_t nlev = ke - kb;
        float *fcur = front + (size_t)cur * ncomp * FRONT_CAP;
        const float *fprev = front + (size_t)(cur ^ 1) * ncomp * FRONT_CAP;
        // ---- issue: hop 1 of level l+2, hop 2 of level l+1 (its donors below kb are final; those in level l come from smem)
        hop1(k2, k3, A);
        hop2(ke, k2, kb, B);
        const int32_t k4 = lp(l + 4);
        // ---- the dependent part of level l ----------------------------------------------------------------------------
        const bool small = Cr.valid;
        if (small) {
            const bool live = slot < nlev;
            for (int c = 0; c < ncomp; ++c) {
                double s = 0.0;
                if (Cr.dp >= 0) {
                    if (Cr.dp >= prev_begin && Cr.dp < prev_end) s = (double)fprev[c * FRONT_CAP + (Cr.dp - prev_begin)];
                    else if (c == 0 && Cr.old_valid) s = (double)Cr.old;
                    else s = (double)a.acc[(int64_t)c * a.n + Cr.dp];
                }
                s += __shfl_xor_sync(0xffffffffu, s, 4, 8);       // warp-shuffle reduction over the 8 donor lanes
                s += __shfl_xor_sync(0xffffffffu, s, 2, 8);
                s += __shfl_xor_sync(0xffffffffu, s, 1, 8);
                if (live && sub == 0) {
                    s += (double)(c == 0 ? Cr.mat : material(a, c, Cr.i));
                    const float f = finish(a, c, kb + slot, Cr.i, s, c == 0 ? Cr.ci : load_cell_in(a, c, Cr.i));
                    fcur[c * FRONT_CAP + slot] = f;
                }
            }
        } else {
            for (int32_t k = kb + tid; k < ke; k += NARROW_THREADS) {
                const int32_t ic = __ldg(a.order + k);
                const int32_t b = __ldg(a.dpos_ptr + k), e = __ldg(a.dpos_ptr + k + 1);
                for (int c = 0; c < ncomp; ++c) {
                    double s = (double)material(a, c, ic);
                    for (int32_t q = b; q < e; ++q) {
                        const int32_t d = __ldg(a.dpos + q);
                        s += (d >= prev_begin && d < prev_end) ? (double)fprev[c * FRONT_CAP + (d - prev_begin)]
                                                               : (double)a.acc[(int64_t)c * a.n + d];
                    }
                    const float f = finish(a, c, k, ic, s, load_cell_in(a, c, ic));
                    fcur[c * FRONT_CAP + (k - kb)] = f;
                }
            }
        }
        prev_begin = kb;
        prev_end = ke;
        cur ^= 1;
        kb = ke; ke = k2; k2 = k3; k3 = k4;
        Cr = B; B = A;
        __syncthreads();
    }
}

// ---- deep-pipelined walk of the tail ------------------------------------------------------------------------------------
// Levels of at most CELLS = 4 * WARPS cells, one routed component.  A level costs one memory round trip as long as it
// consumes what was requested one level earlier, so the requests are issued D = 8 (dependent data) and 2 D = 16 (static
// records) levels ahead with cp.async into 16-entry shared-memory rings -- no registers are held across levels:
//   iteration l:  hop 2 of level l+8  (material, old discharge, kx / fraction of its cells; fluxes of donors that are final, i.e.
//                                      below level l -- younger donors are at most 8 levels back and sit in the flux ring)
//                 compute level l     (8 donor lanes per cell, shuffle reduction in REAL8, recession blend, flux ring)
//                 hop 1 of level l+16 (cell index and donor record from the static arrays)
// Every cp.async result is consumed by the lane that requested it; only the flux ring crosses lanes (one barrier per level).
__device__ __forceinline__ void cp_async4(void *dst, const void *src)
{
    asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"((unsigned)__cvta_generic_to_shared(dst)), "l"(src) : "memory");
}

template <int WARPS>
__global__ void __launch_bounds__(WARPS * 32) k_sweep_tail(const __grid_constant__ SweepArgs a, int32_t l0, int32_t nl)
{
    constexpr int T = WARPS * 32, CELLS = T / 8, D = 8, SLOTS = 2 * D, RINGP = CELLS * SLOTS;
    __shared__ int32_t s_dp[SLOTS][T];
    __shared__ float s_old[SLOTS][T];
    __shared__ int32_t s_i[SLOTS][CELLS];
    __shared__ float s_mat[SLOTS][CELLS], s_qold[SLOTS][CELLS], s_kx[SLOTS][CELLS], s_fr[SLOTS][CELLS], s_qout[SLOTS][CELLS];
    __shared__ float ringf[RINGP];       // fluxes of the last RINGP routing positions (tail positions are consecutive)
    const int tid = threadIdx.x, sub = tid & 7, slot = tid >> 3;
    const int32_t lend = l0 + nl;
    auto KB = [&](int32_t l) { return __ldg(a.level_ptr + (l < l0 ? l0 : (l > lend ? lend : l))); };
    auto barrier = [&]() { if (WARPS == 1) __syncwarp(); else __syncthreads(); };
    const bool has_qold = a.mode == MODE_ROUT || a.mode == MODE_ADV;
    for (int32_t l = l0 - 2 * D; l < lend; ++l) {
        asm volatile("cp.async.wait_group %0;" ::"n"(D - 1) : "memory");      // the requests of iteration l - D have landed
        // ---- hop 2 of level x = l + D: its donors below level l are final and in global memory
        const int32_t x = l + D;
        if (x >= l0 && x < lend) {
            const int32_t kbx = KB(x), nx = KB(x + 1) - kbx, final_below = KB(l);
            if (slot < nx) {
                const int sx = x & (SLOTS - 1);
                const int32_t dp = s_dp[sx][tid];
                if (dp >= 0 && dp < final_below) cp_async4(&s_old[sx][tid], a.acc + dp);
                if (sub == 0) {
                    const int32_t i = s_i[sx][slot];
                    cp_async4(&s_mat[sx][slot], a.m[0] + i);
                    if (has_qold) cp_async4(&s_qold[sx][slot], a.qstate[0] + i);
                    if (a.kx.map) cp_async4(&s_kx[sx][slot], a.kx.map + i);
                    if (a.frac) cp_async4(&s_fr[sx][slot], a.frac + i);
                    if (a.mode == MODE_ADV) cp_async4(&s_qout[sx][slot], a.qout_dense + i);
                }
            }
        }
        // ---- level l
        if (l >= l0) {
            const int32_t kb = KB(l), nlev = KB(l + 1) - kb, fetched_below = KB(l - D);
            const int sl = l & (SLOTS - 1);
            double s = 0.0;
            if (slot < nlev) {
                const int32_t dp = s_dp[sl][tid];
                if (dp >= 0) s = dp < fetched_below ? (double)s_old[sl][tid] : (double)ringf[dp & (RINGP - 1)];
            }
            s += __shfl_xor_sync(0xffffffffu, s, 4, 8);           // warp-shuffle reduction over the 8 donor lanes
            s += __shfl_xor_sync(0xffffffffu, s, 2, 8);
            s += __shfl_xor_sync(0xffffffffu, s, 1, 8);
            if (slot < nlev && sub == 0) {
                const int32_t i = s_i[sl][slot];
                float mat = s_mat[sl][slot];
                if (a.mode == MODE_ROUT) mat = ((mat * 0.001f) * a.cellarea) / 86400.0f;      // routing.py:26
                CellIn ci;
                ci.fr = a.frac ? s_fr[sl][slot] : 1.0f;
                ci.kx = a.kx.map ? s_kx[sl][slot] : a.kx.value;
                ci.qold = has_qold ? s_qold[sl][slot] : 0.0f;
                ci.qout = (a.mode == MODE_ADV && ci.fr == 0.0f) ? s_qout[sl][slot] : 0.0f;
                const float f = finish(a, 0, kb + slot, i, s + (double)mat, ci);
                ringf[(kb + slot) & (RINGP - 1)] = f;
            }
        }
        // ---- hop 1 of level y = l + 2 D: static records (its ring entry was the one of level l, consumed above)
        const int32_t y = l + 2 * D;
        if (y >= l0 && y < lend) {
            const int32_t kby = KB(y), ny = KB(y + 1) - kby;
            if (slot < ny) {
                const int sy = y & (SLOTS - 1);
                const int32_t k = kby + slot;
                cp_async4(&s_dp[sy][tid], a.rec + ((int64_t)(k - a.rec_k0) << 3) + sub);
                if (sub == 0) cp_async4(&s_i[sy][slot], a.order + k);
            }
        }
        asm volatile("cp.async.commit_group;" ::: "memory");
        barrier();
    }
    asm volatile("cp.async.wait_group 0;" ::: "memory");
}

// ---- the narrow end of the level order as heavy-path chains --------------------------------------------------------------
// Level widths never increase, so from the first level of at most CHAIN_MAX cells on (`level0`, routing position k0) the
// rest of the basin -- 17 000 of the 18 000 levels of the 10 000 x 10 000 basin, but only 0.5 % of its cells -- is a
// forest of at most CHAIN_MAX "chains": every cell of level > level0 continues the chain of its HEAVY donor, the donor
// that sets its level (lowest routing position among the donors of level - 1); its other donors either lie below k0
// (final before the walk starts) or are the last cells of other chains.  One warp walks one chain in blocks of 32 cells:
//   parallel part    lane j gathers what cell j receives apart from its heavy donor: own material + the fluxes of the
//                    other donors (a donor inside the tail is awaited: its flux slot holds a sentinel until written);
//   sequential part  32 steps of  tot = lateral_j + flux_{j-1};  flux_j = REAL4(tot)  -- the only truly serial work of
//                    the whole sweep: one shuffle-fed float64 add and one conversion per cell, every lane redundantly;
//   parallel part    lane j finishes cell j (store, recession blend) exactly like the level kernels do.
// The REAL8 sum of a cell adds the same terms as everywhere else (order does not matter: assumption A1), so the result
// stays bit-identical with the oracle -- a level costs ~40 cycles of latency instead of a CTA barrier and 300 instructions.
constexpr int CHAIN_MAX = 1024;      // upper bound; build_chains lowers it until every long chain's CTA is resident at once
constexpr unsigned FLUX_PENDING = 0xFFFFFFFFu;      // no arithmetic produces this NaN payload

struct ChainArgs {
    int32_t n_chains, k0;
    const int32_t *ids;        // chains handled by this launch (n_chains of them)
    const int32_t *start;      // [n_chains + 1] first slot of every chain
    const int32_t *k;          // [slots] routing position
    const int32_t *cell;       // [slots] cell (device order)
    const int32_t *don;        // [slots][8] routing positions of the donors other than the heavy one, -1 padded
};

__device__ __forceinline__ float load_flux_wait(const float *p)
{
    unsigned v;
    for (;;) {
        asm volatile("ld.volatile.global.u32 %0, [%1];" : "=r"(v) : "l"(p));
        if (v != FLUX_PENDING) break;
        __nanosleep(40);
    }
    return __uint_as_float(v);
}

// the lateral inflow of one cell: own material + every donor but the heavy one.  All donor loads are issued before any of
// them is examined, so their latencies overlap; only slots that still hold the sentinel are polled again.
__device__ __forceinline__ double chain_lateral(const float *acc, const int32_t (&d)[8], int32_t k0, double own)
{
    unsigned v[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
        v[j] = 0u;
        if (d[j] >= 0) asm volatile("ld.volatile.global.u32 %0, [%1];" : "=r"(v[j]) : "l"(acc + d[j]));
    }
    double s = own;
#pragma unroll
    for (int j = 0; j < 8; ++j) {
        if (d[j] < 0) continue;
        float f = __uint_as_float(v[j]);
        if (d[j] >= k0 && v[j] == FLUX_PENDING) f = load_flux_wait(acc + d[j]);
        s += (double)f;
    }
    return s;
}

constexpr int CHAIN_WARPS = 4;

__global__ void __launch_bounds__(CHAIN_WARPS * 32) k_sweep_chains(const __grid_constant__ SweepArgs a, const __grid_constant__ ChainArgs ch)
{
    const int lane = threadIdx.x & 31;
    const int32_t wi = blockIdx.x * CHAIN_WARPS + (threadIdx.x >> 5);
    if (wi >= ch.n_chains) return;
    const int32_t w = __ldg(ch.ids + wi);
    const int32_t s0 = __ldg(ch.start + w), len = __ldg(ch.start + w + 1) - s0;
    const bool has_frac = a.frac != nullptr;
    double prev[ROUTE_MAX_COMP];
#pragma unroll
    for (int c = 0; c < ROUTE_MAX_COMP; ++c) prev[c] = 0.0;
    // static record of this lane's cell in the next block (one hop ahead of the flux gathers)
    int32_t nk = 0, ni = 0, nd[8];
    auto hop1 = [&](int32_t b) {
        const bool v = b + lane < len;
        const int32_t s = s0 + b + lane;
        nk = v ? __ldg(ch.k + s) : 0;
        ni = v ? __ldg(ch.cell + s) : 0;
        const int4 d0 = v ? __ldg(reinterpret_cast<const int4 *>(ch.don + (int64_t)s * 8)) : make_int4(-1, -1, -1, -1);
        const int4 d1 = v ? __ldg(reinterpret_cast<const int4 *>(ch.don + (int64_t)s * 8) + 1) : make_int4(-1, -1, -1, -1);
        nd[0] = d0.x; nd[1] = d0.y; nd[2] = d0.z; nd[3] = d0.w; nd[4] = d1.x; nd[5] = d1.y; nd[6] = d1.z; nd[7] = d1.w;
    };
    hop1(0);
    for (int32_t b = 0; b < len; b += 32) {
        const int cnt = min(32, len - b);
        const bool valid = lane < cnt;
        const int32_t k = nk, i = ni;
        int32_t d[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) d[j] = nd[j];
        if (b + 32 < len) hop1(b + 32);
        // ---- parallel: lateral inflow of every cell of the block
        double lat[ROUTE_MAX_COMP];
        CellIn ci[ROUTE_MAX_COMP];
        for (int c = 0; c < a.ncomp; ++c) {
            double s = 0.0;
            if (valid) {
                ci[c] = load_cell_in(a, c, i);
                s = chain_lateral(a.acc + (int64_t)c * a.n, d, ch.k0, (double)material(a, c, i));
            }
            lat[c] = s;
        }
        // ---- sequential: the chain itself
        double mytot[ROUTE_MAX_COMP];
        for (int j = 0; j < cnt; ++j) {
            for (int c = 0; c < a.ncomp; ++c) {
                const double lj = __shfl_sync(0xffffffffu, lat[c], j);
                const float frj = (has_frac || a.mode == MODE_CAPACITY) ? __shfl_sync(0xffffffffu, ci[c].fr, j) : 1.0f;
                const double tot = lj + prev[c];
                if (lane == j) mytot[c] = tot;
                prev[c] = (double)flux_of(a.mode, has_frac, tot, frj);
            }
        }
        // ---- parallel: store, blend
        if (valid)
            for (int c = 0; c < a.ncomp; ++c) finish(a, c, k, i, mytot[c], ci[c]);
    }
}

// Long chains (the main stem: 17 000 cells) with WARP SPECIALISATION: what bounds a chain is the serial float64 add +
// conversion per cell, ~30 cycles -- but a single warp that also gathers the lateral inflows and finishes the cells spends
// ~1000 instructions at ~10 cycles each per block of 32 cells.  Here one CTA walks one chain: four producer/finisher
// warps take every fourth block (gather the laterals into a shared-memory slot, later finish the block's cells from the
// totals), and ONE consumer warp does nothing but the serial recurrence, block after block, out of shared memory.
// Slot s belongs to producer warp s: ready[s] = b + 1 announces the laterals of block b, done[s] = b + 1 its totals.
constexpr int CHAIN_PROD = 7;
constexpr int CHAIN_LONG = 64;             // chains longer than this get a CTA of their own

// NC = 1: one routed map (the usual case) -- every per-component array is a scalar in a register and the 32 serial steps are
// fully unrolled; NC = ROUTE_MAX_COMP: any number of components (arrays indexed at run time live in local memory).
template <int NC>
__global__ void __launch_bounds__((CHAIN_PROD + 1) * 32) k_sweep_chain_long(const __grid_constant__ SweepArgs a,
                                                                            const __grid_constant__ ChainArgs ch)
{
    __shared__ double s_lat[CHAIN_PROD][NC][32];
    __shared__ double s_tot[CHAIN_PROD][NC][32];
    __shared__ float s_fr[CHAIN_PROD][32];
    __shared__ volatile int s_ready[CHAIN_PROD], s_done[CHAIN_PROD];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int32_t w = __ldg(ch.ids + blockIdx.x);
    const int32_t s0 = __ldg(ch.start + w), len = __ldg(ch.start + w + 1) - s0;
    const int32_t nblocks = (len + 31) >> 5;
    const bool has_frac = a.frac != nullptr;
    const bool use_fr = has_frac || a.mode == MODE_CAPACITY;
    if (threadIdx.x < CHAIN_PROD) { s_ready[threadIdx.x] = 0; s_done[threadIdx.x] = 0; }
    __syncthreads();
    const int ncomp = NC == 1 ? 1 : a.ncomp;
    if (warp == CHAIN_PROD) {
        // ---- consumer: the chain itself
        double prev[NC];
#pragma unroll
        for (int c = 0; c < NC; ++c) prev[c] = 0.0;
        for (int32_t b = 0; b < nblocks; ++b) {
            const int slot = b % CHAIN_PROD;
            const int cnt = min(32, len - (b << 5));
            while (s_ready[slot] != b + 1) { }
            __syncwarp();
            double lat[NC], mytot[NC];
#pragma unroll
            for (int c = 0; c < NC; ++c) {
                lat[c] = c < ncomp ? s_lat[slot][c][lane] : 0.0;
                mytot[c] = 0.0;
            }
            const float fr = use_fr ? s_fr[slot][lane] : 1.0f;
            auto step = [&](int j) {
                const float frj = use_fr ? __shfl_sync(0xffffffffu, fr, j) : 1.0f;
#pragma unroll
                for (int c = 0; c < NC; ++c) {
                    if (c >= ncomp) break;
                    const double tot = __shfl_sync(0xffffffffu, lat[c], j) + prev[c];
                    if (lane == j) mytot[c] = tot;
                    prev[c] = (double)flux_of(a.mode, has_frac, tot, frj);
                }
            };
            if (cnt == 32) {
#pragma unroll
                for (int j = 0; j < 32; ++j) step(j);
            } else {
                for (int j = 0; j < cnt; ++j) step(j);
            }
#pragma unroll
            for (int c = 0; c < NC; ++c)
                if (c < ncomp) s_tot[slot][c][lane] = mytot[c];
            __threadfence_block();
            __syncwarp();
            if (lane == 0) s_done[slot] = b + 1;
        }
        return;
    }
    // ---- producer / finisher of every CHAIN_PROD-th block
    for (int32_t b = warp; b < nblocks; b += CHAIN_PROD) {
        const int slot = warp;
        const bool valid = (b << 5) + lane < len;
        const int32_t s = s0 + (b << 5) + lane;
        int32_t k = 0, i = 0, d[8];
        if (valid) {
            k = __ldg(ch.k + s);
            i = __ldg(ch.cell + s);
            const int4 d0 = __ldg(reinterpret_cast<const int4 *>(ch.don + (int64_t)s * 8));
            const int4 d1 = __ldg(reinterpret_cast<const int4 *>(ch.don + (int64_t)s * 8) + 1);
            d[0] = d0.x; d[1] = d0.y; d[2] = d0.z; d[3] = d0.w; d[4] = d1.x; d[5] = d1.y; d[6] = d1.z; d[7] = d1.w;
        }
        CellIn ci[NC];
#pragma unroll
        for (int c = 0; c < NC; ++c) {
            if (c >= ncomp) break;
            double sum = 0.0;
            if (valid) {
                ci[c] = load_cell_in(a, c, i);
                sum = chain_lateral(a.acc + (int64_t)c * a.n, d, ch.k0, (double)material(a, c, i));
            }
            s_lat[slot][c][lane] = sum;
        }
        s_fr[slot][lane] = valid ? ci[0].fr : 1.0f;
        __threadfence_block();
        __syncwarp();
        if (lane == 0) s_ready[slot] = b + 1;
        while (s_done[slot] != b + 1) __nanosleep(32);        // do not compete with the consumer warp for issue slots
        __syncwarp();
        if (valid) {
#pragma unroll
            for (int c = 0; c < NC; ++c)
                if (c < ncomp) finish(a, c, k, i, s_tot[slot][c][lane], ci[c]);
        }
        __syncwarp();
    }
}

// The levels between the per-level launches and the chains (a few hundred levels of 1 000 .. 65 000 cells): ONE cooperative
// launch walks them level by level with a grid barrier in between, instead of one dependent launch per level.
constexpr int MID_LEVEL_MAX = 1 << 16;       // levels up to 65 536 cells join the cooperative launch (one cell per thread on 148 x 512 threads; grid-stride if the grid is smaller)

constexpr int MID_THREADS = 512;

__global__ void __launch_bounds__(MID_THREADS) k_sweep_mid(const __grid_constant__ SweepArgs a, int32_t l0, int32_t l1)
{
    cg::grid_group grid = cg::this_grid();
    const int32_t gt = blockIdx.x * MID_THREADS + threadIdx.x;       // one cell per thread and level (widths <= grid size)
    // Everything that does not depend on the previous level -- the cell, its donors' routing positions, its material, old
    // discharge, kx / fraction -- is fetched one level ahead, before the barrier: behind the barrier a level costs one
    // round trip for the donors' fluxes plus the stores.  (Component 0 is prefetched; further components load in place.)
    struct Pre { int32_t k, i, nd, d[8]; float mat; CellIn ci; bool on, wide; };
    const int32_t gsz = gridDim.x * MID_THREADS;
    auto prefetch = [&](int32_t l) {
        Pre p;
        p.on = p.wide = false; p.k = p.i = p.nd = 0; p.mat = 0.0f; p.ci = CellIn{1.0f, 0.0f, 0.0f, 0.0f};
#pragma unroll
        for (int j = 0; j < 8; ++j) p.d[j] = -1;
        if (l < l1) {
            const int32_t kb = __ldg(a.level_ptr + l), ke = __ldg(a.level_ptr + l + 1);
            if (ke - kb > gsz) {                               // more cells than threads: plain grid-stride pass, no prefetch
                p.wide = true;
                p.k = kb;
                p.i = ke;
            } else if (kb + gt < ke) {
                p.on = true;
                p.k = kb + gt;
                p.i = __ldg(a.order + p.k);
                const int32_t b = __ldg(a.dpos_ptr + p.k), e = __ldg(a.dpos_ptr + p.k + 1);
                p.nd = e - b;
#pragma unroll
                for (int j = 0; j < 8; ++j)
                    if (j < p.nd) p.d[j] = __ldg(a.dpos + b + j);
                p.mat = material(a, 0, p.i);
                p.ci = load_cell_in(a, 0, p.i);
            }
        }
        return p;
    };
    Pre cur = prefetch(l0);
    for (int32_t l = l0; l < l1; ++l) {
        const Pre nxt = prefetch(l + 1);
        if (cur.wide) {
            for (int32_t k = cur.k + gt; k < cur.i; k += gsz) {
                const int32_t i = __ldg(a.order + k);
                const int32_t b = __ldg(a.dpos_ptr + k), e = __ldg(a.dpos_ptr + k + 1);
                for (int c = 0; c < a.ncomp; ++c) {
                    double s = (double)material(a, c, i);
                    const float *acc = a.acc + (int64_t)c * a.n;
                    for (int32_t q = b; q < e; ++q) s += (double)__ldcg(acc + __ldg(a.dpos + q));
                    finish(a, c, k, i, s, load_cell_in(a, c, i));
                }
            }
        } else if (cur.on) {
            for (int c = 0; c < a.ncomp; ++c) {
                const float *acc = a.acc + (int64_t)c * a.n;
                float f[8];
#pragma unroll
                for (int j = 0; j < 8; ++j) f[j] = cur.d[j] >= 0 ? __ldcg(acc + cur.d[j]) : 0.0f;
                double s = (double)(c == 0 ? cur.mat : material(a, c, cur.i));
#pragma unroll
                for (int j = 0; j < 8; ++j)
                    if (cur.d[j] >= 0) s += (double)f[j];
                finish(a, c, cur.k, cur.i, s, c == 0 ? cur.ci : load_cell_in(a, c, cur.i));
            }
        }
        cur = nxt;
        grid.sync();
    }
}

// chain construction (once): heavy donor of every tail position, successor links
__global__ void k_chain_links(const int32_t *__restrict__ level_ptr, int32_t nlevels, const int32_t *__restrict__ dpos_ptr,
                              const int32_t *__restrict__ dpos, int32_t k0, int64_t n, int32_t *__restrict__ heavy, int32_t *__restrict__ next)
{
    const int64_t k = k0 + blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (k >= n) return;
    // level of routing position k: last l with level_ptr[l] <= k
    int32_t lo = 0, hi = nlevels;
    while (hi - lo > 1) {
        const int32_t mid = (lo + hi) >> 1;
        if (level_ptr[mid] <= k) lo = mid; else hi = mid;
    }
    const int32_t below_b = lo > 0 ? level_ptr[lo - 1] : 0, below_e = level_ptr[lo];      // positions of level - 1
    int32_t h = -1;
    for (int32_t q = dpos_ptr[k]; q < dpos_ptr[k + 1]; ++q) {
        const int32_t dp = dpos[q];
        if (dp >= k0 && dp >= below_b && dp < below_e && (h < 0 || dp < h)) h = dp;
    }
    heavy[k - k0] = h;
    if (h >= 0) next[h - k0] = (int32_t)k;
}

// one thread per chain (they start at the positions of level0): number the cells along the chain
__global__ void k_chain_walk(const int32_t *__restrict__ next, int32_t k0, int32_t n_chains, int32_t *__restrict__ chain_of,
                             int32_t *__restrict__ idx_in, int32_t *__restrict__ len)
{
    const int32_t c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= n_chains) return;
    int32_t k = k0 + c, l = 0;
    while (k >= 0) {
        chain_of[k - k0] = c;
        idx_in[k - k0] = l++;
        k = next[k - k0];
    }
    len[c] = l;
}

__global__ void k_chain_tables(const int32_t *__restrict__ order, const int32_t *__restrict__ dpos_ptr, const int32_t *__restrict__ dpos,
                               const int32_t *__restrict__ heavy, const int32_t *__restrict__ chain_of, const int32_t *__restrict__ idx_in,
                               const int32_t *__restrict__ start, int32_t k0, int64_t n, int32_t *__restrict__ ck, int32_t *__restrict__ ccell,
                               int32_t *__restrict__ cdon)
{
    const int64_t k = k0 + blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (k >= n) return;
    const int32_t s = start[chain_of[k - k0]] + idx_in[k - k0];
    ck[s] = (int32_t)k;
    ccell[s] = order[k];
    int j = 0;
    for (int32_t q = dpos_ptr[k]; q < dpos_ptr[k + 1]; ++q) {
        const int32_t dp = dpos[q];
        if (dp != heavy[k - k0] && j < 8) cdon[(int64_t)s * 8 + j++] = dp;
    }
    for (; j < 8; ++j) cdon[(int64_t)s * 8 + j] = -1;
}

__global__ void k_upstream(const float *__restrict__ x, const int32_t *__restrict__ don_ptr, const int32_t *__restrict__ don_idx,
                           int64_t n, float *__restrict__ out)
{
    int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (i >= n) return;
    double s = 0.0;
    for (int32_t q = don_ptr[i]; q < don_ptr[i + 1]; ++q) s += (double)x[don_idx[q]];
    out[i] = (float)s;
}

}  // namespace

static int build_chains_at(sphy_ctx *ctx, cudaStream_t st, int chain_max, int nlv);

// one-time construction of the chain tables (synchronises)
static int build_chains(sphy_ctx *ctx, cudaStream_t st)
{
    ctx->chains_ready = true;
    ctx->n_chains = 0;
    const char *env = getenv("SPHY_SWEEP_CHAINS");              // 0 = keep the level-by-level kernels for the narrow levels
    if (env && env[0] == '0') return SPHY_OK;
    const int nlv = (int)ctx->h_level_ptr.size() - 1;
    int per_sm1 = 0, per_sm8 = 0;
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm1, k_sweep_chain_long<1>, (CHAIN_PROD + 1) * 32, 0);
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm8, k_sweep_chain_long<ROUTE_MAX_COMP>, (CHAIN_PROD + 1) * 32, 0);
    const int resident = ctx->sm_count * (per_sm1 < per_sm8 ? per_sm1 : per_sm8);   // chains wait for one another: all or nothing
    for (int cap = CHAIN_MAX; cap >= 256; cap /= 2) {
        int rc = build_chains_at(ctx, st, cap, nlv);
        if (rc != SPHY_OK) return rc;
        if (ctx->n_chains == 0 || ctx->n_chains_long <= resident - ctx->sm_count) return SPHY_OK;
    }
    ctx->n_chains = 0;                                          // cannot happen on a real device; keep the level kernels
    return SPHY_OK;
}

static int build_chains_at(sphy_ctx *ctx, cudaStream_t st, int chain_max, int nlv)
{
    ctx->n_chains = 0;
    int l0 = -1;
    for (int l = 0; l < nlv; ++l)
        if (ctx->h_level_ptr[l + 1] - ctx->h_level_ptr[l] <= chain_max) { l0 = l; break; }
    if (l0 < 0) return SPHY_OK;
    const int32_t k0 = ctx->h_level_ptr[l0];
    const int64_t T = ctx->n - k0;
    const int32_t nch = ctx->h_level_ptr[l0 + 1] - k0;
    if (T <= 0 || nch <= 0 || T > ((int64_t)1 << 26)) return SPHY_OK;
    int32_t *heavy = nullptr, *next = nullptr, *chain_of = nullptr, *idx_in = nullptr, *len = nullptr;
    void *tmp = nullptr;
    size_t tb = 0;
    cudaError_t e = cudaSuccess;
#define FK(call) do { if (e == cudaSuccess) e = (call); } while (0)
    FK(cudaMalloc(&heavy, sizeof(int32_t) * T));
    FK(cudaMalloc(&next, sizeof(int32_t) * T));
    FK(cudaMalloc(&chain_of, sizeof(int32_t) * T));
    FK(cudaMalloc(&idx_in, sizeof(int32_t) * T));
    FK(cudaMalloc(&len, sizeof(int32_t) * (nch + 1)));
    int rc = SPHY_OK;
    if (e == cudaSuccess &&
        ((rc = dev_alloc(ctx, &ctx->chain_start, (size_t)nch + 1)) || (rc = dev_alloc(ctx, &ctx->chain_ids, (size_t)nch)) ||
         (rc = dev_alloc(ctx, &ctx->chain_k, (size_t)T)) ||
         (rc = dev_alloc(ctx, &ctx->chain_cell, (size_t)T)) || (rc = dev_alloc(ctx, &ctx->chain_don, (size_t)T * 8)))) {
        cudaFree(heavy); cudaFree(next); cudaFree(chain_of); cudaFree(idx_in); cudaFree(len);
        return rc;
    }
    if (e == cudaSuccess) {
        FK(cudaMemsetAsync(next, 0xFF, sizeof(int32_t) * T, st));
        FK(cudaMemsetAsync(len + nch, 0, sizeof(int32_t), st));
        k_chain_links<<<div_up(T, 256), 256, 0, st>>>(ctx->level_ptr, ctx->nlevels, ctx->dpos_ptr, ctx->dpos, k0, ctx->n, heavy, next);
        k_chain_walk<<<div_up(nch, 128), 128, 0, st>>>(next, k0, nch, chain_of, idx_in, len);
        FK(cub::DeviceScan::ExclusiveSum(nullptr, tb, len, ctx->chain_start, nch + 1, st));
        FK(cudaMalloc(&tmp, tb));
        FK(cub::DeviceScan::ExclusiveSum(tmp, tb, len, ctx->chain_start, nch + 1, st));
        k_chain_tables<<<div_up(T, 256), 256, 0, st>>>(ctx->order, ctx->dpos_ptr, ctx->dpos, heavy, chain_of, idx_in, ctx->chain_start, k0,
                                                      ctx->n, ctx->chain_k, ctx->chain_cell, ctx->chain_don);
        FK(cudaGetLastError());
        // short chains (one warp each) first, then the long ones (a warp-specialised CTA each): a short chain only ever
        // waits for other short chains (whatever it waits for ends at a lower level than it does), so the two launches
        // may run one after the other
        std::vector<int32_t> hlen(nch), ids;
        FK(cudaMemcpyAsync(hlen.data(), len, sizeof(int32_t) * nch, cudaMemcpyDeviceToHost, st));
        FK(cudaStreamSynchronize(st));
        if (e == cudaSuccess) {
            for (int32_t c = 0; c < nch; ++c) if (hlen[c] <= CHAIN_LONG) ids.push_back(c);
            ctx->n_chains_short = (int32_t)ids.size();
            for (int32_t c = 0; c < nch; ++c) if (hlen[c] > CHAIN_LONG) ids.push_back(c);
            ctx->n_chains_long = nch - ctx->n_chains_short;
            FK(cudaMemcpyAsync(ctx->chain_ids, ids.data(), sizeof(int32_t) * nch, cudaMemcpyHostToDevice, st));
        }
        FK(cudaStreamSynchronize(st));
    }
#undef FK
    cudaFree(heavy); cudaFree(next); cudaFree(chain_of); cudaFree(idx_in); cudaFree(len); cudaFree(tmp);
    if (e != cudaSuccess) SPHY_FAIL(ctx, SPHY_E_CUDA, "build_chains: %s", cudaGetErrorString(e));
    // the levels in front of the chains that are too narrow to be worth a launch each: one cooperative launch
    ctx->mid_level0 = -1;
    {
        int coop = 0, per_sm = 0;
        cudaDeviceGetAttribute(&coop, cudaDevAttrCooperativeLaunch, ctx->device);
        const char *menv = getenv("SPHY_SWEEP_MID");
        if (coop && !(menv && menv[0] == '0') &&
            cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k_sweep_mid, MID_THREADS, 0) == cudaSuccess && per_sm > 0) {
            ctx->mid_grid = ctx->sm_count;                     // one CTA per SM: the cheaper the barrier
            int lm = l0;
            const int64_t cap = MID_LEVEL_MAX;
            while (lm > 0 && ctx->h_level_ptr[lm] - ctx->h_level_ptr[lm - 1] <= cap) --lm;
            if (lm < l0) ctx->mid_level0 = lm;
        }
    }
    ctx->chain_k0 = k0;
    ctx->chain_level0 = l0;
    ctx->n_chains = nch;
    ctx->n_chain_cells = T;
    return SPHY_OK;
}

// run the level sweep described by `a` following the launch plan built in ldd_build.cu
int route_levels_run(sphy_ctx *ctx, SweepArgs &a, cudaStream_t st)
{
    if (!ctx->chains_ready) {
        int rc = build_chains(ctx, st);
        if (rc != SPHY_OK) return rc;
    }
    a.order = ctx->order;
    a.dpos_ptr = ctx->dpos_ptr;
    a.dpos = ctx->dpos;
    a.level_ptr = ctx->level_ptr;
    a.acc = ctx->acc;
    a.n = ctx->n;
    a.cellarea = ctx->cellarea;
    if (!ctx->sweep_rec_ready) {          // once: donor records of every position from the first small level onwards
        int64_t k0 = ctx->n;
        const int nlv = (int)ctx->h_level_ptr.size() - 1;
        const auto size_of = [&](int l) { return ctx->h_level_ptr[l + 1] - ctx->h_level_ptr[l]; };
        for (int l = 0; l < nlv; ++l)
            if (size_of(l) <= SMALL_LEVEL_MAX) { k0 = ctx->h_level_ptr[l]; break; }
        if (ctx->n - k0 > (int64_t)1 << 25) k0 = ctx->n;      // keep the table under 1 GB: beyond that the tail is not a tail
        ctx->sweep_rec_k0 = (int32_t)k0;
        if (k0 < ctx->n) {
            int rc = dev_alloc(ctx, &ctx->sweep_rec, (size_t)(ctx->n - k0) * 8);
            if (rc != SPHY_OK) return rc;
            k_sweep_records<<<div_up((ctx->n - k0) * 8, 256), 256, 0, st>>>(ctx->dpos_ptr, ctx->dpos, k0, ctx->n, ctx->sweep_rec);
        }
        // launch plan of the sweep: the runs of narrow levels are cut into sub-runs by CTA size
        ctx->sweep_plan.clear();
        for (const RoutePlanSeg &seg : ctx->plan) {
            if (seg.n_levels == 1 && size_of(seg.first_level) >= ROUTE_WIDE_LEVEL) {
                ctx->sweep_plan.push_back({seg.first_level, 1, 0});
                continue;
            }
            const auto cls = [&](int l) {
                const int sz = size_of(l);
                if (ctx->h_level_ptr[l] < k0) return 1024;
                return sz <= 4 ? 32 : (sz <= 32 ? 256 : 1024);
            };
            int l = seg.first_level;
            const int lend = seg.first_level + seg.n_levels;
            while (l < lend) {
                // a sub-run keeps the CTA size of its first level; levels that need a larger CTA end it, and so does a
                // stretch of at least 8 levels that would all fit a smaller one (short dips are not worth a launch)
                const int t = cls(l);
                int e = l + 1;
                while (e < lend && cls(e) <= t) {
                    if (cls(e) < t) {
                        int j = e;
                        while (j < lend && j - e < 8 && cls(j) < t) ++j;
                        if (j - e >= 8) break;
                    }
                    ++e;
                }
                ctx->sweep_plan.push_back({l, e - l, t});
                l = e;
            }
        }
        if (getenv("SPHY_DEBUG_PLAN")) {
            int nseg[4] = {0, 0, 0, 0}, nlev[4] = {0, 0, 0, 0};
            for (const SweepPlanSeg &q : ctx->sweep_plan) {
                const int c = q.threads == 0 ? 0 : (q.threads == 1024 ? 1 : (q.threads == 256 ? 2 : 3));
                ++nseg[c];
                nlev[c] += q.n_levels;
            }
            fprintf(stderr, "sweep plan: wide %d launches; 1024-thread %d runs / %d levels; 256-thread %d / %d; 32-thread %d / %d\n",
                    nseg[0], nseg[1], nlev[1], nseg[2], nlev[2], nseg[3], nlev[3]);
        }
        ctx->sweep_rec_ready = true;
    }
    a.rec = ctx->sweep_rec;
    a.rec_k0 = ctx->sweep_rec_k0;
    if (!ctx->sweep_attr_set) {        // per context (= per device): opt in to the large dynamic shared memory once
        SPHY_CUDA(ctx, cudaFuncSetAttribute(k_sweep_narrow<1024>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                            2 * ROUTE_MAX_COMP * ROUTE_WIDE_LEVEL * (int)sizeof(float)));
        ctx->sweep_attr_set = true;
    }
    bool warmed = false;
    const char *deep_env = getenv("SPHY_SWEEP_DEEP");           // 0 = the two-level pipeline for every narrow run
    const bool deep = !(deep_env && deep_env[0] == '0');
    const int32_t chains_from = ctx->n_chains > 0 ? ctx->chain_level0 : (int32_t)ctx->h_level_ptr.size() - 1;   // chains take over here
    const int32_t level_end = (ctx->n_chains > 0 && ctx->mid_level0 >= 0) ? ctx->mid_level0 : chains_from;       // ... after k_sweep_mid
    for (SweepPlanSeg seg : ctx->sweep_plan) {
        if (seg.first_level >= level_end) break;
        if (seg.first_level + seg.n_levels > level_end) seg.n_levels = level_end - seg.first_level;
        if (!warmed && a.rec && ctx->h_level_ptr[seg.first_level] >= a.rec_k0) {     // the tail starts here
            k_sweep_warm<<<div_up((a.n - a.rec_k0) * 8, 256), 256, 0, st>>>(a);
            warmed = true;
        }
        if (seg.threads == 0) {
            const int32_t k0 = ctx->h_level_ptr[seg.first_level], k1 = ctx->h_level_ptr[seg.first_level + 1];
            k_sweep_wide<<<div_up(k1 - k0, 256), 256, 0, st>>>(a, k0, k1);
        } else if (seg.threads == 1024) {
            k_sweep_narrow<1024><<<1, 1024, (size_t)2 * a.ncomp * ROUTE_WIDE_LEVEL * sizeof(float), st>>>(a, seg.first_level, seg.n_levels);
        } else if (seg.threads == 256) {
            if (a.ncomp == 1 && deep) k_sweep_tail<8><<<1, 256, 0, st>>>(a, seg.first_level, seg.n_levels);
            else k_sweep_narrow<256><<<1, 256, (size_t)2 * a.ncomp * 32 * sizeof(float), st>>>(a, seg.first_level, seg.n_levels);
        } else {
            if (a.ncomp == 1 && deep) k_sweep_tail<1><<<1, 32, 0, st>>>(a, seg.first_level, seg.n_levels);
            else k_sweep_narrow<32><<<1, 32, (size_t)2 * a.ncomp * 4 * sizeof(float), st>>>(a, seg.first_level, seg.n_levels);
        }
    }
    if (ctx->n_chains > 0 && ctx->mid_level0 >= 0) {
        int32_t l0 = ctx->mid_level0, l1 = chains_from;
        void *kargs[] = {(void *)&a, (void *)&l0, (void *)&l1};
        SPHY_CUDA(ctx, cudaLaunchCooperativeKernel((const void *)k_sweep_mid, dim3(ctx->mid_grid), dim3(MID_THREADS), kargs, 0, st));
    }
    if (ctx->n_chains > 0) {
        for (int c = 0; c < a.ncomp; ++c)                      // flux slots of the tail: pending until a chain writes them
            SPHY_CUDA(ctx, cudaMemsetAsync(a.acc + (int64_t)c * a.n + ctx->chain_k0, 0xFF, sizeof(float) * ctx->n_chain_cells, st));
        if (ctx->n_chains_short > 0) {
            ChainArgs ch = {ctx->n_chains_short, ctx->chain_k0, ctx->chain_ids, ctx->chain_start, ctx->chain_k, ctx->chain_cell, ctx->chain_don};
            k_sweep_chains<<<div_up(ctx->n_chains_short, CHAIN_WARPS), CHAIN_WARPS * 32, 0, st>>>(a, ch);
        }
        if (ctx->n_chains_long > 0) {
            ChainArgs ch = {ctx->n_chains_long, ctx->chain_k0, ctx->chain_ids + ctx->n_chains_short, ctx->chain_start, ctx->chain_k,
                            ctx->chain_cell, ctx->chain_don};
            if (a.ncomp == 1) k_sweep_chain_long<1><<<ctx->n_chains_long, (CHAIN_PROD + 1) * 32, 0, st>>>(a, ch);
            else k_sweep_chain_long<ROUTE_MAX_COMP><<<ctx->n_chains_long, (CHAIN_PROD + 1) * 32, 0, st>>>(a, ch);
        }
    }
    SPHY_LAUNCH_CHECK(ctx);
    return SPHY_OK;
}

int route_scan_run(sphy_ctx *ctx, int ncomp, const float *const *runoff_mm, float *const *q_state, sphy_param kx,
                   float one_m_kx, cudaStream_t st);   // route_scan.cu
int route_scan_run_adv(sphy_ctx *ctx, float *material_m3day, float *q_state, sphy_param kx, float one_m_kx, const float *qfrac,
                       const float *qout_dense, float *qday, int n_cut, const int32_t *cut_cell_dev, double *cut_work, cudaStream_t st);

extern "C" int sphy_route_step(sphy_ctx *ctx, int ncomp, const float *const *runoff_mm, float *const *q_state,
                               sphy_param kx, float one_minus_kx, int method, void *stream)
{
    if (!ctx || !runoff_mm || !q_state || ncomp < 1 || ncomp > ROUTE_MAX_COMP) return SPHY_E_INVALID;
    if (!ctx->ldd_ready) SPHY_FAIL(ctx, SPHY_E_STATE, "sphy_route_step before sphy_ldd_build");
    for (int c = 0; c < ncomp; ++c)
        if (!runoff_mm[c] || !q_state[c]) SPHY_FAIL(ctx, SPHY_E_INVALID, "sphy_route_step: component %d has a NULL map", c);
    if (ctx->partitioned) SPHY_FAIL(ctx, SPHY_E_STATE, "partitioned context: use sphy_route_scan_begin / sphy_route_scan_finish");
    if (method == SPHY_ROUTE_SCAN) return route_scan_run(ctx, ncomp, runoff_mm, q_state, kx, one_minus_kx, (cudaStream_t)stream);
    if (method != SPHY_ROUTE_LEVELS) SPHY_FAIL(ctx, SPHY_E_INVALID, "sphy_route_step: unknown method %d", method);
    SweepArgs a = {};
    a.ncomp = ncomp;
    a.mode = MODE_ROUT;
    for (int c = 0; c < ncomp; ++c) {
        a.m[c] = runoff_mm[c];
        a.qstate[c] = q_state[c];
    }
    a.kx = kx;
    a.one_m_kx = one_minus_kx;
    return route_levels_run(ctx, a, (cudaStream_t)stream);
}

extern "C" int sphy_accuflux(sphy_ctx *ctx, const float *mat, const float *fraction, float *flux, int method, void *stream)
{
    if (!ctx || !mat || !flux) return SPHY_E_INVALID;
    if (!ctx->ldd_ready) SPHY_FAIL(ctx, SPHY_E_STATE, "sphy_accuflux before sphy_ldd_build");
    if (ctx->partitioned) SPHY_FAIL(ctx, SPHY_E_STATE, "sphy_accuflux is not available on a partitioned context");
    if (method != SPHY_ROUTE_LEVELS) SPHY_FAIL(ctx, SPHY_E_INVALID, "sphy_accuflux: only SPHY_ROUTE_LEVELS is available");
    SweepArgs a = {};
    a.ncomp = 1;
    a.mode = MODE_PLAIN;
    a.m[0] = mat;
    a.flux_out[0] = flux;
    a.frac = fraction;
    return route_levels_run(ctx, a, (cudaStream_t)stream);
}

extern "C" int sphy_accucapacity(sphy_ctx *ctx, const float *mat, const float *capacity, float *flux, float *state, void *stream)
{
    if (!ctx || !mat || !capacity || (!flux && !state)) return SPHY_E_INVALID;
    if (!ctx->ldd_ready) SPHY_FAIL(ctx, SPHY_E_STATE, "sphy_accucapacity before sphy_ldd_build");
    if (ctx->partitioned) SPHY_FAIL(ctx, SPHY_E_STATE, "sphy_accucapacity is not available on a partitioned context");
    SweepArgs a = {};
    a.ncomp = 1;
    a.mode = MODE_CAPACITY;
    a.m[0] = mat;
    a.flux_out[0] = flux;
    a.qstate[0] = state;
    a.frac = capacity;
    return route_levels_run(ctx, a, (cudaStream_t)stream);
}

extern "C" int sphy_upstream(sphy_ctx *ctx, const float *x, float *out, void *stream)
{
    if (!ctx || !x || !out) return SPHY_E_INVALID;
    if (!ctx->ldd_ready) SPHY_FAIL(ctx, SPHY_E_STATE, "sphy_upstream before sphy_ldd_build");
    if (ctx->partitioned) SPHY_FAIL(ctx, SPHY_E_STATE, "sphy_upstream is not available on a partitioned context");
    k_upstream<<<div_up(ctx->n, 256), 256, 0, (cudaStream_t)stream>>>(x, ctx->don_ptr, ctx->don_idx, ctx->n, out);
    SPHY_LAUNCH_CHECK(ctx);
    return SPHY_OK;
}

// ---- lakes / reservoirs ---------------------------------------------------------------------------------
namespace {

// parameter rows of sphy_reslake.par (SPHY_RL_NPAR x n_struct)
enum {
    RL_KR = 0, RL_B, RL_SMAX,                                   // simple reservoir (reservoirs.py:62-70)
    RL_EVOL, RL_PVOL, RL_MAXFL, RL_DEMFL, RL_FLSTART, RL_FLEND,  // advanced reservoir (reservoirs.py:26-58)
    RL_HS_FUNC, RL_HS_EXPA, RL_HS_EXPB, RL_HS_B, RL_HS_A1, RL_HS_A2, RL_HS_A3,   // lake h(S) (lakes.py:98-131)
    RL_QH_FUNC, RL_QH_EXPA, RL_QH_EXPB, RL_QH_B, RL_QH_A1, RL_QH_A2, RL_QH_A3,   // lake Q(h) (lakes.py:133-158)
    RL_SH_FUNC, RL_SH_EXPA, RL_SH_EXPB, RL_SH_B, RL_SH_A1, RL_SH_A2, RL_SH_A3,   // lake S(h), measured levels (lakes.py:37-63)
};
static_assert(RL_SH_A3 + 1 == SPHY_RL_NPAR, "parameter rows of sphy_reslake.par");

__device__ __forceinline__ float dpowf(float x, float y) { return (float)pow((double)x, (double)y); }
__device__ __forceinline__ float dexpf(float x) { return (float)exp((double)x); }

__device__ float lake_fun(float fn, float ea, float eb, float pb, float a1, float a2, float a3, float x)
{
    if (fn == 1.0f) return ea * dexpf(eb * x);
    if (fn == 2.0f) return a1 * x + pb;
    if (fn == 3.0f) return ((a2 * dpowf(x, 2.0f)) + a1 * x) + pb;
    return (((a3 * dpowf(x, 3.0f)) + (a2 * dpowf(x, 2.0f))) + a1 * x) + pb;
}

// per-structure parameter access
struct RLPar {
    const float *par;
    int ns, j;
    __device__ __forceinline__ float operator()(int row) const { return par[(size_t)row * ns + j]; }
};

// lakes.UpdateLakeHStore, first half (lakes.py:37-65): a gauged lake with a defined level takes its storage from S(h);
// then the WHOLE storage map is clamped at 0 (reservoirs included).  Only on days with a level map.
__device__ __forceinline__ float rl_storage_from_level(const RLPar &P, float S, float measured)
{
    if (measured == measured)
        S = lake_fun(P(RL_SH_FUNC), P(RL_SH_EXPA), P(RL_SH_EXPB), P(RL_SH_B), P(RL_SH_A1), P(RL_SH_A2), P(RL_SH_A3), measured);
    return fmaxf(S, 0.0f);
}
// lakes.UpdateLakeHStore, second half (lakes.py:66-131): the level, measured or h(S)
__device__ __forceinline__ float rl_lake_level(const RLPar &P, float S, float measured)
{
    return measured == measured ? measured
                                : lake_fun(P(RL_HS_FUNC), P(RL_HS_EXPA), P(RL_HS_EXPB), P(RL_HS_B), P(RL_HS_A1), P(RL_HS_A2), P(RL_HS_A3), S);
}
// lakes.QLake (lakes.py:136-158), m3/day
__device__ __forceinline__ float rl_qlake(const RLPar &P, float h)
{
    const float qh = lake_fun(P(RL_QH_FUNC), P(RL_QH_EXPA), P(RL_QH_EXPB), P(RL_QH_B), P(RL_QH_A1), P(RL_QH_A2), P(RL_QH_A3), h);
    return (fmaxf(0.0f, qh) * 3600.0f) * 24.0f;
}
// reservoirs.QSimple (reservoirs.py:62-70)
__device__ __forceinline__ float rl_qsimple(const RLPar &P, float S)
{
    return fmaxf(fminf((P(RL_KR) * S) * dpowf(S / P(RL_SMAX), P(RL_B)), S), S - P(RL_SMAX));
}
// reservoirs.QAdv (reservoirs.py:26-58)
__device__ __forceinline__ float rl_qadv(const RLPar &P, float S, int doy)
{
    const float D = (float)doy, fs = P(RL_FLSTART), fe = P(RL_FLEND);
    bool flood;
    if (fs < fe) flood = D >= fs && D <= fe;
    else if (D >= fe) flood = D >= fs;
    else flood = D <= fe ? D <= fs : false;
    const float avail = fmaxf(S - P(RL_PVOL), 0.0f);
    const float den = P(RL_EVOL) - P(RL_PVOL);
    return fmaxf(flood ? (P(RL_MAXFL) * avail) / den : (P(RL_DEMFL) * avail) / den, S - P(RL_EVOL));
}

__global__ void k_rl_pre(sphy_reslake rl, const float *__restrict__ totr, float vol_factor, int doy, float *__restrict__ qout_dense)
{
    int j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= rl.n_struct) return;
    const int32_t c = rl.cell[j];
    const RLPar P = {rl.par, rl.n_struct, j};
    float S = rl.StorRES[j] + vol_factor * totr[c];             // advanced_routing.py:136-138
    float q = 0.0f;
    const int kind = rl.kind[j];
    float measured = nanf("");
    if (rl.lake_level) {                                        // a measured-level map exists for this day
        if (kind == 3 && rl.update_level[j] != 0) measured = rl.lake_level[j];
        S = rl_storage_from_level(P, S, measured);
    }
    if (kind == 1) q = rl_qsimple(P, S);
    else if (kind == 2) q = rl_qadv(P, S, doy);
    else if (kind == 3) q = rl_qlake(P, rl_lake_level(P, S, measured));
    rl.StorRES[j] = S;
    rl.Qout[j] = q;
    qout_dense[c] = q;
}

// the same functions one at a time, over the structure table (what sphy_b200/modules/lakes.py and reservoirs.py call)
__global__ void k_rl_op(sphy_reslake rl, int op, int doy, const float *__restrict__ in, float *__restrict__ out)
{
    int j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= rl.n_struct) return;
    const RLPar P = {rl.par, rl.n_struct, j};
    const int kind = rl.kind[j];
    float S = rl.StorRES[j];
    float r = nanf("");
    switch (op) {
        case SPHY_RL_OP_UPDATE_LAKE: {                          // lakes.UpdateLakeHStore -> LakeLevel (StorRES updated in place)
            float measured = nanf("");
            if (rl.lake_level) {
                if (kind == 3 && rl.update_level[j] != 0) measured = rl.lake_level[j];
                const float S1 = rl_storage_from_level(P, S, measured);
                if (kind == 3) rl.StorRES[j] = S = S1;          // lakes.py:132: only lake cells keep the new storage
            }
            if (kind == 3) r = rl_lake_level(P, S, measured);
            break;
        }
        case SPHY_RL_OP_QLAKE: r = kind == 3 ? rl_qlake(P, in[j]) : 0.0f; break;            // cover(Qout, 0), lakes.py:157
        case SPHY_RL_OP_QSIMPLE: if (kind == 1) r = rl_qsimple(P, S); break;
        case SPHY_RL_OP_QADV: if (kind == 2) r = rl_qadv(P, S, doy); break;
        case SPHY_RL_OP_QRES: r = kind == 1 ? rl_qsimple(P, S) : (kind == 2 ? rl_qadv(P, S, doy) : 0.0f); break;
        default: break;
    }
    out[j] = r;
}

// The same on a partitioned basin (sphy_route_reslake_finish): the table is replicated on every rank; TotR at the structure
// cells and the discharge of the previous step at their donors come from the owners' exchange blocks.  First the storage
// takes `- Qout + Qin` of the previous step (advanced_routing.py:35-36: upstream(ldd, Q), REAL8 sum in donor order like
// k_rl_post), then the step proceeds like k_rl_pre.
__global__ void k_rl_part_update(sphy_reslake rl, RlPartUpdate u)
{
    const int j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= rl.n_struct) return;
    const double *G = u.gathered;
    if (!G) {                                                   // k_trunk_drift has already advanced the step counter
        const long long cur = *reinterpret_cast<const volatile long long *>(u.p2p_local + u.p2p_world + 1);
        G = u.p2p_recv + (size_t)((cur - 1) & 1) * u.p2p_world * u.p2p_pitch;
    }
    const int32_t *xr = u.x_rank + (int64_t)j * SPHY_RL_EXTRA;
    const int64_t xoff = (int64_t)u.stride - u.n_extra + (int64_t)j * SPHY_RL_EXTRA;
    const RLPar P = {rl.par, rl.n_struct, j};
    float S = rl.StorRES[j];
    if (u.apply_pending) {
        double s = 0.0;
        for (int q = 0; q < SPHY_RL_EXTRA - 1; ++q)
            if (xr[q] >= 0) s += G[(int64_t)xr[q] * u.rank_pitch + xoff + q];
        const float qin = (float)s;
        rl.Qin[j] = qin;
        S = (S - rl.Qout[j]) + qin;                             // advanced_routing.py:36
    }
    const float totr = (float)G[(int64_t)xr[SPHY_RL_EXTRA - 1] * u.rank_pitch + xoff + SPHY_RL_EXTRA - 1];
    S = S + u.vol_factor * totr;                                // advanced_routing.py:136-138
    float q = 0.0f;
    const int kind = rl.kind[j];
    float measured = nanf("");
    if (rl.lake_level) {
        if (kind == 3 && rl.update_level[j] != 0) measured = rl.lake_level[j];
        S = rl_storage_from_level(P, S, measured);
    }
    if (kind == 1) q = rl_qsimple(P, S);
    else if (kind == 2) q = rl_qadv(P, S, u.doy);
    else if (kind == 3) q = rl_qlake(P, rl_lake_level(P, S, measured));
    rl.StorRES[j] = S;
    rl.Qout[j] = q;
}

__global__ void k_rl_material(const float *__restrict__ totr, const float *__restrict__ qfrac, float vol_factor, int64_t n,
                              float *__restrict__ m)
{
    int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (i >= n) return;
    m[i] = qfrac[i] == 0.0f ? 0.0f : vol_factor * totr[i];     // advanced_routing.py:159-161 (second term)
}

// add upstream(ldd, Qout) to the material of each receiver of a structure (once per receiver)
__global__ void k_rl_upstream_add(sphy_reslake rl, const int32_t *__restrict__ down, const int32_t *__restrict__ don_ptr,
                                  const int32_t *__restrict__ don_idx, const float *__restrict__ qfrac,
                                  const float *__restrict__ qout_dense, float *__restrict__ m)
{
    int j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= rl.n_struct) return;
    const int32_t c = rl.cell[j];
    const int32_t r = down[c];
    if (r < 0) return;
    double s = 0.0;
    bool first = true;
    for (int32_t q = don_ptr[r]; q < don_ptr[r + 1]; ++q) {
        const int32_t d = don_idx[q];
        if (qfrac[d] == 0.0f) {
            if (d < c) first = false;                           // a lower-indexed structure donor owns the update
            s += (double)qout_dense[d];
        }
    }
    if (first) m[r] = (float)s + m[r];
}

// qday is indexed by routing position (level sweep: pos != NULL) or by cell (scan path: pos == NULL)
__global__ void k_rl_post(sphy_reslake rl, const int32_t *__restrict__ don_ptr, const int32_t *__restrict__ don_idx,
                          const int32_t *__restrict__ pos, const float *__restrict__ qday)
{
    int j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= rl.n_struct) return;
    const int32_t c = rl.cell[j];
    double s = 0.0;
    for (int32_t q = don_ptr[c]; q < don_ptr[c + 1]; ++q) s += (double)qday[pos ? pos[don_idx[q]] : don_idx[q]];
    const float qin = (float)s;                                 // upstream(ldd, Q) at the structure cell
    rl.Qin[j] = qin;
    rl.StorRES[j] = (rl.StorRES[j] - rl.Qout[j]) + qin;         // advanced_routing.py:36
}

}  // namespace

int reslake_part_update(sphy_ctx *ctx, const sphy_reslake *rl, const RlPartUpdate &u, cudaStream_t st)
{
    k_rl_part_update<<<div_up(rl->n_struct, 128), 128, 0, st>>>(*rl, u);
    SPHY_LAUNCH_CHECK(ctx);
    return SPHY_OK;
}

extern "C" int sphy_route_reslake_step(sphy_ctx *ctx, const sphy_reslake *rl, const float *totr, float *q_state, sphy_param kx,
                                       float one_minus_kx, int doy, int method, void *stream)
{
    if (!ctx || !rl || !totr || !q_state) return SPHY_E_INVALID;
    if (!ctx->ldd_ready) SPHY_FAIL(ctx, SPHY_E_STATE, "sphy_route_reslake_step before sphy_ldd_build");
    if (ctx->partitioned) SPHY_FAIL(ctx, SPHY_E_STATE, "sphy_route_reslake_step is not available on a partitioned context");
    if (rl->n_struct < 0 || !rl->qfrac || (rl->n_struct > 0 && (!rl->cell || !rl->kind || !rl->par || !rl->StorRES || !rl->Qout || !rl->Qin)))
        SPHY_FAIL(ctx, SPHY_E_INVALID, "sphy_route_reslake_step: incomplete structure table");
    cudaStream_t st = (cudaStream_t)stream;
    const float vol_factor = 0.001f * ctx->cellarea;            // 0.001 * cellarea() as REAL4 (advanced_routing.py:137)
    float *qout_dense = ctx->tmp_n[1], *m = ctx->tmp_n[0];
    if (!ctx->qout_zeroed) {                                    // qout_dense is only ever written at structure cells
        SPHY_CUDA(ctx, cudaMemsetAsync(qout_dense, 0, sizeof(float) * ctx->n, st));
        ctx->qout_zeroed = true;
    }
    const int ns = rl->n_struct;
    if (ns > 0) {
        k_rl_pre<<<div_up(ns, 128), 128, 0, st>>>(*rl, totr, vol_factor, doy, qout_dense);
    }
    k_rl_material<<<div_up(ctx->n, 256), 256, 0, st>>>(totr, rl->qfrac, vol_factor, ctx->n, m);
    if (ns > 0) {
        k_rl_upstream_add<<<div_up(ns, 128), 128, 0, st>>>(*rl, ctx->down, ctx->don_ptr, ctx->don_idx, rl->qfrac, qout_dense, m);
    }
    SPHY_LAUNCH_CHECK(ctx);
    if (method == SPHY_ROUTE_SCAN) {
        // structures are listed by ascending cell index == ascending DFS rank, which the cut corrections rely on
        if (ctx->order_mode != SPHY_ORDER_DFS) SPHY_FAIL(ctx, SPHY_E_STATE, "SPHY_ROUTE_SCAN needs SPHY_ORDER_DFS");
        if (ctx->cut_work_cap < ns) {
            int rc0 = dev_alloc(ctx, &ctx->cut_work, (size_t)(ns > 0 ? ns : 1));
            if (rc0 != SPHY_OK) return rc0;
            ctx->cut_work_cap = ns;
        }
        int rc1 = route_scan_run_adv(ctx, m, q_state, kx, one_minus_kx, rl->qfrac, qout_dense, ctx->qsorted, ns, rl->cell, ctx->cut_work, st);
        if (rc1 != SPHY_OK) return rc1;
        if (ns > 0) {
            k_rl_post<<<div_up(ns, 128), 128, 0, st>>>(*rl, ctx->don_ptr, ctx->don_idx, nullptr, ctx->qsorted);
            SPHY_LAUNCH_CHECK(ctx);
        }
        return SPHY_OK;
    }
    if (method != SPHY_ROUTE_LEVELS) SPHY_FAIL(ctx, SPHY_E_INVALID, "sphy_route_reslake_step: unknown method %d", method);
    SweepArgs a = {};
    a.ncomp = 1;
    a.mode = MODE_ADV;
    a.m[0] = m;
    a.qstate[0] = q_state;
    a.frac = rl->qfrac;
    a.qout_dense = qout_dense;
    a.qday = ctx->qsorted;
    a.kx = kx;
    a.one_m_kx = one_minus_kx;
    int rc = route_levels_run(ctx, a, st);
    if (rc != SPHY_OK) return rc;
    if (ns > 0) {
        k_rl_post<<<div_up(ns, 128), 128, 0, st>>>(*rl, ctx->don_ptr, ctx->don_idx, ctx->pos, ctx->qsorted);
        SPHY_LAUNCH_CHECK(ctx);
    }
    return SPHY_OK;
}


// ---- open-water evaporation / precipitation on the structures (advanced_routing.py:167-196) -----------------------
namespace {
__global__ void k_rl_openwater(sphy_reslake rl, sphy_openwater ow, float cellarea)
{
    const int j = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;          // one warp per structure
    const int lane = threadIdx.x & 31;
    if (j >= rl.n_struct) return;
    double se = 0.0, sp = 0.0;
    for (int32_t q = ow.ow_ptr[j] + lane; q < ow.ow_ptr[j + 1]; q += 32) {
        const int32_t c = ow.ow_cell[q];
        const float fr = ow.open_water_frac[c];
        se += (double)((ow.et_open_water[c] * cellarea) * fr);             // areatotal: REAL8 sum of REAL4 terms
        sp += (double)(((ow.prec[c] / ow.pcorr) * cellarea) * fr);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        se += __shfl_xor_sync(0xffffffffu, se, o);
        sp += __shfl_xor_sync(0xffffffffu, sp, o);
    }
    if (lane == 0) {
        const float S = rl.StorRES[j];
        const float eta = S > 0.0f ? fminf((float)se * 0.001f, S) : 0.0f;
        const float pr = S > 0.0f ? (float)sp * 0.001f : 0.0f;
        rl.StorRES[j] = (S - eta) + pr;
    }
}
}  // namespace

extern "C" int sphy_reslake_openwater(sphy_ctx *ctx, const sphy_reslake *rl, const sphy_openwater *ow, void *stream)
{
    if (!ctx || !rl || !ow) return SPHY_E_INVALID;
    if (rl->n_struct <= 0) return SPHY_OK;
    if (!ow->ow_ptr || !ow->ow_cell || !ow->open_water_frac || !ow->et_open_water || !ow->prec || !rl->StorRES)
        SPHY_FAIL(ctx, SPHY_E_INVALID, "sphy_reslake_openwater: incomplete arguments");
    k_rl_openwater<<<div_up((int64_t)rl->n_struct * 32, 128), 128, 0, (cudaStream_t)stream>>>(*rl, *ow, ctx->cellarea);
    SPHY_LAUNCH_CHECK(ctx);
    return SPHY_OK;
}

extern "C" int sphy_reslake_op(sphy_ctx *ctx, const sphy_reslake *rl, int op, int day_of_year, const float *in, float *out, void *stream)
{
    if (!ctx || !rl || !out || op < 0 || op >= SPHY_RL_OP_COUNT) return SPHY_E_INVALID;
    if (rl->n_struct <= 0) return SPHY_OK;
    if (!rl->cell || !rl->kind || !rl->par || !rl->StorRES || (op == SPHY_RL_OP_QLAKE && !in))
        SPHY_FAIL(ctx, SPHY_E_INVALID, "sphy_reslake_op: incomplete structure table or missing input");
    k_rl_op<<<div_up(rl->n_struct, 128), 128, 0, (cudaStream_t)stream>>>(*rl, op, day_of_year, in, out);
    SPHY_LAUNCH_CHECK(ctx);
    return SPHY_OK;
}

