// route_scan.cu -- K4b: flow accumulation as subtree range sums over the DFS cell order (no level latency chain).
//
// accuflux is linear and the drainage network is a forest, so with the cells stored in a depth-first pre-order
// (SPHY_ORDER_DFS, built in ldd_build.cu) the upstream area of cell j is the contiguous index range
// [j, j + sub_size[j]) and its accumulated flux is a difference of prefix sums of the per-cell material.
// One step is a single streaming pass plus two tiny fix-up kernels instead of a walk over ~10^4 dependent levels:
//   main   per tile of 1024 cells (one CTA): stream runoff, convert to m3/s (routing.py:26), float64 inclusive
//          scan inside the tile (cub::BlockScan) kept in shared memory; every cell whose range ends inside the
//          tile (~94 % of them) gets its flux from two shared-memory reads and is blended with Qold and written
//          (routing.py:28) right away -- 16 B of HBM traffic per cell and component;
//   totals exclusive scan of the per-tile totals (one CTA per component);
//   cross  the few cells whose range crosses a tile boundary (a static list built once): tile suffix + whole
//          tiles in between + prefix inside the end tile, from values the main pass parked for them.
// Small subtrees never see the global magnitude, so relative accuracy holds for tiny fluxes too.  The float64
// range sums have no per-cell REAL4 rounding along the flow path, hence this path agrees with the level sweep /
// the oracle to ~1e-6 relative (tested at 1e-5), not bit-for-bit; the level sweep stays the bit-exact path.
#include <cstdlib>
#include <cub/cub.cuh>

#include "common.cuh"

namespace {

constexpr int SCAN_THREADS = 256;
constexpr int SCAN_WARPS = SCAN_THREADS / 32;
constexpr int SCAN_TILE = 1024;                        // cells per tile = one warp's unit of work
constexpr int SCAN_CHUNK = 128;                        // 32 lanes x 4 cells
constexpr int SCAN_CHUNKS = SCAN_TILE / SCAN_CHUNK;
constexpr int SCAN_BATCH = 2;                          // chunks whose phase-2 loads are in flight together
constexpr int SCAN_MIN_CTAS = 3;                       // 3 x 64 KB of shared memory per SM
constexpr int SCAN_SMEM = SCAN_WARPS * SCAN_TILE * (int)sizeof(double);

struct ScanArgs {
    int ncomp;
    const float *m[ROUTE_MAX_COMP];       // runoff mm/day
    float *qstate[ROUTE_MAX_COMP];        // Qold in / Q out
    const int32_t *sub_size;
    // static crossing lists (see build_cross_lists)
    const int32_t *cross_r, *cross_ptr;   // crossing cells sorted by index; per-tile ranges
    const int32_t *cross_e;               // last cell of each crossing cell's upstream range
    const int32_t *end_e, *end_k, *end_ptr;   // their range ends sorted by index, the owning list slot, per-tile ranges
    int32_t n_cross;
    double *park_before;                  // [ncomp][n_cross] tile-local exclusive prefix at the crossing cell
    double *park_end;                     // [ncomp][n_cross] tile-local inclusive prefix at its range end
    double *tile_tot;                     // [ncomp][ntiles+1] tile totals (+ a trailing 0 so the scan yields the grand total)
    double *tile_excl;                    // [ncomp][ntiles+1] exclusive prefix of the totals (within 1024-entry blocks when tile_base != NULL)
    const double *tile_base;              // [ncomp][nblocks] exclusive prefix of the block totals, or NULL
    int32_t tile_nb;
    // published prefixes (trunk cells; on a partitioned basin they are what the ranks exchange)
    int32_t n_pub;
    const int32_t *pub_e;
    double *send;                         // [ncomp][stride]: range total, then range-local inclusive prefix at pub_e[j]
    int32_t stride;
    // lakes / reservoirs on the trunk list: per-structure values that ride at the end of the block (component 0)
    const double *extra;
    int32_t n_extra;
    // lakes / reservoirs (advanced_routing.ROUT): material is m3/day, cut cells take the structure outflow
    int adv;
    float volf;                           // material = volf * m (REAL4): 1 when m is already m3/day, 0.001 * cellarea when m is TotR [mm]
    const float *frac;                    // dense QFRAC (0 at lake / reservoir cells)
    const float *qout_dense;              // dense structure outflow, m3/day
    float *qday;                          // dense blended Q, m3/day (what upstream(ldd, Q) reads at the structures)
    const int32_t *cut_cell, *cut_ptr;    // cut cells (ascending) and their per-tile ranges
    const double *cut_corr;               // float64 material placed at each cut cell (see route_scan_run_adv)
    sphy_param kxc[ROUTE_MAX_COMP];       // recession coefficient per routed map (ensemble members route together, each with its own kx)
    float omkc[ROUTE_MAX_COMP];           // REAL4(1 - kx) when kx is a scalar
    double to_m3s;                        // mm/day -> m3/s: 0.001 * cellarea / 86400 (routing.py:26)
    int64_t n;
    int32_t ntiles;
};

// shared-memory slot of tile-local cell i: within each 128-cell chunk the prefix is stored transposed (cell 4*lane+j of
// the chunk at j*32 + lane), so that the per-lane accesses of both phases are bank-conflict free
__device__ __forceinline__ int ls_slot(int i) { return (i & ~127) | ((i & 3) << 5) | ((i >> 2) & 31); }

// Range sums are differences of float64 prefixes.  Parallel prefixes are only consistent to a few ulp of the prefix, so
// a range without any runoff would come out as +-1e-16 x prefix instead of 0 (and a negative discharge breaks x**1.5
// in the erosion module).  A difference inside that rounding noise IS zero as far as this method can tell.
__device__ __forceinline__ double range_diff(double hi, double lo)
{
    const double d = hi - lo;
    return fabs(d) <= 3.6e-15 * fabs(hi) ? 0.0 : d;            // 32 ulp of the minuend
}

// exclusive prefix of the tile totals at tile t of component c (two-level when the small-CTA scan produced it)
__device__ __forceinline__ double tile_prefix(const ScanArgs &a, int c, int32_t t)
{
    const double x = a.tile_excl[(int64_t)c * (a.ntiles + 1) + t];
    return a.tile_base ? x + a.tile_base[(int64_t)c * a.tile_nb + (t >> 10)] : x;
}

// recession blend of one cell: routing.py:28, or advanced_routing.py:31-37 on a lake/reservoir network
__device__ __forceinline__ float blend(const ScanArgs &a, int64_t cell, float f, float qold, float kx, float omk)
{
    if (!a.adv) return omk * f + kx * qold;
    const float qd = __ldg(a.frac + cell) == 0.0f ? __ldg(a.qout_dense + cell) : omk * f + ((qold * 24.0f) * 3600.0f) * kx;
    a.qday[cell] = qd;
    return qd / 86400.0f;
}

// main pass, warp-autonomous: every WARP owns one 1024-cell tile at a time and keeps the tile's float64 inclusive
// prefix in its own 8 KB slice of shared memory, so there is no CTA barrier and no cross-warp traffic at all.
//   phase 1  stream the runoff of the tile in 8 chunks of 128 cells (4 per lane, 128-bit loads all issued up front),
//            convert to m3/s (routing.py:26; one float64 multiply), lane-serial + shuffle scan with a running carry;
//   park     what crossing cells of / ending in this tile will need; the tile total;
//   phase 2  stream sub_size and Qold, take the range sum of every cell whose range ends inside the tile from two
//            shared-memory reads, blend with Qold (routing.py:28) and write Q.
// Every array is read exactly once (16 B per cell and component).  The kernel is issue-bound, so the variants
// (lake network, per-cell kx) are compiled separately and all in-tile indices are 32-bit.
template <bool ADV, bool KXMAP>
__global__ void __launch_bounds__(SCAN_THREADS, SCAN_MIN_CTAS) k_scan_main(const __grid_constant__ ScanArgs a)
{
    extern __shared__ __align__(16) double scan_smem[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    double *Ls = scan_smem + warp * SCAN_TILE;
    const int32_t nw = gridDim.x * SCAN_WARPS;
    int32_t t = blockIdx.x * SCAN_WARPS + warp;
    if (t >= a.ntiles) return;

    // the material of tile tt, component cc: 8 x 128-bit loads per lane, all in flight together
    auto load_x = [&](int32_t tt, int cc, float4 (&x)[SCAN_CHUNKS]) {
        const int64_t tb = (int64_t)tt * SCAN_TILE;
        const float *m = a.m[cc] + tb;
        const int ln = (int)min((int64_t)SCAN_TILE, a.n - tb);
        if (ln == SCAN_TILE) {
#pragma unroll
            for (int i = 0; i < SCAN_CHUNKS; ++i) x[i] = __ldg(reinterpret_cast<const float4 *>(m) + i * 32 + lane);
        } else {
#pragma unroll
            for (int i = 0; i < SCAN_CHUNKS; ++i) {
                const int o = i * SCAN_CHUNK + lane * 4;
                x[i].x = o < ln ? __ldg(m + o) : 0.0f;
                x[i].y = o + 1 < ln ? __ldg(m + o + 1) : 0.0f;
                x[i].z = o + 2 < ln ? __ldg(m + o + 2) : 0.0f;
                x[i].w = o + 3 < ln ? __ldg(m + o + 3) : 0.0f;
            }
        }
    };
    float4 x[SCAN_CHUNKS];
    load_x(t, 0, x);
    for (; t < a.ntiles; t += nw) {
        const int64_t tbase = (int64_t)t * SCAN_TILE;
        const int len = (int)min((int64_t)SCAN_TILE, a.n - tbase);   // cells in this tile
        const bool full = len == SCAN_TILE;
        const int32_t tb32 = (int32_t)tbase;                   // n < 2^31 (sub_size is int32)
        const int32_t cb = __ldg(a.cross_ptr + t), ce = __ldg(a.cross_ptr + t + 1);
        const int32_t eb = __ldg(a.end_ptr + t), ee = __ldg(a.end_ptr + t + 1);
        int32_t kb = 0, ke = 0;
        if (ADV) { kb = __ldg(a.cut_ptr + t); ke = __ldg(a.cut_ptr + t + 1); }
        const int32_t *szp = a.sub_size + tbase;
        for (int c = 0; c < a.ncomp; ++c) {
            float *qs = a.qstate[c] + tbase;
            int4 sz[SCAN_BATCH];
            float4 q[SCAN_BATCH], kxv[SCAN_BATCH];
            // sub_size / Qold / kx of SCAN_BATCH chunks starting at chunk h
            auto load_batch = [&](int h) {
#pragma unroll
                for (int b = 0; b < SCAN_BATCH; ++b) {
                    const int o = (h + b) * SCAN_CHUNK + lane * 4;
                    if (full) {
                        sz[b] = __ldg(reinterpret_cast<const int4 *>(szp + o));
                        q[b] = *reinterpret_cast<const float4 *>(qs + o);
                        if (KXMAP) kxv[b] = __ldg(reinterpret_cast<const float4 *>(a.kxc[c].map + tbase + o));
                    } else {
                        sz[b].x = o < len ? __ldg(szp + o) : 1;         q[b].x = o < len ? qs[o] : 0.0f;
                        sz[b].y = o + 1 < len ? __ldg(szp + o + 1) : 1; q[b].y = o + 1 < len ? qs[o + 1] : 0.0f;
                        sz[b].z = o + 2 < len ? __ldg(szp + o + 2) : 1; q[b].z = o + 2 < len ? qs[o + 2] : 0.0f;
                        sz[b].w = o + 3 < len ? __ldg(szp + o + 3) : 1; q[b].w = o + 3 < len ? qs[o + 3] : 0.0f;
                        if (KXMAP) {
                            const float *kp = a.kxc[c].map + tbase;
                            kxv[b].x = o < len ? __ldg(kp + o) : 0.0f;         kxv[b].y = o + 1 < len ? __ldg(kp + o + 1) : 0.0f;
                            kxv[b].z = o + 2 < len ? __ldg(kp + o + 2) : 0.0f; kxv[b].w = o + 3 < len ? __ldg(kp + o + 3) : 0.0f;
                        }
                    }
                }
            };
            load_batch(0);                                     // in flight while phase 1 computes
            // ---- phase 1: float64 inclusive prefix of the tile's material -> Ls
            double carry = 0.0;
#pragma unroll
            for (int i = 0; i < SCAN_CHUNKS; ++i) {
                const int o = i * SCAN_CHUNK + lane * 4;
                double v0, v1, v2, v3;
                if (ADV) {
                    v0 = (double)(a.volf * x[i].x); v1 = (double)(a.volf * x[i].y);     // advanced_routing.py:137,159 (REAL4 product)
                    v2 = (double)(a.volf * x[i].z); v3 = (double)(a.volf * x[i].w);
                    for (int32_t k = kb; k < ke; ++k) {        // float64 corrections at the cut cells of this tile
                        const int off = (__ldg(a.cut_cell + k) - tb32) - o;
                        if (off >= 0 && off < 4) {
                            const double corr = a.cut_corr[k];
                            if (off == 0) v0 += corr; else if (off == 1) v1 += corr; else if (off == 2) v2 += corr; else v3 += corr;
                        }
                    }
                } else {
                    v0 = (double)x[i].x * a.to_m3s; v1 = (double)x[i].y * a.to_m3s;
                    v2 = (double)x[i].z * a.to_m3s; v3 = (double)x[i].w * a.to_m3s;
                }
                v1 += v0; v2 += v1; v3 += v2;
                double s = v3;                                 // inclusive scan of the lane totals
#pragma unroll
                for (int d = 1; d < 32; d <<= 1) {
                    const double y = __shfl_up_sync(0xffffffffu, s, d);
                    if (lane >= d) s += y;
                }
                double excl = __shfl_up_sync(0xffffffffu, s, 1);
                excl = (lane == 0 ? 0.0 : excl) + carry;
                v0 += excl; v1 += excl; v2 += excl; v3 += excl;
                double *row = Ls + i * SCAN_CHUNK + lane;      // transposed within the chunk, see ls_slot()
                row[0] = v0; row[32] = v1; row[64] = v2; row[96] = v3;
                carry = __shfl_sync(0xffffffffu, v3, 31);
            }
            // the next material this warp will scan: in flight during parking and phase 2
            if (c + 1 < a.ncomp) load_x(t, c + 1, x);
            else if (t + nw < a.ntiles) load_x(t + nw, 0, x);
            __syncwarp();
            // ---- park what the crossing cells of / ending in this tile will need
            if (ce > cb || ee > eb) {
                double *pb = a.park_before + (int64_t)c * a.n_cross, *pe = a.park_end + (int64_t)c * (a.n_cross + a.n_pub);
                for (int32_t k = cb + lane; k < ce; k += 32) {
                    const int off = __ldg(a.cross_r + k) - tb32;
                    pb[k] = off == 0 ? 0.0 : Ls[ls_slot(off - 1)];
                }
                for (int32_t j = eb + lane; j < ee; j += 32) pe[__ldg(a.end_k + j)] = Ls[ls_slot(__ldg(a.end_e + j) - tb32)];
            }
            if (lane == 0) a.tile_tot[(int64_t)c * (a.ntiles + 1) + t] = carry;
            // ---- phase 2: range sums of the cells whose range ends inside the tile, blended with Qold
#pragma unroll
            for (int h = 0; h < SCAN_CHUNKS; h += SCAN_BATCH) {
                if (h > 0) load_batch(h);
#pragma unroll
                for (int b = 0; b < SCAN_BATCH; ++b) {
                    const int o = (h + b) * SCAN_CHUNK + lane * 4;
                    const double *row = Ls + (h + b) * SCAN_CHUNK + lane;
                    const double p0 = row[0], p1 = row[32], p2 = row[64], p3 = row[96];   // own inclusive prefixes
                    const double p_1 = o == 0 ? 0.0 : (lane == 0 ? row[-1] : row[95]);     // the cell before the lane's first
                    auto cell = [&](int j, int32_t szj, float &qj, float kxj, double before, double own) {
                        const int el = o + j + szj - 1;                // last cell of the upstream range, tile-local
                        const bool inside = el < len;                  // else: crossing cell, left to k_scan_cross
                        const float kx = KXMAP ? kxj : a.kxc[c].value;
                        const float omk = KXMAP ? 1.0f - kx : a.omkc[c];
                        double pe = own;                               // headwater cell: the range is the cell itself
                        if (inside && szj > 1) pe = Ls[ls_slot(el)];   // predicated load: fewer lanes, fewer bank conflicts
                        if (ADV) {
                            // mask value 2 (sphy_reslake_plan.qmask): a listed cell, finished by k_trunk_apply_adv
                            if (inside && __ldg(a.frac + tbase + o + j) != 2.0f)
                                qj = blend(a, tbase + o + j, (float)range_diff(pe, before), qj, kx, omk);
                        } else {
                            const float qn = omk * (float)range_diff(pe, before) + kx * qj;   // routing.py:28
                            qj = inside ? qn : qj;
                        }
                    };
                    cell(0, sz[b].x, q[b].x, kxv[b].x, p_1, p0);
                    cell(1, sz[b].y, q[b].y, kxv[b].y, p0, p1);
                    cell(2, sz[b].z, q[b].z, kxv[b].z, p1, p2);
                    cell(3, sz[b].w, q[b].w, kxv[b].w, p2, p3);
                    if (full) {
                        *reinterpret_cast<float4 *>(qs + o) = q[b];
                    } else {
                        if (o < len) qs[o] = q[b].x;
                        if (o + 1 < len) qs[o + 1] = q[b].y;
                        if (o + 2 < len) qs[o + 2] = q[b].z;
                        if (o + 3 < len) qs[o + 3] = q[b].w;
                    }
                }
            }
            __syncwarp();                                      // Ls is overwritten by the next component / tile
        }
    }
}

// cells whose upstream range crosses a tile boundary
constexpr int CROSS_ILP = 4;       // crossing cells per thread: four independent gather chains in flight
// The kernels behind the main pass run 128-thread CTAs with a modest register budget, so that one CTA of them fits on an SM
// next to the resident CTAs of the NEXT step's water-balance kernel: with SphyModel(overlap_routing=True) this latency-bound
// tail (tile totals, crossing cells, publish / peer push, flag wait, trunk kernels) runs on its own stream underneath it.
constexpr int TAIL_THREADS = 128;

__device__ __forceinline__ void scan_cross_block(const ScanArgs &a, int32_t block)
{
    const int32_t k0 = block * (TAIL_THREADS * CROSS_ILP) + threadIdx.x;
    int32_t r[CROSS_ILP], e[CROSS_ILP];
    bool ok[CROSS_ILP];
#pragma unroll
    for (int j = 0; j < CROSS_ILP; ++j) {
        const int32_t k = k0 + j * TAIL_THREADS;
        ok[j] = k < a.n_cross;
        r[j] = ok[j] ? __ldg(a.cross_r + k) : 0;
        e[j] = ok[j] ? __ldg(a.cross_e + k) : 0;
    }
    for (int c = 0; c < a.ncomp; ++c) {
        const double *T = a.tile_tot + (int64_t)c * (a.ntiles + 1);
        const double *PB = a.park_before + (int64_t)c * a.n_cross;
        const double *PE = a.park_end + (int64_t)c * (a.n_cross + a.n_pub);
        float *qs = a.qstate[c];
        double tt[CROSS_ILP], pb[CROSS_ILP], pe[CROSS_ILP];
        float qold[CROSS_ILP], kx[CROSS_ILP];
#pragma unroll
        for (int j = 0; j < CROSS_ILP; ++j) {                  // all gathers first
            const int32_t k = k0 + j * TAIL_THREADS;
            tt[j] = ok[j] ? T[r[j] / SCAN_TILE] : 0.0;
            pb[j] = ok[j] ? PB[k] : 0.0;
            pe[j] = ok[j] ? PE[k] : 0.0;
            qold[j] = ok[j] ? qs[r[j]] : 0.0f;
            kx[j] = (ok[j] && a.kxc[c].map) ? __ldg(a.kxc[c].map + r[j]) : a.kxc[c].value;
        }
#pragma unroll
        for (int j = 0; j < CROSS_ILP; ++j) {
            if (!ok[j]) continue;
            const int32_t ta = r[j] / SCAN_TILE, te = e[j] / SCAN_TILE;
            // suffix of the first tile + whole tiles in between (none for most crossing cells) + prefix of the last tile
            const double s = range_diff(tt[j], pb[j]) + (te > ta + 1 ? range_diff(tile_prefix(a, c, te), tile_prefix(a, c, ta + 1)) : 0.0) + pe[j];
            const float omk = a.kxc[c].map ? 1.0f - kx[j] : a.omkc[c];
            qs[r[j]] = blend(a, r[j], (float)s, qold[j], kx[j], omk);
        }
    }
}

// Confluence exchange over peer memory (NVLink P2P, CUDA IPC; set up by sphy_p2p_setup_local / sphy_p2p_connect).  Every
// rank owns a receive buffer [2 parities][world][pitch] and, in `local`, one flag per peer followed by a time-out word, the
// exchange step counter and a block counter.  The publishing blocks store this rank's values straight into every peer's
// buffer; the last of them releases flag[rank] = step on every peer with system scope.  The trunk kernel acquires the
// flags of all peers before it reads the blocks (bounded: a dead peer yields NaN discharges and an error, not a hang) and
// advances the step counter when it is done.  Double buffering by step parity is enough: a rank can only be one exchange
// ahead of a peer, because its next push comes after a wait on that peer's current flag.  The step counter lives in device
// memory, so the whole step can be replayed from a CUDA graph.
struct P2PArgs {
    int on;                               // 0: no peer exchange (one GPU, or the NCCL all-gather done by the host)
    int rank, world;
    double *const *peer_recv;
    long long *const *peer_flags;
    long long *local;                     // this rank's flags[world], time-out word, step counter, block counter
    double *recv;                         // this rank's receive buffer
    int64_t pitch;
    long long max_naps;
};

// One launch finishes the ordinary tile-crossing cells (blocks [0, n_cross_blocks)) and publishes the range total and the
// range-local inclusive prefix at every published position (the remaining blocks) -- what the trunk cells are finished
// from; on a partitioned basin that block is the "confluence inflow" message of the rank.
__global__ void __launch_bounds__(TAIL_THREADS, 5) k_scan_cross_pub(const __grid_constant__ ScanArgs a, const __grid_constant__ P2PArgs p,
                                                        int n_cross_blocks)
{
    if ((int)blockIdx.x < n_cross_blocks) {
        scan_cross_block(a, blockIdx.x);
        return;
    }
    const int32_t i = (blockIdx.x - n_cross_blocks) * TAIL_THREADS + threadIdx.x;      // 0: the range total, 1 + j: position j
    long long step = 0;
    size_t slot = 0;
    if (p.on) {
        step = *reinterpret_cast<const volatile long long *>(p.local + p.world + 1) + 1;
        slot = ((size_t)((step - 1) & 1) * p.world + p.rank) * p.pitch;
    }
    if (i <= a.n_pub) {
        for (int c = 0; c < a.ncomp; ++c) {
            const double v = i == 0 ? tile_prefix(a, c, a.ntiles)
                                    : tile_prefix(a, c, a.pub_e[i - 1] / SCAN_TILE) + a.park_end[(int64_t)c * (a.n_cross + a.n_pub) + a.n_cross + i - 1];
            const int64_t o = (int64_t)c * a.stride + i;
            if (a.send) a.send[o] = v;
            if (p.on)
                for (int q = 0; q < p.world; ++q) p.peer_recv[q][slot + o] = v;
        }
    } else if (i - a.n_pub - 1 < a.n_extra) {                  // per-structure extras, at the end of the block
        const int32_t x = i - a.n_pub - 1;
        const double v = a.extra[x];
        const int64_t o = (int64_t)a.stride - a.n_extra + x;
        if (a.send) a.send[o] = v;
        if (p.on)
            for (int q = 0; q < p.world; ++q) p.peer_recv[q][slot + o] = v;
    }
    if (!p.on) return;
    __threadfence_system();
    __syncthreads();
    if (threadIdx.x == 0) {
        unsigned *counter = reinterpret_cast<unsigned *>(p.local + p.world + 2);
        const unsigned nblk = gridDim.x - n_cross_blocks;
        if (atomicAdd(counter, 1u) == nblk - 1) {             // every publishing block has stored and fenced
            *counter = 0u;
            __threadfence_system();
            for (int q = 0; q < p.world; ++q)
                asm volatile("st.release.sys.global.s64 [%0], %1;" ::"l"(p.peer_flags[q] + p.rank), "l"(step) : "memory");
        }
    }
}

// ---- trunk cells: exact range sums + the reference's REAL4 rounding drift -----------------------------------------------
// The reference (PCRaster accuflux, oracle/ldd_oracle.c) rounds every cell's sum to REAL4, a_c = rnd(m_c + sum_d a_d).
// Along the ~10^4-cell main stem of a 10^8-cell basin those roundings add up to ~1e-5 of the flux, which an exact
// (float64) range sum E_c does not contain.  The drift is itself (almost) linear: with r_c = rnd(E_c), the local term
//     g_c = rnd(E_c + sum_{trunk donors t} (r_t - E_t)) - r_c          (a small multiple of ulp(r_c), mostly 0 or +-1 ulp)
// is what cell c's own rounding adds given exactly-rounded inputs, rounding commutes with shifts by whole ulps, and
// therefore a_c = rnd(r_c + D_c) with D_c = sum of g over the trunk cells upstream of c -- again a subtree range sum, over the
// trunk list in its DFS order.  Deviations from the sequential recurrence (binade crossings, confluences of unequal binades)
// are O(1 ulp) per event and do not amplify.  The sums of g are exact in float64 (multiples of REAL4 ulps), so the
// result does not depend on the block decomposition or the number of ranks.
struct TrunkArgs {
    int32_t m, nblk;
    const int32_t *cell, *last, *tdon;
    const uint8_t *trunk;
    const int32_t *brank, *bslot, *erank, *eslot;
    double *E, *Gl, *BS, *BP;
    unsigned *counter;
    const double *gathered;               // [world][rank_pitch]: per rank [ncomp][stride] published blocks
    int64_t rank_pitch, part_begin;
    int32_t stride, world;
    const int *timed_out;                 // peer-memory exchange: set when a bounded flag wait gave up (else NULL)
    P2PArgs p2p;
    // lakes / reservoirs (k_trunk_apply_adv): static tables of sphy_reslake_plan, the structures' outflow of the day
    double *H;                            // [m] global prefix at the end of each entry's range (the noise scale of E), or NULL
    const int32_t *cut_slot, *entry_cut, *entry_ptr, *entry_idx;
    const float *qout;
    float *qday;
};

constexpr int TRUNK_BLOCK = TAIL_THREADS;

__global__ void __launch_bounds__(TRUNK_BLOCK, 5) k_trunk_drift(const __grid_constant__ ScanArgs a, const __grid_constant__ TrunkArgs t)
{
    typedef cub::BlockScan<double, TRUNK_BLOCK> Scan;
    __shared__ typename Scan::TempStorage tmp;
    __shared__ double base[64];
    __shared__ bool is_last;
    const int tid = threadIdx.x;
    const int32_t s = blockIdx.x * TRUNK_BLOCK + tid;
    const double *gathered = t.gathered;
    long long step = 0;
    if (t.p2p.on) {                                            // acquire the blocks the peers stored into this rank's buffer
        const P2PArgs &p = t.p2p;
        step = *reinterpret_cast<const volatile long long *>(p.local + p.world + 1) + 1;
        if (tid < p.world) {
            long long v = 0, it = 0;
            for (; it < p.max_naps; ++it) {
                asm volatile("ld.acquire.sys.global.s64 %0, [%1];" : "=l"(v) : "l"(p.local + tid) : "memory");
                if (v >= step) break;
                __nanosleep(100);
            }
            if (it >= p.max_naps) *reinterpret_cast<volatile int *>(p.local + p.world) = 1;
        }
        gathered = p.recv + (size_t)((step - 1) & 1) * p.world * p.pitch;
    }
    for (int c = 0; c < a.ncomp; ++c) {
        __syncthreads();
        if (tid < t.world) {                                   // global prefix at the start of every rank's range
            double b = 0.0;
            for (int u = 0; u < tid; ++u) b += gathered[(int64_t)u * t.rank_pitch + (int64_t)c * t.stride];
            base[tid] = b;
        }
        __syncthreads();
        const double *V = gathered + (int64_t)c * t.stride + 1;
        double hi_last = 0.0;
        auto eval = [&](int32_t e) {                           // exact accumulated flux of trunk entry e
            const int32_t er = __ldg(t.erank + e), br = __ldg(t.brank + e);
            const double hi = base[er] + V[(int64_t)er * t.rank_pitch + __ldg(t.eslot + e)];
            const double lo = br < 0 ? 0.0 : base[br] + V[(int64_t)br * t.rank_pitch + __ldg(t.bslot + e)];
            hi_last = hi;
            return range_diff(hi, lo);
        };
        double g = 0.0;
        if (s < t.m) {
            const double E = eval(s);
            t.E[(int64_t)c * t.m + s] = E;
            if (t.H && c == 0) t.H[s] = hi_last;
            if (t.trunk[s]) {
                double X = E;
                for (int k = 0; k < 8; ++k) {
                    const int32_t ts = __ldg(t.tdon + (int64_t)s * 8 + k);
                    if (ts < 0) break;
                    const double Et = eval(ts);
                    X += (double)(float)Et - Et;               // the donor hands over rnd(E_t), not E_t
                }
                g = (double)(float)X - (double)(float)E;
            }
        }
        double incl, total;
        Scan(tmp).InclusiveSum(g, incl, total);
        if (s < t.m) t.Gl[(int64_t)c * t.m + s] = incl;
        if (tid == 0) t.BS[(int64_t)c * t.nblk + blockIdx.x] = total;
    }
    // the last block to finish turns the block totals into their exclusive prefix
    __threadfence();
    __syncthreads();
    if (tid == 0) is_last = atomicAdd(t.counter, 1u) == gridDim.x - 1;
    __syncthreads();
    if (!is_last) return;
    __threadfence();
    for (int c = 0; c < a.ncomp; ++c) {
        double carry = 0.0;
        for (int32_t b0 = 0; b0 < t.nblk; b0 += TRUNK_BLOCK) {
            const int32_t b = b0 + tid;
            const double v = b < t.nblk ? __ldcg(t.BS + (int64_t)c * t.nblk + b) : 0.0;
            double ex, tot;
            __syncthreads();
            Scan(tmp).ExclusiveSum(v, ex, tot);
            if (b < t.nblk) t.BP[(int64_t)c * t.nblk + b] = carry + ex;
            carry += tot;
        }
    }
    if (tid == 0) {
        *t.counter = 0u;
        if (t.p2p.on) *(t.p2p.local + t.p2p.world + 1) = step;    // this exchange is consumed: the next push may overwrite its parity
    }
}

__global__ void __launch_bounds__(TAIL_THREADS, 5) k_trunk_apply(const __grid_constant__ ScanArgs a, const __grid_constant__ TrunkArgs t)
{
    const int32_t s = blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= t.m) return;
    const int64_t r = (int64_t)__ldg(t.cell + s) - t.part_begin;
    if (r < 0 || r >= a.n) return;                             // another rank's cell
    const bool drift = t.trunk[s] != 0;
    const int32_t l = __ldg(t.last + s);
    // a peer that never delivered its block: make the failure visible in the data instead of routing stale prefixes
    const bool dead = t.timed_out && *reinterpret_cast<const volatile int *>(t.timed_out) != 0;
    for (int c = 0; c < a.ncomp; ++c) {
        const double *Gl = t.Gl + (int64_t)c * t.m, *BP = t.BP + (int64_t)c * t.nblk;
        double D = 0.0;
        if (drift) D = (BP[l / TRUNK_BLOCK] + Gl[l]) - (s > 0 ? BP[(s - 1) / TRUNK_BLOCK] + Gl[s - 1] : 0.0);
        const float f = (float)((double)(float)t.E[(int64_t)c * t.m + s] + D);
        const float kx = a.kxc[c].map ? __ldg(a.kxc[c].map + r) : a.kxc[c].value;
        const float omk = a.kxc[c].map ? 1.0f - kx : a.omkc[c];
        float *qs = a.qstate[c];
        qs[r] = dead ? nanf("") : omk * f + kx * qs[r];        // routing.py:28
    }
}

// Listed cells of a basin with lakes / reservoirs (sphy_route_reslake_finish).  A structure keeps everything of its upstream
// range and hands on its outflow, so flux(c) = E(c) - sum over the top-level structures k strictly inside the range of c of
// (E(k) - Qout(k)); then the recession blend of advanced_routing.py:31-37 (a structure cell takes its outflow).
__global__ void __launch_bounds__(TAIL_THREADS, 5) k_trunk_apply_adv(const __grid_constant__ ScanArgs a, const __grid_constant__ TrunkArgs t)
{
    const int32_t s = blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= t.m) return;
    const int64_t r = (int64_t)__ldg(t.cell + s) - t.part_begin;
    if (r < 0 || r >= a.n) return;                             // another rank's cell
    const bool dead = t.timed_out && *reinterpret_cast<const volatile int *>(t.timed_out) != 0;
    const double E = t.E[s];
    double dz = 0.0;
    for (int32_t q = __ldg(t.entry_ptr + s); q < __ldg(t.entry_ptr + s + 1); ++q) {
        const int32_t k = __ldg(t.entry_idx + q);
        dz += t.E[__ldg(t.cut_slot + k)] - (double)t.qout[k];
    }
    double F = E - dz;
    if (fabs(F) <= 3.6e-15 * fabs(t.H[s])) F = 0.0;           // inside the rounding noise of the prefixes (see range_diff)
    const float f = (float)F;
    const float kx = a.kxc[0].map ? __ldg(a.kxc[0].map + r) : a.kxc[0].value;
    const float omk = a.kxc[0].map ? 1.0f - kx : a.omkc[0];
    float *qs = a.qstate[0];
    const int32_t ec = __ldg(t.entry_cut + s);
    const float qd = ec >= 0 ? t.qout[ec] : omk * f + ((qs[r] * 24.0f) * 3600.0f) * kx;    // advanced_routing.py:31-33
    t.qday[r] = dead ? nanf("") : qd;
    qs[r] = dead ? nanf("") : qd / 86400.0f;
}

// exclusive prefix of the tile totals (ntiles + 1 entries per component: the trailing 0 makes X[ntiles] the range total);
// one CTA per component: for ranges of at most 4096 tiles (small basins, where launches dominate) one launch instead of
// a library scan per component; larger ranges use cub's single-pass scan
template <int THREADS>
__global__ void __launch_bounds__(THREADS) k_tile_scan(const double *__restrict__ tot, double *__restrict__ excl, int32_t count)
{
    typedef cub::BlockScan<double, THREADS> Scan;
    __shared__ typename Scan::TempStorage tmp;
    const double *T = tot + (int64_t)blockIdx.x * count;
    double *X = excl + (int64_t)blockIdx.x * count;
    constexpr int PER = 4;
    double carry = 0.0;
    for (int32_t b0 = 0; b0 < count; b0 += THREADS * PER) {
        double v[PER], ex[PER], total;
#pragma unroll
        for (int k = 0; k < PER; ++k) {
            const int32_t i = b0 + threadIdx.x * PER + k;
            v[k] = i < count ? T[i] : 0.0;
        }
        __syncthreads();
        Scan(tmp).ExclusiveSum(v, ex, total);
#pragma unroll
        for (int k = 0; k < PER; ++k) {
            const int32_t i = b0 + threadIdx.x * PER + k;
            if (i < count) X[i] = carry + ex[k];
        }
        carry += total;
    }
}

// The same scan for the overlapped tail: blocks of 128 threads scan 1024 entries each (local exclusive prefix) and the last
// block to finish turns the block totals into their exclusive prefix; readers add the two levels (tile_prefix).  Small
// CTAs, many of them: they fit next to the resident CTAs of the water-balance kernel and still finish in microseconds.
__global__ void __launch_bounds__(TAIL_THREADS) k_tile_scan_small(const double *__restrict__ tot, double *__restrict__ excl, int32_t count,
                                                                  double *__restrict__ base, int32_t nb, unsigned *counter)
{
    typedef cub::BlockScan<double, TAIL_THREADS> Scan;
    __shared__ typename Scan::TempStorage tmp;
    __shared__ bool is_last;
    constexpr int PER = 1024 / TAIL_THREADS;
    const int c = blockIdx.y;
    const double *T = tot + (int64_t)c * count;
    double *X = excl + (int64_t)c * count;
    double v[PER], ex[PER], total;
#pragma unroll
    for (int k = 0; k < PER; ++k) {
        const int32_t i = blockIdx.x * 1024 + threadIdx.x * PER + k;
        v[k] = i < count ? T[i] : 0.0;
    }
    Scan(tmp).ExclusiveSum(v, ex, total);
#pragma unroll
    for (int k = 0; k < PER; ++k) {
        const int32_t i = blockIdx.x * 1024 + threadIdx.x * PER + k;
        if (i < count) X[i] = ex[k];
    }
    if (threadIdx.x == 0) base[(int64_t)c * nb + blockIdx.x] = total;          // the block total, replaced by the prefix below
    __threadfence();
    __syncthreads();
    if (threadIdx.x == 0) is_last = atomicAdd(counter, 1u) == gridDim.x * gridDim.y - 1;
    __syncthreads();
    if (!is_last) return;
    __threadfence();
    for (int cc = 0; cc < (int)gridDim.y; ++cc) {
        double carry = 0.0;
        for (int32_t b0 = 0; b0 < nb; b0 += TAIL_THREADS) {
            const int32_t b = b0 + threadIdx.x;
            const double bt = b < nb ? __ldcg(base + (int64_t)cc * nb + b) : 0.0;
            double e, t2;
            __syncthreads();
            Scan(tmp).ExclusiveSum(bt, e, t2);
            if (b < nb) base[(int64_t)cc * nb + b] = carry + e;
            carry += t2;
        }
    }
    if (threadIdx.x == 0) *counter = 0u;
}

// ---- one-time construction of the crossing lists ---------------------------------------------------------
__global__ void k_cross_flag(const int32_t *sub_size, int64_t n, uint8_t *flag)
{
    int64_t r = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (r < n) flag[r] = (r / SCAN_TILE) != ((r + sub_size[r] - 1) / SCAN_TILE);
}

// trunk cells are finished by k_trunk_drift / k_trunk_apply, not by k_scan_cross
__global__ void k_cross_unflag_trunk(const int32_t *sl_cell, int32_t m, int64_t part_begin, int64_t n, uint8_t *flag)
{
    const int32_t s = blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= m) return;
    const int64_t r = (int64_t)sl_cell[s] - part_begin;
    if (r >= 0 && r < n) flag[r] = 0;
}

// sort keys of the "end" list: range ends of the crossing cells (INT32_MAX = ends on another rank: sorted last and
// cut off), followed by the positions published for other ranks (slots n_cross + j)
__global__ void k_cross_ends(const int32_t *cross_r, const int32_t *sub_size, int32_t nc, int64_t n, const int32_t *pub_e,
                             int32_t n_pub, int32_t *e, int32_t *k, int32_t *cross_e)
{
    int32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < nc) {
        const int64_t end = (int64_t)cross_r[i] + sub_size[cross_r[i]] - 1;
        cross_e[i] = (int32_t)end;                 // < 2^31: sub_size is int32 over the whole basin
        e[i] = end < n ? (int32_t)end : INT32_MAX;
        k[i] = i;
    } else if (i < nc + n_pub) {
        e[i] = pub_e[i - nc];
        k[i] = i;
    }
}

// ptr[t] = first position in the sorted key list with key >= t * SCAN_TILE (t = 0..ntiles)
__global__ void k_tile_lower_bound(const int32_t *keys, int32_t nk, int32_t ntiles, int32_t *ptr)
{
    int32_t t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t > ntiles) return;
    const int64_t want = (int64_t)t * SCAN_TILE;
    int32_t lo = 0, hi = nk;
    while (lo < hi) {
        const int32_t mid = (lo + hi) >> 1;
        if (keys[mid] < want) lo = mid + 1; else hi = mid;
    }
    ptr[t] = lo;
}

}  // namespace

static int build_cross_lists(sphy_ctx *ctx, cudaStream_t st)
{
    const int64_t n = ctx->n;
    const int32_t ntiles = (int32_t)((n + SCAN_TILE - 1) / SCAN_TILE);
    ctx->n_tiles = ntiles;
    uint8_t *flag = nullptr;
    int32_t *sel = nullptr, *nsel = nullptr, *e_in = nullptr, *k_in = nullptr;
    void *tmp = nullptr;
    size_t tb = 0;
    cudaError_t e = cudaSuccess;
#define FK(call) do { if (e == cudaSuccess) e = (call); } while (0)
    FK(cudaMalloc(&flag, (size_t)n));
    FK(cudaMalloc(&sel, sizeof(int32_t) * n));
    FK(cudaMalloc(&nsel, sizeof(int32_t)));
    int32_t nc = 0;
    if (e == cudaSuccess) {
        k_cross_flag<<<div_up(n, 256), 256, 0, st>>>(ctx->sub_size_loc, n, flag);
        if (ctx->sl_plan && ctx->sl_n > 0)
            k_cross_unflag_trunk<<<div_up(ctx->sl_n, 256), 256, 0, st>>>(ctx->sl_cell, ctx->sl_n, ctx->part_begin, n, flag);
        cub::CountingInputIterator<int32_t> counting(0);
        FK(cub::DeviceSelect::Flagged(nullptr, tb, counting, flag, sel, nsel, (int)n, st));
        FK(cudaMalloc(&tmp, tb));
        FK(cub::DeviceSelect::Flagged(tmp, tb, counting, flag, sel, nsel, (int)n, st));
        FK(cudaMemcpyAsync(&nc, nsel, sizeof(int32_t), cudaMemcpyDeviceToHost, st));
        FK(cudaStreamSynchronize(st));
    }
    ctx->n_cross = nc;
    const int32_t nend_all = nc + ctx->n_pub;
    const size_t ncs = nc > 0 ? (size_t)nc : 1, nes = nend_all > 0 ? (size_t)nend_all : 1;
    int rc = SPHY_OK;
    if (e == cudaSuccess &&
        ((rc = dev_alloc(ctx, &ctx->cross_r, ncs)) || (rc = dev_alloc(ctx, &ctx->cross_e, ncs)) ||
         (rc = dev_alloc(ctx, &ctx->cross_ptr, (size_t)ntiles + 1)) ||
         (rc = dev_alloc(ctx, &ctx->end_e, nes)) || (rc = dev_alloc(ctx, &ctx->end_k, nes)) ||
         (rc = dev_alloc(ctx, &ctx->end_ptr, (size_t)ntiles + 1)))) {
        cudaFree(flag); cudaFree(sel); cudaFree(nsel); cudaFree(tmp);
        return rc;
    }
    if (e == cudaSuccess && nc > 0) FK(cudaMemcpyAsync(ctx->cross_r, sel, sizeof(int32_t) * nc, cudaMemcpyDeviceToDevice, st));
    if (e == cudaSuccess && nend_all > 0) {
        FK(cudaMalloc(&e_in, sizeof(int32_t) * nend_all));
        FK(cudaMalloc(&k_in, sizeof(int32_t) * nend_all));
        if (e == cudaSuccess) {
            k_cross_ends<<<div_up(nend_all, 256), 256, 0, st>>>(ctx->cross_r, ctx->sub_size_loc, nc, n, ctx->pub_e, ctx->n_pub, e_in, k_in, ctx->cross_e);
            size_t need = 0;
            FK(cub::DeviceRadixSort::SortPairs(nullptr, need, e_in, ctx->end_e, k_in, ctx->end_k, nend_all, 0, 32, st));
            if (need > tb) { cudaFree(tmp); tmp = nullptr; FK(cudaMalloc(&tmp, need)); tb = need; }
            FK(cub::DeviceRadixSort::SortPairs(tmp, need, e_in, ctx->end_e, k_in, ctx->end_k, nend_all, 0, 32, st));
        }
    }
    if (e == cudaSuccess) {
        // keys >= ntiles*TILE (the INT32_MAX remote ends) fall behind end_ptr[ntiles] and are never visited
        k_tile_lower_bound<<<div_up(ntiles + 1, 256), 256, 0, st>>>(ctx->cross_r, nc, ntiles, ctx->cross_ptr);
        k_tile_lower_bound<<<div_up(ntiles + 1, 256), 256, 0, st>>>(ctx->end_e, nend_all, ntiles, ctx->end_ptr);
        FK(cudaGetLastError());
        FK(cudaStreamSynchronize(st));
    }
#undef FK
    cudaFree(flag); cudaFree(sel); cudaFree(nsel); cudaFree(e_in); cudaFree(k_in); cudaFree(tmp);
    if (e != cudaSuccess) SPHY_FAIL(ctx, SPHY_E_CUDA, "build_cross_lists: %s", cudaGetErrorString(e));
    ctx->scan_ready = true;
    return SPHY_OK;
}

static int scan_prepare(sphy_ctx *ctx, int ncomp, const float *const *runoff_mm, float *const *q_state, sphy_param kx,
                        float one_m_kx, cudaStream_t st, ScanArgs &a)
{
    if (ctx->order_mode != SPHY_ORDER_DFS)
        SPHY_FAIL(ctx, SPHY_E_STATE, "SPHY_ROUTE_SCAN needs the cells in SPHY_ORDER_DFS (sphy_set_cell_order)");
    if (!ctx->scan_ready) {                       // one-time list build (synchronises, allocates)
        int rc = build_cross_lists(ctx, st);
        if (rc != SPHY_OK) return rc;
    }
    if (ctx->scan_comp_cap < ncomp) {             // grows only when more components are first routed
        int rc;
        const size_t ncs = ctx->n_cross > 0 ? (size_t)ctx->n_cross : 1;
        if ((rc = dev_alloc(ctx, &ctx->park_before, (size_t)ncomp * ncs)) ||
            (rc = dev_alloc(ctx, &ctx->park_end, (size_t)ncomp * (ncs + ctx->n_pub))) ||
            (rc = dev_alloc(ctx, &ctx->scan_tile, (size_t)ncomp * (ctx->n_tiles + 1))) ||
            (rc = dev_alloc(ctx, &ctx->scan_excl, (size_t)ncomp * (ctx->n_tiles + 1))))
            return rc;
        SPHY_CUDA(ctx, cudaMemsetAsync(ctx->scan_tile, 0, sizeof(double) * (size_t)ncomp * (ctx->n_tiles + 1), st));
        if ((rc = dev_alloc(ctx, &ctx->scan_base, (size_t)ncomp * div_up(ctx->n_tiles + 1, 1024))) ||
            (rc = dev_alloc_once(ctx, &ctx->scan_counter, 1)))
            return rc;
        SPHY_CUDA(ctx, cudaMemsetAsync(ctx->scan_counter, 0, sizeof(unsigned), st));
        size_t need = 0;
        SPHY_CUDA(ctx, cub::DeviceScan::ExclusiveSum(nullptr, need, ctx->scan_tile, ctx->scan_excl, ctx->n_tiles + 1, st));
        if (need > ctx->scan_tmp_bytes) {
            if (ctx->scan_tmp) cudaFree(ctx->scan_tmp);
            ctx->scan_tmp = nullptr;
            SPHY_CUDA(ctx, cudaMalloc(&ctx->scan_tmp, need));
            ctx->scan_tmp_bytes = need;
        }
        ctx->scan_comp_cap = ncomp;
    }
    for (int c = 0; c < ncomp; ++c)
        if ((reinterpret_cast<uintptr_t>(runoff_mm ? runoff_mm[c] : nullptr) | reinterpret_cast<uintptr_t>(q_state[c])) & 15u)
            SPHY_FAIL(ctx, SPHY_E_INVALID, "scan routing: maps must be 16-byte aligned");
    a = ScanArgs{};
    a.ncomp = ncomp;
    for (int c = 0; c < ncomp; ++c) {
        a.m[c] = runoff_mm ? runoff_mm[c] : nullptr;
        a.qstate[c] = q_state[c];
    }
    a.sub_size = ctx->sub_size_loc;
    a.cross_r = ctx->cross_r;
    a.cross_e = ctx->cross_e;
    a.cross_ptr = ctx->cross_ptr;
    a.end_e = ctx->end_e;
    a.end_k = ctx->end_k;
    a.end_ptr = ctx->end_ptr;
    a.n_cross = ctx->n_cross;
    a.park_before = ctx->park_before;
    a.park_end = ctx->park_end;
    a.tile_tot = ctx->scan_tile;
    a.tile_excl = ctx->scan_excl;
    a.n_pub = ctx->n_pub;
    a.pub_e = ctx->pub_e;
    for (int c = 0; c < ncomp; ++c) {
        a.kxc[c] = kx;
        a.omkc[c] = one_m_kx;
    }
    a.to_m3s = (0.001 * (double)ctx->cellarea) / 86400.0;
    a.volf = 1.0f;
    a.n = ctx->n;
    a.ntiles = ctx->n_tiles;
    return SPHY_OK;
}

// main pass -> exclusive scan of the tile totals (cub, one call per component) -> crossing cells
static void prof_mark(sphy_ctx *ctx, cudaStream_t st)
{
    if (!ctx->prof_on || ctx->prof_ev.size() >= 4 * 4096) return;
    cudaEvent_t e;
    if (cudaEventCreate(&e) != cudaSuccess) return;
    cudaEventRecord(e, st);
    ctx->prof_ev.push_back(e);
}

static P2PArgs p2p_args(const sphy_ctx *ctx, bool on)
{
    P2PArgs p = {};
    if (!on) return p;
    static const long long max_naps = [] {
        const char *e = getenv("SPHY_P2P_TIMEOUT_S");
        const double sec = e ? atof(e) : 30.0;
        return (long long)((sec > 0.01 ? sec : 0.01) * 1e7);  // 100 ns naps
    }();
    p.on = 1;
    p.rank = ctx->p2p_rank;
    p.world = ctx->p2p_world;
    p.peer_recv = ctx->p2p_peer_recv;
    p.peer_flags = ctx->p2p_peer_flags;
    p.local = ctx->p2p_flags;
    p.recv = ctx->p2p_recv;
    p.pitch = ctx->p2p_pitch;
    p.max_naps = max_naps;
    return p;
}

// main pass -> exclusive scan of the tile totals -> crossing cells + publishing (+ push into the peers' buffers)
static int scan_launch_local(sphy_ctx *ctx, ScanArgs &a, cudaStream_t st, bool publish = false, bool push = false)
{
    const int resident = ctx->sm_count * SCAN_MIN_CTAS;
    const int want = (ctx->n_tiles + SCAN_WARPS - 1) / SCAN_WARPS;       // one tile per warp
    const int grid = want < resident ? want : resident;
    if (!ctx->scan_smem_opt_in) {                              // per context: the attribute is per device
        SPHY_CUDA(ctx, cudaFuncSetAttribute(k_scan_main<true, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, SCAN_SMEM));
        SPHY_CUDA(ctx, cudaFuncSetAttribute(k_scan_main<true, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, SCAN_SMEM));
        SPHY_CUDA(ctx, cudaFuncSetAttribute(k_scan_main<false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, SCAN_SMEM));
        SPHY_CUDA(ctx, cudaFuncSetAttribute(k_scan_main<false, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, SCAN_SMEM));
        ctx->scan_smem_opt_in = true;
    }
    prof_mark(ctx, st);
    if (a.adv) {
        if (a.kxc[0].map) k_scan_main<true, true><<<grid, SCAN_THREADS, SCAN_SMEM, st>>>(a);
        else k_scan_main<true, false><<<grid, SCAN_THREADS, SCAN_SMEM, st>>>(a);
    } else {
        if (a.kxc[0].map) k_scan_main<false, true><<<grid, SCAN_THREADS, SCAN_SMEM, st>>>(a);
        else k_scan_main<false, false><<<grid, SCAN_THREADS, SCAN_SMEM, st>>>(a);
    }
    prof_mark(ctx, st);
    if (ctx->main_pass_event) SPHY_CUDA(ctx, cudaEventRecord((cudaEvent_t)ctx->main_pass_event, st));
    if (ctx->overlap_tail) {                                   // small CTAs that fit beside the next water-balance kernel
        const int nb = div_up(ctx->n_tiles + 1, 1024);
        k_tile_scan_small<<<dim3(nb, a.ncomp), TAIL_THREADS, 0, st>>>(ctx->scan_tile, ctx->scan_excl, ctx->n_tiles + 1, ctx->scan_base, nb,
                                                                      ctx->scan_counter);
        a.tile_base = ctx->scan_base;
        a.tile_nb = nb;
    } else if (ctx->n_tiles + 1 <= 4 * 1024) {                 // small ranges: one CTA per component, one launch
        k_tile_scan<1024><<<a.ncomp, 1024, 0, st>>>(ctx->scan_tile, ctx->scan_excl, ctx->n_tiles + 1);
    } else {                                                   // single-pass decoupled look-back scan of the library
        for (int c = 0; c < a.ncomp; ++c) {
            size_t need = ctx->scan_tmp_bytes;
            SPHY_CUDA(ctx, cub::DeviceScan::ExclusiveSum(ctx->scan_tmp, need, ctx->scan_tile + (size_t)c * (ctx->n_tiles + 1),
                                                         ctx->scan_excl + (size_t)c * (ctx->n_tiles + 1), ctx->n_tiles + 1, st));
        }
    }
    const int ncb = div_up(ctx->n_cross, TAIL_THREADS * CROSS_ILP);
    const int npb = publish ? div_up(ctx->n_pub + 1 + a.n_extra, TAIL_THREADS) : 0;
    if (ncb + npb > 0) k_scan_cross_pub<<<ncb + npb, TAIL_THREADS, 0, st>>>(a, p2p_args(ctx, push), ncb);
    prof_mark(ctx, st);
    return SPHY_OK;
}

// workspaces of the trunk kernels for `ncomp` routed components
static int trunk_prepare(sphy_ctx *ctx, int ncomp, cudaStream_t st)
{
    if (ctx->sl_comp_cap >= ncomp) return SPHY_OK;
    int rc;
    const size_t m = ctx->sl_n > 0 ? (size_t)ctx->sl_n : 1, nb = ctx->sl_nblk > 0 ? (size_t)ctx->sl_nblk : 1;
    if ((rc = dev_alloc(ctx, &ctx->sl_E, ncomp * m)) || (rc = dev_alloc(ctx, &ctx->sl_Gl, ncomp * m)) ||
        (rc = dev_alloc(ctx, &ctx->sl_BS, ncomp * nb)) || (rc = dev_alloc(ctx, &ctx->sl_BP, ncomp * nb)) ||
        (rc = dev_alloc(ctx, &ctx->sl_V, (size_t)ncomp * ctx->sl_stride)) || (rc = dev_alloc_once(ctx, &ctx->sl_counter, 1)))
        return rc;
    SPHY_CUDA(ctx, cudaMemsetAsync(ctx->sl_counter, 0, sizeof(unsigned), st));
    ctx->sl_comp_cap = ncomp;
    return SPHY_OK;
}

// finish the trunk cells of this context from the published blocks of all ranks
static TrunkArgs trunk_args(sphy_ctx *ctx, const double *gathered, int64_t rank_pitch, bool p2p)
{
    TrunkArgs t = {};
    t.m = ctx->sl_n;
    t.nblk = ctx->sl_nblk;
    t.cell = ctx->sl_cell;
    t.last = ctx->sl_last;
    t.tdon = ctx->sl_tdon;
    t.trunk = ctx->sl_trunk;
    t.brank = ctx->sl_brank;
    t.bslot = ctx->sl_bslot;
    t.erank = ctx->sl_erank;
    t.eslot = ctx->sl_eslot;
    t.E = ctx->sl_E;
    t.Gl = ctx->sl_Gl;
    t.BS = ctx->sl_BS;
    t.BP = ctx->sl_BP;
    t.counter = ctx->sl_counter;
    t.gathered = gathered;
    t.rank_pitch = rank_pitch;
    t.part_begin = ctx->part_begin;
    t.stride = ctx->sl_stride;
    t.world = ctx->sl_world;
    t.p2p = p2p_args(ctx, p2p);
    t.timed_out = p2p ? reinterpret_cast<const int *>(ctx->p2p_flags + ctx->p2p_world) : nullptr;
    return t;
}

static int trunk_finish(sphy_ctx *ctx, const ScanArgs &a, const double *gathered, int64_t rank_pitch, cudaStream_t st, bool p2p = false)
{
    if (ctx->sl_n == 0) return SPHY_OK;
    const TrunkArgs t = trunk_args(ctx, gathered, rank_pitch, p2p);
    k_trunk_drift<<<ctx->sl_nblk, TRUNK_BLOCK, 0, st>>>(a, t);
    k_trunk_apply<<<div_up(ctx->sl_n, TAIL_THREADS), TAIL_THREADS, 0, st>>>(a, t);
    return SPHY_OK;
}

int route_scan_run(sphy_ctx *ctx, int ncomp, const float *const *runoff_mm, float *const *q_state, sphy_param kx,
                   float one_m_kx, cudaStream_t st)
{
    if (ctx->sl_plan && ctx->sl_cuts) SPHY_FAIL(ctx, SPHY_E_STATE, "the trunk list holds lake/reservoir cells: route with sphy_route_reslake_begin/_finish");
    ScanArgs a;
    int rc = scan_prepare(ctx, ncomp, runoff_mm, q_state, kx, one_m_kx, st, a);
    if (rc != SPHY_OK) return rc;
    if (ctx->sl_plan && (rc = trunk_prepare(ctx, ncomp, st)) != SPHY_OK) return rc;
    if (ctx->sl_plan) {
        a.send = ctx->sl_V;
        a.stride = ctx->sl_stride;
    }
    const bool trunk = ctx->sl_plan && ctx->sl_n > 0;
    int rc2 = scan_launch_local(ctx, a, st, trunk);
    if (rc2 != SPHY_OK) return rc2;
    if (trunk && (rc = trunk_finish(ctx, a, ctx->sl_V, (int64_t)ncomp * ctx->sl_stride, st)) != SPHY_OK) return rc;
    prof_mark(ctx, st);
    SPHY_LAUNCH_CHECK(ctx);
    return SPHY_OK;
}

// several runoff maps with a recession coefficient each: ensemble members sharing one LDD route in one pass, so the index
// stream (sub_size, the crossing and trunk lists) is read once for all of them
extern "C" int sphy_route_step_members(sphy_ctx *ctx, int nmaps, const float *const *runoff_mm, float *const *q_state,
                                       const sphy_param *kx, const float *one_minus_kx, void *stream)
{
    if (!ctx || !runoff_mm || !q_state || !kx || !one_minus_kx || nmaps < 1 || nmaps > ROUTE_MAX_COMP) return SPHY_E_INVALID;
    if (!ctx->ldd_ready) SPHY_FAIL(ctx, SPHY_E_STATE, "sphy_route_step_members before sphy_ldd_build");
    if (ctx->partitioned) SPHY_FAIL(ctx, SPHY_E_STATE, "sphy_route_step_members is not available on a partitioned context");
    if (ctx->sl_plan && ctx->sl_cuts) SPHY_FAIL(ctx, SPHY_E_STATE, "the trunk list holds lake/reservoir cells: route with sphy_route_reslake_begin/_finish");
    for (int c = 0; c < nmaps; ++c) {
        if (!runoff_mm[c] || !q_state[c]) SPHY_FAIL(ctx, SPHY_E_INVALID, "sphy_route_step_members: map %d is NULL", c);
        if ((kx[c].map != nullptr) != (kx[0].map != nullptr))
            SPHY_FAIL(ctx, SPHY_E_INVALID, "sphy_route_step_members: kx must be a map for every member or for none");
    }
    cudaStream_t st = (cudaStream_t)stream;
    ScanArgs a;
    int rc = scan_prepare(ctx, nmaps, runoff_mm, q_state, kx[0], one_minus_kx[0], st, a);
    if (rc != SPHY_OK) return rc;
    for (int c = 0; c < nmaps; ++c) {
        a.kxc[c] = kx[c];
        a.omkc[c] = one_minus_kx[c];
    }
    if (ctx->sl_plan && (rc = trunk_prepare(ctx, nmaps, st)) != SPHY_OK) return rc;
    if (ctx->sl_plan) {
        a.send = ctx->sl_V;
        a.stride = ctx->sl_stride;
    }
    const bool trunk = ctx->sl_plan && ctx->sl_n > 0;
    if ((rc = scan_launch_local(ctx, a, st, trunk)) != SPHY_OK) return rc;
    if (trunk && (rc = trunk_finish(ctx, a, ctx->sl_V, (int64_t)nmaps * ctx->sl_stride, st)) != SPHY_OK) return rc;
    prof_mark(ctx, st);
    SPHY_LAUNCH_CHECK(ctx);
    return SPHY_OK;
}

// ---- accufractionflux with QFRAC in {0,1} (lakes / reservoirs) on the scan path ---------------------------------
// A cut cell passes nothing downstream, i.e. its whole upstream range must vanish from every range that contains it.
// With R'(c) = (range sum of c) - sum of R'(c') over the cut cells c' strictly inside c's range, placing the material
// -R'(c) at position c makes every plain range sum equal the accufractionflux (and the cut cell's own flux 0).
namespace {

// one CTA per cut cell: REAL8 sum of the material over its upstream range
__global__ void __launch_bounds__(256) k_cut_range_sums(const int32_t *__restrict__ cut_cell, const int32_t *__restrict__ sub_size,
                                                        const float *__restrict__ m, double *__restrict__ R)
{
    typedef cub::BlockReduce<double, 256> Reduce;
    __shared__ typename Reduce::TempStorage tmp;
    const int64_t c = cut_cell[blockIdx.x];
    const int64_t e = c + sub_size[c];
    double s = 0.0;
    for (int64_t i = c + threadIdx.x; i < e; i += 256) s += (double)m[i];
    s = Reduce(tmp).Sum(s);
    if (threadIdx.x == 0) R[blockIdx.x] = s;
}

// cut cells are listed by ascending index, so nested cuts come after the cut that contains them: walk backwards
__global__ void k_cut_corrections(const int32_t *__restrict__ cut_cell, const int32_t *__restrict__ sub_size, int n_cut, double *R)
{
    if (blockIdx.x != 0 || threadIdx.x != 0) return;
    for (int j = n_cut - 1; j >= 0; --j) {
        const int64_t c = cut_cell[j], e = c + sub_size[c];
        double r = R[j];
        for (int k = j + 1; k < n_cut && cut_cell[k] < e; ++k) r += R[k];      // R[k] already holds -R'(c_k)
        R[j] = -r;                                                             // the material placed at the cut cell
    }
}

}  // namespace

int route_scan_run_adv(sphy_ctx *ctx, float *material_m3day, float *q_state, sphy_param kx, float one_m_kx, const float *qfrac,
                       const float *qout_dense, float *qday, int n_cut, const int32_t *cut_cell_dev, double *cut_work, cudaStream_t st)
{
    if (ctx->partitioned) SPHY_FAIL(ctx, SPHY_E_STATE, "lake/reservoir routing is not available on a partitioned context");
    if (ctx->sl_plan && ctx->sl_n > 0)
        SPHY_FAIL(ctx, SPHY_E_STATE, "lake/reservoir routing on the scan path does not take a trunk plan (sphy_trunk_set_plan)");
    ScanArgs a;
    const float *mm[1] = {material_m3day};
    float *qq[1] = {q_state};
    int rc = scan_prepare(ctx, 1, mm, qq, kx, one_m_kx, st, a);
    if (rc != SPHY_OK) return rc;
    if ((rc = dev_alloc_once(ctx, &ctx->cut_ptr, (size_t)ctx->n_tiles + 1)) != SPHY_OK) return rc;
    if (n_cut > 0) {
        k_cut_range_sums<<<n_cut, 256, 0, st>>>(cut_cell_dev, ctx->sub_size_loc, material_m3day, cut_work);
        k_cut_corrections<<<1, 32, 0, st>>>(cut_cell_dev, ctx->sub_size_loc, n_cut, cut_work);
    }
    k_tile_lower_bound<<<div_up(ctx->n_tiles + 1, 256), 256, 0, st>>>(cut_cell_dev, n_cut, ctx->n_tiles, ctx->cut_ptr);
    a.cut_cell = cut_cell_dev;
    a.cut_ptr = ctx->cut_ptr;
    a.cut_corr = cut_work;
    a.adv = 1;
    a.frac = qfrac;
    a.qout_dense = qout_dense;
    a.qday = qday;
    rc = scan_launch_local(ctx, a, st);
    if (rc != SPHY_OK) return rc;
    SPHY_LAUNCH_CHECK(ctx);
    return SPHY_OK;
}

extern "C" int sphy_set_overlap_tail(sphy_ctx *ctx, int on, void *main_pass_event)
{
    if (!ctx) return SPHY_E_INVALID;
    ctx->overlap_tail = on != 0;
    ctx->main_pass_event = on ? main_pass_event : nullptr;
    return SPHY_OK;
}

// ---- per-phase timing ------------------------------------------------------------------------------------------------
extern "C" int sphy_route_profile(sphy_ctx *ctx, int enable)
{
    if (!ctx) return SPHY_E_INVALID;
    for (cudaEvent_t e : ctx->prof_ev) cudaEventDestroy(e);
    ctx->prof_ev.clear();
    ctx->prof_on = enable != 0;
    return SPHY_OK;
}

// mean milliseconds per step of: the main pass, tile totals + crossing cells (+ publish), everything after that up to the
// end of the trunk kernels (on a partitioned context this includes the exchange), and the whole routing call(s)
extern "C" int sphy_route_profile_read(sphy_ctx *ctx, float *ms4, int *steps)
{
    if (!ctx || !ms4 || !steps) return SPHY_E_INVALID;
    const size_t k = ctx->prof_ev.size() / 4;
    *steps = (int)k;
    ms4[0] = ms4[1] = ms4[2] = ms4[3] = 0.0f;
    if (k == 0) return SPHY_OK;
    SPHY_CUDA(ctx, cudaEventSynchronize(ctx->prof_ev[4 * k - 1]));
    double acc[4] = {0, 0, 0, 0};
    for (size_t i = 0; i < k; ++i) {
        float t = 0.0f;
        for (int j = 0; j < 3; ++j) {
            SPHY_CUDA(ctx, cudaEventElapsedTime(&t, ctx->prof_ev[4 * i + j], ctx->prof_ev[4 * i + j + 1]));
            acc[j] += t;
        }
        SPHY_CUDA(ctx, cudaEventElapsedTime(&t, ctx->prof_ev[4 * i], ctx->prof_ev[4 * i + 3]));
        acc[3] += t;
    }
    for (int j = 0; j < 4; ++j) ms4[j] = (float)(acc[j] / (double)k);
    for (cudaEvent_t e : ctx->prof_ev) cudaEventDestroy(e);
    ctx->prof_ev.clear();
    return SPHY_OK;
}

// ---- multi-GPU -------------------------------------------------------------------------------------------------
extern "C" int sphy_scan_tile_cells(void) { return SCAN_TILE; }

extern "C" int sphy_set_partition(sphy_ctx *ctx, int64_t cell_begin, int64_t cell_end, void *stream)
{
    if (!ctx) return SPHY_E_INVALID;
    if (!ctx->ldd_ready) SPHY_FAIL(ctx, SPHY_E_STATE, "sphy_set_partition before sphy_ldd_build");
    if (ctx->order_mode != SPHY_ORDER_DFS) SPHY_FAIL(ctx, SPHY_E_STATE, "sphy_set_partition needs SPHY_ORDER_DFS");
    if (ctx->partitioned) SPHY_FAIL(ctx, SPHY_E_STATE, "context is already partitioned");
    if (cell_begin < 0 || cell_end > ctx->n_global || cell_begin >= cell_end || (cell_begin % SCAN_TILE) != 0)
        SPHY_FAIL(ctx, SPHY_E_INVALID, "sphy_set_partition: bad range (it must start on a %d-cell tile)", SCAN_TILE);
    // Keep only this rank's slice of what the partitioned path reads (sub_size for the scan, flat_active for the raster
    // gather) and release every whole-basin structure and workspace: the level-sweep order / donor tables and their
    // [8][N] flux workspaces, the canonical Appendix-C arrays, the raster index.  Memory per rank then scales with N / world.
    cudaStream_t st = (cudaStream_t)stream;
    const int64_t nl = cell_end - cell_begin;
    int32_t *sub_l = nullptr, *flat_l = nullptr;
    SPHY_CUDA(ctx, cudaMalloc(&sub_l, sizeof(int32_t) * nl));
    SPHY_CUDA(ctx, cudaMalloc(&flat_l, sizeof(int32_t) * nl));
    SPHY_CUDA(ctx, cudaMemcpyAsync(sub_l, ctx->sub_size + cell_begin, sizeof(int32_t) * nl, cudaMemcpyDeviceToDevice, st));
    SPHY_CUDA(ctx, cudaMemcpyAsync(flat_l, ctx->flat_active + cell_begin, sizeof(int32_t) * nl, cudaMemcpyDeviceToDevice, st));
    SPHY_CUDA(ctx, cudaStreamSynchronize(st));
    void **whole[] = {(void **)&ctx->c_cidx, (void **)&ctx->c_flat_active, (void **)&ctx->c_down, (void **)&ctx->c_order, (void **)&ctx->c_pos,
                      (void **)&ctx->c_don_ptr, (void **)&ctx->c_don_idx, (void **)&ctx->cidx, (void **)&ctx->flat_active, (void **)&ctx->down,
                      (void **)&ctx->level, (void **)&ctx->order, (void **)&ctx->don_ptr, (void **)&ctx->don_idx, (void **)&ctx->indeg,
                      (void **)&ctx->nup, (void **)&ctx->pos, (void **)&ctx->dpos_ptr, (void **)&ctx->dpos, (void **)&ctx->pre,
                      (void **)&ctx->cell_canon, (void **)&ctx->sub_size, (void **)&ctx->acc, (void **)&ctx->qsorted, (void **)&ctx->sweep_rec,
                      (void **)&ctx->tmp_n[0], (void **)&ctx->tmp_n[1], (void **)&ctx->tmp_n[2], (void **)&ctx->tmp_n[3],
                      (void **)&ctx->chain_start, (void **)&ctx->chain_k, (void **)&ctx->chain_cell, (void **)&ctx->chain_don,
                      (void **)&ctx->chain_ids};
    for (void **pp : whole) {
        if (*pp) cudaFree(*pp);
        *pp = nullptr;
    }
    ctx->sub_size = sub_l;                        // owned (freed with the context); the _loc views are what the kernels read
    ctx->flat_active = flat_l;
    ctx->sub_size_loc = sub_l;
    ctx->flat_loc = flat_l;
    ctx->part_begin = cell_begin;
    ctx->n = nl;
    ctx->partitioned = true;
    ctx->scan_ready = false;
    ctx->scan_comp_cap = 0;
    ctx->sl_plan = false;
    return SPHY_OK;
}

// ---- trunk list ------------------------------------------------------------------------------------------------------
namespace {

struct RankBounds { int32_t world; int64_t b[65]; };

__device__ __forceinline__ int rank_of(const RankBounds &rb, int64_t j)
{
    int q = 0;
    while (q + 1 < rb.world && j >= rb.b[q + 1]) ++q;
    return q;
}

// 0: ordinary cell, 1: listed because its upstream range leaves its rank, 2: trunk (carries the rounding drift)
__device__ __forceinline__ int32_t lower_bound_i32(const int32_t *a, int32_t n, int64_t key)
{
    int32_t lo = 0, hi = n;
    while (lo < hi) {
        const int32_t mid = (lo + hi) >> 1;
        if (a[mid] < key) lo = mid + 1; else hi = mid;
    }
    return lo;
}

// cut cells (lakes / reservoirs, ascending): every cell whose upstream range contains one of them is listed as well
__global__ void k_trunk_flag(const int32_t *__restrict__ sub_size, const int32_t *__restrict__ level, const int32_t *__restrict__ cell_canon,
                             int64_t n, int32_t min_level, int64_t min_cells, const __grid_constant__ RankBounds rb,
                             const int32_t *__restrict__ cut, int32_t n_cut, uint8_t *flag)
{
    const int64_t j = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (j >= n) return;
    const int32_t sz = sub_size[j];
    uint8_t f = 0;
    if (sz >= min_cells || level[cell_canon[j]] >= min_level) f = 2;
    else if (rb.world > 1 && rank_of(rb, j) != rank_of(rb, j + sz - 1)) f = 1;
    if (f == 0 && n_cut > 0) {
        const int32_t k = lower_bound_i32(cut, n_cut, j);      // first structure cell at or after j
        if (k < n_cut && cut[k] <= j + sz - 1) f = 1;
    }
    flag[j] = f;
}

__global__ void k_trunk_entries(const int32_t *__restrict__ sl_cell, int32_t m, const int32_t *__restrict__ sub_size,
                                const uint8_t *__restrict__ flag, const int32_t *__restrict__ don_ptr, const int32_t *__restrict__ don_idx,
                                int32_t *sl_sub, int32_t *sl_last, uint8_t *sl_trunk, int32_t *sl_tdon)
{
    const int32_t s = blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= m) return;
    const int32_t j = sl_cell[s];
    const int32_t sz = sub_size[j];
    sl_sub[s] = sz;
    sl_trunk[s] = flag[j] == 2;
    sl_last[s] = lower_bound_i32(sl_cell, m, (int64_t)j + sz) - 1;        // last listed cell inside the upstream range (>= s)
    int k = 0;
    if (flag[j] == 2)
        for (int32_t q = don_ptr[j]; q < don_ptr[j + 1]; ++q) {
            const int32_t d = don_idx[q];
            if (flag[d] == 2 && k < 8) sl_tdon[(int64_t)s * 8 + k++] = lower_bound_i32(sl_cell, m, d);
        }
    for (; k < 8; ++k) sl_tdon[(int64_t)s * 8 + k] = -1;
}

}  // namespace

extern "C" int sphy_trunk_select(sphy_ctx *ctx, int32_t min_level, int64_t min_cells, int world, const int64_t *bounds_host,
                                 int64_t *count_out, void *stream)
{
    return sphy_trunk_select_cuts(ctx, 0, nullptr, min_level, min_cells, world, bounds_host, count_out, stream);
}

extern "C" int sphy_trunk_select_cuts(sphy_ctx *ctx, int32_t n_cut, const int32_t *cut_cell_dev, int32_t min_level, int64_t min_cells,
                                      int world, const int64_t *bounds_host, int64_t *count_out, void *stream)
{
    if (!ctx || !count_out || world < 1 || world > 64 || (world > 1 && !bounds_host) || n_cut < 0 || (n_cut > 0 && !cut_cell_dev))
        return SPHY_E_INVALID;
    if (!ctx->ldd_ready) SPHY_FAIL(ctx, SPHY_E_STATE, "sphy_trunk_select before sphy_ldd_build");
    if (ctx->order_mode != SPHY_ORDER_DFS) SPHY_FAIL(ctx, SPHY_E_STATE, "sphy_trunk_select needs SPHY_ORDER_DFS");
    if (ctx->partitioned) SPHY_FAIL(ctx, SPHY_E_STATE, "sphy_trunk_select after sphy_set_partition (it needs the whole-basin structures)");
    // a trunk cell must leave its scan tile (the main pass skips exactly those cells): sub_size > SCAN_TILE
    if (min_level < SCAN_TILE || min_cells <= SCAN_TILE)
        SPHY_FAIL(ctx, SPHY_E_INVALID, "sphy_trunk_select: thresholds must exceed the %d-cell scan tile", SCAN_TILE);
    cudaStream_t st = (cudaStream_t)stream;
    const int64_t n = ctx->n_global;
    RankBounds rb = {};
    rb.world = world;
    for (int q = 0; q <= world; ++q) rb.b[q] = world > 1 ? bounds_host[q] : (q == 0 ? 0 : n);
    if (world > 1) {
        if (rb.b[0] != 0 || rb.b[world] != n) SPHY_FAIL(ctx, SPHY_E_INVALID, "sphy_trunk_select: bounds must cover [0, %lld)", (long long)n);
        for (int q = 0; q < world; ++q)
            if (rb.b[q] >= rb.b[q + 1] || rb.b[q] % SCAN_TILE) SPHY_FAIL(ctx, SPHY_E_INVALID, "sphy_trunk_select: bounds must ascend on tile starts");
    }
    uint8_t *flag = nullptr;
    int32_t *sel = nullptr, *nsel = nullptr;
    void *tmp = nullptr;
    size_t tb = 0;
    int32_t m = 0;
    cudaError_t e = cudaSuccess;
#define FK(call) do { if (e == cudaSuccess) e = (call); } while (0)
    FK(cudaMalloc(&flag, (size_t)n));
    FK(cudaMalloc(&sel, sizeof(int32_t) * n));
    FK(cudaMalloc(&nsel, sizeof(int32_t)));
    if (e == cudaSuccess) {
        k_trunk_flag<<<div_up(n, 256), 256, 0, st>>>(ctx->sub_size, ctx->level, ctx->cell_canon, n, min_level, min_cells, rb, cut_cell_dev,
                                                     n_cut, flag);
        cub::CountingInputIterator<int32_t> counting(0);
        FK(cub::DeviceSelect::Flagged(nullptr, tb, counting, flag, sel, nsel, (int)n, st));
        FK(cudaMalloc(&tmp, tb));
        FK(cub::DeviceSelect::Flagged(tmp, tb, counting, flag, sel, nsel, (int)n, st));
        FK(cudaMemcpyAsync(&m, nsel, sizeof(int32_t), cudaMemcpyDeviceToHost, st));
        FK(cudaStreamSynchronize(st));
    }
    int rc = SPHY_OK;
    if (e == cudaSuccess) {
        const size_t ms = m > 0 ? (size_t)m : 1;
        if ((rc = dev_alloc(ctx, &ctx->sl_cell, ms)) || (rc = dev_alloc(ctx, &ctx->sl_sub, ms)) || (rc = dev_alloc(ctx, &ctx->sl_last, ms)) ||
            (rc = dev_alloc(ctx, &ctx->sl_trunk, ms)) || (rc = dev_alloc(ctx, &ctx->sl_tdon, ms * 8))) {
            cudaFree(flag); cudaFree(sel); cudaFree(nsel); cudaFree(tmp);
            return rc;
        }
        if (m > 0) {
            FK(cudaMemcpyAsync(ctx->sl_cell, sel, sizeof(int32_t) * m, cudaMemcpyDeviceToDevice, st));
            if (e == cudaSuccess)
                k_trunk_entries<<<div_up(m, 256), 256, 0, st>>>(ctx->sl_cell, m, ctx->sub_size, flag, ctx->don_ptr, ctx->don_idx,
                                                               ctx->sl_sub, ctx->sl_last, ctx->sl_trunk, ctx->sl_tdon);
            FK(cudaGetLastError());
            FK(cudaStreamSynchronize(st));
        }
    }
#undef FK
    cudaFree(flag); cudaFree(sel); cudaFree(nsel); cudaFree(tmp);
    if (e != cudaSuccess) SPHY_FAIL(ctx, SPHY_E_CUDA, "sphy_trunk_select: %s", cudaGetErrorString(e));
    ctx->sl_n = m;
    ctx->sl_nblk = div_up(m, TRUNK_BLOCK);
    ctx->sl_plan = false;
    ctx->sl_comp_cap = 0;
    ctx->sl_cuts = n_cut > 0;                     // such a list only serves sphy_route_reslake_begin / _finish
    *count_out = m;
    return SPHY_OK;
}

namespace {
__global__ void k_cell_donors(const int32_t *__restrict__ cells, int32_t n_cells, const int32_t *__restrict__ don_ptr,
                              const int32_t *__restrict__ don_idx, int32_t *__restrict__ out)
{
    const int32_t j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= n_cells) return;
    const int32_t c = cells[j];
    int k = 0;
    for (int32_t q = don_ptr[c]; q < don_ptr[c + 1] && k < 8; ++q) out[(int64_t)j * 8 + k++] = don_idx[q];
    for (; k < 8; ++k) out[(int64_t)j * 8 + k] = -1;
}
}  // namespace

extern "C" int sphy_cell_donors(sphy_ctx *ctx, int32_t n_cells, const int32_t *cells_dev, int32_t *donors_out_dev, void *stream)
{
    if (!ctx || n_cells < 0 || (n_cells > 0 && (!cells_dev || !donors_out_dev))) return SPHY_E_INVALID;
    if (!ctx->ldd_ready) SPHY_FAIL(ctx, SPHY_E_STATE, "sphy_cell_donors before sphy_ldd_build");
    if (ctx->partitioned) SPHY_FAIL(ctx, SPHY_E_STATE, "sphy_cell_donors after sphy_set_partition (it needs the whole-basin structures)");
    if (n_cells == 0) return SPHY_OK;
    k_cell_donors<<<div_up(n_cells, 128), 128, 0, (cudaStream_t)stream>>>(cells_dev, n_cells, ctx->don_ptr, ctx->don_idx, donors_out_dev);
    SPHY_LAUNCH_CHECK(ctx);
    return SPHY_OK;
}

extern "C" int sphy_trunk_export(sphy_ctx *ctx, const char *name, void *dst, size_t dst_bytes, void *stream)
{
    if (!ctx || !name || !dst) return SPHY_E_INVALID;
    const std::string k(name);
    const void *src = nullptr;
    size_t bytes = 0;
    const size_t m = (size_t)ctx->sl_n;
    if (k == "cell") { src = ctx->sl_cell; bytes = 4 * m; }
    else if (k == "sub_size") { src = ctx->sl_sub; bytes = 4 * m; }
    else if (k == "last") { src = ctx->sl_last; bytes = 4 * m; }
    else if (k == "donors") { src = ctx->sl_tdon; bytes = 32 * m; }
    else if (k == "is_trunk") { src = ctx->sl_trunk; bytes = m; }
    else SPHY_FAIL(ctx, SPHY_E_INVALID, "sphy_trunk_export: unknown array '%s'", name);
    if (dst_bytes < bytes) SPHY_FAIL(ctx, SPHY_E_INVALID, "sphy_trunk_export(%s): need %zu bytes, got %zu", name, bytes, dst_bytes);
    if (bytes) SPHY_CUDA(ctx, cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDeviceToDevice, (cudaStream_t)stream));
    return SPHY_OK;
}

extern "C" int sphy_trunk_set_plan(sphy_ctx *ctx, int32_t n_pub, const int32_t *pub_pos_dev, const int32_t *before_rank_dev,
                                   const int32_t *before_slot_dev, const int32_t *end_rank_dev, const int32_t *end_slot_dev,
                                   int32_t stride, int rank, int world, void *stream)
{
    if (!ctx || n_pub < 0 || (n_pub > 0 && !pub_pos_dev) || stride < 1 + n_pub || rank < 0 || rank >= world || world > 64)
        return SPHY_E_INVALID;
    const int32_t m = ctx->sl_n;
    if (m > 0 && (!before_rank_dev || !before_slot_dev || !end_rank_dev || !end_slot_dev)) return SPHY_E_INVALID;
    if (!ctx->ldd_ready) SPHY_FAIL(ctx, SPHY_E_STATE, "sphy_trunk_set_plan before sphy_ldd_build");
    if ((world > 1) != ctx->partitioned) SPHY_FAIL(ctx, SPHY_E_STATE, "sphy_trunk_set_plan: world %d on a%s context", world, ctx->partitioned ? " partitioned" : "n unpartitioned");
    cudaStream_t st = (cudaStream_t)stream;
    int rc;
    const size_t ms = m > 0 ? (size_t)m : 1;
    if ((rc = dev_alloc(ctx, &ctx->pub_e, (size_t)n_pub)) || (rc = dev_alloc(ctx, &ctx->sl_brank, ms)) || (rc = dev_alloc(ctx, &ctx->sl_bslot, ms)) ||
        (rc = dev_alloc(ctx, &ctx->sl_erank, ms)) || (rc = dev_alloc(ctx, &ctx->sl_eslot, ms)))
        return rc;
    if (n_pub) SPHY_CUDA(ctx, cudaMemcpyAsync(ctx->pub_e, pub_pos_dev, sizeof(int32_t) * n_pub, cudaMemcpyDeviceToDevice, st));
    if (m > 0) {
        SPHY_CUDA(ctx, cudaMemcpyAsync(ctx->sl_brank, before_rank_dev, sizeof(int32_t) * m, cudaMemcpyDeviceToDevice, st));
        SPHY_CUDA(ctx, cudaMemcpyAsync(ctx->sl_bslot, before_slot_dev, sizeof(int32_t) * m, cudaMemcpyDeviceToDevice, st));
        SPHY_CUDA(ctx, cudaMemcpyAsync(ctx->sl_erank, end_rank_dev, sizeof(int32_t) * m, cudaMemcpyDeviceToDevice, st));
        SPHY_CUDA(ctx, cudaMemcpyAsync(ctx->sl_eslot, end_slot_dev, sizeof(int32_t) * m, cudaMemcpyDeviceToDevice, st));
    }
    SPHY_CUDA(ctx, cudaStreamSynchronize(st));
    ctx->n_pub = n_pub;
    ctx->sl_stride = stride;
    ctx->sl_rank = rank;
    ctx->sl_world = world;
    ctx->sl_plan = true;
    ctx->sl_comp_cap = 0;
    ctx->scan_ready = false;                      // the tile-crossing list is rebuilt without the trunk cells
    ctx->scan_comp_cap = 0;
    return SPHY_OK;
}

extern "C" int sphy_route_scan_begin(sphy_ctx *ctx, int ncomp, const float *const *runoff_mm, float *const *q_state, sphy_param kx,
                                     float one_minus_kx, double *send_dev, int32_t stride, void *stream)
{
    if (!ctx || !runoff_mm || !q_state || ncomp < 1 || ncomp > ROUTE_MAX_COMP) return SPHY_E_INVALID;
    if (!send_dev && !ctx->p2p_peer_recv) SPHY_FAIL(ctx, SPHY_E_INVALID, "sphy_route_scan_begin: no send buffer and no peer exchange set up");
    if (!ctx->ldd_ready) SPHY_FAIL(ctx, SPHY_E_STATE, "sphy_route_scan_begin before sphy_ldd_build");
    if (!ctx->sl_plan) SPHY_FAIL(ctx, SPHY_E_STATE, "sphy_route_scan_begin before sphy_trunk_set_plan");
    if (stride != ctx->sl_stride) SPHY_FAIL(ctx, SPHY_E_INVALID, "sphy_route_scan_begin: stride %d != planned %d", stride, ctx->sl_stride);
    if (ctx->sl_cuts) SPHY_FAIL(ctx, SPHY_E_STATE, "the trunk list holds lake/reservoir cells: route with sphy_route_reslake_begin/_finish");
    cudaStream_t st = (cudaStream_t)stream;
    ScanArgs a;
    int rc = scan_prepare(ctx, ncomp, runoff_mm, q_state, kx, one_minus_kx, st, a);
    if (rc != SPHY_OK) return rc;
    if ((rc = trunk_prepare(ctx, ncomp, st)) != SPHY_OK) return rc;
    const bool push = send_dev == nullptr;                 // no host-side gather: the publishing blocks push into the peers' buffers
    if (push && (int64_t)ncomp * stride > ctx->p2p_pitch)
        SPHY_FAIL(ctx, SPHY_E_INVALID, "sphy_route_scan_begin: %lld values exceed the peer buffer pitch", (long long)ncomp * stride);
    a.send = send_dev;
    a.stride = stride;
    rc = scan_launch_local(ctx, a, st, ctx->sl_n > 0, push && ctx->sl_n > 0);
    if (rc != SPHY_OK) return rc;
    SPHY_LAUNCH_CHECK(ctx);
    return SPHY_OK;
}

extern "C" int sphy_route_scan_finish(sphy_ctx *ctx, int ncomp, float *const *q_state, sphy_param kx, float one_minus_kx,
                                      const double *gathered_dev, int32_t stride, int rank, int world, void *stream)
{
    if (!ctx || !q_state || ncomp < 1 || ncomp > ROUTE_MAX_COMP || rank < 0 || rank >= world) return SPHY_E_INVALID;
    if (!gathered_dev && !ctx->p2p_recv) SPHY_FAIL(ctx, SPHY_E_INVALID, "sphy_route_scan_finish: no gathered buffer and no peer exchange set up");
    if (!ctx->scan_ready || !ctx->sl_plan) SPHY_FAIL(ctx, SPHY_E_STATE, "sphy_route_scan_finish before sphy_route_scan_begin");
    if (stride != ctx->sl_stride || rank != ctx->sl_rank || world != ctx->sl_world)
        SPHY_FAIL(ctx, SPHY_E_INVALID, "sphy_route_scan_finish: stride/rank/world differ from the plan");
    if (ctx->sl_n == 0) return SPHY_OK;
    cudaStream_t st = (cudaStream_t)stream;
    ScanArgs a;
    int rc = scan_prepare(ctx, ncomp, nullptr, q_state, kx, one_minus_kx, st, a);
    if (rc != SPHY_OK) return rc;
    // gathered_dev == NULL: what the peers stored into this rank's buffer (the trunk kernel acquires their flags first)
    const bool p2p = gathered_dev == nullptr;
    if ((rc = trunk_finish(ctx, a, gathered_dev, p2p ? ctx->p2p_pitch : (int64_t)ncomp * stride, st, p2p)) != SPHY_OK) return rc;
    prof_mark(ctx, st);
    SPHY_LAUNCH_CHECK(ctx);
    return SPHY_OK;
}

// ---- lakes / reservoirs on the trunk list (a partitioned basin, or one GPU) ---------------------------------------------------
namespace {

// what the replicated structure update needs from this rank's cells: the day's discharge at the donors of every structure
// (of the PREVIOUS step: the main pass is about to overwrite qday) and TotR at the structure cells; 0 for cells of other ranks
__global__ void k_rl_part_gather(int32_t n_extra, const int32_t *__restrict__ x_rank, const int32_t *__restrict__ x_local, int rank,
                                 const float *__restrict__ qday, const float *__restrict__ totr, double *__restrict__ extra)
{
    const int32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_extra) return;
    double v = 0.0;
    if (x_rank[i] == rank) v = (double)((i % SPHY_RL_EXTRA) == SPHY_RL_EXTRA - 1 ? totr[x_local[i]] : qday[x_local[i]]);
    extra[i] = v;
}

int reslake_plan_check(sphy_ctx *ctx, const sphy_reslake *rl, const sphy_reslake_plan *plan, int32_t stride, const char *who)
{
    if (!ctx->ldd_ready) SPHY_FAIL(ctx, SPHY_E_STATE, "%s before sphy_ldd_build", who);
    if (ctx->order_mode != SPHY_ORDER_DFS) SPHY_FAIL(ctx, SPHY_E_STATE, "%s needs SPHY_ORDER_DFS", who);
    if (!ctx->sl_plan) SPHY_FAIL(ctx, SPHY_E_STATE, "%s before sphy_trunk_set_plan", who);
    if (rl->n_struct < 0 || !rl->qfrac || (rl->n_struct > 0 && (!rl->kind || !rl->par || !rl->StorRES || !rl->Qout || !rl->Qin)))
        SPHY_FAIL(ctx, SPHY_E_INVALID, "%s: incomplete structure table", who);
    if (rl->n_struct > 0 && !ctx->sl_cuts) SPHY_FAIL(ctx, SPHY_E_STATE, "%s: the trunk list was selected without the structure cells (sphy_trunk_select_cuts)", who);
    if (!plan->qmask || !plan->qday || !plan->top || !plan->entry_cut || !plan->entry_ptr ||
        (rl->n_struct > 0 && (!plan->cut_slot || !plan->entry_idx || !plan->x_rank || !plan->x_local || !plan->extra)))
        SPHY_FAIL(ctx, SPHY_E_INVALID, "%s: incomplete plan", who);
    if (stride != ctx->sl_stride) SPHY_FAIL(ctx, SPHY_E_INVALID, "%s: stride %d != planned %d", who, stride, ctx->sl_stride);
    if ((int64_t)stride < 1 + (int64_t)ctx->n_pub + (int64_t)SPHY_RL_EXTRA * rl->n_struct)
        SPHY_FAIL(ctx, SPHY_E_INVALID, "%s: stride %d leaves no room for %d structures behind %d prefixes", who, stride, rl->n_struct, ctx->n_pub);
    return SPHY_OK;
}

}  // namespace

extern "C" int sphy_route_reslake_begin(sphy_ctx *ctx, const sphy_reslake *rl, const sphy_reslake_plan *plan, const float *totr, float *q_state,
                                        sphy_param kx, float one_minus_kx, double *send_dev, int32_t stride, void *stream)
{
    if (!ctx || !rl || !plan || !totr || !q_state) return SPHY_E_INVALID;
    int rc = reslake_plan_check(ctx, rl, plan, stride, "sphy_route_reslake_begin");
    if (rc != SPHY_OK) return rc;
    if (!send_dev && ctx->partitioned && !ctx->p2p_peer_recv)
        SPHY_FAIL(ctx, SPHY_E_INVALID, "sphy_route_reslake_begin: no send buffer and no peer exchange set up");
    cudaStream_t st = (cudaStream_t)stream;
    const int32_t n_extra = SPHY_RL_EXTRA * rl->n_struct;
    if (n_extra > 0)
        k_rl_part_gather<<<div_up(n_extra, 128), 128, 0, st>>>(n_extra, plan->x_rank, plan->x_local, ctx->sl_rank, plan->qday, totr, plan->extra);
    SPHY_LAUNCH_CHECK(ctx);
    // The main pass forms the material itself, 0.001 * cellarea * TotR as a REAL4 product (advanced_routing.py:137,159), for every
    // cell -- structure cells included: whatever they add to a range sum is taken out again with E(k) of the structure, which
    // holds the same terms (every cell downstream of a structure subtracts E(k) - Qout(k)).
    ScanArgs a;
    const float *mm[1] = {totr};
    float *qq[1] = {q_state};
    if ((rc = scan_prepare(ctx, 1, mm, qq, kx, one_minus_kx, st, a)) != SPHY_OK) return rc;
    if ((rc = trunk_prepare(ctx, 1, st)) != SPHY_OK) return rc;
    if (ctx->zero_ptr_cap < ctx->n_tiles + 1) {                // the main pass of a lake network walks a per-tile cut list: an empty one
        if ((rc = dev_alloc(ctx, &ctx->zero_ptr, (size_t)ctx->n_tiles + 1)) != SPHY_OK) return rc;
        SPHY_CUDA(ctx, cudaMemsetAsync(ctx->zero_ptr, 0, sizeof(int32_t) * ((size_t)ctx->n_tiles + 1), st));
        ctx->zero_ptr_cap = ctx->n_tiles + 1;
    }
    const bool push = send_dev == nullptr && ctx->partitioned;  // the publishing blocks store into the peers' buffers
    if (push && (int64_t)stride > ctx->p2p_pitch)
        SPHY_FAIL(ctx, SPHY_E_INVALID, "sphy_route_reslake_begin: %d values exceed the peer buffer pitch", stride);
    a.adv = 1;
    a.volf = 0.001f * ctx->cellarea;
    a.frac = plan->qmask;
    a.qout_dense = plan->qmask;                                // never read: no cell of the mask is 0
    a.qday = plan->qday;
    a.cut_ptr = ctx->zero_ptr;
    a.extra = plan->extra;
    a.n_extra = n_extra;
    a.send = send_dev ? send_dev : (ctx->partitioned ? nullptr : ctx->sl_V);
    a.stride = stride;
    if ((rc = scan_launch_local(ctx, a, st, ctx->sl_n > 0, push && ctx->sl_n > 0)) != SPHY_OK) return rc;
    SPHY_LAUNCH_CHECK(ctx);
    return SPHY_OK;
}

extern "C" int sphy_route_reslake_finish(sphy_ctx *ctx, const sphy_reslake *rl, const sphy_reslake_plan *plan, float *q_state, sphy_param kx,
                                         float one_minus_kx, const double *gathered_dev, int32_t stride, int rank, int world, int day_of_year,
                                         int apply_pending, void *stream)
{
    if (!ctx || !rl || !plan || !q_state || rank < 0 || rank >= world) return SPHY_E_INVALID;
    int rc = reslake_plan_check(ctx, rl, plan, stride, "sphy_route_reslake_finish");
    if (rc != SPHY_OK) return rc;
    if (!ctx->scan_ready) SPHY_FAIL(ctx, SPHY_E_STATE, "sphy_route_reslake_finish before sphy_route_reslake_begin");
    if (rank != ctx->sl_rank || world != ctx->sl_world) SPHY_FAIL(ctx, SPHY_E_INVALID, "sphy_route_reslake_finish: rank/world differ from the plan");
    if (!gathered_dev && ctx->partitioned && !ctx->p2p_recv)
        SPHY_FAIL(ctx, SPHY_E_INVALID, "sphy_route_reslake_finish: no gathered buffer and no peer exchange set up");
    if (ctx->sl_n == 0) return SPHY_OK;
    cudaStream_t st = (cudaStream_t)stream;
    ScanArgs a;
    float *qq[1] = {q_state};
    if ((rc = scan_prepare(ctx, 1, nullptr, qq, kx, one_minus_kx, st, a)) != SPHY_OK) return rc;
    const bool p2p = gathered_dev == nullptr && ctx->partitioned;
    const double *gathered = gathered_dev ? gathered_dev : (p2p ? nullptr : ctx->sl_V);
    const int64_t rank_pitch = p2p ? ctx->p2p_pitch : (int64_t)stride;
    TrunkArgs t = trunk_args(ctx, gathered, rank_pitch, p2p);
    t.H = plan->top;
    t.cut_slot = plan->cut_slot;
    t.entry_cut = plan->entry_cut;
    t.entry_ptr = plan->entry_ptr;
    t.entry_idx = plan->entry_idx;
    t.qout = rl->Qout;
    t.qday = plan->qday;
    k_trunk_drift<<<ctx->sl_nblk, TRUNK_BLOCK, 0, st>>>(a, t);    // waits for the peers; E (and its noise scale) of every listed cell
    if (rl->n_struct > 0) {
        RlPartUpdate u = {};
        u.gathered = gathered;
        u.rank_pitch = rank_pitch;
        u.stride = stride;
        u.n_extra = SPHY_RL_EXTRA * rl->n_struct;
        u.x_rank = plan->x_rank;
        u.apply_pending = apply_pending;
        u.vol_factor = 0.001f * ctx->cellarea;
        u.doy = day_of_year;
        if (p2p) {
            u.p2p_local = ctx->p2p_flags;
            u.p2p_world = ctx->p2p_world;
            u.p2p_recv = ctx->p2p_recv;
            u.p2p_pitch = ctx->p2p_pitch;
        }
        if ((rc = reslake_part_update(ctx, rl, u, st)) != SPHY_OK) return rc;
    }
    k_trunk_apply_adv<<<div_up(ctx->sl_n, TAIL_THREADS), TAIL_THREADS, 0, st>>>(a, t);
    prof_mark(ctx, st);
    SPHY_LAUNCH_CHECK(ctx);
    return SPHY_OK;
}

// ---- confluence exchange over peer memory: set-up (the per-step part is in k_scan_cross_pub / k_trunk_drift) ---------------
extern "C" int sphy_p2p_setup_local(sphy_ctx *ctx, int world, int64_t pitch, unsigned char *handle_recv64, unsigned char *handle_flags64)
{
    if (!ctx || world < 2 || world > 64 || pitch < 1 || !handle_recv64 || !handle_flags64) return SPHY_E_INVALID;
    if (ctx->p2p_recv) SPHY_FAIL(ctx, SPHY_E_STATE, "sphy_p2p_setup_local: already set up");
    static_assert(sizeof(cudaIpcMemHandle_t) == 64, "CUDA IPC handle size");
    const size_t nrecv = (size_t)2 * world * pitch;
    SPHY_CUDA(ctx, cudaMalloc(&ctx->p2p_recv, nrecv * sizeof(double)));
    SPHY_CUDA(ctx, cudaMalloc(&ctx->p2p_flags, (size_t)(world + 3) * sizeof(long long)));   // + time-out word, step counter, block counter
    SPHY_CUDA(ctx, cudaMemset(ctx->p2p_recv, 0, nrecv * sizeof(double)));
    SPHY_CUDA(ctx, cudaMemset(ctx->p2p_flags, 0, (size_t)(world + 3) * sizeof(long long)));
    SPHY_CUDA(ctx, cudaDeviceSynchronize());
    cudaIpcMemHandle_t h;
    SPHY_CUDA(ctx, cudaIpcGetMemHandle(&h, ctx->p2p_recv));
    memcpy(handle_recv64, &h, 64);
    SPHY_CUDA(ctx, cudaIpcGetMemHandle(&h, ctx->p2p_flags));
    memcpy(handle_flags64, &h, 64);
    ctx->p2p_world = world;
    ctx->p2p_pitch = pitch;
    return SPHY_OK;
}

extern "C" int sphy_p2p_connect(sphy_ctx *ctx, int rank, int world, const unsigned char *handles_recv, const unsigned char *handles_flags)
{
    if (!ctx || !handles_recv || !handles_flags || rank < 0 || rank >= world) return SPHY_E_INVALID;
    if (!ctx->p2p_recv || world != ctx->p2p_world) SPHY_FAIL(ctx, SPHY_E_STATE, "sphy_p2p_connect before sphy_p2p_setup_local");
    double *recv[64];
    long long *flags[64];
    for (int p = 0; p < world; ++p) {
        if (p == rank) {
            recv[p] = ctx->p2p_recv;
            flags[p] = ctx->p2p_flags;
            continue;
        }
        cudaIpcMemHandle_t h;
        void *ptr = nullptr;
        memcpy(&h, handles_recv + (size_t)p * 64, 64);
        SPHY_CUDA(ctx, cudaIpcOpenMemHandle(&ptr, h, cudaIpcMemLazyEnablePeerAccess));
        recv[p] = (double *)ptr;
        memcpy(&h, handles_flags + (size_t)p * 64, 64);
        SPHY_CUDA(ctx, cudaIpcOpenMemHandle(&ptr, h, cudaIpcMemLazyEnablePeerAccess));
        flags[p] = (long long *)ptr;
        ctx->p2p_opened[ctx->p2p_nopened++] = recv[p];
        ctx->p2p_opened[ctx->p2p_nopened++] = flags[p];
    }
    SPHY_CUDA(ctx, cudaMalloc(&ctx->p2p_peer_recv, sizeof(double *) * world));
    SPHY_CUDA(ctx, cudaMalloc(&ctx->p2p_peer_flags, sizeof(long long *) * world));
    SPHY_CUDA(ctx, cudaMemcpy(ctx->p2p_peer_recv, recv, sizeof(double *) * world, cudaMemcpyHostToDevice));
    SPHY_CUDA(ctx, cudaMemcpy(ctx->p2p_peer_flags, flags, sizeof(long long *) * world, cudaMemcpyHostToDevice));
    ctx->p2p_rank = rank;
    return SPHY_OK;
}

// 1 if a bounded flag wait of the peer-memory exchange gave up since the last call (synchronises the stream)
extern "C" int sphy_p2p_timed_out(sphy_ctx *ctx, void *stream)
{
    if (!ctx || !ctx->p2p_flags) return 0;
    long long w = 0;
    if (cudaMemcpyAsync(&w, ctx->p2p_flags + ctx->p2p_world, sizeof(w), cudaMemcpyDeviceToHost, (cudaStream_t)stream) != cudaSuccess) return 1;
    if (cudaStreamSynchronize((cudaStream_t)stream) != cudaSuccess) return 1;
    return w != 0;
}
