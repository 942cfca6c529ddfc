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
    double *tile_excl;                    // [ncomp][ntiles+1] exclusive prefix of the totals
    // multi-GPU (partitioned context)
    int32_t n_pub, n_remote;
    const int32_t *pub_e, *remote_k, *remote_rank, *remote_slot;
    double *send;                         // [ncomp][stride]: range total, then range-local inclusive prefix at pub_e[j]
    const double *gathered;               // [world][ncomp][stride]
    int32_t stride, rank, world;
    int64_t rank_pitch;                   // elements between two ranks' blocks in `gathered`
    // lakes / reservoirs (advanced_routing.ROUT): material is already m3/day, cut cells take the structure outflow
    int adv;
    const float *frac;                    // dense QFRAC (0 at lake / reservoir cells)
    const float *qout_dense;              // dense structure outflow, m3/day
    float *qday;                          // dense blended Q, m3/day (what upstream(ldd, Q) reads at the structures)
    const int32_t *cut_cell, *cut_ptr;    // cut cells (ascending) and their per-tile ranges
    const double *cut_corr;               // float64 material placed at each cut cell (see route_scan_run_adv)
    sphy_param kx;
    float one_m_kx;
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
                        if (KXMAP) kxv[b] = __ldg(reinterpret_cast<const float4 *>(a.kx.map + tbase + o));
                    } else {
                        sz[b].x = o < len ? __ldg(szp + o) : 1;         q[b].x = o < len ? qs[o] : 0.0f;
                        sz[b].y = o + 1 < len ? __ldg(szp + o + 1) : 1; q[b].y = o + 1 < len ? qs[o + 1] : 0.0f;
                        sz[b].z = o + 2 < len ? __ldg(szp + o + 2) : 1; q[b].z = o + 2 < len ? qs[o + 2] : 0.0f;
                        sz[b].w = o + 3 < len ? __ldg(szp + o + 3) : 1; q[b].w = o + 3 < len ? qs[o + 3] : 0.0f;
                        if (KXMAP) {
                            const float *kp = a.kx.map + tbase;
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
                    v0 = (double)x[i].x; v1 = (double)x[i].y; v2 = (double)x[i].z; v3 = (double)x[i].w;
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
                        const float kx = KXMAP ? kxj : a.kx.value;
                        const float omk = KXMAP ? 1.0f - kx : a.one_m_kx;
                        double pe = own;                               // headwater cell: the range is the cell itself
                        if (inside && szj > 1) pe = Ls[ls_slot(el)];   // predicated load: fewer lanes, fewer bank conflicts
                        if (ADV) {
                            if (inside) qj = blend(a, tbase + o + j, (float)range_diff(pe, before), qj, kx, omk);
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

__global__ void __launch_bounds__(256) k_scan_cross(const __grid_constant__ ScanArgs a)
{
    const int32_t k0 = blockIdx.x * (256 * CROSS_ILP) + threadIdx.x;
    int32_t r[CROSS_ILP], e[CROSS_ILP];
    bool ok[CROSS_ILP];
#pragma unroll
    for (int j = 0; j < CROSS_ILP; ++j) {
        const int32_t k = k0 + j * 256;
        ok[j] = k < a.n_cross;
        r[j] = ok[j] ? __ldg(a.cross_r + k) : 0;
        e[j] = ok[j] ? __ldg(a.cross_e + k) : 0;
    }
    for (int c = 0; c < a.ncomp; ++c) {
        const double *T = a.tile_tot + (int64_t)c * (a.ntiles + 1);
        const double *X = a.tile_excl + (int64_t)c * (a.ntiles + 1);
        double *PB = a.park_before + (int64_t)c * a.n_cross;
        const double *PE = a.park_end + (int64_t)c * (a.n_cross + a.n_pub);
        float *qs = a.qstate[c];
        double tt[CROSS_ILP], pb[CROSS_ILP], pe[CROSS_ILP];
        float qold[CROSS_ILP], kx[CROSS_ILP];
#pragma unroll
        for (int j = 0; j < CROSS_ILP; ++j) {                  // all gathers first
            const int32_t k = k0 + j * 256;
            const bool local = ok[j] && e[j] < a.n;
            tt[j] = ok[j] ? T[r[j] / SCAN_TILE] : 0.0;
            pb[j] = ok[j] ? PB[k] : 0.0;
            pe[j] = local ? PE[k] : 0.0;
            qold[j] = local ? qs[r[j]] : 0.0f;
            kx[j] = (ok[j] && a.kx.map) ? __ldg(a.kx.map + r[j]) : a.kx.value;
        }
#pragma unroll
        for (int j = 0; j < CROSS_ILP; ++j) {
            if (!ok[j]) continue;
            const int32_t k = k0 + j * 256;
            const int32_t ta = r[j] / SCAN_TILE, te = e[j] / SCAN_TILE;
            if (e[j] >= a.n) {                                 // range ends on another rank: keep the part up to the end
                PB[k] = range_diff(tt[j], pb[j]) +             // of this rank's range for k_scan_remote
                        (ta + 1 < a.ntiles ? range_diff(X[a.ntiles], X[ta + 1]) : 0.0);
                continue;
            }
            // suffix of the first tile + whole tiles in between (none for most crossing cells) + prefix of the last tile
            const double s = range_diff(tt[j], pb[j]) + (te > ta + 1 ? range_diff(X[te], X[ta + 1]) : 0.0) + pe[j];
            const float omk = a.kx.map ? 1.0f - kx[j] : a.one_m_kx;
            qs[r[j]] = blend(a, r[j], (float)s, qold[j], kx[j], omk);
        }
    }
}

// multi-GPU: what the other ranks need from this one -- the range total and the range-local inclusive prefix at
// every position one of their trunk cells' upstream range ends at
__global__ void k_scan_publish(const __grid_constant__ ScanArgs a)
{
    const int32_t j = blockIdx.x * blockDim.x + threadIdx.x;
    for (int c = 0; c < a.ncomp; ++c) {
        const double *X = a.tile_excl + (int64_t)c * (a.ntiles + 1);
        double *out = a.send + (int64_t)c * a.stride;
        if (j == 0) out[0] = X[a.ntiles];
        if (j < a.n_pub) out[1 + j] = X[a.pub_e[j] / SCAN_TILE] + a.park_end[(int64_t)c * (a.n_cross + a.n_pub) + a.n_cross + j];
    }
}

// multi-GPU: finish the cells whose upstream range ends on another rank, after the all-gather
__global__ void k_scan_remote(const __grid_constant__ ScanArgs a)
{
    const int32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= a.n_remote) return;
    const int32_t k = a.remote_k[i], q = a.remote_rank[i], slot = a.remote_slot[i];
    const int64_t r = a.cross_r[k];
    const float kx = a.kx.map ? __ldg(a.kx.map + r) : a.kx.value;
    const float omk = a.kx.map ? 1.0f - kx : a.one_m_kx;
    for (int c = 0; c < a.ncomp; ++c) {
        double s = a.park_before[(int64_t)c * a.n_cross + k];                     // r .. end of this rank's range
        for (int u = a.rank + 1; u < q; ++u) s += a.gathered[(int64_t)u * a.rank_pitch + (int64_t)c * a.stride];   // whole ranks in between
        s += a.gathered[(int64_t)q * a.rank_pitch + (int64_t)c * a.stride + 1 + slot];        // start of rank q's range .. range end
        float *qs = a.qstate[c];
        qs[r] = blend(a, r, (float)s, qs[r], kx, omk);
    }
}

// ---- one-time construction of the crossing lists ---------------------------------------------------------
__global__ void k_cross_flag(const int32_t *sub_size, int64_t n, uint8_t *flag)
{
    int64_t r = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (r < n) flag[r] = (r / SCAN_TILE) != ((r + sub_size[r] - 1) / SCAN_TILE);
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
    a.n_remote = ctx->n_remote;
    a.pub_e = ctx->pub_e;
    a.remote_k = ctx->remote_k;
    a.remote_rank = ctx->remote_rank;
    a.remote_slot = ctx->remote_slot;
    a.kx = kx;
    a.one_m_kx = one_m_kx;
    a.to_m3s = (0.001 * (double)ctx->cellarea) / 86400.0;
    a.n = ctx->n;
    a.ntiles = ctx->n_tiles;
    return SPHY_OK;
}

// main pass -> exclusive scan of the tile totals (cub, one call per component) -> crossing cells
static int scan_launch_local(sphy_ctx *ctx, ScanArgs &a, cudaStream_t st)
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
    if (a.adv) {
        if (a.kx.map) k_scan_main<true, true><<<grid, SCAN_THREADS, SCAN_SMEM, st>>>(a);
        else k_scan_main<true, false><<<grid, SCAN_THREADS, SCAN_SMEM, st>>>(a);
    } else {
        if (a.kx.map) k_scan_main<false, true><<<grid, SCAN_THREADS, SCAN_SMEM, st>>>(a);
        else k_scan_main<false, false><<<grid, SCAN_THREADS, SCAN_SMEM, st>>>(a);
    }
    for (int c = 0; c < a.ncomp; ++c) {
        size_t need = ctx->scan_tmp_bytes;
        SPHY_CUDA(ctx, cub::DeviceScan::ExclusiveSum(ctx->scan_tmp, need, ctx->scan_tile + (size_t)c * (ctx->n_tiles + 1),
                                                     ctx->scan_excl + (size_t)c * (ctx->n_tiles + 1), ctx->n_tiles + 1, st));
    }
    if (ctx->n_cross > 0) k_scan_cross<<<div_up(ctx->n_cross, 256 * CROSS_ILP), 256, 0, st>>>(a);
    return SPHY_OK;
}

int route_scan_run(sphy_ctx *ctx, int ncomp, const float *const *runoff_mm, float *const *q_state, sphy_param kx,
                   float one_m_kx, cudaStream_t st)
{
    ScanArgs a;
    int rc = scan_prepare(ctx, ncomp, runoff_mm, q_state, kx, one_m_kx, st, a);
    if (rc != SPHY_OK) return rc;
    int rc2 = scan_launch_local(ctx, a, st);
    if (rc2 != SPHY_OK) return rc2;
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

// ---- multi-GPU -------------------------------------------------------------------------------------------------
extern "C" int sphy_scan_tile_cells(void) { return SCAN_TILE; }

extern "C" int sphy_set_partition(sphy_ctx *ctx, int64_t cell_begin, int64_t cell_end, int32_t n_pub, const int32_t *pub_e_dev,
                                  int32_t n_remote, const int32_t *remote_k_dev, const int32_t *remote_rank_dev,
                                  const int32_t *remote_slot_dev, void *stream)
{
    if (!ctx) return SPHY_E_INVALID;
    if (!ctx->ldd_ready) SPHY_FAIL(ctx, SPHY_E_STATE, "sphy_set_partition before sphy_ldd_build");
    if (ctx->order_mode != SPHY_ORDER_DFS) SPHY_FAIL(ctx, SPHY_E_STATE, "sphy_set_partition needs SPHY_ORDER_DFS");
    if (ctx->partitioned) SPHY_FAIL(ctx, SPHY_E_STATE, "context is already partitioned");
    if (cell_begin < 0 || cell_end > ctx->n_global || cell_begin >= cell_end || (cell_begin % SCAN_TILE) != 0 || n_pub < 0 ||
        n_remote < 0 || (n_pub > 0 && !pub_e_dev) || (n_remote > 0 && (!remote_k_dev || !remote_rank_dev || !remote_slot_dev)))
        SPHY_FAIL(ctx, SPHY_E_INVALID, "sphy_set_partition: bad range or lists (the range must start on a %d-cell tile)", SCAN_TILE);
    cudaStream_t st = (cudaStream_t)stream;
    int rc;
    if ((rc = dev_alloc(ctx, &ctx->pub_e, (size_t)n_pub)) || (rc = dev_alloc(ctx, &ctx->remote_k, (size_t)n_remote)) ||
        (rc = dev_alloc(ctx, &ctx->remote_rank, (size_t)n_remote)) || (rc = dev_alloc(ctx, &ctx->remote_slot, (size_t)n_remote)))
        return rc;
    if (n_pub) SPHY_CUDA(ctx, cudaMemcpyAsync(ctx->pub_e, pub_e_dev, sizeof(int32_t) * n_pub, cudaMemcpyDeviceToDevice, st));
    if (n_remote) {
        SPHY_CUDA(ctx, cudaMemcpyAsync(ctx->remote_k, remote_k_dev, sizeof(int32_t) * n_remote, cudaMemcpyDeviceToDevice, st));
        SPHY_CUDA(ctx, cudaMemcpyAsync(ctx->remote_rank, remote_rank_dev, sizeof(int32_t) * n_remote, cudaMemcpyDeviceToDevice, st));
        SPHY_CUDA(ctx, cudaMemcpyAsync(ctx->remote_slot, remote_slot_dev, sizeof(int32_t) * n_remote, cudaMemcpyDeviceToDevice, st));
    }
    SPHY_CUDA(ctx, cudaStreamSynchronize(st));
    ctx->n_pub = n_pub;
    ctx->n_remote = n_remote;
    ctx->part_begin = cell_begin;
    ctx->n = cell_end - cell_begin;
    ctx->sub_size_loc = ctx->sub_size + cell_begin;
    ctx->flat_loc = ctx->flat_active + cell_begin;
    ctx->partitioned = true;
    ctx->scan_ready = false;
    ctx->scan_comp_cap = 0;
    return SPHY_OK;
}

extern "C" int sphy_route_scan_begin(sphy_ctx *ctx, int ncomp, const float *const *runoff_mm, float *const *q_state, sphy_param kx,
                                     float one_minus_kx, double *send_dev, int32_t stride, void *stream)
{
    if (!ctx || !runoff_mm || !q_state || !send_dev || ncomp < 1 || ncomp > ROUTE_MAX_COMP) return SPHY_E_INVALID;
    if (!ctx->ldd_ready) SPHY_FAIL(ctx, SPHY_E_STATE, "sphy_route_scan_begin before sphy_ldd_build");
    if (stride < 1 + ctx->n_pub) SPHY_FAIL(ctx, SPHY_E_INVALID, "sphy_route_scan_begin: stride %d < 1 + n_pub %d", stride, ctx->n_pub);
    cudaStream_t st = (cudaStream_t)stream;
    ScanArgs a;
    int rc = scan_prepare(ctx, ncomp, runoff_mm, q_state, kx, one_minus_kx, st, a);
    if (rc != SPHY_OK) return rc;
    a.send = send_dev;
    a.stride = stride;
    rc = scan_launch_local(ctx, a, st);
    if (rc != SPHY_OK) return rc;
    k_scan_publish<<<div_up(ctx->n_pub > 0 ? ctx->n_pub : 1, 256), 256, 0, st>>>(a);
    SPHY_LAUNCH_CHECK(ctx);
    return SPHY_OK;
}

extern "C" int sphy_route_scan_finish(sphy_ctx *ctx, int ncomp, float *const *q_state, sphy_param kx, float one_minus_kx,
                                      const double *gathered_dev, int32_t stride, int rank, int world, void *stream)
{
    if (!ctx || !q_state || ncomp < 1 || ncomp > ROUTE_MAX_COMP || rank < 0 || rank >= world) return SPHY_E_INVALID;
    if (!gathered_dev && !ctx->p2p_recv) SPHY_FAIL(ctx, SPHY_E_INVALID, "sphy_route_scan_finish: no gathered buffer and no peer exchange set up");
    if (!ctx->scan_ready) SPHY_FAIL(ctx, SPHY_E_STATE, "sphy_route_scan_finish before sphy_route_scan_begin");
    if (ctx->n_remote == 0) return SPHY_OK;
    cudaStream_t st = (cudaStream_t)stream;
    ScanArgs a;
    int rc = scan_prepare(ctx, ncomp, nullptr, q_state, kx, one_minus_kx, st, a);
    if (rc != SPHY_OK) return rc;
    if (gathered_dev) {
        a.gathered = gathered_dev;
        a.rank_pitch = (int64_t)ncomp * stride;
    } else {                                      // what the peers stored into this rank's buffer (sphy_route_scan_exchange_p2p)
        const int parity = (int)((ctx->p2p_step - 1) & 1);
        a.gathered = ctx->p2p_recv + (size_t)parity * ctx->p2p_world * ctx->p2p_pitch;
        a.rank_pitch = ctx->p2p_pitch;
    }
    a.stride = stride;
    a.rank = rank;
    a.world = world;
    k_scan_remote<<<div_up(ctx->n_remote, 256), 256, 0, st>>>(a);
    SPHY_LAUNCH_CHECK(ctx);
    return SPHY_OK;
}

// ---- confluence exchange over peer memory (NVLink P2P) instead of an NCCL all-gather -------------------------------------
// The message is a few thousand float64 per rank and purely latency-bound.  Every rank owns a receive buffer
// [2 parities][world][pitch] and a flag per peer, allocated with cudaMalloc and opened on the other ranks through CUDA
// IPC.  Per step: k_p2p_push (one CTA per peer) stores this rank's block straight into every peer's buffer and
// releases `flag[rank] = step` with system scope; k_p2p_wait acquires the flags of all peers before the trunk cells
// are finished.  Double buffering by step parity is enough: a rank can only be one exchange ahead of a peer, because
// its next push comes after a wait on that peer's current flag.  The wait is bounded (no hang if a peer dies).
namespace {

__global__ void __launch_bounds__(256) k_p2p_push(const double *__restrict__ send, int count, double *const *peer_recv,
                                                   long long *const *peer_flags, int rank, int world, int64_t pitch, int parity,
                                                   long long step)
{
    const int p = blockIdx.x;
    double *dst = peer_recv[p] + ((size_t)parity * world + rank) * pitch;
    for (int i = threadIdx.x; i < count; i += blockDim.x) dst[i] = send[i];
    __threadfence_system();
    __syncthreads();
    if (threadIdx.x == 0) asm volatile("st.release.sys.global.s64 [%0], %1;" ::"l"(peer_flags[p] + rank), "l"(step) : "memory");
}

__global__ void k_p2p_wait(const long long *flags, int world, long long step, int *timed_out)
{
    const int q = threadIdx.x;
    if (q >= world) return;
    long long v = 0;
    for (long long it = 0; it < 40000000LL; ++it) {             // ~ 4 s of 100 ns naps, then give up
        asm volatile("ld.acquire.sys.global.s64 %0, [%1];" : "=l"(v) : "l"(flags + q) : "memory");
        if (v >= step) return;
        __nanosleep(100);
    }
    *timed_out = 1;
}

}  // namespace

extern "C" int sphy_p2p_setup_local(sphy_ctx *ctx, int world, int64_t pitch, unsigned char *handle_recv64, unsigned char *handle_flags64)
{
    if (!ctx || world < 2 || world > 64 || pitch < 1 || !handle_recv64 || !handle_flags64) return SPHY_E_INVALID;
    if (ctx->p2p_recv) SPHY_FAIL(ctx, SPHY_E_STATE, "sphy_p2p_setup_local: already set up");
    static_assert(sizeof(cudaIpcMemHandle_t) == 64, "CUDA IPC handle size");
    const size_t nrecv = (size_t)2 * world * pitch;
    SPHY_CUDA(ctx, cudaMalloc(&ctx->p2p_recv, nrecv * sizeof(double)));
    SPHY_CUDA(ctx, cudaMalloc(&ctx->p2p_flags, (size_t)(world + 1) * sizeof(long long)));   // + the time-out word
    SPHY_CUDA(ctx, cudaMemset(ctx->p2p_recv, 0, nrecv * sizeof(double)));
    SPHY_CUDA(ctx, cudaMemset(ctx->p2p_flags, 0, (size_t)(world + 1) * sizeof(long long)));
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
    ctx->p2p_step = 0;
    return SPHY_OK;
}

extern "C" int sphy_route_scan_exchange_p2p(sphy_ctx *ctx, int ncomp, const double *send_dev, int32_t stride, void *stream)
{
    if (!ctx || !send_dev || ncomp < 1 || ncomp > ROUTE_MAX_COMP) return SPHY_E_INVALID;
    if (!ctx->p2p_peer_recv) SPHY_FAIL(ctx, SPHY_E_STATE, "sphy_route_scan_exchange_p2p before sphy_p2p_connect");
    const int64_t count = (int64_t)ncomp * stride;
    if (count > ctx->p2p_pitch) SPHY_FAIL(ctx, SPHY_E_INVALID, "sphy_route_scan_exchange_p2p: %lld values exceed the buffer pitch", (long long)count);
    cudaStream_t st = (cudaStream_t)stream;
    const long long step = ++ctx->p2p_step;
    const int parity = (int)((step - 1) & 1);
    k_p2p_push<<<ctx->p2p_world, 256, 0, st>>>(send_dev, (int)count, ctx->p2p_peer_recv, ctx->p2p_peer_flags, ctx->p2p_rank,
                                               ctx->p2p_world, ctx->p2p_pitch, parity, step);
    k_p2p_wait<<<1, 64, 0, st>>>(ctx->p2p_flags, ctx->p2p_world, step, reinterpret_cast<int *>(ctx->p2p_flags + ctx->p2p_world));
    SPHY_LAUNCH_CHECK(ctx);
    return SPHY_OK;
}

// 1 if a bounded wait of sphy_route_scan_exchange_p2p gave up since the last call (synchronises the stream)
extern "C" int sphy_p2p_timed_out(sphy_ctx *ctx, void *stream)
{
    if (!ctx || !ctx->p2p_flags) return 0;
    long long w = 0;
    if (cudaMemcpyAsync(&w, ctx->p2p_flags + ctx->p2p_world, sizeof(w), cudaMemcpyDeviceToHost, (cudaStream_t)stream) != cudaSuccess) return 1;
    if (cudaStreamSynchronize((cudaStream_t)stream) != cudaSuccess) return 1;
    return w != 0;
}
