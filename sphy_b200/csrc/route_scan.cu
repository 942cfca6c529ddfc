// route_scan.cu -- K4b: flow accumulation as subtree range sums over a DFS pre-order (no level latency chain).
//
// accuflux is linear and the drainage network is a forest, so with the cells stored in a depth-first pre-order
// (SPHY_ORDER_DFS, built in ldd_build.cu) the subtree of cell j is the contiguous index range [j, j + sub_size[j]): its accumulated flux is a difference of prefix
// sums of the per-cell material taken in rank order.  One step is therefore two streaming passes instead of a
// walk over ~10^4 dependent levels:
//   pass A  per tile of 2048 cells: stream runoff, convert to m3/s (routing.py:26), float64 inclusive scan inside
//           the tile (cub::BlockScan), tile total;   + a tiny exclusive scan of the tile totals;
//   pass B  per cell: range sum from the tile-local prefixes (+ tile prefix only when the subtree spans tiles, so
//           small subtrees never see the global magnitude), REAL4 flux, kx recession blend (routing.py:28).
// The float64 range sums have no per-cell REAL4 rounding along the flow path, so this path agrees with the
// level sweep / the oracle to ~1e-6 relative (tested at 1e-5), not bit-for-bit; the level sweep stays the
// bit-exact reference path.  Pre-order ranks are built once from the Appendix-C structures.
#include <cub/cub.cuh>

#include "common.cuh"

namespace {

constexpr int SCAN_THREADS = 256;
constexpr int SCAN_ITEMS = 8;
constexpr int SCAN_TILE = SCAN_THREADS * SCAN_ITEMS;   // 2048 ranks

struct ScanArgs {
    int ncomp;
    const float *m[ROUTE_MAX_COMP];       // runoff mm/day, cell order
    float *qstate[ROUTE_MAX_COMP];        // Qold in / Q out, cell order
    const int32_t *sub_size;
    double *local;                        // [ncomp][n] tile-local inclusive prefix, by rank
    double *tile_tot;                     // [ncomp][ntiles+1] tile totals, then exclusive prefix in place
    sphy_param kx;
    float one_m_kx, cellarea;
    int64_t n;
    int64_t stride;                       // per-component stride of `local` (n rounded up to even: 16-byte aligned rows)
    int32_t ntiles;
};

// pass A
__global__ void __launch_bounds__(SCAN_THREADS) k_scan_tiles(const __grid_constant__ ScanArgs a)
{
    typedef cub::BlockScan<double, SCAN_THREADS> BlockScan;
    __shared__ typename BlockScan::TempStorage tmp;
    const int64_t base = (int64_t)blockIdx.x * SCAN_TILE + (int64_t)threadIdx.x * SCAN_ITEMS;
    const bool full = base + SCAN_ITEMS <= a.n;
    for (int c = 0; c < a.ncomp; ++c) {
        double v[SCAN_ITEMS];
        const float *m = a.m[c];
        if (full) {                                        // 2 x 128-bit loads per thread
            const float4 x = __ldg(reinterpret_cast<const float4 *>(m + base));
            const float4 y = __ldg(reinterpret_cast<const float4 *>(m + base + 4));
            const float q[SCAN_ITEMS] = {x.x, x.y, x.z, x.w, y.x, y.y, y.z, y.w};
#pragma unroll
            for (int j = 0; j < SCAN_ITEMS; ++j) v[j] = (double)(((q[j] * 0.001f) * a.cellarea) / 86400.0f);   // routing.py:26
        } else {
#pragma unroll
            for (int j = 0; j < SCAN_ITEMS; ++j)
                v[j] = base + j < a.n ? (double)(((__ldg(m + base + j) * 0.001f) * a.cellarea) / 86400.0f) : 0.0;
        }
        double total;
        BlockScan(tmp).InclusiveSum(v, v, total);
        double *out = a.local + (int64_t)c * a.stride;
        if (full) {
#pragma unroll
            for (int j = 0; j < SCAN_ITEMS; j += 2) *reinterpret_cast<double2 *>(out + base + j) = make_double2(v[j], v[j + 1]);
        } else {
#pragma unroll
            for (int j = 0; j < SCAN_ITEMS; ++j)
                if (base + j < a.n) out[base + j] = v[j];
        }
        if (threadIdx.x == 0) a.tile_tot[(int64_t)c * (a.ntiles + 1) + blockIdx.x] = total;
        __syncthreads();
    }
}

// exclusive scan of the tile totals, one CTA per component (ntiles ~ 5e4 at 10^8 cells)
__global__ void __launch_bounds__(1024) k_scan_tile_totals(const __grid_constant__ ScanArgs a)
{
    typedef cub::BlockScan<double, 1024> BlockScan;
    __shared__ typename BlockScan::TempStorage tmp;
    __shared__ double carry_s;
    double *t = a.tile_tot + (int64_t)blockIdx.x * (a.ntiles + 1);
    if (threadIdx.x == 0) carry_s = 0.0;
    __syncthreads();
    for (int32_t b = 0; b < a.ntiles + 1; b += 1024) {
        const int32_t k = b + threadIdx.x;
        double v = k < a.ntiles ? t[k] : 0.0, ex, tot;
        BlockScan(tmp).ExclusiveSum(v, ex, tot);
        const double carry = carry_s;
        if (k < a.ntiles + 1) t[k] = carry + ex;
        __syncthreads();
        if (threadIdx.x == 0) carry_s = carry + tot;
        __syncthreads();
    }
}

// pass B
__global__ void __launch_bounds__(256) k_scan_apply(const __grid_constant__ ScanArgs a)
{
    const int64_t r = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (r >= a.n) return;
    const int64_t cell = r;                                   // device order IS the pre-order
    const int64_t e = r + __ldg(a.sub_size + r) - 1;          // last index of the subtree
    const int32_t ta = (int32_t)(r / SCAN_TILE), te = (int32_t)(e / SCAN_TILE);
    const bool a_first = (r % SCAN_TILE) == 0;
    const float kx = a.kx.map ? __ldg(a.kx.map + cell) : a.kx.value;
    const float omk = a.kx.map ? 1.0f - kx : a.one_m_kx;
    for (int c = 0; c < a.ncomp; ++c) {
        const double *L = a.local + (int64_t)c * a.stride;
        const double *T = a.tile_tot + (int64_t)c * (a.ntiles + 1);
        const double before = a_first ? 0.0 : L[r - 1];       // tile-local exclusive prefix at r
        double s;
        if (ta == te) {
            s = L[e] - before;
        } else {
            const int64_t ta_last = (int64_t)(ta + 1) * SCAN_TILE - 1;
            s = (L[ta_last] - before) + (T[te] - T[ta + 1]) + L[e];
        }
        const float f = (float)s;
        float *qs = a.qstate[c];
        qs[cell] = omk * f + kx * qs[cell];                    // routing.py:28
    }
}

}  // namespace

int route_scan_run(sphy_ctx *ctx, int ncomp, const float *const *runoff_mm, float *const *q_state, sphy_param kx,
                   float one_m_kx, cudaStream_t st)
{
    if (ctx->order_mode != SPHY_ORDER_DFS)
        SPHY_FAIL(ctx, SPHY_E_STATE, "SPHY_ROUTE_SCAN needs the cells in SPHY_ORDER_DFS (sphy_set_cell_order)");
    ctx->n_tiles = (int32_t)((ctx->n + SCAN_TILE - 1) / SCAN_TILE);
    if (ctx->scan_comp_cap < ncomp) {             // grows only when more components are first routed
        int rc;
        if ((rc = dev_alloc(ctx, &ctx->scan_local, (size_t)ncomp * (ctx->n + 1))) ||
            (rc = dev_alloc(ctx, &ctx->scan_tile, (size_t)ncomp * (ctx->n_tiles + 1))))
            return rc;
        ctx->scan_comp_cap = ncomp;
    }
    ScanArgs a = {};
    a.ncomp = ncomp;
    for (int c = 0; c < ncomp; ++c) {
        a.m[c] = runoff_mm[c];
        a.qstate[c] = q_state[c];
    }
    a.sub_size = ctx->sub_size;
    a.local = ctx->scan_local;
    a.tile_tot = ctx->scan_tile;
    a.kx = kx;
    a.one_m_kx = one_m_kx;
    a.cellarea = ctx->cellarea;
    a.n = ctx->n;
    a.stride = (ctx->n + 1) & ~(int64_t)1;
    a.ntiles = ctx->n_tiles;
    k_scan_tiles<<<ctx->n_tiles, SCAN_THREADS, 0, st>>>(a);
    k_scan_tile_totals<<<ncomp, 1024, 0, st>>>(a);
    k_scan_apply<<<div_up(ctx->n, 256), 256, 0, st>>>(a);
    SPHY_LAUNCH_CHECK(ctx);
    return SPHY_OK;
}

