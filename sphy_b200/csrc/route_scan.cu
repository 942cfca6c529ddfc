// route_scan.cu -- K4b: flow accumulation as subtree range sums over a DFS pre-order (no level latency chain).
//
// accuflux is linear and the drainage network is a forest, so if cells are ranked in a depth-first pre-order the
// subtree of a cell is the contiguous rank range [pre, pre + nup): its accumulated flux is a difference of prefix
// sums of the per-cell material taken in rank order.  One step is therefore two streaming passes instead of a
// walk over ~10^4 dependent levels:
//   pass A  per tile of 2048 ranks: gather runoff, convert to m3/s (routing.py:26), float64 inclusive scan inside
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

__global__ void k_root_val(const int32_t *down, const int64_t *nup, int64_t n, int32_t *val)
{
    int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (i < n) val[i] = down[i] < 0 ? (int32_t)nup[i] : 0;
}

__global__ void k_root_assign(const int32_t *down, const int32_t *off, int64_t n, int32_t *pre)
{
    int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (i < n && down[i] < 0) pre[i] = off[i];
}

// every cell of positions [k0,k1) hands consecutive rank ranges to its donors: first donor gets pre+1
__device__ __forceinline__ void assign_children(int32_t i, const int32_t *don_ptr, const int32_t *don_idx,
                                                const int64_t *nup, int32_t *pre)
{
    int32_t off = pre[i] + 1;
    for (int32_t q = don_ptr[i]; q < don_ptr[i + 1]; ++q) {
        const int32_t d = don_idx[q];
        pre[d] = off;
        off += (int32_t)nup[d];
    }
}

__global__ void k_children_wide(const int32_t *order, int32_t k0, int32_t k1, const int32_t *don_ptr, const int32_t *don_idx,
                                const int64_t *nup, int32_t *pre)
{
    int32_t k = k0 + blockIdx.x * blockDim.x + threadIdx.x;
    if (k < k1) assign_children(order[k], don_ptr, don_idx, nup, pre);
}

// a run of narrow levels, walked from the highest level down by one CTA
__global__ void __launch_bounds__(1024) k_children_narrow(const int32_t *order, const int32_t *level_ptr, int32_t l_hi, int32_t l_lo,
                                                          const int32_t *don_ptr, const int32_t *don_idx, const int64_t *nup,
                                                          int32_t *pre)
{
    for (int32_t l = l_hi; l >= l_lo; --l) {
        const int32_t kb = level_ptr[l], ke = level_ptr[l + 1];
        for (int32_t k = kb + threadIdx.x; k < ke; k += blockDim.x) assign_children(order[k], don_ptr, don_idx, nup, pre);
        __syncthreads();
    }
}

__global__ void k_inverse(const int32_t *pre, const int64_t *nup, int64_t n, int32_t *pre_cell, int32_t *sub_size)
{
    int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (i >= n) return;
    const int32_t r = pre[i];
    pre_cell[r] = (int32_t)i;
    sub_size[r] = (int32_t)nup[i];
}

struct ScanArgs {
    int ncomp;
    const float *m[ROUTE_MAX_COMP];       // runoff mm/day, cell order
    float *qstate[ROUTE_MAX_COMP];        // Qold in / Q out, cell order
    const int32_t *pre_cell, *sub_size;
    double *local;                        // [ncomp][n] tile-local inclusive prefix, by rank
    double *tile_tot;                     // [ncomp][ntiles+1] tile totals, then exclusive prefix in place
    sphy_param kx;
    float one_m_kx, cellarea;
    int64_t n;
    int32_t ntiles;
};

// pass A
__global__ void __launch_bounds__(SCAN_THREADS) k_scan_tiles(const __grid_constant__ ScanArgs a)
{
    typedef cub::BlockScan<double, SCAN_THREADS> BlockScan;
    __shared__ typename BlockScan::TempStorage tmp;
    const int64_t base = (int64_t)blockIdx.x * SCAN_TILE + (int64_t)threadIdx.x * SCAN_ITEMS;
    int32_t cell[SCAN_ITEMS];
#pragma unroll
    for (int j = 0; j < SCAN_ITEMS; ++j) cell[j] = base + j < a.n ? __ldg(a.pre_cell + base + j) : -1;
    for (int c = 0; c < a.ncomp; ++c) {
        double v[SCAN_ITEMS];
#pragma unroll
        for (int j = 0; j < SCAN_ITEMS; ++j) {
            float rr = 0.0f;
            if (cell[j] >= 0) rr = ((__ldg(a.m[c] + cell[j]) * 0.001f) * a.cellarea) / 86400.0f;   // routing.py:26
            v[j] = (double)rr;
        }
        double total;
        BlockScan(tmp).InclusiveSum(v, v, total);
        double *out = a.local + (int64_t)c * a.n;
#pragma unroll
        for (int j = 0; j < SCAN_ITEMS; ++j)
            if (base + j < a.n) out[base + j] = v[j];
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
    const int32_t cell = __ldg(a.pre_cell + r);
    const int64_t e = r + __ldg(a.sub_size + r) - 1;          // last rank of the subtree
    const int32_t ta = (int32_t)(r / SCAN_TILE), te = (int32_t)(e / SCAN_TILE);
    const bool a_first = (r % SCAN_TILE) == 0;
    const float kx = a.kx.map ? __ldg(a.kx.map + cell) : a.kx.value;
    const float omk = a.kx.map ? 1.0f - kx : a.one_m_kx;
    for (int c = 0; c < a.ncomp; ++c) {
        const double *L = a.local + (int64_t)c * a.n;
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

static int build_scan_layout(sphy_ctx *ctx, cudaStream_t st)
{
    const int64_t n = ctx->n;
    int rc;
    if ((rc = dev_alloc(ctx, &ctx->pre, n)) || (rc = dev_alloc(ctx, &ctx->pre_cell, n)) || (rc = dev_alloc(ctx, &ctx->sub_size, n)))
        return rc;
    int32_t *val = nullptr, *off = nullptr;
    void *tmp = nullptr;
    size_t tb = 0;
    SPHY_CUDA(ctx, cudaMalloc(&val, sizeof(int32_t) * n));
    SPHY_CUDA(ctx, cudaMalloc(&off, sizeof(int32_t) * n));
    k_root_val<<<div_up(n, 256), 256, 0, st>>>(ctx->down, ctx->nup, n, val);
    cub::DeviceScan::ExclusiveSum(nullptr, tb, val, off, (int)n, st);
    SPHY_CUDA(ctx, cudaMalloc(&tmp, tb));
    cub::DeviceScan::ExclusiveSum(tmp, tb, val, off, (int)n, st);
    k_root_assign<<<div_up(n, 256), 256, 0, st>>>(ctx->down, off, n, ctx->pre);
    // top-down over the level-sweep plan in reverse
    for (auto it = ctx->plan.rbegin(); it != ctx->plan.rend(); ++it) {
        const int32_t l0 = it->first_level, nl = it->n_levels;
        const int32_t k0 = ctx->h_level_ptr[l0], k1 = ctx->h_level_ptr[l0 + nl];
        if (nl == 1 && k1 - k0 >= ROUTE_WIDE_LEVEL)
            k_children_wide<<<div_up(k1 - k0, 256), 256, 0, st>>>(ctx->order, k0, k1, ctx->don_ptr, ctx->don_idx, ctx->nup, ctx->pre);
        else
            k_children_narrow<<<1, 1024, 0, st>>>(ctx->order, ctx->level_ptr, l0 + nl - 1, l0, ctx->don_ptr, ctx->don_idx, ctx->nup, ctx->pre);
    }
    k_inverse<<<div_up(n, 256), 256, 0, st>>>(ctx->pre, ctx->nup, n, ctx->pre_cell, ctx->sub_size);
    cudaError_t e = cudaGetLastError();
    cudaError_t e2 = cudaStreamSynchronize(st);
    cudaFree(val); cudaFree(off); cudaFree(tmp);
    if (e != cudaSuccess || e2 != cudaSuccess)
        SPHY_FAIL(ctx, SPHY_E_CUDA, "build_scan_layout: %s", cudaGetErrorString(e != cudaSuccess ? e : e2));
    ctx->n_tiles = (int32_t)((n + SCAN_TILE - 1) / SCAN_TILE);
    ctx->scan_ready = true;
    return SPHY_OK;
}

int route_scan_run(sphy_ctx *ctx, int ncomp, const float *const *runoff_mm, float *const *q_state, sphy_param kx,
                   float one_m_kx, cudaStream_t st)
{
    if (!ctx->scan_ready) {                       // one-time layout build (synchronises, allocates)
        int rc = build_scan_layout(ctx, st);
        if (rc != SPHY_OK) return rc;
    }
    if (ctx->scan_comp_cap < ncomp) {             // grows only when more components are first routed
        int rc;
        if ((rc = dev_alloc(ctx, &ctx->scan_local, (size_t)ncomp * ctx->n)) ||
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
    a.pre_cell = ctx->pre_cell;
    a.sub_size = ctx->sub_size;
    a.local = ctx->scan_local;
    a.tile_tot = ctx->scan_tile;
    a.kx = kx;
    a.one_m_kx = one_m_kx;
    a.cellarea = ctx->cellarea;
    a.n = ctx->n;
    a.ntiles = ctx->n_tiles;
    k_scan_tiles<<<ctx->n_tiles, SCAN_THREADS, 0, st>>>(a);
    k_scan_tile_totals<<<ncomp, 1024, 0, st>>>(a);
    k_scan_apply<<<div_up(ctx->n, 256), 256, 0, st>>>(a);
    SPHY_LAUNCH_CHECK(ctx);
    return SPHY_OK;
}

// device copy of the pre-order structures for tests: "pre" i32 N (cell -> rank)
int route_scan_export(sphy_ctx *ctx, const char *name, void *dst, size_t bytes, cudaStream_t st)
{
    if (!ctx->scan_ready) {
        int rc = build_scan_layout(ctx, st);
        if (rc != SPHY_OK) return rc;
    }
    const void *src = nullptr;
    std::string k(name);
    if (k == "pre") src = ctx->pre;
    else if (k == "pre_cell") src = ctx->pre_cell;
    else if (k == "sub_size") src = ctx->sub_size;
    else return SPHY_E_INVALID;
    if (bytes < sizeof(int32_t) * (size_t)ctx->n) return SPHY_E_INVALID;
    SPHY_CUDA(ctx, cudaMemcpyAsync(dst, src, sizeof(int32_t) * (size_t)ctx->n, cudaMemcpyDeviceToDevice, st));
    return SPHY_OK;
}
