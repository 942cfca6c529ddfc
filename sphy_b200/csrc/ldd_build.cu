// ldd_build.cu -- K3: UINT1 LDD raster -> integer routing structures, entirely on device.
//
// Replaces what PCRaster derives implicitly inside accuflux/accufractionflux/upstream from the LDD
// read at /root/reference/modules/routing.py:34.  All structures follow SURVEY.md Appendix C and are
// compared bit-for-bit with oracle/ldd_oracle.c in tests/test_gpu_ldd.py.
#include <cub/cub.cuh>

#include "common.cuh"

namespace {

__constant__ int c_nb_dr[8] = {-1, -1, -1, 0, 0, 1, 1, 1};   // neighbours in ascending raster index
__constant__ int c_nb_dc[8] = {-1, 0, 1, -1, 1, -1, 0, 1};
// LDD code a neighbour at offset (dr,dc) must carry to drain into the centre cell: it steps (-dr,-dc)
// keypad: code = 5 - 3*drow + dcol
__device__ __forceinline__ int code_into_centre(int dr, int dc) { return 5 - 3 * (-dr) + (-dc); }

__global__ void k_active_flag(const uint8_t *ldd, const uint8_t *active, int64_t ncell, int32_t *flag)
{
    int64_t g = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (g >= ncell) return;
    uint8_t c = ldd[g];
    flag[g] = (c >= 1 && c <= 9 && (!active || active[g])) ? 1 : 0;
}

__global__ void k_finalize_cidx(const int32_t *flag, const int32_t *scan, int64_t ncell, int32_t *cidx, int32_t *flat_active)
{
    int64_t g = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (g >= ncell) return;
    if (flag[g]) {
        cidx[g] = scan[g];
        flat_active[scan[g]] = (int32_t)g;
    } else {
        cidx[g] = -1;
    }
}

__global__ void k_down_indeg(const uint8_t *ldd, const int32_t *cidx, const int32_t *flat_active, int nrows, int ncols,
                             int64_t n, int32_t *down, int32_t *indeg32, uint8_t *indeg8)
{
    int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (i >= n) return;
    int32_t g = flat_active[i];
    int r = g / ncols, c = g % ncols;
    int code = ldd[g];
    int32_t d = -1;
    if (code != 5) {
        int dr = (code <= 3) ? 1 : (code >= 7 ? -1 : 0);
        int dc = ((code - 1) % 3) - 1;
        int rr = r + dr, cc = c + dc;
        if (rr >= 0 && rr < nrows && cc >= 0 && cc < ncols) d = cidx[(int64_t)rr * ncols + cc];
    }
    down[i] = d;
    int cnt = 0;
#pragma unroll
    for (int k = 0; k < 8; ++k) {
        int rr = r + c_nb_dr[k], cc = c + c_nb_dc[k];
        if (rr < 0 || rr >= nrows || cc < 0 || cc >= ncols) continue;
        int64_t gn = (int64_t)rr * ncols + cc;
        if (cidx[gn] >= 0 && ldd[gn] == code_into_centre(c_nb_dr[k], c_nb_dc[k])) cnt++;
    }
    indeg32[i] = cnt;
    indeg8[i] = (uint8_t)cnt;
}

__global__ void k_donors_fill(const uint8_t *ldd, const int32_t *cidx, const int32_t *flat_active, int nrows, int ncols,
                              int64_t n, const int32_t *don_ptr, int32_t *don_idx)
{
    int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (i >= n) return;
    int32_t g = flat_active[i];
    int r = g / ncols, c = g % ncols;
    int32_t w = don_ptr[i];
#pragma unroll
    for (int k = 0; k < 8; ++k) {
        int rr = r + c_nb_dr[k], cc = c + c_nb_dc[k];
        if (rr < 0 || rr >= nrows || cc < 0 || cc >= ncols) continue;
        int64_t gn = (int64_t)rr * ncols + cc;
        int32_t j = cidx[gn];
        if (j >= 0 && ldd[gn] == code_into_centre(c_nb_dr[k], c_nb_dc[k])) don_idx[w++] = j;   // ascending index
    }
}

__global__ void k_sources(const int32_t *indeg32, int64_t n, int32_t *frontier, int32_t *count)
{
    int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (i >= n) return;
    if (indeg32[i] == 0) frontier[atomicAdd(count, 1)] = (int32_t)i;
}

// one Kahn round: every frontier cell has all donors finished
__global__ void k_kahn_round(const int32_t *frontier, const int32_t *n_front, int32_t *next, int32_t *n_next,
                             int32_t *rem, const int32_t *down, const int32_t *don_ptr, const int32_t *don_idx,
                             int64_t *nup, int32_t *level, int32_t lev)
{
    int32_t nf = *n_front;
    for (int32_t k = blockIdx.x * blockDim.x + threadIdx.x; k < nf; k += gridDim.x * blockDim.x) {
        int32_t i = frontier[k];
        level[i] = lev;
        int64_t s = 1;
        for (int32_t q = don_ptr[i]; q < don_ptr[i + 1]; ++q) s += nup[don_idx[q]];
        nup[i] = s;
        int32_t d = down[i];
        if (d >= 0 && atomicSub(&rem[d], 1) == 1) next[atomicAdd(n_next, 1)] = d;
    }
}

__global__ void k_iota(int32_t *a, int64_t n)
{
    int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (i < n) a[i] = (int32_t)i;
}

__global__ void k_level_ptr(const int32_t *sorted_level, int64_t n, int32_t nlevels, int32_t *level_ptr)
{
    int64_t k = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (k >= n) return;
    if (k == 0 || sorted_level[k] != sorted_level[k - 1]) level_ptr[sorted_level[k]] = (int32_t)k;
    if (k == n - 1) level_ptr[nlevels] = (int32_t)n;
}

__global__ void k_pos(const int32_t *order, int64_t n, const uint8_t *indeg, int32_t *pos, int32_t *deg_sorted)
{
    int64_t k = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (k >= n) return;
    int32_t i = order[k];
    pos[i] = (int32_t)k;
    deg_sorted[k] = indeg[i];
}

__global__ void k_dpos_fill(const int32_t *order, const int32_t *pos, const int32_t *don_ptr, const int32_t *don_idx,
                            const int32_t *dpos_ptr, int64_t n, int32_t *dpos)
{
    int64_t k = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (k >= n) return;
    int32_t i = order[k];
    int32_t w = dpos_ptr[k];
    for (int32_t q = don_ptr[i]; q < don_ptr[i + 1]; ++q) dpos[w++] = pos[don_idx[q]];
}

}  // namespace

// level-sweep launch plan: wide levels get their own grid launch, runs of narrow levels share one CTA
static void build_plan(sphy_ctx *ctx)
{
    const int WIDE = ROUTE_WIDE_LEVEL;
    ctx->plan.clear();
    int32_t L = ctx->nlevels;
    int32_t l = 0;
    while (l < L) {
        int32_t sz = ctx->h_level_ptr[l + 1] - ctx->h_level_ptr[l];
        if (sz >= WIDE) {
            ctx->plan.push_back({l, 1});
            ++l;
        } else {
            int32_t s = l;
            while (l < L && ctx->h_level_ptr[l + 1] - ctx->h_level_ptr[l] < WIDE) ++l;
            ctx->plan.push_back({s, l - s});
        }
    }
}


// ---- device cell order ------------------------------------------------------------------------------------
// The canonical structures above (c_*) use compact row-major numbering and are what
// sphy_ldd_export returns.  The kernels work in "device order": either the same numbering (SPHY_ORDER_ROWMAJOR)
// or a depth-first pre-order of the drainage forest (SPHY_ORDER_DFS, default), in which the subtree of a cell is
// the contiguous index range [j, j + sub_size[j]) -- the layout the scan routing streams over without gathers.
namespace {

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

// a cell hands consecutive rank ranges to its donors: the first donor gets pre+1
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

__global__ void k_identity(int64_t n, int32_t *pre)
{
    int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (i < n) pre[i] = (int32_t)i;
}

// per device cell j (= pre[i]): inverse map, raster index, receiver, routing position, subtree size, in-degree
__global__ void k_relabel_cells(const int32_t *pre, const int32_t *c_flat, const int32_t *c_down, const int32_t *c_pos,
                                const int64_t *nup, const uint8_t *indeg, int64_t n, int32_t *cell_canon, int32_t *flat,
                                int32_t *down, int32_t *pos, int32_t *sub_size, int32_t *deg)
{
    int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (i >= n) return;
    const int32_t j = pre[i];
    cell_canon[j] = (int32_t)i;
    flat[j] = c_flat[i];
    const int32_t d = c_down[i];
    down[j] = d >= 0 ? pre[d] : -1;
    pos[j] = c_pos[i];
    sub_size[j] = (int32_t)nup[i];
    deg[j] = indeg[i];
}

__global__ void k_relabel_raster(const int32_t *c_cidx, const int32_t *pre, int64_t nraster, int32_t *cidx)
{
    int64_t g = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (g < nraster) cidx[g] = c_cidx[g] >= 0 ? pre[c_cidx[g]] : -1;
}

__global__ void k_relabel_order(const int32_t *c_order, const int32_t *pre, int64_t n, int32_t *order)
{
    int64_t k = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (k < n) order[k] = pre[c_order[k]];
}

__global__ void k_relabel_donors(const int32_t *cell_canon, const int32_t *pre, const int32_t *c_don_ptr, const int32_t *c_don_idx,
                                 const int32_t *don_ptr, int64_t n, int32_t *don_idx)
{
    int64_t j = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (j >= n) return;
    const int32_t i = cell_canon[j];
    int32_t w = don_ptr[j];
    for (int32_t q = c_don_ptr[i]; q < c_don_ptr[i + 1]; ++q) don_idx[w++] = pre[c_don_idx[q]];
}

}  // namespace

static int finalize_cell_order(sphy_ctx *ctx, cudaStream_t st)
{
    const int64_t n = ctx->n;
    const int TB = 256;
    int rc;
    if ((rc = dev_alloc(ctx, &ctx->pre, n)) || (rc = dev_alloc(ctx, &ctx->cell_canon, n)) || (rc = dev_alloc(ctx, &ctx->sub_size, n)) ||
        (rc = dev_alloc(ctx, &ctx->flat_active, n)) || (rc = dev_alloc(ctx, &ctx->down, n)) || (rc = dev_alloc(ctx, &ctx->pos, n)) ||
        (rc = dev_alloc(ctx, &ctx->order, n)) || (rc = dev_alloc(ctx, &ctx->don_ptr, n + 1)) ||
        (rc = dev_alloc(ctx, &ctx->don_idx, (size_t)ctx->n_don)) || (rc = dev_alloc(ctx, &ctx->cidx, (size_t)ctx->nraster)))
        return rc;
    int32_t *val = nullptr, *off = nullptr, *deg = nullptr;
    void *tmp = nullptr;
    size_t tb = 0;
    cudaError_t e = cudaSuccess;
#define FK(call) do { if (e == cudaSuccess) e = (call); } while (0)
    FK(cudaMalloc(&deg, sizeof(int32_t) * (n + 1)));
    if (ctx->order_mode == SPHY_ORDER_DFS) {
        FK(cudaMalloc(&val, sizeof(int32_t) * n));
        FK(cudaMalloc(&off, sizeof(int32_t) * n));
        if (e == cudaSuccess) {
            k_root_val<<<div_up(n, TB), TB, 0, st>>>(ctx->c_down, ctx->nup, n, val);
            FK(cub::DeviceScan::ExclusiveSum(nullptr, tb, val, off, (int)n, st));
            FK(cudaMalloc(&tmp, tb));
            FK(cub::DeviceScan::ExclusiveSum(tmp, tb, val, off, (int)n, st));
            k_root_assign<<<div_up(n, TB), TB, 0, st>>>(ctx->c_down, off, n, ctx->pre);
            // parents before children: walk the level-sweep plan from the highest level down
            for (auto it = ctx->plan.rbegin(); it != ctx->plan.rend(); ++it) {
                const int32_t l0 = it->first_level, nl = it->n_levels;
                const int32_t k0 = ctx->h_level_ptr[l0], k1 = ctx->h_level_ptr[l0 + nl];
                if (nl == 1 && k1 - k0 >= ROUTE_WIDE_LEVEL)
                    k_children_wide<<<div_up(k1 - k0, TB), TB, 0, st>>>(ctx->c_order, k0, k1, ctx->c_don_ptr, ctx->c_don_idx, ctx->nup, ctx->pre);
                else
                    k_children_narrow<<<1, 1024, 0, st>>>(ctx->c_order, ctx->level_ptr, l0 + nl - 1, l0, ctx->c_don_ptr, ctx->c_don_idx,
                                                          ctx->nup, ctx->pre);
            }
        }
    } else {
        k_identity<<<div_up(n, TB), TB, 0, st>>>(n, ctx->pre);
    }
    if (e == cudaSuccess) {
        FK(cudaMemsetAsync(deg + n, 0, sizeof(int32_t), st));
        k_relabel_cells<<<div_up(n, TB), TB, 0, st>>>(ctx->pre, ctx->c_flat_active, ctx->c_down, ctx->c_pos, ctx->nup, ctx->indeg, n,
                                                      ctx->cell_canon, ctx->flat_active, ctx->down, ctx->pos, ctx->sub_size, deg);
        k_relabel_raster<<<div_up(ctx->nraster, TB), TB, 0, st>>>(ctx->c_cidx, ctx->pre, ctx->nraster, ctx->cidx);
        k_relabel_order<<<div_up(n, TB), TB, 0, st>>>(ctx->c_order, ctx->pre, n, ctx->order);
        size_t need = 0;
        FK(cub::DeviceScan::ExclusiveSum(nullptr, need, deg, ctx->don_ptr, (int)(n + 1), st));
        if (need > tb) { cudaFree(tmp); tmp = nullptr; FK(cudaMalloc(&tmp, need)); tb = need; }
        FK(cub::DeviceScan::ExclusiveSum(tmp, need, deg, ctx->don_ptr, (int)(n + 1), st));
        k_relabel_donors<<<div_up(n, TB), TB, 0, st>>>(ctx->cell_canon, ctx->pre, ctx->c_don_ptr, ctx->c_don_idx, ctx->don_ptr, n, ctx->don_idx);
        FK(cudaGetLastError());
        FK(cudaStreamSynchronize(st));
    }
#undef FK
    cudaFree(val); cudaFree(off); cudaFree(deg); cudaFree(tmp);
    if (e != cudaSuccess) SPHY_FAIL(ctx, SPHY_E_CUDA, "finalize_cell_order: %s", cudaGetErrorString(e));
    ctx->sub_size_loc = ctx->sub_size;
    ctx->flat_loc = ctx->flat_active;
    ctx->partitioned = false;
    ctx->part_begin = 0;
    ctx->n_global = n;
    return SPHY_OK;
}

extern "C" int sphy_ldd_build(sphy_ctx *ctx, const uint8_t *ldd, const uint8_t *active, void *stream_)
{
    if (!ctx || !ldd) return SPHY_E_INVALID;
    cudaStream_t st = (cudaStream_t)stream_;
    SPHY_CUDA(ctx, cudaSetDevice(ctx->device));
    const int64_t nc = ctx->nraster;
    const int TB = 256;
    int32_t *flag = nullptr, *scan = nullptr;
    void *cubtmp = nullptr;
    size_t cubbytes = 0;
    int rc = SPHY_OK;
    int32_t *indeg32 = nullptr, *rem = nullptr, *fr_a = nullptr, *fr_b = nullptr, *cnt = nullptr, *keys_in = nullptr,
            *keys_out = nullptr, *vals_in = nullptr, *deg_sorted = nullptr;
#define CLEANUP()                                                                                    \
    do {                                                                                             \
        cudaFree(flag); cudaFree(scan); cudaFree(cubtmp); cudaFree(indeg32); cudaFree(rem);          \
        cudaFree(fr_a); cudaFree(fr_b); cudaFree(cnt); cudaFree(keys_out); cudaFree(vals_in);        \
        cudaFree(deg_sorted);                                                                        \
    } while (0)
#define CK(call)                                                                                     \
    do {                                                                                             \
        cudaError_t _e = (call);                                                                     \
        if (_e != cudaSuccess) {                                                                     \
            CLEANUP();                                                                               \
            SPHY_FAIL(ctx, SPHY_E_CUDA, "%s: %s (%s:%d)", #call, cudaGetErrorString(_e), __FILE__, __LINE__); \
        }                                                                                            \
    } while (0)

    ctx->ldd_ready = false;
    ctx->scan_ready = false;
    ctx->qout_zeroed = false;
    CK(cudaMalloc(&flag, sizeof(int32_t) * nc));
    CK(cudaMalloc(&scan, sizeof(int32_t) * nc));
    if ((rc = dev_alloc(ctx, &ctx->c_cidx, nc)) != SPHY_OK) { CLEANUP(); return rc; }
    k_active_flag<<<div_up(nc, TB), TB, 0, st>>>(ldd, active, nc, flag);
    CK(cudaGetLastError());
    CK(cub::DeviceScan::ExclusiveSum(nullptr, cubbytes, flag, scan, (int)nc, st));
    CK(cudaMalloc(&cubtmp, cubbytes));
    CK(cub::DeviceScan::ExclusiveSum(cubtmp, cubbytes, flag, scan, (int)nc, st));
    int32_t last_scan = 0, last_flag = 0;
    CK(cudaMemcpyAsync(&last_scan, scan + nc - 1, sizeof(int32_t), cudaMemcpyDeviceToHost, st));
    CK(cudaMemcpyAsync(&last_flag, flag + nc - 1, sizeof(int32_t), cudaMemcpyDeviceToHost, st));
    CK(cudaStreamSynchronize(st));
    const int64_t n = (int64_t)last_scan + last_flag;
    ctx->n = n;
    if (n <= 0) { CLEANUP(); SPHY_FAIL(ctx, SPHY_E_INVALID, "LDD has no active cells"); }

    if ((rc = dev_alloc(ctx, &ctx->c_flat_active, n)) || (rc = dev_alloc(ctx, &ctx->c_down, n)) ||
        (rc = dev_alloc(ctx, &ctx->indeg, n)) || (rc = dev_alloc(ctx, &ctx->level, n)) ||
        (rc = dev_alloc(ctx, &ctx->c_order, n)) || (rc = dev_alloc(ctx, &ctx->c_don_ptr, n + 1)) ||
        (rc = dev_alloc(ctx, &ctx->nup, n)) || (rc = dev_alloc(ctx, &ctx->c_pos, n)) ||
        (rc = dev_alloc(ctx, &ctx->dpos_ptr, n + 1))) { CLEANUP(); return rc; }
    k_finalize_cidx<<<div_up(nc, TB), TB, 0, st>>>(flag, scan, nc, ctx->c_cidx, ctx->c_flat_active);
    CK(cudaGetLastError());
    CK(cudaMalloc(&indeg32, sizeof(int32_t) * (n + 1)));
    CK(cudaMemsetAsync(indeg32 + n, 0, sizeof(int32_t), st));
    k_down_indeg<<<div_up(n, TB), TB, 0, st>>>(ldd, ctx->c_cidx, ctx->c_flat_active, ctx->nrows, ctx->ncols, n, ctx->c_down,
                                              indeg32, ctx->indeg);
    CK(cudaGetLastError());
    // donor CSR
    {
        size_t need = 0;
        CK(cub::DeviceScan::ExclusiveSum(nullptr, need, indeg32, ctx->c_don_ptr, (int)(n + 1), st));
        if (need > cubbytes) { cudaFree(cubtmp); cubtmp = nullptr; CK(cudaMalloc(&cubtmp, need)); cubbytes = need; }
        CK(cub::DeviceScan::ExclusiveSum(cubtmp, need, indeg32, ctx->c_don_ptr, (int)(n + 1), st));
    }
    int32_t ndon = 0;
    CK(cudaMemcpyAsync(&ndon, ctx->c_don_ptr + n, sizeof(int32_t), cudaMemcpyDeviceToHost, st));
    CK(cudaStreamSynchronize(st));
    ctx->n_don = ndon;
    if ((rc = dev_alloc(ctx, &ctx->c_don_idx, (size_t)ndon)) || (rc = dev_alloc(ctx, &ctx->dpos, (size_t)ndon))) { CLEANUP(); return rc; }
    k_donors_fill<<<div_up(n, TB), TB, 0, st>>>(ldd, ctx->c_cidx, ctx->c_flat_active, ctx->nrows, ctx->ncols, n, ctx->c_don_ptr,
                                               ctx->c_don_idx);
    CK(cudaGetLastError());

    // Kahn peeling -> level, nup
    const int BATCH = 128;
    CK(cudaMalloc(&rem, sizeof(int32_t) * n));
    CK(cudaMemcpyAsync(rem, indeg32, sizeof(int32_t) * n, cudaMemcpyDeviceToDevice, st));
    CK(cudaMalloc(&fr_a, sizeof(int32_t) * n));
    CK(cudaMalloc(&fr_b, sizeof(int32_t) * n));
    CK(cudaMalloc(&cnt, sizeof(int32_t) * (BATCH + 1)));
    CK(cudaMemsetAsync(cnt, 0, sizeof(int32_t) * (BATCH + 1), st));
    k_sources<<<div_up(n, TB), TB, 0, st>>>(indeg32, n, fr_a, cnt);
    CK(cudaGetLastError());
    std::vector<int32_t> hcnt(BATCH + 1);
    int64_t processed = 0;
    int32_t lev = 0;
    int32_t *cur = fr_a, *nxt = fr_b;
    const int kgrid = ctx->sm_count * 4;
    bool done = false;
    while (!done) {
        for (int j = 0; j < BATCH; ++j) {
            k_kahn_round<<<kgrid, 256, 0, st>>>(cur, cnt + j, nxt, cnt + j + 1, rem, ctx->c_down, ctx->c_don_ptr, ctx->c_don_idx,
                                               ctx->nup, ctx->level, lev + j);
            std::swap(cur, nxt);
        }
        CK(cudaGetLastError());
        CK(cudaMemcpyAsync(hcnt.data(), cnt, sizeof(int32_t) * (BATCH + 1), cudaMemcpyDeviceToHost, st));
        CK(cudaStreamSynchronize(st));
        int j = 0;
        for (; j < BATCH; ++j) {
            if (hcnt[j] == 0) { done = true; break; }
            processed += hcnt[j];
        }
        lev += j;
        if (!done) {
            if (hcnt[BATCH] == 0) { done = true; break; }
            CK(cudaMemsetAsync(cnt, 0, sizeof(int32_t) * (BATCH + 1), st));
            CK(cudaMemcpyAsync(cnt, &hcnt[BATCH], sizeof(int32_t), cudaMemcpyHostToDevice, st));
            CK(cudaStreamSynchronize(st));
        }
    }
    if (processed != n) {
        CLEANUP();
        SPHY_FAIL(ctx, SPHY_E_CYCLE, "unsound LDD: %lld of %lld cells are on a cycle", (long long)(n - processed), (long long)n);
    }
    ctx->nlevels = lev;

    // order = stable sort of cells by level (radix sort keeps ascending index inside a level)
    keys_in = ctx->level;
    CK(cudaMalloc(&keys_out, sizeof(int32_t) * n));
    CK(cudaMalloc(&vals_in, sizeof(int32_t) * n));
    k_iota<<<div_up(n, TB), TB, 0, st>>>(vals_in, n);
    CK(cudaGetLastError());
    {
        int end_bit = 1;
        while ((1ll << end_bit) < (long long)lev + 1 && end_bit < 31) ++end_bit;
        size_t need = 0;
        CK(cub::DeviceRadixSort::SortPairs(nullptr, need, keys_in, keys_out, vals_in, ctx->c_order, (int)n, 0, end_bit, st));
        if (need > cubbytes) { cudaFree(cubtmp); cubtmp = nullptr; CK(cudaMalloc(&cubtmp, need)); cubbytes = need; }
        CK(cub::DeviceRadixSort::SortPairs(cubtmp, need, keys_in, keys_out, vals_in, ctx->c_order, (int)n, 0, end_bit, st));
    }
    if ((rc = dev_alloc(ctx, &ctx->level_ptr, (size_t)lev + 1)) != SPHY_OK) { CLEANUP(); return rc; }
    k_level_ptr<<<div_up(n, TB), TB, 0, st>>>(keys_out, n, lev, ctx->level_ptr);
    CK(cudaGetLastError());
    // routing-order (position) structures
    CK(cudaMalloc(&deg_sorted, sizeof(int32_t) * (n + 1)));
    CK(cudaMemsetAsync(deg_sorted + n, 0, sizeof(int32_t), st));
    k_pos<<<div_up(n, TB), TB, 0, st>>>(ctx->c_order, n, ctx->indeg, ctx->c_pos, deg_sorted);
    CK(cudaGetLastError());
    {
        size_t need = 0;
        CK(cub::DeviceScan::ExclusiveSum(nullptr, need, deg_sorted, ctx->dpos_ptr, (int)(n + 1), st));
        if (need > cubbytes) { cudaFree(cubtmp); cubtmp = nullptr; CK(cudaMalloc(&cubtmp, need)); cubbytes = need; }
        CK(cub::DeviceScan::ExclusiveSum(cubtmp, need, deg_sorted, ctx->dpos_ptr, (int)(n + 1), st));
    }
    k_dpos_fill<<<div_up(n, TB), TB, 0, st>>>(ctx->c_order, ctx->c_pos, ctx->c_don_ptr, ctx->c_don_idx, ctx->dpos_ptr, n, ctx->dpos);
    CK(cudaGetLastError());
    ctx->h_level_ptr.resize((size_t)lev + 1);
    CK(cudaMemcpyAsync(ctx->h_level_ptr.data(), ctx->level_ptr, sizeof(int32_t) * ((size_t)lev + 1), cudaMemcpyDeviceToHost, st));
    CK(cudaStreamSynchronize(st));
    build_plan(ctx);
    // workspaces sized by N
    if ((rc = dev_alloc(ctx, &ctx->acc, (size_t)ROUTE_MAX_COMP * n)) || (rc = dev_alloc(ctx, &ctx->qsorted, (size_t)ROUTE_MAX_COMP * n))) {
        CLEANUP();
        return rc;
    }
    for (int k = 0; k < 4; ++k)
        if ((rc = dev_alloc(ctx, &ctx->tmp_n[k], (size_t)n)) != SPHY_OK) { CLEANUP(); return rc; }
    CLEANUP();
#undef CK
#undef CLEANUP
    rc = finalize_cell_order(ctx, st);
    if (rc != SPHY_OK) return rc;
    ctx->ldd_ready = true;
    return SPHY_OK;
}

extern "C" int64_t sphy_ncells(const sphy_ctx *ctx) { return ctx ? ctx->n : 0; }
extern "C" int32_t sphy_nlevels(const sphy_ctx *ctx) { return ctx ? ctx->nlevels : 0; }

extern "C" int sphy_ldd_export(sphy_ctx *ctx, const char *name, void *dst, size_t dst_bytes, void *stream_)
{
    if (!ctx || !name || !dst) return SPHY_E_INVALID;
    if (!ctx->ldd_ready) SPHY_FAIL(ctx, SPHY_E_STATE, "sphy_ldd_export before sphy_ldd_build");
    if (ctx->partitioned) SPHY_FAIL(ctx, SPHY_E_STATE, "sphy_ldd_export: the whole-basin structures were released by sphy_set_partition");
    cudaStream_t st = (cudaStream_t)stream_;
    std::string k(name);
    const void *src = nullptr;
    size_t bytes = 0;
    const size_t n = (size_t)ctx->n;
    if (k == "cidx") { src = ctx->c_cidx; bytes = 4 * (size_t)ctx->nraster; }
    else if (k == "down") { src = ctx->c_down; bytes = 4 * n; }
    else if (k == "indeg") { src = ctx->indeg; bytes = n; }
    else if (k == "level") { src = ctx->level; bytes = 4 * n; }
    else if (k == "order") { src = ctx->c_order; bytes = 4 * n; }
    else if (k == "level_ptr") { src = ctx->level_ptr; bytes = 4 * ((size_t)ctx->nlevels + 1); }
    else if (k == "don_ptr") { src = ctx->c_don_ptr; bytes = 4 * (n + 1); }
    else if (k == "don_idx") { src = ctx->c_don_idx; bytes = 4 * (size_t)ctx->n_don; }
    else if (k == "nup") { src = ctx->nup; bytes = 8 * n; }
    else if (k == "flat_active") { src = ctx->flat_active; bytes = 4 * n; }       /* device order -> raster index */
    else if (k == "cell_order") { src = ctx->cell_canon; bytes = 4 * n; }         /* device order -> canonical index */
    else if (k == "sub_size") { src = ctx->sub_size; bytes = 4 * n; }
    else SPHY_FAIL(ctx, SPHY_E_INVALID, "sphy_ldd_export: unknown structure '%s'", name);
    if (dst_bytes < bytes) SPHY_FAIL(ctx, SPHY_E_INVALID, "sphy_ldd_export(%s): need %zu bytes, got %zu", name, bytes, dst_bytes);
    SPHY_CUDA(ctx, cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDeviceToDevice, st));
    return SPHY_OK;
}
