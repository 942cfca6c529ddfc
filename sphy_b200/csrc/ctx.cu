// ctx.cu -- context life cycle and the raster <-> compact boundary kernels.
#include "common.cuh"

extern "C" int sphy_ctx_create(sphy_ctx **out, int device, int nrows, int ncols, float cellsize)
{
    if (!out || nrows <= 0 || ncols <= 0 || !(cellsize > 0.f)) return SPHY_E_INVALID;
    if ((int64_t)nrows * ncols >= (1ll << 31)) return SPHY_E_INVALID;
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || device < 0 || device >= ndev) return SPHY_E_CUDA;
    sphy_ctx *ctx = new (std::nothrow) sphy_ctx();
    if (!ctx) return SPHY_E_NOMEM;
    ctx->device = device;
    ctx->nrows = nrows;
    ctx->ncols = ncols;
    ctx->cellsize = cellsize;
    ctx->cellarea = cellsize * cellsize;          // REAL4 cellarea() (routing.py:26)
    ctx->nraster = (int64_t)nrows * ncols;
    cudaDeviceProp prop;
    if (cudaSetDevice(device) != cudaSuccess || cudaGetDeviceProperties(&prop, device) != cudaSuccess) {
        delete ctx;
        return SPHY_E_CUDA;
    }
    ctx->sm_count = prop.multiProcessorCount;
    *out = ctx;
    return SPHY_OK;
}

extern "C" void sphy_ctx_destroy(sphy_ctx *ctx)
{
    if (!ctx) return;
    cudaSetDevice(ctx->device);
    void *ptrs[] = {ctx->c_cidx, ctx->c_flat_active, ctx->c_down, ctx->c_order, ctx->c_pos, ctx->c_don_ptr, ctx->c_don_idx,
                    ctx->cidx, ctx->flat_active, ctx->down, ctx->level, ctx->order, ctx->level_ptr, ctx->don_ptr,
                    ctx->don_idx, ctx->indeg, ctx->nup, ctx->pos, ctx->dpos_ptr, ctx->dpos, ctx->pre, ctx->cell_canon,
                    ctx->sub_size, ctx->cross_r, ctx->cross_e, ctx->cross_ptr, ctx->end_e, ctx->end_k, ctx->end_ptr, ctx->park_before,
                    ctx->park_end, ctx->scan_tile, ctx->scan_excl, ctx->scan_tmp, ctx->scan_base, ctx->scan_counter, ctx->cut_work, ctx->cut_ptr, ctx->zero_ptr, ctx->acc, ctx->qsorted, ctx->pub_e, ctx->sweep_rec, ctx->chain_start, ctx->chain_k, ctx->chain_cell, ctx->chain_don, ctx->chain_ids,
                    ctx->sl_cell, ctx->sl_sub, ctx->sl_last, ctx->sl_tdon, ctx->sl_trunk, ctx->sl_brank, ctx->sl_bslot, ctx->sl_erank, ctx->sl_eslot,
                    ctx->sl_E, ctx->sl_Gl, ctx->sl_BS, ctx->sl_BP, ctx->sl_V, ctx->sl_counter,
                    ctx->ra_table, ctx->tmp_n[0], ctx->tmp_n[1], ctx->tmp_n[2], ctx->tmp_n[3]};
    for (void *p : ptrs)
        if (p) cudaFree(p);
    for (int i = 0; i < ctx->p2p_nopened; ++i) cudaIpcCloseMemHandle(ctx->p2p_opened[i]);
    void *p2p[] = {ctx->p2p_recv, ctx->p2p_flags, ctx->p2p_peer_recv, ctx->p2p_peer_flags};
    for (void *p : p2p)
        if (p) cudaFree(p);
    delete ctx;
}

extern "C" int sphy_set_cell_order(sphy_ctx *ctx, int mode)
{
    if (!ctx || (mode != SPHY_ORDER_ROWMAJOR && mode != SPHY_ORDER_DFS)) return SPHY_E_INVALID;
    if (ctx->ldd_ready) SPHY_FAIL(ctx, SPHY_E_STATE, "sphy_set_cell_order after sphy_ldd_build");
    ctx->order_mode = mode;
    return SPHY_OK;
}

extern "C" const char *sphy_last_error(const sphy_ctx *ctx) { return ctx ? ctx->err.c_str() : "null context"; }
extern "C" int sphy_sm_count(const sphy_ctx *ctx) { return ctx ? ctx->sm_count : 0; }

namespace {
__global__ void k_gather(const float *__restrict__ raster, const int32_t *__restrict__ flat, int64_t n, float *__restrict__ out)
{
    int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (i < n) out[i] = raster[flat[i]];
}
__global__ void k_scatter(const float *__restrict__ comp, const int32_t *__restrict__ cidx, int64_t nraster, float fill,
                          float *__restrict__ raster)
{
    int64_t g = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (g < nraster) {
        int32_t i = cidx[g];
        raster[g] = i >= 0 ? comp[i] : fill;
    }
}
// out[dst[e]] = src[src_idx[e]] with src_idx ascending: `src` may be PINNED HOST memory (unified addressing), in which case
// the device pulls exactly the cells it owns out of the host raster over PCIe -- consecutive raster cells of one rank
// coalesce into full 128-byte reads -- instead of receiving the whole raster and compacting it afterwards
__global__ void __launch_bounds__(256) k_gather_indexed(const float *__restrict__ src, const int32_t *__restrict__ src_idx,
                                                        const int32_t *__restrict__ dst_idx, int64_t n, float *__restrict__ out)
{
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t e0 = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; e0 < n; e0 += 4 * stride) {
        float v[4];
        int32_t d[4];
#pragma unroll
        for (int k = 0; k < 4; ++k) {                          // four independent reads in flight per thread
            const int64_t e = e0 + k * stride;
            if (e < n) {
                v[k] = src[__ldg(src_idx + e)];
                d[k] = __ldg(dst_idx + e);
            }
        }
#pragma unroll
        for (int k = 0; k < 4; ++k)
            if (e0 + k * stride < n) out[d[k]] = v[k];
    }
}
// piecewise-linear interpolation of scattered / gridded values onto the cells: out = REAL4( sum_k w[3i+k] * z[idx[3i+k]] ) in
// float64 in the order k = 0, 1, 2 -- what scipy's LinearNDInterpolator (griddata(method="linear")) evaluates per target
// point; idx < 0 marks a cell outside the convex hull of the source points (missing value)
__global__ void k_interp_linear(const double *__restrict__ z, const int32_t *__restrict__ idx, const double *__restrict__ w, int64_t n,
                                float *__restrict__ out)
{
    const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (i >= n) return;
    const int32_t i0 = idx[3 * i];
    if (i0 < 0) { out[i] = nanf(""); return; }
    double s = w[3 * i] * z[i0];
    s = s + w[3 * i + 1] * z[idx[3 * i + 1]];
    s = s + w[3 * i + 2] * z[idx[3 * i + 2]];
    out[i] = (float)s;
}
__global__ void k_sample(const float *__restrict__ map, const int32_t *__restrict__ cells, int n, float *__restrict__ out)
{
    int j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j < n) out[j] = map[cells[j]];
}
}  // namespace

extern "C" int sphy_gather_f32(sphy_ctx *ctx, const float *raster, float *compact, void *stream)
{
    if (!ctx || !raster || !compact) return SPHY_E_INVALID;
    if (!ctx->ldd_ready) SPHY_FAIL(ctx, SPHY_E_STATE, "sphy_gather_f32 before sphy_ldd_build");
    k_gather<<<div_up(ctx->n, 256), 256, 0, (cudaStream_t)stream>>>(raster, ctx->flat_loc, ctx->n, compact);
    SPHY_LAUNCH_CHECK(ctx);
    return SPHY_OK;
}

extern "C" int sphy_gather_indexed_f32(sphy_ctx *ctx, const float *src, const int32_t *src_idx_dev, const int32_t *dst_idx_dev,
                                       int64_t count, float *out_dev, void *stream)
{
    if (!ctx || !src || !src_idx_dev || !dst_idx_dev || !out_dev || count < 0) return SPHY_E_INVALID;
    if (count == 0) return SPHY_OK;
    const int64_t want = (count + 1023) / 1024;
    const int grid = (int)(want < (int64_t)ctx->sm_count * 8 ? want : (int64_t)ctx->sm_count * 8);
    k_gather_indexed<<<grid, 256, 0, (cudaStream_t)stream>>>(src, src_idx_dev, dst_idx_dev, count, out_dev);
    SPHY_LAUNCH_CHECK(ctx);
    return SPHY_OK;
}

extern "C" int sphy_interp_linear(sphy_ctx *ctx, const double *z_dev, const int32_t *idx3_dev, const double *w3_dev, int64_t count,
                                  float *out_dev, void *stream)
{
    if (!ctx || !z_dev || !idx3_dev || !w3_dev || !out_dev || count < 0) return SPHY_E_INVALID;
    if (count == 0) return SPHY_OK;
    k_interp_linear<<<div_up(count, 256), 256, 0, (cudaStream_t)stream>>>(z_dev, idx3_dev, w3_dev, count, out_dev);
    SPHY_LAUNCH_CHECK(ctx);
    return SPHY_OK;
}

extern "C" int sphy_scatter_f32(sphy_ctx *ctx, const float *compact, float *raster, float fill, void *stream)
{
    if (!ctx || !raster || !compact) return SPHY_E_INVALID;
    if (!ctx->ldd_ready) SPHY_FAIL(ctx, SPHY_E_STATE, "sphy_scatter_f32 before sphy_ldd_build");
    if (ctx->partitioned) SPHY_FAIL(ctx, SPHY_E_STATE, "sphy_scatter_f32 is not available on a partitioned context");
    k_scatter<<<div_up(ctx->nraster, 256), 256, 0, (cudaStream_t)stream>>>(compact, ctx->cidx, ctx->nraster, fill, raster);
    SPHY_LAUNCH_CHECK(ctx);
    return SPHY_OK;
}

extern "C" int sphy_sample(sphy_ctx *ctx, const float *map, const int32_t *cells, int n, float *out, void *stream)
{
    if (!ctx || !map || !cells || !out || n < 0) return SPHY_E_INVALID;
    if (n == 0) return SPHY_OK;
    k_sample<<<div_up(n, 128), 128, 0, (cudaStream_t)stream>>>(map, cells, n, out);
    SPHY_LAUNCH_CHECK(ctx);
    return SPHY_OK;
}

namespace {
__global__ void k_accumulate(float *__restrict__ acc, const float *__restrict__ x, float divisor, int64_t n)
{
    int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (i < n) acc[i] = acc[i] + (divisor != 0.0f ? x[i] / divisor : x[i]);
}
__global__ void k_divide(const float *__restrict__ in, float divisor, float *__restrict__ out, int64_t n)
{
    int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (i < n) out[i] = in[i] / divisor;
}
}  // namespace

extern "C" int sphy_map_accumulate(sphy_ctx *ctx, float *acc, const float *x, float divisor, int64_t count, void *stream)
{
    if (!ctx || !acc || !x || count < 0) return SPHY_E_INVALID;
    if (count == 0) return SPHY_OK;
    k_accumulate<<<div_up(count, 256), 256, 0, (cudaStream_t)stream>>>(acc, x, divisor, count);
    SPHY_LAUNCH_CHECK(ctx);
    return SPHY_OK;
}

extern "C" int sphy_map_divide(sphy_ctx *ctx, const float *in, float divisor, float *out, int64_t count, void *stream)
{
    if (!ctx || !in || !out || count < 0 || divisor == 0.0f) return SPHY_E_INVALID;
    if (count == 0) return SPHY_OK;
    k_divide<<<div_up(count, 256), 256, 0, (cudaStream_t)stream>>>(in, divisor, out, count);
    SPHY_LAUNCH_CHECK(ctx);
    return SPHY_OK;
}

extern "C" size_t sphy_abi_sizeof(int which)
{
    switch (which) {
        case 0: return sizeof(sphy_param);
        case 1: return sizeof(sphy_wb_params);
        case 2: return sizeof(sphy_wb_state);
        case 3: return sizeof(sphy_wb_forcing);
        case 4: return sizeof(sphy_wb_out);
        case 5: return sizeof(sphy_glac_table);
        case 6: return sizeof(sphy_glac_out);
        case 7: return sizeof(sphy_reslake);
        case 8: return sizeof(sphy_openwater);
        case 9: return sizeof(sphy_veg);
        case 10: return sizeof(sphy_mmf);
        case 11: return sizeof(sphy_sedres);
        case 12: return sizeof(sphy_reslake_plan);
        default: return 0;
    }
}
