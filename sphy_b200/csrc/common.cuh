// common.cuh -- context object and small helpers shared by the sm_100a translation units.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#include <string>
#include <vector>

#include "sphy_b200.h"

struct RoutePlanSeg {
    int32_t first_level;   // first level of the segment
    int32_t n_levels;      // 1 for a wide level (grid launch), >1 for a run of narrow levels (one CTA)
};

struct SweepPlanSeg {
    int32_t first_level, n_levels;
    int32_t threads;       // 0: one wide level (grid launch); else the CTA size of k_sweep_narrow for this run
};

struct sphy_ctx {
    int device = 0;
    int nrows = 0, ncols = 0;
    float cellsize = 1.f, cellarea = 1.f;
    int sm_count = 148;
    int64_t nraster = 0;
    int64_t n = 0;             // active cells
    int32_t nlevels = 0;
    bool ldd_ready = false;
    // Appendix-C structures (device), canonical compact row-major numbering (what sphy_ldd_export returns)
    int32_t *c_cidx = nullptr, *c_flat_active = nullptr, *c_down = nullptr, *c_order = nullptr, *c_pos = nullptr;
    int32_t *c_don_ptr = nullptr, *c_don_idx = nullptr;
    int32_t *level = nullptr, *level_ptr = nullptr;
    // the same structures in device cell order (what the kernels use); see finalize_cell_order()
    int order_mode = SPHY_ORDER_DFS;
    int32_t *cidx = nullptr, *flat_active = nullptr, *down = nullptr, *order = nullptr;
    int32_t *don_ptr = nullptr, *don_idx = nullptr;
    int32_t *pre = nullptr;        // canonical cell -> device index
    int32_t *cell_canon = nullptr; // device index -> canonical cell
    int32_t *sub_size = nullptr;   // by device index: upstream cell count incl. self (subtree size)
    uint8_t *indeg = nullptr;
    int64_t *nup = nullptr;
    int64_t n_don = 0;
    // level-sweep layout (routing order = position in `order`): pos[i] = position of cell i,
    // dpos_ptr/dpos = donor CSR expressed in positions, indexed by position
    int32_t *pos = nullptr, *dpos_ptr = nullptr, *dpos = nullptr;
    std::vector<int32_t> h_level_ptr;
    std::vector<RoutePlanSeg> plan;
    std::vector<SweepPlanSeg> sweep_plan;   // route.cu: the narrow runs cut by CTA size
    // scan routing (route_scan.cu): static lists of the cells whose index range crosses a tile boundary + workspaces
    bool scan_ready = false;
    bool scan_smem_opt_in = false;
    bool overlap_tail = false;     // sphy_set_overlap_tail: the routing tail runs under the next step's water-balance kernel
    void *main_pass_event = nullptr;   // caller's cudaEvent_t recorded right behind the main pass (or NULL)
    double *scan_base = nullptr;   // [ncomp][nblocks] second level of the tile-total prefix (small-CTA scan)
    unsigned *scan_counter = nullptr;
    // confluence exchange over peer memory (route_scan.cu)
    double *p2p_recv = nullptr, **p2p_peer_recv = nullptr;
    long long *p2p_flags = nullptr, **p2p_peer_flags = nullptr;
    void *p2p_opened[128] = {};
    int p2p_nopened = 0, p2p_world = 0, p2p_rank = 0;
    int64_t p2p_pitch = 0;
    int32_t *cross_r = nullptr, *cross_e = nullptr, *cross_ptr = nullptr, *end_e = nullptr, *end_k = nullptr, *end_ptr = nullptr;
    int32_t n_cross = 0;
    double *park_before = nullptr, *park_end = nullptr;
    double *scan_tile = nullptr;   // [ncomp][ntiles+1] tile totals (trailing 0)
    void *scan_tmp = nullptr;      // cub temp storage for the tile-total scans
    size_t scan_tmp_bytes = 0;
    double *scan_excl = nullptr;   // [ncomp][ntiles+1] exclusive prefix of the tile totals
    int32_t n_tiles = 0;
    int scan_comp_cap = 0;
    // multi-GPU partition: this context owns the device-index range [part_begin, part_begin + n) of the DFS order
    bool partitioned = false;
    int64_t part_begin = 0, n_global = 0;
    int32_t *sub_size_loc = nullptr, *flat_loc = nullptr;   // views into sub_size / flat_active at part_begin
    int32_t n_pub = 0;
    int32_t *pub_e = nullptr;                                // local indices whose range-local prefix this rank publishes (ascending)
    // trunk list (route_scan.cu): the cells that are finished from published prefixes instead of the tile-crossing list --
    // the downstream-closed set {level >= min_level or sub_size >= min_cells} that carries the reference's REAL4 rounding
    // drift, plus (partitioned basin) every cell whose upstream range crosses a rank boundary.  Global device indices,
    // ascending (= DFS pre-order of the trunk tree); identical on every rank.
    int32_t sl_n = 0;
    int32_t *sl_cell = nullptr, *sl_sub = nullptr, *sl_last = nullptr, *sl_tdon = nullptr;   // sl_tdon: 8 slots per entry, -1 padded
    uint8_t *sl_trunk = nullptr;                             // 1: drift applies (0: listed only because its range leaves the rank)
    bool sl_plan = false;
    bool sl_cuts = false;                                    // the list also holds the cells downstream of lakes / reservoirs
    int32_t *sl_brank = nullptr, *sl_bslot = nullptr, *sl_erank = nullptr, *sl_eslot = nullptr;   // (rank, pub slot) of j-1 and of the range end
    int32_t sl_stride = 0, sl_rank = 0, sl_world = 1, sl_nblk = 0;
    int sl_comp_cap = 0;
    double *sl_E = nullptr, *sl_Gl = nullptr, *sl_BS = nullptr, *sl_BP = nullptr;   // [ncomp][sl_n] / [ncomp][sl_nblk] workspaces
    double *sl_V = nullptr;                                  // unpartitioned context: the published prefixes [ncomp][sl_stride]
    unsigned *sl_counter = nullptr;
    // optional per-phase timing of the scan routing (sphy_route_profile): events recorded on the launching stream
    bool prof_on = false;
    std::vector<cudaEvent_t> prof_ev;                        // 4 per step: before main, after main, after cross, after trunk
    // workspaces
    float *acc = nullptr;          // [ROUTE_MAX_COMP][n] accumulated flux in routing order
    float *qsorted = nullptr;      // scratch [ROUTE_MAX_COMP][n]
    float *ra_table = nullptr;     // per-day Ra by latitude class
    const float *day_block = nullptr;   // caller-owned device copy of the day-of-year scalars (sphy_set_day_block), or NULL
    int32_t ra_table_cap = 0;
    float *tmp_n[4] = {nullptr, nullptr, nullptr, nullptr};
    bool qout_zeroed = false;
    bool sweep_attr_set = false, sweep_rec_ready = false;
    int32_t *sweep_rec = nullptr;   // route.cu: static donor records of the tail levels
    int32_t sweep_rec_k0 = 0;
    // route.cu: the narrow end of the level order as heavy-path chains (k_sweep_chains)
    bool chains_ready = false;
    int32_t chain_k0 = 0, n_chains = 0, chain_level0 = 0;
    int64_t n_chain_cells = 0;
    int32_t *chain_start = nullptr, *chain_k = nullptr, *chain_cell = nullptr, *chain_don = nullptr;
    int32_t *chain_ids = nullptr;  // short chains first, then the long ones (a CTA each)
    int32_t n_chains_short = 0, n_chains_long = 0;
    int32_t mid_level0 = -1;       // levels [mid_level0, chain_level0) run in one cooperative launch (k_sweep_mid); -1: none
    int mid_grid = 0;
    double *cut_work = nullptr;    // scan path with lakes/reservoirs: range sums of the cut cells
    int cut_work_cap = 0;
    int32_t *cut_ptr = nullptr;    // per-tile ranges of the cut-cell list
    int32_t *zero_ptr = nullptr;   // [ntiles + 1] zeros: the empty cut list of sphy_route_reslake_begin
    int32_t zero_ptr_cap = 0;
    std::string err;
};

constexpr int ROUTE_MAX_COMP = 8;
constexpr int ROUTE_WIDE_LEVEL = 2048;   // levels with at least this many cells get their own grid launch

#define SPHY_FAIL(ctx, code, ...)                                 \
    do {                                                          \
        char _b[512];                                             \
        snprintf(_b, sizeof(_b), __VA_ARGS__);                    \
        if (ctx) (ctx)->err = _b;                                 \
        return (code);                                            \
    } while (0)

#define SPHY_CUDA(ctx, call)                                                                       \
    do {                                                                                           \
        cudaError_t _e = (call);                                                                   \
        if (_e != cudaSuccess)                                                                     \
            SPHY_FAIL(ctx, SPHY_E_CUDA, "%s failed: %s (%s:%d)", #call, cudaGetErrorString(_e), __FILE__, __LINE__); \
    } while (0)

#define SPHY_LAUNCH_CHECK(ctx) SPHY_CUDA(ctx, cudaGetLastError())

static inline int div_up(int64_t a, int64_t b) { return (int)((a + b - 1) / b); }

// route.cu: the structure table of a lake / reservoir network updated on every rank from the exchange blocks
// (sphy_route_reslake_finish, route_scan.cu)
struct RlPartUpdate {
    const double *gathered;        // [world][rank_pitch] blocks, or NULL: this rank's peer-memory receive buffer (parity by step)
    int64_t rank_pitch;
    int32_t stride, n_extra;       // the extras sit at [stride - n_extra, stride) of every block
    const int32_t *x_rank;         // [n_struct * SPHY_RL_EXTRA] owner of each extra (-1: none)
    int apply_pending;             // StorRES still has to take - Qout + Qin of the previous step
    float vol_factor;
    int doy;
    const long long *p2p_local;
    int p2p_world;
    const double *p2p_recv;
    int64_t p2p_pitch;
};
int reslake_part_update(sphy_ctx *ctx, const sphy_reslake *rl, const RlPartUpdate &u, cudaStream_t st);

template <typename T>
static inline int dev_alloc_once(sphy_ctx *ctx, T **p, size_t count)
{
    if (*p) return SPHY_OK;
    if (count == 0) count = 1;
    SPHY_CUDA(ctx, cudaMalloc((void **)p, count * sizeof(T)));
    return SPHY_OK;
}

template <typename T>
static inline int dev_alloc(sphy_ctx *ctx, T **p, size_t count)
{
    if (*p) { cudaFree(*p); *p = nullptr; }
    if (count == 0) count = 1;
    SPHY_CUDA(ctx, cudaMalloc((void **)p, count * sizeof(T)));
    return SPHY_OK;
}
