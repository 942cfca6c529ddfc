// route_scan.cu -- K4b: flow accumulation as subtree range sums over a DFS pre-order (placeholder: see DESIGN.md).
#include "common.cuh"

int route_scan_run(sphy_ctx *ctx, int ncomp, const float *const *runoff_mm, float *const *q_state, sphy_param kx,
                   float one_m_kx, cudaStream_t st)
{
    (void)ncomp; (void)runoff_mm; (void)q_state; (void)kx; (void)one_m_kx; (void)st;
    SPHY_FAIL(ctx, SPHY_E_INVALID, "SPHY_ROUTE_SCAN is not built yet");
}
