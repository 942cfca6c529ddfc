// wb_inst_11.cu -- K1 instantiations for SnowFLAG=1, GroundFLAG=1 (see wb_kernel.cuh)
#include "wb_kernel.cuh"

wb::wb_kernel_t wb_pick_11(bool infil, int ra, bool diag, bool usoil, int cpt)
{
    return wb::pick_sg<true, true>(infil, ra, diag, usoil, cpt);
}
