// wb_inst_01.cu -- K1 instantiations for SnowFLAG=0, GroundFLAG=1 (see wb_kernel.cuh)
#include "wb_kernel.cuh"

wb::wb_kernel_t wb_pick_01(bool infil, int ra, bool diag, bool usoil, int cpt)
{
    return wb::pick_sg<false, true>(infil, ra, diag, usoil, cpt);
}
