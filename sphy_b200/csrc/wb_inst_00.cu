// wb_inst_00.cu -- K1 instantiations for SnowFLAG=0, GroundFLAG=0 (see wb_kernel.cuh)
#include "wb_kernel.cuh"

wb::wb_kernel_t wb_pick_00(bool infil, int ra, bool diag, bool usoil, int cpt)
{
    return wb::pick_sg<false, false>(infil, ra, diag, usoil, cpt);
}
