// wb_inst_10.cu -- K1 instantiations for SnowFLAG=1, GroundFLAG=0 (see wb_kernel.cuh)
#include "wb_kernel.cuh"

wb::wb_kernel_t wb_pick_10(bool infil, int ra, bool diag, bool usoil, int cpt)
{
    return wb::pick_sg<true, false>(infil, ra, diag, usoil, cpt);
}
