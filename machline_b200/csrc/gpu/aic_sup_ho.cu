// supersonic higher-order instantiation of the assembly kernel; compiled with -fmad=false as aic_sup.cu (see pair_influence.cuh)
#include "aic_kernels.cuh"

namespace mlgpu {

cudaError_t launch_aic_supersonic_ho(Ctx* c, const AicLaunch& L) {
    if (L.row_nB || L.tile_rows != 8) return cudaErrorInvalidValue;   // Dirichlet rows, 8-row tiles (capi.cu: prepare)
    return launch_aic_t<true, 8, 64, false, true>(c, L);
}

}  // namespace mlgpu
