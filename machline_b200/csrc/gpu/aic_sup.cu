// supersonic instantiations of the assembly kernel; compiled with -fmad=false (see pair_influence.cuh)
#include "aic_kernels.cuh"

namespace mlgpu {

cudaError_t launch_aic_supersonic(Ctx* c, const AicLaunch& L) {
    switch (L.tile_rows) {
        case 32: return launch_aic_t<true, 32, 64>(c, L);
        case 16: return launch_aic_t<true, 16, 64>(c, L);
        case 8: return launch_aic_t<true, 8, 128>(c, L);
        default: return cudaErrorInvalidValue;
    }
}

}  // namespace mlgpu
