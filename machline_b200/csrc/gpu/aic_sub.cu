// subsonic instantiations of the assembly kernel; compiled with FMA contraction (see pair_influence.cuh)
#include "aic_kernels.cuh"

namespace mlgpu {

cudaError_t launch_aic_subsonic(Ctx* c, const AicLaunch& L) {
    if (L.row_nB) {   // Neumann rows
        switch (L.tile_rows) {
            case 32: return launch_aic_t<false, 32, 64, true>(c, L);
            case 16: return launch_aic_t<false, 16, 64, true>(c, L);
            case 8: return launch_aic_t<false, 8, 128, true>(c, L);
            case 4: return launch_aic_t<false, 4, 128, true>(c, L);
            default: return cudaErrorInvalidValue;
        }
    }
    switch (L.tile_rows) {
        case 32: return launch_aic_t<false, 32, 64>(c, L);
        case 16: return launch_aic_t<false, 16, 64>(c, L);
        case 8: return launch_aic_t<false, 8, 128>(c, L);
        case 4: return launch_aic_t<false, 4, 128>(c, L);
        default: return cudaErrorInvalidValue;
    }
}

}  // namespace mlgpu
