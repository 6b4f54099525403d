// placeholder: replaced by the blocked LU below in a later step
#include "ctx.h"
namespace mlgpu {
ml_status lu_solve_device(Ctx* c, int, double*, int, const double*, double*) { return c->fail(ML_UNSUPPORTED, "LU not built yet"); }
ml_status block_jacobi_device(Ctx* c, int, const double*, int, const double*, int, double, double, int, int*, double*) {
    return c->fail(ML_UNSUPPORTED, "BJAC not built yet");
}
}
