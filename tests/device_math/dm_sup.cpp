// supersonic device math on the CPU: no contraction, as in aic_sup.cu (-fmad=false <-> g++ -ffp-contract=off)
#define DM_NAME dm_batch_sup_impl
#define DM_SUP true
#include "dm_impl.h"
extern "C" void dm_batch_sup(const ml_flow* fs, const ml_panel_soa* t, int n_pts, const double* pts, double* d, double* s, unsigned char* in) {
    dm_batch_sup_impl(fs, t, n_pts, pts, d, s, in);
}
extern "C" void dm_batch_sup_ho(const ml_flow* fs, const ml_panel_soa* t, int n_pts, const double* pts, double* d, double* s, unsigned char* in) {
    dm_batch_sup_impl_ho(fs, t, n_pts, pts, d, s, in);
}
