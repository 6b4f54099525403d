// subsonic device math on the CPU: FMA contraction on, as in aic_sub.cu (g++ -mfma -ffp-contract=fast)
#define DM_NAME dm_batch_sub_impl
#define DM_SUP false
#include "dm_impl.h"
extern "C" void dm_batch_sub(const ml_flow* fs, const ml_panel_soa* t, int n_pts, const double* pts, double* d, double* s, unsigned char* in) {
    dm_batch_sub_impl(fs, t, n_pts, pts, d, s, in);
}
extern "C" void dm_batch_sub_ho(const ml_flow* fs, const ml_panel_soa* t, int n_pts, const double* pts, double* d, double* s, unsigned char* in) {
    dm_batch_sub_impl_ho(fs, t, n_pts, pts, d, s, in);
}
