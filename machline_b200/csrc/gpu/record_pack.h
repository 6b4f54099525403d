// Host-side packing of one panel image into the record the assembly kernel reads (panel_record.h).
#pragma once
#include <cmath>
#include <cstring>

#include "panel_record.h"

namespace mlgpu {

struct PanelView {   // raw views of an ml_panel_soa (record r = j + img * n_panels)
    int n_panels;
    const double *centr, *A_g_to_ls, *vertices_ls, *n_hat_ls, *b, *sqrt_b, *J, *vert_g, *T_mu;
};

inline void pack_record(double* rec, int stride, bool sup, const PanelView& t, int j, int img, double sigma_val, int flags) {
    const size_t r = (size_t)j + (size_t)img * t.n_panels;
    std::memset(rec, 0, sizeof(double) * stride);
    for (int k = 0; k < 3; ++k) rec[R_CENTR + k] = t.centr[3 * r + k];
    for (int k = 0; k < 9; ++k) rec[R_A + k] = t.A_g_to_ls[9 * r + k];
    for (int k = 0; k < 6; ++k) rec[R_VLS + k] = t.vertices_ls[6 * r + k];
    for (int k = 0; k < 6; ++k) rec[R_NH + k] = t.n_hat_ls[6 * r + k];
    for (int k = 0; k < 9; ++k) rec[R_T + k] = t.T_mu[9 * r + k];
    rec[R_J] = t.J[r];
    rec[R_SIGMA] = sigma_val;
    int fl[2] = {flags, 0};
    if (sup) {
        for (int k = 0; k < 3; ++k) rec[R_B + k] = t.b[3 * r + k];
        for (int k = 0; k < 3; ++k) rec[R_SB + k] = t.sqrt_b[3 * r + k];
        for (int k = 0; k < 9; ++k) rec[R_VG + k] = t.vert_g[9 * r + k];
    } else {
        const double* v = t.vertices_ls + 6 * r;
        rec[R_AREA2] = std::fabs((v[2] - v[0]) * (v[5] - v[1]) - (v[4] - v[0]) * (v[3] - v[1]));
        // near-edge threshold (pair_influence.cuh): (5 % of the longest edge)^2, as a float in the spare int
        double lmax2 = 0.;
        for (int k = 0; k < 3; ++k) {
            const int n = (k + 1) % 3;
            const double ex = v[2 * n] - v[2 * k], ey = v[2 * n + 1] - v[2 * k + 1];
            lmax2 = std::fmax(lmax2, ex * ex + ey * ey);
        }
        const float near2 = (float)(0.0025 * lmax2);
        std::memcpy(&fl[1], &near2, sizeof near2);
    }
    std::memcpy(rec + R_FLAGS, fl, sizeof fl);
}

// Higher-order extension of a record (panel_record.h): T6 = the image's 6 x 6 padded T_mu, w = its T_sigma applied to the
// known strengths of the panel's source panels.
inline void pack_record_ho(double* rec, bool sup, const double* T6, const double w[3]) {
    double* ext = rec + record_ho_offset(sup);
    for (int k = 0; k < 36; ++k) ext[R_HO_T + k] = T6[k];
    for (int k = 0; k < 3; ++k) ext[R_HO_W + k] = w[k];
}

}  // namespace mlgpu
