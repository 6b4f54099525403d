// TEST INFRASTRUCTURE ONLY: compiles the DEVICE pair evaluation (machline_b200/csrc/gpu/pair_influence.cuh)
// for the CPU so that its arithmetic can be compared with the oracle without a GPU.  The product never loads this.
#include <vector>

#include "../../include/machline_gpu.h"
#include "../../machline_b200/csrc/gpu/pair_influence.cuh"
#include "../../machline_b200/csrc/gpu/record_pack.h"

// out_phi_d[n_pts][n_rec][3], out_phi_s[n_pts][n_rec], out_in[n_pts][n_rec]; records = (j, img) img-major as in ml_panel_soa
static void DM_NAME(const ml_flow* fs, const ml_panel_soa* t, int n_pts, const double* pts, double* out_phi_d, double* out_phi_s,
                    unsigned char* out_in) {
    using namespace mlgpu;
    const bool sup = DM_SUP;
    const int stride = sup ? R_SUP_STRIDE : R_SUB_STRIDE;
    const int n_rec = t->n_panels * t->n_images;
    PanelView v{t->n_panels, t->centr, t->A_g_to_ls, t->vertices_ls, t->n_hat_ls, t->b, t->sqrt_b, t->J, t->vert_g, t->T_mu};
    std::vector<double> recs((size_t)n_rec * stride);
    for (int img = 0; img < t->n_images; ++img)
        for (int j = 0; j < t->n_panels; ++j)
            pack_record(recs.data() + ((size_t)img * t->n_panels + j) * stride, stride, sup, v, j, img, 0., RF_EVAL | (img ? RF_MIRROR : 0));
    FlowConst fc;
    for (int i = 0; i < 3; ++i) fc.c_hat[i] = fs->c_hat_g[i];
    for (int i = 0; i < 9; ++i) fc.C[i] = fs->C_mat_g[i];
    fc.K_inv = fs->K_inv;
    fc.s = (int)fs->s;
    fc.supersonic = fs->supersonic;
#pragma omp parallel for schedule(static)
    for (int p = 0; p < n_pts; ++p) {
        for (int r = 0; r < n_rec; ++r) {
            double ps = 0., pd[3] = {0., 0., 0.};
            const bool ok = pair_influence<DM_SUP>(fc, recs.data() + (size_t)r * stride, pts[3 * p], pts[3 * p + 1], pts[3 * p + 2],
                                                   r >= t->n_panels, ps, pd);
            const size_t o = (size_t)p * n_rec + r;
            out_in[o] = ok;
            out_phi_s[o] = ok ? ps : 0.;
            for (int c = 0; c < 3; ++c) out_phi_d[3 * o + c] = ok ? pd[c] : 0.;
        }
    }
}

// Higher-order tables (t->order2): out_phi_d6[n_pts][n_rec][6], out_phi_s = the pair's contribution to I_known with every
// source strength set to one (= the sum of the oracle's phi_s_S entries).
#define DM_CAT2(a, b) a##b
#define DM_CAT(a, b) DM_CAT2(a, b)
static void DM_CAT(DM_NAME, _ho)(const ml_flow* fs, const ml_panel_soa* t, int n_pts, const double* pts, double* out_phi_d6,
                                 double* out_phi_s, unsigned char* out_in) {
    using namespace mlgpu;
    const bool sup = DM_SUP;
    const int stride = record_stride(sup, true);
    const int n_rec = t->n_panels * t->n_images;
    PanelView v{t->n_panels, t->centr, t->A_g_to_ls, t->vertices_ls, t->n_hat_ls, t->b, t->sqrt_b, t->J, t->vert_g, t->T_mu};
    std::vector<double> recs((size_t)n_rec * stride);
    for (int img = 0; img < t->n_images; ++img)
        for (int j = 0; j < t->n_panels; ++j) {
            const size_t r = (size_t)img * t->n_panels + j;
            double* rec = recs.data() + r * stride;
            pack_record(rec, stride, sup, v, j, img, 0., RF_EVAL | (img ? RF_MIRROR : 0));
            double w[3] = {0., 0., 0.};
            for (int k = 0; k < t->S_dim[j]; ++k)
                for (int a = 0; a < 3; ++a) w[a] += t->T_sigma[r * 12 + 4 * a + k];
            pack_record_ho(rec, sup, t->T_mu6 + 36 * r, w);
        }
    FlowConst fc;
    for (int i = 0; i < 3; ++i) fc.c_hat[i] = fs->c_hat_g[i];
    for (int i = 0; i < 9; ++i) fc.C[i] = fs->C_mat_g[i];
    fc.K_inv = fs->K_inv;
    fc.s = (int)fs->s;
    fc.supersonic = fs->supersonic;
#pragma omp parallel for schedule(static)
    for (int p = 0; p < n_pts; ++p) {
        for (int r = 0; r < n_rec; ++r) {
            double ps = 0., pd[6] = {0., 0., 0., 0., 0., 0.};
            const double* rec = recs.data() + (size_t)r * stride;
            bool ok = true;
            if (DM_SUP) {
                bool e_in[3];
                ok = panel_check_dod(fc, rec, pts[3 * p], pts[3 * p + 1], pts[3 * p + 2], e_in);
                if (ok) pair_eval_supersonic_ho(fc, rec, pts[3 * p], pts[3 * p + 1], pts[3 * p + 2], r >= t->n_panels, e_in, ps, pd);
            } else {
                pair_influence_subsonic_ho(fc, rec, pts[3 * p], pts[3 * p + 1], pts[3 * p + 2], r >= t->n_panels, ps, pd);
            }
            const size_t o = (size_t)p * n_rec + r;
            out_in[o] = ok;
            out_phi_s[o] = ok ? ps : 0.;
            for (int c = 0; c < 6; ++c) out_phi_d6[6 * o + c] = ok ? pd[c] : 0.;
        }
    }
}
