// Post-processing on the device: the lower-order part of panel_solver_calc_cell_velocities,
// calc_pressures, calc_forces and calc_moments (src/panel_solver.f90:2030-2095, 2218-2321, 2440-2528, 2551-2615) from the solution
// the last ml_solve left on the device.
//
//   post_cells_kernel    one thread per cell (= panel image): doublet strengths of the three vertices gathered from x,
//                        mu parameters through T_mu, velocity jump (panel_get_velocity_jump, src/panel.f90:3415-3512), cell
//                        velocity, the selected pressure rules (pressure_rules.hpp, shared with the host library), the cell's
//                        force contribution;
//   post_moments_sums    one CTA: moment contributions (with the reference's use of the un-mirrored panel's force for a mirrored
//                        cell, :2583), then the sums of forces, moments and the pressure extremes in a fixed order (thread-strided
//                        partial sums, one shuffle tree, warp order): deterministic and independent of the grid.
// Built with -fmad=false: the host library is built with -ffp-contract=off, so both evaluate the same IEEE operations
// (tests/test_gpu_post.py compares them; the isentropic rule goes through each side's pow).
#include <cstring>
#include <vector>

#include "../host/pressure_rules.hpp"
#include "ctx.h"

namespace mlgpu {

struct PostArgs {
    int n_cells, n_x;
    const double* x;
    const int* mu_index;
    const double* T_mu;
    const double* A;
    const double* s_dir;
    const int* sigma_index;
    const double* sigma_known;
    const double* v_inner;
    const double* n_g;
    const double* area;
    double U, M_corr;
    int rules, force_rule;
    mlpr::FlowConst fc;
    double* V_cells;                 // [n][3]
    double* C_p[mlpr::RULE_COUNT];   // [n] each (null: not selected)
    double* dC_f;                    // [n][3]
};

__global__ void __launch_bounds__(256) post_cells_kernel(const PostArgs a) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= a.n_cells) return;
    // get_doublet_strengths + mu parameters (panel.f90:3351-3412): mu_params = T_mu mu_verts
    double mu_v[3];
#pragma unroll
    for (int k = 0; k < 3; ++k) {
        const int ix = a.mu_index[3 * i + k];
        mu_v[k] = ix >= 0 ? a.x[ix] : 0.;
    }
    const double* T = a.T_mu + (size_t)9 * i;
    double mu_p[3];
#pragma unroll
    for (int r = 0; r < 3; ++r) {
        double acc = 0.;
#pragma unroll
        for (int k = 0; k < 3; ++k) acc = acc + T[3 * r + k] * mu_v[k];
        mu_p[r] = acc;
    }
    // velocity jump: A^T (mu_x, mu_y, 0) + sigma s_dir (panel.f90:3470-3506)
    const double* A = a.A + (size_t)9 * i;
    const double d0 = mu_p[1], d1 = mu_p[2], d2 = 0.;
    double dv[3];
#pragma unroll
    for (int k = 0; k < 3; ++k) dv[k] = A[k] * d0 + A[3 + k] * d1 + A[6 + k] * d2;
    const double s0 = a.s_dir[3 * i], s1 = a.s_dir[3 * i + 1], s2 = a.s_dir[3 * i + 2];
    if (s0 != 0. || s1 != 0. || s2 != 0.) {
        const int is = a.sigma_index[i];
        const double sg = is >= 0 ? a.x[is] : a.sigma_known[i];
        dv[0] = dv[0] + sg * s0;
        dv[1] = dv[1] + sg * s1;
        dv[2] = dv[2] + sg * s2;
    }
    // cell velocity (panel_solver.f90:2070-2073): U * (V_inner / U + dv)
    double v[3];
#pragma unroll
    for (int k = 0; k < 3; ++k) v[k] = a.U * (a.v_inner[3 * i + k] + dv[k]);
    if (a.V_cells) {
        a.V_cells[3 * i] = v[0];
        a.V_cells[3 * i + 1] = v[1];
        a.V_cells[3 * i + 2] = v[2];
    }
    double cp_force = 0.;
#pragma unroll
    for (int r = 0; r < mlpr::RULE_COUNT; ++r) {
        if (a.rules & (1 << r)) {
            const double cp = mlpr::C_P(a.fc, v, r, a.M_corr);
            if (a.C_p[r]) a.C_p[r][i] = cp;
            if (r == a.force_rule) cp_force = cp;
        }
    }
    // calc_forces :2490-2503: dC_f = -C_p A n_g
    const double f = -cp_force * a.area[i];
    a.dC_f[3 * i] = f * a.n_g[3 * i];
    a.dC_f[3 * i + 1] = f * a.n_g[3 * i + 1];
    a.dC_f[3 * i + 2] = f * a.n_g[3 * i + 2];
}

struct PostSumArgs {
    int n_cells;
    const double* dC_f;
    const double* centr;
    const int* force_cell;
    const double* cp_rep;   // the pressure coefficient whose extremes are reported (may be null)
    double CG[3];
    double* out;            // [8]: C_F sum (3), C_M sum (3), C_p max, C_p min
};

__global__ void __launch_bounds__(1024) post_moments_sums_kernel(const PostSumArgs a) {
    __shared__ double s_part[32][8];
    double acc[6] = {0., 0., 0., 0., 0., 0.};
    double cmax = -1.7976931348623157e308, cmin = 1.7976931348623157e308;
    for (int i = threadIdx.x; i < a.n_cells; i += 1024) {
        const double* f = a.dC_f + (size_t)3 * i;
        acc[0] += f[0];
        acc[1] += f[1];
        acc[2] += f[2];
        const double* fm = a.dC_f + (size_t)3 * a.force_cell[i];   // panel_solver.f90:2575, 2583
        const double r0 = a.centr[3 * i] - a.CG[0], r1 = a.centr[3 * i + 1] - a.CG[1], r2 = a.centr[3 * i + 2] - a.CG[2];
        acc[3] += r1 * fm[2] - r2 * fm[1];
        acc[4] += r2 * fm[0] - r0 * fm[2];
        acc[5] += r0 * fm[1] - r1 * fm[0];
        if (a.cp_rep) {
            const double c = a.cp_rep[i];
            cmax = c > cmax ? c : cmax;
            cmin = c < cmin ? c : cmin;
        }
    }
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
#pragma unroll
        for (int k = 0; k < 6; ++k) acc[k] += __shfl_xor_sync(0xffffffffu, acc[k], o);
        const double om = __shfl_xor_sync(0xffffffffu, cmax, o), on = __shfl_xor_sync(0xffffffffu, cmin, o);
        cmax = om > cmax ? om : cmax;
        cmin = on < cmin ? on : cmin;
    }
    if (lane == 0) {
        for (int k = 0; k < 6; ++k) s_part[warp][k] = acc[k];
        s_part[warp][6] = cmax;
        s_part[warp][7] = cmin;
    }
    __syncthreads();
    if (threadIdx.x < 8) {
        const int k = threadIdx.x;
        double r = s_part[0][k];
        for (int w = 1; w < 32; ++w) {
            const double v = s_part[w][k];
            if (k < 6) r += v;
            else if (k == 6) r = v > r ? v : r;
            else r = v < r ? v : r;
        }
        a.out[k] = r;
    }
}

// host ints / doubles -> device
template <class T>
static cudaError_t upload(Ctx* c, DevBuf<T>& d, const T* h, size_t n) {
    cudaError_t e = d.alloc(n ? n : 1);
    if (e != cudaSuccess || !n) return e;
    c->h2d_bytes += (long long)(n * sizeof(T));
    return cudaMemcpyAsync(d.p, h, n * sizeof(T), cudaMemcpyHostToDevice, c->stream);
}

ml_status post_process(Ctx* c, const ml_post_tables* t, const ml_post_flow* f, const double* x_override, ml_post_out* out) {
    const int n = t->n_cells;
    if (n <= 0) return c->fail(ML_BAD_ARGUMENT, "ml_post_process: no cells");
    if (!t->mu_index || !t->T_mu || !t->A_g_to_ls || !t->s_dir || !t->sigma_index || !t->sigma_known || !t->v_inner || !t->n_g ||
        !t->area || !t->centr || !t->force_cell)
        return c->fail(ML_BAD_ARGUMENT, "ml_post_process: a table pointer is null");
    if (f->force_rule < 0 || f->force_rule >= mlpr::RULE_COUNT || !(f->rules & (1 << f->force_rule)))
        return c->fail(ML_BAD_ARGUMENT, "ml_post_process: the pressure rule of the forces is not among the selected rules");
    if (!x_override && (!c->d_x_last.p || c->n_x_last <= 0))
        return c->fail(ML_BAD_ARGUMENT, "ml_post_process: no solution on the device (call ml_solve first)");
    ML_CUDA(c, cudaSetDevice(c->device));
    DevBuf<double> d_x, d_T, d_A, d_s, d_sk, d_vi, d_n, d_ar, d_ce, d_V, d_f, d_cp[mlpr::RULE_COUNT], d_out;
    DevBuf<int> d_mi, d_si, d_fc;
    const double* x_dev = c->d_x_last.p;
    int n_x = c->n_x_last;
    if (x_override) {
        // the caller's x has the length of the system the tables refer to: the largest index + 1
        int mx = -1;
        for (int i = 0; i < 3 * n; ++i) mx = t->mu_index[i] > mx ? t->mu_index[i] : mx;
        for (int i = 0; i < n; ++i) mx = t->sigma_index[i] > mx ? t->sigma_index[i] : mx;
        n_x = mx + 1;
        ML_CUDA(c, upload(c, d_x, x_override, (size_t)n_x));
        x_dev = d_x.p;
    }
    for (int i = 0; i < 3 * n; ++i)
        if (t->mu_index[i] >= n_x) return c->fail(ML_BAD_ARGUMENT, "ml_post_process: mu_index beyond the solution vector");
    for (int i = 0; i < n; ++i)
        if (t->sigma_index[i] >= n_x || t->force_cell[i] < 0 || t->force_cell[i] >= n)
            return c->fail(ML_BAD_ARGUMENT, "ml_post_process: sigma_index / force_cell out of range");
    ML_CUDA(c, upload(c, d_mi, t->mu_index, (size_t)3 * n));
    ML_CUDA(c, upload(c, d_T, t->T_mu, (size_t)9 * n));
    ML_CUDA(c, upload(c, d_A, t->A_g_to_ls, (size_t)9 * n));
    ML_CUDA(c, upload(c, d_s, t->s_dir, (size_t)3 * n));
    ML_CUDA(c, upload(c, d_si, t->sigma_index, (size_t)n));
    ML_CUDA(c, upload(c, d_sk, t->sigma_known, (size_t)n));
    ML_CUDA(c, upload(c, d_vi, t->v_inner, (size_t)3 * n));
    ML_CUDA(c, upload(c, d_n, t->n_g, (size_t)3 * n));
    ML_CUDA(c, upload(c, d_ar, t->area, (size_t)n));
    ML_CUDA(c, upload(c, d_ce, t->centr, (size_t)3 * n));
    ML_CUDA(c, upload(c, d_fc, t->force_cell, (size_t)n));
    ML_CUDA(c, d_V.alloc((size_t)3 * n));
    ML_CUDA(c, d_f.alloc((size_t)3 * n));
    ML_CUDA(c, d_out.alloc(8));

    PostArgs a;
    std::memset(&a, 0, sizeof a);
    a.n_cells = n;
    a.n_x = n_x;
    a.x = x_dev;
    a.mu_index = d_mi.p;
    a.T_mu = d_T.p;
    a.A = d_A.p;
    a.s_dir = d_s.p;
    a.sigma_index = d_si.p;
    a.sigma_known = d_sk.p;
    a.v_inner = d_vi.p;
    a.n_g = d_n.p;
    a.area = d_ar.p;
    a.U = f->U;
    a.M_corr = f->M_inf_corr;
    a.rules = f->rules;
    a.force_rule = f->force_rule;
    a.fc.U_inv = f->U_inv;
    a.fc.M_inf = f->M_inf;
    a.fc.gamma = f->gamma;
    a.fc.a_ise = f->a_ise;
    a.fc.b_ise = f->b_ise;
    a.fc.c_ise = f->c_ise;
    a.fc.C_P_vac = f->C_P_vac;
    a.fc.C_P_stag = f->C_P_stag;
    for (int k = 0; k < 3; ++k) a.fc.v_inf[k] = f->v_inf[k];
    for (int k = 0; k < 9; ++k) a.fc.A_g_to_c[k] = f->A_g_to_c[k];
    a.V_cells = d_V.p;
    a.dC_f = d_f.p;
    for (int r = 0; r < mlpr::RULE_COUNT; ++r) {
        if (f->rules & (1 << r)) {
            ML_CUDA(c, d_cp[r].alloc((size_t)n));
            a.C_p[r] = d_cp[r].p;
        }
    }
    post_cells_kernel<<<(n + 255) / 256, 256, 0, c->stream>>>(a);
    ML_CUDA(c, cudaGetLastError());
    PostSumArgs s;
    s.n_cells = n;
    s.dC_f = d_f.p;
    s.centr = d_ce.p;
    s.force_cell = d_fc.p;
    // what test/test_machline.py:62-66 reads: the incompressible rule if computed, else the isentropic one
    const int rep = (f->rules & (1 << mlpr::RULE_INCOMPRESSIBLE)) ? mlpr::RULE_INCOMPRESSIBLE
                    : (f->rules & (1 << mlpr::RULE_ISENTROPIC)) ? mlpr::RULE_ISENTROPIC : -1;
    s.cp_rep = rep >= 0 ? d_cp[rep].p : nullptr;
    for (int k = 0; k < 3; ++k) s.CG[k] = f->CG[k];
    s.out = d_out.p;
    post_moments_sums_kernel<<<1, 1024, 0, c->stream>>>(s);
    ML_CUDA(c, cudaGetLastError());
    c->launches += 2;

    double h_out[8];
    ML_CUDA(c, cudaMemcpyAsync(h_out, d_out.p, sizeof h_out, cudaMemcpyDeviceToHost, c->stream));
    long long d2h = sizeof h_out;
    if (out->V_cells) {
        ML_CUDA(c, cudaMemcpyAsync(out->V_cells, d_V.p, (size_t)3 * n * sizeof(double), cudaMemcpyDeviceToHost, c->stream));
        d2h += (long long)3 * n * sizeof(double);
    }
    if (out->dC_f) {
        ML_CUDA(c, cudaMemcpyAsync(out->dC_f, d_f.p, (size_t)3 * n * sizeof(double), cudaMemcpyDeviceToHost, c->stream));
        d2h += (long long)3 * n * sizeof(double);
    }
    for (int r = 0; r < mlpr::RULE_COUNT; ++r) {
        if (out->C_p[r] && (f->rules & (1 << r))) {
            ML_CUDA(c, cudaMemcpyAsync(out->C_p[r], d_cp[r].p, (size_t)n * sizeof(double), cudaMemcpyDeviceToHost, c->stream));
            d2h += (long long)n * sizeof(double);
        }
    }
    ML_CUDA(c, cudaStreamSynchronize(c->stream));
    c->d2h_bytes += d2h;
    // calc_forces :2505-2521, calc_moments :2590-2612
    for (int k = 0; k < 3; ++k) {
        out->C_F[k] = h_out[k] / f->S_ref;
        out->C_M[k] = h_out[3 + k] / f->l_ref;
    }
    if (f->mirrored_symmetric) {
        for (int k = 0; k < 3; ++k) {
            out->C_F[k] = 2. * out->C_F[k];
            if (k == f->mirror_plane - 1) {
                out->C_F[k] = 0.;
                out->C_M[k] = 2. * out->C_M[k];
            } else {
                out->C_M[k] = 0.;
            }
        }
    }
    out->C_p_max = rep >= 0 ? h_out[6] : 0.;
    out->C_p_min = rep >= 0 ? h_out[7] : 0.;
    return ML_OK;
}

}  // namespace mlgpu

extern "C" ml_status ml_post_process(ml_ctx* c, const ml_post_tables* tables, const ml_post_flow* flow, const double* x_override,
                                     ml_post_out* out) {
    if (!c || !tables || !flow || !out) return ML_BAD_ARGUMENT;
    if (c->group) return mlgpu::multi_post_process(c, tables, flow, x_override, out);
    return mlgpu::post_process(c, tables, flow, x_override, out);
}
