// C ABI of the host setup library (include/machline_host.h): flattens the Case object into the
// plain tables that the GPU library takes (ml_flow / ml_panel_soa / ml_system_map), mirroring what
// a Fortran bind(C) shim would do with type(panel) (src/panel.f90:38-70).
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <new>
#include <cstring>
#include <memory>
#include <string>

#include "../../../include/machline_host.h"
#include "model.hpp"
#include "parallel.hpp"

using namespace mlh;

namespace {

// Zero-filled table storage from calloc: for the sizes that matter (220 MB at 280k panels x 2 images) the allocator hands out
// fresh zero pages, so "zero-filling" costs nothing up front and the first touch happens where the records are written -- on
// the host threads (put).  A std::vector would write every byte once serially before that.
template <class T>
struct ZeroBuf {
    T* p = nullptr;
    size_t n = 0;
    ZeroBuf() = default;
    ZeroBuf(const ZeroBuf&) = delete;
    ZeroBuf& operator=(const ZeroBuf&) = delete;
    ~ZeroBuf() { std::free(p); }
    void assign_zero(size_t count) {
        std::free(p);
        p = static_cast<T*>(std::calloc(count ? count : 1, sizeof(T)));
        if (!p) throw std::bad_alloc();
        n = count;
    }
    T* data() { return p; }
    const T* data() const { return p; }
    T& operator[](size_t i) { return p[i]; }
    const T& operator[](size_t i) const { return p[i]; }
};

struct PanelTableStore {
    ZeroBuf<double> centr, A_g_to_ls, vertices_ls, n_hat_ls, b, sqrt_b, J, area, vert_g, T_mu;
    std::vector<int> r, i_vert_d, i_panel_s;
    std::vector<unsigned char> has_sources, image_present;
    int n_panels = 0, n_images = 1, n_cols = 3, in_wake = 0;
    // higher-order tables (include/machline_gpu.h)
    int order2 = 0;
    std::vector<unsigned char> order;
    std::vector<int> M_dim, S_dim, i_panel_s4;
    std::vector<double> T_mu6, T_sigma;

    void reserve_for(int np, int ni, int ncols, int wake, int ho = 0) {
        order2 = ho;
        if (ho) {
            order.assign(np, 1);
            M_dim.assign(np, 3);
            S_dim.assign(np, 1);
            i_panel_s4.assign((size_t)np * 4, -1);
            T_mu6.assign((size_t)np * ni * 36, 0.);
            T_sigma.assign((size_t)np * ni * 12, 0.);
        }
        n_panels = np;
        n_images = ni;
        n_cols = ncols;
        in_wake = wake;
        size_t n_rec = (size_t)np * ni;
        centr.assign_zero(n_rec * 3);
        A_g_to_ls.assign_zero(n_rec * 9);
        vertices_ls.assign_zero(n_rec * 6);
        n_hat_ls.assign_zero(n_rec * 6);
        b.assign_zero(n_rec * 3);
        sqrt_b.assign_zero(n_rec * 3);
        J.assign_zero(n_rec);
        r.assign(n_rec, 1);
        area.assign_zero(np);
        vert_g.assign_zero(n_rec * 9);
        T_mu.assign_zero(n_rec * 9);
        i_vert_d.assign((size_t)np * ncols, -1);
        i_panel_s.assign(np, -1);
        has_sources.assign(np, 0);
        image_present.assign(np, ni > 1 ? 1 : 0);
    }
    void put(int j, const Panel& p, const std::vector<Vertex>& verts, bool with_mirror, int mirror_plane) {
        for (int img = 0; img < (with_mirror ? 2 : 1); ++img) {
            size_t rec = (size_t)j + (size_t)img * n_panels;
            const bool m = img == 1;
            const V3& c = m ? p.centr_mir : p.centr;
            const M33& A = m ? p.A_g_to_ls_mir : p.A_g_to_ls;
            for (int k = 0; k < 3; ++k) centr[rec * 3 + k] = c[k];
            for (int a = 0; a < 3; ++a)
                for (int bb = 0; bb < 3; ++bb) A_g_to_ls[rec * 9 + 3 * a + bb] = A[a][bb];
            for (int k = 0; k < 3; ++k) {
                vertices_ls[rec * 6 + 2 * k + 0] = m ? p.vertices_ls_mir[k][0] : p.vertices_ls[k][0];
                vertices_ls[rec * 6 + 2 * k + 1] = m ? p.vertices_ls_mir[k][1] : p.vertices_ls[k][1];
                n_hat_ls[rec * 6 + 2 * k + 0] = m ? p.n_hat_ls_mir[k][0] : p.n_hat_ls[k][0];
                n_hat_ls[rec * 6 + 2 * k + 1] = m ? p.n_hat_ls_mir[k][1] : p.n_hat_ls[k][1];
                b[rec * 3 + k] = m ? p.b_mir[k] : p.b[k];
                sqrt_b[rec * 3 + k] = m ? p.sqrt_b_mir[k] : p.sqrt_b[k];
                V3 loc = verts[p.iv[k]].loc;
                if (m) loc = mirror_across_plane(loc, mirror_plane);
                for (int cc = 0; cc < 3; ++cc) vert_g[rec * 9 + 3 * k + cc] = loc[cc];
            }
            J[rec] = m ? p.J_mir : p.J;
            r[rec] = m ? p.r_mir : p.r;
            const std::vector<double>& T = m ? p.T_mu_mir : p.T_mu;
            if (p.order == 1)
                for (int k = 0; k < 9 && k < (int)T.size(); ++k) T_mu[rec * 9 + k] = T[k];
            if (order2) {
                for (int a = 0; a < p.mu_dim; ++a)
                    for (int bb = 0; bb < p.M_dim; ++bb) T_mu6[rec * 36 + 6 * a + bb] = T[(size_t)a * p.M_dim + bb];
                const std::vector<double>& Ts = m ? p.T_sigma_mir : p.T_sigma;
                if (p.order == 2 && !Ts.empty()) {
                    for (int a = 0; a < 3; ++a)
                        for (int bb = 0; bb < p.S_dim; ++bb) T_sigma[rec * 12 + 4 * a + bb] = Ts[(size_t)a * p.S_dim + bb];
                } else {
                    T_sigma[rec * 12] = 1.;
                }
            }
        }
        if (order2) {
            order[j] = (unsigned char)p.order;
            M_dim[j] = p.M_dim;
            S_dim[j] = p.i_panel_s.empty() ? 0 : (int)p.i_panel_s.size();
            for (int k = 0; k < 4 && k < (int)p.i_panel_s.size(); ++k) i_panel_s4[(size_t)j * 4 + k] = p.i_panel_s[k];
        }
        area[j] = p.A;
        for (int k = 0; k < n_cols && k < (int)p.i_vert_d.size(); ++k) i_vert_d[(size_t)j * n_cols + k] = p.i_vert_d[k];
        i_panel_s[j] = p.i_panel_s.empty() ? -1 : p.i_panel_s[0];
        has_sources[j] = p.has_sources ? 1 : 0;
    }
    void view(ml_panel_soa* out) const {
        std::memset(out, 0, sizeof *out);
        out->n_panels = n_panels;
        out->n_images = n_images;
        out->n_cols = n_cols;
        out->in_wake = in_wake;
        out->centr = centr.data();
        out->A_g_to_ls = A_g_to_ls.data();
        out->vertices_ls = vertices_ls.data();
        out->n_hat_ls = n_hat_ls.data();
        out->b = b.data();
        out->sqrt_b = sqrt_b.data();
        out->J = J.data();
        out->r = r.data();
        out->area = area.data();
        out->vert_g = vert_g.data();
        out->T_mu = T_mu.data();
        out->i_vert_d = i_vert_d.data();
        out->i_panel_s = i_panel_s.data();
        out->has_sources = has_sources.data();
        out->image_present = image_present.data();
        out->order2 = order2;
        if (order2) {
            out->order = order.data();
            out->M_dim = M_dim.data();
            out->T_mu6 = T_mu6.data();
            out->S_dim = S_dim.data();
            out->i_panel_s4 = i_panel_s4.data();
            out->T_sigma = T_sigma.data();
        }
    }
};

}  // namespace

void mlh::Case::setup() {
    const bool timing = std::getenv("MLH_TIMING") != nullptr;
    auto t0 = std::chrono::steady_clock::now();
    auto lap = [&](const char* what) {
        if (!timing) return;
        auto t1 = std::chrono::steady_clock::now();
        std::fprintf(stderr, "mlh setup: %-16s %8.2f ms\n", what, std::chrono::duration<double, std::milli>(t1 - t0).count());
        t0 = t1;
    };
    init_mesh();
    lap("init_mesh");
    init_with_flow();
    lap("init_with_flow");
    init_solver();
    lap("init_solver");
    pre_solve();
    lap("pre_solve");
}

// tables of ml_post_process (include/machline_gpu.h), one entry per cell
struct PostTableStore {
    std::vector<int> mu_index, sigma_index, force_cell;
    std::vector<double> T_mu, A_g_to_ls, s_dir, sigma_known, v_inner, n_g, area, centr;
};

struct mlh_case {
    Case c;
    PanelTableStore body, wake;
    std::vector<double> cp_loc, cp_n_g;
    std::vector<int> cp_bc, cp_row;
    bool tables_built = false;
    Results last;
    std::vector<double> res_cp, res_v;
    PostTableStore post;
};

static thread_local std::string g_err;

extern "C" const char* mlh_last_error(void) { return g_err.c_str(); }

extern "C" int mlh_case_create(const char* json_text, const char* base_dir, mlh_case** out) {
    if (!json_text || !out) {
        g_err = "null argument";
        return 1;
    }
    try {
        std::unique_ptr<mlh_case> h(new mlh_case());
        h->c.load(json_text, base_dir ? base_dir : "");
        h->c.setup();
        *out = h.release();
        return 0;
    } catch (const std::exception& e) {
        g_err = e.what();
        return 1;
    }
}

extern "C" void mlh_case_destroy(mlh_case* c) { delete c; }

extern "C" int mlh_case_info(const mlh_case* h, mlh_mesh_info* o) {
    if (!h || !o) return 1;
    const Case& c = h->c;
    o->n_body_panels = c.N_panels;
    o->n_body_verts = c.N_verts;
    o->n_wake_panels = c.wake.N_panels;
    o->n_wake_strips = c.wake.N_strips;
    o->n_edges = c.N_edges;
    o->n_cp = c.N_cp;
    o->n_unknown = c.N_unknown;
    o->mirrored = c.mirrored;
    o->asym_flow = c.asym_flow;
    o->mirror_plane = c.mirror_plane;
    o->supersonic = c.freestream.supersonic;
    o->sort_seconds = c.sort_time;
    return 0;
}

static void build_tables(mlh_case* h) {
    Case& c = h->c;
    bool ho = false;
    for (auto& p : c.panels) ho = ho || p.order == 2;
    h->body.reserve_for(c.N_panels, c.mirrored ? 2 : 1, ho ? 6 : 3, 0, ho ? 1 : 0);
    parallel_for(c.N_panels, [&](int j) { h->body.put(j, c.panels[j], c.vertices, c.mirrored, c.mirror_plane); });   // record j only
    // wake: strips flattened in (strip, panel) order, the order of panel_solver.f90:1656-1657
    int nw = c.wake.N_panels;
    bool any_mir = false;
    for (auto& st : c.wake.strips) any_mir = any_mir || st.mirrored;
    h->wake.reserve_for(nw, any_mir ? 2 : 1, 6, 1);
    int j = 0;
    for (auto& st : c.wake.strips)
        for (auto& p : st.panels) {
            h->wake.put(j, p, st.vertices, st.mirrored, c.mirror_plane);
            h->wake.image_present[j] = st.mirrored ? 1 : 0;
            ++j;
        }
    h->cp_loc.assign((size_t)c.N_cp * 3, 0.);
    h->cp_n_g.assign((size_t)c.N_cp * 3, 0.);
    h->cp_bc.assign(c.N_cp, 0);
    h->cp_row.assign(c.N_cp, 0);
    for (int i = 0; i < c.N_cp; ++i) {
        for (int k = 0; k < 3; ++k) {
            h->cp_loc[3 * (size_t)i + k] = c.cp[i].loc[k];
            h->cp_n_g[3 * (size_t)i + k] = c.cp[i].n_g[k];
        }
        h->cp_bc[i] = c.cp[i].bc;
        h->cp_row[i] = c.solver.use_sort_for_cp ? c.P[i] : i;
    }
    h->tables_built = true;
}

extern "C" int mlh_case_tables(mlh_case* h, ml_flow* flow, ml_panel_soa* body, ml_panel_soa* wake, ml_system_map* map,
                               mlh_cp_table* cps) {
    if (!h) return 1;
    try {
        if (!h->tables_built) build_tables(h);
        const Case& c = h->c;
        if (flow) {
            const Flow& f = c.freestream;
            flow->M_inf = f.M_inf;
            flow->B = f.B;
            flow->s = f.s;
            flow->K_inv = f.K_inv;
            for (int i = 0; i < 3; ++i) flow->c_hat_g[i] = f.c_hat_g[i];
            for (int i = 0; i < 3; ++i)
                for (int j = 0; j < 3; ++j) {
                    flow->B_mat_g[3 * i + j] = f.B_mat_g[i][j];
                    flow->C_mat_g[3 * i + j] = f.C_mat_g[i][j];
                }
            flow->supersonic = f.supersonic ? 1 : 0;
            flow->mirror_plane = c.mirrored ? c.mirror_plane : 0;
        }
        if (body) h->body.view(body);
        if (wake) h->wake.view(wake);
        if (map) {
            map->n_cp = c.N_cp;
            map->n_unknown = c.N_unknown;
            map->n_verts = c.N_verts;
            map->n_body_panels = c.N_panels;
            map->n_sigma = c.N_sigma;
            map->mirrored = c.mirrored;
            map->asym_flow = c.asym_flow;
            map->P = c.P.data();
            map->sigma_known = c.sigma_known.data();
            map->i_sigma_in_sys = c.i_sigma_in_sys.data();
            map->sigma = c.sigma.data();
        }
        if (cps) {
            cps->n_cp = c.N_cp;
            cps->loc = h->cp_loc.data();
            cps->bc = h->cp_bc.data();
            cps->n_g = h->cp_n_g.data();
            cps->row_perm = h->cp_row.data();
            cps->BC = c.BC.data();
        }
        return 0;
    } catch (const std::exception& e) {
        g_err = e.what();
        return 1;
    }
}

static int solver_code(const std::string& name) {
    if (name == "LU") return ML_SOLVER_LU;
    if (name == "QRUP") return ML_SOLVER_QRUP;
    if (name == "FQRUP") return ML_SOLVER_FQRUP;
    if (name == "GMRES") return ML_SOLVER_GMRES;
    if (name == "RGMRES") return ML_SOLVER_RGMRES;
    if (name == "PURC") return ML_SOLVER_PURC;
    if (name == "BSSOR") return ML_SOLVER_BSSOR;
    if (name == "BJAC") return ML_SOLVER_BJAC;
    return ML_SOLVER_GMRES;  // invalid name -> GMRES (panel_solver.f90:1969-1973)
}

extern "C" int mlh_case_solver_settings(const mlh_case* h, mlh_solver_settings* o) {
    if (!h || !o) return 1;
    const SolverSettings& s = h->c.solver;
    std::memset(o, 0, sizeof *o);
    o->opts.matrix_solver = solver_code(s.matrix_solver);
    o->opts.preconditioner = (s.preconditioner == "DIAG") ? ML_PREC_DIAG : ML_PREC_NONE;
    o->opts.tol = s.tol;
    o->opts.rel = s.rel;
    o->opts.max_iterations = s.max_iterations;
    o->opts.restart_iterations = s.restart_iterations;
    o->opts.block_size = s.block_size;
    // solver.iterative_solver_output (panel_solver.f90:186); the string lives as long as the case
    o->opts.iteration_file = (s.iteration_file.empty() || s.iteration_file == "none") ? nullptr : s.iteration_file.c_str();
    std::snprintf(o->matrix_solver_name, sizeof o->matrix_solver_name, "%s", s.matrix_solver.c_str());
    std::snprintf(o->formulation, sizeof o->formulation, "%s", s.formulation.c_str());
    o->sort_system = s.sort_system;
    o->write_A_and_b = s.write_A_and_b;
    o->run_checks = h->c.run_checks;
    return 0;
}

extern "C" int mlh_case_inner_points(mlh_case* h, double* pts, int* n_points) {
    if (!h || !n_points) return 1;
    try {
        std::vector<double> p = h->c.inner_points();
        *n_points = (int)(p.size() / 3);
        if (pts) std::memcpy(pts, p.data(), p.size() * sizeof(double));
        return 0;
    } catch (const std::exception& e) {
        g_err = e.what();
        return 1;
    }
}

extern "C" int mlh_case_post(mlh_case* h, const double* x, mlh_results* out) { return mlh_case_post2(h, x, nullptr, out); }

extern "C" int mlh_case_post2(mlh_case* h, const double* x, const double* v_inner, mlh_results* out) {
    if (!h || !x || !out) return 1;
    try {
        const Case& c = h->c;
        std::vector<double> xv(x, x + c.N_unknown);
        h->last = c.post(xv, v_inner);
        const Results& R = h->last;
        const std::vector<double>& rep = c.solver.incompressible_rule ? R.C_p_inc : R.C_p_ise;
        h->res_cp = rep;
        h->res_v.assign((size_t)R.N_cells * 3, 0.);
        for (int i = 0; i < R.N_cells; ++i)
            for (int k = 0; k < 3; ++k) h->res_v[3 * (size_t)i + k] = R.V_cells[i][k];
        out->C_p_max = R.C_p_max;
        out->C_p_min = R.C_p_min;
        for (int k = 0; k < 3; ++k) {
            out->C_F[k] = R.C_F[k];
            out->C_M[k] = R.C_M[k];
        }
        out->n_cells = R.N_cells;
        out->n_mu = (int)R.mu.size();
        out->mu = R.mu.data();
        out->C_p = h->res_cp.data();
        out->V_cells = h->res_v.data();
        return 0;
    } catch (const std::exception& e) {
        g_err = e.what();
        return 1;
    }
}

// Per-rule arrays of the last mlh_case_post / mlh_case_post2: rule = ML_RULE_* (pressure coefficients, [n_cells]) or -1 (the cells'
// force contributions dC_f, [n_cells][3]).  dst may be NULL to query the length; *n = 0 when the rule was not computed.
extern "C" int mlh_case_result_array(mlh_case* h, int rule, double* dst, int* n) {
    if (!h || !n) return 1;
    const Results& R = h->last;
    if (rule == -1) {
        *n = (int)R.dC_f.size() * 3;
        if (dst)
            for (size_t i = 0; i < R.dC_f.size(); ++i)
                for (int k = 0; k < 3; ++k) dst[3 * i + k] = R.dC_f[i][k];
        return 0;
    }
    const std::vector<double>* arrs[ML_RULE_COUNT] = {&R.C_p_inc, &R.C_p_ise, &R.C_p_2nd, &R.C_p_sln, &R.C_p_lin, &R.C_p_pg, &R.C_p_kt, &R.C_p_lai};
    if (rule < 0 || rule >= ML_RULE_COUNT) {
        g_err = "unknown rule";
        return 1;
    }
    *n = (int)arrs[rule]->size();
    if (dst) std::memcpy(dst, arrs[rule]->data(), arrs[rule]->size() * sizeof(double));
    return 0;
}

// The tables and constants ml_post_process takes (lower-order panels only).  v_inner as in mlh_case_post2 (NULL for the Dirichlet
// formulations).  Everything the device kernel reads is prepared with the operations of Case::post, so that both evaluate the same
// arithmetic: mu(i) = x(P(i)) (panel_solver.f90:2018-2020), the mirror shifts of panel_get_velocity_jump (panel.f90:3380-3400,
// 3300-3320), V_inner / U (panel_solver.f90:2070).  Pointers stay valid until the next call or mlh_case_destroy.
extern "C" int mlh_case_post_tables(mlh_case* h, const double* v_inner, ml_post_tables* t, ml_post_flow* f) {
    if (!h || !t || !f) return 1;
    try {
        const Case& c = h->c;
        if (!c.solver.dirichlet && !v_inner)
            throw std::runtime_error("post-processing of a Neumann formulation needs the induced velocities at mlh_case_inner_points");
        for (const Panel& p : c.panels)
            if (p.order != 1) throw std::runtime_error("ml_post_process covers lower-order panels; use mlh_case_post for higher-order distributions");
        const Flow& fs = c.freestream;
        const int Np = c.N_panels, n = c.asym_flow ? 2 * Np : Np;
        PostTableStore& T = h->post;
        T.mu_index.assign((size_t)3 * n, -1);
        T.sigma_index.assign(n, -1);
        T.force_cell.assign(n, 0);
        T.T_mu.assign((size_t)9 * n, 0.);
        T.A_g_to_ls.assign((size_t)9 * n, 0.);
        T.s_dir.assign((size_t)3 * n, 0.);
        T.sigma_known.assign(n, 0.);
        T.v_inner.assign((size_t)3 * n, 0.);
        T.n_g.assign((size_t)3 * n, 0.);
        T.area.assign(n, 0.);
        T.centr.assign((size_t)3 * n, 0.);
        // position in x of body source strength i (unknown sources only): sigma(i_sys_sigma_in_body(k)) = x(P(N_d_unknown + k))
        std::vector<int> sigma_pos(c.sigma.size(), -1);
        for (int k = 0; k < c.N_s_unknown; ++k) sigma_pos[c.i_sys_sigma_in_body[k]] = c.P[c.N_d_unknown + k];
        const int n_mu = c.asym_flow ? 2 * c.N_verts : c.N_verts;
        for (int img = 0; img < (c.asym_flow ? 2 : 1); ++img) {
            const bool mir = img == 1;
            for (int i = 0; i < Np; ++i) {
                const Panel& p = c.panels[i];
                const size_t cell = (size_t)i + (size_t)img * Np;
                if ((int)p.i_vert_d.size() != 3 || p.mu_dim != 3)
                    throw std::runtime_error("ml_post_process: a lower-order panel with other than three doublet vertices");
                for (int k = 0; k < 3; ++k) {
                    const int iv = p.i_vert_d[k];
                    int idx;
                    if (c.asym_flow) idx = mir ? ((iv >= c.N_verts) ? iv - c.N_verts : iv + c.N_verts) : iv;
                    else idx = (iv >= c.N_verts) ? iv - c.N_verts : iv;
                    if (idx < 0 || idx >= n_mu) throw std::runtime_error("ml_post_process: vertex index out of range");
                    T.mu_index[3 * cell + k] = idx < c.N_d_unknown ? c.P[idx] : -1;   // mu beyond the unknowns stays zero (Case::post)
                }
                const std::vector<double>& Tm = mir ? p.T_mu_mir : p.T_mu;
                for (int k = 0; k < 9; ++k) T.T_mu[9 * cell + k] = Tm[k];
                const M33& A = mir ? p.A_g_to_ls_mir : p.A_g_to_ls;
                for (int a = 0; a < 3; ++a)
                    for (int b = 0; b < 3; ++b) T.A_g_to_ls[9 * cell + 3 * a + b] = A[a][b];
                if (p.has_sources) {
                    const V3 s_dir = mir ? p.n_g_mir / inner(p.nu_g_mir, p.n_g_mir) : p.n_g / inner(p.nu_g, p.n_g);
                    for (int k = 0; k < 3; ++k) T.s_dir[3 * cell + k] = s_dir[k];
                    if (p.sigma_dim > 1 || p.i_panel_s.size() != 1) throw std::runtime_error("ml_post_process: linear source distribution on a lower-order panel");
                    const int ip = p.i_panel_s[0];
                    int idx;
                    if (c.asym_flow) idx = mir ? ((ip >= Np) ? ip - Np : ip + Np) : ip;
                    else idx = (ip >= Np) ? ip - Np : ip;
                    if (idx < 0 || idx >= (int)c.sigma.size()) throw std::runtime_error("ml_post_process: source index out of range");
                    T.sigma_index[cell] = sigma_pos[idx];
                    T.sigma_known[cell] = c.sigma[idx];
                }
                // V_cells_inner / U as Case::post forms it (panel_solver.f90:2063-2073)
                V3 V_in = c.inner_flow * fs.U;
                if (!c.solver.dirichlet) {
                    const double* v = v_inner + 3 * cell;
                    V_in = fs.v_inf + fs.U * V3{v[0], v[1], v[2]};
                }
                const V3 vin = V_in / fs.U;
                const V3& ng = mir ? p.n_g_mir : p.n_g;
                const V3& ce = mir ? p.centr_mir : p.centr;
                for (int k = 0; k < 3; ++k) {
                    T.v_inner[3 * cell + k] = vin[k];
                    T.n_g[3 * cell + k] = ng[k];
                    T.centr[3 * cell + k] = ce[k];
                }
                T.area[cell] = p.A;
                T.force_cell[cell] = i;   // the mirrored cell's moment uses the un-mirrored panel's force (panel_solver.f90:2583)
            }
        }
        t->n_cells = n;
        t->mu_index = T.mu_index.data();
        t->T_mu = T.T_mu.data();
        t->A_g_to_ls = T.A_g_to_ls.data();
        t->s_dir = T.s_dir.data();
        t->sigma_index = T.sigma_index.data();
        t->sigma_known = T.sigma_known.data();
        t->v_inner = T.v_inner.data();
        t->n_g = T.n_g.data();
        t->area = T.area.data();
        t->centr = T.centr.data();
        t->force_cell = T.force_cell.data();
        std::memset(f, 0, sizeof *f);
        f->U = fs.U;
        f->U_inv = fs.U_inv;
        f->M_inf = fs.M_inf;
        f->gamma = fs.gamma;
        f->a_ise = fs.a_ise;
        f->b_ise = fs.b_ise;
        f->c_ise = fs.c_ise;
        f->C_P_vac = fs.C_P_vac;
        f->C_P_stag = fs.C_P_stag;
        f->M_inf_corr = c.solver.M_inf_corr;
        for (int k = 0; k < 3; ++k) {
            f->v_inf[k] = fs.v_inf[k];
            f->CG[k] = c.CG[k];
            for (int b = 0; b < 3; ++b) f->A_g_to_c[3 * k + b] = fs.A_g_to_c[k][b];
        }
        f->S_ref = c.S_ref;
        f->l_ref = c.l_ref;
        const SolverSettings& s = c.solver;
        f->rules = (s.incompressible_rule ? 1 << ML_RULE_INCOMPRESSIBLE : 0) | (s.isentropic_rule ? 1 << ML_RULE_ISENTROPIC : 0) |
                   (s.second_order_rule ? 1 << ML_RULE_SECOND_ORDER : 0) | (s.slender_rule ? 1 << ML_RULE_SLENDER_BODY : 0) |
                   (s.linear_rule ? 1 << ML_RULE_LINEAR : 0) | (s.prandtl_glauert ? 1 << ML_RULE_PRANDTL_GLAUERT : 0) |
                   (s.karman_tsien ? 1 << ML_RULE_KARMAN_TSIEN : 0) | (s.laitone ? 1 << ML_RULE_LAITONE : 0);
        f->force_rule = pressure_rule_id(s.pressure_for_forces);
        if (f->force_rule < 0 || !(f->rules & (1 << f->force_rule))) throw std::runtime_error(s.pressure_for_forces + " pressure for forces is not available.");
        f->mirrored_symmetric = (c.mirrored && !c.asym_flow) ? 1 : 0;
        f->mirror_plane = c.mirror_plane;
        return 0;
    } catch (const std::exception& e) {
        g_err = e.what();
        return 1;
    }
}

// Result files (outputs.cpp) of the last mlh_case_post / mlh_case_post2
extern "C" int mlh_case_write_body(mlh_case* h, const char* path, int mirrored) {
    if (!h || !path) return 1;
    try {
        if (h->last.N_cells == 0) throw std::runtime_error("no results yet: call mlh_case_post first");
        write_body_file(h->c, h->last, path, mirrored != 0);
        return 0;
    } catch (const std::exception& e) {
        g_err = e.what();
        return 1;
    }
}

extern "C" int mlh_case_write_wake(mlh_case* h, const char* path, int* exported) {
    if (!h || !path) return 1;
    try {
        if (h->last.N_cells == 0) throw std::runtime_error("no results yet: call mlh_case_post first");
        const bool ok = write_wake_file(h->c, h->last, path);
        if (exported) *exported = ok ? 1 : 0;
        return 0;
    } catch (const std::exception& e) {
        g_err = e.what();
        return 1;
    }
}

extern "C" int mlh_case_write_control_points(mlh_case* h, const char* path, const double* residual) {
    if (!h || !path) return 1;
    try {
        write_control_point_file(h->c, path, residual);
        return 0;
    } catch (const std::exception& e) {
        g_err = e.what();
        return 1;
    }
}

static void add_minmax(FILE* f, const char* label, const std::vector<double>& v, bool& first) {
    if (v.empty()) return;
    double mx = v[0], mn = v[0];
    for (double x : v) {
        if (x > mx) mx = x;
        if (x < mn) mn = x;
    }
    std::fprintf(f, "%s\n        \"%s\": {\n            \"max\": %.16E,\n            \"min\": %.16E\n        }", first ? "" : ",",
                 label, mx, mn);
    first = false;
}

// panel_solver.f90:2618-2746 + main.f90:92-101,140-176
extern "C" int mlh_case_write_report(mlh_case* h, const char* path, const ml_solve_info* info, int solver_stat,
                                     double total_runtime) {
    if (!h || !path) return 1;
    const Case& c = h->c;
    const Results& R = h->last;
    FILE* f = std::fopen(path, "w");
    if (!f) {
        g_err = std::string("cannot write ") + path;
        return 1;
    }
    double l_avg = 0.;
    for (auto& p : c.panels) l_avg = l_avg + std::sqrt(p.A);
    l_avg = l_avg / c.N_panels;
    const double PI = 3.14159265358979323846264338327950288419716939937510;
    std::fprintf(f, "{\n    \"info\": {\n        \"generated_by\": \"machline-b200 (MachLine-compatible report)\"\n    },\n");
    std::fprintf(f,
                 "    \"mesh_info\": {\n        \"N_body_panels\": %d,\n        \"N_body_vertices\": %d,\n        "
                 "\"N_wake_panels\": %d,\n        \"average_characteristic_length\": %.16E,\n        "
                 "\"max_flow_turning_angle\": %.16E\n    },\n",
                 c.N_panels, c.N_verts, c.wake.N_panels, l_avg, std::acos(c.C_min_panel_angle) * 180. / PI);
    std::fprintf(f,
                 "    \"solver_results\": {\n        \"solver_status_code\": %d,\n        \"system_dimension\": %d,\n        "
                 "\"timing\": {\n            \"system_sorting\": %.16E,\n            \"preconditioner\": %.16E,\n            "
                 "\"matrix_solver\": %.16E\n        }",
                 solver_stat, c.N_unknown, c.sort_time, 0.0, info ? info->solve_ms * 1e-3 : 0.0);
    if (solver_stat == 0 && info) {
        if (info->iterations > -1) std::fprintf(f, ",\n        \"iterations\": %d", info->iterations);
        std::fprintf(f, ",\n        \"residual\": {\n            \"max\": %.16E,\n            \"norm\": %.16E\n        }\n    },\n",
                     info->res_max, info->res_norm);
        std::fprintf(f, "    \"pressure_calculations\": {");
        bool first = true;
        add_minmax(f, "incompressible_rule", R.C_p_inc, first);
        add_minmax(f, "isentropic_rule", R.C_p_ise, first);
        add_minmax(f, "second_order_rule", R.C_p_2nd, first);
        add_minmax(f, "slender_body_rule", R.C_p_sln, first);
        add_minmax(f, "linear_rule", R.C_p_lin, first);
        add_minmax(f, "prandtl_glauert", R.C_p_pg, first);
        add_minmax(f, "karman_tsien", R.C_p_kt, first);
        add_minmax(f, "laitone", R.C_p_lai, first);
        std::fprintf(f, "\n    },\n");
        std::fprintf(f, "    \"total_forces\": {\n        \"Cx\": %.16E,\n        \"Cy\": %.16E,\n        \"Cz\": %.16E\n    },\n",
                     R.C_F[0], R.C_F[1], R.C_F[2]);
        std::fprintf(f, "    \"total_moments\": {\n        \"CMx\": %.16E,\n        \"CMy\": %.16E,\n        \"CMz\": %.16E\n    },\n",
                     R.C_M[0], R.C_M[1], R.C_M[2]);
    } else {
        std::fprintf(f, "\n    },\n");
    }
    std::fprintf(f, "    \"total_runtime\": %.16E\n}\n", total_runtime);
    std::fclose(f);
    return 0;
}
