// Per-panel precompute: host restatement of src/panel.f90:188-1115 (geometry, local-scaled
// transform, edge parameters, mirrored twins, T_mu) and the small geometric predicates
// src/panel.f90:1357-1683.  O(N) setup work that produces the panel table consumed by the GPU
// assembly kernel (ml_panel_soa).
#include <cmath>
#include <cstdio>
#include <stdexcept>

#include "model.hpp"

namespace mlh {

static inline V3 vloc(const Panel& p, const std::vector<Vertex>& verts, int k) { return verts[p.iv[k]].loc; }

// panel.f90:265-363
void panel_calc_derived_geom(Panel& p, const std::vector<Vertex>& verts) {
    // calc_normal
    V3 d1 = vloc(p, verts, 1) - vloc(p, verts, 0);
    V3 d2 = vloc(p, verts, 2) - vloc(p, verts, 1);
    p.n_g = cross(d1, d2);
    p.n_g = p.n_g / norm2(p.n_g);
    // calc_area
    p.A = 0.5 * norm2(cross(d1, d2));
    if (p.A < 1.e-12) {
        char msg[128];
        std::snprintf(msg, sizeof msg, "Panel %d has zero area.", p.index + 1);
        throw std::runtime_error(msg);
    }
    // calc_centroid
    V3 sum{0., 0., 0.};
    for (int i = 0; i < 3; ++i) sum = sum + vloc(p, verts, i);
    p.centr = sum / 3.0;  // sum/this%N with integer N -> real division
    // calc_g_edge_vectors
    for (int i = 0; i < 3; ++i) {
        int i_next = (i + 1) % 3;
        V3 d_g = vloc(p, verts, i_next) - vloc(p, verts, i);
        V3 t_hat_g = d_g / norm2(d_g);
        p.n_hat_g[i] = cross(t_hat_g, p.n_g);
    }
}

// panel.f90:188-237
// the part of panel_init that touches shared state: the panel registers itself with its vertices (in panel order: serial)
void panel_init_topology(Panel& p, std::vector<Vertex>& verts, int i1, int i2, int i3, int index, bool in_wake, bool reset) {
    if (reset) p = Panel();   // false: a freshly constructed panel (the mesh loaders: 1.4 KB per panel not written twice)
    p.N = 3;
    p.iv[0] = i1;
    p.iv[1] = i2;
    p.iv[2] = i3;
    p.index = index;
    for (int i = 0; i < 3; ++i) {
        verts[p.iv[i]].panels.push_back(index);
        verts[p.iv[i]].panels_not_across_wake_edge.push_back(index);
    }
    p.in_wake = in_wake;
    p.has_sources = !in_wake;
}

void panel_init(Panel& p, std::vector<Vertex>& verts, int i1, int i2, int i3, int index, bool in_wake) {
    panel_init_topology(p, verts, i1, i2, i3, index, in_wake, true);
    panel_calc_derived_geom(p, verts);
}

// panel.f90:403-489 (original) and :1004-1066 (mirrored twin)
static void calc_g_to_ls(Panel& p, const std::vector<Vertex>& verts, const Flow& fs, bool mir, int mirror_plane) {
    const V3 n_g = mir ? p.n_g_mir : p.n_g;
    V3 u0, v0;
    if (std::fabs(std::fabs(inner(n_g, fs.c_hat_g)) - 1.) < 1e-12) {
        v0 = vloc(p, verts, 1) - vloc(p, verts, 0);
        if (mir) v0 = mirror_across_plane(v0, mirror_plane);
    } else {
        v0 = cross(n_g, fs.c_hat_g);
    }
    v0 = v0 / norm2(v0);
    u0 = cross(v0, n_g);
    u0 = u0 / norm2(u0);

    V3 nu_g = matvec(fs.B_mat_g, n_g);
    double x = inner(n_g, nu_g);
    int r;
    if (!mir) {
        p.nu_g = nu_g;
        if (fs.supersonic && std::fabs(x) < 1.e-12) {
            char msg[128];
            std::snprintf(msg, sizeof msg, "Panel %d is Mach-inclined, which is not allowed.", p.index + 1);
            throw std::runtime_error(msg);
        }
        r = (int)sign(1., x);
        p.r = r;
        if (p.r == -1) {
            char msg[128];
            std::snprintf(msg, sizeof msg, "Panel %d is superinclined, which is not allowed.", p.index + 1);
            throw std::runtime_error(msg);
        }
    } else {
        p.nu_g_mir = nu_g;
        p.r_mir = (int)sign(1., x);
    }
    // The mirrored transform uses the ORIGINAL panel's r here (panel.f90:1034)
    int rs = (int)(p.r * fs.s);

    double y = 1. / std::sqrt(std::fabs(x));
    M33 A;
    V3 Cu = matvec(fs.C_mat_g, u0), Cv = matvec(fs.C_mat_g, v0);
    for (int j = 0; j < 3; ++j) {
        A[0][j] = y * Cu[j];
        A[1][j] = rs / fs.B * Cv[j];  // (rs/B)*C v0 : integer rs promoted
        A[2][j] = fs.B * y * n_g[j];
    }
    double det = det3(A);
    if (!mir) {
        if (std::fabs(det - fs.B * fs.B) > 1.e-10)
            throw std::runtime_error("Calculation of local scaled coordinate transform failed.");
    } else {
        if (std::fabs(det - fs.B * fs.B) >= 1e-10)
            throw std::runtime_error("Calculation of mirrored local scaled coordinate transform failed.");
    }
    M33 Ainv = (fs.M_inf == 0.) ? transpose(A) : matinv3(A);
    double ci = inner(fs.c_hat_g, n_g);
    double J = 1. / (fs.B * std::sqrt(std::fabs(1. - fs.M_inf * fs.M_inf * (ci * ci))));

    const V3 centr = mir ? p.centr_mir : p.centr;
    for (int i = 0; i < 3; ++i) {
        V3 loc = vloc(p, verts, i);
        if (mir) loc = mirror_across_plane(loc, mirror_plane);
        V3 d = loc - centr;
        double xi = A[0][0] * d[0] + A[0][1] * d[1] + A[0][2] * d[2];
        double eta = A[1][0] * d[0] + A[1][1] * d[1] + A[1][2] * d[2];
        if (mir) {
            p.vertices_ls_mir[i][0] = xi;
            p.vertices_ls_mir[i][1] = eta;
        } else {
            p.vertices_ls[i][0] = xi;
            p.vertices_ls[i][1] = eta;
        }
    }
    if (mir) {
        p.A_g_to_ls_mir = A;
        p.A_ls_to_g_mir = Ainv;
        p.J_mir = J;
    } else {
        p.A_g_to_ls = A;
        p.A_ls_to_g = Ainv;
        p.J = J;
    }
}

// panel.f90:492-541 (original) and :1069-1115 (mirrored: traversal direction flipped)
static void calc_ls_edge_vectors(Panel& p, const Flow& fs, bool mir) {
    double t_hat[3][2];
    for (int i = 0; i < 3; ++i) {
        int i_next = (i + 1) % 3;
        double dx, dy;
        if (!mir) {
            dx = p.vertices_ls[i_next][0] - p.vertices_ls[i][0];
            dy = p.vertices_ls[i_next][1] - p.vertices_ls[i][1];
        } else {
            dx = p.vertices_ls_mir[i][0] - p.vertices_ls_mir[i_next][0];
            dy = p.vertices_ls_mir[i][1] - p.vertices_ls_mir[i_next][1];
        }
        double n = norm2_2(dx, dy);
        t_hat[i][0] = dx / n;
        t_hat[i][1] = dy / n;
    }
    double(*nh)[2] = mir ? p.n_hat_ls_mir : p.n_hat_ls;
    double* b = mir ? p.b_mir : p.b;
    double* sb = mir ? p.sqrt_b_mir : p.sqrt_b;
    int r = mir ? p.r_mir : p.r;
    for (int i = 0; i < 3; ++i) {
        nh[i][0] = t_hat[i][1];
        nh[i][1] = -t_hat[i][0];
    }
    for (int i = 0; i < 3; ++i) {
        if (fs.supersonic) {
            if (r > 0) {
                b[i] = (nh[i][0] - nh[i][1]) * (nh[i][0] + nh[i][1]);
                sb[i] = std::sqrt(std::fabs(b[i]));
            } else {
                b[i] = 1.;
                sb[i] = 1.;
            }
        } else {
            b[i] = -1.;
            sb[i] = 1.;
        }
    }
}

// panel.f90:380-400, 972-1001
void panel_init_with_flow(Panel& p, const std::vector<Vertex>& verts, const Flow& fs, bool mirrored, int mirror_plane) {
    calc_g_to_ls(p, verts, fs, false, mirror_plane);
    calc_ls_edge_vectors(p, fs, false);
    if (mirrored) {
        p.n_g_mir = mirror_across_plane(p.n_g, mirror_plane);
        p.centr_mir = mirror_across_plane(p.centr, mirror_plane);
        calc_g_to_ls(p, verts, fs, true, mirror_plane);
        for (int i = 0; i < 3; ++i) p.n_hat_g_mir[i] = mirror_across_plane(p.n_hat_g[i], mirror_plane);
        calc_ls_edge_vectors(p, fs, true);
    }
}

// panel.f90:1593-1624
int panel_get_opposite_vertex(const Panel& p, int i1, int i2) {
    int i_opp = -1;
    for (int i = 0; i < 3; ++i) {
        if (p.iv[i] == i1) {
            for (int j = 0; j < 3; ++j) {
                if (j != i && p.iv[j] != i2) {
                    i_opp = p.iv[j];
                    break;
                }
            }
            break;
        }
    }
    return i_opp;
}

// Small dense products in the order gfortran's inlined MATMUL accumulates (ascending contracted index).
// All matrices row-major: C(m x n) = A(m x k) B(k x n).
static void mm(int m, int k, int n, const double* A, const double* B, double* C) {
    for (int i = 0; i < m; ++i)
        for (int j = 0; j < n; ++j) {
            double acc = 0.;
            for (int l = 0; l < k; ++l) acc = acc + A[i * k + l] * B[l * n + j];
            C[i * n + j] = acc;
        }
}
static void transpose_mn(int m, int n, const double* A, double* At) {
    for (int i = 0; i < m; ++i)
        for (int j = 0; j < n; ++j) At[j * m + i] = A[i * n + j];
}

// panel.f90:696-852: transformation from the strengths that set the distribution ({M}: the panel's vertices and, for a
// quadratic distribution, the vertices opposite its continuous edges) to the distribution parameters {mu}
static void calc_M_mu_transform(Panel& p, const std::vector<Vertex>& body_verts, bool calc_mirror, int mirror_plane) {
    const int md = p.mu_dim, Md = p.M_dim;
    const int N_body_verts = (int)body_verts.size();
    double S_mu[36] = {0.}, S_inv[36] = {0.};   // md <= 6: on the stack (this runs once or twice per panel, on the host threads)
    const double(*vls)[2] = calc_mirror ? p.vertices_ls_mir : p.vertices_ls;
    for (int i = 0; i < md; ++i) S_mu[i * md + 0] = 1.;
    for (int i = 0; i < 3; ++i) {
        S_mu[i * md + 1] = vls[i][0];
        S_mu[i * md + 2] = vls[i][1];
    }
    if (p.order == 2) {
        for (int i = 0; i < 3; ++i) {   // edge midpoints: 0.5*(x + cshift(x, 1))
            const int n = (i + 1) % 3;
            S_mu[(3 + i) * md + 1] = 0.5 * (S_mu[i * md + 1] + S_mu[n * md + 1]);
            S_mu[(3 + i) * md + 2] = 0.5 * (S_mu[i * md + 2] + S_mu[n * md + 2]);
        }
        for (int i = 0; i < md; ++i) {
            const double x = S_mu[i * md + 1], y = S_mu[i * md + 2];
            S_mu[i * md + 3] = 0.5 * (x * x);
            S_mu[i * md + 4] = x * y;
            S_mu[i * md + 5] = 0.5 * (y * y);
        }
    }
    matinv(md, S_mu, S_inv);
    if (p.order == 2) (calc_mirror ? p.S_mu_inv_mir : p.S_mu_inv).assign(S_inv, S_inv + (size_t)md * md);
    std::vector<double>& T = calc_mirror ? p.T_mu_mir : p.T_mu;
    if (p.order == 1) {
        T.assign(S_inv, S_inv + (size_t)md * md);
        return;
    }
    std::vector<double> M_mat((size_t)md * Md, 0.);
    for (int i = 0; i < 3; ++i) M_mat[i * Md + i] = 1.;
    double E[4 * 6] = {0.};
    for (int i = 0; i < 3; ++i)
        for (int k = 0; k < 6; ++k) E[i * 6 + k] = S_mu[i * md + k];
    int j = 3;
    for (int i = 0; i < 3; ++i) {
        if (p.edge_is_discontinuous[i]) {
            // the midpoint strength is the average of the endpoint strengths
            M_mat[(i + 3) * Md + i] = 0.5;
            M_mat[(i + 3) * Md + (i + 1) % 3] = 0.5;
        } else {
            // underdetermined least-squares fit through the three vertices and the vertex opposite the edge
            const int iv = p.i_vert_d[j];
            V3 P_g;
            if (calc_mirror) {
                if (iv >= N_body_verts) P_g = body_verts[iv - N_body_verts].loc;
                else P_g = mirror_across_plane(body_verts[iv].loc, mirror_plane);
            } else {
                if (iv >= N_body_verts) P_g = mirror_across_plane(body_verts[iv - N_body_verts].loc, mirror_plane);
                else P_g = body_verts[iv].loc;
            }
            const V3 P_ls = calc_mirror ? matvec(p.A_g_to_ls_mir, P_g - p.centr_mir) : matvec(p.A_g_to_ls, P_g - p.centr);
            E[3 * 6 + 0] = 1.;
            E[3 * 6 + 1] = P_ls[0];
            E[3 * 6 + 2] = P_ls[1];
            E[3 * 6 + 3] = 0.5 * (P_ls[0] * P_ls[0]);
            E[3 * 6 + 4] = P_ls[0] * P_ls[1];
            E[3 * 6 + 5] = 0.5 * (P_ls[1] * P_ls[1]);
            double Et[6 * 4], EEt[16], EE_inv[16], EtEEinv[6 * 4], M_row[4];
            transpose_mn(4, 6, E, Et);
            mm(4, 6, 4, E, Et, EEt);
            matinv(4, EEt, EE_inv);
            mm(6, 4, 4, Et, EE_inv, EtEEinv);
            mm(1, 6, 4, &S_mu[(i + 3) * md], EtEEinv, M_row);
            for (int k = 0; k < 3; ++k) M_mat[(i + 3) * Md + k] = M_row[k];
            M_mat[(i + 3) * Md + j] = M_row[3];
            ++j;
        }
    }
    T.assign((size_t)md * Md, 0.);
    mm(md, md, Md, S_inv, M_mat.data(), T.data());
}

// panel.f90:855-969: transformation from the source strengths of this panel and its neighbours across continuous edges ({S})
// to the parameters of the linear source distribution {sigma}
static void calc_S_sigma_transform(Panel& p, const std::vector<Panel>& body_panels, bool calc_mirror, int mirror_plane,
                                   bool force_sigma_match) {
    if (p.order != 2) return;
    const int Sd = p.S_dim, N_panels = (int)body_panels.size();
    std::vector<double> S((size_t)Sd * 3, 0.);
    S[0] = 1.;
    V3 P_ls{0., 0., 0.};
    for (int i = 1; i < Sd; ++i) {
        const int ip = p.i_panel_s[i];
        V3 P_g;
        if (calc_mirror) {
            if (ip >= N_panels) P_g = body_panels[ip - N_panels].centr;
            else P_g = body_panels[ip].centr_mir;
            P_ls = matvec(p.A_g_to_ls_mir, P_g - p.centr_mir);
        } else {
            if (ip >= N_panels) P_g = mirror_across_plane(body_panels[ip - N_panels].centr, mirror_plane);
            else P_g = body_panels[ip].centr;
            P_ls = matvec(p.A_g_to_ls, P_g - p.centr);
        }
        S[i * 3 + 0] = 1.;
        S[i * 3 + 1] = P_ls[0];
        S[i * 3 + 2] = P_ls[1];
    }
    std::vector<double> T((size_t)3 * Sd, 0.);
    if (Sd == 4) {
        if (!force_sigma_match) {
            double St[3 * 4], StS[9], SS_inv[9];
            transpose_mn(4, 3, S.data(), St);
            mm(3, 4, 3, St, S.data(), StS);
            matinv(3, StS, SS_inv);
            mm(3, 3, 4, SS_inv, St, T.data());
        } else {
            double Sub[3 * 2], Subt[2 * 3], StS[4], SS_inv[4], A_mat[3 * 4] = {0.}, StA[2 * 4], T23[2 * 4];
            for (int i = 0; i < 3; ++i) {
                Sub[i * 2 + 0] = S[(i + 1) * 3 + 1];
                Sub[i * 2 + 1] = S[(i + 1) * 3 + 2];
            }
            transpose_mn(3, 2, Sub, Subt);
            mm(2, 3, 2, Subt, Sub, StS);
            matinv(2, StS, SS_inv);
            for (int i = 0; i < 3; ++i) A_mat[i * 4 + 0] = -1.;
            A_mat[0 * 4 + 1] = 1.;
            A_mat[1 * 4 + 2] = 1.;
            A_mat[2 * 4 + 3] = 1.;
            mm(2, 3, 4, Subt, A_mat, StA);
            mm(2, 2, 4, SS_inv, StA, T23);
            T[0] = 1.;
            for (int k = 0; k < 4; ++k) {
                T[1 * 4 + k] = T23[k];
                T[2 * 4 + k] = T23[4 + k];
            }
        }
    } else if (Sd == 3) {
        matinv(3, S.data(), T.data());
    } else if (Sd == 2) {
        double A_mat[3 * 2] = {0.}, SA[4], SA_inv[4];
        A_mat[0] = 1.;
        if (std::fabs(P_ls[0]) > std::fabs(P_ls[1])) {
            A_mat[1 * 2 + 1] = 1.;
            A_mat[2 * 2 + 1] = P_ls[1] / P_ls[0];
        } else {
            A_mat[1 * 2 + 1] = P_ls[0] / P_ls[1];
            A_mat[2 * 2 + 1] = 1.;
        }
        mm(2, 3, 2, S.data(), A_mat, SA);
        matinv(2, SA, SA_inv);
        mm(3, 2, 2, A_mat, SA_inv, T.data());
    }
    (calc_mirror ? p.T_sigma_mir : p.T_sigma) = T;
}

// panel.f90:1116-1231: C(i,j) = integral of xi^i eta^j over the panel, from the edge recursions (H by eta, I by xi)
static void calc_C_integrals(Panel& p, bool mir) {
    const int Ni = 3, Nj = 3;
    double d_eta[3], d_xi[3], xi[3], eta[3];
    const double(*v)[2] = mir ? p.vertices_ls_mir : p.vertices_ls;
    for (int k = 0; k < 3; ++k) {
        const int kn = (k + 1) % 3;
        xi[k] = v[k][0];
        eta[k] = v[k][1];
        if (!mir) {
            d_xi[k] = v[kn][0] - v[k][0];
            d_eta[k] = v[kn][1] - v[k][1];
        } else {
            d_xi[kn] = v[k][0] - v[kn][0];
            d_eta[kn] = v[k][1] - v[kn][1];
        }
    }
    static thread_local double II[Ni + 3][Nj + 1][Ni + 2][3];
    for (int i = 0; i <= Ni + 2; ++i)
        for (int e = 0; e < 3; ++e) II[i][0][0][e] = 1. / (i + 1.);
    for (int j = 1; j <= Nj; ++j)
        for (int i = 0; i <= Ni - j; ++i)
            for (int e = 0; e < 3; ++e) II[i][j][0][e] = eta[e] * II[i][j - 1][0][e] + d_eta[e] * II[i + 1][j - 1][0][e];
    for (int j = 0; j <= Nj; ++j)
        for (int k = 1; k <= Ni - j; ++k)
            for (int i = Ni - j; i >= k; --i)
                for (int e = 0; e < 3; ++e) II[i][j][k][e] = xi[e] * II[i - 1][j][k - 1][e] + d_xi[e] * II[i][j][k - 1][e];
    double(*C)[4] = mir ? p.C_mir : p.C;
    for (int i = 0; i <= Ni; ++i)
        for (int j = 0; j <= Nj; ++j) {
            // the reference reads II(i+1, j, i+1, :) for every (i, j) of the 4 x 4 table, including entries its recursions
            // never set (i + j > 2: allocated, not initialised there); only C(i,j) with i + j <= 3 and in fact <= 2 for the
            // pressure average are used.  Entries outside the recursion's range are reported as zero here.
            double acc = 0.;
            const bool set = (i + 1 <= Ni - j);
            if (set)
                for (int e = 0; e < 3; ++e) acc = acc + d_eta[e] * II[i + 1][j][i + 1][e];
            C[i][j] = set ? acc / (i + 1) : 0.;
        }
}

// panel.f90:544-693
void panel_set_distribution(Panel& p, int order, const std::vector<Panel>& body_panels,
                            const std::vector<Vertex>& body_verts, const std::vector<Vertex>& own_verts,
                            bool mirror_needed, int mirror_plane, bool force_sigma_match) {
    if (p.in_wake) {
        p.order = 1;
        p.has_sources = false;
    } else {
        p.order = order;
    }
    if (p.N_discont_edges == 3 && p.order == 2) p.order = 1;
    if (p.order == 1) {
        p.mu_dim = 3;
        p.M_dim = 3;
        p.sigma_dim = 1;
        p.S_dim = 1;
    } else {
        p.mu_dim = 6;
        p.M_dim = 6 - p.N_discont_edges;
        p.sigma_dim = 3;
        p.S_dim = 4 - p.N_discont_edges;
    }
    const int N_body_panels = (int)body_panels.size(), N_body_verts = (int)body_verts.size();
    // set_doublet_verts, panel.f90:605-665
    if (p.in_wake) {
        p.i_vert_d.assign(2 * p.M_dim, -1);
        for (int i = 0; i < 3; ++i) {
            p.i_vert_d[i] = own_verts[p.iv[i]].top_parent;
            p.i_vert_d[i + p.M_dim] = own_verts[p.iv[i]].bot_parent;
        }
    } else {
        p.i_vert_d.assign(p.M_dim, -1);
        for (int i = 0; i < 3; ++i) p.i_vert_d[i] = p.iv[i];
        if (p.order == 2) {
            int j = 3;
            for (int i = 0; i < 3; ++i) {
                if (p.edge_is_discontinuous[i]) continue;
                const int i1 = p.iv[i], i2 = p.iv[(i + 1) % 3];
                if (p.abutting_panels[i] >= N_body_panels) {   // the neighbour is this panel's own mirror image
                    p.i_vert_d[j] = panel_get_opposite_vertex(p, i1, i2) + N_body_verts;
                } else {
                    p.i_vert_d[j] = panel_get_opposite_vertex(body_panels[p.abutting_panels[i]], i1, i2);
                }
                ++j;
            }
        }
    }
    calc_M_mu_transform(p, body_verts, false, mirror_plane);
    if (p.has_sources) {
        p.i_panel_s.assign(p.S_dim, -1);  // set_source_panels, panel.f90:668-693
        p.i_panel_s[0] = p.index;
        if (p.order == 2) {
            int j = 1;
            for (int i = 0; i < 3; ++i)
                if (!p.edge_is_discontinuous[i]) p.i_panel_s[j++] = p.abutting_panels[i];
        }
        calc_S_sigma_transform(p, body_panels, false, mirror_plane, force_sigma_match);
    }
    if (mirror_needed) {
        calc_M_mu_transform(p, body_verts, true, mirror_plane);
        if (p.has_sources) calc_S_sigma_transform(p, body_panels, true, mirror_plane, force_sigma_match);
    }
    if (p.order == 2) {
        calc_C_integrals(p, false);
        if (mirror_needed) calc_C_integrals(p, true);
    }
}

// panel.f90:1357-1401
bool panel_projection_inside(const Panel& p, const std::vector<Vertex>& verts, const V3& point, bool mirrored,
                             int mirror_plane) {
    for (int i = 0; i < 3; ++i) {
        double x;
        if (mirrored) {
            V3 d = point - mirror_across_plane(vloc(p, verts, i), mirror_plane);
            x = inner(d, p.n_hat_g_mir[i]);
        } else {
            V3 d = point - vloc(p, verts, i);
            x = inner(d, p.n_hat_g[i]);
        }
        if (x >= 1.e-16) return false;
    }
    return true;
}

// panel.f90:1437-1465
bool panel_point_above(const Panel& p, const V3& point, bool mirror_panel) {
    double h = mirror_panel ? inner(point - p.centr_mir, p.n_g_mir) : inner(point - p.centr, p.n_g);
    return !(h < 0.);
}

// panel.f90:1468-1510
bool panel_line_passes_through(const Panel& p, const std::vector<Vertex>& verts, const V3& a, const V3& b,
                               bool mirror_panel, int mirror_plane, double& s_star) {
    double d = mirror_panel ? inner(b, p.n_g_mir) : inner(b, p.n_g);
    if (std::fabs(d) < 1.e-16) return false;
    if (mirror_panel) s_star = inner(p.centr_mir - a, p.n_g_mir) / d;
    else s_star = inner(p.centr - a, p.n_g) / d;
    V3 loc = a + s_star * b;
    return panel_projection_inside(p, verts, loc, mirror_panel, mirror_plane);
}

// panel.f90:1513-1569: the corner angle is computed in binary64 and widened; the weighted normal
// n_g*W is a binary128 product.
void panel_weighted_normal_at_corner(const Panel& p, const std::vector<Vertex>& verts, const V3& vert_loc, quad out[3]) {
    quad W = 0;
    for (int i = 0; i < 3; ++i) {
        if (dist(vloc(p, verts, i), vert_loc) < 1.e-12) {
            int i_prev = (i == 0) ? 2 : i - 1;
            double angle = std::acos(inner(-p.n_hat_g[i], p.n_hat_g[i_prev]));
            W = (quad)angle;
            break;
        }
    }
    for (int k = 0; k < 3; ++k) out[k] = (quad)p.n_g[k] * W;
}

}  // namespace mlh
