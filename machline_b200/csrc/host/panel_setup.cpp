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
void panel_init(Panel& p, std::vector<Vertex>& verts, int i1, int i2, int i3, int index, bool in_wake) {
    p = Panel();
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

// panel.f90:696-852, lower-order branch (T_mu = S_mu^-1); the quadratic branch is out of the
// round-1 scope and reported as unsupported by Case::init_with_flow.
static void calc_M_mu_transform(Panel& p, bool calc_mirror) {
    double S_mu[9], S_inv[9];
    for (int i = 0; i < 3; ++i) {
        S_mu[3 * i + 0] = 1.;
        if (calc_mirror) {
            S_mu[3 * i + 1] = p.vertices_ls_mir[i][0];
            S_mu[3 * i + 2] = p.vertices_ls_mir[i][1];
        } else {
            S_mu[3 * i + 1] = p.vertices_ls[i][0];
            S_mu[3 * i + 2] = p.vertices_ls[i][1];
        }
    }
    matinv(3, S_mu, S_inv);
    std::vector<double>& T = calc_mirror ? p.T_mu_mir : p.T_mu;
    T.assign(S_inv, S_inv + 9);
}

// panel.f90:544-693
void panel_set_distribution(Panel& p, int order, const std::vector<Panel>& body_panels,
                            const std::vector<Vertex>& body_verts, const std::vector<Vertex>& own_verts,
                            bool mirror_needed, int mirror_plane, bool force_sigma_match) {
    (void)body_panels;
    (void)body_verts;
    (void)mirror_plane;
    (void)force_sigma_match;
    if (p.in_wake) {
        p.order = 1;
        p.has_sources = false;
    } else {
        p.order = order;
    }
    if (p.N_discont_edges == 3 && p.order == 2) p.order = 1;
    if (p.order != 1) throw std::runtime_error("higher-order singularity distributions are not built yet");
    p.mu_dim = 3;
    p.M_dim = 3;
    p.sigma_dim = 1;
    p.S_dim = 1;
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
    }
    calc_M_mu_transform(p, false);
    if (p.has_sources) {
        p.i_panel_s.assign(p.S_dim, -1);  // set_source_panels, panel.f90:668-693
        p.i_panel_s[0] = p.index;
    }
    if (mirror_needed) calc_M_mu_transform(p, true);
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
