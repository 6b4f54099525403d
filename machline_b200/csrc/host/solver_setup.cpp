// Solver-side setup and post-processing on the host: restatement of the non-hot-path halves of
// src/panel_solver.f90 -- settings (:164-310), Dirichlet init (:313-364, :423-512), control-point
// boundary conditions (:601-648), system permutation (:778-1030), source strengths and BC vector
// (:1104-1200), and after the solve: strengths, cell velocities, pressures, forces, moments
// (:2012-2615) -- plus control-point placement from src/surface_mesh.f90:1473-1892.
// The two hot paths (DoD + influences -> A; the dense solve) are NOT here: they are behind
// include/machline_gpu.h.
#include <algorithm>
#include <chrono>
#include <cmath>
#include <numeric>
#include <stdexcept>

#include "model.hpp"
#include "parallel.hpp"

namespace mlh {

static bool contains(const std::vector<int>& v, int x) { return std::find(v.begin(), v.end(), x) != v.end(); }

// panel_solver.f90:164-310
static void parse_settings(Case& c) {
    static const Json empty_obj = [] {
        Json j;
        j.type = Json::Object;
        return j;
    }();
    const Json* sj = c.input.find("solver");
    const Json* pj = c.input.find("post_processing");
    const Json& s = sj ? *sj : empty_obj;
    const Json& p = pj ? *pj : empty_obj;
    SolverSettings& o = c.solver;
    const Flow& fs = c.freestream;
    o.formulation = s.get("formulation", "dirichlet-morino");
    o.matrix_solver = s.get("matrix_solver", "GMRES");
    o.block_size = s.get("block_size", -1);
    o.tol = s.get("tolerance", 1.e-12);
    o.rel = s.get("relaxation", 0.8);
    o.max_iterations = s.get("max_iterations", 1000);
    o.restart_iterations = s.get("restart_iterations", 20);
    o.preconditioner = s.get("preconditioner", "DIAG");
    o.iteration_file = s.get("iterative_solver_output", "none");
    o.sort_system = s.get("sort_system", fs.supersonic);
    if (o.formulation == "neumann-mass-flux" || o.formulation == "neumann-velocity") {
        o.sort_system = false;
        o.use_sort_for_cp = false;
        o.overdetermined_ls = true;
        o.underdetermined_ls = false;
    } else if (o.formulation == "neumann-doublet-source-mass-flux-ls") {
        o.sort_system = false;
        o.use_sort_for_cp = false;
        o.underdetermined_ls = true;
        o.overdetermined_ls = false;
    } else {
        o.use_sort_for_cp = true;
        o.overdetermined_ls = false;
        o.underdetermined_ls = false;
    }
    o.write_A_and_b = s.get("write_A_and_b", false);
    o.control_point_offset = s.get("control_point_offset", 1.e-7);
    o.control_point_offset_type = s.get("control_point_offset_type", "direct");
    if (o.control_point_offset <= 0.) o.control_point_offset = 1.e-7;

    // parse_processing_settings
    if (fs.M_inf > 0.) {
        o.isentropic_rule = p.get("pressure_rules.isentropic", true);
        o.incompressible_rule = p.get("pressure_rules.incompressible", false);
        if (o.incompressible_rule) {
            o.incompressible_rule = false;
            o.isentropic_rule = true;
        }
    } else {
        o.incompressible_rule = p.get("pressure_rules.incompressible", true);
        o.isentropic_rule = p.get("pressure_rules.isentropic", false);
        if (o.isentropic_rule) {
            o.isentropic_rule = false;
            o.incompressible_rule = true;
        }
    }
    o.second_order_rule = p.get("pressure_rules.second-order", false);
    o.slender_rule = p.get("pressure_rules.slender-body", false);
    o.linear_rule = p.get("pressure_rules.linear", false);
    o.M_inf_corr = p.get("subsonic_pressure_correction.correction_mach_number", 0.0);
    o.prandtl_glauert = p.get("subsonic_pressure_correction.prandtl-glauert", false);
    o.karman_tsien = p.get("subsonic_pressure_correction.karman-tsien", false);
    o.laitone = p.get("subsonic_pressure_correction.laitone", false);
    if (o.M_inf_corr < 0.0 || o.M_inf_corr >= 1.0)
        throw std::runtime_error("The pressure correction Mach number must be between zero and one.");
    bool any_corr = o.prandtl_glauert || o.karman_tsien || o.laitone;
    if (any_corr && fs.M_inf != 0.0)
        throw std::runtime_error("Subsonic pressure corrections require freestream_mach_number = 0.");
    if (any_corr && !o.incompressible_rule) o.incompressible_rule = true;
    const char* dflt = nullptr;
    if (o.incompressible_rule) dflt = "incompressible";
    else if (o.isentropic_rule) dflt = "isentropic";
    else if (o.second_order_rule) dflt = "second-order";
    else if (o.linear_rule) dflt = "linear";
    else if (o.slender_rule) dflt = "slender-body";
    else if (o.prandtl_glauert) dflt = "prandtl-glauert";
    else if (o.karman_tsien) dflt = "karman-tsien";
    else if (o.laitone) dflt = "laitone";
    if (dflt) o.pressure_for_forces = p.get("pressure_for_forces", dflt);
}

// surface_mesh.f90:1473-1680
V3 Case::get_clone_control_point_dir(int i_vert) const {
    const Vertex& v = vertices[i_vert];
    bool found_first = false;
    int i_edge_1 = -1, i_edge_2 = -1;
    for (int i_edge : v.adjacent_edges) {
        const Edge& e = edges[i_edge];
        if (e.sheds_wake) {
            int panel1 = e.panels[0], panel2 = e.panels[1];
            bool in1 = contains(v.panels_not_across_wake_edge, panel1);
            bool in2 = contains(v.panels_not_across_wake_edge, panel2);
            if ((in1 && !in2) || (in2 && !in1)) {
                if (found_first) i_edge_2 = i_edge;
                else {
                    i_edge_1 = i_edge;
                    found_first = true;
                }
            }
        }
    }
    auto opposite_endpoint = [&](const Edge& e) {  // base_geom.f90:481-498
        if (dist(v.loc, vertices[e.top_verts[0]].loc) < 1.e-12) return e.top_verts[1];
        return e.top_verts[0];
    };
    V3 t_avg;
    if (i_edge_2 != -1) {
        V3 t1 = v.loc - vertices[opposite_endpoint(edges[i_edge_1])].loc;
        t1 = t1 / norm2(t1);
        V3 t2 = vertices[opposite_endpoint(edges[i_edge_2])].loc - v.loc;
        t2 = t2 / norm2(t2);
        t_avg = t1 + t2;
        t_avg = t_avg / norm2(t_avg);
    } else {
        t_avg = {0., 0., 0.};
        t_avg[mirror_plane - 1] = 1.;
    }
    bool tp_found = false;
    V3 tp{0., 0., 0.};
    for (int i_panel : v.panels_not_across_wake_edge) {
        const Panel& p = panels[i_panel];
        double l_to_cent = norm2(p.centr - v.loc);
        tp = cross(t_avg, p.n_g);
        tp = tp / norm2(tp);
        if (panel_projection_inside(p, vertices, (0.01 * l_to_cent) * tp + v.loc, false, 0)) {
            tp_found = true;
            break;
        } else if (panel_projection_inside(p, vertices, (-0.01 * l_to_cent) * tp + v.loc, false, 0)) {
            tp_found = true;
            tp = -tp;
            break;
        }
    }
    if (!tp_found) throw std::runtime_error("Failed to find t_p for placing a cloned control point.");
    tp = tp / norm2(tp);
    double C_min = 1.;
    for (size_t j = 0; j < v.panels.size(); ++j) {
        int panel1 = v.panels[j];
        for (size_t k = j + 1; k < v.panels.size(); ++k) {
            double x = inner(panels[panel1].n_g, panels[v.panels[k]].n_g);
            C_min = std::min(C_min, x);
        }
        if (mirrored && v.on_mirror_plane) {
            double x = -panels[panel1].n_g[mirror_plane - 1];
            C_min = std::min(C_min, x);
        }
    }
    V3 n_avg{0., 0., 0.};
    for (int i_panel : v.panels_not_across_wake_edge) {
        quad w[3];
        panel_weighted_normal_at_corner(panels[i_panel], vertices, v.loc, w);
        for (int k = 0; k < 3; ++k) n_avg[k] = (double)((quad)n_avg[k] + w[k]);
    }
    n_avg = n_avg / norm2(n_avg);
    double offset_ratio = 0.5 * std::sqrt(0.5 * (1. + C_min));
    V3 dir = tp - offset_ratio * n_avg;
    if (v.on_mirror_plane) dir[mirror_plane - 1] = 0.;
    dir = dir / norm2(dir);
    return dir;
}

// surface_mesh.f90:1796-1841
bool Case::control_point_outside_mesh(const V3& cp_loc, int i_vert) const {
    const Vertex& v = vertices[i_vert];
    const Panel& p0 = panels[v.panels[0]];
    V3 start = std::sqrt(p0.A) * p0.n_g + p0.centr;
    V3 dir = cp_loc - start;
    int N_crosses = 0;
    for (int i_panel : v.panels) {
        double s_star = 0.;
        if (panel_line_passes_through(panels[i_panel], vertices, start, dir, false, 0, s_star))
            if (s_star <= 1. && s_star >= 0.) ++N_crosses;
        if (v.on_mirror_plane) {
            if (panel_line_passes_through(panels[i_panel], vertices, start, dir, true, mirror_plane, s_star))
                if (s_star <= 1. && s_star >= 0.) ++N_crosses;
        }
    }
    return N_crosses % 2 == 0;
}

// surface_mesh.f90:1683-1769, 1844-1892
void Case::place_internal_vertex_control_points(double offset, const std::string& offset_type) {
    N_cp = asym_flow ? N_verts * 2 : N_verts;
    cp.assign(N_cp, ControlPoint());
    parallel_for(N_verts, [&](int i) {   // reads the mesh, writes cp[i]
        const Vertex& v = vertices[i];
        V3 dir = v.clone ? get_clone_control_point_dir(i) : -v.n_g;
        double this_offset = (offset_type == "local") ? offset * v.l_avg : offset;
        V3 loc = v.loc + this_offset * dir;
        while (control_point_outside_mesh(loc, i)) {
            V3 n_avg{0., 0., 0.};
            for (int i_panel : v.panels) {
                if (panel_point_above(panels[i_panel], loc, false)) {
                    quad w[3];
                    panel_weighted_normal_at_corner(panels[i_panel], vertices, v.loc, w);
                    for (int k = 0; k < 3; ++k) n_avg[k] = (double)((quad)n_avg[k] + w[k]);
                }
            }
            if (v.on_mirror_plane) n_avg[mirror_plane - 1] = 0.;
            if (norm2(n_avg) < 1.e-16) break;
            n_avg = n_avg / norm2(n_avg);
            V3 disp = loc - v.loc;
            V3 new_dir = disp - (1.1 * n_avg) * inner(disp, n_avg);
            new_dir = new_dir / norm2(new_dir);
            loc = v.loc + this_offset * new_dir;
        }
        cp[i].loc = loc;
        cp[i].cp_type = 1;
        cp[i].tied_to_type = TT_VERTEX;
        cp[i].tied_to_index = i;
        cp[i].is_mirror = false;
    });
    if (asym_flow) {
        for (int i = 0; i < N_cp / 2; ++i) {
            ControlPoint& m = cp[i + N_cp / 2];
            m.loc = mirror_across_plane(cp[i].loc, mirror_plane);
            m.cp_type = cp[i].cp_type;
            m.tied_to_type = cp[i].tied_to_type;
            m.tied_to_index = cp[i].tied_to_index;
            m.is_mirror = true;
        }
    }
}

// surface_mesh_place_centroid_control_points (src/surface_mesh.f90:1895-1981): one control point per panel at its centroid,
// moved by `offset` along the panel normal (0: on the surface, > 0: outside); mirrored copies for an asymmetric flow.
void Case::place_centroid_control_points(double offset) {
    if (asym_flow) {
        for (int i = 0; i < N_verts; ++i)
            if (vertices[i].on_mirror_plane && !vertices[i].mirrored_is_unique)
                throw std::runtime_error("Neumann formulations on a mirrored mesh in an asymmetric flow with vertices on the mirror plane "
                                         "(strength-matching control points) are not supported");
    }
    N_cp = asym_flow ? N_panels * 2 : N_panels;
    cp.assign(N_cp, ControlPoint());
    for (int i = 0; i < N_panels; ++i) {
        cp[i].loc = panels[i].centr + panels[i].n_g * offset;      // get_cp_locs_centroid_based, :1772-1793
        cp[i].cp_type = offset == 0. ? 2 : (offset > 0. ? 3 : 1);   // SURFACE / EXTERNAL / INTERNAL
        cp[i].tied_to_type = TT_PANEL;
        cp[i].tied_to_index = i;
        cp[i].is_mirror = false;
    }
    if (asym_flow) {
        for (int i = 0; i < N_panels; ++i) {
            ControlPoint& m = cp[i + N_panels];
            m.loc = mirror_across_plane(cp[i].loc, mirror_plane);
            m.cp_type = cp[i].cp_type;
            m.tied_to_type = TT_PANEL;
            m.tied_to_index = i;
            m.is_mirror = true;
        }
    }
}

// panel_solver.f90:778-1030
void Case::set_permutation() {
    auto t0 = std::chrono::steady_clock::now();
    if (solver.sort_system) {
        std::vector<double> x(N_unknown);
        for (int i = 0; i < N_cp; ++i) {
            V3 loc;
            if (cp[i].is_mirror) {
                if (cp[i].tied_to_type == 1) loc = mirror_across_plane(vertices[cp[i].tied_to_index].loc, mirror_plane);
                else loc = panels[cp[i].tied_to_index].centr_mir;
            } else {
                if (cp[i].tied_to_type == 1) loc = vertices[cp[i].tied_to_index].loc;
                else loc = panels[cp[i].tied_to_index].centr;
            }
            x[i] = -inner(freestream.c_hat_g, loc);
        }
        // insertion_arg_sort (sort.f90:359-388) is a stable ascending sort
        // (keys and indices sorted together: the same order as a stable sort of the indices by x[index], without the scattered
        // reads of the comparator)
        auto stable_arg_sort = [](const std::vector<double>& key, std::vector<int>& order) {
            std::vector<std::pair<double, int>> kv(key.size());
            for (size_t i = 0; i < key.size(); ++i) kv[i] = {key[i], (int)i};
            std::stable_sort(kv.begin(), kv.end(), [](const std::pair<double, int>& a, const std::pair<double, int>& b) { return a.first < b.first; });
            order.resize(key.size());
            for (size_t i = 0; i < key.size(); ++i) order[i] = kv[i].second;
        };
        std::vector<int> P_inv_1;
        stable_arg_sort(x, P_inv_1);

        const double huge = std::numeric_limits<double>::max();
        parallel_for((int)P_inv_1.size(), [&](int i) {   // reads the mesh, writes x[i]
            int i_cp = P_inv_1[i];
            const ControlPoint& q = cp[i_cp];
            const bool mir = q.is_mirror;
            auto key = [&](const V3& l) {
                return -inner(freestream.c_hat_g, mir ? mirror_across_plane(l, mirror_plane) : l);
            };
            if (q.tied_to_type == 1) {
                int i_vert = q.tied_to_index;
                x[i] = huge;
                for (int i_neighbor : vertices[i_vert].adjacent_vertices) x[i] = std::min(x[i], key(vertices[i_neighbor].loc));
                // higher-order distributions reach the vertex opposite each continuous edge as well
                // (panel_solver.f90:859-894 / 931-966)
                for (int i_panel : vertices[i_vert].panels_not_across_wake_edge) {
                    const Panel& pp = panels[i_panel];
                    if (pp.order != 2) continue;
                    int k_v = -1;
                    for (int k = 0; k < 3; ++k)
                        if (pp.iv[k] == i_vert) k_v = k;   // get_opposite_edge, panel.f90:1656-1682
                    if (k_v < 0) continue;
                    const int i_opp_edge = pp.edges[(k_v + 1) % 3];
                    const Edge& e = edges[i_opp_edge];
                    if (e.discontinuous || e.on_mirror_plane) continue;
                    const int i_panel_abutting = (e.panels[0] == i_panel) ? e.panels[1] : e.panels[0];
                    const int i_neighbor = panel_get_opposite_vertex(panels[i_panel_abutting], e.top_verts[0], e.top_verts[1]);
                    if (i_neighbor >= 0) x[i] = std::min(x[i], key(vertices[i_neighbor].loc));
                }
            } else {
                int i_panel = q.tied_to_index;
                x[i] = huge;
                for (int j = 0; j < 3; ++j) x[i] = std::min(x[i], key(vertices[panels[i_panel].iv[j]].loc));
            }
        }, 2048);
        std::vector<int> P_inv_2;
        stable_arg_sort(x, P_inv_2);
        P.assign(N_cp, -1);
        for (int i = 0; i < N_cp; ++i) P[P_inv_1[P_inv_2[i]]] = i;
    } else {
        P.assign(N_unknown, -1);
        for (int i = 0; i < N_verts; ++i) P[i] = vertex_ordering[i];
        int source_start = N_verts;
        if (asym_flow) {
            for (int i = 0; i < N_verts; ++i) P[N_verts + i] = vertex_ordering[i] + N_verts;
            source_start = 2 * N_verts;
        }
        for (int i = 0; i < N_s_unknown; ++i) P[source_start + i] = source_start + i;
    }
    sort_time = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
}

// panel_solver.f90:115-161 without calc_domains_of_dependence (fused into the GPU assembly)
void Case::init_solver() {
    parse_settings(*this);
    const std::string& f = solver.formulation;
    if (f == "dirichlet-morino" || f == "dirichlet-source-free") {
        solver.dirichlet = true;
    } else if (f == "neumann-mass-flux" || f == "neumann-velocity") {
        solver.dirichlet = false;
        init_neumann();
        return;
    } else if (f == "neumann-doublet-source-mass-flux-ls" || f == "neumann-mass-flux-inner-flow" || f == "neumann-doublet-only-mass-flux") {
        throw std::runtime_error("'" + f + "': only the least-squares Neumann formulations neumann-mass-flux and neumann-velocity are built");
    } else {
        throw std::runtime_error("'" + f + "' is not a valid formulation.");
    }
    const bool timing = std::getenv("MLH_TIMING") != nullptr;
    auto t0 = std::chrono::steady_clock::now();
    auto lap = [&](const char* what) {
        if (!timing) return;
        auto t1 = std::chrono::steady_clock::now();
        std::fprintf(stderr, "mlh init_solver: %-24s %8.2f ms\n", what, std::chrono::duration<double, std::milli>(t1 - t0).count());
        t0 = t1;
    };
    // init_dirichlet, panel_solver.f90:313-364
    place_internal_vertex_control_points(solver.control_point_offset, solver.control_point_offset_type);
    lap("control points");
    // set_panel_sources :423-444
    for (auto& p : panels) p.has_sources = (f == "dirichlet-morino") ? true : (p.r < 0);
    // determine_dirichlet_unknowns :447-512
    N_sigma = asym_flow ? N_panels * 2 : N_panels;
    N_d_unknown = asym_flow ? N_verts * 2 : N_verts;
    sigma_known.assign(N_sigma, 1);
    N_s_unknown = N_supinc;
    i_sigma_in_sys.assign(N_sigma, -1);
    i_sys_sigma_in_body.assign(N_s_unknown, -1);
    int j = N_d_unknown;
    for (int i = 0; i < N_panels; ++i) {
        if (panels[i].r < 0) {
            i_sigma_in_sys[i] = j;
            i_sys_sigma_in_body[j - N_d_unknown] = i;
            sigma_known[i] = 0;
            ++j;
        }
    }
    if (asym_flow) {
        for (int i = 0; i < N_panels; ++i) {
            if (panels[i].r_mir < 0) {
                i_sigma_in_sys[i + N_panels] = j;
                i_sys_sigma_in_body[j - N_d_unknown] = i + N_panels;
                sigma_known[i + N_panels] = 0;
                ++j;
            }
        }
    }
    N_unknown = N_d_unknown + N_s_unknown;
    if (N_unknown != N_cp)
        throw std::runtime_error("The number of unknowns is not the same as the number of control points.");
    inner_flow = freestream.c_hat_g;
    if (f == "dirichlet-source-free") inner_flow = inner_flow - matvec(freestream.B_mat_g_inv, freestream.c_hat_g);

    // init_control_point_boundary_conditions :601-648
    for (int i = 0; i < N_cp; ++i) {
        if (cp[i].bc == BC_STRENGTH_MATCHING) continue;
        if (cp[i].is_mirror && cp[i].tied_to_type == TT_VERTEX) {
            if (!vertices[cp[i].tied_to_index].mirrored_is_unique) {
                cp[i].bc = BC_STRENGTH_MATCHING;
                continue;
            }
        }
        cp[i].bc = (f == "dirichlet-morino") ? BC_ZERO_POTENTIAL : BC_SF_POTENTIAL;
    }
    lap("unknowns + conditions");
    set_permutation();
    lap("set_permutation");
}

// init_neumann (panel_solver.f90:367-420) for the two overdetermined least-squares formulations: control points at (above) the
// panel centroids, no sources on subinclined panels (:436-440), one doublet unknown per vertex (:515-563), the panel normal as
// the direction of the condition (:567-598), no sorting (:192-196).
void Case::init_neumann() {
    const std::string& f = solver.formulation;
    place_centroid_control_points(f == "neumann-mass-flux" ? 0. : solver.control_point_offset);
    for (auto& p : panels) p.has_sources = (p.r < 0);
    N_sigma = asym_flow ? N_panels * 2 : N_panels;
    N_d_unknown = asym_flow ? N_verts * 2 : N_verts;
    sigma_known.assign(N_sigma, 1);
    N_s_unknown = 0;
    i_sigma_in_sys.assign(N_sigma, -1);
    i_sys_sigma_in_body.clear();
    N_unknown = N_d_unknown;
    inner_flow = freestream.c_hat_g;
    for (int i = 0; i < N_cp; ++i) {
        cp[i].bc = (f == "neumann-velocity") ? BC_ZERO_NORMAL_VEL : BC_ZERO_NORMAL_MF;
        const Panel& p = panels[cp[i].tied_to_index];
        cp[i].n_g = cp[i].is_mirror ? p.n_g_mir : p.n_g;
    }
    set_permutation();
}

// calc_source_strengths (panel_solver.f90:1162-1200) + assemble_BC_vector (:1104-1159)
void Case::pre_solve() {
    sigma.assign(N_sigma, 0.);
    if (solver.formulation == "dirichlet-morino") {
        for (int i = 0; i < N_panels; ++i) {
            if (sigma_known[i]) sigma[i] = -inner(panels[i].n_g, freestream.c_hat_g);
            if (asym_flow && sigma_known[i + N_panels]) sigma[i + N_panels] = -inner(panels[i].n_g_mir, freestream.c_hat_g);
        }
    }
    BC.assign(N_cp, 0.);
    V3 x = matvec(freestream.B_mat_g_inv, freestream.c_hat_g);
    for (int i = 0; i < N_cp; ++i) {
        int ind = solver.use_sort_for_cp ? P[i] : i;
        switch (cp[i].bc) {
            case BC_SF_POTENTIAL: BC[ind] = -inner(x, cp[i].loc); break;
            case BC_ZERO_NORMAL_MF: BC[ind] = -inner(cp[i].n_g, freestream.c_hat_g); break;
            case BC_MF_INNER_FLOW: BC[ind] = -inner(x, cp[i].loc); break;
            case BC_ZERO_NORMAL_VEL: BC[ind] = -inner(freestream.c_hat_g, cp[i].n_g); break;
            default: BC[ind] = 0.;
        }
    }
}

// panel.f90:3351-3512: velocity jump across a panel at `point` (default: its centroid)
V3 panel_get_velocity_jump(const Panel& p, const Case& c, const std::vector<double>& mu, const std::vector<double>& sigma,
                           bool mirrored, const V3* point) {
    V3 Q_ls{0., 0., 0.};
    if (point) Q_ls = mirrored ? matvec(p.A_g_to_ls_mir, *point - p.centr_mir) : matvec(p.A_g_to_ls, *point - p.centr);
    // get_doublet_strengths, panel.f90:3351-3412 (body panels)
    const int Md = (int)p.i_vert_d.size(), md = p.mu_dim;
    double mu_verts[6] = {0., 0., 0., 0., 0., 0.};
    for (int i = 0; i < Md; ++i) {
        int iv = p.i_vert_d[i];
        int idx;
        if (c.asym_flow) {
            if (mirrored) idx = (iv >= c.N_verts) ? iv - c.N_verts : iv + c.N_verts;
            else idx = iv;
        } else {
            idx = (iv >= c.N_verts) ? iv - c.N_verts : iv;
        }
        mu_verts[i] = mu[idx];
    }
    const std::vector<double>& T = mirrored ? p.T_mu_mir : p.T_mu;
    double mu_params[6] = {0., 0., 0., 0., 0., 0.};
    for (int i = 0; i < md; ++i) {
        double acc = 0.;
        for (int k = 0; k < Md; ++k) acc = acc + T[(size_t)i * Md + k] * mu_verts[k];
        mu_params[i] = acc;
    }
    V3 dv;
    if (p.order == 2) {
        dv = {mu_params[1] + mu_params[3] * Q_ls[0] + mu_params[4] * Q_ls[1],
              mu_params[2] + mu_params[4] * Q_ls[0] + mu_params[5] * Q_ls[1], 0.};
    } else {
        dv = {mu_params[1], mu_params[2], 0.};
    }
    const M33& A = mirrored ? p.A_g_to_ls_mir : p.A_g_to_ls;
    dv = matvec(transpose(A), dv);
    if (p.has_sources) {
        V3 s_dir = mirrored ? p.n_g_mir / inner(p.nu_g_mir, p.n_g_mir) : p.n_g / inner(p.nu_g, p.n_g);
        // get_source_strengths / get_source_parameters, panel.f90:3268-3348
        const int Sd = (int)p.i_panel_s.size();
        double sig[4] = {0., 0., 0., 0.};
        for (int i = 0; i < Sd; ++i) {
            int ip = p.i_panel_s[i];
            int idx;
            if (c.asym_flow) {
                if (mirrored) idx = (ip >= c.N_panels) ? ip - c.N_panels : ip + c.N_panels;
                else idx = ip;
            } else {
                idx = (ip >= c.N_panels) ? ip - c.N_panels : ip;
            }
            sig[i] = sigma[idx];
        }
        double s;
        if (p.sigma_dim > 1) {
            const std::vector<double>& Ts = mirrored ? p.T_sigma_mir : p.T_sigma;
            double sp[3];
            for (int i = 0; i < 3; ++i) {
                double acc = 0.;
                for (int k = 0; k < Sd; ++k) acc = acc + Ts[(size_t)i * Sd + k] * sig[k];
                sp[i] = acc;
            }
            s = sp[0] + sp[1] * Q_ls[0] + sp[2] * Q_ls[1];
        } else {
            s = sig[0];
        }
        dv = dv + s * s_dir;
    }
    return dv;
}

// panel.f90:3541-3606: parameters of the quadratic pressure distribution over an order-2 panel, from the pressure rule applied
// to the velocity at its three vertices and three edge midpoints.  (For the mirrored image the reference evaluates the
// "vertex" velocities at the un-mirrored vertex locations, :3570-3571; restated as it is.)
static void quadratic_pressure_params(const Panel& p, const Case& c, const Results& R, bool mirrored, const V3& inner_flow,
                                      const char* rule, double out[6]) {
    const Flow& fs = c.freestream;
    double C_P[6];
    for (int i = 0; i < 3; ++i) {
        const V3 pt = c.vloc(p, i);
        const V3 dv = panel_get_velocity_jump(p, c, R.mu, R.sigma, mirrored, &pt);
        C_P[i] = fs.get_C_P(fs.U * (inner_flow + dv), rule, c.solver.M_inf_corr);
    }
    for (int i = 0; i < 3; ++i) {
        V3 pt = 0.5 * (c.vloc(p, i) + c.vloc(p, (i + 1) % 3));
        if (mirrored) pt = mirror_across_plane(pt, c.mirror_plane);
        const V3 dv = panel_get_velocity_jump(p, c, R.mu, R.sigma, mirrored, &pt);
        C_P[i + 3] = fs.get_C_P(fs.U * (inner_flow + dv), rule, c.solver.M_inf_corr);
    }
    const std::vector<double>& Si = mirrored ? p.S_mu_inv_mir : p.S_mu_inv;
    for (int i = 0; i < 6; ++i) {
        double acc = 0.;
        for (int k = 0; k < 6; ++k) acc = acc + Si[i * 6 + k] * C_P[k];
        out[i] = acc;
    }
}

// The points just inside every panel (and its mirror image in an asymmetric flow) where calc_cell_velocities evaluates the
// induced velocity for the non-Dirichlet formulations: P = centr - 1e-10 n_g (panel_solver.f90:2063, 2080)
std::vector<double> Case::inner_points() const {
    const int n_cells = asym_flow ? 2 * N_panels : N_panels;
    std::vector<double> pts((size_t)3 * n_cells);
    for (int i = 0; i < N_panels; ++i) {
        V3 P = panels[i].centr - 1.e-10 * panels[i].n_g;
        for (int k = 0; k < 3; ++k) pts[3 * (size_t)i + k] = P[k];
        if (asym_flow) {
            V3 Pm = mirror_across_plane(P, mirror_plane);
            for (int k = 0; k < 3; ++k) pts[3 * (size_t)(i + N_panels) + k] = Pm[k];
        }
    }
    return pts;
}

// panel_solver.f90:2012-2615
Results Case::post(const std::vector<double>& x, const double* v_inner) const {
    // v_inner: [N_cells][3] induced velocity (v_d + v_s, per unit freestream speed) just inside every panel, needed by the
    // formulations without a prescribed inner flow (panel_solver.f90:2063-2066); nullptr for the Dirichlet formulations
    if (!solver.dirichlet && !v_inner)
        throw std::runtime_error("post-processing of a Neumann formulation needs the induced velocities at Case::inner_points()");
    Results R;
    R.mu.assign(asym_flow ? N_verts * 2 : N_verts, 0.);
    for (int i = 0; i < N_d_unknown; ++i) R.mu[i] = x[P[i]];
    R.sigma = sigma;
    for (int i = 0; i < N_s_unknown; ++i) R.sigma[i_sys_sigma_in_body[i]] = x[P[N_d_unknown + i]];

    const Flow& fs = freestream;
    // calc_surface_potentials, panel_solver.f90:2098-2133
    R.Phi_u = R.mu;
    for (int i = 0; i < N_verts; ++i) {
        R.Phi_u[i] = R.Phi_u[i] + inner(inner_flow, vertices[i].loc);
        if (asym_flow) R.Phi_u[i + N_verts] = R.Phi_u[i + N_verts] + inner(inner_flow, mirror_across_plane(vertices[i].loc, mirror_plane));
    }
    for (double& v : R.Phi_u) v = v * fs.U;
    R.N_cells = asym_flow ? 2 * N_panels : N_panels;
    R.V_cells.assign(R.N_cells, V3{});
    R.V_cells_inner.assign(R.N_cells, V3{});
    parallel_for(N_panels, [&](int i) {  // calc_cell_velocities :2030-2095 (a panel writes its own cells)
        R.V_cells_inner[i] = inner_flow * fs.U;
        if (!solver.dirichlet) R.V_cells_inner[i] = fs.v_inf + fs.U * V3{v_inner[3 * i], v_inner[3 * i + 1], v_inner[3 * i + 2]};
        V3 dv = panel_get_velocity_jump(panels[i], *this, R.mu, R.sigma, false);
        R.V_cells[i] = fs.U * ((R.V_cells_inner[i] / fs.U) + dv);
        if (asym_flow) {
            R.V_cells_inner[i + N_panels] = inner_flow * fs.U;
            if (!solver.dirichlet) {
                const double* v = v_inner + 3 * (size_t)(i + N_panels);
                R.V_cells_inner[i + N_panels] = fs.v_inf + fs.U * V3{v[0], v[1], v[2]};
            }
            V3 dvm = panel_get_velocity_jump(panels[i], *this, R.mu, R.sigma, true);
            R.V_cells[i + N_panels] = fs.U * ((R.V_cells_inner[i + N_panels] / fs.U) + dvm);
        }
    }, 1024);
    // calc_pressures :2218-2321; the lower-order average pressure is the rule applied to the
    // centroid velocity (panel.f90:3638-3649), i.e. to V_cells.
    // get_avg_pressure_coef, panel.f90:3609-3675: order 2 integrates the quadratic pressure distribution over the panel
    auto avg_pressure = [&](int i_panel, bool mir, const char* rule) {
        const Panel& p = panels[i_panel];
        const int cell = mir ? i_panel + N_panels : i_panel;
        if (p.order == 1) return fs.get_C_P(R.V_cells[cell], rule, solver.M_inf_corr);
        double q[6];
        quadratic_pressure_params(p, *this, R, mir, R.V_cells_inner[cell] / fs.U, rule, q);
        const double(*C)[4] = mir ? p.C_mir : p.C;
        double avg = C[0][0] * q[0] + C[1][0] * q[1] + C[0][1] * q[2] + 0.5 * C[2][0] * q[3] + C[1][1] * q[4] + 0.5 * C[0][2] * q[5];
        return (mir ? p.J_mir : p.J) * avg / p.A;
    };
    auto fill = [&](std::vector<double>& dst, const char* rule) {
        dst.assign(R.N_cells, 0.);
        parallel_for(N_panels, [&](int i) {
            dst[i] = avg_pressure(i, false, rule);
            if (asym_flow) dst[i + N_panels] = avg_pressure(i, true, rule);
        }, 1024);
    };
    if (solver.incompressible_rule) fill(R.C_p_inc, "incompressible");
    if (solver.isentropic_rule) fill(R.C_p_ise, "isentropic");
    if (solver.second_order_rule) fill(R.C_p_2nd, "second-order");
    if (solver.slender_rule) fill(R.C_p_sln, "slender-body");
    if (solver.linear_rule) fill(R.C_p_lin, "linear");
    if (solver.prandtl_glauert) fill(R.C_p_pg, "prandtl-glauert");
    if (solver.karman_tsien) fill(R.C_p_kt, "karman-tsien");
    if (solver.laitone) fill(R.C_p_lai, "laitone");

    // calc_forces :2440-2528
    const std::vector<double>* pr = nullptr;
    const std::string& pf = solver.pressure_for_forces;
    if (pf == "incompressible") pr = &R.C_p_inc;
    else if (pf == "isentropic") pr = &R.C_p_ise;
    else if (pf == "second-order") pr = &R.C_p_2nd;
    else if (pf == "slender-body") pr = &R.C_p_sln;
    else if (pf == "linear") pr = &R.C_p_lin;
    else if (pf == "prandtl-glauert") pr = &R.C_p_pg;
    else if (pf == "karman-tsien") pr = &R.C_p_kt;
    else if (pf == "laitone") pr = &R.C_p_lai;
    if (!pr || pr->empty()) throw std::runtime_error(pf + " pressure for forces is not available.");
    R.dC_f.assign(R.N_cells, V3{});
    parallel_for(N_panels, [&](int i) {
        R.dC_f[i] = (-(*pr)[i] * panels[i].A) * panels[i].n_g;
        if (asym_flow) R.dC_f[i + N_panels] = (-(*pr)[i + N_panels] * panels[i].A) * panels[i].n_g_mir;
    }, 1024);
    V3 sum{0., 0., 0.};
    for (int i = 0; i < R.N_cells; ++i) sum = sum + R.dC_f[i];
    R.C_F = sum / S_ref;
    if (mirrored && !asym_flow) {
        R.C_F = 2. * R.C_F;
        R.C_F[mirror_plane - 1] = 0.;
    }
    // calc_moments :2551-2615 (order 1: no pressure-variation term)
    V3 msum{0., 0., 0.};
    std::vector<V3> dC_m(R.N_cells, V3{});
    // get_moment_about_centroid, panel.f90:3678-3740 (order 2 only; the reference passes the solver's inner_flow here)
    auto moment_about_centroid = [&](const Panel& p, bool mir) {
        double q[6];
        quadratic_pressure_params(p, *this, R, mir, inner_flow, pf.c_str(), q);
        const double(*C)[4] = mir ? p.C_mir : p.C;
        V3 m{C[1][0] * q[0] + C[2][0] * q[1] + C[1][1] * q[2] + 0.5 * C[3][0] * q[3] + C[2][1] * q[4] + 0.5 * C[1][2] * q[5],
             C[0][1] * q[0] + C[1][1] * q[1] + C[0][2] * q[2] + 0.5 * C[2][1] * q[3] + C[1][2] * q[4] + 0.5 * C[0][3] * q[5], 0.};
        return mir ? p.J_mir * cross(p.n_g_mir, matvec(p.A_ls_to_g_mir, m)) : p.J * cross(p.n_g, matvec(p.A_ls_to_g, m));
    };
    parallel_for(N_panels, [&](int i) {
        dC_m[i] = cross(panels[i].centr - CG, R.dC_f[i]);
        if (panels[i].order == 2) dC_m[i] = dC_m[i] + moment_about_centroid(panels[i], false);
        if (asym_flow) {
            dC_m[i + N_panels] = cross(panels[i].centr_mir - CG, R.dC_f[i]);  // sic (:2583)
            if (panels[i].order == 2) dC_m[i + N_panels] = dC_m[i + N_panels] + moment_about_centroid(panels[i], true);
        }
    }, 1024);
    for (int i = 0; i < R.N_cells; ++i) msum = msum + dC_m[i];
    R.C_M = msum / l_ref;
    if (mirrored && !asym_flow) {
        for (int i = 0; i < 3; ++i) {
            if (i == mirror_plane - 1) R.C_M[i] = 2. * R.C_M[i];
            else R.C_M[i] = 0.;
        }
    }
    // what test/test_machline.py:62-66 reads: incompressible rule if present, else isentropic
    const std::vector<double>& rep = solver.incompressible_rule ? R.C_p_inc : R.C_p_ise;
    if (!rep.empty()) {
        R.C_p_max = *std::max_element(rep.begin(), rep.end());
        R.C_p_min = *std::min_element(rep.begin(), rep.end());
    }
    return R;
}

}  // namespace mlh
