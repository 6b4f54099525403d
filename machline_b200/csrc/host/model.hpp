// Host-side data model: the inputs of the two GPU hot paths are produced by the same setup the
// reference performs in src/surface_mesh.f90, src/panel.f90 (flow-dependent precompute),
// src/wake_mesh.f90, src/wake_strip.f90 and src/panel_solver.f90 (init).  This is setup, not the
// hot path: it runs once per case on the host, in the reference's (serial) order.
// Indices are 0-based here; the reference is 1-based.  "none" indices are -1 (reference: 0).
#pragma once
#include <string>
#include <vector>

#include "geom.hpp"
#include "json_min.hpp"
#include "pressure_rules.hpp"

namespace mlh {

// ---- src/flow.f90:11-29 ------------------------------------------------------------------------
struct Flow {
    V3 v_inf{};
    double M_inf = 0, gamma = 1.4, U = 0, U_inv = 0, B = 0, s = 0, c = 0, mu = 0, C_mu = 0, K = 0, K_inv = 0;
    V3 c_hat_g{};
    bool sym_about[3] = {false, false, false};
    M33 B_mat_g{}, B_mat_c{}, B_mat_g_inv{}, C_mat_g{}, C_mat_c{}, A_g_to_c{}, A_c_to_s{}, A_g_to_s{};
    bool supersonic = false, incompressible = false;
    double a_ise = 0, b_ise = 0, c_ise = 0, C_P_vac = 0, C_P_stag = 0;

    void init(const Json& settings, const std::string& spanwise_axis);  // flow.f90:58-147
    bool point_in_dod(const V3& Q, const V3& P) const;                  // flow.f90:282-310
    double C_g_inner(const V3& a, const V3& b) const { return inner(a, matvec(C_mat_g, b)); }
    // pressure rules flow.f90:313-585
    double get_C_P_inc(const V3& v) const;
    double get_C_P_ise(const V3& v) const;
    V3 get_v_pert_c(const V3& v) const;
    double get_C_P_2nd(const V3& v) const;
    double get_C_P_sln(const V3& v) const;
    double get_C_P_lin(const V3& v) const;
    double get_C_P_crit(double M) const;
    void restrict_pressure(double& C_P) const;
    double get_C_P(const V3& v, const std::string& rule, double M_corr) const;
    mlpr::FlowConst pressure_const() const;   // the constants of the rules as pressure_rules.hpp takes them
};
int pressure_rule_id(const std::string& rule);   // mlpr::Rule of a rule name, -1 if unknown

// ---- src/base_geom.f90:26-61 --------------------------------------------------------------------
struct Vertex {
    V3 loc{}, n_g{}, n_g_mir{};
    double l_avg = 0, l_min = 0;
    std::vector<int> adjacent_vertices, adjacent_edges, panels, panels_not_across_wake_edge;
    int N_wake_edges = 0, index = -1, top_parent = -1, bot_parent = -1;
    bool on_mirror_plane = false, clone = false, mirrored_is_unique = true, convex = true;
    int N_needed_clones = 0;
    void init(const V3& l, int idx) {  // base_geom.f90:136-160
        loc = l;
        index = idx;
        top_parent = bot_parent = -1;
        mirrored_is_unique = true;
        clone = false;
        N_needed_clones = 0;
        on_mirror_plane = false;
        N_wake_edges = 0;
    }
};

// ---- src/base_geom.f90:72-94 --------------------------------------------------------------------
struct Edge {
    int top_verts[2] = {-1, -1}, bot_verts[2] = {-1, -1}, panels[2] = {-1, -1};
    int top_midpoint = -1, bot_midpoint = -1;
    int edge_index_for_panel[2] = {-1, -1};
    bool on_mirror_plane = false, sheds_wake = false, discontinuous = false;
    int get_opposing_panel(int i_panel) const {  // base_geom.f90:389-407
        if (i_panel == panels[0]) return panels[1];
        if (i_panel == panels[1]) return panels[0];
        return -1;
    }
    bool touches_vertex(int i_vert) const { return top_verts[0] == i_vert || top_verts[1] == i_vert; }
    void point_top_to_new_vert(int i_orig, int i_new) {  // base_geom.f90:425-441
        if (top_verts[0] == i_orig) top_verts[0] = i_new;
        else if (top_verts[1] == i_orig) top_verts[1] = i_new;
        else if (top_midpoint == i_orig) top_midpoint = i_new;
    }
    void point_bottom_to_new_vert(int i_orig, int i_new) {  // base_geom.f90:444-460
        if (bot_verts[0] == i_orig) bot_verts[0] = i_new;
        else if (bot_verts[1] == i_orig) bot_verts[1] = i_new;
        else if (bot_midpoint == i_orig) bot_midpoint = i_new;
    }
};

// ---- src/base_geom.f90:115-130 ------------------------------------------------------------------
enum { BC_ZERO_POTENTIAL = 1, BC_SF_POTENTIAL = 2, BC_ZERO_NORMAL_MF = 3, BC_STRENGTH_MATCHING = 4,
       BC_ZERO_NORMAL_VEL = 5, BC_ZERO_X_VEL = 6, BC_MF_INNER_FLOW = 7 };
enum { TT_VERTEX = 1, TT_PANEL = 2 };
struct ControlPoint {
    V3 loc{}, n_g{};
    int cp_type = 1, bc = 0;
    bool is_mirror = false;
    int tied_to_type = TT_VERTEX, tied_to_index = -1;
};

// ---- src/panel.f90:38-70 ------------------------------------------------------------------------
struct Panel {
    int N = 3, index = -1;
    int iv[3] = {-1, -1, -1};  // indices into the owning mesh's vertex array (reference: pointers)
    V3 n_g{}, nu_g{}, n_g_mir{}, nu_g_mir{}, centr{}, centr_mir{};
    M33 A_g_to_ls{}, A_ls_to_g{}, A_g_to_ls_mir{}, A_ls_to_g_mir{};
    double vertices_ls[3][2] = {}, vertices_ls_mir[3][2] = {};  // [vertex][xi,eta]
    V3 n_hat_g[3] = {}, n_hat_g_mir[3] = {};
    double n_hat_ls[3][2] = {}, n_hat_ls_mir[3][2] = {};  // [edge][xi,eta]
    double b[3] = {}, sqrt_b[3] = {}, b_mir[3] = {}, sqrt_b_mir[3] = {};
    double A = 0;
    std::vector<double> T_mu, T_mu_mir;        // mu_dim x M_dim row-major
    std::vector<double> T_sigma, T_sigma_mir;  // sigma_dim x S_dim row-major
    bool in_wake = false;
    int abutting_panels[3] = {-1, -1, -1};
    int edges[3] = {-1, -1, -1};
    int r = 1, r_mir = 1;
    double J = 0, J_mir = 0;
    std::vector<int> i_vert_d, i_panel_s;
    int order = 1, N_discont_edges = 0;
    bool edge_is_discontinuous[3] = {false, false, false};
    bool has_sources = true;
    int mu_dim = 3, M_dim = 3, sigma_dim = 1, S_dim = 1;
    // order 2 only (panel.f90:736-743, 1116-1231): inverse of S_mu (6 x 6 row-major) and the C integrals C(0:3,0:3) that
    // integrate a quadratic pressure distribution over the panel
    std::vector<double> S_mu_inv, S_mu_inv_mir;
    double C[4][4] = {}, C_mir[4][4] = {};
};

struct WakeStrip {  // src/wake_strip.f90:10-26
    std::vector<Vertex> vertices;
    std::vector<Panel> panels;
    int N_verts = 0, N_panels = 0;
    bool mirrored = false, on_mirror_plane = false;
    int mirror_plane = 0;
    int i_top_parent_1 = -1, i_top_parent_2 = -1, i_bot_parent_1 = -1, i_bot_parent_2 = -1;
    int i_top_parent = -1, i_bot_parent = -1;
};

struct WakeMesh {  // src/wake_mesh.f90:20-34
    std::vector<WakeStrip> strips;
    int N_strips = 0, N_max_strip_verts = 0, N_max_strip_panels = 0, N_verts = 0, N_panels = 0;
    bool mirrored = false;
    int mirror_plane = 0;
};

struct SolverSettings {  // src/panel_solver.f90:164-310
    std::string formulation = "dirichlet-morino", matrix_solver = "GMRES", preconditioner = "DIAG",
                iteration_file = "none", pressure_for_forces;
    int block_size = -1, max_iterations = 1000, restart_iterations = 20;
    double tol = 1e-12, rel = 0.8, control_point_offset = 1e-7;
    std::string control_point_offset_type = "direct";
    bool sort_system = false, use_sort_for_cp = true, overdetermined_ls = false, underdetermined_ls = false,
         write_A_and_b = false, dirichlet = true;
    bool incompressible_rule = false, isentropic_rule = false, second_order_rule = false, slender_rule = false,
         linear_rule = false, prandtl_glauert = false, karman_tsien = false, laitone = false;
    double M_inf_corr = 0.0;
};

struct Results {
    std::vector<double> mu, sigma, Phi_u;   // Phi_u: total outer surface potential at the vertices (panel_solver.f90:2098-2133)
    std::vector<V3> V_cells, V_cells_inner, dC_f;
    std::vector<double> C_p_inc, C_p_ise, C_p_2nd, C_p_sln, C_p_lin, C_p_pg, C_p_kt, C_p_lai;
    V3 C_F{}, C_M{};
    double C_p_max = 0, C_p_min = 0;  // of the rule test_machline.py reads (incompressible else isentropic)
    int N_cells = 0;
};

// ---- src/surface_mesh.f90:21-109 + the setup half of src/panel_solver.f90 ------------------------
struct Case {
    // inputs
    Json input;
    std::string base_dir;
    bool verbose = false, run_checks = false;
    std::string spanwise_axis = "+y";

    Flow freestream;
    SolverSettings solver;

    // mesh
    std::vector<Vertex> vertices;
    std::vector<Panel> panels;
    std::vector<Edge> edges;
    int N_verts = 0, N_panels = 0, N_edges = 0, N_cp = 0;
    int N_subinc = 0, N_supinc = 0;
    bool mirrored = false, asym_flow = false, found_wake_edges = false;
    int mirror_plane = 0;
    WakeMesh wake;
    double C_wake_shedding_angle = 0, trefftz_distance = -1, C_min_panel_angle = 1, C_max_cont_angle = 0;
    V3 CG{};
    int N_wake_panels_streamwise = 1;
    bool wake_present = true, append_wake = true;
    std::vector<ControlPoint> cp;
    double S_ref = 1, l_ref = 1;
    std::vector<int> vertex_ordering;
    int initial_panel_order = 1;
    std::string singularity_order = "lower";
    bool force_sigma_match = true;

    // solver bookkeeping (panel_solver.f90:29-49)
    int N_unknown = 0, N_d_unknown = 0, N_s_unknown = 0, N_sigma = 0;
    std::vector<unsigned char> sigma_known;
    std::vector<int> i_sigma_in_sys, i_sys_sigma_in_body;
    std::vector<int> P;
    V3 inner_flow{};
    std::vector<double> sigma;  // body%sigma
    std::vector<double> BC;
    double sort_time = 0;

    // ---- setup entry points, in main.f90 order ----
    void load(const std::string& json_text, const std::string& base_dir);
    void init_mesh();       // surface_mesh_init                     surface_mesh.f90:115-158
    void init_with_flow();  // surface_mesh_init_with_flow           surface_mesh.f90:683-756
    void init_solver();     // panel_solver_init minus the DoD pass  panel_solver.f90:115-161
    void pre_solve();       // calc_source_strengths + assemble_BC_vector  panel_solver.f90:1069,1078
    void setup();   // init_mesh, init_with_flow, init_solver, pre_solve (capi.cpp; MLH_TIMING=1 prints the phases)
    // x -> mu, sigma; cell velocities, pressures, forces, moments (panel_solver.f90:2012-2615)
    Results post(const std::vector<double>& x, const double* v_inner = nullptr) const;
    std::vector<double> inner_points() const;

    // ---- pieces (public for tests) ----
    void load_mesh_file(const std::string& file);
    void find_vertices_on_mirror();
    void locate_adjacent_panels();
    void calc_vertex_geometry();
    void init_panels_with_flow();
    void characterize_edges();
    void set_needed_vertex_clones();
    void clone_vertices();
    void init_wake();
    void place_internal_vertex_control_points(double offset, const std::string& offset_type);
    void place_centroid_control_points(double offset);
    void init_neumann();
    void set_permutation();
    bool is_convex_at_vertex(int i_vert) const;
    V3 get_clone_control_point_dir(int i_vert) const;
    bool control_point_outside_mesh(const V3& cp_loc, int i_vert) const;

    V3 vloc(const Panel& p, int k) const { return vertices[p.iv[k]].loc; }
};

// panel-level procedures (src/panel.f90); `verts` is the vertex array the panel indexes into.
void panel_init(Panel& p, std::vector<Vertex>& verts, int i1, int i2, int i3, int index, bool in_wake);
void panel_init_topology(Panel& p, std::vector<Vertex>& verts, int i1, int i2, int i3, int index, bool in_wake, bool reset);   // panel_init without the geometry
void panel_calc_derived_geom(Panel& p, const std::vector<Vertex>& verts);
void panel_init_with_flow(Panel& p, const std::vector<Vertex>& verts, const Flow& fs, bool mirrored, int mirror_plane);
void panel_set_distribution(Panel& p, int order, const std::vector<Panel>& body_panels,
                            const std::vector<Vertex>& body_verts, const std::vector<Vertex>& own_verts,
                            bool mirror_needed, int mirror_plane, bool force_sigma_match);
bool panel_projection_inside(const Panel& p, const std::vector<Vertex>& verts, const V3& point, bool mirrored, int mirror_plane);
bool panel_point_above(const Panel& p, const V3& point, bool mirror_panel);
bool panel_line_passes_through(const Panel& p, const std::vector<Vertex>& verts, const V3& a, const V3& b,
                               bool mirror_panel, int mirror_plane, double& s_star);
void panel_weighted_normal_at_corner(const Panel& p, const std::vector<Vertex>& verts, const V3& vert_loc, quad out[3]);
int panel_get_opposite_vertex(const Panel& p, int i1, int i2);
V3 panel_get_velocity_jump(const Panel& p, const Case& c, const std::vector<double>& mu, const std::vector<double>& sigma,
                           bool mirrored, const V3* point = nullptr);

std::string read_text_file(const std::string& path);

// outputs.cpp: result files in the reference's legacy-VTK layout
void write_body_file(const Case& c, const Results& R, const std::string& path, bool mirror);
bool write_wake_file(const Case& c, const Results& R, const std::string& path);
void write_control_point_file(const Case& c, const std::string& path, const double* residual);

}  // namespace mlh
