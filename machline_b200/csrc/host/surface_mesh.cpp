// Body-mesh setup: host restatement of src/surface_mesh.f90:115-1470 (settings, adjacency, vertex
// normals, edge characterisation, vertex cloning at wake-shedding edges, convexity) in the
// reference's serial order.  The all-pairs neighbour search (surface_mesh.f90:366-403) is replaced
// by a vertex->panel lookup that visits the same candidate pairs in the same (i asc, j asc) order
// (SURVEY F10 / App. A.18).
#include <algorithm>
#include <chrono>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <stdexcept>

#include "model.hpp"
#include "parallel.hpp"

namespace mlh {

static const double PI = 3.14159265358979323846264338327950288419716939937510;

static bool contains(const std::vector<int>& v, int x) { return std::find(v.begin(), v.end(), x) != v.end(); }
static void erase_all(std::vector<int>& v, int x) { v.erase(std::remove(v.begin(), v.end(), x), v.end()); }

void Case::load(const std::string& json_text, const std::string& dir) {
    input = JsonParser(json_text).parse();
    base_dir = dir;
    const Json* solver_j = input.find("solver");
    const Json* output_j = input.find("output");
    run_checks = solver_j ? solver_j->get("run_checks", false) : false;
    verbose = output_j ? output_j->get("verbose", true) : true;
}

// surface_mesh.f90:115-158 (+ parse_* :161-326)
void Case::init_mesh() {
    const Json* g = input.find("geometry");
    if (!g) throw std::runtime_error("input has no geometry section");
    std::string mesh_file = g->get("file", "");
    // parse_singularity_settings
    singularity_order = g->get("singularity_order", "lower");
    if (singularity_order == "lower") initial_panel_order = 1;
    else if (singularity_order == "higher") initial_panel_order = 2;
    else if (singularity_order == "adaptive") initial_panel_order = 1;
    else {
        singularity_order = "lower";
        initial_panel_order = 1;
    }
    double discont_angle = g->get("max_continuity_angle", 5.);
    C_max_cont_angle = std::cos(PI * discont_angle / 180.);
    force_sigma_match = g->get("force_sigma_match", true);

    const bool timing = std::getenv("MLH_TIMING") != nullptr;
    auto t0 = std::chrono::steady_clock::now();
    auto lap = [&](const char* what) {
        if (!timing) return;
        auto t1 = std::chrono::steady_clock::now();
        std::fprintf(stderr, "mlh init_mesh: %-24s %8.2f ms\n", what, std::chrono::duration<double, std::milli>(t1 - t0).count());
        t0 = t1;
    };
    load_mesh_file(mesh_file);
    lap("load_mesh_file");

    // parse_mirror_settings
    std::string mp = g->get("mirror_about", "none");
    mirrored = true;
    if (mp == "xy") mirror_plane = 3;
    else if (mp == "xz") mirror_plane = 2;
    else if (mp == "yz") mirror_plane = 1;
    else {
        mirror_plane = 0;
        mirrored = false;
    }
    // parse_wake_settings
    wake_present = g->get("wake_model.wake_present", true);
    append_wake = g->get("wake_model.append_wake", wake_present);
    if (!wake_present && append_wake) append_wake = false;
    if (wake_present) {
        double wake_shedding_angle = g->get("wake_model.wake_shedding_angle", 90.);
        C_wake_shedding_angle = std::cos(wake_shedding_angle * PI / 180.);
        if (append_wake) {
            trefftz_distance = g->get("wake_model.trefftz_distance", -1.);
            N_wake_panels_streamwise = g->get("wake_model.N_panels", 1);
        }
    }
    S_ref = g->get("reference.area", 1.);
    l_ref = g->get("reference.length", 1.);
    const Json* cg = g->path("reference.CG");
    CG = {0., 0., 0.};
    if (cg && cg->type == Json::Array && cg->arr.size() == 3)
        for (int i = 0; i < 3; ++i) CG[i] = cg->arr[i].num;

    if (mirrored) find_vertices_on_mirror();
    locate_adjacent_panels();
    lap("locate_adjacent_panels");
    calc_vertex_geometry();
    lap("calc_vertex_geometry");
}

// surface_mesh.f90:329-343, base_geom.f90:217-234
void Case::find_vertices_on_mirror() {
    for (auto& v : vertices)
        if (std::fabs(v.loc[mirror_plane - 1]) < 1e-12) v.on_mirror_plane = true;
}

namespace {
struct EdgeRec {
    int panel1, panel2, vertex1, vertex2, edge_index1, edge_index2;
    bool on_mirror_plane;
};
}  // namespace

// surface_mesh.f90:346-520
// The reference loops over all pairs of panels (O(N^2)); the outcome -- which panels abut across which edge, and the numbering of
// the edges in the order of discovery -- is reproduced in three steps:
//   1 (host threads; reads only): for every panel the panels j > i that share at least two vertex indices with it, ascending -- up
//     to NB_MAX of them in a flat table (a manifold triangle has at most three neighbours; a panel with more candidates, e.g. at
//     a non-manifold edge, is flagged and searched again in step 2);
//   2 (serial, in panel order as the reference's loop, :346-496): the adjacency checks on compact arrays (vertex indices and
//     abutting panels of all panels, 12 bytes per panel each: the 1.4 KB Panel objects and the vertices' little heap arrays are not
//     touched), which number the edges;
//   3 (host threads): the results go into the panels, the edges and the vertices' adjacency lists; the lists are built from the
//     edge records by a counting sort, in edge order, as the appends of a serial loop leave them (base_geom.f90:163-214).
void Case::locate_adjacent_panels() {
    std::vector<EdgeRec> recs;
    recs.reserve((size_t)N_panels * 3 / 2 + 16);
    std::vector<int> ivs((size_t)N_panels * 3), abut((size_t)N_panels * 3, -1);
    parallel_for(N_panels, [&](int i) {
        for (int m = 0; m < 3; ++m) ivs[3 * (size_t)i + m] = panels[i].iv[m];
    }, 4096);
    std::vector<char> on_mirror(N_verts, 0);
    for (int v = 0; v < N_verts; ++v) on_mirror[v] = vertices[v].on_mirror_plane ? 1 : 0;
    auto all_found = [&](int i) { return abut[3 * (size_t)i] != -1 && abut[3 * (size_t)i + 1] != -1 && abut[3 * (size_t)i + 2] != -1; };
    // panel_check_abutting_panel? no: surface_mesh.f90:523-607 (check_panels_adjacent) on the compact arrays
    auto check_adjacent = [&](int i, int j, int ep[2], int& edge_index_i, int& edge_index_j) {
        bool already_found_shared = false;
        int m1 = -1, n1 = -1;
        const int* pi = &ivs[3 * (size_t)i];
        const int* pj = &ivs[3 * (size_t)j];
        for (int m = 0; m < 3; ++m) {
            for (int n = 0; n < 3; ++n) {
                if (pi[m] == pj[n]) {
                    if (already_found_shared) {
                        ep[1] = pi[m];
                        if (m1 == 0 && m == 2) std::swap(ep[0], ep[1]);
                        if ((n1 == 0 && n == 2) || (n == 0 && n1 == 2)) {
                            abut[3 * (size_t)j + 2] = i;
                            edge_index_j = 2;
                        } else {
                            n1 = std::min(n, n1);
                            abut[3 * (size_t)j + n1] = i;
                            edge_index_j = n1;
                        }
                        if (m1 == 0 && m == 2) {
                            abut[3 * (size_t)i + m] = j;
                            edge_index_i = m;
                        } else {
                            abut[3 * (size_t)i + m1] = j;
                            edge_index_i = m1;
                        }
                        return true;
                    } else {
                        already_found_shared = true;
                        ep[0] = pi[m];
                        m1 = m;
                        n1 = n;
                    }
                }
            }
        }
        return false;
    };
    // panel.f90:1293-1354 (check_abutting_mirror_plane) on the compact arrays
    auto check_mirror = [&](int i, int ep[2], int& edge_index) {
        bool already_found = false;
        int m1 = -1;
        const int* pi = &ivs[3 * (size_t)i];
        for (int m = 0; m < 3; ++m) {
            if (on_mirror[pi[m]]) {
                if (already_found) {
                    ep[1] = pi[m];
                    if (m1 == 0 && m == 2) std::swap(ep[0], ep[1]);
                    if (m - m1 == 1) {
                        abut[3 * (size_t)i + m1] = panels[i].index + N_panels;
                        edge_index = m1;
                    } else {
                        abut[3 * (size_t)i + m] = panels[i].index + N_panels;
                        edge_index = m;
                    }
                    return true;
                } else {
                    already_found = true;
                    ep[0] = pi[m];
                    m1 = m;
                }
            }
        }
        return false;
    };
    // ---- step 1 ----
    constexpr int NB_MAX = 4;
    std::vector<int> nb((size_t)N_panels * NB_MAX);
    std::vector<unsigned char> nb_n(N_panels, 0);
    auto candidates_of = [&](int i, std::vector<int>& cand, std::vector<int>& out) {
        cand.clear();
        out.clear();
        for (int m = 0; m < 3; ++m)
            for (int j : vertices[ivs[3 * (size_t)i + m]].panels)
                if (j > i) cand.push_back(j);
        std::sort(cand.begin(), cand.end());
        for (size_t c = 0; c < cand.size();) {
            size_t e = c;
            while (e < cand.size() && cand[e] == cand[c]) ++e;
            if (e - c >= 2) out.push_back(cand[c]);
            c = e;
        }
    };
    {
        const int nt = std::max(1, host_threads());
        std::vector<std::vector<int>> scratch_a(nt), scratch_b(nt);
        const int chunk = (N_panels + nt - 1) / nt;
        parallel_for(nt, [&](int t) {
            std::vector<int>&cand = scratch_a[t], &out = scratch_b[t];
            for (int i = t * chunk; i < std::min(N_panels, (t + 1) * chunk); ++i) {
                candidates_of(i, cand, out);
                if ((int)out.size() > NB_MAX) {
                    nb_n[i] = 255;
                } else {
                    nb_n[i] = (unsigned char)out.size();
                    for (size_t k = 0; k < out.size(); ++k) nb[(size_t)i * NB_MAX + k] = out[k];
                }
            }
        }, 1);
    }
    // ---- step 2 ----
    std::vector<int> cand, over;
    for (int i = 0; i < N_panels; ++i) {
        const int* list = nb.data() + (size_t)i * NB_MAX;
        int n_list = nb_n[i];
        if (n_list == 255) {
            candidates_of(i, cand, over);
            list = over.data();
            n_list = (int)over.size();
        }
        for (int k = 0; k < n_list; ++k) {
            const int j = list[k];
            if (all_found(i)) break;  // surface_mesh.f90:373
            int ep[2], ei, ej;
            if (check_adjacent(i, j, ep, ei, ej)) recs.push_back({i, j, ep[0], ep[1], ei, ej, false});
        }
    }
    if (mirrored) {
        for (int i = 0; i < N_panels; ++i) {
            if (all_found(i)) continue;
            int ep[2], ei;
            if (check_mirror(i, ep, ei)) recs.push_back({i, i + N_panels, ep[0], ep[1], ei, -1, true});
        }
    }
    for (int i = 0; i < N_panels; ++i) {
        for (int j = 0; j < 3; ++j) {
            if (abut[3 * (size_t)i + j] == -1) recs.push_back({i, -1, ivs[3 * (size_t)i + j], ivs[3 * (size_t)i + (j + 1) % 3], j, -1, false});
        }
    }
    // ---- step 3 ----
    N_edges = (int)recs.size();
    std::vector<int> pedge((size_t)N_panels * 3, -1);   // the edge of every panel side (a later edge overwrites, as in a serial loop)
    for (int i = 0; i < N_edges; ++i) {
        pedge[3 * (size_t)recs[i].panel1 + recs[i].edge_index1] = i;
        if (recs[i].panel2 < N_panels && recs[i].panel2 >= 0) pedge[3 * (size_t)recs[i].panel2 + recs[i].edge_index2] = i;
    }
    parallel_for(N_panels, [&](int i) {
        for (int m = 0; m < 3; ++m) {
            panels[i].abutting_panels[m] = abut[3 * (size_t)i + m];
            if (pedge[3 * (size_t)i + m] >= 0) panels[i].edges[m] = pedge[3 * (size_t)i + m];
        }
    }, 4096);
    edges.assign(N_edges, Edge());
    parallel_for(N_edges, [&](int i) {
        Edge& e = edges[i];
        e.top_verts[0] = recs[i].vertex1;
        e.top_verts[1] = recs[i].vertex2;
        e.bot_verts[0] = e.top_verts[0];
        e.bot_verts[1] = e.top_verts[1];
        e.panels[0] = recs[i].panel1;
        e.panels[1] = recs[i].panel2;
        e.on_mirror_plane = recs[i].on_mirror_plane;
        e.edge_index_for_panel[0] = recs[i].edge_index1;
        e.edge_index_for_panel[1] = recs[i].edge_index2;
    }, 4096);
    // vertex adjacency: the incidences (edge, other endpoint) of every vertex in edge order
    std::vector<int> start((size_t)N_verts + 1, 0);
    for (int i = 0; i < N_edges; ++i) {
        ++start[(size_t)recs[i].vertex1 + 1];
        ++start[(size_t)recs[i].vertex2 + 1];
    }
    for (int v = 0; v < N_verts; ++v) start[v + 1] += start[v];
    std::vector<int> fill(start.begin(), start.end() - 1), inc_edge((size_t)2 * N_edges), inc_other((size_t)2 * N_edges);
    for (int i = 0; i < N_edges; ++i) {
        const int a = recs[i].vertex1, b = recs[i].vertex2;
        inc_edge[fill[a]] = i;
        inc_other[fill[a]++] = b;
        inc_edge[fill[b]] = i;
        inc_other[fill[b]++] = a;
    }
    parallel_for(N_verts, [&](int v) {
        Vertex& vx = vertices[v];
        vx.adjacent_edges.assign(inc_edge.begin() + start[v], inc_edge.begin() + start[v + 1]);
        vx.adjacent_vertices.clear();
        for (int k = start[v]; k < start[v + 1]; ++k)
            if (!contains(vx.adjacent_vertices, inc_other[k])) vx.adjacent_vertices.push_back(inc_other[k]);
    }, 2048);
}

// surface_mesh.f90:610-659, base_geom.f90:163-214
void Case::calc_vertex_geometry() {
    parallel_for(N_verts, [&](int i) {   // binary128 sums per vertex (the reference's real(16)); every vertex writes only itself
        Vertex& v = vertices[i];
        quad n_avg[3] = {0, 0, 0};
        for (int j_panel : v.panels) {
            quad w[3];
            panel_weighted_normal_at_corner(panels[j_panel], vertices, v.loc, w);
            for (int k = 0; k < 3; ++k) n_avg[k] = n_avg[k] + w[k];
        }
        if (v.on_mirror_plane) n_avg[mirror_plane - 1] = 0.;
        quad nrm = norm2_gf<quad>(n_avg, 3);
        for (int k = 0; k < 3; ++k) v.n_g[k] = (double)(n_avg[k] / nrm);
        if (mirrored) v.n_g_mir = mirror_across_plane(v.n_g, mirror_plane);
        // set_average_edge_length
        v.l_avg = 0.;
        v.l_min = std::numeric_limits<double>::max();
        int N = 0;
        for (int adj : v.adjacent_vertices) {
            double l_i = dist(v.loc, vertices[adj].loc);
            v.l_min = std::min(v.l_min, l_i);
            if (v.on_mirror_plane && !vertices[adj].on_mirror_plane) {
                v.l_avg = v.l_avg + 2 * l_i;
                N += 2;
            } else {
                v.l_avg = v.l_avg + l_i;
                N += 1;
            }
        }
        if (N > 0) v.l_avg = v.l_avg / N;
        else v.l_avg = 1.;
    });
}

// surface_mesh.f90:759-801
void Case::init_panels_with_flow() {
    parallel_for((int)panels.size(), [&](int i) { panel_init_with_flow(panels[i], vertices, freestream, mirrored, mirror_plane); });
    N_subinc = N_supinc = 0;
    for (auto& p : panels) {
        if (p.r > 0) ++N_subinc; else ++N_supinc;
        if (asym_flow) {
            if (p.r_mir > 0) ++N_subinc; else ++N_supinc;
        }
    }
}

// surface_mesh.f90:804-932
void Case::characterize_edges() {
    found_wake_edges = false;
    double C_min_angle = 1.0;
    for (int k = 0; k < N_edges; ++k) {
        Edge& e = edges[k];
        int i = e.panels[0], j = e.panels[1];
        if (j == -1) {
            e.discontinuous = true;
            panels[i].N_discont_edges += 1;
            panels[i].edge_is_discontinuous[e.edge_index_for_panel[0]] = true;
            continue;
        }
        V3 second_normal = e.on_mirror_plane ? mirror_across_plane(panels[i].n_g, mirror_plane) : panels[j].n_g;
        double C_angle = inner(panels[i].n_g, second_normal);
        if (C_angle < C_max_cont_angle) {
            e.discontinuous = true;
            panels[i].N_discont_edges += 1;
            panels[i].edge_is_discontinuous[e.edge_index_for_panel[0]] = true;
            if (!e.on_mirror_plane) {
                panels[j].N_discont_edges += 1;
                panels[j].edge_is_discontinuous[e.edge_index_for_panel[1]] = true;
            }
        }
        C_min_angle = std::min(C_angle, C_min_angle);
        if (!wake_present) continue;
        if (C_angle < C_wake_shedding_angle) {
            if (inner(panels[i].n_g, freestream.v_inf) > 0.0 || inner(second_normal, freestream.v_inf) > 0.0) {
                int i_vert_1 = e.top_verts[0], i_vert_2 = e.top_verts[1];
                V3 t_hat_g = vertices[i_vert_2].loc - vertices[i_vert_1].loc;
                V3 cross_result = cross(panels[i].n_g, second_normal);
                if (inner(cross_result, t_hat_g) > 0.) {
                    found_wake_edges = true;
                    e.sheds_wake = true;
                    e.discontinuous = true;
                }
            }
        }
    }
    C_min_panel_angle = C_min_angle;
}

// base_geom.f90:237-290
void Case::set_needed_vertex_clones() {
    for (int i = 0; i < N_verts; ++i) {
        Vertex& v = vertices[i];
        v.N_wake_edges = 0;
        v.N_needed_clones = 0;
        int n_on_mirror = 0;
        for (int i_edge : v.adjacent_edges) {
            if (edges[i_edge].sheds_wake) {
                v.N_wake_edges += 1;
                if (edges[i_edge].on_mirror_plane) n_on_mirror += 1;
            }
        }
        if (v.N_wake_edges > 0) {
            if (v.on_mirror_plane) v.N_needed_clones = v.N_wake_edges - n_on_mirror;
            else v.N_needed_clones = v.N_wake_edges - 1;
        }
    }
}

// surface_mesh.f90:1126-1191
static void find_next_wake_edge(const Case& c, int i_start_edge, int i_shared_vert, int i_start_panel, int& i_end_edge,
                                std::vector<int>& i_panels_between) {
    i_panels_between.clear();
    int i_curr_panel = i_start_panel;
    int i_prev_panel = c.edges[i_start_edge].get_opposing_panel(i_start_panel);
    int i_next_panel = -1;
    for (;;) {
        i_panels_between.push_back(i_curr_panel);
        for (int i = 0; i < 3; ++i) {
            i_next_panel = c.panels[i_curr_panel].abutting_panels[i];
            i_end_edge = c.panels[i_curr_panel].edges[i];
            if (i_next_panel == i_prev_panel || i_end_edge == i_start_edge) continue;
            if (c.edges[i_end_edge].touches_vertex(i_shared_vert)) break;
        }
        if (c.edges[i_end_edge].sheds_wake) break;
        if (i_next_panel < 0 || i_next_panel >= c.N_panels) break;
        i_prev_panel = i_curr_panel;
        i_curr_panel = i_next_panel;
    }
}

// panel.f90:1705-1729
static void point_to_new_vertex(Case& c, Panel& p, int i_new) {
    for (int i = 0; i < 3; ++i) {
        if (dist(c.vertices[p.iv[i]].loc, c.vertices[i_new].loc) < 1e-12) {
            p.iv[i] = i_new;
            return;
        }
    }
}

// surface_mesh.f90:1194-1274, base_geom.f90:307-362
static void init_vertex_clone(Case& c, int i_jango, int i_boba, bool mirrored_is_unique,
                              const std::vector<int>& panels_for_this_clone) {
    Vertex& boba = c.vertices[i_boba];
    boba.init(c.vertices[i_jango].loc, i_boba);
    boba.clone = true;
    {
        const Vertex& jango = c.vertices[i_jango];
        boba.N_wake_edges = jango.N_wake_edges;
        boba.on_mirror_plane = jango.on_mirror_plane;
        boba.n_g = jango.n_g;
        boba.n_g_mir = jango.n_g_mir;
        boba.l_avg = jango.l_avg;
        for (int x : jango.panels) boba.panels.push_back(x);
        for (int x : jango.adjacent_vertices) boba.adjacent_vertices.push_back(x);
        for (int x : jango.adjacent_edges) boba.adjacent_edges.push_back(x);
    }
    boba.mirrored_is_unique = mirrored_is_unique;
    for (int i_panel : panels_for_this_clone) {
        if (i_panel != -1) {
            erase_all(c.vertices[i_jango].panels_not_across_wake_edge, i_panel);
            if (!contains(boba.panels_not_across_wake_edge, i_panel)) boba.panels_not_across_wake_edge.push_back(i_panel);
            if (i_panel < c.N_panels) point_to_new_vertex(c, c.panels[i_panel], i_boba);
        }
    }
    for (int i_edge : boba.adjacent_edges) {
        Edge& e = c.edges[i_edge];
        if (e.sheds_wake) continue;
        bool found_edge = false;
        for (int i_panel : panels_for_this_clone) {
            // the reference compares against the zero placeholders too; an edge with no second
            // panel stores 0 there, so a zero entry matches it (surface_mesh.f90:1258).
            if (e.panels[0] == i_panel || e.panels[1] == i_panel) {
                found_edge = true;
                break;
            }
        }
        // zero placeholders of the 20-slot column (surface_mesh.f90:1025,1066)
        if (!found_edge && panels_for_this_clone.size() < 20 && (e.panels[0] == -1 || e.panels[1] == -1))
            found_edge = true;
        if (found_edge) {
            e.point_top_to_new_vert(i_jango, i_boba);
            e.point_bottom_to_new_vert(i_jango, i_boba);
        }
    }
}

// surface_mesh.f90:956-1123
void Case::clone_vertices() {
    if (!found_wake_edges) return;
    int N_clones = 0;
    for (int i = 0; i < N_verts; ++i) N_clones += vertices[i].N_needed_clones;
    // allocate_new_vertices (mesh.f90:70-110): panel vertex references are indices here, so a
    // plain resize keeps them valid.
    const int N_orig = N_verts;
    vertices.resize(N_orig + N_clones);
    N_verts = N_orig + N_clones;

    std::vector<int> i_rearrange_inv(N_verts, -1);
    int j = 0;
    std::vector<int> i_panels_between;
    for (int i_jango = 0; i_jango < N_orig; ++i_jango) {
        i_rearrange_inv[i_jango] = i_jango + j;
        int N_boba = vertices[i_jango].N_needed_clones;
        if (N_boba > 0) {
            std::vector<int> i_start_edge(N_boba + 1, -1), i_end_edge(N_boba + 1, -1);
            for (int i_edge : vertices[i_jango].adjacent_edges) {
                if (edges[i_edge].on_mirror_plane) {
                    i_start_edge[0] = i_edge;
                    break;
                }
                if (edges[i_edge].sheds_wake) i_start_edge[0] = i_edge;
            }
            std::vector<char> mirrored_is_unique(N_boba + 1, 1);
            std::vector<std::vector<int>> i_panels_between_all(N_boba + 1);
            int i_start_panel = edges[i_start_edge[0]].panels[0];
            for (int i = 0; i <= N_boba; ++i) {
                find_next_wake_edge(*this, i_start_edge[i], i_jango, i_start_panel, i_end_edge[i], i_panels_between);
                i_panels_between_all[i] = i_panels_between;
                if (edges[i_start_edge[i]].on_mirror_plane && !edges[i_start_edge[i]].sheds_wake) mirrored_is_unique[i] = 0;
                if (edges[i_end_edge[i]].on_mirror_plane && !edges[i_end_edge[i]].sheds_wake) mirrored_is_unique[i] = 0;
                if (i < N_boba) i_start_edge[i + 1] = i_end_edge[i];
                i_start_panel = edges[i_end_edge[i]].get_opposing_panel(i_panels_between.back());
            }
            vertices[i_jango].mirrored_is_unique = mirrored_is_unique[0] != 0;
            for (int i = 1; i <= N_boba; ++i) {
                j = j + 1;
                int i_boba = N_orig + j - 1;  // position N_verts-N_clones+j (1-based) in the new array
                i_rearrange_inv[i_boba] = i_jango + j;
                init_vertex_clone(*this, i_jango, i_boba, mirrored_is_unique[i] != 0, i_panels_between_all[i]);
                edges[i_start_edge[i]].point_bottom_to_new_vert(i_jango, i_boba);
                if (i < N_boba) edges[i_end_edge[i]].point_top_to_new_vert(i_jango, i_boba);
                else edges[i_end_edge[i]].point_bottom_to_new_vert(i_jango, i_boba);
            }
            vertices[i_jango].clone = true;
        } else {
            if (mirrored && asym_flow && vertices[i_jango].on_mirror_plane && vertices[i_jango].mirrored_is_unique) {
                for (int i_edge : vertices[i_jango].adjacent_edges)
                    if (edges[i_edge].sheds_wake) edges[i_edge].point_bottom_to_new_vert(i_jango, i_jango + N_verts);
            }
        }
    }
    // invert_permutation_vector (helpers.f90:56-72)
    vertex_ordering.assign(N_verts, -1);
    for (int i = 0; i < N_verts; ++i) vertex_ordering[i_rearrange_inv[i]] = i;
}

// surface_mesh.f90:1392-1470
bool Case::is_convex_at_vertex(int i_vert) const {
    bool first = true;
    double s = 1.;
    const Vertex& v = vertices[i_vert];
    for (int i_neighbor : v.panels) {
        for (int j_neighbor : v.adjacent_vertices) {
            double h = inner(panels[i_neighbor].n_g, vertices[j_neighbor].loc - panels[i_neighbor].centr);
            if (std::fabs(h) > 1.e-10) {
                if (first) {
                    s = sign(s, h);
                    first = false;
                } else if (s * h < 0.) {
                    return false;
                }
            }
            if (v.on_mirror_plane) {
                h = inner(panels[i_neighbor].n_g_mir, vertices[j_neighbor].loc - panels[i_neighbor].centr_mir);
                if (std::fabs(h) > 1.e-10) {
                    if (first) {
                        s = sign(s, h);
                        first = false;
                    } else if (s * h < 0.) {
                        return false;
                    }
                }
            }
        }
    }
    return true;
}

// surface_mesh.f90:683-756
void Case::init_with_flow() {
    const Json* flow_j = input.find("flow");
    if (!flow_j) throw std::runtime_error("input has no flow section");
    const Json* g = input.find("geometry");
    spanwise_axis = g->get("spanwise_axis", "+y");
    freestream.init(*flow_j, spanwise_axis);

    asym_flow = false;
    if (mirrored && !freestream.sym_about[mirror_plane - 1]) asym_flow = true;
    const bool timing = std::getenv("MLH_TIMING") != nullptr;
    auto t0 = std::chrono::steady_clock::now();
    auto lap = [&](const char* what) {
        if (!timing) return;
        auto t1 = std::chrono::steady_clock::now();
        std::fprintf(stderr, "mlh init_with_flow: %-24s %8.2f ms\n", what, std::chrono::duration<double, std::milli>(t1 - t0).count());
        t0 = t1;
    };
    init_panels_with_flow();
    lap("init_panels_with_flow");
    characterize_edges();
    lap("characterize_edges");
    if (wake_present) {
        set_needed_vertex_clones();
        clone_vertices();
    }
    lap("clone_vertices");
    if (!found_wake_edges) {
        vertex_ordering.resize(N_verts);
        for (int i = 0; i < N_verts; ++i) vertex_ordering[i] = i;
    }
    parallel_for(N_verts, [&](int i) { vertices[i].convex = is_convex_at_vertex(i); });   // const test, own flag
    lap("is_convex_at_vertex");
    init_wake();
    lap("init_wake");
    // reads the neighbours' vertex indices (fixed by now), writes the panel itself
    parallel_for((int)panels.size(), [&](int i) {
        panel_set_distribution(panels[i], initial_panel_order, panels, vertices, vertices, mirrored, mirror_plane, force_sigma_match);
    });
    lap("panel_set_distribution");
}

}  // namespace mlh
