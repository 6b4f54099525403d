// Wake model setup: host restatement of src/surface_mesh.f90:1277-1389 (init_wake, Trefftz
// distance), src/wake_mesh.f90:38-146 and src/wake_strip.f90:28-283 (one straight strip of
// triangles per wake-shedding edge).  Produces the wake panel table for the GPU assembly.
#include <algorithm>
#include <cmath>

#include "model.hpp"

namespace mlh {

// mesh.f90:30-43
static bool has_zero_area(const std::vector<Vertex>& v, int i1, int i2, int i3) {
    return norm2(cross(v[i3].loc - v[i2].loc, v[i2].loc - v[i1].loc)) < 1.e-12;
}

// wake_strip.f90:28-283
static void wake_strip_init(WakeStrip& st, const Flow& fs, const Edge& starting_edge, bool mirror_start,
                            int mirror_plane, int N_panels_streamwise, double trefftz_dist,
                            const std::vector<Vertex>& body_verts, bool wake_mirrored, int N_body_panels) {
    const int N_body_verts = (int)body_verts.size();
    st.on_mirror_plane = starting_edge.on_mirror_plane;
    st.mirror_plane = mirror_plane;
    st.mirrored = wake_mirrored && !st.on_mirror_plane;
    V3 start_1, start_2;
    if (mirror_start) {
        start_1 = mirror_across_plane(body_verts[starting_edge.top_verts[1]].loc, mirror_plane);
        start_2 = mirror_across_plane(body_verts[starting_edge.top_verts[0]].loc, mirror_plane);
        st.i_top_parent_1 = starting_edge.top_verts[1] + N_body_verts;
        st.i_top_parent_2 = starting_edge.top_verts[0] + N_body_verts;
        st.i_bot_parent_1 = starting_edge.bot_verts[1] + N_body_verts;
        st.i_bot_parent_2 = starting_edge.bot_verts[0] + N_body_verts;
        st.i_top_parent = starting_edge.panels[0] + N_body_panels;
        st.i_bot_parent = starting_edge.panels[1] + N_body_panels;
    } else {
        start_1 = body_verts[starting_edge.top_verts[0]].loc;
        start_2 = body_verts[starting_edge.top_verts[1]].loc;
        st.i_top_parent_1 = starting_edge.top_verts[0];
        st.i_top_parent_2 = starting_edge.top_verts[1];
        st.i_bot_parent_1 = starting_edge.bot_verts[0];
        st.i_bot_parent_2 = starting_edge.bot_verts[1];
        st.i_top_parent = starting_edge.panels[0];
        st.i_bot_parent = starting_edge.panels[1];
    }

    // init_vertices, wake_strip.f90:105-169 (vertex k here is the reference's vertex k+1)
    st.N_verts = N_panels_streamwise * 2 + 2;
    st.vertices.assign(st.N_verts, Vertex());
    st.vertices[0].init(start_1, 0);
    st.vertices[1].init(start_2, 1);
    st.vertices[0].top_parent = st.i_top_parent_1;
    st.vertices[0].bot_parent = st.i_bot_parent_1;
    st.vertices[1].top_parent = st.i_top_parent_2;
    st.vertices[1].bot_parent = st.i_bot_parent_2;
    double d1 = trefftz_dist - inner(start_1, fs.c_hat_g);
    double d2 = trefftz_dist - inner(start_2, fs.c_hat_g);
    double sep_1 = d1 / N_panels_streamwise;
    double sep_2 = d2 / N_panels_streamwise;
    for (int i = 3; i <= st.N_verts; ++i) {  // 1-based i as in the reference
        V3 loc;
        if (i % 2 == 0) loc = start_2 + (sep_2 * (i - 2) / 2) * fs.c_hat_g;  // sep*(i-2)/2*c_hat, left to right
        else loc = start_1 + (sep_1 * (i - 1) / 2) * fs.c_hat_g;
        st.vertices[i - 1].init(loc, i - 1);
        if (i % 2 == 0) {
            st.vertices[i - 1].top_parent = st.i_top_parent_2;
            st.vertices[i - 1].bot_parent = st.i_bot_parent_2;
        } else {
            st.vertices[i - 1].top_parent = st.i_top_parent_1;
            st.vertices[i - 1].bot_parent = st.i_bot_parent_1;
        }
    }

    // init_panels, wake_strip.f90:172-263 (i1, i2 1-based as in the reference)
    int N_panels = N_panels_streamwise * 2;
    std::vector<Panel> tmp(N_panels);
    std::vector<char> skipped(N_panels, 0);
    int i1 = 1, i2 = 2;
    auto init_panel = [&](int i_panel, int a, int b, int c) {  // wake_strip.f90:266-283
        if (has_zero_area(st.vertices, a - 1, b - 1, c - 1)) skipped[i_panel] = 1;
        else panel_init(tmp[i_panel], st.vertices, a - 1, b - 1, c - 1, i_panel, true);
    };
    for (int i = 0; i < N_panels; ++i) {
        int advance;
        if (i1 == st.N_verts - 1) advance = 2;
        else if (i2 == st.N_verts) advance = 1;
        else {
            double h1 = dist(st.vertices[i1 + 2 - 1].loc, st.vertices[i2 - 1].loc);
            double h2 = dist(st.vertices[i1 - 1].loc, st.vertices[i2 + 2 - 1].loc);
            advance = (h1 < h2) ? 1 : 2;
        }
        if (advance == 1) {
            init_panel(i, i1, i1 + 2, i2);
            i1 += 2;
        } else {
            init_panel(i, i1, i2 + 2, i2);
            i2 += 2;
        }
    }
    st.panels.clear();
    for (int i = 0; i < N_panels; ++i)
        if (!skipped[i]) st.panels.push_back(tmp[i]);
    st.N_panels = (int)st.panels.size();

    for (auto& p : st.panels) {
        panel_init_with_flow(p, st.vertices, fs, st.mirrored, mirror_plane);
        panel_set_distribution(p, 1, st.panels, st.vertices, st.vertices, st.mirrored, mirror_plane, true);
    }
}

// surface_mesh.f90:1277-1389, wake_mesh.f90:38-146
void Case::init_wake() {
    wake = WakeMesh();
    if (!(append_wake && found_wake_edges)) return;
    if (trefftz_distance < 0.) {
        if (freestream.supersonic) {  // update_supersonic_trefftz_distance
            double max_dist = 0.;
            for (int i = 0; i < N_verts; ++i) {
                double distance = inner(vertices[i].loc, freestream.c_hat_g);
                max_dist = std::max(distance, max_dist);
                if (asym_flow) {
                    distance = inner(mirror_across_plane(vertices[i].loc, mirror_plane), freestream.c_hat_g);
                    max_dist = std::max(distance, max_dist);
                }
            }
            trefftz_distance = max_dist;
        } else {  // update_subsonic_trefftz_distance
            double front = inner(freestream.c_hat_g, vertices[0].loc);
            double back = front;
            for (int i = 1; i < N_verts; ++i) {
                double x = inner(freestream.c_hat_g, vertices[i].loc);
                front = std::min(front, x);
                back = std::max(back, x);
            }
            trefftz_distance = 20. * std::fabs(front - back);
        }
    }
    wake.mirrored = mirrored && !asym_flow;
    wake.mirror_plane = mirror_plane;
    std::vector<int> wake_shedding_edges;
    wake.N_strips = 0;
    for (int i = 0; i < N_edges; ++i) {
        if (edges[i].sheds_wake) {
            wake.N_strips += 1;
            wake_shedding_edges.push_back(i);
            if (asym_flow && !edges[i].on_mirror_plane) wake.N_strips += 1;
        }
    }
    wake.strips.assign(wake.N_strips, WakeStrip());
    int i = -1, i_strip = 0;
    while (i_strip < wake.N_strips) {
        i += 1;
        int i_start_edge = wake_shedding_edges[i];
        wake_strip_init(wake.strips[i_strip], freestream, edges[i_start_edge], false, mirror_plane,
                        N_wake_panels_streamwise, trefftz_distance, vertices, wake.mirrored, N_panels);
        i_strip += 1;
        if (asym_flow && !edges[i_start_edge].on_mirror_plane) {
            wake_strip_init(wake.strips[i_strip], freestream, edges[i_start_edge], true, mirror_plane,
                            N_wake_panels_streamwise, trefftz_distance, vertices, wake.mirrored, N_panels);
            i_strip += 1;
        }
    }
    wake.N_verts = wake.N_panels = 0;
    for (auto& st : wake.strips) {
        wake.N_max_strip_panels = std::max(wake.N_max_strip_panels, st.N_panels);
        wake.N_max_strip_verts = std::max(wake.N_max_strip_verts, st.N_verts);
        wake.N_verts += st.N_verts;
        wake.N_panels += st.N_panels;
    }
}

}  // namespace mlh
