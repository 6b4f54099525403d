// Mesh readers: ASCII VTK v3/v5 POLYDATA (src/vtk.f90:434-664), ASCII STL (src/stl.f90:11-117),
// Cart3D-style .tri (src/tri.f90:13-102), with the reference's duplicate-vertex collapse
// (src/stl.f90:120-236).  The all-pairs duplicate search is replaced by a spatial hash that
// returns the same answer (first unique vertex within 1e-12 in file order) in O(N).
#include <charconv>
#include <chrono>
#include <cstdlib>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <fstream>
#include <sstream>
#include <stdexcept>
#include <cstring>
#include <unordered_map>

#include "model.hpp"
#include "parallel.hpp"

namespace mlh {

std::string read_text_file(const std::string& path) {
    std::FILE* f = std::fopen(path.c_str(), "rb");
    if (!f) throw std::runtime_error("cannot open file " + path);
    std::string text;
    if (std::fseek(f, 0, SEEK_END) == 0) {
        const long sz = std::ftell(f);
        std::rewind(f);
        if (sz > 0) {
            text.resize((size_t)sz);
            const size_t got = std::fread(&text[0], 1, (size_t)sz, f);
            text.resize(got);
        }
    }
    std::fclose(f);
    return text;
}

namespace {

struct CellKey {
    int64_t x, y, z;
    bool operator==(const CellKey& o) const { return x == o.x && y == o.y && z == o.z; }
};
struct CellHash {
    size_t operator()(const CellKey& k) const {
        uint64_t h = (uint64_t)k.x * 0x9E3779B97F4A7C15ull;
        h ^= (uint64_t)k.y * 0xC2B2AE3D27D4EB4Full + (h << 6) + (h >> 2);
        h ^= (uint64_t)k.z * 0x165667B19E3779F9ull + (h << 6) + (h >> 2);
        return (size_t)h;
    }
};

// stl.f90:120-236.  vertex_locs in file order -> unique vertices + new_ind (0-based).
// The reference compares every vertex with every earlier unique one (O(N^2)): a vertex is a duplicate of the FIRST earlier
// unique vertex within 1e-12.  Here the unique vertices sit in a spatial hash of cell size 1e-6 (flat open addressing, one chain
// per cell); a vertex within 1e-12 of another lies in the same cell or, only if it is within 1e-12 of a cell face, in the
// neighbour across that face -- so one cell is searched as a rule (27 in the first version of this routine).
void collapse_duplicate_vertices(const std::vector<V3>& vertex_locs, std::vector<Vertex>& vertices,
                                 std::vector<int>& new_ind) {
    const int N = (int)vertex_locs.size();
    std::vector<char> is_duplicate(N, 0);
    std::vector<int> duplicate_of(N);
    const double h = 1e-6, tol = 1.e-12;
    size_t cap = 16;
    while (cap < (size_t)N * 2 + 16) cap <<= 1;
    std::vector<int> head(cap, -1);           // slot -> first unique vertex of the cell stored there
    std::vector<CellKey> slot_key(cap);
    std::vector<int> next(N, -1);             // chain of the unique vertices of one cell
    CellHash hasher;
    auto find_slot = [&](const CellKey& k, bool insert) -> long {
        size_t s = hasher(k) & (cap - 1);
        for (;;) {
            if (head[s] < 0) {
                if (!insert) return -1;
                slot_key[s] = k;
                return (long)s;
            }
            if (slot_key[s] == k) return (long)s;
            s = (s + 1) & (cap - 1);
        }
    };
    for (int j = 0; j < N; ++j) {
        duplicate_of[j] = j;
        const V3& p = vertex_locs[j];
        const double q[3] = {p[0] / h, p[1] / h, p[2] / h};
        const double fl[3] = {std::floor(q[0]), std::floor(q[1]), std::floor(q[2])};
        const CellKey c{(int64_t)fl[0], (int64_t)fl[1], (int64_t)fl[2]};
        // which neighbours can hold a point within tol: only across a face the vertex (almost) touches; the margin is a
        // generous multiple of tol / h so that rounding of p / h cannot hide a neighbour
        int64_t lo[3], hi[3];
        for (int d = 0; d < 3; ++d) {
            const double fr = q[d] - fl[d], margin = 1e-3;   // tol / h = 1e-6
            lo[d] = fr < margin ? -1 : 0;
            hi[d] = fr > 1. - margin ? 1 : 0;
        }
        int best = -1;
        for (int64_t dx = lo[0]; dx <= hi[0]; ++dx)
            for (int64_t dy = lo[1]; dy <= hi[1]; ++dy)
                for (int64_t dz = lo[2]; dz <= hi[2]; ++dz) {
                    const long s = find_slot(CellKey{c.x + dx, c.y + dy, c.z + dz}, false);
                    if (s < 0) continue;
                    for (int i = head[s]; i >= 0; i = next[i])
                        if (dist(vertex_locs[i], p) < tol && (best < 0 || i < best)) best = i;
                }
        if (best >= 0) {
            is_duplicate[j] = 1;
            duplicate_of[j] = best;
        } else {
            // only unique vertices are candidates (stl.f90:146,152); appended at the head of its cell's chain
            const long s = find_slot(c, true);
            next[j] = head[s];
            head[s] = j;
        }
    }
    new_ind.assign(N, 0);
    int N_duplicates = 0;
    for (int i = 0; i < N; ++i) {
        if (is_duplicate[i]) {
            new_ind[i] = new_ind[duplicate_of[i]];
            ++N_duplicates;
        } else {
            new_ind[i] = i - N_duplicates;
        }
    }
    vertices.assign(N - N_duplicates, Vertex());
    parallel_for(N, [&](int i) {   // a unique vertex initialises its own slot
        if (!is_duplicate[i]) vertices[new_ind[i]].init(vertex_locs[i], new_ind[i]);
    }, 4096);
}

std::vector<std::string> split_ws(const std::string& line) {
    std::vector<std::string> out;
    std::istringstream ss(line);
    std::string w;
    while (ss >> w) out.push_back(w);
    return out;
}

// N default-constructed panels.  A Panel is 1.4 KB: at 280k panels the storage is 400 MB of fresh pages, and their first touch
// (one page fault per 4 KB, in the kernel) costs more than the construction itself -- so the pages are touched on the host threads
// first (writes of zero into the reserved, not yet constructed storage), then the panels are constructed in place, serially, on
// memory that is already mapped.  The vertices' panel lists get room for the usual valence at the same time.
static void fresh_panels(std::vector<Panel>& panels, int N_panels, std::vector<Vertex>& vertices) {
    panels.clear();
    panels.reserve((size_t)N_panels);
    char* const raw = reinterpret_cast<char*>(panels.data());
    const size_t bytes = (size_t)N_panels * sizeof(Panel), page = 4096;
    const int n_pages = (int)((bytes + page - 1) / page);
    if (raw) parallel_for(n_pages, [&](int k) { raw[(size_t)k * page] = 0; }, 2048);
    panels.resize((size_t)N_panels);
    (void)vertices;
}

// What panel_init does for the triangles tri[3 i .. 3 i + 2] (unique-vertex indices), i = 0 .. N - 1, on freshly constructed panels:
// every panel gets its vertex indices and its geometry (host threads: a panel writes itself), and every vertex the list of the
// panels that use it -- in panel order, as the push_backs of a serial loop over the panels leave it -- built by a counting
// sort over the incidences (streaming passes instead of 3 N appends to N_verts little heap arrays).
static void register_triangles(std::vector<Panel>& panels, std::vector<Vertex>& vertices, const std::vector<int>& tri) {
    const int N = (int)panels.size(), V = (int)vertices.size();
    parallel_for(N, [&](int i) {
        Panel& p = panels[i];
        p.N = 3;
        p.iv[0] = tri[3 * (size_t)i];
        p.iv[1] = tri[3 * (size_t)i + 1];
        p.iv[2] = tri[3 * (size_t)i + 2];
        p.index = i;
        p.in_wake = false;
        p.has_sources = true;
        panel_calc_derived_geom(p, vertices);
    });
    std::vector<int> start((size_t)V + 1, 0);
    for (size_t k = 0; k < tri.size(); ++k) ++start[(size_t)tri[k] + 1];
    for (int v = 0; v < V; ++v) start[v + 1] += start[v];
    std::vector<int> fill(start.begin(), start.end() - 1), inc(tri.size());
    for (int i = 0; i < N; ++i)
        for (int m = 0; m < 3; ++m) inc[fill[tri[3 * (size_t)i + m]]++] = i;
    parallel_for(V, [&](int v) {
        vertices[v].panels.assign(inc.begin() + start[v], inc.begin() + start[v + 1]);
        vertices[v].panels_not_across_wake_edge = vertices[v].panels;
    }, 2048);
}

// One decimal number at p (leading blanks skipped): std::from_chars is correctly rounded like strtod -- the same double for the
// same text -- at a fifth of the time; anything it does not take (a leading '+', "inf", ...) goes through strtod.
static inline double scan_double(const char*& p, const char* end) {
    while (p < end && (*p == ' ' || *p == '\t')) ++p;
    double x = 0.;
    const auto r = std::from_chars(p, end, x);
    if (r.ec == std::errc() && r.ptr != p) {
        p = r.ptr;
        return x;
    }
    char* q = nullptr;
    x = std::strtod(p, &q);
    p = q;
    return x;
}

void load_vtk(const std::string& text, std::vector<Vertex>& vertices, std::vector<Panel>& panels) {
    // Legacy ASCII VTK (vtk.f90:440-640).  The numbers are scanned in place with strtod / strtol (same conversions as a
    // stream would apply, no per-token strings); the few header lines go through a line reader.
    const char* p = text.c_str();
    const char* const end = p + text.size();
    auto next_line = [&]() -> std::string {
        const char* b = p;
        while (p < end && *p != '\n') ++p;
        std::string l(b, p);
        if (p < end) ++p;
        return l;
    };
    const bool timing = std::getenv("MLH_TIMING") != nullptr;
    auto t0 = std::chrono::steady_clock::now();
    auto lap = [&](const char* what) {
        if (!timing) return;
        auto t1 = std::chrono::steady_clock::now();
        std::fprintf(stderr, "mlh load_vtk: %-24s %8.2f ms\n", what, std::chrono::duration<double, std::milli>(t1 - t0).count());
        t0 = t1;
    };
    std::string line = next_line();
    size_t ind = line.find("Version");
    if (ind == std::string::npos) throw std::runtime_error("VTK header has no Version");
    int ver = line[ind + 8] - '0';
    if (ver != 3 && ver != 5) throw std::runtime_error("VTK file version not recognized");
    for (int k = 0; k < 3; ++k) next_line();  // 3 more header lines
    line = next_line();                       // POINTS n float
    auto w = split_ws(line);
    if (w.size() < 2) throw std::runtime_error("VTK: POINTS line not found");
    const int N_verts = std::atoi(w[1].c_str());
    std::vector<V3> locs(N_verts);
    for (int i = 0; i < N_verts; ++i)
        for (int k = 0; k < 3; ++k) {
            while (p < end && (*p == ' ' || *p == '\t' || *p == '\r' || *p == '\n')) ++p;
            const char* const before = p;
            locs[i][k] = scan_double(p, end);
            if (p == before) throw std::runtime_error("VTK: truncated POINTS section");
        }
    lap("points");
    std::vector<int> new_ind;
    collapse_duplicate_vertices(locs, vertices, new_ind);
    lap("collapse_duplicates");
    next_line();  // rest of last coordinate line
    auto next_int = [&]() -> long {
        char* q = nullptr;
        const long v = std::strtol(p, &q, 10);
        if (q == p) throw std::runtime_error("VTK: truncated connectivity");
        p = q;
        return v;
    };
    auto index = [&](long i) -> int {
        if (i < 0 || i >= N_verts) throw std::runtime_error("VTK: vertex index out of range");
        return new_ind[i];
    };
    std::vector<int> tri;
    if (ver == 3) {
        do {
            if (p >= end) throw std::runtime_error("VTK: POLYGONS not found");
            line = next_line();
        } while (line.find("POLYGONS") == std::string::npos);
        w = split_ws(line);
        int N_panels = std::atoi(w.at(1).c_str());
        fresh_panels(panels, N_panels, vertices);
        lap("fresh_panels");
        tri.resize((size_t)3 * N_panels);
        for (int i = 0; i < N_panels; ++i) {
            if (next_int() != 3) throw std::runtime_error("MachLine supports only triangular panels.");
            for (int m = 0; m < 3; ++m) tri[3 * (size_t)i + m] = index(next_int());
        }
    } else {
        do {
            if (p >= end) throw std::runtime_error("VTK: POLYGONS/CELLS not found");
            line = next_line();
        } while (line.find("POLYGONS") == std::string::npos && line.find("CELLS") == std::string::npos);
        w = split_ws(line);
        int N_panels = std::atoi(w.at(1).c_str()) - 1;
        fresh_panels(panels, N_panels, vertices);
        lap("panels.assign");
        do {
            if (p >= end) throw std::runtime_error("VTK: CONNECTIVITY not found");
            line = next_line();
        } while (line.find("CONNECTIVITY") == std::string::npos);
        tri.resize((size_t)3 * N_panels);
        for (int idx = 0; idx < N_panels; ++idx)
            for (int m = 0; m < 3; ++m) tri[3 * (size_t)idx + m] = index(next_int());
    }
    lap("connectivity");
    register_triangles(panels, vertices, tri);
    lap("register_triangles");
}

void load_stl(const std::string& text, std::vector<Vertex>& vertices, std::vector<Panel>& panels) {
    // ASCII STL (stl.f90:14-117): every line whose first word is "vertex" carries one corner.  Scanned in place, in chunks of
    // whole lines on the host threads (parallel.hpp); the chunks' corners are concatenated in file order.
    const char* const begin = text.c_str();
    const char* const end = begin + text.size();
    const char* p0 = begin;
    while (p0 < end && *p0 != '\n') ++p0;   // header line
    const int nt = std::max(1, std::min(host_threads(), (int)((end - p0) / (256 * 1024)) + 1));
    std::vector<const char*> cut(nt + 1);
    cut[0] = p0;
    cut[nt] = end;
    for (int t = 1; t < nt; ++t) {
        const char* c = p0 + (size_t)(end - p0) * t / nt;
        if (c < cut[t - 1]) c = cut[t - 1];
        while (c < end && *c != '\n') ++c;   // a chunk starts at the end of a line
        cut[t] = c;
    }
    std::vector<std::vector<V3>> part(nt);
    parallel_for(nt, [&](int t) {
        const char* p = cut[t];
        const char* const stop = cut[t + 1];
        std::vector<V3>& locs = part[t];
        locs.reserve((size_t)(stop - p) / 60 + 16);
        while (p < stop) {
            while (p < end && (*p == ' ' || *p == '\t' || *p == '\r' || *p == '\n')) ++p;
            if (p >= stop) break;   // the line starting here belongs to the next chunk
            if (end - p > 6 && std::memcmp(p, "vertex", 6) == 0 && (p[6] == ' ' || p[6] == '\t')) {
                p += 6;
                V3 v;
                v[0] = scan_double(p, end);
                v[1] = scan_double(p, end);
                v[2] = scan_double(p, end);
                locs.push_back(v);
            }
            while (p < end && *p != '\n') ++p;
        }
    }, 1);
    const bool timing = std::getenv("MLH_TIMING") != nullptr;
    auto t0 = std::chrono::steady_clock::now();
    auto lap = [&](const char* what) {
        if (!timing) return;
        auto t1 = std::chrono::steady_clock::now();
        std::fprintf(stderr, "mlh load_stl: %-24s %8.2f ms\n", what, std::chrono::duration<double, std::milli>(t1 - t0).count());
        t0 = t1;
    };
    std::vector<V3> locs;
    size_t total = 0;
    for (const auto& v : part) total += v.size();
    locs.reserve(total);
    for (const auto& v : part) locs.insert(locs.end(), v.begin(), v.end());
    int N_panels = (int)locs.size() / 3;
    std::vector<int> new_ind;
    lap("concatenate");
    collapse_duplicate_vertices(locs, vertices, new_ind);
    lap("collapse_duplicates");
    fresh_panels(panels, N_panels, vertices);
    lap("panels.assign");
    new_ind.resize((size_t)3 * N_panels);   // corner k of facet i is file vertex 3 i + k
    register_triangles(panels, vertices, new_ind);
    lap("register_triangles");
}

void load_tri(const std::string& text, std::vector<Vertex>& vertices, std::vector<Panel>& panels) {
    std::istringstream in(text);
    int N_verts = 0, N_panels = 0;
    in >> N_verts >> N_panels;
    std::string line;
    std::getline(in, line);
    std::vector<V3> locs(N_verts);
    for (int i = 0; i < N_verts; ++i) {
        std::getline(in, line);
        auto w = split_ws(line);
        if (w.size() < 3) throw std::runtime_error("tri: bad vertex line");
        locs[i] = {std::strtod(w[0].c_str(), nullptr), std::strtod(w[1].c_str(), nullptr),
                   std::strtod(w[2].c_str(), nullptr)};
    }
    std::vector<int> new_ind;
    collapse_duplicate_vertices(locs, vertices, new_ind);
    fresh_panels(panels, N_panels, vertices);
    for (int i = 0; i < N_panels; ++i) {
        int i1, i2, i3;
        in >> i1 >> i2 >> i3;
        panel_init(panels[i], vertices, new_ind[i1 - 1], new_ind[i2 - 1], new_ind[i3 - 1], i, false);
    }
}

}  // namespace

// surface_mesh.f90:216-264
void Case::load_mesh_file(const std::string& file) {
    std::string path = file;
    if (!base_dir.empty() && !file.empty() && file[0] != '/') path = base_dir + "/" + file;
    size_t loc = file.find('.');
    std::string ext = loc == std::string::npos ? "" : file.substr(loc);
    size_t last = file.rfind('.');
    if (last != std::string::npos) ext = file.substr(last);  // tolerate dots in directories
    std::string text = read_text_file(path);
    if (ext == ".vtk") load_vtk(text, vertices, panels);
    else if (ext == ".stl") load_stl(text, vertices, panels);
    else if (ext == ".tri") load_tri(text, vertices, panels);
    else throw std::runtime_error("MachLine cannot read " + ext + " type mesh files.");
    N_verts = (int)vertices.size();
    N_panels = (int)panels.size();
}

}  // namespace mlh
