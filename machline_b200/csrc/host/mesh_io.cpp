// Mesh readers: ASCII VTK v3/v5 POLYDATA (src/vtk.f90:434-664), ASCII STL (src/stl.f90:11-117),
// Cart3D-style .tri (src/tri.f90:13-102), with the reference's duplicate-vertex collapse
// (src/stl.f90:120-236).  The all-pairs duplicate search is replaced by a spatial hash that
// returns the same answer (first unique vertex within 1e-12 in file order) in O(N).
#include <cmath>
#include <cstdint>
#include <fstream>
#include <sstream>
#include <stdexcept>
#include <cstring>
#include <unordered_map>

#include "model.hpp"

namespace mlh {

std::string read_text_file(const std::string& path) {
    std::ifstream f(path, std::ios::binary);
    if (!f) throw std::runtime_error("cannot open file " + path);
    std::stringstream ss;
    ss << f.rdbuf();
    return ss.str();
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
void collapse_duplicate_vertices(const std::vector<V3>& vertex_locs, std::vector<Vertex>& vertices,
                                 std::vector<int>& new_ind) {
    const int N = (int)vertex_locs.size();
    std::vector<char> is_duplicate(N, 0);
    std::vector<int> duplicate_of(N);
    const double h = 1e-6;  // cell size >> tolerance; neighbours cover straddling
    std::unordered_map<CellKey, std::vector<int>, CellHash> grid;
    grid.reserve((size_t)N * 2);
    for (int j = 0; j < N; ++j) {
        duplicate_of[j] = j;
        const V3& p = vertex_locs[j];
        CellKey c{(int64_t)std::floor(p[0] / h), (int64_t)std::floor(p[1] / h), (int64_t)std::floor(p[2] / h)};
        int best = -1;
        for (int64_t dx = -1; dx <= 1; ++dx)
            for (int64_t dy = -1; dy <= 1; ++dy)
                for (int64_t dz = -1; dz <= 1; ++dz) {
                    auto it = grid.find(CellKey{c.x + dx, c.y + dy, c.z + dz});
                    if (it == grid.end()) continue;
                    for (int i : it->second)
                        if (dist(vertex_locs[i], p) < 1.e-12 && (best < 0 || i < best)) best = i;
                }
        if (best >= 0) {
            is_duplicate[j] = 1;
            duplicate_of[j] = best;
        } else {
            grid[c].push_back(j);  // only unique vertices are candidates (stl.f90:146,152)
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
    for (int i = 0; i < N; ++i)
        if (!is_duplicate[i]) vertices[new_ind[i]].init(vertex_locs[i], new_ind[i]);
}

std::vector<std::string> split_ws(const std::string& line) {
    std::vector<std::string> out;
    std::istringstream ss(line);
    std::string w;
    while (ss >> w) out.push_back(w);
    return out;
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
            char* q = nullptr;
            locs[i][k] = std::strtod(p, &q);
            if (q == p) throw std::runtime_error("VTK: truncated POINTS section");
            p = q;
        }
    std::vector<int> new_ind;
    collapse_duplicate_vertices(locs, vertices, new_ind);
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
    if (ver == 3) {
        do {
            if (p >= end) throw std::runtime_error("VTK: POLYGONS not found");
            line = next_line();
        } while (line.find("POLYGONS") == std::string::npos);
        w = split_ws(line);
        int N_panels = std::atoi(w.at(1).c_str());
        panels.assign(N_panels, Panel());
        for (int i = 0; i < N_panels; ++i) {
            if (next_int() != 3) throw std::runtime_error("MachLine supports only triangular panels.");
            const int i1 = index(next_int()), i2 = index(next_int()), i3 = index(next_int());
            panel_init(panels[i], vertices, i1, i2, i3, i, false);
        }
    } else {
        do {
            if (p >= end) throw std::runtime_error("VTK: POLYGONS/CELLS not found");
            line = next_line();
        } while (line.find("POLYGONS") == std::string::npos && line.find("CELLS") == std::string::npos);
        w = split_ws(line);
        int N_panels = std::atoi(w.at(1).c_str()) - 1;
        panels.assign(N_panels, Panel());
        do {
            if (p >= end) throw std::runtime_error("VTK: CONNECTIVITY not found");
            line = next_line();
        } while (line.find("CONNECTIVITY") == std::string::npos);
        for (int idx = 0; idx < N_panels; ++idx) {
            const int i1 = index(next_int()), i2 = index(next_int()), i3 = index(next_int());
            panel_init(panels[idx], vertices, i1, i2, i3, idx, false);
        }
    }
}

void load_stl(const std::string& text, std::vector<Vertex>& vertices, std::vector<Panel>& panels) {
    // ASCII STL (stl.f90:14-117): every line whose first word is "vertex" carries one corner.  Scanned in place (the
    // coordinates go through the same strtod as before: identical doubles), without a stream or a token vector per line.
    const char* p = text.c_str();
    const char* const end = p + text.size();
    while (p < end && *p != '\n') ++p;   // header line
    std::vector<V3> locs;
    locs.reserve(text.size() / 60);
    while (p < end) {
        while (p < end && (*p == ' ' || *p == '\t' || *p == '\r' || *p == '\n')) ++p;
        if (end - p > 6 && std::memcmp(p, "vertex", 6) == 0 && (p[6] == ' ' || p[6] == '\t')) {
            p += 6;
            char* q = nullptr;
            V3 v;
            v[0] = std::strtod(p, &q);
            p = q;
            v[1] = std::strtod(p, &q);
            p = q;
            v[2] = std::strtod(p, &q);
            p = q;
            locs.push_back(v);
        }
        while (p < end && *p != '\n') ++p;
    }
    int N_panels = (int)locs.size() / 3;
    std::vector<int> new_ind;
    collapse_duplicate_vertices(locs, vertices, new_ind);
    panels.assign(N_panels, Panel());
    for (int i = 0; i < N_panels; ++i)
        panel_init(panels[i], vertices, new_ind[3 * i], new_ind[3 * i + 1], new_ind[3 * i + 2], i, false);
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
    panels.assign(N_panels, Panel());
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
