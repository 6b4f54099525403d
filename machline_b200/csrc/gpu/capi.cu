// C ABI of libmachline_gpu.so (include/machline_gpu.h): context lifecycle, input staging, the
// index resolution of panel_solver_update_system_row (src/panel_solver.f90:1203-1287) done once on
// the host while packing the panel records, and the assembly / solve entry points.
// There is no CPU fallback in this library: every compute entry point launches CUDA kernels or
// returns an error.
#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstring>

#include "ctx.h"

using namespace mlgpu;

void HostPanelTable::copy_from(const ml_panel_soa* t) {
    n_panels = t->n_panels;
    n_images = t->n_images;
    n_cols = t->n_cols;
    in_wake = t->in_wake;
    const size_t nr = (size_t)n_panels * n_images, np = (size_t)n_panels;
    auto cp = [](std::vector<double>& d, const double* s, size_t n) { d.assign(s, s + n); };
    cp(centr, t->centr, nr * 3);
    cp(A_g_to_ls, t->A_g_to_ls, nr * 9);
    cp(vertices_ls, t->vertices_ls, nr * 6);
    cp(n_hat_ls, t->n_hat_ls, nr * 6);
    cp(b, t->b, nr * 3);
    cp(sqrt_b, t->sqrt_b, nr * 3);
    cp(J, t->J, nr);
    cp(area, t->area, np);
    cp(vert_g, t->vert_g, nr * 9);
    cp(T_mu, t->T_mu, nr * 9);
    r.assign(t->r, t->r + nr);
    i_vert_d.assign(t->i_vert_d, t->i_vert_d + np * n_cols);
    if (t->i_panel_s) i_panel_s.assign(t->i_panel_s, t->i_panel_s + np);
    else i_panel_s.assign(np, -1);
    if (t->has_sources) has_sources.assign(t->has_sources, t->has_sources + np);
    else has_sources.assign(np, 0);
    if (t->image_present) image_present.assign(t->image_present, t->image_present + np);
    else image_present.assign(np, n_images > 1 ? 1 : 0);
}

extern "C" int ml_abi_version(void) { return 1; }

extern "C" ml_status ml_ctx_create(ml_ctx** out, int device_id) {
    if (!out) return ML_BAD_ARGUMENT;
    *out = nullptr;
    int n_dev = 0;
    cudaError_t e = cudaGetDeviceCount(&n_dev);
    if (e != cudaSuccess || n_dev == 0) return ML_CUDA_ERROR;  // no GPU: fail loudly, never fall back
    if (device_id < 0 || device_id >= n_dev) return ML_BAD_ARGUMENT;
    if (cudaSetDevice(device_id) != cudaSuccess) return ML_CUDA_ERROR;
    ml_ctx* c = new ml_ctx();
    c->device = device_id;
    cudaDeviceProp prop;
    if (cudaGetDeviceProperties(&prop, device_id) == cudaSuccess) c->num_sms = prop.multiProcessorCount;
    if (cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking) != cudaSuccess ||
        cudaEventCreate(&c->ev0) != cudaSuccess || cudaEventCreate(&c->ev1) != cudaSuccess) {
        delete c;
        return ML_CUDA_ERROR;
    }
    *out = c;
    return ML_OK;
}

extern "C" void ml_ctx_destroy(ml_ctx* c) {
    if (!c) return;
    cudaSetDevice(c->device);
#ifdef ML_HAVE_NCCL
    if (c->comm) ncclCommDestroy(c->comm);
#endif
    c->d_recs.release();
    c->d_cp_xyz.release();
    c->d_A.release();
    c->d_I_known.release();
    c->d_work.release();
    c->d_row_active.release();
    c->d_counter.release();
    c->d_sm_rows.release();
    c->d_sm_colp.release();
    c->d_sm_colm.release();
    if (c->ev0) cudaEventDestroy(c->ev0);
    if (c->ev1) cudaEventDestroy(c->ev1);
    if (c->stream) cudaStreamDestroy(c->stream);
    delete c;
}

extern "C" const char* ml_last_error(const ml_ctx* c) { return c ? c->err.c_str() : "null context"; }
extern "C" long long ml_launch_count(const ml_ctx* c) { return c ? c->launches : 0; }
extern "C" long long ml_pair_count(const ml_ctx* c) { return c ? c->pair_count : 0; }

extern "C" ml_status ml_set_profiling(ml_ctx* c, int on) {
    if (!c) return ML_BAD_ARGUMENT;
    c->profile = on != 0;
    return ML_OK;
}

static void drain_gemv_events(ml_ctx* c) {
    for (size_t i = 0; i + 1 < c->gemv_ev.size(); i += 2) {
        float ms = 0.f;
        if (cudaEventElapsedTime(&ms, c->gemv_ev[i], c->gemv_ev[i + 1]) == cudaSuccess) c->gemv_ms += ms;
        cudaEventDestroy(c->gemv_ev[i]);
        cudaEventDestroy(c->gemv_ev[i + 1]);
    }
    c->gemv_ev.clear();
}

extern "C" ml_status ml_get_profile(ml_ctx* c, ml_profile* out) {
    if (!c || !out) return ML_BAD_ARGUMENT;
    cudaSetDevice(c->device);
    cudaStreamSynchronize(c->stream);
    drain_gemv_events(c);
    out->h2d_bytes = c->h2d_bytes;
    out->d2h_bytes = c->d2h_bytes;
    out->gemv_launches = c->gemv_launches;
    out->gemv_bytes = c->gemv_bytes;
    out->gemv_ms = c->gemv_ms;
    out->assemble_ms = c->assemble_ms;
    return ML_OK;
}

extern "C" ml_status ml_reset_profile(ml_ctx* c) {
    if (!c) return ML_BAD_ARGUMENT;
    cudaSetDevice(c->device);
    cudaStreamSynchronize(c->stream);
    drain_gemv_events(c);
    c->h2d_bytes = c->d2h_bytes = 0;
    c->gemv_launches = c->gemv_bytes = 0;
    c->gemv_ms = 0;
    return ML_OK;
}

extern "C" ml_status ml_set_flow(ml_ctx* c, const ml_flow* f) {
    if (!c || !f) return ML_BAD_ARGUMENT;
    c->flow = *f;
    c->have_flow = true;
    c->dirty = true;
    return ML_OK;
}

extern "C" ml_status ml_set_panels(ml_ctx* c, const ml_panel_soa* body, const ml_panel_soa* wake) {
    if (!c || !body) return ML_BAD_ARGUMENT;
    if (body->n_panels <= 0 || (body->n_images != 1 && body->n_images != 2) || body->n_cols != 3)
        return c->fail(ML_UNSUPPORTED, "body table: only lower-order panels (M_dim = 3) are supported");
    for (size_t i = 0; i < (size_t)body->n_panels * body->n_images; ++i)
        if (body->r[i] != 1) return c->fail(ML_UNSUPPORTED, "superinclined panels are not allowed (panel.f90:439-443)");
    c->body.copy_from(body);
    if (wake && wake->n_panels > 0) {
        if (wake->n_cols != 6) return c->fail(ML_BAD_ARGUMENT, "wake table must carry 6 doublet ids per panel");
        c->wake.copy_from(wake);
    } else {
        c->wake = HostPanelTable();
    }
    c->have_panels = true;
    c->dirty = true;
    return ML_OK;
}

extern "C" ml_status ml_set_control_points(ml_ctx* c, int n_cp, const double* loc, const int* bc, const double* n_g,
                                           const int* row_perm) {
    (void)n_g;
    if (!c || n_cp <= 0 || !loc || !bc || !row_perm) return ML_BAD_ARGUMENT;
    c->n_cp = n_cp;
    c->cp_loc.assign(loc, loc + (size_t)3 * n_cp);
    c->cp_bc.assign(bc, bc + n_cp);
    c->cp_row.assign(row_perm, row_perm + n_cp);
    for (int i = 0; i < n_cp; ++i) {
        int b = bc[i];
        if (b != ML_BC_ZERO_POTENTIAL && b != ML_BC_SF_POTENTIAL && b != ML_BC_STRENGTH_MATCHING)
            return c->fail(ML_UNSUPPORTED, "only Dirichlet and strength-matching control points are supported");
        if (row_perm[i] < 0 || row_perm[i] >= n_cp) return c->fail(ML_BAD_ARGUMENT, "row_perm out of range");
    }
    c->have_cps = true;
    c->dirty = true;
    return ML_OK;
}

extern "C" ml_status ml_set_system_map(ml_ctx* c, const ml_system_map* m) {
    if (!c || !m || !m->P) return ML_BAD_ARGUMENT;
    c->map = *m;
    c->P.assign(m->P, m->P + m->n_unknown);
    c->sigma_known.assign(m->sigma_known, m->sigma_known + m->n_sigma);
    c->i_sigma_in_sys.assign(m->i_sigma_in_sys, m->i_sigma_in_sys + m->n_sigma);
    c->sigma.assign(m->sigma, m->sigma + m->n_sigma);
    for (int i = 0; i < m->n_sigma; ++i)
        if (!c->sigma_known[i]) return c->fail(ML_UNSUPPORTED, "unknown source strengths (superinclined panels) are not supported");
    c->map.P = nullptr;
    c->map.sigma_known = nullptr;
    c->map.i_sigma_in_sys = nullptr;
    c->map.sigma = nullptr;
    c->have_map = true;
    c->dirty = true;
    return ML_OK;
}

extern "C" ml_status ml_set_row_shard(ml_ctx* c, int row0, int nrows) {
    if (!c || row0 < 0 || nrows < 0) return ML_BAD_ARGUMENT;
    c->row0 = row0;
    c->nrows = nrows;
    c->dirty = true;
    return ML_OK;
}

extern "C" ml_status ml_set_communicator(ml_ctx* c, const void* id, int rank, int world) {
    if (!c || !id || world < 1 || rank < 0 || rank >= world) return ML_BAD_ARGUMENT;
#ifdef ML_HAVE_NCCL
    cudaSetDevice(c->device);
    ncclUniqueId uid;
    std::memcpy(&uid, id, sizeof uid);
    if (c->comm) {
        ncclCommDestroy(c->comm);
        c->comm = nullptr;
    }
    ncclResult_t r = ncclCommInitRank(&c->comm, world, uid, rank);
    if (r != ncclSuccess) return c->fail(ML_NCCL_ERROR, std::string("ncclCommInitRank: ") + ncclGetErrorString(r));
    c->rank = rank;
    c->world = world;
    return ML_OK;
#else
    return c->fail(ML_UNSUPPORTED, "library built without NCCL");
#endif
}

// Pack one record (see panel_record.h).  cols[] are the final permuted columns.
static void pack_record(double* rec, int rec_doubles, const HostPanelTable& t, int j, int img, const int cols[6],
                        double sigma_val, int flags) {
    const size_t r = (size_t)j + (size_t)img * t.n_panels;
    std::memset(rec, 0, sizeof(double) * rec_doubles);
    for (int k = 0; k < 3; ++k) rec[R_CENTR + k] = t.centr[3 * r + k];
    for (int k = 0; k < 9; ++k) rec[R_A + k] = t.A_g_to_ls[9 * r + k];
    for (int k = 0; k < 6; ++k) rec[R_VLS + k] = t.vertices_ls[6 * r + k];
    for (int k = 0; k < 6; ++k) rec[R_NH + k] = t.n_hat_ls[6 * r + k];
    for (int k = 0; k < 9; ++k) rec[R_T + k] = t.T_mu[9 * r + k];
    rec[R_J] = t.J[r];
    rec[R_SIGMA] = sigma_val;
    int* ci = reinterpret_cast<int*>(rec + R_COLS);
    for (int k = 0; k < 6; ++k) ci[k] = cols[k];
    int* fl = reinterpret_cast<int*>(rec + R_FLAGS);
    fl[0] = flags;
    fl[1] = 0;
    if (rec_doubles >= R_SUP_DOUBLES) {
        for (int k = 0; k < 3; ++k) rec[R_B + k] = t.b[3 * r + k];
        for (int k = 0; k < 3; ++k) rec[R_SB + k] = t.sqrt_b[3 * r + k];
        for (int k = 0; k < 9; ++k) rec[R_VG + k] = t.vert_g[9 * r + k];
    }
}

// Builds the device tables from the staged inputs.
static ml_status prepare(ml_ctx* c) {
    if (!(c->have_flow && c->have_panels && c->have_cps && c->have_map)) return c->fail(ML_NOT_READY, "inputs incomplete");
    ML_CUDA(c, cudaSetDevice(c->device));
    const ml_system_map& m = c->map;
    if (m.n_cp != c->n_cp) return c->fail(ML_BAD_ARGUMENT, "n_cp mismatch between control points and system map");
    const int N_verts = m.n_verts, N_panels = m.n_body_panels;
    if (N_panels != c->body.n_panels) return c->fail(ML_BAD_ARGUMENT, "n_body_panels mismatch");
    const bool sup = c->flow.supersonic != 0;
    const int REC = sup ? R_SUP_DOUBLES : R_SUB_DOUBLES;
    const std::vector<int>& P = c->P;

    // ---- records in the reference's evaluation order (panel_solver.f90:1445-1476, 1656-1686) ----
    std::vector<double> recs;
    recs.reserve(((size_t)c->body.n_panels * c->body.n_images + (size_t)c->wake.n_panels * c->wake.n_images) * REC);
    std::vector<double> rec(REC);
    int n_rec = 0;
    for (int j = 0; j < c->body.n_panels; ++j) {
        for (int img = 0; img < c->body.n_images; ++img) {
            if (!(c->body.area[j] > 0.)) continue;  // panel.f90:2933
            const bool mirrored_panel = (img == 1) && m.asym_flow;  // panel_solver.f90:1470-1471
            int cols[6] = {-1, -1, -1, -1, -1, -1};
            for (int k = 0; k < 3; ++k) {
                int iv = c->body.i_vert_d[(size_t)j * 3 + k], index;
                if (mirrored_panel) index = (iv >= N_verts) ? iv - N_verts : iv + N_verts;
                else index = (iv >= N_verts) ? iv - N_verts : iv;
                if (index < 0 || index >= m.n_unknown) return c->fail(ML_BAD_ARGUMENT, "doublet index out of range");
                cols[k] = P[index];
            }
            int flags = RF_EVAL | (img == 1 ? RF_MIRROR : 0);
            double sigma_val = 0.;
            if (c->body.has_sources[j]) {
                int ips = c->body.i_panel_s[j], index;
                if (mirrored_panel) index = (ips >= N_panels) ? ips - N_panels : ips + N_panels;
                else index = (ips >= N_panels) ? ips - N_panels : ips;
                if (index < 0 || index >= m.n_sigma) return c->fail(ML_BAD_ARGUMENT, "source index out of range");
                sigma_val = c->sigma[index];
                flags |= RF_SOURCE;
            }
            pack_record(rec.data(), REC, c->body, j, img, cols, sigma_val, flags);
            recs.insert(recs.end(), rec.begin(), rec.end());
            ++n_rec;
        }
    }
    for (int l = 0; l < c->wake.n_panels; ++l) {
        for (int img = 0; img < c->wake.n_images; ++img) {
            if (img == 1 && !c->wake.image_present[l]) continue;
            if (!(c->wake.area[l] > 0.)) continue;
            int cols[6];
            for (int k = 0; k < 6; ++k) {
                int iv = c->wake.i_vert_d[(size_t)l * 6 + k];
                if (iv < 0 || iv >= m.n_unknown) return c->fail(ML_BAD_ARGUMENT, "wake doublet index out of range");
                cols[k] = P[iv];  // panel_solver.f90:1665-1668: no mirror shifting for wake panels
            }
            pack_record(rec.data(), REC, c->wake, l, img, cols, 0., RF_EVAL | (img == 1 ? RF_MIRROR : 0));
            recs.insert(recs.end(), rec.begin(), rec.end());
            ++n_rec;
        }
    }
    c->n_rec = n_rec;
    c->rec_doubles = REC;

    // ---- rows: this context's shard of the permuted system ----
    const int nrows = (c->nrows < 0) ? c->n_cp - c->row0 : c->nrows;
    if (c->row0 + nrows > c->n_cp) return c->fail(ML_BAD_ARGUMENT, "row shard out of range");
    c->n_rows = nrows;
    c->n_rows_pad = ((nrows + 63) / 64) * 64;
    if (c->n_rows_pad == 0) c->n_rows_pad = 64;
    c->ld = c->n_rows_pad;
    c->n_cols = m.n_unknown;
    std::vector<double> xyz((size_t)3 * c->n_rows_pad, 0.);
    std::vector<unsigned char> active(c->n_rows_pad, 0);
    std::vector<int> sm_rows, sm_colp, sm_colm;
    for (int i = 0; i < c->n_cp; ++i) {
        int row = c->cp_row[i] - c->row0;
        if (row < 0 || row >= nrows) continue;
        xyz[row] = c->cp_loc[3 * (size_t)i];
        xyz[(size_t)c->n_rows_pad + row] = c->cp_loc[3 * (size_t)i + 1];
        xyz[(size_t)2 * c->n_rows_pad + row] = c->cp_loc[3 * (size_t)i + 2];
        if (c->cp_bc[i] == ML_BC_STRENGTH_MATCHING) {
            int half = c->n_cp / 2;
            if (i - half < 0) return c->fail(ML_BAD_ARGUMENT, "strength-matching control point in the first half");
            sm_rows.push_back(row);
            sm_colp.push_back(P[i]);
            sm_colm.push_back(P[i - half]);
        } else {
            active[row] = 1;
        }
    }
    c->n_sm_rows = (int)sm_rows.size();

    ML_CUDA(c, c->d_recs.alloc(recs.size() + 2));
    ML_CUDA(c, c->d_cp_xyz.alloc(xyz.size()));
    ML_CUDA(c, c->d_row_active.alloc(active.size()));
    ML_CUDA(c, c->d_counter.alloc(4));
    ML_CUDA(c, c->d_I_known.alloc(c->n_rows_pad));
    ML_CUDA(c, c->d_sm_rows.alloc(sm_rows.size() + 1));
    ML_CUDA(c, c->d_sm_colp.alloc(sm_rows.size() + 1));
    ML_CUDA(c, c->d_sm_colm.alloc(sm_rows.size() + 1));
    ML_CUDA(c, cudaMemcpyAsync(c->d_recs.p, recs.data(), recs.size() * sizeof(double), cudaMemcpyHostToDevice, c->stream));
    ML_CUDA(c, cudaMemcpyAsync(c->d_cp_xyz.p, xyz.data(), xyz.size() * sizeof(double), cudaMemcpyHostToDevice, c->stream));
    ML_CUDA(c, cudaMemcpyAsync(c->d_row_active.p, active.data(), active.size(), cudaMemcpyHostToDevice, c->stream));
    if (!sm_rows.empty()) {
        ML_CUDA(c, cudaMemcpyAsync(c->d_sm_rows.p, sm_rows.data(), sm_rows.size() * sizeof(int), cudaMemcpyHostToDevice, c->stream));
        ML_CUDA(c, cudaMemcpyAsync(c->d_sm_colp.p, sm_colp.data(), sm_rows.size() * sizeof(int), cudaMemcpyHostToDevice, c->stream));
        ML_CUDA(c, cudaMemcpyAsync(c->d_sm_colm.p, sm_colm.data(), sm_rows.size() * sizeof(int), cudaMemcpyHostToDevice, c->stream));
    }
    c->h2d_bytes += (long long)(recs.size() * sizeof(double) + xyz.size() * sizeof(double) + active.size() +
                                3 * sm_rows.size() * sizeof(int) + sizeof(FlowConst));
    ML_CUDA(c, upload_flow_constants(c->flow, c->stream));
    // A: local rows x all columns, column-major
    ML_CUDA(c, c->d_A.alloc((size_t)c->ld * c->n_cols));
    ML_CUDA(c, cudaStreamSynchronize(c->stream));  // host staging vectors go out of scope
    long long active_rows = 0;
    for (int r = 0; r < nrows; ++r) active_rows += active[r];
    c->pair_count = active_rows * n_rec;
    c->dirty = false;
    c->assembled = false;
    return ML_OK;
}

static ml_status run_assembly_kernels(ml_ctx* c) {
    ML_CUDA(c, cudaMemsetAsync(c->d_A.p, 0, (size_t)c->ld * c->n_cols * sizeof(double), c->stream));
    ML_CUDA(c, cudaMemsetAsync(c->d_I_known.p, 0, (size_t)c->n_rows_pad * sizeof(double), c->stream));
    ML_CUDA(c, cudaMemsetAsync(c->d_counter.p, 0, 4 * sizeof(int), c->stream));
    c->launches += 3;
    AicLaunch L{};
    L.recs = c->d_recs.p;
    L.n_rec = c->n_rec;
    L.rec_doubles = c->rec_doubles;
    L.cp_xyz = c->d_cp_xyz.p;
    L.row_active = c->d_row_active.p;
    L.n_rows = c->n_rows;
    L.A = c->d_A.p;
    L.ld = c->ld;
    L.I_known = c->d_I_known.p;
    L.n_cp_tiles = c->n_rows_pad / 32;
    const int TILE = aic_tile_records();
    L.n_tiles = (c->n_rec + TILE - 1) / TILE;
    // enough units for ~8 per resident CTA (2 CTAs/SM), never more segments than tiles
    long long want_units = (long long)c->num_sms * 2 * 8;
    int n_seg = (int)((want_units + L.n_cp_tiles - 1) / L.n_cp_tiles);
    n_seg = std::max(1, std::min(n_seg, L.n_tiles));
    L.tiles_per_seg = (L.n_tiles + n_seg - 1) / n_seg;
    L.n_segments = (L.n_tiles + L.tiles_per_seg - 1) / L.tiles_per_seg;
    L.work_counter = c->d_counter.p;
    if (L.n_tiles > 0) ML_CUDA(c, launch_aic(c, L, c->flow.supersonic != 0));
    ML_CUDA(c, launch_strength_rows(c, c->d_A.p, c->ld, c->d_sm_rows.p, c->d_sm_colp.p, c->d_sm_colm.p, c->n_sm_rows));
    return ML_OK;
}

extern "C" ml_status ml_assemble(ml_ctx* c, double* I_known_out) {
    if (!c) return ML_BAD_ARGUMENT;
    ML_CUDA(c, cudaSetDevice(c->device));
    if (c->dirty) {
        ml_status st = prepare(c);
        if (st != ML_OK) return st;
    }
    ML_CUDA(c, cudaEventRecord(c->ev0, c->stream));
    ml_status st = run_assembly_kernels(c);
    if (st != ML_OK) return st;
    ML_CUDA(c, cudaEventRecord(c->ev1, c->stream));
    c->h_I_known.assign(c->n_rows, 0.);
    ML_CUDA(c, cudaMemcpyAsync(c->h_I_known.data(), c->d_I_known.p, (size_t)c->n_rows * sizeof(double), cudaMemcpyDeviceToHost,
                               c->stream));
    c->d2h_bytes += (long long)c->n_rows * sizeof(double);
    ML_CUDA(c, cudaStreamSynchronize(c->stream));
    float ms = 0.f;
    ML_CUDA(c, cudaEventElapsedTime(&ms, c->ev0, c->ev1));
    c->assemble_ms = ms;
    c->assembled = true;
    if (I_known_out) std::memcpy(I_known_out, c->h_I_known.data(), (size_t)c->n_rows * sizeof(double));
    return ML_OK;
}

extern "C" ml_status ml_assemble_resident(ml_ctx* c, double* device_ms) {
    if (!c) return ML_BAD_ARGUMENT;
    ML_CUDA(c, cudaSetDevice(c->device));
    if (c->dirty) {
        ml_status st = prepare(c);
        if (st != ML_OK) return st;
    }
    ML_CUDA(c, cudaEventRecord(c->ev0, c->stream));
    ml_status st = run_assembly_kernels(c);
    if (st != ML_OK) return st;
    ML_CUDA(c, cudaEventRecord(c->ev1, c->stream));
    ML_CUDA(c, cudaStreamSynchronize(c->stream));
    float ms = 0.f;
    ML_CUDA(c, cudaEventElapsedTime(&ms, c->ev0, c->ev1));
    c->assemble_ms = ms;
    c->assembled = true;
    if (device_ms) *device_ms = ms;
    return ML_OK;
}

extern "C" ml_status ml_get_A(ml_ctx* c, int row0, int nrows, double* dst, int ld) {
    if (!c || !dst || nrows < 0 || ld < nrows) return ML_BAD_ARGUMENT;
    if (!c->assembled) return c->fail(ML_NOT_READY, "ml_get_A before ml_assemble");
    int lr = row0 - c->row0;
    if (lr < 0 || lr + nrows > c->n_rows) return c->fail(ML_BAD_ARGUMENT, "rows outside this context's shard");
    ML_CUDA(c, cudaSetDevice(c->device));
    ML_CUDA(c, cudaMemcpy2DAsync(dst, (size_t)ld * sizeof(double), c->d_A.p + lr, (size_t)c->ld * sizeof(double),
                                 (size_t)nrows * sizeof(double), c->n_cols, cudaMemcpyDeviceToHost, c->stream));
    ML_CUDA(c, cudaStreamSynchronize(c->stream));
    return ML_OK;
}

extern "C" ml_status ml_device_system(ml_ctx* c, double** A_dev, int* ld, int* nrows_local, int* ncols) {
    if (!c) return ML_BAD_ARGUMENT;
    if (!c->assembled) return c->fail(ML_NOT_READY, "system not assembled");
    if (A_dev) *A_dev = c->d_A.p;
    if (ld) *ld = c->ld;
    if (nrows_local) *nrows_local = c->n_rows;
    if (ncols) *ncols = c->n_cols;
    return ML_OK;
}

extern "C" ml_status ml_solve(ml_ctx* c, const ml_solver_opts* opts, const double* BC, double* x_out, ml_solve_info* info) {
    if (!c || !opts || !BC || !x_out) return ML_BAD_ARGUMENT;
    if (!c->assembled) return c->fail(ML_NOT_READY, "ml_solve before ml_assemble");
    ML_CUDA(c, cudaSetDevice(c->device));
    return solve_resident(c, opts, BC, x_out, info);
}

extern "C" ml_status ml_solve_dense(ml_ctx* c, int N, const double* A, const double* b, const ml_solver_opts* opts,
                                    double* x_out, ml_solve_info* info) {
    if (!c || N <= 0 || !A || !b || !opts || !x_out) return ML_BAD_ARGUMENT;
    ML_CUDA(c, cudaSetDevice(c->device));
    DevBuf<double> dA;
    const int ld = ((N + 63) / 64) * 64;
    ML_CUDA(c, dA.alloc((size_t)ld * N));
    ML_CUDA(c, cudaMemsetAsync(dA.p, 0, (size_t)ld * N * sizeof(double), c->stream));
    ML_CUDA(c, cudaMemcpy2DAsync(dA.p, (size_t)ld * sizeof(double), A, (size_t)N * sizeof(double), (size_t)N * sizeof(double),
                                 N, cudaMemcpyHostToDevice, c->stream));
    ml_status st = solve_dense_device(c, N, dA.p, ld, b, opts, x_out, info, true);
    dA.release();
    return st;
}
