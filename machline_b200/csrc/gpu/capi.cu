// C ABI of libmachline_gpu.so (include/machline_gpu.h): context lifecycle, input staging, the
// index resolution of panel_solver_update_system_row (src/panel_solver.f90:1203-1287) done once on
// the host while packing the panel records, and the assembly / solve entry points.
// There is no CPU fallback in this library: every compute entry point launches CUDA kernels or
// returns an error.
#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <thread>

#include "ctx.h"
#include "record_pack.h"

using namespace mlgpu;

void HostPanelTable::copy_from(const ml_panel_soa* t) {
    n_panels = t->n_panels;
    n_images = t->n_images;
    n_cols = t->n_cols;
    in_wake = t->in_wake;
    const size_t nr = (size_t)n_panels * n_images, np = (size_t)n_panels;
    // the copies are memory-bound (13 MB at 40k records): large tables are copied by a few threads, one array each
    std::vector<std::thread> pool;
    const bool par = nr >= 8192 && std::thread::hardware_concurrency() > 2;
    auto cp = [&](std::vector<double>& d, const double* s, size_t n) {
        d.resize(n);
        if (par && n >= 3 * nr) pool.emplace_back([&d, s, n] { std::memcpy(d.data(), s, n * sizeof(double)); });
        else std::memcpy(d.data(), s, n * sizeof(double));
    };
    cp(A_g_to_ls, t->A_g_to_ls, nr * 9);
    cp(vert_g, t->vert_g, nr * 9);
    cp(T_mu, t->T_mu, nr * 9);
    cp(vertices_ls, t->vertices_ls, nr * 6);
    cp(n_hat_ls, t->n_hat_ls, nr * 6);
    cp(centr, t->centr, nr * 3);
    cp(b, t->b, nr * 3);
    cp(sqrt_b, t->sqrt_b, nr * 3);
    cp(J, t->J, nr);
    cp(area, t->area, np);
    struct Join {
        std::vector<std::thread>& p;
        ~Join() {
            for (auto& th : p) th.join();
        }
    } join{pool};
    r.assign(t->r, t->r + nr);
    i_vert_d.assign(t->i_vert_d, t->i_vert_d + np * n_cols);
    if (t->i_panel_s) i_panel_s.assign(t->i_panel_s, t->i_panel_s + np);
    else i_panel_s.assign(np, -1);
    if (t->has_sources) has_sources.assign(t->has_sources, t->has_sources + np);
    else has_sources.assign(np, 0);
    if (t->image_present) image_present.assign(t->image_present, t->image_present + np);
    else image_present.assign(np, n_images > 1 ? 1 : 0);
    order2 = t->order2;
    if (order2) {
        order.assign(t->order, t->order + np);
        M_dim.assign(t->M_dim, t->M_dim + np);
        S_dim.assign(t->S_dim, t->S_dim + np);
        i_panel_s4.assign(t->i_panel_s4, t->i_panel_s4 + np * 4);
        T_mu6.assign(t->T_mu6, t->T_mu6 + nr * 36);
        T_sigma.assign(t->T_sigma, t->T_sigma + nr * 12);
    } else {
        order.clear();
        M_dim.clear();
        S_dim.clear();
        i_panel_s4.clear();
        T_mu6.clear();
        T_sigma.clear();
    }
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
    // solver temporaries come from the default stream-ordered pool and stay cached between solves (ctx.h: PoolScope)
    cudaMemPool_t pool;
    if (cudaDeviceGetDefaultMemPool(&pool, device_id) == cudaSuccess) {
        unsigned long long thr = ~0ull;
        cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &thr);
    }
    *out = c;
    return ML_OK;
}

extern "C" void ml_ctx_destroy(ml_ctx* c) {
    if (!c) return;
    if (c->group) {   // facade of a multi-GPU context: it owns no device resources itself
        mlgpu::multi_destroy(c);
        delete c;
        return;
    }
    cudaSetDevice(c->device);
    if (c->stream) cudaStreamSynchronize(c->stream);
    if (c->h_pinned) cudaFreeHost(c->h_pinned);
    if (c->h_stage) cudaFreeHost(c->h_stage);
    for (auto& e : c->slot_ev)
        if (e) cudaEventDestroy(e);
    mlgpu::p2p_release(c);
#ifdef ML_HAVE_NCCL
    if (c->comm) ncclCommDestroy(c->comm);
#endif
    c->d_x_last.release();
    c->d_recs.release();
    c->d_cp_xyz.release();
    c->d_A.release();
    c->d_I_known.release();
    c->d_work.release();
    c->d_W.release();
    c->d_lists.release();
    c->d_wcol.release();
    c->d_zero_cols.release();
    c->d_row_active.release();
    c->d_row_nB.release();
    c->d_counter.release();
    c->d_sm_rows.release();
    c->d_sm_colp.release();
    c->d_sm_colm.release();
    c->d_g_of_slot.release();
    c->d_slot_of_g.release();
    if (c->stream_hi) cudaStreamDestroy(c->stream_hi);
    for (auto& e : c->ev_la)
        if (e) cudaEventDestroy(e);
    if (c->ev0) cudaEventDestroy(c->ev0);
    if (c->ev1) cudaEventDestroy(c->ev1);
    if (c->stream) cudaStreamDestroy(c->stream);
    trim_default_pool();
    delete c;
}

extern "C" const char* ml_last_error(const ml_ctx* c) { return c ? c->err.c_str() : "null context"; }
extern "C" long long ml_launch_count(const ml_ctx* c) { return !c ? 0 : (c->group ? mlgpu::multi_sum_launches(c) : c->launches); }
extern "C" long long ml_pair_count(const ml_ctx* c) { return !c ? 0 : (c->group ? mlgpu::multi_sum_pairs(c) : c->pair_count); }

extern "C" ml_status ml_set_profiling(ml_ctx* c, int on) {
    if (!c) return ML_BAD_ARGUMENT;
    if (c->group) return mlgpu::multi_profile(c, 0, on, nullptr);
    c->profile = on != 0;
    return ML_OK;
}

static void drain_gemv_events(ml_ctx* c) {
    for (size_t i = 0; i + 1 < c->comm_ev.size(); i += 2) {   // before the gemv pairs: they own the first event of each pair
        float ms = 0.f;
        if (cudaEventElapsedTime(&ms, c->comm_ev[i], c->comm_ev[i + 1]) == cudaSuccess) c->comm_ms += ms;
        cudaEventDestroy(c->comm_ev[i + 1]);
    }
    c->comm_ev.clear();
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
    if (c->group) return mlgpu::multi_profile(c, 1, 0, out);
    cudaSetDevice(c->device);
    cudaStreamSynchronize(c->stream);
    drain_gemv_events(c);
    out->h2d_bytes = c->h2d_bytes;
    out->d2h_bytes = c->d2h_bytes;
    out->gemv_launches = c->gemv_launches;
    out->gemv_bytes = c->gemv_bytes;
    out->gemv_ms = c->gemv_ms;
    out->assemble_ms = c->assemble_ms;
    out->comm_ms = c->comm_ms;
    return ML_OK;
}

extern "C" ml_status ml_reset_profile(ml_ctx* c) {
    if (!c) return ML_BAD_ARGUMENT;
    if (c->group) return mlgpu::multi_profile(c, 2, 0, nullptr);
    cudaSetDevice(c->device);
    cudaStreamSynchronize(c->stream);
    drain_gemv_events(c);
    c->h2d_bytes = c->d2h_bytes = 0;
    c->gemv_launches = c->gemv_bytes = 0;
    c->gemv_ms = 0;
    c->comm_ms = 0;
    return ML_OK;
}

extern "C" ml_status ml_set_flow(ml_ctx* c, const ml_flow* f) {
    if (!c || !f) return ML_BAD_ARGUMENT;
    if (c->group) return mlgpu::multi_set_flow(c, f);
    c->flow = *f;
    c->have_flow = true;
    c->dirty = true;
    c->assembled = false;   // a failed re-preparation must not leave the previous system looking valid
    return ML_OK;
}

extern "C" ml_status ml_set_panels(ml_ctx* c, const ml_panel_soa* body, const ml_panel_soa* wake) {
    if (!c || !body) return ML_BAD_ARGUMENT;
    if (c->group) return mlgpu::multi_set_panels(c, body, wake);
    if (body->n_panels <= 0 || (body->n_images != 1 && body->n_images != 2)) return c->fail(ML_BAD_ARGUMENT, "body table: n_panels / n_images");
    if (body->order2) {   // higher-order table: quadratic doublets / linear sources (panel.f90:544-969)
        if (body->n_cols != 6 || !body->order || !body->M_dim || !body->S_dim || !body->i_panel_s4 || !body->T_mu6 || !body->T_sigma)
            return c->fail(ML_BAD_ARGUMENT, "higher-order body table: n_cols must be 6 and order / M_dim / S_dim / i_panel_s4 / T_mu6 / T_sigma set");
        for (int j = 0; j < body->n_panels; ++j)
            if (body->M_dim[j] < 3 || body->M_dim[j] > 6 || body->S_dim[j] < 0 || body->S_dim[j] > 4)
                return c->fail(ML_BAD_ARGUMENT, "higher-order body table: M_dim must be 3..6 and S_dim 0..4");
    } else if (body->n_cols != 3) {
        return c->fail(ML_BAD_ARGUMENT, "lower-order body table must carry 3 doublet ids per panel");
    }
    for (size_t i = 0; i < (size_t)body->n_panels * body->n_images; ++i)
        if (body->r[i] != 1) return c->fail(ML_UNSUPPORTED, "superinclined panels are not allowed (panel.f90:439-443)");
    c->body.copy_from(body);
    if (wake && wake->n_panels > 0) {
        if (wake->n_cols != 6 || wake->order2) return c->fail(ML_BAD_ARGUMENT, "wake table must be lower order with 6 doublet ids per panel");
        c->wake.copy_from(wake);
    } else {
        c->wake = HostPanelTable();
    }
    c->have_panels = true;
    c->dirty = true;
    c->assembled = false;   // a failed re-preparation must not leave the previous system looking valid
    return ML_OK;
}

extern "C" ml_status ml_set_control_points(ml_ctx* c, int n_cp, const double* loc, const int* bc, const double* n_g,
                                           const int* row_perm) {
    if (!c || n_cp <= 0 || !loc || !bc || !row_perm) return ML_BAD_ARGUMENT;
    if (c->group) return mlgpu::multi_set_control_points(c, n_cp, loc, bc, n_g, row_perm);
    c->n_cp = n_cp;
    c->cp_loc.assign(loc, loc + (size_t)3 * n_cp);
    c->cp_bc.assign(bc, bc + n_cp);
    c->cp_row.assign(row_perm, row_perm + n_cp);
    c->cp_n_g.clear();
    int n_neumann = 0;
    for (int i = 0; i < n_cp; ++i) {
        int b = bc[i];
        if (b == ML_BC_ZERO_NORMAL_MF || b == ML_BC_ZERO_NORMAL_VEL) ++n_neumann;
        else if (b != ML_BC_ZERO_POTENTIAL && b != ML_BC_SF_POTENTIAL && b != ML_BC_STRENGTH_MATCHING)
            return c->fail(ML_UNSUPPORTED, "boundary condition not built (Dirichlet, strength matching, zero normal mass flux / velocity are)");
        if (row_perm[i] < 0 || row_perm[i] >= n_cp) return c->fail(ML_BAD_ARGUMENT, "row_perm out of range");
    }
    if (n_neumann) {   // Neumann rows (panel_solver.f90:1322-1440): every row needs its normal
        if (!n_g) return c->fail(ML_BAD_ARGUMENT, "Neumann control points need n_g");
        if (n_neumann != n_cp) return c->fail(ML_UNSUPPORTED, "Neumann and Dirichlet rows cannot be mixed in one system");
        c->cp_n_g.assign(n_g, n_g + (size_t)3 * n_cp);
    }
    c->have_cps = true;
    c->dirty = true;
    c->assembled = false;   // a failed re-preparation must not leave the previous system looking valid
    return ML_OK;
}

extern "C" ml_status ml_set_system_map(ml_ctx* c, const ml_system_map* m) {
    if (!c || !m || !m->P) return ML_BAD_ARGUMENT;
    if (c->group) return mlgpu::multi_set_system_map(c, m);
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
    c->assembled = false;   // a failed re-preparation must not leave the previous system looking valid
    return ML_OK;
}

extern "C" ml_status ml_set_row_shard_cyclic(ml_ctx* c, int block_rows, int rank, int world) {
    if (!c || block_rows <= 0 || world < 1 || rank < 0 || rank >= world) return ML_BAD_ARGUMENT;
    if (c->group) return c->fail(ML_BAD_ARGUMENT, "a multi-GPU context deals its rows itself (ml_multi_set_dealing)");
    c->cyc_block = block_rows;
    c->cyc_rank = rank;
    c->cyc_world = world;
    c->dirty = true;
    c->assembled = false;   // a failed re-preparation must not leave the previous system looking valid
    return ML_OK;
}

extern "C" ml_status ml_local_rows(ml_ctx* c, int* rows_out, int* n_out) {
    if (!c || !n_out) return ML_BAD_ARGUMENT;
    if (c->group) return mlgpu::multi_local_rows(c, rows_out, n_out);
    if (!c->assembled) return c->fail(ML_NOT_READY, "ml_local_rows before ml_assemble");
    *n_out = (int)c->local_rows.size();
    if (rows_out) std::memcpy(rows_out, c->local_rows.data(), c->local_rows.size() * sizeof(int));
    return ML_OK;
}

extern "C" ml_status ml_set_row_shard(ml_ctx* c, int row0, int nrows) {
    if (!c || row0 < 0 || nrows < 0) return ML_BAD_ARGUMENT;
    if (c->group) return c->fail(ML_BAD_ARGUMENT, "a multi-GPU context deals its rows itself (ml_multi_set_dealing)");
    c->cyc_block = 0;
    c->row0 = row0;
    c->nrows = nrows;
    c->dirty = true;
    c->assembled = false;   // a failed re-preparation must not leave the previous system looking valid
    return ML_OK;
}

extern "C" ml_status ml_set_communicator(ml_ctx* c, const void* id, int rank, int world) {
    if (!c || !id || world < 1 || rank < 0 || rank >= world) return ML_BAD_ARGUMENT;
    if (c->group) return c->fail(ML_BAD_ARGUMENT, "a multi-GPU context owns its communicator");
#ifdef ML_HAVE_NCCL
    cudaSetDevice(c->device);
    ncclUniqueId uid;
    std::memcpy(&uid, id, sizeof uid);
    mlgpu::p2p_release(c);
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

static PanelView view_of(const HostPanelTable& t) {
    return PanelView{t.n_panels, t.centr.data(), t.A_g_to_ls.data(), t.vertices_ls.data(), t.n_hat_ls.data(), t.b.data(),
                     t.sqrt_b.data(), t.J.data(), t.vert_g.data(), t.T_mu.data()};
}

namespace {
struct HostRecord {   // one evaluated (panel, image) in stream order, with its final scatter targets
    const HostPanelTable* table;
    int j, img, flags, n_slots;
    double sigma_val;
    int cols[6];      // body: permuted columns of slots 0..2 (0..M_dim-1 in a higher-order table); wake: permuted columns of
                      // slots 0..5 (3..5 subtract)
    double w[3];      // higher-order table: T_sigma times the known strengths of the panel's source panels (panel_record.h)
};
}  // namespace

// Builds the device tables from the staged inputs.
static ml_status prepare(ml_ctx* c) {
    if (!(c->have_flow && c->have_panels && c->have_cps && c->have_map)) return c->fail(ML_NOT_READY, "inputs incomplete");
    ML_CUDA(c, cudaSetDevice(c->device));
    const ml_system_map& m = c->map;
    if (m.n_cp != c->n_cp) return c->fail(ML_BAD_ARGUMENT, "n_cp mismatch between control points and system map");
    const int N_verts = m.n_verts, N_panels = m.n_body_panels;
    if (N_panels != c->body.n_panels) return c->fail(ML_BAD_ARGUMENT, "n_body_panels mismatch");
    const bool sup = c->flow.supersonic != 0;
    const bool ho = c->body.order2 != 0;
    if (ho && !c->cp_n_g.empty())
        return c->fail(ML_UNSUPPORTED, "velocity (Neumann) rows with higher-order panels are not built (panel.f90:2686-2763 recursions)");
    c->ho = ho;
    const int STRIDE = aic_record_stride(sup, ho);
    const std::vector<int>& P = c->P;

    // ---- rows: this context's shard of the permuted system ----
    // local row -> global row, and its inverse for the rows this context owns
    c->local_rows.clear();
    if (c->cyc_block > 0) {
        for (int b0 = c->cyc_rank * c->cyc_block; b0 < c->n_cp; b0 += c->cyc_world * c->cyc_block)
            for (int r = b0; r < std::min(b0 + c->cyc_block, c->n_cp); ++r) c->local_rows.push_back(r);
    } else {
        const int nr = (c->nrows < 0) ? c->n_cp - c->row0 : c->nrows;
        if (c->row0 + nr > c->n_cp) return c->fail(ML_BAD_ARGUMENT, "row shard out of range");
        for (int r = 0; r < nr; ++r) c->local_rows.push_back(c->row0 + r);
    }
    const int nrows = (int)c->local_rows.size();
    std::vector<int> local_of(c->n_cp, -1);
    for (int lr = 0; lr < nrows; ++lr) local_of[c->local_rows[lr]] = lr;
    c->n_rows = nrows;
    c->n_rows_pad = ((nrows + 63) / 64) * 64;
    if (c->n_rows_pad == 0) c->n_rows_pad = 64;
    c->ld = c->n_rows_pad;
    c->n_cols = m.n_unknown;
    // Tile height: a CTA owns R rows for the whole record stream, so R trades per-chunk overhead against the
    // number of tiles available to the 2 x num_sms resident CTAs (dynamic scheduling wants several per CTA).
    {
        // Measured on B200 (r01e): subsonic, R = 8 is fastest at every size tried (N = 10.5k: 12.6 / 13.7 / 16.9 ms for
        // R = 8 / 16 / 32; N = 29k: 92.3 / 95.7 / 102.9 ms); supersonic, where a warp's DoD branches are uniform only if it
        // holds one record, R = 32 wins once there are a few tiles per resident CTA (SH_320_120: 111.7 / 111.5 / 108.3 ms).
        const long long slots = 2LL * c->num_sms;
        int R = 8;
        if (sup) {
            if (c->n_rows_pad / 32 >= 2 * slots) R = 32;
            else if (c->n_rows_pad / 16 >= 2 * slots) R = 16;
        }
        if (const char* e = std::getenv("MACHLINE_AIC_TILE_ROWS")) {
            int v = std::atoi(e);
            if (v == 4 || v == 8 || v == 16 || v == 32) R = v;
        }
        if (ho) R = 8;   // one instantiation of the higher-order kernels (aic_sub_ho.cu / aic_sup_ho.cu): 8-row tiles, 64-record chunks
        c->tile_rows = R;
        c->chunk_records = ho ? 64 : aic_chunk_records(R);
    }
    const int C = c->chunk_records;

    // ---- records in the reference's evaluation order (panel_solver.f90:1445-1476, 1656-1686) ----
    std::vector<HostRecord> body_recs, wake_recs;
    body_recs.reserve((size_t)c->body.n_panels * c->body.n_images);
    for (int j = 0; j < c->body.n_panels; ++j) {
        for (int img = 0; img < c->body.n_images; ++img) {
            if (!(c->body.area[j] > 0.)) continue;  // panel.f90:2933
            const bool mirrored_panel = (img == 1) && m.asym_flow;  // panel_solver.f90:1470-1471
            HostRecord hr{};
            hr.table = &c->body;
            hr.j = j;
            hr.img = img;
            hr.n_slots = ho ? c->body.M_dim[j] : 3;
            for (int k = 0; k < hr.n_slots; ++k) {
                int iv = c->body.i_vert_d[(size_t)j * c->body.n_cols + k], index;
                if (mirrored_panel) index = (iv >= N_verts) ? iv - N_verts : iv + N_verts;
                else index = (iv >= N_verts) ? iv - N_verts : iv;
                if (index < 0 || index >= m.n_unknown) return c->fail(ML_BAD_ARGUMENT, "doublet index out of range");
                hr.cols[k] = P[index];
            }
            hr.flags = RF_EVAL | (img == 1 ? RF_MIRROR : 0);
            if (ho) {
                // update_system_row over the panel's S_dim source panels (panel_solver.f90:1220-1253), the strengths folded
                // with T_sigma into three weights on the parameter-space influences
                hr.w[0] = hr.w[1] = hr.w[2] = 0.;
                if (c->body.has_sources[j]) {
                    const size_t rec = (size_t)j + (size_t)img * c->body.n_panels;
                    for (int k = 0; k < c->body.S_dim[j]; ++k) {
                        int ips = c->body.i_panel_s4[(size_t)j * 4 + k], index;
                        if (mirrored_panel) index = (ips >= N_panels) ? ips - N_panels : ips + N_panels;
                        else index = (ips >= N_panels) ? ips - N_panels : ips;
                        if (index < 0 || index >= m.n_sigma) return c->fail(ML_BAD_ARGUMENT, "source index out of range");
                        for (int a = 0; a < 3; ++a) hr.w[a] += c->body.T_sigma[rec * 12 + 4 * a + k] * c->sigma[index];
                    }
                    hr.flags |= RF_SOURCE;
                }
            } else if (c->body.has_sources[j]) {
                int ips = c->body.i_panel_s[j], index;
                if (mirrored_panel) index = (ips >= N_panels) ? ips - N_panels : ips + N_panels;
                else index = (ips >= N_panels) ? ips - N_panels : ips;
                if (index < 0 || index >= m.n_sigma) return c->fail(ML_BAD_ARGUMENT, "source index out of range");
                hr.sigma_val = c->sigma[index];
                hr.flags |= RF_SOURCE;
            }
            body_recs.push_back(hr);
        }
    }
    for (int l = 0; l < c->wake.n_panels; ++l) {
        for (int img = 0; img < c->wake.n_images; ++img) {
            if (img == 1 && !c->wake.image_present[l]) continue;
            if (!(c->wake.area[l] > 0.)) continue;
            HostRecord hr{};
            hr.table = &c->wake;
            hr.j = l;
            hr.img = img;
            hr.n_slots = 6;
            for (int k = 0; k < 6; ++k) {
                int iv = c->wake.i_vert_d[(size_t)l * 6 + k];
                if (iv < 0 || iv >= m.n_unknown) return c->fail(ML_BAD_ARGUMENT, "wake doublet index out of range");
                hr.cols[k] = P[iv];  // panel_solver.f90:1665-1668: no mirror shifting for wake panels
            }
            hr.flags = RF_EVAL | (img == 1 ? RF_MIRROR : 0);
            wake_recs.push_back(hr);
        }
    }
    c->n_rec = (int)(body_recs.size() + wake_recs.size());

    // ---- chunks: packed records + ordered scatter lists (panel_record.h) ----
    const int n_body_chunks = (int)((body_recs.size() + C - 1) / C), n_wake_chunks = (int)((wake_recs.size() + C - 1) / C);
    const int n_chunks = n_body_chunks + n_wake_chunks;
    const int LB = aic_list_bytes(C), MAXI = 6 * C;
    // packed straight into pinned host memory (kept by the context), so the two large H2D copies run at PCIe rate
    const size_t n_recs = (size_t)std::max(1, n_chunks) * C * STRIDE, n_lists = (size_t)std::max(1, n_chunks) * LB;
    const size_t recs_bytes = (n_recs * sizeof(double) + 255) / 256 * 256;
    if (c->h_stage_bytes < recs_bytes + n_lists) {
        if (c->h_stage) cudaFreeHost(c->h_stage);
        c->h_stage = nullptr;
        c->h_stage_bytes = 0;
        ML_CUDA(c, cudaMallocHost(&c->h_stage, recs_bytes + n_lists));
        c->h_stage_bytes = recs_bytes + n_lists;
    }
    struct RecsView {
        double* p;
        size_t n;
        double* data() const { return p; }
        size_t size() const { return n; }
    } recs{reinterpret_cast<double*>(c->h_stage), n_recs};
    struct ListsView {
        unsigned char* p;
        size_t n;
        unsigned char* data() const { return p; }
        size_t size() const { return n; }
    } lists{reinterpret_cast<unsigned char*>(c->h_stage) + recs_bytes, n_lists};
    if (n_chunks == 0) std::memset(c->h_stage, 0, recs_bytes + n_lists);
    // Phase 1 (parallel over chunks: they are independent): pack the records, group the chunk's (record, slot) items by
    // target column in order of first appearance.  cols[] holds the raw column here; phase 2 turns it into the final word.
    auto build_chunk = [&](int chunk, const std::vector<HostRecord>& src, size_t first, bool wake, int* slot_of_col) {
        const size_t n_here = std::min((size_t)C, src.size() - first);
        unsigned char* const lb = lists.data() + (size_t)chunk * LB;
        std::memset(lb, 0, LB);
        int* head = reinterpret_cast<int*>(lb);
        unsigned* cols = reinterpret_cast<unsigned*>(head + 4);
        unsigned short* beg = reinterpret_cast<unsigned short*>(head + 4 + MAXI);
        unsigned* item = reinterpret_cast<unsigned*>(beg + MAXI + 2);
        const unsigned item_scale = (unsigned)c->tile_rows * 8u;   // bytes between consecutive staged values of one row
        const int n_stage = (ho ? 6 : 3) + (sup ? 1 : 0);          // staged values per record and row (aic_kernels.cuh: SLOTS)
        // Position of a record inside the chunk: images of the panel itself first, mirror images after them (each group in
        // stream order), so that the records a warp evaluates together share the mirror flag (pair_influence.cuh).  The order
        // of ADDITION is unaffected: it is the order of the items below, which follows the stream.
        int pos[128];
        {
            int n0 = 0;
            for (size_t r = 0; r < n_here; ++r) n0 += (src[first + r].img == 0);
            int p0 = 0, p1 = n0;
            for (size_t r = 0; r < n_here; ++r) pos[r] = (src[first + r].img == 0) ? p0++ : p1++;
        }
        int order[6 * 128], count[6 * 128], n_order = 0;   // columns in order of first appearance, items per column
        for (size_t r0 = 0; r0 < n_here; ++r0) {
            const HostRecord& hr = src[first + r0];
            double* const rec_p = recs.data() + ((size_t)chunk * C + pos[r0]) * STRIDE;
            pack_record(rec_p, STRIDE, sup, view_of(*hr.table), hr.j, hr.img, hr.sigma_val, hr.flags);
            if (ho) {
                if (wake) {   // wake panels stay lower order: their 3 x 3 T_mu in the upper left corner, no sources
                    double T6[36] = {0.};
                    const double* T3 = hr.table->T_mu.data() + 9 * ((size_t)hr.j + (size_t)hr.img * hr.table->n_panels);
                    for (int a = 0; a < 3; ++a)
                        for (int bb = 0; bb < 3; ++bb) T6[6 * a + bb] = T3[3 * a + bb];
                    const double w0[3] = {0., 0., 0.};
                    pack_record_ho(rec_p, sup, T6, w0);
                } else {
                    pack_record_ho(rec_p, sup, hr.table->T_mu6.data() + 36 * ((size_t)hr.j + (size_t)hr.img * hr.table->n_panels), hr.w);
                }
            }
            for (int k = 0; k < hr.n_slots; ++k) {
                const int col = hr.cols[k];
                if (slot_of_col[col] < 0) {
                    slot_of_col[col] = n_order;
                    order[n_order] = col;
                    count[n_order] = 0;
                    ++n_order;
                }
                count[slot_of_col[col]] += 1;
            }
        }
        for (size_t r = n_here; r < (size_t)C; ++r) std::memset(recs.data() + ((size_t)chunk * C + r) * STRIDE, 0, sizeof(double) * STRIDE);
        int n_items = 0;
        for (int i = 0; i < n_order; ++i) {
            cols[i] = (unsigned)order[i];
            beg[i] = (unsigned short)n_items;
            n_items += count[i];
            count[i] = beg[i];   // fill cursor
        }
        beg[n_order] = (unsigned short)n_items;
        bool any_src = false;
        for (size_t r0 = 0; r0 < n_here; ++r0) {
            const HostRecord& hr = src[first + r0];
            any_src = any_src || (hr.flags & RF_SOURCE);
            for (int k = 0; k < hr.n_slots; ++k) {
                // a wake panel's items 3..5 (bottom side) read its slots 0..2 again, negated (panel.f90:2909-2912)
                const int slot = wake ? k % 3 : k;
                item[count[slot_of_col[hr.cols[k]]]++] =
                    (unsigned)(pos[r0] * n_stage + slot) * item_scale | ((wake && k >= 3) ? ITEM_NEG : 0u);
            }
        }
        for (int i = 0; i < n_order; ++i) slot_of_col[order[i]] = -1;
        head[0] = n_order;
        head[1] = n_items;
        head[2] = (wake ? LF_WAKE : 0) | (any_src ? LF_SOURCES : 0);
        head[3] = (int)n_here;
    };
    {
        const int n_thr = std::max(1, std::min({(int)std::thread::hardware_concurrency(), 8, n_chunks / 8 + 1}));
        auto worker = [&](int t) {
            std::vector<int> slot_of_col(m.n_unknown, -1);   // scratch: column -> position in the current chunk's list
            for (int ch = t; ch < n_chunks; ch += n_thr) {
                if (ch < n_body_chunks) build_chunk(ch, body_recs, (size_t)ch * C, false, slot_of_col.data());
                else build_chunk(ch, wake_recs, (size_t)(ch - n_body_chunks) * C, true, slot_of_col.data());
            }
        };
        std::vector<std::thread> pool;
        for (int t = 1; t < n_thr; ++t) pool.emplace_back(worker, t);
        worker(0);
        for (auto& th : pool) th.join();
    }
    // Phase 2 (sequential, cheap): the first chunk of a pass that touches a column starts its sum from zero; wake columns get
    // compact ids in order of first appearance.
    std::vector<unsigned char> col_seen(m.n_unknown, 0);
    std::vector<int> wcol_of(m.n_unknown, -1), wcols;      // compact wake column ids
    std::vector<unsigned char> wseen;
    for (int ch = 0; ch < n_chunks; ++ch) {
        int* head = reinterpret_cast<int*>(lists.data() + (size_t)ch * LB);
        unsigned* cols = reinterpret_cast<unsigned*>(head + 4);
        const bool wake = ch >= n_body_chunks;
        for (int i = 0; i < head[0]; ++i) {
            const int col = (int)cols[i];
            unsigned target;
            bool first_touch;
            if (wake) {
                if (wcol_of[col] < 0) {
                    wcol_of[col] = (int)wcols.size();
                    wcols.push_back(col);
                    wseen.push_back(0);
                }
                target = (unsigned)wcol_of[col];
                first_touch = !wseen[target];
                wseen[target] = 1;
            } else {
                target = (unsigned)col;
                first_touch = !col_seen[col];
                col_seen[col] = 1;
            }
            cols[i] = target | (first_touch ? COL_FIRST : 0u);
        }
    }
    c->n_chunks = n_chunks;
    c->n_wcols = (int)wcols.size();
    std::vector<int> zero_cols;
    for (int col = 0; col < m.n_unknown; ++col)
        if (!col_seen[col]) zero_cols.push_back(col);
    c->n_zero_cols = (int)zero_cols.size();

    // Neumann rows: the direction the induced velocity is projected on, n_g (normal velocity) or B_mat_g^T n_g (normal mass flux:
    // n . (B v) = (B^T n) . v; panel_solver.f90:1335-1336, 1409-1410)
    const bool velocity_rows = !c->cp_n_g.empty();
    std::vector<double> nB(velocity_rows ? (size_t)3 * c->n_rows_pad : 0, 0.);
    std::vector<double> xyz((size_t)3 * c->n_rows_pad, 0.);
    std::vector<unsigned char> active(c->n_rows_pad, 0);
    std::vector<int> sm_rows, sm_colp, sm_colm;
    for (int i = 0; i < c->n_cp; ++i) {
        const int row = local_of[c->cp_row[i]];
        if (row < 0) continue;
        xyz[row] = c->cp_loc[3 * (size_t)i];
        xyz[(size_t)c->n_rows_pad + row] = c->cp_loc[3 * (size_t)i + 1];
        xyz[(size_t)2 * c->n_rows_pad + row] = c->cp_loc[3 * (size_t)i + 2];
        if (velocity_rows) {
            const double* n = c->cp_n_g.data() + 3 * (size_t)i;
            for (int k = 0; k < 3; ++k) {
                double v = n[k];
                if (c->cp_bc[i] == ML_BC_ZERO_NORMAL_MF)
                    v = c->flow.B_mat_g[0 + k] * n[0] + c->flow.B_mat_g[3 + k] * n[1] + c->flow.B_mat_g[6 + k] * n[2];   // (B^T n)_k
                nB[(size_t)k * c->n_rows_pad + row] = v;
            }
        }
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

    ML_CUDA(c, c->d_recs.alloc(recs.size()));
    ML_CUDA(c, c->d_lists.alloc(lists.size()));
    ML_CUDA(c, c->d_cp_xyz.alloc(xyz.size()));
    c->velocity_rows = velocity_rows;
    if (velocity_rows) ML_CUDA(c, c->d_row_nB.alloc(nB.size()));
    ML_CUDA(c, c->d_row_active.alloc(active.size()));
    ML_CUDA(c, c->d_counter.alloc(4));
    ML_CUDA(c, c->d_I_known.alloc(c->n_rows_pad));
    ML_CUDA(c, c->d_sm_rows.alloc(sm_rows.size() + 1));
    ML_CUDA(c, c->d_sm_colp.alloc(sm_rows.size() + 1));
    ML_CUDA(c, c->d_sm_colm.alloc(sm_rows.size() + 1));
    ML_CUDA(c, c->d_wcol.alloc(wcols.size() + 1));
    ML_CUDA(c, c->d_zero_cols.alloc(zero_cols.size() + 1));
    ML_CUDA(c, c->d_W.alloc((size_t)c->ld * std::max<size_t>(1, wcols.size())));
    auto h2d = [&](void* dst, const void* src, size_t bytes) -> cudaError_t {
        c->h2d_bytes += (long long)bytes;
        return bytes ? cudaMemcpyAsync(dst, src, bytes, cudaMemcpyHostToDevice, c->stream) : cudaSuccess;
    };
    ML_CUDA(c, h2d(c->d_recs.p, recs.data(), recs.size() * sizeof(double)));
    ML_CUDA(c, h2d(c->d_lists.p, lists.data(), lists.size()));
    ML_CUDA(c, h2d(c->d_cp_xyz.p, xyz.data(), xyz.size() * sizeof(double)));
    if (velocity_rows) ML_CUDA(c, h2d(c->d_row_nB.p, nB.data(), nB.size() * sizeof(double)));
    ML_CUDA(c, h2d(c->d_row_active.p, active.data(), active.size()));
    ML_CUDA(c, h2d(c->d_sm_rows.p, sm_rows.data(), sm_rows.size() * sizeof(int)));
    ML_CUDA(c, h2d(c->d_sm_colp.p, sm_colp.data(), sm_rows.size() * sizeof(int)));
    ML_CUDA(c, h2d(c->d_sm_colm.p, sm_colm.data(), sm_rows.size() * sizeof(int)));
    ML_CUDA(c, h2d(c->d_wcol.p, wcols.data(), wcols.size() * sizeof(int)));
    ML_CUDA(c, h2d(c->d_zero_cols.p, zero_cols.data(), zero_cols.size() * sizeof(int)));
    // A: local rows x all columns, column-major
    ML_CUDA(c, c->d_A.alloc((size_t)c->ld * c->n_cols));
    ML_CUDA(c, cudaStreamSynchronize(c->stream));  // host staging vectors go out of scope
    long long active_rows = 0;
    for (int r = 0; r < nrows; ++r) active_rows += active[r];
    c->pair_count = active_rows * c->n_rec;
    c->dirty = false;
    c->assembled = false;
    return ML_OK;
}

static ml_status run_assembly_kernels(ml_ctx* c) {
    const bool sup = c->flow.supersonic != 0;
    ML_CUDA(c, cudaMemsetAsync(c->d_counter.p, 0, 4 * sizeof(int), c->stream));
    c->launches += 1;
    if (sup) {
        // the supersonic kernel skips chunks that lie outside every row's domain of dependence, so the sums cannot
        // rely on "first chunk starts from zero": zero-fill instead
        ML_CUDA(c, cudaMemsetAsync(c->d_A.p, 0, (size_t)c->ld * c->n_cols * sizeof(double), c->stream));
        ML_CUDA(c, cudaMemsetAsync(c->d_W.p, 0, (size_t)c->ld * std::max(1, c->n_wcols) * sizeof(double), c->stream));
        c->launches += 2;
    } else {
        ML_CUDA(c, launch_zero_columns(c, c->d_A.p, c->ld, c->d_zero_cols.p, c->n_zero_cols));
    }
    AicLaunch L{};
    L.recs = c->d_recs.p;
    L.lists = c->d_lists.p;
    L.n_chunks = c->n_chunks;
    L.tile_rows = c->tile_rows;
    L.ho = c->ho ? 1 : 0;
    L.cp_xyz = c->d_cp_xyz.p;
    L.row_active = c->d_row_active.p;
    L.row_nB = c->velocity_rows ? c->d_row_nB.p : nullptr;
    L.n_rows = c->n_rows;
    L.n_rows_pad = c->n_rows_pad;
    L.A = c->d_A.p;
    L.ld = c->ld;
    L.I_known = c->d_I_known.p;
    L.W = c->d_W.p;
    L.wcol = c->d_wcol.p;
    L.n_wcols = c->n_wcols;
    L.n_tiles = c->n_rows_pad / c->tile_rows;
    L.work_counter = c->d_counter.p;
    L.fc = make_flow_const(c->flow);
    if (c->n_chunks > 0) {
        ML_CUDA(c, launch_aic(c, L, sup));
    } else {
        ML_CUDA(c, cudaMemsetAsync(c->d_A.p, 0, (size_t)c->ld * c->n_cols * sizeof(double), c->stream));
        ML_CUDA(c, cudaMemsetAsync(c->d_I_known.p, 0, (size_t)c->n_rows_pad * sizeof(double), c->stream));
    }
    ML_CUDA(c, launch_strength_rows(c, c->d_A.p, c->ld, c->d_sm_rows.p, c->d_sm_colp.p, c->d_sm_colm.p, c->n_sm_rows));
    return ML_OK;
}

extern "C" ml_status ml_assemble(ml_ctx* c, double* I_known_out) {
    if (!c) return ML_BAD_ARGUMENT;
    if (c->group) return mlgpu::multi_assemble(c, I_known_out, false, nullptr);
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
    if (c->group) return mlgpu::multi_assemble(c, nullptr, true, device_ms);
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

extern "C" ml_status ml_dod_census(ml_ctx* c, long long* counts4) {
    if (!c || !counts4) return ML_BAD_ARGUMENT;
    if (c->group) return mlgpu::multi_dod_census(c, counts4);
    if (!c->assembled) return c->fail(ML_NOT_READY, "ml_dod_census before ml_assemble");
    ML_CUDA(c, cudaSetDevice(c->device));
    if (!c->flow.supersonic) {   // subsonic: every pair is evaluated with all three edges
        counts4[0] = counts4[1] = counts4[2] = 0;
        counts4[3] = c->pair_count;
        return ML_OK;
    }
    DevBuf<unsigned long long> d;
    ML_CUDA(c, d.alloc(4));
    ML_CUDA(c, cudaMemsetAsync(d.p, 0, 4 * sizeof(unsigned long long), c->stream));
    ML_CUDA(c, launch_dod_census(c, c->d_recs.p, c->n_chunks * c->chunk_records, c->d_cp_xyz.p, c->d_row_active.p, c->n_rows, c->n_rows_pad,
                                 make_flow_const(c->flow), d.p, aic_record_stride(true, c->ho)));
    unsigned long long h[4];
    ML_CUDA(c, cudaMemcpyAsync(h, d.p, sizeof h, cudaMemcpyDeviceToHost, c->stream));
    ML_CUDA(c, cudaStreamSynchronize(c->stream));
    for (int k = 0; k < 4; ++k) counts4[k] = (long long)h[k];
    return ML_OK;
}

// local index of global row `row0` if rows row0 .. row0+nrows-1 are consecutive local rows of this context, else -1
static int local_run(const ml_ctx* c, int row0, int nrows) {
    auto it = std::lower_bound(c->local_rows.begin(), c->local_rows.end(), row0);
    if (nrows == 0) return 0;
    if (it == c->local_rows.end() || *it != row0) return -1;
    const int lr = (int)(it - c->local_rows.begin());
    if (lr + nrows > (int)c->local_rows.size() || c->local_rows[lr + nrows - 1] != row0 + nrows - 1) return -1;
    return lr;
}

extern "C" ml_status ml_get_A(ml_ctx* c, int row0, int nrows, double* dst, int ld) {
    if (!c || !dst || nrows < 0 || ld < nrows) return ML_BAD_ARGUMENT;
    if (c->group) return mlgpu::multi_get_A(c, row0, nrows, dst, ld);
    if (!c->assembled) return c->fail(ML_NOT_READY, "ml_get_A before ml_assemble");
    const int lr = local_run(c, row0, nrows);
    if (lr < 0) return c->fail(ML_BAD_ARGUMENT, "rows outside this context's shard (or not consecutive in it)");
    ML_CUDA(c, cudaSetDevice(c->device));
    ML_CUDA(c, cudaMemcpy2DAsync(dst, (size_t)ld * sizeof(double), c->d_A.p + lr, (size_t)c->ld * sizeof(double),
                                 (size_t)nrows * sizeof(double), c->n_cols, cudaMemcpyDeviceToHost, c->stream));
    ML_CUDA(c, cudaStreamSynchronize(c->stream));
    return ML_OK;
}

extern "C" ml_status ml_set_A(ml_ctx* c, int row0, int nrows, const double* src, int ld) {
    if (!c || !src || nrows < 0 || ld < nrows) return ML_BAD_ARGUMENT;
    if (c->group) return mlgpu::multi_set_A(c, row0, nrows, src, ld);
    if (!c->assembled) return c->fail(ML_NOT_READY, "ml_set_A before ml_assemble");
    const int lr = local_run(c, row0, nrows);
    if (lr < 0) return c->fail(ML_BAD_ARGUMENT, "rows outside this context's shard (or not consecutive in it)");
    ML_CUDA(c, cudaSetDevice(c->device));
    ML_CUDA(c, cudaMemcpy2DAsync(c->d_A.p + lr, (size_t)c->ld * sizeof(double), src, (size_t)ld * sizeof(double),
                                 (size_t)nrows * sizeof(double), c->n_cols, cudaMemcpyHostToDevice, c->stream));
    ML_CUDA(c, cudaStreamSynchronize(c->stream));
    return ML_OK;
}

extern "C" ml_status ml_check_system(ml_ctx* c, const double* BC, int* n_zero_rows, int* n_zero_cols) {
    if (!c || !BC) return ML_BAD_ARGUMENT;
    if (c->group) return mlgpu::multi_check_system(c, BC, n_zero_rows, n_zero_cols);
    if (!c->assembled) return c->fail(ML_NOT_READY, "ml_check_system before ml_assemble");
    ML_CUDA(c, cudaSetDevice(c->device));
    if (n_zero_rows) *n_zero_rows = 0;
    if (n_zero_cols) *n_zero_cols = 0;
    // b = BC - I_known (panel_solver.f90:1818) on this context's rows
    bool nan = false;
    for (int i = 0; i < c->n_rows; ++i) {
        const double b = BC[c->local_rows[i]] - c->h_I_known[i];
        nan = nan || (b != b);
    }
    DevBuf<unsigned char> d_row, d_col;
    DevBuf<int> d_flags;
    ML_CUDA(c, d_row.alloc((size_t)c->n_rows + 1));
    ML_CUDA(c, d_col.alloc((size_t)c->n_cols + 1));
    ML_CUDA(c, d_flags.alloc(4));
    ML_CUDA(c, cudaMemsetAsync(d_row.p, 0, (size_t)c->n_rows + 1, c->stream));
    ML_CUDA(c, cudaMemsetAsync(d_col.p, 0, (size_t)c->n_cols + 1, c->stream));
    ML_CUDA(c, cudaMemsetAsync(d_flags.p, 0, 4 * sizeof(int), c->stream));
    ML_CUDA(c, launch_check_system(c, c->d_A.p, c->ld, c->n_rows, c->n_cols, d_row.p, d_col.p, d_flags.p));
    std::vector<unsigned char> row_nz(c->n_rows), col_nz(c->n_cols);
    int flags[4] = {0, 0, 0, 0};
    // flags[1] = this shard's zero rows; summed over the ranks below.  Columns: a column is zero when it is zero on every shard.
    ML_CUDA(c, cudaMemcpyAsync(row_nz.data(), d_row.p, c->n_rows, cudaMemcpyDeviceToHost, c->stream));
    ML_CUDA(c, cudaMemcpyAsync(flags, d_flags.p, sizeof flags, cudaMemcpyDeviceToHost, c->stream));
    ML_CUDA(c, cudaStreamSynchronize(c->stream));
    int zr = 0;
    for (int i = 0; i < c->n_rows; ++i) zr += !row_nz[i];
    int nan_flag = (flags[0] || nan) ? 1 : 0;
#ifdef ML_HAVE_NCCL
    if (c->world > 1) {
        int h[2] = {nan_flag, zr};
        ML_CUDA(c, cudaMemcpyAsync(d_flags.p, h, sizeof h, cudaMemcpyHostToDevice, c->stream));
        if (ncclAllReduce(d_flags.p, d_flags.p, 2, ncclInt, ncclSum, c->comm, c->stream) != ncclSuccess ||
            ncclAllReduce(d_col.p, d_col.p, c->n_cols, ncclUint8, ncclMax, c->comm, c->stream) != ncclSuccess)
            return c->fail(ML_NCCL_ERROR, "allreduce of the system check");
        ML_CUDA(c, cudaMemcpyAsync(h, d_flags.p, sizeof h, cudaMemcpyDeviceToHost, c->stream));
        ML_CUDA(c, cudaStreamSynchronize(c->stream));
        nan_flag = h[0] != 0;
        zr = h[1];
    }
#endif
    ML_CUDA(c, cudaMemcpyAsync(col_nz.data(), d_col.p, c->n_cols, cudaMemcpyDeviceToHost, c->stream));
    ML_CUDA(c, cudaStreamSynchronize(c->stream));
    int zc = 0;
    for (int j = 0; j < c->n_cols; ++j) zc += !col_nz[j];
    if (nan_flag) return c->fail(ML_NAN_IN_SYSTEM, "invalid value detected in A or b");
    if (n_zero_rows) *n_zero_rows = zr;
    if (n_zero_cols) *n_zero_cols = zc;
    if (zr || zc) {
        char msg[160];
        std::snprintf(msg, sizeof msg, "%d control point(s) not influenced, %d unknown(s) exert no influence", zr, zc);
        return c->fail(ML_UNINFLUENCED, msg);
    }
    return ML_OK;
}

extern "C" ml_status ml_residual(ml_ctx* c, const double* BC, const double* x, double* r_out) {
    if (!c || !BC || !x || !r_out) return ML_BAD_ARGUMENT;
    if (c->group) return mlgpu::multi_residual(c, BC, x, r_out);
    if (!c->assembled) return c->fail(ML_NOT_READY, "ml_residual before ml_assemble");
    ML_CUDA(c, cudaSetDevice(c->device));
    const int n_chunks = 16;
    std::vector<double> b(c->n_rows);
    for (int i = 0; i < c->n_rows; ++i) b[i] = BC[c->local_rows[i]] - c->h_I_known[i];   // panel_solver.f90:1818
    DevBuf<double> d_x, d_b, d_part, d_r;
    ML_CUDA(c, d_x.alloc(c->n_cols));
    ML_CUDA(c, d_b.alloc(c->n_rows + 1));
    ML_CUDA(c, d_part.alloc((size_t)n_chunks * c->n_rows + 1));
    ML_CUDA(c, d_r.alloc(c->n_rows + 1));
    ML_CUDA(c, cudaMemcpyAsync(d_x.p, x, sizeof(double) * c->n_cols, cudaMemcpyHostToDevice, c->stream));
    ML_CUDA(c, cudaMemcpyAsync(d_b.p, b.data(), sizeof(double) * c->n_rows, cudaMemcpyHostToDevice, c->stream));
    ML_CUDA(c, launch_residual(c, c->d_A.p, c->ld, c->n_rows, c->n_cols, d_x.p, d_b.p, d_part.p, n_chunks, d_r.p));
    ML_CUDA(c, cudaMemcpyAsync(r_out, d_r.p, sizeof(double) * c->n_rows, cudaMemcpyDeviceToHost, c->stream));
    ML_CUDA(c, cudaStreamSynchronize(c->stream));
    c->h2d_bytes += (long long)sizeof(double) * (c->n_cols + c->n_rows);
    c->d2h_bytes += (long long)sizeof(double) * c->n_rows;
    return ML_OK;
}

extern "C" ml_status ml_device_system(ml_ctx* c, double** A_dev, int* ld, int* nrows_local, int* ncols) {
    if (!c) return ML_BAD_ARGUMENT;
    if (c->group) return c->fail(ML_UNSUPPORTED, "ml_device_system: a multi-GPU context has one resident shard per device");
    if (!c->assembled) return c->fail(ML_NOT_READY, "system not assembled");
    if (A_dev) *A_dev = c->d_A.p;
    if (ld) *ld = c->ld;
    if (nrows_local) *nrows_local = c->n_rows;
    if (ncols) *ncols = c->n_cols;
    return ML_OK;
}

extern "C" ml_status ml_device_stream(ml_ctx* c, void** stream_out) {
    if (!c || !stream_out) return ML_BAD_ARGUMENT;
    if (c->group) return c->fail(ML_UNSUPPORTED, "ml_device_stream: a multi-GPU context has one stream per device");
    *stream_out = (void*)c->stream;
    return ML_OK;
}

extern "C" ml_status ml_solve(ml_ctx* c, const ml_solver_opts* opts, const double* BC, double* x_out, ml_solve_info* info) {
    if (!c || !opts || !BC || !x_out) return ML_BAD_ARGUMENT;
    if (c->group) return mlgpu::multi_solve(c, opts, BC, x_out, info);
    if (!c->assembled) return c->fail(ML_NOT_READY, "ml_solve before ml_assemble");
    ML_CUDA(c, cudaSetDevice(c->device));
    PoolScope pool(c->stream);
    return solve_resident(c, opts, BC, x_out, info);
}

extern "C" ml_status ml_solve_dense(ml_ctx* c, int N, const double* A, const double* b, const ml_solver_opts* opts,
                                    double* x_out, ml_solve_info* info) {
    if (!c || N <= 0 || !A || !b || !opts || !x_out) return ML_BAD_ARGUMENT;
    if (c->group) return ml_solve_dense(mlgpu::multi_member(c, 0), N, A, b, opts, x_out, info);   // a host system: one device
    ML_CUDA(c, cudaSetDevice(c->device));
    PoolScope pool(c->stream);
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
