// Internal context of libmachline_gpu.so (one per device).  See include/machline_gpu.h for the ABI.
#pragma once
#include <cuda_runtime.h>

#include <string>
#include <vector>

#include "../../../include/machline_gpu.h"
#include "pair_influence.cuh"

#ifdef ML_HAVE_NCCL
#include <nccl.h>
#endif

namespace mlgpu {

struct HostPanelTable {  // deep copy of an ml_panel_soa
    int n_panels = 0, n_images = 1, n_cols = 3, in_wake = 0;
    std::vector<double> centr, A_g_to_ls, vertices_ls, n_hat_ls, b, sqrt_b, J, area, vert_g, T_mu;
    std::vector<int> r, i_vert_d, i_panel_s;
    std::vector<unsigned char> has_sources, image_present;
    // higher-order tables (ml_panel_soa.order2)
    int order2 = 0;
    std::vector<unsigned char> order;
    std::vector<int> M_dim, S_dim, i_panel_s4;
    std::vector<double> T_mu6, T_sigma;
    void copy_from(const ml_panel_soa* t);
};

// Stream-ordered allocation for solver temporaries.  ml_solve / ml_solve_dense set this for the duration of the call
// (PoolScope); DevBuf then allocates from the device's default memory pool on the context's stream, whose release
// threshold ml_ctx_create raises so that freed blocks stay cached: a solve in steady state performs no cudaMalloc /
// cudaFree (both synchronise the device and cost up to milliseconds each; a GMRES solve needs ~10 buffers).
inline thread_local cudaStream_t tl_pool_stream = nullptr;

struct PoolScope {
    cudaStream_t prev;
    explicit PoolScope(cudaStream_t s) : prev(tl_pool_stream) { tl_pool_stream = s; }
    ~PoolScope() { tl_pool_stream = prev; }
};

inline void trim_default_pool() {
    int dev = 0;
    cudaMemPool_t pool;
    if (cudaGetDevice(&dev) == cudaSuccess && cudaDeviceGetDefaultMemPool(&pool, dev) == cudaSuccess) {
        cudaDeviceSynchronize();
        cudaMemPoolTrimTo(pool, 0);
    }
}

template <typename T>
struct DevBuf {
    T* p = nullptr;
    size_t n = 0;
    cudaStream_t pool_stream = nullptr;   // non-null: allocated with cudaMallocAsync on this stream
    DevBuf() = default;
    DevBuf(const DevBuf&) = delete;        // owning: every early return releases
    DevBuf& operator=(const DevBuf&) = delete;
    ~DevBuf() { release(); }
    cudaError_t alloc(size_t count) {
        if (count <= n && p) return cudaSuccess;
        release();
        if (count == 0) return cudaSuccess;
        cudaError_t e;
        if (tl_pool_stream) {
            e = cudaMallocAsync((void**)&p, count * sizeof(T), tl_pool_stream);
            if (e == cudaSuccess) pool_stream = tl_pool_stream;
        } else {
            e = cudaMalloc((void**)&p, count * sizeof(T));
            if (e == cudaErrorMemoryAllocation) {   // cached pool blocks may be holding the memory
                cudaGetLastError();
                trim_default_pool();
                e = cudaMalloc((void**)&p, count * sizeof(T));
            }
        }
        if (e == cudaSuccess) n = count;
        else p = nullptr;
        return e;
    }
    void release() {
        if (p) {
            if (pool_stream) cudaFreeAsync(p, pool_stream);
            else cudaFree(p);
        }
        p = nullptr;
        n = 0;
        pool_stream = nullptr;
    }
};

struct AicLaunch {            // arguments of the assembly kernel (aic_kernels.cu)
    const double* recs;       // packed records, n_chunks * C records of `stride` doubles (panel_record.h)
    const unsigned char* lists;  // per-chunk scatter lists, n_chunks blocks of list_bytes(C)
    int n_chunks;             // body chunks, then wake chunks
    int tile_rows;            // R: rows owned by one CTA (32, 16 or 8)
    int ho;                   // 1: higher-order table (records carry the extension of panel_record.h, six doublet slots per pair)
    const double* cp_xyz;     // [3][n_rows_pad] control-point coordinates by local row
    const unsigned char* row_active;  // [n_rows_pad] 1: row evaluates influences
    const double* row_nB;     // nullptr (potential rows) or [3][n_rows_pad]: direction of the velocity projection of each row
    int n_rows, n_rows_pad;   // local rows, padded to 64
    double* A;                // column-major, ld rows
    int ld;
    double* I_known;          // [n_rows]
    double* W;                // wake side matrix, column-major [n_wcols][ld]
    const int* wcol;          // [n_wcols] column of A each wake column is added to
    int n_wcols;
    int n_tiles;              // n_rows_pad / R
    int* work_counter;
    FlowConst fc;             // freestream constants (kernel-parameter constant bank)
};

struct Group;   // multi.cu: the member contexts behind an ml_ctx_create_multi handle

struct Ctx {
    int device = 0;
    Group* group = nullptr;   // non-null: this handle is the facade of a single-process multi-GPU context (multi.cu)
    cudaStream_t stream = nullptr;
    cudaEvent_t ev0 = nullptr, ev1 = nullptr;
    cudaStream_t stream_hi = nullptr;        // higher-priority stream of the LU look-ahead (created on first use)
    cudaEvent_t ev_la[2] = {nullptr, nullptr};
    std::string err;
    int num_sms = 148;
    long long launches = 0;
    unsigned attr_mask = 0;   // per-context (= per-device) record of the cudaFuncSetAttribute opt-ins already made (bit per call site)
    // accounting for bench.py (ml_get_profile)
    long long h2d_bytes = 0, d2h_bytes = 0;
    bool profile = false;
    std::vector<cudaEvent_t> gemv_ev;   // start/stop pairs around gemv_n_partial launches (profiling only)
    std::vector<cudaEvent_t> comm_ev;   // (end of local matvec, end of all-gather + compaction) pairs; the first is owned by gemv_ev
    double comm_ms = 0;
    // GMRES host pipeline: pinned Hessenberg-column slots and their events, kept across solves
    double* h_pinned = nullptr;
    size_t h_pinned_n = 0;
    cudaEvent_t slot_ev[9] = {};
    long long gemv_launches = 0, gemv_bytes = 0;
    double gemv_ms = 0;

    // ---- host copies of the inputs ----
    bool have_flow = false, have_panels = false, have_cps = false, have_map = false;
    ml_flow flow{};
    HostPanelTable body, wake;
    int n_cp = 0;
    std::vector<double> cp_loc, cp_n_g;   // cp_n_g: empty, or the normals of Neumann rows
    bool velocity_rows = false;
    bool ho = false;            // higher-order body table (ml_panel_soa.order2)
    std::vector<int> cp_bc, cp_row;
    ml_system_map map{};
    std::vector<int> P, i_sigma_in_sys;
    std::vector<unsigned char> sigma_known;
    std::vector<double> sigma;
    int row0 = 0, nrows = -1;   // shard (global permuted rows): contiguous block ...
    int cyc_block = 0, cyc_rank = 0, cyc_world = 1;   // ... or, cyc_block > 0, blocks b = cyc_rank (mod cyc_world) of cyc_block rows
    std::vector<int> local_rows;   // global row of every local row (ascending), set by ml_assemble
    bool dirty = true;          // device tables need rebuilding

    // pinned staging of the packed tables (records + scatter lists): packed in place, then one async H2D each
    void* h_stage = nullptr;
    size_t h_stage_bytes = 0;

    // ---- device tables ----
    DevBuf<double> d_recs, d_cp_xyz, d_row_nB, d_A, d_I_known, d_work, d_W;
    DevBuf<unsigned char> d_row_active, d_lists;
    DevBuf<int> d_counter, d_sm_rows, d_sm_colp, d_sm_colm, d_wcol, d_zero_cols;
    int n_rec = 0, n_sm_rows = 0;
    int tile_rows = 32, chunk_records = 64, n_chunks = 0, n_wcols = 0, n_zero_cols = 0;
    int n_rows = 0, n_rows_pad = 0, ld = 0, n_cols = 0;
    bool assembled = false;
    DevBuf<double> d_x_last;        // the solution of the last ml_solve, kept on the device for ml_post_process (post.cu)
    int n_x_last = 0;
    std::vector<double> h_I_known;  // local rows
    long long pair_count = 0;
    double assemble_ms = 0, solve_ms = 0;

    // ---- multi-GPU ----
    int rank = 0, world = 1;
#ifdef ML_HAVE_NCCL
    ncclComm_t comm = nullptr;
#endif
    std::vector<int> shard_nrows;              // rows per rank
    std::vector<int> g_of_slot, slot_of_g;     // slot = rank * shard_pad + local row  <->  global row (-1: padding)
    int shard_pad = 0;
    DevBuf<int> d_g_of_slot, d_slot_of_g;
    // Peer-memory exchange of the Krylov vector (solve_kernels.cu: p2p_setup): every rank maps every other rank's window
    // (CUDA IPC); after the matvec one CTA per peer stores this rank's rows of w straight into that peer's window over NVLink
    // and raises a flag there.
    static constexpr int P2P_MAX = 8;
    static constexpr int P2P_KR = 2048;  // doubles per rank in a reduction slot of the sharded Arnoldi tail (k_max + 2 <= P2P_KR)
    unsigned xseq = 0;                  // exchanges of the sharded Arnoldi tail so far (identical on every rank)
    double* win = nullptr;              // this rank's window: [2][win_n] doubles, then [2][P2P_MAX] flags
    size_t win_n = 0;                   // doubles per parity buffer
    void* peer_base[P2P_MAX] = {};      // mapped windows (peer_base[rank] == win)
    bool peer_ipc[P2P_MAX] = {};        // mapped through CUDA IPC (another process) rather than peer access inside this process
    unsigned p2p_seq = 0;               // matvecs exchanged so far (identical on every rank)
    bool p2p_ok = false;

    ml_status fail(ml_status st, const std::string& msg) {
        err = msg;
        return st;
    }
    ml_status cuda_fail(cudaError_t e, const char* where) {
        err = std::string(where) + ": " + cudaGetErrorString(e);
        return ML_CUDA_ERROR;
    }
};

#define ML_CUDA(ctx, call)                                            \
    do {                                                              \
        cudaError_t e__ = (call);                                     \
        if (e__ != cudaSuccess) return (ctx)->cuda_fail(e__, #call);  \
    } while (0)

// post.cu
ml_status post_process(Ctx* c, const ml_post_tables* t, const ml_post_flow* f, const double* x_override, ml_post_out* out);

// aic_kernels.cu
FlowConst make_flow_const(const ml_flow& f);
cudaError_t launch_aic(Ctx* c, const AicLaunch& L, bool supersonic);
cudaError_t launch_strength_rows(Ctx* c, double* A, int ld, const int* rows, const int* colp, const int* colm, int n);
cudaError_t launch_zero_columns(Ctx* c, double* A, int ld, const int* cols, int n_cols);
cudaError_t launch_residual(Ctx* c, const double* A, int ld, int n_rows, int n_cols, const double* x, const double* b, double* partial,
                            int n_chunks, double* r);
cudaError_t launch_check_system(Ctx* c, const double* A, int ld, int n_rows, int n_cols, unsigned char* row_nz, unsigned char* col_nz,
                                int* flags);
cudaError_t launch_dod_census(Ctx* c, const double* recs, int n_rec_slots, const double* cp_xyz, const unsigned char* row_active,
                              int n_rows, int n_rows_pad, const FlowConst& fc, unsigned long long* d_counts, int stride);   // aic_sup.cu
int aic_chunk_records(int tile_rows);
int aic_record_stride(bool supersonic, bool ho = false);
cudaError_t launch_aic_subsonic_ho(Ctx* c, const AicLaunch& L);     // aic_sub_ho.cu
cudaError_t launch_aic_supersonic_ho(Ctx* c, const AicLaunch& L);   // aic_sup_ho.cu
int aic_list_bytes(int chunk_records);

// solve_kernels.cu / lu_kernels.cu
void p2p_release(Ctx* c);   // unmap the peer windows (no-op without NCCL)
ml_status solve_resident(Ctx* c, const ml_solver_opts* opts, const double* BC, double* x_out, ml_solve_info* info);
ml_status solve_dense_device(Ctx* c, int N, double* dA, int ld, const double* h_b, const ml_solver_opts* opts,
                             double* x_out, ml_solve_info* info, bool A_is_scratch);

}  // namespace mlgpu

struct ml_ctx : public mlgpu::Ctx {};

// multi.cu: fan-out of the entry points over the member contexts of an ml_ctx_create_multi handle
namespace mlgpu {
void multi_destroy(ml_ctx* c);
ml_status multi_post_process(ml_ctx* c, const ml_post_tables* t, const ml_post_flow* f, const double* x_override, ml_post_out* out);
ml_status multi_set_flow(ml_ctx* c, const ml_flow* f);
ml_status multi_set_panels(ml_ctx* c, const ml_panel_soa* body, const ml_panel_soa* wake);
ml_status multi_set_system_map(ml_ctx* c, const ml_system_map* m);
ml_status multi_set_control_points(ml_ctx* c, int n_cp, const double* loc, const int* bc, const double* n_g, const int* row_perm);
ml_status multi_assemble(ml_ctx* c, double* I_known_out, bool resident, double* device_ms);
ml_status multi_get_A(ml_ctx* c, int row0, int nrows, double* dst, int ld);
ml_status multi_set_A(ml_ctx* c, int row0, int nrows, const double* src, int ld);
ml_status multi_local_rows(ml_ctx* c, int* rows_out, int* n_out);
ml_status multi_solve(ml_ctx* c, const ml_solver_opts* opts, const double* BC, double* x_out, ml_solve_info* info);
ml_status multi_check_system(ml_ctx* c, const double* BC, int* n_zero_rows, int* n_zero_cols);
ml_status multi_residual(ml_ctx* c, const double* BC, const double* x, double* r_out);
ml_status multi_dod_census(ml_ctx* c, long long* counts4);
long long multi_sum_launches(const ml_ctx* c);
long long multi_sum_pairs(const ml_ctx* c);
ml_ctx* multi_member(const ml_ctx* c, int i);
int multi_size(const ml_ctx* c);
ml_status multi_profile(ml_ctx* c, int what, int on, ml_profile* out);
}  // namespace mlgpu

