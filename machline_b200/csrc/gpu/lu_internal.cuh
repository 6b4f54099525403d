// Shared between lu_kernels.cu (single-GPU blocked LU) and lu_sharded.cu (row-sharded LU over NCCL): the building blocks
// of the factorisation as host entry points, so that both drivers launch exactly the same kernels.
#pragma once
#include "ctx.h"

namespace mlgpu {

constexpr int LU_NB = 64;      // panel width
constexpr int LU_GEMM_BM = 128;   // rows per CTA of the trailing update (granularity of the row-block skip flags)

// Scratch of the cooperative panel kernel (candidate slots for up to num_sms CTAs, two parities) and its barrier.
struct LuPanelWork {
    DevBuf<double> pscr;
    DevBuf<int> pidx;
    DevBuf<unsigned> pbar;
    unsigned bar_base = 0;
    bool all_coop = true;   // every panel so far went through the cooperative kernel (which maintains perm)
    int gmax = 0;
    long long* dbg = nullptr;   // MACHLINE_LU_PANEL_DBG: phase stamps of lu_panel_cl2_kernel
    ml_status init(Ctx* c);
    void release();
};

// Factor the panel (rows k0..n, columns k0..k1 of dA): pivots into d_piv[k0..k1), rows interchanged inside the panel,
// perm updated (when W.all_coop stays true).
ml_status lu_panel_factor(Ctx* c, LuPanelWork& W, double* dA, int ld, int n, int k0, int k1, double* d_vv, int* d_piv, int* d_perm,
                          cudaStream_t stream, int max_ctas = 0);
// amax[i] = max_j |A(i, j)| over nr x nc (amax zeroed by the caller; bit pattern of a non-negative double)
void lu_launch_row_amax(Ctx* c, const double* A, int ld, int nr, int nc, double* amax);
// X (64 x ncols, ldx) <- L^-1 X with L the unit lower triangle of a 64 x 64 block
void lu_launch_trsm(Ctx* c, const double* L, int ldl, double* X, int ldx, int ncols);
// x[0..nb) <- solve with the nb x nb diagonal block D (ld): unit lower (upper = 0) or upper triangular (upper = 1)
void lu_launch_trsv_diag(Ctx* c, const double* D, int ld, int nb, double* x, int upper);
// C (M x Nc) -= L (M x 64 k_halves) * U (64 k_halves x Nc) on the FP64 tensor cores; all leading dimensions even, pointers
// 16-byte aligned; row_block_active: optional byte per LU_GEMM_BM rows of C (0 = skip)
void lu_gemm2_launch(Ctx* c, const double* Lp, int ldl, const double* Up, int ldu, double* Cp, int ldc, int M, int Nc,
                     const unsigned char* row_block_active, int k_halves = 1, cudaStream_t stream = nullptr);

// What the block solvers need from a row-sharded system (solve_kernels.cu: Sys): the local rows, their global indices, the
// slot a rank writes its local results to and the exchange that turns the slots of all ranks into the replicated vector.
struct RowShardOps {
    const double* A;          // local rows, column-major
    int ld, n_rows, N;
    const int* g_of_local;    // device: global row of each local row
    double* slot;             // device: this rank's part of the exchange buffer (n_rows doubles)
    void* sys;                // opaque: the Sys behind `exchange`
    ml_status (*exchange)(void* sys, double* y_full);
};
// block_jacobi_solve (common/linalg.f90:601-728) on a row-sharded system: SURVEY 8(e) row 2
ml_status block_jacobi_sharded(Ctx* c, const RowShardOps& R, const double* d_b, int block_size, double tol, double rel, int max_iter,
                               int* iters, double* d_x, double err_scale, const char* iteration_file);

}  // namespace mlgpu
