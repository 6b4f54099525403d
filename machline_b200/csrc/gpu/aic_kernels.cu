// Assembly helpers that are not pair evaluations: strength-matching rows, explicit zero columns, dispatch.
#include "ctx.h"
#include "panel_record.h"

namespace mlgpu {

cudaError_t launch_aic_subsonic(Ctx* c, const AicLaunch& L);    // aic_sub.cu
cudaError_t launch_aic_supersonic(Ctx* c, const AicLaunch& L);  // aic_sup.cu

// Strength-matching rows (bc = 4): A(P(i),P(i)) = 1, A(P(i),P(i-N_cp/2)) = -1 (panel_solver.f90:1317-1320)
__global__ void strength_rows_kernel(double* A, int ld, const int* rows, const int* colp, const int* colm, int n) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) {
        A[rows[i] + (size_t)colp[i] * ld] = 1.;
        A[rows[i] + (size_t)colm[i] * ld] = -1.;
    }
}

// Columns no record feeds (none in a well-formed Dirichlet case) are zero-filled explicitly.
__global__ void zero_columns_kernel(double* A, int ld, const int* cols, int n_cols) {
    double* col = A + (size_t)cols[blockIdx.y] * ld;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < ld; i += gridDim.x * blockDim.x) col[i] = 0.;
}

// check_system (panel_solver.f90:1709-1764): one CTA per column; row_nz[r] / col_nz[c] = 1 when the row / column holds a
// non-zero, flags[0] = 1 when an entry is NaN.
__global__ void check_system_kernel(const double* __restrict__ A, int ld, int n_rows, unsigned char* row_nz, unsigned char* col_nz,
                                    int* flags) {
    const double* col = A + (size_t)blockIdx.x * ld;
    int any = 0, nan = 0;
    for (int r = threadIdx.x; r < n_rows; r += blockDim.x) {
        const double v = col[r];
        if (v != v) nan = 1;
        else if (v != 0.) {
            any = 1;
            row_nz[r] = 1;   // benign race: every writer stores the same value
        }
    }
    any = __syncthreads_or(any);
    nan = __syncthreads_or(nan);
    if (threadIdx.x == 0) {
        col_nz[blockIdx.x] = (unsigned char)any;
        if (nan) atomicOr(flags, 1);
    }
}

cudaError_t launch_check_system(Ctx* c, const double* A, int ld, int n_rows, int n_cols, unsigned char* row_nz, unsigned char* col_nz,
                                int* flags) {
    if (n_cols <= 0) return cudaSuccess;
    check_system_kernel<<<n_cols, 256, 0, c->stream>>>(A, ld, n_rows, row_nz, col_nz, flags);
    c->launches += 1;
    return cudaGetLastError();
}

FlowConst make_flow_const(const ml_flow& f) {
    FlowConst h;
    for (int i = 0; i < 3; ++i) h.c_hat[i] = f.c_hat_g[i];
    for (int i = 0; i < 9; ++i) h.C[i] = f.C_mat_g[i];
    h.K_inv = f.K_inv;
    h.s = (int)f.s;
    h.supersonic = f.supersonic;
    return h;
}

int aic_chunk_records(int tile_rows) { return tile_rows <= 8 ? 128 : 64; }
int aic_record_stride(bool supersonic, bool ho) { return record_stride(supersonic, ho); }
int aic_list_bytes(int chunk_records) { return list_bytes(chunk_records); }

cudaError_t launch_aic(Ctx* c, const AicLaunch& L, bool supersonic) {
    if (L.ho) return supersonic ? launch_aic_supersonic_ho(c, L) : launch_aic_subsonic_ho(c, L);
    return supersonic ? launch_aic_supersonic(c, L) : launch_aic_subsonic(c, L);
}

cudaError_t launch_strength_rows(Ctx* c, double* A, int ld, const int* rows, const int* colp, const int* colm, int n) {
    if (n <= 0) return cudaSuccess;
    strength_rows_kernel<<<(n + 127) / 128, 128, 0, c->stream>>>(A, ld, rows, colp, colm, n);
    c->launches += 1;
    return cudaGetLastError();
}

cudaError_t launch_zero_columns(Ctx* c, double* A, int ld, const int* cols, int n_cols) {
    if (n_cols <= 0) return cudaSuccess;
    dim3 grid((ld + 255) / 256 < 64 ? (ld + 255) / 256 : 64, n_cols);
    zero_columns_kernel<<<grid, 256, 0, c->stream>>>(A, ld, cols, n_cols);
    c->launches += 1;
    return cudaGetLastError();
}

}  // namespace mlgpu
