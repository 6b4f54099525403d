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

// r = A x - b on the local rows (panel_solver.f90:1992): blockIdx.y splits the columns, partial[chunk][row]; the second kernel
// adds the chunks in order (deterministic)
__global__ void residual_partial_kernel(const double* __restrict__ A, int ld, int n_rows, int n_cols, const double* __restrict__ x,
                                        double* __restrict__ partial, int cols_per_chunk) {
    const int row = blockIdx.x * blockDim.x + threadIdx.x;
    if (row >= n_rows) return;
    const int c0 = blockIdx.y * cols_per_chunk, c1 = min(n_cols, c0 + cols_per_chunk);
    double acc = 0.;
    for (int j = c0; j < c1; ++j) acc = fma(A[row + (size_t)j * ld], x[j], acc);
    partial[(size_t)blockIdx.y * n_rows + row] = acc;
}
__global__ void residual_finish_kernel(const double* __restrict__ partial, int n_rows, int n_chunks, const double* __restrict__ b,
                                       double* __restrict__ r) {
    const int row = blockIdx.x * blockDim.x + threadIdx.x;
    if (row >= n_rows) return;
    double acc = 0.;
    for (int k = 0; k < n_chunks; ++k) acc += partial[(size_t)k * n_rows + row];
    r[row] = acc - b[row];
}

cudaError_t launch_residual(Ctx* c, const double* A, int ld, int n_rows, int n_cols, const double* x, const double* b, double* partial,
                            int n_chunks, double* r) {
    if (n_rows <= 0) return cudaSuccess;
    const int cols_per_chunk = (n_cols + n_chunks - 1) / n_chunks;
    dim3 grid((n_rows + 127) / 128, n_chunks);
    residual_partial_kernel<<<grid, 128, 0, c->stream>>>(A, ld, n_rows, n_cols, x, partial, cols_per_chunk);
    residual_finish_kernel<<<(n_rows + 127) / 128, 128, 0, c->stream>>>(partial, n_rows, n_chunks, b, r);
    c->launches += 2;
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
