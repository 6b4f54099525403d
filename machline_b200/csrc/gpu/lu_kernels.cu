// Direct solve on sm_100a: blocked right-looking LU with the reference's implicit-scaled partial
// pivoting (common/linalg.f90:166-280: vv(i) = 1/max_j|A(i,j)|, pivot = LAST row maximising
// vv(i)*|a(i,j)|), forward/back substitution (linalg.f90:283-342), and block Jacobi
// (linalg.f90:376-456, 601-728) on top of the same factorisation.
//
// The reference is an unblocked scalar Crout loop (2/3 N^3 flops through a stride-N inner access).
// Here the matrix is factored in panels of LU_NB columns; everything right of the panel is updated
// by one triangular solve and one rank-LU_NB update C -= L21 * U12 that runs on the FP64 tensor
// cores (mma.sync.m8n8k4.f64 = DMMA) -- the only dense contraction in the project.
#include <algorithm>
#include <cmath>
#include <cstdio>
#include <vector>

#include "ctx.h"

namespace mlgpu {

constexpr int LU_NB = 64;

// ---- implicit row scaling (linalg.f90:193-213) --------------------------------------------------------
__global__ void lu_row_scale_kernel(const double* __restrict__ A, int ld, int n, double* __restrict__ vv, int* __restrict__ flag) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    double amax = 0.;
    for (int j = 0; j < n; ++j) amax = fmax(amax, fabs(A[i + (size_t)j * ld]));
    if (amax <= 1.5e-20) atomicExch(flag, 1);
    vv[i] = 1.0 / amax;
}

// ---- one column of the panel: pivot search + row swap inside the panel ---------------------------------
// Single CTA.  Pivot = last row i >= j maximising vv[i]*|A(i,j)| (">=" in linalg.f90:242).
__global__ void __launch_bounds__(1024) lu_pivot_kernel(double* __restrict__ A, int ld, int n, int j, int k0, int k1,
                                                         double* __restrict__ vv, int* __restrict__ piv) {
    __shared__ double s_val[32];
    __shared__ int s_idx[32];
    __shared__ int s_p;
    double best = -1.;
    int bi = j;
    const double* col = A + (size_t)j * ld;
    for (int i = j + threadIdx.x; i < n; i += 1024) {
        double v = vv[i] * fabs(col[i]);
        if (v > best || (v == best && i > bi)) {
            best = v;
            bi = i;
        }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        double ov = __shfl_xor_sync(0xffffffffu, best, o);
        int oi = __shfl_xor_sync(0xffffffffu, bi, o);
        if (ov > best || (ov == best && oi > bi)) {
            best = ov;
            bi = oi;
        }
    }
    if ((threadIdx.x & 31) == 0) {
        s_val[threadIdx.x >> 5] = best;
        s_idx[threadIdx.x >> 5] = bi;
    }
    __syncthreads();
    if (threadIdx.x < 32) {
        best = s_val[threadIdx.x];
        bi = s_idx[threadIdx.x];
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            double ov = __shfl_xor_sync(0xffffffffu, best, o);
            int oi = __shfl_xor_sync(0xffffffffu, bi, o);
            if (ov > best || (ov == best && oi > bi)) {
                best = ov;
                bi = oi;
            }
        }
        if (threadIdx.x == 0) {
            s_p = bi;
            piv[j] = bi;
            vv[bi] = vv[j];  // linalg.f90:263
        }
    }
    __syncthreads();
    const int p = s_p;
    if (p != j) {
        for (int c = k0 + threadIdx.x; c < k1; c += 1024) {
            double t = A[p + (size_t)c * ld];
            A[p + (size_t)c * ld] = A[j + (size_t)c * ld];
            A[j + (size_t)c * ld] = t;
        }
    }
}

// rows i > j: l = A(i,j)/A(j,j); A(i,c) -= l*A(j,c) for the remaining panel columns c in (j, k1)
__global__ void __launch_bounds__(256) lu_panel_update_kernel(double* __restrict__ A, int ld, int n, int j, int k1) {
    __shared__ double s_row[LU_NB];
    const int nc = k1 - j - 1;
    for (int c = threadIdx.x; c < nc; c += 256) s_row[c] = A[j + (size_t)(j + 1 + c) * ld];
    __syncthreads();
    int i = j + 1 + blockIdx.x * 256 + threadIdx.x;
    if (i >= n) return;
    const double inv = 1.0 / A[j + (size_t)j * ld];
    const double l = A[i + (size_t)j * ld] * inv;  // linalg.f90:272-275 multiplies by the reciprocal
    A[i + (size_t)j * ld] = l;
    for (int c = 0; c < nc; ++c) A[i + (size_t)(j + 1 + c) * ld] = fma(-l, s_row[c], A[i + (size_t)(j + 1 + c) * ld]);
}

// apply the panel's row interchanges to the columns outside the panel
__global__ void __launch_bounds__(256) lu_laswp_kernel(double* __restrict__ A, int ld, int n, int k0, int k1,
                                                        const int* __restrict__ piv) {
    int c = blockIdx.x * 256 + threadIdx.x;
    if (c >= n - (k1 - k0)) return;
    if (c >= k0) c += (k1 - k0);  // skip the panel's own columns
    double* col = A + (size_t)c * ld;
    for (int j = k0; j < k1; ++j) {
        int p = piv[j];
        if (p != j) {
            double t = col[p];
            col[p] = col[j];
            col[j] = t;
        }
    }
}

// U12 = L11^{-1} A12 : one thread per column right of the panel, L11 (unit lower, nb x nb) in shared memory
__global__ void __launch_bounds__(128) lu_trsm_kernel(double* __restrict__ A, int ld, int n, int k0, int k1) {
    __shared__ double sL[LU_NB * (LU_NB + 1)];
    const int nb = k1 - k0;
    for (int t = threadIdx.x; t < nb * nb; t += 128) {
        int r = t % nb, c = t / nb;
        sL[r * (LU_NB + 1) + c] = A[(k0 + r) + (size_t)(k0 + c) * ld];
    }
    __syncthreads();
    int c = k1 + blockIdx.x * 128 + threadIdx.x;
    if (c >= n) return;
    double* col = A + (size_t)c * ld + k0;
    double x[LU_NB];
#pragma unroll 8
    for (int r = 0; r < LU_NB; ++r) x[r] = (r < nb) ? col[r] : 0.;
#pragma unroll 1
    for (int r = 1; r < nb; ++r) {
        double s = x[r];
        for (int k = 0; k < r; ++k) s = fma(-sL[r * (LU_NB + 1) + k], x[k], s);
        x[r] = s;
    }
    for (int r = 0; r < nb; ++r) col[r] = x[r];
}

// ---- trailing update on the FP64 tensor cores ---------------------------------------------------------
// C[M x Nc] -= L[M x 64] * U[64 x Nc]; CTA tile 128 x 64, 8 warps (4 along M x 2 along N), warp tile 32 x 32
// = 4 x 4 mma.m8n8k4 tiles.  Operands are staged once (K = 64 fits) in padded shared memory:
// strides = 4 (mod 16) doubles make every fragment load conflict-free.
constexpr int GM_BM = 128, GM_BN = 64, GM_K = LU_NB;
constexpr int GM_SA = GM_BM + 4;  // sA[k][m]
constexpr int GM_SB = GM_K + 4;   // sB[n][k]

__device__ __forceinline__ void dmma_m8n8k4(double& c0, double& c1, double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}

__global__ void __launch_bounds__(256) lu_gemm_dmma_kernel(double* __restrict__ A, int ld, int n, int k0, int k1) {
    extern __shared__ __align__(16) double smem[];
    double* sA = smem;                  // [GM_K][GM_SA]  holds -L21
    double* sB = smem + GM_K * GM_SA;   // [GM_BN][GM_SB] holds U12
    const int row0 = k1 + blockIdx.x * GM_BM, col0 = k1 + blockIdx.y * GM_BN;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    // stage -L21 tile (rows row0.., cols k0..k1)
    for (int t = tid; t < GM_K * GM_BM; t += 256) {
        int m = t % GM_BM, k = t / GM_BM;
        int r = row0 + m;
        sA[k * GM_SA + m] = (r < n) ? -A[r + (size_t)(k0 + k) * ld] : 0.;
    }
    // stage U12 tile (rows k0..k1, cols col0..)
    for (int t = tid; t < GM_BN * GM_K; t += 256) {
        int k = t % GM_K, nn = t / GM_K;
        int c = col0 + nn;
        sB[nn * GM_SB + k] = (c < n) ? A[(k0 + k) + (size_t)c * ld] : 0.;
    }
    __syncthreads();
    const int wm = (warp & 3) * 32, wn = (warp >> 2) * 32;
    const int g = lane >> 2, q = lane & 3;  // groupID (row of A / col of B), thread-in-group (k)
    double c[4][4][2];
#pragma unroll
    for (int mt = 0; mt < 4; ++mt)
#pragma unroll
        for (int nt = 0; nt < 4; ++nt) {
            int r = row0 + wm + mt * 8 + g;
            int cc = col0 + wn + nt * 8 + 2 * q;
            c[mt][nt][0] = (r < n && cc < n) ? A[r + (size_t)cc * ld] : 0.;
            c[mt][nt][1] = (r < n && cc + 1 < n) ? A[r + (size_t)(cc + 1) * ld] : 0.;
        }
#pragma unroll 4
    for (int ks = 0; ks < GM_K; ks += 4) {
        double a[4], b[4];
#pragma unroll
        for (int mt = 0; mt < 4; ++mt) a[mt] = sA[(ks + q) * GM_SA + wm + mt * 8 + g];
#pragma unroll
        for (int nt = 0; nt < 4; ++nt) b[nt] = sB[(wn + nt * 8 + g) * GM_SB + ks + q];
#pragma unroll
        for (int mt = 0; mt < 4; ++mt)
#pragma unroll
            for (int nt = 0; nt < 4; ++nt) dmma_m8n8k4(c[mt][nt][0], c[mt][nt][1], a[mt], b[nt]);
    }
#pragma unroll
    for (int mt = 0; mt < 4; ++mt)
#pragma unroll
        for (int nt = 0; nt < 4; ++nt) {
            int r = row0 + wm + mt * 8 + g;
            int cc = col0 + wn + nt * 8 + 2 * q;
            if (r < n && cc < n) A[r + (size_t)cc * ld] = c[mt][nt][0];
            if (r < n && cc + 1 < n) A[r + (size_t)(cc + 1) * ld] = c[mt][nt][1];
        }
}

// ---- triangular solves with the factors (blocked TRSV) --------------------------------------------------
// diagonal block: x[k0..k1) solved in place by one CTA
__global__ void __launch_bounds__(64) lu_trsv_diag_kernel(const double* __restrict__ A, int ld, int k0, int k1, double* __restrict__ x,
                                                           int upper) {
    __shared__ double sx[LU_NB];
    const int nb = k1 - k0, t = threadIdx.x;
    if (t < nb) sx[t] = x[k0 + t];
    __syncthreads();
    if (!upper) {
        for (int c = 0; c < nb; ++c) {
            if (t > c && t < nb) sx[t] = fma(-A[(k0 + t) + (size_t)(k0 + c) * ld], sx[c], sx[t]);
            __syncthreads();
        }
    } else {
        for (int c = nb - 1; c >= 0; --c) {
            if (t == c) sx[c] = sx[c] / A[(k0 + c) + (size_t)(k0 + c) * ld];
            __syncthreads();
            if (t < c) sx[t] = fma(-A[(k0 + t) + (size_t)(k0 + c) * ld], sx[c], sx[t]);
            __syncthreads();
        }
    }
    if (t < nb) x[k0 + t] = sx[t];
}

// x[r0..r1) -= A[r0..r1, k0..k1) * x[k0..k1)
__global__ void __launch_bounds__(256) lu_trsv_update_kernel(const double* __restrict__ A, int ld, int r0, int r1, int k0, int k1,
                                                              double* __restrict__ x) {
    __shared__ double sx[LU_NB];
    const int nb = k1 - k0;
    if (threadIdx.x < nb) sx[threadIdx.x] = x[k0 + threadIdx.x];
    __syncthreads();
    int r = r0 + blockIdx.x * 256 + threadIdx.x;
    if (r >= r1) return;
    double acc = 0.;
    for (int c = 0; c < nb; ++c) acc = fma(A[r + (size_t)(k0 + c) * ld], sx[c], acc);
    x[r] -= acc;
}

__global__ void lu_permute_kernel(const double* __restrict__ b, const int* __restrict__ piv, int n, double* __restrict__ x) {
    // sequential interchanges (linalg.f90:311-316 "untangle pivoting"); n is small next to the factorisation
    if (blockIdx.x == 0 && threadIdx.x == 0) {
        for (int i = 0; i < n; ++i) x[i] = b[i];
        for (int i = 0; i < n; ++i) {
            int p = piv[i];
            if (p != i) {
                double t = x[p];
                x[p] = x[i];
                x[i] = t;
            }
        }
    }
}

// ---- host drivers ---------------------------------------------------------------------------------------
// In-place LU of the n x n matrix at dA (leading dimension ld).  piv / vv are device work arrays of length n.
static ml_status lu_factor(Ctx* c, double* dA, int ld, int n, int* d_piv, double* d_vv, int* d_flag) {
    static bool attr_set = false;
    const size_t gemm_smem = (size_t)(GM_K * GM_SA + GM_BN * GM_SB) * sizeof(double);
    if (!attr_set) {
        ML_CUDA(c, cudaFuncSetAttribute(lu_gemm_dmma_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)gemm_smem));
        attr_set = true;
    }
    ML_CUDA(c, cudaMemsetAsync(d_flag, 0, sizeof(int), c->stream));
    lu_row_scale_kernel<<<(n + 127) / 128, 128, 0, c->stream>>>(dA, ld, n, d_vv, d_flag);
    c->launches += 1;
    int flag = 0;
    ML_CUDA(c, cudaMemcpyAsync(&flag, d_flag, sizeof(int), cudaMemcpyDeviceToHost, c->stream));
    ML_CUDA(c, cudaStreamSynchronize(c->stream));
    if (flag) return c->fail(ML_SINGULAR, "lu_decomp: the matrix is singular (a row is zero; linalg.f90:205-208)");
    for (int k0 = 0; k0 < n; k0 += LU_NB) {
        const int k1 = std::min(k0 + LU_NB, n);
        for (int j = k0; j < k1; ++j) {
            lu_pivot_kernel<<<1, 1024, 0, c->stream>>>(dA, ld, n, j, k0, k1, d_vv, d_piv);
            if (j + 1 < n) lu_panel_update_kernel<<<(n - j - 1 + 255) / 256, 256, 0, c->stream>>>(dA, ld, n, j, k1);
            c->launches += 2;
        }
        lu_laswp_kernel<<<(n - (k1 - k0) + 255) / 256 + 1, 256, 0, c->stream>>>(dA, ld, n, k0, k1, d_piv);
        c->launches += 1;
        if (k1 < n) {
            lu_trsm_kernel<<<(n - k1 + 127) / 128, 128, 0, c->stream>>>(dA, ld, n, k0, k1);
            c->launches += 1;
            if (k1 - k0 == LU_NB) {
                dim3 grid((n - k1 + GM_BM - 1) / GM_BM, (n - k1 + GM_BN - 1) / GM_BN);
                lu_gemm_dmma_kernel<<<grid, 256, gemm_smem, c->stream>>>(dA, ld, n, k0, k1);
                c->launches += 1;
            }
        }
        ML_CUDA(c, cudaGetLastError());
    }
    return ML_OK;
}

// x = U^{-1} L^{-1} P b with the factors in dA
static ml_status lu_substitute(Ctx* c, const double* dA, int ld, int n, const int* d_piv, const double* d_b, double* d_x) {
    lu_permute_kernel<<<1, 32, 0, c->stream>>>(d_b, d_piv, n, d_x);
    c->launches += 1;
    for (int k0 = 0; k0 < n; k0 += LU_NB) {
        const int k1 = std::min(k0 + LU_NB, n);
        lu_trsv_diag_kernel<<<1, 64, 0, c->stream>>>(dA, ld, k0, k1, d_x, 0);
        c->launches += 1;
        if (k1 < n) {
            lu_trsv_update_kernel<<<(n - k1 + 255) / 256, 256, 0, c->stream>>>(dA, ld, k1, n, k0, k1, d_x);
            c->launches += 1;
        }
    }
    const int nblk = (n + LU_NB - 1) / LU_NB;
    for (int b = nblk - 1; b >= 0; --b) {
        const int k0 = b * LU_NB, k1 = std::min(k0 + LU_NB, n);
        lu_trsv_diag_kernel<<<1, 64, 0, c->stream>>>(dA, ld, k0, k1, d_x, 1);
        c->launches += 1;
        if (k0 > 0) {
            lu_trsv_update_kernel<<<(k0 + 255) / 256, 256, 0, c->stream>>>(dA, ld, 0, k0, k0, k1, d_x);
            c->launches += 1;
        }
    }
    ML_CUDA(c, cudaGetLastError());
    return ML_OK;
}

// lu_solve (linalg.f90:118-148): dA is overwritten by its factors
ml_status lu_solve_device(Ctx* c, int N, double* dA, int ld, const double* d_b, double* d_x) {
    DevBuf<int> piv, flag;
    DevBuf<double> vv;
    ML_CUDA(c, piv.alloc(N));
    ML_CUDA(c, flag.alloc(1));
    ML_CUDA(c, vv.alloc(N));
    ml_status st = lu_factor(c, dA, ld, N, piv.p, vv.p, flag.p);
    if (st == ML_OK) st = lu_substitute(c, dA, ld, N, piv.p, d_b, d_x);
    if (st == ML_OK) {
        cudaError_t e = cudaStreamSynchronize(c->stream);
        if (e != cudaSuccess) st = c->cuda_fail(e, "lu_solve");
    }
    piv.release();
    flag.release();
    vv.release();
    return st;
}

// ---- block Jacobi (linalg.f90:601-728) --------------------------------------------------------------------
__global__ void bj_init_kernel(const double* __restrict__ A, int ld, const double* __restrict__ b, int n, double* __restrict__ x) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) x[i] = b[i] / A[i + (size_t)i * ld];  // linalg.f90:645-647
}

// out[r] = b[r] - sum_{c not in [xs,xe)} A[r,c] x[c]   for rows r in [r0,r1)
// (block right-hand side, linalg.f90:683-697, and with xs == xe the residual b - A x, :712-713)
__global__ void __launch_bounds__(256) bj_rhs_kernel(const double* __restrict__ A, int ld, int n, int r0, int r1, int xs, int xe,
                                                      const double* __restrict__ b, const double* __restrict__ x,
                                                      double* __restrict__ out) {
    __shared__ double s_part[8][33];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int r = r0 + blockIdx.x * 32 + lane;
    double acc = 0.;
    if (r < r1) {
        for (int c = warp; c < n; c += 8)
            if (c < xs || c >= xe) acc = fma(A[r + (size_t)c * ld], __ldg(x + c), acc);
    }
    s_part[warp][lane] = acc;
    __syncthreads();
    if (warp == 0 && r < r1) {
        double t = 0.;
#pragma unroll
        for (int w = 0; w < 8; ++w) t += s_part[w][lane];
        out[r] = b[r] - t;
    }
}

__global__ void bj_relax_kernel(double rel, const double* __restrict__ x, double* __restrict__ x_new, int n) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) x_new[i] = (1. - rel) * x[i] + rel * x_new[i];  // linalg.f90:709
}

__global__ void __launch_bounds__(256) bj_copy_block_kernel(const double* __restrict__ A, int ld, int s, int nb, double* __restrict__ dst,
                                                             int ldd) {
    int r = blockIdx.x * 256 + threadIdx.x;
    int cidx = blockIdx.y;
    if (r < nb) dst[r + (size_t)cidx * ldd] = A[(s + r) + (size_t)(s + cidx) * ld];
}

__global__ void __launch_bounds__(1024) vec_norm_kernel(const double* __restrict__ w, int n, double* __restrict__ out) {
    __shared__ double s_part[32];
    double acc = 0.;
    for (int i = threadIdx.x; i < n; i += 1024) acc = fma(w[i], w[i], acc);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
    if ((threadIdx.x & 31) == 0) s_part[threadIdx.x >> 5] = acc;
    __syncthreads();
    if (threadIdx.x == 0) {
        double s = 0.;
        for (int k = 0; k < 32; ++k) s += s_part[k];
        *out = sqrt(s);
    }
}

// err_scale = |1/A(N,N)| when the reference's DIAG "preconditioner" scaled the system (the stopping test is on the
// scaled residual, linalg.f90:712-716), else 1.
ml_status block_jacobi_device(Ctx* c, int N, const double* dA, int ld, const double* d_b, int block_size, double tol, double rel,
                              int max_iter, int* iters, double* d_x, double err_scale) {
    if (block_size <= 0 || block_size > N) return c->fail(ML_BAD_ARGUMENT, "block_size out of range");
    int N_blocks = N / block_size;
    if (N % block_size > 0) N_blocks += 1;
    // diagonal blocks are factored in private copies (A_blocks, linalg.f90:407-444)
    std::vector<DevBuf<double>> blocks(N_blocks);
    std::vector<DevBuf<int>> pivs(N_blocks);
    std::vector<int> bs(N_blocks), be(N_blocks), bld(N_blocks);
    DevBuf<double> vv, x_new, bi, nrm;
    DevBuf<int> flag;
    auto cleanup = [&]() {
        for (auto& b : blocks) b.release();
        for (auto& p : pivs) p.release();
        vv.release(); flag.release(); x_new.release(); bi.release(); nrm.release();
    };
    ML_CUDA(c, vv.alloc(block_size + 64));
    ML_CUDA(c, flag.alloc(1));
    ML_CUDA(c, x_new.alloc(N));
    ML_CUDA(c, bi.alloc(N));
    ML_CUDA(c, nrm.alloc(1));
    ml_status st = ML_OK;
    for (int i = 0; i < N_blocks && st == ML_OK; ++i) {
        bs[i] = i * block_size;
        be[i] = (i == N_blocks - 1) ? N : (i + 1) * block_size;
        const int nb = be[i] - bs[i];
        bld[i] = ((nb + 63) / 64) * 64;
        if (blocks[i].alloc((size_t)bld[i] * nb) != cudaSuccess || pivs[i].alloc(nb) != cudaSuccess || vv.alloc(nb) != cudaSuccess) {
            cleanup();
            return c->fail(ML_CUDA_ERROR, "block Jacobi: out of device memory");
        }
        dim3 grid((nb + 255) / 256, nb);
        bj_copy_block_kernel<<<grid, 256, 0, c->stream>>>(dA, ld, bs[i], nb, blocks[i].p, bld[i]);
        c->launches += 1;
        st = lu_factor(c, blocks[i].p, bld[i], nb, pivs[i].p, vv.p, flag.p);
    }
    const int nbk = (N + 255) / 256;
    int iteration = 0;
    double err = tol + 1.;
    if (st == ML_OK) {
        bj_init_kernel<<<nbk, 256, 0, c->stream>>>(dA, ld, d_b, N, d_x);
        c->launches += 1;
    }
    while (st == ML_OK && err >= tol && iteration < max_iter) {
        iteration += 1;
        for (int i = 0; i < N_blocks && st == ML_OK; ++i) {
            const int nb = be[i] - bs[i];
            bj_rhs_kernel<<<(nb + 31) / 32, 256, 0, c->stream>>>(dA, ld, N, bs[i], be[i], bs[i], be[i], d_b, d_x, bi.p);
            c->launches += 1;
            st = lu_substitute(c, blocks[i].p, bld[i], nb, pivs[i].p, bi.p + bs[i], x_new.p + bs[i]);
        }
        if (st != ML_OK) break;
        bj_relax_kernel<<<nbk, 256, 0, c->stream>>>(rel, d_x, x_new.p, N);
        // err = || A x_new - b ||  (linalg.f90:712-713)
        bj_rhs_kernel<<<(N + 31) / 32, 256, 0, c->stream>>>(dA, ld, N, 0, N, 0, 0, d_b, x_new.p, bi.p);
        vec_norm_kernel<<<1, 1024, 0, c->stream>>>(bi.p, N, nrm.p);
        c->launches += 3;
        cudaError_t e = cudaMemcpyAsync(d_x, x_new.p, (size_t)N * sizeof(double), cudaMemcpyDeviceToDevice, c->stream);
        if (e == cudaSuccess) e = cudaMemcpyAsync(&err, nrm.p, sizeof(double), cudaMemcpyDeviceToHost, c->stream);
        if (e == cudaSuccess) e = cudaStreamSynchronize(c->stream);
        if (e != cudaSuccess) { st = c->cuda_fail(e, "block_jacobi iteration"); break; }
        if (!(err == err)) { st = ML_NAN_RESIDUAL; break; }
        err *= err_scale;
    }
    *iters = iteration;
    cleanup();
    return st;
}

// ---- block SSOR (linalg.f90:459-598) ------------------------------------------------------------------------
// x_new(block) = (1-rel) x(block) + rel * D_block^{-1} (b - A(block, left) x_left - A(block, right) x_right), blocks visited
// first..last and last..first in alternation.  In the forward sweep the left part comes from x_new and the right part
// from x (:533-546), in the backward sweep the other way round; because x = x_new at the end of every iteration (:590)
// both are simply the current content of x_new outside the block, so one vector is kept.
__global__ void bssor_relax_kernel(double rel, const double* __restrict__ xi, double* __restrict__ x_new, int n) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) x_new[i] = (1. - rel) * x_new[i] + rel * xi[i];  // linalg.f90:577
}

ml_status block_ssor_device(Ctx* c, int N, const double* dA, int ld, const double* d_b, int block_size, double tol, double rel,
                            int max_iter, int* iters, double* d_x, double err_scale) {
    if (block_size <= 0 || block_size > N) return c->fail(ML_BAD_ARGUMENT, "block_size out of range");
    int N_blocks = N / block_size;
    if (N % block_size > 0) N_blocks += 1;
    std::vector<DevBuf<double>> blocks(N_blocks);
    std::vector<DevBuf<int>> pivs(N_blocks);
    std::vector<int> bs(N_blocks), be(N_blocks), bld(N_blocks);
    DevBuf<double> vv, xi, bi, nrm;
    DevBuf<int> flag;
    auto cleanup = [&]() {
        for (auto& b : blocks) b.release();
        for (auto& p : pivs) p.release();
        vv.release(); flag.release(); xi.release(); bi.release(); nrm.release();
    };
    if (vv.alloc(block_size + 64) != cudaSuccess || flag.alloc(1) != cudaSuccess || xi.alloc(N) != cudaSuccess ||
        bi.alloc(N) != cudaSuccess || nrm.alloc(1) != cudaSuccess) {
        cleanup();
        return c->fail(ML_CUDA_ERROR, "block SSOR: out of device memory");
    }
    ml_status st = ML_OK;
    for (int i = 0; i < N_blocks && st == ML_OK; ++i) {
        bs[i] = i * block_size;
        be[i] = (i == N_blocks - 1) ? N : (i + 1) * block_size;
        const int nb = be[i] - bs[i];
        bld[i] = ((nb + 63) / 64) * 64;
        if (blocks[i].alloc((size_t)bld[i] * nb) != cudaSuccess || pivs[i].alloc(nb) != cudaSuccess || vv.alloc(nb) != cudaSuccess) {
            cleanup();
            return c->fail(ML_CUDA_ERROR, "block SSOR: out of device memory");
        }
        dim3 grid((nb + 255) / 256, nb);
        bj_copy_block_kernel<<<grid, 256, 0, c->stream>>>(dA, ld, bs[i], nb, blocks[i].p, bld[i]);
        c->launches += 1;
        st = lu_factor(c, blocks[i].p, bld[i], nb, pivs[i].p, vv.p, flag.p);
    }
    // x = 0 (linalg.f90:493); d_x plays x_new
    if (st == ML_OK) {
        cudaError_t e = cudaMemsetAsync(d_x, 0, (size_t)N * sizeof(double), c->stream);
        if (e != cudaSuccess) st = c->cuda_fail(e, "block_ssor init");
    }
    int iteration = 0, step = -1;
    double err = tol + 1.;
    while (st == ML_OK && err >= tol && iteration < max_iter) {
        iteration += 1;
        step = -step;   // first sweep runs forward (linalg.f90:516-527)
        for (int n = 0; n < N_blocks && st == ML_OK; ++n) {
            const int i = (step == 1) ? n : N_blocks - 1 - n;
            const int nb = be[i] - bs[i];
            bj_rhs_kernel<<<(nb + 31) / 32, 256, 0, c->stream>>>(dA, ld, N, bs[i], be[i], bs[i], be[i], d_b, d_x, bi.p);
            c->launches += 1;
            st = lu_substitute(c, blocks[i].p, bld[i], nb, pivs[i].p, bi.p + bs[i], xi.p + bs[i]);
            if (st != ML_OK) break;
            bssor_relax_kernel<<<(nb + 255) / 256, 256, 0, c->stream>>>(rel, xi.p + bs[i], d_x + bs[i], nb);
            c->launches += 1;
        }
        if (st != ML_OK) break;
        bj_rhs_kernel<<<(N + 31) / 32, 256, 0, c->stream>>>(dA, ld, N, 0, N, 0, 0, d_b, d_x, bi.p);
        vec_norm_kernel<<<1, 1024, 0, c->stream>>>(bi.p, N, nrm.p);
        c->launches += 2;
        cudaError_t e = cudaMemcpyAsync(&err, nrm.p, sizeof(double), cudaMemcpyDeviceToHost, c->stream);
        if (e == cudaSuccess) e = cudaStreamSynchronize(c->stream);
        if (e != cudaSuccess) { st = c->cuda_fail(e, "block_ssor iteration"); break; }
        if (!(err == err)) { st = ML_NAN_RESIDUAL; break; }
        err *= err_scale;
    }
    *iters = iteration;
    cleanup();
    return st;
}

}  // namespace mlgpu
