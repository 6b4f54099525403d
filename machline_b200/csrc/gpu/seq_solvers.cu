// The reference's "sequential" direct solvers on sm_100a (single device; SURVEY 8(e): replicas only):
//   QR_givens_solve_UP                    common/linalg.f90:882-927   (matrix_solver = QRUP)
//   QR_fast_givens_solve_upper_pentagonal common/linalg.f90:1115-1165 (FQRUP; gen/apply :968-1112)
//   get_lower_bandwidth                   common/linalg.f90:797-835
//   upper_triangular_back_sub             common/linalg.f90:930-965
//   purcell_solve                         common/linalg.f90:731-794   (PURC)
//
// Every floating-point operation is issued in the reference's order with the round-to-nearest intrinsics
// (__dmul_rn / __dadd_rn / __ddiv_rn / __dsqrt_rn are never contracted into FMAs), so on the same matrix these
// solvers return the same bits as a gfortran -O2 build of the reference.  That matters for the ill-conditioned
// sorted systems they are used on (test 20: cond(A) = 4e6, golden produced by FQRUP).
//
// Parallelisation.  A Givens sweep is a dependency chain only along the pivot row j: for one column k the
// rotations (i = j+B_l .. j+1) update a(j,k) one after the other, but different columns k are independent.  So
//   * qr_chain_kernel   (1 CTA)  walks column j and generates the rotation parameters of that column in order
//                                (the chain needs only column j and, for the fast rotations, the scale vector D);
//   * qr_apply_kernel   (N-j threads) thread = column k keeps a(j,k) in a register and applies the chain to it
//                                while streaming a(i,k); the matrix is held TRANSPOSED (row i contiguous in k), so a
//                                warp's accesses to row i are coalesced.  The right-hand side is column k = N.
// The reference never rotates the last matrix column: apply_givens_row_rot is passed N-j-1 for a slice of N-j
// elements (linalg.f90:914, 1152); reproduced.
// Back substitution sums each row in ascending column order exactly like the reference, which serialises on x(i+1);
// the products are formed in parallel, the running difference is one thread's chain.
#include <algorithm>
#include <cmath>
#include <vector>

#include "ctx.h"

namespace mlgpu {

namespace {

constexpr int QR_CHUNK = 1024;   // rotations staged in shared memory at a time
constexpr int QR_APPLY_THREADS = 128;

// At[k + i*ldt] = scale * A[i + k*ld]  (k < N),  At[N + i*ldt] = b[i] * scale   (panel_solver.f90:1862, linalg.f90:1818-1826)
__global__ void __launch_bounds__(256) qr_transpose_kernel(const double* __restrict__ A, int ld, int N, const double* __restrict__ b,
                                                            const double* __restrict__ scale, double* __restrict__ At, int ldt) {
    __shared__ double tile[32][33];
    const double sc = scale ? *scale : 1.0;
    const int i0 = blockIdx.x * 32, k0 = blockIdx.y * 32;
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
    for (int r = ty; r < 32; r += 8) {
        const int i = i0 + tx, k = k0 + r;
        double v = 0.;
        if (i < N && k < N) v = A[i + (size_t)k * ld];
        tile[r][tx] = v;   // tile[k][i]
    }
    __syncthreads();
    for (int r = ty; r < 32; r += 8) {
        const int i = i0 + r, k = k0 + tx;
        if (i < N && k < N) At[k + (size_t)i * ldt] = scale ? __dmul_rn(sc, tile[tx][r]) : tile[tx][r];
    }
    if (blockIdx.y == 0 && threadIdx.x < 32) {
        const int i = i0 + threadIdx.x;
        if (i < N) At[N + (size_t)i * ldt] = scale ? __dmul_rn(b[i], sc) : b[i];
    }
}

// B_l = max over rows i of (i - first column j < i with |A(i,j)| > 1e-12)   (linalg.f90:797-835).  Warp per row.
__global__ void __launch_bounds__(256) qr_bandwidth_kernel(const double* __restrict__ At, int ldt, int N, int* __restrict__ B_l) {
    const int i = blockIdx.x * 8 + (threadIdx.x >> 5), lane = threadIdx.x & 31;
    if (i >= N) return;
    const double* row = At + (size_t)i * ldt;
    for (int j0 = 0; j0 < i; j0 += 32) {
        const int j = j0 + lane;
        const bool nz = (j < i) && fabs(row[j]) > 1.0e-12;
        const unsigned m = __ballot_sync(0xffffffffu, nz);
        if (m) {
            if (lane == 0) atomicMax(B_l, i - (j0 + __ffs(m) - 1));
            return;
        }
    }
}

// Rotation parameters of column j, generated in the reference's order (i descending).  type 0 = no rotation
// (A(i,j) == 0).  QRUP: (p, q) = (c, s), type 5.  FQRUP: (p, q) = (a, b), type 1..4 (linalg.f90:968-1071).
// diag[j] receives the rotated A(j,j); D is the fast-Givens scale vector (updated in place).
template <bool FAST>
__global__ void __launch_bounds__(256) qr_chain_kernel(const double* __restrict__ At, int ldt, int N, int j, int i_hi,
                                                        double* __restrict__ D, double* __restrict__ diag, double* __restrict__ P,
                                                        double* __restrict__ Q, int* __restrict__ T) {
    extern __shared__ double s_y[];   // A(i,j), i = j+1 .. i_hi
    const int n = i_hi - j;
    for (int t = threadIdx.x; t < n; t += blockDim.x) s_y[t] = At[j + (size_t)(j + 1 + t) * ldt];
    __syncthreads();
    if (threadIdx.x != 0) return;
    double xx = At[j + (size_t)j * ldt];
    double D1 = FAST ? D[j] : 1.;
    for (int i = i_hi; i >= j + 1; --i) {
        const double yy = s_y[i - j - 1];
        const int slot = i - j - 1;
        if (yy == 0.) {
            T[slot] = 0;
            continue;
        }
        if (!FAST) {
            // gen_givens_rot (linalg.f90:838-858)
            const double t = __dadd_rn(fabs(xx), fabs(yy));
            const double xt = __ddiv_rn(xx, t), yt = __ddiv_rn(yy, t);
            const double d = __dmul_rn(t, __dsqrt_rn(__dadd_rn(__dmul_rn(xt, xt), __dmul_rn(yt, yt))));
            P[slot] = __ddiv_rn(xx, d);
            Q[slot] = __ddiv_rn(yy, d);
            T[slot] = 5;
            xx = d;
        } else {
            // gen_fast_givens_rot (linalg.f90:968-1071)
            double D2 = D[i];
            const double gamma = __ddiv_rn(D1, D2);
            double ratio, a, b, t, d;
            int type;
            if (xx != 0.) ratio = __ddiv_rn(__dmul_rn(yy, yy), __dmul_rn(xx, xx));
            else ratio = __dadd_rn(gamma, 1.);
            if (D1 >= D2) {
                if (ratio <= gamma) {
                    type = 1;
                    t = __ddiv_rn(yy, xx);
                    b = __ddiv_rn(t, gamma);
                    d = __dadd_rn(1., __dmul_rn(b, t));
                    a = __ddiv_rn(t, d);
                    D1 = __ddiv_rn(D1, d);
                    D2 = __dmul_rn(D2, d);
                    xx = __dmul_rn(xx, d);
                } else {
                    type = 3;
                    a = __ddiv_rn(xx, yy);
                    t = __dmul_rn(a, gamma);
                    d = __dadd_rn(1., __dmul_rn(a, t));
                    b = __ddiv_rn(t, d);
                    const double temp = __dmul_rn(D2, d);
                    D2 = __ddiv_rn(D1, d);
                    D1 = temp;
                    xx = yy;
                }
            } else {
                if (ratio <= gamma) {
                    type = 2;
                    a = __ddiv_rn(yy, xx);
                    t = __ddiv_rn(a, gamma);
                    d = __dadd_rn(1., __dmul_rn(a, t));
                    b = __ddiv_rn(t, d);
                    D1 = __dmul_rn(D1, d);
                    D2 = __ddiv_rn(D2, d);
                } else {
                    type = 4;
                    t = __ddiv_rn(xx, yy);
                    b = __dmul_rn(t, gamma);
                    d = __dadd_rn(1., __dmul_rn(b, t));
                    a = __ddiv_rn(t, d);
                    const double temp = __ddiv_rn(D2, d);
                    D2 = __dmul_rn(D1, d);
                    D1 = temp;
                    xx = __dmul_rn(yy, d);
                }
            }
            D[i] = D2;
            P[slot] = a;
            Q[slot] = b;
            T[slot] = type;
        }
    }
    if (FAST) D[j] = D1;
    diag[j] = xx;
}

// Apply the chain of column j to columns k = j+1 .. N-2 and to the right-hand side (k = N).
__global__ void __launch_bounds__(QR_APPLY_THREADS) qr_apply_kernel(double* __restrict__ At, int ldt, int N, int j, int i_hi,
                                                                     const double* __restrict__ P, const double* __restrict__ Q,
                                                                     const int* __restrict__ T) {
    __shared__ double s_p[QR_CHUNK], s_q[QR_CHUNK];
    __shared__ int s_t[QR_CHUNK];
    // thread -> column: j+1 .. N-2, then the right-hand side stored at k = N
    const int n_cols = (N - 2 - j > 0 ? N - 2 - j : 0) + 1;
    const int c = blockIdx.x * QR_APPLY_THREADS + threadIdx.x;
    const bool live = c < n_cols;
    const int k = (c == n_cols - 1) ? N : j + 1 + c;
    double px = live ? At[k + (size_t)j * ldt] : 0.;
    for (int hi = i_hi; hi >= j + 1; hi -= QR_CHUNK) {
        const int lo = (hi - QR_CHUNK + 1 > j + 1) ? hi - QR_CHUNK + 1 : j + 1;
        const int n = hi - lo + 1;
        __syncthreads();
        for (int t = threadIdx.x; t < n; t += QR_APPLY_THREADS) {
            const int slot = lo + t - j - 1;
            const int ty = T[slot];
            s_t[t] = ty;
            s_p[t] = ty ? P[slot] : 0.;
            s_q[t] = ty ? Q[slot] : 0.;
        }
        __syncthreads();
        if (!live) continue;
#pragma unroll 4
        for (int i = hi; i >= lo; --i) {
            const int ty = s_t[i - lo];
            if (ty == 0) continue;
            const double a = s_p[i - lo], b = s_q[i - lo];
            double* const yp = At + k + (size_t)i * ldt;
            double py = *yp;
            switch (ty) {
                case 5: {   // apply_givens_row_rot (linalg.f90:861-879): t = c x + s y ; y = c y - s x ; x = t
                    const double t = __dadd_rn(__dmul_rn(a, px), __dmul_rn(b, py));
                    py = __dadd_rn(__dmul_rn(a, py), -__dmul_rn(b, px));
                    px = t;
                    break;
                }
                case 1:     // linalg.f90:1089-1091
                    px = __dadd_rn(px, __dmul_rn(b, py));
                    py = __dadd_rn(py, -__dmul_rn(a, px));
                    break;
                case 2:     // :1094-1096
                    py = __dadd_rn(py, -__dmul_rn(a, px));
                    px = __dadd_rn(px, __dmul_rn(b, py));
                    break;
                case 3: {   // :1099-1102
                    const double temp = py;
                    py = __dadd_rn(__dmul_rn(a, py), -px);
                    px = __dadd_rn(temp, -__dmul_rn(b, py));
                    break;
                }
                default: {  // 4, :1105-1108
                    const double temp = px;
                    px = __dadd_rn(__dmul_rn(b, px), py);
                    py = __dadd_rn(__dmul_rn(a, px), -temp);
                    break;
                }
            }
            *yp = py;
        }
    }
    if (live) At[k + (size_t)j * ldt] = px;
}

// upper_triangular_back_sub (linalg.f90:930-965) on the transposed storage: R(i,j) = At[j + i*ldt] (j > i),
// R(i,i) = diag[i], b(i) = At[N + i*ldt].  One CTA; x(i) = b(i) - R(i,i+1) x(i+1) - ... in ascending j.
__global__ void __launch_bounds__(1024) qr_back_sub_kernel(const double* __restrict__ At, int ldt, int N, const double* __restrict__ diag,
                                                            double* __restrict__ x, int* __restrict__ flag) {
    extern __shared__ double s_prod[];   // BS_CHUNK products
    constexpr int BS_CHUNK = 4096;
    __shared__ double s_acc;
    __shared__ int s_bad;
    if (threadIdx.x == 0) s_bad = 0;
    __syncthreads();
    for (int i = N - 1; i >= 0; --i) {
        const double* row = At + (size_t)i * ldt;
        if (threadIdx.x == 0) s_acc = row[N];
        for (int j0 = i + 1; j0 < N; j0 += BS_CHUNK) {
            const int n = (N - j0 < BS_CHUNK) ? N - j0 : BS_CHUNK;
            __syncthreads();   // x(i+1) of the previous row / the previous chunk's sum is complete
            for (int t = threadIdx.x; t < n; t += 1024) s_prod[t] = __dmul_rn(row[j0 + t], x[j0 + t]);
            __syncthreads();
            if (threadIdx.x == 0) {
                double acc = s_acc;
#pragma unroll 8
                for (int t = 0; t < n; ++t) acc = __dadd_rn(acc, -s_prod[t]);
                s_acc = acc;
            }
        }
        if (threadIdx.x == 0) {
            const double d = diag[i];
            if (d != 0.) x[i] = __ddiv_rn(s_acc, d);
            else s_bad = 1;
        }
        __syncthreads();
        if (s_bad) {
            if (threadIdx.x == 0) *flag = 1;
            return;
        }
    }
}

// ---- Purcell (linalg.f90:731-794) -------------------------------------------------------------------------------------
// d(k) = sum_c A(row,c) V(c,k) - b(row) V(N,k), c ascending, one accumulator per k.  One warp per 32 columns; the
// 32 x 32 tile of V is read coalesced (lane = c) and summed by lane = k from shared memory.
__global__ void __launch_bounds__(32) purcell_d_kernel(const double* __restrict__ A, int ld, int N, int row, const double* __restrict__ b,
                                                        const double* __restrict__ scale, const double* __restrict__ V, int M, int n_k,
                                                        double* __restrict__ d) {
    __shared__ double tile[32][33];
    __shared__ double s_a[32];
    const int lane = threadIdx.x, k0 = blockIdx.x * 32;
    const double sc = scale ? *scale : 1.0;
    double acc = 0.;
    for (int c0 = 0; c0 < N; c0 += 32) {
        const int c = c0 + lane;
        double a = (c < N) ? A[row + (size_t)c * ld] : 0.;
        if (scale) a = __dmul_rn(sc, a);
        s_a[lane] = a;
#pragma unroll 8
        for (int kk = 0; kk < 32; ++kk) {
            const int k = k0 + kk;
            tile[kk][lane] = (c < N && k < n_k) ? V[c + (size_t)k * M] : 0.;
        }
        __syncwarp();
        const int n = (N - c0 < 32) ? N - c0 : 32;
        for (int t = 0; t < n; ++t) acc = __dadd_rn(acc, __dmul_rn(s_a[t], tile[lane][t]));
        __syncwarp();
    }
    const int k = k0 + lane;
    if (k < n_k) {
        double bb = b[row];
        if (scale) bb = __dmul_rn(bb, sc);
        d[k] = __dadd_rn(acc, -__dmul_rn(bb, V[N + (size_t)k * M]));
    }
}

// s = first index maximising |d(k)|, k = 0 .. n_k-1 (strict ">" scan, linalg.f90:766 maxloc).  One CTA.
__global__ void __launch_bounds__(1024) purcell_argmax_kernel(const double* __restrict__ d, int n_k, int* __restrict__ s_out) {
    __shared__ double s_v[32];
    __shared__ int s_i[32];
    double best = -1.;
    int bi = 0x7fffffff;
    for (int k = threadIdx.x; k < n_k; k += 1024) {
        const double v = fabs(d[k]);
        if (v > best) {
            best = v;
            bi = k;
        }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        const double ov = __shfl_xor_sync(0xffffffffu, best, o);
        const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
        if (ov > best || (ov == best && oi < bi)) {
            best = ov;
            bi = oi;
        }
    }
    if ((threadIdx.x & 31) == 0) {
        s_v[threadIdx.x >> 5] = best;
        s_i[threadIdx.x >> 5] = bi;
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        for (int w = 1; w < 32; ++w)
            if (s_v[w] > best || (s_v[w] == best && s_i[w] < bi)) {
                best = s_v[w];
                bi = s_i[w];
            }
        *s_out = bi;
    }
}

// V_out(:,k) = alpha_k V_in(:,s) + V_in(:,m(k)),  alpha_k = -d(m(k)) * (1/d(s)),  m(k) = k (k < s) or k+1, k = 0 .. i-1
__global__ void __launch_bounds__(256) purcell_update_kernel(const double* __restrict__ Vin, double* __restrict__ Vout, int M, int n_out,
                                                              const double* __restrict__ d, const int* __restrict__ s_ptr) {
    const int k = blockIdx.x;
    if (k >= n_out) return;
    const int s = *s_ptr;
    const int mk = (k < s) ? k : k + 1;
    const double denom = __ddiv_rn(1., d[s]);
    const double alpha = __dmul_rn(-d[mk], denom);
    const double* vs = Vin + (size_t)s * M;
    const double* vm = Vin + (size_t)mk * M;
    double* vo = Vout + (size_t)k * M;
    for (int r = blockIdx.y * 256 + threadIdx.x; r < M; r += gridDim.y * 256) vo[r] = __dadd_rn(__dmul_rn(alpha, vs[r]), vm[r]);
}

__global__ void purcell_init_kernel(double* __restrict__ V, int M) {
    const size_t n = (size_t)M * M;
    for (size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x; t < n; t += (size_t)gridDim.x * blockDim.x)
        V[t] = (t / M == t % M) ? 1. : 0.;
}

__global__ void purcell_final_kernel(const double* __restrict__ V, int N, double* __restrict__ x) {
    const int r = blockIdx.x * blockDim.x + threadIdx.x;
    if (r < N) x[r] = __ddiv_rn(V[r], V[N]);
}

}  // namespace

// QRUP / FQRUP.  dA is not modified; d_scale = device pointer to 1/A(N,N) when the DIAG "preconditioner" is on.
ml_status qrup_solve_device(Ctx* c, int N, const double* dA, int ld, const double* d_b, const double* d_scale, bool fast,
                            double* d_x) {
    const int ldt = ((N + 1 + 31) / 32) * 32;
    DevBuf<double> At, D, diag, P, Q;
    DevBuf<int> T, scal;
    auto cleanup = [&]() { At.release(); D.release(); diag.release(); P.release(); Q.release(); T.release(); scal.release(); };
#define QR_CUDA(call)                                         \
    do {                                                      \
        cudaError_t e__ = (call);                             \
        if (e__ != cudaSuccess) {                             \
            cleanup();                                        \
            return c->cuda_fail(e__, #call);                  \
        }                                                     \
    } while (0)
    QR_CUDA(At.alloc((size_t)ldt * N));
    QR_CUDA(D.alloc(N));
    QR_CUDA(diag.alloc(N));
    QR_CUDA(P.alloc(N));
    QR_CUDA(Q.alloc(N));
    QR_CUDA(T.alloc(N));
    QR_CUDA(scal.alloc(2));
    dim3 tg((N + 31) / 32, (N + 31) / 32);
    qr_transpose_kernel<<<tg, 256, 0, c->stream>>>(dA, ld, N, d_b, d_scale, At.p, ldt);
    QR_CUDA(cudaMemsetAsync(scal.p, 0, 2 * sizeof(int), c->stream));
    qr_bandwidth_kernel<<<(N + 7) / 8, 256, 0, c->stream>>>(At.p, ldt, N, scal.p);
    c->launches += 2;
    int B_l = 0;
    QR_CUDA(cudaMemcpyAsync(&B_l, scal.p, sizeof(int), cudaMemcpyDeviceToHost, c->stream));
    if (fast) {
        std::vector<double> ones(N, 1.0);
        QR_CUDA(cudaMemcpyAsync(D.p, ones.data(), (size_t)N * sizeof(double), cudaMemcpyHostToDevice, c->stream));
    }
    QR_CUDA(cudaStreamSynchronize(c->stream));
    const size_t chain_smem = (size_t)std::max(1, std::min(B_l, N)) * sizeof(double);
    if (chain_smem > 200 * 1024) {
        cleanup();
        return c->fail(ML_UNSUPPORTED, "QRUP/FQRUP: lower bandwidth above 25600 rows is not supported on the device");
    }
    if (chain_smem > 48 * 1024) {
        QR_CUDA(cudaFuncSetAttribute(qr_chain_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)chain_smem));
        QR_CUDA(cudaFuncSetAttribute(qr_chain_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)chain_smem));
    }
    for (int j = 0; j < N; ++j) {
        const int i_hi = std::min(j + B_l, N - 1);
        if (i_hi <= j) {
            // no rotation in this column: the diagonal is A(j,j) itself
            QR_CUDA(cudaMemcpyAsync(diag.p + j, At.p + j + (size_t)j * ldt, sizeof(double), cudaMemcpyDeviceToDevice, c->stream));
            continue;
        }
        const size_t sm = (size_t)(i_hi - j) * sizeof(double);
        if (fast) qr_chain_kernel<true><<<1, 256, sm, c->stream>>>(At.p, ldt, N, j, i_hi, D.p, diag.p, P.p, Q.p, T.p);
        else qr_chain_kernel<false><<<1, 256, sm, c->stream>>>(At.p, ldt, N, j, i_hi, D.p, diag.p, P.p, Q.p, T.p);
        const int n_cols = std::max(0, N - 2 - j) + 1;
        qr_apply_kernel<<<(n_cols + QR_APPLY_THREADS - 1) / QR_APPLY_THREADS, QR_APPLY_THREADS, 0, c->stream>>>(At.p, ldt, N, j, i_hi,
                                                                                                                 P.p, Q.p, T.p);
        c->launches += 2;
    }
    QR_CUDA(cudaGetLastError());
    QR_CUDA(cudaFuncSetAttribute(qr_back_sub_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 4096 * 8));
    qr_back_sub_kernel<<<1, 1024, 4096 * 8, c->stream>>>(At.p, ldt, N, diag.p, d_x, scal.p + 1);
    c->launches += 1;
    int bad = 0;
    QR_CUDA(cudaMemcpyAsync(&bad, scal.p + 1, sizeof(int), cudaMemcpyDeviceToHost, c->stream));
    QR_CUDA(cudaStreamSynchronize(c->stream));
    cleanup();
#undef QR_CUDA
    if (bad) return c->fail(ML_SINGULAR, "Zero found on the diagonal of R (linalg.f90:956-961)");
    return ML_OK;
}

ml_status purcell_solve_device(Ctx* c, int N, const double* dA, int ld, const double* d_b, const double* d_scale, double* d_x) {
    const int M = N + 1;
    DevBuf<double> V0, V1, d;
    DevBuf<int> s;
    auto cleanup = [&]() { V0.release(); V1.release(); d.release(); s.release(); };
    if (V0.alloc((size_t)M * M) != cudaSuccess || V1.alloc((size_t)M * M) != cudaSuccess || d.alloc(M) != cudaSuccess ||
        s.alloc(1) != cudaSuccess) {
        cleanup();
        return c->fail(ML_CUDA_ERROR, "purcell_solve: out of device memory ((N+1)^2 workspace x 2)");
    }
    purcell_init_kernel<<<c->num_sms * 4, 256, 0, c->stream>>>(V0.p, M);
    c->launches += 1;
    double* Vin = V0.p;
    double* Vout = V1.p;
    for (int i = N; i >= 1; --i) {
        const int row = N - i;
        const int n_k = i + 1;
        purcell_d_kernel<<<(n_k + 31) / 32, 32, 0, c->stream>>>(dA, ld, N, row, d_b, d_scale, Vin, M, n_k, d.p);
        purcell_argmax_kernel<<<1, 1024, 0, c->stream>>>(d.p, n_k, s.p);
        dim3 grid(i, std::min((M + 255) / 256, 8));
        purcell_update_kernel<<<grid, 256, 0, c->stream>>>(Vin, Vout, M, i, d.p, s.p);
        c->launches += 3;
        std::swap(Vin, Vout);
    }
    purcell_final_kernel<<<(N + 255) / 256, 256, 0, c->stream>>>(Vin, N, d_x);
    c->launches += 1;
    cudaError_t e = cudaGetLastError();
    if (e == cudaSuccess) e = cudaStreamSynchronize(c->stream);
    cleanup();
    if (e != cudaSuccess) return c->cuda_fail(e, "purcell_solve");
    return ML_OK;
}

}  // namespace mlgpu
