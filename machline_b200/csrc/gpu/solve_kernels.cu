// Dense solve on sm_100a: replaces the select-case of panel_solver_solve_system
// (src/panel_solver.f90:1802-2027) and the iterative solvers of common/linalg.f90:
//   diagonal_preconditioner  :1798-1831  (reproduced as the uniform scale 1/A(N,N) it really is, folded
//                                          into the matvec instead of copying A into A_p)
//   arnoldi_update / GMRES   :1208-1334
//   restarted_GMRES          :1337-1453
//   block_jacobi_solve       :601-728    (lu_kernels.cu)
// The matrix stays where the assembly kernel built it: column-major, rows sharded across ranks.
// Per iteration the only HBM-heavy kernel is gemv_n_partial (one pass over the local rows of A);
// with several ranks the slices of w are all-gathered over NCCL (8 N bytes) and the
// orthogonalisation is replicated, so no reduction collective is needed.
//
// The Krylov loop is software-pipelined against the host: the Hessenberg column of iteration k is
// copied to pinned host memory asynchronously and iteration k+1 is already enqueued while the host
// applies the Givens rotations of iteration k (linalg.f90:1293-1319), so the device never waits for
// the convergence test; at most one speculative iteration is discarded at the end.
#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <unistd.h>
#include <vector>

#include "ctx.h"
#include "lu_internal.cuh"

namespace mlgpu {

// ---------------------------------------------------------------------------------------------------
// y_part[ks][row] = sum over the ks-th column range of A[row, col] * x[col]      (HBM-bound)
// CTA = 8 warps over the same 64 rows (lane owns 2 rows -> 128-bit loads), warps stride the columns.
// ---------------------------------------------------------------------------------------------------
constexpr int GEMV_THREADS = 256;
constexpr int GEMV_ROWS = 64;
constexpr int GEMV_UNROLL = 8;
constexpr double GEMV_L2_PIN_MB_DEFAULT = 0.;   // MACHLINE_GEMV_L2_PIN_MB overrides

__device__ __forceinline__ double2 ld_stream_d2(const double* p) {
    double2 v;
    asm volatile("ld.global.nc.L1::no_allocate.v2.f64 {%0, %1}, [%2];" : "=d"(v.x), "=d"(v.y) : "l"(p));
    return v;
}
// the same load with an L2 eviction policy (createpolicy): the matvec keeps a fixed slice of A resident in L2 across the
// iterations of a Krylov solve (evict_last) and streams the rest past it (evict_first)
__device__ __forceinline__ double2 ld_stream_d2_hint(const double* p, unsigned long long pol) {
    double2 v;
    asm volatile("ld.global.nc.L1::no_allocate.L2::cache_hint.v2.f64 {%0, %1}, [%2], %3;" : "=d"(v.x), "=d"(v.y) : "l"(p), "l"(pol));
    return v;
}

// Multi-GPU exchange of the matvec result over peer memory (no collective call): every rank maps every other rank's
// window (CUDA IPC).  p2p_push_kernel, one CTA per destination rank, stores this rank's rows of w straight into that
// rank's window at their global row index (NVLink / NVSwitch stores) and then raises this rank's flag there (release,
// system scope); p2p_wait_copy_kernel on the destination spins on the `world` flags of its own window.
// A first version issued the peer stores from the matvec kernel itself (the CTA finishing a 64-row block wrote its rows
// to every window, the last one raised the flags).  Measured on 2 x B200: exchange 19 -> 7 us per iteration but the
// matvec kernel +13 us (every finishing CTA needs a system-scope fence before the block counter and they all finish
// in the last microseconds of the kernel), net zero; one push CTA per peer needs `world` fences in total.
struct P2PArgs {
    double* w[Ctx::P2P_MAX];        // parity buffer of this matvec in every rank's window
    unsigned* flags[Ctx::P2P_MAX];  // flags[p][rank] <- seq
    unsigned seq;
    int rank, n_rows;
    const int* g_of_local;          // global row of each local row of this rank
};

__global__ void __launch_bounds__(512) p2p_push_kernel(const double* __restrict__ y_local, const P2PArgs pp) {
    const int p = blockIdx.x;
    double* dst = pp.w[p];
    for (int i = threadIdx.x; i < pp.n_rows; i += 512) dst[pp.g_of_local[i]] = y_local[i];
    __threadfence_system();
    __syncthreads();
    if (threadIdx.x == 0) asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(pp.flags[p] + pp.rank), "r"(pp.seq) : "memory");
}

// The CTA that finishes a 64-row block last (ticket counter per row block) adds the split partials in split order and
// writes y = alpha * sum: the result does not depend on which CTA that is, and no second kernel is needed.
//
// L2 residency (pin_cols > 0): the same A is streamed once per Krylov iteration and is larger than L2, so a plain stream
// gets nothing from the 126 MB cache.  A fixed slice of the columns (pin_cols, see below) is loaded with an evict_last policy
// and the others with evict_first: after the first iteration that slice (sized by the host to fit beside the Krylov basis)
// is served from L2 in every later matvec of the solve.  The kept columns are spread over the trips of the column loop,
// not put first: a block of hits followed by a block of misses runs at L2 speed and then at DRAM speed (measured: 40 % of
// the ideal gain), interleaved the two overlap.  The order of the additions does not change.
__global__ void __launch_bounds__(GEMV_THREADS) gemv_n_partial_kernel(const double* __restrict__ A, int ld, int n_rows_pad,
                                                                       int n_cols, int cols_per_split,
                                                                       const double* __restrict__ x,
                                                                       double* __restrict__ y_part, unsigned* __restrict__ tickets,
                                                                       int n_rows, const double* __restrict__ alpha_dev,
                                                                       double alpha, double* __restrict__ y, int pin_cols) {
    __shared__ double2 s_acc[GEMV_THREADS / 32][32];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int row = blockIdx.x * GEMV_ROWS + 2 * lane;
    const int c0 = blockIdx.y * cols_per_split;
    const int c1 = min(c0 + cols_per_split, n_cols);
    constexpr int NW = GEMV_THREADS / 32;
    double2 acc = make_double2(0., 0.);
    const double* Ap = A + row;
    int c = c0 + warp;
    if (pin_cols > 0) {
        // pin_cols = 256 * base + f: in every trip of 64 columns the first `base` of the 8 loads of a thread are kept, and one
        // more in f of every 256 trips (evenly spread), so that L2 hits and DRAM reads are in flight together all the time
        unsigned long long pol_keep, pol_stream;
        asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(pol_keep));
        asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(pol_stream));
        const int base = pin_cols >> 8, f256 = pin_cols & 255;
        int trip_no = 0;
        for (; c + (GEMV_UNROLL - 1) * NW < c1; c += GEMV_UNROLL * NW) {
            double2 v[GEMV_UNROLL];
            double xv[GEMV_UNROLL];
            const int npin = base + ((((trip_no + 1) * f256) >> 8) != ((trip_no * f256) >> 8) ? 1 : 0);
            ++trip_no;
#pragma unroll
            for (int u = 0; u < GEMV_UNROLL; ++u)
                v[u] = ld_stream_d2_hint(Ap + (size_t)(c + u * NW) * ld, u < npin ? pol_keep : pol_stream);
#pragma unroll
            for (int u = 0; u < GEMV_UNROLL; ++u) xv[u] = __ldg(x + c + u * NW);
#pragma unroll
            for (int u = 0; u < GEMV_UNROLL; ++u) {
                acc.x = fma(v[u].x, xv[u], acc.x);
                acc.y = fma(v[u].y, xv[u], acc.y);
            }
        }
    }
    // main loop: GEMV_UNROLL independent 16-byte loads in flight per thread
    for (; c + (GEMV_UNROLL - 1) * NW < c1; c += GEMV_UNROLL * NW) {
        double2 v[GEMV_UNROLL];
        double xv[GEMV_UNROLL];
#pragma unroll
        for (int u = 0; u < GEMV_UNROLL; ++u) v[u] = ld_stream_d2(Ap + (size_t)(c + u * NW) * ld);
#pragma unroll
        for (int u = 0; u < GEMV_UNROLL; ++u) xv[u] = __ldg(x + c + u * NW);
#pragma unroll
        for (int u = 0; u < GEMV_UNROLL; ++u) {
            acc.x = fma(v[u].x, xv[u], acc.x);
            acc.y = fma(v[u].y, xv[u], acc.y);
        }
    }
    for (; c < c1; c += NW) {
        double2 v = ld_stream_d2(Ap + (size_t)c * ld);
        double xv = __ldg(x + c);
        acc.x = fma(v.x, xv, acc.x);
        acc.y = fma(v.y, xv, acc.y);
    }
    s_acc[warp][lane] = acc;
    __syncthreads();
    if (warp == 0) {
        double2 s = s_acc[0][lane];
#pragma unroll
        for (int w = 1; w < NW; ++w) {
            s.x += s_acc[w][lane].x;
            s.y += s_acc[w][lane].y;
        }
        double2* out = reinterpret_cast<double2*>(y_part + (size_t)blockIdx.y * n_rows_pad + row);
        __stcg(out, s);
        if (tickets) {
            __threadfence();
            unsigned t = 0;
            if (lane == 0) t = atomicAdd(tickets + blockIdx.x, 1u);
            t = __shfl_sync(0xffffffffu, t, 0);
            if (t == gridDim.y - 1) {
                __threadfence();
                double2 sum = make_double2(0., 0.);
                for (unsigned k = 0; k < gridDim.y; ++k) {
                    const double2 p = __ldcg(reinterpret_cast<const double2*>(y_part + (size_t)k * n_rows_pad + row));
                    sum.x += p.x;
                    sum.y += p.y;
                }
                const double a = alpha_dev ? (*alpha_dev) * alpha : alpha;
                if (row < n_rows) y[row] = a * sum.x;
                if (row + 1 < n_rows) y[row + 1] = a * sum.y;
                if (lane == 0) tickets[blockIdx.x] = 0;   // ready for the next launch
            }
        }
    }
}

// y[row] = alpha * sum_ks y_part[ks][row]   -- fixed order, deterministic
__global__ void gemv_n_finish_kernel(const double* __restrict__ y_part, int n_rows_pad, int n_rows, int n_split,
                                     const double* __restrict__ alpha_dev, double alpha, double* __restrict__ y) {
    int row = blockIdx.x * blockDim.x + threadIdx.x;
    if (row >= n_rows) return;
    double s = 0.;
    for (int k = 0; k < n_split; ++k) s += y_part[(size_t)k * n_rows_pad + row];
    double a = alpha_dev ? (*alpha_dev) * alpha : alpha;
    y[row] = a * s;
}

// h[j] = Q[:, j] . w   for j = 0..ncol-1; one CTA per column, fixed-order block reduction
__global__ void __launch_bounds__(256) gemv_t_kernel(const double* __restrict__ Q, int ldq, int n, const double* __restrict__ w,
                                                      double* __restrict__ h) {
    __shared__ double s_part[8];
    const int j = blockIdx.x;
    const double* q = Q + (size_t)j * ldq;
    double a0 = 0., a1 = 0., a2 = 0., a3 = 0.;
    int i = threadIdx.x;
    for (; i + 768 < n; i += 1024) {
        a0 = fma(q[i], w[i], a0);
        a1 = fma(q[i + 256], w[i + 256], a1);
        a2 = fma(q[i + 512], w[i + 512], a2);
        a3 = fma(q[i + 768], w[i + 768], a3);
    }
    for (; i < n; i += 256) a0 = fma(q[i], w[i], a0);
    double acc = (a0 + a1) + (a2 + a3);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
    if ((threadIdx.x & 31) == 0) s_part[threadIdx.x >> 5] = acc;
    __syncthreads();
    if (threadIdx.x == 0) {
        double s = 0.;
#pragma unroll
        for (int k = 0; k < 8; ++k) s += s_part[k];
        h[j] = s;
    }
}

// w[i] -= sum_j Q[i, j] * h[j]; optionally hsum[j] = h[j] + hprev[j] (second Gram-Schmidt pass)
__global__ void __launch_bounds__(64) gemv_n_sub_kernel(const double* __restrict__ Q, int ldq, int n, int ncol,
                                                         const double* __restrict__ h, double* __restrict__ w,
                                                         const double* __restrict__ hprev, double* __restrict__ hsum) {
    const int i = blockIdx.x * 64 + threadIdx.x;
    if (hsum && blockIdx.x == 0)
        for (int j = threadIdx.x; j < ncol; j += 64) hsum[j] = hprev[j] + h[j];
    if (i >= n) return;
    double a0 = 0., a1 = 0., a2 = 0., a3 = 0.;
    int j = 0;
    for (; j + 3 < ncol; j += 4) {
        a0 = fma(Q[i + (size_t)j * ldq], __ldg(h + j), a0);
        a1 = fma(Q[i + (size_t)(j + 1) * ldq], __ldg(h + j + 1), a1);
        a2 = fma(Q[i + (size_t)(j + 2) * ldq], __ldg(h + j + 2), a2);
        a3 = fma(Q[i + (size_t)(j + 3) * ldq], __ldg(h + j + 3), a3);
    }
    for (; j < ncol; ++j) a0 = fma(Q[i + (size_t)j * ldq], __ldg(h + j), a0);
    w[i] -= (a0 + a1) + (a2 + a3);
}

// ---- device-parallel Gram-Schmidt pass (classical, applied twice) -------------------------------------------------
// K1: partial[rb][j] = sum over the rb-th block of ORTH_RB rows of Q[i, j] * w[i].  Grid (row blocks, column groups of
//     32): every CTA reads its slice of w once into registers and streams 32 columns of Q (L2 resident: N x k doubles).
// K2: h[j] = sum_rb partial[rb][j] (fixed order, every CTA recomputes it into shared memory), w -= Q h for 64 rows per
//     CTA with the 8 warps splitting the columns, and -- optionally -- the per-CTA partial of ||w||^2.
constexpr int ORTH_RB = 512;      // rows per K1 CTA
constexpr int ORTH_CG = 32;       // columns per K1 CTA (4 per warp)

__global__ void __launch_bounds__(256) orth_dot_kernel(const double* __restrict__ Q, int ldq, int n, int ncol,
                                                        const double* __restrict__ w, double* __restrict__ partial, int kpad) {
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int row0 = blockIdx.x * ORTH_RB;
    const int col0 = blockIdx.y * ORTH_CG + warp * 4;
    if (col0 >= ncol) return;
    double wv[ORTH_RB / 32];
#pragma unroll
    for (int j = 0; j < ORTH_RB / 32; ++j) {
        const int r = row0 + lane + 32 * j;
        wv[j] = (r < n) ? w[r] : 0.;
    }
    double acc[4] = {0., 0., 0., 0.};
#pragma unroll
    for (int c = 0; c < 4; ++c) {
        if (col0 + c < ncol) {
            const double* q = Q + (size_t)(col0 + c) * ldq + row0 + lane;
            double a0 = 0., a1 = 0.;
#pragma unroll
            for (int j = 0; j < ORTH_RB / 32; j += 2) {
                const int r = row0 + lane + 32 * j;
                const double q0 = (r < n) ? q[32 * j] : 0.;
                const double q1 = (r + 32 < n) ? q[32 * j + 32] : 0.;
                a0 = fma(q0, wv[j], a0);
                a1 = fma(q1, wv[j + 1], a1);
            }
            acc[c] = a0 + a1;
        }
    }
#pragma unroll
    for (int c = 0; c < 4; ++c) {
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) acc[c] += __shfl_xor_sync(0xffffffffu, acc[c], o);
    }
    if (lane < 4 && col0 + lane < ncol) {
        const double v = lane == 0 ? acc[0] : lane == 1 ? acc[1] : lane == 2 ? acc[2] : acc[3];
        partial[(size_t)blockIdx.x * kpad + col0 + lane] = v;
    }
}

__global__ void __launch_bounds__(256) orth_sub_kernel(const double* __restrict__ Q, int ldq, int n, int ncol,
                                                        const double* __restrict__ partial, int kpad, int n_rb,
                                                        double* __restrict__ w, const double* __restrict__ hprev,
                                                        double* __restrict__ hout, double* __restrict__ norm_partial) {
    extern __shared__ double s_h[];                 // [ncol] coefficients, then [8][64] warp partials, then 2 norm partials
    double* s_acc = s_h + ((ncol + 1) & ~1);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    for (int j = threadIdx.x; j < ncol; j += 256) {
        double s = 0.;
        for (int rb = 0; rb < n_rb; ++rb) s += partial[(size_t)rb * kpad + j];
        s_h[j] = s;
        if (blockIdx.x == 0 && hout) hout[j] = hprev ? hprev[j] + s : s;
    }
    __syncthreads();
    const int row = blockIdx.x * 64 + 2 * lane;     // ldq is a multiple of 64 and Q, w are padded: no row guard needed
    double2 acc = make_double2(0., 0.);
    int j = warp;
    for (; j + 24 < ncol; j += 32) {
        const double2 q0 = *reinterpret_cast<const double2*>(Q + (size_t)j * ldq + row);
        const double2 q1 = *reinterpret_cast<const double2*>(Q + (size_t)(j + 8) * ldq + row);
        const double2 q2 = *reinterpret_cast<const double2*>(Q + (size_t)(j + 16) * ldq + row);
        const double2 q3 = *reinterpret_cast<const double2*>(Q + (size_t)(j + 24) * ldq + row);
        const double h0 = s_h[j], h1 = s_h[j + 8], h2 = s_h[j + 16], h3 = s_h[j + 24];
        acc.x = fma(q0.x, h0, acc.x); acc.y = fma(q0.y, h0, acc.y);
        acc.x = fma(q1.x, h1, acc.x); acc.y = fma(q1.y, h1, acc.y);
        acc.x = fma(q2.x, h2, acc.x); acc.y = fma(q2.y, h2, acc.y);
        acc.x = fma(q3.x, h3, acc.x); acc.y = fma(q3.y, h3, acc.y);
    }
    for (; j < ncol; j += 8) {
        const double2 q0 = *reinterpret_cast<const double2*>(Q + (size_t)j * ldq + row);
        const double h0 = s_h[j];
        acc.x = fma(q0.x, h0, acc.x); acc.y = fma(q0.y, h0, acc.y);
    }
    s_acc[warp * 64 + 2 * lane] = acc.x;
    s_acc[warp * 64 + 2 * lane + 1] = acc.y;
    __syncthreads();
    if (threadIdx.x < 64) {
        double s = 0.;
#pragma unroll
        for (int k = 0; k < 8; ++k) s += s_acc[k * 64 + threadIdx.x];
        const int r = blockIdx.x * 64 + threadIdx.x;
        double v = 0.;
        if (r < n) {
            v = w[r] - s;
            w[r] = v;
        }
        if (norm_partial) {
            double sq = v * v;
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) sq += __shfl_xor_sync(0xffffffffu, sq, o);
            if (lane == 0) s_acc[8 * 64 + warp] = sq;   // warps 0 and 1; a slot of its own (other warps may still read s_acc)
        }
    }
    if (norm_partial) {
        __syncthreads();
        if (threadIdx.x == 0) norm_partial[blockIdx.x] = s_acc[8 * 64] + s_acc[8 * 64 + 1];
    }
}


// ---- fused Arnoldi tail: both Gram-Schmidt passes, the norm and the new basis vector in ONE cooperative launch ----------
// One CTA per SM (1024 threads), four grid barriers:
//   dot(w) | B | h1 = sum partials, w -= Q h1 | B | dot(w) | B | h2, w -= Q h2, ||w||^2 partials | B | q_next = w / ||w||
// Between kernels the per-iteration latency was launch gaps + one wave ramp per kernel (6 launches ~ 80 us at N = 10k
// for ~10 us of L2-resident traffic); inside one launch a barrier costs ~2 us.  Data other CTAs wrote earlier in the
// launch (w, partials) is read with ld.global.cg: L1 is not coherent across SMs.
constexpr int TAIL_THREADS = 1024;
constexpr int TAIL_WARPS = TAIL_THREADS / 32;
constexpr int TAIL_RB = 256;       // rows per dot item (8 per lane)

struct TailArgs {
    const double* Q;
    int ldq, n, k;          // k = number of basis vectors to orthogonalise against
    double* w;              // in: A q (length ldq, pad rows zero); out: orthogonalised, not normalised
    double* partial;        // [n_rb][kpad]
    int kpad, n_rb;
    double* h1;             // first-pass coefficients
    double* hfin;           // hfin[0..k-1] = h1 + h2, hfin[k] = ||w||
    double* npart;          // [gridDim.x]
    double* qnext;          // Q(:, k)
    unsigned* bar;          // monotonic barrier counter (zeroed at the start of a solve)
    unsigned bar_base;      // its value when this launch begins
    int rows_per_cta;       // multiple of 32
};

__device__ __forceinline__ void tail_grid_sync(unsigned* bar, unsigned target) {
    __syncthreads();
    if (threadIdx.x == 0) {
        __threadfence();
        atomicAdd(bar, 1u);
        unsigned v;
        do {
            asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(bar) : "memory");
        } while ((int)(v - target) < 0);
    }
    __syncthreads();
}

__device__ __forceinline__ void tail_dot_phase(const TailArgs& a) {
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int ncg = (a.k + 3) >> 2;
    const int items = a.n_rb * ncg;
    for (int item = blockIdx.x * TAIL_WARPS + warp; item < items; item += gridDim.x * TAIL_WARPS) {
        const int rb = item % a.n_rb, col0 = (item / a.n_rb) * 4;
        const int row0 = rb * TAIL_RB + lane;
        double wv[TAIL_RB / 32];
#pragma unroll
        for (int j = 0; j < TAIL_RB / 32; ++j) {
            const int r = row0 + 32 * j;
            wv[j] = (r < a.n) ? __ldcg(a.w + r) : 0.;
        }
        double acc[4];
#pragma unroll
        for (int c = 0; c < 4; ++c) {
            const int col = min(col0 + c, a.k - 1);
            const double* q = a.Q + (size_t)col * a.ldq + row0;
            double a0 = 0., a1 = 0.;
#pragma unroll
            for (int j = 0; j < TAIL_RB / 32; j += 2) {
                const int r = row0 + 32 * j;
                const double q0 = (r < a.n) ? q[32 * j] : 0.;
                const double q1 = (r + 32 < a.n) ? q[32 * j + 32] : 0.;
                a0 = fma(q0, wv[j], a0);
                a1 = fma(q1, wv[j + 1], a1);
            }
            acc[c] = a0 + a1;
        }
#pragma unroll
        for (int c = 0; c < 4; ++c) {
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) acc[c] += __shfl_xor_sync(0xffffffffu, acc[c], o);
        }
        if (lane < 4 && col0 + lane < a.k) {
            const double v = lane == 0 ? acc[0] : lane == 1 ? acc[1] : lane == 2 ? acc[2] : acc[3];
            __stcg(a.partial + (size_t)rb * a.kpad + col0 + lane, v);
        }
    }
}

// s_h[j] = sum_rb partial[rb][j]; w -= Q s_h over this CTA's rows; returns the CTA's sum of squares of the new w (thread 0)
__device__ __forceinline__ double tail_sub_phase(const TailArgs& a, double* s_h, double* s_acc, bool second) {
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    for (int j = threadIdx.x; j < a.k; j += TAIL_THREADS) {
        double s = 0.;
        for (int rb = 0; rb < a.n_rb; ++rb) s += __ldcg(a.partial + (size_t)rb * a.kpad + j);
        s_h[j] = s;
        if (blockIdx.x == 0) {
            if (second) a.hfin[j] = a.h1[j] + s;
            else a.h1[j] = s;
        }
    }
    __syncthreads();
    const int n_rg = a.rows_per_cta >> 5;                 // 32-row groups of this CTA
    const int n_cs = TAIL_WARPS / n_rg > 0 ? TAIL_WARPS / n_rg : 1;   // column splits
    const int cta_row0 = blockIdx.x * a.rows_per_cta;
    double sq_total = 0.;
    for (int rg0 = 0; rg0 < n_rg; rg0 += TAIL_WARPS) {     // one trip unless a CTA owns more than 1024 rows
        const int rg = rg0 + (n_cs > 1 ? warp % n_rg : warp), cs = (n_cs > 1) ? warp / n_rg : 0;
        const bool busy = rg < n_rg && cs < n_cs;
        const int row = cta_row0 + rg * 32 + lane;
        double a0 = 0., a1 = 0., a2 = 0., a3 = 0.;
        if (busy && row < a.n) {
            const double* q = a.Q + row;
            int j = cs;
            for (; j + 3 * n_cs < a.k; j += 4 * n_cs) {
                const double q0 = q[(size_t)j * a.ldq], q1 = q[(size_t)(j + n_cs) * a.ldq];
                const double q2 = q[(size_t)(j + 2 * n_cs) * a.ldq], q3 = q[(size_t)(j + 3 * n_cs) * a.ldq];
                a0 = fma(q0, s_h[j], a0);
                a1 = fma(q1, s_h[j + n_cs], a1);
                a2 = fma(q2, s_h[j + 2 * n_cs], a2);
                a3 = fma(q3, s_h[j + 3 * n_cs], a3);
            }
            for (; j < a.k; j += n_cs) a0 = fma(q[(size_t)j * a.ldq], s_h[j], a0);
        }
        if (busy) s_acc[cs * a.rows_per_cta + (rg - rg0) * 32 + lane] = (a0 + a1) + (a2 + a3);
        __syncthreads();
        const int rows_here = min(TAIL_WARPS, n_rg - rg0) * 32;
        double sq = 0.;
        if ((int)threadIdx.x < rows_here) {
            double s = 0.;
            for (int c = 0; c < n_cs; ++c) s += s_acc[c * a.rows_per_cta + threadIdx.x];
            const int r = cta_row0 + rg0 * 32 + threadIdx.x;
            if (r < a.n) {
                const double v = __ldcg(a.w + r) - s;
                __stcg(a.w + r, v);
                sq = v * v;
            }
        }
        if (second) {
            // fixed-order block sum of sq (a fixed shuffle tree per warp, then warp order)
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) sq += __shfl_xor_sync(0xffffffffu, sq, o);
            __syncthreads();
            if (lane == 0) s_acc[warp] = sq;
            __syncthreads();
            if (threadIdx.x == 0)
                for (int k2 = 0; k2 < TAIL_WARPS; ++k2) sq_total += s_acc[k2];
        }
        __syncthreads();
    }
    return sq_total;   // meaningful in thread 0
}

__global__ void __launch_bounds__(TAIL_THREADS, 1) arnoldi_tail_kernel(const TailArgs a) {
    extern __shared__ double s_mem[];
    double* s_h = s_mem;                                   // [k rounded up to even]
    double* s_acc = s_mem + ((a.k + 1) & ~1);              // [1024]: column splits x rows per trip <= 32 x 32
    __shared__ double s_norm;
    const unsigned G = gridDim.x;
    tail_dot_phase(a);
    tail_grid_sync(a.bar, a.bar_base + G);
    tail_sub_phase(a, s_h, s_acc, false);
    tail_grid_sync(a.bar, a.bar_base + 2 * G);
    tail_dot_phase(a);
    tail_grid_sync(a.bar, a.bar_base + 3 * G);
    const double sq = tail_sub_phase(a, s_h, s_acc, true);
    if (threadIdx.x == 0) __stcg(a.npart + blockIdx.x, sq);
    tail_grid_sync(a.bar, a.bar_base + 4 * G);
    // ||w|| from the per-CTA partials in CTA order (every CTA computes the same value), q_next = w / ||w||
    if (threadIdx.x < 32) {
        double t = 0.;
        for (unsigned c = threadIdx.x; c < G; c += 32) t += __ldcg(a.npart + c);
        // fixed tree over the 32 lane sums
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) t += __shfl_xor_sync(0xffffffffu, t, o);
        if (threadIdx.x == 0) {
            s_norm = sqrt(t);
            if (blockIdx.x == 0) a.hfin[a.k] = s_norm;
        }
    }
    __syncthreads();
    const double nrm = s_norm;
    const int cta_row0 = blockIdx.x * a.rows_per_cta;
    for (int t = threadIdx.x; t < a.rows_per_cta; t += TAIL_THREADS) {
        const int r = cta_row0 + t;
        if (r < a.n) a.qnext[r] = __ldcg(a.w + r) / nrm;
    }
}

// ---- fused Arnoldi tail, version 3 (default): column-owned dot products ------------------------------------------------
// The first version's dot phase left an [N / 256] x k array of partial sums that every CTA of the subtraction phase then
// added up again, column by column, with dependent L2 loads (29 per column at N = 7376; 7 us of a 19 us pass at k = 535).
// Here a CTA owns whole COLUMNS of Q in the dot phases (groups of CG columns, all 1024 threads over the rows, one fixed
// reduction tree: the finished coefficient is stored once) and ROWS in the subtraction phases (the coefficients are read
// straight from that array, eight independent loads in flight per thread).  Same arithmetic sequence as classical
// Gram-Schmidt with one re-orthogonalisation pass; the sums are deterministic and independent of the grid size.
// CTA 0 finally writes the Hessenberg column into the host's pinned slot (mapped memory): no copy-engine operation
// between the kernels of consecutive iterations.
struct Tail3Args {
    const double* Q;
    int ldq, n, k;
    double* w;
    double* h1;             // first-pass coefficients
    double* h2;             // second-pass corrections
    double* hfin;           // hfin[0..k-1] = h1 + h2, hfin[k] = ||w||
    double* hhost;          // the same k + 1 numbers in mapped pinned host memory (nullptr: none)
    double* npart;          // [gridDim.x]
    double* qnext;          // Q(:, k)
    unsigned* bar;
    unsigned bar_base;
    int rows_per_cta;       // multiple of 32
    long long* dbg;         // MACHLINE_TAIL_DEBUG: [16] accumulated clock64 deltas of CTA 0 per phase boundary, [16] = launches
};

template <int CG>
__device__ __forceinline__ void tail3_dot_phase(const Tail3Args& a, double* s_red, bool second) {
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int ngroups = (a.k + CG - 1) / CG;
    for (int g = blockIdx.x; g < ngroups; g += gridDim.x) {
        const int col0 = g * CG;
        const double* q[CG];
#pragma unroll
        for (int c = 0; c < CG; ++c) q[c] = a.Q + (size_t)min(col0 + c, a.k - 1) * a.ldq;
        double acc[CG];
#pragma unroll
        for (int c = 0; c < CG; ++c) acc[c] = 0.;
        int r = threadIdx.x;
        for (; r + 3 * TAIL_THREADS < a.n; r += 4 * TAIL_THREADS) {
            double wv[4], qv[CG][4];
#pragma unroll
            for (int u = 0; u < 4; ++u) wv[u] = __ldcg(a.w + r + u * TAIL_THREADS);
#pragma unroll
            for (int c = 0; c < CG; ++c)
#pragma unroll
                for (int u = 0; u < 4; ++u) qv[c][u] = __ldcg(q[c] + r + u * TAIL_THREADS);
#pragma unroll
            for (int u = 0; u < 4; ++u)
#pragma unroll
                for (int c = 0; c < CG; ++c) acc[c] = fma(qv[c][u], wv[u], acc[c]);
        }
        {   // the last (partial) trip, guarded: still all loads in flight together
            double wv[4], qv[CG][4];
#pragma unroll
            for (int u = 0; u < 4; ++u) wv[u] = (r + u * TAIL_THREADS < a.n) ? __ldcg(a.w + r + u * TAIL_THREADS) : 0.;
#pragma unroll
            for (int c = 0; c < CG; ++c)
#pragma unroll
                for (int u = 0; u < 4; ++u) qv[c][u] = (r + u * TAIL_THREADS < a.n) ? __ldcg(q[c] + r + u * TAIL_THREADS) : 0.;
#pragma unroll
            for (int u = 0; u < 4; ++u)
#pragma unroll
                for (int c = 0; c < CG; ++c) acc[c] = fma(qv[c][u], wv[u], acc[c]);
        }
#pragma unroll
        for (int c = 0; c < CG; ++c) {
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) acc[c] += __shfl_xor_sync(0xffffffffu, acc[c], o);
            if (lane == 0) s_red[c * TAIL_WARPS + warp] = acc[c];
        }
        __syncthreads();
        if (warp < CG && col0 + warp < a.k) {
            double t = s_red[warp * TAIL_WARPS + lane];
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) t += __shfl_xor_sync(0xffffffffu, t, o);
            if (lane == 0) {
                const int j = col0 + warp;
                if (second) {
                    __stcg(a.h2 + j, t);
                    __stcg(a.hfin + j, __ldcg(a.h1 + j) + t);
                } else {
                    __stcg(a.h1 + j, t);
                }
            }
        }
        __syncthreads();
    }
}

// w -= Q h over this CTA's rows (h = the k coefficients of this pass); returns the CTA's sum of squares of the new w (thread 0)
__device__ __forceinline__ double tail3_sub_phase(const Tail3Args& a, const double* h, double* s_h, double* s_acc, bool second) {
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    for (int j = threadIdx.x; j < a.k; j += TAIL_THREADS) s_h[j] = __ldcg(h + j);
    __syncthreads();
    const int n_rg = a.rows_per_cta >> 5;                 // 32-row groups of this CTA
    const int n_cs = TAIL_WARPS / n_rg > 0 ? TAIL_WARPS / n_rg : 1;   // column splits
    const int cta_row0 = blockIdx.x * a.rows_per_cta;
    double sq_total = 0.;
    for (int rg0 = 0; rg0 < n_rg; rg0 += TAIL_WARPS) {     // one trip unless a CTA owns more than 1024 rows
        const int rg = rg0 + (n_cs > 1 ? warp % n_rg : warp), cs = (n_cs > 1) ? warp / n_rg : 0;
        const bool busy = rg < n_rg && cs < n_cs;
        const int row = cta_row0 + rg * 32 + lane;
        double acc[8];
#pragma unroll
        for (int u = 0; u < 8; ++u) acc[u] = 0.;
        if (busy && row < a.n) {
            const double* q = a.Q + row;
            int j = cs;
            for (; j + 7 * n_cs < a.k; j += 8 * n_cs) {
                double qv[8];
#pragma unroll
                for (int u = 0; u < 8; ++u) qv[u] = __ldcg(q + (size_t)(j + u * n_cs) * a.ldq);
#pragma unroll
                for (int u = 0; u < 8; ++u) acc[u] = fma(qv[u], s_h[j + u * n_cs], acc[u]);
            }
            {
                double qv[8];
#pragma unroll
                for (int u = 0; u < 8; ++u) qv[u] = (j + u * n_cs < a.k) ? __ldcg(q + (size_t)(j + u * n_cs) * a.ldq) : 0.;
#pragma unroll
                for (int u = 0; u < 8; ++u) acc[u] = fma(qv[u], (j + u * n_cs < a.k) ? s_h[j + u * n_cs] : 0., acc[u]);
            }
        }
        if (busy) s_acc[cs * a.rows_per_cta + (rg - rg0) * 32 + lane] =
            ((acc[0] + acc[1]) + (acc[2] + acc[3])) + ((acc[4] + acc[5]) + (acc[6] + acc[7]));
        __syncthreads();
        const int rows_here = min(TAIL_WARPS, n_rg - rg0) * 32;
        double sq = 0.;
        if ((int)threadIdx.x < rows_here) {
            double s = 0.;
            for (int c = 0; c < n_cs; ++c) s += s_acc[c * a.rows_per_cta + threadIdx.x];
            const int r = cta_row0 + rg0 * 32 + threadIdx.x;
            if (r < a.n) {
                const double v = __ldcg(a.w + r) - s;
                __stcg(a.w + r, v);
                sq = v * v;
            }
        }
        if (second) {
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) sq += __shfl_xor_sync(0xffffffffu, sq, o);
            __syncthreads();
            if (lane == 0) s_acc[warp] = sq;
            __syncthreads();
            if (threadIdx.x == 0)
                for (int k2 = 0; k2 < TAIL_WARPS; ++k2) sq_total += s_acc[k2];
        }
        __syncthreads();
    }
    return sq_total;   // meaningful in thread 0
}

__global__ void __launch_bounds__(TAIL_THREADS, 1) arnoldi_tail3_kernel(const Tail3Args a) {
    extern __shared__ double s_mem[];
    double* s_h = s_mem;                                   // [k rounded up to even]
    double* s_acc = s_mem + ((a.k + 1) & ~1);              // [1024]
    __shared__ double s_red[4 * TAIL_WARPS];
    __shared__ double s_norm;
    const unsigned G = gridDim.x;
    const bool wide = a.k > 2 * (int)G;                    // four columns per group once two no longer give every CTA one trip
    long long t_prev = 0;
    int stamp_n = 0;
    const bool stamping = a.dbg != nullptr && blockIdx.x == 0 && threadIdx.x == 0;
#define TAIL3_STAMP()                                                    \
    do {                                                                 \
        if (stamping) {                                                  \
            const long long t_now = clock64();                           \
            if (stamp_n > 0) a.dbg[stamp_n - 1] += t_now - t_prev;       \
            t_prev = t_now;                                              \
            ++stamp_n;                                                   \
        }                                                                \
    } while (0)
    TAIL3_STAMP();
    if (wide) tail3_dot_phase<4>(a, s_red, false); else tail3_dot_phase<2>(a, s_red, false);
    TAIL3_STAMP();
    tail_grid_sync(a.bar, a.bar_base + G);
    TAIL3_STAMP();
    tail3_sub_phase(a, a.h1, s_h, s_acc, false);
    TAIL3_STAMP();
    tail_grid_sync(a.bar, a.bar_base + 2 * G);
    TAIL3_STAMP();
    if (wide) tail3_dot_phase<4>(a, s_red, true); else tail3_dot_phase<2>(a, s_red, true);
    TAIL3_STAMP();
    tail_grid_sync(a.bar, a.bar_base + 3 * G);
    TAIL3_STAMP();
    const double sq = tail3_sub_phase(a, a.h2, s_h, s_acc, true);
    if (threadIdx.x == 0) __stcg(a.npart + blockIdx.x, sq);
    TAIL3_STAMP();
    tail_grid_sync(a.bar, a.bar_base + 4 * G);
    TAIL3_STAMP();
    if (threadIdx.x < 32) {
        double t = 0.;
        for (unsigned c = threadIdx.x; c < G; c += 32) t += __ldcg(a.npart + c);
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) t += __shfl_xor_sync(0xffffffffu, t, o);
        if (threadIdx.x == 0) {
            s_norm = sqrt(t);
            if (blockIdx.x == 0) a.hfin[a.k] = s_norm;
        }
    }
    __syncthreads();
    const double nrm = s_norm;
    const int cta_row0 = blockIdx.x * a.rows_per_cta;
    for (int t = threadIdx.x; t < a.rows_per_cta; t += TAIL_THREADS) {
        const int r = cta_row0 + t;
        if (r < a.n) a.qnext[r] = __ldcg(a.w + r) / nrm;
    }
    if (blockIdx.x == gridDim.x - 1 && a.hhost) {          // the CTA with the fewest rows (often none) reports to the host
        for (int j = threadIdx.x; j <= a.k; j += TAIL_THREADS) a.hhost[j] = j < a.k ? __ldcg(a.hfin + j) : nrm;
    }
    __syncthreads();
    TAIL3_STAMP();
    if (stamping) a.dbg[16] += 1;
#undef TAIL3_STAMP
}

// norm = sqrt(sum of the per-CTA partials) (fixed order); q = w / norm; CTA 0 also stores the norm
__global__ void __launch_bounds__(1024) norm_scale2_kernel(const double* __restrict__ w, int n, const double* __restrict__ norm_partial,
                                                            int n_part, double* __restrict__ norm_out, double* __restrict__ q) {
    __shared__ double s_part[32];
    __shared__ double s_norm;
    double acc = 0.;
    for (int i = threadIdx.x; i < n_part; i += 1024) acc += norm_partial[i];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
    if ((threadIdx.x & 31) == 0) s_part[threadIdx.x >> 5] = acc;
    __syncthreads();
    if (threadIdx.x == 0) {
        double s = 0.;
        for (int k = 0; k < 32; ++k) s += s_part[k];
        s_norm = sqrt(s);
        if (blockIdx.x == 0) *norm_out = s_norm;
    }
    __syncthreads();
    const int i = blockIdx.x * 1024 + threadIdx.x;
    if (i < n) q[i] = w[i] / s_norm;
}

// x[i] (+)= sum_j Q[i, j] * y[j]
__global__ void __launch_bounds__(64) gemv_n_small_kernel(const double* __restrict__ Q, int ldq, int n, int ncol,
                                                           const double* __restrict__ y, double* __restrict__ x, int accumulate) {
    int i = blockIdx.x * 64 + threadIdx.x;
    if (i >= n) return;
    double acc = 0.;
    for (int j = 0; j < ncol; ++j) acc = fma(Q[i + (size_t)j * ldq], __ldg(y + j), acc);
    x[i] = accumulate ? x[i] + acc : acc;
}

// out[0] = sqrt(sum w^2) (single CTA, fixed order); optionally q = w / norm
__global__ void __launch_bounds__(1024) norm_scale_kernel(const double* __restrict__ w, int n, double* __restrict__ norm_out,
                                                           double* __restrict__ q) {
    __shared__ double s_part[32];
    __shared__ double s_norm;
    double acc = 0.;
    for (int i = threadIdx.x; i < n; i += 1024) acc = fma(w[i], w[i], acc);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
    if ((threadIdx.x & 31) == 0) s_part[threadIdx.x >> 5] = acc;
    __syncthreads();
    if (threadIdx.x == 0) {
        double s = 0.;
        for (int k = 0; k < 32; ++k) s += s_part[k];
        s_norm = sqrt(s);
        *norm_out = s_norm;
    }
    __syncthreads();
    if (q) {
        const double nrm = s_norm;
        for (int i = threadIdx.x; i < n; i += 1024) q[i] = w[i] / nrm;
    }
}

// Modified Gram-Schmidt in the reference's order (linalg.f90:1222-1226): one CTA walks the basis.
__global__ void __launch_bounds__(1024) mgs_kernel(const double* __restrict__ Q, int ldq, int n, int ncol,
                                                    double* __restrict__ w, double* __restrict__ h) {
    __shared__ double s_part[32];
    __shared__ double s_h;
    for (int j = 0; j < ncol; ++j) {
        const double* q = Q + (size_t)j * ldq;
        double acc = 0.;
        for (int i = threadIdx.x; i < n; i += 1024) acc = fma(w[i], q[i], acc);
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
        if ((threadIdx.x & 31) == 0) s_part[threadIdx.x >> 5] = acc;
        __syncthreads();
        if (threadIdx.x == 0) {
            double s = 0.;
            for (int k = 0; k < 32; ++k) s += s_part[k];
            s_h = s;
            h[j] = s;
        }
        __syncthreads();
        const double hj = s_h;
        for (int i = threadIdx.x; i < n; i += 1024) w[i] = w[i] - hj * q[i];
        __syncthreads();
    }
}

__global__ void scale_copy_kernel(const double* __restrict__ src, double alpha, const double* alpha_dev, double* __restrict__ dst,
                                  int n) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) dst[i] = (alpha_dev ? *alpha_dev * alpha : alpha) * src[i];
}

__global__ void axpby_kernel(double a, const double* __restrict__ x, double b, const double* __restrict__ y,
                             double* __restrict__ out, int n) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) out[i] = a * x[i] + b * y[i];
}

__global__ void recip_kernel(const double* src, double* dst) { *dst = 1. / *src; }

// ---------------------------------------------------------------------------------------------------
// Host-side drivers
// ---------------------------------------------------------------------------------------------------
// Consumer of the fused exchange: wait until every rank has raised its flag for matvec `seq` in this rank's window, then
// copy the assembled vector out of the window (L2 loads: the window is written by peers, never cached in L1).
__global__ void __launch_bounds__(256) p2p_wait_copy_kernel(const double* __restrict__ win_w, const unsigned* __restrict__ flags, unsigned seq,
                                                             int world, int n, double* __restrict__ y_full) {
    if (threadIdx.x < world) {
        unsigned v;
        do {
            asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(flags + threadIdx.x) : "memory");
        } while ((int)(v - seq) < 0);
    }
    __syncthreads();
    const int i = blockIdx.x * 256 + threadIdx.x;
    if (i < n) y_full[i] = __ldcg(win_w + i);
}

// ---- peer-memory windows (CUDA IPC; one process per GPU on one node) ----------------------------------------------------
// Layout of a window, in doubles (win_n = N rounded up to 64, KR = Ctx::P2P_KR):
//   [0, 2 win_n)            vector, two parities       (Sys::exchange: p2p_push_kernel / p2p_wait_copy_kernel)
//   8 doubles               flags of that exchange     ([2][P2P_MAX] u32)
//   8 doubles               flags of the sharded Arnoldi tail ([2][P2P_MAX] u32, gmres_sharded.cuh)
//   [.., + 2 win_n)         vector of the sharded Arnoldi tail, two parities
//   [.., + 2 P2P_MAX KR)    reduction slots of the sharded Arnoldi tail, two parities
static void p2p_teardown(Ctx* c) {
    for (int r = 0; r < Ctx::P2P_MAX; ++r) {
        if (c->peer_base[r] && r != c->rank && c->peer_ipc[r]) cudaIpcCloseMemHandle(c->peer_base[r]);
        c->peer_base[r] = nullptr;
        c->peer_ipc[r] = false;
    }
    if (c->win) cudaFree(c->win);
    c->win = nullptr;
    c->win_n = 0;
    c->p2p_ok = false;
    cudaGetLastError();
}

void p2p_release(Ctx* c) { p2p_teardown(c); }

static size_t p2p_window_doubles(size_t win_n) { return 4 * win_n + 16 + (size_t)2 * Ctx::P2P_MAX * Ctx::P2P_KR; }
static size_t p2p_window_bytes(size_t win_n) { return p2p_window_doubles(win_n) * sizeof(double); }

// Collective: every rank calls it with the same n.  On any failure on any rank all ranks fall back to ncclAllGather
// (p2p_ok = false).  local_only: one rank, the window is its own memory (the multi-rank kernels on the single-GPU test box).
static ml_status p2p_setup(Ctx* c, int n, bool local_only) {
    static const bool disabled = std::getenv("MACHLINE_NO_P2P") != nullptr;
    const size_t need = ((size_t)n + 63) / 64 * 64;
    if (local_only && c->world == 1) {
        if (c->p2p_ok && c->win_n >= need) return ML_OK;
        ML_CUDA(c, cudaStreamSynchronize(c->stream));
        p2p_teardown(c);
        ML_CUDA(c, cudaMalloc((void**)&c->win, p2p_window_bytes(need)));
        ML_CUDA(c, cudaMemset(c->win, 0, p2p_window_bytes(need)));
        c->peer_base[0] = c->win;
        c->win_n = need;
        c->p2p_seq = 0;
        c->xseq = 0;
        c->p2p_ok = !disabled;
        return ML_OK;
    }
    if (c->world < 2 || c->world > Ctx::P2P_MAX || disabled) { c->p2p_ok = false; return ML_OK; }
#ifdef ML_HAVE_NCCL
    if (c->p2p_ok && c->win_n >= need) return ML_OK;
    ML_CUDA(c, cudaStreamSynchronize(c->stream));
    p2p_teardown(c);
    int ok = 1;
    // What a peer needs to reach this rank's window: the IPC handle (another process), or, for a rank driven by a thread of
    // the SAME process (ml_ctx_create_multi), the pointer itself and the device to enable peer access to -- a process cannot
    // open its own IPC handles.
    struct WinInfo {
        cudaIpcMemHandle_t handle;
        long long pid;
        void* ptr;
        int device;
        int pad;
    } mine;
    std::memset(&mine, 0, sizeof mine);
    if (cudaMalloc((void**)&c->win, p2p_window_bytes(need)) != cudaSuccess || cudaMemset(c->win, 0, p2p_window_bytes(need)) != cudaSuccess ||
        cudaIpcGetMemHandle(&mine.handle, c->win) != cudaSuccess)
        ok = 0;
    mine.pid = (long long)getpid();
    mine.ptr = c->win;
    mine.device = c->device;
    cudaGetLastError();
    // exchange the handles (and whether everyone got this far)
    DevBuf<unsigned char> hs, ha;
    DevBuf<int> dok;
    ML_CUDA(c, hs.alloc(sizeof mine));
    ML_CUDA(c, ha.alloc(sizeof mine * c->world));
    ML_CUDA(c, dok.alloc(1));
    std::vector<unsigned char> all(sizeof mine * c->world);
    ML_CUDA(c, cudaMemcpyAsync(hs.p, &mine, sizeof mine, cudaMemcpyHostToDevice, c->stream));
    ML_CUDA(c, cudaMemcpyAsync(dok.p, &ok, sizeof(int), cudaMemcpyHostToDevice, c->stream));
    if (ncclAllGather(hs.p, ha.p, sizeof mine, ncclChar, c->comm, c->stream) != ncclSuccess ||
        ncclAllReduce(dok.p, dok.p, 1, ncclInt, ncclMin, c->comm, c->stream) != ncclSuccess)
        return c->fail(ML_NCCL_ERROR, "p2p_setup: handle exchange");
    ML_CUDA(c, cudaMemcpyAsync(all.data(), ha.p, all.size(), cudaMemcpyDeviceToHost, c->stream));
    ML_CUDA(c, cudaMemcpyAsync(&ok, dok.p, sizeof(int), cudaMemcpyDeviceToHost, c->stream));
    ML_CUDA(c, cudaStreamSynchronize(c->stream));
    if (ok) {
        for (int r = 0; r < c->world && ok; ++r) {
            if (r == c->rank) { c->peer_base[r] = c->win; continue; }
            WinInfo peer;
            std::memcpy(&peer, all.data() + (size_t)r * sizeof peer, sizeof peer);
            if (peer.pid == mine.pid) {
                int can = 0;
                cudaError_t e = cudaDeviceCanAccessPeer(&can, c->device, peer.device);
                if (e == cudaSuccess && can) {
                    e = cudaDeviceEnablePeerAccess(peer.device, 0);
                    if (e == cudaErrorPeerAccessAlreadyEnabled) e = cudaSuccess;
                }
                cudaGetLastError();
                if (e == cudaSuccess && can) c->peer_base[r] = peer.ptr;
                else ok = 0;
            } else if (cudaIpcOpenMemHandle(&c->peer_base[r], peer.handle, cudaIpcMemLazyEnablePeerAccess) == cudaSuccess) {
                c->peer_ipc[r] = true;
            } else {
                c->peer_base[r] = nullptr;
                ok = 0;
            }
        }
        cudaGetLastError();
    }
    // everyone mapped everyone?
    ML_CUDA(c, cudaMemcpyAsync(dok.p, &ok, sizeof(int), cudaMemcpyHostToDevice, c->stream));
    if (ncclAllReduce(dok.p, dok.p, 1, ncclInt, ncclMin, c->comm, c->stream) != ncclSuccess) return c->fail(ML_NCCL_ERROR, "p2p_setup: agreement");
    ML_CUDA(c, cudaMemcpyAsync(&ok, dok.p, sizeof(int), cudaMemcpyDeviceToHost, c->stream));
    ML_CUDA(c, cudaStreamSynchronize(c->stream));
    hs.release();
    ha.release();
    dok.release();
    if (!ok) {
        p2p_teardown(c);
        return ML_OK;
    }
    c->win_n = need;
    c->p2p_seq = 0;
    c->xseq = 0;
    c->p2p_ok = true;
    return ML_OK;
#else
    c->p2p_ok = false;
    return ML_OK;
#endif
}

// y_full[global row of slot s] = gather[s]   (slot = rank * shard_pad + local row; -1 marks padding)
__global__ void compact_shards_kernel(const double* __restrict__ gather, const int* __restrict__ g_of_slot, int n_slots,
                                      double* __restrict__ y_full) {
    const int s = blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= n_slots) return;
    const int g = g_of_slot[s];
    if (g >= 0) y_full[g] = gather[s];
}

struct Sys {            // the (possibly row-sharded) system seen by the solvers
    Ctx* c;
    const double* A;    // local rows, column-major
    int ld, n_rows, n_rows_pad, N;
    int shard_pad;      // rows per rank in the all-gather layout
    DevBuf<double> y_part, gather;
    DevBuf<unsigned> tickets;   // per 64-row block: splits finished (gemv_n_partial_kernel)
    int n_split = 1, cols_per_split = 0;
    int pin_cols = 0;           // columns of every split kept in L2 across matvecs (gemv_n_partial_kernel)
    bool force_shard = false;   // one rank running the multi-rank code paths (MACHLINE_GMRES_SHARDED=1: the single-GPU test box)

    ml_status init() {
        // enough CTAs for >= 4 per SM; splits of at least 256 columns
        int row_blocks = n_rows_pad / GEMV_ROWS;
        if (row_blocks < 1) row_blocks = 1;
        int want = c->num_sms * 4;
        n_split = std::max(1, std::min((want + row_blocks - 1) / row_blocks, std::max(1, N / 256)));
        cols_per_split = (N + n_split - 1) / n_split;
        n_split = (N + cols_per_split - 1) / cols_per_split;
        {   // L2-resident slice of A: GEMV_L2_PIN_MB of the local matrix when it is larger than the cache
            double pin_mb = GEMV_L2_PIN_MB_DEFAULT;
            if (const char* e = std::getenv("MACHLINE_GEMV_L2_PIN_MB")) pin_mb = std::atof(e);
            const double a_bytes = 8.0 * n_rows_pad * N, col_bytes = 8.0 * n_rows_pad * n_split;
            pin_cols = 0;
            if (pin_mb > 0 && a_bytes > 100e6) {
                // eighths of a trip's loads: `base` whole eighths in every trip + one more in a fraction of the trips
                const double frac8 = std::min(7.0, 8.0 * pin_mb * 1e6 / a_bytes);
                const int base = (int)frac8;
                pin_cols = 256 * base + std::min(255, (int)((frac8 - base) * 256));
                (void)col_bytes;
            }
        }
        ML_CUDA(c, y_part.alloc((size_t)n_split * n_rows_pad));
        ML_CUDA(c, tickets.alloc(row_blocks));
        ML_CUDA(c, cudaMemsetAsync(tickets.p, 0, (size_t)row_blocks * sizeof(unsigned), c->stream));
        if (sharded()) {
            ml_status ps = p2p_setup(c, N, force_shard);
            if (ps != ML_OK) return ps;
            ML_CUDA(c, gather.alloc((size_t)shard_pad * c->world));
        }
        return ML_OK;
    }
    // dst_loc[local rows] = alpha * (alpha_dev ? *alpha_dev : 1) * A_loc x      (x: replicated full-length vector)
    ml_status gemv_local(const double* x, double* dst_loc, double alpha, const double* alpha_dev, cudaEvent_t* end_event = nullptr) {
        dim3 grid(n_rows_pad / GEMV_ROWS, n_split);
        cudaEvent_t e0 = nullptr, e1 = nullptr;
        if (c->profile) {
            ML_CUDA(c, cudaEventCreate(&e0));
            ML_CUDA(c, cudaEventCreate(&e1));
            ML_CUDA(c, cudaEventRecord(e0, c->stream));
        }
        gemv_n_partial_kernel<<<grid, GEMV_THREADS, 0, c->stream>>>(A, ld, n_rows_pad, N, cols_per_split, x, y_part.p, tickets.p,
                                                                     n_rows, alpha_dev, alpha, dst_loc, pin_cols);
        if (c->profile) {
            ML_CUDA(c, cudaEventRecord(e1, c->stream));
            c->gemv_ev.push_back(e0);
            c->gemv_ev.push_back(e1);
            c->gemv_launches += 1;
            c->gemv_bytes += (long long)8 * n_rows * N;
            if (end_event) *end_event = e1;
        }
        c->launches += 1;
        ML_CUDA(c, cudaGetLastError());
        return ML_OK;
    }
    // The exchange step: every rank's local rows (src_loc = gather.p + rank * shard_pad) -> the replicated full vector.
    // Peer-memory push + wait (p2p), else ncclAllGather + compaction.  e1 = end of the producing kernel (profiling).
    ml_status exchange(double* y_full, cudaEvent_t e1) {
        double* src = gather.p + (size_t)c->rank * shard_pad;
        if (sharded() && c->p2p_ok) {
            c->p2p_seq += 1;
            const unsigned par = c->p2p_seq & 1u;
            P2PArgs pp;
            for (int r = 0; r < Ctx::P2P_MAX; ++r) {
                double* base = reinterpret_cast<double*>(c->peer_base[r < c->world ? r : c->rank]);
                pp.w[r] = base + (size_t)par * c->win_n;
                pp.flags[r] = reinterpret_cast<unsigned*>(base + 2 * c->win_n) + par * Ctx::P2P_MAX;
            }
            pp.seq = c->p2p_seq;
            pp.rank = c->rank;
            pp.n_rows = n_rows;
            pp.g_of_local = c->d_g_of_slot.p + (size_t)c->rank * shard_pad;
            p2p_push_kernel<<<c->world, 512, 0, c->stream>>>(src, pp);
            c->launches += 1;
            const double* ww = c->win + (size_t)par * c->win_n;
            const unsigned* fl = reinterpret_cast<const unsigned*>(c->win + 2 * c->win_n) + par * Ctx::P2P_MAX;
            p2p_wait_copy_kernel<<<(N + 255) / 256, 256, 0, c->stream>>>(ww, fl, c->p2p_seq, c->world, N, y_full);
            c->launches += 1;
        } else {
#ifdef ML_HAVE_NCCL
            if (c->world > 1) {
                ncclResult_t r = ncclAllGather(src, gather.p, shard_pad, ncclDouble, c->comm, c->stream);
                if (r != ncclSuccess) return c->fail(ML_NCCL_ERROR, ncclGetErrorString(r));
            }
#endif
            // compact the padded shards into the contiguous full vector: one launch (a memcpy per rank costs ~2.5 us each)
            compact_shards_kernel<<<(shard_pad * c->world + 255) / 256, 256, 0, c->stream>>>(gather.p, c->d_g_of_slot.p, shard_pad * c->world,
                                                                                              y_full);
            c->launches += 1;
        }
        if (c->profile && e1) {   // exchange step timed from the end of the local kernel
            cudaEvent_t e2 = nullptr;
            ML_CUDA(c, cudaEventCreate(&e2));
            ML_CUDA(c, cudaEventRecord(e2, c->stream));
            c->comm_ev.push_back(e1);
            c->comm_ev.push_back(e2);
        }
        ML_CUDA(c, cudaGetLastError());
        return ML_OK;
    }
    bool sharded() const { return c->world > 1 || force_shard; }
    // y_full[N] = alpha * (alpha_dev ? *alpha_dev : 1) * A x      (x, y_full replicated full-length vectors)
    ml_status matvec(const double* x, double* y_full, double alpha, const double* alpha_dev) {
        if (!sharded()) return gemv_local(x, y_full, alpha, alpha_dev);
        cudaEvent_t e1 = nullptr;
        ml_status st = gemv_local(x, gather.p + (size_t)c->rank * shard_pad, alpha, alpha_dev, &e1);
        if (st != ML_OK) return st;
        return exchange(y_full, e1);
    }
    void release() {
        y_part.release();
        gather.release();
        tickets.release();
    }
};

static inline double fsign(double a, double b) { return std::signbit(b) ? -std::fabs(a) : std::fabs(a); }

namespace {
struct GmresWork {  // device buffers of one gmres_device call (pool-allocated); the pinned slots and events live in Ctx
    DevBuf<double> Q, w, r0, ydev, hdev, nrm, opart, npart;
    DevBuf<unsigned> bar;         // grid-barrier counter of the fused Arnoldi tail
    static constexpr int MAX_SLOTS = 9;
    double* h_pinned = nullptr;   // n_slots x (k_max + 2), borrowed from Ctx
    cudaEvent_t* ev = nullptr;    // borrowed from Ctx
    void release() {
        Q.release(); w.release(); r0.release(); ydev.release(); hdev.release(); nrm.release(); opart.release(); npart.release();
        bar.release();
    }
};
}  // namespace

// GMRES / restarted GMRES (linalg.f90:1235-1453) on the scaled system (scale*A) x = scale*b.
// Iteration history as the reference writes it (linalg.f90:1273-1280, 1316; 1376-1383, 1438): list-directed header lines
// (leading blank, default integer width 12), then '(i6, a, ES10.3)' rows.  nullptr / "none": no file.  Rank 0 only.
static FILE* open_iteration_file(Ctx* c, const char* path, const char* method, int N, const char* columns) {
    if (!path || !*path || std::strcmp(path, "none") == 0 || c->rank != 0) return nullptr;
    FILE* f = std::fopen(path, "w");
    if (!f) return nullptr;
    std::fprintf(f, " method\n %s\n N=%12d\n %s\n", method, N, columns);
    return f;
}

struct HistFile {   // closes on every return path
    FILE* f;
    ~HistFile() {
        if (f) std::fclose(f);
    }
};

static ml_status gmres_device(Sys& S, const double* d_b, const double* d_scale, double tol, int max_iter, int restart_iter,
                              bool restarted, bool use_mgs, double* d_x, int* total_iter_out, const char* iteration_file) {
    Ctx* c = S.c;
    const int N = S.N;
    HistFile hist{open_iteration_file(c, iteration_file, "GMRES", N,
                                      restarted ? "iteration,outer iteration,inner iteration,||err||" : "iteration,||err||")};
    int outer_iter = 0;
    const int k_max = restarted ? std::min(restart_iter, N) : std::min(N, max_iter);
    if (k_max < 1) return c->fail(ML_BAD_ARGUMENT, "max_iterations < 1");
    const int hs = k_max + 2;  // stride of one Hessenberg-column slot
    GmresWork W;
    auto fail_cuda = [&](cudaError_t e, const char* where) {
        cudaStreamSynchronize(c->stream);
        W.release();
        return c->cuda_fail(e, where);
    };
#define GM_CUDA(call)                                        \
    do {                                                     \
        cudaError_t e__ = (call);                            \
        if (e__ != cudaSuccess) return fail_cuda(e__, #call); \
    } while (0)
    const int ldq = ((N + 63) / 64) * 64;                    // padded so that 64-row CTAs can use 16-byte loads unguarded
    const int n_row64 = ldq / 64, kpad = k_max + 4;
    // fused Arnoldi tail (one cooperative launch per iteration); MACHLINE_GMRES_TAIL=0 selects the five-kernel sequence
    bool fused_tail = !use_mgs;
    int tail_version = 3;   // 3: column-owned dots + Hessenberg column stored to mapped host memory; 1: the first fused kernel
    if (const char* e = std::getenv("MACHLINE_GMRES_TAIL")) {
        fused_tail = fused_tail && std::atoi(e) != 0;
        if (std::atoi(e) == 1) tail_version = 1;
    }
    // the Gram-Schmidt kernels keep the k coefficients of a pass in shared memory: 200 KB opt-in for the fused tail (k_max <= 24.5k),
    // 64 KB for the five-kernel sequence (k_max <= 7.6k); a Krylov space beyond that is refused, not mis-launched
    constexpr size_t TAIL_SMEM_MAX = 200 * 1024, SUB_SMEM_MAX = 64 * 1024;
    if (!use_mgs) {
        const size_t need_fused = (size_t)(((k_max + 1) & ~1) + 1024) * sizeof(double);
        const size_t need_split = (size_t)(((k_max + 1) & ~1) + 8 * 64 + 2) * sizeof(double);
        if (fused_tail && need_fused > TAIL_SMEM_MAX) fused_tail = false;
        if (!fused_tail && need_split > SUB_SMEM_MAX) {
            if (need_fused <= TAIL_SMEM_MAX) fused_tail = true;
            else return c->fail(ML_UNSUPPORTED, "GMRES: max_iterations (the Krylov dimension) is limited to 24 000 on the device; use RGMRES");
        }
    }
    const int tail_grid = c->num_sms;
    const int tail_rows = 32 * ((N + 32 * tail_grid - 1) / (32 * tail_grid));
    const int n_rb = fused_tail ? (N + TAIL_RB - 1) / TAIL_RB : (N + ORTH_RB - 1) / ORTH_RB;
    unsigned bar_base = 0;
    GM_CUDA(W.Q.alloc((size_t)ldq * (k_max + 1)));
    GM_CUDA(W.w.alloc(ldq));
    GM_CUDA(W.r0.alloc(ldq));
    GM_CUDA(W.opart.alloc((size_t)n_rb * kpad));
    GM_CUDA(W.npart.alloc(std::max(n_row64, tail_grid)));
    GM_CUDA(W.bar.alloc(1));
    GM_CUDA(cudaMemsetAsync(W.bar.p, 0, sizeof(unsigned), c->stream));
    GM_CUDA(cudaMemsetAsync(W.Q.p, 0, (size_t)ldq * (k_max + 1) * sizeof(double), c->stream));
    GM_CUDA(cudaMemsetAsync(W.w.p, 0, (size_t)ldq * sizeof(double), c->stream));
    {
        // the opt-in is per device: remembered per context, not per process
        if (!(c->attr_mask & 1u)) {
            GM_CUDA(cudaFuncSetAttribute(orth_sub_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024));
            GM_CUDA(cudaFuncSetAttribute(arnoldi_tail_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)TAIL_SMEM_MAX));
            GM_CUDA(cudaFuncSetAttribute(arnoldi_tail3_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)TAIL_SMEM_MAX));
            c->attr_mask |= 1u;
        }
    }
    GM_CUDA(W.ydev.alloc(k_max + 1));
    // Look-ahead: Arnoldi steps kept in flight beyond the one whose Hessenberg column the host is examining.  One is
    // enough on a quiet host; a few more make the device immune to host scheduling hiccups (an iteration is ~0.2 ms)
    // at the price of that many discarded steps at convergence.
    int depth = 4;
    if (const char* e = std::getenv("MACHLINE_GMRES_LOOKAHEAD")) depth = std::max(1, std::min(GmresWork::MAX_SLOTS - 1, std::atoi(e)));
    const int n_slots = depth + 1;
    GM_CUDA(W.hdev.alloc((size_t)3 * n_slots * hs));  // per slot: hfin (final column), h1 and h2 (the two Gram-Schmidt passes)
    GM_CUDA(W.nrm.alloc(2));
    if (c->h_pinned_n < (size_t)n_slots * hs) {
        if (c->h_pinned) cudaFreeHost(c->h_pinned);
        c->h_pinned = nullptr;
        c->h_pinned_n = 0;
        GM_CUDA(cudaHostAlloc((void**)&c->h_pinned, (size_t)n_slots * hs * sizeof(double), cudaHostAllocMapped));
        c->h_pinned_n = (size_t)n_slots * hs;
    }
    W.h_pinned = c->h_pinned;
    DevBuf<long long> tail_dbg;
    if (std::getenv("MACHLINE_TAIL_DEBUG")) {
        GM_CUDA(tail_dbg.alloc(17));
        GM_CUDA(cudaMemsetAsync(tail_dbg.p, 0, 17 * sizeof(long long), c->stream));
    }
    double* h_pinned_dev = nullptr;   // the device's address of the pinned slots (tail version 3 stores the Hessenberg column there)
    if (fused_tail && tail_version == 3 && std::getenv("MACHLINE_GMRES_D2H_COPY") == nullptr && cudaHostGetDevicePointer((void**)&h_pinned_dev, c->h_pinned, 0) != cudaSuccess) {
        cudaGetLastError();
        h_pinned_dev = nullptr;
    }
    for (int i = 0; i < n_slots; ++i)
        if (!c->slot_ev[i]) GM_CUDA(cudaEventCreateWithFlags(&c->slot_ev[i], cudaEventDisableTiming));
    W.ev = c->slot_ev;

    const int ldh = k_max + 1;
    std::vector<double> H((size_t)ldh * k_max, 0.), cs(k_max, 0.), sn(k_max, 0.), E(std::max(N, k_max + 1) + 1, 0.);
    GM_CUDA(cudaMemsetAsync(d_x, 0, (size_t)N * sizeof(double), c->stream));
    int total_iter = 0;
    double err = tol + 1;
    const int nb256 = (N + 255) / 256, nb64 = (N + 63) / 64;
    ml_status st = ML_OK;

    // enqueue Arnoldi step kk (0-based): Q(:,kk+1), Hessenberg column -> pinned slot, event
    auto enqueue = [&](int kk) -> ml_status {
        const int k = kk + 1, slot = kk % n_slots;
        double* hfin = W.hdev.p + (size_t)slot * 3 * hs;  // final column h[0..k]
        double* h1 = hfin + hs;                            // first-pass coefficients
        double* h2 = h1 + hs;                              // second-pass corrections (hfin = h1 + h2)
        ml_status s = S.matvec(W.Q.p + (size_t)kk * ldq, W.w.p, 1.0, d_scale);
        if (s != ML_OK) return s;
        if (use_mgs) {
            mgs_kernel<<<1, 1024, 0, c->stream>>>(W.Q.p, ldq, N, k, W.w.p, hfin);
            norm_scale_kernel<<<1, 1024, 0, c->stream>>>(W.w.p, N, hfin + k, W.Q.p + (size_t)k * ldq);
            c->launches += 2;
        } else if (fused_tail && tail_version == 3) {
            Tail3Args ta;
            ta.Q = W.Q.p; ta.ldq = ldq; ta.n = N; ta.k = k;
            ta.w = W.w.p; ta.h1 = h1; ta.h2 = h2; ta.hfin = hfin;
            ta.hhost = h_pinned_dev ? h_pinned_dev + (size_t)slot * hs : nullptr;
            ta.npart = W.npart.p; ta.qnext = W.Q.p + (size_t)k * ldq;
            ta.bar = W.bar.p; ta.bar_base = bar_base; ta.rows_per_cta = tail_rows; ta.dbg = tail_dbg.p;
            bar_base += 4u * (unsigned)tail_grid;
            void* kargs[] = {(void*)&ta};
            const size_t smem = (size_t)(((k + 1) & ~1) + 1024) * sizeof(double);
            cudaError_t e = cudaLaunchCooperativeKernel((const void*)arnoldi_tail3_kernel, dim3(tail_grid), dim3(TAIL_THREADS), kargs, smem,
                                                        c->stream);
            if (e != cudaSuccess) return c->cuda_fail(e, "arnoldi_tail3_kernel");
            c->launches += 1;
            if (h_pinned_dev) {   // the column is already on its way to the host: only the event follows
                e = cudaEventRecord(W.ev[slot], c->stream);
                if (e != cudaSuccess) return c->cuda_fail(e, "gmres enqueue");
                c->d2h_bytes += (long long)(k + 1) * sizeof(double);
                return ML_OK;
            }
        } else if (fused_tail) {
            // classical Gram-Schmidt with one re-orthogonalisation pass (same Krylov subspace and Hessenberg matrix as
            // the reference's modified Gram-Schmidt up to rounding), norm and normalisation: one cooperative launch
            TailArgs ta;
            ta.Q = W.Q.p; ta.ldq = ldq; ta.n = N; ta.k = k;
            ta.w = W.w.p; ta.partial = W.opart.p; ta.kpad = kpad; ta.n_rb = n_rb;
            ta.h1 = h1; ta.hfin = hfin; ta.npart = W.npart.p; ta.qnext = W.Q.p + (size_t)k * ldq;
            ta.bar = W.bar.p; ta.bar_base = bar_base; ta.rows_per_cta = tail_rows;
            bar_base += 4u * (unsigned)tail_grid;
            void* kargs[] = {(void*)&ta};
            const size_t smem = (size_t)(((k + 1) & ~1) + 1024) * sizeof(double);
            cudaError_t e = cudaLaunchCooperativeKernel((const void*)arnoldi_tail_kernel, dim3(tail_grid), dim3(TAIL_THREADS), kargs, smem,
                                                        c->stream);
            if (e != cudaSuccess) return c->cuda_fail(e, "arnoldi_tail_kernel");
            c->launches += 1;
        } else {
            const dim3 gdot(n_rb, (k + ORTH_CG - 1) / ORTH_CG);
            const size_t sub_smem = (size_t)(((k + 1) & ~1) + 8 * 64 + 2) * sizeof(double);
            orth_dot_kernel<<<gdot, 256, 0, c->stream>>>(W.Q.p, ldq, N, k, W.w.p, W.opart.p, kpad);
            orth_sub_kernel<<<n_row64, 256, sub_smem, c->stream>>>(W.Q.p, ldq, N, k, W.opart.p, kpad, n_rb, W.w.p, nullptr, h1, nullptr);
            orth_dot_kernel<<<gdot, 256, 0, c->stream>>>(W.Q.p, ldq, N, k, W.w.p, W.opart.p, kpad);
            orth_sub_kernel<<<n_row64, 256, sub_smem, c->stream>>>(W.Q.p, ldq, N, k, W.opart.p, kpad, n_rb, W.w.p, h1, hfin, W.npart.p);
            norm_scale2_kernel<<<(N + 1023) / 1024, 1024, 0, c->stream>>>(W.w.p, N, W.npart.p, n_row64, hfin + k, W.Q.p + (size_t)k * ldq);
            c->launches += 5;
        }
        cudaError_t e = cudaMemcpyAsync(W.h_pinned + (size_t)slot * hs, hfin, (size_t)(k + 1) * sizeof(double),
                                        cudaMemcpyDeviceToHost, c->stream);
        if (e == cudaSuccess) e = cudaEventRecord(W.ev[slot], c->stream);
        if (e != cudaSuccess) return c->cuda_fail(e, "gmres enqueue");
        c->d2h_bytes += (long long)(k + 1) * sizeof(double);
        return ML_OK;
    };

    bool first_cycle = true;
    while (err > tol && (restarted ? total_iter <= max_iter : first_cycle)) {
        first_cycle = false;
        outer_iter += 1;
        std::fill(H.begin(), H.end(), 0.);
        std::fill(cs.begin(), cs.end(), 0.);
        std::fill(sn.begin(), sn.end(), 0.);
        std::fill(E.begin(), E.end(), 0.);
        E[0] = 1.;
        // r0 = scale*b - (scale*A) x   (x = 0 in the first cycle: r0 = scale*b exactly as linalg.f90:1268)
        scale_copy_kernel<<<nb256, 256, 0, c->stream>>>(d_b, 1.0, d_scale, W.r0.p, N);
        c->launches += 1;
        if (restarted && total_iter > 0) {
            st = S.matvec(d_x, W.w.p, 1.0, d_scale);
            if (st != ML_OK) break;
            axpby_kernel<<<nb256, 256, 0, c->stream>>>(1.0, W.r0.p, -1.0, W.w.p, W.r0.p, N);
            c->launches += 1;
        }
        norm_scale_kernel<<<1, 1024, 0, c->stream>>>(W.r0.p, N, W.nrm.p, W.Q.p);  // beta, Q(:,1) = r0/beta
        c->launches += 1;
        double beta = 0.;
        GM_CUDA(cudaMemcpyAsync(&beta, W.nrm.p, sizeof(double), cudaMemcpyDeviceToHost, c->stream));
        GM_CUDA(cudaStreamSynchronize(c->stream));
        if (!(beta == beta)) {
            st = ML_NAN_IN_SYSTEM;
            break;
        }
        int k = 0;
        const int k_last = k_max - 1;  // the loop condition k < k_max-1 allows steps k = 1..k_max-1
        int next_enq = 0;   // Arnoldi steps enqueued so far in this cycle
        auto fill = [&](int upto) -> ml_status {   // speculatively enqueue steps up to index `upto`
            while (next_enq <= upto && next_enq < k_last) {
                ml_status s = enqueue(next_enq);
                if (s != ML_OK) return s;
                ++next_enq;
            }
            return ML_OK;
        };
        if (err > tol) st = fill(depth - 1);
        if (st != ML_OK) break;
        while (err > tol && k < k_last && k < next_enq) {
            k += 1;
            total_iter += 1;
            const int kk = k - 1, slot = kk % n_slots;
            // keep `depth` steps in flight beyond the one examined now
            st = fill(kk + depth);
            if (st != ML_OK) break;
            GM_CUDA(cudaEventSynchronize(W.ev[slot]));
            const double* hcol = W.h_pinned + (size_t)slot * hs;
            for (int i = 0; i <= k; ++i) H[i + (size_t)kk * ldh] = hcol[i];
            // Givens updates (linalg.f90:1293-1313) on the host: k+1 numbers per iteration
            for (int i = 0; i < kk; ++i) {
                double temp = cs[i] * H[i + (size_t)kk * ldh] + sn[i] * H[(i + 1) + (size_t)kk * ldh];
                H[(i + 1) + (size_t)kk * ldh] = -sn[i] * H[i + (size_t)kk * ldh] + cs[i] * H[(i + 1) + (size_t)kk * ldh];
                H[i + (size_t)kk * ldh] = temp;
            }
            const double hkk = H[kk + (size_t)kk * ldh], hk1 = H[(kk + 1) + (size_t)kk * ldh];
            const double d = std::sqrt(hkk * hkk + hk1 * hk1);
            cs[kk] = std::fabs(hkk) / d;
            sn[kk] = fsign(1., hkk) * hk1 / d;
            H[kk + (size_t)kk * ldh] = cs[kk] * hkk + sn[kk] * hk1;
            H[(kk + 1) + (size_t)kk * ldh] = 0.;
            E[kk + 1] = -sn[kk] * E[kk];
            E[kk] = cs[kk] * E[kk];
            err = beta * std::fabs(E[kk + 1]);
            if (hist.f) {
                if (restarted) std::fprintf(hist.f, "%6d,%6d,%6d,%10.3E\n", total_iter, outer_iter, k, err);
                else std::fprintf(hist.f, "%6d,%10.3E\n", k, err);
            }
            if (!restarted && err < tol) break;
        }
        if (st != ML_OK) break;
        GM_CUDA(cudaStreamSynchronize(c->stream));  // drain the speculative step, if any
        if (k == 0) break;
        // back substitution (linalg.f90:930-965) and x (+)= Q(:,1:k) y
        std::vector<double> y(k);
        bool singular = false;
        for (int i = k - 1; i >= 0; --i) {
            double v = beta * E[i];
            for (int j = i + 1; j < k; ++j) v = v - H[i + (size_t)j * ldh] * y[j];
            if (H[i + (size_t)i * ldh] == 0.) {
                singular = true;
                break;
            }
            y[i] = v / H[i + (size_t)i * ldh];
        }
        if (singular) {
            st = c->fail(ML_SINGULAR, "Zero found on the diagonal of R (linalg.f90:956-961)");
            break;
        }
        GM_CUDA(cudaMemcpyAsync(W.ydev.p, y.data(), (size_t)k * sizeof(double), cudaMemcpyHostToDevice, c->stream));
        gemv_n_small_kernel<<<nb64, 64, 0, c->stream>>>(W.Q.p, ldq, N, k, W.ydev.p, d_x, restarted ? 1 : 0);
        c->launches += 1;
        GM_CUDA(cudaStreamSynchronize(c->stream));
    }
    cudaStreamSynchronize(c->stream);
    if (tail_dbg.p) {
        long long d[17];
        if (cudaMemcpy(d, tail_dbg.p, sizeof(d), cudaMemcpyDeviceToHost) == cudaSuccess && d[16] > 0) {
            static const char* names[] = {"dot1", "bar1", "sub1", "bar2", "dot2", "bar3", "sub2", "bar4", "norm+scale"};
            std::fprintf(stderr, "[tail3] %lld launches, cycles per launch (CTA 0):", d[16]);
            for (int i = 0; i < 9; ++i) std::fprintf(stderr, " %s %.0f", names[i], (double)d[i] / (double)d[16]);
            std::fprintf(stderr, "\n");
        }
    }
    *total_iter_out = total_iter;
    W.release();
#undef GM_CUDA
    return st;
}

}  // namespace mlgpu
#include "gmres_sharded.cuh"
namespace mlgpu {

// local rows of a replicated vector: dst[i] = src[g_of_local[i]]
__global__ void gather_rows_kernel(const double* __restrict__ src, const int* __restrict__ g_of_local, int n_loc, double* __restrict__ dst) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n_loc) dst[i] = src[g_of_local[i]];
}

// GMRES / restarted GMRES on a row-sharded system with the basis sharded by rows (gmres_sharded.cuh).  Same host pipeline as
// gmres_device: the Hessenberg column of step k comes back through a pinned slot while later steps are already enqueued; every
// rank receives bitwise identical columns, takes identical decisions and therefore launches the identical kernel sequence.
static ml_status gmres_sharded_device(Sys& S, const double* d_b, const double* d_scale, double tol, int max_iter, int restart_iter,
                                      bool restarted, double* d_x, int* total_iter_out, const char* iteration_file) {
    Ctx* c = S.c;
    const int N = S.N, n_loc = S.n_rows;
    HistFile hist{open_iteration_file(c, iteration_file, "GMRES", N,
                                      restarted ? "iteration,outer iteration,inner iteration,||err||" : "iteration,||err||")};
    int outer_iter = 0;
    const int k_max = restarted ? std::min(restart_iter, N) : std::min(N, max_iter);
    if (k_max < 1) return c->fail(ML_BAD_ARGUMENT, "max_iterations < 1");
    const int hs = k_max + 3;   // slot stride: h[0..k], the error flag
    const int ldq = S.n_rows_pad;
    const int nfull = ((N + 63) / 64) * 64;
    // one CTA per SM at most; every CTA owns a block of >= 32 local rows
    int grid = std::max(1, std::min(std::min(c->num_sms, 160), (n_loc + 31) / 32));
    const int rows_per_cta = 32 * ((n_loc + 32 * grid - 1) / (32 * grid));
    grid = std::max(1, (n_loc + rows_per_cta - 1) / rows_per_cta);
    const int kpad = k_max + 4;
    DevBuf<double> Q, w, r0, xfull, xloc, ydev, hdev, nrm, part, npart, h1;
    DevBuf<unsigned> sync;   // [0] ticket, [1] ready
    DevBuf<int> err;
    DevBuf<long long> dbg;
    const bool want_dbg = std::getenv("MACHLINE_SHT_DEBUG") != nullptr;
#define GS_CUDA(call)                                            \
    do {                                                         \
        cudaError_t e__ = (call);                                \
        if (e__ != cudaSuccess) {                                \
            cudaStreamSynchronize(c->stream);                    \
            return c->cuda_fail(e__, #call);                     \
        }                                                        \
    } while (0)
    GS_CUDA(Q.alloc((size_t)ldq * (k_max + 1)));
    GS_CUDA(w.alloc(ldq));
    GS_CUDA(r0.alloc(nfull));
    GS_CUDA(xfull.alloc(nfull));
    GS_CUDA(xloc.alloc((size_t)S.shard_pad * c->world));
    const int gpad = ((grid + 31) / 32) * 32;
    GS_CUDA(part.alloc((size_t)gpad * kpad));
    GS_CUDA(npart.alloc(grid));
    GS_CUDA(h1.alloc(kpad));
    GS_CUDA(sync.alloc(2));
    GS_CUDA(err.alloc(1));
    GS_CUDA(ydev.alloc(k_max + 1));
    GS_CUDA(nrm.alloc(2));
    GS_CUDA(cudaMemsetAsync(sync.p, 0, 2 * sizeof(unsigned), c->stream));
    GS_CUDA(cudaMemsetAsync(err.p, 0, sizeof(int), c->stream));
    if (want_dbg) {
        GS_CUDA(dbg.alloc(32));
        GS_CUDA(cudaMemsetAsync(dbg.p, 0xff, 32 * sizeof(long long), c->stream));
    }
    GS_CUDA(cudaMemsetAsync(Q.p, 0, (size_t)ldq * (k_max + 1) * sizeof(double), c->stream));
    GS_CUDA(cudaMemsetAsync(w.p, 0, (size_t)ldq * sizeof(double), c->stream));
    GS_CUDA(cudaMemsetAsync(xloc.p, 0, (size_t)S.shard_pad * c->world * sizeof(double), c->stream));
    if (!(c->attr_mask & 16u)) {
        GS_CUDA(cudaFuncSetAttribute(arnoldi_tail_sharded_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 160 * 1024));
        GS_CUDA(cudaFuncSetAttribute(arnoldi_tail_sharded2_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 160 * 1024));
        c->attr_mask |= 16u;
    }
    // column-owned dot products (gmres_sharded.cuh, version 2) unless MACHLINE_SHT_V1 asks for the row-owned first version
    const bool tail_v2 = std::getenv("MACHLINE_SHT_V1") == nullptr;
    int depth = 4;
    if (const char* e = std::getenv("MACHLINE_GMRES_LOOKAHEAD")) depth = std::max(1, std::min(8, std::atoi(e)));
    const int n_slots = depth + 1;
    GS_CUDA(hdev.alloc((size_t)n_slots * hs));
    if (c->h_pinned_n < (size_t)n_slots * hs) {
        if (c->h_pinned) cudaFreeHost(c->h_pinned);
        c->h_pinned = nullptr;
        c->h_pinned_n = 0;
        GS_CUDA(cudaHostAlloc((void**)&c->h_pinned, (size_t)n_slots * hs * sizeof(double), cudaHostAllocMapped));
        c->h_pinned_n = (size_t)n_slots * hs;
    }
    double* h_pinned_dev = nullptr;   // the device's address of the pinned slots
    if (std::getenv("MACHLINE_GMRES_D2H_COPY") == nullptr &&
        cudaHostGetDevicePointer((void**)&h_pinned_dev, c->h_pinned, 0) != cudaSuccess) {
        cudaGetLastError();
        h_pinned_dev = nullptr;
    }
    for (int i = 0; i < n_slots; ++i)
        if (!c->slot_ev[i]) GS_CUDA(cudaEventCreateWithFlags(&c->slot_ev[i], cudaEventDisableTiming));
    const int* g_of_local = c->d_g_of_slot.p + (size_t)c->rank * S.shard_pad;

    const int ldh = k_max + 1;
    std::vector<double> H((size_t)ldh * k_max, 0.), cs(k_max, 0.), sn(k_max, 0.), E(std::max(N, k_max + 1) + 1, 0.);
    GS_CUDA(cudaMemsetAsync(d_x, 0, (size_t)N * sizeof(double), c->stream));
    int total_iter = 0;
    double err_est = tol + 1;
    const int nb256 = (N + 255) / 256;
    ml_status st = ML_OK;
    unsigned launches_done = 0;

    auto enqueue = [&](int kk) -> ml_status {   // Arnoldi step kk (0-based): w_loc = A_loc q_kk, then the sharded tail
        const int k = kk + 1, slot = kk % n_slots;
        double* hfin = hdev.p + (size_t)slot * hs;
        cudaEvent_t e1 = nullptr;
        ml_status s = S.gemv_local(xfull.p, w.p, 1.0, d_scale, &e1);
        if (s != ML_OK) return s;
        ShTailArgs a{};
        a.Q = Q.p; a.ldq = ldq; a.n_loc = n_loc; a.k = k;
        a.w = w.p; a.partial = part.p; a.gpad = gpad; a.h1 = h1.p; a.hfin = hfin; a.npart = npart.p;
        a.qnext = Q.p + (size_t)k * ldq; a.xfull = xfull.p; a.N = N; a.g_of_local = g_of_local;
        a.P = c->world; a.rank = c->rank; a.kr = Ctx::P2P_KR; a.nv = (int)c->win_n;
        for (int r = 0; r < Ctx::P2P_MAX; ++r) {
            double* base = reinterpret_cast<double*>(c->peer_base[r < c->world ? r : c->rank]);
            a.flags[r] = reinterpret_cast<unsigned*>(base + 2 * c->win_n + 8);
            a.vec[r] = base + 2 * c->win_n + 16;
            a.red[r] = base + 4 * c->win_n + 16;
        }
        a.seq = c->xseq + 1;
        c->xseq += 3;
        a.ticket = sync.p; a.ready = sync.p + 1; a.base = launches_done++;
        a.stages = tail_v2 ? 4u : 3u;
        a.err = err.p; a.rows_per_cta = rows_per_cta; a.dbg = want_dbg ? dbg.p : nullptr;
        a.hhost = (tail_v2 && h_pinned_dev) ? h_pinned_dev + (size_t)slot * hs : nullptr;
        void* kargs[] = {(void*)&a};
        const size_t smem = (size_t)(2 * ((k + 3) & ~1) + SHT_MAXCH * SHT_THREADS + rows_per_cta) * sizeof(double);
        cudaError_t e = cudaLaunchCooperativeKernel(tail_v2 ? (const void*)arnoldi_tail_sharded2_kernel : (const void*)arnoldi_tail_sharded_kernel,
                                                    dim3(grid), dim3(SHT_THREADS), kargs, smem, c->stream);
        if (e != cudaSuccess) return c->cuda_fail(e, "arnoldi_tail_sharded_kernel");
        c->launches += 1;
        if (c->profile && e1) {   // the exchanges live inside the tail: "exchange" = the whole tail, from the end of the local matvec
            cudaEvent_t e2 = nullptr;
            if (cudaEventCreate(&e2) == cudaSuccess && cudaEventRecord(e2, c->stream) == cudaSuccess) {
                c->comm_ev.push_back(e1);
                c->comm_ev.push_back(e2);
            }
        }
        e = a.hhost ? cudaSuccess   // version 2 stored the column into the mapped slot itself: no copy-engine operation between kernels
                    : cudaMemcpyAsync(c->h_pinned + (size_t)slot * hs, hfin, (size_t)(k + 2) * sizeof(double), cudaMemcpyDeviceToHost, c->stream);
        if (e == cudaSuccess) e = cudaEventRecord(c->slot_ev[slot], c->stream);
        if (e != cudaSuccess) return c->cuda_fail(e, "gmres enqueue");
        c->d2h_bytes += (long long)(k + 2) * sizeof(double);
        return ML_OK;
    };

    bool first_cycle = true;
    while (err_est > tol && (restarted ? total_iter <= max_iter : first_cycle)) {
        first_cycle = false;
        outer_iter += 1;
        std::fill(H.begin(), H.end(), 0.);
        std::fill(cs.begin(), cs.end(), 0.);
        std::fill(sn.begin(), sn.end(), 0.);
        std::fill(E.begin(), E.end(), 0.);
        E[0] = 1.;
        // r0 = scale*b - (scale*A) x on the replicated vectors (x = 0 in the first cycle)
        scale_copy_kernel<<<nb256, 256, 0, c->stream>>>(d_b, 1.0, d_scale, r0.p, N);
        c->launches += 1;
        if (restarted && total_iter > 0) {
            st = S.matvec(d_x, xfull.p, 1.0, d_scale);
            if (st != ML_OK) break;
            axpby_kernel<<<nb256, 256, 0, c->stream>>>(1.0, r0.p, -1.0, xfull.p, r0.p, N);
            c->launches += 1;
        }
        norm_scale_kernel<<<1, 1024, 0, c->stream>>>(r0.p, N, nrm.p, xfull.p);  // beta, q_1 = r0/beta (replicated)
        gather_rows_kernel<<<(n_loc + 255) / 256, 256, 0, c->stream>>>(xfull.p, g_of_local, n_loc, Q.p);
        c->launches += 2;
        double beta = 0.;
        GS_CUDA(cudaMemcpyAsync(&beta, nrm.p, sizeof(double), cudaMemcpyDeviceToHost, c->stream));
        GS_CUDA(cudaStreamSynchronize(c->stream));
        if (!(beta == beta)) {
            st = ML_NAN_IN_SYSTEM;
            break;
        }
        int k = 0;
        const int k_last = k_max - 1;
        int next_enq = 0;
        auto fill = [&](int upto) -> ml_status {
            while (next_enq <= upto && next_enq < k_last) {
                ml_status s = enqueue(next_enq);
                if (s != ML_OK) return s;
                ++next_enq;
            }
            return ML_OK;
        };
        if (err_est > tol) st = fill(depth - 1);
        if (st != ML_OK) break;
        while (err_est > tol && k < k_last && k < next_enq) {
            k += 1;
            total_iter += 1;
            const int kk = k - 1, slot = kk % n_slots;
            st = fill(kk + depth);
            if (st != ML_OK) break;
            GS_CUDA(cudaEventSynchronize(c->slot_ev[slot]));
            const double* hcol = c->h_pinned + (size_t)slot * hs;
            if (hcol[k + 1] != 0.) {
                st = c->fail(ML_NCCL_ERROR, "sharded GMRES: a peer did not answer within the spin limit");
                break;
            }
            for (int i = 0; i <= k; ++i) H[i + (size_t)kk * ldh] = hcol[i];
            for (int i = 0; i < kk; ++i) {   // Givens updates, linalg.f90:1293-1313
                double temp = cs[i] * H[i + (size_t)kk * ldh] + sn[i] * H[(i + 1) + (size_t)kk * ldh];
                H[(i + 1) + (size_t)kk * ldh] = -sn[i] * H[i + (size_t)kk * ldh] + cs[i] * H[(i + 1) + (size_t)kk * ldh];
                H[i + (size_t)kk * ldh] = temp;
            }
            const double hkk = H[kk + (size_t)kk * ldh], hk1 = H[(kk + 1) + (size_t)kk * ldh];
            const double d = std::sqrt(hkk * hkk + hk1 * hk1);
            cs[kk] = std::fabs(hkk) / d;
            sn[kk] = fsign(1., hkk) * hk1 / d;
            H[kk + (size_t)kk * ldh] = cs[kk] * hkk + sn[kk] * hk1;
            H[(kk + 1) + (size_t)kk * ldh] = 0.;
            E[kk + 1] = -sn[kk] * E[kk];
            E[kk] = cs[kk] * E[kk];
            err_est = beta * std::fabs(E[kk + 1]);
            if (hist.f) {
                if (restarted) std::fprintf(hist.f, "%6d,%6d,%6d,%10.3E\n", total_iter, outer_iter, k, err_est);
                else std::fprintf(hist.f, "%6d,%10.3E\n", k, err_est);
            }
            if (!restarted && err_est < tol) break;
        }
        if (st != ML_OK) break;
        GS_CUDA(cudaStreamSynchronize(c->stream));  // drain the speculative steps
        if (k == 0) break;
        std::vector<double> y(k);
        bool singular = false;
        for (int i = k - 1; i >= 0; --i) {   // linalg.f90:930-965
            double v = beta * E[i];
            for (int j = i + 1; j < k; ++j) v = v - H[i + (size_t)j * ldh] * y[j];
            if (H[i + (size_t)i * ldh] == 0.) {
                singular = true;
                break;
            }
            y[i] = v / H[i + (size_t)i * ldh];
        }
        if (singular) {
            st = c->fail(ML_SINGULAR, "Zero found on the diagonal of R (linalg.f90:956-961)");
            break;
        }
        // x (+)= Q y: local rows, then the exchange of the local parts (xloc doubles as the gather buffer layout)
        GS_CUDA(cudaMemcpyAsync(ydev.p, y.data(), (size_t)k * sizeof(double), cudaMemcpyHostToDevice, c->stream));
        gemv_n_small_kernel<<<(n_loc + 63) / 64, 64, 0, c->stream>>>(Q.p, ldq, n_loc, k, ydev.p, S.gather.p + (size_t)c->rank * S.shard_pad, 0);
        c->launches += 1;
        st = S.exchange(xfull.p, nullptr);
        if (st != ML_OK) break;
        if (restarted) axpby_kernel<<<nb256, 256, 0, c->stream>>>(1.0, d_x, 1.0, xfull.p, d_x, N);
        else axpby_kernel<<<nb256, 256, 0, c->stream>>>(1.0, xfull.p, 0.0, xfull.p, d_x, N);
        c->launches += 1;
        GS_CUDA(cudaStreamSynchronize(c->stream));
    }
    cudaStreamSynchronize(c->stream);
    if (want_dbg) {   // clock64 stamps of the last launch: CTA 0 and the CTA that was last at stage 0
        long long hd[32];
        if (cudaMemcpy(hd, dbg.p, sizeof hd, cudaMemcpyDeviceToHost) == cudaSuccess) {
            for (int r = 0; r < 2; ++r) {
                std::fprintf(stderr, "sht stamps %s:", r ? "last-CTA" : "CTA0");
                for (int i = 0; i < 16; ++i) std::fprintf(stderr, " %lld", hd[16 * r + i]);
                std::fprintf(stderr, "\n");
            }
        }
    }
    *total_iter_out = total_iter;
#undef GS_CUDA
    return st;
}

ml_status lu_solve_device(Ctx* c, int N, double* dA, int ld, const double* d_b, double* d_x);  // lu_kernels.cu
ml_status lu_solve_sharded(Ctx* c, int N, double* dAloc, int ld, int n_rows, int n_rows_pad, int S, const double* d_b,
                           double* d_x);                                                      // lu_kernels.cu
ml_status block_jacobi_device(Ctx* c, int N, const double* dA, int ld, const double* d_b, int block_size, double tol, double rel,
                              int max_iter, int* iters, double* d_x, double err_scale, const char* iteration_file);   // lu_kernels.cu
ml_status block_ssor_device(Ctx* c, int N, const double* dA, int ld, const double* d_b, int block_size, double tol, double rel,
                            int max_iter, int* iters, double* d_x, double err_scale, const char* iteration_file);     // lu_kernels.cu
ml_status qrup_solve_device(Ctx* c, int N, const double* dA, int ld, const double* d_b, const double* d_scale, bool fast,
                            double* d_x);                                                    // seq_solvers.cu
ml_status purcell_solve_device(Ctx* c, int N, const double* dA, int ld, const double* d_b, const double* d_scale,
                               double* d_x);                                                 // seq_solvers.cu

static bool needs_whole_matrix(int matrix_solver) {   // everything but the Krylov solvers (invalid names are GMRES)
    switch (matrix_solver) {
        case ML_SOLVER_LU: case ML_SOLVER_BJAC: case ML_SOLVER_BSSOR: case ML_SOLVER_QRUP: case ML_SOLVER_FQRUP: case ML_SOLVER_PURC:
            return true;
        default: return false;
    }
}

// Common tail: dispatch.  d_scale points at 1/A(N,N) on the device when the "DIAG" preconditioner is
// selected, else nullptr.  The direct and block solvers are scale-invariant (implicit row scaling), so
// they ignore it.
static ml_status run_solver(Sys& S, const ml_solver_opts* opts, const double* d_b, const double* d_scale, double* d_x,
                            ml_solve_info* info, double* lu_matrix /* full square copy or nullptr */, int lu_ld) {
    Ctx* c = S.c;
    int iters = -1;
    ml_status st = ML_OK;
    const bool use_mgs = std::getenv("MACHLINE_GMRES_MGS") != nullptr;
    int block_size = opts->block_size;
    if (block_size <= 0) block_size = S.N / 5;  // panel_solver.f90:1910-1912
    double err_scale = 1.;
    const int ms = opts->matrix_solver;
    if (d_scale && (ms == ML_SOLVER_BJAC || ms == ML_SOLVER_BSSOR)) {
        cudaError_t e = cudaMemcpyAsync(&err_scale, d_scale, sizeof(double), cudaMemcpyDeviceToHost, c->stream);
        if (e == cudaSuccess) e = cudaStreamSynchronize(c->stream);
        if (e != cudaSuccess) return c->cuda_fail(e, "read 1/A(N,N)");
        err_scale = std::fabs(err_scale);
    }
    switch (opts->matrix_solver) {
        case ML_SOLVER_LU:
            if (!lu_matrix) return c->fail(ML_UNSUPPORTED, "LU needs the full matrix on one device");
            st = lu_solve_device(c, S.N, lu_matrix, lu_ld, d_b, d_x);
            break;
        case ML_SOLVER_BJAC:
            if (S.sharded()) {   // rows stay where they were assembled (lu_kernels.cu: block_jacobi_sharded)
                RowShardOps R{S.A, S.ld, S.n_rows, S.N, c->d_g_of_slot.p + (size_t)c->rank * S.shard_pad,
                              S.gather.p + (size_t)c->rank * S.shard_pad, &S,
                              [](void* sys, double* y_full) { return static_cast<Sys*>(sys)->exchange(y_full, nullptr); }};
                st = block_jacobi_sharded(c, R, d_b, block_size, opts->tol, opts->rel, opts->max_iterations, &iters, d_x, err_scale,
                                          opts->iteration_file);
                break;
            }
            if (!lu_matrix) return c->fail(ML_UNSUPPORTED, "BJAC needs the full matrix on one device");
            st = block_jacobi_device(c, S.N, lu_matrix, lu_ld, d_b, block_size, opts->tol, opts->rel, opts->max_iterations, &iters, d_x,
                                     err_scale, opts->iteration_file);
            break;
        case ML_SOLVER_BSSOR:
            if (!lu_matrix) return c->fail(ML_UNSUPPORTED, "BSSOR needs the full matrix on one device");
            st = block_ssor_device(c, S.N, lu_matrix, lu_ld, d_b, block_size, opts->tol, opts->rel, opts->max_iterations, &iters, d_x,
                                   err_scale, opts->iteration_file);
            break;
        case ML_SOLVER_QRUP:
        case ML_SOLVER_FQRUP:
            if (!lu_matrix) return c->fail(ML_UNSUPPORTED, "QRUP/FQRUP need the full matrix on one device");
            st = qrup_solve_device(c, S.N, lu_matrix, lu_ld, d_b, d_scale, ms == ML_SOLVER_FQRUP, d_x);
            break;
        case ML_SOLVER_PURC:
            if (!lu_matrix) return c->fail(ML_UNSUPPORTED, "PURC needs the full matrix on one device");
            st = purcell_solve_device(c, S.N, lu_matrix, lu_ld, d_b, d_scale, d_x);
            break;
        case ML_SOLVER_RGMRES:
        case ML_SOLVER_GMRES:
        default: {  // invalid names fall back to GMRES (panel_solver.f90:1969-1973)
            const bool restarted = opts->matrix_solver == ML_SOLVER_RGMRES;
            const int k_max = restarted ? std::min(opts->restart_iterations, S.N) : std::min(S.N, opts->max_iterations);
            // row-sharded basis + peer-memory reductions when the windows are mapped; else the replicated basis (NCCL all-gather)
            const bool shard_basis = S.sharded() && c->p2p_ok && !use_mgs && k_max + 3 <= Ctx::P2P_KR && S.n_rows <= c->num_sms * 32 * SHT_MAXCH &&
                                     std::getenv("MACHLINE_GMRES_REPLICATED") == nullptr;
            if (shard_basis)
                st = gmres_sharded_device(S, d_b, d_scale, opts->tol, opts->max_iterations, opts->restart_iterations, restarted, d_x, &iters,
                                          opts->iteration_file);
            else
                st = gmres_device(S, d_b, d_scale, opts->tol, opts->max_iterations, opts->restart_iterations, restarted, use_mgs, d_x, &iters,
                                  opts->iteration_file);
            break;
        }
    }
    if (st != ML_OK) return st;
    if (info) info->iterations = iters;
    return ML_OK;
}

// R = A x - b: max and 2-norm (panel_solver.f90:1992-1998), with the unscaled A
static ml_status residual(Sys& S, const double* d_x, const double* d_b, ml_solve_info* info) {
    Ctx* c = S.c;
    DevBuf<double> Ax;
    ML_CUDA(c, Ax.alloc(S.N));
    ml_status st = S.matvec(d_x, Ax.p, 1.0, nullptr);
    if (st != ML_OK) { Ax.release(); return st; }
    std::vector<double> hAx(S.N), hb(S.N);
    ML_CUDA(c, cudaMemcpyAsync(hAx.data(), Ax.p, (size_t)S.N * sizeof(double), cudaMemcpyDeviceToHost, c->stream));
    ML_CUDA(c, cudaMemcpyAsync(hb.data(), d_b, (size_t)S.N * sizeof(double), cudaMemcpyDeviceToHost, c->stream));
    ML_CUDA(c, cudaStreamSynchronize(c->stream));
    c->d2h_bytes += (long long)2 * S.N * sizeof(double);
    Ax.release();
    double mx = 0., ss = 0.;
    for (int i = 0; i < S.N; ++i) {
        double r = hAx[i] - hb[i];
        mx = std::max(mx, std::fabs(r));
        ss += r * r;
    }
    if (info) {
        info->res_max = mx;
        info->res_norm = std::sqrt(ss);
    }
    if (!(ss == ss)) return ML_NAN_RESIDUAL;
    return ML_OK;
}

// ---- overdetermined least squares (the Neumann formulations; panel_solver.f90:1842-1895) ---------------------------------------
// C = A^T A for A (m x n, column-major, lda): CTA tile 64 x 64, 256 threads x (4 x 4) outputs, the rows of A walked in chunks of
// 16 in ascending order (every C entry is summed in row order, as the reference's matmul(transpose(A), A) does).
__global__ void __launch_bounds__(256) ata_kernel(const double* __restrict__ A, int lda, int m, int n, double* __restrict__ C, int ldc) {
    __shared__ double sI[16][64 + 1], sJ[16][64 + 1];
    const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;
    const int i0 = blockIdx.x * 64, j0 = blockIdx.y * 64;
    double acc[4][4] = {};
    for (int r0 = 0; r0 < m; r0 += 16) {
        for (int t = threadIdx.x; t < 16 * 64; t += 256) {
            const int rr = t & 15, cc = t >> 4;
            const int r = r0 + rr;
            sI[rr][cc] = (r < m && i0 + cc < n) ? A[r + (size_t)(i0 + cc) * lda] : 0.;
            sJ[rr][cc] = (r < m && j0 + cc < n) ? A[r + (size_t)(j0 + cc) * lda] : 0.;
        }
        __syncthreads();
#pragma unroll
        for (int rr = 0; rr < 16; ++rr) {
            double a[4], b[4];
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                a[u] = sI[rr][tx + 16 * u];
                b[u] = sJ[rr][ty + 16 * u];
            }
#pragma unroll
            for (int u = 0; u < 4; ++u)
#pragma unroll
                for (int v = 0; v < 4; ++v) acc[u][v] = fma(a[u], b[v], acc[u][v]);
        }
        __syncthreads();
    }
#pragma unroll
    for (int u = 0; u < 4; ++u)
#pragma unroll
        for (int v = 0; v < 4; ++v) {
            const int i = i0 + tx + 16 * u, j = j0 + ty + 16 * v;
            if (i < n && j < n) C[i + (size_t)j * ldc] = acc[u][v];
        }
}

// x minimises ||A x - b|| through the normal equations A^T A x = A^T b with the input's preconditioner and solver; the residual
// reported is that of the original system (panel_solver.f90:1992-1998).  Single GPU.
// The solution stays on the device for ml_post_process (post.cu): a device-to-device copy of N doubles on the solve's stream.
// The buffer outlives the solve, so it does not come from the stream-ordered pool of the solver temporaries.
static cudaError_t keep_solution(Ctx* c, const double* d_x, int N) {
    c->n_x_last = 0;
    cudaStream_t pool = tl_pool_stream;
    tl_pool_stream = nullptr;
    cudaError_t e = c->d_x_last.alloc((size_t)N);
    tl_pool_stream = pool;
    if (e != cudaSuccess) return e;
    e = cudaMemcpyAsync(c->d_x_last.p, d_x, (size_t)N * sizeof(double), cudaMemcpyDeviceToDevice, c->stream);
    if (e == cudaSuccess) c->n_x_last = N;
    return e;
}

static ml_status solve_least_squares(Ctx* c, const ml_solver_opts* opts, const double* BC, double* x_out, ml_solve_info* info) {
    const int N = c->n_cols, M = c->n_cp;
    if (c->world > 1 || c->n_rows != M) return c->fail(ML_UNSUPPORTED, "least-squares formulations run on one GPU (row shards are not built for them)");
    if (M < N) return c->fail(ML_UNSUPPORTED, "underdetermined least squares (neumann-doublet-source-mass-flux-ls) is not built");
    ML_CUDA(c, cudaEventRecord(c->ev0, c->stream));
    std::vector<double> b(M);
    for (int i = 0; i < M; ++i) b[i] = BC[i] - c->h_I_known[i];
    for (int i = 0; i < M; ++i)
        if (!(b[i] == b[i])) return c->fail(ML_NAN_IN_SYSTEM, "NaN in b");
    const int ldn = ((N + 63) / 64) * 64;
    DevBuf<double> d_b, AtA, Atb, d_x, d_scale, Acopy, Ax;
    ML_CUDA(c, d_b.alloc(c->n_rows_pad));
    ML_CUDA(c, AtA.alloc((size_t)ldn * N));
    ML_CUDA(c, Atb.alloc(ldn));
    ML_CUDA(c, d_x.alloc(ldn));
    ML_CUDA(c, d_scale.alloc(2));
    ML_CUDA(c, cudaMemsetAsync(d_b.p, 0, (size_t)c->n_rows_pad * sizeof(double), c->stream));
    ML_CUDA(c, cudaMemsetAsync(AtA.p, 0, (size_t)ldn * N * sizeof(double), c->stream));
    ML_CUDA(c, cudaMemcpyAsync(d_b.p, b.data(), (size_t)M * sizeof(double), cudaMemcpyHostToDevice, c->stream));
    c->h2d_bytes += (long long)M * sizeof(double);
    ata_kernel<<<dim3((N + 63) / 64, (N + 63) / 64), 256, 0, c->stream>>>(c->d_A.p, c->ld, M, N, AtA.p, ldn);
    gemv_t_kernel<<<N, 256, 0, c->stream>>>(c->d_A.p, c->ld, M, d_b.p, Atb.p);
    c->launches += 2;
    Sys S{};
    S.c = c;
    S.A = AtA.p;
    S.ld = ldn;
    S.n_rows = N;
    S.n_rows_pad = ldn;
    S.N = N;
    S.shard_pad = ldn;
    ml_status st = S.init();
    if (st != ML_OK) return st;
    const double* scale_ptr = nullptr;
    if (opts->preconditioner == ML_PREC_DIAG) {   // diagonal_preconditioner on A^T A: 1 / (A^T A)(N, N)
        recip_kernel<<<1, 1, 0, c->stream>>>(AtA.p + (N - 1) + (size_t)(N - 1) * ldn, d_scale.p);
        c->launches += 1;
        scale_ptr = d_scale.p;
    }
    double* lu_matrix = nullptr;
    if (needs_whole_matrix(opts->matrix_solver)) {
        lu_matrix = AtA.p;
        if (opts->matrix_solver == ML_SOLVER_LU) {
            ML_CUDA(c, Acopy.alloc((size_t)ldn * N));
            ML_CUDA(c, cudaMemcpyAsync(Acopy.p, AtA.p, (size_t)ldn * N * sizeof(double), cudaMemcpyDeviceToDevice, c->stream));
            lu_matrix = Acopy.p;
        }
    }
    st = run_solver(S, opts, Atb.p, scale_ptr, d_x.p, info, lu_matrix, ldn);
    S.release();
    if (st != ML_OK) return st;
    // R = A x - b with the original (rectangular) system
    Sys R{};
    R.c = c;
    R.A = c->d_A.p;
    R.ld = c->ld;
    R.n_rows = M;
    R.n_rows_pad = c->n_rows_pad;
    R.N = N;
    R.shard_pad = c->n_rows_pad;
    st = R.init();
    if (st != ML_OK) return st;
    ML_CUDA(c, Ax.alloc(c->n_rows_pad));
    st = R.gemv_local(d_x.p, Ax.p, 1.0, nullptr);
    if (st != ML_OK) return st;
    std::vector<double> hAx(M);
    ML_CUDA(c, cudaMemcpyAsync(hAx.data(), Ax.p, (size_t)M * sizeof(double), cudaMemcpyDeviceToHost, c->stream));
    ML_CUDA(c, cudaMemcpyAsync(x_out, d_x.p, (size_t)N * sizeof(double), cudaMemcpyDeviceToHost, c->stream));
    ML_CUDA(c, keep_solution(c, d_x.p, N));
    ML_CUDA(c, cudaEventRecord(c->ev1, c->stream));
    ML_CUDA(c, cudaStreamSynchronize(c->stream));
    c->d2h_bytes += (long long)(M + N) * sizeof(double);
    R.release();
    double mx = 0., ss = 0.;
    for (int i = 0; i < M; ++i) {
        const double r = hAx[i] - b[i];
        mx = std::max(mx, std::fabs(r));
        ss += r * r;
    }
    float ms = 0.f;
    cudaEventElapsedTime(&ms, c->ev0, c->ev1);
    c->solve_ms = ms;
    if (info) {
        info->res_max = mx;
        info->res_norm = std::sqrt(ss);
        info->assemble_ms = c->assemble_ms;
        info->solve_ms = ms;
    }
    if (!(ss == ss)) return ML_NAN_RESIDUAL;
    return ML_OK;
}

// Solve with the matrix the assembly left resident (possibly row-sharded).
ml_status solve_resident(Ctx* c, const ml_solver_opts* opts, const double* BC, double* x_out, ml_solve_info* info) {
    const int N = c->n_cols;
    if (c->n_cp != N) return solve_least_squares(c, opts, BC, x_out, info);
    ML_CUDA(c, cudaEventRecord(c->ev0, c->stream));
    Sys S{};
    S.c = c;
    S.A = c->d_A.p;
    S.ld = c->ld;
    S.n_rows = c->n_rows;
    S.n_rows_pad = c->n_rows_pad;
    S.N = N;
    {   // one rank running the multi-rank Krylov code paths (peer-memory exchange, sharded basis): the single-GPU test box
        const int ms = opts->matrix_solver;
        const bool krylov = !(ms == ML_SOLVER_LU || ms == ML_SOLVER_BJAC || ms == ML_SOLVER_BSSOR || ms == ML_SOLVER_QRUP || ms == ML_SOLVER_FQRUP ||
                              ms == ML_SOLVER_PURC);
        S.force_shard = c->world == 1 && ((krylov && std::getenv("MACHLINE_GMRES_SHARDED") != nullptr) ||
                                          (ms == ML_SOLVER_BJAC && std::getenv("MACHLINE_BJAC_SHARDED") != nullptr));
    }
    // Slot tables of the all-gather layout: slot = rank * shard_pad + local row <-> global row.  Every rank contributes the
    // list of rows it assembled, so any dealing of rows to ranks (contiguous blocks, block-cyclic) works the same way.
    {
        c->shard_nrows.assign(c->world, c->n_rows);
        int S_pad = c->n_rows_pad;
        if (c->world > 1) {
#ifdef ML_HAVE_NCCL
            std::vector<int> all(c->world);
            DevBuf<int> dm, da;
            ML_CUDA(c, dm.alloc(1));
            ML_CUDA(c, da.alloc(c->world));
            ML_CUDA(c, cudaMemcpyAsync(dm.p, &c->n_rows, sizeof(int), cudaMemcpyHostToDevice, c->stream));
            if (ncclAllGather(dm.p, da.p, 1, ncclInt, c->comm, c->stream) != ncclSuccess) return c->fail(ML_NCCL_ERROR, "allgather shard sizes");
            ML_CUDA(c, cudaMemcpyAsync(all.data(), da.p, c->world * sizeof(int), cudaMemcpyDeviceToHost, c->stream));
            ML_CUDA(c, cudaStreamSynchronize(c->stream));
            dm.release();
            da.release();
            int mxr = 0;
            for (int r = 0; r < c->world; ++r) {
                c->shard_nrows[r] = all[r];
                mxr = std::max(mxr, all[r]);
            }
            S_pad = ((mxr + 63) / 64) * 64;
#else
            return c->fail(ML_UNSUPPORTED, "library built without NCCL");
#endif
        } else if (c->n_rows != N) {
            return c->fail(ML_BAD_ARGUMENT, "row shard set but no communicator joined");
        }
        c->shard_pad = S_pad;
        const size_t n_slots = (size_t)S_pad * c->world;
        std::vector<int> mine(S_pad, -1);
        for (int i = 0; i < c->n_rows; ++i) mine[i] = c->local_rows[i];
        c->g_of_slot.assign(n_slots, -1);
        ML_CUDA(c, c->d_g_of_slot.alloc(n_slots));
        ML_CUDA(c, c->d_slot_of_g.alloc(N));
        ML_CUDA(c, cudaMemcpyAsync(c->d_g_of_slot.p + (size_t)c->rank * S_pad, mine.data(), (size_t)S_pad * sizeof(int), cudaMemcpyHostToDevice,
                                   c->stream));
#ifdef ML_HAVE_NCCL
        if (c->world > 1 &&
            ncclAllGather(c->d_g_of_slot.p + (size_t)c->rank * S_pad, c->d_g_of_slot.p, S_pad, ncclInt, c->comm, c->stream) != ncclSuccess)
            return c->fail(ML_NCCL_ERROR, "allgather row lists");
#endif
        ML_CUDA(c, cudaMemcpyAsync(c->g_of_slot.data(), c->d_g_of_slot.p, n_slots * sizeof(int), cudaMemcpyDeviceToHost, c->stream));
        ML_CUDA(c, cudaStreamSynchronize(c->stream));
        c->slot_of_g.assign(N, -1);
        for (size_t sl = 0; sl < n_slots; ++sl) {
            const int g = c->g_of_slot[sl];
            if (g < 0) continue;
            if (g >= N || c->slot_of_g[g] != -1) return c->fail(ML_BAD_ARGUMENT, "row shards overlap or exceed the system");
            c->slot_of_g[g] = (int)sl;
        }
        for (int g = 0; g < N; ++g)
            if (c->slot_of_g[g] < 0) return c->fail(ML_BAD_ARGUMENT, "row shards do not cover the system");
        ML_CUDA(c, cudaMemcpyAsync(c->d_slot_of_g.p, c->slot_of_g.data(), (size_t)N * sizeof(int), cudaMemcpyHostToDevice, c->stream));
        S.shard_pad = S_pad;
    }
    ml_status st = S.init();
    if (st != ML_OK) return st;

    // b = BC - I_known (panel_solver.f90:1818); I_known lives sharded -> build the full b on the host
    std::vector<double> b(N);
    std::vector<double> Ik_full(N, 0.);
    if (c->world > 1) {
#ifdef ML_HAVE_NCCL
        DevBuf<double> g;
        ML_CUDA(c, g.alloc((size_t)S.shard_pad * c->world));
        ML_CUDA(c, cudaMemcpyAsync(g.p + (size_t)c->rank * S.shard_pad, c->d_I_known.p, (size_t)c->n_rows * sizeof(double),
                                   cudaMemcpyDeviceToDevice, c->stream));
        if (ncclAllGather(g.p + (size_t)c->rank * S.shard_pad, g.p, S.shard_pad, ncclDouble, c->comm, c->stream) != ncclSuccess)
            return c->fail(ML_NCCL_ERROR, "allgather I_known");
        std::vector<double> tmp((size_t)S.shard_pad * c->world);
        ML_CUDA(c, cudaMemcpyAsync(tmp.data(), g.p, tmp.size() * sizeof(double), cudaMemcpyDeviceToHost, c->stream));
        ML_CUDA(c, cudaStreamSynchronize(c->stream));
        g.release();
        for (size_t sl = 0; sl < tmp.size(); ++sl)
            if (c->g_of_slot[sl] >= 0) Ik_full[c->g_of_slot[sl]] = tmp[sl];
#endif
    } else {
        for (int i = 0; i < N; ++i) Ik_full[c->local_rows[i]] = c->h_I_known[i];
    }
    for (int i = 0; i < N; ++i) b[i] = BC[i] - Ik_full[i];
    for (int i = 0; i < N; ++i)
        if (!(b[i] == b[i])) return c->fail(ML_NAN_IN_SYSTEM, "NaN in b");

    DevBuf<double> d_b, d_x, d_scale;
    ML_CUDA(c, d_b.alloc(N));
    ML_CUDA(c, d_x.alloc(N));
    ML_CUDA(c, d_scale.alloc(2));
    ML_CUDA(c, cudaMemcpyAsync(d_b.p, b.data(), (size_t)N * sizeof(double), cudaMemcpyHostToDevice, c->stream));
    c->h2d_bytes += (long long)N * sizeof(double);
    const double* scale_ptr = nullptr;
    if (opts->preconditioner == ML_PREC_DIAG) {
        // linalg.f90:1813-1816: A_ii_inv(:) = 1/A(N,N).  Row N-1 lives on the rank that owns it.
        const int last_slot = c->slot_of_g[N - 1];
        const int owner = last_slot / S.shard_pad, last_lr = last_slot % S.shard_pad;
        if (owner == c->rank) {
            recip_kernel<<<1, 1, 0, c->stream>>>(c->d_A.p + last_lr + (size_t)(N - 1) * c->ld, d_scale.p);
            c->launches += 1;
        }
#ifdef ML_HAVE_NCCL
        if (c->world > 1) {
            if (ncclBroadcast(d_scale.p, d_scale.p, 1, ncclDouble, owner, c->comm, c->stream) != ncclSuccess)
                return c->fail(ML_NCCL_ERROR, "broadcast scale");
        }
#endif
        scale_ptr = d_scale.p;
    }

    // Direct / block solvers work on a private full copy (the reference's A_p), single device only.
    DevBuf<double> Acopy;
    double* lu_matrix = nullptr;
    static const bool force_sharded_lu = std::getenv("MACHLINE_LU_SHARDED") != nullptr;   // tests: the NCCL algorithm on one rank
    if (opts->matrix_solver == ML_SOLVER_LU && (c->world > 1 || force_sharded_lu)) {
        // row-sharded LU: the local rows (plus the right-hand side as column N) are factored in a scratch copy (A_p)
        ML_CUDA(c, Acopy.alloc((size_t)c->ld * (N + 1)));
        ML_CUDA(c, cudaMemcpyAsync(Acopy.p, c->d_A.p, (size_t)c->ld * N * sizeof(double), cudaMemcpyDeviceToDevice, c->stream));
        st = lu_solve_sharded(c, N, Acopy.p, c->ld, c->n_rows, c->n_rows_pad, S.shard_pad, d_b.p, d_x.p);
        Acopy.release();
        if (st == ML_OK && info) info->iterations = -1;
    } else if (opts->matrix_solver == ML_SOLVER_BJAC && S.sharded()) {
        // block Jacobi on the row shards: no full copy (run_solver)
    } else if (needs_whole_matrix(opts->matrix_solver)) {
        if (c->world > 1)
            return c->fail(ML_UNSUPPORTED, "BSSOR/QRUP/FQRUP/PURC are sequential sweeps (\"replicas only\", SURVEY 8(e)): on a row-sharded system use LU/GMRES/RGMRES/BJAC");
        lu_matrix = c->d_A.p;
        if (opts->matrix_solver == ML_SOLVER_LU) {   // factored in place: work on the reference's A_p copy
            ML_CUDA(c, Acopy.alloc((size_t)c->ld * N));
            ML_CUDA(c, cudaMemcpyAsync(Acopy.p, c->d_A.p, (size_t)c->ld * N * sizeof(double), cudaMemcpyDeviceToDevice, c->stream));
            lu_matrix = Acopy.p;
        }
    }
    if (!(opts->matrix_solver == ML_SOLVER_LU && (c->world > 1 || force_sharded_lu)))
        st = run_solver(S, opts, d_b.p, scale_ptr, d_x.p, info, lu_matrix, c->ld);
    Acopy.release();
    if (st == ML_OK) st = residual(S, d_x.p, d_b.p, info);
    if (st == ML_OK || st == ML_NAN_RESIDUAL) {
        ML_CUDA(c, cudaMemcpyAsync(x_out, d_x.p, (size_t)N * sizeof(double), cudaMemcpyDeviceToHost, c->stream));
        c->d2h_bytes += (long long)N * sizeof(double);
    }
    if (st == ML_OK) ML_CUDA(c, keep_solution(c, d_x.p, N));
    ML_CUDA(c, cudaEventRecord(c->ev1, c->stream));
    ML_CUDA(c, cudaStreamSynchronize(c->stream));
    float ms = 0.f;
    cudaEventElapsedTime(&ms, c->ev0, c->ev1);
    c->solve_ms = ms;
    if (info) {
        info->assemble_ms = c->assemble_ms;
        info->solve_ms = ms;
    }
    d_b.release();
    d_x.release();
    d_scale.release();
    S.release();
    return st;
}

// Stand-alone dense solve of a device matrix (column-major, ld); the matrix itself is left intact.
ml_status solve_dense_device(Ctx* c, int N, double* dA, int ld, const double* h_b, const ml_solver_opts* opts, double* x_out,
                             ml_solve_info* info, bool A_is_scratch) {
    ML_CUDA(c, cudaEventRecord(c->ev0, c->stream));
    Sys S{};
    S.c = c;
    S.A = dA;
    S.ld = ld;
    S.n_rows = N;
    S.n_rows_pad = ld;
    S.N = N;
    S.shard_pad = ld;
    struct WorldGuard {   // a host-supplied dense system is never sharded; restored on every return path
        Ctx* c;
        int saved;
        explicit WorldGuard(Ctx* c_) : c(c_), saved(c_->world) { c->world = 1; }
        ~WorldGuard() { c->world = saved; }
    } world_guard(c);
    ml_status st = S.init();
    if (st != ML_OK) return st;
    DevBuf<double> d_b, d_x, d_scale, Acopy;
    ML_CUDA(c, d_b.alloc(N));
    ML_CUDA(c, d_x.alloc(N));
    ML_CUDA(c, d_scale.alloc(2));
    ML_CUDA(c, cudaMemcpyAsync(d_b.p, h_b, (size_t)N * sizeof(double), cudaMemcpyHostToDevice, c->stream));
    c->h2d_bytes += (long long)N * sizeof(double) + (long long)N * N * sizeof(double);
    const double* scale_ptr = nullptr;
    if (opts->preconditioner == ML_PREC_DIAG) {
        recip_kernel<<<1, 1, 0, c->stream>>>(dA + (N - 1) + (size_t)(N - 1) * ld, d_scale.p);
        c->launches += 1;
        scale_ptr = d_scale.p;
    }
    double* lu_matrix = nullptr;
    if (needs_whole_matrix(opts->matrix_solver)) {
        (void)A_is_scratch;  // the residual below needs the original matrix: LU factors a copy, the others only read it
        lu_matrix = dA;
        if (opts->matrix_solver == ML_SOLVER_LU) {
            ML_CUDA(c, Acopy.alloc((size_t)ld * N));
            ML_CUDA(c, cudaMemcpyAsync(Acopy.p, dA, (size_t)ld * N * sizeof(double), cudaMemcpyDeviceToDevice, c->stream));
            lu_matrix = Acopy.p;
        }
    }
    st = run_solver(S, opts, d_b.p, scale_ptr, d_x.p, info, lu_matrix, ld);
    Acopy.release();
    if (st == ML_OK) st = residual(S, d_x.p, d_b.p, info);
    if (st == ML_OK || st == ML_NAN_RESIDUAL) {
        ML_CUDA(c, cudaMemcpyAsync(x_out, d_x.p, (size_t)N * sizeof(double), cudaMemcpyDeviceToHost, c->stream));
        c->d2h_bytes += (long long)N * sizeof(double);
    }
    ML_CUDA(c, cudaEventRecord(c->ev1, c->stream));
    ML_CUDA(c, cudaStreamSynchronize(c->stream));
    float ms = 0.f;
    cudaEventElapsedTime(&ms, c->ev0, c->ev1);
    c->solve_ms = ms;
    if (info) {
        info->assemble_ms = 0.;
        info->solve_ms = ms;
    }
    d_b.release();
    d_x.release();
    d_scale.release();
    S.release();
    return st;
}

}  // namespace mlgpu
