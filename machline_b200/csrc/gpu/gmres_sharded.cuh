// Row-sharded GMRES iteration for N GPUs (included by solve_kernels.cu): the Krylov basis is sharded by rows like A, the
// Gram-Schmidt reductions and the exchange of the new basis vector go over PEER MEMORY (CUDA-IPC windows, NVLink /
// NVSwitch stores) from inside ONE cooperative kernel per iteration -- no NCCL call, no replicated work.
//
// Reference: arnoldi_update / GMRES, common/linalg.f90:1208-1334 (modified Gram-Schmidt there; classical Gram-Schmidt with
// one re-orthogonalisation here, as on one GPU: same Krylov space and Hessenberg matrix to rounding).
//
// Per rank: Qloc = the local rows of the basis, w_loc = the local rows of A q_k (gemv_n_partial_kernel on the local rows).
// arnoldi_tail_sharded_kernel, grid = one CTA per SM (cooperative launch: all CTAs resident), every CTA owns a fixed block
// of local rows for the whole kernel, so no grid barrier is needed between a subtraction and the next dot products:
//   D1  partial dots of the CTA's rows against Qloc(:, 0..k-1)                    -> ticket
//   R1  LAST CTA to arrive: sums the partials in CTA order, stores the k numbers into every rank's window (peer stores),
//       raises its flag there (st.release.sys), waits for the P flags of its own window (ld.acquire.sys), adds the P
//       contributions in rank order -> h1 (bitwise identical on every rank), publishes (st.release.gpu)
//   S1  every CTA: w_loc -= Qloc h1 on its rows;  D2 second-pass dots of the same rows   -> ticket
//   R2  as R1 -> h2, hfin = h1 + h2
//   S2  w_loc -= Qloc h2, sum of squares of the CTA's rows, and the rows are stored into every rank's vector window at
//       their GLOBAL index (the all-gather of the next Krylov vector, fused)        -> ticket
//   R3  LAST CTA: sum of squares in CTA order -> exchanged like h -> norm = sqrt(sum over ranks), hfin[k] = norm
//   F   Qloc(:, k) = w_loc / norm on the CTA's rows; x_full = window / norm (the operand of the next local matvec)
// Three grid-wide waits and three NVLink round trips per iteration; spin loops give up after ~4 s and raise an error flag
// (a rank that died must not hang the others).
#pragma once

namespace mlgpu {

constexpr int SHT_THREADS = 1024;
constexpr int SHT_WARPS = SHT_THREADS / 32;
constexpr long long SHT_SPIN_LIMIT = 8000000000LL;   // clock64 ticks (~4 s)

struct ShTailArgs {
    const double* Q;        // local rows of the basis, column-major, leading dimension ldq
    int ldq, n_loc, k;      // k = basis vectors to orthogonalise against
    double* w;              // [ldq] local rows of A q (in), orthogonalised (out)
    double* partial;        // [grid][kpad]
    int kpad;
    double* h1;             // [k] coefficients of the pass being applied (global scratch)
    double* hfin;           // [k + 2]: h1 + h2, the norm, and the error flag (as a double) for the host
    double* npart;          // [grid]
    double* qnext;          // Qloc(:, k)
    double* xfull;          // [N] next Krylov vector, replicated (operand of the next matvec)
    int N;
    const int* g_of_local;  // global row of each local row
    int P, rank;
    // windows of every rank (peer memory): reduction slots [2 parities][P2P_MAX][kr], vector [2 parities][nv], flags [2][P2P_MAX]
    double* red[Ctx::P2P_MAX];
    double* vec[Ctx::P2P_MAX];
    unsigned* flags[Ctx::P2P_MAX];
    int kr, nv;
    unsigned seq;           // number of the FIRST exchange of this launch (three per launch: seq, seq + 1, seq + 2); parity = number & 1
    unsigned* ticket;       // grid arrival counter (monotonic over the launches of one solve)
    unsigned* ready;        // published stage counter (monotonic)
    unsigned base;          // launches of this solve before this one
    int* err;               // raised when a spin loop gives up
    int rows_per_cta;       // multiple of 32
};

__device__ __forceinline__ unsigned ld_acquire_sys(const unsigned* p) {
    unsigned v;
    asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ unsigned ld_acquire_gpu(const unsigned* p) {
    unsigned v;
    asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_release_sys(unsigned* p, unsigned v) {
    asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ void st_release_gpu(unsigned* p, unsigned v) {
    asm volatile("st.release.gpu.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
// peer-written data: never through L1
__device__ __forceinline__ double ld_peer(const double* p) {
    double v;
    asm volatile("ld.relaxed.sys.global.f64 %0, [%1];" : "=d"(v) : "l"(p) : "memory");
    return v;
}

// Partial dots of this CTA's rows: partial[blockIdx.x][j] = sum_{i in rows} Q[i, j] * w[i], fixed order.
__device__ __forceinline__ void sht_dot_phase(const ShTailArgs& a, int row0, int nrows) {
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    double* out = a.partial + (size_t)blockIdx.x * a.kpad;
    for (int j = warp; j < a.k; j += SHT_WARPS) {
        const double* q = a.Q + (size_t)j * a.ldq + row0;
        double acc = 0.;
        for (int i = lane; i < nrows; i += 32) acc = fma(q[i], __ldcg(a.w + row0 + i), acc);
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
        if (lane == 0) __stcg(out + j, acc);
    }
}

// w[rows] -= Q[rows, 0..k-1] h; returns the sum of squares of the new w over the CTA's rows (thread 0), fixed order.
__device__ __forceinline__ double sht_sub_phase(const ShTailArgs& a, int row0, int nrows, const double* s_h, double* s_acc,
                                                bool want_norm) {
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    double sq_total = 0.;
    for (int r0 = 0; r0 < nrows; r0 += 32) {              // 32 rows at a time, the 32 warps split the columns
        const int i = r0 + lane;
        double acc = 0.;
        if (i < nrows) {
            const double* q = a.Q + row0 + i;
            double a0 = 0., a1 = 0.;
            int j = warp;
            for (; j + SHT_WARPS < a.k; j += 2 * SHT_WARPS) {
                a0 = fma(q[(size_t)j * a.ldq], s_h[j], a0);
                a1 = fma(q[(size_t)(j + SHT_WARPS) * a.ldq], s_h[j + SHT_WARPS], a1);
            }
            if (j < a.k) a0 = fma(q[(size_t)j * a.ldq], s_h[j], a0);
            acc = a0 + a1;
        }
        s_acc[warp * 32 + lane] = acc;
        __syncthreads();
        if (warp == 0) {
            double s = 0.;
#pragma unroll 8
            for (int c = 0; c < SHT_WARPS; ++c) s += s_acc[c * 32 + lane];
            double v = 0.;
            if (i < nrows) {
                v = __ldcg(a.w + row0 + i) - s;
                __stcg(a.w + row0 + i, v);
            }
            if (want_norm) {
                double sq = v * v;
#pragma unroll
                for (int o = 16; o > 0; o >>= 1) sq += __shfl_xor_sync(0xffffffffu, sq, o);
                sq_total += sq;
            }
        }
        __syncthreads();
    }
    return sq_total;   // meaningful in thread 0
}

// The CTA that arrives last at stage `stage` (0..2) of this launch returns true (in every thread); the others wait until the
// last one has published the stage.
__device__ __forceinline__ bool sht_arrive_is_last(const ShTailArgs& a, int stage, int* s_flag) {
    __syncthreads();
    if (threadIdx.x == 0) {
        __threadfence();
        const unsigned t = atomicAdd(a.ticket, 1u);
        const unsigned target = (a.base * 3u + (unsigned)stage + 1u) * gridDim.x - 1u;
        *s_flag = (t == target);
    }
    __syncthreads();
    return *s_flag != 0;
}
__device__ __forceinline__ void sht_wait_ready(const ShTailArgs& a, int stage) {
    if (threadIdx.x == 0) {
        const unsigned target = a.base * 3u + (unsigned)stage + 1u;
        const long long t0 = clock64();
        while ((int)(ld_acquire_gpu(a.ready) - target) < 0) {
            if (clock64() - t0 > SHT_SPIN_LIMIT) {
                *a.err = 1;
                break;
            }
        }
    }
    __syncthreads();
}
__device__ __forceinline__ void sht_publish(const ShTailArgs& a, int stage) {
    __threadfence();
    __syncthreads();
    if (threadIdx.x == 0) st_release_gpu(a.ready, a.base * 3u + (unsigned)stage + 1u);
}

// LAST CTA: s_vals[0..n) (shared memory, this rank's contribution) -> the reduction slots of exchange `e` in every rank's
// window; flags; wait for the P flags of the own window; sum over ranks in rank order -> s_out[0..n) (shared memory).
__device__ __forceinline__ void sht_exchange(const ShTailArgs& a, unsigned e, const double* s_vals, int n, double* s_out) {
    const size_t roff = (size_t)(e & 1u) * Ctx::P2P_MAX * a.kr, foff = (size_t)(e & 1u) * Ctx::P2P_MAX;
    for (int p = 0; p < a.P; ++p) {
        double* dst = a.red[p] + roff + (size_t)a.rank * a.kr;
        for (int j = threadIdx.x; j < n; j += SHT_THREADS) dst[j] = s_vals[j];
    }
    __threadfence_system();
    __syncthreads();
    if ((int)threadIdx.x < a.P) {
        st_release_sys(a.flags[threadIdx.x] + foff + a.rank, e);
        const unsigned* f = a.flags[a.rank] + foff + threadIdx.x;
        const long long t0 = clock64();
        while ((int)(ld_acquire_sys(f) - e) < 0) {
            if (clock64() - t0 > SHT_SPIN_LIMIT) {
                *a.err = 1;
                break;
            }
        }
    }
    __syncthreads();
    const double* mine = a.red[a.rank] + roff;
    for (int j = threadIdx.x; j < n; j += SHT_THREADS) {
        double s = 0.;
        for (int r = 0; r < a.P; ++r) s += ld_peer(mine + (size_t)r * a.kr + j);
        s_out[j] = s;
    }
    __syncthreads();
}

__global__ void __launch_bounds__(SHT_THREADS, 1) arnoldi_tail_sharded_kernel(const ShTailArgs a) {
    extern __shared__ double s_mem[];
    const int kk = (a.k + 3) & ~1;
    double* s_h = s_mem;                 // [kk] coefficients being applied / this rank's contribution
    double* s_out = s_mem + kk;          // [kk] reduced over ranks
    double* s_acc = s_mem + 2 * kk;      // [1024]
    __shared__ int s_flag;
    __shared__ double s_norm;
    const int row0 = blockIdx.x * a.rows_per_cta;
    const int nrows = max(0, min(a.rows_per_cta, a.n_loc - row0));
    const unsigned G = gridDim.x;

    for (int pass = 0; pass < 2; ++pass) {
        // ---- D: partial dots of this CTA's rows ----
        sht_dot_phase(a, row0, nrows);
        if (sht_arrive_is_last(a, pass, &s_flag)) {
            // ---- R: this rank's coefficients = sum of the CTA partials in CTA order; exchange; sum over ranks ----
            for (int j = threadIdx.x; j < a.k; j += SHT_THREADS) {
                double s = 0.;
                for (unsigned c = 0; c < G; ++c) s += __ldcg(a.partial + (size_t)c * a.kpad + j);
                s_h[j] = s;
            }
            __syncthreads();
            sht_exchange(a, a.seq + (unsigned)pass, s_h, a.k, s_out);
            for (int j = threadIdx.x; j < a.k; j += SHT_THREADS) {
                const double h = s_out[j];
                if (pass == 1) __stcg(a.hfin + j, __ldcg(a.h1 + j) + h);
                __stcg(a.h1 + j, h);     // the coefficients the S phase of this pass applies
            }
            sht_publish(a, pass);
        } else {
            sht_wait_ready(a, pass);
        }
        // ---- S: subtract on this CTA's rows ----
        for (int j = threadIdx.x; j < a.k; j += SHT_THREADS) s_h[j] = __ldcg(a.h1 + j);
        __syncthreads();
        const double sq = sht_sub_phase(a, row0, nrows, s_h, s_acc, pass == 1);
        if (pass == 1) {
            if (threadIdx.x == 0) __stcg(a.npart + blockIdx.x, sq);
            // the all-gather of the next Krylov vector, fused: this CTA's rows go to every rank's window at their global index
            const size_t voff = (size_t)((a.seq + 2u) & 1u) * a.nv;
            for (int p = 0; p < a.P; ++p) {
                double* dst = a.vec[p] + voff;
                for (int i = threadIdx.x; i < nrows; i += SHT_THREADS) dst[a.g_of_local[row0 + i]] = __ldcg(a.w + row0 + i);
            }
            __threadfence_system();
        }
    }
    // ---- R3: the norm ----
    if (sht_arrive_is_last(a, 2, &s_flag)) {
        if (threadIdx.x < 32) {
            double t = 0.;
            for (unsigned c = threadIdx.x; c < G; c += 32) t += __ldcg(a.npart + c);
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) t += __shfl_xor_sync(0xffffffffu, t, o);
            if (threadIdx.x == 0) s_h[0] = t;
        }
        __syncthreads();
        sht_exchange(a, a.seq + 2u, s_h, 1, s_out);
        if (threadIdx.x == 0) {
            __stcg(a.hfin + a.k, sqrt(s_out[0]));
            __stcg(a.hfin + a.k + 1, (double)(*reinterpret_cast<volatile int*>(a.err)));
        }
        sht_publish(a, 2);
    } else {
        sht_wait_ready(a, 2);
    }
    if (threadIdx.x == 0) s_norm = __ldcg(a.hfin + a.k);
    __syncthreads();
    const double nrm = s_norm;
    // ---- F: the new basis vector (local rows) and the operand of the next matvec (all rows, from the window) ----
    for (int i = threadIdx.x; i < nrows; i += SHT_THREADS) a.qnext[row0 + i] = __ldcg(a.w + row0 + i) / nrm;
    const double* mine = a.vec[a.rank] + (size_t)((a.seq + 2u) & 1u) * a.nv;
    for (int g = blockIdx.x * SHT_THREADS + threadIdx.x; g < a.N; g += G * SHT_THREADS) a.xfull[g] = ld_peer(mine + g) / nrm;
}

}  // namespace mlgpu
