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
// Three grid-wide waits and three NVLink round trips per iteration; spin loops give up after ~20 s and raise an error flag
// (a rank that died must not hang the others).
#pragma once

namespace mlgpu {

constexpr int SHT_THREADS = 1024;
constexpr int SHT_WARPS = SHT_THREADS / 32;
constexpr long long SHT_SPIN_LIMIT = 40000000000LL;  // clock64 ticks (~20 s)
constexpr int SHT_MAXCH = 4;                          // a CTA owns at most 32 * SHT_MAXCH rows

struct ShTailArgs {
    const double* Q;        // local rows of the basis, column-major, leading dimension ldq
    int ldq, n_loc, k;      // k = basis vectors to orthogonalise against
    double* w;              // [ldq] local rows of A q (in), orthogonalised (out)
    double* partial;        // [k][gpad]: partial[j * gpad + cta], so that the reducing warp reads one column's partials coalesced
    int gpad;
    double* h1;             // [k] coefficients of the pass being applied (global scratch)
    double* hfin;           // [k + 2]: h1 + h2, the norm, and the error flag (as a double) for the host
    double* hhost;          // version 2: the same k + 2 numbers stored straight into the host's pinned slot (mapped), or nullptr
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
    unsigned stages;        // grid-wide sync points per launch (3: arnoldi_tail_sharded_kernel, 4: arnoldi_tail_sharded2_kernel)
    int* err;               // raised when a spin loop gives up
    int rows_per_cta;       // multiple of 32
    long long* dbg;         // nullptr, or [2][16] clock64 stamps (CTA 0 / the CTA that was last at stage 0) -- MACHLINE_SHT_DEBUG
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

// Partial dots of this CTA's rows: partial[j][blockIdx.x] = sum_{i in rows} Q[i, j] * w[i], fixed order.  A warp takes four
// columns per trip and issues all their loads before the first use (the phase is a chain of L2 round trips otherwise).
template <int NCOL, int NCH>
__device__ __forceinline__ void sht_dot_phase_t(const ShTailArgs& a, int row0, int nrows, const double* s_w) {
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    double* out = a.partial + blockIdx.x;
    double wv[NCH];
#pragma unroll
    for (int c = 0; c < NCH; ++c) wv[c] = (lane + 32 * c < nrows) ? s_w[lane + 32 * c] : 0.;
    for (int j0 = warp * NCOL; j0 < a.k; j0 += SHT_WARPS * NCOL) {
        double qv[NCOL][NCH];
#pragma unroll
        for (int u = 0; u < NCOL; ++u) {
            const double* q = a.Q + (size_t)min(j0 + u, a.k - 1) * a.ldq + row0 + lane;
#pragma unroll
            for (int c = 0; c < NCH; ++c) qv[u][c] = (lane + 32 * c < nrows) ? q[32 * c] : 0.;
        }
#pragma unroll
        for (int u = 0; u < NCOL; ++u) {
            double acc = 0.;
#pragma unroll
            for (int c = 0; c < NCH; ++c) acc = fma(qv[u][c], wv[c], acc);
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
            if (lane == 0 && j0 + u < a.k) __stcg(out + (size_t)(j0 + u) * a.gpad, acc);
        }
    }
}
__device__ __forceinline__ void sht_dot_phase(const ShTailArgs& a, int row0, int nrows, const double* s_w) {
    if (a.rows_per_cta <= 32) sht_dot_phase_t<8, 1>(a, row0, nrows, s_w);
    else if (a.rows_per_cta <= 64) sht_dot_phase_t<8, 2>(a, row0, nrows, s_w);
    else sht_dot_phase_t<4, SHT_MAXCH>(a, row0, nrows, s_w);
}

// w[rows] -= Q[rows, 0..k-1] h; returns the sum of squares of the new w over the CTA's rows (thread 0), fixed order.
// The 32 warps split the columns; a lane owns rows lane, lane + 32, ... (at most SHT_MAXCH of them: rows_per_cta <= 32 SHT_MAXCH),
// so all loads of a thread are independent and one shared-memory reduction over the warps finishes the block.
__device__ __forceinline__ double sht_sub_phase(const ShTailArgs& a, int row0, int nrows, const double* s_h, double* s_acc,
                                                double* s_w, bool want_norm) {
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    double acc[SHT_MAXCH];
#pragma unroll
    for (int c = 0; c < SHT_MAXCH; ++c) acc[c] = 0.;
    const double* q0 = a.Q + row0 + lane;
    for (int j = warp; j < a.k; j += SHT_WARPS) {
        const double hj = s_h[j];
        const double* q = q0 + (size_t)j * a.ldq;
#pragma unroll
        for (int c = 0; c < SHT_MAXCH; ++c)
            if (lane + 32 * c < nrows) acc[c] = fma(q[32 * c], hj, acc[c]);
    }
#pragma unroll
    for (int c = 0; c < SHT_MAXCH; ++c) s_acc[(c * SHT_WARPS + warp) * 32 + lane] = acc[c];
    __syncthreads();
    double sq = 0.;
    if (warp < SHT_MAXCH) {                 // warp c finishes row chunk c
        const int i = warp * 32 + lane;
        double s = 0.;
#pragma unroll 8
        for (int w2 = 0; w2 < SHT_WARPS; ++w2) s += s_acc[(warp * SHT_WARPS + w2) * 32 + lane];
        double v = 0.;
        if (i < nrows) {
            v = s_w[i] - s;
            s_w[i] = v;
            __stcg(a.w + row0 + i, v);
        }
        if (want_norm) {
            sq = v * v;
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) sq += __shfl_xor_sync(0xffffffffu, sq, o);
        }
    }
    __syncthreads();
    double sq_total = 0.;
    if (want_norm) {
        if (warp < SHT_MAXCH && lane == 0) s_acc[warp] = sq;
        __syncthreads();
        if (threadIdx.x == 0)
            for (int c = 0; c < SHT_MAXCH; ++c) sq_total += s_acc[c];
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
        const unsigned target = (a.base * a.stages + (unsigned)stage + 1u) * gridDim.x - 1u;
        *s_flag = (t == target);
    }
    __syncthreads();
    return *s_flag != 0;
}
__device__ __forceinline__ void sht_wait_ready(const ShTailArgs& a, int stage) {
    if (threadIdx.x == 0) {
        const unsigned target = a.base * a.stages + (unsigned)stage + 1u;
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
    __syncthreads();
    if (threadIdx.x == 0) st_release_gpu(a.ready, a.base * a.stages + (unsigned)stage + 1u);
}

// LAST CTA: s_vals[0..n) (shared memory, this rank's contribution) -> the reduction slots of exchange `e` in every rank's
// window; one system-scope fence (cumulative over the CTA's stores through the barrier before it); flags; wait for the P flags
// of the own window.  Afterwards the P contributions sit in this rank's window: sht_rank_sum adds them in rank order.
__device__ __forceinline__ void sht_exchange(const ShTailArgs& a, unsigned e, const double* s_vals, int n) {
    const size_t roff = (size_t)(e & 1u) * Ctx::P2P_MAX * a.kr, foff = (size_t)(e & 1u) * Ctx::P2P_MAX;
    for (int p = 0; p < a.P; ++p) {
        double* dst = a.red[p] + roff + (size_t)a.rank * a.kr;
        for (int j = threadIdx.x; j < n; j += SHT_THREADS) dst[j] = s_vals[j];
    }
    __syncthreads();
    // st.release.sys is cumulative over everything that happens-before it: the CTA's own stores through the barrier above, and
    // the other CTAs' stores to the peers' vector windows through their arrival at the ticket (fence + atomic, gpu scope)
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
}
// s_out[j] = sum over ranks (rank order) of the contributions of exchange e, j = 0..n-1: identical bits on every rank and CTA
__device__ __forceinline__ void sht_rank_sum(const ShTailArgs& a, unsigned e, int n, double* s_out) {
    const double* mine = a.red[a.rank] + (size_t)(e & 1u) * Ctx::P2P_MAX * a.kr;
    for (int j = threadIdx.x; j < n; j += SHT_THREADS) {
        double v[Ctx::P2P_MAX];
#pragma unroll
        for (int r = 0; r < Ctx::P2P_MAX; ++r) v[r] = (r < a.P) ? ld_peer(mine + (size_t)r * a.kr + j) : 0.;
        double s = 0.;
#pragma unroll
        for (int r = 0; r < Ctx::P2P_MAX; ++r) s += v[r];
        s_out[j] = s;
    }
    __syncthreads();
}

__global__ void __launch_bounds__(SHT_THREADS, 1) arnoldi_tail_sharded_kernel(const ShTailArgs a) {
    extern __shared__ double s_mem[];
    const int kk = (a.k + 3) & ~1;
    double* s_h = s_mem;                 // [kk] coefficients being applied / this rank's contribution
    double* s_out = s_mem + kk;          // [kk] reduced over ranks
    double* s_acc = s_mem + 2 * kk;      // [SHT_MAXCH][32 warps][32 lanes]
    double* s_w = s_acc + SHT_MAXCH * SHT_THREADS;   // [rows_per_cta] this CTA's rows of w
    __shared__ int s_flag;
    __shared__ double s_norm;
    const int row0 = blockIdx.x * a.rows_per_cta;
    const int nrows = max(0, min(a.rows_per_cta, a.n_loc - row0));
    const unsigned G = gridDim.x;
    for (int i = threadIdx.x; i < nrows; i += SHT_THREADS) s_w[i] = a.w[row0 + i];
    __syncthreads();
    int stamp_n = 0;
    long long stamps[16];
#define SHT_STAMP() do { if (a.dbg && stamp_n < 16) stamps[stamp_n++] = clock64(); } while (0)
    SHT_STAMP();
    bool was_last0 = false;

    for (int pass = 0; pass < 2; ++pass) {
        // ---- D: partial dots of this CTA's rows ----
        sht_dot_phase(a, row0, nrows, s_w);
        SHT_STAMP();
        const bool last = sht_arrive_is_last(a, pass, &s_flag);
        SHT_STAMP();
        if (pass == 0) was_last0 = last;
        if (last) {
            // ---- R: this rank's coefficients = sum of the CTA partials (a warp per column, lanes over the CTAs: lane l adds
            // CTAs l, l + 32, ... in that order, then a fixed shuffle tree; four columns in flight per warp) ----
            const int lane = threadIdx.x & 31;
            for (int j0 = (threadIdx.x >> 5) * 4; j0 < a.k; j0 += SHT_WARPS * 4) {
                double pv[4][5];   // grid <= 160 CTAs (one per SM): five per lane
#pragma unroll
                for (int u = 0; u < 4; ++u) {
                    const double* pj = a.partial + (size_t)min(j0 + u, a.k - 1) * a.gpad + lane;
#pragma unroll
                    for (int m = 0; m < 5; ++m) pv[u][m] = (lane + 32 * m < (int)G) ? __ldcg(pj + 32 * m) : 0.;
                }
#pragma unroll
                for (int u = 0; u < 4; ++u) {
                    double sv = 0.;
#pragma unroll
                    for (int m = 0; m < 5; ++m) sv += pv[u][m];
#pragma unroll
                    for (int o = 16; o > 0; o >>= 1) sv += __shfl_xor_sync(0xffffffffu, sv, o);
                    if (lane == 0 && j0 + u < a.k) s_h[j0 + u] = sv;
                }
            }
            __syncthreads();
            SHT_STAMP();
            sht_exchange(a, a.seq + (unsigned)pass, s_h, a.k);
            SHT_STAMP();
            sht_publish(a, pass);
        } else {
            sht_wait_ready(a, pass);
        }
        SHT_STAMP();
        // every CTA adds the P contributions itself (rank order): the coefficients this pass applies
        sht_rank_sum(a, a.seq + (unsigned)pass, a.k, s_h);
        if (last) {   // the Hessenberg column for the host, off the critical path of the other CTAs
            for (int j = threadIdx.x; j < a.k; j += SHT_THREADS) {
                if (pass == 0) __stcg(a.h1 + j, s_h[j]);
                else __stcg(a.hfin + j, __ldcg(a.h1 + j) + s_h[j]);
            }
        }
        SHT_STAMP();
        // ---- S: subtract on this CTA's rows ----
        const double sq = sht_sub_phase(a, row0, nrows, s_h, s_acc, s_w, pass == 1);
        SHT_STAMP();
        if (pass == 1) {
            if (threadIdx.x == 0) __stcg(a.npart + blockIdx.x, sq);
            // the all-gather of the next Krylov vector, fused: this CTA's rows go to every rank's window at their global index
            const size_t voff = (size_t)((a.seq + 2u) & 1u) * a.nv;
            for (int p = 0; p < a.P; ++p) {
                double* dst = a.vec[p] + voff;
                for (int i = threadIdx.x; i < nrows; i += SHT_THREADS) dst[a.g_of_local[row0 + i]] = s_w[i];
            }
            // (made visible to the peers by the flag release of the last CTA: see sht_exchange)
        }
    }
    // ---- R3: the norm ----
    const bool last3 = sht_arrive_is_last(a, 2, &s_flag);
    if (last3) {
        if (threadIdx.x < 32) {
            double t = 0.;
            for (unsigned c = threadIdx.x; c < G; c += 32) t += __ldcg(a.npart + c);
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) t += __shfl_xor_sync(0xffffffffu, t, o);
            if (threadIdx.x == 0) s_h[0] = t;
        }
        __syncthreads();
        sht_exchange(a, a.seq + 2u, s_h, 1);
        sht_publish(a, 2);
    } else {
        sht_wait_ready(a, 2);
    }
    sht_rank_sum(a, a.seq + 2u, 1, s_out);
    if (threadIdx.x == 0) {
        s_norm = sqrt(s_out[0]);
        if (last3) {
            __stcg(a.hfin + a.k, s_norm);
            __stcg(a.hfin + a.k + 1, (double)(*reinterpret_cast<volatile int*>(a.err)));
        }
    }
    __syncthreads();
    const double nrm = s_norm;
    // ---- F: the new basis vector (local rows) and the operand of the next matvec (all rows, from the window) ----
    for (int i = threadIdx.x; i < nrows; i += SHT_THREADS) a.qnext[row0 + i] = s_w[i] / nrm;
    const double* mine = a.vec[a.rank] + (size_t)((a.seq + 2u) & 1u) * a.nv;
    for (int g = blockIdx.x * SHT_THREADS + threadIdx.x; g < a.N; g += G * SHT_THREADS) a.xfull[g] = ld_peer(mine + g) / nrm;
    SHT_STAMP();
    if (a.dbg && threadIdx.x == 0 && (blockIdx.x == 0 || was_last0)) {
        long long* d = a.dbg + (was_last0 ? 16 : 0);
        for (int i = 0; i < 16; ++i) d[i] = i < stamp_n ? stamps[i] - stamps[0] : -1;
    }
#undef SHT_STAMP
}

// ---- version 2: COLUMN-owned dot products --------------------------------------------------------------------------------
// In the kernel above a Gram-Schmidt pass costs (measured at n_loc = 7376, k = 535: MACHLINE_SHT_DEBUG) 7 us of row-owned
// partial dots, then 7 us in which ONE CTA adds the 148 x k partials while the others wait, then the exchange.  Here a CTA
// owns whole COLUMNS for the dot products: it reads its ceil(k / grid) columns of Qloc and w_loc once, reduces them inside
// the CTA (fixed order) and stores the finished coefficients straight into every rank's window -- no partial array, no
// serial reduction; the last CTA to arrive only raises and awaits the flags.  The subtraction stays ROW-owned (a CTA keeps
// its rows of w in shared memory), so a pass needs one more grid-wide wait (w complete before the second pass' dots):
//   D1 (columns) | flags | h1 = sum over ranks | S1 (rows) | wait | D2 (columns) | flags | h2 | S2 (rows) + all-gather
//   of the new vector | norm exchange | scale
constexpr int SHT2_NC = 4;   // columns a CTA reduces per trip (1024 threads x 4 rows x 4 columns of loads in flight)

// this rank's dot products of columns [c0, c0 + nc) with w_loc -> slot `rank` of exchange e in every rank's window
__device__ __forceinline__ void sht2_dots(const ShTailArgs& a, unsigned e, int c0, int nc, double* s_red) {
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    double acc[SHT2_NC];
#pragma unroll
    for (int c = 0; c < SHT2_NC; ++c) acc[c] = 0.;
    const double* q0 = a.Q + (size_t)c0 * a.ldq;
    for (int i0 = threadIdx.x; i0 < a.n_loc; i0 += 4 * SHT_THREADS) {
        double wv[4], qv[4][SHT2_NC];
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            const int r = i0 + u * SHT_THREADS;
            const bool on = r < a.n_loc;
            wv[u] = on ? __ldcg(a.w + r) : 0.;
#pragma unroll
            for (int c = 0; c < SHT2_NC; ++c) qv[u][c] = (on && c < nc) ? q0[(size_t)c * a.ldq + r] : 0.;
        }
#pragma unroll
        for (int u = 0; u < 4; ++u)
#pragma unroll
            for (int c = 0; c < SHT2_NC; ++c) acc[c] = fma(qv[u][c], wv[u], acc[c]);
    }
#pragma unroll
    for (int c = 0; c < SHT2_NC; ++c) {
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) acc[c] += __shfl_xor_sync(0xffffffffu, acc[c], o);
        if (lane == 0) s_red[warp * SHT2_NC + c] = acc[c];
    }
    __syncthreads();
    if ((int)threadIdx.x < nc * a.P) {   // thread (column, destination rank): warp order, then one peer store
        const int c = threadIdx.x / a.P, p = threadIdx.x % a.P;
        double v = 0.;
#pragma unroll 8
        for (int w2 = 0; w2 < SHT_WARPS; ++w2) v += s_red[w2 * SHT2_NC + c];
        a.red[p][(size_t)(e & 1u) * Ctx::P2P_MAX * a.kr + (size_t)a.rank * a.kr + c0 + c] = v;
    }
    __syncthreads();
}

// LAST CTA: raise this rank's flag of exchange e in every window (system-scope release: cumulative over the coefficient stores
// of all CTAs, which reached this CTA through their ticket arrival), then wait for the P flags of the own window
__device__ __forceinline__ void sht2_flags(const ShTailArgs& a, unsigned e) {
    const size_t foff = (size_t)(e & 1u) * Ctx::P2P_MAX;
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
}

__global__ void __launch_bounds__(SHT_THREADS, 1) arnoldi_tail_sharded2_kernel(const ShTailArgs a) {
    extern __shared__ double s_mem[];
    const int kk = (a.k + 3) & ~1;
    double* s_h = s_mem;                 // [kk] coefficients being applied
    double* s_out = s_mem + kk;          // [kk] scratch of the norm exchange
    double* s_acc = s_mem + 2 * kk;      // [SHT_MAXCH][32 warps][32 lanes]; the dot phase keeps its warp sums here as well
    double* s_w = s_acc + SHT_MAXCH * SHT_THREADS;   // [rows_per_cta] this CTA's rows of w
    __shared__ int s_flag;
    __shared__ double s_norm;
    const int row0 = blockIdx.x * a.rows_per_cta;
    const int nrows = max(0, min(a.rows_per_cta, a.n_loc - row0));
    const unsigned G = gridDim.x;
    const int npc = (a.k + (int)G - 1) / (int)G;                    // columns per CTA
    const int col_lo = min(a.k, (int)blockIdx.x * npc), col_hi = min(a.k, col_lo + npc);
    int stamp_n = 0;
    long long stamps[16];
#define SHT_STAMP() do { if (a.dbg && stamp_n < 16) stamps[stamp_n++] = clock64(); } while (0)
    SHT_STAMP();
    bool was_last0 = false;

    for (int pass = 0; pass < 2; ++pass) {
        const unsigned e = a.seq + (unsigned)pass;
        // ---- D: dot products of this CTA's columns with all local rows of w ----
        for (int c0 = col_lo; c0 < col_hi; c0 += SHT2_NC) sht2_dots(a, e, c0, min(SHT2_NC, col_hi - c0), s_acc);
        SHT_STAMP();
        const bool last = sht_arrive_is_last(a, 2 * pass, &s_flag);
        if (pass == 0) was_last0 = last;
        if (last) {
            sht2_flags(a, e);
            sht_publish(a, 2 * pass);
        } else {
            sht_wait_ready(a, 2 * pass);
        }
        SHT_STAMP();
        sht_rank_sum(a, e, a.k, s_h);   // the P contributions in rank order: identical bits on every rank and CTA
        if (blockIdx.x == 0) {          // the Hessenberg column for the host
            for (int j = threadIdx.x; j < a.k; j += SHT_THREADS) {
                if (pass == 0) __stcg(a.h1 + j, s_h[j]);
                else {
                    const double hj = __ldcg(a.h1 + j) + s_h[j];
                    __stcg(a.hfin + j, hj);
                    if (a.hhost) a.hhost[j] = hj;
                }
            }
        }
        SHT_STAMP();
        // ---- S: subtract on this CTA's rows ----
        if (pass == 0) {
            for (int i = threadIdx.x; i < nrows; i += SHT_THREADS) s_w[i] = __ldcg(a.w + row0 + i);
            __syncthreads();
        }
        const double sq = sht_sub_phase(a, row0, nrows, s_h, s_acc, s_w, pass == 1);
        SHT_STAMP();
        if (pass == 0) {
            // w must be complete before the second pass' column dots read all of its rows
            const bool lastb = sht_arrive_is_last(a, 1, &s_flag);
            if (lastb) sht_publish(a, 1);
            else sht_wait_ready(a, 1);
            SHT_STAMP();
        } else {
            if (threadIdx.x == 0) __stcg(a.npart + blockIdx.x, sq);
            // the all-gather of the next Krylov vector, fused: this CTA's rows go to every rank's window at their global index
            const size_t voff = (size_t)((a.seq + 2u) & 1u) * a.nv;
            for (int p = 0; p < a.P; ++p) {
                double* dst = a.vec[p] + voff;
                for (int i = threadIdx.x; i < nrows; i += SHT_THREADS) dst[a.g_of_local[row0 + i]] = s_w[i];
            }
        }
    }
    // ---- the norm ----
    const bool last3 = sht_arrive_is_last(a, 3, &s_flag);
    if (last3) {
        if (threadIdx.x < 32) {
            double t = 0.;
            for (unsigned c = threadIdx.x; c < G; c += 32) t += __ldcg(a.npart + c);
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) t += __shfl_xor_sync(0xffffffffu, t, o);
            if (threadIdx.x == 0) s_h[0] = t;
        }
        __syncthreads();
        sht_exchange(a, a.seq + 2u, s_h, 1);
        sht_publish(a, 3);
    } else {
        sht_wait_ready(a, 3);
    }
    sht_rank_sum(a, a.seq + 2u, 1, s_out);
    if (threadIdx.x == 0) {
        s_norm = sqrt(s_out[0]);
        if (last3) {
            const double errd = (double)(*reinterpret_cast<volatile int*>(a.err));
            __stcg(a.hfin + a.k, s_norm);
            __stcg(a.hfin + a.k + 1, errd);
            if (a.hhost) {
                a.hhost[a.k] = s_norm;
                a.hhost[a.k + 1] = errd;
            }
        }
    }
    __syncthreads();
    const double nrm = s_norm;
    for (int i = threadIdx.x; i < nrows; i += SHT_THREADS) a.qnext[row0 + i] = s_w[i] / nrm;
    const double* mine = a.vec[a.rank] + (size_t)((a.seq + 2u) & 1u) * a.nv;
    for (int g = blockIdx.x * SHT_THREADS + threadIdx.x; g < a.N; g += G * SHT_THREADS) a.xfull[g] = ld_peer(mine + g) / nrm;
    SHT_STAMP();
    if (a.dbg && threadIdx.x == 0 && (blockIdx.x == 0 || was_last0)) {
        long long* d = a.dbg + (was_last0 ? 16 : 0);
        for (int i = 0; i < 16; ++i) d[i] = i < stamp_n ? stamps[i] - stamps[0] : -1;
    }
#undef SHT_STAMP
}

}  // namespace mlgpu
