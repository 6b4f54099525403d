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
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#include <cooperative_groups.h>

#include "lu_internal.cuh"

namespace mlgpu {

// ---- implicit row scaling (linalg.f90:193-213) --------------------------------------------------------
// amax(i) = max_j |A(i,j)| over a 2-D grid (rows x column chunks); non-negative doubles order like their bit patterns, so
// the chunks combine with an integer atomicMax.  amax must be zeroed before the launch.
constexpr int RS_COLS = 128;
__global__ void __launch_bounds__(128) lu_row_amax_kernel(const double* __restrict__ A, int ld, int nr, int nc, unsigned long long* __restrict__ amax) {
    const int i = blockIdx.x * 128 + threadIdx.x;
    if (i >= nr) return;
    const int j0 = blockIdx.y * RS_COLS, j1 = min(j0 + RS_COLS, nc);
    double m = 0.;
#pragma unroll 8
    for (int j = j0; j < j1; ++j) m = fmax(m, fabs(A[i + (size_t)j * ld]));
    atomicMax(amax + i, (unsigned long long)__double_as_longlong(m));
}
// vv(i) = 1/amax(i) in place; a zero row flags the matrix singular (linalg.f90:205-208); perm = identity
__global__ void lu_row_scale_finish_kernel(double* __restrict__ vv, int n, int* __restrict__ flag, int* __restrict__ perm) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const double amax = vv[i];
    if (amax <= 1.5e-20) atomicExch(flag, 1);
    vv[i] = 1.0 / amax;
    perm[i] = i;
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

// ---- whole panel in ONE cooperative launch --------------------------------------------------------------
// The per-column kernels above cost two launches and two passes over L2 per column (2 N launches per
// factorisation).  Here the panel (rows k0..n, columns k0..k1) is dealt to the CTAs in blocks of `rpc` row
// positions, each CTA keeps its block in shared memory for the whole panel, and a column costs one grid barrier:
//   before the barrier  every CTA publishes its best candidate (value, position) and that row's panel entries;
//                       the CTA holding position j publishes row j (it moves to the pivot's position);
//   after the barrier   every CTA reduces the G candidates with the reference's rule (largest vv*|a|, ties ->
//                       LAST position, linalg.f90:242), takes the pivot row from the winner's slot, the two
//                       CTAs involved exchange rows j <-> p (all panel columns, as the reference swaps whole
//                       rows, :254-262, and vv(p) = vv(j), :263), and all update their rows below j.
// Candidate slots are double-buffered by column parity: a slot written for column j is next written for
// column j+2, after barrier j+1, i.e. after every CTA finished reading it.  Arithmetic per element is
// a(i,c) = fma(-l, a(j,c), a(i,c)) with l = a(i,j) * (1/a(j,j)), identical to the per-column kernels.
// 256 threads per CTA; 128 (<= 96 registers) when a CTA holds 128 rows, so that it fits on an SM next to a CTA of the
// trailing update: the look-ahead factors panel k+1 under the update of panel k.
#ifndef ML_LUP_THREADS
#define ML_LUP_THREADS 512
#endif
constexpr int LUP_THREADS = ML_LUP_THREADS;   // 16 warps: the per-column work of a CTA is a chain of shared-memory latencies
constexpr int LUP_CAP = 384;   // rows a CTA can hold: 64 columns x 384 rows x 8 B = 192 KB
constexpr int LUP_CAP_OVF = 320;   // shared-memory rows per CTA when the panel overflows (the rest stay in global memory)
constexpr int LUP_RPC_MAX = 2048;  // rows per CTA the overflow variant supports (s_l, s_vv, s_perm are per row)
constexpr int LUP_ROW = LU_NB + 2;   // a published row: LU_NB panel entries, perm, vv

struct LuPanelArgs {
    double* A;
    int ld, n, k0, k1, rpc;
    double* vv;
    int* piv;
    int* perm;          // perm[i] = original row now at position i (the composition of the interchanges so far)
    double* cand_v;     // [2][G]
    int* cand_i;        // [2][G]
    double* cand_row;   // [2][G][LUP_ROW]  panel entries of the candidate row, then its perm entry
    double* rowj;       // [2][LUP_ROW]     panel entries of row j, its perm entry, its vv
    unsigned* bar;
    unsigned bar_base;
};

__device__ __forceinline__ void lup_grid_sync(unsigned* bar, unsigned target) {
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

__device__ __forceinline__ bool lup_better(double v, int i, double best, int bi) { return v > best || (v == best && i > bi); }

// CL: the CTAs form ONE thread-block cluster (<= 16 CTAs): the candidate slots live in every CTA's shared memory and are written
// by the peers through distributed shared memory, the per-column barrier is the hardware cluster barrier instead of an atomic
// counter in global memory (three L2 round trips + a polling loop per column), and the kernel occupies at most 16 SMs, so it
// runs NEXT TO the trailing update of the previous pair of panels (look-ahead).
constexpr int LUP_CL_MAX = 16;   // CTAs of the cluster variant
template <bool OVF, int T, bool CL = false>
__global__ void __launch_bounds__(T, T == 128 ? 5 : 1) lu_panel_coop_kernel(const LuPanelArgs a) {
    namespace cg = cooperative_groups;
    extern __shared__ __align__(16) double lup_smem[];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int G = gridDim.x, bid = blockIdx.x;
    const int nb = a.k1 - a.k0;
    const int r0 = a.k0 + bid * a.rpc;                 // first row position of this CTA
    const int nr = min(a.rpc, a.n - r0);               // > 0 by grid sizing
    // OVF: the panel has more rows than the CTAs' shared memory holds; the first `cap` rows of a CTA live in shared
    // memory, the rest stay where they are in global memory (only their owner CTA ever touches them).
    const int cap = OVF ? LUP_CAP_OVF : a.rpc;
    const int S = cap | 1;                             // odd column stride: row gathers are conflict-free
    double* sP = lup_smem;                             // [LU_NB][S]
    double* const gA = a.A + (size_t)a.k0 * a.ld + r0; // this CTA's rows of the panel in global memory
    auto el = [&](int r, int c) -> double& {
        if (OVF && r >= cap) return gA[(size_t)c * a.ld + r];
        return sP[c * S + r];
    };
    double* s_piv = sP + LU_NB * S;                    // [LU_NB]  pivot row, panel columns
    double* s_oldj = s_piv + LUP_ROW;                  // [LUP_ROW] old row j (+ perm, vv)
    double* s_l = s_oldj + LUP_ROW;                    // [rpc]
    double* s_vv = s_l + a.rpc;                        // [rpc]
    constexpr int NW = T / 32 < 8 ? 8 : T / 32;        // warp partials (at least the 8 slots the layout always had)
    double* s_rv = s_vv + a.rpc;                       // [32]: warp partials, s_rv[NW] = the CTA's best value
    int* s_ri = reinterpret_cast<int*>(s_rv + 32);     // [32] + s_ri[32] = winner
    int* s_perm = s_ri + 34;                           // [rpc]
    // CL: candidate slots of every CTA of the cluster, two parities (written by the peers through DSMEM)
    double* cl_row = reinterpret_cast<double*>(s_perm + ((a.rpc + 1) & ~1));   // [2][LUP_CL_MAX][LUP_ROW]
    double* cl_rowj = cl_row + 2 * LUP_CL_MAX * LUP_ROW;                        // [2][LUP_ROW]
    double* cl_v = cl_rowj + 2 * LUP_ROW;                                       // [2][LUP_CL_MAX]
    int* cl_i = reinterpret_cast<int*>(cl_v + 2 * LUP_CL_MAX);                  // [2][LUP_CL_MAX]
    const int nrb = (nr + 31) >> 5;
    if constexpr (CL) cg::this_cluster().sync();       // every CTA of the cluster is running before the first remote store

    for (int item = warp; item < nb * nrb; item += T / 32) {
        const int c = item / nrb, r = ((item - c * nrb) << 5) + lane;
        if (r < nr && r < cap) sP[c * S + r] = __ldcg(a.A + (size_t)(a.k0 + c) * a.ld + r0 + r);
    }
    for (int r = tid; r < nr; r += T) {
        s_vv[r] = __ldcg(a.vv + r0 + r);
        s_perm[r] = __ldcg(a.perm + r0 + r);
    }
    __syncthreads();

    for (int jj = 0; jj < nb; ++jj) {
        const int j = a.k0 + jj, par = jj & 1;
        // ---- this CTA's candidate for column j ------------------------------------------------------------
        double best = -1.;
        int bi = -1;
        for (int r = tid; r < nr; r += T) {
            const int gp = r0 + r;
            if (gp >= j) {
                const double v = s_vv[r] * fabs(el(r, jj));
                if (lup_better(v, gp, best, bi)) { best = v; bi = gp; }
            }
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            const double ov = __shfl_xor_sync(0xffffffffu, best, o);
            const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
            if (lup_better(ov, oi, best, bi)) { best = ov; bi = oi; }
        }
        if (lane == 0) { s_rv[warp] = best; s_ri[warp] = bi; }
        __syncthreads();
        if (warp == 0) {
            best = lane < T / 32 ? s_rv[lane] : -1.;
            bi = lane < T / 32 ? s_ri[lane] : -1;
#pragma unroll
            for (int o = NW / 2; o > 0; o >>= 1) {
                const double ov = __shfl_xor_sync(0xffffffffu, best, o);
                const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
                if (lup_better(ov, oi, best, bi)) { best = ov; bi = oi; }
            }
            if (lane == 0) {
                s_ri[32] = bi;
                if constexpr (CL) {
                    s_rv[NW] = best;
                } else {
                    __stcg(a.cand_v + par * G + bid, best);
                    __stcg(a.cand_i + par * G + bid, bi);
                }
            }
        }
        __syncthreads();
        bi = s_ri[32];
        const bool own_j = (j >= r0 && j < r0 + nr);
        if constexpr (CL) {
            // this CTA's candidate (value, position, the row's panel entries, its perm entry) and -- from the CTA that holds
            // it -- row j go into the slots of EVERY CTA of the cluster: thread t serves destination t / 80, word t % 80
            cg::cluster_group cluster = cg::this_cluster();
            for (int t = tid; t < G * 80; t += T) {
                const int dst = t / 80, wd = t - dst * 80;
                if (wd < LUP_ROW) {
                    if (bi >= 0 && (wd < nb || wd == LU_NB)) {
                        const double v = wd < nb ? el(bi - r0, wd) : (double)s_perm[bi - r0];
                        cluster.map_shared_rank(cl_row, dst)[(par * LUP_CL_MAX + bid) * LUP_ROW + wd] = v;
                    }
                } else if (wd == LUP_ROW) {
                    cluster.map_shared_rank(cl_v, dst)[par * LUP_CL_MAX + bid] = s_rv[NW];
                } else if (wd == LUP_ROW + 1) {
                    cluster.map_shared_rank(cl_i, dst)[par * LUP_CL_MAX + bid] = bi;
                } else if (own_j && wd - (LUP_ROW + 2) < 12) {
                    // 12 threads per destination copy row j: words 6 q .. 6 q + 5 of its 66-word slot (entries, perm, vv)
                    const int q = wd - (LUP_ROW + 2);
                    double* rj = cluster.map_shared_rank(cl_rowj, dst) + par * LUP_ROW;
                    for (int u2 = 6 * q; u2 < 6 * q + 6 && u2 < LUP_ROW; ++u2) {
                        double v = 0.;
                        if (u2 < nb) v = el(j - r0, u2);
                        else if (u2 == LU_NB) v = (double)s_perm[j - r0];
                        else if (u2 == LU_NB + 1) v = s_vv[j - r0];
                        rj[u2] = v;
                    }
                }
            }
            cluster.sync();
        } else {
            if (bi >= 0) {
                double* slot = a.cand_row + ((size_t)par * G + bid) * LUP_ROW;
                if (tid < nb) __stcg(slot + tid, el(bi - r0, tid));
                if (tid == LU_NB) __stcg(slot + LU_NB, (double)s_perm[bi - r0]);
            }
            if (own_j) {
                double* slot = a.rowj + par * LUP_ROW;
                if (tid < nb) __stcg(slot + tid, el(j - r0, tid));
                if (tid == LU_NB) __stcg(slot + LU_NB, (double)s_perm[j - r0]);
                if (tid == LU_NB + 1) __stcg(slot + LU_NB + 1, s_vv[j - r0]);
            }
            lup_grid_sync(a.bar, a.bar_base + (unsigned)(jj + 1) * (unsigned)G);
        }
        // ---- global pivot: reduce the G candidates (every CTA, redundantly) ---------------------------------
        best = -1.;
        bi = -1;
        for (int g = tid; g < G; g += T) {
            const double v = CL ? cl_v[par * LUP_CL_MAX + g] : __ldcg(a.cand_v + par * G + g);
            const int i = CL ? cl_i[par * LUP_CL_MAX + g] : __ldcg(a.cand_i + par * G + g);
            if (i >= 0 && lup_better(v, i, best, bi)) { best = v; bi = i; }
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            const double ov = __shfl_xor_sync(0xffffffffu, best, o);
            const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
            if (lup_better(ov, oi, best, bi)) { best = ov; bi = oi; }
        }
        if (G > 32) {
            if (lane == 0) { s_rv[warp] = best; s_ri[warp] = bi; }
            __syncthreads();
            if (warp == 0) {
                best = lane < T / 32 ? s_rv[lane] : -1.;
                bi = lane < T / 32 ? s_ri[lane] : -1;
#pragma unroll
                for (int o = NW / 2; o > 0; o >>= 1) {
                    const double ov = __shfl_xor_sync(0xffffffffu, best, o);
                    const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
                    if (lup_better(ov, oi, best, bi)) { best = ov; bi = oi; }
                }
                if (lane == 0) s_ri[32] = bi;
            }
        } else if (tid == 0) {
            s_ri[32] = bi;   // warp 0 saw every candidate
        }
        __syncthreads();
        int p = s_ri[32];
        if (p < 0) p = j;   // a column of NaNs: keep the diagonal (the reference's imax stays at its previous value)
        const int w = (p - a.k0) / a.rpc;
        const bool own_p = (p >= r0 && p < r0 + nr);
        if (tid < nb || tid == LU_NB)
            s_piv[tid] = CL ? cl_row[(par * LUP_CL_MAX + w) * LUP_ROW + tid] : __ldcg(a.cand_row + ((size_t)par * G + w) * LUP_ROW + tid);
        if (p != j && own_p && (tid < nb || tid == LU_NB || tid == LU_NB + 1))
            s_oldj[tid] = CL ? cl_rowj[par * LUP_ROW + tid] : __ldcg(a.rowj + par * LUP_ROW + tid);
        if (bid == 0 && tid == 0) a.piv[j] = p;
        __syncthreads();
        if (p != j) {   // whole-row interchange inside the panel (linalg.f90:254-263)
            if (own_j) {
                if (tid < nb) el(j - r0, tid) = s_piv[tid];
                if (tid == LU_NB) s_perm[j - r0] = (int)s_piv[LU_NB];
            }
            if (own_p) {
                if (tid < nb) el(p - r0, tid) = s_oldj[tid];
                if (tid == LU_NB) s_perm[p - r0] = (int)s_oldj[LU_NB];
                if (tid == LU_NB + 1) s_vv[p - r0] = s_oldj[LU_NB + 1];
            }
            __syncthreads();
        }
        // ---- eliminate column j from the rows below it ------------------------------------------------------
        const double inv = 1.0 / s_piv[jj];
        for (int r = tid; r < nr; r += T) {
            double l = 0.;
            if (r0 + r > j) {
                l = el(r, jj) * inv;
                el(r, jj) = l;
            }
            s_l[r] = l;
        }
        __syncthreads();
        // items = (32-row block, group of 4 columns): four independent load / fma / store chains per thread (one element
        // per item left the loop latency-bound: ~75 cycles per element and warp)
        const int ncr = nb - jj - 1, ncg = (ncr + 3) >> 2;
        for (int item = warp; item < ncg * nrb; item += T / 32) {
            const int cg = item / nrb, r = ((item - cg * nrb) << 5) + lane;
            const int c0 = jj + 1 + (cg << 2);
            if (r < nr && r0 + r > j) {
                const double ml = -s_l[r];
                double v[4];
#pragma unroll
                for (int u = 0; u < 4; ++u)
                    if (c0 + u < nb) v[u] = el(r, c0 + u);
#pragma unroll
                for (int u = 0; u < 4; ++u)
                    if (c0 + u < nb) v[u] = fma(ml, s_piv[c0 + u], v[u]);
#pragma unroll
                for (int u = 0; u < 4; ++u)
                    if (c0 + u < nb) el(r, c0 + u) = v[u];
            }
        }
        __syncthreads();
    }
    for (int item = warp; item < nb * nrb; item += T / 32) {
        const int c = item / nrb, r = ((item - c * nrb) << 5) + lane;
        if (r < nr && r < cap) a.A[(size_t)(a.k0 + c) * a.ld + r0 + r] = sP[c * S + r];
    }
    for (int r = tid; r < nr; r += T) {
        a.vv[r0 + r] = s_vv[r];
        a.perm[r0 + r] = s_perm[r];
    }
}

// ---- cluster panel, second generation ---------------------------------------------------------------------
// Same factorisation, same arithmetic per element (a(i,c) = fma(-l, a(j,c), a(i,c)) for the panel's columns j in ascending order,
// l = a(i,j) * (1/a(j,j)), the reference's pivot rule), organised for latency: lu_panel_coop_kernel<.., CL> spends ~8400 cycles per
// column whatever the panel's height (ncu: 645 warp instructions per warp and column, nine block barriers, 32 % of the stall
// samples in the cluster barrier), and the panel is the critical path of the factorisation below N ~ 15k.
//   * thread = row: a CTA holds up to 512 rows, thread t owns row t (its scale vv and permutation entry live in registers), so
//     the multiplier, the update and the next column's candidate of a row never leave the thread: no s_l array, no barrier
//     between them;
//   * delayed update: columns are grouped in blocks of LUP2_IB = 8.  A column step updates only the (<= 7) columns left in its
//     block; the columns right of the block receive the block's eight rank-1 updates at the block's end in one pass (each
//     element read and written once instead of eight times, eight dependent FMAs in the order of the column steps: bit-identical
//     values).  The pivot rows of a block arrive stale in those columns; every CTA rebuilds them (redundantly, <= 7 FMAs per
//     column: s_U) from the pulled row and the row's own multipliers;
//   * pull, not push: a CTA parks its candidate row (and the CTA that holds position j parks row j) in a slot of its OWN shared
//     memory, double-buffered by column parity; only (value, position) -- 12 bytes -- go to the peers before the cluster
//     barrier.  After it every CTA reads the winner's row from the winner's slot through distributed shared memory (one round
//     trip) instead of every CTA storing 528 bytes into 16 peers per column.  (Pulling the 12 bytes as well makes the barrier
//     cheaper, 1200 -> 880 cycles with no remote store to release, but the second dependent round trip costs more: r5e.)
//   * arg max by three REDUX (lup2_warp_argmax); both reductions (CTA candidate, cluster pivot) are finished redundantly by
//     every warp: one block barrier each.
// Block barriers per column: three (+ one with an interchange inside the CTA, + one per block end); one cluster barrier.
// OVF: rows beyond the shared-memory capacity of a CTA stay in global memory (touched by their owner thread and, for an
// interchange, by the CTA's column threads: L2 loads / stores ordered by the block barriers).
constexpr int LUP2_IB = 8;
constexpr int LUP2_T = 512;     // most threads of a CTA = the most rows a CTA can own
constexpr int LUP2_CAP = 416;   // rows of a CTA held in shared memory: 64 columns x 417 x 8 B = 213.5 KB

// arg max of (v, position) over a warp with the reference's tie rule (largest v, ties -> LAST position), three REDUX
// instead of a five-step shuffle tree on 96-bit items.  key = bit pattern of v (v >= +0: ordered like the values), idx >= 0;
// "no candidate" = (0, -1), which loses against every real candidate (a real v = 0 has idx >= 0 > -1).
__device__ __forceinline__ void lup2_warp_argmax(unsigned long long& key, int& idx) {
    const unsigned hi = (unsigned)(key >> 32), lo = (unsigned)key;
    const unsigned mh = __reduce_max_sync(0xffffffffu, hi);
    const unsigned ml = __reduce_max_sync(0xffffffffu, hi == mh ? lo : 0u);
    idx = __reduce_max_sync(0xffffffffu, (hi == mh && lo == ml) ? idx : -1);
    key = ((unsigned long long)mh << 32) | ml;
}
// (Tried, r5e: one REDUX on the high word + ballot + two shuffles when a single lane holds the maximum -- not faster.)

template <bool OVF>
__global__ void __launch_bounds__(LUP2_T, 1) lu_panel_cl2_kernel(const LuPanelArgs a) {
    namespace cg = cooperative_groups;
    cg::cluster_group cluster = cg::this_cluster();
    extern __shared__ __align__(16) double lup_smem[];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, nwarps = blockDim.x >> 5;
    const int G = gridDim.x, bid = blockIdx.x;
    const int nb = a.k1 - a.k0;
    const int r0 = a.k0 + bid * a.rpc;                 // first row position of this CTA
    const int nr = min(a.rpc, a.n - r0);               // > 0 by grid sizing; <= blockDim.x
    const int cap = OVF ? LUP2_CAP : a.rpc;
    const int S = cap | 1;                             // odd column stride: gathers of a row are conflict-free
    double* const sP = lup_smem;                       // [LU_NB][S]
    double* const s_U = sP + LU_NB * S;                // [LUP2_IB][LU_NB]  pivot rows of the current block, delayed columns
    double* const s_piv = s_U + LUP2_IB * LU_NB;       // [LUP_ROW] pivot row as pulled (+ its perm entry, 1 / pivot)
    double* const s_oldj = s_piv + LUP_ROW;            // [LUP_ROW] row j as pulled (+ perm, vv)
    double* const st_cand = s_oldj + LUP_ROW;          // [2][LUP_ROW] this CTA's candidate row, read by the peers
    double* const st_rowj = st_cand + 2 * LUP_ROW;     // [2][LUP_ROW] row j (+ perm, vv), read by the CTA that holds the pivot
    unsigned long long* const cl_k = reinterpret_cast<unsigned long long*>(st_rowj + 2 * LUP_ROW);   // [2][LUP_CL_MAX] candidate keys of every CTA (written by the peers)
    unsigned long long* const s_rk = cl_k + 2 * LUP_CL_MAX;   // [16] warp partials
    int* const cl_i = reinterpret_cast<int*>(s_rk + 16);      // [2][LUP_CL_MAX]
    int* const s_ri = cl_i + 2 * LUP_CL_MAX;           // [16]
    double* const gA = a.A + (size_t)a.k0 * a.ld + r0; // this CTA's rows of the panel in global memory
    auto ld_el = [&](int r, int c) -> double {
        if (OVF && r >= cap) return __ldcg(gA + (size_t)c * a.ld + r);
        return sP[c * S + r];
    };
    auto st_el = [&](int r, int c, double v) {
        if (OVF && r >= cap) __stcg(gA + (size_t)c * a.ld + r, v);
        else sP[c * S + r] = v;
    };
    // MACHLINE_LU_PANEL_DBG: cycles per phase of a column step, summed by thread 0 of CTA 0 (development aid)
    long long* const dbg = reinterpret_cast<long long*>(a.cand_row);
    const bool stamping = dbg != nullptr && bid == 0 && tid == 0;
    long long t_prev = 0;
    __shared__ long long s_dbg[8];
    if (stamping)
        for (int k = 0; k < 8; ++k) s_dbg[k] = 0;
#define LUP2_STAMP(slot)                                     \
    do {                                                     \
        if (stamping) {                                      \
            const long long t_now = clock64();               \
            s_dbg[slot] += t_now - t_prev;                   \
            t_prev = t_now;                                  \
        }                                                    \
    } while (0)
    const int nrb = (min(nr, cap) + 31) >> 5;
    for (int item = warp; item < nb * nrb; item += nwarps) {
        const int c = item / nrb, r = ((item - c * nrb) << 5) + lane;
        if (r < nr && r < cap) sP[c * S + r] = __ldcg(a.A + (size_t)(a.k0 + c) * a.ld + r0 + r);
    }
    const bool has_row = tid < nr;
    double vv_r = has_row ? __ldcg(a.vv + r0 + tid) : 0.;
    int perm_r = has_row ? __ldcg(a.perm + r0 + tid) : 0;
    cluster.sync();   // the panel is in shared memory; every CTA of the cluster is running before the first remote access
    if (stamping) t_prev = clock64();

    int wj = 0;
    for (int jj = 0; jj < nb; ++jj) {
        const int j = a.k0 + jj, par = jj & 1;
        const int blk0 = jj & ~(LUP2_IB - 1), blkend = min(blk0 + LUP2_IB, nb);
        if (jj >= (wj + 1) * a.rpc) ++wj;   // the CTA that holds position j
        // ---- this CTA's candidate for column j (largest vv*|a|, ties -> last position, linalg.f90:242) -------------
        unsigned long long key = 0ull;
        int bi = -1;
        if (has_row && r0 + tid >= j) {
            const double v = vv_r * fabs(ld_el(tid, jj));
            if (v > -1.) { key = (unsigned long long)__double_as_longlong(v); bi = r0 + tid; }   // a NaN is never a candidate
        }
        lup2_warp_argmax(key, bi);
        if (lane == 0) { s_rk[warp] = key; s_ri[warp] = bi; }
        __syncthreads();
        key = lane < nwarps ? s_rk[lane] : 0ull;
        bi = lane < nwarps ? s_ri[lane] : -1;
        lup2_warp_argmax(key, bi);                         // every warp finishes the reduction: no second barrier
        LUP2_STAMP(0);
        // ---- (value, position) to every CTA of the cluster FIRST: the remote stores' round trips (which the barrier's release
        // waits for) run under the parking of the candidate row and the division below; then park the candidate row and row j
        // in this CTA's slots ----
        if (tid < G) {
            cluster.map_shared_rank(cl_k, tid)[par * LUP_CL_MAX + bid] = key;
            cluster.map_shared_rank(cl_i, tid)[par * LUP_CL_MAX + bid] = bi;
        }
        const bool own_j = (j >= r0 && j < r0 + nr);
        const int ci = bi >= 0 ? bi : (own_j ? j : -1);    // a column of NaNs: row j stands in (the pivot stays on the diagonal)
        if (ci >= 0) {
            if (tid < nb) st_cand[par * LUP_ROW + tid] = ld_el(ci - r0, tid);
            if (tid == ci - r0) st_cand[par * LUP_ROW + LU_NB] = (double)perm_r;
            // 1 / pivot of this candidate, under the cluster barrier's own latency instead of after it
            if (tid == LU_NB + 1) st_cand[par * LUP_ROW + LU_NB + 1] = 1.0 / ld_el(ci - r0, jj);
        }
        if (own_j) {
            if (tid < nb) st_rowj[par * LUP_ROW + tid] = ld_el(j - r0, tid);
            if (tid == j - r0) {
                st_rowj[par * LUP_ROW + LU_NB] = (double)perm_r;
                st_rowj[par * LUP_ROW + LU_NB + 1] = vv_r;
            }
        }
        LUP2_STAMP(1);
        cluster.sync();
        LUP2_STAMP(2);
        // ---- global pivot (every warp, redundantly), then the winner's row from the winner's slot -------------------
        key = lane < G ? cl_k[par * LUP_CL_MAX + lane] : 0ull;
        bi = lane < G ? cl_i[par * LUP_CL_MAX + lane] : -1;
        const int my_i = bi;
        lup2_warp_argmax(key, bi);
        const int p = bi < 0 ? j : bi;   // a column of NaNs: keep the diagonal (the reference's imax stays at its previous value)
        // the CTA that holds the pivot = the lane whose candidate won (positions are unique); no integer division per column
        const unsigned holder = __ballot_sync(0xffffffffu, lane < G && my_i == bi);
        const int w = bi < 0 ? wj : __ffs((int)holder) - 1;
        const bool own_p = (p >= r0 && p < r0 + nr);
        double pulled = 0.;
        if (tid < nb || tid == LU_NB || tid == LU_NB + 1) {
            pulled = cluster.map_shared_rank(st_cand, w)[par * LUP_ROW + tid];
            s_piv[tid] = pulled;
        }
        if (p != j && own_p && (tid < nb || tid == LU_NB || tid == LU_NB + 1))
            s_oldj[tid] = cluster.map_shared_rank(st_rowj, wj)[par * LUP_ROW + tid];
        if (bid == 0 && tid == 0) a.piv[j] = p;
        __syncthreads();
        LUP2_STAMP(3);
        // the pivot row in the delayed columns: it has seen the blocks before this one, not this block's columns (< j) yet
        if (tid >= blkend && tid < nb) {
            // all loads, then the chain of FMAs (the two warps that do this are the last to reach the next block barrier)
            const int jb = jj - blk0;
            double lb[LUP2_IB - 1], ub[LUP2_IB - 1];
#pragma unroll
            for (int b = 0; b < LUP2_IB - 1; ++b) {
                if (b < jb) {
                    lb[b] = s_piv[blk0 + b];
                    ub[b] = s_U[b * LU_NB + tid];
                }
            }
            double u = pulled;
#pragma unroll
            for (int b = 0; b < LUP2_IB - 1; ++b)
                if (b < jb) u = fma(-lb[b], ub[b], u);
            s_U[jb * LU_NB + tid] = u;
        }
        if (p != j && (own_j || own_p)) {   // whole-row interchange inside the panel (linalg.f90:254-263)
            if (own_j) {
                if (tid < nb) st_el(j - r0, tid, s_piv[tid]);
                if (tid == j - r0) perm_r = (int)s_piv[LU_NB];
            }
            if (own_p) {
                if (tid < nb) st_el(p - r0, tid, s_oldj[tid]);
                if (tid == p - r0) {
                    perm_r = (int)s_oldj[LU_NB];
                    vv_r = s_oldj[LU_NB + 1];
                }
            }
            __syncthreads();
        }
        LUP2_STAMP(4);
        // ---- eliminate column j from this thread's row: the multiplier and the columns left in the block -------------
        const double inv = s_piv[LU_NB + 1];   // 1.0 / s_piv[jj], computed by the pivot's CTA
        if (has_row && r0 + tid > j) {
            // loads first, then the FMAs, then the stores: written as a loop of load / fma / store the compiler has to keep the
            // shared-memory accesses in order (it cannot tell sP from s_piv) and the <= 7 columns go one after the other
            const int ncol = blkend - jj - 1;
            double v[LUP2_IB - 1], u[LUP2_IB - 1];
            const double e = ld_el(tid, jj);
#pragma unroll
            for (int k = 0; k < LUP2_IB - 1; ++k) {
                if (k < ncol) {
                    v[k] = ld_el(tid, jj + 1 + k);
                    u[k] = s_piv[jj + 1 + k];
                }
            }
            const double l = e * inv;
            st_el(tid, jj, l);
            const double ml = -l;
#pragma unroll
            for (int k = 0; k < LUP2_IB - 1; ++k)
                if (k < ncol) st_el(tid, jj + 1 + k, fma(ml, u[k], v[k]));
        }
        LUP2_STAMP(5);
        // ---- end of a block: its eight updates of the columns to its right, in the order of the column steps ---------
        if (jj + 1 == blkend && blkend < nb) {   // blkend < nb: the block is complete (LUP2_IB columns)
            __syncthreads();                     // s_U is complete
            if (tid >= blkend && tid < nb) {     // the block's pivot rows take their final values in the delayed columns
#pragma unroll
                for (int b = 0; b < LUP2_IB; ++b) {
                    const int pos = a.k0 + blk0 + b;
                    if (pos >= r0 && pos < r0 + nr) st_el(pos - r0, tid, s_U[b * LU_NB + tid]);
                }
            }
            if (has_row && r0 + tid >= a.k0 + blkend) {
                double ml[LUP2_IB];
#pragma unroll
                for (int b = 0; b < LUP2_IB; ++b) ml[b] = -ld_el(tid, blk0 + b);
                int c = blkend;
                for (; c + 4 <= nb; c += 4) {
                    double v0 = ld_el(tid, c), v1 = ld_el(tid, c + 1), v2 = ld_el(tid, c + 2), v3 = ld_el(tid, c + 3);
#pragma unroll
                    for (int b = 0; b < LUP2_IB; ++b) {
                        const double2 u01 = *reinterpret_cast<const double2*>(s_U + b * LU_NB + c);
                        const double2 u23 = *reinterpret_cast<const double2*>(s_U + b * LU_NB + c + 2);
                        v0 = fma(ml[b], u01.x, v0);
                        v1 = fma(ml[b], u01.y, v1);
                        v2 = fma(ml[b], u23.x, v2);
                        v3 = fma(ml[b], u23.y, v3);
                    }
                    st_el(tid, c, v0);
                    st_el(tid, c + 1, v1);
                    st_el(tid, c + 2, v2);
                    st_el(tid, c + 3, v3);
                }
                for (; c < nb; ++c) {
                    double v = ld_el(tid, c);
#pragma unroll
                    for (int b = 0; b < LUP2_IB; ++b) v = fma(ml[b], s_U[b * LU_NB + c], v);
                    st_el(tid, c, v);
                }
            }
            LUP2_STAMP(6);
        }
    }
#undef LUP2_STAMP
    if (stamping) {
        for (int k = 0; k < 7; ++k) dbg[k] += s_dbg[k];
        dbg[7] += nb;
    }
    cluster.sync();   // no CTA leaves while a peer may still read its slots
    for (int item = warp; item < nb * nrb; item += nwarps) {
        const int c = item / nrb, r = ((item - c * nrb) << 5) + lane;
        if (r < nr && r < cap) a.A[(size_t)(a.k0 + c) * a.ld + r0 + r] = sP[c * S + r];
    }
    if (has_row) {
        a.vv[r0 + tid] = vv_r;
        a.perm[r0 + tid] = perm_r;
    }
}

static size_t lu_panel2_smem(int rpc) {
    const int cap = rpc <= LUP2_CAP ? rpc : LUP2_CAP;
    return (size_t)(LU_NB * (cap | 1) + LUP2_IB * LU_NB + 6 * LUP_ROW + 2 * LUP_CL_MAX + 16) * sizeof(double) +
           (size_t)(2 * LUP_CL_MAX + 16) * sizeof(int);
}

// the per-column fallback keeps only the interchanges: compose them (sequentially) into the row permutation
__global__ void lu_perm_from_piv_kernel(const int* __restrict__ piv, int n, int* __restrict__ perm) {
    if (blockIdx.x == 0 && threadIdx.x == 0) {
        for (int i = 0; i < n; ++i) perm[i] = i;
        for (int i = 0; i < n; ++i) {
            const int p = piv[i];
            if (p != i) {
                const int t = perm[p];
                perm[p] = perm[i];
                perm[i] = t;
            }
        }
    }
}

// apply the panel's row interchanges to columns [c0, c1)
// columns [c0, c1) and [c2, c3) (the second range may be empty): left and right of a pair of panels in one launch
__global__ void __launch_bounds__(256) lu_laswp_kernel(double* __restrict__ A, int ld, int c0, int c1, int c2, int c3, int k0, int k1,
                                                        const int* __restrict__ piv) {
    const int idx = blockIdx.x * 256 + threadIdx.x, n1 = c1 - c0;   // c1 >= c0, c3 >= c2
    const int c = idx < n1 ? c0 + idx : c2 + (idx - n1);
    if (idx >= n1 && c >= c3) return;
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

// U12 = L11^{-1} A12, L11 (unit lower, 64 x 64) in shared memory.  Only full panels reach this kernel (a short last panel has
// nothing to its right).  Column-oriented substitution, x[r] = fma(-L[r][k], x[k], x[r]) for k ascending: per element the
// reference's order.  A warp carries TRSM_CW columns at once, lane = rows (lane, lane + 32) of each: step k broadcasts x[k]
// from its owner lane by shuffle, every lane updates its two rows of every column.  (The first version ran one thread per
// column with the whole column in registers: 2016 dependent shared-memory broadcasts per thread on two warps per SM, 19 us
// per call whatever the width, and three calls per pair of panels sit on the critical path of the factorisation.)
constexpr int TRSM_CW = 4;
constexpr int TRSM_THREADS = 256;
constexpr int TRSM_COLS_PER_CTA = TRSM_THREADS / 32 * TRSM_CW;
__global__ void __launch_bounds__(TRSM_THREADS) lu_trsm_kernel(const double* __restrict__ L, int ldl, double* __restrict__ X, int ldx, int ncols) {
    // L: the 64 x 64 diagonal block (unit lower part used); X: 64 x ncols right-hand sides, overwritten by L^{-1} X
    __shared__ double sL[LU_NB * LU_NB];   // sL[k * LU_NB + r] = L(r, k): column-major, conflict-free fill and reads
    for (int t = threadIdx.x; t < LU_NB * LU_NB; t += TRSM_THREADS) {
        const int r = t % LU_NB, k = t / LU_NB;
        sL[t] = L[r + (size_t)k * ldl];
    }
    __syncthreads();
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int c0 = (blockIdx.x * (TRSM_THREADS / 32) + warp) * TRSM_CW;
    if (c0 >= ncols) return;
    double x0[TRSM_CW], x1[TRSM_CW];
#pragma unroll
    for (int c = 0; c < TRSM_CW; ++c) {
        const double* col = X + (size_t)min(c0 + c, ncols - 1) * ldx;   // a short last group repeats its last column (not stored)
        x0[c] = col[lane];
        x1[c] = col[lane + 32];
    }
#pragma unroll 4
    for (int k = 0; k < 32; ++k) {          // x[k] lives in x0 of lane k
        const double l0 = -sL[k * LU_NB + lane], l1 = -sL[k * LU_NB + lane + 32];
#pragma unroll
        for (int c = 0; c < TRSM_CW; ++c) {
            const double xk = __shfl_sync(0xffffffffu, x0[c], k);
            if (lane > k) x0[c] = fma(l0, xk, x0[c]);
            x1[c] = fma(l1, xk, x1[c]);
        }
    }
#pragma unroll 4
    for (int k = 32; k < LU_NB - 1; ++k) {  // x[k] lives in x1 of lane k - 32
        const double l1 = -sL[k * LU_NB + lane + 32];
#pragma unroll
        for (int c = 0; c < TRSM_CW; ++c) {
            const double xk = __shfl_sync(0xffffffffu, x1[c], k - 32);
            if (lane + 32 > k) x1[c] = fma(l1, xk, x1[c]);
        }
    }
#pragma unroll
    for (int c = 0; c < TRSM_CW; ++c) {
        if (c0 + c < ncols) {
            double* col = X + (size_t)(c0 + c) * ldx;
            col[lane] = x0[c];
            col[lane + 32] = x1[c];
        }
    }
}

// ---- trailing update on the FP64 tensor cores ---------------------------------------------------------
// C[M x Nc] -= L[M x 64] * U[64 x Nc]; CTA tile 128 x 64, 8 warps (4 along M x 2 along N), warp tile 32 x 32
// = 4 x 4 mma.m8n8k4 tiles.  Operands are staged once (K = 64 fits) in padded shared memory:
// strides = 4 (mod 16) doubles make every fragment load conflict-free.
constexpr int GM_BM = LU_GEMM_BM, GM_BN = 64, GM_K = LU_NB;
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

// Second-generation trailing update.  A CTA owns one block of 128 rows: it stages L21 (128 x 64) in shared memory
// ONCE with cp.async and then walks over a run of 32-column tiles.  The CTA is split into two groups of four warps
// that work on alternate tiles with their own double-buffered U12 tiles and their own named barrier, so one group's
// load / store phase overlaps the other group's DMMA phase (with a single CTA-wide barrier per tile all eight warps
// issue their stores, address arithmetic and loads in lockstep and the tensor pipe idles meanwhile).  The U12 tile
// of a group's next step is fetched with cp.async and its next C tile is prefetched into registers while the DMMAs
// of the current tile run.  Same fragment layout and the same accumulation order over k as lu_gemm_dmma_kernel
// (bitwise the same result).
constexpr int G2_THREADS = 256;
constexpr int G2_BN = 32;                                   // columns per group tile
constexpr size_t g2_smem(int nh) { return (size_t)(nh * GM_K * GM_SA + 2 * 2 * G2_BN * GM_SB) * sizeof(double); }

__device__ __forceinline__ void cp_async16(void* smem_dst, const void* gsrc, bool pred) {
    const unsigned d = (unsigned)__cvta_generic_to_shared(smem_dst);
    const int sz = pred ? 16 : 0;   // src-size 0: the 16 bytes are zero-filled
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(d), "l"(gsrc), "r"(sz) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }
__device__ __forceinline__ void group_sync(int id) { asm volatile("bar.sync %0, 128;" ::"r"(id) : "memory"); }

template <int NH>   // NH 64-wide halves of K: 1 = rank-64 update, 2 = rank-128 update (two panels applied in one pass over C)
__global__ void __launch_bounds__(G2_THREADS, 1) lu_gemm2_kernel(const double* __restrict__ Lp, int ldl, const double* __restrict__ Up, int ldu,
                                                                 double* __restrict__ Cp, int ldc, int M, int Nc, int n_col_tiles,
                                                                 int tiles_per_cta, const unsigned char* __restrict__ row_block_active) {
    // C (M x Nc, ldc) -= L (M x 64 NH, ldl) * U (64 NH x Nc, ldu).  Single GPU: the three are windows of the same matrix; the
    // distributed factorisation passes the gathered multipliers of the local rows and the broadcast U12 block row, and
    // a byte per 128-row block that says whether any of its rows is still below the panel.
    // NH = 2: L (128 x 128) stays resident; each C tile takes two steps, one per 64-row half of its U tile, through the
    // same double buffer, and is read and written once - half the C traffic per flop of two rank-64 updates.
    extern __shared__ __align__(16) double smem[];
    if (row_block_active && !row_block_active[blockIdx.x]) return;
    constexpr int KT = GM_K * NH;
    double* sA = smem;                  // [KT][GM_SA]   L21 (negated when the fragments are read)
    const int tid = threadIdx.x, grp = tid >> 7, gtid = tid & 127, warp = gtid >> 5, lane = tid & 31;
    double* sB = smem + KT * GM_SA + grp * (2 * G2_BN * GM_SB);   // this group's [2][G2_BN][GM_SB] U12 half tiles
    const int row0 = blockIdx.x * GM_BM;
    const int ct0 = blockIdx.y * tiles_per_cta + grp, ct1 = min((int)(blockIdx.y + 1) * tiles_per_cta, n_col_tiles);
    const int wm = warp * 32;
    const int g = lane >> 2, q = lane & 3;

    for (int t = tid; t < KT * (GM_BM / 2); t += G2_THREADS) {
        const int k = t / (GM_BM / 2), m2 = (t % (GM_BM / 2)) * 2;
        const int r = row0 + m2;
        cp_async16(sA + k * GM_SA + m2, Lp + (size_t)k * ldl + min(r, M - 1), r < M);
    }
    cp_async_commit();
    cp_async_wait<0>();
    __syncthreads();
    if (ct0 >= ct1) return;
    auto load_u = [&](int buf, int ct, int half) {
        double* dst = sB + buf * (G2_BN * GM_SB);
        const int col0 = ct * G2_BN;
        for (int t = gtid; t < G2_BN * (GM_K / 2); t += 128) {
            const int nn = t / (GM_K / 2), k2 = (t % (GM_K / 2)) * 2;
            const int c = col0 + nn;
            cp_async16(dst + nn * GM_SB + k2, Up + (size_t)min(c, Nc - 1) * ldu + half * GM_K + k2, c < Nc);
        }
    };
    double cn[4][4][2];
    auto load_c = [&](int ct) {
        const int col0 = ct * G2_BN;
#pragma unroll
        for (int mt = 0; mt < 4; ++mt)
#pragma unroll
            for (int nt = 0; nt < 4; ++nt) {
                const int r = row0 + wm + mt * 8 + g;
                const int cc = col0 + nt * 8 + 2 * q;
                cn[mt][nt][0] = (r < M && cc < Nc) ? __ldcs(Cp + r + (size_t)cc * ldc) : 0.;
                cn[mt][nt][1] = (r < M && cc + 1 < Nc) ? __ldcs(Cp + r + (size_t)(cc + 1) * ldc) : 0.;
            }
    };
    load_u(0, ct0, 0);
    cp_async_commit();
    load_c(ct0);
    int buf = 0;
    double c[4][4][2];
    for (int ct = ct0; ct < ct1; ct += 2) {
#pragma unroll
        for (int mt = 0; mt < 4; ++mt)
#pragma unroll
            for (int nt = 0; nt < 4; ++nt) {
                c[mt][nt][0] = cn[mt][nt][0];
                c[mt][nt][1] = cn[mt][nt][1];
            }
#pragma unroll
        for (int half = 0; half < NH; ++half, buf ^= 1) {
            // fetch the next step's U half tile (and, on the last half, the next C tile) before computing this one
            const bool last_half = (half == NH - 1);
            const bool more = !last_half || (ct + 2 < ct1);
            if (more) {
                load_u(buf ^ 1, last_half ? ct + 2 : ct, last_half ? 0 : half + 1);
                cp_async_commit();
                if (last_half) load_c(ct + 2);
                cp_async_wait<1>();
            } else {
                cp_async_wait<0>();
            }
            group_sync(1 + grp);
            const double* sBb = sB + buf * (G2_BN * GM_SB);
            const double* sAh = sA + half * GM_K * GM_SA;
#pragma unroll 4
            for (int ks = 0; ks < GM_K; ks += 4) {
                double a[4], b[4];
#pragma unroll
                for (int mt = 0; mt < 4; ++mt) a[mt] = -sAh[(ks + q) * GM_SA + wm + mt * 8 + g];
#pragma unroll
                for (int nt = 0; nt < 4; ++nt) b[nt] = sBb[(nt * 8 + g) * GM_SB + ks + q];
#pragma unroll
                for (int mt = 0; mt < 4; ++mt)
#pragma unroll
                    for (int nt = 0; nt < 4; ++nt) dmma_m8n8k4(c[mt][nt][0], c[mt][nt][1], a[mt], b[nt]);
            }
            if (last_half) {
                const int col0 = ct * G2_BN;
#pragma unroll
                for (int mt = 0; mt < 4; ++mt)
#pragma unroll
                    for (int nt = 0; nt < 4; ++nt) {
                        const int r = row0 + wm + mt * 8 + g;
                        const int cc = col0 + nt * 8 + 2 * q;
                        if (r < M && cc < Nc) Cp[r + (size_t)cc * ldc] = c[mt][nt][0];
                        if (r < M && cc + 1 < Nc) Cp[r + (size_t)(cc + 1) * ldc] = c[mt][nt][1];
                    }
            }
            group_sync(1 + grp);   // every warp of the group is done with sB[buf] before the next step refills it
        }
    }
}

// Grid shape of the trailing update: rb row blocks x `chunks` runs of 32-column tiles.  Picks the run length that
// minimises (number of waves over the SMs) x (tiles per run + the cost of staging L21, ~3 tiles), so the last wave
// is not mostly empty (82 row blocks x 4 runs = 2.2 waves wasted a quarter of the launch).
static void lu_gemm2_shape(int rb, int ctiles, int num_sms, int k_halves, int* per_out, int* chunks_out) {
    long long best = -1;
    int best_per = ctiles, best_chunks = 1;
    for (int chunks = 1; chunks <= ctiles; ++chunks) {
        int per = (ctiles + chunks - 1) / chunks;
        per = (per + 1) & ~1;   // both warp groups get the same number of tiles
        const int nch = (ctiles + per - 1) / per;
        const long long waves = ((long long)rb * nch + num_sms - 1) / num_sms;
        const long long cost = waves * (per * k_halves + 6 * k_halves);
        if (best < 0 || cost < best) { best = cost; best_per = per; best_chunks = nch; }
        if (per <= 8) break;
    }
    *per_out = best_per;
    *chunks_out = best_chunks;
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

// x = P b: the interchanges of the factorisation (linalg.f90:311-316 "untangle pivoting") composed into one gather
__global__ void lu_permute_kernel(const double* __restrict__ b, const int* __restrict__ perm, int n, double* __restrict__ x) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) x[i] = b[perm[i]];
}

// sum_c a[c * ld] * sx[c] in the order c = 0, 1, ... (one FMA chain, as before), with a full 64-column strip requested
// together: the entries come from HBM (the factors are larger than L2), and a rolled loop kept only a few of them in flight
// (~10 us of exposed latency per step, 2 N / 64 steps per solve).
// The whole strip of a row is requested from HBM at once (prefetch to L2: no registers held), BEFORE the block barrier that
// precedes lu_strip_dot: ptxas keeps a window of only ~10 loads of the FMA chain in flight whatever the source says (and sinks
// prefetches into the chain when they are issued next to it); after the barrier the loads find their lines in L2.
__device__ __forceinline__ void lu_strip_prefetch(const double* __restrict__ a, int ld, int nb) {
    if (nb == LU_NB) {
#pragma unroll
        for (int c = 0; c < LU_NB; ++c) asm volatile("prefetch.global.L2 [%0];" ::"l"(a + (size_t)c * ld));
    }
}
__device__ __forceinline__ double lu_strip_dot(const double* __restrict__ a, int ld, int nb, const double* __restrict__ sx) {
    double acc = 0.;
    if (nb == LU_NB) {
#pragma unroll
        for (int c = 0; c < LU_NB; ++c) acc = fma(__ldg(a + (size_t)c * ld), sx[c], acc);
    } else {
        for (int c = 0; c < nb; ++c) acc = fma(a[(size_t)c * ld], sx[c], acc);
    }
    return acc;
}

// One step of the blocked forward substitution in ONE launch: x[k0..k1) is final; every row below gets
// x[r] -= L(r, k0..k1) . x[k0..k1), and the CTA that holds the next diagonal block solves it straight away (unit lower
// triangular, one warp, column order), so a step costs one launch instead of two.  Same operation order per element
// as a separate update of the rows followed by lu_trsv_diag_kernel on the block.
__global__ void __launch_bounds__(256) lu_fwd_step_kernel(const double* __restrict__ A, int ld, int n, int k0, int k1, double* __restrict__ x) {
    __shared__ double sx[LU_NB];
    __shared__ double sy[LU_NB];
    __shared__ double sL[LU_NB * LU_NB];
    const int nb = k1 - k0, tid = threadIdx.x;
    if (tid < nb) sx[tid] = x[k0 + tid];
    const int nb2 = min(LU_NB, n - k1);
    if (blockIdx.x == 0) {   // the next diagonal block does not depend on x: in flight together with the strip
        for (int t = tid; t < nb2 * nb2; t += 256) {
            const int rr = t % nb2, cc = t / nb2;
            sL[cc * LU_NB + rr] = A[(k1 + rr) + (size_t)(k1 + cc) * ld];
        }
    }
    const int r = k1 + blockIdx.x * 256 + tid;
    if (r < n) lu_strip_prefetch(A + r + (size_t)k0 * ld, ld, nb);
    __syncthreads();
    double xr = 0.;
    if (r < n) xr = x[r] - lu_strip_dot(A + r + (size_t)k0 * ld, ld, nb, sx);
    if (r < n && (blockIdx.x != 0 || tid >= LU_NB)) x[r] = xr;
    if (blockIdx.x != 0) return;
    if (tid < nb2) sy[tid] = xr;
    __syncthreads();
    if (tid < 32) {
        // one warp, rows (lane, lane + 32) in registers, x[c] broadcast by shuffle: same order per row as the column loop over
        // shared memory it replaces (a shared-memory round trip and two warp barriers per column)
        const int lane = tid;
        double y0 = lane < nb2 ? sy[lane] : 0., y1 = lane + 32 < nb2 ? sy[lane + 32] : 0.;
        const int n_lo = min(32, nb2);
#pragma unroll 4
        for (int c = 0; c < n_lo; ++c) {
            const double xc = __shfl_sync(0xffffffffu, y0, c);
            if (lane > c && lane < nb2) y0 = fma(-sL[c * LU_NB + lane], xc, y0);
            if (lane + 32 < nb2) y1 = fma(-sL[c * LU_NB + lane + 32], xc, y1);
        }
#pragma unroll 4
        for (int c = 32; c < nb2; ++c) {
            const double xc = __shfl_sync(0xffffffffu, y1, c - 32);
            if (lane + 32 > c && lane + 32 < nb2) y1 = fma(-sL[c * LU_NB + lane + 32], xc, y1);
        }
        if (lane < nb2) x[k1 + lane] = y0;
        if (lane + 32 < nb2) x[k1 + lane + 32] = y1;
    }
}

// One step of the blocked back substitution: x[k0..k1) is final; rows above get x[r] -= U(r, k0..k1) . x[k0..k1), and the
// CTA holding the diagonal block just above (rows k0-64..k0) solves it (upper triangular, divisions by the diagonal).
__global__ void __launch_bounds__(256) lu_bwd_step_kernel(const double* __restrict__ A, int ld, int k0, int k1, double* __restrict__ x) {
    __shared__ double sx[LU_NB];
    __shared__ double sy[LU_NB];
    __shared__ double sU[LU_NB * LU_NB];
    const int nb = k1 - k0, tid = threadIdx.x;
    if (tid < nb) sx[tid] = x[k0 + tid];
    const int b0 = k0 - LU_NB;   // k0 is a multiple of LU_NB and > 0
    if (blockIdx.x == 0) {       // the diagonal block just above does not depend on x: in flight together with the strip
        for (int t = tid; t < LU_NB * LU_NB; t += 256) {
            const int rr = t % LU_NB, cc = t / LU_NB;
            sU[cc * LU_NB + rr] = A[(b0 + rr) + (size_t)(b0 + cc) * ld];
        }
    }
    // CTA 0 covers rows [k0-256, k0) with its first 64 threads on the block just above the solved one
    const int r = k0 - 1 - (blockIdx.x * 256 + tid);
    if (r >= 0) lu_strip_prefetch(A + r + (size_t)k0 * ld, ld, nb);
    __syncthreads();
    double xr = 0.;
    if (r >= 0) xr = x[r] - lu_strip_dot(A + r + (size_t)k0 * ld, ld, nb, sx);
    if (r >= 0 && (blockIdx.x != 0 || tid >= LU_NB)) x[r] = xr;
    if (blockIdx.x != 0) return;
    if (tid < LU_NB) sy[LU_NB - 1 - tid] = xr;   // thread t holds row k0-1-t
    __syncthreads();
    if (tid < 32) {
        // one warp, rows (lane, lane + 32) in registers, x[c] = y[c] / U(c, c) broadcast by shuffle (see lu_fwd_step_kernel)
        const int lane = tid;
        double y0 = sy[lane], y1 = sy[lane + 32];
#pragma unroll 4
        for (int c = LU_NB - 1; c >= 32; --c) {
            if (lane == c - 32) y1 = y1 / sU[c * LU_NB + c];
            const double xc = __shfl_sync(0xffffffffu, y1, c - 32);
            y0 = fma(-sU[c * LU_NB + lane], xc, y0);
            if (lane + 32 < c) y1 = fma(-sU[c * LU_NB + lane + 32], xc, y1);
        }
#pragma unroll 4
        for (int c = 31; c >= 0; --c) {
            if (lane == c) y0 = y0 / sU[c * LU_NB + c];
            const double xc = __shfl_sync(0xffffffffu, y0, c);
            if (lane < c) y0 = fma(-sU[c * LU_NB + lane], xc, y0);
        }
        x[b0 + lane] = y0;
        x[b0 + lane + 32] = y1;
    }
}

// C (M x Nc) -= L (M x 64) * U (64 x Nc) on the FP64 tensor cores; all leading dimensions even, pointers 16-byte aligned
void lu_gemm2_launch(Ctx* c, const double* Lp, int ldl, const double* Up, int ldu, double* Cp, int ldc, int M, int Nc,
                     const unsigned char* row_block_active, int k_halves, cudaStream_t stream) {
    if (!stream) stream = c->stream;
    if (!(c->attr_mask & 2u)) {   // per device: remembered per context
        cudaError_t e1 = cudaFuncSetAttribute(lu_gemm2_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)g2_smem(1));
        cudaError_t e2 = cudaFuncSetAttribute(lu_gemm2_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)g2_smem(2));
        if (e1 == cudaSuccess && e2 == cudaSuccess) c->attr_mask |= 2u;
        else c->err = std::string("lu_gemm2 shared-memory opt-in: ") + cudaGetErrorString(e1 != cudaSuccess ? e1 : e2);   // the launch below then fails and is reported
    }
    const int rb = (M + GM_BM - 1) / GM_BM, ct32 = (Nc + G2_BN - 1) / G2_BN;
    int per, chunks;
    lu_gemm2_shape(rb, ct32, c->num_sms, k_halves, &per, &chunks);
    if (k_halves == 2)
        lu_gemm2_kernel<2><<<dim3(rb, chunks), G2_THREADS, g2_smem(2), stream>>>(Lp, ldl, Up, ldu, Cp, ldc, M, Nc, ct32, per, row_block_active);
    else
        lu_gemm2_kernel<1><<<dim3(rb, chunks), G2_THREADS, g2_smem(1), stream>>>(Lp, ldl, Up, ldu, Cp, ldc, M, Nc, ct32, per, row_block_active);
    c->launches += 1;
}

static size_t lu_panel_smem(int rpc, bool cluster = false) {
    const int cap = rpc <= LUP_CAP ? rpc : LUP_CAP_OVF;
    size_t b = (size_t)(LU_NB * (cap | 1) + 2 * LUP_ROW + 2 * rpc + 32) * sizeof(double) + (size_t)(34 + ((rpc + 1) & ~1)) * sizeof(int);
    if (cluster) b += (size_t)(2 * LUP_CL_MAX * LUP_ROW + 2 * LUP_ROW + 2 * LUP_CL_MAX) * sizeof(double) + (size_t)2 * LUP_CL_MAX * sizeof(int);
    return b;
}

ml_status LuPanelWork::init(Ctx* c) {
    if (!(c->attr_mask & 4u)) {
        ML_CUDA(c, cudaFuncSetAttribute(lu_panel_coop_kernel<false, LUP_THREADS>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                        (int)lu_panel_smem(LUP_CAP)));
        ML_CUDA(c, cudaFuncSetAttribute(lu_panel_coop_kernel<true, LUP_THREADS>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                        (int)lu_panel_smem(LUP_RPC_MAX)));
        c->attr_mask |= 4u;
    }
    gmax = c->num_sms;
    ML_CUDA(c, pscr.alloc((size_t)2 * gmax * (LUP_ROW + 1) + 2 * LUP_ROW));
    ML_CUDA(c, pidx.alloc((size_t)2 * gmax));
    ML_CUDA(c, pbar.alloc(1));
    ML_CUDA(c, cudaMemsetAsync(pbar.p, 0, sizeof(unsigned), c->stream));
    bar_base = 0;
    all_coop = true;
    if (getenv("MACHLINE_LU_PANEL_DBG")) {
        ML_CUDA(c, cudaMalloc(&dbg, 8 * sizeof(long long)));
        ML_CUDA(c, cudaMemset(dbg, 0, 8 * sizeof(long long)));
    }
    return ML_OK;
}
void LuPanelWork::release() {
    if (dbg) {   // cycles per phase of the cluster panel kernel, per column (thread 0 of CTA 0)
        long long h[8];
        cudaDeviceSynchronize();
        cudaMemcpy(h, dbg, sizeof(h), cudaMemcpyDeviceToHost);
        const double nc = h[7] > 0 ? (double)h[7] : 1.;
        fprintf(stderr, "lu_panel_cl2 cycles per column over %lld columns: candidate %.0f | park+push %.0f | cluster barrier %.0f | pivot+pull %.0f | "
                        "interchange %.0f | eliminate %.0f | block update %.0f\n",
                h[7], h[0] / nc, h[1] / nc, h[2] / nc, h[3] / nc, h[4] / nc, h[5] / nc, h[6] / nc);
        cudaFree(dbg);
        dbg = nullptr;
    }
    pscr.release();
    pidx.release();
    pbar.release();
}

void lu_launch_row_amax(Ctx* c, const double* A, int ld, int nr, int nc, double* amax) {
    if (nr <= 0) return;
    lu_row_amax_kernel<<<dim3((nr + 127) / 128, (nc + RS_COLS - 1) / RS_COLS), 128, 0, c->stream>>>(A, ld, nr, nc, (unsigned long long*)amax);
    c->launches += 1;
}
void lu_launch_trsm(Ctx* c, const double* L, int ldl, double* X, int ldx, int ncols) {
    if (ncols <= 0) return;
    lu_trsm_kernel<<<(ncols + TRSM_COLS_PER_CTA - 1) / TRSM_COLS_PER_CTA, TRSM_THREADS, 0, c->stream>>>(L, ldl, X, ldx, ncols);
    c->launches += 1;
}
void lu_launch_trsv_diag(Ctx* c, const double* D, int ld, int nb, double* x, int upper) {
    lu_trsv_diag_kernel<<<1, 64, 0, c->stream>>>(D, ld, 0, nb, x, upper);
    c->launches += 1;
}

// Factor the panel (rows k0..n, columns k0..k1 of dA): pivots into d_piv[k0..k1), rows interchanged inside the panel.
// One cooperative launch when the panel's rows fit the CTAs' shared memory, else two launches per column.
ml_status lu_panel_factor(Ctx* c, LuPanelWork& W, double* dA, int ld, int n, int k0, int k1, double* d_vv, int* d_piv, int* d_perm,
                          cudaStream_t stream, int max_ctas) {
    static const bool per_column = getenv("MACHLINE_LU_PER_COLUMN") != nullptr;   // the unfused path, kept for A/B timing
    const int m = n - k0;
    static const char* rpc_env = getenv("MACHLINE_LU_PANEL_RPC");   // tests: force a rows-per-CTA value (e.g. the overflow variant)
    // Cluster variant (one thread-block cluster of <= 16 CTAs, candidates through distributed shared memory, hardware cluster
    // barrier per column): whenever the panel's rows fit 16 CTAs' shared memory.  MACHLINE_LU_NO_CLUSTER=1 disables it.
    static const bool no_cluster = getenv("MACHLINE_LU_NO_CLUSTER") != nullptr;
    // Second-generation cluster kernel (thread = row, delayed updates, pulled pivot rows): panels of up to 16 x 512 rows.
    // MACHLINE_LU_PANEL_V1=1 keeps the first cluster kernel (A/B timing, bitwise comparison of the factors).
    static const bool panel_v1 = getenv("MACHLINE_LU_PANEL_V1") != nullptr;
    if (!per_column && !no_cluster && !rpc_env && !panel_v1 && m <= LUP_CL_MAX * LUP2_T) {
        int G = 1;
        while (G < LUP_CL_MAX && G * 64 < m) G *= 2;
        // rows per CTA: exactly ceil(m / G) (nothing in the kernel needs a multiple of 32), so the last CTA always owns rows
        const int rpc_c = (m + G - 1) / G;
        if (rpc_c <= LUP2_T && (long long)rpc_c * (G - 1) < m) {
            if (!(c->attr_mask & 64u)) {
                ML_CUDA(c, cudaFuncSetAttribute(lu_panel_cl2_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)lu_panel2_smem(LUP2_CAP)));
                ML_CUDA(c, cudaFuncSetAttribute(lu_panel_cl2_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)lu_panel2_smem(LUP2_T)));
                ML_CUDA(c, cudaFuncSetAttribute(lu_panel_cl2_kernel<false>, cudaFuncAttributeNonPortableClusterSizeAllowed, 1));
                ML_CUDA(c, cudaFuncSetAttribute(lu_panel_cl2_kernel<true>, cudaFuncAttributeNonPortableClusterSizeAllowed, 1));
                c->attr_mask |= 64u;
            }
            LuPanelArgs pa;
            pa.A = dA; pa.ld = ld; pa.n = n; pa.k0 = k0; pa.k1 = k1; pa.rpc = rpc_c;
            pa.vv = d_vv; pa.piv = d_piv; pa.perm = d_perm;
            pa.cand_v = nullptr; pa.rowj = nullptr; pa.cand_i = nullptr; pa.bar = nullptr; pa.bar_base = 0;
            pa.cand_row = reinterpret_cast<double*>(W.dbg);   // phase stamps (MACHLINE_LU_PANEL_DBG), else null
            cudaLaunchConfig_t cfg = {};
            cfg.gridDim = dim3(G);
            cfg.blockDim = dim3(std::max(128, (rpc_c + 31) & ~31));   // thread = row; at least the 66 threads of the row slots
            cfg.dynamicSmemBytes = lu_panel2_smem(rpc_c);
            cfg.stream = stream;
            cudaLaunchAttribute at[1];
            at[0].id = cudaLaunchAttributeClusterDimension;
            at[0].val.clusterDim.x = G;
            at[0].val.clusterDim.y = 1;
            at[0].val.clusterDim.z = 1;
            cfg.attrs = at;
            cfg.numAttrs = 1;
            if (rpc_c > LUP2_CAP) ML_CUDA(c, cudaLaunchKernelEx(&cfg, lu_panel_cl2_kernel<true>, pa));
            else ML_CUDA(c, cudaLaunchKernelEx(&cfg, lu_panel_cl2_kernel<false>, pa));
            c->launches += 1;
            return ML_OK;
        }
    }
    if (!per_column && !no_cluster && !rpc_env && m <= LUP_CL_MAX * LUP_CAP) {
        int G = 1;
        while (G < LUP_CL_MAX && G * 128 < m) G *= 2;
        while (G < LUP_CL_MAX && (m + G - 1) / G > LUP_CAP) G *= 2;
        int rpc_c = (((m + G - 1) / G) + 31) & ~31;
        while (G > 1 && (long long)rpc_c * (G - 1) >= m) {           // every CTA of the cluster owns at least one row
            G /= 2;
            rpc_c = (((m + G - 1) / G) + 31) & ~31;
        }
        if (rpc_c <= LUP_CAP) {
            auto kern = lu_panel_coop_kernel<false, LUP_THREADS, true>;
            const size_t smem = lu_panel_smem(rpc_c, true);
            if (!(c->attr_mask & 32u)) {
                ML_CUDA(c, cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)lu_panel_smem(LUP_CAP, true)));
                ML_CUDA(c, cudaFuncSetAttribute(kern, cudaFuncAttributeNonPortableClusterSizeAllowed, 1));
                c->attr_mask |= 32u;
            }
            LuPanelArgs pa;
            pa.A = dA; pa.ld = ld; pa.n = n; pa.k0 = k0; pa.k1 = k1; pa.rpc = rpc_c;
            pa.vv = d_vv; pa.piv = d_piv; pa.perm = d_perm;
            pa.cand_v = nullptr; pa.cand_row = nullptr; pa.rowj = nullptr; pa.cand_i = nullptr; pa.bar = nullptr; pa.bar_base = 0;
            cudaLaunchConfig_t cfg = {};
            cfg.gridDim = dim3(G);
            cfg.blockDim = dim3(LUP_THREADS);
            cfg.dynamicSmemBytes = smem;
            cfg.stream = stream;
            cudaLaunchAttribute at[1];
            at[0].id = cudaLaunchAttributeClusterDimension;
            at[0].val.clusterDim.x = G;
            at[0].val.clusterDim.y = 1;
            at[0].val.clusterDim.z = 1;
            cfg.attrs = at;
            cfg.numAttrs = 1;
            ML_CUDA(c, cudaLaunchKernelEx(&cfg, kern, pa));
            c->launches += 1;
            return ML_OK;
        }
    }
    // max_ctas > 0 (look-ahead): the panel runs on a few SMs next to the trailing update of the previous pair
    const int gcap = (max_ctas > 0) ? std::min(max_ctas, W.gmax) : W.gmax;
    int rpc = 128;
    if ((long long)rpc * gcap < m) rpc = (((m + gcap - 1) / gcap) + 31) & ~31;
    if (rpc_env && atoi(rpc_env) >= 128) rpc = std::max(rpc, (atoi(rpc_env) + 31) & ~31);
    if (!per_column && rpc <= LUP_RPC_MAX) {
        LuPanelArgs pa;
        pa.A = dA; pa.ld = ld; pa.n = n; pa.k0 = k0; pa.k1 = k1; pa.rpc = rpc;
        pa.vv = d_vv; pa.piv = d_piv; pa.perm = d_perm;
        pa.cand_v = W.pscr.p;
        pa.cand_row = W.pscr.p + 2 * W.gmax;
        pa.rowj = pa.cand_row + (size_t)2 * W.gmax * LUP_ROW;
        pa.cand_i = W.pidx.p;
        pa.bar = W.pbar.p; pa.bar_base = W.bar_base;
        const int G = (m + rpc - 1) / rpc;
        W.bar_base += (unsigned)(k1 - k0) * (unsigned)G;
        void* kargs[] = {(void*)&pa};
        const void* kern = rpc <= LUP_CAP ? (const void*)lu_panel_coop_kernel<false, LUP_THREADS> : (const void*)lu_panel_coop_kernel<true, LUP_THREADS>;
        // A cooperative launch waits until the whole grid fits at once, i.e. until a concurrent trailing update has drained.
        // With a capped grid an ordinary launch is enough: the CTAs spin on the panel's own barrier, the update's CTAs do not
        // depend on them and retire, and the freed SMs go to this (higher-priority) kernel first: no deadlock.
        if (max_ctas > 0) ML_CUDA(c, cudaLaunchKernel(kern, dim3(G), dim3(LUP_THREADS), kargs, lu_panel_smem(rpc), stream));
        else ML_CUDA(c, cudaLaunchCooperativeKernel(kern, dim3(G), dim3(LUP_THREADS), kargs, lu_panel_smem(rpc), stream));
        c->launches += 1;
    } else {
        W.all_coop = false;
        for (int j = k0; j < k1; ++j) {
            lu_pivot_kernel<<<1, 1024, 0, stream>>>(dA, ld, n, j, k0, k1, d_vv, d_piv);
            if (j + 1 < n) lu_panel_update_kernel<<<(n - j - 1 + 255) / 256, 256, 0, stream>>>(dA, ld, n, j, k1);
            c->launches += 2;
        }
    }
    return ML_OK;
}

// ---- host drivers ---------------------------------------------------------------------------------------
// In-place LU of the n x n matrix at dA (leading dimension ld).  piv / vv are device work arrays of length n.
static ml_status lu_factor(Ctx* c, double* dA, int ld, int n, int* d_piv, double* d_vv, int* d_flag) {
    const size_t gemm_smem = (size_t)(GM_K * GM_SA + GM_BN * GM_SB) * sizeof(double);
    if (!(c->attr_mask & 8u)) {
        ML_CUDA(c, cudaFuncSetAttribute(lu_gemm_dmma_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)gemm_smem));
        c->attr_mask |= 8u;
    }
    ML_CUDA(c, cudaMemsetAsync(d_flag, 0, sizeof(int), c->stream));
    int* d_perm = d_piv + n;   // d_piv holds 2 n ints: the interchanges, then the row permutation they compose to
    ML_CUDA(c, cudaMemsetAsync(d_vv, 0, (size_t)n * sizeof(double), c->stream));
    lu_row_amax_kernel<<<dim3((n + 127) / 128, (n + RS_COLS - 1) / RS_COLS), 128, 0, c->stream>>>(dA, ld, n, n, (unsigned long long*)d_vv);
    lu_row_scale_finish_kernel<<<(n + 255) / 256, 256, 0, c->stream>>>(d_vv, n, d_flag, d_perm);
    c->launches += 2;
    int flag = 0;
    ML_CUDA(c, cudaMemcpyAsync(&flag, d_flag, sizeof(int), cudaMemcpyDeviceToHost, c->stream));
    ML_CUDA(c, cudaStreamSynchronize(c->stream));
    if (flag) return c->fail(ML_SINGULAR, "lu_decomp: the matrix is singular (a row is zero; linalg.f90:205-208)");
    LuPanelWork PW;
    ml_status pst = PW.init(c);
    if (pst != ML_OK) return pst;
    static const bool old_gemm = getenv("MACHLINE_LU_GEMM_V1") != nullptr;
    static const bool no_k128 = getenv("MACHLINE_LU_NO_K128") != nullptr;
    const bool fast_gemm = !old_gemm && (ld & 1) == 0;
    cudaStream_t S0 = c->stream;
    // Look-ahead: the next pair of panels is factored on a second, higher-priority stream (on a few SMs) while the rank-128
    // update of the current pair still runs on the rest of the matrix.  MACHLINE_LU_LOOKAHEAD=0 disables it, =<n> caps the panel
    // kernel at n CTAs (default 32).
    int la_ctas = 32;
    if (const char* e = getenv("MACHLINE_LU_LOOKAHEAD")) la_ctas = atoi(e);
    const bool lookahead = fast_gemm && !no_k128 && la_ctas > 0;
    if (lookahead && !c->stream_hi) {
        int lo = 0, hi = 0;
        ML_CUDA(c, cudaDeviceGetStreamPriorityRange(&lo, &hi));
        ML_CUDA(c, cudaStreamCreateWithPriority(&c->stream_hi, cudaStreamNonBlocking, hi));
        ML_CUDA(c, cudaEventCreateWithFlags(&c->ev_la[0], cudaEventDisableTiming));
        ML_CUDA(c, cudaEventCreateWithFlags(&c->ev_la[1], cudaEventDisableTiming));
    }
    cudaStream_t S1 = lookahead ? c->stream_hi : S0;
    // interchanges of panel [k0, k1) applied to columns [c0, c1) and [c2, c3)
    auto laswp2 = [&](cudaStream_t st, int c0, int c1, int c2, int c3, int k0, int k1) {
        const int nc = std::max(0, c1 - c0) + std::max(0, c3 - c2);
        if (nc > 0) {
            lu_laswp_kernel<<<(nc + 255) / 256, 256, 0, st>>>(dA, ld, c0, std::max(c0, c1), c2, std::max(c2, c3), k0, k1, d_piv);
            c->launches += 1;
        }
    };
    auto laswp = [&](cudaStream_t st, int c0, int c1, int k0, int k1) { laswp2(st, c0, c1, c1, c1, k0, k1); };
    auto trsm = [&](cudaStream_t st, int k0, int c0, int c1) {            // U(k0..k0+64, c0..c1) = L11^-1 A(k0..k0+64, c0..c1)
        if (c1 > c0) {
            lu_trsm_kernel<<<(c1 - c0 + TRSM_COLS_PER_CTA - 1) / TRSM_COLS_PER_CTA, TRSM_THREADS, 0, st>>>(dA + k0 + (size_t)k0 * ld, ld,
                                                                                                        dA + k0 + (size_t)c0 * ld, ld, c1 - c0);
            c->launches += 1;
        }
    };
    // A(r0..r1, c0..c1) -= A(r0..r1, k0..k0+64 kh) A(k0..k0+64 kh, c0..c1)
    auto gemm = [&](cudaStream_t st, int r0, int r1, int c0, int c1, int k0, int kh) {
        if (r1 <= r0 || c1 <= c0) return;
        if (fast_gemm) {
            lu_gemm2_launch(c, dA + r0 + (size_t)k0 * ld, ld, dA + k0 + (size_t)c0 * ld, ld, dA + r0 + (size_t)c0 * ld, ld, r1 - r0, c1 - c0, nullptr, kh,
                            st);
        } else {   // first-generation kernel: whole trailing matrix of one panel (r0 == c0 == k0 + 64, r1 == c1 == n)
            dim3 grid((n - r0 + GM_BM - 1) / GM_BM, (n - r0 + GM_BN - 1) / GM_BN);
            lu_gemm_dmma_kernel<<<grid, 256, gemm_smem, st>>>(dA, ld, n, k0, k0 + LU_NB);
            c->launches += 1;
        }
    };
    // Panels A = [a0, a0 + 64) and B = [a0 + 64, a0 + 128), whose columns have seen every earlier panel: factor A, apply it to B's
    // 64 columns, factor B.  On `st` with at most `cap` CTAs per panel kernel (0: the whole grid, cooperative).
    auto factor_pair = [&](cudaStream_t st, int a0, int cap) -> ml_status {
        const int a1 = a0 + LU_NB, a2 = a1 + LU_NB;
        ml_status ps = lu_panel_factor(c, PW, dA, ld, n, a0, a1, d_vv, d_piv, d_perm, st, cap);
        if (ps != ML_OK) return ps;
        laswp(st, a1, a2, a0, a1);
        trsm(st, a0, a1, a2);
        gemm(st, a1, n, a1, a2, a0, 1);
        return lu_panel_factor(c, PW, dA, ld, n, a1, a2, d_vv, d_piv, d_perm, st, cap);
    };
    int k0 = 0;
    bool have_panel = false;   // panel [k0, k0 + 64) is factored and nothing to its right has seen it yet
    if (fast_gemm && !no_k128 && n > 2 * LU_NB) {
        // ---- pairs of panels, one rank-128 update per pair (C is read and written once for 256 flop per entry) ----
        pst = factor_pair(S0, 0, 0);
        if (pst != ML_OK) { PW.release(); return pst; }
        for (;;) {   // invariant: A = [k0, k1), B = [k1, k2) factored (B has seen A); k2 < n; nothing outside the pair has seen them
            const int k1 = k0 + LU_NB, k2 = k1 + LU_NB;
            // (one fused launch for the four was tried, r5d / r5e: 0.7 ms less at N = 7.4k without interchanges, 9-13 ms more with
            // an interchange in every column)
            laswp2(S0, 0, k0, k2, n, k0, k1);     // A's interchanges, left and right of the pair in one launch
            laswp2(S0, 0, k1, k2, n, k1, k2);     // B's: the left part includes A's own columns (its multipliers move with the rows)
            trsm(S0, k0, k2, n);                  // U rows of A
            gemm(S0, k1, k2, k2, n, k0, 1);       // rows of B's diagonal block: minus L21_A U_A
            trsm(S0, k1, k2, n);                  // U rows of B
            const bool next_pair = (n - k2) > 2 * LU_NB;   // another complete pair with something after it
            if (!next_pair) {
                gemm(S0, k2, n, k2, n, k0, 2);    // everything below: minus [L_A L_B] [U_A; U_B]
                k0 = k2;
                break;
            }
            const int k4 = k2 + 2 * LU_NB;
            gemm(S0, k2, n, k2, k4, k0, 2);       // the next pair's columns first ...
            if (lookahead) {
                ML_CUDA(c, cudaEventRecord(c->ev_la[0], S0));
                ML_CUDA(c, cudaStreamWaitEvent(S1, c->ev_la[0], 0));
            }
            pst = factor_pair(S1, k2, lookahead ? la_ctas : 0);   // ... so that its panels are factored under the rest of the update
            if (pst != ML_OK) { PW.release(); return pst; }
            gemm(S0, k2, n, k4, n, k0, 2);
            if (lookahead) {
                ML_CUDA(c, cudaEventRecord(c->ev_la[1], S1));
                ML_CUDA(c, cudaStreamWaitEvent(S0, c->ev_la[1], 0));
            }
            k0 = k2;
            ML_CUDA(c, cudaGetLastError());
        }
    }
    // ---- single panels: the remaining columns (or the whole matrix on the fallback paths) ----
    if (k0 < n) {
        pst = lu_panel_factor(c, PW, dA, ld, n, k0, std::min(k0 + LU_NB, n), d_vv, d_piv, d_perm, S0);
        if (pst != ML_OK) { PW.release(); return pst; }
        have_panel = true;
    }
    while (have_panel && k0 < n) {   // invariant: panel [k0, k0 + 64) is factored, nothing to its right has seen it yet
        const int k1 = std::min(k0 + LU_NB, n), k2 = std::min(k1 + LU_NB, n);
        laswp2(S0, 0, k0, k1, n, k0, k1);
        if (k1 < n) {
            trsm(S0, k0, k1, n);
            gemm(S0, k1, n, k1, n, k0, 1);
            pst = lu_panel_factor(c, PW, dA, ld, n, k1, k2, d_vv, d_piv, d_perm, S0);
            if (pst != ML_OK) { PW.release(); return pst; }
        }
        k0 = k1;
        ML_CUDA(c, cudaGetLastError());
    }
    if (!PW.all_coop) {
        lu_perm_from_piv_kernel<<<1, 32, 0, c->stream>>>(d_piv, n, d_perm);
        c->launches += 1;
    }
    PW.release();
    return ML_OK;
}

// x = U^{-1} L^{-1} P b with the factors in dA; d_piv holds the interchanges and, after them, the permutation
static ml_status lu_substitute(Ctx* c, const double* dA, int ld, int n, const int* d_piv, const double* d_b, double* d_x) {
    lu_permute_kernel<<<(n + 255) / 256, 256, 0, c->stream>>>(d_b, d_piv + n, n, d_x);
    lu_trsv_diag_kernel<<<1, 64, 0, c->stream>>>(dA, ld, 0, std::min(LU_NB, n), d_x, 0);
    c->launches += 2;
    for (int k0 = 0; k0 + LU_NB < n; k0 += LU_NB) {
        const int k1 = k0 + LU_NB;
        lu_fwd_step_kernel<<<(n - k1 + 255) / 256, 256, 0, c->stream>>>(dA, ld, n, k0, k1, d_x);
        c->launches += 1;
    }
    const int nblk = (n + LU_NB - 1) / LU_NB;
    {
        const int k0 = (nblk - 1) * LU_NB;
        lu_trsv_diag_kernel<<<1, 64, 0, c->stream>>>(dA, ld, k0, n, d_x, 1);
        c->launches += 1;
    }
    for (int b = nblk - 1; b >= 1; --b) {
        const int k0 = b * LU_NB, k1 = std::min(k0 + LU_NB, n);
        lu_bwd_step_kernel<<<(k0 + 255) / 256, 256, 0, c->stream>>>(dA, ld, k0, k1, d_x);
        c->launches += 1;
    }
    ML_CUDA(c, cudaGetLastError());
    return ML_OK;
}

// lu_solve (linalg.f90:118-148): dA is overwritten by its factors
ml_status lu_solve_device(Ctx* c, int N, double* dA, int ld, const double* d_b, double* d_x) {
    DevBuf<int> piv, flag;
    DevBuf<double> vv;
    ML_CUDA(c, piv.alloc(2 * (size_t)N));
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

// out = || a - b ||_2   (||dx|| of the iteration history, linalg.f90:714, 585)
__global__ void __launch_bounds__(1024) vec_diff_norm_kernel(const double* __restrict__ a, const double* __restrict__ b, int n,
                                                              double* __restrict__ out) {
    __shared__ double s_part[32];
    double acc = 0.;
    for (int i = threadIdx.x; i < n; i += 1024) {
        const double d = a[i] - b[i];
        acc = fma(d, d, acc);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
    if ((threadIdx.x & 31) == 0) s_part[threadIdx.x >> 5] = acc;
    __syncthreads();
    if (threadIdx.x == 0) {
        double t = 0.;
        for (int k = 0; k < 32; ++k) t += s_part[k];
        *out = sqrt(t);
    }
}

// Iteration history of the block solvers as the reference writes it (linalg.f90:659-666, 717; 514-520, 587): rank 0 only.
static FILE* open_block_history(Ctx* c, const char* path, const char* method, int N, bool with_n) {
    if (!path || !*path || std::strcmp(path, "none") == 0 || c->rank != 0) return nullptr;
    FILE* f = std::fopen(path, "w");
    if (!f) return nullptr;
    std::fprintf(f, " method\n %s\n", method);
    if (with_n) std::fprintf(f, " N=%12d\n", N);
    std::fprintf(f, " iteration,||dx||,||err||,relaxation\n");
    return f;
}

// err_scale = |1/A(N,N)| when the reference's DIAG "preconditioner" scaled the system (the stopping test is on the
// scaled residual, linalg.f90:712-716), else 1.
ml_status block_jacobi_device(Ctx* c, int N, const double* dA, int ld, const double* d_b, int block_size, double tol, double rel,
                              int max_iter, int* iters, double* d_x, double err_scale, const char* iteration_file) {
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
    ML_CUDA(c, nrm.alloc(2));
    FILE* hist = open_block_history(c, iteration_file, "BJAC", N, true);
    ml_status st = ML_OK;
    for (int i = 0; i < N_blocks && st == ML_OK; ++i) {
        bs[i] = i * block_size;
        be[i] = (i == N_blocks - 1) ? N : (i + 1) * block_size;
        const int nb = be[i] - bs[i];
        bld[i] = ((nb + 63) / 64) * 64;
        if (blocks[i].alloc((size_t)bld[i] * nb) != cudaSuccess || pivs[i].alloc(2 * (size_t)nb) != cudaSuccess || vv.alloc(nb) != cudaSuccess) {
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
        double dx = 0.;
        if (hist) {   // dx = || x - x_new ||  (linalg.f90:714)
            vec_diff_norm_kernel<<<1, 1024, 0, c->stream>>>(d_x, x_new.p, N, nrm.p + 1);
            c->launches += 1;
        }
        cudaError_t e = cudaMemcpyAsync(d_x, x_new.p, (size_t)N * sizeof(double), cudaMemcpyDeviceToDevice, c->stream);
        if (e == cudaSuccess) e = cudaMemcpyAsync(&err, nrm.p, sizeof(double), cudaMemcpyDeviceToHost, c->stream);
        if (e == cudaSuccess && hist) e = cudaMemcpyAsync(&dx, nrm.p + 1, sizeof(double), cudaMemcpyDeviceToHost, c->stream);
        if (e == cudaSuccess) e = cudaStreamSynchronize(c->stream);
        if (e != cudaSuccess) { st = c->cuda_fail(e, "block_jacobi iteration"); break; }
        if (!(err == err)) { st = ML_NAN_RESIDUAL; break; }
        err *= err_scale;
        if (hist) std::fprintf(hist, "%6d,%10.3E,%10.3E,%10.3E\n", iteration, dx, err, rel);
    }
    if (hist) std::fclose(hist);
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

// ---- block Jacobi on a row-sharded system (block_jacobi_solve, common/linalg.f90:601-728) ----------------
// Rows never move.  The diagonal blocks (N_blocks = 5 by default: 20 % of the matrix) are assembled on every rank -- each rank
// stores its rows of a block into a zero-filled copy, one all-reduce per block adds the copies (one non-zero contributor per
// entry: exact) -- and factored redundantly, so the block solves of an iteration need no communication.  Per iteration: the
// right-hand sides b_i - sum_{j outside block} A x of the local rows (one kernel over all local rows, summation order of the
// single-GPU kernel), exchanged into the replicated vector; the block solves and the relaxation on every rank; the residual
// || A x_new - b || of the local rows, exchanged again.  All ranks hold identical vectors, hence take identical decisions.
__global__ void __launch_bounds__(256) bjs_pack_block_kernel(const double* __restrict__ A, int ld, int n_rows, const int* __restrict__ g_of_local,
                                                              int bs, int nb, double* __restrict__ dst, int ldd) {
    const int r = blockIdx.x * 256 + threadIdx.x;
    if (r >= n_rows) return;
    const int g = g_of_local[r];
    if (g >= bs && g < bs + nb) dst[(g - bs) + (size_t)blockIdx.y * ldd] = A[r + (size_t)(bs + blockIdx.y) * ld];
}
__global__ void bjs_init_kernel(const double* __restrict__ A, int ld, int n_rows, const int* __restrict__ g_of_local,
                                const double* __restrict__ b, double* __restrict__ out) {
    const int r = blockIdx.x * blockDim.x + threadIdx.x;
    if (r < n_rows) {
        const int g = g_of_local[r];
        out[r] = b[g] / A[r + (size_t)g * ld];   // linalg.f90:645-647
    }
}
// out[r] = b[g] - sum_{c outside the block of g} A[r, c] x[c] over the local rows (exclude = 0: the residual b - A x);
// warps split the columns and lanes the rows exactly as bj_rhs_kernel does
__global__ void __launch_bounds__(256) bjs_rhs_kernel(const double* __restrict__ A, int ld, int n, int n_rows, const int* __restrict__ g_of_local,
                                                       int block_size, int exclude, const double* __restrict__ b,
                                                       const double* __restrict__ x, double* __restrict__ out) {
    __shared__ double s_part[8][33];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int r = blockIdx.x * 32 + lane;
    double acc = 0.;
    int g = 0;
    if (r < n_rows) {
        g = g_of_local[r];
        int xs = 0, xe = 0;
        if (exclude) {
            xs = (g / block_size) * block_size;
            xe = min(n, xs + block_size);
        }
        for (int c = warp; c < n; c += 8)
            if (c < xs || c >= xe) acc = fma(A[r + (size_t)c * ld], __ldg(x + c), acc);
    }
    s_part[warp][lane] = acc;
    __syncthreads();
    if (warp == 0 && r < n_rows) {
        double t = 0.;
#pragma unroll
        for (int k = 0; k < 8; ++k) t += s_part[k][lane];
        out[r] = b[g] - t;
    }
}

ml_status block_jacobi_sharded(Ctx* c, const RowShardOps& R, const double* d_b, int block_size, double tol, double rel, int max_iter,
                               int* iters, double* d_x, double err_scale, const char* iteration_file) {
    const int N = R.N;
    if (block_size <= 0 || block_size > N) return c->fail(ML_BAD_ARGUMENT, "block_size out of range");
    int N_blocks = N / block_size;
    if (N % block_size > 0) N_blocks += 1;
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
    ML_CUDA(c, nrm.alloc(2));
    FILE* hist = open_block_history(c, iteration_file, "BJAC", N, true);
    ml_status st = ML_OK;
    for (int i = 0; i < N_blocks && st == ML_OK; ++i) {
        bs[i] = i * block_size;
        be[i] = (i == N_blocks - 1) ? N : (i + 1) * block_size;
        const int nb = be[i] - bs[i];
        bld[i] = ((nb + 63) / 64) * 64;
        if (blocks[i].alloc((size_t)bld[i] * nb) != cudaSuccess || pivs[i].alloc(2 * (size_t)nb) != cudaSuccess || vv.alloc(nb) != cudaSuccess) {
            cleanup();
            return c->fail(ML_CUDA_ERROR, "block Jacobi: out of device memory");
        }
        cudaError_t e = cudaMemsetAsync(blocks[i].p, 0, (size_t)bld[i] * nb * sizeof(double), c->stream);
        if (e != cudaSuccess) { cleanup(); return c->cuda_fail(e, "block Jacobi: memset"); }
        if (R.n_rows > 0) {
            dim3 grid((R.n_rows + 255) / 256, nb);
            bjs_pack_block_kernel<<<grid, 256, 0, c->stream>>>(R.A, R.ld, R.n_rows, R.g_of_local, bs[i], nb, blocks[i].p, bld[i]);
            c->launches += 1;
        }
#ifdef ML_HAVE_NCCL
        if (c->world > 1 &&
            ncclAllReduce(blocks[i].p, blocks[i].p, (size_t)bld[i] * nb, ncclDouble, ncclSum, c->comm, c->stream) != ncclSuccess) {
            cleanup();
            return c->fail(ML_NCCL_ERROR, "block Jacobi: all-reduce of a diagonal block");
        }
#endif
        st = lu_factor(c, blocks[i].p, bld[i], nb, pivs[i].p, vv.p, flag.p);
    }
    const int nbk = (N + 255) / 256, lbk = (R.n_rows + 31) / 32;
    int iteration = 0;
    double err = tol + 1.;
    if (st == ML_OK) {
        if (R.n_rows > 0) bjs_init_kernel<<<(R.n_rows + 255) / 256, 256, 0, c->stream>>>(R.A, R.ld, R.n_rows, R.g_of_local, d_b, R.slot);
        c->launches += 1;
        st = R.exchange(R.sys, d_x);
    }
    while (st == ML_OK && err >= tol && iteration < max_iter) {
        iteration += 1;
        if (R.n_rows > 0) bjs_rhs_kernel<<<lbk, 256, 0, c->stream>>>(R.A, R.ld, N, R.n_rows, R.g_of_local, block_size, 1, d_b, d_x, R.slot);
        c->launches += 1;
        st = R.exchange(R.sys, bi.p);
        for (int i = 0; i < N_blocks && st == ML_OK; ++i)
            st = lu_substitute(c, blocks[i].p, bld[i], be[i] - bs[i], pivs[i].p, bi.p + bs[i], x_new.p + bs[i]);
        if (st != ML_OK) break;
        bj_relax_kernel<<<nbk, 256, 0, c->stream>>>(rel, d_x, x_new.p, N);
        if (R.n_rows > 0) bjs_rhs_kernel<<<lbk, 256, 0, c->stream>>>(R.A, R.ld, N, R.n_rows, R.g_of_local, block_size, 0, d_b, x_new.p, R.slot);
        c->launches += 2;
        st = R.exchange(R.sys, bi.p);
        if (st != ML_OK) break;
        vec_norm_kernel<<<1, 1024, 0, c->stream>>>(bi.p, N, nrm.p);
        c->launches += 1;
        double dx = 0.;
        if (hist) {
            vec_diff_norm_kernel<<<1, 1024, 0, c->stream>>>(d_x, x_new.p, N, nrm.p + 1);
            c->launches += 1;
        }
        cudaError_t e = cudaMemcpyAsync(d_x, x_new.p, (size_t)N * sizeof(double), cudaMemcpyDeviceToDevice, c->stream);
        if (e == cudaSuccess) e = cudaMemcpyAsync(&err, nrm.p, sizeof(double), cudaMemcpyDeviceToHost, c->stream);
        if (e == cudaSuccess && hist) e = cudaMemcpyAsync(&dx, nrm.p + 1, sizeof(double), cudaMemcpyDeviceToHost, c->stream);
        if (e == cudaSuccess) e = cudaStreamSynchronize(c->stream);
        if (e != cudaSuccess) { st = c->cuda_fail(e, "block_jacobi iteration"); break; }
        if (!(err == err)) { st = ML_NAN_RESIDUAL; break; }
        err *= err_scale;
        if (hist) std::fprintf(hist, "%6d,%10.3E,%10.3E,%10.3E\n", iteration, dx, err, rel);
    }
    if (hist) std::fclose(hist);
    *iters = iteration;
    cleanup();
    return st;
}

ml_status block_ssor_device(Ctx* c, int N, const double* dA, int ld, const double* d_b, int block_size, double tol, double rel,
                            int max_iter, int* iters, double* d_x, double err_scale, const char* iteration_file) {
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
        bi.alloc(N) != cudaSuccess || nrm.alloc(2) != cudaSuccess) {
        cleanup();
        return c->fail(ML_CUDA_ERROR, "block SSOR: out of device memory");
    }
    ml_status st = ML_OK;
    for (int i = 0; i < N_blocks && st == ML_OK; ++i) {
        bs[i] = i * block_size;
        be[i] = (i == N_blocks - 1) ? N : (i + 1) * block_size;
        const int nb = be[i] - bs[i];
        bld[i] = ((nb + 63) / 64) * 64;
        if (blocks[i].alloc((size_t)bld[i] * nb) != cudaSuccess || pivs[i].alloc(2 * (size_t)nb) != cudaSuccess || vv.alloc(nb) != cudaSuccess) {
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
    FILE* hist = open_block_history(c, iteration_file, "BSSOR", N, false);
    DevBuf<double> x_prev;   // only for the ||dx|| column of the history
    if (hist && x_prev.alloc(N) != cudaSuccess) {
        std::fclose(hist);
        hist = nullptr;
    }
    while (st == ML_OK && err >= tol && iteration < max_iter) {
        iteration += 1;
        step = -step;   // first sweep runs forward (linalg.f90:516-527)
        if (hist) cudaMemcpyAsync(x_prev.p, d_x, (size_t)N * sizeof(double), cudaMemcpyDeviceToDevice, c->stream);
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
        double dx = 0.;
        if (hist) {
            vec_diff_norm_kernel<<<1, 1024, 0, c->stream>>>(x_prev.p, d_x, N, nrm.p + 1);
            c->launches += 1;
        }
        cudaError_t e = cudaMemcpyAsync(&err, nrm.p, sizeof(double), cudaMemcpyDeviceToHost, c->stream);
        if (e == cudaSuccess && hist) e = cudaMemcpyAsync(&dx, nrm.p + 1, sizeof(double), cudaMemcpyDeviceToHost, c->stream);
        if (e == cudaSuccess) e = cudaStreamSynchronize(c->stream);
        if (e != cudaSuccess) { st = c->cuda_fail(e, "block_ssor iteration"); break; }
        if (!(err == err)) { st = ML_NAN_RESIDUAL; break; }
        err *= err_scale;
        if (hist) std::fprintf(hist, "%6d,%10.3E,%10.3E,%10.3E\n", iteration, dx, err, rel);
    }
    if (hist) std::fclose(hist);
    x_prev.release();
    *iters = iteration;
    cleanup();
    return st;
}

}  // namespace mlgpu
