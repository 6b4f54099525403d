// AIC assembly on sm_100a: replaces panel_solver_calc_domains_of_dependence (DoD pre-pass, fused:
// no N_panel x N_cp table), panel_solver_calc_body_influences and panel_solver_calc_wake_influences
// (src/panel_solver.f90:651-775, 1290-1501, 1504-1706).
//
// Mapping.  A work unit is (tile of 32 consecutive rows of the permuted system) x (segment of the
// record stream).  Persistent CTAs pull units from an atomic counter (supersonic rows are very
// uneven in cost).  Inside a unit the record stream -- body panels with their mirror images
// interleaved, then wake panels, i.e. the reference's loop order -- is staged tile by tile in
// shared memory by TMA bulk copies (cp.async.bulk + mbarrier, double buffered); lane = control
// point, warp = record, so every record is a shared-memory broadcast and all 32 lanes of a warp
// take the same DoD/edge branches except right at a Mach-cone boundary.  The three doublet
// coefficients of a pair land in three columns chosen by mesh connectivity: the 32 lanes of a warp
// hit 32 consecutive rows of the same column, so each scatter is one coalesced 256-byte
// red.global.add.f64.
//
// This translation unit is compiled with -fmad=false: branch-exact parity with the reference's
// predicates matters more here than fused multiply-adds (see pair_influence.cuh).
#include <cstdint>

#include "ctx.h"

namespace mlgpu {

__constant__ FlowConst c_flow;

constexpr int AIC_THREADS = 256;
constexpr int AIC_WARPS = AIC_THREADS / 32;
constexpr int AIC_TILE = 64;  // records per shared-memory stage

int aic_tile_records() { return AIC_TILE; }

// ---- TMA bulk copy + mbarrier helpers (PTX ISA: cp.async.bulk, mbarrier) ---------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, int count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void fence_mbar_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void tma_bulk_g2s(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                     smem_u32(dst)),
                 "l"(src), "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    asm volatile(
        "{\n"
        ".reg .pred P1;\n"
        "LAB_WAIT:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n"
        "@P1 bra DONE;\n"
        "bra LAB_WAIT;\n"
        "DONE:\n"
        "}\n" ::"r"(smem_u32(bar)),
        "r"(parity)
        : "memory");
}

template <bool SUP>
__global__ void __launch_bounds__(AIC_THREADS, 2) aic_assemble_kernel(const AicLaunch L) {
    constexpr int REC = SUP ? R_SUP_DOUBLES : R_SUB_DOUBLES;
    extern __shared__ __align__(128) unsigned char smem_raw[];
    double* stage0 = reinterpret_cast<double*>(smem_raw);
    double* stage1 = stage0 + AIC_TILE * REC;
    double* s_red = stage1 + AIC_TILE * REC;  // [AIC_WARPS][32] I_known partials
    __shared__ uint64_t full_bar[2];
    __shared__ int s_unit;

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    if (tid == 0) {
        mbar_init(&full_bar[0], 1);
        mbar_init(&full_bar[1], 1);
        fence_mbar_init();
    }
    __syncthreads();
    uint32_t phase0 = 0, phase1 = 0;
    const int n_units = L.n_cp_tiles * L.n_segments;
    const FlowConst& fc = c_flow;

    for (;;) {
        if (tid == 0) s_unit = atomicAdd(L.work_counter, 1);
        __syncthreads();
        const int u = s_unit;
        __syncthreads();
        if (u >= n_units) break;
        const int ct = u % L.n_cp_tiles, sg = u / L.n_cp_tiles;
        const int row = ct * 32 + lane;
        const bool active = (row < L.n_rows) && L.row_active[row];
        const size_t npad = (size_t)L.n_cp_tiles * 32;
        const double Px = L.cp_xyz[row], Py = L.cp_xyz[npad + row], Pz = L.cp_xyz[2 * npad + row];
        double* const Arow = L.A + row;
        const int t0 = sg * L.tiles_per_seg;
        const int t1 = min(t0 + L.tiles_per_seg, L.n_tiles);

        auto issue = [&](int t, int st) {
            const int nrec = min(AIC_TILE, L.n_rec - t * AIC_TILE);
            const uint32_t bytes = (uint32_t)nrec * REC * 8u;
            uint64_t* bar = &full_bar[st];
            mbar_arrive_expect_tx(bar, bytes);
            tma_bulk_g2s(st ? stage1 : stage0, L.recs + (size_t)t * AIC_TILE * REC, bytes, bar);
        };
        if (tid == 0) issue(t0, 0);
        double Ik = 0.;
        for (int t = t0; t < t1; ++t) {
            const int st = (t - t0) & 1;
            if (tid == 0 && t + 1 < t1) issue(t + 1, st ^ 1);
            if (st == 0) {
                mbar_wait(&full_bar[0], phase0);
                phase0 ^= 1;
            } else {
                mbar_wait(&full_bar[1], phase1);
                phase1 ^= 1;
            }
            const double* base = st ? stage1 : stage0;
            const int nrec = min(AIC_TILE, L.n_rec - t * AIC_TILE);
            if (active) {
                for (int r = warp; r < nrec; r += AIC_WARPS) {
                    const double* rec = base + r * REC;
                    const int flags = reinterpret_cast<const int*>(rec + R_FLAGS)[0];
                    double ps, pd[3];
                    if (pair_influence<SUP>(fc, rec, Px, Py, Pz, (flags & RF_MIRROR) != 0, ps, pd)) {
                        const int* cols = reinterpret_cast<const int*>(rec + R_COLS);
                        if (flags & RF_SOURCE) Ik = Ik + ps * rec[R_SIGMA];
#pragma unroll
                        for (int k = 0; k < 3; ++k) {
                            const int c = cols[k];
                            if (c >= 0) atomicAdd(Arow + (size_t)c * L.ld, pd[k]);
                        }
#pragma unroll
                        for (int k = 0; k < 3; ++k) {
                            const int c = cols[3 + k];
                            if (c >= 0) atomicAdd(Arow + (size_t)c * L.ld, -pd[k]);
                        }
                    }
                }
            }
            __syncthreads();  // stage st may be refilled two iterations from now
        }
        // I_known: fixed-order sum of the per-warp partials, then one atomic per row and segment
        s_red[warp * 32 + lane] = Ik;
        __syncthreads();
        if (warp == 0 && active) {
            double s = 0.;
#pragma unroll
            for (int w = 0; w < AIC_WARPS; ++w) s = s + s_red[w * 32 + lane];
            atomicAdd(L.I_known + row, s);
        }
    }
}

// Strength-matching rows (bc = 4): A(P(i),P(i)) = 1, A(P(i),P(i-N_cp/2)) = -1 (panel_solver.f90:1317-1320)
__global__ void strength_rows_kernel(double* A, int ld, const int* rows, const int* colp, const int* colm, int n) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) {
        A[rows[i] + (size_t)colp[i] * ld] = 1.;
        A[rows[i] + (size_t)colm[i] * ld] = -1.;
    }
}

cudaError_t upload_flow_constants(const ml_flow& f, cudaStream_t s) {
    FlowConst h;
    for (int i = 0; i < 3; ++i) h.c_hat[i] = f.c_hat_g[i];
    for (int i = 0; i < 9; ++i) h.C[i] = f.C_mat_g[i];
    h.K_inv = f.K_inv;
    h.s = (int)f.s;
    h.supersonic = f.supersonic;
    return cudaMemcpyToSymbolAsync(c_flow, &h, sizeof h, 0, cudaMemcpyHostToDevice, s);
}

template <bool SUP>
static cudaError_t launch_aic_t(Ctx* c, const AicLaunch& L) {
    constexpr int REC = SUP ? R_SUP_DOUBLES : R_SUB_DOUBLES;
    const size_t smem = (size_t)2 * AIC_TILE * REC * 8 + (size_t)AIC_WARPS * 32 * 8;
    cudaError_t e = cudaFuncSetAttribute(aic_assemble_kernel<SUP>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    int occ = 1;
    e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, aic_assemble_kernel<SUP>, AIC_THREADS, smem);
    if (e != cudaSuccess) return e;
    if (occ < 1) occ = 1;
    const int n_units = L.n_cp_tiles * L.n_segments;
    int grid = c->num_sms * occ;
    if (grid > n_units) grid = n_units;
    if (grid < 1) grid = 1;
    aic_assemble_kernel<SUP><<<grid, AIC_THREADS, smem, c->stream>>>(L);
    c->launches += 1;
    return cudaGetLastError();
}

cudaError_t launch_aic(Ctx* c, const AicLaunch& L, bool supersonic) {
    return supersonic ? launch_aic_t<true>(c, L) : launch_aic_t<false>(c, L);
}

cudaError_t launch_strength_rows(Ctx* c, double* A, int ld, const int* rows, const int* colp, const int* colm, int n) {
    if (n <= 0) return cudaSuccess;
    strength_rows_kernel<<<(n + 127) / 128, 128, 0, c->stream>>>(A, ld, rows, colp, colm, n);
    c->launches += 1;
    return cudaGetLastError();
}

}  // namespace mlgpu
