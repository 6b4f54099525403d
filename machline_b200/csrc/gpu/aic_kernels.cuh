// AIC assembly on sm_100a: replaces panel_solver_calc_domains_of_dependence (DoD pre-pass, fused:
// no N_panel x N_cp table), panel_solver_calc_body_influences and panel_solver_calc_wake_influences
// (src/panel_solver.f90:651-775, 1290-1501, 1504-1706).
//
// Ownership.  A CTA owns a tile of R consecutive rows of the permuted system for the WHOLE record
// stream (body panels with their mirror images interleaved, then wake panels: the reference's loop
// order), so no other CTA ever touches its part of A: no atomics, and every entry is summed in exactly
// the order the reference adds it (panel_solver.f90:1445-1476, then the separately summed wake row,
// :1650-1697).  Persistent CTAs pull tiles from an atomic counter.  R is 32, 16 or 8 so that there are
// several tiles per resident CTA even for a few thousand control points.
//
// Per chunk of C records:
//   stage  - the chunk's records and its scatter list arrive in shared memory by two TMA bulk copies
//            (cp.async.bulk + mbarrier); the copy of chunk t+1 is issued as soon as phase 1 of chunk t
//            is over, so it overlaps phase 2;
//   phase 1- thread = (record, row): DoD test + influence integrals (pair_influence.cuh) -> the three
//            doublet coefficients of the pair go to a shared-memory stage [record][slot][row];
//   phase 2- thread = (column, row): for every column the chunk feeds, walk the column's (record, slot)
//            items in reference order, acc = A[row, col] (or 0 on the column's first chunk), acc += item,
//            one coalesced store back (R consecutive rows of a column-major column).
// Wake chunks accumulate into a compact side matrix W (rows x wake columns) that is added to A at the
// end of the tile: the reference sums the wake influences of a row separately and then adds the row.
//
// The kernel template is instantiated in two translation units: aic_sup.cu (supersonic, -fmad=false: the
// reference's predicates are evaluated on identical IEEE values) and aic_sub.cu (subsonic, FMA contraction on);
// see pair_influence.cuh.
#pragma once
#include <cstdint>

#include "ctx.h"

namespace mlgpu {

#ifndef ML_AIC_SUB_CTAS
#define ML_AIC_SUB_CTAS 2
#endif
constexpr int AIC_THREADS = 256;
constexpr int AIC_WARPS = AIC_THREADS / 32;

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
// generic-proxy reads of a buffer must be ordered before the async proxy overwrites it
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

template <bool SUP, int C, bool HO = false>
struct AicSmem {
    static constexpr int STRIDE = record_stride(SUP, HO);
    static constexpr int ND = HO ? 6 : 3;       // doublet coefficients per pair (M_dim <= 6 for higher-order tables)
    static constexpr int REC_BYTES = C * STRIDE * 8;
    static constexpr int LIST_BYTES = list_bytes(C);
    static constexpr int MAXI = list_max_items(C);
    static constexpr int REC_OFF = 0;
    static constexpr int LIST_OFF = (REC_BYTES + 127) / 128 * 128;
    static constexpr int STAGE_OFF = (LIST_OFF + 2 * LIST_BYTES + 127) / 128 * 128;
    static constexpr int SLOTS = ND + (SUP ? 1 : 0);   // staged values per (record, row): the doublet coefficients (+ the source term)
    static constexpr int stage_bytes(int R) { return C * SLOTS * R * 8; }
    static constexpr int queue_bytes(int R) { return SUP ? R * C * 2 : 0; }   // per-row queues of in-DoD records (u16)
    static constexpr int total(int R) { return STAGE_OFF + stage_bytes(R) + queue_bytes(R); }
};

// VEL: Neumann rows (velocity influences projected on the row's direction, L.row_nB) -- a separate instantiation, so that the
// Dirichlet kernel carries neither the extra registers nor the branch
// HO: higher-order table (records with the extension of panel_record.h, six doublet coefficients per pair)
template <bool SUP, int R, int C, bool VEL = false, bool HO = false>
__global__ void __launch_bounds__(AIC_THREADS, SUP ? 2 : ML_AIC_SUB_CTAS) aic_assemble_kernel(const AicLaunch L) {
    using S = AicSmem<SUP, C, HO>;
    constexpr int ND = S::ND;
    static_assert(!(HO && VEL), "velocity influences of higher-order panels are not built");
    constexpr int SUBS = AIC_THREADS / R;   // records evaluated concurrently by the CTA
    constexpr int CPW = 32 / R;             // columns one warp handles concurrently in phase 2
    extern __shared__ __align__(128) unsigned char smem_raw[];
    double* const s_rec = reinterpret_cast<double*>(smem_raw + S::REC_OFF);
    unsigned char* const s_list = smem_raw + S::LIST_OFF;
    double* const s_stage = reinterpret_cast<double*>(smem_raw + S::STAGE_OFF);
    unsigned short* const s_q = reinterpret_cast<unsigned short*>(smem_raw + S::STAGE_OFF + C * S::SLOTS * R * 8);   // [R][C], SUP only
    __shared__ uint64_t full_bar[2];
    __shared__ int s_tile;
    __shared__ int s_qn[R];          // SUP: entries in each row's queue
    double* const s_red = s_stage;   // [AIC_THREADS] partial sums of I_known: the stage is free at the end of a tile

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int row_l = tid & (R - 1), sub0 = tid / R;
    if (tid == 0) {
        mbar_init(&full_bar[0], 1);
        mbar_init(&full_bar[1], 1);
        fence_mbar_init();
    }
    if (tid < R) s_qn[tid] = 0;
    __syncthreads();
    uint32_t phase[2] = {0, 0};
    const FlowConst& fc = L.fc;   // kernel parameters live in the constant bank
    const size_t ld = (size_t)L.ld;

    auto issue = [&](int t) {   // one thread: records + list of chunk t -> shared memory
        uint64_t* bar = &full_bar[t & 1];
        mbar_arrive_expect_tx(bar, (uint32_t)(S::REC_BYTES + S::LIST_BYTES));
        tma_bulk_g2s(s_rec, L.recs + (size_t)t * C * S::STRIDE, S::REC_BYTES, bar);
        tma_bulk_g2s(s_list + (t & 1) * S::LIST_BYTES, L.lists + (size_t)t * S::LIST_BYTES, S::LIST_BYTES, bar);
    };

    for (;;) {
        if (tid == 0) s_tile = atomicAdd(L.work_counter, 1);
        __syncthreads();
        const int tile = s_tile;
        if (tile >= L.n_tiles) break;
        const int row = tile * R + row_l;   // < n_rows_pad: the coordinate arrays are padded
        const bool active = (row < L.n_rows) && L.row_active[row];
        const double Px = L.cp_xyz[row], Py = L.cp_xyz[L.n_rows_pad + row], Pz = L.cp_xyz[2 * (size_t)L.n_rows_pad + row];
        // Neumann rows: the direction of the velocity projection (kernel-uniform choice: all rows of a system are of one kind)
        double nBv[3] = {0., 0., 0.};
        if constexpr (VEL) {
            nBv[0] = L.row_nB[row];
            nBv[1] = L.row_nB[L.n_rows_pad + row];
            nBv[2] = L.row_nB[2 * (size_t)L.n_rows_pad + row];
        }
        const double* const nB = VEL ? nBv : nullptr;
        if (tid == 0) issue(0);
        double Ik = 0.;

        for (int t = 0; t < L.n_chunks; ++t) {
            const int b = t & 1;
            mbar_wait(&full_bar[b], phase[b]);
            phase[b] ^= 1;
            int live;
            if constexpr (!SUP) {
                // ---- phase 1: pair influences -> stage ------------------------------------------------------
#pragma unroll 1
                for (int r = sub0; r < C; r += SUBS) {
                    const double* rec = s_rec + r * S::STRIDE;
                    const int flags = reinterpret_cast<const int*>(rec + R_FLAGS)[0];
                    double ps = 0., pd[ND] = {};
                    if (active && (flags & RF_EVAL)) {
                        if constexpr (HO) {
                            pair_influence_subsonic_ho(fc, rec, Px, Py, Pz, (flags & RF_MIRROR) != 0, ps, pd);
                            if (flags & RF_SOURCE) Ik = Ik + ps;   // the known strengths are folded into the record
                        } else {
                            pair_influence<false>(fc, rec, Px, Py, Pz, (flags & RF_MIRROR) != 0, ps, pd, nB);
                            if (flags & RF_SOURCE) Ik = Ik + ps * rec[R_SIGMA];   // panel_solver.f90:1245-1246
                        }
                    }
                    double* st = s_stage + (size_t)(r * S::SLOTS) * R + row_l;
#pragma unroll
                    for (int k = 0; k < ND; ++k) st[k * R] = pd[k];
                }
                __syncthreads();
                live = 1;
            } else {
                // ---- phase 1a: domain-of-dependence test of every (record, row) pair of the chunk; the pairs inside are
                // compacted with a warp ballot into one queue per row, so that phase 1b runs the influence integrals on
                // full warps whatever fraction of the chunk is culled (panel_check_dod, src/panel.f90:1732-1901) -----
                const bool chunk_src = (reinterpret_cast<const int*>(s_list + b * S::LIST_BYTES)[2] & LF_SOURCES) != 0;
                constexpr unsigned ROWMASK = R == 32 ? 0x1u : R == 16 ? 0x00010001u : R == 8 ? 0x01010101u : 0x11111111u;
#pragma unroll 1
                for (int r = sub0; r < C; r += SUBS) {
                    const double* rec = s_rec + r * S::STRIDE;
                    const int flags = reinterpret_cast<const int*>(rec + R_FLAGS)[0];
                    bool e_in[3] = {false, false, false};
                    bool in = false;
                    if (active && (flags & RF_EVAL)) in = panel_check_dod(fc, rec, Px, Py, Pz, e_in);
                    const unsigned mine = __ballot_sync(0xffffffffu, in) & (ROWMASK << (lane & (R - 1)));   // lanes of my row
                    int base = 0;
                    if (lane < R && mine) base = atomicAdd(&s_qn[row_l], __popc(mine));
                    base = __shfl_sync(0xffffffffu, base, lane & (R - 1));
                    if (in) {
                        s_q[row_l * C + base + __popc(mine & ((1u << lane) - 1u))] =
                            (unsigned short)(r | ((int)e_in[0] << 8) | ((int)e_in[1] << 9) | ((int)e_in[2] << 10));
                    } else {
                        double* st = s_stage + (size_t)(r * S::SLOTS) * R + row_l;
#pragma unroll
                        for (int k = 0; k < ND; ++k) st[k * R] = 0.;
                        if (chunk_src) st[ND * R] = 0.;
                    }
                }
                __syncthreads();
                // ---- phase 1b: influence integrals of the queued pairs -> stage --------------------------------
                const int qn = s_qn[row_l];
#pragma unroll 1
                for (int k = sub0; k < qn; k += SUBS) {
                    const unsigned e = s_q[row_l * C + k];
                    const int r = (int)(e & 0xffu);
                    const double* rec = s_rec + r * S::STRIDE;
                    const int flags = reinterpret_cast<const int*>(rec + R_FLAGS)[0];
                    const bool e_in[3] = {(e & 0x100u) != 0, (e & 0x200u) != 0, (e & 0x400u) != 0};
                    double ps = 0., pd[ND] = {};
                    if constexpr (HO) pair_eval_supersonic_ho(fc, rec, Px, Py, Pz, (flags & RF_MIRROR) != 0, e_in, ps, pd);
                    else pair_eval_supersonic(fc, rec, Px, Py, Pz, (flags & RF_MIRROR) != 0, e_in, ps, pd, nB);
                    double* st = s_stage + (size_t)(r * S::SLOTS) * R + row_l;
#pragma unroll
                    for (int k = 0; k < ND; ++k) st[k * R] = pd[k];
                    if (chunk_src) st[ND * R] = (flags & RF_SOURCE) ? (HO ? ps : ps * rec[R_SIGMA]) : 0.;
                }
                // a chunk entirely outside every row's domain of dependence adds only zeros -> phase 2 is skipped
                live = __syncthreads_or(qn > 0);
                if (tid < R) s_qn[tid] = 0;   // next written after the barrier that ends phase 2
                if (live && chunk_src) {
                    // known-source terms of this chunk, in the fixed order (records sub0, sub0 + SUBS, ...) of the subsonic kernel
#pragma unroll 1
                    for (int r = sub0; r < C; r += SUBS) Ik = Ik + s_stage[(size_t)(r * S::SLOTS + ND) * R + row_l];   // panel_solver.f90:1245-1246
                }
            }
            if (tid == 0 && t + 1 < L.n_chunks) {
                fence_proxy_async();
                issue(t + 1);
            }
            // ---- phase 2: ordered accumulation into A (or the wake side matrix) ------------------------------
            if (live) {
                const int* lst = reinterpret_cast<const int*>(s_list + b * S::LIST_BYTES);
                const int n_cols = lst[0];
                const bool wake = (lst[2] & LF_WAKE) != 0;
                const unsigned* cols = reinterpret_cast<const unsigned*>(lst + 4);
                const unsigned short* beg = reinterpret_cast<const unsigned short*>(lst + 4 + S::MAXI);
                const unsigned* item = reinterpret_cast<const unsigned*>(beg + S::MAXI + 2);
                double* const base = (wake ? L.W : L.A) + (size_t)tile * R + (lane & (R - 1));
                const char* const stage_lane = reinterpret_cast<const char*>(s_stage + (lane & (R - 1)));
                for (int ci = warp * CPW + lane / R; ci < n_cols; ci += AIC_WARPS * CPW) {
                    const unsigned cw = cols[ci];
                    double* dst = base + (size_t)(cw & ~COL_FIRST) * ld;
                    // first chunk of the pass that feeds this column: the sum starts from zero (no memset of A needed);
                    // the supersonic variant may have skipped that chunk, so it always reads (A is zero-filled).
                    double acc = (!SUP && (cw & COL_FIRST)) ? 0. : __ldcg(dst);
                    const int i1 = beg[ci + 1];
                    for (int it = beg[ci]; it < i1; ++it) {
                        const unsigned u = item[it];   // byte offset of the staged value | sign
                        const double v = *reinterpret_cast<const double*>(stage_lane + (u & ~ITEM_NEG));
                        acc = acc + __hiloint2double(__double2hiint(v) ^ (int)(u & ITEM_NEG), __double2loint(v));
                    }
                    __stcg(dst, acc);
                }
            }
            __syncthreads();   // stage and list buffer b are free again; A / W writes visible to the CTA
        }

        // ---- tile epilogue: A += W over the wake columns (panel_solver.f90:1691-1697), I_known ------------------
        if (L.n_wcols > 0) {
            double* const a = L.A + (size_t)tile * R + (lane & (R - 1));
            const double* const w = L.W + (size_t)tile * R + (lane & (R - 1));
            for (int wc = warp * CPW + lane / R; wc < L.n_wcols; wc += AIC_WARPS * CPW) {
                double* dst = a + (size_t)L.wcol[wc] * ld;
                __stcg(dst, __ldcg(dst) + __ldcg(w + (size_t)wc * ld));
            }
        }
        s_red[tid] = Ik;
        __syncthreads();
        if (tid < R) {
            // fixed-order sum of the per-thread partials (thread s*R + row handled records s, s+SUBS, ...)
            double s = 0.;
            for (int k = 0; k < SUBS; ++k) s = s + s_red[k * R + tid];
            if (tile * R + tid < L.n_rows) L.I_known[tile * R + tid] = s;
        }
        // s_tile / s_red are rewritten only after the next barrier pair
        __syncthreads();
    }
}

template <bool SUP, int R, int C, bool VEL = false, bool HO = false>
static cudaError_t launch_aic_t(Ctx* c, const AicLaunch& L) {
    using S = AicSmem<SUP, C, HO>;
    const size_t smem = S::total(R);
    auto kern = aic_assemble_kernel<SUP, R, C, VEL, HO>;
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    int occ = 1;
    e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kern, AIC_THREADS, smem);
    if (e != cudaSuccess) return e;
    if (occ < 1) occ = 1;
    int grid = c->num_sms * occ;
    if (grid > L.n_tiles) grid = L.n_tiles;
    if (grid < 1) grid = 1;
    kern<<<grid, AIC_THREADS, smem, c->stream>>>(L);
    c->launches += 1;
    return cudaGetLastError();
}

}  // namespace mlgpu
