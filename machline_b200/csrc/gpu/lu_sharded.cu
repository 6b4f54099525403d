// LU of a row-sharded matrix over NCCL; the single-GPU factorisation is in lu_kernels.cu and both use the
// same panel, TRSM and tensor-core update kernels (lu_internal.cuh).  Replaces common/linalg.f90:118-342 (lu_solve) for a
// matrix whose rows live on several GPUs.
#include <algorithm>
#include <string>
#include <vector>

#include "lu_internal.cuh"

namespace mlgpu {

// ---- LU of a row-sharded matrix (NCCL over NVLink) ---------------------------------------------------
// Every rank keeps the rows it assembled (n_rows x N, column-major) for the whole factorisation: rows never move
// between GPUs.  The reference's row interchanges act on POSITIONS; perm[pos] = slot (rank * S + local row) of the
// row currently at position pos is replicated on every rank and is the only thing an interchange changes.
// Per panel of 64 columns:
//   1. ncclAllGather of the panel's columns (local rows x 64) -> every rank has the whole panel; a gather kernel puts
//      it into position order;
//   2. every rank factors the panel redundantly with the same cooperative kernel as the single-GPU path (identical
//      inputs -> identical pivots and multipliers everywhere; no pivot broadcast, no per-column collective);
//   3. the 64 pivot rows' trailing entries are packed by their owners into a zero-filled 64 x W block and summed with
//      ncclAllReduce (each entry has exactly one non-zero contributor) -> U-row block on every rank; TRSM with the
//      panel's diagonal block, redundantly; owners store their rows back;
//   4. each rank updates the rows it owns that are still below the panel: C -= L21(local) * U12 on the FP64 tensor
//      cores (blocks of 128 local rows without such a row are skipped).
// The right-hand side travels as column N of the local matrix, so forward elimination costs nothing extra and L is
// never stored.  Back substitution walks the panels backwards: owners contribute their 64 y entries (allreduce), the
// diagonal block (kept from step 2, replicated) is solved redundantly, and every rank updates its local y.
__global__ void dist_init_perm_kernel(int* __restrict__ perm, int* __restrict__ pos_of_lr, int N, const int* __restrict__ slot_of_g, int S,
                                      int rank, const double* __restrict__ b, double* __restrict__ ycol) {
    // initially position = global row: perm[pos] = its slot; the owner notes the position of its local row and takes b
    const int pos = blockIdx.x * blockDim.x + threadIdx.x;
    if (pos >= N) return;
    const int s = slot_of_g[pos];
    perm[pos] = s;
    if (s / S == rank) {
        pos_of_lr[s - rank * S] = pos;
        ycol[s - rank * S] = b[pos];
    }
}
// vv(pos) = 1 / amax(slot at pos); a zero row flags the matrix singular
__global__ void dist_vv_kernel(const double* __restrict__ amax_all, const int* __restrict__ perm, int N, double* __restrict__ vv,
                               int* __restrict__ flag) {
    const int pos = blockIdx.x * blockDim.x + threadIdx.x;
    if (pos >= N) return;
    const double a = amax_all[perm[pos]];
    if (a <= 1.5e-20) atomicExch(flag, 1);
    vv[pos] = 1.0 / a;
}
// Pbuf(pos, c) = gathered panel entry of the row at position pos, for pos >= k0
__global__ void __launch_bounds__(256) dist_gather_panel_kernel(const double* __restrict__ Gall, int S, int nb, const int* __restrict__ perm,
                                                                 int k0, int N, double* __restrict__ Pbuf, int NP) {
    const int pos = k0 + blockIdx.x * 256 + threadIdx.x;
    const int c = blockIdx.y;
    if (pos >= N) return;
    const int s = perm[pos], r = s / S, lr = s - r * S;
    Pbuf[pos + (size_t)c * NP] = Gall[((size_t)r * nb + c) * S + lr];
}
__global__ void dist_pos_of_row_kernel(const int* __restrict__ perm, int k0, int N, int S, int rank, int* __restrict__ pos_of_lr) {
    const int pos = k0 + blockIdx.x * blockDim.x + threadIdx.x;
    if (pos >= N) return;
    const int s = perm[pos];
    if (s / S == rank) pos_of_lr[s - rank * S] = pos;
}
// the per-column fallback of the panel factorisation records interchanges only: apply them to perm
__global__ void dist_apply_piv_kernel(const int* __restrict__ piv, int k0, int k1, int* __restrict__ perm) {
    if (blockIdx.x == 0 && threadIdx.x == 0)
        for (int j = k0; j < k1; ++j) {
            const int p = piv[j];
            if (p != j) {
                const int t = perm[p];
                perm[p] = perm[j];
                perm[j] = t;
            }
        }
}
// Usend(j, c') = A_loc(row at position k0+j, c0+c') if this rank owns that row, else 0   (j < 64, c' < W)
__global__ void __launch_bounds__(256) dist_pack_u_kernel(const double* __restrict__ Aloc, int ld, const int* __restrict__ perm, int k0, int nb,
                                                           int S, int rank, int c0, int W, double* __restrict__ Usend) {
    const int j = threadIdx.x & 63;
    const int cc = blockIdx.x * 4 + (threadIdx.x >> 6);
    if (cc >= W) return;
    double v = 0.;
    if (j < nb) {
        const int s = perm[k0 + j];
        if (s / S == rank) v = Aloc[(s - rank * S) + (size_t)(c0 + cc) * ld];
    }
    Usend[j + (size_t)cc * LU_NB] = v;
}
__global__ void __launch_bounds__(256) dist_unpack_u_kernel(double* __restrict__ Aloc, int ld, const int* __restrict__ perm, int k0, int nb,
                                                             int S, int rank, int c0, int W, const double* __restrict__ Ubuf) {
    const int j = threadIdx.x & 63;
    const int cc = blockIdx.x * 4 + (threadIdx.x >> 6);
    if (cc >= W || j >= nb) return;
    const int s = perm[k0 + j];
    if (s / S == rank) Aloc[(s - rank * S) + (size_t)(c0 + cc) * ld] = Ubuf[j + (size_t)cc * LU_NB];
}
// Lloc(lr, c) = multiplier of local row lr in the factored panel (0 for rows at or above the panel's last pivot)
__global__ void __launch_bounds__(256) dist_gather_l_kernel(const double* __restrict__ Pbuf, int NP, const int* __restrict__ pos_of_lr, int n_rows,
                                                             int n_rows_pad, int k1, double* __restrict__ Lloc,
                                                             unsigned char* __restrict__ rb_active) {
    const int lr = blockIdx.x * 256 + threadIdx.x;
    const int c = blockIdx.y;
    if (lr >= n_rows_pad) return;
    const int pos = lr < n_rows ? pos_of_lr[lr] : -1;
    const bool active = pos >= k1;
    Lloc[lr + (size_t)c * n_rows_pad] = active ? Pbuf[pos + (size_t)c * NP] : 0.;
    if (active && c == 0) rb_active[lr / LU_GEMM_BM] = 1;
}
__global__ void dist_pack_y_kernel(const double* __restrict__ ycol, const int* __restrict__ perm, int k0, int nb, int S, int rank,
                                   double* __restrict__ ysend) {
    const int j = threadIdx.x;
    double v = 0.;
    if (j < nb) {
        const int s = perm[k0 + j];
        if (s / S == rank) v = ycol[s - rank * S];
    }
    ysend[j] = v;
}
// y_loc -= A_loc(:, k0..k1) x_k  (rows already solved receive garbage that is never read again)
__global__ void __launch_bounds__(256) dist_y_update_kernel(const double* __restrict__ Aloc, int ld, int n_rows, int k0, int nb,
                                                             const double* __restrict__ xk, double* __restrict__ ycol, double* __restrict__ x_out) {
    __shared__ double sx[LU_NB];
    if (threadIdx.x < nb) {
        sx[threadIdx.x] = xk[threadIdx.x];
        if (blockIdx.x == 0) x_out[k0 + threadIdx.x] = xk[threadIdx.x];
    }
    __syncthreads();
    const int r = blockIdx.x * 256 + threadIdx.x;
    if (r >= n_rows) return;
    double acc = 0.;
    for (int c = 0; c < nb; ++c) acc = fma(Aloc[r + (size_t)(k0 + c) * ld], sx[c], acc);
    ycol[r] -= acc;
}
__global__ void dist_copy_block_kernel(const double* __restrict__ Pbuf, int NP, int k0, int nb, double* __restrict__ D) {
    // D (64 x 64, zero-initialised) <- the panel's nb x nb diagonal block [L11 \ U11]
    const int r = threadIdx.x, cc = blockIdx.x;
    if (r < nb && cc < nb) D[r + cc * LU_NB] = Pbuf[(k0 + r) + (size_t)cc * NP];
}

#ifdef ML_HAVE_NCCL
#define DIST_NCCL(call)                                                                        \
    do {                                                                                       \
        ncclResult_t r__ = (call);                                                             \
        if (r__ != ncclSuccess) { st = c->fail(ML_NCCL_ERROR, std::string(#call) + ": " + ncclGetErrorString(r__)); goto done; } \
    } while (0)
#endif
#define DIST_CUDA(call)                                                   \
    do {                                                                  \
        cudaError_t e__ = (call);                                         \
        if (e__ != cudaSuccess) { st = c->cuda_fail(e__, #call); goto done; } \
    } while (0)

// dAloc: this rank's rows (n_rows x (N+1) columns allocated, leading dimension ld, a scratch copy: overwritten);
// d_b: the whole right-hand side on every rank; d_x: the whole solution on every rank.
ml_status lu_solve_sharded(Ctx* c, int N, double* dAloc, int ld, int n_rows, int n_rows_pad, int S, const double* d_b, double* d_x) {
    const int world = c->world, rank = c->rank;
    if ((int)c->shard_nrows.size() != world || c->d_slot_of_g.p == nullptr || c->shard_pad != S) return c->fail(ML_NOT_READY, "row shards unknown");
    if (ld & 1) return c->fail(ML_BAD_ARGUMENT, "leading dimension must be even");
    const int NP = ((N + 63) / 64) * 64;
    const int npan = (N + LU_NB - 1) / LU_NB;
    ml_status st = ML_OK;
    DevBuf<int> perm, piv, pos_of_lr, flag;
    DevBuf<double> vv, amax_loc, amax_all, Gsend, Gall, Pbuf, Usend, Ubuf, Dall, Lloc, yvec;
    DevBuf<unsigned char> rb_active;
    LuPanelWork PW;
    const int n_rb = (n_rows_pad + LU_GEMM_BM - 1) / LU_GEMM_BM;
    double* ycol = dAloc + (size_t)N * ld;
    int h_flag = 0;
    auto allgather = [&](const double* send, double* recv, size_t count) -> bool {
        if (world == 1) return cudaMemcpyAsync(recv, send, count * sizeof(double), cudaMemcpyDeviceToDevice, c->stream) == cudaSuccess;
#ifdef ML_HAVE_NCCL
        return ncclAllGather(send, recv, count, ncclDouble, c->comm, c->stream) == ncclSuccess;
#else
        return false;
#endif
    };
    auto allreduce = [&](const double* send, double* recv, size_t count) -> bool {
        if (world == 1) return cudaMemcpyAsync(recv, send, count * sizeof(double), cudaMemcpyDeviceToDevice, c->stream) == cudaSuccess;
#ifdef ML_HAVE_NCCL
        return ncclAllReduce(send, recv, count, ncclDouble, ncclSum, c->comm, c->stream) == ncclSuccess;
#else
        return false;
#endif
    };
    DIST_CUDA(perm.alloc(N));
    DIST_CUDA(piv.alloc(N));
    DIST_CUDA(pos_of_lr.alloc(std::max(1, S)));
    DIST_CUDA(flag.alloc(1));
    DIST_CUDA(vv.alloc(N));
    DIST_CUDA(amax_loc.alloc(S));
    DIST_CUDA(amax_all.alloc((size_t)S * world));
    DIST_CUDA(Gsend.alloc((size_t)S * LU_NB));
    DIST_CUDA(Gall.alloc((size_t)S * LU_NB * world));
    DIST_CUDA(Pbuf.alloc((size_t)NP * LU_NB));
    DIST_CUDA(Usend.alloc((size_t)LU_NB * (N + 1)));
    DIST_CUDA(Ubuf.alloc((size_t)LU_NB * (N + 1)));
    DIST_CUDA(Dall.alloc((size_t)LU_NB * LU_NB * npan));
    DIST_CUDA(Lloc.alloc((size_t)n_rows_pad * LU_NB));
    DIST_CUDA(yvec.alloc(2 * LU_NB));
    DIST_CUDA(rb_active.alloc(n_rb));
    st = PW.init(c);
    if (st != ML_OK) goto done;
    DIST_CUDA(cudaMemsetAsync(flag.p, 0, sizeof(int), c->stream));
    DIST_CUDA(cudaMemsetAsync(Dall.p, 0, (size_t)LU_NB * LU_NB * npan * sizeof(double), c->stream));
    DIST_CUDA(cudaMemsetAsync(amax_loc.p, 0, (size_t)S * sizeof(double), c->stream));
    // perm = identity in slot terms; the right-hand side becomes column N of the local rows
    dist_init_perm_kernel<<<(N + 255) / 256, 256, 0, c->stream>>>(perm.p, pos_of_lr.p, N, c->d_slot_of_g.p, S, rank, d_b, ycol);
    lu_launch_row_amax(c, dAloc, ld, n_rows, N, amax_loc.p);
    c->launches += 1;
    if (!allgather(amax_loc.p, amax_all.p, S)) { st = c->fail(ML_NCCL_ERROR, "allgather row maxima"); goto done; }
    dist_vv_kernel<<<(N + 255) / 256, 256, 0, c->stream>>>(amax_all.p, perm.p, N, vv.p, flag.p);
    c->launches += 1;
    DIST_CUDA(cudaMemcpyAsync(&h_flag, flag.p, sizeof(int), cudaMemcpyDeviceToHost, c->stream));
    DIST_CUDA(cudaStreamSynchronize(c->stream));
    if (h_flag) { st = c->fail(ML_SINGULAR, "lu_decomp: the matrix is singular (a row is zero; linalg.f90:205-208)"); goto done; }

    for (int k = 0; k < npan; ++k) {
        const int k0 = k * LU_NB, k1 = std::min(k0 + LU_NB, N), nb = k1 - k0;
        const int W = N + 1 - k1;   // trailing columns + the right-hand side
        // 1. the whole panel on every rank, in position order
        if (n_rows > 0)
            DIST_CUDA(cudaMemcpy2DAsync(Gsend.p, (size_t)S * sizeof(double), dAloc + (size_t)k0 * ld, (size_t)ld * sizeof(double),
                                        (size_t)n_rows * sizeof(double), nb, cudaMemcpyDeviceToDevice, c->stream));
        if (!allgather(Gsend.p, Gall.p, (size_t)S * nb)) { st = c->fail(ML_NCCL_ERROR, "allgather panel"); goto done; }
        dist_gather_panel_kernel<<<dim3((N - k0 + 255) / 256, nb), 256, 0, c->stream>>>(Gall.p, S, nb, perm.p, k0, N, Pbuf.p, NP);
        c->launches += 1;
        // 2. factor it (replicated); the kernel addresses column k0 + c at A + (k0 + c) * ld
        st = lu_panel_factor(c, PW, Pbuf.p - (size_t)k0 * NP, NP, N, k0, k1, vv.p, piv.p, perm.p, c->stream);
        if (st != ML_OK) goto done;
        if (!PW.all_coop) {
            dist_apply_piv_kernel<<<1, 32, 0, c->stream>>>(piv.p, k0, k1, perm.p);
            c->launches += 1;
            PW.all_coop = true;
        }
        dist_pos_of_row_kernel<<<(N - k0 + 255) / 256, 256, 0, c->stream>>>(perm.p, k0, N, S, rank, pos_of_lr.p);
        dist_copy_block_kernel<<<nb, 64, 0, c->stream>>>(Pbuf.p, NP, k0, nb, Dall.p + (size_t)k * LU_NB * LU_NB);
        // 3. U-row block: owners pack, sum over ranks, solve with L11, owners store back
        dist_pack_u_kernel<<<(W + 3) / 4, 256, 0, c->stream>>>(dAloc, ld, perm.p, k0, nb, S, rank, k1, W, Usend.p);
        c->launches += 3;
        if (!allreduce(Usend.p, Ubuf.p, (size_t)LU_NB * W)) { st = c->fail(ML_NCCL_ERROR, "allreduce U rows"); goto done; }
        lu_launch_trsm(c, Dall.p + (size_t)k * LU_NB * LU_NB, LU_NB, Ubuf.p, LU_NB, W);
        dist_unpack_u_kernel<<<(W + 3) / 4, 256, 0, c->stream>>>(dAloc, ld, perm.p, k0, nb, S, rank, k1, W, Ubuf.p);
        c->launches += 1;
        // 4. local trailing update
        if (k1 < N && n_rows > 0) {
            DIST_CUDA(cudaMemsetAsync(rb_active.p, 0, n_rb, c->stream));
            dist_gather_l_kernel<<<dim3((n_rows_pad + 255) / 256, LU_NB), 256, 0, c->stream>>>(Pbuf.p, NP, pos_of_lr.p, n_rows, n_rows_pad, k1,
                                                                                               Lloc.p, rb_active.p);
            c->launches += 1;
            lu_gemm2_launch(c, Lloc.p, n_rows_pad, Ubuf.p, LU_NB, dAloc + (size_t)k1 * ld, ld, n_rows, W, rb_active.p, 1);
        }
        DIST_CUDA(cudaGetLastError());
    }
    // back substitution, last panel first
    for (int k = npan - 1; k >= 0; --k) {
        const int k0 = k * LU_NB, k1 = std::min(k0 + LU_NB, N), nb = k1 - k0;
        dist_pack_y_kernel<<<1, 64, 0, c->stream>>>(ycol, perm.p, k0, nb, S, rank, yvec.p);
        if (!allreduce(yvec.p, yvec.p + LU_NB, LU_NB)) { st = c->fail(ML_NCCL_ERROR, "allreduce y block"); goto done; }
        lu_launch_trsv_diag(c, Dall.p + (size_t)k * LU_NB * LU_NB, LU_NB, nb, yvec.p + LU_NB, 1);
        dist_y_update_kernel<<<std::max(1, (n_rows + 255) / 256), 256, 0, c->stream>>>(dAloc, ld, n_rows, k0, nb, yvec.p + LU_NB, ycol, d_x);
        c->launches += 2;
    }
    DIST_CUDA(cudaGetLastError());
    DIST_CUDA(cudaStreamSynchronize(c->stream));
done:
    PW.release();
    perm.release(); piv.release(); pos_of_lr.release(); flag.release();
    vv.release(); amax_loc.release(); amax_all.release(); Gsend.release(); Gall.release(); Pbuf.release();
    Usend.release(); Ubuf.release(); Dall.release(); Lloc.release(); yvec.release(); rb_active.release();
    return st;
}
#undef DIST_CUDA

}  // namespace mlgpu
