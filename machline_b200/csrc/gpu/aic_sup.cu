// supersonic instantiations of the assembly kernel; compiled with -fmad=false (see pair_influence.cuh)
#include "aic_kernels.cuh"

namespace mlgpu {

cudaError_t launch_aic_supersonic(Ctx* c, const AicLaunch& L) {
    if (L.row_nB) {   // Neumann rows
        switch (L.tile_rows) {
            case 32: return launch_aic_t<true, 32, 64, true>(c, L);
            case 16: return launch_aic_t<true, 16, 64, true>(c, L);
            case 8: return launch_aic_t<true, 8, 128, true>(c, L);
            case 4: return launch_aic_t<true, 4, 128, true>(c, L);
            default: return cudaErrorInvalidValue;
        }
    }
    switch (L.tile_rows) {
        case 32: return launch_aic_t<true, 32, 64>(c, L);
        case 16: return launch_aic_t<true, 16, 64>(c, L);
        case 8: return launch_aic_t<true, 8, 128>(c, L);
        case 4: return launch_aic_t<true, 4, 128>(c, L);
        default: return cudaErrorInvalidValue;
    }
}

}  // namespace mlgpu

// ---- census of the supersonic pair classes (bench.py: algorithmic flops of a supersonic assembly) --------
// counts[0] = pairs outside the domain of dependence (culled), counts[e] = evaluated pairs with e = 1..3 edges inside.
namespace mlgpu {

__global__ void __launch_bounds__(256) dod_census_kernel(const double* __restrict__ recs, int n_rec_slots, int stride,
                                                          const double* __restrict__ cp_xyz, const unsigned char* __restrict__ row_active,
                                                          int n_rows, int n_rows_pad, FlowConst fc, unsigned long long* counts) {
    __shared__ unsigned long long s_cnt[4];
    if (threadIdx.x < 4) s_cnt[threadIdx.x] = 0ull;
    __syncthreads();
    const int row = blockIdx.x * blockDim.x + threadIdx.x;
    unsigned long long c[4] = {0ull, 0ull, 0ull, 0ull};
    if (row < n_rows && row_active[row]) {
        const double Px = cp_xyz[row], Py = cp_xyz[n_rows_pad + row], Pz = cp_xyz[2 * (size_t)n_rows_pad + row];
        for (int r = blockIdx.y; r < n_rec_slots; r += gridDim.y) {
            const double* rec = recs + (size_t)r * stride;
            if (!(reinterpret_cast<const int*>(rec + R_FLAGS)[0] & RF_EVAL)) continue;
            bool e_in[3];
            const bool in = panel_check_dod(fc, rec, Px, Py, Pz, e_in);
            const int e = in ? (int)e_in[0] + (int)e_in[1] + (int)e_in[2] : 0;
            c[e] += 1;
        }
    }
    for (int k = 0; k < 4; ++k) atomicAdd(&s_cnt[k], c[k]);
    __syncthreads();
    if (threadIdx.x < 4) atomicAdd(counts + threadIdx.x, s_cnt[threadIdx.x]);
}

cudaError_t launch_dod_census(Ctx* c, const double* recs, int n_rec_slots, const double* cp_xyz, const unsigned char* row_active,
                              int n_rows, int n_rows_pad, const FlowConst& fc, unsigned long long* d_counts, int stride) {
    dim3 grid((n_rows + 255) / 256, 64);
    dod_census_kernel<<<grid, 256, 0, c->stream>>>(recs, n_rec_slots, stride, cp_xyz, row_active, n_rows, n_rows_pad, fc, d_counts);
    c->launches += 1;
    return cudaGetLastError();
}

}  // namespace mlgpu
