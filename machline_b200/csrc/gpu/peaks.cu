// Measured denominators for the rooflines that MEASURED_PEAKS.json does not carry (no FP64
// peak was measured by the driver): a register-resident DFMA loop over all SMs (FP64 vector pipe) and
// a streaming copy (HBM).  Used by bench.py only.
#include <cstdlib>

#include "ctx.h"

namespace mlgpu {

__global__ void __launch_bounds__(256) dfma_peak_kernel(double* out, int iters, double seed) {
    // 16 independent accumulator chains per thread: enough ILP to cover the FP64 pipe latency
    double a[16];
#pragma unroll
    for (int k = 0; k < 16; ++k) a[k] = seed + k * 1e-3 + threadIdx.x * 1e-6;
    const double m = 1.0000001, c = 1e-9;
    for (int i = 0; i < iters; ++i) {
#pragma unroll
        for (int k = 0; k < 16; ++k) a[k] = fma(a[k], m, c);
    }
    double s = 0.;
#pragma unroll
    for (int k = 0; k < 16; ++k) s += a[k];
    if (s == 123.456) out[blockIdx.x * blockDim.x + threadIdx.x] = s;  // never true; keeps the loop alive
}

__global__ void __launch_bounds__(256) copy_kernel(const double2* __restrict__ src, double2* __restrict__ dst, size_t n) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    size_t stride = (size_t)gridDim.x * blockDim.x;
    for (; i < n; i += stride) dst[i] = src[i];
}

// FP64 tensor pipe: register-resident mma.sync.m8n8k4.f64 (DMMA) loop, 16 independent accumulator tiles per warp
// (the trailing update of the blocked LU keeps 16 per warp as well).  512 flop per instruction per warp.
__global__ void __launch_bounds__(256) dmma_peak_kernel(double* out, int iters, double seed) {
    double c[16][2];
#pragma unroll
    for (int k = 0; k < 16; ++k) {
        c[k][0] = seed + k * 1e-3;
        c[k][1] = seed - k * 1e-3;
    }
    double a[4], b[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        a[k] = 1e-9 * (threadIdx.x + k);
        b[k] = 1e-9 * (threadIdx.x - k);
    }
    for (int i = 0; i < iters; ++i) {
#pragma unroll
        for (int m = 0; m < 4; ++m)
#pragma unroll
            for (int n = 0; n < 4; ++n)
                asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                             : "+d"(c[m * 4 + n][0]), "+d"(c[m * 4 + n][1])
                             : "d"(a[m]), "d"(b[n]));
    }
    double s = 0.;
#pragma unroll
    for (int k = 0; k < 16; ++k) s += c[k][0] + c[k][1];
    if (s == 123.456) out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

}  // namespace mlgpu

using namespace mlgpu;

extern "C" ml_status ml_measure_dmma_peak(ml_ctx* c, double* tflops) {
    if (!c || !tflops) return ML_BAD_ARGUMENT;
    if (c->group) return ml_measure_dmma_peak(mlgpu::multi_member(c, 0), tflops);
    ML_CUDA(c, cudaSetDevice(c->device));
    DevBuf<double> out;
    ML_CUDA(c, out.alloc(1024));
    const char* e = getenv("ML_DMMA_BLOCKS_PER_SM");
    const int iters = 4000, blocks = c->num_sms * (e ? atoi(e) : 4);
    float best = 1e30f;
    for (int rep = 0; rep < 4; ++rep) {
        ML_CUDA(c, cudaEventRecord(c->ev0, c->stream));
        dmma_peak_kernel<<<blocks, 256, 0, c->stream>>>(out.p, iters, 1.0);
        ML_CUDA(c, cudaEventRecord(c->ev1, c->stream));
        ML_CUDA(c, cudaStreamSynchronize(c->stream));
        float ms;
        ML_CUDA(c, cudaEventElapsedTime(&ms, c->ev0, c->ev1));
        if (rep > 0 && ms < best) best = ms;
    }
    c->launches += 4;
    *tflops = 512.0 * 16.0 * iters * 8.0 * blocks / (best * 1e-3) / 1e12;
    out.release();
    return ML_OK;
}

extern "C" ml_status ml_measure_peaks(ml_ctx* c, double* fp64_tflops, double* hbm_gbs) {
    if (!c) return ML_BAD_ARGUMENT;
    if (c->group) return ml_measure_peaks(mlgpu::multi_member(c, 0), fp64_tflops, hbm_gbs);
    ML_CUDA(c, cudaSetDevice(c->device));
    DevBuf<double> out;
    ML_CUDA(c, out.alloc(1024));
    const int iters = 20000, blocks = c->num_sms * 8;
    float best = 1e30f;
    for (int rep = 0; rep < 4; ++rep) {
        ML_CUDA(c, cudaEventRecord(c->ev0, c->stream));
        dfma_peak_kernel<<<blocks, 256, 0, c->stream>>>(out.p, iters, 1.0);
        ML_CUDA(c, cudaEventRecord(c->ev1, c->stream));
        ML_CUDA(c, cudaStreamSynchronize(c->stream));
        float ms;
        ML_CUDA(c, cudaEventElapsedTime(&ms, c->ev0, c->ev1));
        if (rep > 0 && ms < best) best = ms;
    }
    c->launches += 4;
    if (fp64_tflops) *fp64_tflops = 2.0 * 16.0 * iters * 256.0 * blocks / (best * 1e-3) / 1e12;
    out.release();
    if (hbm_gbs) {
        const size_t n = (size_t)1 << 27;  // 2 GiB per buffer of double2
        DevBuf<double> a, b;
        ML_CUDA(c, a.alloc(2 * n));
        ML_CUDA(c, b.alloc(2 * n));
        ML_CUDA(c, cudaMemsetAsync(a.p, 0, 2 * n * sizeof(double), c->stream));
        best = 1e30f;
        for (int rep = 0; rep < 5; ++rep) {
            ML_CUDA(c, cudaEventRecord(c->ev0, c->stream));
            copy_kernel<<<c->num_sms * 16, 256, 0, c->stream>>>((const double2*)a.p, (double2*)b.p, n);
            ML_CUDA(c, cudaEventRecord(c->ev1, c->stream));
            ML_CUDA(c, cudaStreamSynchronize(c->stream));
            float ms;
            ML_CUDA(c, cudaEventElapsedTime(&ms, c->ev0, c->ev1));
            if (rep > 0 && ms < best) best = ms;
        }
        c->launches += 5;
        *hbm_gbs = 2.0 * n * 16.0 / (best * 1e-3) / 1e9;
        a.release();
        b.release();
    }
    return ML_OK;
}

extern "C" ml_status ml_nccl_unique_id(void* out128) {
    if (!out128) return ML_BAD_ARGUMENT;
#ifdef ML_HAVE_NCCL
    ncclUniqueId id;
    if (ncclGetUniqueId(&id) != ncclSuccess) return ML_NCCL_ERROR;
    static_assert(sizeof(id) == 128, "ncclUniqueId is 128 bytes");
    memcpy(out128, &id, sizeof id);
    return ML_OK;
#else
    return ML_UNSUPPORTED;
#endif
}
