// microbench.cu -- FP64 issue-rate microbenchmarks (DMMA.8x8x4 and DFMA).  The measured
// DMMA rate is the roofline denominator bench.py uses for the factorisation kernels,
// because MEASURED_PEAKS.json carries no fp64 figure.
#include "../../include/gpb200.h"
#include "common.cuh"

namespace {
template <bool USE_DMMA>
__global__ void __launch_bounds__(256) fp64_peak_kernel(double* out, const double* in, int iters) {
    const double a = in[threadIdx.x & 7], b = in[8 + (threadIdx.x & 7)];
    double c[16][2];
#pragma unroll
    for (int q = 0; q < 16; q++) c[q][0] = c[q][1] = (double)q;
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int q = 0; q < 16; q++) {
            if (USE_DMMA) dmma884(c[q][0], c[q][1], a, b);
            else { c[q][0] = fma(a, b, c[q][0]); c[q][1] = fma(b, a, c[q][1]); }
        }
    }
    double s = 0;
#pragma unroll
    for (int q = 0; q < 16; q++) s += c[q][0] + c[q][1];
    out[(size_t)blockIdx.x * blockDim.x + threadIdx.x] = s;
}
}  // namespace

extern "C" int gpb_microbench_fp64(int use_dmma, int iters, double* tflops, double* ms_out) {
    const int ctas = 148 * 8, threads = 256;
    double *out = nullptr, *in = nullptr;
    GPB_CUDA(cudaMalloc(&out, (size_t)ctas * threads * 8));
    GPB_CUDA(cudaMalloc(&in, 16 * 8));
    double h[16];
    for (int i = 0; i < 16; i++) h[i] = 1e-3 * (i + 1);
    GPB_CUDA(cudaMemcpy(in, h, sizeof(h), cudaMemcpyHostToDevice));
    cudaEvent_t e0, e1;
    GPB_CUDA(cudaEventCreate(&e0));
    GPB_CUDA(cudaEventCreate(&e1));
    float best = 1e30f;
    for (int rep = 0; rep < 4; rep++) {
        GPB_CUDA(cudaEventRecord(e0, 0));
        if (use_dmma) fp64_peak_kernel<true><<<ctas, threads>>>(out, in, iters);
        else fp64_peak_kernel<false><<<ctas, threads>>>(out, in, iters);
        GPB_LAUNCH_CHECK("fp64_peak_kernel");
        GPB_CUDA(cudaEventRecord(e1, 0));
        GPB_CUDA(cudaEventSynchronize(e1));
        float ms;
        GPB_CUDA(cudaEventElapsedTime(&ms, e0, e1));
        if (rep > 0 && ms < best) best = ms;
    }
    const double warps = (double)ctas * threads / 32.0;
    const double flops_per_warp_iter = use_dmma ? 16.0 * 512.0 : 32.0 * 64.0;
    *tflops = warps * flops_per_warp_iter * iters / (best * 1e-3) / 1e12;
    *ms_out = best;
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    cudaFree(out);
    cudaFree(in);
    return GPB_OK;
}
