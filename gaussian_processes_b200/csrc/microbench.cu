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

// ---- dependent-issue latencies (cycles) of the instructions on the factorisation's serial paths ----
namespace {
__global__ void latency_kernel(double* out, double seed) {
    const int lane = threadIdx.x;
    double x = seed + lane * 1e-9, y = 1.0 + seed;
    long long t0, t1;
    constexpr int R = 256;
    // DFMA chain
    t0 = clock64();
#pragma unroll 16
    for (int i = 0; i < R; i++) x = fma(x, y, 1e-9);
    t1 = clock64();
    out[0] = (double)(t1 - t0) / R;
    // rsqrt chain
    double r = 2.0 + x * 1e-30;
    t0 = clock64();
#pragma unroll 8
    for (int i = 0; i < R; i++) r = rsqrt(r) + 1.5;
    t1 = clock64();
    out[1] = (double)(t1 - t0) / R;     // includes one DADD
    // 64-bit shuffle chain
    double s = r;
    t0 = clock64();
#pragma unroll 16
    for (int i = 0; i < R; i++) s = __shfl_sync(0xffffffffu, s, (lane + 1) & 31);
    t1 = clock64();
    out[2] = (double)(t1 - t0) / R;
    // dependent DMMA chain
    double c0 = s, c1 = x;
    t0 = clock64();
#pragma unroll 16
    for (int i = 0; i < R; i++) dmma884(c0, c1, 1e-3, 1e-3);
    t1 = clock64();
    out[3] = (double)(t1 - t0) / R;
    // 8 independent DMMA accumulators from one warp: issue interval
    double a[8][2];
#pragma unroll
    for (int q = 0; q < 8; q++) a[q][0] = a[q][1] = c0 + q;
    t0 = clock64();
#pragma unroll 4
    for (int i = 0; i < R / 8; i++)
#pragma unroll
        for (int q = 0; q < 8; q++) dmma884(a[q][0], a[q][1], 1e-3, 1e-3);
    t1 = clock64();
    out[4] = (double)(t1 - t0) / R;
    // sqrt + divide chain (what rsqrt replaced)
    double q2 = 2.0 + c1 * 1e-30;
    t0 = clock64();
#pragma unroll 8
    for (int i = 0; i < R; i++) q2 = 1.0 / sqrt(q2) + 1.5;
    t1 = clock64();
    out[5] = (double)(t1 - t0) / R;
    // shared-memory store -> load round trip
    __shared__ double buf[64];
    double z = q2;
    t0 = clock64();
#pragma unroll 8
    for (int i = 0; i < R; i++) {
        buf[lane] = z;
        __syncwarp();
        z = buf[(lane + 1) & 31] + 1.0;
        __syncwarp();
    }
    t1 = clock64();
    out[6] = (double)(t1 - t0) / R;
    double sum = x + r + s + c0 + c1 + q2 + z;
#pragma unroll
    for (int q = 0; q < 8; q++) sum += a[q][0] + a[q][1];
    out[8 + lane] = sum;
}
}  // namespace

extern "C" int gpb_microbench_latency(double* out7) {
    double* d = nullptr;
    GPB_CUDA(cudaMalloc(&d, 64 * 8));
    for (int rep = 0; rep < 2; rep++) {
        latency_kernel<<<1, 32>>>(d, 0.5);
        GPB_LAUNCH_CHECK("latency_kernel");
    }
    GPB_CUDA(cudaMemcpy(out7, d, 7 * 8, cudaMemcpyDeviceToHost));
    cudaFree(d);
    return GPB_OK;
}
