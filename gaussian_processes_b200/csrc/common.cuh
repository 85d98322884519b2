// common.cuh -- shared device helpers for the sm_100a GP hot path.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <math.h>

#define GPB_NB 128            // block size of the blocked factorisation / GEMM tile edge
#define GPB_MAX_SLICES 13     // 1 + n_p + n_p^2 for the periodic kernel

// MIN = log(exp2(-1022 + 4))  (reference: gaussian_c.pyx:15, gp_c.pyx:14, gp.py:17)
#define GPB_MIN_LOG (-705.6238298100243)

// status codes of the C ABI
#define GPB_OK 0
#define GPB_ERR_ARG (-1)
#define GPB_ERR_CUDA (-2)

void gpb_set_error(const char* fmt, ...);
int gpb_check_cuda(cudaError_t e, const char* what);

#define GPB_CUDA(call)                                              \
    do {                                                            \
        int _st = gpb_check_cuda((call), #call);                    \
        if (_st != GPB_OK) return _st;                              \
    } while (0)

extern long long g_gpb_launches;   // kernels launched by this library (bench.py gpu_launches)
#define GPB_LAUNCH_CHECK(name)                                      \
    do {                                                            \
        g_gpb_launches++;                                           \
        GPB_CUDA(cudaPeekAtLastError());                            \
    } while (0)

// ---------------------------------------------------------------------------
// optional per-kernel-class device timing (bench.py roofline): when enabled, every
// launch of a class is bracketed by CUDA events on its own stream.
// ---------------------------------------------------------------------------
enum GpbKernelClass { GPB_KC_GEMM = 0, GPB_KC_DIAG = 1, GPB_KC_BUILD = 2, GPB_KC_SOLVE = 3,
                      GPB_KC_REDUCE = 4, GPB_KC_MISC = 5, GPB_KC_COUNT = 6 };
extern int g_gpb_profile;
void gpb_prof_begin(int cls, cudaStream_t st);
void gpb_prof_end(int cls, cudaStream_t st);
struct GpbProfScope {
    int cls;
    cudaStream_t st;
    GpbProfScope(int c, cudaStream_t s) : cls(c), st(s) { if (g_gpb_profile) gpb_prof_begin(cls, st); }
    ~GpbProfScope() { if (g_gpb_profile) gpb_prof_end(cls, st); }
};

#define GPB_REQUIRE(cond, msg)                                      \
    do {                                                            \
        if (!(cond)) {                                              \
            gpb_set_error("%s: %s", __func__, msg);                 \
            return GPB_ERR_ARG;                                     \
        }                                                           \
    } while (0)

// ---------------------------------------------------------------------------
// kernel-function parameters: everything the element formulas need, precomputed
// on the host once per (kernel, theta).  kind 0 = Gaussian, 1 = Periodic.
// ---------------------------------------------------------------------------
struct KParams {
    int kind;
    int pad_;
    double c1;        // gaussian: -0.5/w^2 ; periodic: -2/w^2
    double half_ip;   // periodic: 0.5/p
    double k0;        // K coefficient
    double j[3][2];   // jacobian coefficients
    double h[6][4];   // hessian coefficients (unique slices)
    double s2;        // s^2 (index-diagonal add for Kxx), 0 if unused
    double s;         // noise standard deviation (gradient row of s)
};

#ifdef __CUDACC__

__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// block-wide sum; result valid in thread 0.  `sh` must hold >= 32 doubles.
__device__ __forceinline__ double block_sum(double v, double* sh) {
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    v = warp_sum(v);
    __syncthreads();
    if (lane == 0) sh[wid] = v;
    __syncthreads();
    const int nw = (blockDim.x + 31) >> 5;
    v = (threadIdx.x < nw) ? sh[threadIdx.x] : 0.0;
    if (wid == 0) v = warp_sum(v);
    return v;
}

__device__ __forceinline__ void cp_async16(void* smem, const void* gmem) {
    unsigned s = (unsigned)__cvta_generic_to_shared(smem);
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(s), "l"(gmem));
}
__device__ __forceinline__ void cp_async8(void* smem, const void* gmem) {
    unsigned s = (unsigned)__cvta_generic_to_shared(smem);
    asm volatile("cp.async.ca.shared.global [%0], [%1], 8;\n" ::"r"(s), "l"(gmem));
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;\n" ::); }
template <int N>
__device__ __forceinline__ void cp_async_wait() {
    asm volatile("cp.async.wait_group %0;\n" ::"n"(N));
}

// FP64 tensor-core MMA: D(8x8) += A(8x4, row) * B(4x8, col).  SASS: DMMA.8x8x4.
// lane = 4*g + t :  a = A[g][t],  b = B[t][g],  c0,c1 = C[g][2t], C[g][2t+1]
__device__ __forceinline__ void dmma884(double& c0, double& c1, double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
                 : "+d"(c0), "+d"(c1)
                 : "d"(a), "d"(b));
}

__device__ __forceinline__ int ld_acquire(const int* p) {
    int v;
    asm volatile("ld.acquire.gpu.global.s32 %0, [%1];\n" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_release(int* p, int v) {
    asm volatile("st.release.gpu.global.s32 [%0], %1;\n" ::"l"(p), "r"(v) : "memory");
}

#endif  // __CUDACC__
