// reduce.cu -- HBM-bound O(N^2) helpers over materialised matrices: GEMV,
// trace of a product and quadratic forms.  They serve d2lh_dtheta2 / dm_dtheta and
// the host-buffer drop-ins of gp_c.pyx, where the caller supplies dense arrays.
// Reductions are two-stage (per-block partials, then one block) so results are
// deterministic for a given shape.
#include "common.cuh"
#include "launch.h"

namespace {

// y = alpha * A x + beta * y ; one warp per row
__global__ void __launch_bounds__(256)
gemv_kernel(const double* A, long long rows, long long cols, long long lda, const double* x, double* y,
            double alpha, double beta) {
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    for (long long r = (long long)blockIdx.x * 8 + wid; r < rows; r += (long long)gridDim.x * 8) {
        const double* row = A + r * lda;
        double s = 0.0;
        for (long long c = lane; c < cols; c += 32) s += row[c] * x[c];
        s = warp_sum(s);
        if (lane == 0) y[r] = alpha * s + (beta != 0.0 ? beta * y[r] : 0.0);
    }
}

// partial[blk] = sum over a 32x32 tile set of A[a,b] * B[b,a]
__global__ void __launch_bounds__(256)
trace_prod_kernel(const double* A, long long lda, const double* B, long long ldb, long long n, double* partial) {
    __shared__ double tile[32][33];
    __shared__ double red[32];
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;      // 32 x 8
    const long long nt = (n + 31) / 32;
    double s = 0.0;
    for (long long tidx = blockIdx.x; tidx < nt * nt; tidx += gridDim.x) {
        const long long ta = tidx / nt, tb = tidx % nt;
        // load B tile (rows tb*32.., cols ta*32..) transposed through shared memory
        __syncthreads();
        for (int r = ty; r < 32; r += 8) {
            const long long gr = tb * 32 + r, gc = ta * 32 + tx;
            tile[r][tx] = (gr < n && gc < n) ? B[gr * ldb + gc] : 0.0;
        }
        __syncthreads();
        for (int r = ty; r < 32; r += 8) {
            const long long ga = ta * 32 + r, gb = tb * 32 + tx;
            if (ga < n && gb < n) s += A[ga * lda + gb] * tile[tx][r];
        }
    }
    s = block_sum(s, red);
    if (threadIdx.x == 0) partial[blockIdx.x] = s;
}

// partial[blk] = sum_r u[r] * sum_c M[r,c] v[c]
__global__ void __launch_bounds__(256)
quadform_kernel(const double* u, const double* M, long long ldm, const double* v, long long n, double* partial) {
    __shared__ double red[32];
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    double s = 0.0;
    for (long long r = (long long)blockIdx.x * 8 + wid; r < n; r += (long long)gridDim.x * 8) {
        const double* row = M + r * ldm;
        double q = 0.0;
        for (long long c = lane; c < n; c += 32) q += row[c] * v[c];
        s += u[r] * q;
    }
    s = block_sum(s, red);
    if (threadIdx.x == 0) partial[blockIdx.x] = s;
}

__global__ void __launch_bounds__(256) sum1_kernel(const double* partial, int nblk, double* out) {
    __shared__ double red[32];
    double s = 0.0;
    for (int k = threadIdx.x; k < nblk; k += blockDim.x) s += partial[k];
    s = block_sum(s, red);
    if (threadIdx.x == 0) *out = s;
}

}  // namespace

int gpb_reduce_blocks(long long n) {
    long long nb = (n + 7) / 8;
    if (nb > 148 * 4) nb = 148 * 4;
    if (nb < 1) nb = 1;
    return (int)nb;
}

int gpb_launch_gemv(const double* A, long long rows, long long cols, long long lda, const double* x,
                    double* y, double alpha, double beta, cudaStream_t st) {
    if (rows == 0) return GPB_OK;
    gemv_kernel<<<gpb_reduce_blocks(rows), 256, 0, st>>>(A, rows, cols, lda, x, y, alpha, beta);
    GPB_LAUNCH_CHECK("gemv_kernel");
    return GPB_OK;
}

int gpb_launch_trace_prod(const double* A, long long lda, const double* B, long long ldb, long long n,
                          double* partial, double* out, cudaStream_t st) {
    const int nb = gpb_reduce_blocks(n);
    trace_prod_kernel<<<nb, 256, 0, st>>>(A, lda, B, ldb, n, partial);
    GPB_LAUNCH_CHECK("trace_prod_kernel");
    sum1_kernel<<<1, 256, 0, st>>>(partial, nb, out);
    GPB_LAUNCH_CHECK("sum1_kernel");
    return GPB_OK;
}

int gpb_launch_quadform(const double* u, const double* M, long long ldm, const double* v, long long n,
                        double* partial, double* out, cudaStream_t st) {
    const int nb = gpb_reduce_blocks(n);
    quadform_kernel<<<nb, 256, 0, st>>>(u, M, ldm, v, n, partial);
    GPB_LAUNCH_CHECK("quadform_kernel");
    sum1_kernel<<<1, 256, 0, st>>>(partial, nb, out);
    GPB_LAUNCH_CHECK("sum1_kernel");
    return GPB_OK;
}
