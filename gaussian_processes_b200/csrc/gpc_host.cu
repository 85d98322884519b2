// gpc_host.cu -- the five reductions of gp/ext/gp_c.pyx as host-pointer C entry points with exactly the
// reference's argument lists (gp_c.pyx:17, 34, 52, 70, 114): C-contiguous float64 host arrays in, outputs
// written in place.  A maintainer binds these from cgo / ctypes / Cython without any Python or torch on the
// path; the library stages the arrays into its own device pool, runs the same
// kernels the GP object uses on resident data, and copies the small results back.
//
// Algebra (nothing below is an N^3 chain unless the reference's quantity needs one):
//   a = Kiy (given), v = Ki y, b_i = dK_i a, with dK_i = Kj[i] (i < n_p) and dK_s = 2 s I
//   log_lh        Cholesky of K (log|K| = 2 sum log L_ii instead of the LU slogdet), MIN clamp
//   dloglh / dlh  t0_i = y^T Ki b_i (quadratic form), t1_i = tr(Ki dK_i) = sum(Ki o dK_i^T)        O(N^2)
//   d2lh          P_i = Ki dK_i on the DMMA GEMM (n_p products), then
//                 t1a = -u^T dK_j (Ki b_i) with u = Ki^T y,  t1b = a^T d2K_ij a,  t1c = -a^T dK_i (Ki dK_j v),
//                 tr(dKi_j dK_i) = -tr(P_j P_i),  tr(Ki d2K_ij)                                     O(N^3 n_p)
//   dm            dm_i = dKxox_i v - Kxox (Ki (dK_i v))                                             O(N^2 + M N)
#include <mutex>
#include <vector>
#include "../../include/gpb200.h"
#include "common.cuh"
#include "launch.h"

std::recursive_mutex& gpb_api_mutex();                 // api.cu: the library-wide lock of the staging state

namespace {

// ---- device arena (grow-only; used under the API lock) ---------------------------------------------------
char* g_arena = nullptr;
size_t g_arena_cap = 0;
struct Arena {
    size_t off = 0;
    double* take(size_t doubles) {
        double* p = reinterpret_cast<double*>(g_arena + off);
        off += (doubles * 8 + 255) / 256 * 256;
        return p;
    }
};
int arena_reserve(size_t bytes) {
    if (bytes <= g_arena_cap) return GPB_OK;
    if (g_arena) GPB_CUDA(cudaFree(g_arena));
    g_arena = nullptr; g_arena_cap = 0;
    GPB_CUDA(cudaMalloc(&g_arena, bytes));
    g_arena_cap = bytes;
    return GPB_OK;
}
inline long long rup(long long n) { return (n + GPB_NB - 1) / GPB_NB * GPB_NB; }

__global__ void pad_diag_kernel(double* A, long long n, long long npad, double v) {
    const long long i = n + (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < npad) A[i * npad + i] = v;
}
__global__ void transpose_kernel(const double* A, double* T, long long n) {       // T = A^T, n x n, ld = n
    __shared__ double tile[32][33];
    const long long bx = (long long)blockIdx.x * 32, by = (long long)blockIdx.y * 32;
    for (int r = threadIdx.y; r < 32; r += 8) {
        const long long gr = by + r, gc = bx + threadIdx.x;
        if (gr < n && gc < n) tile[r][threadIdx.x] = A[gr * n + gc];
    }
    __syncthreads();
    for (int r = threadIdx.y; r < 32; r += 8) {
        const long long gr = bx + r, gc = by + threadIdx.x;
        if (gr < n && gc < n) T[gr * n + gc] = tile[threadIdx.x][r];
    }
}
__global__ void __launch_bounds__(256) trace_dot_kernel(const double* A, long long ld, const double* u, const double* v,
                                                        long long n, double* out2) {
    // out2[0] = tr(A) (A may be null), out2[1] = u . v (u may be null)
    __shared__ double red[32];
    double st = 0.0, sd = 0.0;
    for (long long i = threadIdx.x; i < n; i += blockDim.x) {
        if (A) st += A[i * ld + i];
        if (u) sd += u[i] * v[i];
    }
    st = block_sum(st, red);
    sd = block_sum(sd, red);
    if (threadIdx.x == 0) { out2[0] = st; out2[1] = sd; }
}

// host [rows, cols] (contiguous) -> device [prows, pcols] zero padded (+ `diag` on the pad diagonal of a square)
int upload(double* dst, long long prows, long long pcols, const double* src, long long rows, long long cols,
           double pad_diag, cudaStream_t st) {
    if (prows != rows || pcols != cols) GPB_CUDA(cudaMemsetAsync(dst, 0, (size_t)prows * pcols * 8, st));
    if (rows && cols)
        GPB_CUDA(cudaMemcpy2DAsync(dst, (size_t)pcols * 8, src, (size_t)cols * 8, (size_t)cols * 8, (size_t)rows,
                                   cudaMemcpyHostToDevice, st));
    if (pad_diag != 0.0 && prows == pcols && prows > rows) {
        pad_diag_kernel<<<(unsigned)((prows - rows + 255) / 256), 256, 0, st>>>(dst, rows, prows, pad_diag);
        GPB_LAUNCH_CHECK("pad_diag_kernel");
    }
    return GPB_OK;
}
int transpose(const double* A, double* T, long long n, cudaStream_t st) {
    dim3 grid((unsigned)((n + 31) / 32), (unsigned)((n + 31) / 32));
    transpose_kernel<<<grid, dim3(32, 8), 0, st>>>(A, T, n);
    GPB_LAUNCH_CHECK("transpose_kernel");
    return GPB_OK;
}
int trace_dot(const double* A, long long ld, const double* u, const double* v, long long n, double* out2, cudaStream_t st) {
    trace_dot_kernel<<<1, 256, 0, st>>>(A, ld, u, v, n, out2);
    GPB_LAUNCH_CHECK("trace_dot_kernel");
    return GPB_OK;
}
#define GPC_TRY(call) do { int _s = (call); if (_s != GPB_OK) return _s; } while (0)

// (y^T Ki dK_i Kiy, tr(Ki dK_i)) for i < n_p and the noise row
int brackets(const double* y, const double* Ki, const double* Kj, const double* Kiy, double s, long long n_p,
             long long n, std::vector<double>& t0, std::vector<double>& t1) {
    cudaStream_t st = 0;
    const size_t mat = (size_t)n * n;
    GPC_TRY(arena_reserve((2 * mat + 8 * (size_t)n + 4096 + (size_t)gpb_reduce_blocks(n) * 2) * 8 + 16 * 256));
    Arena ar;
    double *dKi = ar.take(mat), *dK = ar.take(mat), *dy = ar.take(n), *da = ar.take(n), *w = ar.take(n);
    double *part = ar.take(gpb_reduce_blocks(n) + 8), *res = ar.take(2 * (n_p + 1) + 2);
    GPC_TRY(upload(dKi, n, n, Ki, n, n, 0.0, st));
    GPC_TRY(upload(dy, 1, n, y, 1, n, 0.0, st));
    GPC_TRY(upload(da, 1, n, Kiy, 1, n, 0.0, st));
    for (long long i = 0; i < n_p; i++) {
        GPC_TRY(upload(dK, n, n, Kj + i * mat, n, n, 0.0, st));
        GPC_TRY(gpb_launch_gemv(dK, n, n, n, da, w, 1.0, 0.0, st));                    // b_i = dK_i a
        GPC_TRY(gpb_launch_quadform(dy, dKi, n, w, n, part, res + 2 * i, st));         // y^T Ki b_i
        GPC_TRY(gpb_launch_trace_prod(dKi, n, dK, n, n, part, res + 2 * i + 1, st));   // tr(Ki dK_i)
    }
    GPC_TRY(gpb_launch_quadform(dy, dKi, n, da, n, part, res + 2 * n_p, st));          // y^T Ki a
    GPC_TRY(trace_dot(dKi, n, nullptr, nullptr, n, res + 2 * n_p + 1, st));            // tr Ki (+ unused dot)
    std::vector<double> h(2 * (n_p + 1) + 2);
    GPB_CUDA(cudaMemcpyAsync(h.data(), res, h.size() * 8, cudaMemcpyDeviceToHost, st));
    GPB_CUDA(cudaStreamSynchronize(st));
    t0.resize(n_p + 1); t1.resize(n_p + 1);
    for (long long i = 0; i < n_p; i++) { t0[i] = h[2 * i]; t1[i] = h[2 * i + 1]; }
    t0[n_p] = 2.0 * s * h[2 * n_p];                    // dK_s = 2 s I  (gp_c.pyx:45)
    t1[n_p] = 2.0 * s * h[2 * n_p + 1];
    return GPB_OK;
}

}  // namespace

extern "C" {

// gp_c.log_lh(y, K, Kiy) -> float   (gp_c.pyx:17-31)
int gpb_gp_c_log_lh(const double* y, const double* K, const double* Kiy, int64_t n, double* llh) {
    std::lock_guard<std::recursive_mutex> lk(gpb_api_mutex());
    GPB_REQUIRE(y && K && Kiy && llh && n >= 1, "bad argument");
    cudaStream_t st = 0;
    const long long np_ = rup(n);
    const size_t mat = (size_t)np_ * np_;
    GPC_TRY(arena_reserve((2 * mat + 4 * (size_t)np_ + 64) * 8 + 16 * 256));
    Arena ar;
    double *L = ar.take(mat), *W = ar.take(mat), *dy = ar.take(np_), *da = ar.take(np_), *out3 = ar.take(4);
    int* info = reinterpret_cast<int*>(ar.take(2));
    GPC_TRY(upload(L, np_, np_, K, n, n, 1.0, st));
    GPC_TRY(upload(dy, 1, np_, y, 1, n, 0.0, st));
    GPC_TRY(upload(da, 1, np_, Kiy, 1, n, 0.0, st));
    GPC_TRY(gpb_launch_potrf(L, np_, np_, 0, 1, W, np_, 0, nullptr, 0, 0, info, st, n, true, true));
    GPC_TRY(gpb_launch_loglh(L, n, np_, 0, 1, dy, 0, da, 0, info, out3, st));
    double h[3];
    GPB_CUDA(cudaMemcpyAsync(h, out3, 24, cudaMemcpyDeviceToHost, st));
    GPB_CUDA(cudaStreamSynchronize(st));
    *llh = h[0];
    return GPB_OK;
}

// gp_c.dloglh_dtheta(y, Ki, Kj, Kiy, s, dloglh)   (gp_c.pyx:34-49); Kj: [n_p, n, n], dloglh: [n_p + 1]
int gpb_gp_c_dloglh_dtheta(const double* y, const double* Ki, const double* Kj, const double* Kiy, double s,
                           int64_t n_p, int64_t n, double* dloglh) {
    std::lock_guard<std::recursive_mutex> lk(gpb_api_mutex());
    GPB_REQUIRE(y && Ki && (Kj || n_p == 0) && Kiy && dloglh && n >= 1 && n_p >= 0, "bad argument");
    std::vector<double> t0, t1;
    GPC_TRY(brackets(y, Ki, Kj, Kiy, s, n_p, n, t0, t1));
    for (int64_t i = 0; i <= n_p; i++) dloglh[i] = 0.5 * t0[i] + -0.5 * t1[i];
    return GPB_OK;
}

// gp_c.dlh_dtheta(y, Ki, Kj, Kiy, s, lh, dlh)   (gp_c.pyx:52-67)
int gpb_gp_c_dlh_dtheta(const double* y, const double* Ki, const double* Kj, const double* Kiy, double s, double lh,
                        int64_t n_p, int64_t n, double* dlh) {
    std::lock_guard<std::recursive_mutex> lk(gpb_api_mutex());
    GPB_REQUIRE(y && Ki && (Kj || n_p == 0) && Kiy && dlh && n >= 1 && n_p >= 0, "bad argument");
    std::vector<double> t0, t1;
    GPC_TRY(brackets(y, Ki, Kj, Kiy, s, n_p, n, t0, t1));
    for (int64_t i = 0; i <= n_p; i++) dlh[i] = 0.5 * lh * (t0[i] - t1[i]);
    return GPB_OK;
}

// gp_c.d2lh_dtheta2(y, Ki, Kj, Kh, Kiy, s, lh, dlh, d2lh)   (gp_c.pyx:70-111)
// Kj: [n_p, n, n], Kh: [n_p, n_p, n, n], dlh: [n_p + 1], d2lh: [n_p + 1, n_p + 1]
int gpb_gp_c_d2lh_dtheta2(const double* y, const double* Ki, const double* Kj, const double* Kh, const double* Kiy,
                          double s, double lh, const double* dlh, int64_t n_p, int64_t n, double* d2lh) {
    std::lock_guard<std::recursive_mutex> lk(gpb_api_mutex());
    GPB_REQUIRE(y && Ki && (Kj || n_p == 0) && (Kh || n_p == 0) && Kiy && dlh && d2lh && n >= 1 && n_p >= 0 && n_p <= 8,
                "bad argument");
    cudaStream_t st = 0;
    const long long np_ = rup(n), nth = n_p + 1;
    const size_t mat = (size_t)np_ * np_;
    const size_t nres = (size_t)(4 * nth + 5 * nth * nth + 8);
    GPC_TRY(arena_reserve(((3 + 2 * (size_t)n_p) * mat + (8 + 4 * (size_t)nth) * np_ + nres + gpb_reduce_blocks(np_) + 64) * 8 + 64 * 256));
    Arena ar;
    double *dKi = ar.take(mat), *T = ar.take(mat), *H = ar.take(mat);
    std::vector<double*> dK(n_p), P(n_p);
    for (auto& p : dK) p = ar.take(mat);
    for (auto& p : P) p = ar.take(mat);
    double *dy = ar.take(np_), *da = ar.take(np_), *u = ar.take(np_), *v = ar.take(np_), *tmp = ar.take(np_);
    std::vector<double*> b(nth), c(nth), f(nth);           // b_i = dK_i a ; c_i = Ki b_i ; f_j = Ki dK_j v
    for (auto& p : b) p = ar.take(np_);
    for (auto& p : c) p = ar.take(np_);
    for (auto& p : f) p = ar.take(np_);
    double *part = ar.take(gpb_reduce_blocks(np_) + 8), *res = ar.take(nres);
    GPB_CUDA(cudaMemsetAsync(res, 0, nres * 8, st));
    GPC_TRY(upload(dKi, np_, np_, Ki, n, n, 0.0, st));
    GPC_TRY(upload(dy, 1, np_, y, 1, n, 0.0, st));
    GPC_TRY(upload(da, 1, np_, Kiy, 1, n, 0.0, st));
    for (long long i = 0; i < n_p; i++) GPC_TRY(upload(dK[i], np_, np_, Kj + i * (size_t)n * n, n, n, 0.0, st));
    GPC_TRY(transpose(dKi, T, np_, st));
    GPC_TRY(gpb_launch_gemv(T, np_, np_, np_, dy, u, 1.0, 0.0, st));                   // u = Ki^T y
    GPC_TRY(gpb_launch_gemv(dKi, np_, np_, np_, dy, v, 1.0, 0.0, st));                 // v = Ki y
    for (long long i = 0; i < nth; i++) {
        if (i < n_p) {
            GPC_TRY(gpb_launch_gemv(dK[i], np_, np_, np_, da, b[i], 1.0, 0.0, st));
            GPC_TRY(gpb_launch_gemv(dK[i], np_, np_, np_, v, tmp, 1.0, 0.0, st));
        } else {
            GPB_CUDA(cudaMemcpyAsync(b[i], da, np_ * 8, cudaMemcpyDeviceToDevice, st));
            GPB_CUDA(cudaMemcpyAsync(tmp, v, np_ * 8, cudaMemcpyDeviceToDevice, st));
        }
        // the noise row carries its factor 2 s on the host side (b_s = 2 s a, dK_s v = 2 s v)
        GPC_TRY(gpb_launch_gemv(dKi, np_, np_, np_, b[i], c[i], 1.0, 0.0, st));        // c_i = Ki b_i
        GPC_TRY(gpb_launch_gemv(dKi, np_, np_, np_, tmp, f[i], 1.0, 0.0, st));         // f_i = Ki dK_i v
    }
    // P_i = Ki dK_i  (NT product with B = dK_i^T)
    for (long long i = 0; i < n_p; i++) {
        GPC_TRY(transpose(dK[i], T, np_, st));
        GpbGemm g = gpb_gemm_default();
        g.A = dKi; g.lda = np_; g.B = T; g.ldb = np_; g.C = P[i]; g.ldc = np_;
        g.M = g.N = g.K = (int)np_;
        GPC_TRY(gpb_launch_gemm(g, 1, st));
    }
    // scalars.  layout of res: q0[i] = y^T Ki b_i | tr1[i] = tr(Ki dK_i) | then per (i,j): t1a, t1b, t1c, trPP, trKH
    double* q0 = res; double* tr1 = res + nth; double* ij = res + 2 * nth; double* misc = res + 2 * nth + 5 * nth * nth;
    for (long long i = 0; i < nth; i++) {
        GPC_TRY(gpb_launch_quadform(dy, dKi, np_, b[i], n, part, q0 + i, st));
        if (i < n_p) GPC_TRY(gpb_launch_trace_prod(dKi, np_, dK[i], np_, n, part, tr1 + i, st));
    }
    GPC_TRY(trace_dot(dKi, np_, da, da, n, misc, st));                                 // tr Ki, a . a
    GPC_TRY(gpb_launch_trace_prod(dKi, np_, dKi, np_, n, part, misc + 2, st));         // tr(Ki Ki)
    for (long long i = 0; i < nth; i++)
        for (long long j = 0; j < nth; j++) {
            double* o = ij + 5 * (i * nth + j);
            // t1a = -u^T dK_j c_i
            if (j < n_p) {
                GPC_TRY(gpb_launch_quadform(u, dK[j], np_, c[i], n, part, o + 0, st));
            } else {
                GPC_TRY(trace_dot(nullptr, 0, u, c[i], n, misc + 4, st));
                GPB_CUDA(cudaMemcpyAsync(o + 0, misc + 5, 8, cudaMemcpyDeviceToDevice, st));
            }
            // t1c = -a^T dK_i f_j
            if (i < n_p) {
                GPC_TRY(gpb_launch_quadform(da, dK[i], np_, f[j], n, part, o + 2, st));
            } else {
                GPC_TRY(trace_dot(nullptr, 0, da, f[j], n, misc + 4, st));
                GPB_CUDA(cudaMemcpyAsync(o + 2, misc + 5, 8, cudaMemcpyDeviceToDevice, st));
            }
            if (i < n_p && j < n_p) {
                GPC_TRY(upload(H, np_, np_, Kh + (size_t)(i * n_p + j) * n * n, n, n, 0.0, st));
                GPC_TRY(gpb_launch_quadform(da, H, np_, da, n, part, o + 1, st));      // t1b = a^T d2K a
                GPC_TRY(gpb_launch_trace_prod(dKi, np_, H, np_, n, part, o + 4, st));  // tr(Ki d2K)
                GPC_TRY(gpb_launch_trace_prod(P[j], np_, P[i], np_, n, part, o + 3, st));   // tr(P_j P_i)
            } else if (i < n_p) {          // j = s: P_s = 2 s Ki -> tr(Ki P_i) (factor on the host)
                GPC_TRY(gpb_launch_trace_prod(dKi, np_, P[i], np_, n, part, o + 3, st));
            } else if (j < n_p) {
                GPC_TRY(gpb_launch_trace_prod(P[j], np_, dKi, np_, n, part, o + 3, st));
            }
        }
    std::vector<double> h(nres);
    GPB_CUDA(cudaMemcpyAsync(h.data(), res, nres * 8, cudaMemcpyDeviceToHost, st));
    GPB_CUDA(cudaStreamSynchronize(st));
    const double trKi = h[2 * nth + 5 * nth * nth], aa = h[2 * nth + 5 * nth * nth + 1], trKK = h[2 * nth + 5 * nth * nth + 2];
    const double s2 = 2.0 * s;
    for (long long i = 0; i < nth; i++) {
        const double fi = (i < n_p) ? 1.0 : s2;                    // factor of dK_i carried on the host for the noise row
        const double ydKy = fi * h[i];                             // y^T Ki dK_i Kiy
        const double trP = (i < n_p) ? h[nth + i] : s2 * trKi;     // tr(Ki dK_i)
        const double r_i = ydKy - trP;                             // gp_c.pyx:92-93
        for (long long j = 0; j < nth; j++) {
            const double fj = (j < n_p) ? 1.0 : s2;
            const double* o = &h[2 * nth + 5 * (i * nth + j)];
            const double t1a = -fi * fj * o[0];
            const double t1c = -fi * fj * o[2];
            double t1b, trPP, trKH;
            if (i < n_p && j < n_p) { t1b = o[1]; trPP = o[3]; trKH = o[4]; }
            else if (i == n_p && j == n_p) { t1b = 2.0 * aa; trPP = s2 * s2 * trKK; trKH = 2.0 * trKi; }   // d2k = 2 I
            else { t1b = 0.0; trPP = s2 * o[3]; trKH = 0.0; }
            const double t1 = lh * (t1a + t1b + t1c - (-trPP + trKH));
            d2lh[i * nth + j] = 0.5 * (dlh[j] * r_i + t1);
        }
    }
    return GPB_OK;
}

// gp_c.dm_dtheta(y, Ki, Kj, Kjxo, Kxox, s, dm)   (gp_c.pyx:114-131)
// Kj: [n_p, n, n], Kjxo: [n_p, m, n], Kxox: [m, n], dm: [n_p + 1, m]
int gpb_gp_c_dm_dtheta(const double* y, const double* Ki, const double* Kj, const double* Kjxo, const double* Kxox,
                       double s, int64_t n_p, int64_t n, int64_t m, double* dm) {
    std::lock_guard<std::recursive_mutex> lk(gpb_api_mutex());
    GPB_REQUIRE(y && Ki && (Kj || n_p == 0) && (Kjxo || n_p == 0 || m == 0) && (Kxox || m == 0) && (dm || m == 0) &&
                n >= 1 && n_p >= 0 && m >= 0, "bad argument");
    if (m == 0) return GPB_OK;
    cudaStream_t st = 0;
    const size_t mat = (size_t)n * n, rect = (size_t)m * n;
    GPC_TRY(arena_reserve((2 * mat + 2 * rect + 6 * (size_t)n + (size_t)(n_p + 2) * m + 64) * 8 + 32 * 256));
    Arena ar;
    double *dKi = ar.take(mat), *dK = ar.take(mat), *dX = ar.take(rect), *dJ = ar.take(rect);
    double *dy = ar.take(n), *v = ar.take(n), *bb = ar.take(n), *cc = ar.take(n), *out = ar.take((size_t)(n_p + 1) * m);
    GPC_TRY(upload(dKi, n, n, Ki, n, n, 0.0, st));
    GPC_TRY(upload(dX, m, n, Kxox, m, n, 0.0, st));
    GPC_TRY(upload(dy, 1, n, y, 1, n, 0.0, st));
    GPC_TRY(gpb_launch_gemv(dKi, n, n, n, dy, v, 1.0, 0.0, st));                        // v = Ki y
    for (int64_t i = 0; i <= n_p; i++) {
        double* o = out + i * m;
        if (i < n_p) {
            GPC_TRY(upload(dK, n, n, Kj + i * mat, n, n, 0.0, st));
            GPC_TRY(upload(dJ, m, n, Kjxo + i * rect, m, n, 0.0, st));
            GPC_TRY(gpb_launch_gemv(dK, n, n, n, v, bb, 1.0, 0.0, st));                 // dK_i v
            GPC_TRY(gpb_launch_gemv(dKi, n, n, n, bb, cc, 1.0, 0.0, st));               // Ki dK_i v
            GPC_TRY(gpb_launch_gemv(dJ, m, n, n, v, o, 1.0, 0.0, st));                  // dKxox_i v
            GPC_TRY(gpb_launch_gemv(dX, m, n, n, cc, o, -1.0, 1.0, st));                // - Kxox (Ki dK_i v)
        } else {
            GPC_TRY(gpb_launch_gemv(dKi, n, n, n, v, cc, 1.0, 0.0, st));                // Ki v
            GPC_TRY(gpb_launch_gemv(dX, m, n, n, cc, o, -2.0 * s, 0.0, st));            // dK_s = 2 s I, dKxox_s = 0
        }
    }
    GPB_CUDA(cudaMemcpyAsync(dm, out, (size_t)(n_p + 1) * m * 8, cudaMemcpyDeviceToHost, st));
    GPB_CUDA(cudaStreamSynchronize(st));
    return GPB_OK;
}

}  // extern "C"
