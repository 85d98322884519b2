// launch.h -- host-side launchers shared between translation units (internal).
#pragma once
#include "common.cuh"

// C = beta*C + alpha * A * B^T over a tile-level k-range (see gemm.cu)
struct GpbGemm {
    const double* A;
    const double* B;
    double* C;
    double* Ct;              // optional mirrored (transposed) store, may alias C
    long long lda, ldb, ldc, ldct;
    long long sA, sB, sC, sCt;   // outer batch strides (blockIdx.z / nb1: independent GPs / candidates)
    long long tA, tB, tC, tCt;   // inner batch strides (blockIdx.z % nb1: sibling blocks of one matrix)
    int nb1;                     // inner batch count (>= 1)
    int M, N, K;
    double alpha, beta;
    int a_tri, b_tri;        // 0 dense | 1 lower (nonzero k <= row + off) | 2 upper (k >= row + off)
    int a_off, b_off;
    int lower_only;          // only tiles with tj <= ti (square tile grid); off-diagonal tiles mirrored into Ct
};

inline GpbGemm gpb_gemm_default() {
    GpbGemm g;
    g.A = g.B = nullptr; g.C = g.Ct = nullptr;
    g.lda = g.ldb = g.ldc = g.ldct = 0;
    g.sA = g.sB = g.sC = g.sCt = 0;
    g.tA = g.tB = g.tC = g.tCt = 0;
    g.nb1 = 1;
    g.M = g.N = g.K = 0;
    g.alpha = 1.0; g.beta = 0.0;
    g.a_tri = g.b_tri = 0; g.a_off = g.b_off = 0;
    g.lower_only = 0;
    return g;
}

int gpb_launch_gemm(const GpbGemm& p, int batch, cudaStream_t st);

#define GPB_RED_MAXS 6      // slices per reduction pass
#define GPB_RED_WIDTH 16    // doubles per partial row: t0[6] | t1[6] | tr | a.a | 0 0

// Gradient brackets in the epilogue of the lauum GEMM (Ki = V V^T, lower tiles): while a tile of K^-1 is in registers
// the kernel-derivative tiles are regenerated from x and  t0[q] += a_r a_c dK_q(r,c),  t1[q] += Ki_rc dK_q(r,c),
// tr += Ki_rr,  aa += a_r^2  are accumulated (gp/ext/gp_c.pyx:41-49); one row of 16 partials per CTA.
struct GpbLauumFuse {
    int kind;                // GPB_GAUSSIAN | GPB_PERIODIC
    int store;               // 0: K^-1 itself is not needed, its tiles are not written
    long long n;             // valid extent (identity pad excluded)
    KParams P;               // parameters by value (one object) ...
    const KParams* Pb;       // ... or one entry per batch member
    const double* x;
    const double* alpha;
    long long astride;
    double* partial;         // [batch][CTAs per matrix][16]
};
// Ki = V V^T with the fused gradient reduction; out16[b] = {t0[6] | t1[6] | tr | a.a | 0 0}.  Returns GPB_ERR_ARG
// (nothing launched) when the fused path is not available (gemm_bm / gemm_impl options): use lauum + grad_reduce.
bool gpb_lauum_grad_available();
int gpb_launch_lauum_grad(const double* V, long long n, long long ldv, long long sV, int batch, double* Ki,
                          long long ldk, long long sK, const GpbLauumFuse& f, double* out16, cudaStream_t st);
int gpb_launch_sum_partials(const double* partial, int nblk, int batch, double* out16, cudaStream_t st);

// run-time tuning knobs (api.cu): "eval_streams", "gemm_bm", "potrf_inner"; 0 = default
int gpb_get_option(const char* name);

int gpb_launch_build(int kind, const KParams* P, const KParams* Pb, int batch, const double* x1,
                     long long n1, const double* x2, long long n2, long long rows, long long cols,
                     double* const* out, long long ld, long long bstride, int add_diag,
                     int pad_identity, cudaStream_t st, int lower_only = 0);

int gpb_launch_fused_matvec(int kind, const KParams* P, const KParams* Pb, int batch,
                            const double* x1, long long n1, const double* x2, long long n2,
                            int npairs, const int* slice, const int* outidx, const double* coef,
                            const double* const* vec, int nout, double* const* out,
                            long long vstride, long long ostride, cudaStream_t st);

int gpb_grad_reduce_blocks(long long n);
int gpb_launch_grad_reduce(int kind, const KParams* P, const KParams* Pb, int batch, const double* x,
                           long long n, const double* Ki, long long ldk, long long kstride,
                           const double* alpha, long long astride, int nsl, const int* slices,
                           double* partial, double* out16, cudaStream_t st);

int gpb_launch_post_var(int kind, const KParams* P, const double* Z, long long ldz, long long m, long long n,
                        double* out, cudaStream_t st);

// potrf.cu
int gpb_launch_potrf(double* A, long long n, long long ld, long long sA, int batch, double* W,
                     long long ldw, long long sW, double* V, long long ldv, long long sV, int* info,
                     cudaStream_t st, long long n_valid = 0, bool zero_blocks = true, bool single_chain = false);
// n_valid: rows that are not identity pad (0 = n); zero_blocks = false: the caller writes the structural
// zeros of the inverted diagonal blocks itself (gpb_launch_small_tail does); single_chain: one matrix whose
// factorisation is not overlapped with others (right-looking panel steps, see potrf.cu)
// chain.cu: one matrix (2 <= n/128 <= 64) factored by a single persistent dataflow launch
bool gpb_potrf_dataflow_ok(long long n, int batch);
int gpb_launch_potrf_dataflow(double* A, long long n, long long ld, double* W, long long ldw, double* V,
                              long long ldv, int* info, cudaStream_t st, long long n_valid);
int gpb_launch_trtri(const double* L, long long n, long long ld, long long sL, int batch, double* W,
                     long long ldw, long long sW, double* V, long long ldv, long long sV, double* T,
                     long long ldt, long long sT, cudaStream_t st);
int gpb_launch_lauum(const double* V, long long n, long long ldv, long long sV, int batch, double* Ki,
                     long long ldk, long long sK, cudaStream_t st);
int gpb_launch_potrs(const double* L, const double* W, long long n, long long ld, long long ldw,
                     long long sL, long long sW, int batch, const double* y, long long sy, double* z,
                     double* alpha, long long svec, int* flags, cudaStream_t st);
int gpb_launch_loglh(const double* L, long long n_valid, long long ld, long long sL, int batch,
                     const double* y, long long sy, const double* alpha, long long svec,
                     const int* info, double* out3, cudaStream_t st);
int gpb_launch_tril(double* A, long long n, long long ld, long long sA, int batch, cudaStream_t st);
int gpb_launch_tril_copy(double* dst, long long ldd, const double* src, long long lds, long long n, cudaStream_t st);
int gpb_launch_copy2d(double* dst, long long ldd, const double* src, long long lds, long long rows,
                      long long cols, long long sD, long long sS, int batch, cudaStream_t st);

// small.cu: everything after the factorisation of a one-block GP (n <= 128) in one launch
int gpb_launch_small_tail(int kind, const KParams* P, const KParams* Pb, int batch, const double* x, long long n,
                          const double* y, long long sy, const double* L, long long ldl, long long sL,
                          const double* W, long long ldw, long long sW, double* Ki, long long ldk, long long sK,
                          double* z, double* alpha, long long svec, const int* info, double* out3,
                          double* out16, double* Wz, double* Vz, long long ldv, long long sV, double* pack,
                          cudaStream_t st);
// Wz / Vz (optional): W and V again, writable -- the kernel then writes their structural zeros (the
// strictly upper / lower 32x32 sub-blocks) instead of a zero_diag_blocks launch.  pack (optional):
// the 24-double read-back block of gpb_gp_stages for batch 1.

// cov(xo) of a one-block GP (n <= 128) at m <= 128 test points in one launch, when
// gpb_small_cov_smem(m, n) <= 200 KB
size_t gpb_small_cov_smem(long long m, long long n);
int gpb_launch_small_cov(int kind, const KParams* P, const double* xo, long long m, const double* x, long long n,
                         const double* W, long long ldw, double* out, long long ldo, cudaStream_t st);

// reduce.cu
int gpb_launch_gemv(const double* A, long long rows, long long cols, long long lda, const double* x,
                    double* y, double alpha, double beta, cudaStream_t st);
int gpb_launch_trace_prod(const double* A, long long lda, const double* B, long long ldb, long long n,
                          double* partial, double* out, cudaStream_t st);
int gpb_launch_quadform(const double* u, const double* M, long long ldm, const double* v, long long n,
                        double* partial, double* out, cudaStream_t st);
int gpb_reduce_blocks(long long n);
