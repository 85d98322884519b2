/*
 * gpb200.h -- C ABI of libgpb200.so: the B200 (sm_100a) GP-regression hot path.
 *
 * Drop-in boundary for jhamrick/gaussian_processes v1.0.5.  The reference's native
 * layer is three Cython modules (gp/ext/gaussian_c.pyx, periodic_c.pyx, gp_c.pyx)
 * plus the LAPACK calls gp/gp.py makes through scipy/numpy; every entry point
 * below names the reference interface it replaces (file:line under the reference
 * tree).  Plain pointers and sizes only; no torch / numpy types.
 *
 * Conventions
 *   - all matrices are row-major (C order) fp64, like the reference's numpy arrays;
 *   - "*_host" / "gpb_gaussian_*" / "gpb_periodic_*" entry points take HOST pointers and
 *     do their own host<->device copies (what a ctypes/cffi binding of the
 *     reference's Cython signatures would call);
 *   - all other entry points take DEVICE pointers and a cudaStream_t (as void*),
 *     launch asynchronously and never retain pointers;
 *   - device matrices used by the factorisation are padded to a multiple of 128
 *     (GPB_NB) with an identity pad, see DESIGN.md "Data layout";
 *   - return value: 0 ok, <0 error (gpb_last_error() has the text).  Cholesky
 *     failure is NOT an error code: it is reported through the device `info` word
 *     (LAPACK convention: index+1 of the first non-positive pivot), which the Python
 *     layer turns into numpy.linalg.LinAlgError exactly where gp/gp.py:294 raises.
 *   - threading: entry points that use the library's process-global staging state (every
 *     *_host entry point, gpb_gp_eval, gpb_gp_stages, gpb_download_2d, options, profiling)
 *     serialise on one internal lock, so concurrent host threads are safe (and do not
 *     overlap inside those calls).  The stream-only entry points (gpb_kernel_build,
 *     gpb_potrf, gpb_gemm_nt, ...) take no lock; use one host thread per stream for them.
 */
#ifndef GPB200_H
#define GPB200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define GPB_KIND_GAUSSIAN 0   /* gp/kernels/gaussian.py:14, params (h, w)    */
#define GPB_KIND_PERIODIC 1   /* gp/kernels/periodic.py:14, params (h, w, p) */
#define GPB_BLOCK 128         /* padding / blocking unit of the device matrices */

/* ---- library ------------------------------------------------------------ */
int gpb_version(void);
const char* gpb_last_error(void);
/* MIN = log(exp2(minexp+4)): gaussian_c.pyx:15, gp_c.pyx:14, gp.py:17 */
double gpb_min_log(void);

/* ---- kernel-matrix builders, device pointers -------------------------------
 * Slice ids (the order jacobian()/hessian() stack them in the reference):
 *   Gaussian: 0 K | 1 dK_dh 2 dK_dw | 3 hh 4 hw 5 wh 6 ww
 *   Periodic: 0 K | 1 dh 2 dw 3 dp | 4 hh 5 hw 6 hp 7 wh 8 ww 9 wp 10 ph 11 pw 12 pp
 * Every slice selected in `slice_mask` is written, in slice-id order, to
 * out + q * slice_stride (q = rank among the selected slices) with row stride ld.
 * rows/cols >= n1/n2: the pad region is 0, or the identity for slice 0 when
 * pad_identity.  add_diag adds s^2 on the index diagonal of slice 0 (gp.py:265).
 * theta = (h, w[, p]) on the HOST.
 * Replaces gaussian_c.K, jacobian, hessian, dK_dX, d2K_dXdY (gaussian_c.pyx:18-164) and
 * periodic_c.* (periodic_c.pyx:18-235) in one fused pass.                        */
int gpb_kernel_build(int kind, const double* theta, double s, const double* x1, int64_t n1,
                     const double* x2, int64_t n2, int64_t rows, int64_t cols,
                     unsigned slice_mask, double* out, int64_t ld, int64_t slice_stride,
                     int add_diag, int pad_identity, void* stream);

/* out[g][r] = sum_p coef[p] * sum_c slice[p](x1[r], x2[c]) * vec[p][c]   (never
 * materialises the n1 x n2 kernel matrix).  Serves GP.mean (gp.py:597), the
 * mat-vecs of gp_c.dm_dtheta (gp_c.pyx:122-131) and dK_i * alpha.  npairs <= 8, nout <= 4. */
int gpb_kernel_matvec(int kind, const double* theta, const double* x1, int64_t n1,
                      const double* x2, int64_t n2, int npairs, const int* slice,
                      const int* outidx, const double* coef, const double* const* vec,
                      int nout, double* const* out, void* stream);

/* var[r] = k(x*_r, x*_r) - sum_c Z[r][c]^2 for Z = K(xo, x) L^-T (m x n, row stride ldz): the
 * diagonal of GP.cov (gp.py:625) -- all GP.plot needs (gp.py:692-693) -- without the m x m matrix. */
int gpb_post_var(int kind, const double* theta, const double* Z, int64_t ldz, int64_t m, int64_t n,
                 double* out, void* stream);

/* ---- factorisation and solves, device pointers (n multiple of 128) ---------- */
/* In-place lower Cholesky of the batch of n x n matrices A.  Only the lower triangle is read
 * and written at 32-column granularity: the strict upper part of the 32x32 diagonal
 * sub-blocks is zeroed, everything else above the diagonal is left untouched -- see
 * gpb_tril for the dense lower-triangular matrix scipy returns.  Also
 * writes the inverted 128x128 diagonal blocks into W (lower) and, if V != NULL, their
 * transposes into V.  The tiles of W strictly below its diagonal blocks are WORKSPACE of the call
 * (a single matrix factored by the dataflow launch keeps L tiles there until it copies them home);
 * gpb_trtri overwrites every one of them.  info[b] = 0 or first failing column + 1.
 * Replaces scipy.linalg.cholesky(lower=True) at gp/gp.py:294.                       */
int gpb_potrf(double* A, int64_t n, int64_t ld, int64_t stride_a, int batch, double* W,
              int64_t ldw, int64_t stride_w, double* V, int64_t ldv, int64_t stride_v,
              int* info, void* stream);
/* alpha = (L L^T)^-1 y by substitution; z = scratch [batch][n]; flags = int scratch
 * [2*batch*n/128 + 2].  Replaces scipy.linalg.cho_solve at gp/gp.py:332-334.          */
int gpb_potrs(const double* L, const double* W, int64_t n, int64_t ld, int64_t ldw,
              int64_t stride_l, int64_t stride_w, int batch, const double* y, int64_t stride_y,
              double* z, double* alpha, int64_t stride_vec, int* flags, void* stream);
/* W = L^-1 (lower), V = W^T (upper) completed from the diagonal blocks left by
 * gpb_potrf; T = n x n scratch.  Replaces np.linalg.inv(Lxx) at gp/gp.py:311.         */
int gpb_trtri(const double* L, int64_t n, int64_t ld, int64_t stride_l, int batch, double* W,
              int64_t ldw, int64_t stride_w, double* V, int64_t ldv, int64_t stride_v,
              double* T, int64_t ldt, int64_t stride_t, void* stream);
/* Ki = V V^T = inv(L)^T inv(L), full symmetric.  Replaces np.dot at gp/gp.py:312.     */
int gpb_lauum(const double* V, int64_t n, int64_t ldv, int64_t stride_v, int batch, double* Ki,
              int64_t ldk, int64_t stride_k, void* stream);
/* zero the strict upper triangle (what scipy returns for Lxx).                          */
int gpb_tril(double* A, int64_t n, int64_t ld, void* stream);
/* dst = tril(src) without reading the strict upper triangle of src (which gpb_potrf never
 * writes): how GP.Lxx (gp/gp.py:278-294) is extracted from the factorisation buffer.          */
int gpb_tril_copy(double* dst, int64_t ldd, const double* src, int64_t lds, int64_t n, void* stream);
int gpb_copy2d(double* dst, int64_t ldd, const double* src, int64_t lds, int64_t rows,
               int64_t cols, void* stream);
/* rows x cols block, device -> HOST.  dst_pinned != 0: the destination is page-locked, one
 * asynchronous strided DMA is enqueued on `stream` (the caller synchronises).  Otherwise the
 * block is pipelined through pinned staging slots (DMA of chunk k+1 overlaps the host copy
 * of chunk k) and the call returns when dst_host is complete.  This is how matrix-valued
 * results (GP.cov gp.py:625, Kxx, Lxx, inv_Kxx, kernel slices) reach numpy arrays.          */
int gpb_download_2d(double* dst_host, int64_t ld_host, const double* src_dev, int64_t ld_dev,
                    int64_t rows, int64_t cols, int dst_pinned, void* stream);

/* C = beta*C + alpha * A * B^T on FP64 tensor cores (DMMA); M, N multiples of 128, K of 16.
 * a_tri/b_tri: 0 dense, 1 lower (A[i][k] = 0 for k > i), 2 upper (k < i): zero tiles
 * are skipped.  lower_only: only tiles on/below the diagonal, mirrored into Ct.
 * The products of GP.cov (gp.py:625) and gp_c.d2lh_dtheta2 (gp_c.pyx:81-111) run here. */
int gpb_gemm_nt(const double* A, int64_t lda, const double* B, int64_t ldb, double* C,
                int64_t ldc, double* Ct, int64_t ldct, int64_t M, int64_t N, int64_t K,
                double alpha, double beta, int a_tri, int b_tri, int lower_only, void* stream);

/* ---- reductions, device pointers --------------------------------------------- */
/* out3 = {log_lh, logdet, y.alpha}; log_lh = -inf if info != 0 or logdet < MIN.
 * logdet = 2 sum log L_ii.  Replaces gp_c.log_lh (gp_c.pyx:17-31).                     */
int gpb_loglh(const double* L, int64_t n_valid, int64_t ld, const double* y,
              const double* alpha, const int* info, double* out3, void* stream);
/* For each requested slice S_q (q < nslices <= 6; any jacobian/hessian slice id):
 *   out16[q] = a^T S_q a,  out16[6+q] = sum(Ki o S_q);  out16[12] = tr Ki, out16[13] = a.a.
 * S_q is regenerated from x on the fly (one pass over Ki is the only HBM traffic).
 * partial = scratch [gpb_grad_partial_doubles(n)].  The O(N^2) core of
 * gp_c.dloglh_dtheta / dlh_dtheta (gp_c.pyx:34-67) and of the trace and quadratic-form
 * terms of gp_c.d2lh_dtheta2 (gp_c.pyx:104-110).                                        */
int gpb_slice_reduce(int kind, const double* theta, const double* x, int64_t n, const double* Ki,
                     int64_t ldk, const double* alpha, int nslices, const int* slices,
                     double* partial, double* out16, void* stream);
int64_t gpb_grad_partial_doubles(int64_t n);
/* y = alpha * A x + beta * y                                                          */
int gpb_gemv(const double* A, int64_t rows, int64_t cols, int64_t lda, const double* x, double* y,
             double alpha, double beta, void* stream);
/* *out = sum_ab A[a,b] B[b,a] = trace(A B)   (np.trace(np.dot(..)) in gp_c.pyx:49,110)  */
int gpb_trace_prod(const double* A, int64_t lda, const double* B, int64_t ldb, int64_t n,
                   double* partial, double* out, void* stream);
/* *out = u^T M v                                                                      */
int gpb_quadform(const double* u, const double* M, int64_t ldm, const double* v, int64_t n,
                 double* partial, double* out, void* stream);

/* ---- fused evaluator: log_lh + dloglh_dtheta for a batch of hyperparameter
 * candidates on fixed (x, y) -- the unit of the headline metric and of fit_MLII.
 * thetas: HOST [batch][n_theta] rows (kernel params..., s).  x, y: DEVICE [n].
 * result: DEVICE [batch][8] = {log_lh, dloglh[0..3] (unused = 0), logdet, y.alpha, info}.
 * dloglh is NaN and log_lh -inf for a candidate whose Cholesky fails (gp.py:362-365,
 * 424-428).  want_grad = 0 skips inverse + gradient.
 * Replaces, per candidate, GP.Kxx/Lxx/inv_Kxx_y/inv_Kxx/log_lh/dloglh_dtheta
 * (gp/gp.py:242-433).                                                                 */
size_t gpb_eval_workspace_bytes(int64_t n, int batch, int want_grad);
int gpb_gp_eval(int kind, const double* thetas, int batch, const double* x, const double* y,
                int64_t n, int want_grad, void* workspace, size_t workspace_bytes,
                double* result, void* stream);
/* same with HOST x, y, result: uploads, evaluates, downloads, synchronises.            */
int gpb_gp_eval_host(int kind, const double* thetas, int batch, const double* x, const double* y,
                     int64_t n, int want_grad, double* result);

/* ---- one GP object, resident buffers, staged: the property chain of gp/gp.py:242-433 for
 * ONE theta, enqueued by one call.  The workspace is the batch-1 evaluator workspace; its
 * layout (byte offsets of L, W, V, Ki, z, alpha, ypad, partial, out3, out16, params, info,
 * flags, pack, then the total size; noff >= 15) comes from gpb_eval_layout so the caller can
 * keep using L / W / V / Ki / alpha afterwards (posterior, second derivatives).
 * stages (bit mask, run in this order):
 *   1  Kxx + s^2 I -> Cholesky (gp.py:265,294), alpha (gp.py:332-334), log_lh (gp_c.pyx:17-31)
 *   2  W = L^-1, V = L^-T                      4  Ki = V V^T (gp.py:311-312)
 *   8  gradient brackets (gp_c.pyx:41-49): out16 as in gpb_slice_reduce for the Jacobian slices
 * A GP of n <= 128 (one block) completes all four stages whenever stage 1 is asked for.
 * theta: HOST [n_theta] (kernel params..., s).  x: DEVICE [n].  ypad: DEVICE [roundup(n,128)],
 * zero beyond n.  host_out: HOST [24] or NULL = {log_lh, logdet, y.alpha | out16[16] | info |
 * mask of the stages this call completed | pad};
 * when given, the call synchronises the stream (the ONE blocking read-back of the chain).     */
int gpb_eval_layout(int64_t n, int64_t* off, int noff);
int gpb_gp_stages(int kind, const double* theta, const double* x, const double* ypad, int64_t n,
                  unsigned stages, void* workspace, size_t workspace_bytes, double* host_out,
                  void* stream);
/* Posterior mean K(xo, x) alpha (gp.py:574-597) with HOST xo / out: upload, fused
 * kernel-times-vector (K(xo, x) is never materialised), download, synchronise.
 * scratch: DEVICE, >= 2 * roundup(m, 32) doubles; may be NULL for m <= 2048 (the kernel then reads
 * xo and writes the result in page-locked host memory directly, no copy engine involved).      */
int gpb_post_mean_host(int kind, const double* theta, const double* xo_host, int64_t m,
                       const double* x, int64_t n, const double* alpha, double* scratch,
                       double* out_host, void* stream);
/* Posterior covariance (gp.py:599-625) with HOST xo / out for small test sets:
 * K(xo,xo) - Z Z^T, Z = K(xo,x) L^-T, W = L^-1 [roundup(n,128)]^2 (stage 2 above).
 * scratch: DEVICE, 256-byte aligned, >= gpb_post_cov_scratch_doubles(m, n) doubles; that is 0
 * (scratch may be NULL) for a GP of n <= 128 at m <= 128 test points, which is one launch.      */
size_t gpb_post_cov_scratch_doubles(int64_t m, int64_t n);
int gpb_post_cov_host(int kind, const double* theta, const double* xo_host, int64_t m,
                      const double* x, int64_t n, const double* W, int64_t ldw, double* scratch,
                      double* out_host, int64_t ld_out, void* stream);

/* ---- host-buffer drop-ins for the reference's Cython signatures ----------------
 * f(out, x1, x2, h, w[, p]) with caller-allocated C-contiguous out, exactly the
 * arguments of gaussian_c.pyx:18,39,44,51,72,95,116,139,143 and
 * periodic_c.pyx:18,33,39,53,68,83,99,...,223 (sizes made explicit).                   */
int gpb_kernel_slices_host(int kind, unsigned slice_mask, double* out, const double* x1,
                           int64_t n1, const double* x2, int64_t n2, const double* theta);
int gpb_gaussian_K(double* out, const double* x1, int64_t n1, const double* x2, int64_t n2, double h, double w);
int gpb_gaussian_jacobian(double* out, const double* x1, int64_t n1, const double* x2, int64_t n2, double h, double w);
int gpb_gaussian_hessian(double* out, const double* x1, int64_t n1, const double* x2, int64_t n2, double h, double w);
int gpb_gaussian_dK_dh(double* out, const double* x1, int64_t n1, const double* x2, int64_t n2, double h, double w);
int gpb_gaussian_dK_dw(double* out, const double* x1, int64_t n1, const double* x2, int64_t n2, double h, double w);
int gpb_gaussian_d2K_dhdh(double* out, const double* x1, int64_t n1, const double* x2, int64_t n2, double h, double w);
int gpb_gaussian_d2K_dhdw(double* out, const double* x1, int64_t n1, const double* x2, int64_t n2, double h, double w);
int gpb_gaussian_d2K_dwdh(double* out, const double* x1, int64_t n1, const double* x2, int64_t n2, double h, double w);
int gpb_gaussian_d2K_dwdw(double* out, const double* x1, int64_t n1, const double* x2, int64_t n2, double h, double w);
int gpb_periodic_K(double* out, const double* x1, int64_t n1, const double* x2, int64_t n2, double h, double w, double p);
int gpb_periodic_jacobian(double* out, const double* x1, int64_t n1, const double* x2, int64_t n2, double h, double w, double p);
int gpb_periodic_hessian(double* out, const double* x1, int64_t n1, const double* x2, int64_t n2, double h, double w, double p);
int gpb_periodic_dK_dh(double* out, const double* x1, int64_t n1, const double* x2, int64_t n2, double h, double w, double p);
int gpb_periodic_dK_dw(double* out, const double* x1, int64_t n1, const double* x2, int64_t n2, double h, double w, double p);
int gpb_periodic_dK_dp(double* out, const double* x1, int64_t n1, const double* x2, int64_t n2, double h, double w, double p);
int gpb_periodic_d2K_dhdh(double* out, const double* x1, int64_t n1, const double* x2, int64_t n2, double h, double w, double p);
int gpb_periodic_d2K_dhdw(double* out, const double* x1, int64_t n1, const double* x2, int64_t n2, double h, double w, double p);
int gpb_periodic_d2K_dhdp(double* out, const double* x1, int64_t n1, const double* x2, int64_t n2, double h, double w, double p);
int gpb_periodic_d2K_dwdh(double* out, const double* x1, int64_t n1, const double* x2, int64_t n2, double h, double w, double p);
int gpb_periodic_d2K_dwdw(double* out, const double* x1, int64_t n1, const double* x2, int64_t n2, double h, double w, double p);
int gpb_periodic_d2K_dwdp(double* out, const double* x1, int64_t n1, const double* x2, int64_t n2, double h, double w, double p);
int gpb_periodic_d2K_dpdh(double* out, const double* x1, int64_t n1, const double* x2, int64_t n2, double h, double w, double p);
int gpb_periodic_d2K_dpdw(double* out, const double* x1, int64_t n1, const double* x2, int64_t n2, double h, double w, double p);
int gpb_periodic_d2K_dpdp(double* out, const double* x1, int64_t n1, const double* x2, int64_t n2, double h, double w, double p);

/* ---- gp/ext/gp_c.pyx as host-pointer entry points --------------------------------------
 * Exactly the reference's argument lists (C-contiguous float64 host arrays; the shapes the
 * Cython signatures imply are passed as n_p = kernel parameters, n = observations, m = test
 * points), outputs written in place, no Python / torch on the path:
 *   gp_c.log_lh(y, K, Kiy) -> float                                  gp_c.pyx:17-31
 *   gp_c.dloglh_dtheta(y, Ki, Kj, Kiy, s, dloglh)                    gp_c.pyx:34-49
 *   gp_c.dlh_dtheta(y, Ki, Kj, Kiy, s, lh, dlh)                      gp_c.pyx:52-67
 *   gp_c.d2lh_dtheta2(y, Ki, Kj, Kh, Kiy, s, lh, dlh, d2lh)          gp_c.pyx:70-111
 *   gp_c.dm_dtheta(y, Ki, Kj, Kjxo, Kxox, s, dm)                     gp_c.pyx:114-131
 * Kj: [n_p, n, n]; Kh: [n_p, n_p, n, n]; Kjxo: [n_p, m, n]; Kxox: [m, n]; dloglh, dlh: [n_p + 1];
 * d2lh: [n_p + 1, n_p + 1]; dm: [n_p + 1, m].  log|K| comes from a Cholesky of K (the reference
 * takes an LU slogdet, gp_c.pyx:21): a K that is not positive definite gives -inf.            */
int gpb_gp_c_log_lh(const double* y, const double* K, const double* Kiy, int64_t n, double* llh);
int gpb_gp_c_dloglh_dtheta(const double* y, const double* Ki, const double* Kj, const double* Kiy, double s,
                           int64_t n_p, int64_t n, double* dloglh);
int gpb_gp_c_dlh_dtheta(const double* y, const double* Ki, const double* Kj, const double* Kiy, double s, double lh,
                        int64_t n_p, int64_t n, double* dlh);
int gpb_gp_c_d2lh_dtheta2(const double* y, const double* Ki, const double* Kj, const double* Kh, const double* Kiy,
                          double s, double lh, const double* dlh, int64_t n_p, int64_t n, double* d2lh);
int gpb_gp_c_dm_dtheta(const double* y, const double* Ki, const double* Kj, const double* Kjxo, const double* Kxox,
                       double s, int64_t n_p, int64_t n, int64_t m, double* dm);

/* ---- measurement helpers ----------------------------------------------------------
 * FP64 tensor (DMMA.8x8x4) and FP64 FMA issue-rate microbenchmarks: the roofline
 * denominator for the factorisation (MEASURED_PEAKS.json has no fp64 figure).          */
int gpb_microbench_fp64(int use_dmma, int iters, double* tflops, double* ms);
/* Dependent-issue latencies in SM cycles, one warp: out7 = {DFMA chain, rsqrt(+DADD) chain, 64-bit
 * shuffle chain, dependent DMMA chain, DMMA issue interval with 8 independent accumulators,
 * 1/sqrt(+DADD) chain, shared-memory store->load round trip}.  These bound the serial paths
 * of the diagonal-block factorisation.                                         */
int gpb_microbench_latency(double* out7);
/* number of kernel launches issued by this library since load (bench.py gpu_launches) */
int64_t gpb_launch_count(void);
/* Tuning knobs (0 = built-in default; also readable from the environment):
 *   "eval_streams" (GPB_EVAL_STREAMS) candidate groups evaluated on concurrent streams by gpb_gp_eval
 *   "gemm_bm"      (GPB_GEMM_BM)      64 (two CTAs per SM) or 128 row tiles in the DMMA GEMM
 *   "potrf_inner"  (GPB_POTRF_INNER)  128-columns per outer Cholesky panel
 *   "potrf_lookahead" (GPB_POTRF_LOOKAHEAD) panel look-ahead of a single factorisation: 0 auto (N >= 6144), 1 on, 2 off
 *   "gemm_impl"    (GPB_GEMM_IMPL)    0 TMA + mbarrier operand pipeline, 1 cp.async pipeline
 *   "potrf_dataflow" (GPB_POTRF_DATAFLOW) one matrix by the persistent dataflow launch: 0 auto (256 <= N <= 8192), 1 whenever batch == 1, 2 never
 *   "chain_group"  (GPB_CHAIN_GROUP)  0 = pipelined chain group (sweeping CTA + 8 helpers + inverter), 8 or 4 = the first
 *                                     chain group of that many CTAs (TRSM + SYRK of the chain tiles after the diagonal block)
 *   "chain_diag"   (GPB_CHAIN_DIAG)   2 = the chain CTA factors diagonal blocks with the 256-thread body
 *   "chain_sched"  (GPB_CHAIN_SCHED)  workers: 0 = most urgent runnable half tile first, 1 = in-order task lists
 *   "chain_mform"  (GPB_CHAIN_MFORM)  M form of a tile's last worker update(s): 0/1 = last step, 3 / 4 = more, 2 = off
 *   "chain_fuse"   (GPB_CHAIN_FUSE)   most backlog steps of a half tile applied by one worker task (default 8)
 *   "chain_horizon" (GPB_CHAIN_HORIZON) updates of tiles not needed for this many steps yield while their group has a tile
 *                                     due within "chain_imminent" steps (defaults 3 and 1; 100 = off)
 *   "chain_fuse_guard", "chain_express"  scheduling experiments (csrc/chain.cu)
 *   "lauum_fuse"   (GPB_LAUUM_FUSE)   1 = gradient brackets in the epilogue of the K^-1 = V V^T GEMM (measured slower: off)
 *   "stage_overlap" (GPB_STAGE_OVERLAP) 2 = gpb_gp_stages keeps the triangular solves on the caller's stream
 * None of them changes a result beyond rounding; the dataflow variants are compared in tests/test_parity_gpu_r2.py. */
int gpb_set_option(const char* name, int value);
/* Per-kernel-class device timing: while enabled, each launch group of a class is bracketed
 * by CUDA events on its stream.  Classes: 0 DMMA GEMM, 1 diagonal-block factor, 2 kernel
 * builders / fused mat-vecs, 3 triangular solves, 4 reductions, 5 misc.                 */
void gpb_profile_enable(int on);
int gpb_profile_read(int kernel_class, double* ms, int64_t* launches);

#ifdef __cplusplus
}
#endif
#endif /* GPB200_H */
