// potrf.cu -- blocked Cholesky of Kxx + s^2 I, the triangular inverse / K^-1 built
// from it, the Cholesky solves and the log-likelihood reduction.
//
// Replaces the LAPACK calls the reference makes through scipy/numpy:
//   scipy.linalg.cholesky (gp/gp.py:294), cho_solve (gp.py:332-334),
//   np.linalg.inv + np.dot (gp.py:311-312), np.linalg.slogdet (gp_c.pyx:21).
//
// Structure (row-major, lower):  right-looking blocked algorithm with NB = 128.
//   per block step k:  potrf_diag_kernel   L_kk = chol(A_kk),  W_kk = L_kk^-1, V_kk = W_kk^T
//                      GEMM (gemm.cu)      A_ik <- A_ik W_kk^T            (TRSM as a DMMA product)
//                      GEMM                A_ij -= A_ik A_jk^T, i>=j>k    (SYRK, lower tiles)
//   trtri: pairwise merges  W21 = -W22 (L21 W11), each level one batched GEMM pair
//   lauum: Ki = V V^T (V = L^-T), one GEMM launch, lower tiles + mirrored store
// Every kernel carries a batch dimension (independent GPs / MLII candidates).
#include <vector>
#include "common.cuh"
#include "launch.h"
#include "diag_block.cuh"

namespace {

__global__ void __launch_bounds__(256, DIAG_MIN_CTAS)
potrf_diag_kernel(double* A, long long ld, long long sA, double* W, long long ldw, long long sW,
                  double* V, long long ldv, long long sV, int* info, int col0, int nsub) {
    extern __shared__ __align__(16) double sm[];
    diag_block_body<false>(A + (long long)blockIdx.x * sA, ld, W + (long long)blockIdx.x * sW, ldw,
                           V ? V + (long long)blockIdx.x * sV : nullptr, ldv, info + blockIdx.x, col0, nsub, sm);
}

// ===========================================================================

// Cholesky solves with one right-hand side: dataflow over 128-row blocks.
// CTA i accumulates  sum_j L_ij z_j  as soon as z_j is published (release/acquire
// flags), then applies the inverted diagonal block.  Tickets make the start order
// match the dependency order, so spinning CTAs can never starve their producers.
//
// The serial chain (block i needs z_{i-1}) is what bounds a single solve, so nothing on
// it may wait for HBM: the L blocks stream through a cp.async ring of 128 x 32 (forward)
// or 32 x 128 (backward) chunks that runs AHEAD of the flag waits -- L does not depend
// on the flags, only the multiply does -- and the inverted diagonal block W_ii sits in
// registers (64 doubles per thread) from the start of the CTA.  When z_{i-1} lands, block
// L(i,i-1) is already in shared memory: the step costs one flag round trip, two 128 x 128
// mat-vecs from on-chip data and one publish.
// ===========================================================================
constexpr int TS = GPB_NB;
constexpr int TR_RING = 4;                    // chunks resident per CTA
constexpr int TR_FLD = 34;                    // forward chunk row stride (doubles): 16-byte rows, conflict-free 128-bit reads
constexpr int TR_FCH = TS * TR_FLD;           // forward chunk: 128 rows x 32 columns (padded)
constexpr int TR_BCH = 32 * TS;               // backward chunk: 32 rows x 128 columns
constexpr int TRSV_FWD_SMEM = (TR_RING * TR_FCH + 4 * TS) * 8;
constexpr int TRSV_BWD_SMEM = (TR_RING * TR_BCH + 4 * TS) * 8;

__global__ void __launch_bounds__(256, 1)
trsv_fwd_kernel(const double* L, long long ld, long long sL, const double* W, long long ldw, long long sW,
                const double* y, long long sy, double* z, long long svec, int T, int* flags, int* counter) {
    extern __shared__ __align__(16) double tsm[];
    double* ring = tsm;                        // TR_RING chunks
    double* zs = tsm + TR_RING * TR_FCH;       // current z_j (128)
    double* red = zs + TS;                     // [2][128] partial sums of the two column halves
    double* rhs = red + 2 * TS;                // 128
    __shared__ int s_ticket;
    const int tid = threadIdx.x;
    if (tid == 0) s_ticket = atomicAdd(counter, 1);
    __syncthreads();
    const int b = s_ticket / T, i = s_ticket % T;
    L += b * sL; W += b * sW; y += b * sy; z += b * svec; flags += (long long)b * T;
    const int row = tid & (TS - 1), half = tid >> 7;

    // chunk n = (block j = n / 4, columns j*128 + (n % 4)*32 ..): 8 x 16-byte pieces per thread
    const int nchunks = 4 * i;
    const double* Lrow0 = L + (long long)i * TS * ld;
    auto issue = [&](int n) {
        if (n < nchunks) {
            const double* src = Lrow0 + (long long)(n >> 2) * TS + (n & 3) * 32;
            double* dst = ring + (n % TR_RING) * TR_FCH;
#pragma unroll
            for (int q = 0; q < 8; q++) {
                const int p = tid + q * 256, r = p >> 4, pc = p & 15;
                cp_async16(dst + r * TR_FLD + pc * 2, src + (long long)r * ld + pc * 2);
            }
        }
        cp_async_commit();
    };
#pragma unroll
    for (int n = 0; n < TR_RING - 1; n++) issue(n);

    // W_ii[row][half*64 .. +64) -> registers (independent of every flag)
    double wreg[64];
    {
        const double* Wp = W + ((long long)i * TS + row) * ldw + (long long)i * TS + half * 64;
#pragma unroll
        for (int k = 0; k < 64; k += 2) {
            const double2 v = *reinterpret_cast<const double2*>(Wp + k);
            wreg[k] = v.x; wreg[k + 1] = v.y;
        }
    }

    double acc = 0.0;
    for (int n = 0; n < nchunks; n++) {
        cp_async_wait<TR_RING - 2>();
        if ((n & 3) == 0) {                    // new block j: its z must be published
            const int j = n >> 2;
            if (tid == 0) while (ld_acquire(flags + j) == 0) {}
            __syncthreads();
            if (tid < TS) zs[tid] = __ldcg(z + (long long)j * TS + tid);
        }
        __syncthreads();                       // chunk n landed for all threads; zs visible; slot (n-1) free
        issue(n + TR_RING - 1);
        const double* ch = ring + (n % TR_RING) * TR_FCH + row * TR_FLD + half * 16;
        const double* zp = zs + (n & 3) * 32 + half * 16;
#pragma unroll
        for (int k = 0; k < 16; k += 2) {
            const double2 v = *reinterpret_cast<const double2*>(ch + k);
            acc = fma(v.x, zp[k], acc);
            acc = fma(v.y, zp[k + 1], acc);
        }
    }
    cp_async_wait<0>();
    red[half * TS + row] = acc;
    __syncthreads();
    if (tid < TS) rhs[tid] = y[(long long)i * TS + tid] - (red[tid] + red[TS + tid]);
    __syncthreads();
    {   // z_i = W_ii * rhs, W_ii from registers
        double s0 = 0.0, s1 = 0.0;
        const double* rp = rhs + half * 64;
#pragma unroll
        for (int k = 0; k < 64; k += 2) {
            s0 = fma(wreg[k], rp[k], s0);
            s1 = fma(wreg[k + 1], rp[k + 1], s1);
        }
        red[half * TS + row] = s0 + s1;
    }
    __syncthreads();
    if (tid < TS) z[(long long)i * TS + tid] = red[tid] + red[TS + tid];
    __threadfence();
    __syncthreads();
    if (tid == 0) st_release(flags + i, 1);
}

__global__ void __launch_bounds__(256, 1)
trsv_bwd_kernel(const double* L, long long ld, long long sL, const double* W, long long ldw, long long sW,
                const double* z, double* alpha, long long svec, int T, int* flags, int* counter) {
    extern __shared__ __align__(16) double tsm[];
    double* ring = tsm;
    double* as = tsm + TR_RING * TR_BCH;       // current alpha_j (128)
    double* red = as + TS;                     // [2][128] partial sums of the two row halves
    double* rhs = red + 2 * TS;
    __shared__ int s_ticket;
    const int tid = threadIdx.x;
    if (tid == 0) s_ticket = atomicAdd(counter, 1);
    __syncthreads();
    const int b = s_ticket / T, i = T - 1 - (s_ticket % T);
    L += b * sL; W += b * sW; z += b * svec; alpha += b * svec; flags += (long long)b * T;
    const int col = tid & (TS - 1), rhalf = tid >> 7;

    // chunk n = (block j = T-1 - n/4, rows j*128 + (n % 4)*32 .. of column block i)
    const int nchunks = 4 * (T - 1 - i);
    auto issue = [&](int n) {
        if (n < nchunks) {
            const int j = T - 1 - (n >> 2);
            const double* src = L + ((long long)j * TS + (n & 3) * 32) * ld + (long long)i * TS;
            double* dst = ring + (n % TR_RING) * TR_BCH;
#pragma unroll
            for (int q = 0; q < 8; q++) {
                const int p = tid + q * 256, r = p >> 6, pc = p & 63;
                cp_async16(dst + r * TS + pc * 2, src + (long long)r * ld + pc * 2);
            }
        }
        cp_async_commit();
    };
#pragma unroll
    for (int n = 0; n < TR_RING - 1; n++) issue(n);

    // W_ii[rhalf*64 .. +64)[col] -> registers: alpha_i = W_ii^T rhs
    double wreg[64];
    {
        const double* Wp = W + ((long long)i * TS + rhalf * 64) * ldw + (long long)i * TS + col;
#pragma unroll
        for (int k = 0; k < 64; k++) wreg[k] = Wp[(long long)k * ldw];
    }

    double acc = 0.0;
    for (int n = 0; n < nchunks; n++) {
        cp_async_wait<TR_RING - 2>();
        if ((n & 3) == 0) {
            const int j = T - 1 - (n >> 2);
            if (tid == 0) while (ld_acquire(flags + j) == 0) {}
            __syncthreads();
            if (tid < TS) as[tid] = __ldcg(alpha + (long long)j * TS + tid);
        }
        __syncthreads();
        issue(n + TR_RING - 1);
        const double* ch = ring + (n % TR_RING) * TR_BCH + (rhalf * 16) * TS + col;
        const double* ap = as + (n & 3) * 32 + rhalf * 16;
#pragma unroll
        for (int k = 0; k < 16; k++) acc = fma(ch[k * TS], ap[k], acc);
    }
    cp_async_wait<0>();
    red[rhalf * TS + col] = acc;
    __syncthreads();
    if (tid < TS) rhs[tid] = z[(long long)i * TS + tid] - (red[tid] + red[TS + tid]);
    __syncthreads();
    {
        double s0 = 0.0, s1 = 0.0;
        const double* rp = rhs + rhalf * 64;
#pragma unroll
        for (int k = 0; k < 64; k += 2) {
            s0 = fma(wreg[k], rp[k], s0);
            s1 = fma(wreg[k + 1], rp[k + 1], s1);
        }
        red[rhalf * TS + col] = s0 + s1;
    }
    __syncthreads();
    if (tid < TS) alpha[(long long)i * TS + tid] = red[tid] + red[TS + tid];
    __threadfence();
    __syncthreads();
    if (tid == 0) st_release(flags + i, 1);
}

// log_lh = -1/2 y.alpha - 1/2 logdet - n/2 log 2pi  with the reference's clamps
// (gp_c.pyx:21-29: -inf when logdet < MIN; gp.py:362-365: -inf when not PD).
// logdet comes from the Cholesky diagonal -- no second (LU) factorisation.
__global__ void __launch_bounds__(256)
loglh_kernel(const double* L, long long n, long long ld, long long sL, const double* y, long long sy,
             const double* alpha, long long svec, const int* info, double* out3) {
    __shared__ double red[32];
    const int b = blockIdx.x;
    L += b * sL; y += b * sy; alpha += b * svec;
    double sl = 0.0, sq = 0.0;
    for (long long i = threadIdx.x; i < n; i += blockDim.x) {
        sl += log(L[i * ld + i]);
        sq += y[i] * alpha[i];
    }
    sl = block_sum(sl, red);
    sq = block_sum(sq, red);
    if (threadIdx.x == 0) {
        const double logdet = 2.0 * sl;
        double llh;
        if ((info && info[b] != 0) || logdet < GPB_MIN_LOG) llh = -INFINITY;
        else llh = -0.5 * sq + -0.5 * logdet + -0.5 * (double)n * log(2.0 * M_PI);
        out3[b * 3 + 0] = llh;
        out3[b * 3 + 1] = logdet;
        out3[b * 3 + 2] = sq;
    }
}

// structural zeros of the inverted diagonal blocks: W_kk strictly-upper and V_kk strictly-lower
// 32x32 sub-blocks (one parallel launch per factorisation instead of stores on every
// diagonal-block kernel's serial path)
__global__ void __launch_bounds__(256) zero_diag_blocks_kernel(double* W, long long ldw, long long sW, double* V,
                                                               long long ldv, long long sV) {
    const long long o = (long long)blockIdx.x * GPB_NB;
    double* Wp = W + (long long)blockIdx.y * sW + o * ldw + o;
    double* Vp = V ? V + (long long)blockIdx.y * sV + o * ldv + o : nullptr;
    for (int e = threadIdx.x; e < GPB_NB * GPB_NB / 2; e += 256) {
        const int r = e >> 6, c2 = (e & 63) * 2;
        if ((c2 >> 5) > (r >> 5)) *reinterpret_cast<double2*>(Wp + (long long)r * ldw + c2) = make_double2(0.0, 0.0);
        if (Vp && (c2 >> 5) < (r >> 5)) *reinterpret_cast<double2*>(Vp + (long long)r * ldv + c2) = make_double2(0.0, 0.0);
    }
}

__global__ void tril_kernel(double* A, long long n, long long ld, long long sA) {
    A += (long long)blockIdx.z * sA;
    const long long c = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const long long r = (long long)blockIdx.y * blockDim.y + threadIdx.y;
    if (r < n && c < n && c > r) A[r * ld + c] = 0.0;
}

// dst = tril(src): the strict upper triangle of src is never read (the factorisation never writes it)
__global__ void tril_copy_kernel(double* dst, long long ldd, const double* src, long long lds, long long n) {
    const long long c = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    for (long long r = (long long)blockIdx.y * blockDim.y + threadIdx.y; r < n; r += (long long)gridDim.y * blockDim.y)
        if (c < n) dst[r * ldd + c] = (c <= r) ? src[r * lds + c] : 0.0;
}

__global__ void copy2d_kernel(double* dst, long long ldd, const double* src, long long lds,
                              long long rows, long long cols, long long sD, long long sS) {
    dst += (long long)blockIdx.z * sD;
    src += (long long)blockIdx.z * sS;
    const long long c = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    for (long long r = (long long)blockIdx.y * blockDim.y + threadIdx.y; r < rows;
         r += (long long)gridDim.y * blockDim.y)
        if (c < cols) dst[r * ldd + c] = src[r * lds + c];
}

}  // namespace

// ---------------------------------------------------------------------------
// drivers
// ---------------------------------------------------------------------------
#ifdef GPB_DIAG_CLK
extern "C" int gpb_debug_diag_clk(long long* out) {
    GPB_CUDA(cudaMemcpyFromSymbol(out, g_diag_clk, sizeof(long long) * 64));
    return GPB_OK;
}
#endif

int gpb_launch_potrf(double* A, long long n, long long ld, long long sA, int batch, double* W,
                     long long ldw, long long sW, double* V, long long ldv, long long sV, int* info,
                     cudaStream_t st, long long n_valid, bool zero_blocks, bool single_chain) {
    GPB_REQUIRE(n > 0 && n % GPB_NB == 0, "n must be a positive multiple of 128");
    GPB_REQUIRE(ld >= n && ldw >= n && (!V || ldv >= n), "leading dimension too small");
    if (n_valid <= 0 || n_valid > n) n_valid = n;
    GPB_REQUIRE(A && W && info, "null pointer");
    GPB_REQUIRE(ld % 2 == 0 && ldw % 2 == 0 && (!V || ldv % 2 == 0), "leading dimensions must be even");
    GPB_REQUIRE((reinterpret_cast<uintptr_t>(W) & 15) == 0 && (reinterpret_cast<uintptr_t>(V) & 15) == 0, "W and V must be 16-byte aligned");
    static bool attr_set = false;
    if (!attr_set) {
        GPB_CUDA(cudaFuncSetAttribute(potrf_diag_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, DIAG_SMEM));
        attr_set = true;
    }
    GPB_CUDA(cudaMemsetAsync(info, 0, sizeof(int) * batch, st));
    const int T = (int)(n / GPB_NB);
    if (zero_blocks) {
        zero_diag_blocks_kernel<<<dim3((unsigned)T, (unsigned)batch), 256, 0, st>>>(W, ldw, sW, V, ldv, sV);
        GPB_LAUNCH_CHECK("zero_diag_blocks_kernel");
    }
    // One matrix on its own, up to N = 8192: a single persistent dataflow launch (chain.cu) instead of the
    // launch sequence below.  (option potrf_dataflow: 0 = as described, 1 = also when the caller did not
    // ask for the single-matrix schedule, 2 = never)
    {
        const int df = gpb_get_option("potrf_dataflow");
        if (df != 2 && (single_chain || df == 1) && gpb_potrf_dataflow_ok(n, batch))
            return gpb_launch_potrf_dataflow(A, n, ld, W, ldw, V, ldv, info, st, n_valid);
    }
    // Two-level blocking: an outer panel of `inner` 128-columns is factored left-looking
    // (skinny column updates with K = q*128), then ONE trailing update with K = inner*128
    // touches the rest of the matrix -- half (or a quarter) as many passes over the trailing
    // matrix as a plain 128-wide right-looking sweep, each with a longer DMMA main loop.
    int inner = gpb_get_option("potrf_inner");
    if (inner < 1) inner = 4;

    // Look-ahead for a single matrix (a batch keeps the GPU busy by itself): the serial panel
    // factorisation runs on a high-priority side stream `ps`, and each trailing update is split in
    //   SYRK_a  the columns of the NEXT panel (on `ps`, the only part the next panel needs), and
    //   SYRK_b  the rest of the trailing matrix (on `st`, overlapping the next panel's serial work;
    //           the panel kernels' CTAs take freed SM slots first because of their stream priority).
    static thread_local cudaStream_t la_stream = nullptr;      // one per host thread (one thread per stream, gpb200.h)
    static thread_local cudaEvent_t la_ev[4];
    // Measured (tests/gpu_potrf_la.py): N = 8192 / 16384 / 32768 gain 11 / 11 / 4 % (34.2 TFLOP/s at
    // 32768); at N <= 4096 the trailing updates are too short to pay for the co-scheduling.
    const int la_opt = gpb_get_option("potrf_lookahead");
    const bool lookahead = (batch == 1) && (T > inner) && (la_opt != 2) && (T >= 48 || la_opt == 1);
    if (lookahead && !la_stream) {
        int lo = 0, hi = 0;
        GPB_CUDA(cudaDeviceGetStreamPriorityRange(&lo, &hi));
        GPB_CUDA(cudaStreamCreateWithPriority(&la_stream, cudaStreamNonBlocking, hi));
        for (auto& e : la_ev) GPB_CUDA(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
    }
    // One matrix on its own (single_chain: the GP-object chain and the public gpb_potrf with batch 1): the
    // panel steps are a serial chain of kernel latencies -> right-looking inside the panel.  The batched
    // evaluator never sets it, so a candidate's arithmetic does not depend on how the batch was grouped.
    // (option potrf_panel_rl: 0 = as the caller says, 1 = always, 2 = never)
    const int rl_opt = gpb_get_option("potrf_panel_rl");
    // Up to N = 8192 only: the K = 128 updates are HBM-bound (6 flop/B), and once the panel is tall the
    // three extra passes over it cost more than the shorter chain saves (N = 32768: 355 -> 397 ms).
    const bool panel_rl = (rl_opt == 1) || (rl_opt == 0 && single_chain && batch == 1 && T <= 64);
    cudaStream_t ps = lookahead ? la_stream : st;
    cudaEvent_t ev_start = la_ev[0], ev_panel = la_ev[1], ev_b = la_ev[2], ev_end = la_ev[3];
    bool have_b = false;
    if (lookahead) {
        GPB_CUDA(cudaEventRecord(ev_start, st));
        GPB_CUDA(cudaStreamWaitEvent(ps, ev_start, 0));
    }
    for (int k0 = 0; k0 < T; k0 += inner) {
        const int kw = (T - k0 < inner) ? (T - k0) : inner;
        const long long p0 = (long long)k0 * GPB_NB;
        for (int q = 0; q < kw; q++) {
            const int k = k0 + q;
            const long long o = (long long)k * GPB_NB;
            if (q > 0 && !panel_rl) {                  // A[k:, k] -= L[k:, k0..k) * L[k, k0..k)^T
                GpbGemm c = gpb_gemm_default();
                c.A = A + o * ld + p0; c.lda = ld; c.sA = sA;
                c.B = A + o * ld + p0; c.ldb = ld; c.sB = sA;
                c.C = A + o * ld + o; c.ldc = ld; c.sC = sA;
                c.M = (T - k) * GPB_NB; c.N = GPB_NB; c.K = q * GPB_NB;
                c.alpha = -1.0; c.beta = 1.0;
                int stt = gpb_launch_gemm(c, batch, ps);
                if (stt != GPB_OK) return stt;
            }
            {
                GpbProfScope prof(GPB_KC_DIAG, ps);
                // sub-blocks of this 128-block that hold observations (the rest is identity pad)
                const long long valid = n_valid - o;
                const int nsub = valid >= GPB_NB ? 4 : (valid <= 0 ? 0 : (int)((valid + SB - 1) / SB));
                potrf_diag_kernel<<<batch, 256, DIAG_SMEM, ps>>>(A + o * ld + o, ld, sA, W + o * ldw + o, ldw, sW,
                                                                 V ? V + o * ldv + o : nullptr, ldv, sV, info,
                                                                 (int)o, nsub);
                GPB_LAUNCH_CHECK("potrf_diag_kernel");
            }
            const int rem = (T - 1 - k) * GPB_NB;
            if (rem == 0) break;
            GpbGemm g = gpb_gemm_default();            // panel: A_ik <- A_ik * W_kk^T (in place)
            g.A = A + (o + GPB_NB) * ld + o; g.lda = ld; g.sA = sA;
            g.B = W + o * ldw + o; g.ldb = ldw; g.sB = sW;
            g.C = A + (o + GPB_NB) * ld + o; g.ldc = ld; g.sC = sA;
            g.M = rem; g.N = GPB_NB; g.K = GPB_NB;
            g.b_tri = 1;                               // W_kk lower: k <= j
            int stt = gpb_launch_gemm(g, batch, ps);
            if (stt != GPB_OK) return stt;
            const int ncols = (kw - 1 - q) * GPB_NB;
            if (panel_rl && ncols > 0) {
                // right-looking inside the panel: the new column updates the panel's remaining columns at
                // once (K = 128), so no step of the serial chain carries a K = q*128 product
                GpbGemm r = gpb_gemm_default();
                r.A = A + (o + GPB_NB) * ld + o; r.lda = ld; r.sA = sA;
                r.B = A + (o + GPB_NB) * ld + o; r.ldb = ld; r.sB = sA;
                r.C = A + (o + GPB_NB) * ld + (o + GPB_NB); r.ldc = ld; r.sC = sA;
                r.M = rem; r.N = ncols; r.K = GPB_NB;
                r.alpha = -1.0; r.beta = 1.0;
                stt = gpb_launch_gemm(r, batch, ps);
                if (stt != GPB_OK) return stt;
            }
        }
        const long long t0 = (long long)(k0 + kw) * GPB_NB;
        const int rem2 = (T - k0 - kw) * GPB_NB;
        if (rem2 <= 0) continue;
        GpbGemm u = gpb_gemm_default();                // trailing: A_ij -= P_i P_j^T (lower tiles)
        u.A = A + t0 * ld + p0; u.lda = ld; u.sA = sA;
        u.B = A + t0 * ld + p0; u.ldb = ld; u.sB = sA;
        u.C = A + t0 * ld + t0; u.ldc = ld; u.sC = sA;
        u.M = rem2; u.N = rem2; u.K = kw * GPB_NB;
        u.alpha = -1.0; u.beta = 1.0; u.lower_only = 1;
        if (!lookahead) {
            int stt = gpb_launch_gemm(u, batch, st);
            if (stt != GPB_OK) return stt;
            continue;
        }
        const int w = ((rem2 / GPB_NB < inner) ? rem2 / GPB_NB : inner) * GPB_NB;     // next panel's width
        // SYRK_b (rows/cols beyond the next panel) on st, as soon as this panel's P is complete
        if (rem2 > w) {
            GPB_CUDA(cudaEventRecord(ev_panel, ps));
            GPB_CUDA(cudaStreamWaitEvent(st, ev_panel, 0));
            GpbGemm ub = u;
            ub.A = u.A + (long long)w * ld; ub.B = ub.A;
            ub.C = u.C + (long long)w * ld + w;
            ub.M = ub.N = rem2 - w;
            int stt = gpb_launch_gemm(ub, batch, st);
            if (stt != GPB_OK) return stt;
        }
        // SYRK_a on the panel stream; it overwrites columns the previous SYRK_b also wrote
        if (have_b) GPB_CUDA(cudaStreamWaitEvent(ps, ev_b, 0));
        {
            GpbGemm ua = u;                             // top w x w square, lower tiles
            ua.M = ua.N = w;
            int stt = gpb_launch_gemm(ua, batch, ps);
            if (stt != GPB_OK) return stt;
            if (rem2 > w) {                             // rows below it, all w columns
                GpbGemm ur = u;
                ur.A = u.A + (long long)w * ld;
                ur.C = u.C + (long long)w * ld;
                ur.M = rem2 - w; ur.N = w; ur.lower_only = 0;
                stt = gpb_launch_gemm(ur, batch, ps);
                if (stt != GPB_OK) return stt;
            }
        }
        if (rem2 > w) {
            GPB_CUDA(cudaEventRecord(ev_b, st));
            have_b = true;
        }
    }
    if (lookahead) {
        GPB_CUDA(cudaEventRecord(ev_end, ps));
        GPB_CUDA(cudaStreamWaitEvent(st, ev_end, 0));
    }
    return GPB_OK;
}

// W = L^-1 (lower) and V = W^T (upper) from the inverted 128-blocks left by potrf.
// Adjacent blocks are merged pairwise; all equal-sized pairs of a level share one launch.
int gpb_launch_trtri(const double* L, long long n, long long ld, long long sL, int batch, double* W,
                     long long ldw, long long sW, double* V, long long ldv, long long sV, double* T,
                     long long ldt, long long sT, cudaStream_t st) {
    GPB_REQUIRE(n > 0 && n % GPB_NB == 0, "n must be a positive multiple of 128");
    GPB_REQUIRE(L && W && V && T, "null pointer");
    std::vector<long long> bounds;
    for (long long b = 0; b <= n; b += GPB_NB) bounds.push_back(b);
    while (bounds.size() > 2) {
        std::vector<long long> nb;
        const size_t npairs = (bounds.size() - 1) / 2;
        size_t p = 0;
        while (p < npairs) {
            const long long lo = bounds[2 * p], mid = bounds[2 * p + 1], hi = bounds[2 * p + 2];
            const long long n1 = mid - lo, n2 = hi - mid;
            size_t q = p + 1;     // extend over following pairs with the same shape
            while (q < npairs && bounds[2 * q + 1] - bounds[2 * q] == n1 &&
                   bounds[2 * q + 2] - bounds[2 * q + 1] == n2 && bounds[2 * q] - bounds[2 * (q - 1)] == n1 + n2)
                q++;
            const int cnt = (int)(q - p);
            const long long step = n1 + n2;
            // Tt[c, r] = sum_j V11[c, j] L21[r, j]      (c in block 1, r in block 2)
            GpbGemm g = gpb_gemm_default();
            g.A = V + lo * ldv + lo; g.lda = ldv; g.sA = sV; g.tA = step * (ldv + 1);
            g.B = L + mid * ld + lo; g.ldb = ld; g.sB = sL; g.tB = step * (ld + 1);
            g.C = T + lo * ldt + mid; g.ldc = ldt; g.sC = sT; g.tC = step * (ldt + 1);
            g.M = (int)n1; g.N = (int)n2; g.K = (int)n1; g.a_tri = 2; g.nb1 = cnt;
            int stt = gpb_launch_gemm(g, batch, st);
            if (stt != GPB_OK) return stt;
            // W21[r, c] = -sum_q W22[r, q] Tt[c, q]   and the mirrored V12[c, r]
            GpbGemm h = gpb_gemm_default();
            h.A = W + mid * ldw + mid; h.lda = ldw; h.sA = sW; h.tA = step * (ldw + 1);
            h.B = g.C; h.ldb = ldt; h.sB = sT; h.tB = step * (ldt + 1);
            h.C = W + mid * ldw + lo; h.ldc = ldw; h.sC = sW; h.tC = step * (ldw + 1);
            h.Ct = V + lo * ldv + mid; h.ldct = ldv; h.sCt = sV; h.tCt = step * (ldv + 1);
            h.M = (int)n2; h.N = (int)n1; h.K = (int)n2; h.a_tri = 1; h.alpha = -1.0; h.nb1 = cnt;
            stt = gpb_launch_gemm(h, batch, st);
            if (stt != GPB_OK) return stt;
            p = q;
        }
        for (size_t i = 0; i + 2 < bounds.size(); i += 2) nb.push_back(bounds[i]);
        if ((bounds.size() - 1) % 2 == 1) nb.push_back(bounds[bounds.size() - 2]);
        nb.push_back(bounds.back());
        bounds.swap(nb);
    }
    return GPB_OK;
}

// Ki = V V^T = L^-T L^-1  (the reference's inv(L).T @ inv(L), gp.py:311-312)
int gpb_launch_lauum(const double* V, long long n, long long ldv, long long sV, int batch, double* Ki,
                     long long ldk, long long sK, cudaStream_t st) {
    GPB_REQUIRE(n > 0 && n % GPB_NB == 0, "n must be a positive multiple of 128");
    GpbGemm g = gpb_gemm_default();
    g.A = V; g.lda = ldv; g.sA = sV;
    g.B = V; g.ldb = ldv; g.sB = sV;
    g.C = Ki; g.ldc = ldk; g.sC = sK;
    g.Ct = Ki; g.ldct = ldk; g.sCt = sK;
    g.M = g.N = g.K = (int)n;
    g.a_tri = 2; g.b_tri = 2; g.lower_only = 1;
    return gpb_launch_gemm(g, batch, st);
}

// alpha = K^-1 y by forward + backward substitution (cho_solve, gp.py:332-334).
// flags: int workspace of at least 2 * batch * T + 2 entries.
int gpb_launch_potrs(const double* L, const double* W, long long n, long long ld, long long ldw,
                     long long sL, long long sW, int batch, const double* y, long long sy, double* z,
                     double* alpha, long long svec, int* flags, cudaStream_t st) {
    GPB_REQUIRE(n > 0 && n % GPB_NB == 0, "n must be a positive multiple of 128");
    GPB_REQUIRE(ld % 2 == 0 && ldw % 2 == 0 && sL % 2 == 0 && sW % 2 == 0, "strides must be even");
    const int T = (int)(n / GPB_NB);
    const size_t nfl = (size_t)2 * batch * T + 2;
    static bool attr_set = false;
    if (!attr_set) {
        GPB_CUDA(cudaFuncSetAttribute(trsv_fwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, TRSV_FWD_SMEM));
        GPB_CUDA(cudaFuncSetAttribute(trsv_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, TRSV_BWD_SMEM));
        attr_set = true;
    }
    GPB_REQUIRE(((reinterpret_cast<uintptr_t>(L) | reinterpret_cast<uintptr_t>(W)) & 15) == 0, "L and W must be 16-byte aligned");
    GPB_CUDA(cudaMemsetAsync(flags, 0, nfl * sizeof(int), st));
    GpbProfScope prof(GPB_KC_SOLVE, st);
    int* fF = flags + 2;
    int* fB = fF + (size_t)batch * T;
    trsv_fwd_kernel<<<batch * T, 256, TRSV_FWD_SMEM, st>>>(L, ld, sL, W, ldw, sW, y, sy, z, svec, T, fF, flags);
    GPB_LAUNCH_CHECK("trsv_fwd_kernel");
    trsv_bwd_kernel<<<batch * T, 256, TRSV_BWD_SMEM, st>>>(L, ld, sL, W, ldw, sW, z, alpha, svec, T, fB, flags + 1);
    GPB_LAUNCH_CHECK("trsv_bwd_kernel");
    return GPB_OK;
}

int gpb_launch_loglh(const double* L, long long n_valid, long long ld, long long sL, int batch,
                     const double* y, long long sy, const double* alpha, long long svec,
                     const int* info, double* out3, cudaStream_t st) {
    GpbProfScope prof(GPB_KC_REDUCE, st);
    loglh_kernel<<<batch, 256, 0, st>>>(L, n_valid, ld, sL, y, sy, alpha, svec, info, out3);
    GPB_LAUNCH_CHECK("loglh_kernel");
    return GPB_OK;
}

int gpb_launch_tril(double* A, long long n, long long ld, long long sA, int batch, cudaStream_t st) {
    dim3 block(32, 8);
    dim3 grid((unsigned)((n + 31) / 32), (unsigned)((n + 7) / 8), (unsigned)batch);
    tril_kernel<<<grid, block, 0, st>>>(A, n, ld, sA);
    GPB_LAUNCH_CHECK("tril_kernel");
    return GPB_OK;
}

int gpb_launch_tril_copy(double* dst, long long ldd, const double* src, long long lds, long long n, cudaStream_t st) {
    if (n == 0) return GPB_OK;
    dim3 block(32, 8);
    long long gy = (n + 7) / 8;
    if (gy > 16384) gy = 16384;
    dim3 grid((unsigned)((n + 31) / 32), (unsigned)gy, 1);
    tril_copy_kernel<<<grid, block, 0, st>>>(dst, ldd, src, lds, n);
    GPB_LAUNCH_CHECK("tril_copy_kernel");
    return GPB_OK;
}

int gpb_launch_copy2d(double* dst, long long ldd, const double* src, long long lds, long long rows,
                      long long cols, long long sD, long long sS, int batch, cudaStream_t st) {
    if (rows == 0 || cols == 0) return GPB_OK;
    dim3 block(32, 8);
    long long gy = (rows + 7) / 8;
    if (gy > 16384) gy = 16384;
    dim3 grid((unsigned)((cols + 31) / 32), (unsigned)gy, (unsigned)batch);
    copy2d_kernel<<<grid, block, 0, st>>>(dst, ldd, src, lds, rows, cols, sD, sS);
    GPB_LAUNCH_CHECK("copy2d_kernel");
    return GPB_OK;
}
