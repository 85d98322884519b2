// small.cu -- everything after the factorisation for a GP that fits ONE 128-block
// (N <= 128: the reference's own test-suite scale, gp/tests/*.py use N <= 50).
//
// At this size the multi-kernel chain (two substitution kernels, log-likelihood reduction,
// lauum GEMM, two gradient-reduction kernels, result packing) is pure launch latency, so one
// CTA per GP does all of it from shared memory:
//     alpha = W^T (W y)                      cho_solve, gp/gp.py:332-334   (W = L^-1 from potrf_diag_kernel)
//     log_lh with the reference's clamps     gp/ext/gp_c.pyx:17-31         (logdet = 2 sum log L_ii)
//     Ki = W^T W                             gp/gp.py:311-312              (DMMA 16x32 strips)
//     a^T dK_i a, sum(Ki o dK_i), tr Ki, a.a gp/ext/gp_c.pyx:41-49         (fused into the Ki epilogue:
//                                            dK regenerated from x while Ki is still in registers)
// Outputs are written in the padded [128 x 128] layout the rest of the library uses, so the
// posterior / second-derivative paths run unchanged on them.
#include "kfunctors.cuh"
#include "launch.h"

namespace {

constexpr int TB = 32;                 // sub-block edge
constexpr int TLD = 132;               // 132 = 4 (mod 16): DMMA fragment loads stay at ~the 2-wavefront minimum
constexpr int TAIL_SMEM = (GPB_NB * TLD + 5 * GPB_NB + 64) * 8;

struct TailArgs {
    KParams P;
    const KParams* Pb;
    const double* x;        // [n]
    const double* y;        // [>= n] (+ b * sy)
    long long sy;
    int n;                  // observations (<= 128)
    const double* L;        // factor: only the diagonal is read
    long long ldl, sL;
    const double* W;        // L^-1, lower, [128 x 128]
    long long ldw, sW;
    double* Ki;             // out [128 x 128] or nullptr (no inverse / gradient)
    long long ldk, sK;
    double* z;              // out [128] (W y)
    double* alpha;          // out [128]
    long long svec;
    const int* info;
    double* out3;           // [batch][3]
    double* out16;          // [batch][16]
    double* Wz;             // optional: W / V again, writable: this kernel writes their structural zeros
    double* Vz;
    long long ldv, sV;
    double* pack;           // optional (batch 1): {out3 | out16 | info | 0...}, the block gpb_gp_stages reads back
};

// acc (rows hf*16 .. +16 of a 32 x 32 block) += X Y with X[m][k] = xp[m*xsm + k*xsk],
// Y[k][c] = yp[k*ysk + c*ysn]: eight independent DMMA accumulators per warp.
__device__ __forceinline__ void strip_mm_g(double (&acc)[2][4][2], const double* xp, int xsm, int xsk,
                                           const double* yp, int ysk, int ysn, int hf, int g, int t) {
#pragma unroll
    for (int kk = 0; kk < 8; kk++) {
        double a[2], b[4];
#pragma unroll
        for (int rt = 0; rt < 2; rt++) a[rt] = xp[(hf * 16 + rt * 8 + g) * xsm + (kk * 4 + t) * xsk];
#pragma unroll
        for (int ct = 0; ct < 4; ct++) b[ct] = yp[(kk * 4 + t) * ysk + (ct * 8 + g) * ysn];
#pragma unroll
        for (int rt = 0; rt < 2; rt++)
#pragma unroll
            for (int ct = 0; ct < 4; ct++) dmma884(acc[rt][ct][0], acc[rt][ct][1], a[rt], b[ct]);
    }
}

template <int KIND>
__global__ void __launch_bounds__(256, 1) small_tail_kernel(const TailArgs a) {
    extern __shared__ __align__(16) double sm[];
    double* Ws = sm;                          // W, row-major, stride TLD
    double* xs = sm + GPB_NB * TLD;           // x (0 beyond n)
    double* ys = xs + GPB_NB;
    double* zs = ys + GPB_NB;
    double* as = zs + GPB_NB;
    double* red = as + GPB_NB;                // 32 + slack
    __shared__ KParams sP;
    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5, g = lane >> 2, t = lane & 3;
    const int b = blockIdx.x;
    const int n = a.n;
    const int nb = (n + TB - 1) / TB, nn = nb * TB;
    const double* W = a.W + b * a.sW;
    const double* L = a.L + b * a.sL;
    const double* y = a.y + b * a.sy;

    if (a.Pb) {
        const double* src = reinterpret_cast<const double*>(a.Pb + b);
        double* dst = reinterpret_cast<double*>(&sP);
        for (int i = tid; i < (int)(sizeof(KParams) / 8); i += 256) dst[i] = src[i];
    } else if (tid == 0) {
        sP = a.P;
    }
    // ---- W (nn x nn) -> shared memory; sub-blocks above the diagonal are structural zeros ----
    for (int e = tid; e < nn * (nn / 2); e += 256) {
        const int r = e / (nn / 2), c2 = (e % (nn / 2)) * 2;
        if (c2 / TB > r / TB) *reinterpret_cast<double2*>(Ws + r * TLD + c2) = make_double2(0.0, 0.0);
        else cp_async16(Ws + r * TLD + c2, W + (long long)r * a.ldw + c2);
    }
    cp_async_commit();
    if (a.Wz) {
        // ... which this kernel also writes to global W (upper sub-blocks) and V = W^T (lower ones)
        double* Wz = a.Wz + b * a.sW;
        double* Vz = a.Vz ? a.Vz + b * a.sV : nullptr;
        for (int e = tid; e < GPB_NB * (GPB_NB / 2); e += 256) {
            const int r = e / (GPB_NB / 2), c2 = (e % (GPB_NB / 2)) * 2;
            if (c2 / TB > r / TB) {
                *reinterpret_cast<double2*>(Wz + (long long)r * a.ldw + c2) = make_double2(0.0, 0.0);
            } else if (c2 / TB < r / TB && Vz) {
                *reinterpret_cast<double2*>(Vz + (long long)r * a.ldv + c2) = make_double2(0.0, 0.0);
            }
        }
    }
    if (tid < GPB_NB) {
        xs[tid] = (tid < n) ? a.x[tid] : 0.0;
        ys[tid] = (tid < n) ? y[tid] : 0.0;
    }
    cp_async_wait<0>();
    __syncthreads();

    // ---- z = W y: one warp per row, lanes over the columns -------------------------------
    for (int r = wid; r < nn; r += 8) {
        double s = 0.0;
        for (int c = lane; c <= r; c += 32) s = fma(Ws[r * TLD + c], ys[c], s);
        s = warp_sum(s);
        if (lane == 0) zs[r] = s;
    }
    __syncthreads();
    // ---- alpha = W^T z: thread = column (conflict-free), two partial sums ------------------
    if (tid < GPB_NB) {
        double s0 = 0.0, s1 = 0.0;
        if (tid < nn) {
            int r = tid;
            for (; r + 1 < nn; r += 2) {
                s0 = fma(Ws[r * TLD + tid], zs[r], s0);
                s1 = fma(Ws[(r + 1) * TLD + tid], zs[r + 1], s1);
            }
            if (r < nn) s0 = fma(Ws[r * TLD + tid], zs[r], s0);
        }
        const double al = s0 + s1;
        as[tid] = al;
        a.alpha[b * a.svec + tid] = al;
        a.z[b * a.svec + tid] = (tid < nn) ? zs[tid] : 0.0;
    }
    __syncthreads();

    // ---- log_lh (same thread mapping and reduction as loglh_kernel) --------------------------
    double sl = 0.0, sq = 0.0, saa = 0.0;
    if (tid < n) {
        sl = log(L[(long long)tid * a.ldl + tid]);
        sq = ys[tid] * as[tid];
        saa = as[tid] * as[tid];
    }
    sl = block_sum(sl, red);
    sq = block_sum(sq, red);
    saa = block_sum(saa, red);
    if (tid == 0) {
        const double logdet = 2.0 * sl;
        double llh;
        if ((a.info && a.info[b] != 0) || logdet < GPB_MIN_LOG) llh = -INFINITY;
        else llh = -0.5 * sq + -0.5 * logdet + -0.5 * (double)n * log(2.0 * M_PI);
        a.out3[b * 3 + 0] = llh;
        a.out3[b * 3 + 1] = logdet;
        a.out3[b * 3 + 2] = sq;
        if (a.pack) {
            a.pack[0] = llh; a.pack[1] = logdet; a.pack[2] = sq;
            a.pack[19] = a.info ? (double)a.info[b] : 0.0;
            for (int i = 20; i < 24; i++) a.pack[i] = 0.0;
        }
    }
    if (!a.Ki) return;

    // ---- Ki = W^T W on DMMA, gradient brackets fused into the epilogue ------------------------
    double* Ki = a.Ki + b * a.sK;
    constexpr int NP = (KIND == GPB_GAUSSIAN) ? 2 : 3;
    constexpr unsigned NEED = (KIND == GPB_GAUSSIAN) ? 0x6u : 0xEu;      // unique ids of the Jacobian slices
    double t0[NP], t1[NP], tr = 0.0;
#pragma unroll
    for (int q = 0; q < NP; q++) t0[q] = t1[q] = 0.0;
    const int nitems = nb * (nb + 1);                 // (block pairs bj <= bi) x 2 half-strips
    for (int item = wid; item < nitems; item += 8) {
        const int pr = item >> 1, hf = item & 1;
        int bi = 0;
        while ((bi + 1) * (bi + 2) / 2 <= pr) bi++;
        const int bj = pr - bi * (bi + 1) / 2;
        double acc[2][4][2];
#pragma unroll
        for (int rt = 0; rt < 2; rt++)
#pragma unroll
            for (int ct = 0; ct < 4; ct++) acc[rt][ct][0] = acc[rt][ct][1] = 0.0;
        // Ki[bi, bj] = sum_{kb >= bi} W[kb, bi]^T W[kb, bj]
        for (int kb = bi; kb < nb; kb++)
            strip_mm_g(acc, Ws + kb * TB * TLD + bi * TB, 1, TLD, Ws + kb * TB * TLD + bj * TB, TLD, 1, hf, g, t);
        const double wgt = (bi == bj) ? 1.0 : 2.0;    // the mirrored block is not visited
#pragma unroll
        for (int rt = 0; rt < 2; rt++)
#pragma unroll
            for (int ct = 0; ct < 4; ct++) {
                const int r = bi * TB + hf * 16 + rt * 8 + g, c = bj * TB + ct * 8 + 2 * t;
                const double v0 = acc[rt][ct][0], v1 = acc[rt][ct][1];
                *reinterpret_cast<double2*>(Ki + (long long)r * a.ldk + c) = make_double2(v0, v1);
                if (bi != bj) {
                    Ki[(long long)c * a.ldk + r] = v0;
                    Ki[(long long)(c + 1) * a.ldk + r] = v1;
                }
                if (r < n) {
                    const double ar = as[r];
#pragma unroll
                    for (int e = 0; e < 2; e++) {
                        const int cc = c + e;
                        if (cc >= n) continue;
                        const double kv = e ? v1 : v0;
                        double u[10];
                        gpb_eval_unique<KIND>(sP, xs[r] - xs[cc], NEED, u);
                        const double aw = wgt * ar * as[cc], kw = wgt * kv;
#pragma unroll
                        for (int q = 0; q < NP; q++) {
                            t0[q] = fma(aw, u[1 + q], t0[q]);
                            t1[q] = fma(kw, u[1 + q], t1[q]);
                        }
                        if (r == cc) tr += kv;
                    }
                }
            }
    }
    // identity in the pad region of Ki (rows / columns nn .. 127)
    for (int e = tid; e < GPB_NB * GPB_NB; e += 256) {
        const int r = e >> 7, c = e & 127;
        if (r >= nn || c >= nn) Ki[(long long)r * a.ldk + c] = (r == c) ? 1.0 : 0.0;
    }
    double* o = a.out16 + (long long)b * 16;
#pragma unroll
    for (int q = 0; q < NP; q++) {
        const double v0 = block_sum(t0[q], red);
        const double v1 = block_sum(t1[q], red);
        if (tid == 0) { o[q] = v0; o[6 + q] = v1; }
    }
    tr = block_sum(tr, red);
    if (tid == 0) {
        for (int q = NP; q < 6; q++) o[q] = o[6 + q] = 0.0;
        o[12] = tr;
        o[13] = saa;
        o[14] = o[15] = 0.0;
        if (a.pack)
            for (int i = 0; i < 16; i++) a.pack[3 + i] = o[i];
    }
}

// ---------------------------------------------------------------------------
// Posterior covariance of a one-block GP at up to 128 test points in ONE launch (gp/gp.py:599-625):
//   Kx = K(xo, x) generated into shared memory, Z = Kx W^T (W = L^-1 read through L2),
//   cov = K(xo, xo) - Z Z^T with K(xo, xo) generated in the epilogue (lower blocks + mirror).
// The four-launch path (two builders, two TMA GEMMs) spends ~47 us of kernel latency on what is
// 2 MFLOP of work at the reference's test-suite scale.
// ---------------------------------------------------------------------------
struct CovArgs {
    KParams P;
    const double* xo;       // [m] device
    const double* x;        // [n] device
    int m, n;
    const double* W;        // [128 x 128] lower
    long long ldw;
    double* out;            // [m x m], row stride ldo
    long long ldo;
    int ldz;                // shared-memory row stride (= roundup(n, 32) + 4)
};

template <int KIND>
__global__ void __launch_bounds__(256, 1) small_cov_kernel(const CovArgs a) {
    extern __shared__ __align__(16) double sm[];
    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5, g = lane >> 2, t = lane & 3;
    const int m = a.m, n = a.n, LD = a.ldz;
    const int nb = (n + TB - 1) / TB, nn = nb * TB;          // k / column blocks of Kx and Z
    const int mb = (m + TB - 1) / TB, mm = mb * TB;          // row blocks
    double* Kx = sm;                                         // [mm][LD]
    double* Zs = sm + (size_t)mm * LD;                       // [mm][LD]
    double* xs = Zs + (size_t)mm * LD;                       // x  (nn)
    double* xos = xs + GPB_NB;                               // xo (mm)
    __shared__ KParams sP;
    if (tid == 0) sP = a.P;
    if (tid < GPB_NB) {
        xs[tid] = (tid < n) ? a.x[tid] : 0.0;
        xos[tid] = (tid < m) ? a.xo[tid] : 0.0;
    }
    __syncthreads();
    // ---- Kx[r][c] = k(xo_r - x_c), zero in the pad ----------------------------------------
    for (int e = tid; e < mm * (nn / 2); e += 256) {
        const int r = e / (nn / 2), c = (e % (nn / 2)) * 2;
        double u[2][10];
        const double dd[2] = {xos[r] - xs[c], xos[r] - xs[c + 1]};
        gpb_eval_unique_v<KIND, 2>(sP, dd, 1u, u);
        const bool vr = r < m;
        Kx[r * LD + c] = (vr && c < n) ? u[0][0] : 0.0;
        Kx[r * LD + c + 1] = (vr && c + 1 < n) ? u[1][0] : 0.0;
    }
    __syncthreads();
    // ---- Z[rows of block bi][block bj] = sum_{kb <= bj} Kx[bi, kb] W[bj, kb]^T ----------------
    for (int item = wid; item < mb * nb * 2; item += 8) {
        const int hf = item & 1, blkid = item >> 1, bi = blkid / nb, bj = blkid % nb;
        double acc[2][4][2];
#pragma unroll
        for (int rt = 0; rt < 2; rt++)
#pragma unroll
            for (int ct = 0; ct < 4; ct++) acc[rt][ct][0] = acc[rt][ct][1] = 0.0;
        for (int kb = 0; kb <= bj; kb++)
            strip_mm_g(acc, Kx + bi * TB * LD + kb * TB, LD, 1,
                       a.W + (long long)bj * TB * a.ldw + kb * TB, 1, (int)a.ldw, hf, g, t);
#pragma unroll
        for (int rt = 0; rt < 2; rt++)
#pragma unroll
            for (int ct = 0; ct < 4; ct++) {
                double* p = Zs + (bi * TB + hf * 16 + rt * 8 + g) * LD + bj * TB + ct * 8 + 2 * t;
                p[0] = acc[rt][ct][0];
                p[1] = acc[rt][ct][1];
            }
    }
    __syncthreads();
    // ---- cov[bi, bj] = K(xo, xo) - Z[bi] Z[bj]^T for bj <= bi, mirrored ----------------------
    const int npair = mb * (mb + 1) / 2;
    for (int item = wid; item < npair * 2; item += 8) {
        const int pr = item >> 1, hf = item & 1;
        int bi = 0;
        while ((bi + 1) * (bi + 2) / 2 <= pr) bi++;
        const int bj = pr - bi * (bi + 1) / 2;
        double acc[2][4][2];
#pragma unroll
        for (int rt = 0; rt < 2; rt++)
#pragma unroll
            for (int ct = 0; ct < 4; ct++) acc[rt][ct][0] = acc[rt][ct][1] = 0.0;
        for (int kb = 0; kb < nb; kb++)
            strip_mm_g(acc, Zs + bi * TB * LD + kb * TB, LD, 1, Zs + bj * TB * LD + kb * TB, 1, LD, hf, g, t);
#pragma unroll
        for (int rt = 0; rt < 2; rt++)
#pragma unroll
            for (int ct = 0; ct < 4; ct++) {
                const int r = bi * TB + hf * 16 + rt * 8 + g, c = bj * TB + ct * 8 + 2 * t;
                if (r >= m) continue;
                double u[2][10];
                const double dd[2] = {xos[r] - xos[c], xos[r] - xos[c + 1]};
                gpb_eval_unique_v<KIND, 2>(sP, dd, 1u, u);
#pragma unroll
                for (int e = 0; e < 2; e++) {
                    const int cc = c + e;
                    if (cc >= m || cc > r) continue;          // lower triangle; the mirror makes it exactly symmetric
                    const double v = u[e][0] - acc[rt][ct][e];
                    a.out[(long long)r * a.ldo + cc] = v;
                    if (cc != r) a.out[(long long)cc * a.ldo + r] = v;
                }
            }
    }
}

}  // namespace

size_t gpb_small_cov_smem(long long m, long long n) {
    const long long mm = (m + TB - 1) / TB * TB, ldz = (n + TB - 1) / TB * TB + 4;
    return (size_t)(2 * mm * ldz + 2 * GPB_NB) * 8;
}

// cov(xo) of a one-block GP in one launch; returns GPB_ERR_ARG (without setting an error text the
// caller would show) when the problem does not fit the shared-memory budget -> use the GEMM path.
int gpb_launch_small_cov(int kind, const KParams* P, const double* xo, long long m, const double* x, long long n,
                         const double* W, long long ldw, double* out, long long ldo, cudaStream_t st) {
    GPB_REQUIRE(kind == GPB_GAUSSIAN || kind == GPB_PERIODIC, "unknown kernel kind");
    GPB_REQUIRE(m >= 1 && m <= GPB_NB && n >= 1 && n <= GPB_NB && P && xo && x && W && out, "bad argument");
    const size_t smem = gpb_small_cov_smem(m, n);
    GPB_REQUIRE(smem <= (size_t)200 * 1024, "problem too large for the one-launch covariance");
    static bool attr_set = false;
    if (!attr_set) {
        GPB_CUDA(cudaFuncSetAttribute(small_cov_kernel<GPB_GAUSSIAN>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
        GPB_CUDA(cudaFuncSetAttribute(small_cov_kernel<GPB_PERIODIC>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
        attr_set = true;
    }
    CovArgs a;
    a.P = *P; a.xo = xo; a.x = x; a.m = (int)m; a.n = (int)n; a.W = W; a.ldw = ldw; a.out = out; a.ldo = ldo;
    a.ldz = (int)((n + TB - 1) / TB * TB + 4);
    GpbProfScope prof(GPB_KC_GEMM, st);
    if (kind == GPB_GAUSSIAN) small_cov_kernel<GPB_GAUSSIAN><<<1, 256, smem, st>>>(a);
    else small_cov_kernel<GPB_PERIODIC><<<1, 256, smem, st>>>(a);
    GPB_LAUNCH_CHECK("small_cov_kernel");
    return GPB_OK;
}

// One CTA per GP: alpha, log_lh and (when Ki != nullptr) K^-1 + the gradient brackets.
// W must hold the complete L^-1 of a single 128-block (potrf with n == 128).
int gpb_launch_small_tail(int kind, const KParams* P, const KParams* Pb, int batch, const double* x, long long n,
                          const double* y, long long sy, const double* L, long long ldl, long long sL,
                          const double* W, long long ldw, long long sW, double* Ki, long long ldk, long long sK,
                          double* z, double* alpha, long long svec, const int* info, double* out3,
                          double* out16, double* Wz, double* Vz, long long ldv, long long sV, double* pack,
                          cudaStream_t st) {
    GPB_REQUIRE(kind == GPB_GAUSSIAN || kind == GPB_PERIODIC, "unknown kernel kind");
    GPB_REQUIRE(n >= 1 && n <= GPB_NB && batch >= 1, "small tail needs 1 <= n <= 128");
    GPB_REQUIRE(x && y && L && W && z && alpha && out3 && (!Ki || out16), "null pointer");
    GPB_REQUIRE((reinterpret_cast<uintptr_t>(W) & 15) == 0 && ldw % 2 == 0 && sW % 2 == 0, "W must be 16-byte aligned");
    GPB_REQUIRE(!Ki || ((reinterpret_cast<uintptr_t>(Ki) & 15) == 0 && ldk % 2 == 0 && sK % 2 == 0), "Ki must be 16-byte aligned");
    static bool attr_set = false;
    if (!attr_set) {
        GPB_CUDA(cudaFuncSetAttribute(small_tail_kernel<GPB_GAUSSIAN>, cudaFuncAttributeMaxDynamicSharedMemorySize, TAIL_SMEM));
        GPB_CUDA(cudaFuncSetAttribute(small_tail_kernel<GPB_PERIODIC>, cudaFuncAttributeMaxDynamicSharedMemorySize, TAIL_SMEM));
        attr_set = true;
    }
    TailArgs a;
    if (P) a.P = *P; else memset(&a.P, 0, sizeof(KParams));
    a.Pb = Pb; a.x = x; a.y = y; a.sy = sy; a.n = (int)n;
    a.L = L; a.ldl = ldl; a.sL = sL; a.W = W; a.ldw = ldw; a.sW = sW;
    a.Ki = Ki; a.ldk = ldk; a.sK = sK; a.z = z; a.alpha = alpha; a.svec = svec;
    a.info = info; a.out3 = out3; a.out16 = out16;
    a.Wz = Wz; a.Vz = Vz; a.ldv = ldv; a.sV = sV; a.pack = (batch == 1) ? pack : nullptr;
    GPB_REQUIRE(!Vz || ((reinterpret_cast<uintptr_t>(Vz) & 15) == 0 && ldv % 2 == 0 && sV % 2 == 0), "V must be 16-byte aligned");
    GpbProfScope prof(GPB_KC_REDUCE, st);
    if (kind == GPB_GAUSSIAN) small_tail_kernel<GPB_GAUSSIAN><<<batch, 256, TAIL_SMEM, st>>>(a);
    else small_tail_kernel<GPB_PERIODIC><<<batch, 256, TAIL_SMEM, st>>>(a);
    GPB_LAUNCH_CHECK("small_tail_kernel");
    return GPB_OK;
}
