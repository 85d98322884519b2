// build.cu -- kernel-matrix builders and the kernels that consume kernel values
// without materialising them (fused kernel-times-vector products, fused gradient
// reductions over K^-1).  All fp64, HBM-write-bound (builders) or FP64-ALU-bound
// (fused consumers).  Replaces gp/ext/gaussian_c.pyx and periodic_c.pyx.
#include "kfunctors.cuh"
#include "launch.h"

// ---------------------------------------------------------------------------
// 1. slice builder: any subset of {K, jacobian slices, hessian slices} in one pass
// ---------------------------------------------------------------------------
struct BuildArgs {
    KParams P;
    const KParams* Pb;     // per-batch parameters (device) or nullptr -> P
    const double* x1;
    const double* x2;
    long long n1, n2;      // valid extents
    long long rows, cols;  // extents to fill (>= n1, n2): the pad region gets 0 / identity
    double* out[GPB_MAX_SLICES];
    long long ld;
    long long bstride;     // batch stride of every out[] slice
    int add_diag;          // slice 0: += s^2 on the index diagonal (gp.py:265)
    int pad_identity;      // slice 0: 1.0 on the diagonal of the pad region
    int lower_only;        // square x1 == x2 build for the factorisation: only the 32 x 64 tiles that touch
                           // the lower triangle (whole 64 x 64 diagonal blocks included) are generated
};

template <int KIND>
__device__ __forceinline__ void load_kparams(KParams* sP, const KParams& byval, const KParams* Pb, int b) {
    if (Pb) {
        const double* src = reinterpret_cast<const double*>(Pb + b);
        double* dst = reinterpret_cast<double*>(sP);
        for (int i = threadIdx.x + threadIdx.y * blockDim.x; i < (int)(sizeof(KParams) / 8);
             i += blockDim.x * blockDim.y)
            dst[i] = src[i];
    } else if (threadIdx.x == 0 && threadIdx.y == 0) {
        *sP = byval;
    }
    __syncthreads();
}

// block (32, 8): each thread owns 2 adjacent columns and 4 rows (stride 8) of a 32 x 64 tile.
// K only (slice 0), the common case (Kxx for the factorisation, K(xo, x), K(xo, xo)): no slice loop, the
// eight separations of a thread evaluated together.  The generic kernel below spends ~100 issued
// instructions per element on slice bookkeeping (ncu: FP64 pipe 38 % active); this one ~35.
// Periodic kernel: sin / cos of the tile's points, once per block (96 sincos for 2048 elements)
// sin / cos of a difference from the per-point tables, with separately rounded products (no FMA contraction):
// S(i,j) = -S(j,i) and C(i,j) = C(j,i) hold bit for bit, so K(x, x) stays exactly symmetric like the
// d = x_i - x_j form it replaces.
__device__ __forceinline__ double sin_diff(double si, double ci, double sj, double cj) {
    return __dsub_rn(__dmul_rn(si, cj), __dmul_rn(ci, sj));
}
__device__ __forceinline__ double cos_diff(double si, double ci, double sj, double cj) {
    return __dadd_rn(__dmul_rn(ci, cj), __dmul_rn(si, sj));
}

template <int KIND>
__device__ __forceinline__ void stage_sincos(const BuildArgs& a, const KParams& P, double* s_col, double* c_col,
                                             double* s_row, double* c_row) {
    if (KIND != GPB_PERIODIC) return;
    const int tid = threadIdx.y * 32 + threadIdx.x;
    if (tid < 64) {
        const long long j = (long long)blockIdx.x * 64 + tid;
        sincos((j < a.n2 ? a.x2[j] : 0.0) * P.half_ip, &s_col[tid], &c_col[tid]);
    } else if (tid < 96) {
        const long long i = (long long)blockIdx.y * 32 + tid - 64;
        sincos((i < a.n1 ? a.x1[i] : 0.0) * P.half_ip, &s_row[tid - 64], &c_row[tid - 64]);
    }
    __syncthreads();
}

template <int KIND>
__global__ void __launch_bounds__(256) build_k_kernel(const BuildArgs a) {
    __shared__ KParams sP;
    __shared__ double s_col[64], c_col[64], s_row[32], c_row[32];
    load_kparams<KIND>(&sP, a.P, a.Pb, blockIdx.z);
    if (a.lower_only && blockIdx.x > blockIdx.y / 2) return;
    stage_sincos<KIND>(a, sP, s_col, c_col, s_row, c_row);
    const long long j0 = ((long long)blockIdx.x * 32 + threadIdx.x) * 2;
    if (j0 >= a.cols) return;
    const bool ja = j0 < a.n2, jb = j0 + 1 < a.n2;
    const double xa = ja ? a.x2[j0] : 0.0, xb = jb ? a.x2[j0 + 1] : 0.0;
    double* o = a.out[0] + (long long)blockIdx.z * a.bstride;
    const long long i0 = (long long)blockIdx.y * 32 + threadIdx.y;
    double d[8], u[8][10];
#pragma unroll
    for (int r = 0; r < 4; r++) {
        const long long i = i0 + 8 * r;
        const double xi = (i < a.n1) ? a.x1[i] : 0.0;
        d[2 * r] = xi - xa;
        d[2 * r + 1] = xi - xb;
    }
    if (KIND == GPB_PERIODIC) {
        // K = k0 exp(c1 S^2), S = sin((x_i - x_j) / 2p) = s_i c_j - c_i s_j from the staged tables
        double arg[8], ev[8];
        const double sa = s_col[2 * threadIdx.x], ca = c_col[2 * threadIdx.x];
        const double sb = s_col[2 * threadIdx.x + 1], cb = c_col[2 * threadIdx.x + 1];
#pragma unroll
        for (int r = 0; r < 4; r++) {
            const double si = s_row[threadIdx.y + 8 * r], ci = c_row[threadIdx.y + 8 * r];
            const double S0 = sin_diff(si, ci, sa, ca), S1 = sin_diff(si, ci, sb, cb);
            arg[2 * r] = sP.c1 * (S0 * S0);
            arg[2 * r + 1] = sP.c1 * (S1 * S1);
        }
        gpb_expv<8>(arg, ev);
#pragma unroll
        for (int e = 0; e < 8; e++) u[e][0] = sP.k0 * ev[e];
    } else {
        gpb_eval_unique_v<KIND, 8>(sP, d, 1u, u);
    }
#pragma unroll
    for (int r = 0; r < 4; r++) {
        const long long i = i0 + 8 * r;
        if (i >= a.rows) break;
        const bool vi = i < a.n1;
        double v0 = (vi && ja) ? u[2 * r][0] : 0.0, v1 = (vi && jb) ? u[2 * r + 1][0] : 0.0;
        if (a.add_diag) {
            if (vi && ja && i == j0) v0 += sP.s2;
            if (vi && jb && i == j0 + 1) v1 += sP.s2;
        }
        if (a.pad_identity) {
            if (!(vi && ja) && i == j0) v0 = 1.0;
            if (!(vi && jb) && i == j0 + 1) v1 = 1.0;
        }
        *reinterpret_cast<double2*>(o + i * a.ld + j0) = make_double2(v0, v1);
    }
}

template <int KIND, bool VEC2>
__global__ void __launch_bounds__(256) build_kernel(const BuildArgs a) {
    __shared__ KParams sP;
    __shared__ double s_col[64], c_col[64], s_row[32], c_row[32];
    load_kparams<KIND>(&sP, a.P, a.Pb, blockIdx.z);
    constexpr int NS = (KIND == GPB_GAUSSIAN) ? 7 : 13;

    unsigned need = 0;
#pragma unroll
    for (int s = 0; s < NS; s++)
        if (a.out[s]) need |= 1u << gpb_slice_to_unique(KIND, s);

    if (a.lower_only && blockIdx.x > blockIdx.y / 2) return;     // tile strictly above the 64-block diagonal
    stage_sincos<KIND>(a, sP, s_col, c_col, s_row, c_row);
    const long long j0 = ((long long)blockIdx.x * 32 + threadIdx.x) * 2;
    if (j0 >= a.cols) return;
    const bool two = (j0 + 1 < a.cols);
    const double xa = (j0 < a.n2) ? a.x2[j0] : 0.0;
    const double xb = (j0 + 1 < a.n2) ? a.x2[j0 + 1] : 0.0;
    const long long boff = (long long)blockIdx.z * a.bstride;

#pragma unroll
    for (int r = 0; r < 4; r++) {
        const long long i = (long long)blockIdx.y * 32 + threadIdx.y + 8 * r;
        if (i >= a.rows) break;
        double ua[10], ub[10];
        const bool vi = i < a.n1;
        const bool va = vi && (j0 < a.n2), vb = vi && (j0 + 1 < a.n2);
        if (va | vb) {
            const double xi = a.x1[i];
            if (KIND == GPB_PERIODIC) {
                const double si = s_row[threadIdx.y + 8 * r], ci = c_row[threadIdx.y + 8 * r];
                const double sa = s_col[2 * threadIdx.x], ca = c_col[2 * threadIdx.x];
                const double sb = s_col[2 * threadIdx.x + 1], cb = c_col[2 * threadIdx.x + 1];
                gpb_eval_periodic_sc(sP, xi - xa, sin_diff(si, ci, sa, ca), cos_diff(si, ci, sa, ca), need, ua);
                gpb_eval_periodic_sc(sP, xi - xb, sin_diff(si, ci, sb, cb), cos_diff(si, ci, sb, cb), need, ub);
            } else {
                gpb_eval_unique<KIND>(sP, xi - xa, need, ua);
                gpb_eval_unique<KIND>(sP, xi - xb, need, ub);
            }
        }
#pragma unroll
        for (int s = 0; s < NS; s++) {
            double* o = a.out[s];
            if (!o) continue;
            const int u = gpb_slice_to_unique(KIND, s);
            double v0 = va ? ua[u] : 0.0, v1 = vb ? ub[u] : 0.0;
            if (s == 0) {
                if (a.add_diag) {
                    if (va && i == j0) v0 += sP.s2;
                    if (vb && i == j0 + 1) v1 += sP.s2;
                }
                if (a.pad_identity) {
                    if (!va && i == j0) v0 = 1.0;
                    if (!vb && i == j0 + 1) v1 = 1.0;
                }
            }
            double* p = o + boff + i * a.ld + j0;
            if (VEC2) {
                *reinterpret_cast<double2*>(p) = make_double2(v0, v1);   // cols is even when VEC2
            } else {
                p[0] = v0;
                if (two) p[1] = v1;
            }
        }
    }
}

int gpb_launch_build(int kind, const KParams* P, const KParams* Pb, int batch, const double* x1,
                     long long n1, const double* x2, long long n2, long long rows, long long cols,
                     double* const* out, long long ld, long long bstride, int add_diag,
                     int pad_identity, cudaStream_t st, int lower_only) {
    GPB_REQUIRE(kind == GPB_GAUSSIAN || kind == GPB_PERIODIC, "unknown kernel kind");
    GPB_REQUIRE(rows >= n1 && cols >= n2 && ld >= cols, "bad extents");
    GPB_REQUIRE(!lower_only || (rows == cols && n1 == n2), "lower_only needs a square build");
    if (rows == 0 || cols == 0) return GPB_OK;
    BuildArgs a;
    if (P) a.P = *P; else memset(&a.P, 0, sizeof(KParams));
    a.Pb = Pb;
    a.x1 = x1; a.x2 = x2; a.n1 = n1; a.n2 = n2; a.rows = rows; a.cols = cols;
    const int ns = gpb_n_slices(kind);
    bool vec = (ld % 2 == 0) && (cols % 2 == 0) && (bstride % 2 == 0);
    bool any = false;
    for (int s = 0; s < GPB_MAX_SLICES; s++) {
        a.out[s] = (s < ns) ? out[s] : nullptr;
        if (a.out[s]) {
            any = true;
            if (reinterpret_cast<uintptr_t>(a.out[s]) % 16) vec = false;
        }
    }
    if (!any) return GPB_OK;
    a.ld = ld; a.bstride = bstride; a.add_diag = add_diag; a.pad_identity = pad_identity; a.lower_only = lower_only;
    dim3 block(32, 8);
    dim3 grid((unsigned)((cols + 63) / 64), (unsigned)((rows + 31) / 32), (unsigned)batch);
    GPB_REQUIRE(grid.y <= 65535 && batch <= 65535, "extent too large for the launch grid");
    GpbProfScope prof(GPB_KC_BUILD, st);
    bool k_only = vec && a.out[0] != nullptr;
    for (int s = 1; s < GPB_MAX_SLICES; s++) k_only = k_only && a.out[s] == nullptr;
    if (k_only) {
        if (kind == GPB_GAUSSIAN) build_k_kernel<GPB_GAUSSIAN><<<grid, block, 0, st>>>(a);
        else build_k_kernel<GPB_PERIODIC><<<grid, block, 0, st>>>(a);
    } else if (kind == GPB_GAUSSIAN) {
        if (vec) build_kernel<GPB_GAUSSIAN, true><<<grid, block, 0, st>>>(a);
        else build_kernel<GPB_GAUSSIAN, false><<<grid, block, 0, st>>>(a);
    } else {
        if (vec) build_kernel<GPB_PERIODIC, true><<<grid, block, 0, st>>>(a);
        else build_kernel<GPB_PERIODIC, false><<<grid, block, 0, st>>>(a);
    }
    GPB_LAUNCH_CHECK("build_kernel");
    return GPB_OK;
}

// ---------------------------------------------------------------------------
// 2. fused kernel-times-vector: out[g][r] = sum_p coef_p * sum_c slice_p(x1[r], x2[c]) * vec_p[c]
//    (posterior mean gp.py:597, dm_dtheta gp_c.pyx:122-131, dK_i * alpha for d2lh)
//    -- the M x N kernel matrix is never written to HBM.
// ---------------------------------------------------------------------------
#define GPB_MV_MAXP 8
#define GPB_MV_MAXO 4
struct MatvecArgs {
    KParams P;
    const KParams* Pb;
    const double* x1;
    const double* x2;
    long long n1, n2;
    int npairs, nout;
    int slice[GPB_MV_MAXP];
    int outidx[GPB_MV_MAXP];
    double coef[GPB_MV_MAXP];
    const double* vec[GPB_MV_MAXP];
    double* out[GPB_MV_MAXO];
    long long vstride, ostride;   // batch strides of vec / out
};

// one warp per row, 8 rows per block; lanes stride the columns.
template <int KIND>
__global__ void __launch_bounds__(256) fused_matvec_kernel(const MatvecArgs a) {
    __shared__ KParams sP;
    load_kparams<KIND>(&sP, a.P, a.Pb, blockIdx.z);
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    unsigned need = 0;
    for (int p = 0; p < a.npairs; p++) need |= 1u << gpb_slice_to_unique(KIND, a.slice[p]);
    const long long voff = (long long)blockIdx.z * a.vstride;
    const long long ooff = (long long)blockIdx.z * a.ostride;

    for (long long r = (long long)blockIdx.x * 8 + wid; r < a.n1; r += (long long)gridDim.x * 8) {
        const double xi = a.x1[r];
        double acc[GPB_MV_MAXO] = {0, 0, 0, 0};
        for (long long c = lane; c < a.n2; c += 32) {
            double u[10];
            gpb_eval_unique<KIND>(sP, xi - a.x2[c], need, u);
#pragma unroll
            for (int p = 0; p < GPB_MV_MAXP; p++) {
                if (p < a.npairs) {
                    const double v = gpb_pick(u, gpb_slice_to_unique(KIND, a.slice[p])) * a.vec[p][voff + c];
                    const double t = a.coef[p] * v;
#pragma unroll
                    for (int g = 0; g < GPB_MV_MAXO; g++)
                        if (a.outidx[p] == g) acc[g] += t;
                }
            }
        }
#pragma unroll
        for (int g = 0; g < GPB_MV_MAXO; g++) {
            if (g < a.nout) {
                const double s = warp_sum(acc[g]);
                if (lane == 0) a.out[g][ooff + r] = s;
            }
        }
    }
}

// Posterior-mean fast path (gp.py:597, one pair: slice 0, coefficient 1): the slice is a
// compile-time constant, a warp owns two rows so every x2 / vector load feeds two kernel
// evaluations, and lanes take two adjacent columns per 128-bit load.
// Periodic posterior mean: the block's 16 rows against all columns, the columns' sin / cos staged in
// shared memory by chunks of MEAN_CH points (one sincos per column per block instead of one per element):
// S = s_i c_j - c_i s_j, K = k0 exp(c1 S^2).
#define MEAN_CH 1024
__global__ void __launch_bounds__(256) mean_periodic_kernel(const MatvecArgs a) {
    __shared__ KParams sP;
    __shared__ double s_c[MEAN_CH], c_c[MEAN_CH], v_c[MEAN_CH];
    load_kparams<GPB_PERIODIC>(&sP, a.P, a.Pb, blockIdx.z);
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    const double* vec = a.vec[0] + (long long)blockIdx.z * a.vstride;
    double* out = a.out[0] + (long long)blockIdx.z * a.ostride;
    for (long long g0 = (long long)blockIdx.x * 16; g0 < a.n1; g0 += (long long)gridDim.x * 16) {
        const long long r0 = g0 + 2 * wid;
        const bool va = r0 < a.n1, vb = r0 + 1 < a.n1;
        double sa = 0.0, ca = 1.0, sb = 0.0, cb = 1.0;
        if (va) sincos(a.x1[r0] * sP.half_ip, &sa, &ca);
        if (vb) sincos(a.x1[r0 + 1] * sP.half_ip, &sb, &cb);
        double acc_a = 0.0, acc_b = 0.0;
        for (long long c0 = 0; c0 < a.n2; c0 += MEAN_CH) {
            const int nc = (int)((a.n2 - c0 < MEAN_CH) ? a.n2 - c0 : MEAN_CH);
            __syncthreads();
            for (int e = threadIdx.x; e < MEAN_CH; e += 256) {
                double sv = 0.0, cv = 1.0, vv = 0.0;
                if (e < nc) {
                    sincos(a.x2[c0 + e] * sP.half_ip, &sv, &cv);
                    vv = vec[c0 + e];
                }
                s_c[e] = sv; c_c[e] = cv; v_c[e] = vv;       // beyond the chunk: weight 0
            }
            __syncthreads();
            // each lane: 4 columns per step for both rows -> 8 independent exponentials
            for (int e = 4 * lane; e < nc; e += 128) {
                double arg[8], ev[8];
#pragma unroll
                for (int q = 0; q < 4; q++) {
                    const double sj = s_c[e + q], cj = c_c[e + q];
                    const double S0 = sin_diff(sa, ca, sj, cj), S1 = sin_diff(sb, cb, sj, cj);
                    arg[q] = sP.c1 * (S0 * S0);
                    arg[4 + q] = sP.c1 * (S1 * S1);
                }
                gpb_expv<8>(arg, ev);
#pragma unroll
                for (int q = 0; q < 4; q++) {
                    acc_a = fma(ev[q], v_c[e + q], acc_a);
                    acc_b = fma(ev[4 + q], v_c[e + q], acc_b);
                }
            }
        }
        acc_a = warp_sum(acc_a) * sP.k0;
        acc_b = warp_sum(acc_b) * sP.k0;
        if (lane == 0) {
            if (va) out[r0] = acc_a;
            if (vb) out[r0 + 1] = acc_b;
        }
    }
}

template <int KIND>
__global__ void __launch_bounds__(256) mean_kernel(const MatvecArgs a) {
    __shared__ KParams sP;
    load_kparams<KIND>(&sP, a.P, a.Pb, blockIdx.z);
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    const double* vec = a.vec[0] + (long long)blockIdx.z * a.vstride;
    double* out = a.out[0] + (long long)blockIdx.z * a.ostride;
    const bool vec_ok = ((reinterpret_cast<uintptr_t>(a.x2) | reinterpret_cast<uintptr_t>(vec)) & 15) == 0;
    const long long n2v = vec_ok ? (a.n2 & ~1LL) : 0;
    for (long long r0 = ((long long)blockIdx.x * 8 + wid) * 2; r0 < a.n1; r0 += (long long)gridDim.x * 16) {
        const bool two = r0 + 1 < a.n1;
        const double xa = a.x1[r0], xb = two ? a.x1[r0 + 1] : xa;
        double acc_a = 0.0, acc_b = 0.0;
        for (long long c = 2 * lane; c < n2v; c += 64) {
            const double2 xc = *reinterpret_cast<const double2*>(a.x2 + c);
            const double2 vc = *reinterpret_cast<const double2*>(vec + c);
            double u[4][10];
            const double dd[4] = {xa - xc.x, xa - xc.y, xb - xc.x, xb - xc.y};
            gpb_eval_unique_v<KIND, 4>(sP, dd, 1u, u);
            acc_a = fma(u[0][0], vc.x, acc_a);
            acc_a = fma(u[1][0], vc.y, acc_a);
            acc_b = fma(u[2][0], vc.x, acc_b);
            acc_b = fma(u[3][0], vc.y, acc_b);
        }
        for (long long c = n2v + lane; c < a.n2; c += 32) {
            double u[10];
            const double xc = a.x2[c], vc = vec[c];
            gpb_eval_unique<KIND>(sP, xa - xc, 1u, u); acc_a = fma(u[0], vc, acc_a);
            gpb_eval_unique<KIND>(sP, xb - xc, 1u, u); acc_b = fma(u[0], vc, acc_b);
        }
        acc_a = warp_sum(acc_a);
        acc_b = warp_sum(acc_b);
        if (lane == 0) {
            out[r0] = acc_a;
            if (two) out[r0 + 1] = acc_b;
        }
    }
}

int gpb_launch_fused_matvec(int kind, const KParams* P, const KParams* Pb, int batch,
                            const double* x1, long long n1, const double* x2, long long n2,
                            int npairs, const int* slice, const int* outidx, const double* coef,
                            const double* const* vec, int nout, double* const* out,
                            long long vstride, long long ostride, cudaStream_t st) {
    GPB_REQUIRE(kind == GPB_GAUSSIAN || kind == GPB_PERIODIC, "unknown kernel kind");
    GPB_REQUIRE(npairs >= 1 && npairs <= GPB_MV_MAXP && nout >= 1 && nout <= GPB_MV_MAXO, "bad pair count");
    if (n1 == 0) return GPB_OK;
    MatvecArgs a;
    if (P) a.P = *P; else memset(&a.P, 0, sizeof(KParams));
    a.Pb = Pb; a.x1 = x1; a.x2 = x2; a.n1 = n1; a.n2 = n2; a.npairs = npairs; a.nout = nout;
    for (int p = 0; p < GPB_MV_MAXP; p++) {
        a.slice[p] = p < npairs ? slice[p] : 0;
        a.outidx[p] = p < npairs ? outidx[p] : 0;
        a.coef[p] = p < npairs ? coef[p] : 0.0;
        a.vec[p] = p < npairs ? vec[p] : nullptr;
        if (p < npairs) GPB_REQUIRE(slice[p] >= 0 && slice[p] < gpb_n_slices(kind) && outidx[p] >= 0 && outidx[p] < nout, "bad pair");
    }
    for (int g = 0; g < GPB_MV_MAXO; g++) a.out[g] = g < nout ? out[g] : nullptr;
    a.vstride = vstride; a.ostride = ostride;
    long long nb = (n1 + 7) / 8;
    if (nb > 148 * 16) nb = 148 * 16;
    dim3 grid((unsigned)nb, 1, (unsigned)batch);
    GpbProfScope prof(GPB_KC_BUILD, st);
    if (npairs == 1 && nout == 1 && slice[0] == 0 && coef[0] == 1.0) {
        long long nbm = (n1 + 15) / 16;
        if (nbm > 148 * 16) nbm = 148 * 16;
        const dim3 gm((unsigned)nbm, 1, (unsigned)batch);
        if (kind == GPB_GAUSSIAN) mean_kernel<GPB_GAUSSIAN><<<gm, 256, 0, st>>>(a);
        else mean_periodic_kernel<<<gm, 256, 0, st>>>(a);
    } else if (kind == GPB_GAUSSIAN) fused_matvec_kernel<GPB_GAUSSIAN><<<grid, 256, 0, st>>>(a);
    else fused_matvec_kernel<GPB_PERIODIC><<<grid, 256, 0, st>>>(a);
    GPB_LAUNCH_CHECK("fused_matvec_kernel");
    return GPB_OK;
}

// ---------------------------------------------------------------------------
// 3. fused slice reduction over K^-1 (gp_c.pyx:34-49 without the (n_p+1) N^3 GEMMs,
//    and the trace / quadratic-form terms of gp_c.pyx:104-110):
//    for every requested slice S_q (q < nsl <= 6, any jacobian or hessian slice):
//        t0[q] = a^T S_q a ,   t1[q] = sum(Ki o S_q)
//    plus tr(Ki) and a.a for the noise rows (dK_s = 2 s I).  Slice tiles are
//    regenerated from x on the fly: one pass over Ki is the only HBM traffic.
//    Output layout (16 doubles): t0[0..5], t1[0..5], tr(Ki), a.a, 0, 0
// ---------------------------------------------------------------------------
struct GradArgs {
    KParams P;
    const KParams* Pb;
    const double* x;
    long long n;            // valid extent (pad rows/cols are excluded)
    const double* Ki;
    long long ldk, kstride;
    const double* alpha;
    long long astride;
    double* partial;        // [batch][gridDim.x][16]
    int nsl;
    int uq[GPB_RED_MAXS];   // unique slice ids
};

template <int KIND>
__global__ void __launch_bounds__(256) grad_reduce_kernel(const GradArgs a) {
    __shared__ KParams sP;
    __shared__ double red[32];
    load_kparams<KIND>(&sP, a.P, a.Pb, blockIdx.z);
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    const double* Ki = a.Ki + (long long)blockIdx.z * a.kstride;
    const double* al = a.alpha + (long long)blockIdx.z * a.astride;
    unsigned need = 0;
    for (int q = 0; q < a.nsl; q++) need |= 1u << a.uq[q];

    // Ki and every slice S are symmetric (x1 == x2): visit only c <= r and weight the strict
    // lower part twice -- half the exp / sincos evaluations and half the Ki traffic.
    double t0[GPB_RED_MAXS], t1[GPB_RED_MAXS], tr = 0, aa = 0;
#pragma unroll
    for (int q = 0; q < GPB_RED_MAXS; q++) t0[q] = t1[q] = 0.0;
    for (long long r = (long long)blockIdx.x * 8 + wid; r < a.n; r += (long long)gridDim.x * 8) {
        const double xi = a.x[r], ar = al[r];
        const double* row = Ki + r * a.ldk;
        double q0[GPB_RED_MAXS];
#pragma unroll
        for (int q = 0; q < GPB_RED_MAXS; q++) q0[q] = 0.0;
        for (long long c = lane; c <= r; c += 32) {
            double u[10];
            gpb_eval_unique<KIND>(sP, xi - a.x[c], need, u);
            const double wgt = (c < r) ? 2.0 : 1.0;
            const double k = wgt * row[c], ac = wgt * al[c];
#pragma unroll
            for (int q = 0; q < GPB_RED_MAXS; q++) {
                if (q < a.nsl) {
                    const double v = gpb_pick(u, a.uq[q]);
                    q0[q] += v * ac;
                    t1[q] += v * k;
                }
            }
            if (c == r) tr += row[c];
        }
#pragma unroll
        for (int q = 0; q < GPB_RED_MAXS; q++) t0[q] += ar * q0[q];
        if (lane == 0) aa += ar * ar;
    }
    double* out = a.partial + ((long long)blockIdx.z * gridDim.x + blockIdx.x) * GPB_RED_WIDTH;
#pragma unroll
    for (int q = 0; q < GPB_RED_MAXS; q++) {
        const double s0 = block_sum(t0[q], red);
        const double s1 = block_sum(t1[q], red);
        if (threadIdx.x == 0) { out[q] = s0; out[GPB_RED_MAXS + q] = s1; }
    }
    {
        const double s0 = block_sum(tr, red);
        const double s1 = block_sum(aa, red);
        if (threadIdx.x == 0) { out[12] = s0; out[13] = s1; out[14] = 0.0; out[15] = 0.0; }
    }
}

// Gradient fast path (the per-candidate reduction of the batched evaluator): the requested
// slices are exactly the Jacobian slices in order, so the slice set is a compile-time constant
// (no run-time picks), two adjacent columns per lane with 128-bit loads, and a register budget
// that lets three CTAs share an SM.
// ncu (B=8, N=4096): 46 us per candidate, FP64 pipe 34 % active, top stall long_scoreboard (latency of
// the Ki / x / alpha loads at 24 warps per SM).  A four-columns-per-lane variant (six 128-bit loads in
// flight, 114 registers, two CTAs per SM) was built and measured: -14 % at N=4096 (1.38 -> 1.18 ms per
// 32 candidates, 0.3 % of the step) but +15 % at N=1024 where rows are short -- not kept.
template <int KIND>
__global__ void __launch_bounds__(256, 3) grad_jac_kernel(const GradArgs a) {
    constexpr int NP = (KIND == GPB_GAUSSIAN) ? 2 : 3;
    constexpr unsigned NEED = (KIND == GPB_GAUSSIAN) ? 0x6u : 0xEu;
    __shared__ KParams sP;
    __shared__ double red[32];
    load_kparams<KIND>(&sP, a.P, a.Pb, blockIdx.z);
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    const double* Ki = a.Ki + (long long)blockIdx.z * a.kstride;
    const double* al = a.alpha + (long long)blockIdx.z * a.astride;
    const bool vec_ok = ((reinterpret_cast<uintptr_t>(a.x) | reinterpret_cast<uintptr_t>(al) |
                          reinterpret_cast<uintptr_t>(Ki)) & 15) == 0 && (a.ldk & 1) == 0;

    double t0[NP], t1[NP], tr = 0, aa = 0;
#pragma unroll
    for (int q = 0; q < NP; q++) t0[q] = t1[q] = 0.0;
    for (long long r = (long long)blockIdx.x * 8 + wid; r < a.n; r += (long long)gridDim.x * 8) {
        const double xi = a.x[r], ar = al[r];
        const double* row = Ki + r * a.ldk;
        double q0[NP];
#pragma unroll
        for (int q = 0; q < NP; q++) q0[q] = 0.0;
        for (long long c = 2 * lane; c <= r; c += 64) {
            double xc0, xc1 = 0.0, k0, k1 = 0.0, a0, a1 = 0.0;
            const bool two = (c + 1 <= r);
            if (two && vec_ok) {
                const double2 xv = *reinterpret_cast<const double2*>(a.x + c);
                const double2 kv = *reinterpret_cast<const double2*>(row + c);
                const double2 av = *reinterpret_cast<const double2*>(al + c);
                xc0 = xv.x; xc1 = xv.y; k0 = kv.x; k1 = kv.y; a0 = av.x; a1 = av.y;
            } else {
                xc0 = a.x[c]; k0 = row[c]; a0 = al[c];
                if (two) { xc1 = a.x[c + 1]; k1 = row[c + 1]; a1 = al[c + 1]; }
            }
            double u0[10], u1[10];
            gpb_eval_unique<KIND>(sP, xi - xc0, NEED, u0);
            // strict lower part counts twice (symmetry), the diagonal once
            const double w0 = (c < r) ? 2.0 : 1.0;
            if (c == r) tr += k0;
            k0 *= w0; a0 *= w0;
#pragma unroll
            for (int q = 0; q < NP; q++) { q0[q] += u0[1 + q] * a0; t1[q] += u0[1 + q] * k0; }
            if (two) {
                gpb_eval_unique<KIND>(sP, xi - xc1, NEED, u1);
                const double w1 = (c + 1 < r) ? 2.0 : 1.0;
                if (c + 1 == r) tr += k1;
                k1 *= w1; a1 *= w1;
#pragma unroll
                for (int q = 0; q < NP; q++) { q0[q] += u1[1 + q] * a1; t1[q] += u1[1 + q] * k1; }
            }
        }
#pragma unroll
        for (int q = 0; q < NP; q++) t0[q] += ar * q0[q];
        if (lane == 0) aa += ar * ar;
    }
    double* out = a.partial + ((long long)blockIdx.z * gridDim.x + blockIdx.x) * GPB_RED_WIDTH;
#pragma unroll
    for (int q = 0; q < GPB_RED_MAXS; q++) {
        // only the NP live slots are reduced (each block_sum is two barriers + two shuffle trees)
        const double s0 = (q < NP) ? block_sum(t0[q < NP ? q : 0], red) : 0.0;
        const double s1 = (q < NP) ? block_sum(t1[q < NP ? q : 0], red) : 0.0;
        if (threadIdx.x == 0) { out[q] = s0; out[GPB_RED_MAXS + q] = s1; }
    }
    {
        const double s0 = block_sum(tr, red);
        const double s1 = block_sum(aa, red);
        if (threadIdx.x == 0) { out[12] = s0; out[13] = s1; out[14] = 0.0; out[15] = 0.0; }
    }
}

// deterministic second stage: out[b][q] = sum_k partial[b][k][q]
__global__ void __launch_bounds__(256) sum_partials_kernel(const double* partial, int nblk, int width, double* out) {
    __shared__ double red[32];
    const double* p = partial + (long long)blockIdx.x * nblk * width;
    for (int q = 0; q < width; q++) {
        double s = 0;
        for (int k = threadIdx.x; k < nblk; k += blockDim.x) s += p[(long long)k * width + q];
        s = block_sum(s, red);
        if (threadIdx.x == 0) out[(long long)blockIdx.x * width + q] = s;
    }
}

// CTAs of one matrix's reduction pass = rows of `partial` to provide (capacity; a batched launch uses fewer)
int gpb_grad_reduce_blocks(long long n) {
    long long nb = (n + 7) / 8;
    if (nb > 148 * 4) nb = 148 * 4;
    if (nb < 1) nb = 1;
    // the fused lauum epilogue writes one row per CTA of its lower-triangular 64 x 128 tile grid
    const long long tn = (n + GPB_NB - 1) / GPB_NB;
    if (tn * (tn + 1) > nb) nb = tn * (tn + 1);
    return (int)nb;
}

int gpb_launch_sum_partials(const double* partial, int nblk, int batch, double* out16, cudaStream_t st) {
    sum_partials_kernel<<<batch, 256, 0, st>>>(partial, nblk, GPB_RED_WIDTH, out16);
    GPB_LAUNCH_CHECK("sum_partials_kernel");
    return GPB_OK;
}
static int grad_reduce_blocks_used(long long n, bool one_object) {
    // batches: four rows per warp (strided, so the triangular row lengths balance): with one row per warp the
    // block-wide reduction at the end of every CTA cost as much as its 8 rows at N = 1024.
    // One GP object (parameters by value; the batched evaluator passes a parameter array even for one candidate, so
    // a candidate's sums never depend on how the batch was split): one row per warp -- 128 CTAs left the GPU at 7
    // warps per SM (136 us at N = 4096, 0.5 TB/s).
    if (one_object) {
        long long nb1 = (n + 7) / 8;
        if (nb1 > 148 * 4) nb1 = 148 * 4;
        return (int)(nb1 < 1 ? 1 : nb1);
    }
    long long nb = (n + 31) / 32;
    if (nb > 148 * 4) nb = 148 * 4;
    if (nb < 1) nb = 1;
    return (int)nb;
}

// slices: slice ids (NOT unique ids); nsl <= 6
int gpb_launch_grad_reduce(int kind, const KParams* P, const KParams* Pb, int batch, const double* x,
                           long long n, const double* Ki, long long ldk, long long kstride,
                           const double* alpha, long long astride, int nsl, const int* slices,
                           double* partial, double* out16, cudaStream_t st) {
    GPB_REQUIRE(kind == GPB_GAUSSIAN || kind == GPB_PERIODIC, "unknown kernel kind");
    GPB_REQUIRE(nsl >= 0 && nsl <= GPB_RED_MAXS, "at most 6 slices per pass");
    GradArgs a;
    if (P) a.P = *P; else memset(&a.P, 0, sizeof(KParams));
    a.Pb = Pb; a.x = x; a.n = n; a.Ki = Ki; a.ldk = ldk; a.kstride = kstride;
    a.alpha = alpha; a.astride = astride; a.partial = partial; a.nsl = nsl;
    for (int q = 0; q < GPB_RED_MAXS; q++) {
        a.uq[q] = 0;
        if (q < nsl) {
            GPB_REQUIRE(slices[q] >= 0 && slices[q] < gpb_n_slices(kind), "bad slice id");
            a.uq[q] = gpb_slice_to_unique(kind, slices[q]);
        }
    }
    const int nb = grad_reduce_blocks_used(n, Pb == nullptr && batch == 1);
    dim3 grid((unsigned)nb, 1, (unsigned)batch);
    GpbProfScope prof(GPB_KC_REDUCE, st);
    bool jac = (nsl == gpb_n_kparams(kind));
    for (int q = 0; q < nsl; q++) jac = jac && (a.uq[q] == q + 1);
    if (jac) {
        if (kind == GPB_GAUSSIAN) grad_jac_kernel<GPB_GAUSSIAN><<<grid, 256, 0, st>>>(a);
        else grad_jac_kernel<GPB_PERIODIC><<<grid, 256, 0, st>>>(a);
    } else {
        if (kind == GPB_GAUSSIAN) grad_reduce_kernel<GPB_GAUSSIAN><<<grid, 256, 0, st>>>(a);
        else grad_reduce_kernel<GPB_PERIODIC><<<grid, 256, 0, st>>>(a);
    }
    GPB_LAUNCH_CHECK("grad_reduce_kernel");
    sum_partials_kernel<<<batch, 256, 0, st>>>(partial, nb, GPB_RED_WIDTH, out16);
    GPB_LAUNCH_CHECK("sum_partials_kernel");
    return GPB_OK;
}

// ---------------------------------------------------------------------------
// 4. predictive variance: var[r] = K(xo_r, xo_r) - sum_c Z[r][c]^2 with Z = K(xo, x) L^-T, i.e.
//    diag(GP.cov) (gp.py:625, what GP.plot needs, gp.py:692-693) without the M x M matrix.
// ---------------------------------------------------------------------------
template <int KIND>
__global__ void __launch_bounds__(256) post_var_kernel(const KParams P, const double* Z, long long ldz, long long m,
                                                       long long n, double* out) {
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    double u[10];
    gpb_eval_unique<KIND>(P, 0.0, 1u, u);        // k(x*, x*): both kernels are stationary
    const double k0 = u[0];
    for (long long r = (long long)blockIdx.x * 8 + wid; r < m; r += (long long)gridDim.x * 8) {
        const double* zr = Z + r * ldz;
        double s0 = 0.0, s1 = 0.0;
        const long long nv = ((reinterpret_cast<uintptr_t>(zr) & 15) == 0) ? (n & ~1LL) : 0;
        for (long long c = 2 * lane; c < nv; c += 64) {
            const double2 v = *reinterpret_cast<const double2*>(zr + c);
            s0 = fma(v.x, v.x, s0);
            s1 = fma(v.y, v.y, s1);
        }
        for (long long c = nv + lane; c < n; c += 32) s0 = fma(zr[c], zr[c], s0);
        const double s = warp_sum(s0 + s1);
        if (lane == 0) out[r] = k0 - s;
    }
}

int gpb_launch_post_var(int kind, const KParams* P, const double* Z, long long ldz, long long m, long long n,
                        double* out, cudaStream_t st) {
    GPB_REQUIRE(kind == GPB_GAUSSIAN || kind == GPB_PERIODIC, "unknown kernel kind");
    GPB_REQUIRE(Z && out && ldz >= n, "bad argument");
    if (m == 0) return GPB_OK;
    long long nb = (m + 7) / 8;
    if (nb > 148 * 8) nb = 148 * 8;
    GpbProfScope prof(GPB_KC_REDUCE, st);
    if (kind == GPB_GAUSSIAN) post_var_kernel<GPB_GAUSSIAN><<<(unsigned)nb, 256, 0, st>>>(*P, Z, ldz, m, n, out);
    else post_var_kernel<GPB_PERIODIC><<<(unsigned)nb, 256, 0, st>>>(*P, Z, ldz, m, n, out);
    GPB_LAUNCH_CHECK("post_var_kernel");
    return GPB_OK;
}
