// kfunctors.cuh -- element formulas of the Gaussian and Periodic kernels and all
// their first/second parameter derivatives, evaluated from ONE exp (Gaussian) or
// ONE exp + ONE sincos (Periodic) per (x1[i], x2[j]) pair.
//
// Reference formulas: gp/ext/gaussian_c.pyx:18-164, gp/ext/periodic_c.pyx:18-235
// (SURVEY.md Appendix A).  The reference re-evaluates exp/sin/cos for every term
// of every slice; here every slice shares them.
//
// Slice ids (the order the reference's jacobian()/hessian() stack them):
//   Gaussian: 0 K | 1 dh 2 dw | 3 hh 4 hw 5 wh 6 ww
//   Periodic: 0 K | 1 dh 2 dw 3 dp | 4 hh 5 hw 6 hp 7 wh 8 ww 9 wp 10 ph 11 pw 12 pp
// "unique" ids drop the symmetric duplicates:
//   Gaussian: 0 K 1 h 2 w 3 hh 4 hw 5 ww
//   Periodic: 0 K 1 h 2 w 3 p 4 hh 5 hw 6 hp 7 ww 8 wp 9 pp
#pragma once
#include "common.cuh"

#define GPB_GAUSSIAN 0
#define GPB_PERIODIC 1

__host__ __device__ __forceinline__ int gpb_n_kparams(int kind) { return kind == GPB_GAUSSIAN ? 2 : 3; }
__host__ __device__ __forceinline__ int gpb_n_slices(int kind) { return kind == GPB_GAUSSIAN ? 7 : 13; }

__host__ __device__ constexpr int gpb_slice_to_unique(int kind, int s) {
    return kind == GPB_GAUSSIAN
               ? (s <= 4 ? s : s - 1)
               : (s <= 6 ? s : s == 7 ? 5 : s == 8 ? 7 : s == 9 ? 8 : s == 10 ? 6 : s == 11 ? 8 : 9);
}

// host: fill KParams from (h, w[, p], s).  Coefficient expressions keep the
// reference's operation order (e.g. gaussian_c.pyx:83-84: 0.5*S*h2/w**4).
inline void gpb_make_kparams(KParams* P, int kind, const double* theta, double s) {
    const double S = sqrt(2.0 / M_PI);
    const double h = theta[0], w = theta[1];
    const double h2 = h * h, w2 = w * w;
    P->kind = kind;
    P->pad_ = 0;
    P->s2 = s * s;
    P->s = s;
    for (int i = 0; i < 3; i++) P->j[i][0] = P->j[i][1] = 0;
    for (int i = 0; i < 6; i++)
        for (int k = 0; k < 4; k++) P->h[i][k] = 0;
    if (kind == GPB_GAUSSIAN) {
        P->c1 = -0.5 / w2;
        P->half_ip = 0;
        P->k0 = 0.5 * S * h2 / w;                       // gaussian_c.pyx:27
        P->j[0][0] = S * h / w;                         // :61
        P->j[1][0] = 0.5 * S * h2 / pow(w, 4);          // :84 (c3)
        P->j[1][1] = 0.5 * S * h2 / w2;                 // :83 (c2)
        P->h[0][0] = S / w;                             // :105
        P->h[1][0] = S * h / pow(w, 4);                 // :128 (c3)
        P->h[1][1] = S * h / w2;                        // :127 (c2)
        P->h[2][0] = 0.5 * S * h2 / pow(w, 7);          // :156 (c4)
        P->h[2][1] = 2.5 * S * h2 / pow(w, 5);          // :155 (c3)
        P->h[2][2] = S * h2 / pow(w, 3);                // :154 (c2)
    } else {
        const double p = theta[2], p2 = p * p;
        P->c1 = -2.0 / w2;
        P->half_ip = 0.5 / p;
        P->k0 = h2;                                     // periodic_c.pyx:30
        P->j[0][0] = 2.0 * h;                           // :65
        P->j[1][0] = 4.0 * h2 / pow(w, 3);              // :80
        P->j[2][0] = 2.0 * h2 / (p2 * w2);              // :96
        P->h[0][0] = 2.0;                               // :111
        P->h[1][0] = 8.0 * h / pow(w, 3);               // :126
        P->h[2][0] = 4.0 * h / (p2 * w2);               // :142
        P->h[3][0] = -12.0 * h2 / pow(w, 4);            // :172
        P->h[3][1] = 16.0 * h2 / pow(w, 6);
        P->h[4][0] = -4.0 * h2 / (p2 * pow(w, 3));      // :188
        P->h[4][1] = 8.0 * h2 / (p2 * pow(w, 5));
        P->h[5][0] = h2 / (pow(p, 4) * w2);             // :235
        P->h[5][1] = 4.0 / w2;
        P->h[5][2] = 4.0 * h2 / (pow(p, 3) * w2);
    }
}

#ifdef __CUDACC__

// exp for the kernel functors.  libdevice's exp() rebuilds its 11 polynomial coefficients from
// 64-bit immediates on every call (two MOV/UMOV per coefficient: cuobjdump of the previous build showed
// ~115 issued instructions per kernel-matrix element of which 24 were FP64), so the element kernels were
// bound by instruction issue, not by the FP64 pipe.  Here the coefficients are constant-bank operands of the
// DFMAs themselves: k = rint(x log2 e) by the 1.5*2^52 shift, two-term Cody-Waite reduction
// r = x - k ln2 (|r| <= 0.3466), degree-13 Taylor/Horner (truncation 4e-18 relative), 2^k added straight
// into the exponent field.  Valid while the result is a normal number; arguments outside (-700, 700)
// (or NaN) take the library routine.  Max error measured against libm over the Gaussian / periodic
// argument ranges: < 1 ulp (tests/test_parity_gpu.py keeps the 1e-13 builder tolerance).
static __constant__ double gpb_exp_c[16] = {
    1.0 / 6227020800.0, 1.0 / 479001600.0, 1.0 / 39916800.0, 1.0 / 3628800.0, 1.0 / 362880.0, 1.0 / 40320.0,
    1.0 / 5040.0, 1.0 / 720.0, 1.0 / 120.0, 1.0 / 24.0, 1.0 / 6.0, 0.5, 1.0, 1.0,
    1.4426950408889634074, 6755399441055744.0};
static __constant__ double gpb_exp_ln2[2] = {-6.93147180559945286227e-01, -2.31904681384629955842e-17};

// (The same treatment of sincos for the periodic functor -- Cody-Waite + minimax kernels from a
// constant table -- was built and measured: no change, N=8192 K-only build 0.297 ms either way; the
// periodic element kernels already run ~50 FP64 operations per element at ~60 % of the FP64 pipe.)
// W independent arguments at once: the Horner steps run coefficient-major, so each coefficient is
// fetched once per W DFMAs (the element kernels evaluate 2-8 independent separations per thread).
template <int W>
__device__ __forceinline__ void gpb_expv(const double (&x)[W], double (&e)[W]) {
    double r[W], p[W];
    int k[W];
    bool slow = false;
#pragma unroll
    for (int w = 0; w < W; w++) {
        slow |= !(fabs(x[w]) < 700.0);
        const double t = fma(x[w], gpb_exp_c[14], gpb_exp_c[15]);
        k[w] = __double2loint(t);
        const double kf = t - gpb_exp_c[15];
        r[w] = fma(kf, gpb_exp_ln2[1], fma(kf, gpb_exp_ln2[0], x[w]));
        p[w] = gpb_exp_c[0];
    }
#pragma unroll
    for (int i = 1; i < 14; i++) {
        const double c = gpb_exp_c[i];
#pragma unroll
        for (int w = 0; w < W; w++) p[w] = fma(p[w], r[w], c);
    }
#pragma unroll
    for (int w = 0; w < W; w++) e[w] = __hiloint2double(__double2hiint(p[w]) + (k[w] << 20), __double2loint(p[w]));
    if (slow) {                      // rare: result not a normal number (or NaN / inf argument)
#pragma unroll
        for (int w = 0; w < W; w++)
            if (!(fabs(x[w]) < 700.0)) e[w] = exp(x[w]);
    }
}

__device__ __forceinline__ double gpb_exp(double x) {
    const double xv[1] = {x};
    double ev[1];
    gpb_expv<1>(xv, ev);
    return ev[0];
}

// u[idx] for a run-time idx without spilling u[] to local memory
__device__ __forceinline__ double gpb_pick(const double* u, int idx) {
    double v = 0.0;
#pragma unroll
    for (int q = 0; q < 10; q++)
        if (idx == q) v = u[q];
    return v;
}

// Periodic slices from S = sin(d / 2p), C = cos(d / 2p) (periodic_c.pyx:18-235 with one exp per element).
// The builders and the fused mean obtain S, C from per-point tables, S = s_i c_j - c_i s_j,
// C = c_i c_j + s_i s_j with (s_i, c_i) = sincos(x_i / 2p) staged in shared memory: 4 FMAs per element
// instead of a sincos (~45 FP64 operations).  The two forms differ by rounding in the argument only,
// |dS| ~ 3e-16 for |x / 2p| <= pi: far inside the 1e-13 normwise tolerance of the builders.
__device__ __forceinline__ void gpb_eval_periodic_sc(const KParams& P, double d, double S, double C, unsigned need, double* u) {
    const double S2 = S * S;
    const double E = gpb_exp(P.c1 * S2);
    if (need & 1u) u[0] = P.k0 * E;
    if (need & 2u) u[1] = P.j[0][0] * E;
    const double ES2 = E * S2;
    if (need & 4u) u[2] = P.j[1][0] * ES2;
    const double dESC = d * E * S * C;
    if (need & 8u) u[3] = P.j[2][0] * dESC;
    if (need & 16u) u[4] = P.h[0][0] * E;
    if (need & 32u) u[5] = P.h[1][0] * ES2;
    if (need & 64u) u[6] = P.h[2][0] * dESC;
    if (need & 128u) u[7] = ES2 * (P.h[3][0] + P.h[3][1] * S2);
    if (need & 256u) u[8] = dESC * (P.h[4][0] + P.h[4][1] * S2);
    if (need & 512u) {
        const double C2 = C * C;
        u[9] = P.h[5][0] * (d * d) * E * (S2 - C2 + P.h[5][1] * S2 * C2) - P.h[5][2] * dESC;
    }
}

// Evaluate the unique slice values selected by `need` (bit u set = wanted) at
// separation d = x1[i] - x2[j].  Unselected entries of u[] are left untouched.
template <int KIND>
__device__ __forceinline__ void gpb_eval_unique(const KParams& P, double d, unsigned need, double* u) {
    if (KIND == GPB_GAUSSIAN) {
        const double d2 = d * d;
        const double e = P.c1 * d2;
        // entries whose exponent is below MIN are exactly 0 in every slice (gaussian_c.pyx:33-34)
        const double ex = (e < GPB_MIN_LOG) ? 0.0 : gpb_exp(e);
        if (need & 1u) u[0] = P.k0 * ex;
        if (need & 2u) u[1] = P.j[0][0] * ex;
        if (need & 4u) u[2] = ex * (P.j[1][0] * d2 - P.j[1][1]);
        if (need & 8u) u[3] = P.h[0][0] * ex;
        if (need & 16u) u[4] = ex * (P.h[1][0] * d2 - P.h[1][1]);
        if (need & 32u) u[5] = ex * (P.h[2][0] * (d2 * d2) - P.h[2][1] * d2 + P.h[2][2]);
    } else {
        double S, C;
        sincos(d * P.half_ip, &S, &C);
        gpb_eval_periodic_sc(P, d, S, C, need, u);
    }
}

// W separations at once (same arithmetic per element as gpb_eval_unique, shared coefficient fetches)
template <int KIND, int W>
__device__ __forceinline__ void gpb_eval_unique_v(const KParams& P, const double (&d)[W], unsigned need,
                                                  double (&u)[W][10]) {
    if (KIND == GPB_GAUSSIAN) {
        double e[W], ex[W], d2[W];
#pragma unroll
        for (int w = 0; w < W; w++) {
            d2[w] = d[w] * d[w];
            const double ee = P.c1 * d2[w];
            e[w] = (ee < GPB_MIN_LOG) ? 0.0 : ee;            // placeholder argument; zeroed below
            ex[w] = ee;
        }
        double ev[W];
        gpb_expv<W>(e, ev);
#pragma unroll
        for (int w = 0; w < W; w++) {
            // entries whose exponent is below MIN are exactly 0 in every slice (gaussian_c.pyx:33-34)
            const double x_ = (ex[w] < GPB_MIN_LOG) ? 0.0 : ev[w];
            if (need & 1u) u[w][0] = P.k0 * x_;
            if (need & 2u) u[w][1] = P.j[0][0] * x_;
            if (need & 4u) u[w][2] = x_ * (P.j[1][0] * d2[w] - P.j[1][1]);
            if (need & 8u) u[w][3] = P.h[0][0] * x_;
            if (need & 16u) u[w][4] = x_ * (P.h[1][0] * d2[w] - P.h[1][1]);
            if (need & 32u) u[w][5] = x_ * (P.h[2][0] * (d2[w] * d2[w]) - P.h[2][1] * d2[w] + P.h[2][2]);
        }
    } else {
#pragma unroll
        for (int w = 0; w < W; w++) gpb_eval_unique<KIND>(P, d[w], need, u[w]);
    }
}

#endif  // __CUDACC__
