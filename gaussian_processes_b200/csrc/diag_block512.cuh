// diag_block512.cuh -- the diagonal block of the dataflow factorisation's chain CTA: 128 x 128 Cholesky +
// inverse by 512 threads with look-ahead inside the block.
//
// The 256-thread body (diag_block.cuh) runs its phases one after the other: per 32-column step the 32x32
// sweep (one warp, 8.5 k cycles: 32 dependent pivot -> rsqrt -> scale -> shuffle links), then the inverse
// of the 32x32 factor and the substitution of the rows below (2.9 k), then the in-tile DMMA update
// (3.5-9 k), and at the end the 128-level inverse (11.5 k): 78 k cycles of which only the sweeps are
// inherently serial.  In the chain CTA that time is the critical path of the whole factorisation, so
// here everything except the sweeps is taken off it:
//   S(bb)  warp 0 sweeps the 32x32 diagonal sub-block and publishes each finished column (shared memory
//          + a progress counter); warps 1..3 take the rows below THROUGH THE SAME COLUMN STREAM
//          (right-looking substitution, lane = row, one column behind the sweep), so X = A L^-T is complete
//          ~100 cycles after the sweep instead of 2.9 k; meanwhile warp 5 inverts the previous 32x32
//          factor, warps 6, 7 stream the previous L column to global memory, warps 9-11, 13-15 finish the
//          part of the previous in-tile update the next sweep does not need (and, during the last sweep,
//          the first level of the 128-level inverse).  Warps 4, 8, 12 idle: they share the sweep's SM
//          sub-partition, and anything they issue delays its dependent chain;
//   C(bb)  only the block column the next sweep reads is updated between two sweeps (DMMA strips);
//   tail   last 32x32 inverse, then two short DMMA phases complete W = L^-1.
// Only full blocks (no identity pad) take this body; a padded last block uses the 256-thread one.
#pragma once
#include "diag_block.cuh"

namespace {

constexpr int LTLD = 34;                      // column-stream row stride (16-byte aligned pairs)
constexpr int NSCR = 7;                       // scratch sub-blocks of the 128-level inverse
constexpr int DIAG512_DOUBLES = (NBLK + 4 + NSCR) * SBSZ + SB * LTLD + 4 * SB + 2;
constexpr int DIAG512_SMEM = DIAG512_DOUBLES * 8;

// One 16 x 32 strip (rows hf*16..) of   C = Cin + sgn * (A1 B1 [+ A2 B2])   for 32 x 32 blocks in shared memory.
// EVERY DMMA phase of the body goes through ONE copy of this code (the body is a loop over phases with the
// strip section at its end): the chain CTA executes ~150 KB of straight-line code per step (sweep,
// substitution, inverse), so code that runs once per phase is cold every time it is reached -- measured 7 k
// cycles for a 64-DMMA strip whose arithmetic takes 1.1 k -- while one shared copy is fetched once per step.
//   A operands row-major, stride SLD.  B element (k, n) at B + k * bks + n * bns:
//     bks = 1, bns = SLD  -> B = (row-major block)^T   (the in-tile update's L_jb^T)
//     bks = SLD, bns = 1  -> B = row-major block       (products of the 128-level inverse)
struct D512Strip {
    const double *A1, *B1, *A2, *B2;      // A2 == nullptr: single product
    int bks, bns;
    double sgn;
    const double* Cin; int cin_s;          // nullptr: start from zero
    double* Cout; int cout_s;              // nullptr: no shared-memory result
    int wi, wj, hf;                        // wi >= 0: also block (wi, wj) of the global W and V = W^T
    long long* dbg;                        // optional: clock stamps of this strip (thread 0)
};

template <int KKU>
__device__ __forceinline__ void d512_strip(const D512Strip& s, double* W, long long ldw, double* V, long long ldv) {
    const int lane = threadIdx.x & 31, g = lane >> 2, t = lane & 3, hf = s.hf;
    if (s.dbg && threadIdx.x == 0) s.dbg[0] = clock64();
    double acc[2][4][2];
#pragma unroll
    for (int rt = 0; rt < 2; rt++)
#pragma unroll
        for (int ct = 0; ct < 4; ct++) {
            if (s.Cin) {
                const double* p = s.Cin + (hf * 16 + rt * 8 + g) * s.cin_s + ct * 8 + 2 * t;
                acc[rt][ct][0] = p[0];
                acc[rt][ct][1] = p[1];
            } else {
                acc[rt][ct][0] = acc[rt][ct][1] = 0.0;
            }
        }
    if (s.dbg && threadIdx.x == 0) s.dbg[1] = clock64();
    const double* Ap = s.A1 + (hf * 16 + g) * SLD + t;
    const double* Bp = s.B1 + t * s.bks + g * s.bns;
    const int nprod = s.A2 ? 2 : 1;
#pragma unroll 1
    for (int pr = 0; pr < nprod; pr++) {
        // (rolled: 20 instructions fetched once instead of 160 cold ones -- instruction fetch, not the DMMA
        //  pipe, bounds a strip that runs once per phase)
#pragma unroll KKU
        for (int kk = 0; kk < 8; kk++) {
            double a[2], b[4];
#pragma unroll
            for (int rt = 0; rt < 2; rt++) a[rt] = s.sgn * Ap[rt * 8 * SLD + kk * 4];
#pragma unroll
            for (int ct = 0; ct < 4; ct++) b[ct] = Bp[kk * 4 * s.bks + ct * 8 * s.bns];
#pragma unroll
            for (int rt = 0; rt < 2; rt++)
#pragma unroll
                for (int ct = 0; ct < 4; ct++) dmma884(acc[rt][ct][0], acc[rt][ct][1], a[rt], b[ct]);
        }
        if (s.A2) {
            Ap = s.A2 + (hf * 16 + g) * SLD + t;
            Bp = s.B2 + t * s.bks + g * s.bns;
        }
    }
    if (s.dbg && threadIdx.x == 0) s.dbg[2] = clock64();
    if (s.Cout) {
#pragma unroll
        for (int rt = 0; rt < 2; rt++)
#pragma unroll
            for (int ct = 0; ct < 4; ct++) {
                double* p = s.Cout + (hf * 16 + rt * 8 + g) * s.cout_s + ct * 8 + 2 * t;
                p[0] = acc[rt][ct][0];
                p[1] = acc[rt][ct][1];
            }
    }
    if (s.wi >= 0) strip_to_wv(acc, W, ldw, V, ldv, s.wi, s.wj, hf, g, t);
    if (s.dbg && threadIdx.x == 0) s.dbg[3] = clock64();
}

// in-tile update  C(bi,bj) -= L(bi,bb) L(bj,bb)^T, one strip
__device__ __forceinline__ D512Strip d512_update(double* Lb, int bi, int bj, int bb, int hf) {
    D512Strip s;
    s.A1 = Lb + blk(bi, bb) * SBSZ; s.B1 = Lb + blk(bj, bb) * SBSZ; s.A2 = s.B2 = nullptr;
    s.bks = 1; s.bns = SLD; s.sgn = -1.0;
    double* Cb = Lb + blk(bi, bj) * SBSZ;
    s.Cin = Cb; s.Cout = Cb; s.cin_s = s.cout_s = (bi == bj) ? DLD : SLD;
    s.wi = -1; s.wj = 0; s.hf = hf; s.dbg = nullptr;
    return s;
}

// strip of  sgn * (A1 B1 [+ A2 B2])  (row-major operands) -> shared block `out` and / or block (wi, wj) of W, V
__device__ __forceinline__ D512Strip d512_prod(const double* A1, const double* B1, const double* A2, const double* B2,
                                               double sgn, double* out, int wi, int wj, int hf) {
    D512Strip s;
    s.A1 = A1; s.B1 = B1; s.A2 = A2; s.B2 = B2;
    s.bks = SLD; s.bns = 1; s.sgn = sgn;
    s.Cin = nullptr; s.cin_s = 0; s.Cout = out; s.cout_s = SLD;
    s.wi = wi; s.wj = wj; s.hf = hf; s.dbg = nullptr;
    return s;
}

template <int KKU>
__device__ __forceinline__ void diag_block_body512(double* A, long long ld, double* W, long long ldw, double* V,
                                                   long long ldv, int* info, int col0, double* sm,
                                                   long long* dclk = nullptr, int* lpub = nullptr) {
    // lpub != nullptr (pipelined chain group, chain.cu): after each 32-column step the block column of L and the
    // inverted 32x32 diagonal sub-block go to global memory at once and lpub[bb] is released -- the helper CTAs run
    // the next tile's TRSM / SYRK one 32-column block behind the sweeps -- and the 128-level inverse (everything
    // after the last 32x32 inverse) is left to the group's inverter CTA.
#define D512_STAMP(i) do { if (dclk && threadIdx.x == 0) dclk[i] = clock64(); } while (0)
    D512_STAMP(0);
    double* Lb = sm;                          // 10 lower sub-blocks (off-diagonal: stride SLD, diagonal: DLD)
    double* Wd = sm + NBLK * SBSZ;            // inverses of the 4 diagonal sub-blocks (stride SLD)
    double* Xs = Wd + 4 * SBSZ;               // scratch blocks: S10 W10 S32 W32 S20 S21 S30 (S31 reuses S10)
    double* Lt = Xs + NSCR * SBSZ;            // column stream of the running sweep: Lt[k][row], stride LTLD
    double* invd = Lt + SB * LTLD;            // 1 / L_kk, 4 x 32
    volatile int* prog = reinterpret_cast<volatile int*>(invd + 4 * SB);     // columns the sweeps have published so far (32 bb + k)
    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;

    // ---- load the lower block-triangle ----------------------------------------------------------
    for (int bi = 0; bi < 4; bi++)
        for (int bj = 0; bj <= bi; bj++) {
            double* dst = Lb + blk(bi, bj) * SBSZ;
            if (bi == bj) {
                for (int e = tid; e < SB * SB / 2; e += 512) {
                    const int r = e >> 4, c2 = (e & 15) * 2;
                    const double2 v = __ldcg(reinterpret_cast<const double2*>(A + (long long)(bi * SB + r) * ld + bj * SB + c2));
                    dst[r * DLD + c2] = v.x;
                    dst[r * DLD + c2 + 1] = v.y;
                }
            } else {
                for (int e = tid; e < SB * SB / 2; e += 512) {
                    const int r = e >> 4, c2 = (e & 15) * 2;
                    cp_async16(dst + r * SLD + c2, A + (long long)(bi * SB + r) * ld + bj * SB + c2);
                }
            }
        }
    cp_async_commit();
    cp_async_wait<0>();
    if (tid == 0) *prog = 0;
    __syncthreads();
    D512_STAMP(1);

    // (both copies read eight elements before they store them: one warp's dependent load -> store pairs cost ~60
    //  cycles each, 4 k cycles for the 32x32 W / V block, and the release fence then waits for the stores)
    auto store_Lcol = [&](int cb, int first, int nthr) {
        for (int bi = cb; bi < 4; bi++) {
            const double* ls = Lb + blk(bi, cb) * SBSZ;
            const int stride = (bi == cb) ? DLD : SLD;
            double* dst = A + (long long)(bi * SB) * ld + cb * SB;
            for (int e0 = first; e0 < SB * SB; e0 += nthr * 8) {
                double v[8];
#pragma unroll
                for (int q = 0; q < 8; q++) {
                    const int e = e0 + q * nthr;
                    v[q] = (e < SB * SB) ? ls[(e >> 5) * stride + (e & 31)] : 0.0;
                }
#pragma unroll
                for (int q = 0; q < 8; q++) {
                    const int e = e0 + q * nthr;
                    if (e < SB * SB) dst[(long long)(e >> 5) * ld + (e & 31)] = v[q];
                }
            }
        }
    };
    auto store_Wdiag = [&](int cb, int first, int nthr) {
        const double* ws = Wd + cb * SBSZ;
        double* wdst = W + (long long)(cb * SB) * ldw + cb * SB;
        double* vdst = V ? V + (long long)(cb * SB) * ldv + cb * SB : nullptr;
        for (int e0 = first; e0 < SB * SB; e0 += nthr * 8) {
            double v[8], u[8];
#pragma unroll
            for (int q = 0; q < 8; q++) {
                const int e = e0 + q * nthr, r = e >> 5, c = e & 31;
                v[q] = (e < SB * SB) ? ws[r * SLD + c] : 0.0;
                u[q] = (e < SB * SB && vdst) ? ws[c * SLD + r] : 0.0;
            }
#pragma unroll
            for (int q = 0; q < 8; q++) {
                const int e = e0 + q * nthr, r = e >> 5, c = e & 31;
                if (e < SB * SB) {
                    wdst[(long long)r * ldw + c] = v[q];
                    if (vdst) vdst[(long long)r * ldv + c] = u[q];
                }
            }
        }
    };

    double* const S10 = Xs, * const W10 = Xs + SBSZ, * const S32 = Xs + 2 * SBSZ, * const W32 = Xs + 3 * SBSZ;
    double* const S20 = Xs + 4 * SBSZ, * const S21 = Xs + 5 * SBSZ, * const S30 = Xs + 6 * SBSZ, * const S31 = Xs;

    // Phases (one CTA barrier after each):
    //   0, 2, 4, 6  S(bb)   sweep of sub-block bb + rows below (streamed) + off-path work of step bb-1
    //   1, 3, 5     C(bb)   block column bb+1 -= L(.,bb) L(bb+1,bb)^T: what the next sweep and its rows below read
    //   7           S(4)    no sweep: inverse of L_33, L column 3 out, S_32, S_2j, S_3j
    //   8           T2      W_32, W_2j, inverted diagonal blocks out
    //   9           T3      W_3j
    const int nph = lpub ? 8 : 10;
#pragma unroll 1
    for (int ph = 0; ph < nph; ph++) {
        if (dclk && ph == 3 && tid == 0) dclk[25] = clock64();
        const bool is_s = (ph <= 7) && ((ph & 1) == 0 || ph == 7);
        const int bb = (ph <= 6) ? (ph >> 1) : 4;               // S / C index (S(4) = the inverse-only phase)
        D512Strip st0, st1;
        st0.A1 = st0.B1 = st0.A2 = st0.B2 = nullptr; st0.Cin = nullptr; st0.Cout = nullptr;
        st0.bks = st0.bns = st0.cin_s = st0.cout_s = 0; st0.sgn = 0.0; st0.wi = -1; st0.wj = st0.hf = 0; st0.dbg = nullptr;
        st1 = st0;
        if (dclk && ph == 3 && tid == 0) dclk[26] = clock64();
        int nst = 0;                                             // strips this warp runs at the end of the phase
        bool pair_sync = false;                                  // two dependent strips of a warp pair (W_10)
        if (lpub && wid == 15 && lane == 0 && (ph == 3 || ph == 5 || ph == 7)) {
            // block column (ph - 3) / 2 of L and its inverted diagonal block were stored in the previous phase (the
            // phase barrier ordered those stores before this fence)
            __threadfence();
            st_release(lpub + (ph - 3) / 2, 1);
        }
        if (is_s) {
            double* invb = invd + (bb & 3) * SB;
            if (wid == 0) {
                if (bb < 4) {
                    // ---- the sweep: 32x32 Cholesky in registers (lane = row), software-pipelined as in the
                    //      256-thread body; every finished column goes to Lt[k][.] and is announced through *prog
                    double* Ld = Lb + blk(bb, bb) * SBSZ;        // stride DLD
                    double row[SB];
#pragma unroll
                    for (int k = 0; k < SB; k++) row[k] = Ld[lane * DLD + k];
                    int fail = 0;
                    double piv = __shfl_sync(0xffffffffu, row[0], 0);
                    double lprev = 0.0;
#pragma unroll
                    for (int k = 0; k < SB; k++) {
                        if (!(piv > GPB_PIVOT_MIN) && fail == 0) fail = k + 1;   // not positive definite (or NaN)
                        const double id = rsqrt_pivot(piv);
                        const double d = piv * id;
                        const double lik = (lane == k) ? d : row[k] * id;
                        row[k] = lik;
                        if (k + 1 < SB) {
                            const double own = fma(-lik, lik, row[k + 1]);          // valid in lane k+1: next pivot
                            piv = __shfl_sync(0xffffffffu, own, k + 1);
                            const double lnext = __shfl_sync(0xffffffffu, lik, k + 1);
                            row[k + 1] = fma(-lik, lnext, row[k + 1]);
                        }
                        if (k + 2 < SB) {
                            const double lnext2 = __shfl_sync(0xffffffffu, lik, k + 2);
                            row[k + 2] = fma(-lik, lnext2, row[k + 2]);
                        }
                        if (k >= 1 && k + 2 < SB) {
                            // bulk update of column k-1: rows' elements j >= k+2 (published last iteration)
                            const double* cb = Lt + (k - 1) * LTLD;
                            if ((k + 2) & 1) row[k + 2] = fma(-lprev, cb[k + 2], row[k + 2]);
#pragma unroll
                            for (int j = (k + 3) & ~1; j + 1 < SB; j += 2) {
                                const double2 v = *reinterpret_cast<const double2*>(cb + j);
                                row[j] = fma(-lprev, v.x, row[j]);
                                row[j + 1] = fma(-lprev, v.y, row[j + 1]);
                            }
                        }
                        Lt[k * LTLD + lane] = lik;
                        if (lane == k) invb[k] = id;
                        __syncwarp();
                        // The column stream is a flag protocol inside one CTA: data stores by the whole warp, __syncwarp, then the
                        // counter by lane 0; readers poll the counter and then read the column.  It relies on shared-memory
                        // accesses of one SM being performed in issue order (they are: one LSU pipe per SM); racecheck reports
                        // these lines because it only understands barriers.  -DGPB_STREAM_FENCE adds the __threadfence_block
                        // pair the PTX memory model formally asks for: measured +12 % on the diagonal block (67.5 k -> 75.8 k
                        // cycles, N = 4096 1.61 -> 1.67 ms), same results bit for bit -- off by default.
                        if (lane == 0) {
#ifdef GPB_STREAM_FENCE
                            __threadfence_block();
#endif
                            *prog = bb * SB + k + 1;
                        }
                        lprev = lik;
                    }
#pragma unroll
                    for (int k = 0; k < SB; k++) Ld[lane * DLD + k] = (k <= lane) ? row[k] : 0.0;
                    if (fail && lane == 0 && *info == 0) *info = col0 + bb * SB + fail;
                }
            } else if (wid <= 3) {
                if (bb < 4 && wid > bb) {
                    // ---- rows below, block (wid, bb): X = A L_bb^-T through the column stream (lane = row):
                    //      x_l = a_l / L_ll, then a_c -= x_l L_cl for c > l -- the same right-looking step the
                    //      sweep applies to its own rows, one column behind it.  (Block row = warp index: from S1 on
                    //      warp 1 is free, so the 32x32 inverse of warp 5 has their shared sub-partition to itself.)
                    const int bi = wid;
                    double* Ab = Lb + blk(bi, bb) * SBSZ;
                    double* St = Wd + bi * SBSZ;              // private staging (this inverse slot is still unused)
                    for (int r = 0; r < SB; r++) St[r * DLD + lane] = Ab[r * SLD + lane];
                    __syncwarp();
                    double a[SB];
#pragma unroll
                    for (int c = 0; c < SB; c++) a[c] = St[lane * DLD + c];
                    int seen = 0;                              // columns known to be published (the sweep is usually ahead)
#pragma unroll
                    for (int l = 0; l < SB; l++) {
                        if (seen <= bb * SB + l) {
                            int p = 0;
                            if (lane == 0) { do { p = *prog; } while (p <= bb * SB + l); }
#ifdef GPB_STREAM_FENCE
                            __threadfence_block();
#endif
                            seen = __shfl_sync(0xffffffffu, p, 0);
                        }
                        const double x = a[l] * invb[l];
                        a[l] = x;
                        const double* col = Lt + l * LTLD;
                        if ((l + 1) & 1) { if (l + 1 < SB) a[l + 1] = fma(-x, col[l + 1], a[l + 1]); }
#pragma unroll
                        for (int c = (l + 2) & ~1; c + 1 < SB; c += 2) {
                            const double2 v = *reinterpret_cast<const double2*>(col + c);
                            a[c] = fma(-x, v.x, a[c]);
                            a[c + 1] = fma(-x, v.y, a[c + 1]);
                        }
                    }
#pragma unroll
                    for (int c = 0; c < SB; c++) St[lane * DLD + c] = a[c];
                    __syncwarp();
                    for (int r = 0; r < SB; r++) Ab[r * SLD + lane] = St[r * DLD + lane];
                } else if (wid == 1 && bb > 0) {
                    // previous L column -> global, first half (second half: warp 5 after its inverse).  Both sit on
                    // sub-partition 1, which has no rows-below warp from S1 on: next to one, the store loop was starved
                    // (its few instructions are refetched from L2 while 25 KB of straight-line code stream by)
                    store_Lcol(bb - 1, lane, 64);
                }
            } else if (wid == 5) {
                if (bb > 0) {
                    // ---- inverse of the previous 32x32 factor, two levels: both 16x16 diagonal blocks at once by forward
                    //      substitution (lanes 0-15: column cc of W_11 from L_11, lanes 16-31: of W_22 from L_22; 120 FMAs
                    //      per lane), then W_21 = -W_22 (L_21 W_11) as two 16x16x16 DMMA products.  (The 32-row substitution
                    //      -- 496 dependent-ish FMAs of straight-line code per lane -- took 10-15 k cycles and was the last
                    //      warp of every sweep phase.)
                    const double* Lp = Lb + blk(bb - 1, bb - 1) * SBSZ;       // stride DLD
                    const double* ip = invd + (bb - 1) * SB;
                    double* Wo = Wd + (bb - 1) * SBSZ;                        // stride SLD
                    const int hi = lane >> 4, cc = lane & 15;
                    if (dclk && lane == 0) dclk[200 + ph * 6 + 0] = clock64();
                    const double* Lh = Lp + (hi * 16) * DLD + hi * 16;
                    double w[16];
#pragma unroll
                    for (int r = 0; r < 16; r++) {
                        double s0 = 0.0, s1 = 0.0;
#pragma unroll
                        for (int k = 0; k + 1 < r; k += 2) {
                            s0 = fma(Lh[r * DLD + k], w[k], s0);
                            s1 = fma(Lh[r * DLD + k + 1], w[k + 1], s1);
                        }
                        if (r & 1) s0 = fma(Lh[r * DLD + r - 1], w[r - 1], s0);
                        w[r] = (((r == cc) ? 1.0 : 0.0) - (s0 + s1)) * ip[hi * 16 + r];
                    }
#pragma unroll
                    for (int r = 0; r < 16; r++) Wo[(hi * 16 + r) * SLD + lane] = w[r];
                    __syncwarp();
                    if (dclk && lane == 0) dclk[200 + ph * 6 + 1] = clock64();
                    {
                        const int g = lane >> 2, t = lane & 3;
                        double* Tt = Wo + 16;                                  // scratch: the (zero) block W_12
                        // T = L_21 W_11   (W_11 lower triangular: k >= column; all 16 k are walked, the zeros are stored)
                        double c[2][2][2];
#pragma unroll
                        for (int ti = 0; ti < 2; ti++)
#pragma unroll
                            for (int tj = 0; tj < 2; tj++) c[ti][tj][0] = c[ti][tj][1] = 0.0;
#pragma unroll
                        for (int kk = 0; kk < 4; kk++) {
                            double av[2], bv[2];
#pragma unroll
                            for (int ti = 0; ti < 2; ti++) av[ti] = Lp[(16 + ti * 8 + g) * DLD + kk * 4 + t];
#pragma unroll
                            for (int tj = 0; tj < 2; tj++) bv[tj] = Wo[(kk * 4 + t) * SLD + tj * 8 + g];
#pragma unroll
                            for (int ti = 0; ti < 2; ti++)
#pragma unroll
                                for (int tj = 0; tj < 2; tj++) dmma884(c[ti][tj][0], c[ti][tj][1], av[ti], bv[tj]);
                        }
#pragma unroll
                        for (int ti = 0; ti < 2; ti++)
#pragma unroll
                            for (int tj = 0; tj < 2; tj++) {
                                Tt[(ti * 8 + g) * SLD + tj * 8 + 2 * t] = c[ti][tj][0];
                                Tt[(ti * 8 + g) * SLD + tj * 8 + 2 * t + 1] = c[ti][tj][1];
                                c[ti][tj][0] = c[ti][tj][1] = 0.0;
                            }
                        __syncwarp();
                        // W_21 = -W_22 T
#pragma unroll
                        for (int kk = 0; kk < 4; kk++) {
                            double av[2], bv[2];
#pragma unroll
                            for (int ti = 0; ti < 2; ti++) av[ti] = -Wo[(16 + ti * 8 + g) * SLD + 16 + kk * 4 + t];
#pragma unroll
                            for (int tj = 0; tj < 2; tj++) bv[tj] = Tt[(kk * 4 + t) * SLD + tj * 8 + g];
#pragma unroll
                            for (int ti = 0; ti < 2; ti++)
#pragma unroll
                                for (int tj = 0; tj < 2; tj++) dmma884(c[ti][tj][0], c[ti][tj][1], av[ti], bv[tj]);
                        }
                        __syncwarp();
#pragma unroll
                        for (int ti = 0; ti < 2; ti++)
#pragma unroll
                            for (int tj = 0; tj < 2; tj++) {
                                Wo[(16 + ti * 8 + g) * SLD + tj * 8 + 2 * t] = c[ti][tj][0];
                                Wo[(16 + ti * 8 + g) * SLD + tj * 8 + 2 * t + 1] = c[ti][tj][1];
                                Tt[(ti * 8 + g) * SLD + tj * 8 + 2 * t] = 0.0;
                                Tt[(ti * 8 + g) * SLD + tj * 8 + 2 * t + 1] = 0.0;
                            }
                    }
                    if (dclk && lane == 0) dclk[200 + ph * 6 + 2] = clock64();
                    if (lpub) {
                        __syncwarp();
                        store_Wdiag(bb - 1, lane, 32);
                    }
                    store_Lcol(bb - 1, 32 + lane, 64);
                    if (dclk && lane == 0) dclk[200 + ph * 6 + 3] = clock64();
                }
            } else if (wid >= 9 && (wid & 3) != 0) {
                // (warps 4, 8, 12 share the sweep's sub-partition and stay idle while it runs)
                const int slot = wid - 9 - (wid > 12);                  // warps 9 10 11 13 14 15 -> 0..5
                if (bb == 1 || bb == 2) {
                    // ---- the rest of the previous in-tile update: blocks (i, j), j >= bb + 1
                    const int nb = 3 - bb, npairs = nb * (nb + 1) / 2;
                    if (slot < npairs * 2) {
                        const int pr = slot >> 1;
                        int ii = 0;
                        while ((ii + 1) * (ii + 2) / 2 <= pr) ii++;
                        const int jj = pr - ii * (ii + 1) / 2;
                        st0 = d512_update(Lb, bb + 1 + ii, bb + 1 + jj, bb - 1, slot & 1);
                        nst = 1;
                    }
                } else if (bb == 0 && slot == 0) {
                    // nothing to update yet: one strip on scratch blocks keeps the strip code in the instruction
                    // cache for C(0) (the sweep streams 36 KB of straight-line code through the 32 KB cache)
                    st0 = d512_prod(S20, S21, nullptr, nullptr, 0.0, S30, -1, 0, 0);
                    nst = 1;
                } else if (bb == 3 && slot < 2 && !lpub) {
                    // ---- first level of the 128-level inverse, left half: W_10 = -W_11 (L_10 W_00)
                    st0 = d512_prod(Lb + blk(1, 0) * SBSZ, Wd, nullptr, nullptr, 1.0, S10, -1, 0, slot);
                    st1 = d512_prod(Wd + SBSZ, S10, nullptr, nullptr, -1.0, W10, 1, 0, slot);
                    nst = 2;
                    pair_sync = true;
                }
            }
            if (lpub && bb == 4 && (wid == 1 || wid == 5)) {
                // last block column of L (warps 6, 7) and its inverted diagonal block (warp 5) are stored: publish.
                // (ONE barrier instruction for the three warps: synccheck rejects a named barrier whose
                //  participants arrive from different instructions.)  Columns 0..2 are released by warp 15 at the
                // start of the NEXT phase instead: the release fence waits 2-5 k cycles for the stores to be
                // acknowledged, and inside the sweep phase that made warp 5 the last warp at every phase barrier.
                __syncwarp();
                asm volatile("bar.sync 6, 64;\n" ::: "memory");
                if (dclk && wid == 5 && lane == 0) dclk[200 + ph * 6 + 4] = clock64();
                if (wid == 5 && lane == 0) {
                    __threadfence();
                    st_release(lpub + (bb - 1), 1);
                }
                if (dclk && wid == 5 && lane == 0) dclk[200 + ph * 6 + 5] = clock64();
            }
            if (bb == 4 && !lpub) {
                // S(4): S_32 = L_32 W_22 (warps 9, 10) and S_ij = sum_k L_ik W_kj, i = 2,3; j = 0,1 (eight strips)
                int item = -1;
                if (wid == 9 || wid == 10) {
                    st0 = d512_prod(Lb + blk(3, 2) * SBSZ, Wd + 2 * SBSZ, nullptr, nullptr, 1.0, S32, -1, 0, wid - 9);
                    nst = 1;
                } else if (wid < 4) item = wid;
                else if (wid >= 11 && wid <= 14) item = wid - 7;
                if (item >= 0) {
                    const int i = 2 + (item >> 2), j = (item >> 1) & 1, hf = item & 1;
                    double* Sdst = (i == 2) ? (j ? S21 : S20) : (j ? S31 : S30);
                    if (j == 0) st0 = d512_prod(Lb + blk(i, 0) * SBSZ, Wd, Lb + blk(i, 1) * SBSZ, W10, 1.0, Sdst, -1, 0, hf);
                    else st0 = d512_prod(Lb + blk(i, 1) * SBSZ, Wd + SBSZ, nullptr, nullptr, 1.0, Sdst, -1, 0, hf);
                    nst = 1;
                }
            }
        } else if (ph < 7) {
            // C(bb): the block column the next sweep and its rows-below warps read
            if (dclk && ph == 3 && tid == 0) dclk[27] = clock64();
            if (wid < 2 * (3 - bb)) { st0 = d512_update(Lb, bb + 1 + (wid >> 1), bb + 1, bb, wid & 1); nst = 1; }
            if (dclk && ph == 3) st0.dbg = dclk + 16;
        } else if (ph == 8) {
            // T2: W_32 = -W_33 S_32 | W_2j = -W_22 S_2j | inverted diagonal blocks out
            if (wid < 2) { st0 = d512_prod(Wd + 3 * SBSZ, S32, nullptr, nullptr, -1.0, W32, 3, 2, wid); nst = 1; }
            else if (wid < 6) {
                const int j = (wid - 2) >> 1, hf = (wid - 2) & 1;
                st0 = d512_prod(Wd + 2 * SBSZ, j ? S21 : S20, nullptr, nullptr, -1.0, nullptr, 2, j, hf);
                nst = 1;
            } else {
                for (int cb = 0; cb < 4; cb++) store_Wdiag(cb, tid - 6 * 32, 320);
            }
        } else {
            // T3: W_3j = -(W_32 S_2j + W_33 S_3j)
            if (wid < 4) {
                const int j = wid >> 1, hf = wid & 1;
                st0 = d512_prod(W32, j ? S21 : S20, Wd + 3 * SBSZ, j ? S31 : S30, -1.0, nullptr, 3, j, hf);
                nst = 1;
            }
        }
        // Off-path strips of a sweep phase start when the sweep is nearly through: they finish with it, and the
        // strip code is still in the instruction cache when the C phase -- on the critical path -- needs it.
        if (is_s && bb < 4 && nst > 0) {
            const int thr = pair_sync ? 14 : 24;
            if (lane == 0) while (*prog < bb * SB + thr) {}
            __syncwarp();
        }
        // ---- the phase's DMMA strips (the one shared copy of the strip code)
#pragma unroll 1
        if (dclk && ph == 3 && tid == 0) dclk[20] = clock64();
        for (int q = 0; q < nst; q++) {
            d512_strip<KKU>(st0, W, ldw, V, ldv);
            if (pair_sync && q == 0) asm volatile("bar.sync 5, 64;\n" ::: "memory");
            st0 = st1;
        }
        if (dclk && lane == 0) dclk[32 + ph * 16 + wid] = clock64();      // this warp's arrival at the phase barrier
        __syncthreads();
        D512_STAMP(2 + ph);
        if (dclk && ph == 2 && tid == 0) dclk[24] = clock64();
    }
#undef D512_STAMP
}

}  // namespace
