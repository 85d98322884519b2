// diag_block.cuh -- factorisation + inversion of one 128 x 128 diagonal block by 256 threads
// (device body shared by potrf_diag_kernel in potrf.cu and the dataflow factorisation in chain.cu).
#pragma once
#include "common.cuh"

namespace {

// ===========================================================================
// diagonal-block kernel: one CTA factors and inverts one 128 x 128 block held in
// shared memory as ten 32 x 32 sub-blocks (lower block-triangle).
//
// Resource shape is part of the design: 126.75 KB of shared memory and <= 128 registers
// x 256 threads, so that a diagonal-block CTA fits on an SM next to ONE resident 64x128 TMA
// GEMM CTA (97 KB, 288 x 96 registers) of another candidate group instead of waiting for a
// completely idle SM -- this is what lets the serial part of one group's factorisation overlap
// the trailing updates of the others.  (Stream priorities for this kernel were measured: no gain.)
//
// Measured trade-offs behind the structure (tests/gpu_diag_clk.py, profiles/README.md):
//   * the 32x32 sweeps (P1) and substitutions (P2/P3) are FULLY UNROLLED although that makes the
//     kernel ~150 KB of SASS against a 32 KB L1.5 instruction cache (warp sampling: a third of the
//     active issue slots wait on instruction fetch, and the small DMMA phases pay for it).  A fully
//     rolled variant (shifted register slots, 24 KB of code) was built and measured: the DMMA
//     phases got 2x faster, but the single-warp phases got slower by more (P1 8.5k -> 10.4k,
//     P2/P3 2.9k -> 8.1k cycles per step: twice the FMAs, no triangular savings), 105k vs 78k cycles.
//   * one warp issues dependent DFMA every 8 cycles, a 64-bit shuffle costs 26, rsqrt 66: the
//     column sweep is bound by that chain (~265 cycles per column), not by throughput.
//
// Per 32-column step bb:
//   P1  warp 0        Cholesky of the 32x32 diagonal sub-block in registers (lane = row,
//                     pivots / columns by warp shuffle, rsqrt on the critical path)
//   P2  warp 0        its inverse (lane = column, forward substitution)      } concurrently
//   P3  warps 1..3    rows below: X = A L_bb^-T by substitution, lane = row   }
//   P4  all warps     trailing update inside the tile on DMMA
// then W = L^-1 is completed block by block (DMMA), results streamed to W / V = W^T.
// ===========================================================================
#ifndef DIAG_MIN_CTAS
#define DIAG_MIN_CTAS 2
#endif
constexpr int SB = 32;            // sub-block edge
constexpr int SLD = 36;           // padded stride: 36 = 4 (mod 16) -> conflict-free DMMA fragment loads
constexpr int SBSZ = SB * SLD;
constexpr int NBLK = 10;
constexpr int DLD = 33;           // odd stride for lane-per-row accesses
constexpr int DIAG_SMEM = ((NBLK + 4) * SBSZ + SB + 2 * SB) * 8;   // L blocks, diagonal inverses, 1/diag, column broadcast (126.75 KB)

__device__ __forceinline__ int blk(int bi, int bj) { return bi * (bi + 1) / 2 + bj; }

// 1/sqrt(x) for a positive, normal pivot: MUFU.RSQ64H (2^-22.9) + one third-order correction
// (error ~ e^3 = 2^-66).  libdevice's rsqrt() adds a special-value path behind a CALL; in the fully
// unrolled sweep that call costs a register spill + reload and a BRA.DIV convergence check per
// column, all on the serial pivot chain (cuobjdump of the previous build).  Pivots that are not
// positive normal numbers are reported as a failed factorisation by the caller instead.
__device__ __forceinline__ double rsqrt_pivot(double x) {
    double y;
    asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(x));
    const double e = fma(-x * y, y, 1.0);
    return fma(y * e, fma(e, 0.375, 0.5), y);
}
// smallest pivot taken as positive definite (far below anything a kernel matrix + s^2 I produces)
#define GPB_PIVOT_MIN 1e-290

// optional phase timing of the diagonal-block kernel (build with -DGPB_DIAG_CLK; read back with
// gpb_debug_diag_clk): clock64 stamps of CTA 0 / thread 0 at the phase boundaries.
#ifdef GPB_DIAG_CLK
__device__ long long g_diag_clk[64];
#define DIAG_STAMP(i) do { if (blockIdx.x == 0 && threadIdx.x == 0) g_diag_clk[i] = clock64(); } while (0)
#else
#define DIAG_STAMP(i) do { } while (0)
#endif

// 16 x 32 strip (rows hf*16 .. +16) of  acc += sgn * Ablk * Bblk  for 32 x 32 blocks held row-major
// with stride SLD: eight independent DMMA accumulators per warp, so the ~100-cycle latency of a
// dependent DMMA chain is covered by issue from the other seven tiles.
__device__ __forceinline__ void strip_mm(double (&acc)[2][4][2], const double* Ablk, const double* Bblk,
                                         int hf, int g, int t, double sgn) {
    const double* Ap = Ablk + (hf * 16 + g) * SLD + t;
    const double* Bp = Bblk + t * SLD + g;
#pragma unroll
    for (int kk = 0; kk < 8; kk++) {
        double a[2], b[4];
#pragma unroll
        for (int rt = 0; rt < 2; rt++) a[rt] = sgn * Ap[rt * 8 * SLD + kk * 4];
#pragma unroll
        for (int ct = 0; ct < 4; ct++) b[ct] = Bp[kk * 4 * SLD + ct * 8];
#pragma unroll
        for (int rt = 0; rt < 2; rt++)
#pragma unroll
            for (int ct = 0; ct < 4; ct++) dmma884(acc[rt][ct][0], acc[rt][ct][1], a[rt], b[ct]);
    }
}

__device__ __forceinline__ void strip_zero(double (&acc)[2][4][2]) {
#pragma unroll
    for (int rt = 0; rt < 2; rt++)
#pragma unroll
        for (int ct = 0; ct < 4; ct++) acc[rt][ct][0] = acc[rt][ct][1] = 0.0;
}

// strip -> shared block (row-major, stride SLD)
__device__ __forceinline__ void strip_to_smem(const double (&acc)[2][4][2], double* blkp, int hf, int g, int t) {
#pragma unroll
    for (int rt = 0; rt < 2; rt++)
#pragma unroll
        for (int ct = 0; ct < 4; ct++) {
            double* p = blkp + (hf * 16 + rt * 8 + g) * SLD + ct * 8 + 2 * t;
            p[0] = acc[rt][ct][0];
            p[1] = acc[rt][ct][1];
        }
}

// strip of W_ij -> global W (row-major) and its transpose into V
__device__ __forceinline__ void strip_to_wv(const double (&acc)[2][4][2], double* W, long long ldw, double* V,
                                            long long ldv, int i, int j, int hf, int g, int t) {
#pragma unroll
    for (int rt = 0; rt < 2; rt++)
#pragma unroll
        for (int ct = 0; ct < 4; ct++) {
            const long long gr = i * SB + hf * 16 + rt * 8 + g, gc = j * SB + ct * 8 + 2 * t;
            *reinterpret_cast<double2*>(W + gr * ldw + gc) = make_double2(acc[rt][ct][0], acc[rt][ct][1]);
            if (V) {
                V[gc * ldv + gr] = acc[rt][ct][0];
                V[(gc + 1) * ldv + gr] = acc[rt][ct][1];
            }
        }
}

// NAMED = true: the caller is a wider CTA (the dataflow factorisation's chain CTA) and only its first 256
// threads run the body; they synchronise on named barrier 1 instead of the CTA-wide barrier.
template <bool NAMED>
__device__ __forceinline__ void diag_bar() {
    if (NAMED) asm volatile("bar.sync 1, 256;\n" ::: "memory");
    else __syncthreads();
}

template <bool NAMED>
__device__ __forceinline__ void diag_block_body(double* A, long long ld, double* W, long long ldw, double* V,
                                                long long ldv, int* info, int col0, int nsub, double* sm) {
    double* Lb = sm;                          // 10 lower sub-blocks of the tile: off-diagonal ones with stride
                                              // SLD (DMMA fragments), diagonal ones with stride DLD (lane = row)
    double* Wd = sm + NBLK * SBSZ;            // inverses of the 4 diagonal sub-blocks (stride SLD)
    double* invd = Wd + 4 * SBSZ;
    double* colbuf = invd + SB;               // 2 x 32: column k of the 32x32 factor, broadcast to all lanes
    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5, g = lane >> 2, t = lane & 3;
    DIAG_STAMP(0);
    // ---- load the lower block-triangle (all copies in flight before the single wait) -------
    const bool vec_ok = ((reinterpret_cast<uintptr_t>(A) & 15) == 0) && ((ld & 1) == 0);
    for (int bi = 0; bi < 4; bi++)
        for (int bj = 0; bj <= bi; bj++) {
            double* dst = Lb + blk(bi, bj) * SBSZ;
            if (bi == bj) {
                for (int e = tid; e < SB * SB; e += 256) {
                    const int r = e >> 5, c = e & 31;
                    cp_async8(dst + r * DLD + c, A + (long long)(bi * SB + r) * ld + bj * SB + c);
                }
            } else if (vec_ok) {
                for (int e = tid; e < SB * SB / 2; e += 256) {
                    const int r = e >> 4, c2 = (e & 15) * 2;
                    cp_async16(dst + r * SLD + c2, A + (long long)(bi * SB + r) * ld + bj * SB + c2);
                }
            } else {
                for (int e = tid; e < SB * SB; e += 256) {
                    const int r = e >> 5, c = e & 31;
                    dst[r * SLD + c] = A[(long long)(bi * SB + r) * ld + bj * SB + c];
                }
            }
        }
    cp_async_commit();
    cp_async_wait<0>();
    diag_bar<NAMED>();

    // column `cb` of the tile is final: stream out L (cb..3, cb) and the inverted diagonal block
    auto store_column = [&](int cb, int first, int nthr) {
        for (int bi = cb; bi < 4; bi++) {
            const double* ls = Lb + blk(bi, cb) * SBSZ;
            const int stride = (bi == cb) ? DLD : SLD;
            for (int e = first; e < SB * SB; e += nthr) {
                const int r = e >> 5, c = e & 31;
                A[(long long)(bi * SB + r) * ld + cb * SB + c] = ls[r * stride + c];
            }
        }
        const double* ws = Wd + cb * SBSZ;
        for (int e = first; e < SB * SB; e += nthr) {
            const int r = e >> 5, c = e & 31;
            const long long gr = cb * SB + r, gc = cb * SB + c;
            W[gr * ldw + gc] = ws[r * SLD + c];
            if (V) V[gr * ldv + gc] = ws[c * SLD + r];
        }
    };

    DIAG_STAMP(1);
    for (int bb = 0; bb < 4; bb++) {
        double* Ld = Lb + blk(bb, bb) * SBSZ;        // stride DLD
        if (bb >= nsub) {
            // identity pad (rows/cols beyond the observations): the sub-column is (I, 0, ..) already and
            // every update it would apply is zero -- skip the sweep, W_bb = I.  At the reference's
            // test-suite scale (N = 50 -> 2 of 4 sweeps) this halves the kernel.
            if (wid == 0) {
                double* Wo = Wd + bb * SBSZ;
#pragma unroll 4
                for (int r = 0; r < SB; r++) Wo[r * SLD + lane] = (r == lane) ? 1.0 : 0.0;
            } else if (bb > 0) {
                store_column(bb - 1, tid - 32, 224);
            }
            diag_bar<NAMED>();
            continue;
        }
        if (wid == 0) {
            // ---- P1 (warp 0): 32x32 Cholesky in registers, in place (lane = row) ---------------
            double row[SB];
#pragma unroll
            for (int k = 0; k < SB; k++) row[k] = Ld[lane * DLD + k];
            int fail = 0;
            double myinv = 0.0;
            // Software-pipelined sweep: column k's critical part (pivot -> rsqrt -> scale -> the two rows
            // the next two pivots need, by shuffle) is issued BEFORE the bulk update of column k-1
            // (columns >= k+2, through the shared-memory broadcast), so the bulk FMAs and the
            // store -> load round trip fill the latency of the serial chain instead of extending it.
            // The next pivot is formed in lane k+1 from its own values (no wait for the row shuffle).
            double piv = __shfl_sync(0xffffffffu, row[0], 0);
            double lprev = 0.0;
#pragma unroll
            for (int k = 0; k < SB; k++) {
                if (!(piv > GPB_PIVOT_MIN) && fail == 0) fail = k + 1;   // not positive definite (or NaN)
                const double id = rsqrt_pivot(piv);
                const double d = piv * id;
                const double lik = (lane == k) ? d : row[k] * id;
                if (lane == k) myinv = id;
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
                    // bulk update of column k-1: rows' elements j >= k+2
                    const double* cb = colbuf + ((k - 1) & 1) * SB;
                    __syncwarp();
                    if ((k + 2) & 1) row[k + 2] = fma(-lprev, cb[k + 2], row[k + 2]);
#pragma unroll
                    for (int j = (k + 3) & ~1; j + 1 < SB; j += 2) {
                        const double2 v = *reinterpret_cast<const double2*>(cb + j);
                        row[j] = fma(-lprev, v.x, row[j]);
                        row[j + 1] = fma(-lprev, v.y, row[j + 1]);
                    }
                }
                // column k for the next iteration's bulk update (this buffer was last read two
                // iterations ago, with a __syncwarp in between)
                if (k + 3 < SB) colbuf[(k & 1) * SB + lane] = lik;
                lprev = lik;
            }
#pragma unroll
            for (int k = 0; k < SB; k++) Ld[lane * DLD + k] = (k <= lane) ? row[k] : 0.0;
            invd[lane] = myinv;
            if (fail && lane == 0 && *info == 0) *info = col0 + bb * SB + fail;
        } else if (bb > 0) {
            // ---- warps 1..7 meanwhile: the previous column is final, stream it out -------------
            store_column(bb - 1, tid - 32, 224);
        }
        diag_bar<NAMED>();
        DIAG_STAMP(2 + bb * 3);

        const int nbelow = 3 - bb;
        if (wid == 0) {
            // ---- P2 (warp 0): inverse of L_bb, lane c solves L w = e_c (rows of L broadcast) ----
            double w[SB];
#pragma unroll
            for (int r = 0; r < SB; r++) {
                double s0 = 0.0, s1 = 0.0;
#pragma unroll
                for (int k = 0; k + 1 < r; k += 2) {
                    s0 = fma(Ld[r * DLD + k], w[k], s0);
                    s1 = fma(Ld[r * DLD + k + 1], w[k + 1], s1);
                }
                if (r & 1) s0 = fma(Ld[r * DLD + r - 1], w[r - 1], s0);
                w[r] = (((r == lane) ? 1.0 : 0.0) - (s0 + s1)) * invd[r];
            }
            double* Wo = Wd + bb * SBSZ;
#pragma unroll
            for (int r = 0; r < SB; r++) Wo[r * SLD + lane] = w[r];
        } else if (wid <= nbelow && bb + wid < nsub) {      // (rows in the identity pad stay zero)
            // ---- P3 (warps 1..nbelow): sub-block (bb+wid, bb): X = A L_bb^-T, lane = row.
            //      The still unused inverse slot of that block row is the warp's private staging
            //      area, so every shared-memory access below is conflict-free.
            const int bi = bb + wid;
            double* Ab = Lb + blk(bi, bb) * SBSZ;
            double* St = Wd + bi * SBSZ;
            for (int r = 0; r < SB; r++) St[r * DLD + lane] = Ab[r * SLD + lane];
            __syncwarp();
            double xr[SB];
#pragma unroll
            for (int c = 0; c < SB; c++) xr[c] = St[lane * DLD + c];
#pragma unroll
            for (int c = 0; c < SB; c++) {
                double s0 = 0.0, s1 = 0.0;
#pragma unroll
                for (int l = 0; l + 1 < c; l += 2) {
                    s0 = fma(xr[l], Ld[c * DLD + l], s0);
                    s1 = fma(xr[l + 1], Ld[c * DLD + l + 1], s1);
                }
                if (c & 1) s0 = fma(xr[c - 1], Ld[c * DLD + c - 1], s0);
                xr[c] = (xr[c] - (s0 + s1)) * invd[c];
            }
#pragma unroll
            for (int c = 0; c < SB; c++) St[lane * DLD + c] = xr[c];
            __syncwarp();
            for (int r = 0; r < SB; r++) Ab[r * SLD + lane] = St[r * DLD + lane];
        }
        diag_bar<NAMED>();
        DIAG_STAMP(3 + bb * 3);

        // ---- P4: trailing update A_ij -= L_ib L_jb^T, bb < j <= i, on DMMA: one 16 x 32 strip
        //      (eight independent accumulators) per warp item ------------------------------------
        const int npairs = nbelow * (nbelow + 1) / 2;
        for (int item = wid; item < npairs * 2; item += 8) {
            const int pr = item >> 1, hf = item & 1;
            int ii = 0;
            while ((ii + 1) * (ii + 2) / 2 <= pr) ii++;
            const int jj = pr - ii * (ii + 1) / 2;
            const int bi = bb + 1 + ii, bj = bb + 1 + jj;
            if (bi >= nsub) continue;                 // L_ib = 0 in the identity pad: nothing to subtract
            double* Cb = Lb + blk(bi, bj) * SBSZ;
            const int cs = (bi == bj) ? DLD : SLD;
            double acc[2][4][2];
#pragma unroll
            for (int rt = 0; rt < 2; rt++)
#pragma unroll
                for (int ct = 0; ct < 4; ct++) {
                    const double* p = Cb + (hf * 16 + rt * 8 + g) * cs + ct * 8 + 2 * t;
                    acc[rt][ct][0] = p[0];
                    acc[rt][ct][1] = p[1];
                }
            // B operand "col" layout: B[k][n] = L_jb[n][k]
            const double* Ap = Lb + blk(bi, bb) * SBSZ + (hf * 16 + g) * SLD + t;
            const double* Bp = Lb + blk(bj, bb) * SBSZ + g * SLD + t;
#pragma unroll
            for (int kk = 0; kk < 8; kk++) {
                double a[2], b[4];
#pragma unroll
                for (int rt = 0; rt < 2; rt++) a[rt] = -Ap[rt * 8 * SLD + kk * 4];
#pragma unroll
                for (int ct = 0; ct < 4; ct++) b[ct] = Bp[ct * 8 * SLD + kk * 4];
#pragma unroll
                for (int rt = 0; rt < 2; rt++)
#pragma unroll
                    for (int ct = 0; ct < 4; ct++) dmma884(acc[rt][ct][0], acc[rt][ct][1], a[rt], b[ct]);
            }
#pragma unroll
            for (int rt = 0; rt < 2; rt++)
#pragma unroll
                for (int ct = 0; ct < 4; ct++) {
                    double* p = Cb + (hf * 16 + rt * 8 + g) * cs + ct * 8 + 2 * t;
                    p[0] = acc[rt][ct][0];
                    p[1] = acc[rt][ct][1];
                }
        }
        diag_bar<NAMED>();
        DIAG_STAMP(4 + bb * 3);
    }

    // last column (L_33, W_33); afterwards the four diagonal L slots are scratch for the inverse
    store_column(3, tid, 256);
    diag_bar<NAMED>();
    DIAG_STAMP(14);

    // ---- P5: off-diagonal sub-blocks of W = L^-1 by pairwise merges (same recursion as trtri):
    //      level 1  W_10 = -W_11 (L_10 W_00),  W_32 = -W_33 (L_32 W_22)
    //      level 2  W[2:4,0:2] = -W[2:4,2:4] (L[2:4,0:2] W[0:2,0:2])        (64 x 64 blocks)
    //      scratch: S_10 -> diag slot 0, S_32 -> diag slot 1, W_10 -> diag slot 2, W_32 -> diag slot 3,
    //      then S_20, S_21 -> diag slots 0, 1 and S_30, S_31 -> the dead L_10, L_32 slots.
    double* const dg0 = Lb + blk(0, 0) * SBSZ;
    double* const dg1 = Lb + blk(1, 1) * SBSZ;
    double* const W10 = Lb + blk(2, 2) * SBSZ;
    double* const W32 = Lb + blk(3, 3) * SBSZ;
    double acc[2][4][2];
    if (wid < 4) {                                    // level 1a
        const int q = wid >> 1, hf = wid & 1, i = q ? 3 : 1, j = q ? 2 : 0;
        strip_zero(acc);
        if (i < nsub) strip_mm(acc, Lb + blk(i, j) * SBSZ, Wd + j * SBSZ, hf, g, t, 1.0);   // else L_ij = 0
        strip_to_smem(acc, q ? dg1 : dg0, hf, g, t);
    }
    diag_bar<NAMED>();
    if (wid < 4) {                                    // level 1b
        const int q = wid >> 1, hf = wid & 1, i = q ? 3 : 1, j = q ? 2 : 0;
        strip_zero(acc);
        if (i < nsub) strip_mm(acc, Wd + i * SBSZ, q ? dg1 : dg0, hf, g, t, -1.0);
        strip_to_smem(acc, q ? W32 : W10, hf, g, t);
        strip_to_wv(acc, W, ldw, V, ldv, i, j, hf, g, t);
    }
    diag_bar<NAMED>();
    DIAG_STAMP(15);
    {                                                 // level 2a: S_ij = sum_k L_ik W_kj, k = j..1
        const int i = 2 + (wid >> 2), j = (wid >> 1) & 1, hf = wid & 1;
        strip_zero(acc);
        if (i >= nsub) {
            // block row i of L is zero left of the diagonal: S_ij = 0
        } else if (j == 0) {
            strip_mm(acc, Lb + blk(i, 0) * SBSZ, Wd, hf, g, t, 1.0);
            strip_mm(acc, Lb + blk(i, 1) * SBSZ, W10, hf, g, t, 1.0);
        } else {
            strip_mm(acc, Lb + blk(i, 1) * SBSZ, Wd + SBSZ, hf, g, t, 1.0);
        }
        // destinations (diag slots 0/1, dead L_10 / L_32) are not read by anyone in this phase
        double* Sdst = (i == 2) ? (j ? dg1 : dg0) : (j ? Lb + blk(3, 2) * SBSZ : Lb + blk(1, 0) * SBSZ);
        strip_to_smem(acc, Sdst, hf, g, t);
    }
    diag_bar<NAMED>();
    DIAG_STAMP(16);
    {                                                 // level 2b: W_ij = -sum_k W_ik S_kj, k = 2..i
        const int i = 2 + (wid >> 2), j = (wid >> 1) & 1, hf = wid & 1;
        const double* S2 = j ? dg1 : dg0;
        const double* S3 = j ? Lb + blk(3, 2) * SBSZ : Lb + blk(1, 0) * SBSZ;
        strip_zero(acc);
        if (i >= nsub) {
            // W_ij = 0
        } else if (i == 2) {
            strip_mm(acc, Wd + 2 * SBSZ, S2, hf, g, t, -1.0);
        } else {
            strip_mm(acc, W32, S2, hf, g, t, -1.0);
            strip_mm(acc, Wd + 3 * SBSZ, S3, hf, g, t, -1.0);
        }
        strip_to_wv(acc, W, ldw, V, ldv, i, j, hf, g, t);
    }
    DIAG_STAMP(17);
}

}  // namespace
