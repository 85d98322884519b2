// gemm.cu -- the one dense contraction of the GP hot path, on FP64 tensor cores.
//
//   C[i,j] = beta * C[i,j] + alpha * sum_{k in [kbeg(i,j), kend(i,j))} A[i,k] * B[j,k]
//
// "NT" form: both operands are row-major with the contraction index contiguous,
// which is exactly the m8n8k4 DMMA fragment layout (A row-major, B "col").  Every
// O(N^3) step of the path is phrased this way:
//   TRSM by inverted diagonal block   P_ik = A_ik * W_kk^T
//   SYRK trailing update              A_ij -= P_ik * P_jk^T           (lower tiles only)
//   trtri combine                     Tt = V11 * L21^T ;  W21 = -W22 * Tt^T  (+ mirrored V12)
//   lauum                             Ki = V * V^T                    (lower + mirror)
//   posterior cov                     Z = Kxox * W^T ;  C = Kxoxo - Z Z^T
// Triangular operands are exploited at tile granularity through the k-range.
//
// Tile (64 | 128) x 128 x 16 with 32 x 32 warp tiles (4 x 4 DMMA tiles, 32 fp64 accumulators
// per thread, ~100 registers): the DMMA pipe needs >= 3 resident warps per SM sub-partition to
// saturate (measured: 1 warp/SMSP 44 %, 2 warps 87 % -- profiles/), so the default 64 x 128 CTA
// has 8 warps and two CTAs share an SM (4 warps per SMSP).  cp.async multi-stage pipeline.
// All extents are multiples of the tile (buffers are padded by the host layer), so there is
// no edge predication in the main loop.
#include <cuda.h>
#include <stdlib.h>
#include "common.cuh"
#include "kfunctors.cuh"
#include "launch.h"

namespace {

constexpr int BN = 128, BK = 16;
constexpr int LDS = BK + 4;   // padded row stride (doubles): conflict-free 8-byte fragment loads

// Two tile shapes share the code:
//   BM = 128: 512 threads (4 x 4 warps), 4 stages (160 KB smem), one CTA per SM
//   BM =  64: 256 threads (2 x 4 warps), 3 stages ( 90 KB smem), two CTAs per SM -- the epilogue /
//             prologue of one CTA overlaps the main loop of the other (short-K trailing updates)
template <int BM> struct TileCfg {
    static constexpr int THREADS = BM * 4;
    static constexpr int STAGES = (BM == 128) ? 4 : 3;
    static constexpr int A_DOUBLES = BM * LDS;
    static constexpr int B_DOUBLES = BN * LDS;
    static constexpr int SMEM_BYTES = STAGES * (A_DOUBLES + B_DOUBLES) * 8;
    static constexpr int MIN_CTAS = (BM == 128) ? 1 : 2;
};

template <int ROWS, int THREADS>
__device__ __forceinline__ void load_tile(double* sdst, const double* g, long long ld, int tid) {
    // ROWS x 16 doubles = ROWS*8 chunks of 16 B
#pragma unroll
    for (int q = 0; q < ROWS * 8 / THREADS; q++) {
        const int c = tid + q * THREADS;
        const int row = c >> 3, ch = c & 7;
        cp_async16(sdst + row * LDS + ch * 2, g + (long long)row * ld + ch * 2);
    }
}

template <int BM>
__global__ void __launch_bounds__(TileCfg<BM>::THREADS, TileCfg<BM>::MIN_CTAS) gemm_nt_kernel(const GpbGemm p) {
    using Cfg = TileCfg<BM>;
    constexpr int STAGES = Cfg::STAGES, THREADS = Cfg::THREADS;
    extern __shared__ __align__(16) double smem[];
    double* As = smem;
    double* Bs = smem + STAGES * Cfg::A_DOUBLES;

    // ---- tile coordinates -------------------------------------------------
    int ti, tj;
    bool diag_tile = false;
    int diag_sub = -1;          // BM = 64: which half (0 = rows 0-63, 1 = rows 64-127) of a diagonal 128-block
    if (p.lower_only) {
        // blockIdx.x enumerates the tiles that touch the lower triangle, by 128-row groups:
        // group q holds (128/BM) row tiles with q+1 column tiles each.
        constexpr int RPG = 128 / BM;
        const int t = blockIdx.x / RPG, sub = blockIdx.x % RPG;
        int i = (int)((sqrt(8.0 * t + 1.0) - 1.0) * 0.5);
        while ((long long)(i + 1) * (i + 2) / 2 <= t) i++;
        while ((long long)i * (i + 1) / 2 > t) i--;
        tj = t - i * (i + 1) / 2;
        ti = i * RPG + sub;
        diag_tile = (tj == i);
        if (diag_tile && RPG == 2) diag_sub = sub;
    } else {
        // longest k-range first: for a lower-triangular A the work grows with the row, for an upper-triangular one it
        // shrinks.  The hardware issues CTAs x-fastest, so with a triangular A the row tile is taken from the SLOW
        // part of the linear CTA index: all column tiles of the longest row first, ... (one matrix, 512 CTAs = 1.7
        // waves of the last trtri level: the long CTAs of the last columns no longer start in the tail)
        if (p.a_tri) {
            const unsigned b = blockIdx.x + gridDim.x * blockIdx.y;
            ti = (int)(b / gridDim.y);
            tj = (int)(b % gridDim.y);
            if (p.a_tri == 1) ti = (int)gridDim.x - 1 - ti;
        } else {
            ti = blockIdx.x;
            tj = blockIdx.y;
        }
    }
    const int m0 = ti * BM, n0 = tj * BN;
    const long long bz = blockIdx.z / p.nb1, bt = blockIdx.z % p.nb1;
    const double* A = p.A + bz * p.sA + bt * p.tA + (long long)m0 * p.lda;
    const double* B = p.B + bz * p.sB + bt * p.tB + (long long)n0 * p.ldb;

    // ---- k-range from the triangular structure ------------------------------
    int kbeg = 0, kend = p.K;
    if (p.a_tri == 1) kend = min(kend, m0 + BM + p.a_off);        // lower: k <= i + off
    else if (p.a_tri == 2) kbeg = max(kbeg, m0 + p.a_off);        // upper: k >= i + off
    if (p.b_tri == 1) kend = min(kend, n0 + BN + p.b_off);
    else if (p.b_tri == 2) kbeg = max(kbeg, n0 + p.b_off);
    kbeg = max(kbeg, 0) & ~(BK - 1);
    kend = min((kend + BK - 1) & ~(BK - 1), p.K);
    const int nk = (kend > kbeg) ? (kend - kbeg) / BK : 0;

    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    const int wm = wid >> 2, wn = wid & 3;          // (BM/32) x 4 warps, warp tile 32 x 32
    const int g = lane >> 2, t = lane & 3;

    // upper-right 64x64 quadrant of a diagonal block: strictly above the diagonal, never needed
    // (its values are the mirror of the lower-left quadrant).  Those warps skip the DMMA work and
    // leave the pipe to the co-resident CTA.
    const bool skip_mma = (diag_sub == 0) && (wn >= 2);

    double acc[4][4][2];
#pragma unroll
    for (int mi = 0; mi < 4; mi++)
#pragma unroll
        for (int ni = 0; ni < 4; ni++) acc[mi][ni][0] = acc[mi][ni][1] = 0.0;

    // ---- prologue -------------------------------------------------------------
#pragma unroll
    for (int s = 0; s < STAGES - 1; s++) {
        if (s < nk) {
            load_tile<BM, THREADS>(As + s * Cfg::A_DOUBLES, A + kbeg + s * BK, p.lda, tid);
            load_tile<BN, THREADS>(Bs + s * Cfg::B_DOUBLES, B + kbeg + s * BK, p.ldb, tid);
        }
        cp_async_commit();
    }

    // ---- main loop --------------------------------------------------------------
    for (int kt = 0; kt < nk; kt++) {
        cp_async_wait<STAGES - 2>();
        __syncthreads();
        {
            const int nx = kt + STAGES - 1;
            if (nx < nk) {
                const int slot = nx % STAGES;
                load_tile<BM, THREADS>(As + slot * Cfg::A_DOUBLES, A + kbeg + nx * BK, p.lda, tid);
                load_tile<BN, THREADS>(Bs + slot * Cfg::B_DOUBLES, B + kbeg + nx * BK, p.ldb, tid);
            }
            cp_async_commit();
        }
        const double* as = As + (kt % STAGES) * Cfg::A_DOUBLES + (wm * 32 + g) * LDS + t;
        const double* bs = Bs + (kt % STAGES) * Cfg::B_DOUBLES + (wn * 32 + g) * LDS + t;
        if (skip_mma) continue;
#pragma unroll
        for (int kk = 0; kk < BK / 4; kk++) {
            double a[4], b[4];
#pragma unroll
            for (int mi = 0; mi < 4; mi++) a[mi] = as[mi * 8 * LDS + kk * 4];
#pragma unroll
            for (int ni = 0; ni < 4; ni++) b[ni] = bs[ni * 8 * LDS + kk * 4];
#pragma unroll
            for (int mi = 0; mi < 4; mi++)
#pragma unroll
                for (int ni = 0; ni < 4; ni++) dmma884(acc[mi][ni][0], acc[mi][ni][1], a[mi], b[ni]);
        }
    }
    cp_async_wait<0>();

    // ---- epilogue -----------------------------------------------------------------
    double* C = p.C + bz * p.sC + bt * p.tC;
    double* Ct = p.Ct ? p.Ct + bz * p.sCt + bt * p.tCt : nullptr;
    // mirrored store: off-diagonal tiles always; inside a diagonal block only the lower-left
    // quadrant (second half-tile, columns 0-63) is mirrored into the skipped upper-right one.
    bool mirror = (Ct != nullptr) && !(p.lower_only && diag_tile && Ct == C);
    if (Ct != nullptr && diag_sub == 1 && wn < 2) mirror = true;
    if (skip_mma) return;
    const double alpha = p.alpha, beta = p.beta;
#pragma unroll
    for (int mi = 0; mi < 4; mi++) {
        const int r = m0 + wm * 32 + mi * 8 + g;
#pragma unroll
        for (int ni = 0; ni < 4; ni++) {
            const int c = n0 + wn * 32 + ni * 8 + 2 * t;
            double2* dst = reinterpret_cast<double2*>(C + (long long)r * p.ldc + c);
            double v0 = alpha * acc[mi][ni][0], v1 = alpha * acc[mi][ni][1];
            if (beta != 0.0) {
                const double2 old = *dst;
                v0 += beta * old.x;
                v1 += beta * old.y;
            }
            *dst = make_double2(v0, v1);
            if (mirror) {
                Ct[(long long)c * p.ldct + r] = v0;
                Ct[(long long)(c + 1) * p.ldct + r] = v1;
            }
        }
    }
}


// ===========================================================================
// TMA + mbarrier variant (default).  Same tile and warp layout as above, but the operand
// panels are staged by the TMA unit (cp.async.bulk.tensor, one elected producer thread)
// into 128B-swizzled shared memory, and consumers synchronise per stage through
// full/empty mbarriers -- there is no CTA-wide barrier in the main loop, so a DMMA warp
// only ever waits for data, never for its siblings.
//   stage = A box {16 k, BM rows} + B box {16 k, 128 rows}, rows of 128 bytes, 16-byte chunk
//   index XOR-ed with (row & 7): fragment loads (8 rows x 4 k per instruction) touch every
//   bank group exactly twice = the 2-wavefront minimum of a 256-byte warp load.
// ===========================================================================
__device__ __forceinline__ unsigned smem_u32(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(unsigned bar, int count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;\n" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(unsigned bar, unsigned bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;\n" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(unsigned bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];\n" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_wait(unsigned bar, unsigned parity) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "WAIT_LOOP:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra WAIT_DONE;\n"
        "bra WAIT_LOOP;\n"
        "WAIT_DONE:\n"
        "}\n" ::"r"(bar), "r"(parity) : "memory");
}
__device__ __forceinline__ void tma_load_2d(unsigned dst, const CUtensorMap* map, int c0, int c1, unsigned bar) {
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];\n"
        ::"r"(dst), "l"(map), "r"(c0), "r"(c1), "r"(bar) : "memory");
}

// Two variants of this kernel were built and measured and are deliberately NOT used:
//   * 112 registers per thread (no spill headroom needed): a sub-partition holds 16384 registers and the
//     2 x 9 warps of two co-resident CTAs put up to 5 warps on one of them, so anything above 96
//     registers silently drops the SM to one resident CTA (458 -> 402 evals/s);
//   * a persistent grid (2 CTAs per SM walking the tile list, stage ring running on across tiles):
//     removes the per-tile pipeline fill, but the resident CTAs then hold every SM slot for the whole
//     launch, so the serial kernels of the other candidate groups (and the look-ahead panel stream)
//     can no longer slip in between tiles, and static tile assignment loses the hardware's dynamic
//     balance: 458 -> 435 evals/s, N=16384 potrf 52.7 -> 57.7 ms.
template <int BM> struct TmaCfg {
    static constexpr int CONSUMERS = BM * 4;               // threads: (BM/32) x 4 warps of 32 x 32
    static constexpr int THREADS = CONSUMERS + 32;         // + one producer warp
    static constexpr int STAGES = (BM == 128) ? 6 : 4;
    static constexpr int A_BYTES = BM * 128, B_BYTES = BN * 128;
    static constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
    static constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + 1024 /* alignment slack */ + 2 * STAGES * 8;
    static constexpr int MIN_CTAS = (BM == 128) ? 1 : 2;
};

// FUSE: 0 = plain GEMM; 1 + kernel kind = lauum with the gradient brackets in the epilogue (GpbLauumFuse, launch.h)
template <int BM, int FUSE>
__device__ __forceinline__ void gemm_nt_tma_body(const GpbGemm& p, const CUtensorMap& mapA, const CUtensorMap& mapB,
                                                 const GpbLauumFuse* fp) {
    using Cfg = TmaCfg<BM>;
    constexpr int STAGES = Cfg::STAGES;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const unsigned base = (smem_u32(smem_raw) + 1023u) & ~1023u;          // swizzle atoms are 1024-byte aligned
    const unsigned char* tiles = smem_raw + (base - smem_u32(smem_raw));  // generic-space view for fragment loads
    const unsigned bar_full = base + STAGES * Cfg::STAGE_BYTES;           // STAGES x 8 bytes
    const unsigned bar_empty = bar_full + STAGES * 8;

    // ---- tile coordinates (identical to the cp.async kernel) -------------------------------
    int ti, tj;
    bool diag_tile = false;
    int diag_sub = -1;
    if (p.lower_only) {
        constexpr int RPG = 128 / BM;
        const int t = blockIdx.x / RPG, sub = blockIdx.x % RPG;
        int i = (int)((sqrt(8.0 * t + 1.0) - 1.0) * 0.5);
        while ((long long)(i + 1) * (i + 2) / 2 <= t) i++;
        while ((long long)i * (i + 1) / 2 > t) i--;
        tj = t - i * (i + 1) / 2;
        ti = i * RPG + sub;
        diag_tile = (tj == i);
        if (diag_tile && RPG == 2) diag_sub = sub;
    } else {
        if (p.a_tri) {
            const unsigned b = blockIdx.x + gridDim.x * blockIdx.y;
            ti = (int)(b / gridDim.y);
            tj = (int)(b % gridDim.y);
            if (p.a_tri == 1) ti = (int)gridDim.x - 1 - ti;
        } else {
            ti = blockIdx.x;
            tj = blockIdx.y;
        }
    }
    const int m0 = ti * BM, n0 = tj * BN;
    const long long bz = blockIdx.z / p.nb1, bt = blockIdx.z % p.nb1;

    int kbeg = 0, kend = p.K;
    if (p.a_tri == 1) kend = min(kend, m0 + BM + p.a_off);
    else if (p.a_tri == 2) kbeg = max(kbeg, m0 + p.a_off);
    if (p.b_tri == 1) kend = min(kend, n0 + BN + p.b_off);
    else if (p.b_tri == 2) kbeg = max(kbeg, n0 + p.b_off);
    kbeg = max(kbeg, 0) & ~(BK - 1);
    kend = min((kend + BK - 1) & ~(BK - 1), p.K);
    const int nk = (kend > kbeg) ? (kend - kbeg) / BK : 0;

    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    if (tid == 0) {
#pragma unroll
        for (int s = 0; s < STAGES; s++) {
            mbar_init(bar_full + s * 8, 1);                       // producer's expect_tx arrival (+ TMA bytes)
            mbar_init(bar_empty + s * 8, Cfg::CONSUMERS / 32);    // one arrival per consumer warp
        }
        asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
    }
    __syncthreads();

    if (wid == Cfg::CONSUMERS / 32) {
        // ===== producer warp: one elected lane drives the TMA unit =====
        if (lane == 0) {
            // batch offsets become (row, column) coordinates of the tensor map (base = p.A / p.B)
            const long long offA = bz * p.sA + bt * p.tA, offB = bz * p.sB + bt * p.tB;
            const int arow = (int)(offA / p.lda) + m0, acol = (int)(offA % p.lda) + kbeg;
            const int brow = (int)(offB / p.ldb) + n0, bcol = (int)(offB % p.ldb) + kbeg;
            for (int kt = 0; kt < nk; kt++) {
                const int s = kt % STAGES;
                const unsigned ph = (unsigned)(kt / STAGES) & 1u;
                mbar_wait(bar_empty + s * 8, ph ^ 1u);            // first round passes immediately
                mbar_expect_tx(bar_full + s * 8, Cfg::STAGE_BYTES);
                const unsigned sa = base + s * Cfg::STAGE_BYTES;
                tma_load_2d(sa, &mapA, acol + kt * BK, arow, bar_full + s * 8);
                tma_load_2d(sa + Cfg::A_BYTES, &mapB, bcol + kt * BK, brow, bar_full + s * 8);
            }
        }
        return;
    }

    // ===== consumer warps =====
    const int wm = wid >> 2, wn = wid & 3;
    const int g = lane >> 2, t = lane & 3;
    // Triangular structure below the tile level: a warp tile (32 x 32) strictly above the diagonal of a
    // lower-only diagonal block is never needed (its mirror image is computed), and triangular operands
    // shorten the k-range of each warp individually.  Skipping warps still walk the barriers; their DMMA
    // slots go to the other warps of the SM.
    const int row0_w = m0 + wm * 32, col0_w = n0 + wn * 32;
    const bool in_diag = p.lower_only && diag_tile && (BM == 64);
    const bool skip_mma = in_diag && (col0_w > row0_w);
    int kb_w = kbeg, ke_w = kend;
    if (p.a_tri == 1) ke_w = min(ke_w, row0_w + 32 + p.a_off);
    else if (p.a_tri == 2) kb_w = max(kb_w, row0_w + p.a_off);
    if (p.b_tri == 1) ke_w = min(ke_w, col0_w + 32 + p.b_off);
    else if (p.b_tri == 2) kb_w = max(kb_w, col0_w + p.b_off);
    const int kt_lo = max(kb_w - kbeg, 0) / BK, kt_hi = min((max(ke_w - kbeg, 0) + BK - 1) / BK, nk);

    double acc[4][4][2];
#pragma unroll
    for (int mi = 0; mi < 4; mi++)
#pragma unroll
        for (int ni = 0; ni < 4; ni++) acc[mi][ni][0] = acc[mi][ni][1] = 0.0;

    // swizzled byte offset of element (row = 8*q + g, k = 4*kk + t) inside a tile: row*128 + sw[kk]
    unsigned sw[4];
#pragma unroll
    for (int kk = 0; kk < 4; kk++) sw[kk] = (unsigned)((((kk * 2 + (t >> 1)) ^ g) << 4) + ((t & 1) << 3));
    const unsigned a_row = (unsigned)((wm * 32 + g) * 128), b_row = (unsigned)(Cfg::A_BYTES + (wn * 32 + g) * 128);

    for (int kt = 0; kt < nk; kt++) {
        const int s = kt % STAGES;
        mbar_wait(bar_full + s * 8, (unsigned)(kt / STAGES) & 1u);
        if (!skip_mma && kt >= kt_lo && kt < kt_hi) {
            const unsigned char* sa = tiles + s * Cfg::STAGE_BYTES + a_row;
            const unsigned char* sb = tiles + s * Cfg::STAGE_BYTES + b_row;
#pragma unroll
            for (int kk = 0; kk < BK / 4; kk++) {
                double a[4], b[4];
#pragma unroll
                for (int mi = 0; mi < 4; mi++) a[mi] = *reinterpret_cast<const double*>(sa + sw[kk] + mi * 1024);
#pragma unroll
                for (int ni = 0; ni < 4; ni++) b[ni] = *reinterpret_cast<const double*>(sb + sw[kk] + ni * 1024);
#pragma unroll
                for (int mi = 0; mi < 4; mi++)
#pragma unroll
                    for (int ni = 0; ni < 4; ni++) dmma884(acc[mi][ni][0], acc[mi][ni][1], a[mi], b[ni]);
            }
        }
        __syncwarp();
        if (lane == 0) mbar_arrive(bar_empty + s * 8);            // this warp is done with the stage
    }

    // ---- epilogue (registers -> global, optional mirrored store) ------------------------------
    double* C = p.C + bz * p.sC + bt * p.tC;
    double* Ct = p.Ct ? p.Ct + bz * p.sCt + bt * p.tCt : nullptr;
    // mirrored store: off-diagonal tiles always; inside a lower-only diagonal block every warp tile strictly
    // below the diagonal is mirrored into the skipped one above it, the diagonal warp tiles only when the
    // mirror is a different matrix
    bool mirror = (Ct != nullptr) && !(p.lower_only && diag_tile && Ct == C);
    if (Ct != nullptr && in_diag && col0_w < row0_w) mirror = true;
    if constexpr (FUSE > 0) {
        // ---- gradient brackets of this tile (all 8 consumer warps take part in the reduction; the producer warp is gone)
        constexpr int KIND = FUSE - 1;
        constexpr int NP = (KIND == GPB_GAUSSIAN) ? 2 : 3;
        constexpr unsigned NEED = (KIND == GPB_GAUSSIAN) ? 0x6u : 0xEu;
        const GpbLauumFuse& f = *fp;
        __shared__ KParams sP;
        __shared__ double sred[8][2 * NP + 2];
        if (f.Pb) {
            const double* src = reinterpret_cast<const double*>(f.Pb + bz);
            double* dstp = reinterpret_cast<double*>(&sP);
            for (int i = tid; i < (int)(sizeof(KParams) / 8); i += Cfg::CONSUMERS) dstp[i] = src[i];
        } else if (tid == 0) {
            sP = f.P;
        }
        asm volatile("bar.sync 1, %0;\n" ::"n"(Cfg::CONSUMERS) : "memory");
        const double* al = f.alpha + bz * f.astride;
        double v[2 * NP + 2];
#pragma unroll
        for (int q = 0; q < 2 * NP + 2; q++) v[q] = 0.0;
        if (!skip_mma) {
#pragma unroll
            for (int mi = 0; mi < 4; mi++) {
                const int r = m0 + wm * 32 + mi * 8 + g;
                if (r >= f.n) continue;
                const double xr = f.x[r], ar = al[r];
#pragma unroll
                for (int ni = 0; ni < 4; ni++) {
#pragma unroll
                    for (int e = 0; e < 2; e++) {
                        const int c = n0 + wn * 32 + ni * 8 + 2 * t + e;
                        if (c > r) continue;
                        // strict lower part counts twice (symmetry), the diagonal once
                        const double wgt = (c < r) ? 2.0 : 1.0;
                        const double k = p.alpha * acc[mi][ni][e] * wgt;
                        const double aw = ar * al[c] * wgt;
                        double u[10];
                        gpb_eval_unique<KIND>(sP, xr - f.x[c], NEED, u);
#pragma unroll
                        for (int q = 0; q < NP; q++) { v[q] += u[1 + q] * aw; v[NP + q] += u[1 + q] * k; }
                        if (c == r) { v[2 * NP] += k; v[2 * NP + 1] += ar * ar; }
                    }
                }
            }
        }
#pragma unroll
        for (int q = 0; q < 2 * NP + 2; q++) {
            const double sq = warp_sum(v[q]);
            if (lane == 0) sred[wid][q] = sq;
        }
        asm volatile("bar.sync 1, %0;\n" ::"n"(Cfg::CONSUMERS) : "memory");
        if (tid < GPB_RED_WIDTH) {
            // partial row: t0[6] | t1[6] | tr | a.a | 0 0
            int q = -1;
            if (tid < NP) q = tid;
            else if (tid >= GPB_RED_MAXS && tid < GPB_RED_MAXS + NP) q = NP + (tid - GPB_RED_MAXS);
            else if (tid == 12) q = 2 * NP;
            else if (tid == 13) q = 2 * NP + 1;
            double sq = 0.0;
            if (q >= 0)
                for (int w = 0; w < Cfg::CONSUMERS / 32; w++) sq += sred[w][q];
            f.partial[((long long)blockIdx.z * gridDim.x + blockIdx.x) * GPB_RED_WIDTH + tid] = sq;
        }
        if (!f.store) return;
    }
    if (skip_mma) return;
    const double alpha = p.alpha, beta = p.beta;
#pragma unroll
    for (int mi = 0; mi < 4; mi++) {
        const int r = m0 + wm * 32 + mi * 8 + g;
#pragma unroll
        for (int ni = 0; ni < 4; ni++) {
            const int c = n0 + wn * 32 + ni * 8 + 2 * t;
            double2* dst = reinterpret_cast<double2*>(C + (long long)r * p.ldc + c);
            double v0 = alpha * acc[mi][ni][0], v1 = alpha * acc[mi][ni][1];
            if (beta != 0.0) {
                const double2 old = *dst;
                v0 += beta * old.x;
                v1 += beta * old.y;
            }
            *dst = make_double2(v0, v1);
            if (mirror) {
                Ct[(long long)c * p.ldct + r] = v0;
                Ct[(long long)(c + 1) * p.ldct + r] = v1;
            }
        }
    }
}

template <int BM>
__global__ void __launch_bounds__(TmaCfg<BM>::THREADS, TmaCfg<BM>::MIN_CTAS)
gemm_nt_tma_kernel(const GpbGemm p, const __grid_constant__ CUtensorMap mapA, const __grid_constant__ CUtensorMap mapB) {
    gemm_nt_tma_body<BM, 0>(p, mapA, mapB, nullptr);
}

// lauum (Ki = V V^T, 64 x 128 tiles) with the gradient brackets in the epilogue: its own instantiation, so the plain
// GEMM keeps its 96 registers
template <int FUSE>
__global__ void __launch_bounds__(TmaCfg<64>::THREADS, TmaCfg<64>::MIN_CTAS)
lauum_grad_tma_kernel(const GpbGemm p, const __grid_constant__ CUtensorMap mapA, const __grid_constant__ CUtensorMap mapB,
                      const __grid_constant__ GpbLauumFuse f) {
    gemm_nt_tma_body<64, FUSE>(p, mapA, mapB, &f);
}

}  // namespace


template <int BM>
static int launch_cfg(const GpbGemm& p, int batch, cudaStream_t st) {
    using Cfg = TileCfg<BM>;
    static bool attr_set = false;
    if (!attr_set) {
        GPB_CUDA(cudaFuncSetAttribute(gemm_nt_kernel<BM>, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::SMEM_BYTES));
        attr_set = true;
    }
    dim3 grid;
    const int tm = p.M / BM, tn = p.N / BN;
    if (p.lower_only) {
        GPB_REQUIRE(p.M == p.N, "lower_only needs a square output");
        grid = dim3((unsigned)((long long)tn * (tn + 1) / 2 * (128 / BM)), 1, (unsigned)(batch * p.nb1));
    } else {
        GPB_REQUIRE(tn <= 65535, "too many column tiles");
        grid = dim3((unsigned)tm, (unsigned)tn, (unsigned)(batch * p.nb1));
    }
    GpbProfScope prof(GPB_KC_GEMM, st);
    gemm_nt_kernel<BM><<<grid, Cfg::THREADS, Cfg::SMEM_BYTES, st>>>(p);
    GPB_LAUNCH_CHECK("gemm_nt_kernel");
    return GPB_OK;
}

// ---- TMA descriptors ---------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static EncodeTiledFn g_encode = nullptr;
static int encode_init() {
    if (g_encode) return GPB_OK;
    void* fn = nullptr;
    cudaDriverEntryPointQueryResult qres;
    GPB_CUDA(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres));
    GPB_REQUIRE(fn != nullptr && qres == cudaDriverEntryPointSuccess, "cuTensorMapEncodeTiled not available in this driver");
    g_encode = (EncodeTiledFn)fn;
    return GPB_OK;
}
// rows x ld view starting at `ptr`; box = {16 doubles (one 128-byte swizzle row), box_rows}
static int make_map(CUtensorMap* m, const double* ptr, long long ld, long long rows, int box_rows) {
    const cuuint64_t dims[2] = {(cuuint64_t)ld, (cuuint64_t)rows};
    const cuuint64_t strides[1] = {(cuuint64_t)ld * 8};
    const cuuint32_t box[2] = {(cuuint32_t)BK, (cuuint32_t)box_rows};
    const cuuint32_t estr[2] = {1, 1};
    const CUresult r = g_encode(m, CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 2, (void*)ptr, dims, strides, box, estr,
                                CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                                CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) {
        gpb_set_error("cuTensorMapEncodeTiled failed (%d): ptr %p ld %lld rows %lld", (int)r, (const void*)ptr, ld, rows);
        return GPB_ERR_CUDA;
    }
    return GPB_OK;
}

template <int BM>
static int launch_tma(const GpbGemm& p, int batch, cudaStream_t st) {
    using Cfg = TmaCfg<BM>;
    static bool attr_set = false;
    if (!attr_set) {
        GPB_CUDA(cudaFuncSetAttribute(gemm_nt_tma_kernel<BM>, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::SMEM_BYTES));
        attr_set = true;
    }
    int stt = encode_init();
    if (stt) return stt;
    // the map's row extent covers the furthest batch member: offsets are turned into (row, column)
    const long long offA = (long long)(batch - 1) * p.sA + (long long)(p.nb1 - 1) * p.tA;
    const long long offB = (long long)(batch - 1) * p.sB + (long long)(p.nb1 - 1) * p.tB;
    CUtensorMap mapA, mapB;
    stt = make_map(&mapA, p.A, p.lda, offA / p.lda + p.M, BM);
    if (stt) return stt;
    stt = make_map(&mapB, p.B, p.ldb, offB / p.ldb + p.N, BN);
    if (stt) return stt;
    dim3 grid;
    const int tm = p.M / BM, tn = p.N / BN;
    if (p.lower_only) grid = dim3((unsigned)((long long)tn * (tn + 1) / 2 * (128 / BM)), 1, (unsigned)(batch * p.nb1));
    else grid = dim3((unsigned)tm, (unsigned)tn, (unsigned)(batch * p.nb1));
    GpbProfScope prof(GPB_KC_GEMM, st);
    gemm_nt_tma_kernel<BM><<<grid, Cfg::THREADS, Cfg::SMEM_BYTES, st>>>(p, mapA, mapB);
    GPB_LAUNCH_CHECK("gemm_nt_tma_kernel");
    return GPB_OK;
}

// `batch` = outer batch count; the grid's z extent is batch * p.nb1
int gpb_launch_gemm(const GpbGemm& p, int batch, cudaStream_t st) {
    GPB_REQUIRE(p.M > 0 && p.N > 0 && p.K >= 0, "empty problem");
    GPB_REQUIRE(p.M % 128 == 0 && p.N % BN == 0 && p.K % BK == 0, "extents must be multiples of the 128x128x16 tile");
    GPB_REQUIRE(p.lda % 2 == 0 && p.ldb % 2 == 0 && p.ldc % 2 == 0, "leading dimensions must be even");
    GPB_REQUIRE(((uintptr_t)p.A % 16 == 0) && ((uintptr_t)p.B % 16 == 0) && ((uintptr_t)p.C % 16 == 0), "operands must be 16-byte aligned");
    GPB_REQUIRE(p.sA % 2 == 0 && p.sB % 2 == 0 && p.sC % 2 == 0, "batch strides must be even");
    GPB_REQUIRE(p.tA % 2 == 0 && p.tB % 2 == 0 && p.tC % 2 == 0 && p.nb1 >= 1, "inner batch strides must be even");
    GPB_REQUIRE(batch >= 1 && (long long)batch * p.nb1 <= 65535, "bad batch");
    if (p.lower_only) GPB_REQUIRE(p.M == p.N, "lower_only needs a square output");
    else GPB_REQUIRE(p.N / BN <= 65535, "too many column tiles");
    int bm = gpb_get_option("gemm_bm");
    if (bm != 64 && bm != 128) bm = 64;     // default: two co-resident 64x128 CTAs per SM
    // operand staging: TMA (default) whenever batch offsets map onto (row, column) coordinates of one
    // tensor map -- always the case for the factorisation's own calls; gemm_impl = 1 forces cp.async
    const long long offA = (long long)(batch - 1) * p.sA + (long long)(p.nb1 - 1) * p.tA;
    const long long offB = (long long)(batch - 1) * p.sB + (long long)(p.nb1 - 1) * p.tB;
    const bool tma_ok = p.K > 0 && p.sA >= 0 && p.tA >= 0 && p.sB >= 0 && p.tB >= 0 &&
                        (offA % p.lda) + p.K <= p.lda && (offB % p.ldb) + p.K <= p.ldb &&
                        offA / p.lda + p.M < (1LL << 31) && offB / p.ldb + p.N < (1LL << 31) && p.lda < (1LL << 31) && p.ldb < (1LL << 31);
    if (tma_ok && gpb_get_option("gemm_impl") != 1)
        return bm == 128 ? launch_tma<128>(p, batch, st) : launch_tma<64>(p, batch, st);
    return bm == 128 ? launch_cfg<128>(p, batch, st) : launch_cfg<64>(p, batch, st);
}

// Measured on B200 (bench.py, same box): fused / separate  headline 452.8 / 461.5 evals/s, C4 19.6 k / 21.1 k evals/s, one GP
// object 3.39 / 3.33 ms.  The separate grad_jac_kernel overlaps the GEMMs of the other evaluator streams and K^-1 stays in
// L2 between the two kernels; in the epilogue the same exp work sits on the SMs the DMMA pipe is waiting for (and the 32
// accumulators spill around it at 96 registers).  So the fused path is OFF unless "lauum_fuse" = 1.
bool gpb_lauum_grad_available() {
    int bm = gpb_get_option("gemm_bm");
    return (bm == 0 || bm == 64) && gpb_get_option("gemm_impl") != 1 && gpb_get_option("lauum_fuse") == 1;
}

int gpb_launch_lauum_grad(const double* V, long long n, long long ldv, long long sV, int batch, double* Ki,
                          long long ldk, long long sK, const GpbLauumFuse& f, double* out16, cudaStream_t st) {
    GPB_REQUIRE(n > 0 && n % GPB_NB == 0, "n must be a positive multiple of 128");
    GPB_REQUIRE(gpb_lauum_grad_available(), "fused lauum not available with these options");
    GPB_REQUIRE(f.kind == GPB_GAUSSIAN || f.kind == GPB_PERIODIC, "unknown kernel kind");
    GPB_REQUIRE(f.x && f.alpha && f.partial && out16 && (Ki || !f.store), "null pointer");
    GPB_REQUIRE(batch >= 1 && batch <= 65535 && ldv % 2 == 0 && sV % 2 == 0 && n < (1LL << 31) && ldv < (1LL << 31), "bad extents");
    using Cfg = TmaCfg<64>;
    static bool attr_set = false;
    if (!attr_set) {
        GPB_CUDA(cudaFuncSetAttribute(lauum_grad_tma_kernel<1 + GPB_GAUSSIAN>, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::SMEM_BYTES));
        GPB_CUDA(cudaFuncSetAttribute(lauum_grad_tma_kernel<1 + GPB_PERIODIC>, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::SMEM_BYTES));
        attr_set = true;
    }
    int stt = encode_init();
    if (stt) return stt;
    GpbGemm g = gpb_gemm_default();
    g.A = V; g.lda = ldv; g.sA = sV;
    g.B = V; g.ldb = ldv; g.sB = sV;
    g.C = Ki; g.ldc = ldk; g.sC = sK;
    g.Ct = Ki; g.ldct = ldk; g.sCt = sK;
    g.M = g.N = g.K = (int)n;
    g.a_tri = 2; g.b_tri = 2; g.lower_only = 1;
    const long long off = (long long)(batch - 1) * sV;
    GPB_REQUIRE((off % ldv) + n <= ldv && off / ldv + n < (1LL << 31), "batch stride does not map onto one tensor map");
    CUtensorMap mapA, mapB;
    stt = make_map(&mapA, V, ldv, off / ldv + n, 64);
    if (stt) return stt;
    stt = make_map(&mapB, V, ldv, off / ldv + n, BN);
    if (stt) return stt;
    const int tn = (int)(n / BN);
    const int nblk = tn * (tn + 1) / 2 * 2;
    dim3 grid((unsigned)nblk, 1, (unsigned)batch);
    {
        GpbProfScope prof(GPB_KC_GEMM, st);
        if (f.kind == GPB_GAUSSIAN) lauum_grad_tma_kernel<1 + GPB_GAUSSIAN><<<grid, Cfg::THREADS, Cfg::SMEM_BYTES, st>>>(g, mapA, mapB, f);
        else lauum_grad_tma_kernel<1 + GPB_PERIODIC><<<grid, Cfg::THREADS, Cfg::SMEM_BYTES, st>>>(g, mapA, mapB, f);
        GPB_LAUNCH_CHECK("lauum_grad_tma_kernel");
    }
    return gpb_launch_sum_partials(f.partial, nblk, batch, out16, st);
}
