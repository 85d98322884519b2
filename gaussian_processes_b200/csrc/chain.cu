// chain.cu -- Cholesky of ONE matrix as a dataflow over 128 x 128 tiles in a single persistent launch.
//
// Why: a single factorisation (the GP-object chain, gp/gp.py:294 -> scipy.linalg.cholesky) has nothing to
// overlap its serial part with.  As a sequence of launches (potrf.cu) every 128-column step costs
// diagonal block -> TRSM -> panel update, three dependent kernels of 13-56 us that each use a fraction of
// the GPU; at N = 4096 that chain is 2.3 of the 3.1 ms while the O(N^3) work needs 0.6 ms at the DMMA peak.
//
// Here ONE cooperative launch holds one CTA per SM for the whole factorisation:
//   CTA 0, the chain:   for k = 0..T-1:  [k > 0: L(k,k-1) = A(k,k-1) W_{k-1}^T ;  A(k,k) -= L(k,k-1) L(k,k-1)^T]
//                       then L_kk = chol(A_kk), W_kk = L_kk^-1, V_kk = W_kk^T          (diag_block.cuh)
//                       -- the critical path POTRF(k) -> TRSM(k+1,k) -> SYRK(k+1,k+1) -> POTRF(k+1) never
//                       leaves the SM and never waits for a kernel boundary;
//   CTAs 1.., workers:  every other tile task, statically owned (tile (i,j) -> one worker, so the updates of a
//                       tile are applied in step order without any lock):
//                         TRSM(i,k)   L(i,k) = A(i,k) W_kk^T                  i >= k+2   needs DIAG[k]
//                         UPD(i,j,k)  A(i,j) -= L(i,k) L(j,k)^T   i >= j >= k+1, not (k+1,k+1)
//                                                                             needs LREADY[i][k], LREADY[j][k]
//   Dependencies travel through release/acquire flags in global memory (DIAG[k], LREADY[i][k], CNT[i][j] =
//   number of updates tile (i,j) has received).  Every CTA works through its tasks in one global
//   topological order (by step k, then column), taking a TRSM of its own as soon as it is runnable, so the
//   earliest unfinished task is always runnable by its owner: no deadlock; all CTAs are co-resident
//   (cooperative launch), spin loops carry a clock-based time-out that aborts the launch through an error flag.
//
// Tile product: 128 x 128 x 128 on DMMA.8x8x4, 16 warps of 32 x 32, operands staged by a 4-stage cp.async.cg
// ring (L2 -> shared memory, never through L1: tiles are produced by other SMs).  The warp -> (row, column)
// block map is skewed so that every SM sub-partition holds all four column blocks: triangular skipping
// (TRSM: k <= column) then shortens every sub-partition's DMMA queue alike.
#include <map>
#include <mutex>
#include <vector>
#include "common.cuh"
#include "launch.h"
#include "diag_block.cuh"

namespace {

constexpr int CT = GPB_NB;                  // tile edge
constexpr int CBK = 16;                     // k-chunk of the ring
constexpr int CLD = CBK + 4;                // padded row stride (doubles): conflict-free 8-byte fragment loads
constexpr int CSTAGES = 4;
constexpr int CTHREADS = 512;
constexpr int C_OPER = CT * CLD;            // doubles per operand chunk
constexpr int C_RING_BYTES = CSTAGES * 2 * C_OPER * 8;          // 163840
constexpr int C_SMEM_0 = (C_RING_BYTES > DIAG_SMEM) ? C_RING_BYTES : DIAG_SMEM;
constexpr int C_STRIP_MAX = (GPB_NB + 32) * (GPB_NB + 4) * 8;      // second operand + a 32-row strip
constexpr int C_SMEM = (C_SMEM_0 > C_STRIP_MAX) ? C_SMEM_0 : C_STRIP_MAX;
constexpr long long C_TIMEOUT = 4000000000LL;                   // ~2 s of SM clocks

enum { TASK_TRSM = 0, TASK_UPD = 1 };
struct ChainTask { short type, i, j, k; };

// lower 32x32 blocks of a diagonal tile spread over the sub-partitions (warp w runs on sub-partition w & 3)
__constant__ signed char c_diag_m[16] = {0, 1, 1, 2, 2, 2, 3, 3, 3, 3, -1, -1, -1, -1, -1, -1};
__constant__ signed char c_diag_n[16] = {0, 0, 1, 0, 1, 2, 0, 1, 2, 3, 0, 0, 0, 0, 0, 0};

struct ChainArgs {
    double* A; long long ld;
    double* W; long long ldw;
    double* V; long long ldv;
    int* info;
    int T; int n_valid;
    int NG;                     // CTAs of the chain group (CTA 0 = the chain, 1..NG-1 its helpers)
    int* flags;                 // [0] error | DIAG[T] | LREADY[T*T] | CNT[T*T] | TP[T] | SP[T]
    const ChainTask* bulk; const int* bulk_off;     // per CTA: UPD tasks in (k, j, i) order
    const ChainTask* trsm; const int* trsm_off;     // per CTA: TRSM tasks in (k, i) order
    long long* clk;             // [T][8] phase clocks of the chain CTA (gpb_debug_chain_clocks)
    long long* wclk;            // [grid][4] per CTA: cycles waiting | in TRSM tasks | in UPD tasks | tasks
};
#define CHAIN_STAMP(k, p) do { if (tid == 0) a.clk[(k) * 8 + (p)] = clock64(); } while (0)

__device__ __forceinline__ int* f_diag(const ChainArgs& a, int k) { return a.flags + 1 + k; }
__device__ __forceinline__ int* f_lready(const ChainArgs& a, int i, int k) { return a.flags + 1 + a.T + i * a.T + k; }
__device__ __forceinline__ int* f_cnt(const ChainArgs& a, int i, int j) { return a.flags + 1 + a.T + a.T * a.T + i * a.T + j; }
__device__ __forceinline__ int* f_tp(const ChainArgs& a, int k) { return a.flags + 1 + a.T + 2 * a.T * a.T + k; }
__device__ __forceinline__ int* f_sp(const ChainArgs& a, int k) { return a.flags + 1 + 2 * a.T + 2 * a.T * a.T + k; }
__host__ __device__ inline size_t chain_flag_words(int T) { return 1 + 3 * (size_t)T + 2 * (size_t)T * T; }

// thread 0 only: spin until *flag >= target; false on time-out or when another CTA raised the error flag
__device__ __forceinline__ bool spin_ge(const int* flag, int target, int* err) {
    if (ld_acquire(flag) >= target) return true;
    const long long t0 = clock64();
    unsigned it = 0;
    while (ld_acquire(flag) < target) {
        if ((++it & 255u) == 0) {
            if (ld_acquire(err) != 0) return false;
            if (clock64() - t0 > C_TIMEOUT) { atomicExch(err, 2); return false; }
        }
    }
    return true;
}

template <int ROWS>
__device__ __forceinline__ void ring_load(double* sdst, const double* g, long long ld, int tid) {
#pragma unroll
    for (int q = 0; q < ROWS * 8 / CTHREADS; q++) {
        const int c = tid + q * CTHREADS;
        const int row = c >> 3, ch = c & 7;
        cp_async16(sdst + row * CLD + ch * 2, g + (long long)row * ld + ch * 2);
    }
}

// acc (+)= sum_k Ag[r, k] * Bg[c, k] over K = 128 for this warp's 32 x 32 block (wm, wn); the warp multiplies
// only k-chunks [0, kt_hi) (triangular B), but every warp walks all chunks (loads and barriers are CTA-wide).
__device__ __forceinline__ void tile_mm(double (&acc)[4][4][2], const double* Ag, long long lda, const double* Bg,
                                        long long ldb, int wm, int wn, int kt_hi, double* ring, int tid) {
    constexpr int NK = CT / CBK;
    const int lane = tid & 31, g = lane >> 2, t = lane & 3;
#pragma unroll
    for (int s = 0; s < CSTAGES - 1; s++) {
        ring_load<CT>(ring + s * 2 * C_OPER, Ag + s * CBK, lda, tid);
        ring_load<CT>(ring + s * 2 * C_OPER + C_OPER, Bg + s * CBK, ldb, tid);
        cp_async_commit();
    }
    for (int kt = 0; kt < NK; kt++) {
        cp_async_wait<CSTAGES - 2>();
        __syncthreads();
        {
            const int nx = kt + CSTAGES - 1;
            if (nx < NK) {
                const int slot = nx % CSTAGES;
                ring_load<CT>(ring + slot * 2 * C_OPER, Ag + nx * CBK, lda, tid);
                ring_load<CT>(ring + slot * 2 * C_OPER + C_OPER, Bg + nx * CBK, ldb, tid);
            }
            cp_async_commit();
        }
        if (kt >= kt_hi) continue;
        const double* as = ring + (kt % CSTAGES) * 2 * C_OPER + (wm * 32 + g) * CLD + t;
        const double* bs = ring + (kt % CSTAGES) * 2 * C_OPER + C_OPER + (wn * 32 + g) * CLD + t;
#pragma unroll
        for (int kk = 0; kk < CBK / 4; kk++) {
            double a[4], b[4];
#pragma unroll
            for (int mi = 0; mi < 4; mi++) a[mi] = as[mi * 8 * CLD + kk * 4];
#pragma unroll
            for (int ni = 0; ni < 4; ni++) b[ni] = bs[ni * 8 * CLD + kk * 4];
#pragma unroll
            for (int mi = 0; mi < 4; mi++)
#pragma unroll
                for (int ni = 0; ni < 4; ni++) dmma884(acc[mi][ni][0], acc[mi][ni][1], a[mi], b[ni]);
        }
    }
    cp_async_wait<0>();
    __syncthreads();                 // the ring may be reused (next task / the diagonal-block body)
}

// L(i,k) = A(i,k) W_kk^T, in place.  W_kk is lower triangular: column block wn needs k < 32 (wn + 1).
__device__ __forceinline__ void task_trsm(const ChainArgs& a, int i, int k, double* ring, int tid) {
    const int wid = tid >> 5, lane = tid & 31, g = lane >> 2, t = lane & 3;
    const int wm = wid >> 2, wn = (wid + wm) & 3;
    double* At = a.A + (long long)i * CT * a.ld + (long long)k * CT;
    const double* Wk = a.W + (long long)k * CT * a.ldw + (long long)k * CT;
    double acc[4][4][2];
#pragma unroll
    for (int mi = 0; mi < 4; mi++)
#pragma unroll
        for (int ni = 0; ni < 4; ni++) acc[mi][ni][0] = acc[mi][ni][1] = 0.0;
    tile_mm(acc, At, a.ld, Wk, a.ldw, wm, wn, (wn + 1) * 2, ring, tid);
    // every thread's loads of the tile are complete (barrier at the end of tile_mm): safe to overwrite
#pragma unroll
    for (int mi = 0; mi < 4; mi++)
#pragma unroll
        for (int ni = 0; ni < 4; ni++) {
            double* dst = At + (long long)(wm * 32 + mi * 8 + g) * a.ld + wn * 32 + ni * 8 + 2 * t;
            *reinterpret_cast<double2*>(dst) = make_double2(acc[mi][ni][0], acc[mi][ni][1]);
        }
}

// A(i,j) -= L(i,k) L(j,k)^T.  Diagonal tiles: only the 32x32 blocks on or below the diagonal.
__device__ __forceinline__ void task_upd(const ChainArgs& a, int i, int j, int k, double* ring, int tid) {
    const int wid = tid >> 5, lane = tid & 31, g = lane >> 2, t = lane & 3;
    int wm, wn;
    bool active = true;
    if (i == j) {
        wm = c_diag_m[wid]; wn = c_diag_n[wid];
        if (wm < 0) { active = false; wm = 0; wn = 0; }
    } else {
        wm = wid >> 2; wn = (wid + wm) & 3;
    }
    const double* Ai = a.A + (long long)i * CT * a.ld + (long long)k * CT;
    const double* Aj = a.A + (long long)j * CT * a.ld + (long long)k * CT;
    double* Ct = a.A + (long long)i * CT * a.ld + (long long)j * CT;
    // accumulators start at -C: the loads overlap the ring's prologue, the result is -(acc)
    double acc[4][4][2];
#pragma unroll
    for (int mi = 0; mi < 4; mi++)
#pragma unroll
        for (int ni = 0; ni < 4; ni++) {
            if (active) {
                const double2 c = __ldcg(reinterpret_cast<const double2*>(
                    Ct + (long long)(wm * 32 + mi * 8 + g) * a.ld + wn * 32 + ni * 8 + 2 * t));
                acc[mi][ni][0] = -c.x; acc[mi][ni][1] = -c.y;
            } else {
                acc[mi][ni][0] = acc[mi][ni][1] = 0.0;
            }
        }
    tile_mm(acc, Ai, a.ld, Aj, a.ld, wm, wn, active ? CT / CBK : 0, ring, tid);
    if (!active) return;
#pragma unroll
    for (int mi = 0; mi < 4; mi++)
#pragma unroll
        for (int ni = 0; ni < 4; ni++) {
            double* dst = Ct + (long long)(wm * 32 + mi * 8 + g) * a.ld + wn * 32 + ni * 8 + 2 * t;
            *reinterpret_cast<double2*>(dst) = make_double2(-acc[mi][ni][0], -acc[mi][ni][1]);
        }
}

// make this CTA's global stores visible, then publish
// (stores of all threads -> CTA barrier -> one gpu-scope fence + release store: cumulativity carries the
//  other threads' stores; a fence per thread costs ~1k cycles more)
__device__ __forceinline__ void publish(int* flag, int value, int tid) {
    __syncthreads();
    if (tid == 0) {
        __threadfence();
        st_release(flag, value);
    }
}

// ===========================================================================
// The critical path between two diagonal blocks, TRSM(k,k-1) then SYRK(k,k), is 0.625 * 2 * 128^3 * 2 flop:
// 21 us on ONE SM at the DMMA peak -- as long as the diagonal block itself.  It is therefore split over
// the NG CTAs of the chain group by rows: each takes a strip of CR = 128 / NG rows of the tile, holds the
// whole second operand in shared memory (one load phase, no ring: the products are latency-bound, K = 128
// with two accumulator tiles per warp) and the group meets through two arrival counters per step.
// ===========================================================================
constexpr int CSLD = CT + 4;                     // strip operand row stride (doubles)

// rows [r0, r0 + CR) of  X = A(k,k-1) W_{k-1}^T  (in place); X strip also left in shared memory (As)
template <int CR>
__device__ __forceinline__ void strip_trsm(const ChainArgs& a, int k, int r0, double* Bs, double* As, int tid) {
    const int wid = tid >> 5, lane = tid & 31, g = lane >> 2, t = lane & 3;
    double* At = a.A + ((long long)k * CT + r0) * a.ld + (long long)(k - 1) * CT;
    const double* Wk = a.W + (long long)(k - 1) * CT * a.ldw + (long long)(k - 1) * CT;
    // W rows c, columns [0, 32 * (c / 32 + 1)): 16-byte pieces
    for (int e = tid; e < CT * (CT / 2); e += CTHREADS) {
        const int c = e >> 6, p2 = (e & 63) * 2;
        if (p2 < ((c >> 5) + 1) * 32) cp_async16(Bs + c * CSLD + p2, Wk + (long long)c * a.ldw + p2);
    }
    for (int e = tid; e < CR * (CT / 2); e += CTHREADS) {
        const int r = e >> 6, p2 = (e & 63) * 2;
        cp_async16(As + r * CSLD + p2, At + (long long)r * a.ld + p2);
    }
    cp_async_commit();
    cp_async_wait<0>();
    __syncthreads();
    // warp w: columns [8w, 8w + 8), all CR rows; k < 8w + 8
    double acc[CR / 8][2];
#pragma unroll
    for (int mi = 0; mi < CR / 8; mi++) acc[mi][0] = acc[mi][1] = 0.0;
    const double* ap = As + g * CSLD + t;
    const double* bp = Bs + (wid * 8 + g) * CSLD + t;
    const int nkk = 2 * wid + 2;
#pragma unroll 4
    for (int kk = 0; kk < nkk; kk++) {
        const double b = bp[kk * 4];
#pragma unroll
        for (int mi = 0; mi < CR / 8; mi++) dmma884(acc[mi][0], acc[mi][1], ap[mi * 8 * CSLD + kk * 4], b);
    }
    __syncthreads();                 // every warp has read the strip: overwrite it (shared and global)
#pragma unroll
    for (int mi = 0; mi < CR / 8; mi++) {
        const int r = mi * 8 + g, c = wid * 8 + 2 * t;
        *reinterpret_cast<double2*>(At + (long long)r * a.ld + c) = make_double2(acc[mi][0], acc[mi][1]);
        *reinterpret_cast<double2*>(As + r * CSLD + c) = make_double2(acc[mi][0], acc[mi][1]);
    }
}

// rows [r0, r0 + CR) of  A(k,k) -= X X^T, columns up to the end of the 32-block that holds the diagonal
template <int CR>
__device__ __forceinline__ void strip_syrk(const ChainArgs& a, int k, int r0, double* Bs, const double* As, int tid) {
    const int wid = tid >> 5, lane = tid & 31, g = lane >> 2, t = lane & 3;
    const double* Xt = a.A + (long long)k * CT * a.ld + (long long)(k - 1) * CT;
    double* Ct = a.A + ((long long)k * CT + r0) * a.ld + (long long)k * CT;
    const int ncol = ((r0 + CR + 31) >> 5) << 5;          // columns (= rows of X) needed
    for (int e = tid; e < ncol * (CT / 2); e += CTHREADS) {
        const int c = e >> 6, p2 = (e & 63) * 2;
        cp_async16(Bs + c * CSLD + p2, Xt + (long long)c * a.ld + p2);
    }
    cp_async_commit();
    const bool active = wid * 8 < ncol;
    double acc[CR / 8][2];
#pragma unroll
    for (int mi = 0; mi < CR / 8; mi++) {
        if (active) {
            const double2 c = __ldcg(reinterpret_cast<const double2*>(Ct + (long long)(mi * 8 + g) * a.ld + wid * 8 + 2 * t));
            acc[mi][0] = -c.x; acc[mi][1] = -c.y;
        } else {
            acc[mi][0] = acc[mi][1] = 0.0;
        }
    }
    cp_async_wait<0>();
    __syncthreads();
    if (!active) return;
    const double* ap = As + g * CSLD + t;
    const double* bp = Bs + (wid * 8 + g) * CSLD + t;
#pragma unroll 4
    for (int kk = 0; kk < CT / 4; kk++) {
        const double b = bp[kk * 4];
#pragma unroll
        for (int mi = 0; mi < CR / 8; mi++) dmma884(acc[mi][0], acc[mi][1], ap[mi * 8 * CSLD + kk * 4], b);
    }
#pragma unroll
    for (int mi = 0; mi < CR / 8; mi++)
        *reinterpret_cast<double2*>(Ct + (long long)(mi * 8 + g) * a.ld + wid * 8 + 2 * t) =
            make_double2(-acc[mi][0], -acc[mi][1]);
}

// this CTA's part is stored: count it; the last of `total` arrivals publishes `done_flag` (optional)
__device__ __forceinline__ void arrive(int* counter, int total, int* done_flag, int tid) {
    __syncthreads();
    if (tid == 0) {
        __threadfence();
        const int old = atomicAdd(counter, 1);
        if (done_flag && old == total - 1) {
            __threadfence();
            st_release(done_flag, 1);
        }
    }
}

template <int CR>
__device__ __forceinline__ void chain_group(const ChainArgs& a, double* csm, int* s_act, int tid) {
    int* err = a.flags;
    const int c = blockIdx.x, NG = a.NG, T = a.T;
    double* Bs = csm;
    double* As = csm + CT * CSLD;
    for (int k = 0; k < T; k++) {
        if (c == 0) CHAIN_STAMP(k, 0);
        if (k > 0) {
            if (tid == 0)
                s_act[0] = (spin_ge(f_diag(a, k - 1), 1, err) && spin_ge(f_cnt(a, k, k - 1), k - 1, err)) ? 1 : 0;
            __syncthreads();
            if (!s_act[0]) { if (tid == 0 && c == 0) *a.info = -999; return; }
            if (c == 0) CHAIN_STAMP(k, 1);
            strip_trsm<CR>(a, k, c * CR, Bs, As, tid);
            if (c == 0) CHAIN_STAMP(k, 2);
            arrive(f_tp(a, k), NG, f_lready(a, k, k - 1), tid);
            if (c == 0) CHAIN_STAMP(k, 3);
            if (tid == 0)
                s_act[1] = (spin_ge(f_tp(a, k), NG, err) && spin_ge(f_cnt(a, k, k), k - 1, err)) ? 1 : 0;
            __syncthreads();
            if (!s_act[1]) { if (tid == 0 && c == 0) *a.info = -999; return; }
            if (c == 0) CHAIN_STAMP(k, 4);
            strip_syrk<CR>(a, k, c * CR, Bs, As, tid);
            arrive(f_sp(a, k), NG, nullptr, tid);
        }
        if (c != 0) continue;
        if (k > 0) {
            if (tid == 0) s_act[2] = spin_ge(f_sp(a, k), NG, err) ? 1 : 0;
            __syncthreads();
            if (!s_act[2]) { if (tid == 0) *a.info = -999; return; }
        }
        CHAIN_STAMP(k, 5);
        if (tid < 256) {
            const long long o = (long long)k * CT;
            const long long valid = (long long)a.n_valid - o;
            const int nsub = valid >= CT ? 4 : (valid <= 0 ? 0 : (int)((valid + SB - 1) / SB));
            diag_block_body<true>(a.A + o * a.ld + o, a.ld, a.W + o * a.ldw + o, a.ldw,
                                  a.V ? a.V + o * a.ldv + o : nullptr, a.ldv, a.info, (int)o, nsub, csm);
        }
        CHAIN_STAMP(k, 6);
        publish(f_diag(a, k), 1, tid);
        CHAIN_STAMP(k, 7);
    }
}

__global__ void __launch_bounds__(CTHREADS, 1) potrf_dataflow_kernel(const ChainArgs a) {
    extern __shared__ __align__(16) double csm[];
    __shared__ int s_act[4];
    const int tid = threadIdx.x;
    int* err = a.flags;
    const int T = a.T;

    if ((int)blockIdx.x < a.NG) {
        // ===================== the chain group =====================
        if (a.NG == 8) chain_group<16>(a, csm, s_act, tid);
        else chain_group<32>(a, csm, s_act, tid);
        return;
    }

    // ===================== workers =====================
    int bt = a.bulk_off[blockIdx.x];
    const int bend = a.bulk_off[blockIdx.x + 1];
    int qt = a.trsm_off[blockIdx.x];
    const int qend = a.trsm_off[blockIdx.x + 1];
    long long w_wait = 0, w_trsm = 0, w_upd = 0, w_n = 0;
    while (bt < bend || qt < qend) {
        const long long tw0 = clock64();
        if (tid == 0) {
            // next action: an own TRSM as soon as it is runnable (it unblocks other CTAs), else the next
            // update in order; spin on both conditions until one holds
            int act = -1;
            const long long t0 = clock64();
            unsigned it = 0;
            for (;;) {
                if (qt < qend) {
                    const ChainTask r = a.trsm[qt];
                    if (ld_acquire(f_cnt(a, r.i, r.k)) >= r.k && ld_acquire(f_diag(a, r.k)) != 0) { act = 0; break; }
                }
                if (bt < bend) {
                    const ChainTask u = a.bulk[bt];
                    if (ld_acquire(f_lready(a, u.i, u.k)) != 0 && ld_acquire(f_lready(a, u.j, u.k)) != 0) { act = 1; break; }
                }
                if ((++it & 63u) == 0) {
                    if (ld_acquire(err) != 0) break;
                    if (clock64() - t0 > C_TIMEOUT) { atomicExch(err, 3); break; }
                }
            }
            s_act[0] = act;
        }
        __syncthreads();
        const int act = s_act[0];
        if (act < 0) return;
        const long long tw1 = clock64();
        w_wait += tw1 - tw0;
        if (act == 0) {
            const ChainTask r = a.trsm[qt++];
            task_trsm(a, r.i, r.k, csm, tid);
            publish(f_lready(a, r.i, r.k), 1, tid);
            w_trsm += clock64() - tw1;
        } else {
            const ChainTask u = a.bulk[bt++];
            task_upd(a, u.i, u.j, u.k, csm, tid);
            publish(f_cnt(a, u.i, u.j), u.k + 1, tid);
            w_upd += clock64() - tw1;
        }
        w_n++;
    }
    if (tid == 0) {
        a.wclk[blockIdx.x * 4 + 0] = w_wait; a.wclk[blockIdx.x * 4 + 1] = w_trsm;
        a.wclk[blockIdx.x * 4 + 2] = w_upd; a.wclk[blockIdx.x * 4 + 3] = w_n;
    }
}

// ---- host side: task lists per (T, grid) ------------------------------------------------------------
struct ChainPlan {
    ChainTask* bulk = nullptr; int* bulk_off = nullptr;
    ChainTask* trsm = nullptr; int* trsm_off = nullptr;
};
std::map<std::pair<std::pair<int, int>, int>, ChainPlan> g_plans;
std::map<cudaStream_t, std::pair<int*, size_t>> g_flag_pool;
std::map<cudaStream_t, int> g_last_T;
std::mutex g_plan_mu;

int build_plan(int T, int G, int NG, ChainPlan* out) {
    std::lock_guard<std::mutex> lk(g_plan_mu);
    auto it = g_plans.find({{T, G}, NG});
    if (it != g_plans.end()) { *out = it->second; return GPB_OK; }
    const int nw = G - NG;
    // Tile (i,j) receives j updates (steps 0..j-1), and at step k exactly the tiles with j > k are live.
    // Dealing the tiles out in order of decreasing j (boustrophedon over the workers) therefore balances
    // every live set, i.e. every step of the factorisation, to within one tile per worker.
    std::vector<int> own((size_t)T * T, NG);
    {
        long long n = 0;
        for (int j = T - 1; j >= 0; j--)
            for (int i = j; i < T; i++) {
                const long long round = n / nw, pos = n % nw;
                own[(size_t)i * T + j] = NG + (int)((round & 1) ? nw - 1 - pos : pos);
                n++;
            }
    }
    auto owner = [&](int i, int j) { return own[(size_t)i * T + j]; };
    std::vector<std::vector<ChainTask>> bulk(G), trsm(G);
    for (int k = 0; k + 1 < T; k++) {
        for (int i = k + 2; i < T; i++) trsm[owner(i, k)].push_back({TASK_TRSM, (short)i, (short)k, (short)k});
        for (int j = k + 1; j < T; j++)
            for (int i = j; i < T; i++) {
                if (i == k + 1 && j == k + 1) continue;            // the chain's own update
                bulk[owner(i, j)].push_back({TASK_UPD, (short)i, (short)j, (short)k});
            }
    }
    auto upload = [&](std::vector<std::vector<ChainTask>>& lists, ChainTask** dt, int** doff) -> int {
        std::vector<ChainTask> flat;
        std::vector<int> off(G + 1, 0);
        for (int g = 0; g < G; g++) {
            off[g] = (int)flat.size();
            flat.insert(flat.end(), lists[g].begin(), lists[g].end());
        }
        off[G] = (int)flat.size();
        GPB_CUDA(cudaMalloc(dt, (flat.size() + 1) * sizeof(ChainTask)));
        GPB_CUDA(cudaMalloc(doff, (G + 1) * sizeof(int)));
        if (!flat.empty()) GPB_CUDA(cudaMemcpy(*dt, flat.data(), flat.size() * sizeof(ChainTask), cudaMemcpyHostToDevice));
        GPB_CUDA(cudaMemcpy(*doff, off.data(), (G + 1) * sizeof(int), cudaMemcpyHostToDevice));
        return GPB_OK;
    };
    ChainPlan p;
    int stt = upload(bulk, &p.bulk, &p.bulk_off);
    if (stt) return stt;
    stt = upload(trsm, &p.trsm, &p.trsm_off);
    if (stt) return stt;
    g_plans[{{T, G}, NG}] = p;
    *out = p;
    return GPB_OK;
}

}  // namespace

// The largest T the dataflow factorisation takes (N = 8192): beyond it the trailing matrix no longer
// lives in L2 and the K = 128 updates re-stream it from HBM every step; potrf.cu's look-ahead panels win.
bool gpb_potrf_dataflow_ok(long long n, int batch) {
    const long long T = n / GPB_NB;
    return batch == 1 && T >= 2 && T <= 64;
}

// zero_blocks and info initialisation are the caller's (gpb_launch_potrf) business
int gpb_launch_potrf_dataflow(double* A, long long n, long long ld, double* W, long long ldw, double* V,
                              long long ldv, int* info, cudaStream_t st, long long n_valid) {
    const int T = (int)(n / GPB_NB);
    static int num_sms = 0;
    static bool attr_set = false;
    if (!attr_set) {
        int dev = 0, coop = 0;
        GPB_CUDA(cudaGetDevice(&dev));
        GPB_CUDA(cudaDeviceGetAttribute(&num_sms, cudaDevAttrMultiProcessorCount, dev));
        GPB_CUDA(cudaDeviceGetAttribute(&coop, cudaDevAttrCooperativeLaunch, dev));
        GPB_REQUIRE(coop != 0, "device does not support cooperative launches");
        GPB_CUDA(cudaFuncSetAttribute(potrf_dataflow_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, C_SMEM));
        attr_set = true;
    }
    const long long ntiles = (long long)T * (T + 1) / 2;
    int NG = gpb_get_option("chain_group");          // CTAs sharing the critical path (4 or 8)
    if (NG != 4 && NG != 8) NG = 8;
    GPB_REQUIRE(num_sms >= 2 * NG, "device too small for the dataflow factorisation");
    int G = (int)((ntiles + NG < (long long)num_sms) ? ntiles + NG : num_sms);
    if (G < NG + 1) G = NG + 1;
    ChainPlan plan;
    int stt = build_plan(T, G, NG, &plan);
    if (stt) return stt;
    // flag words: one grow-only set per stream (two factorisations on one stream are serialised anyway)
    // (+ the chain's phase clocks, 8 per step, behind the flags)
    const size_t nfl = (chain_flag_words(T) + 1) / 2 * 2 + (size_t)T * 16 + (size_t)num_sms * 8;
    int* flags = nullptr;
    {
        std::lock_guard<std::mutex> lk(g_plan_mu);
        auto& e = g_flag_pool[st];
        if (nfl > e.second) {
            if (e.first) GPB_CUDA(cudaFree(e.first));
            e.first = nullptr; e.second = 0;
            GPB_CUDA(cudaMalloc(&e.first, nfl * sizeof(int)));
            e.second = nfl;
        }
        flags = e.first;
        g_last_T[st] = T;
    }
    GPB_CUDA(cudaMemsetAsync(flags, 0, nfl * sizeof(int), st));
    ChainArgs a;
    a.A = A; a.ld = ld; a.W = W; a.ldw = ldw; a.V = V; a.ldv = ldv; a.info = info;
    a.T = T; a.n_valid = (int)n_valid; a.flags = flags;
    a.bulk = plan.bulk; a.bulk_off = plan.bulk_off; a.trsm = plan.trsm; a.trsm_off = plan.trsm_off;
    a.clk = reinterpret_cast<long long*>(flags + (chain_flag_words(T) + 1) / 2 * 2);
    a.wclk = a.clk + (size_t)T * 8;
    a.NG = NG;
    void* args[] = {(void*)&a};
    GpbProfScope prof(GPB_KC_GEMM, st);
    GPB_CUDA(cudaLaunchCooperativeKernel((const void*)potrf_dataflow_kernel, dim3((unsigned)G), dim3(CTHREADS), args,
                                         (size_t)C_SMEM, st));
    GPB_LAUNCH_CHECK("potrf_dataflow_kernel");
    return GPB_OK;
}

// phase clocks (SM cycles) of the chain CTA in the last dataflow factorisation on `stream`:
// out[k*8 + p], p = 0 step start | 1 sub-diagonal tile ready | 2 TRSM done | 3 published | 4 diagonal tile
// ready | 5 SYRK done | 6 diagonal block done | 7 published.  Returns T (steps) or < 0.
extern "C" int gpb_debug_chain_clocks(void* stream, long long* out, int max_steps) {
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
    int* flags = nullptr;
    int T = 0;
    {
        std::lock_guard<std::mutex> lk(g_plan_mu);
        auto it = g_flag_pool.find(st);
        if (it == g_flag_pool.end() || !g_last_T.count(st)) { gpb_set_error("no dataflow factorisation ran on this stream"); return GPB_ERR_ARG; }
        flags = it->second.first;
        T = g_last_T[st];
    }
    GPB_CUDA(cudaStreamSynchronize(st));
    const int n = T < max_steps ? T : max_steps;
    const long long* clk = reinterpret_cast<const long long*>(flags + (chain_flag_words(T) + 1) / 2 * 2);
    GPB_CUDA(cudaMemcpy(out, clk, (size_t)n * 8 * sizeof(long long), cudaMemcpyDeviceToHost));
    return T;
}

// per-CTA accounting of the same launch: out[cta*4 + {0 cycles waiting, 1 in TRSM tasks, 2 in UPD tasks, 3 tasks}]
extern "C" int gpb_debug_chain_workers(void* stream, long long* out, int max_ctas) {
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
    int* flags = nullptr;
    int T = 0;
    {
        std::lock_guard<std::mutex> lk(g_plan_mu);
        auto it = g_flag_pool.find(st);
        if (it == g_flag_pool.end() || !g_last_T.count(st)) { gpb_set_error("no dataflow factorisation ran on this stream"); return GPB_ERR_ARG; }
        flags = it->second.first;
        T = g_last_T[st];
    }
    GPB_CUDA(cudaStreamSynchronize(st));
    const long long* clk = reinterpret_cast<const long long*>(flags + (chain_flag_words(T) + 1) / 2 * 2) + (size_t)T * 8;
    GPB_CUDA(cudaMemcpy(out, clk, (size_t)max_ctas * 4 * sizeof(long long), cudaMemcpyDeviceToHost));
    return T;
}
