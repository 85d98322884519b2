// chain.cu -- Cholesky of ONE matrix as a dataflow over 128 x 128 tiles in a single persistent launch.
//
// Why: a single factorisation (the GP-object chain, gp/gp.py:294 -> scipy.linalg.cholesky) has nothing to
// overlap its serial part with.  As a sequence of launches (potrf.cu) every 128-column step costs
// diagonal block -> TRSM -> panel update, three dependent kernels of 13-56 us that each use a fraction of
// the GPU; at N = 4096 that chain is 2.3 of the 3.1 ms while the O(N^3) work needs 0.6 ms at the DMMA peak.
//
// Here ONE cooperative launch holds one CTA (512 threads) per SM for the whole factorisation:
//
//   CTAs 0..NG-1, the chain group.  The critical path POTRF(k) -> TRSM(k+1,k) -> SYRK(k+1,k+1) -> POTRF(k+1)
//       never waits for a kernel boundary.  CTA 0 factors and inverts the diagonal block (diag_block.cuh);
//       the two tile products between two diagonal blocks -- 21 us on one SM at the DMMA peak, as long as the
//       diagonal block itself -- are split over the NG CTAs by strips of 128 / NG rows, each CTA holding the
//       whole second operand in shared memory (one load phase, no ring: K = 128, latency-bound), meeting
//       through two arrival counters per step.
//
//   CTAs NG.., the workers.  Every other tile task, in HALF tiles (64 x 128) so that one SM runs two
//       independent warp groups of 8 warps: a task's fixed part (C half-tile in, operand prologue, C out,
//       fence + publish: L2-bandwidth and latency, measured 12.7 k cycles per full tile against 32.8 k of
//       DMMA work) overlaps the other group's DMMA main loop -- what two co-resident CTAs per SM do for the
//       GEMM of gemm.cu, inside a kernel whose CTAs must stay one per SM.
//         TRSM(i,h,k)   L(i,k)[half h] = A(i,k)[half h] W_kk^T              i >= k+2   needs DIAG[k]
//         UPD(i,h,j,k)  A(i,j)[half h] -= L(i,k)[half h] L(j,k)^T   i >= j >= k+1, not (k+1,k+1)
//       Half tiles are statically owned (one group applies all updates of a half tile, in step order, without
//       any lock), dealt out in order of decreasing column so that every step's live set is balanced.
//
//   Dependencies travel through release/acquire flags in global memory: DIAG[k], LRH[i][h][k] (half h of
//   L(i,k) final), CNT[i][h][j] (updates received).  Every group works through its tasks in one global
//   topological order (step, column, row), taking a TRSM of its own as soon as it is runnable, so the earliest
//   unfinished task is always runnable by its owner: no deadlock; all CTAs are co-resident (cooperative
//   launch); spin loops carry a clock-based time-out that aborts the launch through an error flag (info = -999).
//
// Tile product: DMMA.8x8x4, warps of 32 x 32, operands staged by a 3-stage cp.async.cg ring (L2 -> shared
// memory, never through L1: tiles are produced by other SMs).  The warp -> (row, column) block map pairs column
// blocks (0,3), (1,2) on each SM sub-partition: triangular skipping (TRSM: k <= column) then shortens every
// sub-partition's DMMA queue alike.
#include <algorithm>
#include <map>
#include <mutex>
#include <vector>
#include "common.cuh"
#include "launch.h"
#include "diag_block.cuh"
#include "diag_block512.cuh"

namespace {

constexpr int CT = GPB_NB;                  // tile edge
constexpr int HR = 64;                      // rows of a half tile
#ifndef GPB_CBK
#define GPB_CBK 16
#define GPB_GSTAGES 3
#endif
constexpr int CBK = GPB_CBK;                // k-chunk of the ring
constexpr int CLD = CBK + 4;                // padded row stride (doubles): conflict-free 8-byte fragment loads
constexpr int GSTAGES = GPB_GSTAGES;
constexpr int CTHREADS = 512;
constexpr int GTHREADS = 256;               // one worker group
constexpr int G_A = HR * CLD, G_B = CT * CLD, G_STAGE = G_A + G_B;      // doubles
constexpr int G_RING = GSTAGES * G_STAGE;                               // doubles per group (92160 bytes)
constexpr int CSLD = CT + 4;                // strip operand row stride (doubles)
constexpr int C_STRIP_MAX = (CT + 32) * CSLD * 8;                       // second operand + a 32-row strip
constexpr int C_SMEM_0 = (2 * G_RING * 8 > DIAG_SMEM) ? 2 * G_RING * 8 : DIAG_SMEM;
constexpr int C_SMEM_1 = (C_SMEM_0 > C_STRIP_MAX) ? C_SMEM_0 : C_STRIP_MAX;
constexpr int C_SMEM = (C_SMEM_1 > DIAG512_SMEM) ? C_SMEM_1 : DIAG512_SMEM;
constexpr long long C_TIMEOUT = 4000000000LL;                           // ~2 s of SM clocks

enum { TASK_TRSM = 0, TASK_UPD = 1, TASK_MUPD = 2 };
constexpr int NH2 = 8;                  // helpers of the pipelined chain group
constexpr int C_COMPLETE = 1 << 20;     // CNT value of a half tile that has received all its worker-side updates
constexpr int g_dclk_off = 2 * 160 * 4;       // phase clocks of one diagonal block behind the worker accounting
constexpr int g_hclk_off = g_dclk_off + 256;  // helper 0: 8 clock64 stamps per step, then 4 %globaltimer stamps per step
                                              // (one thread, one store each; read back by gpb_debug_chain_workers for
                                              //  tests/gpu_potrf_dataflow.py -- the timeline in DESIGN.md comes from these)
struct ChainTask { unsigned char type, h; short i, j, k; };
// One half tile owned by a worker group (scheduler 0): it receives the updates of steps 0..nupd-1 from its owner,
// then (has_trsm) the owner's TRSM at step j.  (The last one or two steps of the tiles next to the diagonal belong to
// the chain group.)
// dl: step that consumes the finished tile.  start / handoff: a tile of the diagonal band is owned by a regular group for
// its steps [0, nupd) (handoff = 1: that entry publishes its count, never "complete") and by a band group -- alone on its
// SM -- from step `start` on (it waits for CNT >= start first).
struct ChainTile { short i, j; unsigned char h, nupd, has_trsm, has_m; short dl; unsigned char start, handoff; };

struct ChainArgs {
    double* A; long long ld;
    double* W; long long ldw;
    double* V; long long ldv;
    int* info;
    int T; int n_valid;
    int NG;                     // CTAs of the chain group (CTA 0 = the chain, 1..NG-1 its helpers)
    int diag512;                // 1: full diagonal blocks by the 512-thread body (diag_block512.cuh)
    int pipelined;              // 1: chain group v2 (c0 publishes every 32-column block; helpers one block behind; inverter CTA)
    int imminent;               // "due soon": deadline <= step + imminent
    int horizon;                // a far tile (deadline > step + horizon) yields while the group has an imminent tile
    int fuse_guard;             // ... unless a more urgent tile of the group is due within this many steps
    int fuse;                   // most steps of one half tile's backlog applied in one task (K = 128 * steps)
    int mform;                  // 1: last worker update of a tile in M form, worker TRSMs out of place (see worker_group_edf)
    long long* tclk;            // trace (%globaltimer): [2T][T][4] per half tile: last update start | complete | TRSM start | done; then [T][4]
    double* M;                  // [3][T] tiles [128][128]: M_{k+c,k} = L(k+c,k) W_k  (c = 1..3)
    int* flags;                 // [0] error | DIAG[T] | TP[T] | SP[T] | LRH[2T*T] | CNT[2T*T] | LPUB[4T] | XP[4T] | UX[T] | MP[3T]
    const ChainTask* bulk; const int* bulk_off;     // per worker group: UPD tasks in (k, j, i, h) order
    const ChainTask* trsm; const int* trsm_off;     // per worker group: TRSM tasks in (k, i, h) order
    const ChainTile* tiles; const int* tile_off;    // per worker group: owned half tiles in (j, i, h) order (<= 32), or null
    long long* clk;             // [T][8] phase clocks of the chain CTA (gpb_debug_chain_clocks)
    long long* wclk;            // [2 * grid][4] per worker group: cycles waiting | in TRSM | in UPD | tasks
};
#define CHAIN_STAMP(k, p) do { if (tid == 0) a.clk[(k) * 8 + (p)] = clock64(); } while (0)

__device__ __forceinline__ int* f_diag(const ChainArgs& a, int k) { return a.flags + 1 + k; }
__device__ __forceinline__ int* f_tp(const ChainArgs& a, int k) { return a.flags + 1 + a.T + k; }
__device__ __forceinline__ int* f_sp(const ChainArgs& a, int k) { return a.flags + 1 + 2 * a.T + k; }
__device__ __forceinline__ int* f_lrh(const ChainArgs& a, int i, int h, int k) { return a.flags + 1 + 3 * a.T + (2 * i + h) * a.T + k; }
__device__ __forceinline__ int* f_cnt(const ChainArgs& a, int i, int h, int j) { return a.flags + 1 + 3 * a.T + 2 * a.T * a.T + (2 * i + h) * a.T + j; }
__device__ __forceinline__ int* f_lpub(const ChainArgs& a, int d, int bb) { return a.flags + 1 + 3 * a.T + 4 * a.T * a.T + d * 4 + bb; }
__device__ __forceinline__ int* f_xp(const ChainArgs& a, int k, int bb) { return a.flags + 1 + 7 * a.T + 4 * a.T * a.T + k * 4 + bb; }
__device__ __forceinline__ int* f_ux(const ChainArgs& a, int r) { return a.flags + 1 + 11 * a.T + 4 * a.T * a.T + r; }
__device__ __forceinline__ int* f_mp(const ChainArgs& a, int c, int k) { return a.flags + 1 + 12 * a.T + 4 * a.T * a.T + (c - 1) * a.T + k; }
__host__ __device__ inline size_t chain_flag_words(int T) { return 1 + 15 * (size_t)T + 4 * (size_t)T * T; }
__device__ __forceinline__ double* m_tile(const ChainArgs& a, int c, int k) { return a.M + ((long long)(c - 1) * a.T + k) * (CT * CT); }
// where L(i,k) lives while the factorisation runs: worker TRSMs (i >= k + 3) write to the unused strictly-lower tile (i,k)
// of W in M-form mode (the tile of A keeps A'(i,k) for the M-form readers and is overwritten at the very end)
__device__ __forceinline__ const double* l_tile(const ChainArgs& a, int i, int k, long long& ld) {
    if (a.mform && i >= k + 3) { ld = a.ldw; return a.W + (long long)i * CT * a.ldw + (long long)k * CT; }
    ld = a.ld;
    return a.A + (long long)i * CT * a.ld + (long long)k * CT;
}

__device__ __forceinline__ long long gtimer_ns() {
    long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    return t;
}

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

// ===========================================================================
// workers: half-tile tasks by one group of 8 warps (2 x 4 blocks of 32 x 32)
// ===========================================================================
__device__ __forceinline__ void group_bar(int grp) {
    asm volatile("bar.sync %0, 256;\n" ::"r"(2 + grp) : "memory");
}

template <int ROWS>
__device__ __forceinline__ void ring_load(double* sdst, const double* g, long long ld, int ltid) {
#pragma unroll
    constexpr int PPR = CBK / 2;            // 16-byte pieces per row
    for (int q = 0; q < ROWS * PPR / GTHREADS; q++) {
        const int c = ltid + q * GTHREADS;
        const int row = c / PPR, ch = c % PPR;
        cp_async16(sdst + row * CLD + ch * 2, g + (long long)row * ld + ch * 2);
    }
}

// acc += sum_k Ag[r, k] * Bg[c, k], K = 128, for this warp's 32 x 32 block (wm, wn) of a 64 x 128 half tile;
// the warp multiplies only k-chunks [0, kt_hi), but walks all chunks (loads and barriers are group-wide)
__device__ __forceinline__ void half_mm(double (&acc)[4][4][2], const double* Ag, long long lda, const double* Bg,
                                        long long ldb, int wm, int wn, int kt_hi, double* ring, int ltid, int grp,
                                        int NK = CT / CBK) {
    const int lane = ltid & 31, g = lane >> 2, t = lane & 3;
#pragma unroll
    for (int s = 0; s < GSTAGES - 1; s++) {
        ring_load<HR>(ring + s * G_STAGE, Ag + s * CBK, lda, ltid);
        ring_load<CT>(ring + s * G_STAGE + G_A, Bg + s * CBK, ldb, ltid);
        cp_async_commit();
    }
    for (int kt = 0; kt < NK; kt++) {
        cp_async_wait<GSTAGES - 2>();
        group_bar(grp);
        {
            const int nx = kt + GSTAGES - 1;
            if (nx < NK) {
                const int slot = nx % GSTAGES;
                ring_load<HR>(ring + slot * G_STAGE, Ag + nx * CBK, lda, ltid);
                ring_load<CT>(ring + slot * G_STAGE + G_A, Bg + nx * CBK, ldb, ltid);
            }
            cp_async_commit();
        }
        if (kt >= kt_hi) continue;
        const double* as = ring + (kt % GSTAGES) * G_STAGE + (wm * 32 + g) * CLD + t;
        const double* bs = ring + (kt % GSTAGES) * G_STAGE + G_A + (wn * 32 + g) * CLD + t;
#pragma unroll
        for (int kk = 0; kk < CBK / 4; kk++) {
            double av[4], bv[4];
#pragma unroll
            for (int mi = 0; mi < 4; mi++) av[mi] = as[mi * 8 * CLD + kk * 4];
#pragma unroll
            for (int ni = 0; ni < 4; ni++) bv[ni] = bs[ni * 8 * CLD + kk * 4];
#pragma unroll
            for (int mi = 0; mi < 4; mi++)
#pragma unroll
                for (int ni = 0; ni < 4; ni++) dmma884(acc[mi][ni][0], acc[mi][ni][1], av[mi], bv[ni]);
        }
    }
    cp_async_wait<0>();
    group_bar(grp);                  // the ring may be reused by this group's next task
}

// warp lw (0..7) of a group -> block (wm, wn); sub-partition lw & 3 holds column blocks {s, 3 - s}
__device__ __forceinline__ void warp_block(int lw, int& wm, int& wn) {
    wm = lw >> 2;
    wn = wm ? 3 - (lw & 3) : (lw & 3);
}

// L(i,k)[half h] = A(i,k)[half h] W_kk^T, in place.  W_kk lower triangular: column block wn needs k < 32 (wn + 1)
__device__ __forceinline__ void task_trsm(const ChainArgs& a, int i, int h, int k, double* ring, int ltid, int grp) {
    const int lw = ltid >> 5, lane = ltid & 31, g = lane >> 2, t = lane & 3;
    int wm, wn;
    warp_block(lw, wm, wn);
    double* At = a.A + ((long long)i * CT + h * HR) * a.ld + (long long)k * CT;
    const double* Wk = a.W + (long long)k * CT * a.ldw + (long long)k * CT;
    double acc[4][4][2];
#pragma unroll
    for (int mi = 0; mi < 4; mi++)
#pragma unroll
        for (int ni = 0; ni < 4; ni++) acc[mi][ni][0] = acc[mi][ni][1] = 0.0;
    half_mm(acc, At, a.ld, Wk, a.ldw, wm, wn, (wn + 1) * 32 / CBK, ring, ltid, grp);
    // every thread's loads of the half tile are complete (barrier at the end of half_mm): safe to overwrite
    long long ldd;
    double* Dt = const_cast<double*>(l_tile(a, i, k, ldd)) + (long long)h * HR * ldd;
#pragma unroll
    for (int mi = 0; mi < 4; mi++)
#pragma unroll
        for (int ni = 0; ni < 4; ni++) {
            double* dst = Dt + (long long)(wm * 32 + mi * 8 + g) * ldd + wn * 32 + ni * 8 + 2 * t;
            *reinterpret_cast<double2*>(dst) = make_double2(acc[mi][ni][0], acc[mi][ni][1]);
        }
}

// A(i,j)[half h] -= L(i,k)[half h] L(j,k)^T.  Diagonal tiles: only the 32x32 blocks on or below the diagonal.
//   mform: A(i,j)[half h] -= A'(i,k)[half h] M^T with M = L(j,k) W_k (the same product without L(i,k))
//   nk > 1: steps k..k+nk-1 in one pass (K = 128 nk; the operands of consecutive steps are adjacent columns)
__device__ __forceinline__ void task_upd(const ChainArgs& a, int i, int h, int j, int k, double* ring, int ltid, int grp,
                                         bool mform = false, int nk = 1) {
    const int lw = ltid >> 5, lane = ltid & 31, g = lane >> 2, t = lane & 3;
    int wm, wn;
    warp_block(lw, wm, wn);
    const bool active = (i != j) || (wn <= 2 * h + wm);
    long long ldi, ldj;
    const double* Ai;
    const double* Aj;
    if (mform) {
        Ai = a.A + ((long long)i * CT + h * HR) * a.ld + (long long)k * CT; ldi = a.ld;
        Aj = m_tile(a, j - k, k); ldj = CT;
    } else {
        Ai = l_tile(a, i, k, ldi) + (long long)h * HR * ldi;
        Aj = l_tile(a, j, k, ldj);
    }
    double* Ct = a.A + ((long long)i * CT + h * HR) * a.ld + (long long)j * CT;
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
    half_mm(acc, Ai, ldi, Aj, ldj, wm, wn, active ? nk * (CT / CBK) : 0, ring, ltid, grp, nk * (CT / CBK));
    if (!active) return;
#pragma unroll
    for (int mi = 0; mi < 4; mi++)
#pragma unroll
        for (int ni = 0; ni < 4; ni++) {
            double* dst = Ct + (long long)(wm * 32 + mi * 8 + g) * a.ld + wn * 32 + ni * 8 + 2 * t;
            *reinterpret_cast<double2*>(dst) = make_double2(-acc[mi][ni][0], -acc[mi][ni][1]);
        }
}

// stores of all threads of the group -> group barrier -> one gpu-scope fence + release store (cumulativity
// carries the other threads' stores; a fence per thread costs ~1k cycles more)
__device__ __forceinline__ void publish_group(int* flag, int value, int ltid, int grp) {
    group_bar(grp);
    if (ltid == 0) {
        __threadfence();
        st_release(flag, value);
    }
}

__device__ __forceinline__ void worker_group(const ChainArgs& a, double* ring, volatile int* s_act, int ltid, int grp) {
    int* err = a.flags;
    // worker group index: consecutive indices sit on DIFFERENT SMs (the plan deals the two halves of a tile to
    // consecutive groups; as siblings on one SM the two halves of an urgent tile ran at half speed each, 25 us
    // instead of 12, while most other SMs were idle)
    const int v = grp * ((int)gridDim.x - a.NG) + ((int)blockIdx.x - a.NG);
    int bt = a.bulk_off[v];
    const int bend = a.bulk_off[v + 1];
    int qt = a.trsm_off[v];
    const int qend = a.trsm_off[v + 1];
    long long w_wait = 0, w_trsm = 0, w_upd = 0, w_n = 0;
    while (bt < bend || qt < qend) {
        const long long tw0 = clock64();
        if (ltid < 32) {
            // The group's first warp picks the next action: an own TRSM as soon as it is runnable (it
            // unblocks other groups), else the next update in order.  One poll = one parallel round of
            // acquire loads: lanes 0-1 the TRSM's conditions, lanes 2-4 the update's.
            const int lane = ltid;
            ChainTask r = {0, 0, 0, 0, 0}, u = {0, 0, 0, 0, 0};
            const bool have_r = qt < qend, have_u = bt < bend;
            if (have_r) r = a.trsm[qt];
            if (have_u) u = a.bulk[bt];
            const int* fp = nullptr;
            int target = 0;
            if (have_r && lane == 0) { fp = f_cnt(a, r.i, r.h, r.k); target = r.k; }
            if (have_r && lane == 1) { fp = f_diag(a, r.k); target = 1; }
            if (have_u && lane == 2) { fp = f_lrh(a, u.i, u.h, u.k); target = 1; }
            if (have_u && lane == 3) { fp = f_lrh(a, u.j, 0, u.k); target = 1; }
            // the first half of a diagonal tile only multiplies by rows 0..63 of L(j,k)
            if (have_u && lane == 4 && !(u.i == u.j && u.h == 0)) { fp = f_lrh(a, u.j, 1, u.k); target = 1; }
            // (the tile's earlier steps: normally this group's own, but an express group takes over a chain tile)
            if (have_u && lane == 5) { fp = f_cnt(a, u.i, u.h, u.j); target = u.k; }
            int act = -1;
            const long long t0 = clock64();
            unsigned it = 0;
            for (;;) {
                const bool ok = fp ? (ld_acquire(fp) >= target) : true;
                const unsigned m = __ballot_sync(0xffffffffu, ok);
                if (have_r && (m & 3u) == 3u) { act = 0; break; }
                if (have_u && (m & 60u) == 60u) { act = 1; break; }
                if ((++it & 63u) == 0) {
                    int stop = 0;
                    if (lane == 0) {
                        if (ld_acquire(err) != 0) stop = 1;
                        else if (clock64() - t0 > C_TIMEOUT) { atomicExch(err, 3); stop = 1; }
                    }
                    if (__shfl_sync(0xffffffffu, stop, 0)) break;
                }
            }
            if (lane == 0) s_act[grp] = act;
        }
        group_bar(grp);
        const int act = s_act[grp];
        if (act < 0) return;
        const long long tw1 = clock64();
        w_wait += tw1 - tw0;
        if (act == 0) {
            const ChainTask r = a.trsm[qt++];
            task_trsm(a, r.i, r.h, r.k, ring, ltid, grp);
            publish_group(f_lrh(a, r.i, r.h, r.k), 1, ltid, grp);
            w_trsm += clock64() - tw1;
        } else {
            const ChainTask u = a.bulk[bt++];
            task_upd(a, u.i, u.h, u.j, u.k, ring, ltid, grp);
            publish_group(f_cnt(a, u.i, u.h, u.j), u.k + 1, ltid, grp);
            w_upd += clock64() - tw1;
        }
        w_n++;
    }
    if (ltid == 0) {
        a.wclk[v * 4 + 0] = w_wait; a.wclk[v * 4 + 1] = w_trsm;
        a.wclk[v * 4 + 2] = w_upd; a.wclk[v * 4 + 3] = w_n;
    }
}


// Scheduler 0 ("most urgent runnable tile first").  A group owns at most 32 half tiles; lane t of its first warp keeps
// tile t's progress (the next step to apply) in a register and polls that tile's next task -- the update of step `next`
// (L(i,next) and L(j,next) final) or, after the last update, the TRSM (DIAG[j]) -- so ONE round of acquire loads covers
// every task the group could run.  Tiles are sorted by column: the lowest runnable lane is the task whose result is
// needed first (column j is consumed at step j).  With the in-order lists of scheduler 1 the step-s update of the tile
// NEXT to the current column -- the only thing between TRSM(i,s) and TRSM(i,s+1) on row i's serial chain -- queued
// behind the group's backlog of far-off tiles from step s-1.
// No deadlock: a group never waits while one of its tasks is runnable, and the globally earliest unfinished task
// (in step order) always is.
// L(i,k)[half h] and both halves of L(j,k) final?  (three independent acquire loads)
__device__ __forceinline__ bool lrh3(const ChainArgs& a, const ChainTile& t, int k) {
    const int f0 = ld_acquire(f_lrh(a, t.i, t.h, k)), f1 = ld_acquire(f_lrh(a, t.j, 0, k)), f2 = ld_acquire(f_lrh(a, t.j, 1, k));
    return f0 >= 1 && f1 >= 1 && ((t.i == t.j && t.h == 0) || f2 >= 1);
}

__device__ __forceinline__ void worker_group_edf(const ChainArgs& a, double* ring, volatile int* s_act, volatile int* s_task,
                                                 int ltid, int grp) {
    int* err = a.flags;
    // worker group index: consecutive indices sit on DIFFERENT SMs (the plan deals the two halves of a tile to
    // consecutive groups; as siblings on one SM the two halves of an urgent tile ran at half speed each, 25 us
    // instead of 12, while most other SMs were idle)
    const int v = grp * ((int)gridDim.x - a.NG) + ((int)blockIdx.x - a.NG);
    const int t0 = a.tile_off[v], nt = a.tile_off[v + 1] - t0;
    ChainTile my = {0, 0, 0, 0, 0, 0, 0, 0, 0};
    int next = 0;                  // L-form updates applied (steps 0..next-1)
    int s_cur = 0;                 // diagonal blocks known to be published
    bool m1_done = true, m2_done = true;   // the M-form updates of steps j-1 (has_m bit 0) and j-2 (bit 1) applied, or none
    bool live = false;
    bool handed = true;            // (band group) the regular owner has applied its part of the tile
    if (ltid < 32 && ltid < nt) {
        my = a.tiles[t0 + ltid];
        m1_done = !(my.has_m & 1);
        m2_done = !(my.has_m & 2);
        next = my.start;
        handed = my.start == 0;
        live = (my.nupd > next) || my.has_m || my.has_trsm;
        // nothing to receive: complete from the start (the M-form readers of this tile poll it)
        if (my.nupd == 0 && !my.has_m && !my.handoff) st_release(f_cnt(a, my.i, my.h, my.j), C_COMPLETE);
    }
    long long w_wait = 0, w_trsm = 0, w_upd = 0, w_n = 0;
    for (;;) {
        const long long tw0 = clock64();
        if (ltid < 32) {
            const int lane = ltid;
            const long long tp0 = clock64();
            unsigned it = 0;
            int act = -1, kind = 0, nfuse = 1;
            for (;;) {
                // this tile's runnable task: 1 = M-form update of step j-2, 2 = of step j-1 (A'(i,kappa) complete and
                // M_{j,kappa} published), 3 = TRSM (nothing else left, DIAG[j]), 4 = L-form update of step `next`
                int cand = 0;
                nfuse = 1;
                if (live && !handed && ld_acquire(f_cnt(a, my.i, my.h, my.j)) >= (int)my.start) handed = true;
                // the group's most urgent live tile: a long fused task must not start when that tile's last updates
                // are about to become runnable (tasks are not preempted)
                const int dmin = (int)__reduce_min_sync(0xffffffffu, live ? (unsigned)my.dl : 0x7fffffffu);
                // how far the factorisation is (diagonal blocks published; at most one step per poll round)
                if (s_cur < a.T && ld_acquire(f_diag(a, s_cur)) >= 1) s_cur++;
                if (live && handed) {
                    const bool l_left = next < (int)my.nupd;
                    // (a tile's updates are applied in step order -- L form 0..nupd-1, then j-2, then j-1 -- whatever the
                    //  timing: the factor is bit-reproducible from run to run)
                    // (the flags of one task are loaded together: one L2 round trip per poll, not one per flag)
                    if (!l_left && !m2_done) {
                        const int f0 = ld_acquire(f_cnt(a, my.i, my.h, my.j - 2)), f1 = ld_acquire(f_mp(a, 2, my.j - 2));
                        if (f0 >= C_COMPLETE && f1 >= NH2) cand = 1;
                    } else if (!l_left && !m1_done) {
                        const int f0 = ld_acquire(f_cnt(a, my.i, my.h, my.j - 1)), f1 = ld_acquire(f_mp(a, 1, my.j - 1));
                        if (f0 >= C_COMPLETE && f1 >= NH2) cand = 2;
                    } else if (!l_left) {
                        if (ld_acquire(f_diag(a, my.j)) >= 1) cand = 3;
                    } else if (my.dl != dmin && dmin <= s_cur + a.imminent && my.dl > s_cur + a.horizon) {
                        // Tasks are not preempted: while the group's most urgent tile is due within a step, a tile
                        // that is not needed for `horizon` more steps waits (its backlog is absorbed -- fused -- in the
                        // second half of the factorisation, where the workers have time).  Everything an imminent task
                        // depends on has a deadline <= its own, so nothing it waits for is ever held back here.
                    } else if (lrh3(a, my, next)) {
                        cand = 4;
                        // backlog: the following steps too, while their operands are final and live in the same
                        // buffer as this step's (M-form mode: L(r,s) is in W for r >= s + 3, the chain's rows in A)
                        while (nfuse < a.fuse && next + nfuse < (int)my.nupd && (!a.mform || next + nfuse <= my.j - 3) &&
                               (my.dl == dmin || next + nfuse + a.fuse_guard <= dmin) &&
                               lrh3(a, my, next + nfuse))
                            nfuse++;
                    }
                }
                // earliest deadline first over ALL runnable tasks: an update inherits its tile's deadline, a TRSM is
                // consumed one step after its tile (by the step-j updates of row i); ties: M form / TRSM, then lane
                unsigned key = 0xffffffffu;
                if (cand) key = ((unsigned)(my.dl + (cand == 3 ? 1 : 0)) << 10) | (cand == 4 ? 256u : 0u) | ((unsigned)cand << 5) | (unsigned)lane;
                const unsigned best = __reduce_min_sync(0xffffffffu, key);
                const unsigned lv = __ballot_sync(0xffffffffu, live);
                if (lv == 0u) { act = -2; break; }
                if (best != 0xffffffffu) { act = (int)(best & 31u); kind = (int)((best >> 5) & 7u); break; }
                if ((++it & 63u) == 0) {
                    int stop = 0;
                    if (lane == 0) {
                        if (ld_acquire(err) != 0) stop = 1;
                        else if (clock64() - tp0 > C_TIMEOUT) { atomicExch(err, 3); stop = 1; }
                    }
                    if (__shfl_sync(0xffffffffu, stop, 0)) break;
                }
            }
            if (act >= 0 && lane == act) {
                int type, k;
                if (kind == 1) { type = TASK_MUPD; k = my.j - 2; m2_done = true; }
                else if (kind == 2) { type = TASK_MUPD; k = my.j - 1; m1_done = true; }
                else if (kind == 3) { type = TASK_TRSM; k = my.j; live = false; }
                else { type = TASK_UPD; k = next; next += nfuse; }
                const bool complete = m1_done && m2_done && next >= (int)my.nupd;
                if (type != TASK_TRSM && complete && !my.has_trsm) live = false;
                s_task[grp * 6 + 0] = type;
                s_task[grp * 6 + 1] = ((int)my.i << 16) | (int)my.j;
                s_task[grp * 6 + 2] = (int)my.h;
                s_task[grp * 6 + 3] = k;
                s_task[grp * 6 + 4] = (complete && !my.handoff) ? C_COMPLETE : next;
                s_task[grp * 6 + 5] = (type == TASK_UPD) ? nfuse : 1;
            }
            if (lane == 0) s_act[grp] = act;
        }
        group_bar(grp);
        const int act = s_act[grp];
        if (act < 0) {
            if (act == -1) return;          // aborted
            break;
        }
        const int type = s_task[grp * 6 + 0], ij = s_task[grp * 6 + 1], h = s_task[grp * 6 + 2], k = s_task[grp * 6 + 3];
        const int cnt = s_task[grp * 6 + 4];
        const int i = ij >> 16, j = ij & 0xffff;
        const long long tw1 = clock64();
        w_wait += tw1 - tw0;
        long long* tc = a.tclk + ((long long)(2 * i + h) * a.T + j) * 4;
        if (type == TASK_TRSM) {
            if (ltid == 0) tc[2] = gtimer_ns();
            task_trsm(a, i, h, k, ring, ltid, grp);
            publish_group(f_lrh(a, i, h, k), 1, ltid, grp);
            if (ltid == 0) tc[3] = gtimer_ns();
            w_trsm += clock64() - tw1;
        } else {
            if (ltid == 0 && cnt == C_COMPLETE) tc[0] = gtimer_ns();
            task_upd(a, i, h, j, k, ring, ltid, grp, type == TASK_MUPD, s_task[grp * 6 + 5]);
            publish_group(f_cnt(a, i, h, j), cnt, ltid, grp);
            if (ltid == 0 && cnt == C_COMPLETE) tc[1] = gtimer_ns();
            w_upd += clock64() - tw1;
        }
        w_n++;
    }
    if (ltid == 0) {
        a.wclk[v * 4 + 0] = w_wait; a.wclk[v * 4 + 1] = w_trsm;
        a.wclk[v * 4 + 2] = w_upd; a.wclk[v * 4 + 3] = w_n;
    }
    if (!a.mform) return;
    group_bar(grp);                 // every thread has read the loop's last s_act before it is reused below
    // M-form mode: L(i,j) of this group's TRSMs sits in W's tile (i,j); move it home once the readers of A'(i,j) -- the
    // M-form updates of tiles (i, j+1) and (i, j+2) -- are done.
    for (int tix = 0; tix < nt; tix++) {
        const ChainTile tl = a.tiles[t0 + tix];
        if (!tl.has_trsm) continue;
        if (ltid == 0) {
            const bool ok = spin_ge(f_cnt(a, tl.i, tl.h, tl.j + 1), C_COMPLETE, err) && spin_ge(f_cnt(a, tl.i, tl.h, tl.j + 2), C_COMPLETE, err);
            s_act[grp] = ok ? 1 : 0;
        }
        group_bar(grp);
        if (!s_act[grp]) return;
        const double* src = a.W + ((long long)tl.i * CT + tl.h * HR) * a.ldw + (long long)tl.j * CT;
        double* dst = a.A + ((long long)tl.i * CT + tl.h * HR) * a.ld + (long long)tl.j * CT;
        for (int e = ltid; e < HR * (CT / 2); e += 256) {
            const int r = e >> 6, c2 = (e & 63) * 2;
            *reinterpret_cast<double2*>(dst + (long long)r * a.ld + c2) =
                __ldcg(reinterpret_cast<const double2*>(src + (long long)r * a.ldw + c2));
        }
        group_bar(grp);
    }
}

// ===========================================================================
// chain group: strips of CR = 128 / NG rows
// ===========================================================================
// rows [r0, r0 + CR) of  X = A(k,k-1) W_{k-1}^T  (in place); X strip also left in shared memory (As)
template <int CR>
__device__ __forceinline__ void strip_trsm(const ChainArgs& a, int k, int r0, double* Bs, double* As, int tid, int kc = -1) {
    // tile (k, kc) with the inverted diagonal block kc (kc = k - 1 unless given)
    if (kc < 0) kc = k - 1;
    const int wid = tid >> 5, lane = tid & 31, g = lane >> 2, t = lane & 3;
    double* At = a.A + ((long long)k * CT + r0) * a.ld + (long long)kc * CT;
    const double* Wk = a.W + (long long)kc * CT * a.ldw + (long long)kc * CT;
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
__device__ __forceinline__ void strip_syrk(const ChainArgs& a, int k, int r0, double* Bs, const double* As, int tid, int kc = -1) {
    // A(k,k)[strip] -= X[strip] X^T with X = tile (k, kc) (kc = k - 1 unless given)
    if (kc < 0) kc = k - 1;
    const int wid = tid >> 5, lane = tid & 31, g = lane >> 2, t = lane & 3;
    const double* Xt = a.A + (long long)k * CT * a.ld + (long long)kc * CT;
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

// rows [r0, r0 + CR) of  A(r,j) -= X1 L(j,c)^T  with X1 = this strip of L(r,c) in shared memory (As1); the result
// also goes to Aout (shared strip): it is the helper's own input of the next pipelined TRSM
template <int CR>
__device__ __forceinline__ void strip_upd(const ChainArgs& a, int r, int j, int c, int r0, double* Bs, const double* As1,
                                          double* Aout, int tid) {
    const int wid = tid >> 5, lane = tid & 31, g = lane >> 2, t = lane & 3;
    const double* Lj = a.A + (long long)j * CT * a.ld + (long long)c * CT;
    double* Ct = a.A + ((long long)r * CT + r0) * a.ld + (long long)j * CT;
    for (int e = tid; e < CT * (CT / 2); e += CTHREADS) {
        const int q = e >> 6, p2 = (e & 63) * 2;
        cp_async16(Bs + q * CSLD + p2, Lj + (long long)q * a.ld + p2);
    }
    cp_async_commit();
    double acc[CR / 8][2];
#pragma unroll
    for (int mi = 0; mi < CR / 8; mi++) {
        const double2 cv = __ldcg(reinterpret_cast<const double2*>(Ct + (long long)(mi * 8 + g) * a.ld + wid * 8 + 2 * t));
        acc[mi][0] = -cv.x; acc[mi][1] = -cv.y;
    }
    cp_async_wait<0>();
    __syncthreads();
    const double* ap = As1 + g * CSLD + t;
    const double* bp = Bs + (wid * 8 + g) * CSLD + t;
#pragma unroll 4
    for (int kk = 0; kk < CT / 4; kk++) {
        const double b = bp[kk * 4];
#pragma unroll
        for (int mi = 0; mi < CR / 8; mi++) dmma884(acc[mi][0], acc[mi][1], ap[mi * 8 * CSLD + kk * 4], b);
    }
#pragma unroll
    for (int mi = 0; mi < CR / 8; mi++) {
        const int rr = mi * 8 + g, cc = wid * 8 + 2 * t;
        *reinterpret_cast<double2*>(Ct + (long long)rr * a.ld + cc) = make_double2(-acc[mi][0], -acc[mi][1]);
        Aout[rr * CSLD + cc] = -acc[mi][0];
        Aout[rr * CSLD + cc + 1] = -acc[mi][1];
    }
}

// rows [r0, r0 + CR) of  M_{.,c} = L W_c  (L strip in shared memory, As; W_c is loaded into Bs unless it is there) ->
// tile `slot` of step c in a.M, then this CTA's arrival on MP[slot][c]
template <int CR>
__device__ __forceinline__ void strip_m(const ChainArgs& a, int c, int r0, double* Bs, const double* As, int tid, int slot,
                                        bool load_w) {
    const int wid = tid >> 5, lane = tid & 31, g = lane >> 2, t = lane & 3;
    if (load_w) {
        const double* Wk = a.W + (long long)c * CT * a.ldw + (long long)c * CT;
        for (int e = tid; e < CT * (CT / 2); e += CTHREADS) {
            const int q = e >> 6, p2 = (e & 63) * 2;
            if (p2 < ((q >> 5) + 1) * 32) cp_async16(Bs + q * CSLD + p2, Wk + (long long)q * a.ldw + p2);
        }
        cp_async_commit();
        cp_async_wait<0>();
        __syncthreads();
    }
    // warp w: columns [8w, 8w + 8); W_c is lower triangular: rows p >= 8w only
    double acc[CR / 8][2];
#pragma unroll
    for (int mi = 0; mi < CR / 8; mi++) acc[mi][0] = acc[mi][1] = 0.0;
    const double* ap = As + g * CSLD + t;
    const double* bp = Bs + t * CSLD + wid * 8 + g;
#pragma unroll 4
    for (int kk = 2 * wid; kk < CT / 4; kk++) {
        const double b = bp[kk * 4 * CSLD];
#pragma unroll
        for (int mi = 0; mi < CR / 8; mi++) dmma884(acc[mi][0], acc[mi][1], ap[mi * 8 * CSLD + kk * 4], b);
    }
    double* Mt = m_tile(a, slot, c) + (long long)r0 * CT;
#pragma unroll
    for (int mi = 0; mi < CR / 8; mi++)
        *reinterpret_cast<double2*>(Mt + (mi * 8 + g) * CT + wid * 8 + 2 * t) = make_double2(acc[mi][0], acc[mi][1]);
    __syncthreads();
    if (tid == 0) {
        __threadfence();
        atomicAdd(f_mp(a, slot, c), 1);
        if (blockIdx.x == 1) a.tclk[(long long)2 * a.T * a.T * 4 + c * 4 + slot] = gtimer_ns();
    }
}

__device__ __forceinline__ void publish(int* flag, int value, int tid) {
    __syncthreads();
    if (tid == 0) {
        __threadfence();
        st_release(flag, value);
    }
}

// this CTA's part is stored: count it; the last of `total` arrivals publishes the two `done` flags (optional)
__device__ __forceinline__ void arrive(int* counter, int total, int* done0, int* done1, int tid) {
    __syncthreads();
    if (tid == 0) {
        __threadfence();
        const int old = atomicAdd(counter, 1);
        if (done0 && old == total - 1) {
            __threadfence();
            st_release(done0, 1);
            st_release(done1, 1);
        }
    }
}

template <int CR>
__device__ __forceinline__ void chain_group(const ChainArgs& a, double* csm, volatile int* s_act, int tid) {
    int* err = a.flags;
    const int c = blockIdx.x, NG = a.NG, T = a.T;
    const int r0 = c * CR, h = r0 / HR;
    double* Bs = csm;
    double* As = csm + CT * CSLD;
    for (int k = 0; k < T; k++) {
        if (c == 0) CHAIN_STAMP(k, 0);
        if (k > 0) {
            if (tid == 0)
                s_act[0] = (spin_ge(f_diag(a, k - 1), 1, err) && spin_ge(f_cnt(a, k, h, k - 1), k - 1, err)) ? 1 : 0;
            __syncthreads();
            if (!s_act[0]) { if (tid == 0 && c == 0) *a.info = -999; return; }
            if (c == 0) CHAIN_STAMP(k, 1);
            strip_trsm<CR>(a, k, r0, Bs, As, tid);
            if (c == 0) CHAIN_STAMP(k, 2);
            arrive(f_tp(a, k), NG, f_lrh(a, k, 0, k - 1), f_lrh(a, k, 1, k - 1), tid);
            if (c == 0) CHAIN_STAMP(k, 3);
            if (tid == 0)
                s_act[1] = (spin_ge(f_tp(a, k), NG, err) && spin_ge(f_cnt(a, k, h, k), k - 1, err)) ? 1 : 0;
            __syncthreads();
            if (!s_act[1]) { if (tid == 0 && c == 0) *a.info = -999; return; }
            if (c == 0) CHAIN_STAMP(k, 4);
            strip_syrk<CR>(a, k, r0, Bs, As, tid);
            arrive(f_sp(a, k), NG, nullptr, nullptr, tid);
        }
        if (c != 0) continue;
        if (k > 0) {
            if (tid == 0) s_act[2] = spin_ge(f_sp(a, k), NG, err) ? 1 : 0;
            __syncthreads();
            if (!s_act[2]) { if (tid == 0) *a.info = -999; return; }
        }
        CHAIN_STAMP(k, 5);
        {
            const long long o = (long long)k * CT;
            const long long valid = (long long)a.n_valid - o;
            const int nsub = valid >= CT ? 4 : (valid <= 0 ? 0 : (int)((valid + SB - 1) / SB));
            long long* dclk = (k == 1) ? a.wclk + (size_t)g_dclk_off : nullptr;
            double* Ak = a.A + o * a.ld + o;
            double* Wk = a.W + o * a.ldw + o;
            double* Vk = a.V ? a.V + o * a.ldv + o : nullptr;
            if (nsub == 4 && a.diag512 == 1)     // full block: the 512-thread body with look-ahead inside the block
                diag_block_body512<8>(Ak, a.ld, Wk, a.ldw, Vk, a.ldv, a.info, (int)o, csm, dclk);
            else if (nsub == 4 && a.diag512 == 3) diag_block_body512<1>(Ak, a.ld, Wk, a.ldw, Vk, a.ldv, a.info, (int)o, csm, dclk);
            else if (nsub == 4 && a.diag512 == 4) diag_block_body512<2>(Ak, a.ld, Wk, a.ldw, Vk, a.ldv, a.info, (int)o, csm, dclk);
            else if (nsub == 4 && a.diag512 == 5) diag_block_body512<4>(Ak, a.ld, Wk, a.ldw, Vk, a.ldv, a.info, (int)o, csm, dclk);
            else if (tid < 256)             // identity-padded last block
                diag_block_body<true>(a.A + o * a.ld + o, a.ld, a.W + o * a.ldw + o, a.ldw,
                                      a.V ? a.V + o * a.ldv + o : nullptr, a.ldv, a.info, (int)o, nsub, csm);
        }
        CHAIN_STAMP(k, 6);
        publish(f_diag(a, k), 1, tid);
        CHAIN_STAMP(k, 7);
    }
}

// ===========================================================================
// chain group v2 ("pipelined"): CTA 0 sweeps, CTAs 1..8 helpers, CTA 9 inverter.
//
// In v1 the group runs diagonal block (76 k cycles, 22 k of them the 128-level inverse) -> TRSM strips -> SYRK
// strips one after the other.  Here CTA 0 releases every 32-column block of L_kk (+ its inverted 32x32 diagonal
// sub-block) as soon as it is final (diag_block512.cuh, lpub), and the helpers run the next tile's TRSM and SYRK
// block column by block column ONE BLOCK BEHIND the sweeps:
//     X[:, bb] = (A[:, bb] - sum_{b' < bb} X[:, b'] L(bb,b')^T) W_bb^T          (16-row strip per helper)
//     C strip  -= X[:, bb] X_all[:, bb]^T                                        (after the group exchanged X[:, bb])
// so that after the last sweep only the last block column is left.  The 128-level inverse W_kk = L_kk^-1 is only
// needed by the WORKERS' TRSM tasks (one step of slack): the inverter CTA completes it from global memory and
// publishes DIAG[k]; CTA 0 goes straight to the next diagonal block.
// ===========================================================================
constexpr int CR2 = CT / NH2;           // 16 rows each
constexpr int XLD = SB + 4;             // stride of 32-column operands (conflict-free fragment loads)

__device__ __forceinline__ void helper_v2(const ChainArgs& a, double* csm, volatile int* s_act, int tid, int h) {
    int* err = a.flags;
    const int T = a.T;
    const int wid = tid >> 5, lane = tid & 31, g = lane >> 2, t = lane & 3;
    const int r0 = h * CR2, half = r0 / HR;
    double* As = csm;                               // [CR2][CSLD]   the strip: A, block column by block column -> X
    double* As1 = As + CR2 * CSLD;                  // [CR2][CSLD]   urgent phase: this strip of L(k, k-2)
    double* Bs = As1 + CR2 * CSLD;                  // [CT][CSLD]    urgent phase: second operand (overlays the next three)
    double* Rs = Bs;                                // [CR2][XLD]    residual of the current block column
    double* Lk = Rs + CR2 * XLD;                    // 4 blocks [SB][SLD]: L(bb,0..2) and W_bb
    double* Xa = Lk + 4 * SBSZ;                     // [CT][XLD]     X[:, bb] of the whole tile
    const int mi = wid & 1, ni = wid >> 1;          // warps 0..7: tile (mi, ni) of a [16 x 32] block
    const int ncol = ((r0 + CR2 + 31) >> 5) << 5;   // rows of X (= columns of C) this strip needs: lower triangle
#define HSTAMP(i) do { if (h == 0 && tid == 0) a.wclk[g_hclk_off + k * 8 + (i)] = clock64(); } while (0)
    for (int k = 1; k < T; k++) {
        const int d = k - 1;
        HSTAMP(0);
        double* At = a.A + ((long long)k * CT + r0) * a.ld + (long long)d * CT;         // strip of tile (k, k-1)
        const double* Ld = a.A + (long long)d * CT * a.ld + (long long)d * CT;          // diagonal block d
        const double* Wdg = a.W + (long long)d * CT * a.ldw + (long long)d * CT;
        double* Ct = a.A + ((long long)k * CT + r0) * a.ld + (long long)k * CT;          // strip of tile (k, k)
        if (d >= 1) {
            // ---- urgent phase (while CTA 0 starts diagonal block d): step d-1 of this row's two chain tiles.
            //      The workers' updates reach these tiles a full step late (inverter -> TRSM -> update, each behind
            //      whatever task its owner is running), so the group applies the last step itself:
            //        L(k, d-1) = A(k, d-1) W_{d-1}^T ;  A(k, d) -= L(k, d-1) L(d, d-1)^T ;  A(k, k) -= L(k, d-1) L(k, d-1)^T
            const int c = d - 1;
            if (tid == 0) {
                const bool ok0 = spin_ge(f_diag(a, c), 1, err);
                if (h == 0) a.wclk[g_hclk_off + 8 * 64 + k * 4 + 0] = gtimer_ns();
                s_act[3] = ok0 ? 1 : 0;
            }
            __syncthreads();
            if (!s_act[3]) return;
            // M-form: M_{c+1,c} = L(c+1,c) W_c first -- every row's update of its tile (i, c+1) waits for it.  This CTA's
            // strip of L(c+1,c) = L(d,c) is still in As from the previous iteration's pipelined columns.
            if (a.mform) strip_m<CR2>(a, c, r0, Bs, As, tid, 1, true);
            if (tid == 0) s_act[0] = spin_ge(f_cnt(a, k, half, c), c, err) ? 1 : 0;
            __syncthreads();
            if (!s_act[0]) return;
            HSTAMP(1);
            strip_trsm<CR2>(a, k, r0, Bs, As1, tid, c);
            // M_{c+2,c} = L(k,c) W_c for the last worker-side update of tile (c+3, c+2); W_c is still in Bs
            if (a.mform >= 3) { __syncthreads(); strip_m<CR2>(a, c, r0, Bs, As1, tid, 2, false); }
            HSTAMP(2);
            arrive(f_ux(a, k), NH2, f_lrh(a, k, 0, c), f_lrh(a, k, 1, c), tid);
            if (tid == 0) s_act[1] = (spin_ge(f_cnt(a, k, half, d), d - 1, err) && spin_ge(f_xp(a, d, 3), NH2, err)) ? 1 : 0;
            __syncthreads();
            if (!s_act[1]) return;
            HSTAMP(3);
            strip_upd<CR2>(a, k, d, c, r0, Bs, As1, As, tid);
            HSTAMP(4);
            if (tid == 0) {
                const bool ok0 = spin_ge(f_ux(a, k), NH2, err);
                if (h == 0) a.wclk[g_hclk_off + 8 * 64 + k * 4 + 1] = gtimer_ns();
                s_act[2] = (ok0 && spin_ge(f_cnt(a, k, half, k), d - 1, err)) ? 1 : 0;
            }
            __syncthreads();
            if (!s_act[2]) return;
            HSTAMP(5);
            strip_syrk<CR2>(a, k, r0, Bs, As1, tid, c);
            __syncthreads();
            HSTAMP(6);
        } else {
            for (int e = tid; e < CR2 * (CT / 2); e += CTHREADS) {
                const int r = e >> 6, p2 = (e & 63) * 2;
                cp_async16(As + r * CSLD + p2, At + (long long)r * a.ld + p2);
            }
        }
        cp_async_commit();
        double acc[CR2 / 8][2];
#pragma unroll
        for (int m = 0; m < CR2 / 8; m++) acc[m][0] = acc[m][1] = 0.0;
        for (int bb = 0; bb < 4; bb++) {
            if (tid == 0) s_act[1] = spin_ge(f_lpub(a, d, bb), 1, err) ? 1 : 0;
            __syncthreads();
            if (!s_act[1]) return;
            // L(bb, b') for b' < bb and W_bb -> shared memory
            for (int e = tid; e < (bb + 1) * SB * (SB / 2); e += CTHREADS) {
                const int blkq = e / (SB * SB / 2), q = e % (SB * SB / 2), r = q >> 4, c2 = (q & 15) * 2;
                const double* src = (blkq < bb) ? Ld + (long long)(bb * SB + r) * a.ld + blkq * SB + c2
                                                : Wdg + (long long)(bb * SB + r) * a.ldw + bb * SB + c2;
                cp_async16(Lk + (blkq < bb ? blkq : 3) * SBSZ + r * SLD + c2, src);
            }
            cp_async_commit();
            cp_async_wait<0>();
            __syncthreads();
            if (wid < 8) {
                // R = A[:, bb] - sum_{b' < bb} X[:, b'] L(bb,b')^T      (tile (mi, ni) of [16 x 32])
                double r0v = As[(mi * 8 + g) * CSLD + bb * SB + ni * 8 + 2 * t], r1v = As[(mi * 8 + g) * CSLD + bb * SB + ni * 8 + 2 * t + 1];
                for (int bp = 0; bp < bb; bp++) {
                    const double* ap = As + (mi * 8 + g) * CSLD + bp * SB + t;
                    const double* bp_ = Lk + bp * SBSZ + (ni * 8 + g) * SLD + t;
#pragma unroll
                    for (int kk = 0; kk < 8; kk++) dmma884(r0v, r1v, -ap[kk * 4], bp_[kk * 4]);
                }
                Rs[(mi * 8 + g) * XLD + ni * 8 + 2 * t] = r0v;
                Rs[(mi * 8 + g) * XLD + ni * 8 + 2 * t + 1] = r1v;
            }
            __syncthreads();
            if (wid < 8) {
                // X[:, bb] = R W_bb^T   (W_bb lower triangular: k <= column)
                double x0 = 0.0, x1 = 0.0;
                const double* ap = Rs + (mi * 8 + g) * XLD + t;
                const double* bp_ = Lk + 3 * SBSZ + (ni * 8 + g) * SLD + t;
                for (int kk = 0; kk < 2 * ni + 2; kk++) dmma884(x0, x1, ap[kk * 4], bp_[kk * 4]);
                const int r = mi * 8 + g, c = bb * SB + ni * 8 + 2 * t;
                *reinterpret_cast<double2*>(At + (long long)r * a.ld + c) = make_double2(x0, x1);
                As[r * CSLD + c] = x0;
                As[r * CSLD + c + 1] = x1;
            }
            // this strip's X[:, bb] is out: count it; the last arrival of the last block publishes L(k, k-1)
            __syncthreads();
            if (tid == 0) {
                __threadfence();
                const int old = atomicAdd(f_xp(a, k, bb), 1);
                if (bb == 3 && old == NH2 - 1) {
                    __threadfence();
                    st_release(f_lrh(a, k, 0, d), 1);
                    st_release(f_lrh(a, k, 1, d), 1);
                }
                s_act[2] = spin_ge(f_xp(a, k, bb), NH2, err) ? 1 : 0;
            }
            __syncthreads();
            if (!s_act[2]) return;
            // X[:, bb] of the rows this strip's part of the lower triangle needs
            const double* Xt = a.A + (long long)k * CT * a.ld + (long long)d * CT + bb * SB;
            for (int e = tid; e < ncol * (SB / 2); e += CTHREADS) {
                const int r = e >> 4, c2 = (e & 15) * 2;
                cp_async16(Xa + r * XLD + c2, Xt + (long long)r * a.ld + c2);
            }
            cp_async_commit();
            cp_async_wait<0>();
            __syncthreads();
            if (wid * 8 < ncol) {
                const double* ap = As + g * CSLD + bb * SB + t;
                const double* bp_ = Xa + (wid * 8 + g) * XLD + t;
#pragma unroll
                for (int kk = 0; kk < 8; kk++) {
                    const double b = bp_[kk * 4];
#pragma unroll
                    for (int m = 0; m < CR2 / 8; m++) dmma884(acc[m][0], acc[m][1], ap[m * 8 * CSLD + kk * 4], b);
                }
            }
            __syncthreads();            // Lk / Xa are reloaded by the next block column
            if (bb == 0 && a.mform >= 3 && k >= 3) {
                // M-form mode: step k-3 of the diagonal tile is this group's too (a TRSM -> update chain of the workers
                // inside one step otherwise): acc += L(k,k-3)[strip] L(k,k-3)^T, in the idle time before block column 1.
                // L(k,k-3) is a worker TRSM: it sits in W's tile.
                const int c3 = k - 3;
                if (tid == 0) s_act[1] = (spin_ge(f_lrh(a, k, 0, c3), 1, err) && spin_ge(f_lrh(a, k, 1, c3), 1, err)) ? 1 : 0;
                __syncthreads();
                if (!s_act[1]) return;
                const double* Ls = a.W + (long long)k * CT * a.ldw + (long long)c3 * CT;
                for (int e = tid; e < CR2 * (CT / 2); e += CTHREADS) {
                    const int r = e >> 6, p2 = (e & 63) * 2;
                    cp_async16(As1 + r * CSLD + p2, Ls + (long long)(r0 + r) * a.ldw + p2);
                }
                for (int e = tid; e < ncol * (CT / 2); e += CTHREADS) {
                    const int r = e >> 6, p2 = (e & 63) * 2;
                    cp_async16(Bs + r * CSLD + p2, Ls + (long long)r * a.ldw + p2);
                }
                cp_async_commit();
                cp_async_wait<0>();
                __syncthreads();
                if (wid * 8 < ncol) {
                    const double* ap = As1 + g * CSLD + t;
                    const double* bp_ = Bs + (wid * 8 + g) * CSLD + t;
#pragma unroll 4
                    for (int kk = 0; kk < CT / 4; kk++) {
                        const double b = bp_[kk * 4];
#pragma unroll
                        for (int m = 0; m < CR2 / 8; m++) dmma884(acc[m][0], acc[m][1], ap[m * 8 * CSLD + kk * 4], b);
                    }
                }
                __syncthreads();
            }
        }
        // C strip -= sum_bb X[:, bb] X_all[:, bb]^T   (the tile's earlier steps: workers, then the urgent phase above)
        if (wid * 8 < ncol) {
#pragma unroll
            for (int m = 0; m < CR2 / 8; m++) {
                double2* p = reinterpret_cast<double2*>(Ct + (long long)(m * 8 + g) * a.ld + wid * 8 + 2 * t);
                const double2 c = __ldcg(p);
                *p = make_double2(c.x - acc[m][0], c.y - acc[m][1]);
            }
        }
        arrive(f_sp(a, k), NH2, nullptr, nullptr, tid);
        HSTAMP(7);
    }
#undef HSTAMP
}

// the 128-level inverse of diagonal block k from its published 32x32 pieces, as early as they arrive:
//   after block column 1:  W_10 = -W_11 (L_10 W_00)
//   after block column 2:  S_ij = sum_k L_ik W_kj (i = 2,3; j = 0,1), S_32 = L_32 W_22, W_2j = -W_22 S_2j,
//                          U_j = S_3j - S_32 S_2j
//   after block column 3:  W_32 = -W_33 S_32, W_3j = -W_33 U_j      -- ONE phase behind the last 32x32 inverse
__device__ __forceinline__ void inverter_v2(const ChainArgs& a, double* csm, volatile int* s_act, int tid) {
    int* err = a.flags;
    const int T = a.T, wid = tid >> 5;
    double* Lb = csm;                               // 10 block slots: the six off-diagonal L blocks; the diagonal slots hold W_bb
    double* Xs = csm + NBLK * SBSZ;                 // S10 W10 S32 S20 S21 S30 S31 U0 U1
    auto Wdp = [&](int b) { return Lb + blk(b, b) * SBSZ; };
    double* const S10 = Xs, * const W10 = Xs + SBSZ, * const S32 = Xs + 2 * SBSZ;
    double* const S20 = Xs + 3 * SBSZ, * const S21 = Xs + 4 * SBSZ, * const S30 = Xs + 5 * SBSZ, * const S31 = Xs + 6 * SBSZ;
    double* const U0 = Xs + 7 * SBSZ, * const U1 = Xs + 8 * SBSZ;
    for (int k = 0; k < T; k++) {
        const long long o = (long long)k * CT;
        if ((long long)a.n_valid - o < CT) break;                   // identity-padded last block: CTA 0 publishes it itself
        const double* Ak = a.A + o * a.ld + o;
        double* Wk = a.W + o * a.ldw + o;
        double* Vk = a.V ? a.V + o * a.ldv + o : nullptr;
        auto load_L = [&](int bi, int bj) {
            for (int e = tid; e < SB * (SB / 2); e += CTHREADS) {
                const int r = e >> 4, c2 = (e & 15) * 2;
                cp_async16(Lb + blk(bi, bj) * SBSZ + r * SLD + c2, Ak + (long long)(bi * SB + r) * a.ld + bj * SB + c2);
            }
        };
        auto load_W = [&](int b) {
            for (int e = tid; e < SB * (SB / 2); e += CTHREADS) {
                const int r = e >> 4, c2 = (e & 15) * 2;
                cp_async16(Wdp(b) + r * SLD + c2, Wk + (long long)(b * SB + r) * a.ldw + b * SB + c2);
            }
        };
#pragma unroll 1
        for (int ph = 0; ph < 5; ph++) {
            // phases 0,1 after column 1; 2,3 after column 2; 4 after column 3
            if (ph == 0 || ph == 2 || ph == 4) {
                const int col = (ph == 0) ? 1 : (ph == 2 ? 2 : 3);
                if (tid == 0) s_act[0] = spin_ge(f_lpub(a, k, col), 1, err) ? 1 : 0;
                __syncthreads();
                if (!s_act[0]) return;
                if (ph == 0) { load_L(1, 0); load_W(0); load_W(1); }
                else if (ph == 2) { load_L(2, 0); load_L(2, 1); load_L(3, 0); load_L(3, 1); load_L(3, 2); load_W(2); }
                else load_W(3);
                cp_async_commit();
                cp_async_wait<0>();
                __syncthreads();
            }
            D512Strip st;
            bool have = false;
            if (ph == 0 && wid < 2) {
                st = d512_prod(Lb + blk(1, 0) * SBSZ, Wdp(0), nullptr, nullptr, 1.0, S10, -1, 0, wid);
                have = true;
            } else if (ph == 1 && wid < 2) {
                st = d512_prod(Wdp(1), S10, nullptr, nullptr, -1.0, W10, 1, 0, wid);
                have = true;
            } else if (ph == 2 && wid < 10) {
                if (wid < 8) {
                    const int i = 2 + (wid >> 2), j = (wid >> 1) & 1, hf = wid & 1;
                    double* Sdst = (i == 2) ? (j ? S21 : S20) : (j ? S31 : S30);
                    st = (j == 0) ? d512_prod(Lb + blk(i, 0) * SBSZ, Wdp(0), Lb + blk(i, 1) * SBSZ, W10, 1.0, Sdst, -1, 0, hf)
                                  : d512_prod(Lb + blk(i, 1) * SBSZ, Wdp(1), nullptr, nullptr, 1.0, Sdst, -1, 0, hf);
                } else {
                    st = d512_prod(Lb + blk(3, 2) * SBSZ, Wdp(2), nullptr, nullptr, 1.0, S32, -1, 0, wid - 8);
                }
                have = true;
            } else if (ph == 3 && wid < 8) {
                const int j = (wid >> 1) & 1, hf = wid & 1;
                if (wid < 4) {
                    st = d512_prod(Wdp(2), j ? S21 : S20, nullptr, nullptr, -1.0, nullptr, 2, j, hf);
                } else {                                  // U_j = S_3j - S_32 S_2j
                    st = d512_prod(S32, j ? S21 : S20, nullptr, nullptr, -1.0, j ? U1 : U0, -1, 0, hf);
                    st.Cin = j ? S31 : S30; st.cin_s = SLD;
                }
                have = true;
            } else if (ph == 4 && wid < 6) {
                if (wid < 2) st = d512_prod(Wdp(3), S32, nullptr, nullptr, -1.0, nullptr, 3, 2, wid);
                else st = d512_prod(Wdp(3), ((wid - 2) >> 1) ? U1 : U0, nullptr, nullptr, -1.0, nullptr, 3, (wid - 2) >> 1, wid & 1);
                have = true;
            }
            if (have) d512_strip<8>(st, Wk, a.ldw, Vk, a.ldv);
            __syncthreads();
        }
        if (tid == 0) {
            __threadfence();
            st_release(f_diag(a, k), 1);
            a.tclk[(long long)2 * a.T * a.T * 4 + k * 4 + 0] = gtimer_ns();
            a.wclk[g_hclk_off + 8 * 64 + k * 4 + 2] = gtimer_ns();
        }
    }
}

// CTA 0 of the pipelined group: diagonal blocks only
__device__ __forceinline__ void chain0_v2(const ChainArgs& a, double* csm, volatile int* s_act, int tid) {
    int* err = a.flags;
    const int T = a.T;
    for (int k = 0; k < T; k++) {
        CHAIN_STAMP(k, 0);
        if (k > 0) {
            if (tid == 0) s_act[0] = spin_ge(f_sp(a, k), NH2, err) ? 1 : 0;
            __syncthreads();
            if (!s_act[0]) { if (tid == 0) *a.info = -999; return; }
        }
        CHAIN_STAMP(k, 5);
        if (tid == 0) a.tclk[(long long)2 * a.T * a.T * 4 + k * 4 + 3] = gtimer_ns();
        const long long o = (long long)k * CT;
        const long long valid = (long long)a.n_valid - o;
        const int nsub = valid >= CT ? 4 : (valid <= 0 ? 0 : (int)((valid + SB - 1) / SB));
        long long* dclk = (k == 1) ? a.wclk + (size_t)g_dclk_off : nullptr;
        double* Ak = a.A + o * a.ld + o;
        double* Wk = a.W + o * a.ldw + o;
        double* Vk = a.V ? a.V + o * a.ldv + o : nullptr;
        if (nsub == 4) {
            diag_block_body512<8>(Ak, a.ld, Wk, a.ldw, Vk, a.ldv, a.info, (int)o, csm, dclk, f_lpub(a, k, 0));
            __syncthreads();
        } else {
            if (tid < 256) diag_block_body<true>(Ak, a.ld, Wk, a.ldw, Vk, a.ldv, a.info, (int)o, nsub, csm);
            __syncthreads();
            if (tid == 0) {                      // a padded block is the last one: publish everything at once
                __threadfence();
                for (int bb = 0; bb < 4; bb++) st_release(f_lpub(a, k, bb), 1);
                st_release(f_diag(a, k), 1);
                a.tclk[(long long)2 * a.T * a.T * 4 + k * 4 + 0] = gtimer_ns();
            a.tclk[(long long)2 * a.T * a.T * 4 + k * 4 + 0] = gtimer_ns();
            }
        }
        CHAIN_STAMP(k, 6);
        if (tid == 0) a.wclk[g_hclk_off + 8 * 64 + k * 4 + 3] = gtimer_ns();
        CHAIN_STAMP(k, 7);
    }
}

__global__ void __launch_bounds__(CTHREADS, 1) potrf_dataflow_kernel(const ChainArgs a) {
    extern __shared__ __align__(16) double csm[];
    __shared__ int s_act[4];
    __shared__ int s_task[12];
    const int tid = threadIdx.x;
    if ((int)blockIdx.x < a.NG) {
        if (a.pipelined) {
            if (blockIdx.x == 0) chain0_v2(a, csm, s_act, tid);
            else if ((int)blockIdx.x <= NH2) helper_v2(a, csm, s_act, tid, (int)blockIdx.x - 1);
            else inverter_v2(a, csm, s_act, tid);
            return;
        }
        if (a.NG == 8) chain_group<16>(a, csm, s_act, tid);
        else chain_group<32>(a, csm, s_act, tid);
        return;
    }
    const int grp = tid >> 8;
    if (a.tiles) worker_group_edf(a, csm + grp * G_RING, s_act, s_task, tid & 255, grp);
    else worker_group(a, csm + grp * G_RING, s_act, tid & 255, grp);
}

// ---- host side: task lists per (T, grid, NG) ------------------------------------------------------------
struct ChainPlan {
    ChainTask* bulk = nullptr; int* bulk_off = nullptr;
    ChainTask* trsm = nullptr; int* trsm_off = nullptr;
    ChainTile* tiles = nullptr; int* tile_off = nullptr;    // null when a group would own more than 32 half tiles
};
std::map<std::pair<std::pair<int, int>, int>, ChainPlan> g_plans;
std::map<cudaStream_t, std::pair<int*, size_t>> g_flag_pool;
std::map<cudaStream_t, int> g_last_T;
std::mutex g_plan_mu;

int build_plan(int T, int G, int NG, int mform, int band, ChainPlan* out) {
    const bool pipelined = (NG == NH2 + 2);
    const bool express = gpb_get_option("chain_express") == 1;
    int band_x = gpb_get_option("chain_band_x");
    if (band_x <= 0) band_x = 2;
    NG += 1000 * mform + 10000 * band + 1000000 * band_x;              // (plan cache key only)
    std::lock_guard<std::mutex> lk(g_plan_mu);
    auto it = g_plans.find({{T, G}, NG + (express ? 100 : 0)});
    if (it != g_plans.end()) { *out = it->second; return GPB_OK; }
    const int nv = 2 * (G - NG % 1000);     // worker groups
    const int nW = nv / 2;                   // worker CTAs; group index v = grp * nW + cta (see worker_group_edf)
    // pipelined group: worker groups 0..3 are EXPRESS groups.  Per step s they run exactly the tasks the chain group
    // waits for one step later -- TRSM(s+3, s) and the step-s updates of tiles (s+3, s+1), (s+3, s+2), (s+3, s+3) -- and
    // nothing else, so those tasks never queue behind a regular group's 20-40 k-cycle backlog.
    // Six express roles, each ALONE on its SM (group 0 of the first six worker CTAs; their group 1 stays empty, so an
    // express task has the SM's DMMA pipes to itself): TRSM halves | tile (s+3, s+1) halves | tiles (s+3, s+2) then
    // (s+3, s+3) halves.
    // Measured (N = 1024 / 2048 / 4096): 0.348 / 0.721 / 1.85 ms with the express groups against 0.345 / 0.694 / 1.81
    // without -- the helpers' waits only move to the next tile back along the row (every row's TRSM -> update chain
    // runs at the regular groups' pace before it reaches the express zone) -- so they are OFF unless asked for.
    const int nexp = (pipelined && express && nv >= 24) ? 12 : 0;
    // Half tile (i,h,j) receives j updates (steps 0..j-1), and at step k exactly the tiles with j > k are
    // live.  Dealing the half tiles out in order of decreasing j (boustrophedon over the groups) therefore
    // balances every live set, i.e. every step of the factorisation, to within one task per group.
    std::vector<int> own((size_t)2 * T * T, 0);
    {
        long long n = 0;
        for (int j = T - 1; j >= 0; j--)
            for (int i = j; i < T; i++)
                for (int h = 0; h < 2; h++) {
                    const int nreg = nv - nexp;
                    const long long round = n / nreg, pos = n % nreg;
                    own[((size_t)2 * i + h) * T + j] = nexp + (int)((round & 1) ? nreg - 1 - pos : pos);
                    n++;
                }
    }
    auto owner = [&](int i, int h, int j) { return own[((size_t)2 * i + h) * T + j]; };
    std::vector<std::vector<ChainTask>> bulk(nv), trsm(nv);
    for (int k = 0; k + 1 < T; k++) {
        // pipelined group: the helpers also run the TRSM of row k+2 and the last worker-side step of the two chain
        // tiles of every row (their "urgent phase")
        for (int i = k + (pipelined ? 3 : 2); i < T; i++)
            for (int h = 0; h < 2; h++)
                trsm[(nexp && i == k + 3) ? 2 * h : owner(i, h, k)].push_back({TASK_TRSM, (unsigned char)h, (short)i, (short)k, (short)k});
        for (int j = k + 1; j < T; j++)
            for (int i = j; i < T; i++) {
                if (i == k + 1 && j == k + 1) continue;            // the chain group's own update
                if (pipelined && ((i == j && k == j - 2) || (i == j + 1 && k == j - 1))) continue;
                for (int h = 0; h < 2; h++) {
                    int who = owner(i, h, j);
                    if (nexp && i == j + 2 && k == j - 1) who = 2 * (2 + h);  // tile (s+3, s+1) at step s: input of the urgent TRSM
                    if (nexp && i == j + 1 && k == j - 2) who = 2 * (4 + h);  // tile (s+3, s+2) at step s
                    if (nexp && i == j && k == j - 3) who = 2 * (4 + h);      // tile (s+3, s+3) at step s
                    bulk[who].push_back({TASK_UPD, (unsigned char)h, (short)i, (short)j, (short)k});
                }
            }
    }
    auto upload = [&](std::vector<std::vector<ChainTask>>& lists, ChainTask** dt, int** doff) -> int {
        std::vector<ChainTask> flat;
        std::vector<int> off(nv + 1, 0);
        for (int g = 0; g < nv; g++) {
            off[g] = (int)flat.size();
            flat.insert(flat.end(), lists[g].begin(), lists[g].end());
        }
        off[nv] = (int)flat.size();
        GPB_CUDA(cudaMalloc(dt, (flat.size() + 1) * sizeof(ChainTask)));
        GPB_CUDA(cudaMalloc(doff, (nv + 1) * sizeof(int)));
        if (!flat.empty()) GPB_CUDA(cudaMemcpy(*dt, flat.data(), flat.size() * sizeof(ChainTask), cudaMemcpyHostToDevice));
        GPB_CUDA(cudaMemcpy(*doff, off.data(), (nv + 1) * sizeof(int), cudaMemcpyHostToDevice));
        return GPB_OK;
    };
    ChainPlan p;
    int stt = upload(bulk, &p.bulk, &p.bulk_off);
    if (stt) return stt;
    stt = upload(trsm, &p.trsm, &p.trsm_off);
    if (stt) return stt;
    // scheduler 0: the same ownership as tile lists, most urgent (lowest column) first
    if (!nexp) {
        std::vector<ChainTile> flat;
        std::vector<int> off(nv + 1, 0);
        std::vector<std::vector<ChainTile>> tl(nv);
        size_t most = 0;
        // band groups: group 0 of the first `band` worker CTAs (their group 1 stays empty); owner() deals over the rest
        std::vector<int> regular;
        for (int g = 0; g < nv; g++)
            if (!(band > 0 && (g % nW) < band)) regular.push_back(g);
        if (band > 0) {
            long long n = 0;
            const int nreg = (int)regular.size();
            for (int j = T - 1; j >= 0; j--)
                for (int i = j; i < T; i++)
                    for (int h = 0; h < 2; h++) {
                        const long long round = n / nreg, pos = n % nreg;
                        own[((size_t)2 * i + h) * T + j] = (int)((round & 1) ? nreg - 1 - pos : pos);
                        n++;
                    }
        }
        // a tile's deadline = the step that consumes its last worker-side update: column j for most, one or two
        // steps earlier for the tiles the chain group finishes itself
        std::vector<std::pair<std::pair<int, int>, std::pair<int, int>>> order;       // ((deadline, j), (i, h))
        for (int j = 0; j < T; j++)
            for (int i = j; i < T; i++)
                for (int h = 0; h < 2; h++) {
                    int dl = j;
                    if (pipelined) dl = j - (i == j ? (mform >= 3 ? 3 : 2) : (i == j + 1 ? 1 : 0));
                    else dl = j - (i == j ? 1 : 0);
                    order.push_back({{dl, j}, {i, h}});
                }
        std::sort(order.begin(), order.end());
        long long band_next = 0;          // band tiles are dealt round-robin in deadline order: every step's finishing tiles spread over all band groups
        for (auto& o : order) {
            const int j = o.first.second, i = o.second.first, h = o.second.second;
            int nupd = j, has_trsm, has_m = 0;
            if (pipelined) {
                has_trsm = i >= j + 3;
                if (i == j) nupd = j - 2;           // steps j-2 (urgent phase) and j-1 (pipelined columns) are the helpers'
                if (i == j + 1) nupd = j - 1;       // step j-1 is the helpers' urgent phase
                // M form: the tile's last worker-side step (kappa = nupd - 1) reads A'(i,kappa) instead of L(i,kappa)
                // M form (level 1): step j-1 of the tiles below the first sub-diagonal reads A'(i,j-1) and M_{j,j-1}
                // instead of L(i,j-1).  Level 3: also step j-2 of the first sub-diagonal tiles (M_{j,j-2}), and step j-3
                // of the diagonal tiles moves to the helpers (folded into their pipelined columns).  Level 4: step j-2
                // of every tile below the diagonal in M form.
                if (mform >= 3 && i == j) nupd = j - 3;
                if (mform >= 4 && i >= j + 1) {
                    nupd = j - 2;
                    if (j >= 2) has_m |= 2;
                    if (i >= j + 2 && j >= 1) has_m |= 1;
                } else if (mform >= 1) {
                    if (i >= j + 2 && j >= 1) { nupd = j - 1; has_m |= 1; }
                    if (mform >= 3 && i == j + 1 && j >= 2) { nupd = j - 2; has_m |= 2; }
                }
            } else {
                has_trsm = i >= j + 2;
                if (i == j) nupd = j - 1;
            }
            if (nupd < 0) nupd = 0;
            const int dl = o.first.first < 0 ? 0 : o.first.first;
            if (band > 0 && i - j <= 3) {
                // Diagonal band (tiles (k,k) ... (k,k-3): everything the helpers' urgent phase reads, plus the TRSM of
                // row k-3+3 that feeds it).  The regular owner applies all but the last band_x L-form steps and hands the
                // tile over; a band group -- alone on its SM, so a task takes 12 us instead of 23 and never queues behind
                // a backlog task -- applies the rest, the M-form step and the TRSM.
                const int R = nupd > band_x ? nupd - band_x : 0;
                if (R > 0) {
                    auto& l = tl[regular[owner(i, h, j)]];
                    l.push_back({(short)i, (short)j, (unsigned char)h, (unsigned char)R, 0, 0, (short)(dl > band_x ? dl - band_x : 0), 0, 1});
                    if (l.size() > most) most = l.size();
                }
                auto& l = tl[(band_next++) % band];
                l.push_back({(short)i, (short)j, (unsigned char)h, (unsigned char)nupd, (unsigned char)has_trsm, (unsigned char)has_m,
                             (short)dl, (unsigned char)R, 0});
                if (l.size() > most) most = l.size();
                continue;
            }
            auto& l = tl[regular[owner(i, h, j)]];
            l.push_back({(short)i, (short)j, (unsigned char)h, (unsigned char)nupd, (unsigned char)has_trsm, (unsigned char)has_m,
                         (short)dl, 0, 0});
            if (l.size() > most) most = l.size();
        }
        if (most <= 32) {
            for (int g = 0; g < nv; g++) {
                off[g] = (int)flat.size();
                flat.insert(flat.end(), tl[g].begin(), tl[g].end());
            }
            off[nv] = (int)flat.size();
            GPB_CUDA(cudaMalloc(&p.tiles, (flat.size() + 1) * sizeof(ChainTile)));
            GPB_CUDA(cudaMalloc(&p.tile_off, (nv + 1) * sizeof(int)));
            GPB_CUDA(cudaMemcpy(p.tiles, flat.data(), flat.size() * sizeof(ChainTile), cudaMemcpyHostToDevice));
            GPB_CUDA(cudaMemcpy(p.tile_off, off.data(), (nv + 1) * sizeof(int), cudaMemcpyHostToDevice));
        }
    }
    if (band > 0 && !p.tiles) { gpb_set_error("dataflow plan: diagonal band does not fit 32 tiles per group"); return GPB_ERR_ARG; }
    g_plans[{{T, G}, NG + (express ? 100 : 0)}] = p;
    *out = p;
    return GPB_OK;
}

int g_num_sms = 0;
int chain_init() {
    static bool done = false;
    if (done) return GPB_OK;
    int dev = 0, coop = 0;
    GPB_CUDA(cudaGetDevice(&dev));
    GPB_CUDA(cudaDeviceGetAttribute(&g_num_sms, cudaDevAttrMultiProcessorCount, dev));
    GPB_CUDA(cudaDeviceGetAttribute(&coop, cudaDevAttrCooperativeLaunch, dev));
    GPB_REQUIRE(coop != 0, "device does not support cooperative launches");
    GPB_CUDA(cudaFuncSetAttribute(potrf_dataflow_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, C_SMEM));
    done = true;
    return GPB_OK;
}

// flags and clocks (zeroed before every launch), then the M tiles
size_t chain_zero_words(int T) { return (chain_flag_words(T) + 3) / 4 * 4 + (size_t)T * 16 + (size_t)(g_hclk_off + 12 * 64 + 8) * 2; }
size_t chain_trace_lls(int T) { return (size_t)2 * T * T * 4 + (size_t)T * 4; }
size_t chain_pool_words(int T) { return chain_zero_words(T) + (size_t)3 * T * CT * CT * 2 + chain_trace_lls(T) * 2; }

}  // namespace

// The largest T the dataflow factorisation takes (N = 6144; measured 4.80 vs 5.72 ms there, 9.9 vs 9.3 ms at
// N = 8192): beyond it the trailing matrix no longer
// lives in L2 and the K = 128 updates re-stream it from HBM every step; potrf.cu's look-ahead panels win.
bool gpb_potrf_dataflow_ok(long long n, int batch) {
    const long long T = n / GPB_NB;
    return batch == 1 && T >= 2 && T <= 64;
}

// zero_blocks and info initialisation are the caller's (gpb_launch_potrf) business
int gpb_launch_potrf_dataflow(double* A, long long n, long long ld, double* W, long long ldw, double* V,
                              long long ldv, int* info, cudaStream_t st, long long n_valid) {
    const int T = (int)(n / GPB_NB);
    GPB_REQUIRE((reinterpret_cast<uintptr_t>(A) & 15) == 0 && ld % 2 == 0, "A must be 16-byte aligned with an even leading dimension");
    int stt = chain_init();
    if (stt) return stt;
    const int num_sms = g_num_sms;
    const long long nhalf = (long long)T * (T + 1);                // half tiles
    int NG = gpb_get_option("chain_group");          // CTAs sharing the critical path: 10 = pipelined (default), 8 or 4 = v1
    const int pipelined = (NG != 4 && NG != 8) ? 1 : 0;
    if (pipelined) NG = NH2 + 2;
    GPB_REQUIRE(num_sms >= 2 * NG, "device too small for the dataflow factorisation");
    int G = (int)(((nhalf + 1) / 2 + NG < (long long)num_sms) ? (nhalf + 1) / 2 + NG : num_sms);
    if (G < NG + 1) G = NG + 1;
    ChainPlan plan;
    const bool lists = gpb_get_option("chain_sched") == 1;       // 1 = in-order task lists, 0 = most urgent runnable tile first
    // 0 = M form whenever the tile scheduler and the pipelined group run, 2 = off
    int mform = gpb_get_option("chain_mform");      // 0 -> level 1; 2 = off; 1, 3, 4 = levels (build_plan)
    if (mform == 0) mform = 1;
    if (mform == 2 || !pipelined || lists || gpb_get_option("chain_express") == 1) mform = 0;
    // diagonal band on dedicated SMs: chain_band = number of band SMs (0 -> default, 1 = off)
    int band = gpb_get_option("chain_band");
    if (band == 0) band = 1;
    if (band == 1 || mform != 1 || T < 12 || T > 40 || G < num_sms) band = 0;
    stt = build_plan(T, G, NG, mform, band, &plan);
    if (stt) return stt;
    // flag words: one grow-only set per stream (two factorisations on one stream are serialised anyway)
    // (+ the chain's phase clocks, 8 per step, and the workers' accounting behind the flags)
    const size_t nfl = chain_pool_words(T);
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
    GPB_CUDA(cudaMemsetAsync(flags, 0, chain_zero_words(T) * sizeof(int), st));
    ChainArgs a;
    a.A = A; a.ld = ld; a.W = W; a.ldw = ldw; a.V = V; a.ldv = ldv; a.info = info;
    a.T = T; a.n_valid = (int)n_valid; a.flags = flags;
    a.bulk = plan.bulk; a.bulk_off = plan.bulk_off; a.trsm = plan.trsm; a.trsm_off = plan.trsm_off;
    a.tiles = lists ? nullptr : plan.tiles; a.tile_off = lists ? nullptr : plan.tile_off;
    a.clk = reinterpret_cast<long long*>(flags + (chain_flag_words(T) + 3) / 4 * 4);
    a.wclk = a.clk + (size_t)T * 8;
    a.M = reinterpret_cast<double*>(flags + chain_zero_words(T));
    a.tclk = reinterpret_cast<long long*>(a.M + (size_t)3 * T * CT * CT);
    a.mform = a.tiles ? mform : 0;
    a.fuse = gpb_get_option("chain_fuse");            // 0 -> default
    a.imminent = gpb_get_option("chain_imminent");
    if (a.imminent <= 0) a.imminent = 1;
    a.horizon = gpb_get_option("chain_horizon");          // 0 -> default; 100 = off
    if (a.horizon <= 0) a.horizon = 3;
    if (a.horizon >= 100) a.horizon = 1 << 20;
    a.fuse_guard = gpb_get_option("chain_fuse_guard");
    if (a.fuse_guard <= 0) a.fuse_guard = 2;
    if (a.fuse_guard >= 100) a.fuse_guard = -1000;       // (off)
    // measured before the yield policy: 1 / 2 / 3 / 4 / 8 at N = 4096: 1.68 / 1.65 / 1.60 / 1.59 / 1.59 ms, N = 6144: 4.15 / 3.83 / 3.72 /
    // 3.69 / 3.79; with it (far tiles no longer start right before urgent ones): 4 / 8 / 12 at N = 6144: 3.54 / 3.49 / 3.48,
    // N = 8192: 7.46 / 7.29 / 7.29
    if (a.fuse <= 0) a.fuse = 8;
    a.NG = NG;
    a.pipelined = pipelined;
    a.diag512 = gpb_get_option("chain_diag");           // 0/1 default body, 2 the 256-thread body, 3.. experiments
    if (a.diag512 == 0) a.diag512 = 1;
    void* args[] = {(void*)&a};
    GpbProfScope prof(GPB_KC_GEMM, st);
    GPB_CUDA(cudaLaunchCooperativeKernel((const void*)potrf_dataflow_kernel, dim3((unsigned)G), dim3(CTHREADS), args,
                                         (size_t)C_SMEM, st));
    GPB_LAUNCH_CHECK("potrf_dataflow_kernel");
    return GPB_OK;
}

static int chain_debug_pool(void* stream, int** flags, int* T) {
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
    {
        std::lock_guard<std::mutex> lk(g_plan_mu);
        auto it = g_flag_pool.find(st);
        if (it == g_flag_pool.end() || !g_last_T.count(st)) { gpb_set_error("no dataflow factorisation ran on this stream"); return GPB_ERR_ARG; }
        *flags = it->second.first;
        *T = g_last_T[st];
    }
    GPB_CUDA(cudaStreamSynchronize(st));
    return GPB_OK;
}

// phase clocks (SM cycles) of the chain CTA in the last dataflow factorisation on `stream`:
// out[k*8 + p], p = 0 step start | 1 sub-diagonal tile ready | 2 TRSM done | 3 published | 4 diagonal tile
// ready | 5 SYRK done | 6 diagonal block done | 7 published.  Returns T (steps) or < 0.
extern "C" int gpb_debug_chain_clocks(void* stream, long long* out, int max_steps) {
    int* flags = nullptr;
    int T = 0;
    int stt = chain_debug_pool(stream, &flags, &T);
    if (stt) return stt;
    const int n = T < max_steps ? T : max_steps;
    const long long* clk = reinterpret_cast<const long long*>(flags + (chain_flag_words(T) + 3) / 4 * 4);
    GPB_CUDA(cudaMemcpy(out, clk, (size_t)n * 8 * sizeof(long long), cudaMemcpyDeviceToHost));
    return T;
}

// %globaltimer trace of the same launch: [2T][T][4] per half tile (last update start | complete | TRSM start | TRSM done), then
// [T][4] per step (DIAG published | helper 0's M1 | M2 arrival | the chain CTA starts the diagonal block).  Returns T.
extern "C" int gpb_debug_chain_tiles(void* stream, long long* out, long long max_lls) {
    int* flags = nullptr;
    int T = 0;
    int stt = chain_debug_pool(stream, &flags, &T);
    if (stt) return stt;
    size_t n = chain_trace_lls(T);
    if ((long long)n > max_lls) n = (size_t)max_lls;
    const long long* src = reinterpret_cast<const long long*>(reinterpret_cast<const double*>(flags + chain_zero_words(T)) + (size_t)3 * T * CT * CT);
    GPB_CUDA(cudaMemcpy(out, src, n * sizeof(long long), cudaMemcpyDeviceToHost));
    return T;
}

// per-worker-group accounting of the same launch: out[v*4 + {0 cycles waiting, 1 in TRSM, 2 in UPD, 3 tasks}]
extern "C" int gpb_debug_chain_workers(void* stream, long long* out, int max_groups) {
    int* flags = nullptr;
    int T = 0;
    int stt = chain_debug_pool(stream, &flags, &T);
    if (stt) return stt;
    if (max_groups > (g_hclk_off + 12 * 64) / 4) max_groups = (g_hclk_off + 12 * 64) / 4;
    const long long* clk = reinterpret_cast<const long long*>(flags + (chain_flag_words(T) + 3) / 4 * 4) + (size_t)T * 8;
    GPB_CUDA(cudaMemcpy(out, clk, (size_t)max_groups * 4 * sizeof(long long), cudaMemcpyDeviceToHost));
    return T;
}

// ---- micro-benchmark of the half-tile task (no flags): both groups of every CTA repeat UPD / TRSM ----------
namespace {
__global__ void __launch_bounds__(CTHREADS, 1) chain_tile_bench_kernel(const ChainArgs a, int reps, int mode, long long* out) {
    extern __shared__ __align__(16) double csm[];
    const int tid = threadIdx.x, grp = tid >> 8, ltid = tid & 255;
    // tiles of CTA c: rows 3c + {0, 1, 2}; L_i = (2,0), L_j = (1,0), C = (2,1) of a [3 * grid * 128, 384] matrix
    ChainArgs b = a;
    b.A = a.A + (long long)blockIdx.x * 3 * CT * a.ld;
    if (grp == 1 && (mode & 2)) {           // stagger the groups by half a task
        const long long ts = clock64();
        while (clock64() - ts < 20000) {}
    }
    const long long t0 = clock64();
    for (int r = 0; r < reps; r++) {
        if ((mode & 1) == 0) task_upd(b, 2, grp, 1, 0, csm + grp * G_RING, ltid, grp);
        else task_trsm(b, 2, grp, 0, csm + grp * G_RING, ltid, grp);
        publish_group(a.flags + 1 + blockIdx.x * 2 + grp, r, ltid, grp);
    }
    if (ltid == 0) out[blockIdx.x * 2 + grp] = clock64() - t0;
}
}  // namespace

// cycles for `reps` half-tile tasks per worker group.  A: device [3 * grid * 128, 384] (ld >= 384), W: [128, >= 128]
extern "C" int gpb_debug_tile_bench(double* A, long long ld, double* W, long long ldw, int grid, int reps, int mode,
                                    int* flags, long long* out_dev, void* stream) {
    static bool attr_set = false;
    if (!attr_set) {
        GPB_CUDA(cudaFuncSetAttribute(chain_tile_bench_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, C_SMEM));
        attr_set = true;
    }
    ChainArgs a;
    a.tiles = nullptr; a.tile_off = nullptr; a.mform = 0; a.fuse = 1; a.fuse_guard = 2; a.horizon = 1 << 20; a.imminent = 1; a.M = nullptr; a.tclk = nullptr;
    a.A = A; a.ld = ld; a.W = W; a.ldw = ldw; a.V = nullptr; a.ldv = 0; a.info = nullptr; a.T = 3; a.n_valid = 0;
    a.NG = 8; a.pipelined = 0; a.diag512 = 1; a.flags = flags; a.bulk = nullptr; a.bulk_off = nullptr; a.trsm = nullptr; a.trsm_off = nullptr;
    a.clk = nullptr; a.wclk = nullptr;
    chain_tile_bench_kernel<<<grid, CTHREADS, C_SMEM, reinterpret_cast<cudaStream_t>(stream)>>>(a, reps, mode, out_dev);
    GPB_LAUNCH_CHECK("chain_tile_bench_kernel");
    return GPB_OK;
}
