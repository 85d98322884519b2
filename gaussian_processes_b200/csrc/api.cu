// api.cu -- extern "C" surface of libgpb200.so (declared in include/gpb200.h).
#include <stdarg.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <vector>
#include "../../include/gpb200.h"
#include "kfunctors.cuh"
#include "launch.h"

#include <mutex>
long long g_gpb_launches = 0;
static thread_local char g_err[512] = "";

// The library keeps process-global staging state (pinned rings and slots, the device pool, the fork/join
// events of the candidate groups, profiling lists).  ctypes drops the GIL during a call, so two host
// threads can be inside the library at once: every entry point that touches that state holds this lock
// for its duration (recursive: the host-buffer entry points call the device-pointer ones).  Entry points
// that only enqueue kernels on the caller's stream do not take it.
static std::recursive_mutex g_api_mu;
#define GPB_API_LOCK std::lock_guard<std::recursive_mutex> gpb_api_guard(g_api_mu)
std::recursive_mutex& gpb_api_mutex() { return g_api_mu; }     // for the other translation units' host entry points

void gpb_set_error(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}

int gpb_check_cuda(cudaError_t e, const char* what) {
    if (e == cudaSuccess) return GPB_OK;
    gpb_set_error("CUDA error %d (%s) at %s", (int)e, cudaGetErrorString(e), what);
    return GPB_ERR_CUDA;
}

// ---- per-class timing -------------------------------------------------------------
int g_gpb_profile = 0;
struct ProfPair { cudaEvent_t a, b; int cls; };
static std::vector<ProfPair> g_prof_pairs;
static std::vector<cudaEvent_t> g_prof_open[GPB_KC_COUNT];
static double g_prof_ms[GPB_KC_COUNT];
static long long g_prof_n[GPB_KC_COUNT];

void gpb_prof_begin(int cls, cudaStream_t st) {
    cudaEvent_t e;
    if (cudaEventCreate(&e) != cudaSuccess) return;
    cudaEventRecord(e, st);
    g_prof_open[cls].push_back(e);
}
void gpb_prof_end(int cls, cudaStream_t st) {
    if (g_prof_open[cls].empty()) return;
    cudaEvent_t b;
    if (cudaEventCreate(&b) != cudaSuccess) return;
    cudaEventRecord(b, st);
    ProfPair p = {g_prof_open[cls].back(), b, cls};
    g_prof_open[cls].pop_back();
    g_prof_pairs.push_back(p);
}
static void prof_collect() {
    for (auto& p : g_prof_pairs) {
        float ms = 0.f;
        if (cudaEventSynchronize(p.b) == cudaSuccess && cudaEventElapsedTime(&ms, p.a, p.b) == cudaSuccess) {
            g_prof_ms[p.cls] += ms;
            g_prof_n[p.cls] += 1;
        }
        cudaEventDestroy(p.a);
        cudaEventDestroy(p.b);
    }
    g_prof_pairs.clear();
}

// ---- run-time options -------------------------------------------------------------
struct GpbOption { const char* name; const char* env; int value; bool env_read; };
static GpbOption g_options[] = {
    {"eval_streams", "GPB_EVAL_STREAMS", 0, false},   // candidate groups evaluated on concurrent streams
    {"gemm_bm", "GPB_GEMM_BM", 0, false},             // 64 (2 CTAs/SM) or 128 row tiles
    {"potrf_inner", "GPB_POTRF_INNER", 0, false},     // 128-columns per outer Cholesky panel
    {"potrf_lookahead", "GPB_POTRF_LOOKAHEAD", 0, false},   // 0 = panel look-ahead for one matrix of N >= 6144, 1 = always, 2 = off
    {"potrf_panel_rl", "GPB_POTRF_PANEL_RL", 0, false},     // 0 = right-looking panel steps for one matrix, 1 = always, 2 = never
    {"gemm_impl", "GPB_GEMM_IMPL", 0, false},               // 0 = TMA + mbarrier pipeline, 1 = cp.async pipeline
    {"potrf_dataflow", "GPB_POTRF_DATAFLOW", 0, false},     // 0 = one matrix of 256 <= N <= 6144 by the persistent dataflow launch, 1 = whenever batch == 1, 2 = never
    {"chain_diag", "GPB_CHAIN_DIAG", 0, false},             // 2 = the chain CTA factors diagonal blocks with the 256-thread body
    {"chain_express", "GPB_CHAIN_EXPRESS", 0, false},       // 1 = six express worker groups for the chain-adjacent tiles (measured: no gain)
    {"chain_mform", "GPB_CHAIN_MFORM", 0, false},           // M form of the workers' last update(s) of a tile: 0/1 = step j-1, 3, 4 = more (chain.cu build_plan), 2 = off
    {"chain_fuse", "GPB_CHAIN_FUSE", 0, false},             // most backlog steps of a half tile applied by one worker task (default 8)
    {"chain_fuse_guard", "GPB_CHAIN_FUSE_GUARD", 0, false}, // no fused task while a more urgent tile of the group is due within this many steps (default 2, 100 = off)
    {"stage_overlap", "GPB_STAGE_OVERLAP", 0, false},       // 2 = gpb_gp_stages keeps the triangular solves on the caller's stream
    {"chain_horizon", "GPB_CHAIN_HORIZON", 0, false},       // far tiles (deadline > step + horizon) yield while their group has an imminent tile (default 3, 100 = off)
    {"chain_imminent", "GPB_CHAIN_IMMINENT", 0, false},     // ... "imminent": due within this many steps (default 1)
    {"chain_band", "GPB_CHAIN_BAND", 0, false},             // SMs dedicated to the tiles next to the diagonal (1 = off)
    {"chain_band_x", "GPB_CHAIN_BAND_X", 0, false},         // L-form steps of a band tile taken over by its band group (default 2)
    {"lauum_fuse", "GPB_LAUUM_FUSE", 0, false},             // 1 = gradient brackets in the epilogue of the lauum GEMM (K^-1 not stored in the batched evaluator); measured slower, default off
    {"chain_sched", "GPB_CHAIN_SCHED", 0, false},           // workers: 0 = most urgent runnable half tile first, 1 = in-order task lists
    {"chain_group", "GPB_CHAIN_GROUP", 0, false},           // 0 = pipelined chain group (sweeping CTA + 8 helpers + inverter), 8 or 4 = the first chain group of that many CTAs
};
int gpb_get_option(const char* name) {
    for (auto& o : g_options)
        if (strcmp(o.name, name) == 0) {
            if (!o.env_read) {
                o.env_read = true;
                const char* e = getenv(o.env);
                if (e && o.value == 0) o.value = atoi(e);
            }
            return o.value;
        }
    return 0;
}

// ---- pinned staging ring: small host->device parameter uploads without stalling ------
struct PinSlot { void* host; size_t cap; cudaEvent_t ev; bool used; };
static PinSlot g_pin[32];
static int g_pin_next = 0;
static int pin_acquire(size_t bytes, int* slot, void** host) {
    const int sidx = g_pin_next++ % 32;
    PinSlot& p = g_pin[sidx];
    if (p.used) GPB_CUDA(cudaEventSynchronize(p.ev));
    if (p.cap < bytes) {
        if (p.host) GPB_CUDA(cudaFreeHost(p.host));
        p.host = nullptr; p.cap = 0;
        GPB_CUDA(cudaHostAlloc(&p.host, bytes, cudaHostAllocDefault));
        p.cap = bytes;
    }
    if (!p.ev) GPB_CUDA(cudaEventCreateWithFlags(&p.ev, cudaEventDisableTiming));
    *slot = sidx;
    *host = p.host;
    return GPB_OK;
}
static int pin_release(int slot, cudaStream_t st) {
    GPB_CUDA(cudaEventRecord(g_pin[slot].ev, st));
    g_pin[slot].used = true;
    return GPB_OK;
}

// one persistent, grow-only page-locked buffer for the synchronous host-buffer posterior calls
// (the ring above re-allocates a slot whenever a request outgrows it: ~1 ms per MB)
static void* g_post_pin = nullptr;
static size_t g_post_pin_cap = 0;
static int post_pin_reserve(size_t bytes, void** host) {
    if (bytes > g_post_pin_cap) {
        if (g_post_pin) GPB_CUDA(cudaFreeHost(g_post_pin));
        g_post_pin = nullptr; g_post_pin_cap = 0;
        size_t cap = (size_t)1 << 16;
        while (cap < bytes) cap <<= 1;
        GPB_CUDA(cudaHostAlloc(&g_post_pin, cap, cudaHostAllocDefault));
        g_post_pin_cap = cap;
    }
    *host = g_post_pin;
    return GPB_OK;
}

// ---- internal streams for concurrent candidate groups ---------------------------------
#define GPB_MAX_GROUPS 8
static cudaStream_t g_gstream[GPB_MAX_GROUPS];
static cudaEvent_t g_fork_ev, g_join_ev[GPB_MAX_GROUPS];
static bool g_gstream_init = false;
static int gstreams_init() {
    if (g_gstream_init) return GPB_OK;
    for (int i = 0; i < GPB_MAX_GROUPS; i++) {
        GPB_CUDA(cudaStreamCreateWithFlags(&g_gstream[i], cudaStreamNonBlocking));
        GPB_CUDA(cudaEventCreateWithFlags(&g_join_ev[i], cudaEventDisableTiming));
    }
    GPB_CUDA(cudaEventCreateWithFlags(&g_fork_ev, cudaEventDisableTiming));
    g_gstream_init = true;
    return GPB_OK;
}
static int eval_groups(int64_t n, int batch) {
    int g = gpb_get_option("eval_streams");
    if (g < 1) g = 4;
    if (g > GPB_MAX_GROUPS) g = GPB_MAX_GROUPS;
    if (n < 512) g = 1;                  // tiny problems: one launch group is already latency-bound
    if (g > batch) g = batch;
    return g;
}

static inline cudaStream_t S(void* s) { return reinterpret_cast<cudaStream_t>(s); }
static inline long long roundup(long long n, long long m) { return (n + m - 1) / m * m; }

// ---------------------------------------------------------------------------
// eval finalisation: assemble {log_lh, dloglh[...], logdet, quad, info} per candidate
// ---------------------------------------------------------------------------
__global__ void eval_finalize_kernel(const double* out3, const double* out8, const int* info,
                                     const KParams* Pb, int kind, int want_grad, int batch,
                                     double* result) {
    const int b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= batch) return;
    const int np = kind == GPB_GAUSSIAN ? 2 : 3;
    double* r = result + (long long)b * 8;
    const bool bad = info[b] != 0;
    r[0] = out3[b * 3 + 0];
    for (int i = 1; i <= 4; i++) r[i] = 0.0;
    if (want_grad) {
        const double* q = out8 + (long long)b * 16;
        const double s = Pb[b].s;
        for (int i = 0; i < np; i++)      // gp_c.pyx:47-49: 0.5*y^T(Ki dK)Kiy - 0.5*tr(Ki dK)
            r[1 + i] = bad ? NAN : 0.5 * q[i] + -0.5 * q[6 + i];
        // noise row: dK = 2 s I  (gp_c.pyx:45)
        r[1 + np] = bad ? NAN : 0.5 * (2.0 * s * q[13]) + -0.5 * (2.0 * s * q[12]);
    }
    r[5] = out3[b * 3 + 1];
    r[6] = out3[b * 3 + 2];
    r[7] = (double)info[b];
}

#define GPB_STAGE_PACK 24      // doubles read back by gpb_gp_stages: out3 | out16 | info | pad
struct EvalWs {
    double *L, *W, *V, *Ki, *z, *alpha, *ypad, *partial, *out3, *out8, *pack;
    KParams* Pb;
    int *info, *flags;
    size_t bytes;
};

static EvalWs carve(char* base, long long n, int batch, int want_grad) {
    const long long np_ = roundup(n, GPB_NB), T = np_ / GPB_NB;
    const size_t mat = (size_t)batch * np_ * np_ * 8;
    auto al = [](size_t v) { return (v + 255) / 256 * 256; };
    size_t off = 0;
    EvalWs w;
    auto take = [&](size_t bytes) { char* p = base ? base + off : nullptr; off += al(bytes); return p; };
    w.L = (double*)take(mat);
    w.W = (double*)take(mat);
    w.V = (double*)take(want_grad ? mat : 0);
    w.Ki = (double*)take(want_grad ? mat : 0);
    w.z = (double*)take((size_t)batch * np_ * 8);
    w.alpha = (double*)take((size_t)batch * np_ * 8);
    w.ypad = (double*)take((size_t)np_ * 8);
    w.partial = (double*)take((size_t)batch * gpb_grad_reduce_blocks(n) * 16 * 8);
    w.out3 = (double*)take((size_t)batch * 3 * 8);
    w.out8 = (double*)take((size_t)batch * 16 * 8);
    w.Pb = (KParams*)take((size_t)batch * sizeof(KParams));
    w.info = (int*)take((size_t)batch * 4);
    w.flags = (int*)take(((size_t)2 * batch * T + 2) * 4);
    w.pack = (double*)take(GPB_STAGE_PACK * 8);
    w.bytes = off;
    return w;
}

// host-API device memory pool (grow-only)
static void* g_pool = nullptr;
static size_t g_pool_bytes = 0;
static int pool_reserve(size_t bytes) {
    if (bytes <= g_pool_bytes) return GPB_OK;
    if (g_pool) GPB_CUDA(cudaFree(g_pool));
    g_pool = nullptr;
    g_pool_bytes = 0;
    GPB_CUDA(cudaMalloc(&g_pool, bytes));
    g_pool_bytes = bytes;
    return GPB_OK;
}

// ---- device -> host downloads ---------------------------------------------------------
// A pageable destination is filled through three pinned staging slots: the DMA of chunk
// k+1 runs while the host copies chunk k out of its slot, so the PCIe link never waits for
// the driver's own (slower, serialised) pageable path.
// host-side copy out of a pinned slot, split over a few persistent helper threads: one core moves
// ~10 GB/s, the PCIe link delivers 25-50
#include <atomic>
#include <condition_variable>
#include <mutex>
#include <thread>
namespace {
struct CopyPool {
    static constexpr int NT = 3;                 // helpers (the caller copies a share too)
    std::thread th[NT];
    std::mutex mu;
    std::condition_variable cv, done_cv;
    struct Task { char* d; const char* s; size_t n; } task[NT];
    unsigned gen = 0;
    int pending = 0;
    bool started = false, stop = false;
    void worker(int id) {
        unsigned seen = 0;
        for (;;) {
            Task t;
            {
                std::unique_lock<std::mutex> lk(mu);
                cv.wait(lk, [&] { return stop || gen != seen; });
                if (stop) return;
                seen = gen;
                t = task[id];
            }
            if (t.n) memcpy(t.d, t.s, t.n);
            {
                std::lock_guard<std::mutex> lk(mu);
                if (--pending == 0) done_cv.notify_one();
            }
        }
    }
    void copy(char* d, const char* s, size_t n) {
        if (n < ((size_t)1 << 20)) { memcpy(d, s, n); return; }
        if (!started) {
            started = true;
            for (int i = 0; i < NT; i++) th[i] = std::thread(&CopyPool::worker, this, i);
        }
        const size_t share = (n / (NT + 1)) & ~(size_t)4095;
        {
            std::lock_guard<std::mutex> lk(mu);
            for (int i = 0; i < NT; i++) task[i] = {d + (size_t)(i + 1) * share, s + (size_t)(i + 1) * share,
                                                    (i == NT - 1) ? n - (size_t)(NT) * share : share};
            pending = NT;
            gen++;
        }
        cv.notify_all();
        memcpy(d, s, share);
        std::unique_lock<std::mutex> lk(mu);
        done_cv.wait(lk, [&] { return pending == 0; });
    }
    ~CopyPool() {
        if (!started) return;
        { std::lock_guard<std::mutex> lk(mu); stop = true; }
        cv.notify_all();
        for (auto& t : th) if (t.joinable()) t.join();
    }
};
CopyPool g_copy_pool;
}  // namespace

#define GPB_STAGE_SLOTS 3
#define GPB_STAGE_BYTES ((size_t)8 << 20)
static void* g_stage[GPB_STAGE_SLOTS];
static cudaEvent_t g_stage_ev[GPB_STAGE_SLOTS];
static int stage_init() {
    if (g_stage[0]) return GPB_OK;
    for (int i = 0; i < GPB_STAGE_SLOTS; i++) {
        GPB_CUDA(cudaHostAlloc(&g_stage[i], GPB_STAGE_BYTES, cudaHostAllocDefault));
        GPB_CUDA(cudaEventCreateWithFlags(&g_stage_ev[i], cudaEventDisableTiming));
    }
    return GPB_OK;
}
static int d2h_staged(double* dst, long long ldd, const double* src, long long lds, long long rows,
                      long long cols, cudaStream_t st) {
    if (rows <= 0 || cols <= 0) return GPB_OK;
    int stt = stage_init();
    if (stt) return stt;
    const long long slot_doubles = (long long)(GPB_STAGE_BYTES / 8);
    if (ldd == cols && lds == cols && rows > 1 && (cols > slot_doubles || cols < 4096)) {
        // contiguous on both sides: re-tile as rows of one slot quarter (plus a remainder row)
        const long long total = rows * cols, w = slot_doubles / 4, full = total / w;
        if (full > 0) { stt = d2h_staged(dst, w, src, w, full, w, st); if (stt) return stt; }
        if (total > full * w) return d2h_staged(dst + full * w, total - full * w, src + full * w, total - full * w, 1, total - full * w, st);
        return GPB_OK;
    }
    if (cols > slot_doubles) {           // a single row longer than a slot: split by columns
        for (long long c0 = 0; c0 < cols; c0 += slot_doubles) {
            const long long cw = (cols - c0 < slot_doubles) ? cols - c0 : slot_doubles;
            stt = d2h_staged(dst + c0, ldd, src + c0, lds, rows, cw, st);
            if (stt) return stt;
        }
        return GPB_OK;
    }
    const long long rpc = slot_doubles / cols;     // rows per chunk (>= 1)
    const long long nchunks = (rows + rpc - 1) / rpc;
    auto drain = [&](long long c) -> int {
        const int slot = (int)(c % GPB_STAGE_SLOTS);
        GPB_CUDA(cudaEventSynchronize(g_stage_ev[slot]));
        const long long r0 = c * rpc, nr = (rows - r0 < rpc) ? rows - r0 : rpc;
        const double* h = (const double*)g_stage[slot];
        if (ldd == cols) g_copy_pool.copy((char*)(dst + r0 * ldd), (const char*)h, (size_t)nr * cols * 8);
        else for (long long r = 0; r < nr; r++) memcpy(dst + (r0 + r) * ldd, h + r * cols, (size_t)cols * 8);
        return GPB_OK;
    };
    for (long long c = 0; c < nchunks; c++) {
        if (c >= GPB_STAGE_SLOTS) { stt = drain(c - GPB_STAGE_SLOTS); if (stt) return stt; }
        const int slot = (int)(c % GPB_STAGE_SLOTS);
        const long long r0 = c * rpc, nr = (rows - r0 < rpc) ? rows - r0 : rpc;
        GPB_CUDA(cudaMemcpy2DAsync(g_stage[slot], (size_t)cols * 8, src + r0 * lds, (size_t)lds * 8,
                                   (size_t)cols * 8, (size_t)nr, cudaMemcpyDeviceToHost, st));
        GPB_CUDA(cudaEventRecord(g_stage_ev[slot], st));
    }
    for (long long c = (nchunks > GPB_STAGE_SLOTS ? nchunks - GPB_STAGE_SLOTS : 0); c < nchunks; c++) {
        stt = drain(c);
        if (stt) return stt;
    }
    return GPB_OK;
}

extern "C" {

int gpb_version(void) { return 100; }
const char* gpb_last_error(void) { return g_err; }
double gpb_min_log(void) { return GPB_MIN_LOG; }
int64_t gpb_launch_count(void) { return g_gpb_launches; }

void gpb_profile_enable(int on) {
    GPB_API_LOCK;
    prof_collect();
    g_gpb_profile = on;
    if (on) for (int c = 0; c < GPB_KC_COUNT; c++) { g_prof_ms[c] = 0.0; g_prof_n[c] = 0; }
}
int gpb_set_option(const char* name, int value) {
    GPB_API_LOCK;
    for (auto& o : g_options)
        if (strcmp(o.name, name) == 0) { o.value = value; o.env_read = true; return GPB_OK; }
    gpb_set_error("gpb_set_option: unknown option %s", name);
    return GPB_ERR_ARG;
}
int gpb_profile_read(int cls, double* ms, int64_t* launches) {
    GPB_API_LOCK;
    GPB_REQUIRE(cls >= 0 && cls < GPB_KC_COUNT, "bad kernel class");
    prof_collect();
    *ms = g_prof_ms[cls];
    *launches = g_prof_n[cls];
    return GPB_OK;
}

int gpb_kernel_build(int kind, const double* theta, double s, const double* x1, int64_t n1,
                     const double* x2, int64_t n2, int64_t rows, int64_t cols, unsigned slice_mask,
                     double* out, int64_t ld, int64_t slice_stride, int add_diag, int pad_identity,
                     void* stream) {
    GPB_REQUIRE(kind == GPB_GAUSSIAN || kind == GPB_PERIODIC, "unknown kernel kind");
    GPB_REQUIRE(theta && out, "null pointer");
    KParams P;
    gpb_make_kparams(&P, kind, theta, s);
    double* outs[GPB_MAX_SLICES];
    int q = 0;
    for (int sidx = 0; sidx < GPB_MAX_SLICES; sidx++) {
        if (sidx < gpb_n_slices(kind) && (slice_mask >> sidx) & 1u) outs[sidx] = out + (long long)(q++) * slice_stride;
        else outs[sidx] = nullptr;
    }
    return gpb_launch_build(kind, &P, nullptr, 1, x1, n1, x2, n2, rows, cols, outs, ld, 0, add_diag,
                            pad_identity, S(stream));
}

int gpb_kernel_matvec(int kind, const double* theta, const double* x1, int64_t n1, const double* x2,
                      int64_t n2, int npairs, const int* slice, const int* outidx, const double* coef,
                      const double* const* vec, int nout, double* const* out, void* stream) {
    GPB_REQUIRE(kind == GPB_GAUSSIAN || kind == GPB_PERIODIC, "unknown kernel kind");
    KParams P;
    gpb_make_kparams(&P, kind, theta, 0.0);
    return gpb_launch_fused_matvec(kind, &P, nullptr, 1, x1, n1, x2, n2, npairs, slice, outidx, coef, vec,
                                   nout, out, 0, 0, S(stream));
}

int gpb_post_var(int kind, const double* theta, const double* Z, int64_t ldz, int64_t m, int64_t n,
                 double* out, void* stream) {
    GPB_REQUIRE(kind == GPB_GAUSSIAN || kind == GPB_PERIODIC, "unknown kernel kind");
    GPB_REQUIRE(theta, "null pointer");
    KParams P;
    gpb_make_kparams(&P, kind, theta, 0.0);
    return gpb_launch_post_var(kind, &P, Z, ldz, m, n, out, S(stream));
}

int gpb_potrf(double* A, int64_t n, int64_t ld, int64_t stride_a, int batch, double* W, int64_t ldw,
              int64_t stride_w, double* V, int64_t ldv, int64_t stride_v, int* info, void* stream) {
    return gpb_launch_potrf(A, n, ld, stride_a, batch, W, ldw, stride_w, V, ldv, stride_v, info, S(stream), 0, true,
                            batch == 1);
}

int gpb_potrs(const double* L, const double* W, int64_t n, int64_t ld, int64_t ldw, int64_t stride_l,
              int64_t stride_w, int batch, const double* y, int64_t stride_y, double* z, double* alpha,
              int64_t stride_vec, int* flags, void* stream) {
    return gpb_launch_potrs(L, W, n, ld, ldw, stride_l, stride_w, batch, y, stride_y, z, alpha, stride_vec,
                            flags, S(stream));
}

int gpb_trtri(const double* L, int64_t n, int64_t ld, int64_t stride_l, int batch, double* W,
              int64_t ldw, int64_t stride_w, double* V, int64_t ldv, int64_t stride_v, double* T,
              int64_t ldt, int64_t stride_t, void* stream) {
    return gpb_launch_trtri(L, n, ld, stride_l, batch, W, ldw, stride_w, V, ldv, stride_v, T, ldt,
                            stride_t, S(stream));
}

int gpb_lauum(const double* V, int64_t n, int64_t ldv, int64_t stride_v, int batch, double* Ki,
              int64_t ldk, int64_t stride_k, void* stream) {
    return gpb_launch_lauum(V, n, ldv, stride_v, batch, Ki, ldk, stride_k, S(stream));
}

int gpb_tril(double* A, int64_t n, int64_t ld, void* stream) {
    return gpb_launch_tril(A, n, ld, 0, 1, S(stream));
}

int gpb_tril_copy(double* dst, int64_t ldd, const double* src, int64_t lds, int64_t n, void* stream) {
    GPB_REQUIRE(dst && src && n >= 0 && ldd >= n && lds >= n, "bad argument");
    return gpb_launch_tril_copy(dst, ldd, src, lds, n, S(stream));
}

int gpb_copy2d(double* dst, int64_t ldd, const double* src, int64_t lds, int64_t rows, int64_t cols,
               void* stream) {
    return gpb_launch_copy2d(dst, ldd, src, lds, rows, cols, 0, 0, 1, S(stream));
}

int gpb_download_2d(double* dst_host, int64_t ld_host, const double* src_dev, int64_t ld_dev,
                    int64_t rows, int64_t cols, int dst_pinned, void* stream) {
    GPB_API_LOCK;
    GPB_REQUIRE(rows >= 0 && cols >= 0 && ld_host >= cols && ld_dev >= cols, "bad extents");
    if (rows == 0 || cols == 0) return GPB_OK;
    GPB_REQUIRE(dst_host && src_dev, "null pointer");
    if (dst_pinned) {          // page-locked destination: one asynchronous strided DMA
        GPB_CUDA(cudaMemcpy2DAsync(dst_host, (size_t)ld_host * 8, src_dev, (size_t)ld_dev * 8, (size_t)cols * 8,
                                   (size_t)rows, cudaMemcpyDeviceToHost, S(stream)));
        return GPB_OK;
    }
    return d2h_staged(dst_host, ld_host, src_dev, ld_dev, rows, cols, S(stream));
}

int gpb_gemm_nt(const double* A, int64_t lda, const double* B, int64_t ldb, double* C, int64_t ldc,
                double* Ct, int64_t ldct, int64_t M, int64_t N, int64_t K, double alpha, double beta,
                int a_tri, int b_tri, int lower_only, void* stream) {
    GpbGemm g = gpb_gemm_default();
    g.A = A; g.lda = lda; g.B = B; g.ldb = ldb; g.C = C; g.ldc = ldc; g.Ct = Ct; g.ldct = ldct;
    g.M = (int)M; g.N = (int)N; g.K = (int)K; g.alpha = alpha; g.beta = beta;
    g.a_tri = a_tri; g.b_tri = b_tri; g.lower_only = lower_only;
    return gpb_launch_gemm(g, 1, S(stream));
}

int gpb_loglh(const double* L, int64_t n_valid, int64_t ld, const double* y, const double* alpha,
              const int* info, double* out3, void* stream) {
    return gpb_launch_loglh(L, n_valid, ld, 0, 1, y, 0, alpha, 0, info, out3, S(stream));
}

int64_t gpb_grad_partial_doubles(int64_t n) { return (int64_t)gpb_grad_reduce_blocks(n) * 16; }

int gpb_slice_reduce(int kind, const double* theta, const double* x, int64_t n, const double* Ki,
                     int64_t ldk, const double* alpha, int nslices, const int* slices,
                     double* partial, double* out16, void* stream) {
    GPB_REQUIRE(kind == GPB_GAUSSIAN || kind == GPB_PERIODIC, "unknown kernel kind");
    KParams P;
    gpb_make_kparams(&P, kind, theta, 0.0);
    return gpb_launch_grad_reduce(kind, &P, nullptr, 1, x, n, Ki, ldk, 0, alpha, 0, nslices, slices,
                                  partial, out16, S(stream));
}

int gpb_gemv(const double* A, int64_t rows, int64_t cols, int64_t lda, const double* x, double* y,
             double alpha, double beta, void* stream) {
    return gpb_launch_gemv(A, rows, cols, lda, x, y, alpha, beta, S(stream));
}

int gpb_trace_prod(const double* A, int64_t lda, const double* B, int64_t ldb, int64_t n,
                   double* partial, double* out, void* stream) {
    return gpb_launch_trace_prod(A, lda, B, ldb, n, partial, out, S(stream));
}

int gpb_quadform(const double* u, const double* M, int64_t ldm, const double* v, int64_t n,
                 double* partial, double* out, void* stream) {
    return gpb_launch_quadform(u, M, ldm, v, n, partial, out, S(stream));
}

size_t gpb_eval_workspace_bytes(int64_t n, int batch, int want_grad) {
    const int G = eval_groups(n, batch);
    const int gb = (batch + G - 1) / G;
    return (size_t)G * carve(nullptr, n, gb, want_grad).bytes;
}

// one group of candidates on one stream
static int eval_group(int kind, const double* thetas, int batch, const double* x, const double* y, int64_t n,
                      int want_grad, char* wsbase, double* result, cudaStream_t st) {
    EvalWs w = carve(wsbase, n, batch, want_grad);
    const long long np_ = roundup(n, GPB_NB);
    const long long mstride = np_ * np_;
    const int nth = gpb_n_kparams(kind) + 1;

    int slot;
    void* hostp;
    int stt = pin_acquire(sizeof(KParams) * batch, &slot, &hostp);
    if (stt) return stt;
    KParams* hp = (KParams*)hostp;
    for (int b = 0; b < batch; b++) gpb_make_kparams(&hp[b], kind, thetas + (long long)b * nth, thetas[(long long)b * nth + nth - 1]);
    GPB_CUDA(cudaMemcpyAsync(w.Pb, hp, sizeof(KParams) * batch, cudaMemcpyHostToDevice, st));
    stt = pin_release(slot, st);
    if (stt) return stt;
    GPB_CUDA(cudaMemsetAsync(w.ypad, 0, np_ * 8, st));
    GPB_CUDA(cudaMemcpyAsync(w.ypad, y, n * 8, cudaMemcpyDeviceToDevice, st));

    // Kxx + s^2 I straight into the factorisation buffer, identity in the pad
    double* outs[GPB_MAX_SLICES] = {nullptr};
    outs[0] = w.L;
    // the factorisation reads the lower triangle only: skip the tiles above it (half the exp work)
    stt = gpb_launch_build(kind, nullptr, w.Pb, batch, x, n, x, n, np_, np_, outs, np_, mstride, 1, 1, st, 1);
    if (stt) return stt;
    const bool one_block = (np_ == GPB_NB);
    stt = gpb_launch_potrf(w.L, np_, np_, mstride, batch, w.W, np_, mstride, want_grad ? w.V : nullptr, np_, mstride, w.info, st, n,
                           !one_block);
    if (stt) return stt;
    if (one_block) {
        // one 128-block per candidate: solves, log_lh, K^-1 and the gradient brackets in ONE launch
        stt = gpb_launch_small_tail(kind, nullptr, w.Pb, batch, x, n, w.ypad, 0, w.L, np_, mstride, w.W, np_, mstride,
                                    want_grad ? w.Ki : nullptr, np_, mstride, w.z, w.alpha, np_, w.info, w.out3,
                                    w.out8, w.W, want_grad ? w.V : nullptr, np_, mstride, nullptr, st);
        if (stt) return stt;
    } else {
        stt = gpb_launch_potrs(w.L, w.W, np_, np_, np_, mstride, mstride, batch, w.ypad, 0, w.z, w.alpha, np_, w.flags, st);
        if (stt) return stt;
        stt = gpb_launch_loglh(w.L, n, np_, mstride, batch, w.ypad, 0, w.alpha, np_, w.info, w.out3, st);
        if (stt) return stt;
        if (want_grad) {
            stt = gpb_launch_trtri(w.L, np_, np_, mstride, batch, w.W, np_, mstride, w.V, np_, mstride, w.Ki, np_, mstride, st);
            if (stt) return stt;
            if (gpb_lauum_grad_available()) {
                // K^-1 = V V^T never reaches memory: the gradient brackets are taken from its tiles in the GEMM's
                // epilogue (nothing else reads K^-1 in the batched evaluator)
                GpbLauumFuse f;
                f.kind = kind; f.store = 0; f.n = n; memset(&f.P, 0, sizeof(KParams)); f.Pb = w.Pb; f.x = x;
                f.alpha = w.alpha; f.astride = np_; f.partial = w.partial;
                stt = gpb_launch_lauum_grad(w.V, np_, np_, mstride, batch, w.Ki, np_, mstride, f, w.out8, st);
                if (stt) return stt;
            } else {
                stt = gpb_launch_lauum(w.V, np_, np_, mstride, batch, w.Ki, np_, mstride, st);
                if (stt) return stt;
                const int jsl[3] = {1, 2, 3};
                stt = gpb_launch_grad_reduce(kind, nullptr, w.Pb, batch, x, n, w.Ki, np_, mstride, w.alpha, np_,
                                             gpb_n_kparams(kind), jsl, w.partial, w.out8, st);
                if (stt) return stt;
            }
        }
    }
    eval_finalize_kernel<<<(batch + 127) / 128, 128, 0, st>>>(w.out3, w.out8, w.info, w.Pb, kind, want_grad, batch, result);
    GPB_LAUNCH_CHECK("eval_finalize_kernel");
    return GPB_OK;
}

int gpb_gp_eval(int kind, const double* thetas, int batch, const double* x, const double* y, int64_t n,
                int want_grad, void* workspace, size_t workspace_bytes, double* result, void* stream) {
    GPB_API_LOCK;
    GPB_REQUIRE(kind == GPB_GAUSSIAN || kind == GPB_PERIODIC, "unknown kernel kind");
    GPB_REQUIRE(batch >= 1 && n >= 1 && thetas && x && y && workspace && result, "bad argument");
    GPB_REQUIRE(gpb_eval_workspace_bytes(n, batch, want_grad) <= workspace_bytes, "workspace too small (see gpb_eval_workspace_bytes)");
    GPB_REQUIRE((uintptr_t)workspace % 256 == 0, "workspace must be 256-byte aligned");
    cudaStream_t st = S(stream);
    const int nth = gpb_n_kparams(kind) + 1;
    const int G = eval_groups(n, batch);
    if (G == 1) return eval_group(kind, thetas, batch, x, y, n, want_grad, (char*)workspace, result, st);
    // Candidates are independent: groups run on concurrent streams so that one group's serial
    // diagonal-block / panel kernels overlap another group's trailing-update GEMMs.
    int stt = gstreams_init();
    if (stt) return stt;
    const int gb = (batch + G - 1) / G;
    const size_t gbytes = carve(nullptr, n, gb, want_grad).bytes;
    GPB_CUDA(cudaEventRecord(g_fork_ev, st));
    for (int g = 0; g < G; g++) {
        const int b0 = g * gb, b1 = (b0 + gb < batch) ? b0 + gb : batch;
        if (b1 <= b0) break;
        GPB_CUDA(cudaStreamWaitEvent(g_gstream[g], g_fork_ev, 0));
        stt = eval_group(kind, thetas + (long long)b0 * nth, b1 - b0, x, y, n, want_grad,
                         (char*)workspace + (size_t)g * gbytes, result + (long long)b0 * 8, g_gstream[g]);
        if (stt) return stt;
        GPB_CUDA(cudaEventRecord(g_join_ev[g], g_gstream[g]));
        GPB_CUDA(cudaStreamWaitEvent(st, g_join_ev[g], 0));
    }
    return GPB_OK;
}

int gpb_gp_eval_host(int kind, const double* thetas, int batch, const double* x, const double* y,
                     int64_t n, int want_grad, double* result) {
    GPB_API_LOCK;
    GPB_REQUIRE(batch >= 1 && n >= 1 && thetas && x && y && result, "bad argument");
    const size_t wsb = gpb_eval_workspace_bytes(n, batch, want_grad);
    const size_t extra = (size_t)(2 * n + 8 * (size_t)batch) * 8 + 1024;
    int stt = pool_reserve(wsb + extra);
    if (stt) return stt;
    char* base = (char*)g_pool;
    double* dx = (double*)(base + wsb);
    double* dy = dx + n;
    double* dres = dy + n;
    GPB_CUDA(cudaMemcpyAsync(dx, x, n * 8, cudaMemcpyHostToDevice, 0));
    GPB_CUDA(cudaMemcpyAsync(dy, y, n * 8, cudaMemcpyHostToDevice, 0));
    stt = gpb_gp_eval(kind, thetas, batch, dx, dy, n, want_grad, base, wsb, dres, nullptr);
    if (stt) return stt;
    GPB_CUDA(cudaMemcpyAsync(result, dres, (size_t)batch * 8 * 8, cudaMemcpyDeviceToHost, 0));
    GPB_CUDA(cudaStreamSynchronize(0));
    return GPB_OK;
}

// ---- one GP, caller-resident buffers, staged: what GP.log_lh / dloglh_dtheta / cov need ----
// The property chain of a single GP object used to be ~12 library calls and 3 blocking read-backs
// from Python; at the reference's own test-suite scale (N = 50) that host latency IS the cost.
// Here the whole chain is enqueued by one call and everything the host needs comes back in one
// 192-byte transfer.
__global__ void stage_pack_kernel(const double* out3, const double* out16, const int* info, double* pack) {
    const int t = threadIdx.x;
    if (t < 3) pack[t] = out3[t];
    else if (t < 19) pack[t] = out16 ? out16[t - 3] : 0.0;      // null until the gradient stage has run
    else if (t == 19) pack[t] = (double)info[0];
    else if (t < GPB_STAGE_PACK) pack[t] = 0.0;
}

int gpb_eval_layout(int64_t n, int64_t* off, int noff) {
    GPB_REQUIRE(n >= 1 && off && noff >= 15, "bad argument");
    char* base = reinterpret_cast<char*>((uintptr_t)1 << 20);
    EvalWs w = carve(base, n, 1, 1);
    const void* p[14] = {w.L, w.W, w.V, w.Ki, w.z, w.alpha, w.ypad, w.partial, w.out3, w.out8, w.Pb, w.info, w.flags, w.pack};
    for (int i = 0; i < 14; i++) off[i] = (int64_t)((const char*)p[i] - base);
    off[14] = (int64_t)w.bytes;
    return GPB_OK;
}

int gpb_gp_stages(int kind, const double* theta, const double* x, const double* ypad, int64_t n,
                  unsigned stages, void* workspace, size_t workspace_bytes, double* host_out, void* stream) {
    GPB_API_LOCK;
    GPB_REQUIRE(kind == GPB_GAUSSIAN || kind == GPB_PERIODIC, "unknown kernel kind");
    GPB_REQUIRE(n >= 1 && theta && x && ypad && workspace, "bad argument");
    GPB_REQUIRE((uintptr_t)workspace % 256 == 0, "workspace must be 256-byte aligned");
    EvalWs w = carve((char*)workspace, n, 1, 1);
    GPB_REQUIRE(w.bytes <= workspace_bytes, "workspace too small (see gpb_eval_layout)");
    cudaStream_t st = S(stream);
    const long long np_ = roundup(n, GPB_NB);
    const int nkp = gpb_n_kparams(kind);
    KParams P;
    gpb_make_kparams(&P, kind, theta, theta[nkp]);
    int stt;
    unsigned done = 0;
    bool packed = false, side = false;
    // the read-back block is written by the last kernel straight into page-locked host memory
    // (device-addressable under unified addressing): no copy-engine round trip after the chain
    double* pack_dst = w.pack;
    if (host_out) {
        void* hp;
        stt = post_pin_reserve(GPB_STAGE_PACK * 8, &hp);
        if (stt) return stt;
        pack_dst = (double*)hp;
    }
    if (np_ == GPB_NB) {
        // a GP that fits one 128-block: build, factor + invert (one CTA), then everything else in
        // one more launch -- all four stages at once, whatever subset was asked for
        if (stages & 1u) {
            double* outs[GPB_MAX_SLICES] = {nullptr};
            outs[0] = w.L;
            stt = gpb_launch_build(kind, &P, nullptr, 1, x, n, x, n, np_, np_, outs, np_, 0, 1, 1, st, 1);
            if (stt) return stt;
            stt = gpb_launch_potrf(w.L, np_, np_, 0, 1, w.W, np_, 0, w.V, np_, 0, w.info, st, n, false);
            if (stt) return stt;
            stt = gpb_launch_small_tail(kind, &P, nullptr, 1, x, n, ypad, 0, w.L, np_, 0, w.W, np_, 0, w.Ki, np_, 0,
                                        w.z, w.alpha, np_, w.info, w.out3, w.out8, w.W, w.V, np_, 0, pack_dst, st);
            if (stt) return stt;
            done = 15u;
            packed = true;
        }
        stages = 0;
    }
    if (stages & 1u) {          // Kxx + s^2 I (lower tiles) -> L, W_kk, V_kk, info; alpha; log_lh
        double* outs[GPB_MAX_SLICES] = {nullptr};
        outs[0] = w.L;
        stt = gpb_launch_build(kind, &P, nullptr, 1, x, n, x, n, np_, np_, outs, np_, 0, 1, 1, st, 1);
        if (stt) return stt;
        stt = gpb_launch_potrf(w.L, np_, np_, 0, 1, w.W, np_, 0, w.V, np_, 0, w.info, st, n, true, true);
        if (stt) return stt;
        // The two triangular solves are latency-bound chains of 32 CTAs (2 x 119 us at N = 4096) and touch only L, the
        // diagonal blocks of W, z and alpha: when the inverse is asked for in the same call they run on a side
        // stream under the trtri GEMMs and join before the gradient reduction / the read-back.
        cudaStream_t ss = st;
        if ((stages & 2u) && np_ >= 1024 && gpb_get_option("stage_overlap") != 2) {
            stt = gstreams_init();
            if (stt) return stt;
            ss = g_gstream[0];
            GPB_CUDA(cudaEventRecord(g_fork_ev, st));
            GPB_CUDA(cudaStreamWaitEvent(ss, g_fork_ev, 0));
            side = true;
        }
        stt = gpb_launch_potrs(w.L, w.W, np_, np_, np_, 0, 0, 1, ypad, 0, w.z, w.alpha, np_, w.flags, ss);
        if (stt) return stt;
        stt = gpb_launch_loglh(w.L, n, np_, 0, 1, ypad, 0, w.alpha, np_, w.info, w.out3, ss);
        if (stt) return stt;
        if (side) GPB_CUDA(cudaEventRecord(g_join_ev[0], ss));
    }
    if (stages & 2u) {          // W = L^-1, V = L^-T
        stt = gpb_launch_trtri(w.L, np_, np_, 0, 1, w.W, np_, 0, w.V, np_, 0, w.Ki, np_, 0, st);
        if (stt) return stt;
    }
    bool grad_fused = false;
    if (stages & 4u) {          // Ki = V V^T  (with the gradient brackets in its epilogue when both are asked for)
        if ((stages & 8u) && gpb_lauum_grad_available()) {
            if (side) {             // alpha of the side stream
                GPB_CUDA(cudaStreamWaitEvent(st, g_join_ev[0], 0));
                side = false;
            }
            GpbLauumFuse f;
            f.kind = kind; f.store = 1; f.n = n; f.P = P; f.Pb = nullptr; f.x = x;
            f.alpha = w.alpha; f.astride = np_; f.partial = w.partial;
            stt = gpb_launch_lauum_grad(w.V, np_, np_, 0, 1, w.Ki, np_, 0, f, w.out8, st);
            if (stt) return stt;
            grad_fused = true;
        } else {
            stt = gpb_launch_lauum(w.V, np_, np_, 0, 1, w.Ki, np_, 0, st);
            if (stt) return stt;
        }
    }
    if (side) {                 // alpha, out3 of the side stream
        GPB_CUDA(cudaStreamWaitEvent(st, g_join_ev[0], 0));
        side = false;
    }
    if ((stages & 8u) && !grad_fused) {          // a^T dK_i a, sum(Ki o dK_i), tr Ki, a.a
        const int jsl[3] = {1, 2, 3};
        stt = gpb_launch_grad_reduce(kind, &P, nullptr, 1, x, n, w.Ki, np_, 0, w.alpha, np_, nkp, jsl, w.partial,
                                     w.out8, st);
        if (stt) return stt;
    }
    done |= stages & 15u;
    if (host_out) {
        if (!packed) {
            stage_pack_kernel<<<1, 32, 0, st>>>(w.out3, (stages & 8u) ? w.out8 : nullptr, w.info, pack_dst);
            GPB_LAUNCH_CHECK("stage_pack_kernel");
        }
        GPB_CUDA(cudaStreamSynchronize(st));
        memcpy(host_out, pack_dst, GPB_STAGE_PACK * 8);
        host_out[20] = (double)done;       // stages this call completed (a one-block GP completes all four)
    }
    return GPB_OK;
}

// posterior mean with host buffers in and out: K(xo, x) alpha, the M x N kernel matrix never
// materialised (gp.py:574-597).  scratch: DEVICE, >= 2 * roundup(m, 32) doubles.
int gpb_post_mean_host(int kind, const double* theta, const double* xo_host, int64_t m, const double* x,
                       int64_t n, const double* alpha, double* scratch, double* out_host, void* stream) {
    GPB_API_LOCK;
    GPB_REQUIRE(kind == GPB_GAUSSIAN || kind == GPB_PERIODIC, "unknown kernel kind");
    GPB_REQUIRE(m >= 0 && n >= 1 && theta && x && alpha, "bad argument");
    if (m == 0) return GPB_OK;
    GPB_REQUIRE(xo_host && out_host && (scratch || m <= 2048), "null pointer");
    cudaStream_t st = S(stream);
    const long long mr = roundup(m, 32);
    KParams P;
    gpb_make_kparams(&P, kind, theta, 0.0);
    // small test sets go through a page-locked slot in both directions: a pageable cudaMemcpyAsync
    // costs ~12 us per direction in the driver, the whole kernel runs ~4 us
    const bool pinned = (size_t)m * 16 <= ((size_t)64 << 20);
    void* hp = nullptr;
    if (pinned) {
        int stt = post_pin_reserve((size_t)(m + mr) * 8, &hp);
        if (stt) return stt;
        memcpy(hp, xo_host, (size_t)m * 8);
    }
    if (pinned && m <= 2048) {
        // Few test points: no copy engine at all.  Page-locked host memory is device-addressable
        // (unified addressing), so the kernel reads xo and writes the means straight across PCIe --
        // one launch and one synchronise instead of two DMA round trips (~20 us) around a 4 us kernel.
        const int sl[1] = {0}, oi[1] = {0};
        const double cf[1] = {1.0};
        const double* vec[1] = {alpha};
        double* out[1] = {(double*)hp + mr};
        int stt = gpb_launch_fused_matvec(kind, &P, nullptr, 1, (const double*)hp, m, x, n, 1, sl, oi, cf, vec, 1, out, 0, 0, st);
        if (stt) return stt;
        GPB_CUDA(cudaStreamSynchronize(st));
        memcpy(out_host, (double*)hp + mr, (size_t)m * 8);
        return GPB_OK;
    }
    GPB_CUDA(cudaMemcpyAsync(scratch, pinned ? hp : xo_host, (size_t)m * 8, cudaMemcpyHostToDevice, st));
    const int sl[1] = {0}, oi[1] = {0};
    const double cf[1] = {1.0};
    const double* vec[1] = {alpha};
    double* out[1] = {scratch + mr};
    int stt = gpb_launch_fused_matvec(kind, &P, nullptr, 1, scratch, m, x, n, 1, sl, oi, cf, vec, 1, out, 0, 0, st);
    if (stt) return stt;
    double* dst = pinned ? (double*)hp + m : out_host;
    GPB_CUDA(cudaMemcpyAsync(dst, scratch + mr, (size_t)m * 8, cudaMemcpyDeviceToHost, st));
    GPB_CUDA(cudaStreamSynchronize(st));
    if (pinned) memcpy(out_host, dst, (size_t)m * 8);
    return GPB_OK;
}

// posterior covariance with host buffers in and out, for test sets whose M x M result is small
// enough that pipelining its download does not pay (gp.py:599-625):
//   cov = K(xo,xo) - Z Z^T,  Z = K(xo,x) L^-T  (W = L^-1 from gpb_trtri / gpb_gp_stages).
// scratch: DEVICE, >= gpb_post_cov_scratch_doubles(m, n) doubles, 256-byte aligned.
static bool post_cov_one_launch(int64_t m, int64_t n) {
    return roundup(n, GPB_NB) == GPB_NB && m >= 1 && m <= GPB_NB && gpb_small_cov_smem(m, n) <= (size_t)200 * 1024;
}
size_t gpb_post_cov_scratch_doubles(int64_t m, int64_t n) {
    if (post_cov_one_launch(m, n)) return 0;       // one-block GP, one block of test points: no device scratch
    const size_t mp = (size_t)roundup(m, GPB_NB), np_ = (size_t)roundup(n, GPB_NB);
    return mp + 2 * mp * np_ + mp * mp;
}
int gpb_post_cov_host(int kind, const double* theta, const double* xo_host, int64_t m, const double* x,
                      int64_t n, const double* W, int64_t ldw, double* scratch, double* out_host,
                      int64_t ld_out, void* stream) {
    GPB_API_LOCK;
    GPB_REQUIRE(kind == GPB_GAUSSIAN || kind == GPB_PERIODIC, "unknown kernel kind");
    GPB_REQUIRE(m >= 0 && n >= 1 && theta && x && W, "bad argument");
    if (m == 0) return GPB_OK;
    GPB_REQUIRE(xo_host && out_host && ld_out >= m && (scratch || post_cov_one_launch(m, n)), "bad argument");
    GPB_REQUIRE((uintptr_t)scratch % 256 == 0, "scratch must be 256-byte aligned");
    cudaStream_t st = S(stream);
    const long long mp = roundup(m, GPB_NB), np_ = roundup(n, GPB_NB);
    double* dxo = scratch;
    double* Kxox = dxo + mp;
    double* Z = Kxox + mp * np_;
    double* C = Z + mp * np_;
    KParams P;
    gpb_make_kparams(&P, kind, theta, 0.0);
    if (post_cov_one_launch(m, n)) {
        // one-block GP, one block of test points: the whole covariance in a single launch that reads
        // xo from and writes the m x m result to page-locked host memory directly (<= 128 KB)
        void* hp;
        int stt = post_pin_reserve((size_t)(GPB_NB + m * m) * 8, &hp);
        if (stt) return stt;
        double* hxo = (double*)hp;
        double* hC = hxo + GPB_NB;
        memcpy(hxo, xo_host, (size_t)m * 8);
        stt = gpb_launch_small_cov(kind, &P, hxo, m, x, n, W, ldw, hC, m, st);
        if (stt) return stt;
        GPB_CUDA(cudaStreamSynchronize(st));
        if (ld_out == m) memcpy(out_host, hC, (size_t)m * m * 8);
        else for (int64_t r = 0; r < m; r++) memcpy(out_host + r * ld_out, hC + r * m, (size_t)m * 8);
        return GPB_OK;
    }
    {
        void* hp;
        int stt = post_pin_reserve((size_t)m * 8, &hp);      // the call synchronises before it returns
        if (stt) return stt;
        memcpy(hp, xo_host, (size_t)m * 8);
        GPB_CUDA(cudaMemcpyAsync(dxo, hp, (size_t)m * 8, cudaMemcpyHostToDevice, st));
    }
    double* outs[GPB_MAX_SLICES] = {nullptr};
    outs[0] = Kxox;
    int stt = gpb_launch_build(kind, &P, nullptr, 1, dxo, m, x, n, mp, np_, outs, np_, 0, 0, 0, st);
    if (stt) return stt;
    GpbGemm g = gpb_gemm_default();                 // Z = Kxox W^T   (W lower)
    g.A = Kxox; g.lda = np_; g.B = W; g.ldb = ldw; g.C = Z; g.ldc = np_;
    g.M = (int)mp; g.N = (int)np_; g.K = (int)np_; g.b_tri = 1;
    stt = gpb_launch_gemm(g, 1, st);
    if (stt) return stt;
    outs[0] = C;
    stt = gpb_launch_build(kind, &P, nullptr, 1, dxo, m, dxo, m, mp, mp, outs, mp, 0, 0, 0, st);
    if (stt) return stt;
    GpbGemm u = gpb_gemm_default();                 // C -= Z Z^T (lower tiles, mirrored)
    u.A = Z; u.lda = np_; u.B = Z; u.ldb = np_; u.C = C; u.ldc = mp; u.Ct = C; u.ldct = mp;
    u.M = (int)mp; u.N = (int)mp; u.K = (int)np_; u.alpha = -1.0; u.beta = 1.0; u.lower_only = 1;
    stt = gpb_launch_gemm(u, 1, st);
    if (stt) return stt;
    stt = d2h_staged(out_host, ld_out, C, mp, m, m, st);
    if (stt) return stt;
    GPB_CUDA(cudaStreamSynchronize(st));
    return GPB_OK;
}

// ---- host-buffer drop-ins for the Cython signatures --------------------------------
int gpb_kernel_slices_host(int kind, unsigned slice_mask, double* out, const double* x1, int64_t n1,
                           const double* x2, int64_t n2, const double* theta) {
    GPB_API_LOCK;
    GPB_REQUIRE(kind == GPB_GAUSSIAN || kind == GPB_PERIODIC, "unknown kernel kind");
    GPB_REQUIRE(out && x1 && x2 && theta && n1 >= 0 && n2 >= 0, "bad argument");
    if (n1 == 0 || n2 == 0) return GPB_OK;
    int ns = 0;
    for (int s = 0; s < gpb_n_slices(kind); s++) ns += (slice_mask >> s) & 1u;
    GPB_REQUIRE(ns > 0, "empty slice mask");
    const size_t xb = (size_t)roundup((n1 + n2) * 8, 256);
    const size_t ob = (size_t)ns * n1 * n2 * 8;
    int stt = pool_reserve(xb + ob);
    if (stt) return stt;
    double* dx1 = (double*)g_pool;
    double* dx2 = dx1 + n1;
    double* dout = (double*)((char*)g_pool + xb);
    GPB_CUDA(cudaMemcpyAsync(dx1, x1, n1 * 8, cudaMemcpyHostToDevice, 0));
    GPB_CUDA(cudaMemcpyAsync(dx2, x2, n2 * 8, cudaMemcpyHostToDevice, 0));
    stt = gpb_kernel_build(kind, theta, 0.0, dx1, n1, dx2, n2, n1, n2, slice_mask, dout, n2, n1 * n2, 0, 0, nullptr);
    if (stt) return stt;
    // ns slices of n1 x n2 are contiguous on both sides: one [ns*n1, n2] strided download
    stt = d2h_staged(out, n2, dout, n2, (long long)ns * n1, n2, 0);
    if (stt) return stt;
    GPB_CUDA(cudaStreamSynchronize(0));
    return GPB_OK;
}

#define GPB_G(name, mask)                                                                              \
    int gpb_gaussian_##name(double* out, const double* x1, int64_t n1, const double* x2, int64_t n2,  \
                            double h, double w) {                                                      \
        const double th[2] = {h, w};                                                                   \
        return gpb_kernel_slices_host(GPB_GAUSSIAN, mask, out, x1, n1, x2, n2, th);                    \
    }
#define GPB_P(name, mask)                                                                              \
    int gpb_periodic_##name(double* out, const double* x1, int64_t n1, const double* x2, int64_t n2,  \
                            double h, double w, double p) {                                            \
        const double th[3] = {h, w, p};                                                                \
        return gpb_kernel_slices_host(GPB_PERIODIC, mask, out, x1, n1, x2, n2, th);                    \
    }
GPB_G(K, 0x01u)
GPB_G(jacobian, 0x06u)
GPB_G(hessian, 0x78u)
GPB_G(dK_dh, 1u << 1)
GPB_G(dK_dw, 1u << 2)
GPB_G(d2K_dhdh, 1u << 3)
GPB_G(d2K_dhdw, 1u << 4)
GPB_G(d2K_dwdh, 1u << 5)
GPB_G(d2K_dwdw, 1u << 6)
GPB_P(K, 0x0001u)
GPB_P(jacobian, 0x000Eu)
GPB_P(hessian, 0x1FF0u)
GPB_P(dK_dh, 1u << 1)
GPB_P(dK_dw, 1u << 2)
GPB_P(dK_dp, 1u << 3)
GPB_P(d2K_dhdh, 1u << 4)
GPB_P(d2K_dhdw, 1u << 5)
GPB_P(d2K_dhdp, 1u << 6)
GPB_P(d2K_dwdh, 1u << 7)
GPB_P(d2K_dwdw, 1u << 8)
GPB_P(d2K_dwdp, 1u << 9)
GPB_P(d2K_dpdh, 1u << 10)
GPB_P(d2K_dpdw, 1u << 11)
GPB_P(d2K_dpdp, 1u << 12)

}  // extern "C"
