// libpkb200.so -- host side of the C ABI declared in include/pkb200.h.
//
// Orchestrates the sm_100a kernels of phase1.cuh (per-day dispersal kernels,
// ParasitoidModel.py) and chain.cuh (daily convolution chain, CalcSol.py /
// cuda_lib.py).  Everything runs on one CUDA stream per context; the only
// host<->device synchronisations inside a solve are the two places where the
// reference's algorithm is data dependent in *size*: the accumulation-window
// radius after the drift pass and the crop radii (-> torus size) after phase 1.
//
// There is no CPU code path in this library.  (tests/emul builds these same
// sources against a CPU fiber emulation of CUDA purely to unit-test indexing
// logic in the GPU-less build container; see tests/emul/emul_cuda.h.)
#include "../../include/pkb200.h"
#include "phase1.cuh"
#include "chain.cuh"
#include "bchain.cuh"
#include "project.cuh"
#include "dist.cuh"

#include <algorithm>
#include <array>
#include <chrono>
#include <condition_variable>
#include <deque>
#include <exception>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <map>
#include <memory>
#include <mutex>
#include <string>
#include <thread>
#include <vector>

using namespace pkb;

static_assert(sizeof(pkb_day_meta) == sizeof(DayMeta), "pkb_day_meta layout");
static_assert(sizeof(pkb_step_meta) == sizeof(StepMeta), "pkb_step_meta layout");

// ---------------------------------------------------------------------------
// errors
// ---------------------------------------------------------------------------
static thread_local std::string g_err;

static int fail(int code, const char* fmt, ...) {
    char buf[1024];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof buf, fmt, ap);
    va_end(ap);
    g_err = buf;
    return code;
}

#define CU(call)                                                                           \
    do {                                                                                   \
        cudaError_t e_ = (call);                                                           \
        if (e_ != cudaSuccess) return fail(PKB_ECUDA, "%s failed: %s (%s:%d)", #call,       \
                                           cudaGetErrorString(e_), __FILE__, __LINE__);    \
    } while (0)
#define TRY(call)            \
    do {                     \
        int rc_ = (call);    \
        if (rc_) return rc_; \
    } while (0)

extern "C" const char* pkb_last_error(void) { return g_err.c_str(); }
extern "C" int pkb_version(void) { return 100; }

// ---------------------------------------------------------------------------
// context: stream, caching allocators, FFT plans
// ---------------------------------------------------------------------------
struct PlanRec {
    FftPlan plan;
    cplx* tw0;
    cplx* twb;
    int2* pair;
    int* perm;
    int* slot;
    int4* spair;
};

struct pkb_ctx {
    int device;
    cudaStream_t stream;
    cudaStream_t own_stream;   // the stream created with the context (ctx->stream unless pkb_set_stream lent another one)
    cudaStream_t aux;       // side stream: output emission overlapped with the next chain step
    cudaStream_t cp;        // copy stream: per-day COO compaction + D2H while the chain is still running
    cudaStream_t cp2;       // second copy stream: the days alternate, so one day's compaction overlaps the previous day's copies
    cudaEvent_t ev_cp2;
    std::vector<cudaEvent_t> day_events;
    std::vector<cudaEvent_t> emit_events;   // day d of the fused solve has been emitted (the COO worker thread waits for it)
    int fin_clusters;       // k_day_finalize as clusters of 8 CTAs per problem when a launch has fewer problems than SMs (option "fin_clusters", default 1)
    int cohort_lanes;       // population model with a release of several days: the cohort back-solves of day n run on a child context next to the
                            // main chain's step n + 1 (option "cohort_lanes", default 1)
    int coo_thread;         // the per-day COO / CSR compaction + D2H of pkb_solve is driven by a helper host thread (option "coo_thread", default 1)
    std::vector<cudaEvent_t> win_events;
    cudaEvent_t ev_cp;
    size_t coo_hint;        // triplets of the last solve (initial size of the next one's host buffers)
    cudaEvent_t ev_step[2], ev_emit[2];
    long long launches;
    std::map<int, PlanRec> plans;
    // size-bucketed free lists: repeated solves (MCMC proposals, bench steps)
    // reuse their buffers instead of paying cudaMalloc/cudaMallocHost again
    std::map<size_t, std::vector<void*> > dev_free, host_free;
    std::map<void*, size_t> dev_live, host_live;
    cudaEvent_t ev[5];
    cudaEvent_t marks[8];
    double timing[4];
    // optional per-kernel device timing (pkb_profile_*): event pairs around launches
    bool prof_on;
    struct ProfRec { const char* name; cudaEvent_t a, b; };
    std::vector<ProfRec> prof_pending;
    std::vector<cudaEvent_t> prof_pool;
    std::map<std::string, std::pair<long long, double> > prof_acc;
    int stencil_max_radius;
    int fft_threads;
    int use_windows;        // fused solve: support-window steps (option "windows", default on)
    int use_fusion;         // fused solve: inverse row pass + next forward row pass in one kernel (option "fuse_rows")
    int use_trunc_torus;    // steps from a truncated (flagged) state on a torus >= D + 2m (option "trunc_torus")
    double ring_tol;        // ring-growth decisions closer than this to cdf_eps are re-taken in the reference's summation order (option "ring_tol")
    int use_tau_windows;    // fused solve: support windows follow the NUMERICAL support (cells >= 1e-15) instead of the exact one (option "tau_windows")
    int tau_lag;            // ... using the extent measured this many steps back (option "tau_lag": the host never waits for the step in flight)
    std::vector<cudaEvent_t> box_events;
    cudaEvent_t ring_events[4];
    int use_rowwin;         // ... restricted to the rows that can hold anything (option "spectral_rows", chain.cuh PKB_SPEC_TAU)
    int spec_min_reach;     // ... armed only if the exact support stays inside the domain for this many steps (option "spectral_min_reach")
    int use_spectral;       // fused solve: spectral-resident steps while nothing of consequence lies outside the domain (option "spectral")
    int emit_ctas;          // fused solve: side-stream emission as this many persistent 64-thread CTAs (option "emit_ctas"; 0, the
                            // default: one 256-thread CTA per row -- the small persistent CTAs measured slower)
    int rows_desc;          // k_rows_inv takes its jobs in descending order (short fold jobs last; option "rows_desc", default 1)
    int occ_cap;            // resident CTAs per SM the persistent grids are sized for (4; tuning hook PKB_FFT_OCC)
    int use_step_torus;     // whole-torus steps on the smallest 7-smooth torus >= P + 2m of THAT day's kernel (option "step_torus")
    int batch_group;        // pkb_solve_batch: proposals per kernel-construction group (option "batch_group", default PKB_BATCH_GROUP)
    int batch_lanes;        // pkb_solve_batch: proposals in flight at once, each on its own child context (option "batch_lanes")
    int batch_chain;        // pkb_solve_batch: step n of every proposal of a group in ONE launch per pass (bchain.cuh; option "batch_chain", default 1)
    int batch_occ;          // ... with this many resident CTAs per SM (option "batch_occ", default PKB_BCH_B)
    int batch_threads;      // ... enqueued by one host thread per lane (option "batch_threads", default 1 = yes): a Kalbar-sized
                            // chain is ~90 launches of 20-60 us kernels, so one thread issuing four lanes is the bottleneck
    std::vector<pkb_ctx*> lanes;     // child contexts (own streams, pools and plans) of the likelihood batch
    cudaEvent_t ev_lane;
    cudaEvent_t ev_kr[2];      // fused solve: kernel row spectra batched on the side stream (before / after)
    int sm_count;
    int max_smem;
};

static size_t bucket(size_t n) {
    if (n < 4096) return 4096;
    size_t p = 4096;
    while (p < n) p <<= 1;          // next power of two ...
    const size_t step = p >> 3;     // ... refined to eighths
    return ((n + step - 1) / step) * step;
}

// (the pools of a context are touched by its COO worker thread as well as by the thread that owns the context)
static std::mutex g_pool_mu;
static int dev_alloc(pkb_ctx* ctx, size_t bytes, void** out) {
    std::lock_guard<std::mutex> lock(g_pool_mu);
    const size_t b = bucket(bytes);
    std::vector<void*>& fl = ctx->dev_free[b];
    if (!fl.empty()) {
        *out = fl.back();
        fl.pop_back();
    } else {
        void* p = nullptr;
        cudaError_t e = cudaMalloc(&p, b);
        if (e != cudaSuccess) {
            // drop the cache and retry once
            for (auto& kv : ctx->dev_free) {
                for (void* q : kv.second) cudaFree(q);
                kv.second.clear();
            }
            e = cudaMalloc(&p, b);
            if (e != cudaSuccess) return fail(PKB_ENOMEM, "cudaMalloc(%zu bytes) failed: %s", b, cudaGetErrorString(e));
        }
        *out = p;
    }
    ctx->dev_live[*out] = b;
    return 0;
}
static void dev_release(pkb_ctx* ctx, void* p) {
    if (!p) return;
    std::lock_guard<std::mutex> lock(g_pool_mu);
    auto it = ctx->dev_live.find(p);
    if (it == ctx->dev_live.end()) return;
    ctx->dev_free[it->second].push_back(p);
    ctx->dev_live.erase(it);
}
static int host_alloc(pkb_ctx* ctx, size_t bytes, void** out) {
    std::lock_guard<std::mutex> lock(g_pool_mu);
    const size_t b = bucket(bytes);
    std::vector<void*>& fl = ctx->host_free[b];
    if (!fl.empty()) {
        *out = fl.back();
        fl.pop_back();
    } else {
        void* p = nullptr;
        cudaError_t e = cudaMallocHost(&p, b);
        if (e != cudaSuccess) return fail(PKB_ENOMEM, "cudaMallocHost(%zu bytes) failed: %s", b, cudaGetErrorString(e));
        *out = p;
    }
    ctx->host_live[*out] = b;
    return 0;
}
static void host_release(pkb_ctx* ctx, void* p) {
    if (!p) return;
    std::lock_guard<std::mutex> lock(g_pool_mu);
    auto it = ctx->host_live.find(p);
    if (it == ctx->host_live.end()) return;
    ctx->host_free[it->second].push_back(p);
    ctx->host_live.erase(it);
}

// typed handle on a pooled device buffer
template <class T>
struct DBuf {
    pkb_ctx* ctx = nullptr;
    T* p = nullptr;
    size_t n = 0;
    int alloc(pkb_ctx* c, size_t count) {
        release();
        ctx = c;
        void* q = nullptr;
        int rc = dev_alloc(c, std::max<size_t>(count, 1) * sizeof(T), &q);
        if (rc) return rc;
        p = static_cast<T*>(q);
        n = count;
        return 0;
    }
    void release() {
        if (p && ctx) dev_release(ctx, p);
        p = nullptr;
        n = 0;
    }
    ~DBuf() { release(); }
    DBuf() {}
    DBuf(const DBuf&) = delete;
    DBuf& operator=(const DBuf&) = delete;
};
template <class T>
struct HBuf {
    pkb_ctx* ctx = nullptr;
    T* p = nullptr;
    size_t n = 0;
    int alloc(pkb_ctx* c, size_t count) {
        release();
        ctx = c;
        void* q = nullptr;
        int rc = host_alloc(c, std::max<size_t>(count, 1) * sizeof(T), &q);
        if (rc) return rc;
        p = static_cast<T*>(q);
        n = count;
        return 0;
    }
    void release() {
        if (p && ctx) host_release(ctx, p);
        p = nullptr;
        n = 0;
    }
    ~HBuf() { release(); }
    HBuf() {}
    HBuf(const HBuf&) = delete;
    HBuf& operator=(const HBuf&) = delete;
};

static cudaEvent_t prof_event(pkb_ctx* ctx) {
    cudaEvent_t e = nullptr;
    if (!ctx->prof_pool.empty()) {
        e = ctx->prof_pool.back();
        ctx->prof_pool.pop_back();
    } else {
        cudaEventCreate(&e);
    }
    return e;
}
static void prof_begin(pkb_ctx* ctx, const char* name, cudaStream_t strm) {
    pkb_ctx::ProfRec r;
    r.name = name;
    r.a = prof_event(ctx);
    r.b = prof_event(ctx);
    cudaEventRecord(r.a, strm);
    ctx->prof_pending.push_back(r);
}
static void prof_end(pkb_ctx* ctx, cudaStream_t strm) { cudaEventRecord(ctx->prof_pending.back().b, strm); }
// fold finished event pairs into the accumulators (call after a stream sync)
static void prof_collect(pkb_ctx* ctx) {
    for (auto& r : ctx->prof_pending) {
        float ms = 0.f;
        if (cudaEventElapsedTime(&ms, r.a, r.b) == cudaSuccess) {
            auto& acc = ctx->prof_acc[r.name];
            acc.first += 1;
            acc.second += ms;
        }
        ctx->prof_pool.push_back(r.a);
        ctx->prof_pool.push_back(r.b);
    }
    ctx->prof_pending.clear();
}

#define LAUNCH_ON(ctx, strm, kern, grid, block, smem, ...)                    \
    do {                                                                      \
        if ((ctx)->prof_on) prof_begin((ctx), #kern, (strm));                 \
        PKB_LAUNCH(kern, grid, block, smem, (strm), __VA_ARGS__);             \
        if ((ctx)->prof_on) prof_end((ctx), (strm));                          \
        (ctx)->launches++;                                                    \
    } while (0)
#define LAUNCH(ctx, kern, grid, block, smem, ...) LAUNCH_ON(ctx, (ctx)->stream, kern, grid, block, smem, __VA_ARGS__)
// same launch, accounted under another name in the per-kernel profile (support-window steps)
#define LAUNCH_AS(ctx, name, kern, grid, block, smem, ...)                    \
    do {                                                                      \
        if ((ctx)->prof_on) prof_begin((ctx), (name), (ctx)->stream);         \
        PKB_LAUNCH(kern, grid, block, smem, (ctx)->stream, __VA_ARGS__);      \
        if ((ctx)->prof_on) prof_end((ctx), (ctx)->stream);                   \
        (ctx)->launches++;                                                    \
    } while (0)

static int check_launches(pkb_ctx* ctx, const char* where) {
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return fail(PKB_ECUDA, "kernel launch failed in %s: %s", where, cudaGetErrorString(e));
    (void)ctx;
    return 0;
}
static int sync_check(pkb_ctx* ctx, const char* where) {
    TRY(check_launches(ctx, where));
    cudaError_t e = cudaStreamSynchronize(ctx->stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(ctx->aux);
    if (e == cudaSuccess) e = cudaStreamSynchronize(ctx->cp);
    if (e == cudaSuccess) e = cudaStreamSynchronize(ctx->cp2);
    if (e != cudaSuccess) return fail(PKB_ECUDA, "stream synchronize failed in %s: %s", where, cudaGetErrorString(e));
    if (!ctx->prof_pending.empty()) prof_collect(ctx);
    return 0;
}

static const int kMaxSmem = 232448;   // 227 KB opt-in per CTA on sm_100
static const int kStaticSmemReserve = 8192 + 1024;   // largest static __shared__ use of any kernel here

// allow a kernel the full opt-in dynamic shared memory (minus its static use)
template <class F>
static int opt_in_smem(F kern) {
    cudaFuncAttributes fa;
    CU(cudaFuncGetAttributes(&fa, kern));
    CU(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, kMaxSmem - (int)fa.sharedSizeBytes));
    // always configure the full shared-memory carveout: left to its default the
    // driver sized it for ONE resident CTA of k_cols (ncu: 102 KB config)
    CU(cudaFuncSetAttribute(kern, cudaFuncAttributePreferredSharedMemoryCarveout, 100));
    return 0;
}

extern "C" int pkb_create(int device, pkb_ctx** out) {
    if (!out) return fail(PKB_EINVAL, "pkb_create: out is NULL");
    *out = nullptr;
    int ndev = 0;
    cudaError_t e = cudaGetDeviceCount(&ndev);
    if (e != cudaSuccess || ndev <= 0)
        return fail(PKB_ECUDA, "pkb_create: no CUDA device available (%s); this library has no CPU path",
                    e != cudaSuccess ? cudaGetErrorString(e) : "device count is 0");
    if (device < 0 || device >= ndev) return fail(PKB_EINVAL, "pkb_create: device %d out of range [0, %d)", device, ndev);
    CU(cudaSetDevice(device));
    pkb_ctx* ctx = new pkb_ctx();
    ctx->device = device;
    ctx->launches = 0;
    ctx->stencil_max_radius = 3;
    ctx->fft_threads = PKB_ROWS_T;
    ctx->use_windows = 1;
    ctx->use_fusion = 1;
    ctx->batch_lanes = 4;
    ctx->batch_threads = 1;
    ctx->coo_thread = 1;
    ctx->cohort_lanes = 1;
    ctx->fin_clusters = 1;
    ctx->batch_chain = 1;
    ctx->batch_occ = PKB_BCH_B;
    ctx->batch_group = 32;
    ctx->use_step_torus = 1;
    ctx->use_trunc_torus = 1;
    ctx->use_spectral = 1;
    ctx->spec_min_reach = 4;
    ctx->use_rowwin = 1;
    ctx->use_tau_windows = 1;
    ctx->tau_lag = 1;
    for (int i = 0; i < 4; ++i) CU(cudaEventCreateWithFlags(&ctx->ring_events[i], cudaEventDisableTiming));
    ctx->ring_tol = 1e-12;
    ctx->occ_cap = 4;
    if (const char* env = getenv("PKB_FFT_OCC")) ctx->occ_cap = std::max(1, std::min(16, atoi(env)));
    CU(cudaEventCreateWithFlags(&ctx->ev_lane, cudaEventDisableTiming));
    for (int i = 0; i < 2; ++i) CU(cudaEventCreateWithFlags(&ctx->ev_kr[i], cudaEventDisableTiming));
    CU(cudaDeviceGetAttribute(&ctx->sm_count, cudaDevAttrMultiProcessorCount, device));
    ctx->max_smem = kMaxSmem - kStaticSmemReserve;
    ctx->emit_ctas = 0;
    ctx->rows_desc = 1;
    ctx->prof_on = false;
    for (int i = 0; i < 4; ++i) ctx->timing[i] = 0.0;
    CU(cudaStreamCreate(&ctx->stream));
    ctx->own_stream = ctx->stream;
    {
        // the side streams carry short kernels that must slip in between the persistent FFT kernels
        // of the main stream: highest priority, so their CTAs are placed first whenever SM slots free up
        int lo = 0, hi = 0;
        CU(cudaDeviceGetStreamPriorityRange(&lo, &hi));
        CU(cudaStreamCreateWithPriority(&ctx->aux, cudaStreamNonBlocking, hi));
        CU(cudaStreamCreateWithPriority(&ctx->cp, cudaStreamNonBlocking, hi));
        CU(cudaStreamCreateWithPriority(&ctx->cp2, cudaStreamNonBlocking, hi));
    }
    CU(cudaEventCreateWithFlags(&ctx->ev_cp, cudaEventDisableTiming));
    CU(cudaEventCreateWithFlags(&ctx->ev_cp2, cudaEventDisableTiming));
    ctx->coo_hint = 0;
    for (int i = 0; i < 2; ++i) {
        CU(cudaEventCreateWithFlags(&ctx->ev_step[i], cudaEventDisableTiming));
        CU(cudaEventCreateWithFlags(&ctx->ev_emit[i], cudaEventDisableTiming));
    }
    for (int i = 0; i < 5; ++i) CU(cudaEventCreate(&ctx->ev[i]));
    for (int i = 0; i < 8; ++i) CU(cudaEventCreate(&ctx->marks[i]));
    TRY(opt_in_smem(k_rows_fwd));
    TRY(opt_in_smem(k_kernel_rows));
    TRY(opt_in_smem(k_kernel_rows_batch));
    TRY(opt_in_smem(k_cols));
    TRY(opt_in_smem(k_rows_inv));
    TRY(opt_in_smem(kb_rows_fwd));
    TRY(opt_in_smem(kb_cols));
    TRY(opt_in_smem(kb_rows_inv));
    TRY(opt_in_smem(k_cols_dist));
    TRY(opt_in_smem(k_rows_inv_dist));
    TRY(opt_in_smem(k_fft_test));
    TRY(opt_in_smem(k_period));
    TRY(opt_in_smem(k_hprob));
    TRY(opt_in_smem(k_stencil));
    *out = ctx;
    return 0;
}

extern "C" int pkb_destroy(pkb_ctx* ctx) {
    if (!ctx) return 0;
    cudaSetDevice(ctx->device);
    cudaStreamSynchronize(ctx->stream);
    for (pkb_ctx* lane : ctx->lanes) pkb_destroy(lane);
    cudaEventDestroy(ctx->ev_lane);
    for (int i = 0; i < 2; ++i) cudaEventDestroy(ctx->ev_kr[i]);
    for (auto& kv : ctx->plans) {
        cudaFree(kv.second.tw0);
        cudaFree(kv.second.twb);
        cudaFree(kv.second.pair);
        cudaFree(kv.second.perm);
        cudaFree(kv.second.slot);
        cudaFree(kv.second.spair);
    }
    for (auto& kv : ctx->dev_free)
        for (void* p : kv.second) cudaFree(p);
    for (auto& kv : ctx->dev_live) cudaFree(kv.first);
    for (auto& kv : ctx->host_free)
        for (void* p : kv.second) cudaFreeHost(p);
    for (auto& kv : ctx->host_live) cudaFreeHost(kv.first);
    for (int i = 0; i < 5; ++i) cudaEventDestroy(ctx->ev[i]);
    for (int i = 0; i < 8; ++i) cudaEventDestroy(ctx->marks[i]);
    for (auto& r : ctx->prof_pending) { cudaEventDestroy(r.a); cudaEventDestroy(r.b); }
    for (cudaEvent_t e : ctx->prof_pool) cudaEventDestroy(e);
    for (int i = 0; i < 2; ++i) { cudaEventDestroy(ctx->ev_step[i]); cudaEventDestroy(ctx->ev_emit[i]); }
    for (cudaEvent_t e : ctx->day_events) cudaEventDestroy(e);
    for (cudaEvent_t e : ctx->emit_events) cudaEventDestroy(e);
    for (cudaEvent_t e : ctx->win_events) cudaEventDestroy(e);
    for (cudaEvent_t e : ctx->box_events) cudaEventDestroy(e);
    for (int i = 0; i < 4; ++i) cudaEventDestroy(ctx->ring_events[i]);
    cudaEventDestroy(ctx->ev_cp);
    cudaEventDestroy(ctx->ev_cp2);
    cudaStreamDestroy(ctx->cp2);
    cudaStreamDestroy(ctx->cp);
    cudaStreamDestroy(ctx->aux);
    cudaStreamDestroy(ctx->own_stream);
    delete ctx;
    return 0;
}

extern "C" int pkb_sync(pkb_ctx* ctx) {
    if (!ctx) return fail(PKB_EINVAL, "pkb_sync: ctx is NULL");
    return sync_check(ctx, "pkb_sync");
}

extern "C" int pkb_set_option(pkb_ctx* ctx, const char* key, double value) {
    if (!ctx || !key) return fail(PKB_EINVAL, "pkb_set_option: NULL argument");
    if (!strcmp(key, "stencil_max_radius")) {
        if (value < -1 || value > 24) return fail(PKB_EINVAL, "stencil_max_radius must be in [-1, 24]");
        ctx->stencil_max_radius = (int)value;
        return 0;
    }
    if (!strcmp(key, "fuse_rows")) {
        ctx->use_fusion = value != 0;
        return 0;
    }
    if (!strcmp(key, "rows_desc")) {
        ctx->rows_desc = value != 0;
        return 0;
    }
    if (!strcmp(key, "emit_ctas")) {
        if (value < 0 || value > 65535) return fail(PKB_EINVAL, "emit_ctas must be 0..65535");
        ctx->emit_ctas = (int)value;
        return 0;
    }
    if (!strcmp(key, "trunc_torus")) {
        ctx->use_trunc_torus = value != 0;
        return 0;
    }
    if (!strcmp(key, "ring_tol")) {
        if (!(value >= 0)) return fail(PKB_EINVAL, "ring_tol must be >= 0");
        ctx->ring_tol = value;
        return 0;
    }
    if (!strcmp(key, "tau_windows")) {
        ctx->use_tau_windows = value != 0;
        return 0;
    }
    if (!strcmp(key, "tau_lag")) {
        if (value < 1 || value > 8) return fail(PKB_EINVAL, "tau_lag must be 1..8");
        ctx->tau_lag = (int)value;
        return 0;
    }
    if (!strcmp(key, "spectral_rows")) {
        ctx->use_rowwin = value != 0;
        return 0;
    }
    if (!strcmp(key, "spectral_min_reach")) {
        ctx->spec_min_reach = (int)value;
        return 0;
    }
    if (!strcmp(key, "spectral")) {
        ctx->use_spectral = value != 0;
        return 0;
    }
    if (!strcmp(key, "step_torus")) {
        ctx->use_step_torus = value != 0;
        return 0;
    }
    if (!strcmp(key, "batch_group")) {
        if (value < 1 || value > 1024) return fail(PKB_EINVAL, "batch_group must be 1..1024");
        ctx->batch_group = (int)value;
        return 0;
    }
    if (!strcmp(key, "batch_occ")) {
        if (value < 1 || value > 16) return fail(PKB_EINVAL, "batch_occ must be 1..16");
        ctx->batch_occ = (int)value;
        return 0;
    }
    if (!strcmp(key, "batch_chain")) {
        ctx->batch_chain = value != 0;
        return 0;
    }
    if (!strcmp(key, "fin_clusters")) {
        ctx->fin_clusters = value != 0;
        return 0;
    }
    if (!strcmp(key, "cohort_lanes")) {
        ctx->cohort_lanes = value != 0;
        return 0;
    }
    if (!strcmp(key, "coo_thread")) {
        ctx->coo_thread = value != 0;
        return 0;
    }
    if (!strcmp(key, "batch_threads")) {
        ctx->batch_threads = value != 0;
        return 0;
    }
    if (!strcmp(key, "batch_lanes")) {
        if (value < 1 || value > 8) return fail(PKB_EINVAL, "batch_lanes must be 1..8");
        ctx->batch_lanes = (int)value;
        return 0;
    }
    if (!strcmp(key, "windows")) {
        ctx->use_windows = value != 0;
        return 0;
    }
    if (!strcmp(key, "fft_threads")) {
        const int t = (int)value;
        if (t < 32 || t > PKB_ROWS_T || (t & 31)) return fail(PKB_EINVAL, "fft_threads must be a multiple of 32 in [32, %d]", PKB_ROWS_T);
        ctx->fft_threads = t;
        return 0;
    }
    return fail(PKB_EINVAL, "pkb_set_option: unknown key '%s'", key);
}

extern "C" int pkb_timing(pkb_ctx* ctx, double out_ms[4]) {
    if (!ctx || !out_ms) return fail(PKB_EINVAL, "pkb_timing: NULL argument");
    for (int i = 0; i < 4; ++i) out_ms[i] = ctx->timing[i];
    return 0;
}

extern "C" long long pkb_launch_count(pkb_ctx* ctx) { return ctx ? ctx->launches : -1; }

extern "C" int pkb_mark(pkb_ctx* ctx, int slot) {
    if (!ctx || slot < 0 || slot >= 8) return fail(PKB_EINVAL, "pkb_mark: bad argument");
    CU(cudaSetDevice(ctx->device));
    CU(cudaEventRecord(ctx->marks[slot], ctx->stream));
    return 0;
}

extern "C" int pkb_elapsed_ms(pkb_ctx* ctx, int slot_a, int slot_b, double* ms) {
    if (!ctx || !ms || slot_a < 0 || slot_a >= 8 || slot_b < 0 || slot_b >= 8) return fail(PKB_EINVAL, "pkb_elapsed_ms: bad argument");
    CU(cudaSetDevice(ctx->device));
    CU(cudaEventSynchronize(ctx->marks[slot_b]));
    float f = 0.f;
    CU(cudaEventElapsedTime(&f, ctx->marks[slot_a], ctx->marks[slot_b]));
    *ms = f;
    return 0;
}

extern "C" int pkb_profile_enable(pkb_ctx* ctx, int on) {
    if (!ctx) return fail(PKB_EINVAL, "pkb_profile_enable: NULL context");
    TRY(sync_check(ctx, "pkb_profile_enable"));
    ctx->prof_on = on != 0;
    return 0;
}

extern "C" int pkb_profile_reset(pkb_ctx* ctx) {
    if (!ctx) return fail(PKB_EINVAL, "pkb_profile_reset: NULL context");
    TRY(sync_check(ctx, "pkb_profile_reset"));
    ctx->prof_acc.clear();
    return 0;
}

extern "C" int pkb_profile_get(pkb_ctx* ctx, const char* kernel, long long* count, double* total_ms) {
    if (!ctx || !kernel) return fail(PKB_EINVAL, "pkb_profile_get: NULL argument");
    TRY(sync_check(ctx, "pkb_profile_get"));
    auto it = ctx->prof_acc.find(kernel);
    if (count) *count = it == ctx->prof_acc.end() ? 0 : it->second.first;
    if (total_ms) *total_ms = it == ctx->prof_acc.end() ? 0.0 : it->second.second;
    return 0;
}

// ---------------------------------------------------------------------------
// FFT plans
// ---------------------------------------------------------------------------
static bool is_smooth(int n) {
    if (n < 1) return false;
    const int pr[4] = {2, 3, 5, 7};
    for (int p : pr)
        while (n % p == 0) n /= p;
    return n == 1;
}

extern "C" int pkb_smooth_len(int n) {
    if (n < 1) n = 1;
    while (!is_smooth(n)) ++n;
    return n;
}

// fewest factors of n from the register radices (iterative deepening; n is 7-smooth)
static const int kRadices[] = {12, 10, 9, 8, 7, 6, 5, 4, 3, 2};
static bool factor_depth(int n, int depth, int max_r, std::vector<int>& out) {
    if (n == 1) return true;
    if (depth == 0) return false;
    for (int r : kRadices) {
        if (r > max_r || n % r) continue;
        out.push_back(r);
        if (factor_depth(n / r, depth - 1, r, out)) return true;
        out.pop_back();
    }
    return false;
}
static bool min_factor(int n, std::vector<int>& out) {
    if (n == 1) { out.assign(1, 2); return false; }
    for (int depth = 1; depth <= PKB_FFT_MAX_STAGES; ++depth) {
        out.clear();
        if (factor_depth(n, depth, 12, out)) return true;
    }
    return false;
}

static int get_plan(pkb_ctx* ctx, int N, FftPlan* out) {
    auto it = ctx->plans.find(N);
    if (it != ctx->plans.end()) {
        *out = it->second.plan;
        return 0;
    }
    if (!is_smooth(N)) return fail(PKB_EINVAL, "FFT length %d is not 7-smooth", N);
    PlanRec rec;
    FftPlan& p = rec.plan;
    p.N = N;
    p.nstage = 0;
    // fewest passes over shared memory: factor N into the fewest register radices
    std::vector<int> fac;
    if (!min_factor(N, fac)) return fail(PKB_EINVAL, "FFT length %d cannot be factored into the supported radices", N);
    std::sort(fac.begin(), fac.end(), [](int a, int b) { return a > b; });
    bool forced = false;
    if (const char* env = getenv("PKB_FFT_PLAN")) {
        // tuning hook: explicit radix sequence "12,8,7,7" (used as is when its product is N)
        std::vector<int> f2;
        long long prod = 1;
        for (const char* q = env; *q;) {
            const int r = atoi(q);
            if (r >= 2 && r <= 12 && r != 11) { f2.push_back(r); prod *= r; }
            while (*q && *q != ',') ++q;
            if (*q == ',') ++q;
        }
        if (prod == N && !f2.empty()) { fac = f2; forced = true; }
    }
    // last radix (k_cols keeps whole last-stage blocks per thread, stride R_last
    // complex between threads): the largest odd one is bank-conflict free
    int last = -1;
    for (size_t i = 0; i < fac.size() && last < 0; ++i)
        if (fac[i] & 1) last = (int)i;
    if (last < 0) last = 0;
    if (forced) last = (int)fac.size() - 1;
    const int rl = fac[last];
    fac.erase(fac.begin() + last);
    fac.push_back(rl);
    if ((int)fac.size() > PKB_FFT_MAX_STAGES) return fail(PKB_ELIMIT, "FFT length %d needs too many stages", N);
    // Threads per CTA.  The pure transform runs faster with MORE resident CTAs per SM (independent
    // CTAs sit in different phases and hide each other's shared-memory latency and barriers: core
    // benchmark at N = 2352, 2 x 256 threads 6.1 k cycles per transform, 4 x 128 threads 4.8 k), and
    // the register file holds 512 threads of these kernels; so the CTA shrinks as more transforms fit
    // into shared memory: 2 x 256, 3 x 160 or 4 x 128 threads.
    {
        int ntw = 0, M = N / fac[0];
        for (size_t i = 1; i < fac.size(); ++i) { ntw += M / fac[i]; M /= fac[i]; }
        const size_t foot = (size_t)(N + ntw) * sizeof(cplx) + 1024;      // + the per-CTA reservation
        const int fit = (int)(233472 / foot);
        p.threads = std::min(ctx->fft_threads, fit >= 4 ? 128 : (fit == 3 ? 160 : 256));
        if (const char* env = getenv("PKB_FFT_T")) {          // tuning hook
            const int t = atoi(env);
            if (t >= 32 && t <= 256 && t % 32 == 0) p.threads = t;
        }
    }
    // k_cols geometry: threads per column and last-stage blocks per thread
    {
        const int nbl = N / rl;
        int best_t = 0, best_kb = 0;
        // at least 4 warps per column whenever there is work for them (a 1-warp CTA would "waste" the least)
        // (when three or four transforms fit an SM the column kernel simply takes the row kernels' CTA size:
        // measured at N = 4704, 3 x 160 threads 306 us against 313 us for 3 x 128 and 326 us for 2 x 224)
        for (int t = p.threads < 256 ? p.threads : (nbl >= 128 ? 128 : 32); t <= std::min(PKB_COLS_TMAX, p.threads); t += 32) {
            const int kb = (nbl + t - 1) / t;
            // least idle work first, then the most threads
            if (!best_t || t * kb < best_t * best_kb || (t * kb == best_t * best_kb && t > best_t)) { best_t = t; best_kb = kb; }
        }
        if (const char* env = getenv("PKB_COLS_T")) {         // tuning hook
            const int t = atoi(env);
            if (t >= 32 && t <= PKB_COLS_TMAX && t % 32 == 0) { best_t = t; best_kb = (nbl + t - 1) / t; }
        }
        p.cols_threads = best_t;
        p.cols_kb = best_kb;
    }
    p.rpack = 0;
    for (int r : fac) {
        p.rpack |= (unsigned long long)r << (4 * p.nstage);
        ++p.nstage;
    }
    // base twiddles exp(-2 pi i k / M_s), k < M_s / R_s: stage 0 -> tw0 (global), stages 1.. -> twb (shared memory)
    const long double twopi = 2.0L * 3.14159265358979323846264338327950288L;
    std::vector<cplx> tw0, twb;
    {
        int M = N;
        for (int s = 0; s < p.nstage; ++s) {
            const int Ms = M / fac[s];
            for (int k = 0; k < Ms; ++k) {
                const long double a = -twopi * (long double)k / (long double)M;
                (s == 0 ? tw0 : twb).push_back(cmake((double)cosl(a), (double)sinl(a)));
            }
            M = Ms;
        }
        if (twb.empty()) twb.push_back(cmake(1.0, 0.0));
    }
    p.ntw = p.nstage > 1 ? (int)twb.size() : 0;
    std::vector<int> perm(N);
    for (int k = 0; k < N; ++k) {
        int f = k, M = N, pos = 0;
        for (int s = 0; s < p.nstage; ++s) {
            const int R = fac[s], Ms = M / R;
            pos += (f % R) * Ms;
            f /= R;
            M = Ms;
        }
        perm[k] = pos;
    }
    std::vector<int2> pair(N);
    for (int k = 0; k < N; ++k) pair[k] = make_int2(perm[k], perm[(N - k) % N]);
    CU(cudaMalloc((void**)&rec.tw0, sizeof(cplx) * tw0.size()));
    CU(cudaMemcpy(rec.tw0, tw0.data(), sizeof(cplx) * tw0.size(), cudaMemcpyHostToDevice));
    p.tw0 = rec.tw0;
    CU(cudaMalloc((void**)&rec.twb, sizeof(cplx) * twb.size()));
    CU(cudaMalloc((void**)&rec.pair, sizeof(int2) * (N + 2)));
    CU(cudaMemset(rec.pair, 0, sizeof(int2) * (N + 2)));
    CU(cudaMalloc((void**)&rec.perm, sizeof(int) * N));
    CU(cudaMemcpy(rec.twb, twb.data(), sizeof(cplx) * twb.size(), cudaMemcpyHostToDevice));
    CU(cudaMemcpy(rec.pair, pair.data(), sizeof(int2) * N, cudaMemcpyHostToDevice));
    CU(cudaMemcpy(rec.perm, perm.data(), sizeof(int) * N, cudaMemcpyHostToDevice));
    p.twb = rec.twb;
    p.pair = rec.pair;
    p.perm = rec.perm;
    {
        // Slot order of the Hermitian pack / unpack loops (FftPlan::slot).  Windows of 64 (up to 1024) column pairs
        // are re-ordered greedily so that every aligned run of 8 slots (a quarter-warp, one 128-byte
        // shared-memory wavefront) takes pairs whose pos(2c) fall into 8 different 16-byte bank groups;
        // kept only if it lowers the wavefront count over all four accesses of a slot.
        const int Nc = N / 2 + 1, npair = (Nc + 1) / 2;
        auto quad = [&](int c) {
            const int k0 = 2 * c, k1 = std::min(2 * c + 1, N - 1);
            return make_int4(perm[k0], perm[(N - k0) % N], perm[k1], perm[(N - k1) % N]);
        };
        auto cost = [&](const std::vector<int>& order) {
            long long tot = 0;
            for (int s0 = 0; s0 < npair; s0 += 8) {
                int cnt[4][8] = {{0}};
                for (int s1 = s0; s1 < std::min(npair, s0 + 8); ++s1) {
                    const int4 q = quad(order[s1]);
                    cnt[0][q.x & 7]++; cnt[1][q.y & 7]++; cnt[2][q.z & 7]++; cnt[3][q.w & 7]++;
                }
                for (int a = 0; a < 4; ++a) tot += *std::max_element(cnt[a], cnt[a] + 8);
            }
            return tot;
        };
        std::vector<int> ident(npair), order;
        for (int c = 0; c < npair; ++c) ident[c] = c;
        auto greedy = [&](int win) {
            std::vector<int> ord;
            ord.reserve(npair);
            for (int w0 = 0; w0 < npair; w0 += win) {
                const int w1 = std::min(npair, w0 + win);
                std::vector<int> bucket[8];
                for (int c = w1 - 1; c >= w0; --c) bucket[perm[2 * c] & 7].push_back(c);      // (popped from the back: ascending)
                int left = w1 - w0;
                while (left > 0) {
                    int ids[8] = {0, 1, 2, 3, 4, 5, 6, 7};
                    std::stable_sort(ids, ids + 8, [&](int a, int b) { return bucket[a].size() > bucket[b].size(); });
                    const int take = std::min(8, left);
                    int got = 0;
                    for (int i = 0; i < 8 && got < take; ++i)
                        if (!bucket[ids[i]].empty()) { ord.push_back(bucket[ids[i]].back()); bucket[ids[i]].pop_back(); ++got; }
                    while (got < take)                     // fewer than 8 bank groups left: the run cannot be conflict free
                        for (int i = 0; i < 8 && got < take; ++i)
                            if (!bucket[ids[i]].empty()) { ord.push_back(bucket[ids[i]].back()); bucket[ids[i]].pop_back(); ++got; }
                    left -= take;
                }
            }
            return ord;
        };
        // smallest window that gets within 1.75 x the conflict-free count (the bank group of pos(2c) may only
        // change every R_0 R_1 bins, e.g. 80 at 10.8.8.7), else the best one found
        const long long ideal = 4LL * ((npair + 7) / 8);
        order = ident;
        long long best = cost(ident);
        if (!getenv("PKB_NO_SLOTS"))
            for (int win = 64; win <= 1024 && 4 * best > 7 * ideal; win *= 2) {
                std::vector<int> cand = greedy(win);
                const long long cst = cost(cand);
                if (cst < best) { best = cst; order.swap(cand); }
            }
        std::vector<int4> sp(npair);
        for (int s1 = 0; s1 < npair; ++s1) sp[s1] = quad(order[s1]);
        CU(cudaMalloc((void**)&rec.slot, sizeof(int) * npair));
        CU(cudaMalloc((void**)&rec.spair, sizeof(int4) * npair));
        CU(cudaMemcpy(rec.slot, order.data(), sizeof(int) * npair, cudaMemcpyHostToDevice));
        CU(cudaMemcpy(rec.spair, sp.data(), sizeof(int4) * npair, cudaMemcpyHostToDevice));
        p.slot = rec.slot;
        p.spair = rec.spair;
        if (getenv("PKB_PLAN_DEBUG"))
            fprintf(stderr, "plan N=%d: pack/unpack wavefronts per 4 accesses: natural %lld, slots %lld (ideal %lld)\n", N, cost(ident), cost(order),
                    ideal);
    }
    p.grid_rows = p.grid_cols = 0;
    if (fft_smem_bytes(p) <= (size_t)ctx->max_smem) {
        // resident CTAs per SM of the persistent kernels at this plan's footprint (at most 4)
        const size_t sm1 = fft_smem_bytes(p);
        int occ_r = 0, occ_i = 0, occ_c = 0;
        CU(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ_r, k_rows_fwd, p.threads, sm1));
        CU(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ_i, k_rows_inv, p.threads, sm1));
        CU(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ_c, k_cols, p.cols_threads, sm1));
        p.grid_rows = std::min(ctx->occ_cap, std::min(occ_r, occ_i)) * ctx->sm_count;
        p.grid_cols = std::min(ctx->occ_cap, occ_c) * ctx->sm_count;
    }
    ctx->plans[N] = rec;
    *out = p;
    return 0;
}

extern "C" int pkb_debug_fft(pkb_ctx* ctx, int n, const double* in, double* out, int inverse) {
    if (!ctx || !in || !out) return fail(PKB_EINVAL, "pkb_debug_fft: NULL argument");
    if (n < 1 || !is_smooth(n)) return fail(PKB_EINVAL, "pkb_debug_fft: n = %d is not 7-smooth", n);
    CU(cudaSetDevice(ctx->device));
    FftPlan plan;
    TRY(get_plan(ctx, n, &plan));
    if (fft_smem_bytes(plan) > (size_t)ctx->max_smem) return fail(PKB_ELIMIT, "pkb_debug_fft: n = %d exceeds shared memory", n);
    DBuf<cplx> a, b;
    TRY(a.alloc(ctx, n));
    TRY(b.alloc(ctx, n));
    CU(cudaMemcpyAsync(a.p, in, sizeof(cplx) * n, cudaMemcpyHostToDevice, ctx->stream));
    LAUNCH(ctx, k_fft_test, 1, plan.threads, fft_smem_bytes(plan), a.p, b.p, inverse, plan);
    CU(cudaMemcpyAsync(out, b.p, sizeof(cplx) * n, cudaMemcpyDeviceToHost, ctx->stream));
    return sync_check(ctx, "pkb_debug_fft");
}

// ---------------------------------------------------------------------------
// phase 1
// ---------------------------------------------------------------------------
struct pkb_kset {
    pkb_ctx* ctx;
    int nprob, periods, racc, W, rad_res;
    bool keep_pre;
    DBuf<double> acc, acc_lo, pre, hprob, loss_t;      // acc_lo: low plane of the exact accumulator (phase1.cuh, acc_add_exact)
    DBuf<PeriodInfo> pinfo;
    DBuf<DayMeta> dmeta;
    DBuf<DayParams> ddp;
    DBuf<BvnPar> bvn;
    std::vector<DayMeta> hmeta;
    std::vector<DayParams> hdp;
};

static int check_dparams(const double d[3], const char* which) {
    // Dmat's assertions (ParasitoidModel.py:276-278)
    if (!(d[0] > 0)) return fail(PKB_EINVAL, "sig_x must be positive (%s)", which);
    if (!(d[1] > 0)) return fail(PKB_EINVAL, "sig_y must be positive (%s)", which);
    if (!(-1 <= d[2] && d[2] <= 1)) return fail(PKB_EINVAL, "correlation must be between -1 and 1 (%s)", which);
    return 0;
}

// k_day_finalize: one CTA per problem when the launch has enough problems to fill the GPU, else one thread-block CLUSTER of
// PKB_FIN_SLICES CTAs per problem (distributed shared memory carries the slice sums, phase1.cuh) -- same arithmetic either way.
static int launch_day_finalize(pkb_ctx* ctx, int nprob, const DayParams* ddp, const BvnPar* bvn, int periods, double* acc, const double* acc_lo, int racc,
                               const double* loss_t, DayMeta* dmeta, double negval, double* pre, const PeriodInfo* pinfo) {
#if PKB_IS_EMUL
    const int csize = 1;
    LAUNCH(ctx, k_day_finalize, nprob, 1024, 0, ddp, bvn, periods, acc, acc_lo, racc, loss_t, dmeta, negval, pre, pinfo, csize);
#else
    // (measured: 18 / 30 problems with 5 MB windows 0.75 -> 0.25 ms; 60 problems with 1 MB windows gain nothing from 480 CTAs)
    const int csize = (ctx->fin_clusters && nprob * PKB_FIN_SLICES <= 2 * ctx->sm_count) ? PKB_FIN_SLICES : 1;
    if (csize == 1) {
        LAUNCH(ctx, k_day_finalize, nprob, 1024, 0, ddp, bvn, periods, acc, acc_lo, racc, loss_t, dmeta, negval, pre, pinfo, csize);
        return 0;
    }
    cudaLaunchConfig_t cfg;
    memset(&cfg, 0, sizeof cfg);
    cfg.gridDim = dim3(nprob * csize);
    cfg.blockDim = dim3(1024);
    cfg.dynamicSmemBytes = 0;
    cfg.stream = ctx->stream;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = csize;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    if (ctx->prof_on) prof_begin(ctx, "k_day_finalize", ctx->stream);
    CU(cudaLaunchKernelEx(&cfg, k_day_finalize, ddp, bvn, periods, acc, acc_lo, racc, loss_t, dmeta, negval, pre, pinfo, csize));
    if (ctx->prof_on) prof_end(ctx, ctx->stream);
    ctx->launches++;
#endif
    return 0;
}

// wind_dev: device pointer [nd_wind][periods][3]
static int kernels_build_dev(pkb_ctx* ctx, const double* wind_dev, int nd_wind, int periods, const pkb_day_args* args, int nprob,
                             int keep_pre, pkb_kset** out) {
    if (nprob < 1 || periods < 1 || nd_wind < 1) return fail(PKB_EINVAL, "pkb_kernels_build: empty problem set");
    pkb_kset* ks = new pkb_kset();
    ks->ctx = ctx;
    ks->nprob = nprob;
    ks->periods = periods;
    ks->keep_pre = keep_pre != 0;
    ks->rad_res = args[0].rad_res;
    struct Guard {
        pkb_kset* k;
        ~Guard() { delete k; }
    } guard{ks};

    std::vector<DayParams>& hdp = ks->hdp;
    hdp.resize(nprob);
    std::vector<double> dpar(6 * (size_t)nprob), cells(2 * (size_t)nprob);
    for (int i = 0; i < nprob; ++i) {
        const pkb_day_args& a = args[i];
        if (a.rad_res != ks->rad_res) return fail(PKB_EINVAL, "pkb_kernels_build: all problems must share rad_res");
        if (a.rad_res < 1 || !(a.rad_dist > 0)) return fail(PKB_EINVAL, "pkb_kernels_build: bad domain (rad_dist %g, rad_res %d)", a.rad_dist, a.rad_res);
        if (a.kind != 0 && a.kind != 1) return fail(PKB_EINVAL, "pkb_kernels_build: kind must be 0 (day) or 1 (spread kernel)");
        if (!a.kind && a.n_periods < 1) return fail(PKB_EINVAL, "pkb_kernels_build: n_periods must be >= 1");
        if (!a.kind && (a.wind_day < 0 || a.wind_day >= nd_wind)) return fail(PKB_EINVAL, "pkb_kernels_build: wind_day %d out of range", a.wind_day);
        if (a.kind && !(a.sprd_factor >= 0 && a.sprd_factor <= 1)) return fail(PKB_EINVAL, "pkb_kernels_build: sprd_factor must be in [0, 1]");
        TRY(check_dparams(a.dparams, "Dparams"));
        TRY(check_dparams(a.dlparams, "Dlparams"));
        DayParams& d = hdp[i];
        d.lam = a.hparams[0]; d.aw = a.hparams[1]; d.bw = a.hparams[2];
        d.a1 = a.hparams[3]; d.b1 = a.hparams[4]; d.a2 = a.hparams[5]; d.b2 = a.hparams[6];
        d.mu_r = a.mu_r;
        d.cell = a.rad_dist / a.rad_res;                     // ParasitoidModel.py:411
        d.n_periods = a.n_periods;
        d.rad_res = a.rad_res;
        d.single = a.single ? 1 : 0;
        const int P = d.single ? 1 : periods;
        d.start_indx = 0;
        if (a.start_time >= 0) {
            const double s = floor(a.start_time * P);         // :431-433
            d.start_indx = s > P ? P : (int)s;
        }
        d.wind_day = a.wind_day;
        d.has_next = (a.wind_day + 1 < nd_wind) ? 1 : 0;
        d.bvn_S = 2 * i;
        d.bvn_Sl = 2 * i + 1;
        d.kind = a.kind;
        d.sprd_factor = a.sprd_factor;
        d.sdx = a.sprd_drift[0];
        d.sdy = a.sprd_drift[1];
        if (a.kind) { d.n_periods = 1; d.wind_day = 0; d.has_next = 0; d.single = 0; d.start_indx = 0; }
        for (int k = 0; k < 3; ++k) { dpar[6 * i + k] = a.dparams[k]; dpar[6 * i + 3 + k] = a.dlparams[k]; }
        cells[2 * i] = cells[2 * i + 1] = d.cell;
    }
    DBuf<double> d_dpar, d_cells;
    TRY(d_dpar.alloc(ctx, dpar.size()));
    TRY(d_cells.alloc(ctx, cells.size()));
    TRY(ks->bvn.alloc(ctx, 2 * (size_t)nprob));
    TRY(ks->ddp.alloc(ctx, nprob));
    TRY(ks->dmeta.alloc(ctx, nprob));
    TRY(ks->hprob.alloc(ctx, (size_t)nprob * periods));
    TRY(ks->loss_t.alloc(ctx, (size_t)nprob * periods));
    TRY(ks->pinfo.alloc(ctx, (size_t)nprob * periods));
    CU(cudaMemcpyAsync(d_dpar.p, dpar.data(), dpar.size() * sizeof(double), cudaMemcpyHostToDevice, ctx->stream));
    CU(cudaMemcpyAsync(d_cells.p, cells.data(), cells.size() * sizeof(double), cudaMemcpyHostToDevice, ctx->stream));
    CU(cudaMemcpyAsync(ks->ddp.p, hdp.data(), sizeof(DayParams) * nprob, cudaMemcpyHostToDevice, ctx->stream));
    CU(cudaMemsetAsync(ks->dmeta.p, 0, sizeof(DayMeta) * nprob, ctx->stream));
    CU(cudaMemsetAsync(ks->loss_t.p, 0, sizeof(double) * (size_t)nprob * periods, ctx->stream));
    CU(cudaMemsetAsync(ks->hprob.p, 0, sizeof(double) * (size_t)nprob * periods, ctx->stream));

    LAUNCH(ctx, k_bvn_setup, 2 * nprob, 32, 0, ks->bvn.p, d_dpar.p, d_cells.p, 0);
    if ((size_t)4 * periods * sizeof(double) > (size_t)ctx->max_smem) return fail(PKB_ELIMIT, "too many periods per day (%d)", periods);
    LAUNCH(ctx, k_hprob, nprob, 256, 4 * (size_t)periods * sizeof(double), ks->ddp.p, wind_dev, periods, ks->hprob.p, ks->dmeta.p,
           (double*)nullptr, (double*)nullptr);
    LAUNCH(ctx, k_drift, nprob, 256, 0, ks->ddp.p, ks->bvn.p, wind_dev, periods, ks->pinfo.p, ks->dmeta.p, ctx->ring_tol);

    // size of the accumulation window: needs the drift extents (one small D2H)
    ks->hmeta.resize(nprob);
    std::vector<BvnPar> hbvn(2 * (size_t)nprob);
    CU(cudaMemcpyAsync(ks->hmeta.data(), ks->dmeta.p, sizeof(DayMeta) * nprob, cudaMemcpyDeviceToHost, ctx->stream));
    CU(cudaMemcpyAsync(hbvn.data(), ks->bvn.p, sizeof(BvnPar) * 2 * nprob, cudaMemcpyDeviceToHost, ctx->stream));
    TRY(sync_check(ctx, "phase 1 drift pass"));
    int racc = 1, nmax = 4;
    for (int i = 0; i < nprob; ++i) {
        const int h0 = hbvn[2 * i].h0, hl = hbvn[2 * i + 1].h0;
        if (h0 < 0 || hl < 0) return fail(PKB_ELIMIT, "BVN support half-width exceeds 4096 cells (problem %d)", i);
        if (hl > ks->rad_res) return fail(PKB_ELIMIT, "local-diffusion blob (half-width %d) is larger than the domain (rad_res %d)", hl, ks->rad_res);
        racc = std::max(racc, std::min(ks->rad_res, std::max(ks->hmeta[i].ext, hl)));
        nmax = std::max(nmax, 2 * (h0 + 1) + 2);
    }
    if (nmax > PKB_LATTICE_CAP / 2) return fail(PKB_ELIMIT, "BVN support of %d cells per side exceeds the lattice tile", nmax);
    ks->racc = racc;
    ks->W = 2 * racc + 1;
    const size_t nel = (size_t)ks->W * ks->W;
    TRY(ks->acc.alloc(ctx, nel * nprob));
    TRY(ks->acc_lo.alloc(ctx, nel * nprob));
    CU(cudaMemsetAsync(ks->acc.p, 0, sizeof(double) * nel * nprob, ctx->stream));
    CU(cudaMemsetAsync(ks->acc_lo.p, 0, sizeof(double) * nel * nprob, ctx->stream));
    if (keep_pre) TRY(ks->pre.alloc(ctx, nel * nprob));

    // lattice tile: the whole (nmax x nmax) corner lattice when it fits, so that small supports leave room for more CTAs per SM
    int tile_target = PKB_TILE_TARGET;
    if (const char* env = getenv("PKB_TILE_TARGET")) tile_target = std::max(256, std::min(PKB_LATTICE_CAP, atoi(env)));      // tuning hook
    const int tile_cap = (int)std::min<size_t>(nmax > tile_target / 2 ? PKB_LATTICE_CAP : tile_target, (size_t)nmax * nmax);
    const size_t smem = (6 * (size_t)nmax + tile_cap) * sizeof(double);
    // CTA size: a period is a short, latency-bound job (set-up, lattice, differencing, with barriers in between), so
    // small supports get small CTAs and more of them per SM (C4, 48 x 48 lattice: 256 threads 1.83 ms, 64 threads 1.23 ms)
    const int lattice_items = nmax * ((nmax + PKB_BVN_SEG - 1) / PKB_BVN_SEG);
    const int period_threads = lattice_items <= 256 ? 64 : (lattice_items <= 1024 ? 128 : 256);
    LAUNCH(ctx, k_period, dim3(periods, nprob), period_threads, smem, ks->ddp.p, ks->bvn.p, ks->pinfo.p, ks->hprob.p, periods, nmax, tile_cap, ks->acc.p, ks->acc_lo.p,
           racc, ks->loss_t.p, ks->dmeta.p);
    TRY(launch_day_finalize(ctx, nprob, ks->ddp.p, ks->bvn.p, periods, ks->acc.p, ks->acc_lo.p, racc, ks->loss_t.p, ks->dmeta.p, 1e-8,
                            keep_pre ? ks->pre.p : (double*)nullptr, ks->pinfo.p));
    CU(cudaMemcpyAsync(ks->hmeta.data(), ks->dmeta.p, sizeof(DayMeta) * nprob, cudaMemcpyDeviceToHost, ctx->stream));
    TRY(sync_check(ctx, "phase 1"));
    guard.k = nullptr;
    *out = ks;
    return 0;
}

extern "C" int pkb_kernels_build(pkb_ctx* ctx, const double* wind, int nd_wind, int periods, const pkb_day_args* args, int nprob,
                                 int keep_pre, pkb_kset** out) {
    if (!ctx || !wind || !args || !out) return fail(PKB_EINVAL, "pkb_kernels_build: NULL argument");
    *out = nullptr;
    if (nd_wind < 1 || periods < 1) return fail(PKB_EINVAL, "pkb_kernels_build: empty wind array");
    CU(cudaSetDevice(ctx->device));
    DBuf<double> dw;
    const size_t nw = (size_t)nd_wind * periods * 3;
    TRY(dw.alloc(ctx, nw));
    CU(cudaMemcpyAsync(dw.p, wind, nw * sizeof(double), cudaMemcpyHostToDevice, ctx->stream));
    return kernels_build_dev(ctx, dw.p, nd_wind, periods, args, nprob, keep_pre, out);
}

extern "C" int pkb_kset_meta(pkb_kset* ks, int i, pkb_day_meta* out) {
    if (!ks || !out || i < 0 || i >= ks->nprob) return fail(PKB_EINVAL, "pkb_kset_meta: bad argument");
    memcpy(out, &ks->hmeta[i], sizeof(DayMeta));
    return 0;
}

extern "C" int pkb_kset_racc(pkb_kset* ks) { return ks ? ks->racc : -1; }

// copy the centred (2r+1)^2 sub-window of a W x W device window to the host
static int copy_window(pkb_ctx* ctx, const double* src, int W, int r, double* out) {
    const int c = W / 2, k = 2 * r + 1;
    for (int row = 0; row < k; ++row)
        CU(cudaMemcpyAsync(out + (size_t)row * k, src + (size_t)(c - r + row) * W + (c - r), sizeof(double) * k, cudaMemcpyDeviceToHost,
                           ctx->stream));
    return sync_check(ctx, "copy_window");
}

extern "C" int pkb_kset_get(pkb_kset* ks, int i, double* out) {
    if (!ks || !out || i < 0 || i >= ks->nprob) return fail(PKB_EINVAL, "pkb_kset_get: bad argument");
    CU(cudaSetDevice(ks->ctx->device));
    const size_t nel = (size_t)ks->W * ks->W;
    return copy_window(ks->ctx, ks->acc.p + nel * i, ks->W, ks->hmeta[i].rad, out);
}

extern "C" int pkb_kset_get_pre(pkb_kset* ks, int i, double* out) {
    if (!ks || !out || i < 0 || i >= ks->nprob) return fail(PKB_EINVAL, "pkb_kset_get_pre: bad argument");
    if (!ks->keep_pre) return fail(PKB_ESTATE, "pkb_kset_get_pre: kernel set was built without keep_pre");
    CU(cudaSetDevice(ks->ctx->device));
    const size_t nel = (size_t)ks->W * ks->W;
    CU(cudaMemcpyAsync(out, ks->pre.p + nel * i, sizeof(double) * nel, cudaMemcpyDeviceToHost, ks->ctx->stream));
    return sync_check(ks->ctx, "pkb_kset_get_pre");
}

extern "C" int pkb_kset_periods(pkb_kset* ks, int i, int* rch, double* hprob) {
    if (!ks || i < 0 || i >= ks->nprob) return fail(PKB_EINVAL, "pkb_kset_periods: bad argument");
    pkb_ctx* ctx = ks->ctx;
    CU(cudaSetDevice(ctx->device));
    const int P = ks->periods;
    if (hprob) CU(cudaMemcpyAsync(hprob, ks->hprob.p + (size_t)i * P, sizeof(double) * P, cudaMemcpyDeviceToHost, ctx->stream));
    std::vector<PeriodInfo> pi(P);
    if (rch) CU(cudaMemcpyAsync(pi.data(), ks->pinfo.p + (size_t)i * P, sizeof(PeriodInfo) * P, cudaMemcpyDeviceToHost, ctx->stream));
    TRY(sync_check(ctx, "pkb_kset_periods"));
    if (rch) {
        const DayParams& d = ks->hdp[i];
        const int np = d.single ? 1 : P;
        for (int t = 0; t < P; ++t) {
            const bool live = t >= d.start_indx && t < np;
            rch[3 * t] = live ? pi[t].row_c : -1;
            rch[3 * t + 1] = live ? pi[t].col_c : -1;
            rch[3 * t + 2] = live ? pi[t].h : -1;
        }
    }
    return 0;
}

extern "C" int pkb_kset_destroy(pkb_kset* ks) {
    if (!ks) return 0;
    cudaSetDevice(ks->ctx->device);
    cudaStreamSynchronize(ks->ctx->stream);
    delete ks;
    return 0;
}

extern "C" int pkb_wind_interp(pkb_ctx* ctx, const double* raw, int nd, int npts, int interp_num, int half_hour_start, double* out) {
    if (!ctx || !raw || !out) return fail(PKB_EINVAL, "pkb_wind_interp: NULL argument");
    if (nd < 1 || npts < 1 || interp_num < 1) return fail(PKB_EINVAL, "pkb_wind_interp: bad sizes");
    CU(cudaSetDevice(ctx->device));
    DBuf<double> din, dout;
    const size_t nin = (size_t)nd * npts * 3, nout = nin * interp_num;
    TRY(din.alloc(ctx, nin));
    TRY(dout.alloc(ctx, nout));
    CU(cudaMemcpyAsync(din.p, raw, nin * sizeof(double), cudaMemcpyHostToDevice, ctx->stream));
    const int periods = npts * interp_num;
    LAUNCH(ctx, k_wind_interp, dim3((periods + 255) / 256, nd), 256, 0, (const double*)din.p, nd, npts, interp_num, half_hour_start ? 1 : 0, dout.p);
    CU(cudaMemcpyAsync(out, dout.p, nout * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
    return sync_check(ctx, "pkb_wind_interp");
}

// h_flight_prob alone: a one-problem hprob launch
extern "C" int pkb_hprob(pkb_ctx* ctx, const double* wind, int periods, int single, const double hparams[7], double* out,
                         double* f_out, double* g_out) {
    if (!ctx || !wind || !hparams || !out) return fail(PKB_EINVAL, "pkb_hprob: NULL argument");
    if (periods < 1) return fail(PKB_EINVAL, "pkb_hprob: periods must be >= 1");
    CU(cudaSetDevice(ctx->device));
    if (single) periods = 1;
    if ((size_t)4 * periods * sizeof(double) > (size_t)ctx->max_smem) return fail(PKB_ELIMIT, "too many periods per day (%d)", periods);
    DayParams d;
    memset(&d, 0, sizeof d);
    d.lam = hparams[0]; d.aw = hparams[1]; d.bw = hparams[2];
    d.a1 = hparams[3]; d.b1 = hparams[4]; d.a2 = hparams[5]; d.b2 = hparams[6];
    d.single = single ? 1 : 0;
    d.start_indx = periods;   // no range check here: prob_mass asserts, h_flight_prob does not
    DBuf<double> dw, dh, df, dg;
    DBuf<DayParams> dd;
    DBuf<DayMeta> dm;
    TRY(dw.alloc(ctx, (size_t)periods * 3));
    TRY(dh.alloc(ctx, periods));
    TRY(df.alloc(ctx, periods));
    TRY(dg.alloc(ctx, periods));
    TRY(dd.alloc(ctx, 1));
    TRY(dm.alloc(ctx, 1));
    CU(cudaMemcpyAsync(dw.p, wind, sizeof(double) * 3 * periods, cudaMemcpyHostToDevice, ctx->stream));
    CU(cudaMemcpyAsync(dd.p, &d, sizeof d, cudaMemcpyHostToDevice, ctx->stream));
    CU(cudaMemsetAsync(dm.p, 0, sizeof(DayMeta), ctx->stream));
    LAUNCH(ctx, k_hprob, 1, 256, 4 * (size_t)periods * sizeof(double), dd.p, dw.p, periods, dh.p, dm.p, df.p, dg.p);
    CU(cudaMemcpyAsync(out, dh.p, sizeof(double) * periods, cudaMemcpyDeviceToHost, ctx->stream));
    if (f_out) CU(cudaMemcpyAsync(f_out, df.p, sizeof(double) * periods, cudaMemcpyDeviceToHost, ctx->stream));
    if (g_out) CU(cudaMemcpyAsync(g_out, dg.p, sizeof(double) * periods, cudaMemcpyDeviceToHost, ctx->stream));
    return sync_check(ctx, "pkb_hprob");
}

extern "C" int pkb_mvn_cdf(pkb_ctx* ctx, double cell_length, const double mu[2], const double cov[3], double* out, int cap, int* h_out) {
    if (!ctx || !mu || !cov || !out || !h_out) return fail(PKB_EINVAL, "pkb_mvn_cdf: NULL argument");
    if (!(cell_length > 0) || !(cov[0] > 0) || !(cov[1] > 0)) return fail(PKB_EINVAL, "pkb_mvn_cdf: cell length and variances must be positive");
    if (cap < 1) return fail(PKB_EINVAL, "pkb_mvn_cdf: cap must be positive");
    CU(cudaSetDevice(ctx->device));
    DBuf<BvnPar> bp;
    DBuf<double> dpar, dcell, dout;
    DBuf<int> dh;
    TRY(bp.alloc(ctx, 1));
    TRY(dpar.alloc(ctx, 3));
    TRY(dcell.alloc(ctx, 1));
    TRY(dout.alloc(ctx, cap));
    TRY(dh.alloc(ctx, 1));
    CU(cudaMemcpyAsync(dpar.p, cov, sizeof(double) * 3, cudaMemcpyHostToDevice, ctx->stream));
    CU(cudaMemcpyAsync(dcell.p, &cell_length, sizeof(double), cudaMemcpyHostToDevice, ctx->stream));
    LAUNCH(ctx, k_bvn_setup, 1, 32, 0, bp.p, dpar.p, dcell.p, 1);
    LAUNCH(ctx, k_mvn_cdf, 1, 256, 0, bp.p, cell_length, mu[0], mu[1], dout.p, cap, dh.p, ctx->ring_tol);
    int h = -1;
    CU(cudaMemcpyAsync(&h, dh.p, sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
    TRY(sync_check(ctx, "pkb_mvn_cdf"));
    if (h < 0) return fail(PKB_ELIMIT, "pkb_mvn_cdf: support exceeds the output capacity (%d values)", cap);
    const size_t n = (size_t)(2 * h + 1) * (2 * h + 1);
    CU(cudaMemcpyAsync(out, dout.p, sizeof(double) * n, cudaMemcpyDeviceToHost, ctx->stream));
    TRY(sync_check(ctx, "pkb_mvn_cdf"));
    *h_out = h;
    return 0;
}

// ---------------------------------------------------------------------------
// phase 2: the chain
// ---------------------------------------------------------------------------
#define PKB_MAX_COHORTS 16
#ifndef PKB_EMIT_T
#define PKB_EMIT_T 256     // threads of the side-stream emission CTAs
#endif

struct pkb_chain {
    pkb_ctx* ctx;
    ChainDims d;
    int mmax;
    FftPlan plan;
    DBuf<double> S[2];
    int cur;
    DBuf<cplx> Yt, Wt, Krt;
    DBuf<cplx> Krt_t;       // kernel row spectra on the truncated-source torus (TruncGeom), when a caller brings none
    DBuf<cplx> cscr;        // k_cols: per-CTA parking space for the filter column spectrum
    DBuf<cplx> Shat;        // spectral-resident state (ChainCtrl::spec): Nc slices of hstride complex, allocated on first use
    size_t hstride;
    bool fixed_torus;       // every whole-torus step runs on the chain's own torus (spectral-resident steps need one torus)
    DBuf<int> done;         // k_rows_inv: CTAs finished (the last one finalises the step)
    DBuf<int> colflag;      // [P] columns in which a support-window step saw a cell >= PKB_SPEC_TAU (cleared by its finalize)
    StepMeta* meta_main;    // fused solve: where the main chain's next step writes its StepMeta (the per-day array; saves a copy per step)
    int* host_box;          // tau windows: mapped pinned host slot the next window step reports its extent to (+ ticket), or NULL
    int host_ticket;
    size_t cscr_per_cta;
    int grid_rows, grid_cols;
    DBuf<RowStats> rstat;
    DBuf<ChainCtrl> ctrl;   // [0] main state, [1 + j] cohort j
    DBuf<StepMeta> meta;    // same indexing
    DBuf<double> dout;      // D*D staging
    DBuf<double> kup;       // uploaded filter window, (2*mmax+1)^2
    DBuf<double> coh[PKB_MAX_COHORTS];
    DBuf<cplx> kcache[PKB_MAX_COHORTS];   // cached row spectra of the release-day filters (fused solve)
    DBuf<cplx> kcache_t[PKB_MAX_COHORTS]; // the same on each filter's truncated-source torus (trunc_torus)
    int kcache_m[PKB_MAX_COHORTS];
    double negval;          // threshold the row statistics were taken with
    double flag_thresh;     // boundary-flag threshold of the last finalize (1e-8, or negval on the cuda_lib.get_cursol path)
    bool stats_valid;
};

static int roundup(int v, int q) { return (v + q - 1) / q * q; }

static int chain_create(pkb_ctx* ctx, int D, int mmax, pkb_chain** out) {
    if (D < 1 || mmax < 0) return fail(PKB_EINVAL, "pkb_chain_create: bad sizes (dom_len %d, filter radius %d)", D, mmax);
    pkb_chain* ch = new pkb_chain();
    struct Guard {
        pkb_chain* c;
        ~Guard() { delete c; }
    } guard{ch};
    ch->ctx = ctx;
    ch->mmax = mmax;
    ChainDims& d = ch->d;
    d.D = D;
    d.P = D + mmax;                         // CalcSol.py:20-21 with max_shape = 2*mmax + 1
    d.N = pkb_smooth_len(std::max(2, d.P + 2 * mmax));
    d.Nc = d.N / 2 + 1;
    d.ldS = roundup(d.P, 16);
    d.ldY = roundup(d.P, 2);
    d.ldW = roundup(d.N, 2);
    d.ldK = roundup(2 * mmax + 1, 2);
    d.win = d.wr0 = d.wc0 = d.wn = 0;
    TRY(get_plan(ctx, d.N, &ch->plan));
    if (fft_smem_bytes(ch->plan) > (size_t)ctx->max_smem)
        return fail(PKB_ELIMIT, "torus side %d (domain %d + filter radius %d) exceeds the shared-memory FFT limit of %d points", d.N, D,
                    mmax, (int)(ctx->max_smem / sizeof(cplx)));
    if (ch->plan.grid_rows < 1 || ch->plan.grid_cols < 1) return fail(PKB_ELIMIT, "FFT kernels cannot be resident at torus side %d", d.N);
    ch->grid_rows = ch->plan.grid_rows;
    ch->grid_cols = ch->plan.grid_cols;
    // k_cols parking space: any plan up to this torus side needs at most N + R_last * threads slots per CTA
    ch->cscr_per_cta = (size_t)d.N + 21 * 256 + 256;
    TRY(ch->cscr.alloc(ctx, (size_t)ctx->occ_cap * ctx->sm_count * ch->cscr_per_cta));
    const size_t ns = (size_t)d.P * d.ldS;
    TRY(ch->S[0].alloc(ctx, ns));
    TRY(ch->S[1].alloc(ctx, ns));
    TRY(ch->Yt.alloc(ctx, spec_size(d.Nc + 1, d.ldY)));
    TRY(ch->Wt.alloc(ctx, spec_size(d.Nc + 1, d.ldW)));
    TRY(ch->Krt.alloc(ctx, spec_size(d.Nc + 1, d.ldK)));
    TRY(ch->rstat.alloc(ctx, d.P));
    TRY(ch->done.alloc(ctx, 1));
    CU(cudaMemsetAsync(ch->done.p, 0, sizeof(int), ctx->stream));
    TRY(ch->colflag.alloc(ctx, d.P));
    CU(cudaMemsetAsync(ch->colflag.p, 0, sizeof(int) * d.P, ctx->stream));
    CU(cudaMemsetAsync(ch->rstat.p, 0, sizeof(RowStats) * d.P, ctx->stream));    // rows a windowed step never touches are zero rows
    TRY(ch->ctrl.alloc(ctx, 1 + PKB_MAX_COHORTS));
    TRY(ch->meta.alloc(ctx, 1 + PKB_MAX_COHORTS));
    TRY(ch->dout.alloc(ctx, (size_t)D * D));
    TRY(ch->kup.alloc(ctx, (size_t)(2 * mmax + 1) * (2 * mmax + 1)));
    CU(cudaMemsetAsync(ch->S[0].p, 0, ns * sizeof(double), ctx->stream));
    CU(cudaMemsetAsync(ch->ctrl.p, 0, sizeof(ChainCtrl) * (1 + PKB_MAX_COHORTS), ctx->stream));
    CU(cudaMemsetAsync(ch->meta.p, 0, sizeof(StepMeta) * (1 + PKB_MAX_COHORTS), ctx->stream));
    ch->cur = 0;
    ch->negval = 1e-8;
    ch->stats_valid = false;
    ch->flag_thresh = 1e-8;
    ch->hstride = 0;
    ch->fixed_torus = false;
    ch->host_box = nullptr;
    ch->host_ticket = 0;
    ch->meta_main = nullptr;
    for (int i = 0; i < PKB_MAX_COHORTS; ++i) ch->kcache_m[i] = -1;
    guard.c = nullptr;
    *out = ch;
    return 0;
}

extern "C" int pkb_chain_create(pkb_ctx* ctx, int dom_len, int max_shape, pkb_chain** out) {
    if (!ctx || !out) return fail(PKB_EINVAL, "pkb_chain_create: NULL argument");
    *out = nullptr;
    if (max_shape < 0) return fail(PKB_EINVAL, "pkb_chain_create: max_shape must be >= 0");
    CU(cudaSetDevice(ctx->device));
    return chain_create(ctx, dom_len, max_shape / 2, out);
}

extern "C" int pkb_chain_destroy(pkb_chain* ch) {
    if (!ch) return 0;
    cudaSetDevice(ch->ctx->device);
    cudaStreamSynchronize(ch->ctx->stream);
    delete ch;
    return 0;
}

extern "C" int pkb_chain_dims(pkb_chain* ch, int* D, int* P, int* N) {
    if (!ch) return fail(PKB_EINVAL, "pkb_chain_dims: NULL chain");
    if (D) *D = ch->d.D;
    if (P) *P = ch->d.P;
    if (N) *N = ch->d.N;
    return 0;
}

// One convolution step: dst = src (*) K on the P torus, followed by the step's
// flag / sums -> ctrl[slot], meta[slot] (apply_trunc: mark a flagged state as
// truncated to the domain, CalcSol.py:200-201; readers of a state honour
// ctrl->trunc, so the pad is only physically zeroed where a caller can see it).
// K is a device window Wk x Wk with support radius m.  krt: row spectra buffer to
// (re)use; krt_ready: it already holds the spectra of K.
// Torus of a whole-torus step with a filter of radius m.  The linear convolution of the P x P state
// with a (2m+1)^2 kernel needs N >= P + 2m for THIS m, not for the chain's largest kernel: days
// with a small kernel run on a smaller 7-smooth torus (same fold mod P afterwards, same result to
// rounding).  Falls back to the chain's torus when the smaller plan cannot use the chain's buffers.
static int step_torus(pkb_chain* ch, int m, ChainDims* d, FftPlan* plan) {
    pkb_ctx* ctx = ch->ctx;
    *d = ch->d;
    *plan = ch->plan;
    if (!ctx->use_step_torus || ch->fixed_torus) return 0;
    const int Nd = pkb_smooth_len(std::max(2, ch->d.P + 2 * m));
    if (Nd >= ch->d.N) return 0;
    FftPlan p;
    TRY(get_plan(ctx, Nd, &p));
    if (p.grid_rows < 1 || p.grid_cols < 1) return 0;
    if ((size_t)p.cols_kb * plan_radix(p, p.nstage - 1) * p.cols_threads > ch->cscr_per_cta) return 0;
    d->N = Nd;
    d->Nc = Nd / 2 + 1;
    d->ldW = (Nd + 1) / 2 * 2;
    *plan = p;
    return 0;
}

// Geometry for a truncated source (TruncGeom in chain.cuh) next to the step's own (d, plan): torus
// >= D + 2m, used by the kernels when the source state turns out truncated.  tg->N == 0: none.
static int trunc_torus(pkb_chain* ch, int m, const ChainDims& d, const FftPlan& plan, TruncGeom* tg, FftPlan* plan_t) {
    pkb_ctx* ctx = ch->ctx;
    memset(tg, 0, sizeof *tg);
    *plan_t = plan;
    if (!ctx->use_trunc_torus) return 0;
    const int Nt = pkb_smooth_len(std::max(2, d.D + 2 * m));
    if (Nt >= d.N) return 0;
    FftPlan p;
    TRY(get_plan(ctx, Nt, &p));
    if (p.grid_rows < 1 || p.grid_cols < 1) return 0;
    // the launch is configured for the step's own plan: CTA sizes, shared memory and column scratch must cover this one too
    const int RL = plan_radix(p, p.nstage - 1);
    const int kb = (Nt / RL + plan.cols_threads - 1) / plan.cols_threads;
    if ((size_t)kb * RL * plan.cols_threads > ch->cscr_per_cta) return 0;
    tg->N = Nt;
    tg->Nc = Nt / 2 + 1;
    tg->ldW = (Nt + 1) / 2 * 2;
    tg->cols_kb = kb;
    *plan_t = p;
    return 0;
}
// ChainDims of the truncated-source torus (kernel row spectra are built with these)
static ChainDims trunc_dims(const ChainDims& d, const TruncGeom& tg) {
    ChainDims t = d;
    t.N = tg.N; t.Nc = tg.Nc; t.ldW = tg.ldW;
    return t;
}

static int conv_step(pkb_chain* ch, const double* src, const ChainCtrl* src_ctrl, double* dst, const double* K, int Wk, int m,
                     cplx* krt, bool krt_ready, int slot, int apply_trunc, const int* win = nullptr, bool fuse_next = false,
                     int pre_m = -1, bool allow_trunc = false, cplx* krt_t = nullptr, bool krt_t_ready = true, int spec_try = 0,
                     cudaEvent_t krt_event = nullptr) {
    // krt_event: the kernel row spectra are being computed on another stream; the column pass waits for this event (the
    // forward row pass does not need them and starts at once)
    // spec_try: 0 exact steps only; 1 spectral-resident steps allowed; 2 ... with row windows (ChainCtrl::er0, probability model)
    pkb_ctx* ctx = ch->ctx;
    if (m > ch->mmax) return fail(PKB_ELIMIT, "filter radius %d exceeds the chain's max_shape//2 = %d", m, ch->mmax);
    if (2 * m > ch->d.P) return fail(PKB_ELIMIT, "filter radius %d does not fit the %d-cell padded domain", m, ch->d.P);
    StepMeta* meta_dst = (slot == 0 && ch->meta_main) ? ch->meta_main : ch->meta.p + slot;
    if (m <= ctx->stencil_max_radius) {
        const ChainDims& d = ch->d;
        const size_t smem = ((size_t)(8 + 2 * m) * (32 + 2 * m) + (size_t)(2 * m + 1) * (2 * m + 1)) * sizeof(double);
        LAUNCH(ctx, k_stencil, dim3((d.P + 31) / 32, (d.P + 7) / 8), dim3(32, 8), smem, src, K, Wk, m, d, src_ctrl, dst);
        LAUNCH(ctx, k_row_stats, d.P, 256, 0, (const double*)dst, d, ch->rstat.p, ch->negval);
        LAUNCH(ctx, k_step_finalize, 1, 256, 0, (const RowStats*)ch->rstat.p, d, ch->ctrl.p + slot, meta_dst, apply_trunc, 1e-8);
        return 0;
    }
    // geometry of this step: the chain's torus, or a smaller one around the state's support window
    ChainDims d;
    FftPlan plan;
    TRY(step_torus(ch, m, &d, &plan));
    if (win) {
        d = ch->d;
        d.win = 1; d.wr0 = win[0]; d.wc0 = win[1]; d.wn = win[2];
        d.N = pkb_smooth_len(std::max(2, d.wn + 2 * m));
        d.Nc = d.N / 2 + 1;
        d.ldY = roundup(d.wn, 2);
        d.ldW = roundup(d.N, 2);
        d.ldK = roundup(2 * m + 1, 2);
        TRY(get_plan(ctx, d.N, &plan));
        if (plan.grid_rows < 1 || plan.grid_cols < 1) return fail(PKB_ELIMIT, "FFT kernels cannot be resident at torus side %d", d.N);
        if (!krt_ready) krt = ch->Krt.p;     // (spectra prepared for another torus do not apply: callers pass per-window ones)
    }
    // second geometry for a truncated source (main chain steps only)
    TruncGeom tg;
    FftPlan plan_t;
    memset(&tg, 0, sizeof tg);
    plan_t = plan;
    if (!win && allow_trunc) TRY(trunc_torus(ch, m, d, plan, &tg, &plan_t));
    // persistent grids: (resident CTAs per SM) x (SM count), capped by the job count
    const int T = plan.threads;
    const size_t sm1 = std::max(fft_smem_bytes(plan), fft_smem_bytes(plan_t));
    const int rows_in = win ? d.wn : d.P;
    if ((size_t)plan.cols_kb * plan_radix(plan, plan.nstage - 1) * plan.cols_threads > ch->cscr_per_cta)
        return fail(PKB_ELIMIT, "column scratch too small for torus side %d", d.N);
    const int njobs = win ? (d.wn + 2 * m + 1) / 2 : std::max(rows_inv_jobs(d.P, m), tg.N ? rows_inv_jobs_trunc(d.P, d.D, m) : 0);
    // spectral-resident steps: main chain, the chain's own torus (ChainCtrl::spec; the kernels decide on the device)
    cplx* shat = nullptr;
    if (spec_try && !win && slot == 0 && d.N == ch->d.N) {
        const size_t hs = (size_t)plan.cols_kb * plan_radix(plan, plan.nstage - 1) * plan.cols_threads;
        if (!ch->Shat.p || ch->hstride != hs) {
            ch->hstride = hs;
            TRY(ch->Shat.alloc(ctx, hs * d.Nc));
        }
        shat = ch->Shat.p;
    }
    ChainCtrl* src_ctrl_w = const_cast<ChainCtrl*>(src_ctrl);      // (k_cols leaves its `stored` message there)
    const int rowwin_ok = (shat && spec_try >= 2 && ctx->use_rowwin) ? 1 : 0;
    if (win) {
        if (!krt_ready) LAUNCH_AS(ctx, "k_kernel_rows_win", k_kernel_rows, std::min(m + 1, plan.grid_rows), T, sm1, K, Wk, m, d, krt, plan);
        LAUNCH_AS(ctx, "k_rows_fwd_win", k_rows_fwd, std::min((rows_in + 1) / 2, plan.grid_rows), T, sm1, src, d, src_ctrl, ch->Yt.p, plan, -1, tg, plan_t, 0);
        if (krt_event) CU(cudaStreamWaitEvent(ctx->stream, krt_event, 0));
        LAUNCH_AS(ctx, "k_cols_win", k_cols, std::min(d.Nc, plan.grid_cols), plan.cols_threads, sm1, (const cplx*)ch->Yt.p, (const cplx*)krt,
                  m, d, src_ctrl_w, ch->Wt.p, ch->cscr.p, plan, tg, plan_t, (const cplx*)krt, (cplx*)nullptr, (size_t)0, 0);
        LAUNCH_AS(ctx, "k_rows_inv_win", k_rows_inv, std::min(njobs, plan.grid_rows), T, sm1, (const cplx*)ch->Wt.p, m, d, dst, ch->rstat.p,
                  ch->negval, plan, ch->done.p, ch->ctrl.p + slot, meta_dst, apply_trunc, (cplx*)nullptr, src_ctrl, tg, plan_t, ctx->rows_desc, 0, ch->colflag.p,
                  ch->host_box, ch->host_ticket);
        return 0;
    }
    if (!krt_ready) LAUNCH(ctx, k_kernel_rows, std::min(m + 1, plan.grid_rows), T, sm1, K, Wk, m, d, krt, plan);
    if (tg.N && !krt_t) {
        // no spectra for the truncated-source torus from the caller: built here (whether they are needed is only known on the device)
        if (!ch->Krt_t.p) TRY(ch->Krt_t.alloc(ctx, spec_size(ch->d.Nc + 1, ch->d.ldK)));
        krt_t = ch->Krt_t.p;
        krt_t_ready = false;
    }
    if (tg.N && !krt_t_ready)
        LAUNCH(ctx, k_kernel_rows, std::min(m + 1, plan_t.grid_rows), plan_t.threads, fft_smem_bytes(plan_t), K, Wk, m, trunc_dims(d, tg), krt_t, plan_t);
    if (!tg.N) krt_t = krt;
    LAUNCH(ctx, k_rows_fwd, std::min((rows_in + 1) / 2, plan.grid_rows), T, sm1, src, d, src_ctrl, ch->Yt.p, plan, pre_m, tg, plan_t, shat ? 1 : 0);
    LAUNCH(ctx, k_cols, std::min(d.Nc, plan.grid_cols), plan.cols_threads, sm1, (const cplx*)ch->Yt.p, (const cplx*)krt, m, d, src_ctrl_w,
           ch->Wt.p, ch->cscr.p, plan, tg, plan_t, (const cplx*)krt_t, shat, ch->hstride, rowwin_ok);
    LAUNCH(ctx, k_rows_inv, std::min(njobs, plan.grid_rows), T, sm1, (const cplx*)ch->Wt.p, m, d, dst, ch->rstat.p, ch->negval, plan,
           ch->done.p, ch->ctrl.p + slot, meta_dst, apply_trunc, fuse_next ? ch->Yt.p : (cplx*)nullptr, src_ctrl, tg, plan_t, ctx->rows_desc,
           rowwin_ok, (int*)nullptr, (int*)nullptr, 0);
    return 0;
}

// flag / sums of a state whose row statistics are in ch->rstat -> ctrl[0], meta[0] (re-thresholding)
static void finalize_state(pkb_chain* ch, double flag_thresh) {
    pkb_ctx* ctx = ch->ctx;
    LAUNCH(ctx, k_step_finalize, 1, 256, 0, (const RowStats*)ch->rstat.p, ch->d, ch->ctrl.p, ch->meta.p, 0, flag_thresh);
}

extern "C" int pkb_chain_set_state(pkb_chain* ch, const double* A) {
    if (!ch || !A) return fail(PKB_EINVAL, "pkb_chain_set_state: NULL argument");
    pkb_ctx* ctx = ch->ctx;
    const ChainDims& d = ch->d;
    CU(cudaSetDevice(ctx->device));
    double* S = ch->S[ch->cur].p;
    CU(cudaMemsetAsync(S, 0, (size_t)d.P * d.ldS * sizeof(double), ctx->stream));
    CU(cudaMemcpyAsync(ch->dout.p, A, (size_t)d.D * d.D * sizeof(double), cudaMemcpyHostToDevice, ctx->stream));
    LAUNCH(ctx, k_load_state, d.D, 256, 0, (const double*)ch->dout.p, d, S);
    LAUNCH(ctx, k_set_ctrl, 1, 32, 0, ch->ctrl.p, 1, 0);
    ch->stats_valid = false;
    return sync_check(ctx, "pkb_chain_set_state");
}

static int set_state_kernel_dev(pkb_chain* ch, const double* K, int Wk, int m) {
    pkb_ctx* ctx = ch->ctx;
    const ChainDims& d = ch->d;
    if (m > d.D / 2) return fail(PKB_ELIMIT, "kernel radius %d is larger than the domain radius %d", m, d.D / 2);
    double* S = ch->S[ch->cur].p;
    CU(cudaMemsetAsync(S, 0, (size_t)d.P * d.ldS * sizeof(double), ctx->stream));
    LAUNCH(ctx, k_place_kernel, 2 * m + 1, 128, 0, K, Wk, m, d, S);
    LAUNCH(ctx, k_set_ctrl, 1, 32, 0, ch->ctrl.p, 1, 0);
    ch->stats_valid = false;
    return 0;
}

extern "C" int pkb_chain_set_state_kernel(pkb_chain* ch, pkb_kset* ks, int i) {
    if (!ch || !ks || i < 0 || i >= ks->nprob) return fail(PKB_EINVAL, "pkb_chain_set_state_kernel: bad argument");
    CU(cudaSetDevice(ch->ctx->device));
    if (2 * ks->rad_res + 1 != ch->d.D) return fail(PKB_EINVAL, "kernel set domain (%d) does not match the chain (%d)", 2 * ks->rad_res + 1, ch->d.D);
    TRY(set_state_kernel_dev(ch, ks->acc.p + (size_t)ks->W * ks->W * i, ks->W, ks->hmeta[i].rad));
    return sync_check(ch->ctx, "pkb_chain_set_state_kernel");
}

// upload the support window of a host filter (k x k, odd) into ch->kup; returns its radius
static int upload_filter(pkb_chain* ch, const double* B, int k, int* m_out) {
    pkb_ctx* ctx = ch->ctx;
    if (k < 1 || !(k & 1)) return fail(PKB_EINVAL, "filters must be square with an odd side (got %d)", k);
    const int c = k / 2;
    int m = 0;
    for (int r = 0; r < k; ++r)
        for (int q = 0; q < k; ++q)
            if (B[(size_t)r * k + q] != 0.0) m = std::max(m, std::max(std::abs(r - c), std::abs(q - c)));
    if (m > ch->mmax) return fail(PKB_ELIMIT, "filter support radius %d exceeds the chain's max_shape//2 = %d", m, ch->mmax);
    const int w = 2 * m + 1;
    std::vector<double> win((size_t)w * w);
    for (int r = 0; r < w; ++r) memcpy(&win[(size_t)r * w], B + (size_t)(c - m + r) * k + (c - m), sizeof(double) * w);
    CU(cudaMemcpyAsync(ch->kup.p, win.data(), win.size() * sizeof(double), cudaMemcpyHostToDevice, ctx->stream));
    CU(cudaStreamSynchronize(ctx->stream));   // `win` is pageable and about to go out of scope
    *m_out = m;
    return 0;
}

static int chain_conv_main(pkb_chain* ch, const double* K, int Wk, int m, int apply_trunc, cplx* krt = nullptr, const int* win = nullptr,
                           bool fuse_next = false, int pre_m = -1, cplx* krt_t = nullptr, int spec_try = 0, cudaEvent_t krt_event = nullptr) {
    const int nxt = ch->cur ^ 1;
    TRY(conv_step(ch, ch->S[ch->cur].p, ch->ctrl.p, ch->S[nxt].p, K, Wk, m, krt ? krt : ch->Krt.p, krt != nullptr, 0, apply_trunc, win,
                  fuse_next, pre_m, true, krt_t, true, spec_try, krt_event));
    ch->cur = nxt;
    return 0;
}

extern "C" int pkb_chain_conv(pkb_chain* ch, const double* B, int k) {
    if (!ch || !B) return fail(PKB_EINVAL, "pkb_chain_conv: NULL argument");
    CU(cudaSetDevice(ch->ctx->device));
    int m = 0;
    TRY(upload_filter(ch, B, k, &m));
    TRY(chain_conv_main(ch, ch->kup.p, 2 * m + 1, m, 0));
    ch->stats_valid = true;
    ch->flag_thresh = 1e-8;
    return sync_check(ch->ctx, "pkb_chain_conv");
}

extern "C" int pkb_chain_conv_kernel(pkb_chain* ch, pkb_kset* ks, int i) {
    if (!ch || !ks || i < 0 || i >= ks->nprob) return fail(PKB_EINVAL, "pkb_chain_conv_kernel: bad argument");
    CU(cudaSetDevice(ch->ctx->device));
    TRY(chain_conv_main(ch, ks->acc.p + (size_t)ks->W * ks->W * i, ks->W, ks->hmeta[i].rad, 0));
    ch->stats_valid = true;
    ch->flag_thresh = 1e-8;
    return sync_check(ch->ctx, "pkb_chain_conv_kernel");
}

extern "C" int pkb_chain_get_cursol(pkb_chain* ch, double negval, int mode, int apply_trunc, double* out, pkb_step_meta* meta) {
    if (!ch) return fail(PKB_EINVAL, "pkb_chain_get_cursol: NULL chain");
    if (mode < 0 || mode > 2) return fail(PKB_EINVAL, "pkb_chain_get_cursol: mode must be 0, 1 or 2");
    pkb_ctx* ctx = ch->ctx;
    const ChainDims& d = ch->d;
    CU(cudaSetDevice(ctx->device));
    double* S = ch->S[ch->cur].p;
    // cuda_lib.get_cursol (mode 1) keeps v > negval and raises the flag when anything outside the domain survives that
    // threshold (cuda_lib.py:117-130); the CPU path's ifft2 (modes 0, 2) compares with 1e-8 (CalcSol.py:36-37)
    const double thresh = mode == 1 ? negval : 1e-8;
    if (!ch->stats_valid || negval != ch->negval || thresh != ch->flag_thresh) {
        ch->negval = negval;
        ch->flag_thresh = thresh;
        LAUNCH(ctx, k_row_stats, d.P, 256, 0, (const double*)S, d, ch->rstat.p, negval);
        finalize_state(ch, thresh);
        ch->stats_valid = true;
    }
    if (out) {
        if (mode == 0) LAUNCH(ctx, k_copy_domain, d.D, 256, 0, (const double*)S, d, ch->dout.p, (const StepMeta*)nullptr);
        else LAUNCH(ctx, k_emit_dense, d.D, 256, 0, (const double*)S, d, (const StepMeta*)ch->meta.p, negval, mode == 2 ? 1 : 0, mode == 1 ? 1 : 0, ch->dout.p, (int*)nullptr, 0);
        CU(cudaMemcpyAsync(out, ch->dout.p, (size_t)d.D * d.D * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
    }
    if (apply_trunc) {
        LAUNCH(ctx, k_apply_trunc, 1, 32, 0, ch->ctrl.p);
        LAUNCH(ctx, k_zero_pad, d.P, 256, 0, S, d, (const ChainCtrl*)ch->ctrl.p);
    }
    StepMeta hm;
    CU(cudaMemcpyAsync(&hm, ch->meta.p, sizeof hm, cudaMemcpyDeviceToHost, ctx->stream));
    TRY(sync_check(ctx, "pkb_chain_get_cursol"));
    if (meta) memcpy(meta, &hm, sizeof hm);
    return 0;
}

// cohorts of earlier release days: cohort j = state (*) F[nf-1] (*) ... (*) F[j]
// F[j]: device window (Wk[j], radius m[j]); krt[j]/ready[j]: optional cached spectra.
// src0 / ctrl0 (optional): the state to start from and its control block, when it is not this chain's own current state
// (cohort lanes of the fused solve: the scratch, cohort buffers and stream of `ch`, the state of the main chain)
static int back_solve_dev(pkb_chain* ch, const double* const* F, const int* Wk, const int* m, int nf, cplx* const* krt, bool* ready,
                          cplx* const* krt_t = nullptr, bool* ready_t = nullptr, const double* src0 = nullptr, const ChainCtrl* ctrl0 = nullptr) {
    if (nf > PKB_MAX_COHORTS) return fail(PKB_ELIMIT, "at most %d earlier release days are supported", PKB_MAX_COHORTS);
    pkb_ctx* ctx = ch->ctx;
    const ChainDims& d = ch->d;
    const double* src = src0 ? src0 : ch->S[ch->cur].p;
    const ChainCtrl* src_ctrl = src0 ? ctrl0 : ch->ctrl.p;
    for (int j = nf - 1; j >= 0; --j) {
        if (!ch->coh[j].p) TRY(ch->coh[j].alloc(ctx, (size_t)d.P * d.ldS));
        cplx* kr = krt ? krt[j] : ch->Krt.p;
        const bool rdy = krt && ready && ready[j];
        // CalcSol.py:103-105 (same-shape re-FFT); with cached spectra on the truncated-source torus that geometry travels along too
        const bool tt = krt_t && ready_t && krt_t[j];
        TRY(conv_step(ch, src, src_ctrl, ch->coh[j].p, F[j], Wk[j], m[j], kr, rdy, 1 + j, 1, nullptr, false, -1, tt, tt ? krt_t[j] : nullptr,
                      tt && ready_t[j]));
        if (ready) ready[j] = true;
        if (tt) ready_t[j] = true;
        src = ch->coh[j].p;
        src_ctrl = ch->ctrl.p + 1 + j;
    }
    return 0;
}

extern "C" int pkb_chain_back_solve(pkb_chain* ch, const double* const* filters, const int* ks, int nf, double threshold, double* out,
                                    int* flags) {
    if (!ch || (nf > 0 && (!filters || !ks))) return fail(PKB_EINVAL, "pkb_chain_back_solve: NULL argument");
    if (nf > PKB_MAX_COHORTS) return fail(PKB_ELIMIT, "at most %d earlier release days are supported", PKB_MAX_COHORTS);
    pkb_ctx* ctx = ch->ctx;
    const ChainDims& d = ch->d;
    CU(cudaSetDevice(ctx->device));
    const double* src = ch->S[ch->cur].p;
    const ChainCtrl* src_ctrl = ch->ctrl.p;
    const size_t nd = (size_t)d.D * d.D;
    for (int j = nf - 1; j >= 0; --j) {
        if (!ch->coh[j].p) TRY(ch->coh[j].alloc(ctx, (size_t)d.P * d.ldS));
        int m = 0;
        TRY(upload_filter(ch, filters[j], ks[j], &m));
        TRY(conv_step(ch, src, src_ctrl, ch->coh[j].p, ch->kup.p, 2 * m + 1, m, ch->Krt.p, false, 1 + j, 1));
        if (out) {
            if (threshold < 0) LAUNCH(ctx, k_copy_domain, d.D, 256, 0, (const double*)ch->coh[j].p, d, ch->dout.p, (const StepMeta*)nullptr);
            else LAUNCH(ctx, k_emit_dense, d.D, 256, 0, (const double*)ch->coh[j].p, d, (const StepMeta*)(ch->meta.p + 1 + j), threshold, 0, 1, ch->dout.p, (int*)nullptr, 0);
            CU(cudaMemcpyAsync(out + nd * j, ch->dout.p, nd * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
        }
        src = ch->coh[j].p;
        src_ctrl = ch->ctrl.p + 1 + j;
    }
    std::vector<StepMeta> hm(nf > 0 ? nf : 1);
    if (nf > 0) CU(cudaMemcpyAsync(hm.data(), ch->meta.p + 1, sizeof(StepMeta) * nf, cudaMemcpyDeviceToHost, ctx->stream));
    TRY(sync_check(ctx, "pkb_chain_back_solve"));
    if (flags)
        for (int j = 0; j < nf; ++j) flags[j] = hm[j].flag;
    return 0;
}

// Population-model output from the cohorts of the last pkb_chain_back_solve
// (cohorts 0..ncoh-2) plus the current state (cohort ncoh-1).
extern "C" int pkb_chain_population(pkb_chain* ch, int ncoh, const double* weights, double r_number, double centre_extra, int add_centre,
                                    double negval, int first_day, double* out, double* pre) {
    if (!ch || !weights || !out) return fail(PKB_EINVAL, "pkb_chain_population: NULL argument");
    if (ncoh < 1 || ncoh > PKB_MAX_COHORTS) return fail(PKB_ELIMIT, "pkb_chain_population: 1..%d cohorts are supported", PKB_MAX_COHORTS);
    pkb_ctx* ctx = ch->ctx;
    const ChainDims& d = ch->d;
    CU(cudaSetDevice(ctx->device));
    CohortArgs ca;
    memset(&ca, 0, sizeof ca);
    ca.n = ncoh;
    for (int c = 0; c < ncoh; ++c) {
        if (c < ncoh - 1 && !ch->coh[c].p) return fail(PKB_ESTATE, "pkb_chain_population: cohort %d has not been computed (call pkb_chain_back_solve)", c);
        ca.S[c] = c < ncoh - 1 ? ch->coh[c].p : ch->S[ch->cur].p;
        ca.w[c] = weights[c];
    }
    DBuf<double> dpre;
    const size_t nd = (size_t)d.D * d.D;
    if (pre) TRY(dpre.alloc(ctx, nd));
    LAUNCH(ctx, k_emit_population, d.D, 256, 0, ca, d, r_number, centre_extra, add_centre, negval, first_day, ch->dout.p,
           pre ? dpre.p : (double*)nullptr);
    CU(cudaMemcpyAsync(out, ch->dout.p, nd * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
    if (pre) CU(cudaMemcpyAsync(pre, dpre.p, nd * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
    return sync_check(ctx, "pkb_chain_population");
}

extern "C" int pkb_chain_get_state(pkb_chain* ch, double* out) {
    if (!ch || !out) return fail(PKB_EINVAL, "pkb_chain_get_state: NULL argument");
    pkb_ctx* ctx = ch->ctx;
    const ChainDims& d = ch->d;
    CU(cudaSetDevice(ctx->device));
    const double* S = ch->S[ch->cur].p;
    for (int r = 0; r < d.P; ++r)
        CU(cudaMemcpyAsync(out + (size_t)r * d.P, S + (size_t)r * d.ldS, sizeof(double) * d.P, cudaMemcpyDeviceToHost, ctx->stream));
    return sync_check(ctx, "pkb_chain_get_state");
}

// ---------------------------------------------------------------------------
// one solve over the GPUs of a box (csrc/dist.cuh)
// ---------------------------------------------------------------------------
extern "C" int pkb_set_stream(pkb_ctx* ctx, void* stream) {
    if (!ctx) return fail(PKB_EINVAL, "pkb_set_stream: NULL context");
    CU(cudaSetDevice(ctx->device));
    TRY(sync_check(ctx, "pkb_set_stream"));
    ctx->stream = stream ? (cudaStream_t)stream : ctx->own_stream;
    return 0;
}

// dst[r][c] = centred (Wd x Wd) window of the (Ws x Ws) source, zero where the source has no cell.  grid = Wd, block = 128
__global__ void k_copy_window(const double* __restrict__ src, int Ws, double* __restrict__ dst, int Wd) {
    const int r = blockIdx.x, off = Ws / 2 - Wd / 2;
    for (int c = threadIdx.x; c < Wd; c += blockDim.x) {
        const int sr = r + off, sc = c + off;
        dst[(size_t)r * Wd + c] = (sr >= 0 && sr < Ws && sc >= 0 && sc < Ws) ? src[(size_t)sr * Ws + sc] : 0.0;
    }
}

extern "C" int pkb_kset_export_device(pkb_kset* ks, int i, void* dst_dev, int Wdst) {
    if (!ks || !dst_dev || i < 0 || i >= ks->nprob) return fail(PKB_EINVAL, "pkb_kset_export_device: bad argument");
    if (Wdst < 2 * ks->hmeta[i].rad + 1 || !(Wdst & 1)) return fail(PKB_EINVAL, "pkb_kset_export_device: window side %d does not hold radius %d", Wdst, ks->hmeta[i].rad);
    pkb_ctx* ctx = ks->ctx;
    CU(cudaSetDevice(ctx->device));
    LAUNCH(ctx, k_copy_window, Wdst, 128, 0, (const double*)(ks->acc.p + (size_t)ks->W * ks->W * i), ks->W, (double*)dst_dev, Wdst);
    return check_launches(ctx, "pkb_kset_export_device");
}

extern "C" int pkb_kset_from_device(pkb_ctx* ctx, const void* windows_dev, int n, int W, const int* rads, int rad_res, pkb_kset** out) {
    if (!ctx || !windows_dev || !rads || !out) return fail(PKB_EINVAL, "pkb_kset_from_device: NULL argument");
    if (n < 1 || W < 1 || !(W & 1) || rad_res < 1) return fail(PKB_EINVAL, "pkb_kset_from_device: bad sizes");
    CU(cudaSetDevice(ctx->device));
    pkb_kset* ks = new pkb_kset();
    ks->ctx = ctx;
    ks->nprob = n;
    ks->periods = 0;
    ks->keep_pre = false;
    ks->rad_res = rad_res;
    ks->racc = W / 2;
    ks->W = W;
    ks->hmeta.resize(n);
    ks->hdp.resize(n);
    for (int i = 0; i < n; ++i) {
        memset(&ks->hmeta[i], 0, sizeof(DayMeta));
        memset(&ks->hdp[i], 0, sizeof(DayParams));
        if (rads[i] < 0 || 2 * rads[i] + 1 > W) { delete ks; return fail(PKB_EINVAL, "pkb_kset_from_device: radius %d does not fit the window", rads[i]); }
        ks->hmeta[i].rad = rads[i];
    }
    const size_t nel = (size_t)W * W * n;
    int rc = ks->acc.alloc(ctx, nel);
    if (rc) { delete ks; return rc; }
    cudaError_t e = cudaMemcpyAsync(ks->acc.p, windows_dev, nel * sizeof(double), cudaMemcpyDeviceToDevice, ctx->stream);
    if (e != cudaSuccess) { delete ks; return fail(PKB_ECUDA, "pkb_kset_from_device: copy failed: %s", cudaGetErrorString(e)); }
    *out = ks;
    return 0;
}

struct pkb_dist {
    pkb_ctx* ctx;
    pkb_kset* ks;          // borrowed
    int nd;
    ChainDims d;
    FftPlan plan;
    DistGeom g;
    DBuf<double> S0;       // the placed kernel of day 0 (full real state, only for the first forward row pass)
    DBuf<cplx> Yt, Krt, cscr, Shat;
    size_t hstride;
    DBuf<double> Sloc;     // [Jr][ldL]
    int ldL;
    DBuf<RowStats> rstat;
    DBuf<int> done;
    DBuf<double> meta;     // [nd][4]
    DBuf<double> worst;
    DBuf<ChainCtrl> ctrl;
    cplx *send, *recv;
    double *stats, *allstats, *out;
    bool first;            // the next column pass starts from Yt (day 1)
};

static int dist_geom(pkb_ctx* ctx, int D, int mmax, int world, int rank, ChainDims* d, FftPlan* plan, DistGeom* g) {
    if (D < 1 || mmax < 0 || world < 1 || rank < 0 || rank >= world) return fail(PKB_EINVAL, "pkb_dist: bad sizes");
    memset(d, 0, sizeof *d);
    d->D = D;
    d->P = D + mmax;
    d->N = pkb_smooth_len(std::max(2, d->P + 2 * mmax));
    d->Nc = d->N / 2 + 1;
    d->ldS = roundup(d->P, 16);
    d->ldY = roundup(d->P, 2);
    d->ldW = roundup(d->N, 2);
    d->ldK = roundup(2 * mmax + 1, 2);
    TRY(get_plan(ctx, d->N, plan));
    if (fft_smem_bytes(*plan) > (size_t)ctx->max_smem || plan->grid_rows < 1 || plan->grid_cols < 1)
        return fail(PKB_ELIMIT, "torus side %d exceeds the shared-memory FFT limit", d->N);
    memset(g, 0, sizeof *g);
    g->G = world;
    g->rank = rank;
    g->Cg = roundup((d->Nc + world - 1) / world, PKB_CB);
    g->c0 = std::min(d->Nc, rank * g->Cg);
    g->c1 = std::min(d->Nc, (rank + 1) * g->Cg);
    g->mmax = mmax;
    g->J = d->P + 2 * mmax;
    g->Jr = roundup((g->J + world - 1) / world, 2);
    g->j0 = std::min(g->J, rank * g->Jr);
    g->j1 = std::min(g->J, (rank + 1) * g->Jr);
    g->jr_magic = g->Jr == 1 ? 0u : 0xFFFFFFFFu / (unsigned)g->Jr + 1u;
    g->blk = spec_size(g->Cg, g->Jr);
    if (g->J >= 65536 || d->Nc >= 65536) return fail(PKB_ELIMIT, "pkb_dist: torus too large for the 16-bit index arithmetic");
    return 0;
}

extern "C" int pkb_dist_plan(pkb_ctx* ctx, int dom_len, int mmax, int world, long long* xchg_elems, int* rows_per_rank, int* P, int* N) {
    if (!ctx) return fail(PKB_EINVAL, "pkb_dist_plan: NULL context");
    CU(cudaSetDevice(ctx->device));
    ChainDims d;
    FftPlan plan;
    DistGeom g;
    TRY(dist_geom(ctx, dom_len, mmax, world, 0, &d, &plan, &g));
    if (xchg_elems) *xchg_elems = (long long)g.blk * world;
    if (rows_per_rank) *rows_per_rank = g.Jr;
    if (P) *P = d.P;
    if (N) *N = d.N;
    return 0;
}

extern "C" int pkb_dist_create(pkb_ctx* ctx, pkb_kset* ks, int ndays, int rank, int world, void* send, void* recv, void* stats,
                               void* allstats, void* out, pkb_dist** handle) {
    if (!ctx || !ks || !send || !recv || !stats || !allstats || !out || !handle) return fail(PKB_EINVAL, "pkb_dist_create: NULL argument");
    *handle = nullptr;
    if (ndays < 1 || ndays > ks->nprob) return fail(PKB_EINVAL, "pkb_dist_create: ndays %d exceeds the kernel set", ndays);
    CU(cudaSetDevice(ctx->device));
    int mmax = 0;
    for (int i = 0; i < ndays; ++i) mmax = std::max(mmax, ks->hmeta[i].rad);
    const int D = 2 * ks->rad_res + 1;
    pkb_dist* h = new pkb_dist();
    struct Guard {
        pkb_dist* p;
        ~Guard() { delete p; }
    } guard{h};
    h->ctx = ctx;
    h->ks = ks;
    h->nd = ndays;
    TRY(dist_geom(ctx, D, mmax, world, rank, &h->d, &h->plan, &h->g));
    const ChainDims& d = h->d;
    const FftPlan& plan = h->plan;
    for (int i = 0; i < ndays; ++i)
        if (ks->hmeta[i].rad <= ctx->stencil_max_radius || ks->hmeta[i].rad > D / 2)
            return fail(PKB_ELIMIT, "pkb_dist_create: kernel radius %d is outside the range the distributed chain handles", ks->hmeta[i].rad);
    h->send = (cplx*)send; h->recv = (cplx*)recv;
    h->stats = (double*)stats; h->allstats = (double*)allstats; h->out = (double*)out;
    const int RL = plan_radix(plan, plan.nstage - 1);
    h->hstride = (size_t)plan.cols_kb * RL * plan.cols_threads;
    const size_t ns = (size_t)d.P * d.ldS;
    TRY(h->S0.alloc(ctx, ns));
    TRY(h->Yt.alloc(ctx, spec_size(d.Nc + 1, d.ldY)));
    TRY(h->Krt.alloc(ctx, spec_size(d.Nc + 1, d.ldK)));
    TRY(h->cscr.alloc(ctx, (size_t)ctx->occ_cap * ctx->sm_count * ((size_t)d.N + 21 * 256 + 256)));
    TRY(h->Shat.alloc(ctx, h->hstride * std::max(1, h->g.c1 - h->g.c0)));
    h->ldL = roundup(D, 16);
    TRY(h->Sloc.alloc(ctx, (size_t)h->g.Jr * h->ldL));
    TRY(h->rstat.alloc(ctx, h->g.Jr));
    TRY(h->done.alloc(ctx, 1));
    TRY(h->meta.alloc(ctx, (size_t)ndays * 4));
    TRY(h->worst.alloc(ctx, 1));
    TRY(h->ctrl.alloc(ctx, 1));
    CU(cudaMemsetAsync(h->done.p, 0, sizeof(int), ctx->stream));
    CU(cudaMemsetAsync(h->meta.p, 0, sizeof(double) * 4 * ndays, ctx->stream));
    CU(cudaMemsetAsync(h->worst.p, 0, sizeof(double), ctx->stream));
    CU(cudaMemsetAsync(h->ctrl.p, 0, sizeof(ChainCtrl), ctx->stream));
    CU(cudaMemsetAsync(h->S0.p, 0, ns * sizeof(double), ctx->stream));
    CU(cudaMemsetAsync(h->send, 0, sizeof(cplx) * h->g.blk * world, ctx->stream));      // (padding rows / columns of the blocks stay zero)
    // day 0: the first kernel recentred on the domain (Run.py:454-458); its row spectra feed the first column pass
    const size_t nW = (size_t)ks->W * ks->W;
    LAUNCH(ctx, k_place_kernel, 2 * ks->hmeta[0].rad + 1, 128, 0, (const double*)ks->acc.p, ks->W, ks->hmeta[0].rad, d, h->S0.p);
    TruncGeom tg;
    memset(&tg, 0, sizeof tg);
    const size_t sm1 = fft_smem_bytes(plan);
    LAUNCH(ctx, k_rows_fwd, std::min((d.P + 1) / 2, plan.grid_rows), plan.threads, sm1, (const double*)h->S0.p, d, (const ChainCtrl*)h->ctrl.p, h->Yt.p,
           plan, -1, tg, plan, 0);
    // this rank's rows of day 0 (un-thresholded copy, CalcSol.get_solutions leaves modelsol[0] as it is)
    {
        const int r0 = h->g.j0, r1 = std::min(h->g.j1, D);
        CU(cudaMemsetAsync(h->out, 0, sizeof(double) * (size_t)h->g.Jr * D, ctx->stream));
        if (r1 > r0)
            CU(cudaMemcpy2DAsync(h->out, sizeof(double) * D, h->S0.p + (size_t)r0 * d.ldS, sizeof(double) * d.ldS, sizeof(double) * D, r1 - r0,
                                 cudaMemcpyDeviceToDevice, ctx->stream));
    }
    (void)nW;
    h->first = true;
    TRY(check_launches(ctx, "pkb_dist_create"));
    guard.p = nullptr;
    *handle = h;
    return 0;
}

extern "C" int pkb_dist_step_cols(pkb_dist* h, int day) {
    if (!h || day < 1 || day >= h->nd) return fail(PKB_EINVAL, "pkb_dist_step_cols: bad argument");
    pkb_ctx* ctx = h->ctx;
    const ChainDims& d = h->d;
    const FftPlan& plan = h->plan;
    pkb_kset* ks = h->ks;
    const int m = ks->hmeta[day].rad;
    const double* K = ks->acc.p + (size_t)ks->W * ks->W * day;
    const size_t sm1 = fft_smem_bytes(plan);
    LAUNCH(ctx, k_kernel_rows, std::min(m + 1, plan.grid_rows), plan.threads, sm1, K, ks->W, m, d, h->Krt.p, plan);
    const int ncol = h->g.c1 - h->g.c0;
    if (ncol > 0)
        LAUNCH(ctx, k_cols_dist, std::min(ncol, plan.grid_cols), plan.cols_threads, sm1, (const cplx*)h->Yt.p, (const cplx*)h->Krt.p, m, d, h->send,
               h->cscr.p, plan, h->Shat.p, h->hstride, h->g, h->first ? 1 : 0);
    h->first = false;
    return check_launches(ctx, "pkb_dist_step_cols");
}

extern "C" int pkb_dist_step_rows(pkb_dist* h) {
    if (!h) return fail(PKB_EINVAL, "pkb_dist_step_rows: NULL handle");
    pkb_ctx* ctx = h->ctx;
    const FftPlan& plan = h->plan;
    const int njobs = (h->g.j1 - h->g.j0 + 1) / 2;
    LAUNCH(ctx, k_rows_inv_dist, std::max(1, std::min(njobs, plan.grid_rows)), plan.threads, fft_smem_bytes(plan), (const cplx*)h->recv, h->d, h->Sloc.p,
           h->ldL, h->rstat.p, 1e-8, plan, h->done.p, h->stats, h->g);
    return check_launches(ctx, "pkb_dist_step_rows");
}

extern "C" int pkb_dist_step_emit(pkb_dist* h, int day) {
    if (!h || day < 1 || day >= h->nd) return fail(PKB_EINVAL, "pkb_dist_step_emit: bad argument");
    pkb_ctx* ctx = h->ctx;
    LAUNCH(ctx, k_emit_dist, h->g.Jr, 256, 0, (const double*)h->Sloc.p, h->ldL, h->d, (const double*)h->allstats, h->g, 1e-8,
           h->out + (size_t)day * h->g.Jr * h->d.D, h->meta.p + 4 * (size_t)day, h->worst.p);
    return check_launches(ctx, "pkb_dist_step_emit");
}

extern "C" int pkb_dist_finish(pkb_dist* h, double* meta, int* ok) {
    if (!h || !ok) return fail(PKB_EINVAL, "pkb_dist_finish: NULL argument");
    pkb_ctx* ctx = h->ctx;
    std::vector<double> hm((size_t)h->nd * 4);
    CU(cudaMemcpyAsync(hm.data(), h->meta.p, sizeof(double) * hm.size(), cudaMemcpyDeviceToHost, ctx->stream));
    TRY(sync_check(ctx, "pkb_dist_finish"));
    // the criterion of chain.cuh (PKB_SPEC_EPS per day, PKB_SPEC_BUDGET accumulated) on the global maxima
    double emax = 0.0, esum = 0.0;
    int good = 1;
    for (int n = 1; n < h->nd; ++n) {
        const double pa = hm[4 * n + 3], pm = hm[4 * n + 2];
        emax = std::max(emax, pa);
        esum += 2.0 * emax;
        if (!(pa <= PKB_SPEC_EPS) || pm > 1e-8 || !(esum <= PKB_SPEC_BUDGET)) good = 0;
    }
    if (meta) memcpy(meta, hm.data(), sizeof(double) * hm.size());
    *ok = good;
    return 0;
}

extern "C" int pkb_dist_destroy(pkb_dist* h) {
    if (!h) return 0;
    cudaSetDevice(h->ctx->device);
    cudaStreamSynchronize(h->ctx->stream);
    delete h;
    return 0;
}

// ---------------------------------------------------------------------------
// fused forward solve
// ---------------------------------------------------------------------------
struct pkb_result {
    pkb_ctx* ctx;
    int ndays, D, P, N, max_shape;
    int window_steps;       // chain steps run on a support-window torus (ChainDims::win)
    DBuf<double> dense;     // [ndays][D][D]
    DBuf<double> pre;       // optional [ndays][D][D] un-thresholded (parity export)
    std::vector<DayMeta> kmeta;
    std::vector<StepMeta> smeta;
    std::vector<StepMeta> cmeta;   // [ndays][PKB_MAX_COHORTS]: back_solve steps of each day (population model, r_dur > 1)
    DBuf<int> rownnz;       // [ndays][D] non-zeros per output row (COO compaction)
    DBuf<long long> rowoff; // [ndays][D] exclusive scan of rownnz within each day
    DBuf<long long> daytot; // [ndays][2]: (0, non-zeros of the day)
    HBuf<long long> tot_host;
    std::vector<char> counted;   // days whose rownnz the emission kernel already filled
    std::vector<char> day_ready; // days whose row counts / event have been enqueued (coo_day_ready)
    HBuf<long long> dayoff;
    HBuf<long long> rowoff_host;   // CSR output: [ndays][D] first triplet of every row, relative to the day's start
    bool csr;                      // the result holds CSR pieces (want_coo == 2): no row array
    HBuf<int> rows, cols;
    HBuf<double> vals;
    bool have_coo;
    const StepMeta* win_meta;   // during the solve: per-day step records, when the dense days are only an intermediate of the compaction
                                // (k_emit_dense sparse_only: nothing outside a day's computed window was written)
};

// ---- COO output, pipelined with the chain ------------------------------------
// As soon as a day's dense solution has been emitted its rows are counted and
// scanned on the emitting stream (coo_day_ready); once the whole chain has been
// ENQUEUED the host walks the days in order, learns each day's size from an
// 16-byte D2H, and queues that day's compaction + triplet copy on the copy stream
// (coo_pump) -- so the 16 bytes/non-zero cross PCIe while later days are still
// being computed.
// wl != nullptr: called from the COO worker thread -- launches are counted there instead of in the context (whose counters
// and profile records belong to the thread that runs the chain)
#define COO_LAUNCH(wl, ctx, strm, kern, grid, block, smem, ...)                 \
    do {                                                                        \
        if (wl) { PKB_LAUNCH(kern, grid, block, smem, (strm), __VA_ARGS__); ++*(wl); } \
        else LAUNCH_ON(ctx, strm, kern, grid, block, smem, __VA_ARGS__);        \
    } while (0)
static int coo_day_ready(pkb_ctx* ctx, pkb_result* r, int day, cudaStream_t strm, long long* wl = nullptr) {
    const int D = r->D;
    const size_t nD = (size_t)D * D;
    if (!r->counted[day])
        COO_LAUNCH(wl, ctx, strm, k_row_nnz, D, 256, 0, (const double*)(r->dense.p + nD * day), D, r->rownnz.p + (size_t)D * day);
    COO_LAUNCH(wl, ctx, strm, k_row_scan, 1, 1024, 0, (const int*)(r->rownnz.p + (size_t)D * day), D, 1, r->rowoff.p + (size_t)D * day,
               r->daytot.p + 2 * day);
    CU(cudaMemcpyAsync(r->tot_host.p + 2 * day, r->daytot.p + 2 * day, 2 * sizeof(long long), cudaMemcpyDeviceToHost, strm));
    CU(cudaEventRecord(ctx->day_events[day], strm));
    r->day_ready[day] = 1;
    return 0;
}

template <class T>
static int hbuf_grow(pkb_ctx* ctx, HBuf<T>& b, size_t used, size_t newcap) {
    HBuf<T> nb;
    TRY(nb.alloc(ctx, newcap));
    if (used) memcpy(nb.p, b.p, used * sizeof(T));
    std::swap(nb.p, b.p);
    std::swap(nb.n, b.n);
    std::swap(nb.ctx, b.ctx);
    return 0;       // nb's destructor returns the old block to the pool
}

// Incremental form: coo_pump(block = false) handles the days whose dense solution already exists and returns; the
// chain loop calls it after every step, so the triplets cross PCIe while later days are still being computed even
// when the host paces the chain itself (tau windows wait for an older step's extent before every step).
struct CooState {
    DBuf<int> srow[2], scol[2];      // two staging areas, one per copy stream (days alternate)
    DBuf<double> sval[2];
    size_t cap = 0;
    int next_day = 0;
    bool started = false;
};

// upto: handle days < upto only; finish: join the copy streams into the main stream (the last call)
static int coo_pump(pkb_ctx* ctx, pkb_result* r, CooState* st, bool block, int upto = 1 << 30, bool finish = true, long long* wl = nullptr) {
    const int D = r->D, nd = std::min(r->ndays, upto);
    const size_t nD = (size_t)D * D;
    if (!st->started) {
        TRY(r->dayoff.alloc(ctx, r->ndays + 1));
        r->dayoff.p[0] = 0;
        st->cap = std::max<size_t>(ctx->coo_hint + ctx->coo_hint / 16, (size_t)r->ndays * D * 64);
        if (r->csr) TRY(r->rowoff_host.alloc(ctx, (size_t)r->ndays * D));
        else TRY(r->rows.alloc(ctx, st->cap));
        TRY(r->cols.alloc(ctx, st->cap));
        TRY(r->vals.alloc(ctx, st->cap));
        // one day's triplets at a time through a device staging area (worst case D*D entries)
        for (int b = 0; b < 2; ++b) {
            TRY(st->srow[b].alloc(ctx, nD));
            TRY(st->scol[b].alloc(ctx, nD));
            TRY(st->sval[b].alloc(ctx, nD));
        }
        st->started = true;
    }
    for (; st->next_day < nd; ++st->next_day) {
        const int day = st->next_day;
        if (!r->day_ready[day]) {
            if (block) return fail(PKB_ESTATE, "COO collection: day %d was never emitted", day);
            break;
        }
        cudaError_t e = block ? cudaEventSynchronize(ctx->day_events[day]) : cudaEventQuery(ctx->day_events[day]);
        if (e == cudaErrorNotReady) break;
        if (e != cudaSuccess) return fail(PKB_ECUDA, "waiting for day %d of the solve failed: %s", day, cudaGetErrorString(e));
        const long long tot = r->tot_host.p[2 * day + 1];
        const size_t off = (size_t)r->dayoff.p[day];
        r->dayoff.p[day + 1] = (long long)(off + tot);
        if (off + tot > st->cap) {
            CU(cudaStreamSynchronize(ctx->cp));                   // copies into the old blocks must have landed
            CU(cudaStreamSynchronize(ctx->cp2));
            const size_t newcap = std::max<size_t>(off + tot + (off + tot) / 4, st->cap * 2);
            if (!r->csr) TRY(hbuf_grow(ctx, r->rows, off, newcap));
            TRY(hbuf_grow(ctx, r->cols, off, newcap));
            TRY(hbuf_grow(ctx, r->vals, off, newcap));
            st->cap = newcap;
        }
        const int b = day & 1;
        cudaStream_t cs = b ? ctx->cp2 : ctx->cp;
        CU(cudaStreamWaitEvent(cs, ctx->day_events[day], 0));
        COO_LAUNCH(wl, ctx, cs, k_coo_write, D, 256, 0, (const double*)(r->dense.p + nD * day), D,
                   (const long long*)(r->rowoff.p + (size_t)D * day), (const int*)(r->rownnz.p + (size_t)D * day),
                   r->csr ? (int*)nullptr : st->srow[b].p, st->scol[b].p, st->sval[b].p, r->win_meta ? r->win_meta + day : (const StepMeta*)nullptr);
        if (r->csr)
            CU(cudaMemcpyAsync(r->rowoff_host.p + (size_t)D * day, r->rowoff.p + (size_t)D * day, sizeof(long long) * D, cudaMemcpyDeviceToHost, cs));
        if (tot > 0) {
            if (!r->csr) CU(cudaMemcpyAsync(r->rows.p + off, st->srow[b].p, sizeof(int) * tot, cudaMemcpyDeviceToHost, cs));
            CU(cudaMemcpyAsync(r->cols.p + off, st->scol[b].p, sizeof(int) * tot, cudaMemcpyDeviceToHost, cs));
            CU(cudaMemcpyAsync(r->vals.p + off, st->sval[b].p, sizeof(double) * tot, cudaMemcpyDeviceToHost, cs));
        }
    }
    if (block && finish) {
        CU(cudaEventRecord(ctx->ev_cp, ctx->cp));
        CU(cudaStreamWaitEvent(ctx->stream, ctx->ev_cp, 0));
        CU(cudaEventRecord(ctx->ev_cp2, ctx->cp2));
        CU(cudaStreamWaitEvent(ctx->stream, ctx->ev_cp2, 0));
        ctx->coo_hint = (size_t)r->dayoff.p[nd];
        r->have_coo = true;
    }
    return 0;
}

// The per-day output work of the fused solve -- row scan, a 16-byte D2H to learn the day's size, compaction kernel, the copies
// of the day's triplets -- is ~10 runtime calls per day.  The chain of a probability-model solve is paced by its own host
// thread (tau windows: step n is sized from the extent step n - 1 reports), so on the chain's thread those calls sat between
// one step's ticket and the next step's launch (C4: chain phase 13.2 -> 20.7 ms with CSR output).  Here a helper thread does
// them: the chain's thread records one event per emitted day and hands the day over.
struct CooWorker {
    pkb_ctx* ctx = nullptr;
    pkb_result* r = nullptr;
    CooState* st = nullptr;
    std::thread th;
    std::mutex mu;
    std::condition_variable cv;
    std::deque<int> q;
    bool done = false, running = false;
    int rc = 0;
    std::string msg;
    long long launches = 0;
    void start(pkb_ctx* c, pkb_result* res, CooState* s) {
        ctx = c; r = res; st = s;
        running = true;
        th = std::thread([this]() { run(); });
    }
    void push(int day) {
        { std::lock_guard<std::mutex> g(mu); q.push_back(day); }
        cv.notify_one();
    }
    void run() {
        cudaSetDevice(ctx->device);
        std::vector<int> days;
        for (;;) {
            days.clear();
            {
                std::unique_lock<std::mutex> lk(mu);
                cv.wait(lk, [this]() { return done || !q.empty(); });
                if (q.empty()) return;
                days.assign(q.begin(), q.end());      // everything handed over so far
                q.clear();
            }
            if (rc) continue;                   // (drain the queue after an error)
            // Day by day: the wait for the day's size, its compaction and copies.  The counting / scan / size copy of the NEXT day
            // (the other copy stream) is enqueued first, so that it runs under that wait -- but not further ahead: a later day of the
            // same stream would queue its wait for a far emission in front of this day's compaction.
            int e = 0;
            size_t readied = 0;
            auto ready_to = [&](size_t upto) {
                for (; readied < upto && readied < days.size() && !e; ++readied) {
                    const int day = days[readied];
                    cudaStream_t cs = (day & 1) ? ctx->cp2 : ctx->cp;
                    if (cudaStreamWaitEvent(cs, ctx->emit_events[day], 0) != cudaSuccess) e = fail(PKB_ECUDA, "COO worker: waiting for day %d failed", day);
                    if (!e) e = coo_day_ready(ctx, r, day, cs, &launches);
                }
            };
            for (size_t i = 0; i < days.size() && !e; ++i) {
                ready_to(i + 2);
                if (!e) e = coo_pump(ctx, r, st, true, days[i] + 1, false, &launches);
            }
            if (e) { rc = e; msg = g_err; }
        }
    }
    // all days handed over: wait for the worker; its error (if any) becomes this thread's
    int join() {
        if (!running) return 0;
        { std::lock_guard<std::mutex> g(mu); done = true; }
        cv.notify_one();
        th.join();
        running = false;
        ctx->launches += launches;
        if (rc) g_err = msg;
        return rc;
    }
    ~CooWorker() { join(); }
};

static int check_solve_args(const pkb_solve_args* a) {
    if (a->ndays < 1 || a->ndays > a->nd_wind) return fail(PKB_EINVAL, "pkb_solve: ndays %d must be in [1, %d]", a->ndays, a->nd_wind);
    if (a->sprd && !(a->sprd_factor >= 0 && a->sprd_factor <= 1)) return fail(PKB_EINVAL, "pkb_solve: sprd_factor must be in [0, 1]");
    if (!a->prob_model) {
        if (a->r_dur < 1 || a->r_dur > a->ndays + (a->sprd ? 1 : 0)) return fail(PKB_EINVAL, "pkb_solve: r_dur %d must be in [1, ndays]", a->r_dur);
        if (a->r_dur > PKB_MAX_COHORTS) return fail(PKB_ELIMIT, "pkb_solve: r_dur is limited to %d days", PKB_MAX_COHORTS);
        if (!a->r_dist) return fail(PKB_EINVAL, "pkb_solve: r_dist is NULL");
    }
    return 0;
}

// per-day prob_mass arguments of one solve (Run.py:412-425); with a->sprd the day-0 spread kernel of
// Bayes_Run.py:245-270 comes first: ndays + 1 problems
static int solve_nkernels(const pkb_solve_args* a) { return a->ndays + (a->sprd ? 1 : 0); }
static void solve_day_args(const pkb_solve_args* a, pkb_day_args* dargs) {
    const int lead = a->sprd ? 1 : 0;
    if (lead) {
        dargs[0] = a->day;
        dargs[0].kind = 1;
        dargs[0].wind_day = 0;
        dargs[0].single = 0;
        dargs[0].start_time = -1.0;
        dargs[0].sprd_factor = a->sprd_factor;
        dargs[0].sprd_drift[0] = a->sprd_drift[0];
        dargs[0].sprd_drift[1] = a->sprd_drift[1];
    }
    for (int i = 0; i < a->ndays; ++i) {
        pkb_day_args& d = dargs[lead + i];
        d = a->day;
        d.kind = 0;
        d.wind_day = i;
        d.single = 0;
        d.start_time = (!a->prob_model && i == 0 && a->r_start >= 0) ? a->r_start : -1.0;   // Run.py:418-421
    }
}

// Phase 2 and outputs of one solve whose per-day kernels are problems k0 .. k0 + ndays - 1 of ks.
// Likelihood batches read the model only at K sample cells: with a sink the emission kernels evaluate
// just those cells into out[nd][K] (device), no dense day is materialised, no result object is returned
// and the chain is left enqueued on ctx's streams (the caller synchronises once per group of proposals).
// a child context works with its parent's options
static void lane_inherit(const pkb_ctx* ctx, pkb_ctx* lane) {
    lane->stencil_max_radius = ctx->stencil_max_radius;
    lane->fft_threads = ctx->fft_threads;
    lane->use_windows = ctx->use_windows;
    lane->use_fusion = ctx->use_fusion;
    lane->use_step_torus = ctx->use_step_torus;
    lane->use_trunc_torus = ctx->use_trunc_torus;
    lane->use_spectral = ctx->use_spectral;
    lane->spec_min_reach = ctx->spec_min_reach;
    lane->use_rowwin = ctx->use_rowwin;
    lane->use_tau_windows = ctx->use_tau_windows;
    lane->tau_lag = ctx->tau_lag;
    lane->ring_tol = ctx->ring_tol;
    lane->rows_desc = ctx->rows_desc;
    lane->prof_on = ctx->prof_on;
    lane->batch_chain = ctx->batch_chain;
    lane->batch_occ = ctx->batch_occ;
}

struct SampleSink {
    const int* cells;       // device [K][2]
    int K;
    double* out;            // device [nd][K]
};
static int solve_chain(pkb_ctx* ctx, const pkb_solve_args* a, pkb_kset* ks, int k0, pkb_result** out, const SampleSink* sink = nullptr) {
    // chain days: with a->sprd the spread kernel is day 0 of the chain and its output is dropped (Bayes_Run.py:288-296);
    // output day o is chain day o + lead
    const int lead = a->sprd ? 1 : 0;
    const int nd = a->ndays + lead, nout = a->ndays;
    if (sink && a->want_coo) return fail(PKB_EINVAL, "sample-cell emission excludes COO output");
    const int sgrid = sink ? (sink->K + 255) / 256 : 0;
    const double negval = a->negval > 0 ? a->negval : 1e-8;
    CU(cudaEventRecord(ctx->ev[1], ctx->stream));

    int mmax = 0;
    for (int i = 0; i < nd; ++i) mmax = std::max(mmax, ks->hmeta[k0 + i].rad);      // Run.py:426-429
    const int D = 2 * ks->rad_res + 1;

    pkb_result* res = new pkb_result();
    struct RGuard {
        pkb_result* r;
        ~RGuard() { delete r; }
    } rguard{res};
    res->ctx = ctx;
    res->ndays = nout;
    res->D = D;
    res->max_shape = 2 * mmax + 1;
    res->have_coo = false;
    res->win_meta = nullptr;
    res->csr = a->want_coo == 2;
    res->window_steps = 0;
    res->kmeta.assign(ks->hmeta.begin() + k0 + lead, ks->hmeta.begin() + k0 + nd);
    res->smeta.assign(nout, StepMeta());
    for (auto& sm : res->smeta) memset(&sm, 0, sizeof sm);

    // ---- phase 2 -----------------------------------------------------------
    pkb_chain* ch = nullptr;
    TRY(chain_create(ctx, D, mmax, &ch));
    struct CGuard {
        pkb_chain* c;
        ~CGuard() { delete c; }
    } cguard{ch};
    ch->negval = negval;
    // spectral-resident steps on the main chain (probability model, or a one-day release: no cohorts to back-solve)
    // Spectral-resident steps on the main chain (probability model, or a one-day release: no cohorts to back-solve).
    // They need ONE torus for every whole-torus step, which gives up the per-step torus, so they are only armed for
    // solves that look like the in-domain regime: the exact support (first kernel, grown by every later radius) stays
    // inside the domain for at least PKB_SPEC_MIN_REACH steps.  (A performance heuristic only: the device-side
    // criterion of chain.cuh decides whether any step actually goes spectral.)
    int reach = 0;
    {
        int lo = D / 2 - ks->hmeta[k0].rad, hi = D / 2 + ks->hmeta[k0].rad;
        for (int n = (a->prob_model ? 1 : a->r_dur); n < nd; ++n, ++reach) {
            lo -= ks->hmeta[k0 + n].rad;
            hi += ks->hmeta[k0 + n].rad;
            if (lo < 0 || hi >= D) break;
        }
    }
    const bool spec_arm = ctx->use_spectral && (a->prob_model || a->r_dur == 1) && reach >= ctx->spec_min_reach;
    const int spec_try = spec_arm ? (a->prob_model ? 2 : 1) : 0;      // (row windows: the emission of the probability model knows about them)
    ch->fixed_torus = spec_arm;
    const ChainDims d = ch->d;
    res->P = d.P;
    res->N = d.N;
    const size_t nD = (size_t)D * D, nW = (size_t)ks->W * ks->W;
    if (!sink) TRY(res->dense.alloc(ctx, nD * nout));
    if (!sink && a->keep_pre_device) TRY(res->pre.alloc(ctx, nD * nout));      // parity export (pkb_result_pre)
    res->counted.assign(nout, 0);
    res->day_ready.assign(nout, 0);
    CooState coo;
    if (a->want_coo) {
        TRY(res->rownnz.alloc(ctx, (size_t)nout * D));
        TRY(res->rowoff.alloc(ctx, (size_t)nout * D));
        TRY(res->daytot.alloc(ctx, 2 * (size_t)nout));
        TRY(res->tot_host.alloc(ctx, 2 * (size_t)nout));
        while ((int)ctx->day_events.size() < nout) {
            cudaEvent_t e;
            CU(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
            ctx->day_events.push_back(e);
        }
    }
    // (day: chain day; nothing is emitted for the dropped leading day)
#ifdef PKB_EMUL
    bool coo_threaded = false;                          // (the CPU emulation of CUDA blocks is not re-entrant)
#else
    bool coo_threaded = a->want_coo && ctx->coo_thread && !ctx->prof_on;      // (the per-kernel profile belongs to this thread)
#endif
    CooWorker worker;
    if (coo_threaded) {
        while ((int)ctx->emit_events.size() < nout) {
            cudaEvent_t e;
            CU(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
            ctx->emit_events.push_back(e);
        }
        try {
            worker.start(ctx, res, &coo);
        } catch (const std::exception&) {               // no thread to be had: the chain's own thread does the output work
            worker.running = false;
            coo_threaded = false;
        }
    }
    auto emitted = [&](int day, cudaStream_t strm) -> int {
        if (!a->want_coo || day < lead) return 0;
        if (coo_threaded) {
            CU(cudaEventRecord(ctx->emit_events[day - lead], strm));
            worker.push(day - lead);
            return 0;
        }
        TRY(coo_day_ready(ctx, res, day - lead, strm));
        return coo_pump(ctx, res, &coo, false);         // whatever is ready by now goes to the host
    };
    DBuf<StepMeta> dsm, dcm;
    TRY(dsm.alloc(ctx, nd));
    CU(cudaMemsetAsync(dsm.p, 0, sizeof(StepMeta) * nd, ctx->stream));
    // the dense days are only an intermediate of the COO / CSR compaction: the emission writes, and the compaction reads, nothing
    // outside the rows and columns a day's step computed (probability model; k_emit_population writes whole days)
    const bool sparse_only = a->prob_model && a->want_coo && !a->want_dense_host && !a->keep_dense_device && !a->keep_pre_device && !sink;
    res->win_meta = sparse_only ? dsm.p + lead : nullptr;
    struct WmGuard {        // (the step records live as long as this call)
        pkb_result* r;
        ~WmGuard() { r->win_meta = nullptr; }
    } wm_guard{res};
    const bool want_cmeta = !sink && !a->prob_model && a->r_dur > 1;
    if (want_cmeta) {
        TRY(dcm.alloc(ctx, (size_t)nd * PKB_MAX_COHORTS));
        CU(cudaMemsetAsync(dcm.p, 0, sizeof(StepMeta) * nd * PKB_MAX_COHORTS, ctx->stream));
    }
    // back_solve steps of `day` (cohorts 0 .. nc-1, slots 1 .. nc of the chain's meta block) -> cmeta[day][.]
    auto keep_cmeta = [&](int day, int nc) -> int {
        if (want_cmeta && nc > 0)
            CU(cudaMemcpyAsync(dcm.p + (size_t)day * PKB_MAX_COHORTS, ch->meta.p + 1, sizeof(StepMeta) * nc, cudaMemcpyDeviceToDevice, ctx->stream));
        return 0;
    };
    auto kern = [&](int i) { return (const double*)(ks->acc.p + nW * (k0 + i)); };
    auto krad = [&](int i) { return ks->hmeta[k0 + i].rad; };

    // Row spectra of the daily kernels, a block of days per launch, ahead of the
    // chain (phase 1 left every kernel on the device): off the critical path.
    const size_t krt_stride = spec_size(d.Nc + 1, d.ldK);
    const int kr_chunk = (int)std::max<size_t>(1, std::min<size_t>(PKB_KR_MAXD, ((size_t)4 << 30) / (krt_stride * sizeof(cplx))));
    DBuf<cplx> krt_all, krt_all_t;      // (_t: on the truncated-source torus of each day, trunc_torus)
    TRY(krt_all.alloc(ctx, krt_stride * std::min(kr_chunk, nd)));
    if (ctx->use_trunc_torus) TRY(krt_all_t.alloc(ctx, krt_stride * std::min(kr_chunk, nd)));
    int kr_first = -1;      // first day held in krt_all
    bool kr_wait = false;   // the current block was launched on the side stream: the chain waits for ev_kr[1] before its first use
    struct KrJoin {         // (also on error returns: the main stream joins the side-stream batch before krt_all* are released)
        pkb_ctx* c;
        bool* w;
        ~KrJoin() { if (*w) cudaStreamWaitEvent(c->stream, c->ev_kr[1], 0); }
    } kr_join{ctx, &kr_wait};
    auto day_spectra = [&](int n, cplx** out, cplx** out_t, bool side = false) -> int {
        // spectra of day n (nullptr: the stencil path needs none; *out_t nullptr: no smaller truncated-source torus)
        *out = nullptr;
        *out_t = nullptr;
        if (krad(n) <= ctx->stencil_max_radius) return 0;
        if (kr_wait && !side) {     // first request by the chain since a block was launched on the side stream: join it BEFORE
            // anything else touches krt_all* (a relaunch below would otherwise race with that block's writes)
            CU(cudaStreamWaitEvent(ctx->stream, ctx->ev_kr[1], 0));
            kr_wait = false;
        }
        if (kr_first < 0 || n >= kr_first + kr_chunk) {
            kr_first = n;
            const int cnt = std::min(kr_chunk, nd - n);
            cudaStream_t strm = ctx->stream;
            if (side) {      // ahead of the chain, next to its first (support-window) steps: the kernels exist once phase 1 is done
                CU(cudaEventRecord(ctx->ev_kr[0], ctx->stream));
                CU(cudaStreamWaitEvent(ctx->aux, ctx->ev_kr[0], 0));
                strm = ctx->aux;
            }
            // one launch per distinct torus among the block's days (step_torus / trunc_torus: a handful of sizes)
            std::vector<int> tor(cnt), tor_t(cnt);
            std::vector<ChainDims> dims(cnt), dims_t(cnt);
            std::vector<FftPlan> plans(cnt), plans_t(cnt);
            for (int i = 0; i < cnt; ++i) {
                TRY(step_torus(ch, krad(n + i), &dims[i], &plans[i]));
                tor[i] = dims[i].N;
                TruncGeom tg;
                TRY(trunc_torus(ch, krad(n + i), dims[i], plans[i], &tg, &plans_t[i]));
                tor_t[i] = krt_all_t.p ? tg.N : 0;
                dims_t[i] = tg.N ? trunc_dims(dims[i], tg) : dims[i];
            }
            auto batch = [&](const std::vector<int>& tr, const std::vector<ChainDims>& dm, const std::vector<FftPlan>& pls, cplx* dst) -> int {
                std::vector<int> sizes(tr);
                std::sort(sizes.begin(), sizes.end());
                sizes.erase(std::unique(sizes.begin(), sizes.end()), sizes.end());
                for (int Nd : sizes) {
                    if (Nd == 0) continue;
                    KrBatch kb;
                    memset(&kb, 0, sizeof kb);
                    kb.nd = cnt;
                    int rep = -1;
                    for (int i = 0; i < cnt; ++i) {
                        kb.m[i] = krad(n + i);
                        const bool mine = tr[i] == Nd && kb.m[i] > ctx->stencil_max_radius;
                        if (mine) rep = i;
                        kb.job0[i + 1] = kb.job0[i] + (mine ? kb.m[i] + 1 : 0);
                    }
                    if (rep < 0 || kb.job0[cnt] == 0) continue;
                    const FftPlan& pl = pls[rep];
                    LAUNCH_ON(ctx, strm, k_kernel_rows_batch, std::min(kb.job0[cnt], pl.grid_rows), pl.threads, fft_smem_bytes(pl), kern(n), nW, ks->W,
                              kb, dm[rep], dst, krt_stride, pl);
                }
                return 0;
            };
            TRY(batch(tor, dims, plans, krt_all.p));
            if (krt_all_t.p) TRY(batch(tor_t, dims_t, plans_t, krt_all_t.p));
            if (side) {
                CU(cudaEventRecord(ctx->ev_kr[1], ctx->aux));
                kr_wait = true;
            }
        }
        *out = krt_all.p + krt_stride * (n - kr_first);
        if (krt_all_t.p) {
            ChainDims dd;
            FftPlan pp, pt;
            TruncGeom tg;
            TRY(step_torus(ch, krad(n), &dd, &pp));
            TRY(trunc_torus(ch, krad(n), dd, pp, &tg, &pt));
            if (tg.N) *out_t = krt_all_t.p + krt_stride * (n - kr_first);
        }
        return 0;
    };

    // Support window of the state (host-side bookkeeping, exact: the first state is a kernel of
    // radius m0 at the domain centre and every step grows the support by the day's radius).
    // While the grown window stays inside the domain the step runs on a torus sized for the
    // window instead of the whole padded domain (ChainDims::win).
    int wr0 = D / 2 - krad(0), wn = 2 * krad(0) + 1;
    bool wmode = ctx->use_windows != 0;
    if (wmode) CU(cudaMemsetAsync(ch->S[1].p, 0, (size_t)d.P * d.ldS * sizeof(double), ctx->stream));
    int win[3];
    // "tau windows" (probability model): the window follows the NUMERICAL support -- the extent of the cells with
    // |value| >= PKB_SPEC_TAU = 1e-15, measured on the device by every window step (ChainCtrl::er0..ec1) and read by the
    // host `tau_lag` steps late (pinned copy + event, so the host never waits for the step in flight).  A cell
    // farther than m from that extent stays below TAU after the next convolution (a weighted average with unit mass),
    // so state n - 1 has nothing above TAU outside  extent(n - L) grown by m_{n-L+1} .. m_{n-1} ; what a window step
    // leaves out is charged to the same budget as the spectral-resident steps (<= 2 TAU per step).  The windows may
    // shrink and move with the plume, so a day's emission only reads the region its step wrote (StepMeta::wr0..wc1).
    struct Box {
        int r0, r1, c0, c1;
    };
    const bool tau_mode_ok = wmode && ctx->use_tau_windows && a->prob_model;
    bool tau_mode = tau_mode_ok;
    std::vector<Box> reg(nd), meas(nd);
    std::vector<char> have(nd, 0);
    HBuf<int> hbox;
    DBuf<cplx> krt_ring;
    const int kRing = 4;
    if (tau_mode) {
        TRY(hbox.alloc(ctx, 8 * (size_t)nd));
        memset(hbox.p, 0, sizeof(int) * 8 * (size_t)nd);
        TRY(krt_ring.alloc(ctx, krt_stride * kRing));
        while ((int)ctx->box_events.size() < nd) {
            cudaEvent_t e;
            CU(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
            ctx->box_events.push_back(e);
        }
        const int c0 = D / 2 - krad(0);
        reg[0] = meas[0] = Box{c0, c0 + 2 * krad(0) + 1, c0, c0 + 2 * krad(0) + 1};
        have[0] = 1;
        CU(cudaEventRecord(ctx->ev_step[0], ctx->stream));            // the kernels exist (phase 1 ran on the main stream)
        CU(cudaStreamWaitEvent(ctx->aux, ctx->ev_step[0], 0));
    }
    int ring_used = 0;
    cudaEvent_t tau_event = nullptr;
    // window of step n in tau mode (nullptr: the step cannot run as a window step any more -> whole-torus steps from here)
    auto tau_window = [&](int n, int* rc) -> const int* {
        *rc = 0;
        const int m = krad(n), k = n - 1;
        if (m <= ctx->stencil_max_radius) return nullptr;
        const int src = std::max(0, n - std::max(1, ctx->tau_lag));
        if (src >= 1 && !have[src]) {
            // the finalising CTA of step `src` writes its extent and then the ticket src + 1 straight into this pinned slot
            volatile int* hv = hbox.p + 8 * src;
            const auto t0 = std::chrono::steady_clock::now();
            for (long spins = 0; hv[4] != src + 1; ++spins) {
#if defined(__x86_64__) && !defined(PKB_EMUL)
                __builtin_ia32_pause();
#endif
                if ((spins & 1023) == 1023) {
                    if (cudaEventQuery(ctx->box_events[src]) == cudaSuccess && hv[4] != src + 1) {
                        // (emulation build / no mapped write seen: fall back to the device copy of the control block)
                        int tmp[4];
                        if (cudaMemcpy(tmp, &ch->ctrl.p->er0, sizeof tmp, cudaMemcpyDeviceToHost) != cudaSuccess) break;
                        hv[0] = tmp[0]; hv[1] = tmp[1]; hv[2] = tmp[2]; hv[3] = tmp[3]; hv[4] = src + 1;
                        break;
                    }
                    if (std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count() > 30.0) {
                        *rc = fail(PKB_ECUDA, "timed out waiting for the extent of day %d", src);
                        return nullptr;
                    }
                }
            }
            const int* hb = hbox.p + 8 * src;
            meas[src] = (hb[1] > hb[0] && hb[3] > hb[2]) ? Box{hb[0], hb[1], hb[2], hb[3]} : reg[src];
            have[src] = 1;
        }
        Box b = have[src] ? meas[src] : reg[src];
        for (int j = src + 1; j <= k; ++j) {
            const int mj = krad(j);
            b = Box{std::max(reg[j].r0, b.r0 - mj), std::min(reg[j].r1, b.r1 + mj), std::max(reg[j].c0, b.c0 - mj), std::min(reg[j].c1, b.c1 + mj)};
        }
        // square window of side s inside the region state k was written to
        const Box& rk = reg[k];
        const int s = std::min(std::max(b.r1 - b.r0, b.c1 - b.c0), std::min(rk.r1 - rk.r0, rk.c1 - rk.c0));
        auto place = [&](int lo, int hi, int rlo, int rhi) {       // origin of a length-s interval covering [lo, hi) inside [rlo, rhi)
            int o = lo - (s - (hi - lo)) / 2;
            o = std::max(rlo, std::min(o, rhi - s));
            return o;
        };
        const int or0 = place(b.r0, b.r1, rk.r0, rk.r1), oc0 = place(b.c0, b.c1, rk.c0, rk.c1);
        if (or0 - m < 0 || oc0 - m < 0 || or0 + s + m > D || oc0 + s + m > D) return nullptr;
        if (pkb_smooth_len(s + 2 * m) >= d.N) return nullptr;
        win[0] = or0; win[1] = oc0; win[2] = s;
        reg[n] = Box{or0 - m, or0 + s + m, oc0 - m, oc0 + s + m};
        res->window_steps++;
        return win;
    };
    // kernel row spectra of a tau-window step: just in time on the side stream, a ring of kRing slots
    auto tau_spectra = [&](int n, const int* wp, cplx** out) -> int {
        const int m = krad(n), slot = ring_used % kRing;
        ChainDims dw = d;
        FftPlan pw;
        dw.N = pkb_smooth_len(std::max(2, wp[2] + 2 * m));
        dw.Nc = dw.N / 2 + 1;
        dw.ldK = roundup(2 * m + 1, 2);
        TRY(get_plan(ctx, dw.N, &pw));
        if (pw.grid_rows < 1) return fail(PKB_ELIMIT, "FFT kernels cannot be resident at torus side %d", dw.N);
        if (ring_used >= kRing) CU(cudaStreamWaitEvent(ctx->aux, ctx->ring_events[slot], 0));      // the step that read this slot last is done
        if (ctx->prof_on) prof_begin(ctx, "k_kernel_rows_win", ctx->aux);
        PKB_LAUNCH(k_kernel_rows, std::min(m + 1, pw.grid_rows), pw.threads, fft_smem_bytes(pw), ctx->aux, kern(n), ks->W, m, dw,
                   krt_ring.p + krt_stride * slot, pw);
        if (ctx->prof_on) prof_end(ctx, ctx->aux);
        ctx->launches++;
        while ((int)ctx->win_events.size() < kRing) {
            cudaEvent_t e;
            CU(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
            ctx->win_events.push_back(e);
        }
        CU(cudaEventRecord(ctx->win_events[slot], ctx->aux));
        tau_event = ctx->win_events[slot];              // (the step's column pass waits for it, conv_step)
        *out = krt_ring.p + krt_stride * slot;
        return 0;
    };
    auto tau_after_step = [&](int n) -> int {          // the step has been enqueued: its slot may be reused, its extent travels to the host
        CU(cudaEventRecord(ctx->ring_events[ring_used % kRing], ctx->stream));
        ++ring_used;
        CU(cudaEventRecord(ctx->box_events[n], ctx->stream));
        ch->host_box = nullptr;
        return 0;
    };
    // leaving tau mode before step n: older windows may have left cells outside the region of state n - 1
    auto tau_leave = [&](int n) -> int {
        const Box& rk = reg[n - 1];
        LAUNCH(ctx, k_zero_outside, d.P, 256, 0, ch->S[ch->cur].p, d, rk.r0, rk.r1, rk.c0, rk.c1);
        tau_mode = false;
        wmode = false;
        return 0;
    };
    auto step_window = [&](int n) -> const int* {
        const int m = krad(n);
        if (wmode && !tau_mode_ok && wr0 - m >= 0 && wr0 + wn + m <= D && pkb_smooth_len(wn + 2 * m) < d.N) {
            win[0] = win[1] = wr0; win[2] = wn;
            wr0 -= m; wn += 2 * m;
            if (m <= ctx->stencil_max_radius) return nullptr;       // the stencil path works on the whole torus
            res->window_steps++;
            return win;
        }
        wmode = false;      // the state is no longer confined: whole-torus steps from here on
        return nullptr;
    };

    // Row spectra of the kernels of the window steps: every such step has its own torus, so they
    // cannot share the batched launch above; they are launched on the side stream now (dry run of
    // the window bookkeeping) and the chain waits on a per-step event.
    std::vector<int> win_slot(nd, -1);
    DBuf<cplx> krt_win;
    {
        const int first = a->prob_model ? 1 : a->r_dur;
        std::vector<std::array<int, 3> > wsteps;     // (day, window side, torus side)
        if ((a->prob_model || a->r_dur == 1) && !tau_mode_ok) {
            const int sv_wr0 = wr0, sv_wn = wn, sv_ws = res->window_steps;
            const bool sv_mode = wmode;
            for (int n = first; n < nd; ++n) {
                const int side = wn;
                if (step_window(n)) wsteps.push_back({n, side, pkb_smooth_len(std::max(2, side + 2 * krad(n)))});
                if (!wmode) break;
            }
            wr0 = sv_wr0; wn = sv_wn; wmode = sv_mode; res->window_steps = sv_ws;
        }
        if (!wsteps.empty()) {
            TRY(krt_win.alloc(ctx, krt_stride * wsteps.size()));
            while (ctx->win_events.size() < wsteps.size()) {
                cudaEvent_t e;
                CU(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
                ctx->win_events.push_back(e);
            }
            CU(cudaEventRecord(ctx->ev_step[0], ctx->stream));            // the kernels exist (phase 1 ran on the main stream)
            CU(cudaStreamWaitEvent(ctx->aux, ctx->ev_step[0], 0));
            for (size_t i = 0; i < wsteps.size(); ++i) {
                const int n = wsteps[i][0], m = krad(n);
                ChainDims dw = d;
                FftPlan pw;
                dw.N = wsteps[i][2];
                dw.Nc = dw.N / 2 + 1;
                dw.ldK = roundup(2 * m + 1, 2);
                TRY(get_plan(ctx, dw.N, &pw));
                if (pw.grid_rows < 1) return fail(PKB_ELIMIT, "FFT kernels cannot be resident at torus side %d", dw.N);
                if (ctx->prof_on) prof_begin(ctx, "k_kernel_rows_win", ctx->aux);
                PKB_LAUNCH(k_kernel_rows, std::min(m + 1, pw.grid_rows), pw.threads, fft_smem_bytes(pw), ctx->aux, kern(n), ks->W, m, dw,
                           krt_win.p + krt_stride * i, pw);
                if (ctx->prof_on) prof_end(ctx, ctx->aux);
                ctx->launches++;
                CU(cudaEventRecord(ctx->win_events[i], ctx->aux));
                win_slot[n] = (int)i;
            }
        }
    }
    // steps n and n + 1 both whole-torus FFT steps on the SAME torus: the inverse row pass of step n can
    // also run the forward row pass of step n + 1
    auto fusable = [&](int n) -> bool {
        if (!ctx->use_fusion || krad(n) <= ctx->stencil_max_radius || krad(n + 1) <= ctx->stencil_max_radius) return false;
        ChainDims da, db;
        FftPlan pa, pb;
        if (step_torus(ch, krad(n), &da, &pa) || step_torus(ch, krad(n + 1), &db, &pb)) return false;
        return da.N == db.N && rows_fusable(da.N, da.P);
    };
    int fused_m = -1;       // >= 0: the previous step's k_rows_inv already transformed the interior row pairs (its filter radius)
    auto window_spectra = [&](int n, cplx** out) -> int {
        *out = nullptr;
        if (win_slot[n] < 0) return 0;
        CU(cudaStreamWaitEvent(ctx->stream, ctx->win_events[win_slot[n]], 0));
        *out = krt_win.p + krt_stride * win_slot[n];
        return 0;
    };

    {   // row spectra of the first block of days: on the side stream, hidden behind the first steps of the chain
        const int first = a->prob_model ? 1 : a->r_dur;
        int nf = -1;
        for (int n = first; n < nd && nf < 0; ++n)
            if (krad(n) > ctx->stencil_max_radius && win_slot[n] < 0) nf = n;
        // (the block starts at the first day the chain will ask these spectra for -- the first FFT step that is not a
        // support-window step, known from the dry run above -- exactly as the lazy path would)
        if (nf >= 0 && !tau_mode_ok) {      // (tau windows: which day leaves window mode depends on the data)
            cplx *k0 = nullptr, *k1 = nullptr;
            TRY(day_spectra(nf, &k0, &k1, true));
        }
    }
    if (a->prob_model) {
        // modelsol[0] = first kernel re-centred on the domain (Run.py:454-458)
        TRY(set_state_kernel_dev(ch, kern(0), ks->W, krad(0)));
        if (!lead) {
            if (sink) LAUNCH(ctx, k_copy_domain_cells, sgrid, 256, 0, (const double*)ch->S[ch->cur].p, d, sink->cells, sink->K, sink->out);
            else LAUNCH(ctx, k_copy_domain, D, 256, 0, (const double*)ch->S[ch->cur].p, d, res->dense.p, (const StepMeta*)nullptr);
            if (res->pre.p) LAUNCH(ctx, k_copy_domain, D, 256, 0, (const double*)ch->S[ch->cur].p, d, res->pre.p, (const StepMeta*)nullptr);
            TRY(emitted(0, ctx->stream));
        }
        for (int n = 1; n < nd; ++n) {                                          // CalcSol.py:191-201
            const int* wp = nullptr;
            cplx* krt = nullptr;
            cplx* krt_t = nullptr;
            bool tau_step = false;
            if (tau_mode) {
                int rcw = 0;
                wp = tau_window(n, &rcw);
                TRY(rcw);
                if (wp) {
                    TRY(tau_spectra(n, wp, &krt));
                    tau_step = true;
                    ch->host_box = hbox.p + 8 * n;      // this step reports its extent here (ticket n + 1)
                    ch->host_ticket = n + 1;
                }
                else TRY(tau_leave(n));
            } else {
                wp = step_window(n);
            }
            if (!wp) TRY(day_spectra(n, &krt, &krt_t));
            else if (!tau_step) TRY(window_spectra(n, &krt));
            // step n overwrites the state buffer that the emission of day n-2 reads
            if (n >= 3 && !sink) CU(cudaStreamWaitEvent(ctx->stream, ctx->ev_emit[n & 1], 0));
            // whole-torus step followed by another one: its inverse row pass also does the next step's forward row pass
            const bool fuse = !wp && !wmode && n + 1 < nd && fusable(n);
            ch->meta_main = dsm.p + n;          // the step's flag / sums / windows go straight to the day's record
            TRY(chain_conv_main(ch, kern(n), ks->W, krad(n), 1, krt, wp, fuse, fused_m, krt_t, spec_try, tau_step ? tau_event : (cudaEvent_t)nullptr));
            fused_m = fuse ? krad(n) : -1;
            if (tau_step) TRY(tau_after_step(n));
            if (res->pre.p) LAUNCH(ctx, k_copy_domain, D, 256, 0, (const double*)ch->S[ch->cur].p, d, res->pre.p + nD * (n - lead), (const StepMeta*)(dsm.p + n));
            if (sink) {
                // (a few hundred cells: on the chain's own stream, no events -- the likelihood batch is bound by the host's launch rate)
                LAUNCH(ctx, k_emit_dense_cells, sgrid, 256, 0, (const double*)ch->S[ch->cur].p, d, (const StepMeta*)(dsm.p + n), negval,
                       1, 0, sink->cells, sink->K, sink->out + (size_t)sink->K * (n - lead));
                continue;
            }
            // r_small_vals + dense output on the side stream, overlapped with step n+1
            CU(cudaEventRecord(ctx->ev_step[n & 1], ctx->stream));
            CU(cudaStreamWaitEvent(ctx->aux, ctx->ev_step[n & 1], 0));
            LAUNCH_ON(ctx, ctx->aux, k_emit_dense, ctx->emit_ctas > 0 ? std::min(D, ctx->emit_ctas) : D, ctx->emit_ctas > 0 ? 64 : PKB_EMIT_T, 0,
                      (const double*)ch->S[ch->cur].p, d, (const StepMeta*)(dsm.p + n), negval, 1, 0,
                      res->dense.p + nD * (n - lead), a->want_coo ? res->rownnz.p + (size_t)D * (n - lead) : (int*)nullptr, sparse_only ? 1 : 0);
            if (a->want_coo) res->counted[n - lead] = 1;
            CU(cudaEventRecord(ctx->ev_emit[n & 1], ctx->aux));
            TRY(emitted(n, ctx->aux));
        }
        for (int i = 0; i < 2 && !sink; ++i)
            if (nd - 1 - i >= 1) CU(cudaStreamWaitEvent(ctx->stream, ctx->ev_emit[(nd - 1 - i) & 1], 0));
    } else {
        const int rd = a->r_dur;
        const double rn = a->r_number;
        const double* F[PKB_MAX_COHORTS];
        int Wk[PKB_MAX_COHORTS], mm[PKB_MAX_COHORTS];
        cplx* krt[PKB_MAX_COHORTS];
        cplx* krt_t[PKB_MAX_COHORTS];
        bool ready[PKB_MAX_COHORTS], ready_t[PKB_MAX_COHORTS];
        for (int j = 0; j < rd; ++j) {
            F[j] = kern(j); Wk[j] = ks->W; mm[j] = krad(j); ready[j] = false; krt[j] = nullptr; ready_t[j] = false; krt_t[j] = nullptr;
            if (krad(j) > D / 2) return fail(PKB_ELIMIT, "kernel radius %d is larger than the domain radius", krad(j));
        }
        for (int j = 0; j + 1 < rd; ++j) {
            TRY(ch->kcache[j].alloc(ctx, spec_size(d.Nc + 1, d.ldK)));
            krt[j] = ch->kcache[j].p;
            if (ctx->use_trunc_torus) {
                TRY(ch->kcache_t[j].alloc(ctx, spec_size(d.Nc + 1, d.ldK)));
                krt_t[j] = ch->kcache_t[j].p;
            }
        }
        // r_spread[j] as a state (Run.py:469-474); `spread` holds the latest one
        DBuf<double> spread;
        TRY(spread.alloc(ctx, (size_t)d.P * d.ldS));
        CohortArgs ca;
        memset(&ca, 0, sizeof ca);
        // day 0 (CalcSol.py:236-237)
        TRY(set_state_kernel_dev(ch, kern(0), ks->W, krad(0)));
        ca.n = 1; ca.S[0] = ch->S[ch->cur].p; ca.w[0] = a->r_dist[0];
        auto emit_pop = [&](int cday, double centre_extra, int add_centre, int first_day) -> int {
            const int day = cday - lead;
            if (day < 0) return 0;               // the leading spread day is not part of the output
            if (sink)
                LAUNCH(ctx, k_emit_population_cells, sgrid, 256, 0, ca, d, rn, centre_extra, add_centre, negval, first_day, sink->cells, sink->K,
                       sink->out + (size_t)sink->K * day);
            else
                LAUNCH(ctx, k_emit_population, D, 256, 0, ca, d, rn, centre_extra, add_centre, negval, first_day, res->dense.p + nD * day,
                       res->pre.p ? res->pre.p + nD * day : (double*)nullptr);
            return 0;
        };
        TRY(emit_pop(0, rn * (1 - a->r_dist[0]), 1, 1));
        TRY(emitted(0, ctx->stream));
        // release days (CalcSol.py:296-306)
        for (int day = 1; day < rd; ++day) {
            TRY(set_state_kernel_dev(ch, kern(day), ks->W, krad(day)));
            TRY(back_solve_dev(ch, F, Wk, mm, day, krt, ready, krt_t, ready_t));
            TRY(keep_cmeta(day, day));
            double wsum = 0.0;
            for (int c = 0; c <= day; ++c) {
                ca.S[c] = c < day ? ch->coh[c].p : ch->S[ch->cur].p;
                ca.w[c] = a->r_dist[c];
                wsum += a->r_dist[c];
            }
            ca.n = day + 1;
            TRY(emit_pop(day, (1 - wsum) * rn, 1, 0));
            TRY(emitted(day, ctx->stream));
        }
        // Cohort lanes.  After the release the main chain's step n + 1 does not depend on the cohort back-solves of day n (they
        // start from S_n and end in that day's output), and every convolution of an 801^2-sized domain is a one-wave kernel
        // set that leaves SMs idle: the back-solves and the emission of day n run on one of two child contexts (own stream,
        // scratch, cohort buffers and cached filter spectra) while the main chain goes on.  S_n is read by its lane until the
        // lane's event; the main chain waits for it before the step that overwrites that buffer (two days later, same lane).
        struct CohLane {
            pkb_ctx* lc = nullptr;
            pkb_chain* ch = nullptr;
            cplx* krt[PKB_MAX_COHORTS];
            cplx* krt_t[PKB_MAX_COHORTS];
            bool ready[PKB_MAX_COHORTS], ready_t[PKB_MAX_COHORTS];
            bool busy = false;
        } lane[2];
        struct LaneGuard {
            CohLane* l;
            ~LaneGuard() {
                for (int i = 0; i < 2; ++i)
                    if (l[i].ch) {
                        cudaStreamSynchronize(l[i].lc->stream);
                        delete l[i].ch;
                    }
            }
        } lane_guard{lane};
        int nlane = 0;
#ifndef PKB_EMUL
        if (rd > 1 && nd > rd + 1 && ctx->cohort_lanes && !ctx->prof_on && !sink) nlane = 2;
#endif
        for (int i = 0; i < nlane; ++i) {
            while ((int)ctx->lanes.size() <= i) {
                pkb_ctx* lc = nullptr;
                TRY(pkb_create(ctx->device, &lc));
                ctx->lanes.push_back(lc);
            }
            CohLane& ln = lane[i];
            ln.lc = ctx->lanes[i];
            lane_inherit(ctx, ln.lc);
            TRY(chain_create(ln.lc, D, mmax, &ln.ch));
            ln.ch->negval = negval;
            for (int j = 0; j < PKB_MAX_COHORTS; ++j) { ln.krt[j] = nullptr; ln.krt_t[j] = nullptr; ln.ready[j] = false; ln.ready_t[j] = false; }
            for (int j = 0; j + 1 < rd; ++j) {
                TRY(ln.ch->kcache[j].alloc(ln.lc, spec_size(d.Nc + 1, d.ldK)));
                ln.krt[j] = ln.ch->kcache[j].p;
                if (ctx->use_trunc_torus) {
                    TRY(ln.ch->kcache_t[j].alloc(ln.lc, spec_size(d.Nc + 1, d.ldK)));
                    ln.krt_t[j] = ln.ch->kcache_t[j].p;
                }
            }
        }
        // post-release days (CalcSol.py:308-323)
        for (int n = rd; n < nd; ++n) {
            CohLane* ln = nlane ? &lane[n & 1] : nullptr;
            if (ln && ln->busy) CU(cudaStreamWaitEvent(ctx->stream, ln->lc->ev_kr[0], 0));      // day n - 2 is done with the buffer this step writes
            const int* wp = rd == 1 ? step_window(n) : nullptr;
            cplx* kday = nullptr;
            cplx* kday_t = nullptr;
            if (!wp) TRY(day_spectra(n, &kday, &kday_t));
            else TRY(window_spectra(n, &kday));
            const bool fuse = rd == 1 && !wp && !wmode && n + 1 < nd && fusable(n);
            ch->meta_main = dsm.p + n;
            TRY(chain_conv_main(ch, kern(n), ks->W, krad(n), 1, kday, wp, fuse, fused_m, kday_t, spec_try));
            fused_m = fuse ? krad(n) : -1;
            if (ln) {
                pkb_ctx* lc = ln->lc;
                // (the next main step rewrites ch->ctrl[0] while this day's first back-solve step may still be queued: the lane
                // reads a snapshot)
                CU(cudaMemcpyAsync(ln->ch->ctrl.p, ch->ctrl.p, sizeof(ChainCtrl), cudaMemcpyDeviceToDevice, ctx->stream));
                CU(cudaEventRecord(lc->ev_lane, ctx->stream));
                CU(cudaStreamWaitEvent(lc->stream, lc->ev_lane, 0));
                TRY(back_solve_dev(ln->ch, F, Wk, mm, rd - 1, ln->krt, ln->ready, ln->krt_t, ln->ready_t, ch->S[ch->cur].p, ln->ch->ctrl.p));
                if (want_cmeta)
                    CU(cudaMemcpyAsync(dcm.p + (size_t)n * PKB_MAX_COHORTS, ln->ch->meta.p + 1, sizeof(StepMeta) * (rd - 1), cudaMemcpyDeviceToDevice,
                                       lc->stream));
                for (int c = 0; c < rd; ++c) {
                    ca.S[c] = c < rd - 1 ? ln->ch->coh[c].p : ch->S[ch->cur].p;
                    ca.w[c] = a->r_dist[c];
                }
                ca.n = rd;
                const int day = n - lead;
                if (day >= 0)
                    LAUNCH(lc, k_emit_population, D, 256, 0, ca, d, rn, 0.0, 0, negval, 0, res->dense.p + nD * day,
                           res->pre.p ? res->pre.p + nD * day : (double*)nullptr);
                TRY(emitted(n, lc->stream));
                CU(cudaEventRecord(lc->ev_kr[0], lc->stream));
                ln->busy = true;
                continue;
            }
            TRY(back_solve_dev(ch, F, Wk, mm, rd - 1, krt, ready, krt_t, ready_t));
            TRY(keep_cmeta(n, rd - 1));
            for (int c = 0; c < rd; ++c) {
                ca.S[c] = c < rd - 1 ? ch->coh[c].p : ch->S[ch->cur].p;
                ca.w[c] = a->r_dist[c];
            }
            ca.n = rd;
            TRY(emit_pop(n, 0.0, 0, 0));
            TRY(emitted(n, ctx->stream));
        }
        for (int i = 0; i < nlane; ++i)
            if (lane[i].busy) {
                CU(cudaStreamWaitEvent(ctx->stream, lane[i].lc->ev_kr[0], 0));
                ctx->launches += lane[i].lc->launches;
                lane[i].lc->launches = 0;
            }
    }
    if (kr_wait) {      // spectra launched on the side stream and never used: still join before their buffers are released
        CU(cudaStreamWaitEvent(ctx->stream, ctx->ev_kr[1], 0));
        kr_wait = false;
    }
    CU(cudaEventRecord(ctx->ev[2], ctx->stream));
    if (sink) {
        // everything above is ordered on ctx->stream (the side streams were joined): the buffers released
        // on return are reused stream-ordered by the next chain of this context
        TRY(check_launches(ctx, "pkb_solve_batch chain"));
        if (out) *out = nullptr;
        return 0;
    }

    // ---- outputs (the chain is enqueued, not finished: compaction and D2H overlap it) ----
    if (coo_threaded) TRY(worker.join());
    if (a->want_coo) TRY(coo_pump(ctx, res, &coo, true));
    // (pageable destination: this copy blocks the host until the chain has finished, so it comes last)
    CU(cudaMemcpyAsync(res->smeta.data(), dsm.p + lead, sizeof(StepMeta) * nout, cudaMemcpyDeviceToHost, ctx->stream));
    if (want_cmeta) {
        res->cmeta.resize((size_t)nout * PKB_MAX_COHORTS);
        CU(cudaMemcpyAsync(res->cmeta.data(), dcm.p + (size_t)lead * PKB_MAX_COHORTS, sizeof(StepMeta) * nout * PKB_MAX_COHORTS, cudaMemcpyDeviceToHost,
                           ctx->stream));
    }
    CU(cudaEventRecord(ctx->ev[3], ctx->stream));
    TRY(sync_check(ctx, "pkb_solve outputs"));
    float ms = 0.f;
    CU(cudaEventElapsedTime(&ms, ctx->ev[0], ctx->ev[1])); ctx->timing[0] = ms;
    CU(cudaEventElapsedTime(&ms, ctx->ev[1], ctx->ev[2])); ctx->timing[1] = ms;
    CU(cudaEventElapsedTime(&ms, ctx->ev[2], ctx->ev[3])); ctx->timing[2] = ms;
    CU(cudaEventElapsedTime(&ms, ctx->ev[0], ctx->ev[3])); ctx->timing[3] = ms;
    if (!a->keep_dense_device && !a->want_dense_host) res->dense.release();
    rguard.r = nullptr;
    *out = res;
    return 0;
}

// ---------------------------------------------------------------------------
// Batched chains of a likelihood group (bchain.cuh): step n of every proposal in one launch per pass.
//
// Handles what the likelihood batch asks for (sample-cell emission; probability model, or population model with a
// one-day release, Bayes_Run.py's Kalbar setting) when every step of the proposal is an FFT step whose plans run with
// PKB_BT threads; `rest` receives the proposals left for the per-proposal path (solve_chain) -- longer releases, tori too
// large for four resident CTAs, and everything the per-proposal path reports as an error.  (Days whose kernel is small
// enough for the direct stencil take the FFT path here: the same convolution to rounding, and a proposal with one calm
// day stays in the group.)
// The geometry of every step follows solve_chain / conv_step exactly (support windows while the exact support fits the
// domain, then the step's own torus >= P + 2m with the truncated-source torus >= D + 2m beside it), minus the
// spectral-resident steps and the tau windows, which need a host decision per proposal and step.
// Everything is enqueued on lc->stream; nothing here waits for the device.
struct BGeom {
    ChainDims d;
    int mmax;
    std::vector<BStep> steps;      // steps 1 .. nd-1, pointers not yet set
    size_t oS, oYt, oWt, oKrt, oRs;
};
static int solve_chains_batched(pkb_ctx* lc, const pkb_solve_args* sa, int np, pkb_kset* ks, int nk, const int* cells_dev, int K,
                                double* dgroup, std::vector<int>* rest) {
    rest->clear();
    const pkb_solve_args& a0 = sa[0];
    const int lead = a0.sprd ? 1 : 0, nd = a0.ndays + lead, nout = a0.ndays;
    const int D = 2 * ks->rad_res + 1;
    const double negval = a0.negval > 0 ? a0.negval : 1e-8;
    if (!lc->batch_chain || nd < 2 || !(a0.prob_model || a0.r_dur == 1)) {
        for (int p = 0; p < np; ++p) rest->push_back(p);
        return 0;
    }
    std::vector<FftPlan> plans;
    std::map<int, int> plan_of;         // torus side -> index (-1: not usable here)
    auto plan_index = [&](int N) -> int {
        auto it = plan_of.find(N);
        if (it != plan_of.end()) return it->second;
        FftPlan pl;
        int idx = -1;
        if (get_plan(lc, N, &pl) == 0 && pl.grid_rows >= 1 && pl.grid_cols >= 1 && pl.threads == PKB_BT && pl.cols_threads == PKB_BT) {
            idx = (int)plans.size();
            plans.push_back(pl);
        }
        plan_of[N] = idx;
        return idx;
    };
    std::vector<BGeom> geo;
    std::vector<int> who;               // proposal (index in the group) of geo[i]
    geo.reserve(np);
    for (int p = 0; p < np; ++p) {
        const int k0 = p * nk;
        auto krad = [&](int i) { return ks->hmeta[k0 + i].rad; };
        BGeom g;
        bool ok = true;
        int mmax = 0;
        for (int i = 0; i < nd; ++i) mmax = std::max(mmax, krad(i));
        ChainDims& d = g.d;
        memset(&d, 0, sizeof d);
        d.D = D;
        d.P = D + mmax;
        d.N = pkb_smooth_len(std::max(2, d.P + 2 * mmax));
        d.Nc = d.N / 2 + 1;
        d.ldS = roundup(d.P, 16);
        d.ldY = roundup(d.P, 2);
        d.ldW = roundup(d.N, 2);
        d.ldK = roundup(2 * mmax + 1, 2);
        g.mmax = mmax;
        ok = ok && krad(0) <= D / 2 && d.N > 0 && plan_index(d.N) >= 0;
        int wr0 = D / 2 - krad(0), wn = 2 * krad(0) + 1;
        bool wmode = lc->use_windows != 0;
        for (int n = 1; n < nd && ok; ++n) {
            const int m = krad(n);
            if (2 * m > d.P) { ok = false; break; }
            BStep s;
            memset(&s, 0, sizeof s);
            s.m = m;
            s.Wk = ks->W;
            s.d = d;
            if (wmode && wr0 - m >= 0 && wr0 + wn + m <= D && pkb_smooth_len(wn + 2 * m) < d.N) {
                s.d.win = 1; s.d.wr0 = s.d.wc0 = wr0; s.d.wn = wn;
                s.d.N = pkb_smooth_len(std::max(2, wn + 2 * m));
                s.d.Nc = s.d.N / 2 + 1;
                s.d.ldY = roundup(wn, 2);
                s.d.ldW = roundup(s.d.N, 2);
                s.d.ldK = roundup(2 * m + 1, 2);
                wr0 -= m; wn += 2 * m;
                s.plan = s.plan_t = plan_index(s.d.N);
            } else {
                wmode = false;
                if (lc->use_step_torus) {
                    const int Nd = pkb_smooth_len(std::max(2, d.P + 2 * m));
                    if (Nd < d.N && plan_index(Nd) >= 0) {
                        s.d.N = Nd;
                        s.d.Nc = Nd / 2 + 1;
                        s.d.ldW = roundup(Nd, 2);
                    }
                }
                s.plan = s.plan_t = plan_index(s.d.N);
                if (lc->use_trunc_torus) {
                    const int Nt = pkb_smooth_len(std::max(2, D + 2 * m));
                    const int it = Nt < s.d.N ? plan_index(Nt) : -1;
                    if (it >= 0) {
                        s.tg.N = Nt; s.tg.Nc = Nt / 2 + 1; s.tg.ldW = roundup(Nt, 2);
                        s.tg.cols_kb = plans[it].cols_kb;
                        s.plan_t = it;
                    }
                }
            }
            if (s.plan < 0) { ok = false; break; }
            s.cols_kb = plans[s.plan].cols_kb;
            g.steps.push_back(s);
        }
        if (!ok) { rest->push_back(p); continue; }
        geo.push_back(std::move(g));
        who.push_back(p);
    }
    g_err.clear();                      // (plans that could not be made only mean "not here")
    const int nb = (int)geo.size();
    if (nb == 0) return 0;
    CU(cudaEventRecord(lc->ev[1], lc->stream));

    // one slab per kind of buffer, every proposal at its own offset (32-byte aligned: the spectra are accessed as 32-byte pairs)
    size_t tS = 0, tYt = 0, tWt = 0, tKrt = 0, tRs = 0;
    int mmax_all = 0;
    for (BGeom& g : geo) {
        const ChainDims& d = g.d;
        g.oS = tS;   tS += 2 * (size_t)d.P * d.ldS;
        g.oYt = tYt; tYt += spec_size(d.Nc + 1, d.ldY);       // (multiples of PKB_CB elements)
        g.oWt = tWt; tWt += spec_size(d.Nc + 1, d.ldW);
        g.oKrt = tKrt; tKrt += spec_size(d.Nc + 1, d.ldK);
        g.oRs = tRs; tRs += (size_t)d.P;
        mmax_all = std::max(mmax_all, g.mmax);
    }
    DBuf<double> S;
    DBuf<cplx> Yt, Wt, Krt, scr;
    DBuf<RowStats> rstat;
    DBuf<ChainCtrl> ctrl;
    DBuf<StepMeta> dsm;
    DBuf<int> hint;
    TRY(hint.alloc(lc, nb));
    CU(cudaMemsetAsync(hint.p, 0, sizeof(int) * nb, lc->stream));
    TRY(S.alloc(lc, tS));
    TRY(Yt.alloc(lc, tYt));
    TRY(Wt.alloc(lc, tWt));
    TRY(Krt.alloc(lc, tKrt));
    TRY(rstat.alloc(lc, tRs));
    TRY(ctrl.alloc(lc, nb));
    TRY(dsm.alloc(lc, (size_t)nb * nd));
    CU(cudaMemsetAsync(S.p, 0, tS * sizeof(double), lc->stream));
    CU(cudaMemsetAsync(rstat.p, 0, tRs * sizeof(RowStats), lc->stream));       // rows a windowed step never touches are zero rows
    CU(cudaMemsetAsync(ctrl.p, 0, sizeof(ChainCtrl) * nb, lc->stream));
    CU(cudaMemsetAsync(dsm.p, 0, sizeof(StepMeta) * nb * nd, lc->stream));

    // descriptors: steps [n - 1][i], job tables [n - 1][pass][i], emissions [n][i], day-0 placement [i]
    const int ns = nd - 1;
    std::vector<BStep> hsteps((size_t)ns * nb);
    const int jstride = 2 * (nb + 1) + (2 * nb + 1);      // per step: forward rows [nb + 1], columns [nb + 1], inverse rows [2 nb + 1] (two passes)
    std::vector<int> hjobs((size_t)ns * jstride, 0);
    std::vector<BEmit> hemit((size_t)nd * nb);
    std::vector<BInit> hinit(nb);
    for (int i = 0; i < nb; ++i) {
        const BGeom& g = geo[i];
        const pkb_solve_args& a = sa[who[i]];
        const ChainDims& d = g.d;
        const int k0 = who[i] * nk;
        const size_t nW = (size_t)ks->W * ks->W;
        double* Sb[2] = {S.p + g.oS, S.p + g.oS + (size_t)d.P * d.ldS};
        double* outp = dgroup + (size_t)who[i] * nout * K;
        const double w0 = (!a.prob_model && a.r_dist) ? a.r_dist[0] : 1.0;
        BInit& bi = hinit[i];
        bi.K = ks->acc.p + nW * k0; bi.S = Sb[0]; bi.ctrl = ctrl.p + i; bi.Wk = ks->W; bi.m = ks->hmeta[k0].rad; bi.ldS = d.ldS; bi.D = D;
        BEmit& e0 = hemit[i];
        e0.S = Sb[0]; e0.meta = dsm.p + (size_t)i * nd; e0.out = outp; e0.ldS = d.ldS;
        e0.mode = lead ? -1 : (a.prob_model ? 0 : 2);
        e0.w0 = w0; e0.centre_extra = a.r_number * (1 - w0);
        int cur = 0;
        for (int n = 1; n < nd; ++n) {
            BStep s = g.steps[n - 1];
            s.K = ks->acc.p + nW * (k0 + n);
            s.src = Sb[cur]; s.dst = Sb[cur ^ 1];
            cur ^= 1;
            s.Yt = Yt.p + g.oYt; s.Wt = Wt.p + g.oWt; s.Krt = Krt.p + g.oKrt;
            s.rstat = rstat.p + g.oRs; s.ctrl = ctrl.p + i; s.meta = dsm.p + (size_t)i * nd + n; s.hint = hint.p + i;
            hsteps[(size_t)(n - 1) * nb + i] = s;
            int* jt = hjobs.data() + (size_t)(n - 1) * jstride;
            const int rows_in = s.d.win ? s.d.wn : d.P;
            const int jf = s.m + 1 + (rows_in + 1) / 2;
            const int jc = s.d.Nc;
            // inverse rows: first pass every job of the step's own geometry (on the truncated-source torus only the (D + 1) / 2 row
            // pairs that start inside the domain), second pass the rest of the truncated-source torus (kb_rows_inv)
            const int ndom_t = (d.D + 1) / 2;
            const int ji = s.d.win ? (s.d.wn + 2 * s.m + 1) / 2 : std::max(rows_inv_jobs(d.P, s.m), s.tg.N ? ndom_t : 0);
            const int ji2 = (!s.d.win && s.tg.N) ? std::max(0, rows_inv_jobs_trunc(d.P, d.D, s.m) - ndom_t) : 0;
            jt[0 * (nb + 1) + i + 1] = jf;
            jt[1 * (nb + 1) + i + 1] = jc;
            jt[2 * (nb + 1) + i + 1] = ji;
            jt[2 * (nb + 1) + nb + i + 1] = ji2;
            BEmit& e = hemit[(size_t)n * nb + i];
            e.S = s.dst; e.meta = s.meta; e.ldS = d.ldS;
            e.mode = n < lead ? -1 : (a.prob_model ? 1 : 3);
            e.out = outp + (size_t)std::max(0, n - lead) * K;
            e.w0 = w0; e.centre_extra = 0.0;
        }
    }
    for (int n = 0; n < ns; ++n) {
        int* jt = hjobs.data() + (size_t)n * jstride;
        for (int k = 0; k < 2; ++k)
            for (int i = 0; i < nb; ++i) jt[k * (nb + 1) + i + 1] += jt[k * (nb + 1) + i];
        for (int i = 0; i < 2 * nb; ++i) jt[2 * (nb + 1) + i + 1] += jt[2 * (nb + 1) + i];
    }
    DBuf<BStep> dsteps;
    DBuf<int> djobs;
    DBuf<BEmit> demit;
    DBuf<BInit> dinit;
    DBuf<FftPlan> dplans;
    TRY(dsteps.alloc(lc, hsteps.size()));
    TRY(djobs.alloc(lc, hjobs.size()));
    TRY(demit.alloc(lc, hemit.size()));
    TRY(dinit.alloc(lc, hinit.size()));
    TRY(dplans.alloc(lc, plans.size()));
    // (pageable sources: the copies are staged before these calls return)
    CU(cudaMemcpyAsync(dsteps.p, hsteps.data(), sizeof(BStep) * hsteps.size(), cudaMemcpyHostToDevice, lc->stream));
    CU(cudaMemcpyAsync(djobs.p, hjobs.data(), sizeof(int) * hjobs.size(), cudaMemcpyHostToDevice, lc->stream));
    CU(cudaMemcpyAsync(demit.p, hemit.data(), sizeof(BEmit) * hemit.size(), cudaMemcpyHostToDevice, lc->stream));
    CU(cudaMemcpyAsync(dinit.p, hinit.data(), sizeof(BInit) * hinit.size(), cudaMemcpyHostToDevice, lc->stream));
    CU(cudaMemcpyAsync(dplans.p, plans.data(), sizeof(FftPlan) * plans.size(), cudaMemcpyHostToDevice, lc->stream));

    // persistent grids at the largest footprint of the group
    size_t smem = 1024, scr_per_cta = 0;
    for (const FftPlan& pl : plans) {
        smem = std::max(smem, fft_smem_bytes(pl));
        scr_per_cta = std::max(scr_per_cta, (size_t)pl.cols_kb * plan_radix(pl, pl.nstage - 1) * PKB_BT);
    }
    int occ_f = 0, occ_c = 0, occ_i = 0;
    CU(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ_f, kb_rows_fwd, PKB_BT, smem));
    CU(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ_c, kb_cols, PKB_BT, smem));
    CU(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ_i, kb_rows_inv, PKB_BT, smem));
    if (occ_f < 1 || occ_c < 1 || occ_i < 1) return fail(PKB_ELIMIT, "batched chain kernels cannot be resident with %zu bytes of shared memory", smem);
    const int cap = std::max(lc->batch_occ, 1);
    const int gmax_f = std::min(occ_f, cap) * lc->sm_count, gmax_c = std::min(occ_c, cap) * lc->sm_count, gmax_i = std::min(occ_i, cap) * lc->sm_count;
    TRY(scr.alloc(lc, (size_t)gmax_c * scr_per_cta));

    if (getenv("PKB_BCHAIN_DEBUG")) {
        long long jf = 0, jc = 0, ji = 0, wsteps = 0, sumN = 0, sumP = 0, summ = 0;
        for (int n = 0; n < ns; ++n) {
            const int* hj = hjobs.data() + (size_t)n * jstride;
            jf += hj[nb]; jc += hj[(nb + 1) + nb]; ji += hj[2 * (nb + 1) + 2 * nb];
        }
        for (const BStep& s : hsteps) { wsteps += s.d.win; sumN += s.d.N; sumP += s.d.P; summ += s.m; }
        fprintf(stderr, "bchain: %d of %d proposals, %d steps each (%lld window steps), mean N %.0f P %.0f m %.0f, %zu plans, smem %zu, occ %d/%d/%d (cap %d), "
                "jobs fwd %lld cols %lld inv %lld, slabs S %.0f MB Yt %.0f Wt %.0f Krt %.0f\n", nb, np, ns, wsteps, (double)sumN / hsteps.size(),
                (double)sumP / hsteps.size(), (double)summ / hsteps.size(), plans.size(), smem, occ_f, occ_c, occ_i, cap, jf, jc, ji, tS * 8e-6, tYt * 16e-6,
                tWt * 16e-6, tKrt * 16e-6);
    }
    const double rn = a0.r_number;
    LAUNCH(lc, kb_init, dim3(2 * mmax_all + 1, nb), 128, 0, (const BInit*)dinit.p);
    LAUNCH(lc, kb_finish, nb, PKB_BT, 0, (const BStep*)nullptr, (const BEmit*)demit.p, cells_dev, K, D, rn, negval);
    for (int n = 1; n < nd; ++n) {
        const BStep* st = dsteps.p + (size_t)(n - 1) * nb;
        const int* jt = djobs.p + (size_t)(n - 1) * jstride;
        const int* hj = hjobs.data() + (size_t)(n - 1) * jstride;
        LAUNCH(lc, kb_rows_fwd, std::min(hj[0 * (nb + 1) + nb], gmax_f), PKB_BT, smem, st, (const FftPlan*)dplans.p, jt, nb);
        LAUNCH(lc, kb_cols, std::min(hj[1 * (nb + 1) + nb], gmax_c), PKB_BT, smem, st, (const FftPlan*)dplans.p, jt + (nb + 1), nb, scr.p, scr_per_cta);
        LAUNCH(lc, kb_rows_inv, std::min(hj[2 * (nb + 1) + 2 * nb], gmax_i), PKB_BT, smem, st, (const FftPlan*)dplans.p, jt + 2 * (nb + 1), nb, negval);
        LAUNCH(lc, kb_finish, nb, PKB_BT, 0, st, (const BEmit*)(demit.p + (size_t)n * nb), cells_dev, K, D, rn, negval);
    }
    CU(cudaEventRecord(lc->ev[2], lc->stream));
    return check_launches(lc, "pkb_solve_batch batched chains");
}

extern "C" int pkb_solve(pkb_ctx* ctx, const pkb_solve_args* a, pkb_result** out) {
    if (!ctx || !a || !out || !a->wind) return fail(PKB_EINVAL, "pkb_solve: NULL argument");
    *out = nullptr;
    TRY(check_solve_args(a));
    CU(cudaSetDevice(ctx->device));
    CU(cudaEventRecord(ctx->ev[0], ctx->stream));

    // ---- phase 1 (Run.py:412-425) ------------------------------------------
    DBuf<double> dwind;
    const double* wind_dev = a->wind;
    if (!a->wind_on_device) {
        const size_t nw = (size_t)a->nd_wind * a->periods * 3;
        TRY(dwind.alloc(ctx, nw));
        CU(cudaMemcpyAsync(dwind.p, a->wind, nw * sizeof(double), cudaMemcpyHostToDevice, ctx->stream));
        wind_dev = dwind.p;
    }
    std::vector<pkb_day_args> dargs(solve_nkernels(a));
    solve_day_args(a, dargs.data());
    pkb_kset* ks = nullptr;
    TRY(kernels_build_dev(ctx, wind_dev, a->nd_wind, a->periods, dargs.data(), (int)dargs.size(), 0, &ks));
    struct KGuard {
        pkb_kset* k;
        ~KGuard() { delete k; }
    } kguard{ks};
    return solve_chain(ctx, a, ks, 0, out);
}

// ---- likelihood projection (Bayes_funcs.py) -----------------------------------------------------------------
struct DevProjection {
    DBuf<int> set_ptr, set_cells, row_ptr, grp_ptr, term_day, term_set;
    DBuf<double> term_w;
    ProjTables t;
};
static int upload_projection(pkb_ctx* ctx, const pkb_projection* p, int ndays, int K, DevProjection* d) {
    if (!p || !p->set_ptr || !p->row_ptr || !p->grp_ptr) return fail(PKB_EINVAL, "projection: NULL table");
    if (p->nsets < 0 || p->nrows < 1 || p->ngroups < 0) return fail(PKB_EINVAL, "projection: bad sizes");
    const int ncell = p->set_ptr[p->nsets], nterm = p->grp_ptr[p->ngroups];
    if (p->row_ptr[p->nrows] != p->ngroups) return fail(PKB_EINVAL, "projection: row_ptr does not cover the groups");
    if ((ncell > 0 && !p->set_cells) || (nterm > 0 && (!p->term_day || !p->term_set || !p->term_w))) return fail(PKB_EINVAL, "projection: NULL table");
    for (int i = 0; i < p->nsets; ++i)
        if (p->set_ptr[i + 1] < p->set_ptr[i]) return fail(PKB_EINVAL, "projection: set_ptr is not monotone");
    for (int i = 0; i < ncell; ++i)
        if (p->set_cells[i] < 0 || p->set_cells[i] >= K) return fail(PKB_EINVAL, "projection: sample-cell index %d out of range", p->set_cells[i]);
    for (int i = 0; i < p->nrows; ++i)
        if (p->row_ptr[i + 1] < p->row_ptr[i]) return fail(PKB_EINVAL, "projection: row_ptr is not monotone");
    for (int i = 0; i < p->ngroups; ++i)
        if (p->grp_ptr[i + 1] < p->grp_ptr[i]) return fail(PKB_EINVAL, "projection: grp_ptr is not monotone");
    for (int i = 0; i < nterm; ++i) {
        if (p->term_day[i] < 0 || p->term_day[i] >= ndays) return fail(PKB_EINVAL, "projection: model day %d out of range [0, %d)", p->term_day[i], ndays);
        if (p->term_set[i] < 0 || p->term_set[i] >= p->nsets) return fail(PKB_EINVAL, "projection: set index %d out of range", p->term_set[i]);
    }
    auto up_i = [&](DBuf<int>& b, const int* src, size_t n) -> int {
        TRY(b.alloc(ctx, n));
        if (n) CU(cudaMemcpyAsync(b.p, src, n * sizeof(int), cudaMemcpyHostToDevice, ctx->stream));
        return 0;
    };
    TRY(up_i(d->set_ptr, p->set_ptr, (size_t)p->nsets + 1));
    TRY(up_i(d->set_cells, p->set_cells, ncell));
    TRY(up_i(d->row_ptr, p->row_ptr, (size_t)p->nrows + 1));
    TRY(up_i(d->grp_ptr, p->grp_ptr, (size_t)p->ngroups + 1));
    TRY(up_i(d->term_day, p->term_day, nterm));
    TRY(up_i(d->term_set, p->term_set, nterm));
    TRY(d->term_w.alloc(ctx, nterm));
    if (nterm) CU(cudaMemcpyAsync(d->term_w.p, p->term_w, (size_t)nterm * sizeof(double), cudaMemcpyHostToDevice, ctx->stream));
    CU(cudaStreamSynchronize(ctx->stream));      // the caller's (pageable) tables may go away after this call
    d->t.set_ptr = d->set_ptr.p; d->t.set_cells = d->set_cells.p; d->t.row_ptr = d->row_ptr.p; d->t.grp_ptr = d->grp_ptr.p;
    d->t.term_day = d->term_day.p; d->t.term_set = d->term_set.p; d->t.term_w = d->term_w.p; d->t.nrows = p->nrows;
    return 0;
}
static void launch_project(pkb_ctx* ctx, const DevProjection& d, const double* samples, int nd, int K, int nprop, double* out) {
    const long long n = (long long)nprop * d.t.nrows;
    LAUNCH(ctx, k_project, (unsigned)((n + 127) / 128), 128, 0, d.t, samples, nd, K, nprop, out);
}

extern "C" int pkb_project(pkb_ctx* ctx, const pkb_projection* proj, const double* samples, int nprop, int ndays, int K, double* out) {
    if (!ctx || !proj || !samples || !out) return fail(PKB_EINVAL, "pkb_project: NULL argument");
    if (nprop < 1 || ndays < 1 || K < 1) return fail(PKB_EINVAL, "pkb_project: bad sizes");
    CU(cudaSetDevice(ctx->device));
    DevProjection dp;
    TRY(upload_projection(ctx, proj, ndays, K, &dp));
    DBuf<double> ds, dout;
    const size_t ns = (size_t)nprop * ndays * K, no = (size_t)nprop * proj->nrows;
    TRY(ds.alloc(ctx, ns));
    TRY(dout.alloc(ctx, no));
    CU(cudaMemcpyAsync(ds.p, samples, ns * sizeof(double), cudaMemcpyHostToDevice, ctx->stream));
    launch_project(ctx, dp, ds.p, ndays, K, nprop, dout.p);
    CU(cudaMemcpyAsync(out, dout.p, no * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
    return sync_check(ctx, "pkb_project");
}

// Likelihood batch: `nprop` independent proposals of the same solve (shared wind, domain and release
// settings; the 15 block variables of Bayes_Run.py:186-196 differ), returning the model at K sample
// cells for every proposal and day.  Kernel construction is batched over groups of proposals (one
// launch set for up to PKB_BATCH_GROUP x ndays (proposal, day) problems instead of one per proposal).
// The chains of a group are then enqueued round-robin on `batch_lanes` child contexts (own streams,
// buffer pools and plans): a Kalbar-sized chain step is a handful of short persistent kernels whose
// last wave leaves most SMs idle, and a second chain in flight fills those tails.  Every chain emits
// only the K sample cells (SampleSink); the host synchronises once per group.
#define PKB_BATCH_GROUP 32
static int solve_batch_impl(pkb_ctx* ctx, const pkb_solve_args* base, const double* proposals, int nprop, const int* cells, int K,
                            const pkb_projection* proj, double* out, int* status) {
    if (!ctx || !base || !base->wind || !proposals || !cells || !out) return fail(PKB_EINVAL, "pkb_solve_batch: NULL argument");
    if (nprop < 0 || K < 1) return fail(PKB_EINVAL, "pkb_solve_batch: bad sizes");
    TRY(check_solve_args(base));
    CU(cudaSetDevice(ctx->device));
    const int nd = base->ndays;
    const size_t dom = 2 * (size_t)base->day.rad_res + 1;
    for (int k = 0; k < 2 * K; ++k)
        if (cells[k] < 0 || cells[k] >= (int)dom) return fail(PKB_EINVAL, "pkb_solve_batch: cell index %d outside the domain", cells[k]);
    DBuf<double> dwind;
    const double* wind_dev = base->wind;
    if (!base->wind_on_device) {
        const size_t nw = (size_t)base->nd_wind * base->periods * 3;
        TRY(dwind.alloc(ctx, nw));
        CU(cudaMemcpyAsync(dwind.p, base->wind, nw * sizeof(double), cudaMemcpyHostToDevice, ctx->stream));
        wind_dev = dwind.p;
    }
    DBuf<int> dcells;
    TRY(dcells.alloc(ctx, 2 * (size_t)K));
    CU(cudaMemcpyAsync(dcells.p, cells, sizeof(int) * 2 * K, cudaMemcpyHostToDevice, ctx->stream));
    // optional likelihood projection of every proposal's samples, on the device (Bayes_funcs.py)
    DevProjection dproj;
    DBuf<double> dprojout;
    if (proj) TRY(upload_projection(ctx, proj, nd, K, &dproj));
    // lanes: child contexts on the same device, created on first use and kept
    const int nlanes = std::max(1, std::min(ctx->batch_lanes, nprop));
    while ((int)ctx->lanes.size() < nlanes) {
        pkb_ctx* lane = nullptr;
        TRY(pkb_create(ctx->device, &lane));
        ctx->lanes.push_back(lane);
    }
    for (int l = 0; l < nlanes; ++l) {
        lane_inherit(ctx, ctx->lanes[l]);
    }
    // after an error or at the end of a group: drain the lanes, fold their launch counts and per-kernel
    // profile into the parent (what pkb_launch_count / pkb_profile_get report)
    auto drain = [&]() -> int {
        int rc = 0;
        for (int l = 0; l < nlanes; ++l) {
            pkb_ctx* lane = ctx->lanes[l];
            const int r1 = sync_check(lane, "pkb_solve_batch lane");
            if (r1 && !rc) rc = r1;
            // pkb_timing: chain phase of the last chain of lane 0, kernel construction of the last group
            float ms = 0.f;
            if (l == 0 && !r1 && cudaEventElapsedTime(&ms, lane->ev[1], lane->ev[2]) == cudaSuccess) {
                ctx->timing[1] = ms;
                ctx->timing[2] = 0.0;
                if (cudaEventElapsedTime(&ms, ctx->ev[0], ctx->ev[1]) == cudaSuccess) ctx->timing[0] = ms;
                ctx->timing[3] = ctx->timing[0] + ctx->timing[1];
            } else {
                cudaGetLastError();      // (an unrecorded event is not a launch failure)
            }
            ctx->launches += lane->launches;
            lane->launches = 0;
            for (auto& kv : lane->prof_acc) {
                auto& acc = ctx->prof_acc[kv.first];
                acc.first += kv.second.first;
                acc.second += kv.second.second;
            }
            lane->prof_acc.clear();
        }
        return rc;
    };
    // bound the group so that the accumulation windows (worst case the whole domain per problem) stay below ~8 GB
    int group = (int)std::max<size_t>(1, std::min<size_t>(ctx->batch_group, ((size_t)8 << 30) / (2 * dom * dom * sizeof(double) * (nd + 1))));
    // Groups are pipelined: the kernels of group g+1 are built on the parent's stream (and its small
    // sizing D2H waited for) while the lanes still run the chains of group g; only then are the lanes
    // drained and group g's samples copied out.  Two kernel sets and two output buffers are alive at a time.
    DBuf<double> dout[2];
    const size_t out_group = (size_t)std::min(group, std::max(nprop, 1)) * nd * K;
    TRY(dout[0].alloc(ctx, out_group));
    if (nprop > group) TRY(dout[1].alloc(ctx, out_group));
    if (proj) TRY(dprojout.alloc(ctx, (size_t)std::min(group, std::max(nprop, 1)) * proj->nrows));
    struct Pending {
        pkb_kset* ks = nullptr;
        int p0 = 0, np = 0, buf = 0;
        int rc = 0;           // first error while its chains were enqueued
        bool live = false;
        // the group's chains are enqueued by one host thread per lane; what they read lives here until they are joined
        std::vector<std::thread> workers;
        std::shared_ptr<std::vector<pkb_solve_args> > sa;
        std::mutex mu;
        std::string msg;
    } pend;
    auto join_workers = [&]() {
        for (auto& t : pend.workers)
            if (t.joinable()) t.join();
        pend.workers.clear();
        if (pend.rc && !pend.msg.empty()) g_err = pend.msg;
    };
    // wait for the group in flight, copy its samples out, release its kernels
    auto finish = [&]() -> int {
        if (!pend.live) return 0;
        join_workers();
        int rc = drain();
        if (pend.rc) {
            rc = pend.rc;
            if (!pend.msg.empty()) g_err = pend.msg;
        }
        if (!rc) {
            cudaError_t e;
            const cudaMemcpyKind kind = base->out_on_device ? cudaMemcpyDeviceToDevice : cudaMemcpyDeviceToHost;
            if (proj) {
                launch_project(ctx, dproj, dout[pend.buf].p, nd, K, pend.np, dprojout.p);
                e = cudaMemcpyAsync(out + (size_t)pend.p0 * proj->nrows, dprojout.p, sizeof(double) * pend.np * proj->nrows, kind, ctx->stream);
            } else
                e = cudaMemcpyAsync(out + (size_t)pend.p0 * nd * K, dout[pend.buf].p, sizeof(double) * pend.np * nd * K, kind, ctx->stream);
            if (e != cudaSuccess) rc = fail(PKB_ECUDA, "pkb_solve_batch: output copy failed: %s", cudaGetErrorString(e));
            else rc = sync_check(ctx, "pkb_solve_batch outputs");
        }
        delete pend.ks;
        pend.ks = nullptr; pend.p0 = pend.np = pend.buf = 0; pend.rc = 0; pend.live = false; pend.sa.reset(); pend.msg.clear();
        return rc;
    };
    int gi = 0;
    for (int p0 = 0; p0 < nprop; p0 += group, ++gi) {
        const int np = std::min(group, nprop - p0);
        auto sa_ptr = std::make_shared<std::vector<pkb_solve_args> >(np, *base);
        std::vector<pkb_solve_args>& sa = *sa_ptr;
        const int nk = solve_nkernels(base), lead = nk - nd;      // kernels per proposal (leading spread kernel, Bayes_Run.py:245-270)
        std::vector<pkb_day_args> dargs((size_t)np * nk);
        for (int p = 0; p < np; ++p) {
            const double* q = proposals + 15 * (size_t)(p0 + p);      // g_aw g_bw f_a1 f_b1 f_a2 f_b2 sig_x sig_y corr sig_xl sig_yl corr_l lam n_periods mu_r
            pkb_solve_args& s = sa[p];
            s.wind = wind_dev;
            s.wind_on_device = 1;
            s.day.hparams[0] = q[12];
            for (int j = 0; j < 6; ++j) s.day.hparams[1 + j] = q[j];
            for (int j = 0; j < 3; ++j) { s.day.dparams[j] = q[6 + j]; s.day.dlparams[j] = q[9 + j]; }
            s.day.n_periods = (int)llround(q[13]);
            s.day.mu_r = q[14];
            if (base->sprd && base->sprd_factors) {
                s.sprd_factor = base->sprd_factors[p0 + p];
                if (!(s.sprd_factor >= 0 && s.sprd_factor <= 1)) return fail(PKB_EINVAL, "pkb_solve_batch: sprd_factor of proposal %d outside [0, 1]", p0 + p);
            }
            s.want_coo = 0; s.want_dense_host = 0; s.keep_dense_device = 0;
            solve_day_args(&s, dargs.data() + (size_t)p * nk);
        }
        cudaEventRecord(ctx->ev[0], ctx->stream);
        pkb_kset* ks = nullptr;
        const int rcb = kernels_build_dev(ctx, wind_dev, base->nd_wind, base->periods, dargs.data(), np * nk, 0, &ks);
        if (rcb) {
            const std::string msg = g_err;      // (finish() may overwrite the message)
            finish();
            g_err = msg;
            return rcb;
        }
        if (status)
            for (int p = 0; p < np; ++p)
                for (int i = 0; i < nd; ++i) status[(size_t)(p0 + p) * nd + i] = ks->hmeta[(size_t)p * nk + lead + i].status;
        cudaEventRecord(ctx->ev[1], ctx->stream);
        const int rcf = finish();              // the previous group (its chains overlapped the kernel construction above)
        if (rcf) { delete ks; return rcf; }
        // the lanes start once the kernels (and, first group, the cells) are on the device
        cudaEventRecord(ctx->ev_lane, ctx->stream);
        for (int l = 0; l < nlanes; ++l) cudaStreamWaitEvent(ctx->lanes[l]->stream, ctx->ev_lane, 0);
        pend.ks = ks; pend.p0 = p0; pend.np = np; pend.buf = gi & 1; pend.live = true; pend.sa = sa_ptr;
        double* dgroup = dout[pend.buf].p;
        const int* dcell_p = dcells.p;
        // step n of every proposal in one launch per pass (bchain.cuh), on lane 0; what that path does not take is left in `rest`
        auto rest_ptr = std::make_shared<std::vector<int> >();
        {
            const int rcc = solve_chains_batched(ctx->lanes[0], sa.data(), np, ks, nk, dcell_p, K, dgroup, rest_ptr.get());
            if (rcc) {
                const std::string msg = g_err;
                finish();
                g_err = msg;
                return rcc;
            }
        }
        const int nrest = (int)rest_ptr->size();
        auto lane_work = [&pend, ctx, ks, sa_ptr, rest_ptr, nrest, nk, nd, K, nlanes, dgroup, dcell_p](int l) {
            cudaSetDevice(ctx->device);                 // (the current device is per host thread)
            for (int ir = l; ir < nrest; ir += nlanes) {
                const int p = (*rest_ptr)[ir];
                { std::lock_guard<std::mutex> g(pend.mu); if (pend.rc) return; }
                SampleSink sink = {dcell_p, K, dgroup + (size_t)p * nd * K};
                const int rc = solve_chain(ctx->lanes[l], &(*sa_ptr)[p], ks, p * nk, nullptr, &sink);
                if (rc) {
                    std::lock_guard<std::mutex> g(pend.mu);
                    if (!pend.rc) { pend.rc = rc; pend.msg = g_err; }
                    return;
                }
            }
        };
#ifdef PKB_EMUL
        const bool threaded = false;                    // (the CPU emulation of CUDA blocks is not re-entrant)
#else
        const bool threaded = ctx->batch_threads && nlanes > 1 && nrest > 1;
#endif
        if (threaded) {
            // the workers keep enqueueing while this thread goes on to build the next group's kernels; finish() joins them
            for (int l = 0; l < nlanes; ++l) pend.workers.emplace_back(lane_work, l);
        } else {
            for (int ir = 0; ir < nrest && !pend.rc; ++ir) {
                const int p = (*rest_ptr)[ir];
                SampleSink sink = {dcell_p, K, dgroup + (size_t)p * nd * K};
                pend.rc = solve_chain(ctx->lanes[ir % nlanes], &sa[p], ks, p * nk, nullptr, &sink);
                if (pend.rc) pend.msg = g_err;
            }
            if (pend.rc) {
                const int rc = pend.rc;
                finish();
                return rc;
            }
        }
    }
    return finish();
}

extern "C" int pkb_solve_batch(pkb_ctx* ctx, const pkb_solve_args* base, const double* proposals, int nprop, const int* cells, int K,
                               double* out, int* status) {
    return solve_batch_impl(ctx, base, proposals, nprop, cells, K, nullptr, out, status);
}

extern "C" int pkb_solve_batch_projected(pkb_ctx* ctx, const pkb_solve_args* base, const double* proposals, int nprop, const int* cells,
                                         int K, const pkb_projection* proj, double* out, int* status) {
    if (!proj) return fail(PKB_EINVAL, "pkb_solve_batch_projected: NULL projection");
    return solve_batch_impl(ctx, base, proposals, nprop, cells, K, proj, out, status);
}

extern "C" int pkb_result_info(pkb_result* r, int* ndays, int* dom_len, int* P, int* N, int* max_shape) {
    if (!r) return fail(PKB_EINVAL, "pkb_result_info: NULL result");
    if (ndays) *ndays = r->ndays;
    if (dom_len) *dom_len = r->D;
    if (P) *P = r->P;
    if (N) *N = r->N;
    if (max_shape) *max_shape = r->max_shape;
    return 0;
}

extern "C" int pkb_result_window_steps(pkb_result* r, int* n) {
    if (!r || !n) return fail(PKB_EINVAL, "pkb_result_window_steps: NULL argument");
    *n = r->window_steps;
    return 0;
}

extern "C" int pkb_result_day_meta(pkb_result* r, int day, pkb_day_meta* kmeta, pkb_step_meta* smeta) {
    if (!r || day < 0 || day >= r->ndays) return fail(PKB_EINVAL, "pkb_result_day_meta: bad argument");
    if (kmeta) memcpy(kmeta, &r->kmeta[day], sizeof(DayMeta));
    if (smeta) memcpy(smeta, &r->smeta[day], sizeof(StepMeta));
    return 0;
}

extern "C" int pkb_result_cohort_meta(pkb_result* r, int day, int cohort, pkb_step_meta* smeta) {
    if (!r || !smeta || day < 0 || day >= r->ndays || cohort < 0 || cohort >= PKB_MAX_COHORTS)
        return fail(PKB_EINVAL, "pkb_result_cohort_meta: bad argument");
    if (r->cmeta.empty()) memset(smeta, 0, sizeof(StepMeta));
    else memcpy(smeta, &r->cmeta[(size_t)day * PKB_MAX_COHORTS + cohort], sizeof(StepMeta));
    return 0;
}

extern "C" int pkb_result_dense(pkb_result* r, int day, double* out) {
    if (!r || !out || day < 0 || day >= r->ndays) return fail(PKB_EINVAL, "pkb_result_dense: bad argument");
    if (!r->dense.p) return fail(PKB_ESTATE, "pkb_result_dense: dense solutions were not kept (want_dense_host / keep_dense_device)");
    pkb_ctx* ctx = r->ctx;
    CU(cudaSetDevice(ctx->device));
    const size_t nD = (size_t)r->D * r->D;
    CU(cudaMemcpyAsync(out, r->dense.p + nD * day, nD * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
    return sync_check(ctx, "pkb_result_dense");
}

extern "C" int pkb_result_pre(pkb_result* r, int day, double* out) {
    if (!r || !out || day < 0 || day >= r->ndays) return fail(PKB_EINVAL, "pkb_result_pre: bad argument");
    if (!r->pre.p) return fail(PKB_ESTATE, "pkb_result_pre: the solve was run without keep_pre_device");
    pkb_ctx* ctx = r->ctx;
    CU(cudaSetDevice(ctx->device));
    const size_t nD = (size_t)r->D * r->D;
    CU(cudaMemcpyAsync(out, r->pre.p + nD * day, nD * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
    return sync_check(ctx, "pkb_result_pre");
}

extern "C" int pkb_result_csr(pkb_result* r, const long long** day_offsets, const long long** row_offsets, const int** cols, const double** vals) {
    if (!r) return fail(PKB_EINVAL, "pkb_result_csr: NULL result");
    if (!r->have_coo || !r->csr) return fail(PKB_ESTATE, "pkb_result_csr: solve was run without want_coo = 2");
    if (day_offsets) *day_offsets = r->dayoff.p;
    if (row_offsets) *row_offsets = r->rowoff_host.p;
    if (cols) *cols = r->cols.p;
    if (vals) *vals = r->vals.p;
    return 0;
}

extern "C" int pkb_result_coo(pkb_result* r, const long long** day_offsets, const int** rows, const int** cols, const double** vals) {
    if (!r) return fail(PKB_EINVAL, "pkb_result_coo: NULL result");
    if (!r->have_coo || r->csr) return fail(PKB_ESTATE, "pkb_result_coo: solve was run without want_coo = 1");
    if (day_offsets) *day_offsets = r->dayoff.p;
    if (rows) *rows = r->rows.p;
    if (cols) *cols = r->cols.p;
    if (vals) *vals = r->vals.p;
    return 0;
}

extern "C" int pkb_result_sample(pkb_result* r, const int* cells, int K, double* out) {
    if (!r || !cells || !out || K < 1) return fail(PKB_EINVAL, "pkb_result_sample: bad argument");
    if (!r->dense.p) return fail(PKB_ESTATE, "pkb_result_sample: dense solutions were not kept");
    pkb_ctx* ctx = r->ctx;
    CU(cudaSetDevice(ctx->device));
    for (int k = 0; k < 2 * K; ++k)
        if (cells[k] < 0 || cells[k] >= r->D) return fail(PKB_EINVAL, "pkb_result_sample: cell index %d outside the domain", cells[k]);
    DBuf<int> dc;
    DBuf<double> dv;
    TRY(dc.alloc(ctx, 2 * (size_t)K));
    TRY(dv.alloc(ctx, (size_t)K * r->ndays));
    CU(cudaMemcpyAsync(dc.p, cells, sizeof(int) * 2 * K, cudaMemcpyHostToDevice, ctx->stream));
    LAUNCH(ctx, k_sample, r->ndays, 256, 0, (const double*)r->dense.p, r->D, (const int*)dc.p, K, dv.p);
    CU(cudaMemcpyAsync(out, dv.p, sizeof(double) * K * r->ndays, cudaMemcpyDeviceToHost, ctx->stream));
    return sync_check(ctx, "pkb_result_sample");
}

extern "C" int pkb_result_project(pkb_result* r, const int* cells, int K, const pkb_projection* proj, double* out) {
    if (!r || !cells || !proj || !out || K < 1) return fail(PKB_EINVAL, "pkb_result_project: bad argument");
    if (!r->dense.p) return fail(PKB_ESTATE, "pkb_result_project: dense solutions were not kept");
    pkb_ctx* ctx = r->ctx;
    CU(cudaSetDevice(ctx->device));
    for (int k = 0; k < 2 * K; ++k)
        if (cells[k] < 0 || cells[k] >= r->D) return fail(PKB_EINVAL, "pkb_result_project: cell index %d outside the domain", cells[k]);
    DevProjection dp;
    TRY(upload_projection(ctx, proj, r->ndays, K, &dp));
    DBuf<int> dc;
    DBuf<double> dv, dout;
    TRY(dc.alloc(ctx, 2 * (size_t)K));
    TRY(dv.alloc(ctx, (size_t)K * r->ndays));
    TRY(dout.alloc(ctx, proj->nrows));
    CU(cudaMemcpyAsync(dc.p, cells, sizeof(int) * 2 * K, cudaMemcpyHostToDevice, ctx->stream));
    LAUNCH(ctx, k_sample, r->ndays, 256, 0, (const double*)r->dense.p, r->D, (const int*)dc.p, K, dv.p);
    launch_project(ctx, dp, dv.p, r->ndays, K, 1, dout.p);
    CU(cudaMemcpyAsync(out, dout.p, sizeof(double) * proj->nrows, cudaMemcpyDeviceToHost, ctx->stream));
    return sync_check(ctx, "pkb_result_project");
}

extern "C" int pkb_result_device_ptr(pkb_result* r, void** dptr) {
    if (!r || !dptr) return fail(PKB_EINVAL, "pkb_result_device_ptr: NULL argument");
    if (!r->dense.p) return fail(PKB_ESTATE, "pkb_result_device_ptr: dense solutions were not kept");
    *dptr = r->dense.p;
    return 0;
}

extern "C" int pkb_result_destroy(pkb_result* r) {
    if (!r) return 0;
    cudaSetDevice(r->ctx->device);
    cudaStreamSynchronize(r->ctx->stream);
    delete r;
    return 0;
}
