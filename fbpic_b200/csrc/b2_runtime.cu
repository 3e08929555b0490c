// b2_runtime.cu -- context, memory, streams, events, CUDA graphs (host side of the C ABI).
// Replaces the cupy array/pool plumbing of the reference GPU path (fbpic/utils/cuda.py:101-182).
#include "b2_common.cuh"
#include <cstdlib>
#include <cstring>
#include <map>
#include <mutex>

std::atomic<uint64_t> g_b2_launches{0};
thread_local char g_b2_err[512] = "no error";

int b2_fail(int code, const char *what, const char *file, int line) {
    snprintf(g_b2_err, sizeof(g_b2_err), "%s (code %d) at %s:%d", what, code, file, line);
    return code ? code : -1;
}

// ---- profiler -------------------------------------------------------------------------
#include <vector>
bool g_b2_prof_on = false;
struct ProfRec { int slot; cudaEvent_t a, b; };
static std::vector<ProfRec> g_prof_recs;
static std::vector<cudaEvent_t> g_prof_pool;
static const char *g_prof_names[B2P_NSLOTS] = {"cell_index", "sort", "permute", "gather", "push", "gather_push",
    "deposit_rho", "deposit_J", "fft", "dht", "spectral", "elementwise", "comm"};
static cudaEvent_t prof_event() {
    if (!g_prof_pool.empty()) { cudaEvent_t e = g_prof_pool.back(); g_prof_pool.pop_back(); return e; }
    cudaEvent_t e; cudaEventCreate(&e); return e;
}
B2Prof::B2Prof(int slot, cudaStream_t stream) : idx(-1), s(stream) {
    if (!g_b2_prof_on) return;
    ProfRec r; r.slot = slot; r.a = prof_event(); r.b = prof_event();
    cudaEventRecord(r.a, s);
    g_prof_recs.push_back(r);
    idx = (int)g_prof_recs.size() - 1;
}
B2Prof::~B2Prof() { if (idx >= 0) cudaEventRecord(g_prof_recs[idx].b, s); }

extern "C" {

int b2_profile_enable(int on) { g_b2_prof_on = (on != 0); return 0; }
int b2_profile_reset(void) {
    for (auto &r : g_prof_recs) { g_prof_pool.push_back(r.a); g_prof_pool.push_back(r.b); }
    g_prof_recs.clear();
    return 0;
}
int b2_profile_slots(void) { return B2P_NSLOTS; }
const char *b2_profile_name(int slot) { return (slot >= 0 && slot < B2P_NSLOTS) ? g_prof_names[slot] : ""; }
int b2_profile_read(int slot, double *total_ms, uint64_t *count) {
    double t = 0.; uint64_t n = 0;
    for (auto &r : g_prof_recs) {
        if (r.slot != slot) continue;
        B2_CUDA(cudaEventSynchronize(r.b));
        float ms = 0.f;
        B2_CUDA(cudaEventElapsedTime(&ms, r.a, r.b));
        t += ms; ++n;
    }
    *total_ms = t; *count = n;
    return 0;
}

const char *b2_error_string(void) { return g_b2_err; }
const char *b2_version(void) { return "fbpic_b200 0.1 (sm_100a)"; }
uint64_t b2_launch_count(void) { return g_b2_launches.load(); }

int b2_device_count(int *count) {
    B2_CUDA(cudaGetDeviceCount(count));
    return 0;
}

int b2_ctx_create(int device, b2_ctx **out) {
    B2_CUDA(cudaSetDevice(device));
    cudaDeviceProp prop;
    B2_CUDA(cudaGetDeviceProperties(&prop, device));
    if (prop.major < 10)
        return b2_fail(-2, "fbpic_b200 needs an sm_100a (B200) device", __FILE__, __LINE__);
    cudaStream_t stream;
    B2_CUDA(cudaStreamCreate(&stream));
    b2_ctx *ctx = new b2_ctx();
    ctx->device = device;
    ctx->stream = stream;
    for (int i = 0; i < 4; ++i) { ctx->scratch[i] = nullptr; ctx->scratch_bytes[i] = 0; }
    ctx->last_idx32 = nullptr; ctx->last_keys_sorted = nullptr; ctx->last_sort_n = -1; ctx->part_n = -1; ctx->last_sort_prefix = nullptr;
    ctx->nccl_comm = nullptr;
    ctx->nccl_rank = 0;
    ctx->nccl_size = 1;
    ctx->sm_count = prop.multiProcessorCount;
    *out = ctx;
    return 0;
}

int b2_ctx_destroy(b2_ctx *ctx) {
    if (!ctx) return 0;
    cudaSetDevice(ctx->device);
    cudaStreamSynchronize(ctx->stream);
    for (auto &kv : ctx->fft_plans) cufftDestroy(kv.second);
    for (int l = 1; l < ctx->fft_lanes; ++l) {
        cudaStreamDestroy(ctx->fft_lane[l]);
        cudaEventDestroy(ctx->fft_join[l]);
    }
    if (ctx->fft_lanes) cudaEventDestroy(ctx->fft_fork);
    for (int i = 0; i < 4; ++i) if (ctx->scratch[i]) cudaFree(ctx->scratch[i]);
    cudaStreamDestroy(ctx->stream);
    delete ctx;
    return 0;
}

void *b2_ctx_stream(b2_ctx *ctx) { return (void *)ctx->stream; }

// Device blocks are recycled through a size-keyed free list (the reference leans on cupy's memory pool for the
// same reason, fbpic/utils/cuda.py): Simulation.step() drops and re-creates every state array at its entry / exit,
// and cudaMalloc / cudaFree (a device-wide synchronisation each) of ~40 blocks of 16-134 MB took as long as the
// PCIe copies themselves.  Reuse is safe in stream order: everything this library launches is ordered on the
// context stream (the FFT lanes and NCCL join back into it).  B2_POOL_MB caps the parked bytes (default 65536);
// when cudaMalloc runs out of memory the parked blocks are released and the call is retried.
static std::mutex g_pool_mutex;
static std::multimap<size_t, void *> g_pool_free;        // parked blocks by size
static std::map<void *, size_t> g_pool_size;             // size of every block handed out or parked
static size_t g_pool_bytes = 0;
static size_t pool_round(size_t n) {
    if (n < 16) n = 16;
    if (n < (1u << 20)) return (n + 255) & ~(size_t)255;
    // above 1 MB: multiples of 1/16 of the power of two below n, so that arrays whose length drifts from call to call
    // (moving window, injection) find their blocks again
    size_t g = (size_t)1 << 20;
    while ((g << 5) <= n) g <<= 1;
    return (n + g - 1) / g * g;
}
static void pool_release_all() {
    for (auto &kv : g_pool_free) { g_pool_size.erase(kv.second); cudaFree(kv.second); }
    g_pool_free.clear();
    g_pool_bytes = 0;
}
int b2_malloc(void **p, size_t n) {
    const size_t cap = pool_round(n);
    std::lock_guard<std::mutex> lk(g_pool_mutex);
    auto it = g_pool_free.find(cap);
    if (it != g_pool_free.end()) {
        *p = it->second;
        g_pool_bytes -= cap;
        g_pool_free.erase(it);
        return 0;
    }
    cudaError_t err = cudaMalloc(p, cap);
    if (err == cudaErrorMemoryAllocation) {
        cudaGetLastError();
        pool_release_all();
        err = cudaMalloc(p, cap);
    }
    B2_CUDA(err);
    g_pool_size[*p] = cap;
    return 0;
}
int b2_free(void *p) {
    if (!p) return 0;
    b2_dht_forget(p);
    static const size_t limit = []() { const char *e = getenv("B2_POOL_MB"); return (size_t)(e ? atoll(e) : 65536) << 20; }();
    std::lock_guard<std::mutex> lk(g_pool_mutex);
    auto it = g_pool_size.find(p);
    if (it != g_pool_size.end() && g_pool_bytes + it->second <= limit) {
        g_pool_free.emplace(it->second, p);
        g_pool_bytes += it->second;
        return 0;
    }
    if (it != g_pool_size.end()) g_pool_size.erase(it);
    B2_CUDA(cudaFree(p));
    return 0;
}
int b2_host_alloc(void **p, size_t n) { B2_CUDA(cudaMallocHost(p, n ? n : 16)); return 0; }
int b2_host_free(void *p) { B2_CUDA(cudaFreeHost(p)); return 0; }
int b2_memcpy_h2d(void *d, const void *h, size_t n, void *s) {
    b2_dht_forget(d);
    B2_CUDA(cudaMemcpyAsync(d, h, n, cudaMemcpyHostToDevice, (cudaStream_t)s)); return 0; }
int b2_memcpy_d2h(void *h, const void *d, size_t n, void *s) {
    B2_CUDA(cudaMemcpyAsync(h, d, n, cudaMemcpyDeviceToHost, (cudaStream_t)s)); return 0; }
int b2_memcpy_d2d(void *dst, const void *src, size_t n, void *s) {
    b2_dht_forget(dst);
    B2_CUDA(cudaMemcpyAsync(dst, src, n, cudaMemcpyDeviceToDevice, (cudaStream_t)s));
    g_b2_launches.fetch_add(1); return 0; }
int b2_memset(void *p, int v, size_t n, void *s) {
    b2_dht_forget(p);
    B2_CUDA(cudaMemsetAsync(p, v, n, (cudaStream_t)s)); g_b2_launches.fetch_add(1); return 0; }
int b2_stream_sync(void *s) { B2_CUDA(cudaStreamSynchronize((cudaStream_t)s)); return 0; }
int b2_device_sync(void) { B2_CUDA(cudaDeviceSynchronize()); return 0; }
int b2_stream_create(void **s) { cudaStream_t st; B2_CUDA(cudaStreamCreateWithFlags(&st, cudaStreamNonBlocking)); *s = (void *)st; return 0; }
int b2_stream_destroy(void *s) { B2_CUDA(cudaStreamDestroy((cudaStream_t)s)); return 0; }
int b2_stream_wait_event(void *s, void *e) { B2_CUDA(cudaStreamWaitEvent((cudaStream_t)s, (cudaEvent_t)e, 0)); return 0; }
int b2_event_create(void **e) { cudaEvent_t ev; B2_CUDA(cudaEventCreate(&ev)); *e = (void *)ev; return 0; }
int b2_event_destroy(void *e) { B2_CUDA(cudaEventDestroy((cudaEvent_t)e)); return 0; }
int b2_event_record(void *e, void *s) { B2_CUDA(cudaEventRecord((cudaEvent_t)e, (cudaStream_t)s)); return 0; }
int b2_event_elapsed_ms(void *a, void *b, float *ms) {
    B2_CUDA(cudaEventSynchronize((cudaEvent_t)b));
    B2_CUDA(cudaEventElapsedTime(ms, (cudaEvent_t)a, (cudaEvent_t)b)); return 0; }

int b2_graph_begin(b2_ctx *ctx) {
    B2_CUDA(cudaStreamBeginCapture(ctx->stream, cudaStreamCaptureModeThreadLocal)); return 0; }
int b2_graph_end(b2_ctx *ctx, void **exec) {
    cudaGraph_t g;
    B2_CUDA(cudaStreamEndCapture(ctx->stream, &g));
    cudaGraphExec_t ge;
    B2_CUDA(cudaGraphInstantiate(&ge, g, 0));
    B2_CUDA(cudaGraphDestroy(g));
    *exec = (void *)ge;
    return 0;
}
int b2_graph_launch(b2_ctx *ctx, void *exec) {
    B2_CUDA(cudaGraphLaunch((cudaGraphExec_t)exec, ctx->stream)); return 0; }
int b2_graph_destroy(void *exec) { B2_CUDA(cudaGraphExecDestroy((cudaGraphExec_t)exec)); return 0; }

}  // extern "C"

int b2_scratch(b2_ctx *ctx, int slot, size_t nbytes, void **ptr) {
    if (ctx->scratch_bytes[slot] < nbytes) {
        if (ctx->scratch[slot]) {
            B2_CUDA(cudaStreamSynchronize(ctx->stream));
            B2_CUDA(cudaFree(ctx->scratch[slot]));
        }
        size_t cap = nbytes + nbytes / 8 + 256;
        B2_CUDA(cudaMalloc(&ctx->scratch[slot], cap));
        ctx->scratch_bytes[slot] = cap;
    }
    *ptr = ctx->scratch[slot];
    return 0;
}
