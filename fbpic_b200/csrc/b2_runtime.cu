// b2_runtime.cu -- context, memory, streams, events, CUDA graphs (host side of the C ABI).
// Replaces the cupy array/pool plumbing of the reference GPU path (fbpic/utils/cuda.py:101-182).
#include "b2_common.cuh"
#include <cstring>

std::atomic<uint64_t> g_b2_launches{0};
thread_local char g_b2_err[512] = "no error";

int b2_fail(int code, const char *what, const char *file, int line) {
    snprintf(g_b2_err, sizeof(g_b2_err), "%s (code %d) at %s:%d", what, code, file, line);
    return code ? code : -1;
}

extern "C" {

const char *b2_error_string(void) { return g_b2_err; }
const char *b2_version(void) { return "fbpic_b200 0.1 (sm_100a)"; }
uint64_t b2_launch_count(void) { return g_b2_launches.load(); }

int b2_device_count(int *count) {
    B2_CUDA(cudaGetDeviceCount(count));
    return 0;
}

int b2_ctx_create(int device, b2_ctx **out) {
    B2_CUDA(cudaSetDevice(device));
    b2_ctx *ctx = new b2_ctx();
    ctx->device = device;
    B2_CUDA(cudaStreamCreate(&ctx->stream));
    for (int i = 0; i < 4; ++i) { ctx->scratch[i] = nullptr; ctx->scratch_bytes[i] = 0; }
    ctx->nccl_comm = nullptr;
    ctx->nccl_rank = 0;
    ctx->nccl_size = 1;
    cudaDeviceProp prop;
    B2_CUDA(cudaGetDeviceProperties(&prop, device));
    ctx->sm_count = prop.multiProcessorCount;
    if (prop.major < 10)
        return b2_fail(-2, "fbpic_b200 needs an sm_100a (B200) device", __FILE__, __LINE__);
    *out = ctx;
    return 0;
}

int b2_ctx_destroy(b2_ctx *ctx) {
    if (!ctx) return 0;
    cudaSetDevice(ctx->device);
    cudaStreamSynchronize(ctx->stream);
    for (auto &kv : ctx->fft_plans) cufftDestroy(kv.second);
    for (int i = 0; i < 4; ++i) if (ctx->scratch[i]) cudaFree(ctx->scratch[i]);
    cudaStreamDestroy(ctx->stream);
    delete ctx;
    return 0;
}

void *b2_ctx_stream(b2_ctx *ctx) { return (void *)ctx->stream; }

int b2_malloc(void **p, size_t n) { B2_CUDA(cudaMalloc(p, n ? n : 16)); return 0; }
int b2_free(void *p) { B2_CUDA(cudaFree(p)); return 0; }
int b2_host_alloc(void **p, size_t n) { B2_CUDA(cudaMallocHost(p, n ? n : 16)); return 0; }
int b2_host_free(void *p) { B2_CUDA(cudaFreeHost(p)); return 0; }
int b2_memcpy_h2d(void *d, const void *h, size_t n, void *s) {
    B2_CUDA(cudaMemcpyAsync(d, h, n, cudaMemcpyHostToDevice, (cudaStream_t)s)); return 0; }
int b2_memcpy_d2h(void *h, const void *d, size_t n, void *s) {
    B2_CUDA(cudaMemcpyAsync(h, d, n, cudaMemcpyDeviceToHost, (cudaStream_t)s)); return 0; }
int b2_memcpy_d2d(void *dst, const void *src, size_t n, void *s) {
    B2_CUDA(cudaMemcpyAsync(dst, src, n, cudaMemcpyDeviceToDevice, (cudaStream_t)s));
    g_b2_launches.fetch_add(1); return 0; }
int b2_memset(void *p, int v, size_t n, void *s) {
    B2_CUDA(cudaMemsetAsync(p, v, n, (cudaStream_t)s)); g_b2_launches.fetch_add(1); return 0; }
int b2_stream_sync(void *s) { B2_CUDA(cudaStreamSynchronize((cudaStream_t)s)); return 0; }
int b2_device_sync(void) { B2_CUDA(cudaDeviceSynchronize()); return 0; }
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
