// capi.cu -- the C ABI of libmsegment_b200.so (include/msegment.h): context, workspace, host staging,
// stream-ordered pipelines.  No CPU implementation of any operator lives here: every entry point
// ends in the CUDA kernels of k_*.cu or fails.
#include <math.h>
#include <stdarg.h>
#include <stdio.h>
#include <stdlib.h>

#include "msg_internal.h"

static char g_create_err[512] = "";

int msg_fail(msg_ctx* ctx, int code, const char* fmt, ...)
{
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(ctx ? ctx->err : g_create_err, 512, fmt, ap);
    va_end(ap);
    return code;
}

int msg_reserve(msg_ctx* ctx, void** p, size_t* cap, size_t bytes)
{
    if (*cap >= bytes && *p) return MSG_OK;
    if (*p) {
        MSG_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
        MSG_CUDA(ctx, cudaFree(*p));
        *p = nullptr;
        *cap = 0;
    }
    ctx->ws_epoch++;
    size_t want = bytes + bytes / 8 + 256;  // head-room against ping-pong regrowth
    cudaError_t e = cudaMalloc(p, want);
    if (e != cudaSuccess) {
        cudaGetLastError();
        *p = nullptr;
        return msg_fail(ctx, MSG_ENOMEM, "cudaMalloc(%zu) failed: %s", want, cudaGetErrorString(e));
    }
    *cap = want;
    return MSG_OK;
}

int msg_func_smem(msg_ctx* ctx, const void* func, size_t smem)
{
    for (int i = 0; i < ctx->n_attr; i++)
        if (ctx->attr_cache[i].func == func) {
            if (ctx->attr_cache[i].smem >= smem) return MSG_OK;
            MSG_CUDA(ctx, cudaFuncSetAttribute(func, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
            ctx->attr_cache[i].smem = smem;
            return MSG_OK;
        }
    MSG_CUDA(ctx, cudaFuncSetAttribute(func, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    if (ctx->n_attr < 64) {
        ctx->attr_cache[ctx->n_attr].func = func;
        ctx->attr_cache[ctx->n_attr].smem = smem;
        ctx->n_attr++;
    }
    return MSG_OK;
}

// ---- options: one table maps names to msg_tuning fields; msg_create seeds them from MSG_<NAME> once
struct opt_entry { const char* name; size_t off; int dflt; };
#define OPT(n, f, d) {n, offsetof(msg_tuning, f), d}
static const opt_entry g_opts[] = {
    OPT("tile_w", tile_w, 0),           OPT("acc", acc, -1),           OPT("pitch_res", pitch_res, -1),
    OPT("tma", use_tma, 1),             OPT("no_order", no_order, 0),  OPT("merge_scalar", merge_scalar, 0), OPT("merge_strips", merge_strips, 1),
    OPT("merge_small_max", merge_small_max, -1), OPT("merge_grid", merge_grid, 0), OPT("merge_medium_only", merge_medium_only, 0), OPT("no_graph", no_graph, 0), OPT("graph_debug", graph_debug, 0),
    OPT("ccl_legacy", ccl_legacy, 0),   OPT("ccl_quad", ccl_quad, 1),   OPT("gray_compat", gray_compat, 0), OPT("dt_fixed", dt_fixed, 0), OPT("dt_legacy", dt_legacy, 0),
    OPT("labels_canonical", labels_canonical, 0),
    OPT("staging", staging, 1),
};
#undef OPT
static int* opt_field(msg_tuning* t, const opt_entry& e) { return (int*)((char*)t + e.off); }

static void tuning_from_env(msg_tuning* t)
{
    for (const opt_entry& e : g_opts) {
        *opt_field(t, e) = e.dflt;
        char env[64] = "MSG_";
        size_t k = 4;
        for (const char* c = e.name; *c && k < sizeof(env) - 1; c++) env[k++] = (char)((*c >= 'a' && *c <= 'z') ? *c - 32 : *c);
        env[k] = 0;
        if (const char* v = getenv(env)) *opt_field(t, e) = *v ? atoi(v) : 1;     // MSG_NO_GRAPH= (empty) counts as set
    }
}

static void prof_fold(msg_ctx* ctx, int level)
{
    if (!ctx->prof_pending[level]) return;
    float a = 0, b = 0;
    cudaEventSynchronize(ctx->prof_ev[level][2]);
    cudaEventElapsedTime(&a, ctx->prof_ev[level][0], ctx->prof_ev[level][1]);
    cudaEventElapsedTime(&b, ctx->prof_ev[level][1], ctx->prof_ev[level][2]);
    ctx->prof.tile_ms[level] += a;
    ctx->prof.overflow_ms[level] += b;
    ctx->prof.launches[level] += 1;
    ctx->prof_pending[level] = 0;
}

void msg_prof_begin(msg_ctx* ctx, int level)
{
    prof_fold(ctx, level);
    cudaEventRecord(ctx->prof_ev[level][0], ctx->stream);
}

void msg_prof_end(msg_ctx* ctx, int level, int slot)
{
    cudaEventRecord(ctx->prof_ev[level][1 + slot], ctx->stream);
    if (slot == 1) ctx->prof_pending[level] = 1;
}

#define CTX_ENTER(ctx)                                                                           \
    do {                                                                                         \
        if (!(ctx)) return MSG_EINVAL;                                                           \
        if ((ctx)->cuda_failed) return msg_fail((ctx), MSG_ECUDA, "context is in a failed CUDA state"); \
        MSG_CUDA((ctx), cudaSetDevice((ctx)->device));                                           \
    } while (0)

extern "C" {

int msg_version(void) { return MSG_VERSION; }

int msg_device_count(void)
{
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) { cudaGetLastError(); return 0; }
    return n;
}

const char* msg_last_error(const msg_ctx* ctx) { return ctx ? ctx->err : g_create_err; }

int msg_create(int device, msg_ctx** out)
{
    if (!out) return MSG_EINVAL;
    *out = nullptr;
    int n = 0;
    cudaError_t e = cudaGetDeviceCount(&n);
    if (e != cudaSuccess || n == 0) {
        cudaGetLastError();
        return msg_fail(nullptr, MSG_ECUDA, "no CUDA device available (%s); this library has no CPU fallback",
                        e != cudaSuccess ? cudaGetErrorString(e) : "device count 0");
    }
    if (device < 0 || device >= n) return msg_fail(nullptr, MSG_EINVAL, "device %d out of range [0,%d)", device, n);
    msg_ctx* ctx = (msg_ctx*)calloc(1, sizeof(msg_ctx));
    if (!ctx) return msg_fail(nullptr, MSG_ENOMEM, "out of host memory");
    ctx->device = device;
#define CR(call)                                                                                  \
    do {                                                                                          \
        cudaError_t e2 = (call);                                                                  \
        if (e2 != cudaSuccess) {                                                                  \
            msg_fail(nullptr, MSG_ECUDA, "%s failed: %s", #call, cudaGetErrorString(e2));         \
            msg_destroy(ctx);   /* frees whatever was created so far (streams, events, allocations) */ \
            return MSG_ECUDA;                                                                     \
        }                                                                                         \
    } while (0)
    CR(cudaSetDevice(device));
    cudaDeviceProp prop;
    CR(cudaGetDeviceProperties(&prop, device));
    if (prop.major < 10) {
        msg_fail(nullptr, MSG_ECUDA, "device %d is sm_%d%d; this library is built for sm_100a (B200) only", device,
                 prop.major, prop.minor);
        free(ctx);
        return MSG_ECUDA;
    }
    tuning_from_env(&ctx->tune);
    ctx->sm_count = prop.multiProcessorCount;
    ctx->max_smem_optin = (int)prop.sharedMemPerBlockOptin;
    CR(cudaStreamCreateWithFlags(&ctx->own_stream, cudaStreamNonBlocking));
    ctx->stream = ctx->own_stream;
    for (int i = 0; i < 8; i++) CR(cudaEventCreate(&ctx->ev[i]));
    CR(cudaMalloc((void**)&ctx->d_counters, 64 * sizeof(int32_t)));
    CR(cudaMemset(ctx->d_counters, 0, 64 * sizeof(int32_t)));
    CR(cudaMallocHost((void**)&ctx->h_counters, 64 * sizeof(int32_t)));
    memset(ctx->h_counters, 0, 64 * sizeof(int32_t));
    CR(cudaMalloc((void**)&ctx->d_work, MSG_MAX_LEVELS * 4 * sizeof(unsigned long long)));
    CR(cudaMemset(ctx->d_work, 0, MSG_MAX_LEVELS * 4 * sizeof(unsigned long long)));
    for (int l = 0; l < MSG_MAX_LEVELS; l++)
        for (int k = 0; k < 3; k++) CR(cudaEventCreate(&ctx->prof_ev[l][k]));
    for (int i = 0; i < MSG_MAX_INFLIGHT; i++) {
        // msg_wait blocks on this event: a sleeping wait (not a spin) keeps the host cores free for the submitting threads when
        // many contexts share few cores (8 GPUs x 6 contexts on a 16-core host)
        CR(cudaEventCreateWithFlags(&ctx->pend[i].done, cudaEventDisableTiming | cudaEventBlockingSync));
        CR(cudaEventCreateWithFlags(&ctx->pend[i].ev_in, cudaEventDisableTiming));
        CR(cudaEventCreateWithFlags(&ctx->pend[i].ev_core, cudaEventDisableTiming));
        ctx->pend[i].n_regions_host = ctx->h_counters + 32 + i;
        // every in-flight frame owns its counters (region count, mean-shift statistics) and their pinned mirror
        CR(cudaMalloc((void**)&ctx->pend[i].d_cnt, 16 * sizeof(int32_t)));
        CR(cudaMemset(ctx->pend[i].d_cnt, 0, 16 * sizeof(int32_t)));
        CR(cudaMallocHost((void**)&ctx->pend[i].h_cnt, 16 * sizeof(int32_t)));
        memset(ctx->pend[i].h_cnt, 0, 16 * sizeof(int32_t));
    }
    for (int i = 0; i < MSG_RING_CHUNKS; i++) CR(cudaEventCreateWithFlags(&ctx->ring_ev[i], cudaEventDisableTiming));
    CR(cudaStreamCreateWithFlags(&ctx->h2d_stream, cudaStreamNonBlocking));
    CR(cudaStreamCreateWithFlags(&ctx->d2h_stream, cudaStreamNonBlocking));
#undef CR
    *out = ctx;
    return MSG_OK;
}

void msg_destroy(msg_ctx* ctx)
{
    // also the cleanup path of a failed msg_create: every member may still be null
    if (!ctx) return;
    cudaSetDevice(ctx->device);
    if (ctx->stream) cudaStreamSynchronize(ctx->stream);
    if (ctx->h2d_stream) cudaStreamSynchronize(ctx->h2d_stream);
    if (ctx->d2h_stream) cudaStreamSynchronize(ctx->d2h_stream);
    for (int i = 0; i < ctx->n_reg; i++) cudaHostUnregister((void*)ctx->reg[i].base);
    for (int i = 0; i < MSG_MAX_INFLIGHT; i++) {
        msg_ctx::pending& q = ctx->pend[i];
        cudaFree(q.d_in); cudaFree(q.d_filt); cudaFree(q.d_ren); cudaFree(q.d_lab); cudaFree(q.d_lab16); cudaFree(q.d_cnt);
        if (q.h_cnt) cudaFreeHost(q.h_cnt);
        if (q.h_out) cudaFreeHost(q.h_out);
        if (q.ev_in) cudaEventDestroy(q.ev_in);
        if (q.ev_core) cudaEventDestroy(q.ev_core);
        if (q.done) cudaEventDestroy(q.done);
        if (q.g_exec) cudaGraphExecDestroy(q.g_exec);
    }
    if (ctx->h2d_stream) cudaStreamDestroy(ctx->h2d_stream);
    if (ctx->d2h_stream) cudaStreamDestroy(ctx->d2h_stream);
    cudaFree(ctx->d_in); cudaFree(ctx->d_out); cudaFree(ctx->d_out2); cudaFree(ctx->d_labels);
    cudaFree(ctx->d_planes); cudaFree(ctx->d_ovf); cudaFree(ctx->d_scratch); cudaFree(ctx->d_counters);
    cudaFree(ctx->d_colors); cudaFree(ctx->d_work); cudaFree(ctx->d_cells); cudaFree(ctx->d_aux); cudaFree(ctx->d_small);
    cudaFree(ctx->d_ws); cudaFree(ctx->d_ccl);
    for (int l = 0; l < MSG_MAX_LEVELS; l++)
        for (int k = 0; k < 3; k++)
            if (ctx->prof_ev[l][k]) cudaEventDestroy(ctx->prof_ev[l][k]);
    if (ctx->h_counters) cudaFreeHost(ctx->h_counters);
    if (ctx->h_ring) cudaFreeHost(ctx->h_ring);
    for (int i = 0; i < MSG_RING_CHUNKS; i++)
        if (ctx->ring_ev[i]) cudaEventDestroy(ctx->ring_ev[i]);
    for (int i = 0; i < 8; i++)
        if (ctx->ev[i]) cudaEventDestroy(ctx->ev[i]);
    if (ctx->own_stream) cudaStreamDestroy(ctx->own_stream);
    cudaGetLastError();
    free(ctx);
}

int msg_set_stream(msg_ctx* ctx, void* cuda_stream)
{
    CTX_ENTER(ctx);
    MSG_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    ctx->stream = cuda_stream ? (cudaStream_t)cuda_stream : ctx->own_stream;
    return MSG_OK;
}

int msg_synchronize(msg_ctx* ctx)
{
    CTX_ENTER(ctx);
    MSG_CUDA(ctx, cudaStreamSynchronize(ctx->h2d_stream));
    MSG_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    MSG_CUDA(ctx, cudaStreamSynchronize(ctx->d2h_stream));
    ctx->st.merge_rounds = (uint64_t)ctx->h_counters[10];    // device-side callers: k_merge mirrors the round count asynchronously
    return MSG_OK;
}

int msg_set_option(msg_ctx* ctx, const char* name, int value)
{
    if (!ctx || !name) return MSG_EINVAL;
    for (const opt_entry& e : g_opts)
        if (!strcmp(e.name, name)) {
            if (*opt_field(&ctx->tune, e) != value) ctx->ws_epoch++;     // captured graphs must not outlive an option change
            *opt_field(&ctx->tune, e) = value;
            return MSG_OK;
        }
    return msg_fail(ctx, MSG_EINVAL, "unknown option '%s'", name);
}

int msg_get_option(msg_ctx* ctx, const char* name, int* value)
{
    if (!ctx || !name || !value) return MSG_EINVAL;
    for (const opt_entry& e : g_opts)
        if (!strcmp(e.name, name)) { *value = *opt_field(&ctx->tune, e); return MSG_OK; }
    return msg_fail(ctx, MSG_EINVAL, "unknown option '%s'", name);
}

int msg_register_host(msg_ctx* ctx, void* ptr, size_t bytes)
{
    CTX_ENTER(ctx);
    if (!ptr || !bytes) return msg_fail(ctx, MSG_EINVAL, "register_host: null range");
    if (ctx->n_reg >= 16) return msg_fail(ctx, MSG_ESTATE, "register_host: 16 ranges already registered");
    cudaError_t e = cudaHostRegister(ptr, bytes, cudaHostRegisterPortable);
    if (e != cudaSuccess) {
        cudaGetLastError();
        return msg_fail(ctx, MSG_ENOMEM, "cudaHostRegister(%zu bytes) failed: %s", bytes, cudaGetErrorString(e));
    }
    ctx->reg[ctx->n_reg].base = (const uint8_t*)ptr;
    ctx->reg[ctx->n_reg].bytes = bytes;
    ctx->n_reg++;
    return MSG_OK;
}

int msg_unregister_host(msg_ctx* ctx, void* ptr)
{
    CTX_ENTER(ctx);
    for (int i = 0; i < ctx->n_reg; i++)
        if (ctx->reg[i].base == (const uint8_t*)ptr) {
            MSG_TRY(msg_synchronize(ctx));                       // no copy of ours may still be using the range
            cudaHostUnregister(ptr);
            cudaGetLastError();
            ctx->reg[i] = ctx->reg[--ctx->n_reg];
            return MSG_OK;
        }
    return msg_fail(ctx, MSG_ESTATE, "unregister_host: range was not registered with this context");
}

void* msg_alloc_pinned(size_t bytes)
{
    void* p = nullptr;
    if (cudaMallocHost(&p, bytes ? bytes : 1) != cudaSuccess) { cudaGetLastError(); return nullptr; }
    return p;
}

void msg_free_pinned(void* p)
{
    if (p) cudaFreeHost(p);
}

void msg_segment_params_default(msg_segment_params* p)
{
    if (!p) return;
    p->sp = 10.0; p->sr = 10.0; p->max_level = 1;
    p->term_type = MSG_TERM_COUNT | MSG_TERM_EPS; p->max_count = 5; p->eps = 1.0;
    p->lo_diff = 2; p->min_size = 0; p->color_dist = 0; p->render_depth = 0; p->connectivity = 4;
    p->labels_type = MSG_LABELS_32S;
}

int msg_get_timings(msg_ctx* ctx, msg_timings* out)
{
    if (!ctx || !out) return MSG_EINVAL;
    *out = ctx->tm;
    return MSG_OK;
}

}  // extern "C"

// ============================================================================ mean-shift pipeline

struct ms_config {
    double sp0;
    int isr2, isr22, max_level, max_count, ieps;
};

static int ms_validate(msg_ctx* ctx, int w, int h, double sp, double sr, int max_level, int term_type, int max_count,
                       double eps, ms_config* cfg)
{
    if (w <= 0 || h <= 0) return msg_fail(ctx, MSG_EINVAL, "image size %dx%d is empty", w, h);
    if (max_level < 0 || max_level > 8)
        return msg_fail(ctx, MSG_EINVAL, "The number of pyramid levels is too large or negative (maxLevel=%d)", max_level);
    if (isnan(sp) || isnan(sr) || isnan(eps)) return msg_fail(ctx, MSG_EINVAL, "NaN parameter");
    if (!(term_type & MSG_TERM_COUNT)) max_count = 5;
    if (max_count < 1) max_count = 1;
    if (max_count > 100) max_count = 100;
    if (!(term_type & MSG_TERM_EPS)) eps = 1.0;
    if (eps < 0) eps = 0;
    double sr2 = sr * sr;
    cfg->isr2 = sr2 >= 2147483647.0 ? 2147483647 : (int)lrint(sr2);   // cvRound(sr*sr)
    cfg->isr22 = cfg->isr2 > 16 ? cfg->isr2 : 16;
    cfg->sp0 = sp;
    cfg->max_level = max_level;
    cfg->max_count = max_count;
    cfg->ieps = eps >= 2147483647.0 ? 2147483647 : (int)floor(eps);
    return MSG_OK;
}

// Builds planes for rows [r0, r1) (global, level 0) of a w x hfull image whose BGR rows start at d_src
// (row r0 first), runs all levels, leaves the result in ctx->D[0].
static int ms_run(msg_ctx* ctx, const uint8_t* d_src, size_t sstep, int w, int hfull, int r0, int r1, const ms_config& cfg)
{
    const int L = cfg.max_level;
    int lw[MSG_MAX_LEVELS], lh[MSG_MAX_LEVELS], ly0[MSG_MAX_LEVELS], ly1[MSG_MAX_LEVELS];
    lw[0] = w; lh[0] = hfull; ly0[0] = r0; ly1[0] = r1;
    for (int l = 1; l <= L; l++) {
        lw[l] = (lw[l - 1] + 1) / 2; lh[l] = (lh[l - 1] + 1) / 2;
        ly0[l] = ly0[l - 1] / 2; ly1[l] = (ly1[l - 1] + 1) / 2;
    }
    size_t total = 0, offs[MSG_MAX_LEVELS];
    for (int l = 0; l <= L; l++) {
        offs[l] = total;
        total += 2 * (size_t)msg_align_up(lw[l], 32) * (size_t)(ly1[l] - ly0[l]);
    }
    MSG_TRY(msg_reserve(ctx, (void**)&ctx->d_planes, &ctx->d_planes_cap, total * sizeof(uint32_t)));
    for (int l = 0; l <= L; l++) {
        msg_plane s;
        s.w = lw[l]; s.rows = ly1[l] - ly0[l]; s.y0 = ly0[l]; s.hfull = lh[l]; s.pitch = msg_align_up(lw[l], 32);
        s.p = ctx->d_planes + offs[l];
        msg_plane d = s;
        d.p = s.p + (size_t)s.pitch * s.rows;
        ctx->S[l] = s; ctx->D[l] = d;
    }
    ctx->last_levels = L;
    MSG_CUDA(ctx, cudaMemsetAsync(ctx->d_counters, 0, 8 * sizeof(int32_t), ctx->stream));
    MSG_TRY(k_bgr_to_plane(ctx, d_src, sstep, ctx->S[0]));
    for (int l = 1; l <= L; l++) MSG_TRY(k_pyr_down(ctx, ctx->S[l - 1], ctx->S[l]));
    for (int l = L; l >= 0; l--) {
        msg_ms_params prm;
        float sp = (float)(cfg.sp0 / (double)(1 << l));
        if (!(sp >= 1.f)) sp = 1.f;
        prm.sp = sp;
        prm.radius = (int)ceilf(sp);
        prm.isr2 = cfg.isr2;
        prm.max_count = cfg.max_count;
        prm.ieps = cfg.ieps;
        prm.use_mask = l < L;
        if (l < L) {
            // per-cell activity counts (32x32 cells) for the heavy-first tile order of this level's mean-shift kernel
            int cells_x = (ctx->D[l].w + 31) / 32, cells_y = (ctx->D[l].rows + 31) / 32;
            size_t ncell = (size_t)cells_x * cells_y;
            MSG_TRY(msg_reserve(ctx, (void**)&ctx->d_cells, &ctx->d_cells_cap, (2 * ncell + 4096) * sizeof(int32_t)));
            MSG_CUDA(ctx, cudaMemsetAsync(ctx->d_cells, 0, ncell * sizeof(int32_t), ctx->stream));
            MSG_TRY(k_pyr_up_mask(ctx, ctx->D[l + 1], ctx->D[l], cfg.isr22, ctx->d_cells, cells_x));
        }
        MSG_TRY(k_meanshift_level(ctx, ctx->S[l], ctx->D[l], prm, l));
    }
    return MSG_OK;
}

static int fetch_ms_stats(msg_ctx* ctx)
{
    // counters [2..3] active items (u64), [4..5] overflow items (u64)
    MSG_CUDA(ctx, cudaMemcpyAsync(ctx->h_counters, ctx->d_counters, 8 * sizeof(int32_t), cudaMemcpyDeviceToHost, ctx->stream));
    return MSG_OK;
}

static void publish_ms_stats(msg_ctx* ctx)
{
    unsigned long long a, o;
    memcpy(&a, ctx->h_counters + 2, 8);
    memcpy(&o, ctx->h_counters + 4, 8);
    ctx->st.ms_active_items = a;
    ctx->st.ms_overflow_items = o;
    ctx->st.merge_rounds = (uint64_t)ctx->h_counters[10];
}

static float ev_ms(cudaEvent_t a, cudaEvent_t b)
{
    float ms = 0;
    if (cudaEventElapsedTime(&ms, a, b) != cudaSuccess) { cudaGetLastError(); return 0; }
    return ms;
}

// ---- host buffers.  cv::Mat storage is pageable (fastMalloc); handing it to cudaMemcpyAsync makes the driver stage it
// synchronously, which serialises the asynchronous path.  Pageable buffers therefore go through the context's pinned ring
// (uploads, synchronous downloads) or through the frame's pinned output staging (msg_submit_segment); pinned memory
// (msg_alloc_pinned, cudaHostAlloc, cudaHostRegister, msg_register_host) is used in place.
static bool host_is_pinned(msg_ctx* ctx, const void* p)
{
    if (!ctx->tune.staging) return true;          // option staging = 0: behave as round 1 (driver staging)
    const uint8_t* b = (const uint8_t*)p;
    for (int i = 0; i < ctx->n_reg; i++)
        if (b >= ctx->reg[i].base && b < ctx->reg[i].base + ctx->reg[i].bytes) return true;
    cudaPointerAttributes at;
    if (cudaPointerGetAttributes(&at, p) != cudaSuccess) { cudaGetLastError(); return false; }
    return at.type == cudaMemoryTypeHost || at.type == cudaMemoryTypeManaged || at.type == cudaMemoryTypeDevice;
}

static int ring_acquire(msg_ctx* ctx, int* idx)
{
    if (!ctx->h_ring) {
        cudaError_t e = cudaMallocHost((void**)&ctx->h_ring, MSG_RING_CHUNKS * MSG_RING_CHUNK_BYTES);
        if (e != cudaSuccess) { cudaGetLastError(); ctx->h_ring = nullptr; return msg_fail(ctx, MSG_ENOMEM, "pinned staging ring: %s", cudaGetErrorString(e)); }
    }
    const int i = ctx->ring_next;
    ctx->ring_next = (i + 1) % MSG_RING_CHUNKS;
    if (ctx->ring_busy[i]) {
        MSG_CUDA(ctx, cudaEventSynchronize(ctx->ring_ev[i]));
        ctx->ring_busy[i] = 0;
    }
    *idx = i;
    return MSG_OK;
}

// bytes [off, off + n) of the dense row-major image <-> the strided host image
static void host_gather(uint8_t* dst, const uint8_t* host, size_t hstep, size_t row_bytes, size_t off, size_t n)
{
    if (hstep == row_bytes) { memcpy(dst, host + off, n); return; }
    size_t r = off / row_bytes, c = off % row_bytes;
    while (n) {
        size_t k = row_bytes - c < n ? row_bytes - c : n;
        memcpy(dst, host + r * hstep + c, k);
        dst += k; n -= k; r++; c = 0;
    }
}

static void host_scatter(uint8_t* host, size_t hstep, size_t row_bytes, size_t off, const uint8_t* src, size_t n)
{
    if (hstep == row_bytes) { memcpy(host + off, src, n); return; }
    size_t r = off / row_bytes, c = off % row_bytes;
    while (n) {
        size_t k = row_bytes - c < n ? row_bytes - c : n;
        memcpy(host + r * hstep + c, src, k);
        src += k; n -= k; r++; c = 0;
    }
}

// upload into an already reserved dense device buffer
static int upload_on(msg_ctx* ctx, cudaStream_t st, const void* host, size_t hstep, size_t row_bytes, int rows, uint8_t* d)
{
    const size_t total = row_bytes * (size_t)rows;
    if (host_is_pinned(ctx, host)) {
        if (hstep == row_bytes)      // continuous Mat: one linear copy (the 2-D form is issued row by row by the copy engine)
            MSG_CUDA(ctx, cudaMemcpyAsync(d, host, total, cudaMemcpyHostToDevice, st));
        else
            MSG_CUDA(ctx, cudaMemcpy2DAsync(d, row_bytes, host, hstep, row_bytes, rows, cudaMemcpyHostToDevice, st));
    } else {
        // pageable: chunks are copied into the pinned ring by this thread while the DMA of the previous chunks runs
        for (size_t off = 0; off < total; off += MSG_RING_CHUNK_BYTES) {
            const size_t n = total - off < MSG_RING_CHUNK_BYTES ? total - off : MSG_RING_CHUNK_BYTES;
            int i;
            MSG_TRY(ring_acquire(ctx, &i));
            uint8_t* chunk = ctx->h_ring + (size_t)i * MSG_RING_CHUNK_BYTES;
            host_gather(chunk, (const uint8_t*)host, hstep, row_bytes, off, n);
            MSG_CUDA(ctx, cudaMemcpyAsync(d + off, chunk, n, cudaMemcpyHostToDevice, st));
            MSG_CUDA(ctx, cudaEventRecord(ctx->ring_ev[i], st));
            ctx->ring_busy[i] = 1;
        }
        ctx->st.staged_bytes += total;
    }
    ctx->st.h2d_bytes += total;
    return MSG_OK;
}

static int copy_in_on(msg_ctx* ctx, cudaStream_t st, const void* host, size_t hstep, size_t row_bytes, int rows, uint8_t** d,
                      size_t* dcap)
{
    MSG_TRY(msg_reserve(ctx, (void**)d, dcap, row_bytes * (size_t)rows));
    return upload_on(ctx, st, host, hstep, row_bytes, rows, *d);
}

// Download on `st`.  Pinned destination: asynchronous.  Pageable destination: chunks are DMA'd into the pinned ring and copied
// out by this thread as they land -- the call returns when the host buffer is complete (the synchronous entry points
// synchronise right after it anyway; msg_submit_segment uses the frame's own output staging instead, see there).
static int copy_out_on(msg_ctx* ctx, cudaStream_t st, void* host, size_t hstep, const void* d, size_t row_bytes, int rows)
{
    const size_t total = row_bytes * (size_t)rows;
    ctx->st.d2h_bytes += total;
    if (host_is_pinned(ctx, host)) {
        if (hstep == row_bytes)
            MSG_CUDA(ctx, cudaMemcpyAsync(host, d, total, cudaMemcpyDeviceToHost, st));
        else
            MSG_CUDA(ctx, cudaMemcpy2DAsync(host, hstep, d, row_bytes, row_bytes, rows, cudaMemcpyDeviceToHost, st));
        return MSG_OK;
    }
    struct { int slot; size_t off, n; } fifo[MSG_RING_CHUNKS];
    int head = 0, cnt = 0;
    auto retire = [&]() -> int {
        MSG_CUDA(ctx, cudaEventSynchronize(ctx->ring_ev[fifo[head].slot]));
        ctx->ring_busy[fifo[head].slot] = 0;
        host_scatter((uint8_t*)host, hstep, row_bytes, fifo[head].off, ctx->h_ring + (size_t)fifo[head].slot * MSG_RING_CHUNK_BYTES,
                     fifo[head].n);
        head = (head + 1) % MSG_RING_CHUNKS;
        cnt--;
        return MSG_OK;
    };
    for (size_t off = 0; off < total; off += MSG_RING_CHUNK_BYTES) {
        const size_t n = total - off < MSG_RING_CHUNK_BYTES ? total - off : MSG_RING_CHUNK_BYTES;
        if (cnt == MSG_RING_CHUNKS) MSG_TRY(retire());
        int i;
        MSG_TRY(ring_acquire(ctx, &i));
        MSG_CUDA(ctx, cudaMemcpyAsync(ctx->h_ring + (size_t)i * MSG_RING_CHUNK_BYTES, (const uint8_t*)d + off, n, cudaMemcpyDeviceToHost, st));
        MSG_CUDA(ctx, cudaEventRecord(ctx->ring_ev[i], st));
        ctx->ring_busy[i] = 1;
        const int tail = (head + cnt) % MSG_RING_CHUNKS;
        fifo[tail].slot = i; fifo[tail].off = off; fifo[tail].n = n;
        cnt++;
    }
    while (cnt) MSG_TRY(retire());
    ctx->st.staged_bytes += total;
    return MSG_OK;
}

static int copy_in(msg_ctx* ctx, const void* host, size_t hstep, size_t row_bytes, int rows, uint8_t** d, size_t* dcap)
{
    return copy_in_on(ctx, ctx->stream, host, hstep, row_bytes, rows, d, dcap);
}

static int copy_out(msg_ctx* ctx, void* host, size_t hstep, const void* d, size_t row_bytes, int rows)
{
    return copy_out_on(ctx, ctx->stream, host, hstep, d, row_bytes, rows);
}

static int check_img(msg_ctx* ctx, const void* p, size_t step, int w, int h, int elem, const char* what)
{
    if (w <= 0 || h <= 0) return msg_fail(ctx, MSG_EINVAL, "%s: empty image %dx%d", what, w, h);
    if (!p) return msg_fail(ctx, MSG_EINVAL, "%s: null pointer", what);
    if (step < (size_t)w * elem) return msg_fail(ctx, MSG_EINVAL, "%s: step %zu < row bytes %zu", what, step, (size_t)w * elem);
    if ((long long)w * h > 0x7fffffffLL - 1) return msg_fail(ctx, MSG_EINVAL, "%s: more than 2^31-2 pixels", what);
    if (h > 65535) return msg_fail(ctx, MSG_EINVAL, "%s: more than 65535 rows (image rows map to gridDim.y)", what);
    return MSG_OK;
}

extern "C" {

int msg_meanshift_filter_dev(msg_ctx* ctx, const uint8_t* d_src, size_t sstep, uint8_t* d_dst, size_t dstep, int w,
                             int h, double sp, double sr, int max_level, int term_type, int max_count, double eps)
{
    CTX_ENTER(ctx);
    MSG_TRY(check_img(ctx, d_src, sstep, w, h, 3, "meanshift src"));
    MSG_TRY(check_img(ctx, d_dst, dstep, w, h, 3, "meanshift dst"));
    ms_config cfg;
    MSG_TRY(ms_validate(ctx, w, h, sp, sr, max_level, term_type, max_count, eps, &cfg));
    MSG_TRY(ms_run(ctx, d_src, sstep, w, h, 0, h, cfg));
    MSG_TRY(k_plane_to_bgr(ctx, ctx->D[0], 0, h, d_dst, dstep));
    return MSG_OK;
}

int msg_meanshift_filter(msg_ctx* ctx, const uint8_t* src, size_t sstep, uint8_t* dst, size_t dstep, int w, int h,
                         double sp, double sr, int max_level, int term_type, int max_count, double eps)
{
    CTX_ENTER(ctx);
    MSG_TRY(check_img(ctx, src, sstep, w, h, 3, "meanshift src"));
    MSG_TRY(check_img(ctx, dst, dstep, w, h, 3, "meanshift dst"));
    ms_config cfg;
    MSG_TRY(ms_validate(ctx, w, h, sp, sr, max_level, term_type, max_count, eps, &cfg));
    size_t rb = (size_t)w * 3;
    MSG_CUDA(ctx, cudaEventRecord(ctx->ev[0], ctx->stream));
    MSG_TRY(copy_in(ctx, src, sstep, rb, h, &ctx->d_in, &ctx->d_in_cap));
    MSG_CUDA(ctx, cudaEventRecord(ctx->ev[1], ctx->stream));
    MSG_TRY(msg_reserve(ctx, (void**)&ctx->d_out, &ctx->d_out_cap, rb * h));
    MSG_TRY(ms_run(ctx, ctx->d_in, rb, w, h, 0, h, cfg));
    MSG_TRY(k_plane_to_bgr(ctx, ctx->D[0], 0, h, ctx->d_out, rb));
    MSG_CUDA(ctx, cudaEventRecord(ctx->ev[2], ctx->stream));
    MSG_TRY(copy_out(ctx, dst, dstep, ctx->d_out, rb, h));
    MSG_TRY(fetch_ms_stats(ctx));
    MSG_CUDA(ctx, cudaEventRecord(ctx->ev[3], ctx->stream));
    MSG_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    publish_ms_stats(ctx);
    memset(&ctx->tm, 0, sizeof(ctx->tm));
    ctx->tm.h2d_ms = ev_ms(ctx->ev[0], ctx->ev[1]);
    ctx->tm.filter_ms = ev_ms(ctx->ev[1], ctx->ev[2]);
    ctx->tm.d2h_ms = ev_ms(ctx->ev[2], ctx->ev[3]);
    ctx->tm.total_ms = ev_ms(ctx->ev[0], ctx->ev[3]);
    return MSG_OK;
}

// ============================================================================ labelling

}  // extern "C"

// labels 1..n of an 8UC3 device image into a dense int32 map.  Tile-local path: the kernels read the BGR bytes themselves
// (no packed plane); legacy path (option ccl_legacy): plane + row-run union-find + separate relabel.
static int label_from_bgr(msg_ctx* ctx, const uint8_t* d_bgr, size_t step, int w, int h, int lo_diff, int connectivity,
                          int32_t* d_dense_labels, int32_t* d_n)
{
    if (!ctx->tune.ccl_legacy)
        return k_label_canonical_src(ctx, d_bgr, step, 2, w, h, lo_diff, connectivity, d_dense_labels, d_n);
    msg_plane s;
    s.w = w; s.rows = h; s.y0 = 0; s.hfull = h; s.pitch = msg_align_up(w, 32);
    MSG_TRY(msg_reserve(ctx, (void**)&ctx->d_planes, &ctx->d_planes_cap, (size_t)s.pitch * h * sizeof(uint32_t)));
    s.p = ctx->d_planes;
    MSG_TRY(k_bgr_to_plane(ctx, d_bgr, step, s));
    MSG_TRY(k_ccl_color(ctx, s.p, s.pitch, w, h, lo_diff, connectivity, d_dense_labels, -1, w));
    return k_relabel_canonical(ctx, d_dense_labels, w, h, 1, d_n, 0);
}

extern "C" {

int msg_label_regions_dev(msg_ctx* ctx, const uint8_t* d_bgr, size_t step, int32_t* d_labels, size_t lstep, int w, int h,
                          int lo_diff, int connectivity, int32_t* d_n)
{
    CTX_ENTER(ctx);
    MSG_TRY(check_img(ctx, d_bgr, step, w, h, 3, "label src"));
    MSG_TRY(check_img(ctx, d_labels, lstep, w, h, 4, "labels"));
    if (lo_diff < 0) return msg_fail(ctx, MSG_EINVAL, "lo_diff must be >= 0");
    if (connectivity != 4 && connectivity != 8) return msg_fail(ctx, MSG_EINVAL, "connectivity must be 4 or 8");
    if (lstep % 4) return msg_fail(ctx, MSG_EINVAL, "labels step must be a multiple of 4");
    bool dense = lstep == (size_t)w * 4;
    int32_t* work = d_labels;
    if (!dense) {
        MSG_TRY(msg_reserve(ctx, (void**)&ctx->d_labels, &ctx->d_labels_cap, (size_t)w * h * 4));
        work = ctx->d_labels;
    }
    MSG_TRY(label_from_bgr(ctx, d_bgr, step, w, h, lo_diff, connectivity, work, d_n));
    if (!dense) MSG_TRY(k_copy_labels_2d(ctx, work, (size_t)w * 4, d_labels, lstep, w, h));
    return MSG_OK;
}

int msg_connected_components_dev(msg_ctx* ctx, const uint8_t* d_mask, size_t step, int32_t* d_labels, size_t lstep,
                                 int w, int h, int connectivity, int32_t* d_n)
{
    CTX_ENTER(ctx);
    MSG_TRY(check_img(ctx, d_mask, step, w, h, 1, "connectedComponents image"));
    MSG_TRY(check_img(ctx, d_labels, lstep, w, h, 4, "labels"));
    if (connectivity != 4 && connectivity != 8) return msg_fail(ctx, MSG_EINVAL, "connectivity must be 4 or 8");
    if (lstep % 4) return msg_fail(ctx, MSG_EINVAL, "labels step must be a multiple of 4");
    bool dense = lstep == (size_t)w * 4;
    int32_t* work = d_labels;
    if (!dense) {
        MSG_TRY(msg_reserve(ctx, (void**)&ctx->d_labels, &ctx->d_labels_cap, (size_t)w * h * 4));
        work = ctx->d_labels;
    }
    MSG_TRY(k_cc_canonical(ctx, d_mask, step, w, h, connectivity, work, d_n));   // count includes the background, as OpenCV's
    if (!dense) MSG_TRY(k_copy_labels_2d(ctx, work, (size_t)w * 4, d_labels, lstep, w, h));
    return MSG_OK;
}

int msg_merge_regions_dev(msg_ctx* ctx, const uint8_t* d_bgr, size_t step, int32_t* d_labels, size_t lstep, int w, int h,
                          int min_size, int color_dist, int32_t* d_n)
{
    CTX_ENTER(ctx);
    MSG_TRY(check_img(ctx, d_bgr, step, w, h, 3, "merge image"));
    MSG_TRY(check_img(ctx, d_labels, lstep, w, h, 4, "labels"));
    if (min_size < 0 || color_dist < 0) return msg_fail(ctx, MSG_EINVAL, "min_size and color_dist must be >= 0");
    if (lstep % 4) return msg_fail(ctx, MSG_EINVAL, "labels step must be a multiple of 4");
    msg_plane s;
    s.w = w; s.rows = h; s.y0 = 0; s.hfull = h; s.pitch = msg_align_up(w, 32);
    MSG_TRY(msg_reserve(ctx, (void**)&ctx->d_planes, &ctx->d_planes_cap, (size_t)s.pitch * h * sizeof(uint32_t)));
    s.p = ctx->d_planes;
    MSG_TRY(k_bgr_to_plane(ctx, d_bgr, step, s));
    bool dense = lstep == (size_t)w * 4;
    int32_t* work = d_labels;
    if (!dense) {
        MSG_TRY(msg_reserve(ctx, (void**)&ctx->d_labels, &ctx->d_labels_cap, (size_t)w * h * 4));
        work = ctx->d_labels;
        MSG_TRY(k_copy_labels_2d(ctx, d_labels, lstep, work, (size_t)w * 4, w, h));
    }
    // option "labels_canonical": the caller vouches that the labels are 1..n in raster order of first pixel (what
    // msg_label_regions / msg_connected_components write) with n = *d_n on entry: no validation, no renumbering pass
    if (ctx->tune.labels_canonical && !d_n)
        return msg_fail(ctx, MSG_EINVAL, "merge: option labels_canonical needs the label count in *d_n_regions on entry");
    MSG_TRY(k_merge(ctx, s.p, s.pitch, work, w, h, min_size, color_dist, ctx->tune.labels_canonical ? d_n : nullptr, d_n));
    if (!dense) MSG_TRY(k_copy_labels_2d(ctx, work, (size_t)w * 4, d_labels, lstep, w, h));
    return MSG_OK;
}

int msg_render_labels_dev(msg_ctx* ctx, const int32_t* d_labels, size_t lstep, uint8_t* d_dst, size_t dstep, int w, int h,
                          int depth, const uint8_t* d_colors)
{
    CTX_ENTER(ctx);
    MSG_TRY(check_img(ctx, d_labels, lstep, w, h, 4, "labels"));
    MSG_TRY(check_img(ctx, d_dst, dstep, w, h, 3, "render dst"));
    return k_render(ctx, d_labels, lstep, d_dst, dstep, w, h, depth, d_colors);
}

int msg_synth_bgr_dev(msg_ctx* ctx, uint8_t* d_dst, size_t step, int w, int h, uint64_t seed)
{
    CTX_ENTER(ctx);
    MSG_TRY(check_img(ctx, d_dst, step, w, h, 3, "synth dst"));
    return k_synth(ctx, d_dst, step, w, h, 0, h, seed);
}

int msg_synth_bgr_rows_dev(msg_ctx* ctx, uint8_t* d_dst, size_t step, int w, int full_h, int row0, int rows, uint64_t seed)
{
    CTX_ENTER(ctx);
    if (row0 < 0 || rows <= 0 || row0 + rows > full_h) return msg_fail(ctx, MSG_EINVAL, "synth rows: bad row range");
    MSG_TRY(check_img(ctx, d_dst, step, w, rows, 3, "synth dst"));
    return k_synth(ctx, d_dst, step, w, full_h, row0, rows, seed);
}

// ---- host-buffer label operators

static int finish_count(msg_ctx* ctx, int32_t* n_out)
{
    MSG_CUDA(ctx, cudaMemcpyAsync(ctx->h_counters + 16, ctx->d_counters + 16, sizeof(int32_t), cudaMemcpyDeviceToHost,
                                  ctx->stream));
    MSG_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    if (n_out) *n_out = ctx->h_counters[16];
    return MSG_OK;
}

int msg_label_regions(msg_ctx* ctx, const uint8_t* bgr, size_t step, int32_t* labels, size_t lstep, int w, int h,
                      int lo_diff, int up_diff, int connectivity, int32_t* n_regions)
{
    CTX_ENTER(ctx);
    MSG_TRY(check_img(ctx, bgr, step, w, h, 3, "label src"));
    MSG_TRY(check_img(ctx, labels, lstep, w, h, 4, "labels"));
    if (lo_diff != up_diff)
        return msg_fail(ctx, MSG_EINVAL, "lo_diff (%d) != up_diff (%d): asymmetric floodFill ranges are order dependent", lo_diff, up_diff);
    if (connectivity != 4 && connectivity != 8) return msg_fail(ctx, MSG_EINVAL, "connectivity must be 4 or 8 (got %d)", connectivity);
    if (lo_diff < 0) return msg_fail(ctx, MSG_EINVAL, "lo_diff must be >= 0");
    size_t rb = (size_t)w * 3;
    MSG_CUDA(ctx, cudaEventRecord(ctx->ev[0], ctx->stream));
    MSG_TRY(copy_in(ctx, bgr, step, rb, h, &ctx->d_in, &ctx->d_in_cap));
    MSG_CUDA(ctx, cudaEventRecord(ctx->ev[1], ctx->stream));
    MSG_TRY(msg_reserve(ctx, (void**)&ctx->d_labels, &ctx->d_labels_cap, (size_t)w * h * 4));
    MSG_TRY(label_from_bgr(ctx, ctx->d_in, rb, w, h, lo_diff, connectivity, ctx->d_labels, ctx->d_counters + 16));
    MSG_CUDA(ctx, cudaEventRecord(ctx->ev[2], ctx->stream));
    MSG_TRY(copy_out(ctx, labels, lstep, ctx->d_labels, (size_t)w * 4, h));
    MSG_CUDA(ctx, cudaEventRecord(ctx->ev[3], ctx->stream));
    MSG_TRY(finish_count(ctx, n_regions));
    memset(&ctx->tm, 0, sizeof(ctx->tm));
    ctx->tm.h2d_ms = ev_ms(ctx->ev[0], ctx->ev[1]);
    ctx->tm.label_ms = ev_ms(ctx->ev[1], ctx->ev[2]);
    ctx->tm.d2h_ms = ev_ms(ctx->ev[2], ctx->ev[3]);
    ctx->tm.total_ms = ev_ms(ctx->ev[0], ctx->ev[3]);
    return MSG_OK;
}

int msg_connected_components(msg_ctx* ctx, const uint8_t* mask, size_t step, int32_t* labels, size_t lstep, int w, int h,
                             int connectivity, int32_t* n_labels)
{
    CTX_ENTER(ctx);
    MSG_TRY(check_img(ctx, mask, step, w, h, 1, "connectedComponents image"));
    MSG_TRY(check_img(ctx, labels, lstep, w, h, 4, "labels"));
    if (connectivity != 4 && connectivity != 8) return msg_fail(ctx, MSG_EINVAL, "connectivity must be 4 or 8 (got %d)", connectivity);
    MSG_CUDA(ctx, cudaEventRecord(ctx->ev[0], ctx->stream));
    MSG_TRY(copy_in(ctx, mask, step, (size_t)w, h, &ctx->d_in, &ctx->d_in_cap));
    MSG_CUDA(ctx, cudaEventRecord(ctx->ev[1], ctx->stream));
    MSG_TRY(msg_reserve(ctx, (void**)&ctx->d_labels, &ctx->d_labels_cap, (size_t)w * h * 4));
    MSG_TRY(k_cc_canonical(ctx, ctx->d_in, (size_t)w, w, h, connectivity, ctx->d_labels, ctx->d_counters + 16));
    MSG_CUDA(ctx, cudaEventRecord(ctx->ev[2], ctx->stream));
    MSG_TRY(copy_out(ctx, labels, lstep, ctx->d_labels, (size_t)w * 4, h));
    MSG_CUDA(ctx, cudaEventRecord(ctx->ev[3], ctx->stream));
    MSG_TRY(finish_count(ctx, n_labels));
    memset(&ctx->tm, 0, sizeof(ctx->tm));
    ctx->tm.h2d_ms = ev_ms(ctx->ev[0], ctx->ev[1]);
    ctx->tm.label_ms = ev_ms(ctx->ev[1], ctx->ev[2]);
    ctx->tm.d2h_ms = ev_ms(ctx->ev[2], ctx->ev[3]);
    ctx->tm.total_ms = ev_ms(ctx->ev[0], ctx->ev[3]);
    return MSG_OK;
}

int msg_merge_regions(msg_ctx* ctx, const uint8_t* bgr, size_t step, int32_t* labels, size_t lstep, int w, int h,
                      int min_size, int color_dist, int32_t* n_regions)
{
    CTX_ENTER(ctx);
    MSG_TRY(check_img(ctx, bgr, step, w, h, 3, "merge image"));
    MSG_TRY(check_img(ctx, labels, lstep, w, h, 4, "labels"));
    if (min_size < 0 || color_dist < 0) return msg_fail(ctx, MSG_EINVAL, "min_size and color_dist must be >= 0");
    size_t rb = (size_t)w * 3;
    MSG_CUDA(ctx, cudaEventRecord(ctx->ev[0], ctx->stream));
    MSG_TRY(copy_in(ctx, bgr, step, rb, h, &ctx->d_in, &ctx->d_in_cap));
    MSG_TRY(msg_reserve(ctx, (void**)&ctx->d_labels, &ctx->d_labels_cap, (size_t)w * h * 4));
    MSG_TRY(upload_on(ctx, ctx->stream, labels, lstep, (size_t)w * 4, h, (uint8_t*)ctx->d_labels));
    MSG_CUDA(ctx, cudaEventRecord(ctx->ev[1], ctx->stream));
    msg_plane s;
    s.w = w; s.rows = h; s.y0 = 0; s.hfull = h; s.pitch = msg_align_up(w, 32);
    MSG_TRY(msg_reserve(ctx, (void**)&ctx->d_planes, &ctx->d_planes_cap, (size_t)s.pitch * h * sizeof(uint32_t)));
    s.p = ctx->d_planes;
    MSG_TRY(k_bgr_to_plane(ctx, ctx->d_in, rb, s));
    MSG_TRY(k_merge(ctx, s.p, s.pitch, ctx->d_labels, w, h, min_size, color_dist, nullptr, ctx->d_counters + 16));
    MSG_CUDA(ctx, cudaEventRecord(ctx->ev[2], ctx->stream));
    MSG_TRY(copy_out(ctx, labels, lstep, ctx->d_labels, (size_t)w * 4, h));
    MSG_CUDA(ctx, cudaEventRecord(ctx->ev[3], ctx->stream));
    MSG_TRY(finish_count(ctx, n_regions));
    ctx->st.merge_rounds = (uint64_t)ctx->h_counters[10];
    memset(&ctx->tm, 0, sizeof(ctx->tm));
    ctx->tm.h2d_ms = ev_ms(ctx->ev[0], ctx->ev[1]);
    ctx->tm.merge_ms = ev_ms(ctx->ev[1], ctx->ev[2]);
    ctx->tm.d2h_ms = ev_ms(ctx->ev[2], ctx->ev[3]);
    ctx->tm.total_ms = ev_ms(ctx->ev[0], ctx->ev[3]);
    return MSG_OK;
}

int msg_render_labels(msg_ctx* ctx, const int32_t* labels, size_t lstep, uint8_t* dst, size_t dstep, int w, int h,
                      int depth, const uint8_t* colors)
{
    CTX_ENTER(ctx);
    MSG_TRY(check_img(ctx, labels, lstep, w, h, 4, "labels"));
    MSG_TRY(check_img(ctx, dst, dstep, w, h, 3, "render dst"));
    if (depth < 0) depth = 0;
    size_t rb = (size_t)w * 3;
    MSG_CUDA(ctx, cudaEventRecord(ctx->ev[0], ctx->stream));
    MSG_TRY(msg_reserve(ctx, (void**)&ctx->d_labels, &ctx->d_labels_cap, (size_t)w * h * 4));
    MSG_TRY(upload_on(ctx, ctx->stream, labels, lstep, (size_t)w * 4, h, (uint8_t*)ctx->d_labels));
    const uint8_t* d_colors = nullptr;
    if (colors && depth > 0) {
        MSG_TRY(msg_reserve(ctx, (void**)&ctx->d_colors, &ctx->d_colors_cap, (size_t)depth * 3));
        MSG_CUDA(ctx, cudaMemcpyAsync(ctx->d_colors, colors, (size_t)depth * 3, cudaMemcpyHostToDevice, ctx->stream));
        d_colors = ctx->d_colors;
    }
    MSG_CUDA(ctx, cudaEventRecord(ctx->ev[1], ctx->stream));
    MSG_TRY(msg_reserve(ctx, (void**)&ctx->d_out2, &ctx->d_out2_cap, rb * h));
    MSG_TRY(k_render(ctx, ctx->d_labels, (size_t)w * 4, ctx->d_out2, rb, w, h, depth, d_colors));
    MSG_CUDA(ctx, cudaEventRecord(ctx->ev[2], ctx->stream));
    MSG_TRY(copy_out(ctx, dst, dstep, ctx->d_out2, rb, h));
    MSG_CUDA(ctx, cudaEventRecord(ctx->ev[3], ctx->stream));
    MSG_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    memset(&ctx->tm, 0, sizeof(ctx->tm));
    ctx->tm.h2d_ms = ev_ms(ctx->ev[0], ctx->ev[1]);
    ctx->tm.render_ms = ev_ms(ctx->ev[1], ctx->ev[2]);
    ctx->tm.d2h_ms = ev_ms(ctx->ev[2], ctx->ev[3]);
    ctx->tm.total_ms = ev_ms(ctx->ev[0], ctx->ev[3]);
    return MSG_OK;
}

// ============================================================================ fused pipeline

// Device core of the fused pipeline.  All pointers are device pointers; any output may be NULL.  d_labels is int32_t* or,
// with labels_type = MSG_LABELS_16U, uint16_t* (lstep in bytes).  d_stats (optional): 4 int32 receiving the mean-shift
// statistics of THIS frame (active items u64, overflow items u64) in stream order, so that pipelined frames do not read each
// other's counters.
// Events: ev[1] start, ev[2] after filter, ev[3] after label, ev[4] after merge, ev[5] after render.
static int segment_core_dev(msg_ctx* ctx, const uint8_t* d_src, size_t sstep, int w, int h, const msg_segment_params* p,
                            const ms_config& cfg, uint8_t* d_filtered, size_t fstep, void* d_labels, size_t lstep,
                            uint8_t* d_rendered, size_t rstep, int32_t* d_n, int32_t* d_stats)
{
    const bool l16 = p->labels_type == MSG_LABELS_16U;
    const bool do_label = p->lo_diff >= 0 && (d_labels || d_rendered || d_n);
    const bool do_merge = do_label && (p->min_size > 0 || p->color_dist > 0);
    const bool do_render = do_label && p->render_depth >= 0 && d_rendered;
    cudaStream_t st = ctx->stream;
    if (d_labels && lstep % (l16 ? 2 : 4)) return msg_fail(ctx, MSG_EINVAL, "labels step must be a multiple of the label size");
    if (!ctx->no_events) MSG_CUDA(ctx, cudaEventRecord(ctx->ev[1], st));
    MSG_TRY(ms_run(ctx, d_src, sstep, w, h, 0, h, cfg));
    if (d_stats) MSG_CUDA(ctx, cudaMemcpyAsync(d_stats, ctx->d_counters + 2, 4 * sizeof(int32_t), cudaMemcpyDeviceToDevice, st));
    if (d_filtered) MSG_TRY(k_plane_to_bgr(ctx, ctx->D[0], 0, h, d_filtered, fstep));
    if (!ctx->no_events) MSG_CUDA(ctx, cudaEventRecord(ctx->ev[2], st));
    int32_t* n_dev = d_n ? d_n : ctx->d_counters + 17;
    int32_t* work = nullptr;
    const bool dense = d_labels && !l16 && lstep == (size_t)w * 4;
    if (do_label) {
        work = (int32_t*)d_labels;
        if (!dense) {
            MSG_TRY(msg_reserve(ctx, (void**)&ctx->d_labels, &ctx->d_labels_cap, (size_t)w * h * 4));
            work = ctx->d_labels;
        }
        MSG_TRY(k_label_canonical(ctx, ctx->D[0].p, ctx->D[0].pitch, w, h, p->lo_diff, p->connectivity == 8 ? 8 : 4, work, n_dev));
    } else if (d_n) {
        MSG_CUDA(ctx, cudaMemsetAsync(d_n, 0, sizeof(int32_t), st));
    }
    if (!ctx->no_events) MSG_CUDA(ctx, cudaEventRecord(ctx->ev[3], st));
    if (do_merge) MSG_TRY(k_merge(ctx, ctx->D[0].p, ctx->D[0].pitch, work, w, h, p->min_size, p->color_dist, n_dev, n_dev));
    if (!ctx->no_events) MSG_CUDA(ctx, cudaEventRecord(ctx->ev[4], st));
    if (do_label && d_labels && l16) MSG_TRY(k_labels_to_u16(ctx, work, w, h, (uint16_t*)d_labels, lstep));
    else if (do_label && d_labels && !dense) MSG_TRY(k_copy_labels_2d(ctx, work, (size_t)w * 4, (int32_t*)d_labels, lstep, w, h));
    if (do_render) {
        int depth = p->render_depth > 0 ? p->render_depth : 0x7fffffff;   // 0: every region renders
        MSG_TRY(k_render(ctx, work, (size_t)w * 4, d_rendered, rstep, w, h, depth, nullptr));
    }
    if (!ctx->no_events) MSG_CUDA(ctx, cudaEventRecord(ctx->ev[5], st));
    return MSG_OK;
}

static int segment_validate(msg_ctx* ctx, const uint8_t* src, size_t sstep, int w, int h, const msg_segment_params* p,
                            uint8_t* filtered, size_t fstep, void* labels, size_t lstep, uint8_t* rendered, size_t rstep,
                            ms_config* cfg)
{
    MSG_TRY(check_img(ctx, src, sstep, w, h, 3, "segment src"));
    if (!p) return msg_fail(ctx, MSG_EINVAL, "segment: params is NULL");
    if (p->labels_type != MSG_LABELS_32S && p->labels_type != MSG_LABELS_16U)
        return msg_fail(ctx, MSG_EINVAL, "segment: labels_type must be MSG_LABELS_32S or MSG_LABELS_16U");
    if (filtered) MSG_TRY(check_img(ctx, filtered, fstep, w, h, 3, "segment filtered"));
    if (labels) MSG_TRY(check_img(ctx, labels, lstep, w, h, p->labels_type == MSG_LABELS_16U ? 2 : 4, "segment labels"));
    if (rendered) MSG_TRY(check_img(ctx, rendered, rstep, w, h, 3, "segment rendered"));
    MSG_TRY(ms_validate(ctx, w, h, p->sp, p->sr, p->max_level, p->term_type, p->max_count, p->eps, cfg));
    if (p->min_size < 0 || p->color_dist < 0) return msg_fail(ctx, MSG_EINVAL, "min_size and color_dist must be >= 0");
    if (p->connectivity != 0 && p->connectivity != 4 && p->connectivity != 8) return msg_fail(ctx, MSG_EINVAL, "connectivity must be 4 or 8");
    return MSG_OK;
}

static int segment_enqueue(msg_ctx* ctx, const uint8_t* src, size_t sstep, int w, int h, const msg_segment_params* p,
                           uint8_t* filtered, size_t fstep, void* labels, size_t lstep, uint8_t* rendered,
                           size_t rstep, int32_t* h_n_slot)
{
    ms_config cfg;
    MSG_TRY(segment_validate(ctx, src, sstep, w, h, p, filtered, fstep, labels, lstep, rendered, rstep, &cfg));
    const bool do_label = p->lo_diff >= 0;
    const bool do_render = do_label && p->render_depth >= 0 && rendered;
    const bool l16 = p->labels_type == MSG_LABELS_16U;
    const size_t rb = (size_t)w * 3, lb = (size_t)w * (l16 ? 2 : 4);
    cudaStream_t st = ctx->stream;
    // every allocation before the first enqueue: a failed reservation must not leave a DMA reading the caller's buffer
    MSG_TRY(msg_reserve(ctx, (void**)&ctx->d_in, &ctx->d_in_cap, rb * h));
    if (filtered) MSG_TRY(msg_reserve(ctx, (void**)&ctx->d_out, &ctx->d_out_cap, rb * h));
    if (do_label) MSG_TRY(msg_reserve(ctx, (void**)&ctx->d_labels, &ctx->d_labels_cap, (size_t)w * h * 4));
    if (do_label && labels && l16) MSG_TRY(msg_reserve(ctx, (void**)&ctx->d_aux, &ctx->d_aux_cap, lb * h));
    if (do_render) MSG_TRY(msg_reserve(ctx, (void**)&ctx->d_out2, &ctx->d_out2_cap, rb * h));
    MSG_CUDA(ctx, cudaEventRecord(ctx->ev[0], st));
    MSG_TRY(upload_on(ctx, st, src, sstep, rb, h, ctx->d_in));
    int32_t* d_n = ctx->d_counters + 16;
    void* d_lab_out = !do_label ? nullptr : (l16 ? (labels ? (void*)ctx->d_aux : nullptr) : (void*)ctx->d_labels);
    int rc = segment_core_dev(ctx, ctx->d_in, rb, w, h, p, cfg, filtered ? ctx->d_out : nullptr, rb, d_lab_out, lb,
                              do_render ? ctx->d_out2 : nullptr, rb, d_n, nullptr);
    if (rc != MSG_OK) { cudaStreamSynchronize(st); return rc; }      // the upload may still be reading the caller's buffer
    // all downloads last, so the stage timings are clean
    if (filtered) MSG_TRY(copy_out(ctx, filtered, fstep, ctx->d_out, rb, h));
    if (do_label && labels) MSG_TRY(copy_out(ctx, labels, lstep, d_lab_out, lb, h));
    if (do_render) MSG_TRY(copy_out(ctx, rendered, rstep, ctx->d_out2, rb, h));
    MSG_CUDA(ctx, cudaMemcpyAsync(h_n_slot, d_n, sizeof(int32_t), cudaMemcpyDeviceToHost, st));
    MSG_TRY(fetch_ms_stats(ctx));
    MSG_CUDA(ctx, cudaEventRecord(ctx->ev[6], st));
    return MSG_OK;
}

int msg_segment_dev(msg_ctx* ctx, const uint8_t* d_src, size_t sstep, int w, int h, const msg_segment_params* p,
                    uint8_t* d_filtered, size_t fstep, void* d_labels, size_t lstep, uint8_t* d_rendered, size_t rstep,
                    int32_t* d_n)
{
    CTX_ENTER(ctx);
    ms_config cfg;
    MSG_TRY(segment_validate(ctx, d_src, sstep, w, h, p, d_filtered, fstep, d_labels, lstep, d_rendered, rstep, &cfg));
    return segment_core_dev(ctx, d_src, sstep, w, h, p, cfg, d_filtered, fstep, d_labels, lstep, d_rendered, rstep, d_n, nullptr);
}

int msg_segment(msg_ctx* ctx, const uint8_t* src, size_t sstep, int w, int h, const msg_segment_params* p,
                uint8_t* filtered, size_t fstep, void* labels, size_t lstep, uint8_t* rendered, size_t rstep,
                int32_t* n_regions)
{
    CTX_ENTER(ctx);
    MSG_TRY(segment_enqueue(ctx, src, sstep, w, h, p, filtered, fstep, labels, lstep, rendered, rstep, ctx->h_counters + 16));
    MSG_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    publish_ms_stats(ctx);
    if (n_regions) *n_regions = ctx->h_counters[16];
    memset(&ctx->tm, 0, sizeof(ctx->tm));
    ctx->tm.h2d_ms = ev_ms(ctx->ev[0], ctx->ev[1]);
    ctx->tm.filter_ms = ev_ms(ctx->ev[1], ctx->ev[2]);
    ctx->tm.label_ms = ev_ms(ctx->ev[2], ctx->ev[3]);
    ctx->tm.merge_ms = ev_ms(ctx->ev[3], ctx->ev[4]);
    ctx->tm.render_ms = ev_ms(ctx->ev[4], ctx->ev[5]);
    ctx->tm.d2h_ms = ev_ms(ctx->ev[5], ctx->ev[6]);
    ctx->tm.total_ms = ev_ms(ctx->ev[0], ctx->ev[6]);
    if (labels && p->labels_type == MSG_LABELS_16U && ctx->h_counters[16] > 65535)
        return msg_fail(ctx, MSG_ERANGE, "segment: %d regions do not fit 16-bit labels", ctx->h_counters[16]);
    return MSG_OK;
}

int msg_submit_segment(msg_ctx* ctx, const uint8_t* src, size_t sstep, int w, int h, const msg_segment_params* p,
                       uint8_t* filtered, size_t fstep, void* labels, size_t lstep, uint8_t* rendered, size_t rstep,
                       int* ticket)
{
    CTX_ENTER(ctx);
    if (!ticket) return msg_fail(ctx, MSG_EINVAL, "submit: ticket is NULL");
    int slot = -1;
    for (int i = 0; i < MSG_MAX_INFLIGHT; i++)
        if (!ctx->pend[i].used) { slot = i; break; }
    if (slot < 0) return msg_fail(ctx, MSG_ESTATE, "submit: %d submissions already in flight; call msg_wait", MSG_MAX_INFLIGHT);
    // three-stage pipeline over consecutive submissions: upload on h2d_stream, kernels on the context stream, downloads
    // on d2h_stream, chained by events; the frame owns its device buffers until msg_wait returns
    ms_config cfg;
    MSG_TRY(segment_validate(ctx, src, sstep, w, h, p, filtered, fstep, labels, lstep, rendered, rstep, &cfg));
    msg_ctx::pending& q = ctx->pend[slot];
    const bool do_label = p->lo_diff >= 0;
    const bool do_render = do_label && p->render_depth >= 0 && rendered;
    const bool l16 = p->labels_type == MSG_LABELS_16U;
    const size_t rb = (size_t)w * 3, lb = (size_t)w * (l16 ? 2 : 4);
    // ---- every reservation first (device buffers of the frame, pinned staging for pageable destinations): a failure here
    //      returns before anything touches the caller's buffers
    MSG_TRY(msg_reserve(ctx, (void**)&q.d_in, &q.d_in_cap, rb * h));
    if (filtered) MSG_TRY(msg_reserve(ctx, (void**)&q.d_filt, &q.d_filt_cap, rb * h));
    if (do_label && !l16) MSG_TRY(msg_reserve(ctx, (void**)&q.d_lab, &q.d_lab_cap, (size_t)w * h * 4));
    if (do_label && labels && l16) MSG_TRY(msg_reserve(ctx, (void**)&q.d_lab16, &q.d_lab16_cap, lb * h));
    if (do_render) MSG_TRY(msg_reserve(ctx, (void**)&q.d_ren, &q.d_ren_cap, rb * h));
    struct out_desc { void* host; size_t hstep; const void* dev; size_t row_bytes; bool staged; size_t off; } outs[3];
    int n_out = 0;
    size_t stage_bytes = 0;
    auto add_out = [&](void* host, size_t hstep, const void* dev, size_t row_bytes) {
        out_desc& o = outs[n_out++];
        o.host = host; o.hstep = hstep; o.dev = dev; o.row_bytes = row_bytes;
        o.staged = !host_is_pinned(ctx, host);
        o.off = stage_bytes;
        if (o.staged) stage_bytes += (row_bytes * (size_t)h + 255) & ~(size_t)255;
    };
    if (filtered) add_out(filtered, fstep, q.d_filt, rb);
    if (do_label && labels) add_out(labels, lstep, l16 ? (const void*)q.d_lab16 : (const void*)q.d_lab, lb);
    if (do_render) add_out(rendered, rstep, q.d_ren, rb);
    if (stage_bytes > q.h_out_cap) {
        if (q.h_out) { cudaFreeHost(q.h_out); q.h_out = nullptr; q.h_out_cap = 0; }
        cudaError_t e = cudaMallocHost((void**)&q.h_out, stage_bytes);
        if (e != cudaSuccess) { cudaGetLastError(); q.h_out = nullptr; return msg_fail(ctx, MSG_ENOMEM, "pinned output staging (%zu bytes): %s", stage_bytes, cudaGetErrorString(e)); }
        q.h_out_cap = stage_bytes;
    }
    // ---- upload
    MSG_TRY(upload_on(ctx, ctx->h2d_stream, src, sstep, rb, h, q.d_in));
    MSG_CUDA(ctx, cudaEventRecord(q.ev_in, ctx->h2d_stream));
    int32_t* d_n = q.d_cnt;
    MSG_CUDA(ctx, cudaStreamWaitEvent(ctx->stream, q.ev_in, 0));
    // The kernel sequence of a frame is fixed by (geometry, parameters, buffers): replay it as a CUDA graph.  The first
    // submission of a configuration runs eagerly (it sizes the workspace), the second is captured, later ones replay.
    unsigned char key[sizeof(q.g_key)];
    memset(key, 0, sizeof(key));
    void* d_lab_out = !do_label ? nullptr : (l16 ? (labels ? (void*)q.d_lab16 : nullptr) : (void*)q.d_lab);
    {
        size_t o = 0;
        auto put = [&](const void* v, size_t nbytes) { memcpy(key + o, v, nbytes); o += nbytes; };
        const void* ptrs[7] = {q.d_in, filtered ? q.d_filt : nullptr, (do_label && !l16) ? q.d_lab : nullptr, d_lab_out,
                               do_render ? q.d_ren : nullptr, (void*)ctx->stream, d_n};
        const double dv[3] = {p->sp, p->sr, p->eps};
        const int iv[11] = {w, h, p->max_level, p->term_type, p->max_count, p->lo_diff, p->min_size, p->color_dist,
                            p->render_depth, p->connectivity, p->labels_type};
        put(dv, sizeof(dv)); put(iv, sizeof(iv)); put(ptrs, sizeof(ptrs));     // field by field: no struct padding in the key
        static_assert(sizeof(dv) + sizeof(iv) + sizeof(ptrs) <= sizeof(q.g_key), "graph key too small");
    }
    const bool graphs_off = ctx->tune.no_graph != 0;
    const bool same = q.g_state > 0 && memcmp(key, q.g_key, sizeof(key)) == 0 && q.g_epoch == ctx->ws_epoch;
    auto run_core = [&]() {
        // with 16-bit labels the int32 working map stays in the context workspace; only the converted map belongs to the frame
        return segment_core_dev(ctx, q.d_in, rb, w, h, p, cfg, filtered ? q.d_filt : nullptr, rb,
                                l16 ? d_lab_out : (void*)(do_label ? q.d_lab : nullptr), lb, do_render ? q.d_ren : nullptr, rb, d_n,
                                q.d_cnt + 2);
    };
    ctx->no_events = 1;
    int rc = MSG_OK;
    if (graphs_off || ctx->profiling || q.g_state < 0 || !same) {
        if (q.g_state == 2) { cudaGraphExecDestroy(q.g_exec); q.g_exec = nullptr; }
        rc = run_core();
        if (q.g_state >= 0) { q.g_state = (rc == MSG_OK && !graphs_off && !ctx->profiling) ? 1 : 0; }
        memcpy(q.g_key, key, sizeof(key));
        q.g_epoch = ctx->ws_epoch;                   // after the run: it may have grown the workspace
    } else {
        if (q.g_state == 1) {
            cudaGraph_t graph = nullptr;
            const uint64_t l0 = ctx->st.kernel_launches;
            cudaError_t e = cudaStreamBeginCapture(ctx->stream, cudaStreamCaptureModeThreadLocal);
            int crc = e == cudaSuccess ? run_core() : MSG_ECUDA;
            cudaError_t e2 = e == cudaSuccess ? cudaStreamEndCapture(ctx->stream, &graph) : e;
            if (crc == MSG_OK && e2 == cudaSuccess && graph && cudaGraphInstantiate(&q.g_exec, graph, 0) == cudaSuccess) {
                q.g_state = 2;
                q.g_launches = ctx->st.kernel_launches - l0;
                ctx->st.kernel_launches = l0;        // counted again below, when the graph is launched
            } else {                                 // capture not possible here: stay on the eager path for this slot
                if (ctx->tune.graph_debug)
                    fprintf(stderr, "[msegment] graph capture failed: begin=%s core=%d end=%s\n", cudaGetErrorString(e), crc,
                            cudaGetErrorString(e2));
                cudaGetLastError();
                ctx->cuda_failed = 0;
                ctx->st.kernel_launches = l0;
                q.g_state = -1;
            }
            if (graph) cudaGraphDestroy(graph);
        }
        if (q.g_state == 2) {
            cudaError_t e = cudaGraphLaunch(q.g_exec, ctx->stream);
            if (e != cudaSuccess) {
                ctx->no_events = 0; ctx->cuda_failed = 1;
                cudaStreamSynchronize(ctx->h2d_stream);
                return msg_fail(ctx, MSG_ECUDA, "cudaGraphLaunch failed: %s", cudaGetErrorString(e));
            }
            ctx->st.kernel_launches += q.g_launches;
        } else {
            rc = run_core();
        }
    }
    ctx->no_events = 0;
    if (rc != MSG_OK) {
        cudaStreamSynchronize(ctx->h2d_stream);      // the caller is told the call failed: no DMA may still read its buffer
        cudaStreamSynchronize(ctx->stream);
        return rc;
    }
    MSG_CUDA(ctx, cudaEventRecord(q.ev_core, ctx->stream));
    cudaStream_t ds = ctx->d2h_stream;
    MSG_CUDA(ctx, cudaStreamWaitEvent(ds, q.ev_core, 0));
    q.n_defer = 0;
    for (int k = 0; k < n_out; k++) {
        const out_desc& o = outs[k];
        const size_t total = o.row_bytes * (size_t)h;
        if (!o.staged) {
            if (o.hstep == o.row_bytes) MSG_CUDA(ctx, cudaMemcpyAsync(o.host, o.dev, total, cudaMemcpyDeviceToHost, ds));
            else MSG_CUDA(ctx, cudaMemcpy2DAsync(o.host, o.hstep, o.dev, o.row_bytes, o.row_bytes, h, cudaMemcpyDeviceToHost, ds));
        } else {   // pageable destination: DMA into the frame's pinned staging now, memcpy to the caller inside msg_wait
            MSG_CUDA(ctx, cudaMemcpyAsync(q.h_out + o.off, o.dev, total, cudaMemcpyDeviceToHost, ds));
            q.defer[q.n_defer].dst = o.host; q.defer[q.n_defer].dstep = o.hstep; q.defer[q.n_defer].off = o.off;
            q.defer[q.n_defer].row_bytes = o.row_bytes; q.defer[q.n_defer].rows = h;
            q.n_defer++;
            ctx->st.staged_bytes += total;
        }
        ctx->st.d2h_bytes += total;
    }
    MSG_CUDA(ctx, cudaMemcpyAsync(q.h_cnt, q.d_cnt, 8 * sizeof(int32_t), cudaMemcpyDeviceToHost, ds));
    MSG_CUDA(ctx, cudaEventRecord(q.done, ds));
    q.l16 = (labels && l16) ? 1 : 0;
    q.used = 1;
    *ticket = slot;
    return MSG_OK;
}

int msg_wait(msg_ctx* ctx, int ticket, int32_t* n_regions)
{
    CTX_ENTER(ctx);
    if (ticket < 0 || ticket >= MSG_MAX_INFLIGHT || !ctx->pend[ticket].used)
        return msg_fail(ctx, MSG_ESTATE, "wait: invalid ticket %d", ticket);
    msg_ctx::pending& q = ctx->pend[ticket];
    q.used = 0;                                      // whatever happens below, the slot is free again
    MSG_CUDA(ctx, cudaEventSynchronize(q.done));
    for (int k = 0; k < q.n_defer; k++)
        host_scatter((uint8_t*)q.defer[k].dst, q.defer[k].dstep, q.defer[k].row_bytes, 0, q.h_out + q.defer[k].off,
                     q.defer[k].row_bytes * (size_t)q.defer[k].rows);
    q.n_defer = 0;
    const int32_t n = q.h_cnt[0];
    if (n_regions) *n_regions = n;
    unsigned long long a, o;
    memcpy(&a, q.h_cnt + 2, 8);
    memcpy(&o, q.h_cnt + 4, 8);
    ctx->st.ms_active_items = a;
    ctx->st.ms_overflow_items = o;
    if (q.l16 && n > 65535) return msg_fail(ctx, MSG_ERANGE, "wait: %d regions do not fit 16-bit labels", n);
    return MSG_OK;
}

// ============================================================================ strip sharding, round 2: device-side tables, merge

int msg_strip_resolve_dense_dev(msg_ctx* ctx, const int32_t* d_gathered, int n_strips, int width, const int* row0, int32_t* d_tables,
                                size_t tables_capacity_ints)
{
    CTX_ENTER(ctx);
    if (!d_gathered || !row0 || !d_tables || n_strips < 1 || n_strips > MSG_MAX_STRIPS || width < 1)
        return msg_fail(ctx, MSG_EINVAL, "strip resolve: bad arguments");
    if (tables_capacity_ints < (size_t)MSG_SHARD_TABLE_HEADER + 2 * (size_t)n_strips * width)
        return msg_fail(ctx, MSG_EINVAL, "strip resolve: d_tables needs MSG_SHARD_TABLE_HEADER + 2 * n_strips * width ints");
    for (int s = 1; s < n_strips; s++)
        if (row0[s] <= row0[s - 1]) return msg_fail(ctx, MSG_EINVAL, "strip resolve: row0 must be ascending");
    return k_strip_resolve(ctx, d_gathered, n_strips, width, row0, d_tables);
}

int msg_strip_finalize_tables_dev(msg_ctx* ctx, int32_t* d_labels, size_t lstep, int w, int rows, int row0, int full_w, int strip,
                                  int n_strips, const int32_t* d_tables)
{
    CTX_ENTER(ctx);
    MSG_TRY(check_img(ctx, d_labels, lstep, w, rows, 4, "labels"));
    if (lstep % 4 || full_w != w || !d_tables || strip < 0 || strip >= n_strips || n_strips > MSG_MAX_STRIPS)
        return msg_fail(ctx, MSG_EINVAL, "strip finalize: bad arguments");
    if (!ctx->d_scratch) return msg_fail(ctx, MSG_ESTATE, "strip finalize: msg_strip_rank_dev must run first on this context");
    return k_strip_finalize_tables(ctx, d_labels, lstep, w, rows, (long long)row0 * full_w, d_tables, n_strips * w, strip);
}

int msg_strip_merge_stats_dev(msg_ctx* ctx, const uint8_t* d_bgr, size_t step, const int32_t* d_labels, size_t lstep, int w, int rows,
                              const int32_t* d_up_row_labels, int n_total, uint32_t* d_area, unsigned long long* d_sum,
                              int32_t* d_pairs, long long pair_cap, int32_t* d_npairs)
{
    CTX_ENTER(ctx);
    MSG_TRY(check_img(ctx, d_bgr, step, w, rows, 3, "strip merge image"));
    MSG_TRY(check_img(ctx, d_labels, lstep, w, rows, 4, "strip merge labels"));
    if (lstep != (size_t)w * 4) return msg_fail(ctx, MSG_EINVAL, "strip merge: labels must be dense");
    if (n_total < 0 || !d_area || !d_sum || !d_pairs || !d_npairs || pair_cap < 0)
        return msg_fail(ctx, MSG_EINVAL, "strip merge: bad table arguments");
    msg_plane s;
    s.w = w; s.rows = rows; s.y0 = 0; s.hfull = rows; s.pitch = msg_align_up(w, 32);
    MSG_TRY(msg_reserve(ctx, (void**)&ctx->d_planes, &ctx->d_planes_cap, (size_t)s.pitch * rows * sizeof(uint32_t)));
    s.p = ctx->d_planes;
    MSG_TRY(k_bgr_to_plane(ctx, d_bgr, step, s));
    return k_strip_merge_stats(ctx, s.p, s.pitch, d_labels, w, rows, d_up_row_labels, n_total, d_area, d_sum, d_pairs, pair_cap, d_npairs);
}

int msg_strip_merge_finish_dev(msg_ctx* ctx, int32_t* d_labels, size_t lstep, int w, int rows, long long full_pixels, int n_total,
                               uint32_t* d_area, unsigned long long* d_sum, const int32_t* d_all_pairs, long long n_all_pairs,
                               int min_size, int color_dist, int32_t* d_n_out)
{
    CTX_ENTER(ctx);
    MSG_TRY(check_img(ctx, d_labels, lstep, w, rows, 4, "strip merge labels"));
    if (lstep != (size_t)w * 4) return msg_fail(ctx, MSG_EINVAL, "strip merge: labels must be dense");
    if (n_total < 0 || !d_area || !d_sum || n_all_pairs < 0 || (n_all_pairs > 0 && !d_all_pairs) || min_size < 0 || color_dist < 0)
        return msg_fail(ctx, MSG_EINVAL, "strip merge: bad arguments");
    if (min_size <= 0 && color_dist <= 0) {
        if (d_n_out) {
            ctx->h_counters[20] = n_total;
            MSG_CUDA(ctx, cudaMemcpyAsync(d_n_out, ctx->h_counters + 20, sizeof(int32_t), cudaMemcpyHostToDevice, ctx->stream));
        }
        return MSG_OK;
    }
    return k_strip_merge_finish(ctx, d_labels, w, rows, full_pixels, n_total, d_area, d_sum, d_all_pairs, n_all_pairs, min_size,
                                color_dist, d_n_out);
}

// ============================================================================ watershed (f1)

int msg_watershed_batch_dev(msg_ctx* ctx, const uint8_t* d_bgr, size_t step, size_t image_stride, int32_t* d_markers, size_t mstep,
                            size_t markers_stride, int w, int h, int count, unsigned long long* d_pops)
{
    CTX_ENTER(ctx);
    if (count < 1 || count > 65535) return msg_fail(ctx, MSG_EINVAL, "watershed: count must be in [1, 65535]");
    MSG_TRY(check_img(ctx, d_bgr, step, w, h, 3, "watershed image"));
    MSG_TRY(check_img(ctx, d_markers, mstep, w, h, 4, "watershed markers"));
    if (mstep % 4 || markers_stride % 4) return msg_fail(ctx, MSG_EINVAL, "watershed: marker steps must be multiples of 4");
    if (count > 1 && (image_stride < step * (size_t)h || markers_stride < mstep * (size_t)h))
        return msg_fail(ctx, MSG_EINVAL, "watershed: batch strides smaller than one image");
    return k_watershed(ctx, d_bgr, step, image_stride, d_markers, mstep, markers_stride, w, h, count, d_pops);
}

int msg_watershed_dev(msg_ctx* ctx, const uint8_t* d_bgr, size_t step, int32_t* d_markers, size_t mstep, int w, int h)
{
    return msg_watershed_batch_dev(ctx, d_bgr, step, 0, d_markers, mstep, 0, w, h, 1, nullptr);
}

int msg_watershed(msg_ctx* ctx, const uint8_t* bgr, size_t step, int32_t* markers, size_t mstep, int w, int h)
{
    CTX_ENTER(ctx);
    MSG_TRY(check_img(ctx, bgr, step, w, h, 3, "watershed image"));
    MSG_TRY(check_img(ctx, markers, mstep, w, h, 4, "watershed markers"));
    const size_t rb = (size_t)w * 3, lb = (size_t)w * 4;
    MSG_TRY(msg_reserve(ctx, (void**)&ctx->d_in, &ctx->d_in_cap, rb * h));
    MSG_TRY(msg_reserve(ctx, (void**)&ctx->d_labels, &ctx->d_labels_cap, lb * h));
    cudaStream_t st = ctx->stream;
    MSG_CUDA(ctx, cudaEventRecord(ctx->ev[0], st));
    MSG_TRY(upload_on(ctx, st, bgr, step, rb, h, ctx->d_in));
    MSG_TRY(upload_on(ctx, st, markers, mstep, lb, h, (uint8_t*)ctx->d_labels));
    MSG_CUDA(ctx, cudaEventRecord(ctx->ev[1], st));
    int rc = k_watershed(ctx, ctx->d_in, rb, 0, ctx->d_labels, lb, 0, w, h, 1, nullptr);
    if (rc != MSG_OK) { cudaStreamSynchronize(st); return rc; }
    MSG_CUDA(ctx, cudaEventRecord(ctx->ev[2], st));
    MSG_TRY(copy_out(ctx, markers, mstep, ctx->d_labels, lb, h));
    MSG_CUDA(ctx, cudaEventRecord(ctx->ev[3], st));
    MSG_CUDA(ctx, cudaStreamSynchronize(st));
    memset(&ctx->tm, 0, sizeof(ctx->tm));
    ctx->tm.h2d_ms = ev_ms(ctx->ev[0], ctx->ev[1]);
    ctx->tm.filter_ms = ev_ms(ctx->ev[1], ctx->ev[2]);
    ctx->tm.d2h_ms = ev_ms(ctx->ev[2], ctx->ev[3]);
    ctx->tm.total_ms = ev_ms(ctx->ev[0], ctx->ev[3]);
    return MSG_OK;
}

int msg_set_profiling(msg_ctx* ctx, int enable)
{
    CTX_ENTER(ctx);
    MSG_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    for (int l = 0; l < MSG_MAX_LEVELS; l++) ctx->prof_pending[l] = 0;
    if (enable) {
        memset(&ctx->prof, 0, sizeof(ctx->prof));
        MSG_CUDA(ctx, cudaMemset(ctx->d_work, 0, MSG_MAX_LEVELS * 4 * sizeof(unsigned long long)));
    }
    ctx->profiling = enable ? 1 : 0;
    return MSG_OK;
}

int msg_get_kernel_profile(msg_ctx* ctx, msg_kernel_profile* out)
{
    CTX_ENTER(ctx);
    if (!out) return MSG_EINVAL;
    MSG_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    for (int l = 0; l < MSG_MAX_LEVELS; l++) prof_fold(ctx, l);
    unsigned long long wk[MSG_MAX_LEVELS * 4];
    MSG_CUDA(ctx, cudaMemcpy(wk, ctx->d_work, sizeof(wk), cudaMemcpyDeviceToHost));
    for (int l = 0; l < MSG_MAX_LEVELS; l++) {
        ctx->prof.tile_tests[l] = wk[4 * l]; ctx->prof.tile_hits[l] = wk[4 * l + 1];
        ctx->prof.overflow_tests[l] = wk[4 * l + 2]; ctx->prof.overflow_hits[l] = wk[4 * l + 3];
    }
    *out = ctx->prof;
    return MSG_OK;
}

int msg_get_stats(msg_ctx* ctx, msg_stats* out)
{
    if (!ctx || !out) return MSG_EINVAL;
    *out = ctx->st;
    return MSG_OK;
}

int msg_debug_get_plane(msg_ctx* ctx, int kind, int level, uint32_t* host_out, size_t cap, int* w, int* h)
{
    CTX_ENTER(ctx);
    if (level < 0 || level > ctx->last_levels || (kind != 0 && kind != 1) || !ctx->d_planes)
        return msg_fail(ctx, MSG_EINVAL, "debug_get_plane: bad kind/level");
    msg_plane p = kind ? ctx->D[level] : ctx->S[level];
    if (w) *w = p.w;
    if (h) *h = p.rows;
    if (!host_out) return MSG_OK;
    if (cap < (size_t)p.w * p.rows) return msg_fail(ctx, MSG_EINVAL, "debug_get_plane: buffer too small");
    MSG_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    MSG_CUDA(ctx, cudaMemcpy2D(host_out, (size_t)p.w * 4, p.p, (size_t)p.pitch * 4, (size_t)p.w * 4, p.rows, cudaMemcpyDeviceToHost));
    return MSG_OK;
}

// ============================================================================ strip sharding

int msg_meanshift_halo_rows(double sp, int max_level, int term_type, int max_count)
{
    if (max_level < 0 || max_level > 8) return -1;
    if (!(term_type & MSG_TERM_COUNT)) max_count = 5;
    if (max_count < 1) max_count = 1;
    if (max_count > 100) max_count = 100;
    // Dependency of an output row on source rows (DESIGN.md "strip sharding").  need(l) = rows of S[l] (level-l
    // units) that a pixel of the level-l result D[l] can depend on, either side:
    //   * its own windows: the centre moves <= r_l = ceil(sp_l) per iteration and max_count windows are evaluated, the k-th one
    //     reaching k * r_l rows from the pixel  ->  max_count * r_l   (windows read the SOURCE plane S[l], never D[l]);
    //   * below the top level, its initial value and its change-mask bit come from D[l+1] within 3 level-(l+1) rows (pyrUp
    //     taps +-1, the 8-neighbour test of the mask rule, the 3x3 dilate); a row of D[l+1] depends on S[l+1] within
    //     need(l+1) rows, a row of S[l+1] on S[l] within 2 rows of its double (pyrDown taps)  ->  2 * (need(l+1) + 3) + 2.
    // The two are alternatives, not a chain: need(l) = max of them.  Defaults: max(50, 2 * (25 + 3) + 2) = 58.
    // The bound is proven tight enough by the bit-identity gates (tests/test_gpu_sharded.py: strips == unsharded call for
    // maxLevel 0 / 1 / 2, several sp and termcrits; tools/shard_large_image.py --verify on real ranks).
    long need = 0;
    for (int l = max_level; l >= 0; l--) {
        double spl = sp / (double)(1 << l);
        if (!(spl >= 1.0)) spl = 1.0;
        const long own = (long)max_count * (long)ceil(spl);
        const long via_up = 2 * (need + 3) + 2;
        need = (l == max_level || own > via_up) ? own : via_up;
    }
    const long a = 1L << max_level;
    long halo0 = (need + a - 1) / a * a + a;   // multiple of 2^max_level (strip origins keep the pyramid phase), one unit of slack
    return (int)halo0;
}

int msg_meanshift_filter_strip_dev(msg_ctx* ctx, const uint8_t* d_src_rows, size_t sstep, int halo_row0, int halo_row1,
                                   uint8_t* d_dst, size_t dstep, int w, int full_h, int row0, int row1, double sp,
                                   double sr, int max_level, int term_type, int max_count, double eps)
{
    CTX_ENTER(ctx);
    if (!(0 <= halo_row0 && halo_row0 <= row0 && row0 < row1 && row1 <= halo_row1 && halo_row1 <= full_h))
        return msg_fail(ctx, MSG_EINVAL, "strip: need 0 <= halo_row0 <= row0 < row1 <= halo_row1 <= full_height");
    MSG_TRY(check_img(ctx, d_src_rows, sstep, w, halo_row1 - halo_row0, 3, "strip src"));
    MSG_TRY(check_img(ctx, d_dst, dstep, w, row1 - row0, 3, "strip dst"));
    ms_config cfg;
    MSG_TRY(ms_validate(ctx, w, full_h, sp, sr, max_level, term_type, max_count, eps, &cfg));
    int a = 1 << max_level;
    if (halo_row0 % a || row0 % a) return msg_fail(ctx, MSG_EINVAL, "strip: row0 and halo_row0 must be multiples of 2^max_level");
    MSG_TRY(ms_run(ctx, d_src_rows, sstep, w, full_h, halo_row0, halo_row1, cfg));
    MSG_TRY(k_plane_to_bgr(ctx, ctx->D[0], row0 - halo_row0, row1 - row0, d_dst, dstep));
    return MSG_OK;
}

int msg_label_strip_dev(msg_ctx* ctx, const uint8_t* d_bgr_rows, size_t step, int32_t* d_labels, size_t lstep, int w,
                        int rows, int row0, int full_w, int lo_diff)
{
    CTX_ENTER(ctx);
    MSG_TRY(check_img(ctx, d_bgr_rows, step, w, rows, 3, "label strip src"));
    MSG_TRY(check_img(ctx, d_labels, lstep, w, rows, 4, "labels"));
    if (lstep != (size_t)w * 4) return msg_fail(ctx, MSG_EINVAL, "label strip: labels must be dense");
    if (lo_diff < 0 || full_w != w) return msg_fail(ctx, MSG_EINVAL, "label strip: lo_diff >= 0 and full_width == width required");
    if (((long long)row0 + rows) * (long long)full_w > 0x7fffffffLL - 1)
        return msg_fail(ctx, MSG_EINVAL, "label strip: global label would overflow int32");
    msg_plane s;
    s.w = w; s.rows = rows; s.y0 = 0; s.hfull = rows; s.pitch = msg_align_up(w, 32);
    MSG_TRY(msg_reserve(ctx, (void**)&ctx->d_planes, &ctx->d_planes_cap, (size_t)s.pitch * rows * sizeof(uint32_t)));
    s.p = ctx->d_planes;
    MSG_TRY(k_bgr_to_plane(ctx, d_bgr_rows, step, s));
    return k_ccl_color(ctx, s.p, s.pitch, w, rows, lo_diff, 4, d_labels, (int64_t)row0 * full_w, full_w);
}

int msg_seam_pairs_dev(msg_ctx* ctx, const uint8_t* up_bgr, const int32_t* up_lab, const uint8_t* lo_bgr,
                       const int32_t* lo_lab, int w, int lo_diff, int32_t* d_pairs, int32_t* d_count)
{
    CTX_ENTER(ctx);
    if (!up_bgr || !up_lab || !lo_bgr || !lo_lab || !d_pairs || !d_count || w <= 0 || lo_diff < 0)
        return msg_fail(ctx, MSG_EINVAL, "seam_pairs: bad argument");
    return k_seam_pairs(ctx, up_bgr, up_lab, lo_bgr, lo_lab, w, lo_diff, d_pairs, d_count);
}

int msg_apply_label_map_dev(msg_ctx* ctx, int32_t* d_labels, size_t lstep, int w, int rows, const int32_t* d_from,
                            const int32_t* d_to, int n_map)
{
    CTX_ENTER(ctx);
    MSG_TRY(check_img(ctx, d_labels, lstep, w, rows, 4, "labels"));
    if (n_map < 0 || (n_map > 0 && (!d_from || !d_to))) return msg_fail(ctx, MSG_EINVAL, "apply_label_map: bad map");
    return k_apply_map(ctx, d_labels, lstep, w, rows, d_from, d_to, n_map);
}

int msg_strip_rank_dev(msg_ctx* ctx, const int32_t* d_labels, size_t lstep, int w, int rows, int row0, int full_w,
                       int32_t* d_count)
{
    CTX_ENTER(ctx);
    MSG_TRY(check_img(ctx, d_labels, lstep, w, rows, 4, "labels"));
    if (!d_count || full_w != w || row0 < 0 || lstep % 4) return msg_fail(ctx, MSG_EINVAL, "strip_rank: bad argument");
    return k_strip_rank(ctx, d_labels, lstep, w, rows, (long long)row0 * full_w, d_count);
}

int msg_strip_query_dense_dev(msg_ctx* ctx, const int32_t* d_query, int nq, int w, int rows, int row0, int full_w, int offset,
                              int32_t* d_out)
{
    CTX_ENTER(ctx);
    if (nq < 0 || (nq > 0 && (!d_query || !d_out)) || full_w != w) return msg_fail(ctx, MSG_EINVAL, "strip_query: bad argument");
    if (!ctx->d_scratch) return msg_fail(ctx, MSG_ESTATE, "strip_query: call msg_strip_rank_dev first");
    return k_strip_query(ctx, d_query, nq, w, rows, (long long)row0 * full_w, offset, d_out);
}

int msg_strip_apply_dense_dev(msg_ctx* ctx, int32_t* d_labels, size_t lstep, int w, int rows, int row0, int full_w, int offset,
                              const int32_t* d_rlab, const int32_t* d_rdense, int nr)
{
    CTX_ENTER(ctx);
    MSG_TRY(check_img(ctx, d_labels, lstep, w, rows, 4, "labels"));
    if (nr < 0 || (nr > 0 && (!d_rlab || !d_rdense)) || full_w != w || lstep % 4) return msg_fail(ctx, MSG_EINVAL, "strip_apply_dense: bad argument");
    if (!ctx->d_scratch) return msg_fail(ctx, MSG_ESTATE, "strip_apply_dense: call msg_strip_rank_dev first");
    return k_strip_apply_dense(ctx, d_labels, lstep, w, rows, (long long)row0 * full_w, offset, d_rlab, d_rdense, nr);
}

int msg_seam_quads_dev(msg_ctx* ctx, const uint8_t* up_bgr, const int32_t* up_lab, const int32_t* up_rank1, const uint8_t* lo_bgr,
                       const int32_t* lo_lab, int w, int lo_diff, int rows, int row0, int full_w, int32_t* d_quads, int32_t* d_count)
{
    CTX_ENTER(ctx);
    if (!up_bgr || !up_lab || !up_rank1 || !lo_bgr || !lo_lab || !d_quads || !d_count || w <= 0 || lo_diff < 0 || rows <= 0 ||
        row0 < 0 || full_w != w)
        return msg_fail(ctx, MSG_EINVAL, "seam_quads: bad argument");
    if (!ctx->d_scratch) return msg_fail(ctx, MSG_ESTATE, "seam_quads: call msg_strip_rank_dev first");
    return k_seam_quads(ctx, up_bgr, up_lab, up_rank1, lo_bgr, lo_lab, w, lo_diff, rows, (long long)row0 * full_w, d_quads, d_count);
}

int msg_strip_finalize_dense_dev(msg_ctx* ctx, int32_t* d_labels, size_t lstep, int w, int rows, int row0, int full_w, int offset,
                                 const int32_t* d_frm, const int32_t* d_dense, int n_map, int frm_lo)
{
    CTX_ENTER(ctx);
    MSG_TRY(check_img(ctx, d_labels, lstep, w, rows, 4, "labels"));
    if (n_map < 0 || (n_map > 0 && (!d_frm || !d_dense)) || frm_lo < 0 || frm_lo > n_map || full_w != w || lstep % 4)
        return msg_fail(ctx, MSG_EINVAL, "strip_finalize_dense: bad argument");
    if (!ctx->d_scratch) return msg_fail(ctx, MSG_ESTATE, "strip_finalize_dense: call msg_strip_rank_dev first");
    return k_strip_finalize(ctx, d_labels, lstep, w, rows, (long long)row0 * full_w, offset, d_frm, d_dense, n_map, frm_lo);
}

// ============================================================================ pre-filters (8(f2))

static int filter_host(msg_ctx* ctx, const uint8_t* src, size_t sstep, int in_ch, uint8_t* dst, size_t dstep, int out_ch, int w,
                       int h, int which, const int8_t* taps, int krows, int kcols, int ksize)
{
    size_t rin = (size_t)w * in_ch, rout = (size_t)w * out_ch;
    MSG_TRY(copy_in(ctx, src, sstep, rin, h, &ctx->d_in, &ctx->d_in_cap));
    MSG_TRY(msg_reserve(ctx, (void**)&ctx->d_out, &ctx->d_out_cap, rout * h));
    if (which == 0) MSG_TRY(k_sharpen(ctx, ctx->d_in, rin, ctx->d_out, rout, w, h, taps, krows, kcols));
    else if (which == 1) MSG_TRY(k_gray(ctx, ctx->d_in, rin, ctx->d_out, rout, w, h));
    else MSG_TRY(k_median(ctx, ctx->d_in, rin, ctx->d_out, rout, w, h, ksize));
    MSG_TRY(copy_out(ctx, dst, dstep, ctx->d_out, rout, h));
    MSG_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return MSG_OK;
}

int msg_laplacian_sharpen(msg_ctx* ctx, const uint8_t* src, size_t sstep, uint8_t* dst, size_t dstep, int w, int h,
                          const int8_t* taps, int krows, int kcols)
{
    CTX_ENTER(ctx);
    MSG_TRY(check_img(ctx, src, sstep, w, h, 3, "sharpen src"));
    MSG_TRY(check_img(ctx, dst, dstep, w, h, 3, "sharpen dst"));
    if (!taps || krows < 1 || kcols < 1 || !(krows & 1) || !(kcols & 1) || krows * kcols > 1024)
        return msg_fail(ctx, MSG_EINVAL, "sharpen: kernel must be odd x odd with at most 1024 taps");
    return filter_host(ctx, src, sstep, 3, dst, dstep, 3, w, h, 0, taps, krows, kcols, 0);
}

int msg_bgr2gray(msg_ctx* ctx, const uint8_t* src, size_t sstep, uint8_t* dst, size_t dstep, int w, int h)
{
    CTX_ENTER(ctx);
    MSG_TRY(check_img(ctx, src, sstep, w, h, 3, "cvtColor src"));
    MSG_TRY(check_img(ctx, dst, dstep, w, h, 1, "cvtColor dst"));
    return filter_host(ctx, src, sstep, 3, dst, dstep, 1, w, h, 1, nullptr, 0, 0, 0);
}

int msg_median_blur(msg_ctx* ctx, const uint8_t* src, size_t sstep, uint8_t* dst, size_t dstep, int w, int h, int ksize)
{
    CTX_ENTER(ctx);
    MSG_TRY(check_img(ctx, src, sstep, w, h, 1, "medianBlur src"));
    MSG_TRY(check_img(ctx, dst, dstep, w, h, 1, "medianBlur dst"));
    if (ksize < 1 || ksize > 127 || !(ksize & 1)) return msg_fail(ctx, MSG_EINVAL, "medianBlur: ksize must be odd and in [1,127] (got %d)", ksize);
    return filter_host(ctx, src, sstep, 1, dst, dstep, 1, w, h, 2, nullptr, 0, 0, ksize);
}

// ============================================================================ shape-method seeds (8(f3), row a7)

static int canny_thresholds(msg_ctx* ctx, double low, double high, int* lo, int* hi)
{
    if (!(low == low) || !(high == high)) return msg_fail(ctx, MSG_EINVAL, "Canny: thresholds must be numbers");
    if (low > high) { double t = low; low = high; high = t; }                 // cv::Canny swaps them
    *lo = (int)floor(fmin(fmax(low, -1.0), 1e9));
    *hi = (int)floor(fmin(fmax(high, -1.0), 1e9));
    return MSG_OK;
}

// d_src: gray plane (step sstep); d_dst: edges 0/255.  Uses d_aux[0..2n) (classes, flags) and d_labels as scratch.
static int canny_dev(msg_ctx* ctx, const uint8_t* d_src, size_t sstep, uint8_t* d_dst, size_t dstep, int w, int h, int lo, int hi,
                     uint8_t* d_cls, uint8_t* d_flag)
{
    MSG_TRY(msg_reserve(ctx, (void**)&ctx->d_labels, &ctx->d_labels_cap, (size_t)w * h * 4));
    MSG_TRY(k_canny_nms(ctx, d_src, sstep, d_cls, (size_t)w, w, h, lo, hi));
    return k_hysteresis(ctx, d_cls, w, h, ctx->d_labels, d_flag, d_dst, dstep);
}

int msg_canny(msg_ctx* ctx, const uint8_t* src, size_t sstep, uint8_t* dst, size_t dstep, int w, int h, double low, double high)
{
    CTX_ENTER(ctx);
    MSG_TRY(check_img(ctx, src, sstep, w, h, 1, "Canny src"));
    MSG_TRY(check_img(ctx, dst, dstep, w, h, 1, "Canny dst"));
    int lo, hi;
    MSG_TRY(canny_thresholds(ctx, low, high, &lo, &hi));
    size_t n = (size_t)w * h;
    MSG_TRY(copy_in(ctx, src, sstep, (size_t)w, h, &ctx->d_in, &ctx->d_in_cap));
    MSG_TRY(msg_reserve(ctx, (void**)&ctx->d_out, &ctx->d_out_cap, n));
    MSG_TRY(msg_reserve(ctx, (void**)&ctx->d_aux, &ctx->d_aux_cap, 2 * n));
    MSG_TRY(canny_dev(ctx, ctx->d_in, (size_t)w, ctx->d_out, (size_t)w, w, h, lo, hi, ctx->d_aux, ctx->d_aux + n));
    MSG_TRY(copy_out(ctx, dst, dstep, ctx->d_out, (size_t)w, h));
    MSG_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return MSG_OK;
}

int msg_dilate(msg_ctx* ctx, const uint8_t* src, size_t sstep, uint8_t* dst, size_t dstep, int w, int h, int kw, int kh)
{
    CTX_ENTER(ctx);
    MSG_TRY(check_img(ctx, src, sstep, w, h, 1, "dilate src"));
    MSG_TRY(check_img(ctx, dst, dstep, w, h, 1, "dilate dst"));
    if (kw < 1 || kh < 1 || kw > 63 || kh > 63) return msg_fail(ctx, MSG_EINVAL, "dilate: kernel must be 1..63 x 1..63 (got %dx%d)", kw, kh);
    MSG_TRY(copy_in(ctx, src, sstep, (size_t)w, h, &ctx->d_in, &ctx->d_in_cap));
    MSG_TRY(msg_reserve(ctx, (void**)&ctx->d_out, &ctx->d_out_cap, (size_t)w * h));
    MSG_TRY(k_dilate(ctx, ctx->d_in, (size_t)w, ctx->d_out, (size_t)w, w, h, kw, kh));
    MSG_TRY(copy_out(ctx, dst, dstep, ctx->d_out, (size_t)w, h));
    MSG_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return MSG_OK;
}

int msg_subtract(msg_ctx* ctx, const uint8_t* a, size_t astep, const uint8_t* b, size_t bstep, uint8_t* dst, size_t dstep, int w,
                 int h)
{
    CTX_ENTER(ctx);
    MSG_TRY(check_img(ctx, a, astep, w, h, 1, "subtract src1"));
    MSG_TRY(check_img(ctx, b, bstep, w, h, 1, "subtract src2"));
    MSG_TRY(check_img(ctx, dst, dstep, w, h, 1, "subtract dst"));
    size_t n = (size_t)w * h;
    MSG_TRY(copy_in(ctx, a, astep, (size_t)w, h, &ctx->d_in, &ctx->d_in_cap));
    MSG_TRY(msg_reserve(ctx, (void**)&ctx->d_aux, &ctx->d_aux_cap, n));
    MSG_TRY(upload_on(ctx, ctx->stream, b, bstep, (size_t)w, h, ctx->d_aux));
    MSG_TRY(msg_reserve(ctx, (void**)&ctx->d_out, &ctx->d_out_cap, n));
    MSG_TRY(k_subtract(ctx, ctx->d_in, (size_t)w, ctx->d_aux, (size_t)w, ctx->d_out, (size_t)w, w, h));
    MSG_TRY(copy_out(ctx, dst, dstep, ctx->d_out, (size_t)w, h));
    MSG_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return MSG_OK;
}

int msg_copy_masked(msg_ctx* ctx, const uint8_t* src, size_t sstep, const uint8_t* mask, size_t mstep, uint8_t* dst, size_t dstep,
                    int w, int h)
{
    CTX_ENTER(ctx);
    MSG_TRY(check_img(ctx, src, sstep, w, h, 3, "copyTo src"));
    MSG_TRY(check_img(ctx, mask, mstep, w, h, 1, "copyTo mask"));
    MSG_TRY(check_img(ctx, dst, dstep, w, h, 3, "copyTo dst"));
    const size_t n = (size_t)w * h;
    MSG_TRY(msg_reserve(ctx, (void**)&ctx->d_in, &ctx->d_in_cap, 3 * n));
    MSG_TRY(msg_reserve(ctx, (void**)&ctx->d_aux, &ctx->d_aux_cap, n));
    MSG_TRY(msg_reserve(ctx, (void**)&ctx->d_out, &ctx->d_out_cap, 3 * n));
    MSG_TRY(upload_on(ctx, ctx->stream, src, sstep, (size_t)w * 3, h, ctx->d_in));
    MSG_TRY(upload_on(ctx, ctx->stream, mask, mstep, (size_t)w, h, ctx->d_aux));
    MSG_TRY(k_copy_masked(ctx, ctx->d_in, (size_t)w * 3, ctx->d_aux, (size_t)w, ctx->d_out, (size_t)w * 3, w, h));
    MSG_TRY(copy_out(ctx, dst, dstep, ctx->d_out, (size_t)w * 3, h));
    MSG_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return MSG_OK;
}

// d_stages (optional): 4 dense planes of w*h bytes: blurred gray, Canny edges, dilate-dilate-subtract band, its 3x3 median
int msg_shape_seeds_dev(msg_ctx* ctx, const uint8_t* d_bgr, size_t step, int w, int h, int median_ksize, double low, double high,
                        int32_t* d_markers, size_t lstep, int32_t* d_n, uint8_t* d_stages)
{
    CTX_ENTER(ctx);
    MSG_TRY(check_img(ctx, d_bgr, step, w, h, 3, "shape seeds src"));
    MSG_TRY(check_img(ctx, d_markers, lstep, w, h, 4, "markers"));
    if (median_ksize < 1 || median_ksize > 127 || !(median_ksize & 1))
        return msg_fail(ctx, MSG_EINVAL, "shape seeds: median ksize must be odd and in [1,127] (got %d)", median_ksize);
    if (!d_n) return msg_fail(ctx, MSG_EINVAL, "shape seeds: null count pointer");
    int lo, hi;
    MSG_TRY(canny_thresholds(ctx, low, high, &lo, &hi));
    size_t n = (size_t)w * h;
    MSG_TRY(msg_reserve(ctx, (void**)&ctx->d_aux, &ctx->d_aux_cap, 7 * n));
    uint8_t* a = ctx->d_aux;                     // [cls][flag][gray / d3][blurred][edges][dde][dde3]
    uint8_t *cls = a, *flag = a + n, *gray = a + 2 * n;
    uint8_t* blurred = d_stages ? d_stages : a + 3 * n;
    uint8_t* edges = d_stages ? d_stages + n : a + 4 * n;
    uint8_t* dde = d_stages ? d_stages + 2 * n : a + 5 * n;
    uint8_t* dde3 = d_stages ? d_stages + 3 * n : a + 6 * n;
    MSG_TRY(k_gray(ctx, d_bgr, step, gray, (size_t)w, w, h));                                   // PictureService.java:405
    MSG_TRY(k_median(ctx, gray, (size_t)w, blurred, (size_t)w, w, h, median_ksize));            // :408
    MSG_TRY(canny_dev(ctx, blurred, (size_t)w, edges, (size_t)w, w, h, lo, hi, cls, flag));     // :416
    uint8_t* d3 = gray;                                                                         // gray is dead from here on
    MSG_TRY(k_dilate(ctx, edges, (size_t)w, d3, (size_t)w, w, h, 3, 3));                        // :428
    MSG_TRY(k_dilate(ctx, d3, (size_t)w, cls, (size_t)w, w, h, 5, 5));                          // :429 (classes are dead)
    MSG_TRY(k_subtract(ctx, cls, (size_t)w, d3, (size_t)w, dde, (size_t)w, w, h));              // :430
    MSG_TRY(k_median(ctx, dde, (size_t)w, dde3, (size_t)w, w, h, 3));                           // :436
    return msg_connected_components_dev(ctx, dde3, (size_t)w, d_markers, lstep, w, h, 8, d_n);  // :441-442
}

int msg_shape_seeds(msg_ctx* ctx, const uint8_t* src, size_t sstep, int w, int h, int median_ksize, double low, double high,
                    int32_t* markers, size_t lstep, int32_t* n_labels, uint8_t* stages, size_t stage_step)
{
    CTX_ENTER(ctx);
    MSG_TRY(check_img(ctx, src, sstep, w, h, 3, "shape seeds src"));
    MSG_TRY(check_img(ctx, markers, lstep, w, h, 4, "markers"));
    if (stages) MSG_TRY(check_img(ctx, stages, stage_step, w, h, 1, "shape seeds stages"));
    size_t n = (size_t)w * h;
    MSG_TRY(copy_in(ctx, src, sstep, (size_t)w * 3, h, &ctx->d_in, &ctx->d_in_cap));
    MSG_TRY(msg_reserve(ctx, (void**)&ctx->d_out, &ctx->d_out_cap, 4 * n));
    MSG_TRY(msg_reserve(ctx, (void**)&ctx->d_out2, &ctx->d_out2_cap, 4 * n));
    int32_t* d_mark = (int32_t*)ctx->d_out2;
    MSG_TRY(msg_shape_seeds_dev(ctx, ctx->d_in, (size_t)w * 3, w, h, median_ksize, low, high, d_mark, (size_t)w * 4,
                                ctx->d_counters + 16, ctx->d_out));
    MSG_TRY(copy_out(ctx, markers, lstep, d_mark, (size_t)w * 4, h));
    if (stages)
        for (int k = 0; k < 4; k++)
            MSG_TRY(copy_out(ctx, stages + (size_t)k * stage_step * h, stage_step, ctx->d_out + (size_t)k * n, (size_t)w, h));
    return finish_count(ctx, n_labels);
}

// ============================================================================ colour-method seeds (8(f3), rows a6 / a4)

// small device tables live in d_small: [0,1024) histogram, [1024, 1024+64) scalars (Otsu threshold, max, min/max pair),
// [2048, ...) circle spans / bilateral tables
static int small_reserve(msg_ctx* ctx, size_t extra)
{
    return msg_reserve(ctx, (void**)&ctx->d_small, &ctx->d_small_cap, 2048 + extra);
}
#define SMALL_HIST(ctx) ((unsigned*)(ctx)->d_small)
#define SMALL_THRESH(ctx) ((int32_t*)((ctx)->d_small + 1024))
#define SMALL_MAX(ctx) ((float*)((ctx)->d_small + 1024 + 8))
#define SMALL_MM(ctx) ((float*)((ctx)->d_small + 1024 + 16))
#define SMALL_TABLES(ctx) ((ctx)->d_small + 2048)

int msg_white_to_black(msg_ctx* ctx, const uint8_t* src, size_t sstep, uint8_t* dst, size_t dstep, int w, int h)
{
    CTX_ENTER(ctx);
    MSG_TRY(check_img(ctx, src, sstep, w, h, 3, "white_to_black src"));
    MSG_TRY(check_img(ctx, dst, dstep, w, h, 3, "white_to_black dst"));
    size_t rb = (size_t)w * 3;
    MSG_TRY(copy_in(ctx, src, sstep, rb, h, &ctx->d_in, &ctx->d_in_cap));
    MSG_TRY(msg_reserve(ctx, (void**)&ctx->d_out, &ctx->d_out_cap, rb * h));
    MSG_TRY(k_white_to_black(ctx, ctx->d_in, rb, ctx->d_out, rb, w, h));
    MSG_TRY(copy_out(ctx, dst, dstep, ctx->d_out, rb, h));
    MSG_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return MSG_OK;
}

static int clamp_to_u8_range(double v) { return (int)floor(fmin(fmax(v, -1.0), 256.0)); }

int msg_threshold(msg_ctx* ctx, const uint8_t* src, size_t sstep, uint8_t* dst, size_t dstep, int w, int h, double thresh,
                  double maxval, int type, double* used)
{
    CTX_ENTER(ctx);
    MSG_TRY(check_img(ctx, src, sstep, w, h, 1, "threshold src"));
    MSG_TRY(check_img(ctx, dst, dstep, w, h, 1, "threshold dst"));
    if (type != MSG_THRESH_BINARY && type != (MSG_THRESH_BINARY | MSG_THRESH_OTSU))
        return msg_fail(ctx, MSG_EINVAL, "threshold: only THRESH_BINARY, optionally | THRESH_OTSU (got type %d)", type);
    if (!(thresh == thresh) || !(maxval == maxval)) return msg_fail(ctx, MSG_EINVAL, "threshold: thresh / maxval must be numbers");
    int imax = (int)lrint(fmin(fmax(maxval, 0.0), 255.0));          // saturate_cast<uchar>(cvRound(maxval))
    MSG_TRY(small_reserve(ctx, 0));
    MSG_TRY(copy_in(ctx, src, sstep, (size_t)w, h, &ctx->d_in, &ctx->d_in_cap));
    MSG_TRY(msg_reserve(ctx, (void**)&ctx->d_out, &ctx->d_out_cap, (size_t)w * h));
    int t = clamp_to_u8_range(thresh);
    if (type & MSG_THRESH_OTSU) {
        MSG_TRY(k_otsu(ctx, ctx->d_in, (size_t)w, w, h, SMALL_HIST(ctx), SMALL_THRESH(ctx)));
        MSG_TRY(k_threshold_u8(ctx, ctx->d_in, (size_t)w, ctx->d_out, (size_t)w, w, h, SMALL_THRESH(ctx), 0, imax));
        MSG_CUDA(ctx, cudaMemcpyAsync(&t, SMALL_THRESH(ctx), sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
    } else {
        MSG_TRY(k_threshold_u8(ctx, ctx->d_in, (size_t)w, ctx->d_out, (size_t)w, w, h, nullptr, t, imax));
    }
    MSG_TRY(copy_out(ctx, dst, dstep, ctx->d_out, (size_t)w, h));
    MSG_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    if (used) *used = (double)t;
    return MSG_OK;
}

int msg_distance_transform_max_width(msg_ctx* ctx) { return ctx ? k_distance_transform_max_width(ctx) : 0; }

int msg_distance_transform(msg_ctx* ctx, const uint8_t* src, size_t sstep, float* dst, size_t dstep, int w, int h, int dist_type,
                           int mask_size)
{
    CTX_ENTER(ctx);
    MSG_TRY(check_img(ctx, src, sstep, w, h, 1, "distanceTransform src"));
    MSG_TRY(check_img(ctx, dst, dstep, w, h, 4, "distanceTransform dst"));
    if (dist_type != 2 || mask_size != 5)
        return msg_fail(ctx, MSG_EINVAL, "distanceTransform: only CV_DIST_L2 (2) with mask size 5 (got %d, %d)", dist_type, mask_size);
    MSG_TRY(small_reserve(ctx, 0));
    MSG_TRY(copy_in(ctx, src, sstep, (size_t)w, h, &ctx->d_in, &ctx->d_in_cap));
    MSG_TRY(msg_reserve(ctx, (void**)&ctx->d_out, &ctx->d_out_cap, (size_t)w * h * 4));
    MSG_CUDA(ctx, cudaEventRecord(ctx->ev[0], ctx->stream));
    MSG_TRY(k_distance_transform(ctx, ctx->d_in, (size_t)w, (float*)ctx->d_out, w, h, SMALL_MAX(ctx)));
    MSG_CUDA(ctx, cudaEventRecord(ctx->ev[1], ctx->stream));
    MSG_TRY(copy_out(ctx, dst, dstep, ctx->d_out, (size_t)w * 4, h));
    MSG_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    memset(&ctx->tm, 0, sizeof(ctx->tm));
    ctx->tm.filter_ms = ev_ms(ctx->ev[0], ctx->ev[1]);          // the kernel alone (msg_get_timings)
    return MSG_OK;
}

// 32F plane in, 32F plane out through d_in / d_out (which: 0 normalise, 1 threshold, 2 dilate)
static int f32_host(msg_ctx* ctx, const float* src, size_t sstep, float* dst, size_t dstep, int w, int h, int which, double a,
                    double b, int kw, int kh)
{
    size_t rb = (size_t)w * 4;
    MSG_TRY(small_reserve(ctx, 0));
    MSG_TRY(copy_in(ctx, src, sstep, rb, h, &ctx->d_in, &ctx->d_in_cap));
    MSG_TRY(msg_reserve(ctx, (void**)&ctx->d_out, &ctx->d_out_cap, rb * h));
    if (which == 0) MSG_TRY(k_normalize_minmax_f32(ctx, (const float*)ctx->d_in, (float*)ctx->d_out, w, h, a, b, SMALL_MM(ctx)));
    else if (which == 1) MSG_TRY(k_threshold_f32(ctx, (const float*)ctx->d_in, (float*)ctx->d_out, w, h, (float)a, (float)b));
    else MSG_TRY(k_dilate_f32(ctx, (const float*)ctx->d_in, (float*)ctx->d_out, w, h, kw, kh));
    MSG_TRY(copy_out(ctx, dst, dstep, ctx->d_out, rb, h));
    MSG_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return MSG_OK;
}

int msg_normalize_minmax(msg_ctx* ctx, const float* src, size_t sstep, float* dst, size_t dstep, int w, int h, double alpha,
                         double beta)
{
    CTX_ENTER(ctx);
    MSG_TRY(check_img(ctx, src, sstep, w, h, 4, "normalize src"));
    MSG_TRY(check_img(ctx, dst, dstep, w, h, 4, "normalize dst"));
    if (!(alpha == alpha) || !(beta == beta)) return msg_fail(ctx, MSG_EINVAL, "normalize: alpha / beta must be numbers");
    return f32_host(ctx, src, sstep, dst, dstep, w, h, 0, alpha, beta, 0, 0);
}

int msg_threshold_f32(msg_ctx* ctx, const float* src, size_t sstep, float* dst, size_t dstep, int w, int h, double thresh,
                      double maxval)
{
    CTX_ENTER(ctx);
    MSG_TRY(check_img(ctx, src, sstep, w, h, 4, "threshold src"));
    MSG_TRY(check_img(ctx, dst, dstep, w, h, 4, "threshold dst"));
    if (!(thresh == thresh) || !(maxval == maxval)) return msg_fail(ctx, MSG_EINVAL, "threshold: thresh / maxval must be numbers");
    return f32_host(ctx, src, sstep, dst, dstep, w, h, 1, thresh, maxval, 0, 0);
}

int msg_dilate_f32(msg_ctx* ctx, const float* src, size_t sstep, float* dst, size_t dstep, int w, int h, int kw, int kh)
{
    CTX_ENTER(ctx);
    MSG_TRY(check_img(ctx, src, sstep, w, h, 4, "dilate src"));
    MSG_TRY(check_img(ctx, dst, dstep, w, h, 4, "dilate dst"));
    if (kw < 1 || kh < 1 || kw > 63 || kh > 63) return msg_fail(ctx, MSG_EINVAL, "dilate: kernel must be 1..63 x 1..63 (got %dx%d)", kw, kh);
    return f32_host(ctx, src, sstep, dst, dstep, w, h, 2, 0, 0, kw, kh);
}

int msg_convert_f32_to_u8(msg_ctx* ctx, const float* src, size_t sstep, uint8_t* dst, size_t dstep, int w, int h)
{
    CTX_ENTER(ctx);
    MSG_TRY(check_img(ctx, src, sstep, w, h, 4, "convertTo src"));
    MSG_TRY(check_img(ctx, dst, dstep, w, h, 1, "convertTo dst"));
    MSG_TRY(copy_in(ctx, src, sstep, (size_t)w * 4, h, &ctx->d_in, &ctx->d_in_cap));
    MSG_TRY(msg_reserve(ctx, (void**)&ctx->d_out, &ctx->d_out_cap, (size_t)w * h));
    MSG_TRY(k_f32_to_u8(ctx, (const float*)ctx->d_in, ctx->d_out, (size_t)w, w, h));
    MSG_TRY(copy_out(ctx, dst, dstep, ctx->d_out, (size_t)w, h));
    MSG_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return MSG_OK;
}

int msg_contour_markers(msg_ctx* ctx, const uint8_t* mask, size_t step, int32_t* markers, size_t mstep, int w, int h,
                        int32_t* n_contours)
{
    CTX_ENTER(ctx);
    MSG_TRY(check_img(ctx, mask, step, w, h, 1, "findContours image"));
    MSG_TRY(check_img(ctx, markers, mstep, w, h, 4, "markers"));
    MSG_TRY(copy_in(ctx, mask, step, (size_t)w, h, &ctx->d_in, &ctx->d_in_cap));
    MSG_TRY(msg_reserve(ctx, (void**)&ctx->d_out2, &ctx->d_out2_cap, (size_t)w * h * 4));
    int32_t n = 0;
    MSG_TRY(k_contour_markers(ctx, ctx->d_in, (size_t)w, w, h, (int32_t*)ctx->d_out2, (size_t)w * 4, &n));
    MSG_TRY(copy_out(ctx, markers, mstep, ctx->d_out2, (size_t)w * 4, h));
    MSG_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    if (n_contours) *n_contours = n;
    return MSG_OK;
}

static int circle_dev(msg_ctx* ctx, int32_t* d_img, size_t step, int w, int h, int cx, int cy, int radius, int32_t value)
{
    if (radius < 0 || radius > 16384) return msg_fail(ctx, MSG_EINVAL, "circle: radius must be in [0, 16384] (got %d)", radius);
    size_t ints = (size_t)12 * (radius + 2);
    MSG_TRY(small_reserve(ctx, ints * sizeof(int)));
    int* h_spans = (int*)malloc(ints * sizeof(int));
    if (!h_spans) return msg_fail(ctx, MSG_ENOMEM, "circle: out of host memory");
    int rc = k_circle_filled_i32(ctx, d_img, step, w, h, cx, cy, radius, value, (int*)SMALL_TABLES(ctx), h_spans);
    if (rc == MSG_OK && cudaStreamSynchronize(ctx->stream) != cudaSuccess) rc = msg_fail(ctx, MSG_ECUDA, "circle: stream sync failed");
    free(h_spans);
    return rc;
}

int msg_circle_filled(msg_ctx* ctx, int32_t* img, size_t step, int w, int h, int cx, int cy, int radius, int32_t value)
{
    CTX_ENTER(ctx);
    MSG_TRY(check_img(ctx, img, step, w, h, 4, "circle img"));
    MSG_TRY(copy_in(ctx, img, step, (size_t)w * 4, h, &ctx->d_in, &ctx->d_in_cap));
    MSG_TRY(circle_dev(ctx, (int32_t*)ctx->d_in, (size_t)w * 4, w, h, cx, cy, radius, value));
    MSG_TRY(copy_out(ctx, img, step, ctx->d_in, (size_t)w * 4, h));
    MSG_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return MSG_OK;
}

int msg_color_seeds_dev(msg_ctx* ctx, const uint8_t* d_bgr, size_t step, int w, int h, const int8_t* taps, int krows, int kcols,
                        double peak_thresh, int32_t* d_markers, size_t mstep, int32_t* n_contours, uint8_t* d_sharp, uint8_t* d_bw,
                        float* d_dist, uint8_t* d_peaks)
{
    CTX_ENTER(ctx);
    MSG_TRY(check_img(ctx, d_bgr, step, w, h, 3, "colour seeds src"));
    MSG_TRY(check_img(ctx, d_markers, mstep, w, h, 4, "markers"));
    if (!taps || krows < 1 || kcols < 1 || !(krows & 1) || !(kcols & 1) || krows * kcols > 1024)
        return msg_fail(ctx, MSG_EINVAL, "colour seeds: sharpen kernel must be odd x odd with at most 1024 taps");
    if (!(peak_thresh == peak_thresh)) return msg_fail(ctx, MSG_EINVAL, "colour seeds: peak threshold must be a number");
    if (!n_contours) return msg_fail(ctx, MSG_EINVAL, "colour seeds: null count pointer");
    const size_t n = (size_t)w * h;
    MSG_TRY(small_reserve(ctx, 256));
    // d_aux: [black 3n][sharp 3n][gray n][bw n][peaks n][pad][dist 4n][tmpA 4n][tmpB 4n]
    MSG_TRY(msg_reserve(ctx, (void**)&ctx->d_aux, &ctx->d_aux_cap, 21 * n + 64));
    uint8_t* a = ctx->d_aux;
    uint8_t* black = a;
    uint8_t* sharp = d_sharp ? d_sharp : a + 3 * n;
    uint8_t* gray = a + 6 * n;
    uint8_t* bw = d_bw ? d_bw : a + 7 * n;
    uint8_t* peaks = d_peaks ? d_peaks : a + 8 * n;
    float* f0 = (float*)(a + ((9 * n + 15) & ~(size_t)15));
    float* dist = d_dist ? d_dist : f0;
    float* tmpA = f0 + n;
    float* tmpB = f0 + 2 * n;
    const size_t rb3 = (size_t)w * 3;
    MSG_TRY(k_white_to_black(ctx, d_bgr, step, black, rb3, w, h));                                   // PictureService.java:309-318
    MSG_TRY(k_sharpen(ctx, black, rb3, sharp, rb3, w, h, taps, krows, kcols));                        // :323-333
    MSG_TRY(k_gray(ctx, sharp, rb3, gray, (size_t)w, w, h));                                          // :940
    MSG_TRY(k_otsu(ctx, gray, (size_t)w, w, h, SMALL_HIST(ctx), SMALL_THRESH(ctx)));                  // :941
    MSG_TRY(k_threshold_u8(ctx, gray, (size_t)w, bw, (size_t)w, w, h, SMALL_THRESH(ctx), 0, 255));
    MSG_TRY(k_distance_transform(ctx, bw, (size_t)w, tmpA, w, h, SMALL_MAX(ctx)));                    // :1020
    MSG_TRY(k_normalize_minmax_f32(ctx, tmpA, dist, w, h, 0., 1., SMALL_MM(ctx)));                    // :1021
    MSG_TRY(k_threshold_f32(ctx, dist, tmpA, w, h, (float)peak_thresh, 1.f));                         // :348 (raw distances are dead)
    MSG_TRY(k_dilate_f32(ctx, tmpA, tmpB, w, h, 3, 3));                                               // :349-350
    MSG_TRY(k_f32_to_u8(ctx, tmpB, peaks, (size_t)w, w, h));                                          // :355-356
    MSG_TRY(k_contour_markers(ctx, peaks, (size_t)w, w, h, d_markers, mstep, n_contours));            // :360-365
    return circle_dev(ctx, d_markers, mstep, w, h, 5, 5, 3, 255);                                     // :366
}

int msg_color_seeds(msg_ctx* ctx, const uint8_t* src, size_t sstep, int w, int h, const int8_t* taps, int krows, int kcols,
                    double peak_thresh, int32_t* markers, size_t mstep, int32_t* n_contours, uint8_t* sharp, size_t sharp_step,
                    uint8_t* bw, size_t bw_step, float* dist, size_t dist_step, uint8_t* peaks, size_t peaks_step)
{
    CTX_ENTER(ctx);
    MSG_TRY(check_img(ctx, src, sstep, w, h, 3, "colour seeds src"));
    MSG_TRY(check_img(ctx, markers, mstep, w, h, 4, "markers"));
    if (sharp) MSG_TRY(check_img(ctx, sharp, sharp_step, w, h, 3, "colour seeds sharp"));
    if (bw) MSG_TRY(check_img(ctx, bw, bw_step, w, h, 1, "colour seeds bw"));
    if (dist) MSG_TRY(check_img(ctx, dist, dist_step, w, h, 4, "colour seeds dist"));
    if (peaks) MSG_TRY(check_img(ctx, peaks, peaks_step, w, h, 1, "colour seeds peaks"));
    const size_t n = (size_t)w * h;
    MSG_TRY(copy_in(ctx, src, sstep, (size_t)w * 3, h, &ctx->d_in, &ctx->d_in_cap));
    // d_out: [sharp 3n][bw n][peaks n][pad][dist 4n];  d_out2: markers
    MSG_TRY(msg_reserve(ctx, (void**)&ctx->d_out, &ctx->d_out_cap, 9 * n + 64));
    MSG_TRY(msg_reserve(ctx, (void**)&ctx->d_out2, &ctx->d_out2_cap, 4 * n));
    uint8_t* d_sharp = ctx->d_out;
    uint8_t* d_bw = ctx->d_out + 3 * n;
    uint8_t* d_peaks = ctx->d_out + 4 * n;
    float* d_dist = (float*)(ctx->d_out + (((5 * n) + 15) & ~(size_t)15));
    int32_t cnt = 0;
    MSG_TRY(msg_color_seeds_dev(ctx, ctx->d_in, (size_t)w * 3, w, h, taps, krows, kcols, peak_thresh, (int32_t*)ctx->d_out2,
                                (size_t)w * 4, &cnt, d_sharp, d_bw, d_dist, d_peaks));
    MSG_TRY(copy_out(ctx, markers, mstep, ctx->d_out2, (size_t)w * 4, h));
    if (sharp) MSG_TRY(copy_out(ctx, sharp, sharp_step, d_sharp, (size_t)w * 3, h));
    if (bw) MSG_TRY(copy_out(ctx, bw, bw_step, d_bw, (size_t)w, h));
    if (dist) MSG_TRY(copy_out(ctx, dist, dist_step, d_dist, (size_t)w * 4, h));
    if (peaks) MSG_TRY(copy_out(ctx, peaks, peaks_step, d_peaks, (size_t)w, h));
    MSG_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    if (n_contours) *n_contours = cnt;
    return MSG_OK;
}

// ============================================================================ bilateral filter (8(f2), row a5)

int msg_bilateral_filter(msg_ctx* ctx, const uint8_t* src, size_t sstep, uint8_t* dst, size_t dstep, int w, int h, int cn, int d,
                         double sigma_color, double sigma_space)
{
    CTX_ENTER(ctx);
    if (cn != 1 && cn != 3) return msg_fail(ctx, MSG_EINVAL, "bilateralFilter: 1 or 3 channels (got %d)", cn);
    MSG_TRY(check_img(ctx, src, sstep, w, h, cn, "bilateralFilter src"));
    MSG_TRY(check_img(ctx, dst, dstep, w, h, cn, "bilateralFilter dst"));
    if (!(sigma_color == sigma_color) || !(sigma_space == sigma_space))
        return msg_fail(ctx, MSG_EINVAL, "bilateralFilter: sigmas must be numbers");
    if (sigma_color <= 0) sigma_color = 1;
    if (sigma_space <= 0) sigma_space = 1;
    const double gc = -0.5 / (sigma_color * sigma_color), gs = -0.5 / (sigma_space * sigma_space);
    int radius = d <= 0 ? (int)lrint(sigma_space * 1.5) : d / 2;
    if (radius < 1) radius = 1;
    if (radius > 32) return msg_fail(ctx, MSG_EINVAL, "bilateralFilter: window radius %d > 32", radius);
    const int dd = 2 * radius + 1;
    // tables exactly as cv::bilateralFilter builds them (double exp, rounded to float)
    size_t tbytes = (size_t)dd * dd * (sizeof(float) + 2 * sizeof(short)) + (size_t)256 * cn * sizeof(float);
    uint8_t* host = (uint8_t*)malloc(tbytes);
    if (!host) return msg_fail(ctx, MSG_ENOMEM, "bilateralFilter: out of host memory");
    float* sw = (float*)host;
    float* cw = sw + dd * dd;
    short* ofs = (short*)(cw + 256 * cn);
    for (int i = 0; i < 256 * cn; i++) cw[i] = (float)exp(i * i * gc);
    int maxk = 0;
    for (int i = -radius; i <= radius; i++)
        for (int j = -radius; j <= radius; j++) {
            double r = sqrt((double)i * i + (double)j * j);
            if (r > radius) continue;
            sw[maxk] = (float)exp(r * r * gs);
            ofs[2 * maxk] = (short)j; ofs[2 * maxk + 1] = (short)i;
            maxk++;
        }
    int rc = small_reserve(ctx, tbytes);
    if (rc == MSG_OK && cudaMemcpyAsync(SMALL_TABLES(ctx), host, tbytes, cudaMemcpyHostToDevice, ctx->stream) != cudaSuccess)
        rc = msg_fail(ctx, MSG_ECUDA, "bilateralFilter: table upload failed");
    if (rc == MSG_OK && cudaStreamSynchronize(ctx->stream) != cudaSuccess) rc = msg_fail(ctx, MSG_ECUDA, "bilateralFilter: sync failed");
    free(host);
    MSG_TRY(rc);
    const float* d_sw = (const float*)SMALL_TABLES(ctx);
    const float* d_cw = d_sw + dd * dd;
    const short* d_ofs = (const short*)(d_cw + 256 * cn);
    size_t rb = (size_t)w * cn;
    MSG_TRY(copy_in(ctx, src, sstep, rb, h, &ctx->d_in, &ctx->d_in_cap));
    MSG_TRY(msg_reserve(ctx, (void**)&ctx->d_out, &ctx->d_out_cap, rb * h));
    MSG_TRY(k_bilateral(ctx, ctx->d_in, rb, ctx->d_out, rb, w, h, cn, radius, maxk, d_sw, d_ofs, d_cw));
    MSG_TRY(copy_out(ctx, dst, dstep, ctx->d_out, rb, h));
    MSG_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return MSG_OK;
}

}  // extern "C"
