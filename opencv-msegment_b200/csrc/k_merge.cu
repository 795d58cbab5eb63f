// k_merge.cu -- K2b region statistics + merge rounds, K2c label rendering.
//
// K2b has no OpenCV counterpart; its specification is DESIGN.md "K2b" (restated by the oracle's
// orc_merge_regions): per round every participating region selects the 4-adjacent region with the
// smallest (dist2 of rounded mean colours, label) key, accepted selections are united
// simultaneously (union by smallest label); phase A = colour fuse, phase B = min-size prune.
// K2c = PictureService.colorByIndexes (PictureService.java:913-936).
#include "msg_internal.h"

namespace {

constexpr int MT = 256;
constexpr int MERGE_MAX_ROUNDS = 64;

struct merge_tables {
    unsigned int* area;            // [nl]
    unsigned long long* sum;       // [3*nl]  B,G,R
    uint32_t* mean;                // [nl]   packed B | G<<8 | R<<16
    unsigned long long* best;      // [nl]   (dist2 << 32) | neighbour label
    int32_t* par;                  // [nl]
};

__global__ void __launch_bounds__(MT) max_label_kernel(const int32_t* __restrict__ L, size_t n, int32_t* __restrict__ out)
{
    int m = 0;
    for (size_t i = (size_t)blockIdx.x * MT + threadIdx.x; i < n; i += (size_t)gridDim.x * MT) m = max(m, L[i]);
#pragma unroll
    for (int o = 16; o; o >>= 1) m = max(m, __shfl_xor_sync(0xffffffffu, m, o));
    if ((threadIdx.x & 31) == 0 && m > 0) atomicMax(out, m);
}

__global__ void __launch_bounds__(MT) init_tables_kernel(merge_tables t, int nl)
{
    int i = blockIdx.x * MT + threadIdx.x;
    if (i >= nl) return;
    t.area[i] = 0;
    t.sum[3 * (size_t)i] = 0; t.sum[3 * (size_t)i + 1] = 0; t.sum[3 * (size_t)i + 2] = 0;
    t.best[i] = ~0ull;
    t.par[i] = i;
}

// per-pixel statistics with warp-level run aggregation (pixels of one warp lie in one row)
__global__ void __launch_bounds__(MT) stats_kernel(const uint32_t* __restrict__ plane, int pitch,
                                                   const int32_t* __restrict__ L, int w, int h, merge_tables t)
{
    int x = blockIdx.x * MT + threadIdx.x;
    int y = blockIdx.y;
    int lane = threadIdx.x & 31;
    int lab = 0;
    unsigned cnt = 0, b = 0, g = 0, r = 0;
    if (x < w) {
        lab = L[(size_t)y * w + x];
        if (lab > 0) {
            uint32_t c = __ldg(plane + (size_t)y * pitch + x);
            cnt = 1; b = c & 0xFF; g = (c >> 8) & 0xFF; r = (c >> 16) & 0xFF;
        }
    }
    int prev = __shfl_up_sync(0xffffffffu, lab, 1);
    bool head = lane == 0 || prev != lab;
    unsigned heads = __ballot_sync(0xffffffffu, head);
    unsigned above = lane == 31 ? 0u : (heads >> (lane + 1));
    int seg_end = above ? lane + __ffs(above) - 1 : 31;   // last lane of my run
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        unsigned c2 = __shfl_down_sync(0xffffffffu, cnt, o);
        unsigned b2 = __shfl_down_sync(0xffffffffu, b, o);
        unsigned g2 = __shfl_down_sync(0xffffffffu, g, o);
        unsigned r2 = __shfl_down_sync(0xffffffffu, r, o);
        if (lane + o <= seg_end) { cnt += c2; b += b2; g += g2; r += r2; }
    }
    if (head && lab > 0) {
        atomicAdd(t.area + lab, cnt);
        atomicAdd(t.sum + 3 * (size_t)lab, (unsigned long long)b);
        atomicAdd(t.sum + 3 * (size_t)lab + 1, (unsigned long long)g);
        atomicAdd(t.sum + 3 * (size_t)lab + 2, (unsigned long long)r);
    }
}

__global__ void __launch_bounds__(MT) mean_kernel(merge_tables t, int nl)
{
    int i = blockIdx.x * MT + threadIdx.x;
    if (i >= nl || i == 0) return;
    unsigned long long a = t.area[i];
    if (!a) { t.mean[i] = 0; return; }
    uint32_t m = 0;
#pragma unroll
    for (int c = 0; c < 3; c++) m |= (uint32_t)((2ull * t.sum[3 * (size_t)i + c] + a) / (2ull * a)) << (8 * c);
    t.mean[i] = m;
}

__global__ void __launch_bounds__(MT) edges_kernel(const int32_t* __restrict__ L, int w, int h, merge_tables t,
                                                   long long size_thr)
{
    int x = blockIdx.x * MT + threadIdx.x;
    int y = blockIdx.y;
    if (x >= w) return;
    size_t p = (size_t)y * w + x;
    int lab = L[p];
    if (lab <= 0 || (long long)t.area[lab] >= size_thr) return;
    uint32_t ml = t.mean[lab];
    unsigned long long bestk = ~0ull;
    int q[4];
    q[0] = x > 0 ? L[p - 1] : 0;
    q[1] = x + 1 < w ? L[p + 1] : 0;
    q[2] = y > 0 ? L[p - w] : 0;
    q[3] = y + 1 < h ? L[p + w] : 0;
#pragma unroll
    for (int k = 0; k < 4; k++) {
        if (q[k] <= 0 || q[k] == lab) continue;
        uint32_t e = __vabsdiffu4(ml, t.mean[q[k]]);
        unsigned long long key = ((unsigned long long)__dp4a(e, e, 0u) << 32) | (unsigned)q[k];
        bestk = key < bestk ? key : bestk;
    }
    if (bestk != ~0ull && bestk < t.best[lab]) atomicMin(t.best + lab, bestk);
}

__device__ __forceinline__ int pf_find(const int32_t* P, int a)
{
    int p = __ldcg(P + a);
    while (p != a) { a = p; p = __ldcg(P + a); }
    return a;
}

__global__ void __launch_bounds__(MT) select_kernel(merge_tables t, int nl, long long dist_limit, int32_t* __restrict__ accepted)
{
    int i = blockIdx.x * MT + threadIdx.x;
    if (i >= nl || i == 0) return;
    unsigned long long k = t.best[i];
    if (k == ~0ull) return;
    if ((long long)(k >> 32) > dist_limit) return;
    int a = i, b = (int)(k & 0xffffffffu);
    for (;;) {
        a = pf_find(t.par, a);
        b = pf_find(t.par, b);
        if (a == b) break;
        if (a < b) { int s = a; a = b; b = s; }
        int old = atomicMin(t.par + a, b);
        if (old == a) break;
        a = old;
    }
    atomicAdd(accepted, 1);
}

__global__ void __launch_bounds__(MT) apply_par_kernel(int32_t* __restrict__ L, size_t n, const int32_t* __restrict__ par)
{
    size_t i = (size_t)blockIdx.x * MT + threadIdx.x;
    if (i >= n) return;
    int v = L[i];
    if (v > 0) {
        int r = pf_find(par, v);
        if (r != v) L[i] = r;
    }
}

__global__ void __launch_bounds__(MT) render_kernel(const int32_t* __restrict__ L, size_t lstep, uint8_t* __restrict__ dst,
                                                    size_t dstep, int w, int depth, const uint8_t* __restrict__ colors)
{
    int x = blockIdx.x * MT + threadIdx.x;
    int y = blockIdx.y;
    if (x >= w) return;
    int v = ((const int32_t*)((const char*)L + (size_t)y * lstep))[x];
    uint8_t c0 = 0, c1 = 0, c2 = 0;
    if (v > 0 && v <= depth) {
        if (colors) { c0 = colors[3 * (size_t)(v - 1)]; c1 = colors[3 * (size_t)(v - 1) + 1]; c2 = colors[3 * (size_t)(v - 1) + 2]; }
        else c0 = c1 = c2 = 255;
    }
    uint8_t* p = dst + (size_t)y * dstep + 3 * (size_t)x;
    p[0] = c0; p[1] = c1; p[2] = c2;
}

}  // namespace

int k_render(msg_ctx* ctx, const int32_t* d_labels, size_t lstep, uint8_t* d_dst, size_t dstep, int w, int h, int depth,
             const uint8_t* d_colors)
{
    dim3 grid((w + MT - 1) / MT, h);
    render_kernel<<<grid, MT, 0, ctx->stream>>>(d_labels, lstep, d_dst, dstep, w, depth, d_colors);
    MSG_LAUNCHED(ctx);
    MSG_CHECK_LAUNCH(ctx);
    return MSG_OK;
}

// d_counters: [8] max label, [9] accepted
int k_merge(msg_ctx* ctx, const uint32_t* d_plane, int pitch, int32_t* d_labels, int w, int h, int min_size,
            int color_dist, int32_t* d_n_out)
{
    cudaStream_t st = ctx->stream;
    size_t n = (size_t)w * h;
    int32_t* d_max = ctx->d_counters + 8;
    int32_t* d_acc = ctx->d_counters + 9;
    ctx->st.merge_rounds = 0;
    MSG_CUDA(ctx, cudaMemsetAsync(d_max, 0, sizeof(int32_t), st));
    max_label_kernel<<<ctx->sm_count * 8, MT, 0, st>>>(d_labels, n, d_max);
    MSG_LAUNCHED(ctx);
    MSG_CUDA(ctx, cudaMemcpyAsync(ctx->h_counters + 8, d_max, sizeof(int32_t), cudaMemcpyDeviceToHost, st));
    MSG_CUDA(ctx, cudaStreamSynchronize(st));
    long long maxl = ctx->h_counters[8];
    if (maxl > (long long)n) return msg_fail(ctx, MSG_EINVAL, "merge: labels must be <= width*height (max label %lld)", maxl);
    if (maxl > 0 && (min_size > 0 || color_dist > 0)) {
        int nl = (int)maxl + 1;
        size_t bytes = (size_t)nl * (4 + 24 + 4 + 8 + 4) + 256;
        // tables live after the relabel scratch region? keep them in their own allocation: d_colors is small,
        // so use a dedicated grow-only buffer carved from d_ovf (unused outside mean shift).
        MSG_TRY(msg_reserve(ctx, (void**)&ctx->d_ovf, &ctx->d_ovf_cap, bytes));
        char* base = (char*)ctx->d_ovf;
        merge_tables t;
        t.sum = (unsigned long long*)base;                         base += (size_t)nl * 24;
        t.best = (unsigned long long*)base;                        base += (size_t)nl * 8;
        t.area = (unsigned int*)base;                              base += (size_t)nl * 4;
        t.mean = (uint32_t*)base;                                  base += (size_t)nl * 4;
        t.par = (int32_t*)base;
        dim3 pgrid((w + MT - 1) / MT, h);
        unsigned lgrid = (unsigned)((nl + MT - 1) / MT);
        const long long INF = 1ll << 40;
        for (int phase = 0; phase < 2; phase++) {
            long long size_thr, dist_limit;
            if (phase == 0) { if (color_dist <= 0) continue; size_thr = INF; dist_limit = (long long)color_dist * color_dist; }
            else { if (min_size <= 0) continue; size_thr = min_size; dist_limit = INF; }
            for (int round = 0; round < MERGE_MAX_ROUNDS; round++) {
                init_tables_kernel<<<lgrid, MT, 0, st>>>(t, nl);
                MSG_LAUNCHED(ctx);
                stats_kernel<<<pgrid, MT, 0, st>>>(d_plane, pitch, d_labels, w, h, t);
                MSG_LAUNCHED(ctx);
                mean_kernel<<<lgrid, MT, 0, st>>>(t, nl);
                MSG_LAUNCHED(ctx);
                edges_kernel<<<pgrid, MT, 0, st>>>(d_labels, w, h, t, size_thr);
                MSG_LAUNCHED(ctx);
                MSG_CUDA(ctx, cudaMemsetAsync(d_acc, 0, sizeof(int32_t), st));
                select_kernel<<<lgrid, MT, 0, st>>>(t, nl, dist_limit, d_acc);
                MSG_LAUNCHED(ctx);
                MSG_CUDA(ctx, cudaMemcpyAsync(ctx->h_counters + 9, d_acc, sizeof(int32_t), cudaMemcpyDeviceToHost, st));
                MSG_CUDA(ctx, cudaStreamSynchronize(st));
                ctx->st.merge_rounds++;
                if (ctx->h_counters[9] == 0) break;
                apply_par_kernel<<<(unsigned)((n + MT - 1) / MT), MT, 0, st>>>(d_labels, n, t.par);
                MSG_LAUNCHED(ctx);
            }
        }
        MSG_CHECK_LAUNCH(ctx);
    }
    return k_relabel_canonical(ctx, d_labels, w, h, 0, d_n_out, 0);
}
