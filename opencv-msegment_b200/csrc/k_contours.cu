// k_contours.cu -- contour labelling of the colour-method marker generator:
//     findContours(mask, RETR_CCOMP, CHAIN_APPROX_NONE) and then, for i = 0 .. n-1,
//     drawContours(markers, contours, i, Scalar.all(i + 1), FILLED, 8, hierarchy, INT_MAX)      PictureService.java:360-364
// restated without contour tracing (oracle: orc_contour_markers, pinned on cv2 4.13):
//   * outer contours  = 8-connected components of the non-zero pixels; holes = 4-connected components of the zero pixels
//     that do not touch the image border; a hole belongs to the component of the pixel left of its first (raster) pixel;
//   * contour index   : components in REVERSE raster order of their first pixel, each followed by its holes in reverse
//     raster order of their first pixel;
//   * painting an outer contour covers the component; painting a hole covers the hole, everything enclosed by it and the
//     component's pixels 4-adjacent to the hole; larger indices overwrite smaller ones, so a pixel inside any hole takes the
//     index of the OUTERMOST hole around it.
// Both labellings are the union-find of k_ccl.cu (roots = first pixels).  The per-contour bookkeeping (ordered compaction of
// the roots, a scan over the components, a stable sort of the holes by parent) uses CUB's device primitives.
#pragma GCC diagnostic ignored "-Wdeprecated-declarations"
#include <cub/device/device_radix_sort.cuh>
#include <cub/device/device_scan.cuh>
#include <cub/device/device_select.cuh>
#include <cub/iterator/counting_input_iterator.cuh>

#include "msg_internal.h"

namespace {

constexpr int CT_THREADS = 256;
inline unsigned blocks_for(size_t n, int per) { return (unsigned)((n + per - 1) / per); }

__global__ void __launch_bounds__(CT_THREADS) invert_mask_kernel(const uint8_t* __restrict__ src, size_t sstep,
                                                                 uint8_t* __restrict__ inv, int w)
{
    int x = blockIdx.x * CT_THREADS + threadIdx.x, y = blockIdx.y;
    if (x < w) inv[(size_t)y * w + x] = src[(size_t)y * sstep + x] ? 0 : 1;
}

// open[root] = 1 for zero-pixel components that touch the image border
__global__ void __launch_bounds__(CT_THREADS) mark_open_kernel(const int32_t* __restrict__ BG, int w, int h,
                                                               uint8_t* __restrict__ open)
{
    int i = blockIdx.x * CT_THREADS + threadIdx.x;
    int per = 2 * w + 2 * h;
    if (i >= per) return;
    int x, y;
    if (i < w) { x = i; y = 0; }
    else if (i < 2 * w) { x = i - w; y = h - 1; }
    else if (i < 2 * w + h) { x = 0; y = i - 2 * w; }
    else { x = w - 1; y = i - 2 * w - h; }
    int r = BG[(size_t)y * w + x];
    if (r >= 0) open[r] = 1;
}

struct is_fg_root {
    const int32_t* FG;
    __device__ bool operator()(int i) const { return FG[i] == i; }
};
struct is_hole_root {
    const int32_t* BG;
    const uint8_t* open;
    __device__ bool operator()(int i) const { return BG[i] == i && !open[i]; }
};

// holes: parent component (root pixel) as the sort key, hole count per component
__global__ void __launch_bounds__(CT_THREADS) hole_parent_kernel(const int32_t* __restrict__ holes, int nh,
                                                                 const int32_t* __restrict__ FG, int32_t* __restrict__ keys,
                                                                 int32_t* __restrict__ nholes)
{
    int i = blockIdx.x * CT_THREADS + threadIdx.x;
    if (i >= nh) return;
    int c = FG[holes[i] - 1];                    // the pixel left of a hole's first pixel lies in the surrounding component
    keys[i] = c;
    atomicAdd(&nholes[c], 1);
}

__global__ void __launch_bounds__(CT_THREADS) comp_weight_kernel(const int32_t* __restrict__ comps, int nc,
                                                                 const int32_t* __restrict__ nholes, int32_t* __restrict__ wgt)
{
    int i = blockIdx.x * CT_THREADS + threadIdx.x;
    if (i < nc) wgt[i] = 1 + nholes[comps[i]];
}

// idx[root of component i] = total - inclusive(i);  hbefore[root] = holes of the components discovered before it
__global__ void __launch_bounds__(CT_THREADS) comp_index_kernel(const int32_t* __restrict__ comps, int nc,
                                                                const int32_t* __restrict__ wgt, const int32_t* __restrict__ excl,
                                                                int total, int32_t* __restrict__ idx, int32_t* __restrict__ hbefore)
{
    int i = blockIdx.x * CT_THREADS + threadIdx.x;
    if (i >= nc) return;
    int r = comps[i];
    idx[r] = total - (excl[i] + wgt[i]);
    hbefore[r] = excl[i] - i;
}

// holes sorted by (parent, first pixel): idx[hole root] = idx[parent] + nholes[parent] - ordinal among the siblings
__global__ void __launch_bounds__(CT_THREADS) hole_index_kernel(const int32_t* __restrict__ skeys, const int32_t* __restrict__ sholes,
                                                                int nh, const int32_t* __restrict__ nholes,
                                                                const int32_t* __restrict__ hbefore, int32_t* __restrict__ idx)
{
    int i = blockIdx.x * CT_THREADS + threadIdx.x;
    if (i >= nh) return;
    int c = skeys[i];
    idx[sholes[i]] = idx[c] + nholes[c] - (i - hbefore[c]);
}

// top[hole root] = 1 + index of the outermost hole around it
__global__ void __launch_bounds__(CT_THREADS) hole_top_kernel(const int32_t* __restrict__ holes, int nh,
                                                              const int32_t* __restrict__ FG, const int32_t* __restrict__ BG,
                                                              const uint8_t* __restrict__ open, int w,
                                                              const int32_t* __restrict__ idx, int32_t* __restrict__ top)
{
    int i = blockIdx.x * CT_THREADS + threadIdx.x;
    if (i >= nh) return;
    int hole = holes[i];
    for (;;) {
        int r = FG[hole - 1];                    // surrounding component (root = its first pixel)
        if (r % w == 0) break;
        int b = BG[r - 1];                       // the background region left of the component's first pixel
        if (open[b]) break;
        hole = b;
    }
    top[holes[i]] = idx[hole] + 1;
}

__global__ void __launch_bounds__(CT_THREADS) paint_kernel(const int32_t* __restrict__ FG, const int32_t* __restrict__ BG,
                                                           const uint8_t* __restrict__ open, const int32_t* __restrict__ idx,
                                                           const int32_t* __restrict__ top, int w, int h,
                                                           int32_t* __restrict__ out, size_t ostep)
{
    int x = blockIdx.x * CT_THREADS + threadIdx.x, y = blockIdx.y;
    if (x >= w) return;
    size_t p = (size_t)y * w + x;
    int r = FG[p], v = 0;
    if (r >= 0) {
        int b = (r % w) ? BG[r - 1] : -1;
        if (b >= 0 && !open[b]) v = top[b];                    // the component is an island inside a hole
        else {
            int best = idx[r];
            if (x > 0) { int q = BG[p - 1]; if (q >= 0 && !open[q]) best = max(best, idx[q]); }
            if (x + 1 < w) { int q = BG[p + 1]; if (q >= 0 && !open[q]) best = max(best, idx[q]); }
            if (y > 0) { int q = BG[p - w]; if (q >= 0 && !open[q]) best = max(best, idx[q]); }
            if (y + 1 < h) { int q = BG[p + w]; if (q >= 0 && !open[q]) best = max(best, idx[q]); }
            v = best + 1;
        }
    } else {
        int b = BG[p];
        if (!open[b]) v = top[b];
    }
    *(int32_t*)((uint8_t*)out + (size_t)y * ostep + (size_t)x * 4) = v;
}

}  // namespace

// Scratch (ctx->d_scratch, grown here): [FG n][BG n][idx n][aux n (nholes at component roots, top at hole roots)]
// [hbefore n][lists: comps n/2+1, holes n/2+1, keys, sorted keys, sorted holes, wgt, excl][open n bytes][inv n bytes][cub temp]
int k_contour_markers(msg_ctx* ctx, const uint8_t* d_mask, size_t step, int w, int h, int32_t* d_markers, size_t mstep,
                      int32_t* n_contours_host)
{
    const size_t n = (size_t)w * h;
    const size_t half = n / 2 + 2;                 // roots of either kind are never adjacent in a row: at most ~n/2
    cudaStream_t st = ctx->stream;
    // CUB temporary storage: the maximum over the calls below
    size_t t_sel = 0, t_scan = 0, t_sort = 0;
    {
        cub::CountingInputIterator<int> it(0);
        is_fg_root pr{nullptr};
        cub::DeviceSelect::If(nullptr, t_sel, it, (int32_t*)nullptr, (int32_t*)nullptr, (int)n, pr, st);
        cub::DeviceScan::ExclusiveSum(nullptr, t_scan, (int32_t*)nullptr, (int32_t*)nullptr, (int)half, st);
        cub::DeviceRadixSort::SortPairs(nullptr, t_sort, (int32_t*)nullptr, (int32_t*)nullptr, (int32_t*)nullptr,
                                        (int32_t*)nullptr, (int)half, 0, 32, st);
    }
    size_t t_cub = t_sel > t_scan ? t_sel : t_scan;
    if (t_sort > t_cub) t_cub = t_sort;
    size_t need = (5 * n + 7 * half + 16) * sizeof(int32_t) + 2 * n + 64 + t_cub + 256;
    MSG_TRY(msg_reserve(ctx, (void**)&ctx->d_scratch, &ctx->d_scratch_cap, need));
    int32_t* FG = (int32_t*)ctx->d_scratch;
    int32_t* BG = FG + n;
    int32_t* idx = BG + n;
    int32_t* aux = idx + n;
    int32_t* hbefore = aux + n;
    int32_t* comps = hbefore + n;
    int32_t* holes = comps + half;
    int32_t* keys = holes + half;
    int32_t* skeys = keys + half;
    int32_t* sholes = skeys + half;
    int32_t* wgt = sholes + half;
    int32_t* excl = wgt + half;
    int32_t* d_cnt = excl + half;                  // [0] components, [1] holes
    uint8_t* open = (uint8_t*)(d_cnt + 16);
    uint8_t* inv = open + n;
    void* d_tmp = (void*)(((uintptr_t)(inv + n) + 255) & ~(uintptr_t)255);

    dim3 grid2((w + CT_THREADS - 1) / CT_THREADS, h);
    // the two labellings (k_ccl_binary leaves union-find parents: flatten to roots)
    MSG_TRY(k_ccl_binary(ctx, d_mask, step, w, h, 8, FG));
    MSG_TRY(k_ccl_flatten(ctx, FG, n));
    invert_mask_kernel<<<grid2, CT_THREADS, 0, st>>>(d_mask, step, inv, w);
    MSG_LAUNCHED(ctx);
    MSG_TRY(k_ccl_binary(ctx, inv, (size_t)w, w, h, 4, BG));
    MSG_TRY(k_ccl_flatten(ctx, BG, n));
    MSG_CUDA(ctx, cudaMemsetAsync(open, 0, n, st));
    MSG_CUDA(ctx, cudaMemsetAsync(aux, 0, n * sizeof(int32_t), st));
    mark_open_kernel<<<blocks_for((size_t)2 * w + 2 * h, CT_THREADS), CT_THREADS, 0, st>>>(BG, w, h, open);
    MSG_LAUNCHED(ctx);
    // ordered lists of the component roots and of the hole roots
    cub::CountingInputIterator<int> it(0);
    size_t tb = t_cub;
    MSG_CUDA(ctx, cub::DeviceSelect::If(d_tmp, tb, it, comps, d_cnt, (int)n, is_fg_root{FG}, st));
    MSG_LAUNCHED(ctx);
    tb = t_cub;
    MSG_CUDA(ctx, cub::DeviceSelect::If(d_tmp, tb, it, holes, d_cnt + 1, (int)n, is_hole_root{BG, open}, st));
    MSG_LAUNCHED(ctx);
    int cnt[2] = {0, 0};
    MSG_CUDA(ctx, cudaMemcpyAsync(cnt, d_cnt, sizeof(cnt), cudaMemcpyDeviceToHost, st));
    MSG_CUDA(ctx, cudaStreamSynchronize(st));
    const int nc = cnt[0], nh = cnt[1];
    if (nc < 0 || nh < 0 || (size_t)nc > half || (size_t)nh > half)
        return msg_fail(ctx, MSG_ECUDA, "contour labelling: inconsistent root counts %d / %d", nc, nh);
    if (nh) {
        hole_parent_kernel<<<blocks_for(nh, CT_THREADS), CT_THREADS, 0, st>>>(holes, nh, FG, keys, aux);
        MSG_LAUNCHED(ctx);
    }
    if (nc) {
        comp_weight_kernel<<<blocks_for(nc, CT_THREADS), CT_THREADS, 0, st>>>(comps, nc, aux, wgt);
        MSG_LAUNCHED(ctx);
        tb = t_cub;
        MSG_CUDA(ctx, cub::DeviceScan::ExclusiveSum(d_tmp, tb, wgt, excl, nc, st));
        MSG_LAUNCHED(ctx);
        comp_index_kernel<<<blocks_for(nc, CT_THREADS), CT_THREADS, 0, st>>>(comps, nc, wgt, excl, nc + nh, idx, hbefore);
        MSG_LAUNCHED(ctx);
    }
    if (nh) {
        tb = t_cub;
        MSG_CUDA(ctx, cub::DeviceRadixSort::SortPairs(d_tmp, tb, keys, skeys, holes, sholes, nh, 0, 32, st));
        MSG_LAUNCHED(ctx);
        hole_index_kernel<<<blocks_for(nh, CT_THREADS), CT_THREADS, 0, st>>>(skeys, sholes, nh, aux, hbefore, idx);
        MSG_LAUNCHED(ctx);
        // aux at hole roots (never a component root) now receives the outermost-hole label
        hole_top_kernel<<<blocks_for(nh, CT_THREADS), CT_THREADS, 0, st>>>(holes, nh, FG, BG, open, w, idx, aux);
        MSG_LAUNCHED(ctx);
    }
    paint_kernel<<<grid2, CT_THREADS, 0, st>>>(FG, BG, open, idx, aux, w, h, d_markers, mstep);
    MSG_LAUNCHED(ctx);
    MSG_CHECK_LAUNCH(ctx);
    *n_contours_host = nc + nh;
    return MSG_OK;
}
