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
// Both labellings are the union-find of k_ccl.cu (roots = first pixels).  The per-contour bookkeeping is three small
// primitives written here: an ordered compaction of the roots (ballot ranks inside 4096-pixel chunks + a scan of the chunk
// totals), an exclusive scan over the components and a bitonic sort of the holes by (parent, first pixel).
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

// ---------------------------------------------------------------- ordered compaction of root pixels
// KIND 0: component roots (FG[i] == i); KIND 1: hole roots (BG[i] == i and the region does not touch the border)
template <int KIND>
__device__ __forceinline__ bool is_root(const int32_t* __restrict__ L, const uint8_t* __restrict__ open, size_t i)
{
    return KIND == 0 ? L[i] == (int32_t)i : (L[i] == (int32_t)i && !open[i]);
}

constexpr int CT_CHUNK = 4096, CT_SWEEPS = CT_CHUNK / CT_THREADS, CT_WARPS = CT_THREADS / 32;

// rank of every root among the roots of its 4096-pixel chunk (raster order) + the chunk totals
template <int KIND>
__global__ void __launch_bounds__(CT_THREADS) rank_roots_kernel(const int32_t* __restrict__ L, const uint8_t* __restrict__ open,
                                                                size_t n, int32_t* __restrict__ block_sums,
                                                                int32_t* __restrict__ lrank)
{
    __shared__ unsigned s_ballot[CT_SWEEPS * CT_WARPS];
    __shared__ int s_prefix[CT_SWEEPS * CT_WARPS];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const size_t cbase = (size_t)blockIdx.x * CT_CHUNK;
#pragma unroll
    for (int k = 0; k < CT_SWEEPS; k++) {
        size_t i = cbase + (size_t)k * CT_THREADS + threadIdx.x;
        bool r = i < n && is_root<KIND>(L, open, i);
        unsigned bal = __ballot_sync(0xffffffffu, r);
        if (lane == 0) s_ballot[k * CT_WARPS + warp] = bal;
    }
    __syncthreads();
    if (warp == 0) {
        int carry = 0;
#pragma unroll
        for (int q = 0; q < CT_SWEEPS * CT_WARPS / 32; q++) {
            int c = __popc(s_ballot[q * 32 + lane]);
            int incl = c;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                int t = __shfl_up_sync(0xffffffffu, incl, o);
                if (lane >= o) incl += t;
            }
            s_prefix[q * 32 + lane] = carry + incl - c;
            carry += __shfl_sync(0xffffffffu, incl, 31);
        }
        if (lane == 0) block_sums[blockIdx.x] = carry;
    }
    __syncthreads();
#pragma unroll
    for (int k = 0; k < CT_SWEEPS; k++) {
        unsigned bal = s_ballot[k * CT_WARPS + warp];
        if ((bal >> lane) & 1u) {
            size_t i = cbase + (size_t)k * CT_THREADS + threadIdx.x;
            lrank[i] = s_prefix[k * CT_WARPS + warp] + __popc(bal & ((1u << lane) - 1));
        }
    }
}

template <int KIND>
__global__ void __launch_bounds__(CT_THREADS) scatter_roots_kernel(const int32_t* __restrict__ L, const uint8_t* __restrict__ open,
                                                                   size_t n, const int32_t* __restrict__ block_offs,
                                                                   const int32_t* __restrict__ lrank, int32_t* __restrict__ list)
{
    size_t i = (size_t)blockIdx.x * CT_THREADS + threadIdx.x;
    if (i < n && is_root<KIND>(L, open, i)) list[block_offs[i / CT_CHUNK] + lrank[i]] = (int32_t)i;
}

// exclusive scan of `count` ints by one CTA (in == out allowed); the total goes to *total
__global__ void __launch_bounds__(1024) exclusive_scan_kernel(const int32_t* in, int32_t* out, int count, int32_t* __restrict__ total)
{
    __shared__ int s_warp[32];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    int carry = 0;
    for (int base = 0; base < count; base += 1024) {
        int i = base + threadIdx.x;
        int v = i < count ? in[i] : 0;
        int incl = v;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            int t = __shfl_up_sync(0xffffffffu, incl, o);
            if (lane >= o) incl += t;
        }
        if (lane == 31) s_warp[warp] = incl;
        __syncthreads();
        if (warp == 0) {
            int w = s_warp[lane];
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                int t = __shfl_up_sync(0xffffffffu, w, o);
                if (lane >= o) w += t;
            }
            s_warp[lane] = w;
        }
        __syncthreads();
        int prefix = warp ? s_warp[warp - 1] : 0;
        if (i < count) out[i] = carry + prefix + incl - v;
        carry += s_warp[31];
        __syncthreads();
    }
    if (threadIdx.x == 0 && total) *total = carry;
}

// ---------------------------------------------------------------- bitonic sort of 64-bit keys (N a power of two >= 2048)
constexpr int BS_TILE = 2048, BS_THREADS = BS_TILE / 2;

__device__ __forceinline__ void bitonic_cx(unsigned long long& x, unsigned long long& y, bool asc)
{
    if ((x > y) == asc) { unsigned long long t = x; x = y; y = t; }
}

// all steps with partner distance < BS_TILE of the merge stages k_lo .. k_hi, inside shared memory
__global__ void __launch_bounds__(BS_THREADS) bitonic_tile_kernel(unsigned long long* __restrict__ a, int k_lo, int k_hi)
{
    __shared__ unsigned long long s[BS_TILE];
    const int t = threadIdx.x;
    const size_t base = (size_t)blockIdx.x * BS_TILE;
    s[t] = a[base + t];
    s[t + BS_THREADS] = a[base + t + BS_THREADS];
    for (long long k = k_lo; k <= k_hi; k <<= 1) {
        for (int j = (int)(k >> 1 < BS_THREADS ? k >> 1 : BS_THREADS); j > 0; j >>= 1) {
            __syncthreads();
            int i = ((t & ~(j - 1)) << 1) | (t & (j - 1)), l = i | j;
            bool asc = (((base + i) & (size_t)k) == 0);
            unsigned long long x = s[i], y = s[l];
            bitonic_cx(x, y, asc);
            s[i] = x; s[l] = y;
        }
    }
    __syncthreads();
    a[base + t] = s[t];
    a[base + t + BS_THREADS] = s[t + BS_THREADS];
}

// one compare-exchange step with partner distance j >= BS_TILE
__global__ void __launch_bounds__(256) bitonic_step_kernel(unsigned long long* __restrict__ a, size_t half_n, long long k, long long j)
{
    size_t t = (size_t)blockIdx.x * 256 + threadIdx.x;
    if (t >= half_n) return;
    size_t i = ((t & ~(size_t)(j - 1)) << 1) | (t & (size_t)(j - 1)), l = i | (size_t)j;
    unsigned long long x = a[i], y = a[l];
    bool asc = ((i & (size_t)k) == 0);
    if ((x > y) == asc) { a[i] = y; a[l] = x; }
}

// holes: sort key (parent component root << 32 | first pixel) -- padding keys sort last -- and the hole count per component
__global__ void __launch_bounds__(CT_THREADS) hole_parent_kernel(const int32_t* __restrict__ holes, int nh, int npad,
                                                                 const int32_t* __restrict__ FG, unsigned long long* __restrict__ keys,
                                                                 int32_t* __restrict__ nholes)
{
    int i = blockIdx.x * CT_THREADS + threadIdx.x;
    if (i >= npad) return;
    if (i >= nh) { keys[i] = ~0ull; return; }
    int hole = holes[i];
    int c = FG[hole - 1];                        // the pixel left of a hole's first pixel lies in the surrounding component
    keys[i] = ((unsigned long long)(unsigned)c << 32) | (unsigned)hole;
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
__global__ void __launch_bounds__(CT_THREADS) hole_index_kernel(const unsigned long long* __restrict__ skeys, int nh,
                                                                const int32_t* __restrict__ nholes,
                                                                const int32_t* __restrict__ hbefore, int32_t* __restrict__ idx)
{
    int i = blockIdx.x * CT_THREADS + threadIdx.x;
    if (i >= nh) return;
    unsigned long long key = skeys[i];
    int c = (int)(key >> 32), hole = (int)(key & 0xffffffffu);
    idx[hole] = idx[c] + nholes[c] - (i - hbefore[c]);
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

// Scratch (ctx->d_scratch, grown here): [FG n][BG n][idx n (also the compaction ranks)][aux n (nholes at component roots,
// top at hole roots)][hbefore n][comps, holes, wgt, excl: n/2+2 each][chunk sums][counts][sort keys (64 bit)][open n][inv n]
int k_contour_markers(msg_ctx* ctx, const uint8_t* d_mask, size_t step, int w, int h, int32_t* d_markers, size_t mstep,
                      int32_t* n_contours_host)
{
    const size_t n = (size_t)w * h;
    const size_t half = n / 2 + 2;                 // roots of either kind are never adjacent in a row: at most ~n/2
    const int nb = (int)blocks_for(n, CT_CHUNK);
    size_t npad_max = BS_TILE;
    while (npad_max < half) npad_max <<= 1;
    cudaStream_t st = ctx->stream;
    size_t need = (5 * n + 4 * half + (size_t)nb + 64) * sizeof(int32_t) + npad_max * sizeof(unsigned long long) + 2 * n + 512;
    MSG_TRY(msg_reserve(ctx, (void**)&ctx->d_scratch, &ctx->d_scratch_cap, need));
    int32_t* FG = (int32_t*)ctx->d_scratch;
    int32_t* BG = FG + n;
    int32_t* idx = BG + n;
    int32_t* aux = idx + n;
    int32_t* hbefore = aux + n;
    int32_t* comps = hbefore + n;
    int32_t* holes = comps + half;
    int32_t* wgt = holes + half;
    int32_t* excl = wgt + half;
    int32_t* bsum = excl + half;                   // nb chunk totals
    int32_t* d_cnt = bsum + nb + 8;                // [0] components, [1] holes
    unsigned long long* keys = (unsigned long long*)(((uintptr_t)(d_cnt + 16) + 15) & ~(uintptr_t)15);
    uint8_t* open = (uint8_t*)(keys + npad_max);
    uint8_t* inv = open + n;

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
    // ordered lists of the component roots and of the hole roots (idx serves as the rank scratch: it is written later)
    rank_roots_kernel<0><<<nb, CT_THREADS, 0, st>>>(FG, open, n, bsum, idx);
    exclusive_scan_kernel<<<1, 1024, 0, st>>>(bsum, bsum, nb, d_cnt);
    scatter_roots_kernel<0><<<blocks_for(n, CT_THREADS), CT_THREADS, 0, st>>>(FG, open, n, bsum, idx, comps);
    rank_roots_kernel<1><<<nb, CT_THREADS, 0, st>>>(BG, open, n, bsum, idx);
    exclusive_scan_kernel<<<1, 1024, 0, st>>>(bsum, bsum, nb, d_cnt + 1);
    scatter_roots_kernel<1><<<blocks_for(n, CT_THREADS), CT_THREADS, 0, st>>>(BG, open, n, bsum, idx, holes);
    ctx->st.kernel_launches += 6;
    int cnt[2] = {0, 0};
    MSG_CUDA(ctx, cudaMemcpyAsync(cnt, d_cnt, sizeof(cnt), cudaMemcpyDeviceToHost, st));
    MSG_CUDA(ctx, cudaStreamSynchronize(st));
    const int nc = cnt[0], nh = cnt[1];
    if (nc < 0 || nh < 0 || (size_t)nc > half || (size_t)nh > half)
        return msg_fail(ctx, MSG_ECUDA, "contour labelling: inconsistent root counts %d / %d", nc, nh);
    size_t npad = BS_TILE;
    while (npad < (size_t)nh) npad <<= 1;
    if (nh) {
        hole_parent_kernel<<<blocks_for(npad, CT_THREADS), CT_THREADS, 0, st>>>(holes, nh, (int)npad, FG, keys, aux);
        MSG_LAUNCHED(ctx);
    }
    if (nc) {
        comp_weight_kernel<<<blocks_for(nc, CT_THREADS), CT_THREADS, 0, st>>>(comps, nc, aux, wgt);
        MSG_LAUNCHED(ctx);
        exclusive_scan_kernel<<<1, 1024, 0, st>>>(wgt, excl, nc, nullptr);
        MSG_LAUNCHED(ctx);
        comp_index_kernel<<<blocks_for(nc, CT_THREADS), CT_THREADS, 0, st>>>(comps, nc, wgt, excl, nc + nh, idx, hbefore);
        MSG_LAUNCHED(ctx);
    }
    if (nh) {
        // holes by (parent, first pixel): bitonic sort, partner distances below 2048 inside shared memory
        const unsigned tiles = (unsigned)(npad / BS_TILE);
        bitonic_tile_kernel<<<tiles, BS_THREADS, 0, st>>>(keys, 2, BS_TILE);
        MSG_LAUNCHED(ctx);
        for (size_t k = 2 * (size_t)BS_TILE; k <= npad; k <<= 1) {
            for (size_t j = k >> 1; j >= (size_t)BS_TILE; j >>= 1) {
                bitonic_step_kernel<<<blocks_for(npad / 2, 256), 256, 0, st>>>(keys, npad / 2, (long long)k, (long long)j);
                MSG_LAUNCHED(ctx);
            }
            bitonic_tile_kernel<<<tiles, BS_THREADS, 0, st>>>(keys, (int)k, (int)k);
            MSG_LAUNCHED(ctx);
        }
        hole_index_kernel<<<blocks_for(nh, CT_THREADS), CT_THREADS, 0, st>>>(keys, nh, aux, hbefore, idx);
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
