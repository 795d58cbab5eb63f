// k_filters.cu -- the pre-filters the reference really calls around its segmentation stage ("next" row):
//   Laplacian sharpen chain  filter2D(CV_32F, K) + convertTo + subtract + convertTo(8U)   PictureService.java:323-333
//   cvtColor(BGR2GRAY)                                                                     PictureService.java:405, :940
//   medianBlur(8UC1, k)                                                                    PictureService.java:408, :436
// Exact integer forms; one pass each, HBM-bound (median: shared-memory tile + bisection on the value).
#include "msg_internal.h"

namespace {

__device__ __forceinline__ int reflect101(int p, int n)
{
    if (n == 1) return 0;
    while (p < 0 || p >= n) p = p < 0 ? -p : 2 * (n - 1) - p;
    return p;
}

// sharpen kernel taps (krows * kcols <= 1024), passed BY VALUE as a kernel argument: several contexts per GPU may run
// different kernels concurrently, so a shared __constant__ symbol would race between them
struct sharpen_taps { int8_t t[1024]; };

// dst = saturate_u8(src - sum_K taps * src), correlation, anchor at the kernel centre, BORDER_REFLECT_101
__global__ void __launch_bounds__(256) sharpen_kernel(const uint8_t* __restrict__ src, size_t sstep, uint8_t* __restrict__ dst,
                                                      size_t dstep, int w, int h, int krows, int kcols,
                                                      const __grid_constant__ sharpen_taps taps)
{
    int x = blockIdx.x * blockDim.x + threadIdx.x;
    int y = blockIdx.y;
    if (x >= w) return;
    const int ay = krows / 2, ax = kcols / 2;
    int a0 = 0, a1 = 0, a2 = 0;
    for (int a = 0; a < krows; a++) {
        const uint8_t* row = src + (size_t)reflect101(y + a - ay, h) * sstep;
        for (int b = 0; b < kcols; b++) {
            const uint8_t* p = row + 3 * (size_t)reflect101(x + b - ax, w);
            int t = taps.t[a * kcols + b];
            a0 += t * p[0]; a1 += t * p[1]; a2 += t * p[2];
        }
    }
    const uint8_t* c = src + (size_t)y * sstep + 3 * (size_t)x;
    uint8_t* o = dst + (size_t)y * dstep + 3 * (size_t)x;
    o[0] = (uint8_t)min(max((int)c[0] - a0, 0), 255);
    o[1] = (uint8_t)min(max((int)c[1] - a1, 0), 255);
    o[2] = (uint8_t)min(max((int)c[2] - a2, 0), 255);
}

__global__ void __launch_bounds__(256) gray_kernel(const uint8_t* __restrict__ src, size_t sstep, uint8_t* __restrict__ dst,
                                                   size_t dstep, int w, int compat342)
{
    int x = blockIdx.x * blockDim.x + threadIdx.x;
    int y = blockIdx.y;
    if (x >= w) return;
    const uint8_t* p = src + (size_t)y * sstep + 3 * (size_t)x;
    // OpenCV 4.x: 15-bit coefficients; OpenCV 3.4.2 (the version the reference binds, pom.xml:39-43): 14-bit ones
    int g = compat342 ? (1868 * p[0] + 9617 * p[1] + 4899 * p[2] + 8192) >> 14
                      : (3735 * p[0] + 19235 * p[1] + 9798 * p[2] + 16384) >> 15;
    dst[(size_t)y * dstep + x] = (uint8_t)g;
}

// median of the k x k window (BORDER_REPLICATE): tile 32x8 + halo in shared memory, then 8 bisection steps on the value
// (count of window samples <= mid), i.e. the smallest v with #(samples <= v) >= k*k/2 + 1.
constexpr int MED_TW = 32, MED_TH = 8;
__global__ void __launch_bounds__(MED_TW * MED_TH) median_kernel(const uint8_t* __restrict__ src, size_t sstep,
                                                                uint8_t* __restrict__ dst, size_t dstep, int w, int h, int k)
{
    extern __shared__ uint8_t tile[];
    const int r = k / 2;
    const int sw = MED_TW + 2 * r, sh = MED_TH + 2 * r;
    const int x0 = blockIdx.x * MED_TW - r, y0 = blockIdx.y * MED_TH - r;
    for (int i = threadIdx.x; i < sw * sh; i += MED_TW * MED_TH) {
        int yy = y0 + i / sw, xx = x0 + i % sw;
        yy = min(max(yy, 0), h - 1); xx = min(max(xx, 0), w - 1);
        tile[i] = src[(size_t)yy * sstep + xx];
    }
    __syncthreads();
    const int tx = threadIdx.x % MED_TW, ty = threadIdx.x / MED_TW;
    const int x = blockIdx.x * MED_TW + tx, y = blockIdx.y * MED_TH + ty;
    if (x >= w || y >= h) return;
    const int need = (k * k) / 2 + 1;
    int lo = 0, hi = 255;
    while (lo < hi) {
        int mid = (lo + hi) >> 1, cnt = 0;
        for (int a = 0; a < k; a++) {
            const uint8_t* row = tile + (ty + a) * sw + tx;
            for (int b = 0; b < k; b++) cnt += row[b] <= mid;
        }
        if (cnt >= need) hi = mid; else lo = mid + 1;
    }
    dst[(size_t)y * dstep + x] = (uint8_t)lo;
}

}  // namespace

int k_sharpen(msg_ctx* ctx, const uint8_t* d_src, size_t sstep, uint8_t* d_dst, size_t dstep, int w, int h, const int8_t* taps,
              int krows, int kcols)
{
    sharpen_taps t;
    memset(&t, 0, sizeof(t));
    memcpy(t.t, taps, (size_t)krows * kcols);
    dim3 grid((w + 255) / 256, h);
    sharpen_kernel<<<grid, 256, 0, ctx->stream>>>(d_src, sstep, d_dst, dstep, w, h, krows, kcols, t);
    MSG_LAUNCHED(ctx);
    MSG_CHECK_LAUNCH(ctx);
    return MSG_OK;
}

int k_gray(msg_ctx* ctx, const uint8_t* d_src, size_t sstep, uint8_t* d_dst, size_t dstep, int w, int h)
{
    dim3 grid((w + 255) / 256, h);
    gray_kernel<<<grid, 256, 0, ctx->stream>>>(d_src, sstep, d_dst, dstep, w, ctx->tune.gray_compat);
    MSG_LAUNCHED(ctx);
    MSG_CHECK_LAUNCH(ctx);
    return MSG_OK;
}

int k_median(msg_ctx* ctx, const uint8_t* d_src, size_t sstep, uint8_t* d_dst, size_t dstep, int w, int h, int k)
{
    int r = k / 2;
    size_t smem = (size_t)(MED_TW + 2 * r) * (MED_TH + 2 * r);
    MSG_TRY(msg_func_smem(ctx, (const void*)median_kernel, smem));
    dim3 grid((w + MED_TW - 1) / MED_TW, (h + MED_TH - 1) / MED_TH);
    median_kernel<<<grid, MED_TW * MED_TH, smem, ctx->stream>>>(d_src, sstep, d_dst, dstep, w, h, k);
    MSG_LAUNCHED(ctx);
    MSG_CHECK_LAUNCH(ctx);
    return MSG_OK;
}
