// k_seeds.cu -- the shape-method marker generator of the reference, the caller side of the
// labelling stage:   Canny(5, 50)  PictureService.java:416     dilate 3x3, dilate 5x5  :428-429     subtract  :430
// (cvtColor, medianBlur and connectedComponents of the same chain live in k_filters.cu / k_ccl.cu).
// Exact integer restatements of the OpenCV functions (oracle: orc_canny / orc_dilate_rect / orc_subtract_u8, pinned on cv2):
//   Canny   : Sobel 3x3 with BORDER_REPLICATE, magnitude |dx|+|dy| (zero outside the image), sector test with the
//             fixed-point tan(22.5 deg) = 13573 / 2^15, local maxima > low are candidates, > high are strong;
//             hysteresis = 8-connected components of the candidates that hold a strong one (union-find in k_ccl.cu,
//             order-independent, so the stack order of the CPU implementation does not matter).
//   dilate  : kw x kh all-ones kernel, anchor at the centre, pixels outside the image ignored.
// All three are one pass over an 8-bit plane: HBM-bound, shared-memory tile with halo.
#include "msg_internal.h"

namespace {

constexpr int CN_TW = 32, CN_TH = 8;     // 256 threads; gray tile + 2 halo, magnitude tile + 1 halo

__global__ void __launch_bounds__(CN_TW * CN_TH) canny_nms_kernel(const uint8_t* __restrict__ src, size_t sstep,
                                                                  uint8_t* __restrict__ cls, size_t cstep, int w, int h,
                                                                  int low, int high)
{
    constexpr int GW = CN_TW + 4, GH = CN_TH + 4, MW = CN_TW + 2, MH = CN_TH + 2;
    __shared__ uint8_t s_g[GH][GW];
    __shared__ short s_m[MH][MW];
    const int bx = blockIdx.x * CN_TW, by = blockIdx.y * CN_TH;
    for (int i = threadIdx.x; i < GW * GH; i += CN_TW * CN_TH) {
        int yy = min(max(by - 2 + i / GW, 0), h - 1), xx = min(max(bx - 2 + i % GW, 0), w - 1);   // BORDER_REPLICATE
        s_g[i / GW][i % GW] = src[(size_t)yy * sstep + xx];
    }
    __syncthreads();
    for (int i = threadIdx.x; i < MW * MH; i += CN_TW * CN_TH) {
        int my = i / MW, mx = i % MW;                      // magnitude at image pixel (bx - 1 + mx, by - 1 + my)
        int x = bx - 1 + mx, y = by - 1 + my;
        int m = 0;
        if (x >= 0 && x < w && y >= 0 && y < h) {
            const uint8_t* r0 = &s_g[my][mx];              // gray (x-1, y-1) sits at s_g[my][mx]
            const uint8_t* r1 = &s_g[my + 1][mx];
            const uint8_t* r2 = &s_g[my + 2][mx];
            int gx = (r0[2] + 2 * r1[2] + r2[2]) - (r0[0] + 2 * r1[0] + r2[0]);
            int gy = (r2[0] + 2 * r2[1] + r2[2]) - (r0[0] + 2 * r0[1] + r0[2]);
            m = abs(gx) + abs(gy);
        }
        s_m[my][mx] = (short)m;
    }
    __syncthreads();
    const int tx = threadIdx.x % CN_TW, ty = threadIdx.x / CN_TW;
    const int x = bx + tx, y = by + ty;
    if (x >= w || y >= h) return;
    const int v = s_m[ty + 1][tx + 1];
    int out = 0;
    if (v > low) {
        const uint8_t* r0 = &s_g[ty + 1][tx + 1];
        const uint8_t* r1 = &s_g[ty + 2][tx + 1];
        const uint8_t* r2 = &s_g[ty + 3][tx + 1];
        int xs = (r0[2] + 2 * r1[2] + r2[2]) - (r0[0] + 2 * r1[0] + r2[0]);
        int ys = (r2[0] + 2 * r2[1] + r2[2]) - (r0[0] + 2 * r0[1] + r0[2]);
        int ax = abs(xs), ay = abs(ys) << 15;
        int tg22 = ax * 13573;
        bool keep;
        if (ay < tg22) keep = v > s_m[ty + 1][tx] && v >= s_m[ty + 1][tx + 2];
        else {
            int tg67 = tg22 + (ax << 16);
            if (ay > tg67) keep = v > s_m[ty][tx + 1] && v >= s_m[ty + 2][tx + 1];
            else {
                int s = (xs ^ ys) < 0 ? -1 : 1;
                keep = v > s_m[ty][tx + 1 - s] && v > s_m[ty + 2][tx + 1 + s];
            }
        }
        if (keep) out = v > high ? 2 : 1;
    }
    cls[(size_t)y * cstep + x] = (uint8_t)out;
}

constexpr int DL_TW = 32, DL_TH = 8;
__global__ void __launch_bounds__(DL_TW * DL_TH) dilate_kernel(const uint8_t* __restrict__ src, size_t sstep,
                                                               uint8_t* __restrict__ dst, size_t dstep, int w, int h, int kw,
                                                               int kh)
{
    extern __shared__ uint8_t s_t[];
    const int sw = DL_TW + kw - 1, sh = DL_TH + kh - 1;
    const int x0 = blockIdx.x * DL_TW - kw / 2, y0 = blockIdx.y * DL_TH - kh / 2;
    for (int i = threadIdx.x; i < sw * sh; i += DL_TW * DL_TH) {
        int yy = y0 + i / sw, xx = x0 + i % sw;
        s_t[i] = (yy >= 0 && yy < h && xx >= 0 && xx < w) ? src[(size_t)yy * sstep + xx] : 0;   // outside: neutral for max
    }
    __syncthreads();
    const int tx = threadIdx.x % DL_TW, ty = threadIdx.x / DL_TW;
    const int x = blockIdx.x * DL_TW + tx, y = blockIdx.y * DL_TH + ty;
    if (x >= w || y >= h) return;
    int best = 0;
    for (int a = 0; a < kh; a++) {
        const uint8_t* row = s_t + (ty + a) * sw + tx;
        for (int b = 0; b < kw; b++) best = max(best, (int)row[b]);
    }
    dst[(size_t)y * dstep + x] = (uint8_t)best;
}

__global__ void __launch_bounds__(256) subtract_kernel(const uint8_t* __restrict__ a, size_t astep, const uint8_t* __restrict__ b,
                                                       size_t bstep, uint8_t* __restrict__ dst, size_t dstep, int w)
{
    int x = blockIdx.x * blockDim.x + threadIdx.x;
    int y = blockIdx.y;
    if (x >= w) return;
    int v = (int)a[(size_t)y * astep + x] - (int)b[(size_t)y * bstep + x];
    dst[(size_t)y * dstep + x] = (uint8_t)max(v, 0);
}

// Mat.copyTo(dst, mask) onto a zero Mat (PictureService.java:417-418, the "borders" step): dst = mask ? src : 0
__global__ void __launch_bounds__(256) copy_masked_kernel(const uint8_t* __restrict__ src, size_t sstep, const uint8_t* __restrict__ mask,
                                                          size_t mstep, uint8_t* __restrict__ dst, size_t dstep, int w)
{
    int x = blockIdx.x * blockDim.x + threadIdx.x;
    int y = blockIdx.y;
    if (x >= w) return;
    const bool on = mask[(size_t)y * mstep + x] != 0;
    const uint8_t* s = src + (size_t)y * sstep + 3 * (size_t)x;
    uint8_t* d = dst + (size_t)y * dstep + 3 * (size_t)x;
    d[0] = on ? s[0] : 0; d[1] = on ? s[1] : 0; d[2] = on ? s[2] : 0;
}

}  // namespace

int k_copy_masked(msg_ctx* ctx, const uint8_t* d_src, size_t sstep, const uint8_t* d_mask, size_t mstep, uint8_t* d_dst, size_t dstep,
                  int w, int h)
{
    dim3 grid((w + 255) / 256, h);
    copy_masked_kernel<<<grid, 256, 0, ctx->stream>>>(d_src, sstep, d_mask, mstep, d_dst, dstep, w);
    MSG_LAUNCHED(ctx);
    MSG_CHECK_LAUNCH(ctx);
    return MSG_OK;
}

int k_canny_nms(msg_ctx* ctx, const uint8_t* d_src, size_t sstep, uint8_t* d_cls, size_t cstep, int w, int h, int low, int high)
{
    dim3 grid((w + CN_TW - 1) / CN_TW, (h + CN_TH - 1) / CN_TH);
    canny_nms_kernel<<<grid, CN_TW * CN_TH, 0, ctx->stream>>>(d_src, sstep, d_cls, cstep, w, h, low, high);
    MSG_LAUNCHED(ctx);
    MSG_CHECK_LAUNCH(ctx);
    return MSG_OK;
}

int k_dilate(msg_ctx* ctx, const uint8_t* d_src, size_t sstep, uint8_t* d_dst, size_t dstep, int w, int h, int kw, int kh)
{
    size_t smem = (size_t)(DL_TW + kw - 1) * (DL_TH + kh - 1);
    MSG_TRY(msg_func_smem(ctx, (const void*)dilate_kernel, smem));
    dim3 grid((w + DL_TW - 1) / DL_TW, (h + DL_TH - 1) / DL_TH);
    dilate_kernel<<<grid, DL_TW * DL_TH, smem, ctx->stream>>>(d_src, sstep, d_dst, dstep, w, h, kw, kh);
    MSG_LAUNCHED(ctx);
    MSG_CHECK_LAUNCH(ctx);
    return MSG_OK;
}

int k_subtract(msg_ctx* ctx, const uint8_t* d_a, size_t astep, const uint8_t* d_b, size_t bstep, uint8_t* d_dst, size_t dstep,
               int w, int h)
{
    dim3 grid((w + 255) / 256, h);
    subtract_kernel<<<grid, 256, 0, ctx->stream>>>(d_a, astep, d_b, bstep, d_dst, dstep, w);
    MSG_LAUNCHED(ctx);
    MSG_CHECK_LAUNCH(ctx);
    return MSG_OK;
}
