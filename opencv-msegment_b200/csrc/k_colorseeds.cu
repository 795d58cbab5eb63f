// k_colorseeds.cu -- the colour-method marker generator of the reference, the caller side
// of the labelling stage, plus the bilateral pre-filter of row a5:
//   white -> black loop            PictureService.java:309-318
//   threshold(40,255,BINARY|OTSU)  :941         (histogram + the sequential double-precision Otsu scan, one thread)
//   distanceTransform(L2, 5)       :1020        (two-pass 5x5 chamfer, float metrics 1 / 1.4 / 2.1969 accumulated in float)
//   normalize(0,1,NORM_MINMAX)     :1021        threshold(.4,1.,BINARY) :348   dilate 3x3 :349-350   convertTo 8U :355-356
//   circle((5,5),3,255,FILLED)     :366         bilateralFilter :490
// (contour labelling, :360-364, lives in k_contours.cu).  Oracle: orc_* functions of the same names, pinned on cv2 4.13.
//
// The chamfer transform is a sequential algorithm: row y of a pass depends on the two previous rows of the same pass and,
// inside the row, on its left (right) neighbour.  The dependence between rows is kept (one CTA walks the rows); inside a row
// the recurrence d[x] = min(t[x], d[x-1] + 1.0f) is evaluated as a scan: every thread runs its chunk sequentially, the
// chunk summaries (value at the chunk end, chunk length) form a monoid under
//       (c1,n1) o (c2,n2) = (min(c2, F^n2(c1)), n1 + n2),      F(v) = fl(v + 1.0f)
// and F^n is evaluated exactly (float additions of 1.0f are exact inside a binade and round once per binade crossing), so
// the result is the sequential float recurrence bit for bit.
#include <float.h>
#include <math.h>

#include "msg_internal.h"

namespace {

// ---------------------------------------------------------------- white -> black (Java loop, :309-318)
__global__ void __launch_bounds__(256) white_to_black_kernel(const uint8_t* __restrict__ src, size_t sstep,
                                                             uint8_t* __restrict__ dst, size_t dstep, int w)
{
    int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y;
    if (x >= w) return;
    const uint8_t* p = src + (size_t)y * sstep + 3 * (size_t)x;
    uint8_t* q = dst + (size_t)y * dstep + 3 * (size_t)x;
    uint8_t b = p[0], g = p[1], r = p[2];
    if (b == 255 && g == 255 && r == 255) b = g = r = 0;
    q[0] = b; q[1] = g; q[2] = r;
}

// ---------------------------------------------------------------- Otsu
__global__ void __launch_bounds__(256) hist_kernel(const uint8_t* __restrict__ src, size_t sstep, int w, int h,
                                                   unsigned* __restrict__ hist)
{
    __shared__ unsigned s_h[256];
    s_h[threadIdx.x] = 0;
    __syncthreads();
    for (int y = blockIdx.x; y < h; y += gridDim.x) {
        const uint8_t* row = src + (size_t)y * sstep;
        for (int x = threadIdx.x; x < w; x += 256) atomicAdd(&s_h[row[x]], 1u);
    }
    __syncthreads();
    if (s_h[threadIdx.x]) atomicAdd(&hist[threadIdx.x], s_h[threadIdx.x]);
}

// getThreshVal_Otsu_8u: the order of the double operations is part of the specification (explicit _rn intrinsics: no
// fused multiply-add contraction)
__global__ void otsu_kernel(const unsigned* __restrict__ hist, int w, int h, int32_t* __restrict__ out)
{
    if (threadIdx.x || blockIdx.x) return;
    double mu = 0, scale = __ddiv_rn(1., (double)(w * h));
    for (int i = 0; i < 256; i++) mu = __dadd_rn(mu, __dmul_rn((double)i, (double)hist[i]));
    mu = __dmul_rn(mu, scale);
    double mu1 = 0, q1 = 0, max_sigma = 0;
    int max_val = 0;
    const double eps = 1.1920928955078125e-07;
    for (int i = 0; i < 256; i++) {
        double p_i = __dmul_rn((double)hist[i], scale);
        mu1 = __dmul_rn(mu1, q1);
        q1 = __dadd_rn(q1, p_i);
        double q2 = __dsub_rn(1., q1);
        if (fmin(q1, q2) < eps || fmax(q1, q2) > 1. - eps) continue;
        mu1 = __ddiv_rn(__dadd_rn(mu1, __dmul_rn((double)i, p_i)), q1);
        double mu2 = __ddiv_rn(__dsub_rn(mu, __dmul_rn(q1, mu1)), q2);
        double dm = __dsub_rn(mu1, mu2);
        double sigma = __dmul_rn(__dmul_rn(__dmul_rn(q1, q2), dm), dm);
        if (sigma > max_sigma) { max_sigma = sigma; max_val = i; }
    }
    *out = max_val;
}

__global__ void __launch_bounds__(256) threshold_u8_kernel(const uint8_t* __restrict__ src, size_t sstep,
                                                           uint8_t* __restrict__ dst, size_t dstep, int w,
                                                           const int32_t* __restrict__ d_thresh, int thresh, int maxval)
{
    int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y;
    if (x >= w) return;
    int t = d_thresh ? *d_thresh : thresh;
    dst[(size_t)y * dstep + x] = (uint8_t)(src[(size_t)y * sstep + x] > t ? maxval : 0);
}

// ---------------------------------------------------------------- chamfer distance transform (L2, 5x5, float metrics)
constexpr int DT_THREADS = 1024;
constexpr float DT_A = 1.0f, DT_B = 1.4f, DT_C = 2.1969f;

// F^n(v), F(v) = fl(v + 1.0f), for v = 0 or v >= 1
__device__ __forceinline__ float dt_advance(float v, int n)
{
    while (n > 0) {
        if (!(v < 16777216.f)) return v;                      // the "infinite" value (FLT_MAX): v + 1 rounds back to v
        if (v < 1.f) { v = __fadd_rn(v, 1.f); n--; continue; }
        int e = (__float_as_int(v) >> 23) - 127;              // v in [2^e, 2^(e+1))
        float lim = __int_as_float((e + 128) << 23);
        int k = (int)ceilf(__fsub_rn(lim, v)) - 1;            // additions that stay below lim are exact
        if (n <= k) return __fadd_rn(v, (float)n);
        v = __fadd_rn(v, (float)k);
        v = __fadd_rn(v, 1.f);                                // the addition that enters the next binade rounds
        n -= k + 1;
    }
    return v;
}

struct dt_sum { float c; int n; };
__device__ __forceinline__ dt_sum dt_compose(dt_sum l, dt_sum r)
{
    dt_sum o;
    o.c = fminf(r.c, dt_advance(l.c, r.n));
    o.n = l.n + r.n;
    return o;
}

// One pass over the rows.  DIR = +1: forward (top -> bottom, left -> right, init = 0 on zero pixels / infinite elsewhere),
// DIR = -1: backward (bottom -> top, right -> left, init = forward value).  `dist` (dense w floats per row) receives the
// forward values and is updated in place by the backward pass; the backward pass also reduces the maximum.
template <int DIR>
__device__ void dt_pass(const uint8_t* __restrict__ src, size_t sstep, float* __restrict__ dist, int w, int h, float* smem,
                        float* s_scan_c, int* s_scan_n, float* d_max)
{
    const int pw = w + 4;                          // two "infinite" pixels on either side
    float* ring = smem;                            // 3 rows of pw floats
    float* init = smem + 3 * pw;                   // 2 rows of w floats (double buffer of the row's initial values)
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int P = (w + DT_THREADS - 1) / DT_THREADS;
    for (int i = tid; i < 3 * pw; i += DT_THREADS) ring[i] = FLT_MAX;
    const int ystart = DIR > 0 ? 0 : h - 1;
    // initial values of the first row
    for (int x = tid; x < w; x += DT_THREADS)
        init[x] = DIR > 0 ? (src[(size_t)ystart * sstep + x] ? FLT_MAX : 0.f) : dist[(size_t)ystart * w + x];
    float vmax = 0.f;
    __syncthreads();
    for (int it = 0; it < h; it++) {
        const int y = DIR > 0 ? it : h - 1 - it;
        float* cur = ring + (it % 3) * pw + 2;
        const float* r1 = ring + ((it + 2) % 3) * pw + 2;     // previous row of this pass
        const float* r2 = ring + ((it + 1) % 3) * pw + 2;     // the row before it
        const float* ini = init + (it & 1) * w;
        float* ini_next = init + ((it + 1) & 1) * w;
        // (a) store the previous row (final for this pass), prefetch the next row's initial values
        if (it > 0) {
            const int yp = y - DIR;
            for (int x = tid; x < w; x += DT_THREADS) {
                float v = r1[x];
                dist[(size_t)yp * w + x] = v;
                if (DIR < 0) vmax = fmaxf(vmax, v);
            }
        }
        if (it + 1 < h) {
            const int yn = y + DIR;
            for (int x = tid; x < w; x += DT_THREADS)
                ini_next[x] = DIR > 0 ? (src[(size_t)yn * sstep + x] ? FLT_MAX : 0.f) : dist[(size_t)yn * w + x];
        }
        // (b) the thread's chunk, sequentially, with an unknown ("infinite") carry
        const int m0 = tid * P, m1 = min(w, m0 + P);           // chunk in pass order (mirrored for the backward pass)
        float run = FLT_MAX;
        for (int m = m0; m < m1; m++) {
            const int x = DIR > 0 ? m : w - 1 - m;
            float t = ini[x];
            if (t > DT_A) {
                t = fminf(t, __fadd_rn(r2[x - 1], DT_C));
                t = fminf(t, __fadd_rn(r2[x + 1], DT_C));
                t = fminf(t, __fadd_rn(r1[x - 2], DT_C));
                t = fminf(t, __fadd_rn(r1[x - 1], DT_B));
                t = fminf(t, __fadd_rn(r1[x], DT_A));
                t = fminf(t, __fadd_rn(r1[x + 1], DT_B));
                t = fminf(t, __fadd_rn(r1[x + 2], DT_C));
                t = fminf(t, __fadd_rn(run, DT_A));
            }
            run = t;
            cur[x] = t;
        }
        // (c) scan of the chunk summaries
        dt_sum s;
        s.c = m0 < m1 ? run : FLT_MAX;
        s.n = max(m1 - m0, 0);
        dt_sum inc = s;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            dt_sum l;
            l.c = __shfl_up_sync(0xffffffffu, inc.c, o);
            l.n = __shfl_up_sync(0xffffffffu, inc.n, o);
            if (lane >= o) inc = dt_compose(l, inc);
        }
        dt_sum exc;                                            // exclusive prefix inside the warp
        exc.c = __shfl_up_sync(0xffffffffu, inc.c, 1);
        exc.n = __shfl_up_sync(0xffffffffu, inc.n, 1);
        if (lane == 0) { exc.c = FLT_MAX; exc.n = 0; }
        if (lane == 31) { s_scan_c[warp] = inc.c; s_scan_n[warp] = inc.n; }
        __syncthreads();
        if (warp == 0) {
            dt_sum v;
            v.c = s_scan_c[lane]; v.n = s_scan_n[lane];
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                dt_sum l;
                l.c = __shfl_up_sync(0xffffffffu, v.c, o);
                l.n = __shfl_up_sync(0xffffffffu, v.n, o);
                if (lane >= o) v = dt_compose(l, v);
            }
            float ec = __shfl_up_sync(0xffffffffu, v.c, 1);    // exclusive: what enters warp `lane`
            if (lane == 0) ec = FLT_MAX;
            s_scan_c[32 + lane] = ec;
        }
        __syncthreads();
        // (d) fold the carry into the chunk
        {
            float carry = fminf(exc.c, dt_advance(s_scan_c[32 + warp], exc.n));
            if (carry < 16777216.f)
                for (int m = m0; m < m1; m++) {
                    const int x = DIR > 0 ? m : w - 1 - m;
                    carry = __fadd_rn(carry, DT_A);
                    if (carry < cur[x]) cur[x] = carry;
                    else break;                                // monotone: the carry cannot matter further right
                }
        }
        __syncthreads();
    }
    // the last row
    {
        const int y = DIR > 0 ? h - 1 : 0;
        const float* last = ring + ((h - 1) % 3) * pw + 2;
        for (int x = tid; x < w; x += DT_THREADS) {
            float v = last[x];
            dist[(size_t)y * w + x] = v;
            if (DIR < 0) vmax = fmaxf(vmax, v);
        }
    }
    if (DIR < 0) {
#pragma unroll
        for (int o = 16; o; o >>= 1) vmax = fmaxf(vmax, __shfl_xor_sync(0xffffffffu, vmax, o));
        __syncthreads();
        if (lane == 0) s_scan_c[warp] = vmax;
        __syncthreads();
        if (warp == 0) {
            float v = s_scan_c[lane];
#pragma unroll
            for (int o = 16; o; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
            if (lane == 0) *d_max = v;
        }
    }
    __syncthreads();
}

__global__ void __launch_bounds__(DT_THREADS, 1) dt_kernel(const uint8_t* __restrict__ src, size_t sstep,
                                                           float* __restrict__ dist, int w, int h, float* d_max)
{
    extern __shared__ float dt_smem[];
    __shared__ float s_scan_c[64];
    __shared__ int s_scan_n[32];
    dt_pass<1>(src, sstep, dist, w, h, dt_smem, s_scan_c, s_scan_n, d_max);
    __threadfence_block();
    dt_pass<-1>(src, sstep, dist, w, h, dt_smem, s_scan_c, s_scan_n, d_max);
}

// ---------------------------------------------------------------- float planes: normalise, threshold, dilate, convert
// Core.normalize(src, dst, alpha, beta, NORM_MINMAX) on 32F with the source extrema on the device: scale and shift in
// double as cv::normalize computes them, dst = fma(src, (float)scale, (float)shift) (cv2's vector path fuses).
__global__ void __launch_bounds__(256) minmax_f32_kernel(const float* __restrict__ src, size_t n, float* __restrict__ mm)
{
    __shared__ float s_lo[8], s_hi[8];
    float lo = FLT_MAX, hi = -FLT_MAX;
    for (size_t i = (size_t)blockIdx.x * 256 + threadIdx.x; i < n; i += (size_t)gridDim.x * 256) {
        float v = src[i];
        lo = fminf(lo, v); hi = fmaxf(hi, v);
    }
#pragma unroll
    for (int o = 16; o; o >>= 1) {
        lo = fminf(lo, __shfl_xor_sync(0xffffffffu, lo, o));
        hi = fmaxf(hi, __shfl_xor_sync(0xffffffffu, hi, o));
    }
    if ((threadIdx.x & 31) == 0) { s_lo[threadIdx.x >> 5] = lo; s_hi[threadIdx.x >> 5] = hi; }
    __syncthreads();
    if (threadIdx.x == 0) {
        for (int i = 1; i < 8; i++) { lo = fminf(lo, s_lo[i]); hi = fmaxf(hi, s_hi[i]); }
        // float min / max through the order-preserving integer view (all values finite)
        int ilo = __float_as_int(lo), ihi = __float_as_int(hi);
        ilo = ilo >= 0 ? ilo : ilo ^ 0x7fffffff;
        ihi = ihi >= 0 ? ihi : ihi ^ 0x7fffffff;
        atomicMin((int*)mm, ilo);
        atomicMax((int*)mm + 1, ihi);
    }
}

__global__ void __launch_bounds__(256) normalize_f32_kernel(const float* __restrict__ src, float* __restrict__ dst, size_t n,
                                                            const float* __restrict__ mm, double alpha, double beta)
{
    size_t i = (size_t)blockIdx.x * 256 + threadIdx.x;
    if (i >= n) return;
    int ilo = ((const int*)mm)[0], ihi = ((const int*)mm)[1];
    ilo = ilo >= 0 ? ilo : ilo ^ 0x7fffffff;
    ihi = ihi >= 0 ? ihi : ihi ^ 0x7fffffff;
    double smin = (double)__int_as_float(ilo), smax = (double)__int_as_float(ihi);
    double dmin = fmin(alpha, beta), dmax = fmax(alpha, beta);
    double range = __dsub_rn(smax, smin);
    double scale = __dmul_rn(__dsub_rn(dmax, dmin), range > 2.220446049250313e-16 ? __ddiv_rn(1., range) : 0.);
    double shift = __dsub_rn(dmin, __dmul_rn(smin, scale));
    dst[i] = __fmaf_rn(src[i], __double2float_rn(scale), __double2float_rn(shift));
}

__global__ void __launch_bounds__(256) threshold_f32_kernel(const float* __restrict__ src, float* __restrict__ dst, size_t n,
                                                            float thresh, float maxval)
{
    size_t i = (size_t)blockIdx.x * 256 + threadIdx.x;
    if (i < n) dst[i] = src[i] > thresh ? maxval : 0.f;
}

__global__ void __launch_bounds__(256) dilate_f32_kernel(const float* __restrict__ src, float* __restrict__ dst, int w, int h,
                                                         int kw, int kh)
{
    int x = blockIdx.x * 32 + (threadIdx.x & 31), y = blockIdx.y * 8 + (threadIdx.x >> 5);
    if (x >= w || y >= h) return;
    float best = -FLT_MAX;
    for (int yy = y - kh / 2; yy < y - kh / 2 + kh; yy++)
        for (int xx = x - kw / 2; xx < x - kw / 2 + kw; xx++)
            if (yy >= 0 && yy < h && xx >= 0 && xx < w) best = fmaxf(best, src[(size_t)yy * w + xx]);
    dst[(size_t)y * w + x] = best;
}

__global__ void __launch_bounds__(256) f32_to_u8_kernel(const float* __restrict__ src, uint8_t* __restrict__ dst, size_t dstep,
                                                        int w, int h)
{
    int x = blockIdx.x * 256 + threadIdx.x, y = blockIdx.y;
    if (x >= w) return;
    int v = __float2int_rn(src[(size_t)y * w + x]);           // saturate_cast<uchar>(cvRound(v))
    dst[(size_t)y * dstep + x] = (uint8_t)min(max(v, 0), 255);
}

// ---------------------------------------------------------------- filled circle (spans computed on the host)
__global__ void fill_spans_i32_kernel(int32_t* __restrict__ img, size_t step, const int* __restrict__ spans, int nspans,
                                      int32_t value)
{
    int s = blockIdx.x;
    if (s >= nspans) return;
    int y = spans[3 * s], xa = spans[3 * s + 1], xb = spans[3 * s + 2];
    int32_t* row = (int32_t*)((uint8_t*)img + (size_t)y * step);
    for (int x = xa + threadIdx.x; x <= xb; x += blockDim.x) row[x] = value;
}

// ---------------------------------------------------------------- bilateral filter (8UC1 / 8UC3, BORDER_REFLECT_101)
__device__ __forceinline__ int reflect101(int p, int n)
{
    if (n == 1) return 0;
    while (p < 0 || p >= n) p = p < 0 ? -p : 2 * (n - 1) - p;
    return p;
}

template <int CN>
__global__ void __launch_bounds__(256) bilateral_kernel(const uint8_t* __restrict__ src, size_t sstep, uint8_t* __restrict__ dst,
                                                        size_t dstep, int w, int h, int radius, int maxk,
                                                        const float* __restrict__ space_w, const short2* __restrict__ space_ofs,
                                                        const float* __restrict__ color_w)
{
    extern __shared__ uint8_t bl_tile[];
    const int tw = 32 + 2 * radius, th = 8 + 2 * radius;
    const int bx = blockIdx.x * 32 - radius, by = blockIdx.y * 8 - radius;
    for (int i = threadIdx.x; i < tw * th; i += 256) {
        int yy = reflect101(by + i / tw, h), xx = reflect101(bx + i % tw, w);
        const uint8_t* p = src + (size_t)yy * sstep + (size_t)xx * CN;
#pragma unroll
        for (int c = 0; c < CN; c++) bl_tile[i * CN + c] = p[c];
    }
    __syncthreads();
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
    const int x = blockIdx.x * 32 + tx, y = blockIdx.y * 8 + ty;
    if (x >= w || y >= h) return;
    const uint8_t* c0 = bl_tile + ((ty + radius) * tw + tx + radius) * CN;
    float sum[CN], wsum = 0.f;
#pragma unroll
    for (int c = 0; c < CN; c++) sum[c] = 0.f;
    for (int k = 0; k < maxk; k++) {
        short2 o = __ldg(space_ofs + k);
        const uint8_t* p = c0 + (o.y * tw + o.x) * CN;
        int diff = 0;
#pragma unroll
        for (int c = 0; c < CN; c++) diff += abs((int)p[c] - (int)c0[c]);
        float wgt = __fmul_rn(__ldg(space_w + k), __ldg(color_w + diff));
#pragma unroll
        for (int c = 0; c < CN; c++) sum[c] = __fadd_rn(sum[c], __fmul_rn((float)p[c], wgt));
        wsum = __fadd_rn(wsum, wgt);
    }
    uint8_t* q = dst + (size_t)y * dstep + (size_t)x * CN;
    if (CN == 1) q[0] = (uint8_t)__float2int_rn(__fdiv_rn(sum[0], wsum));
    else {
        float inv = __fdiv_rn(1.f, wsum);
#pragma unroll
        for (int c = 0; c < CN; c++) q[c] = (uint8_t)__float2int_rn(__fmul_rn(sum[c], inv));
    }
}

inline unsigned blocks_for(size_t n, int per) { return (unsigned)((n + per - 1) / per); }

}  // namespace

int k_white_to_black(msg_ctx* ctx, const uint8_t* d_src, size_t sstep, uint8_t* d_dst, size_t dstep, int w, int h)
{
    dim3 grid((w + 255) / 256, h);
    white_to_black_kernel<<<grid, 256, 0, ctx->stream>>>(d_src, sstep, d_dst, dstep, w);
    MSG_LAUNCHED(ctx);
    MSG_CHECK_LAUNCH(ctx);
    return MSG_OK;
}

// d_hist: 256 unsigned (scratch); d_thresh: the selected threshold (device)
int k_otsu(msg_ctx* ctx, const uint8_t* d_src, size_t sstep, int w, int h, unsigned* d_hist, int32_t* d_thresh)
{
    MSG_CUDA(ctx, cudaMemsetAsync(d_hist, 0, 256 * sizeof(unsigned), ctx->stream));
    int nb = min(h, 4 * ctx->sm_count);
    hist_kernel<<<nb, 256, 0, ctx->stream>>>(d_src, sstep, w, h, d_hist);
    MSG_LAUNCHED(ctx);
    otsu_kernel<<<1, 32, 0, ctx->stream>>>(d_hist, w, h, d_thresh);
    MSG_LAUNCHED(ctx);
    MSG_CHECK_LAUNCH(ctx);
    return MSG_OK;
}

int k_threshold_u8(msg_ctx* ctx, const uint8_t* d_src, size_t sstep, uint8_t* d_dst, size_t dstep, int w, int h,
                   const int32_t* d_thresh, int thresh, int maxval)
{
    dim3 grid((w + 255) / 256, h);
    threshold_u8_kernel<<<grid, 256, 0, ctx->stream>>>(d_src, sstep, d_dst, dstep, w, d_thresh, thresh, maxval);
    MSG_LAUNCHED(ctx);
    MSG_CHECK_LAUNCH(ctx);
    return MSG_OK;
}

int k_distance_transform_max_width(msg_ctx* ctx) { return ((ctx->max_smem_optin - 1024) / 4 - 12) / 5; }

// d_dist: dense w*h floats; d_max (device float, may not be NULL): maximum of the result
int k_distance_transform(msg_ctx* ctx, const uint8_t* d_src, size_t sstep, float* d_dist, int w, int h, float* d_max)
{
    size_t smem = ((size_t)3 * (w + 4) + 2 * (size_t)w) * sizeof(float);
    if (w > k_distance_transform_max_width(ctx))
        return msg_fail(ctx, MSG_EINVAL, "distanceTransform supports rows up to %d pixels", k_distance_transform_max_width(ctx));
    MSG_CUDA(ctx, cudaFuncSetAttribute(dt_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    dt_kernel<<<1, DT_THREADS, smem, ctx->stream>>>(d_src, sstep, d_dist, w, h, d_max);
    MSG_LAUNCHED(ctx);
    MSG_CHECK_LAUNCH(ctx);
    return MSG_OK;
}

// d_mm: 2 floats of scratch (device)
int k_normalize_minmax_f32(msg_ctx* ctx, const float* d_src, float* d_dst, int w, int h, double alpha, double beta, float* d_mm)
{
    size_t n = (size_t)w * h;
    const int init[2] = {0x7f7fffff, (int)0x80000000};         // FLT_MAX / -FLT_MAX in the order-preserving integer view
    MSG_CUDA(ctx, cudaMemcpyAsync(d_mm, init, sizeof(init), cudaMemcpyHostToDevice, ctx->stream));
    minmax_f32_kernel<<<min(blocks_for(n, 256), (unsigned)(8 * ctx->sm_count)), 256, 0, ctx->stream>>>(d_src, n, d_mm);
    MSG_LAUNCHED(ctx);
    normalize_f32_kernel<<<blocks_for(n, 256), 256, 0, ctx->stream>>>(d_src, d_dst, n, d_mm, alpha, beta);
    MSG_LAUNCHED(ctx);
    MSG_CHECK_LAUNCH(ctx);
    return MSG_OK;
}

int k_threshold_f32(msg_ctx* ctx, const float* d_src, float* d_dst, int w, int h, float thresh, float maxval)
{
    size_t n = (size_t)w * h;
    threshold_f32_kernel<<<blocks_for(n, 256), 256, 0, ctx->stream>>>(d_src, d_dst, n, thresh, maxval);
    MSG_LAUNCHED(ctx);
    MSG_CHECK_LAUNCH(ctx);
    return MSG_OK;
}

int k_dilate_f32(msg_ctx* ctx, const float* d_src, float* d_dst, int w, int h, int kw, int kh)
{
    dim3 grid((w + 31) / 32, (h + 7) / 8);
    dilate_f32_kernel<<<grid, 256, 0, ctx->stream>>>(d_src, d_dst, w, h, kw, kh);
    MSG_LAUNCHED(ctx);
    MSG_CHECK_LAUNCH(ctx);
    return MSG_OK;
}

int k_f32_to_u8(msg_ctx* ctx, const float* d_src, uint8_t* d_dst, size_t dstep, int w, int h)
{
    dim3 grid((w + 255) / 256, h);
    f32_to_u8_kernel<<<grid, 256, 0, ctx->stream>>>(d_src, d_dst, dstep, w, h);
    MSG_LAUNCHED(ctx);
    MSG_CHECK_LAUNCH(ctx);
    return MSG_OK;
}

// cv::circle(img, (cx,cy), radius, value, FILLED) on 32SC1: OpenCV's midpoint circle, spans clipped to the image.
// d_spans: device scratch for 3 * 4 * (radius + 1) ints.
int k_circle_filled_i32(msg_ctx* ctx, int32_t* d_img, size_t step, int w, int h, int cx, int cy, int radius, int32_t value,
                        int* d_spans, int* h_spans)
{
    int ns = 0;
    int err = 0, dx = radius, dy = 0, plus = 1, minus = (radius << 1) - 1;
    while (dx >= dy) {
        const int ys[4] = {cy - dy, cy + dy, cy - dx, cy + dx};
        const int xa[4] = {cx - dx, cx - dx, cx - dy, cx - dy}, xb[4] = {cx + dx, cx + dx, cx + dy, cx + dy};
        for (int k = 0; k < 4; k++) {
            if (ys[k] < 0 || ys[k] >= h) continue;
            int a = xa[k] < 0 ? 0 : xa[k], b = xb[k] >= w ? w - 1 : xb[k];
            if (a > b) continue;
            h_spans[3 * ns] = ys[k]; h_spans[3 * ns + 1] = a; h_spans[3 * ns + 2] = b;
            ns++;
        }
        dy++;
        err += plus;
        plus += 2;
        int mask = (err <= 0) - 1;
        err -= minus & mask;
        dx += mask;
        minus -= mask & 2;
    }
    if (!ns) return MSG_OK;
    MSG_CUDA(ctx, cudaMemcpyAsync(d_spans, h_spans, (size_t)ns * 3 * sizeof(int), cudaMemcpyHostToDevice, ctx->stream));
    fill_spans_i32_kernel<<<ns, 128, 0, ctx->stream>>>(d_img, step, d_spans, ns, value);
    MSG_LAUNCHED(ctx);
    MSG_CHECK_LAUNCH(ctx);
    return MSG_OK;
}

// d_tables: device scratch for maxk floats + maxk short2 + 256*cn floats, filled from the host arrays
int k_bilateral(msg_ctx* ctx, const uint8_t* d_src, size_t sstep, uint8_t* d_dst, size_t dstep, int w, int h, int cn, int radius,
                int maxk, const float* d_space_w, const short* d_space_ofs, const float* d_color_w)
{
    size_t smem = (size_t)(32 + 2 * radius) * (8 + 2 * radius) * cn;
    dim3 grid((w + 31) / 32, (h + 7) / 8);
    if (cn == 1) {
        MSG_CUDA(ctx, cudaFuncSetAttribute(bilateral_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        bilateral_kernel<1><<<grid, 256, smem, ctx->stream>>>(d_src, sstep, d_dst, dstep, w, h, radius, maxk, d_space_w,
                                                             (const short2*)d_space_ofs, d_color_w);
    } else {
        MSG_CUDA(ctx, cudaFuncSetAttribute(bilateral_kernel<3>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        bilateral_kernel<3><<<grid, 256, smem, ctx->stream>>>(d_src, sstep, d_dst, dstep, w, h, radius, maxk, d_space_w,
                                                             (const short2*)d_space_ofs, d_color_w);
    }
    MSG_LAUNCHED(ctx);
    MSG_CHECK_LAUNCH(ctx);
    return MSG_OK;
}
