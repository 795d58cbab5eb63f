// k_dt_fixed.cu -- Imgproc.distanceTransform(src, dst, CV_DIST_L2, 5) (PictureService.java:1020) in the arithmetic of a
// NON-IPP OpenCV build (distransform.cpp, distanceTransform_5x5: 16.16 fixed point, metrics 65536 / 91750 / 143976), selected
// with msg_set_option("dt_fixed", 1).  This is what the openpnp 3.4.2 natives the reference binds compute; the default mode
// (k_colorseeds.cu, dt_wave_kernel) reproduces the float arithmetic of the IPP-backed cv2 4.13 build instead.
//
// Integer min-plus is exact and associative, so -- unlike the float mode, whose additions do not associate and whose published
// pixel order therefore has to be kept -- the two-pass recurrence can be replaced by what it computes: the 5x5 chamfer
// distance to the nearest zero pixel, dist(p) = min over zero pixels z of f(p - z), with f the gauge of the chamfer mask
//     f(dx, dy) = lo * LONG + (hi - 2 lo) * HV          if 2 lo <= hi        (knight moves + axis moves)
//                 (hi - lo) * LONG + (2 lo - hi) * DIAG  otherwise            (knight moves + diagonal moves)
// with lo = min(|dx|, |dy|), hi = max(|dx|, |dy|).  (The unit-cost points of the sixteen mask directions form a convex polygon
// for these metrics -- LONG < HV + DIAG, 2 LONG >= 4 HV, 2 LONG >= 3 DIAG -- so a shortest path uses two adjacent directions,
// stays inside the bounding box of its end points, i.e. inside the image, and can be ordered "forward moves first", which is
// the path the two raster passes find.)  f is non-decreasing in |dy| for a fixed dx, so the nearest zero of a COLUMN is the
// only one of that column that matters:
//   1. dtf_band_kernel / dtf_vertical_kernel (whole GPU): g(y, x) = rows to the nearest zero pixel of column x (u16, 0xFFFF =
//      the column has none), through per-band first / last zero tables so that no thread walks a whole column;
//   2. dtf_row_kernel (whole GPU, one thread per pixel): dist = min over x' of f(x - x', g(y, x')), searched outwards from x
//      and stopped as soon as HV * |x - x'| reaches the best value so far (work per pixel = 2 * dist, 5-20 tests on the
//      pipeline's Otsu masks).
// An image without any zero pixel reads DIST_MAX * 2^-16 = 65533.805 everywhere (cv2 4.13 clamps there).
// Checked bit for bit against the CPU oracle's two-pass fixed-point recurrence, itself pinned on cv2 with IPP switched off
// (tests/golden/dt_fixed.npz, tests/test_color_seeds.py).
#include "msg_internal.h"

namespace {

constexpr unsigned DTF_HV = 65536u, DTF_DIAG = 91750u, DTF_LONG = 143976u;
constexpr unsigned DTF_DIST_MAX = 0xFFFFFFFFu - DTF_LONG;
constexpr int DTF_BAND = 32;               // rows per band
constexpr unsigned DTF_NONE = 0xFFFFu;

// tables: first[band][x], last[band][x] = image row of the first / last zero pixel of column x inside the band, -1 = none
__global__ void __launch_bounds__(128) dtf_band_kernel(const uint8_t* __restrict__ src, size_t sstep, int w, int h,
                                                       int16_t* __restrict__ first, int16_t* __restrict__ last, int* __restrict__ any_zero)
{
    const int x = blockIdx.x * 128 + threadIdx.x, band = blockIdx.y;
    int f = -1, l = -1;
    if (x < w) {
        const int y0 = band * DTF_BAND, y1 = min(h, y0 + DTF_BAND);
#pragma unroll 8
        for (int y = y0; y < y1; y++)
            if (!src[(size_t)y * sstep + x]) {
                if (f < 0) f = y;
                l = y;
            }
        first[(size_t)band * w + x] = (int16_t)f;
        last[(size_t)band * w + x] = (int16_t)l;
    }
    if (__syncthreads_or(f >= 0) && threadIdx.x == 0) *any_zero = 1;
}

__global__ void __launch_bounds__(128) dtf_vertical_kernel(const uint8_t* __restrict__ src, size_t sstep, int w, int h, int nbands,
                                                           const int16_t* __restrict__ first, const int16_t* __restrict__ last,
                                                           uint16_t* __restrict__ g)
{
    const int x = blockIdx.x * 128 + threadIdx.x, band = blockIdx.y;
    if (x >= w) return;
    const int y0 = band * DTF_BAND, y1 = min(h, y0 + DTF_BAND);
    int above = -1, below = -1;                     // rows of the nearest zero above / below the band
    for (int b = band - 1; b >= 0 && above < 0; b--) above = last[(size_t)b * w + x];
    for (int b = band + 1; b < nbands && below < 0; b++) below = first[(size_t)b * w + x];
    constexpr int BIG = 1 << 20;
    int d[DTF_BAND];
    int run = above >= 0 ? y0 - 1 - above : BIG;    // distance of row y0 - 1 to the nearest zero at or above it
#pragma unroll
    for (int i = 0; i < DTF_BAND; i++) {
        const int y = y0 + i;
        const bool zero = y < y1 && !src[(size_t)y * sstep + x];
        run = zero ? 0 : run + 1;
        d[i] = run;
    }
    run = below >= 0 ? below - (y0 + DTF_BAND) : BIG;   // distance of row y0 + BAND to the nearest zero at or below it
#pragma unroll
    for (int i = DTF_BAND - 1; i >= 0; i--) {
        run = d[i] == 0 ? 0 : run + 1;
        const int v = min(d[i], run);
        if (y0 + i < y1) g[(size_t)(y0 + i) * w + x] = (uint16_t)(v >= (int)DTF_NONE ? DTF_NONE : v);
    }
}

__device__ __forceinline__ unsigned dtf_cost(unsigned d, unsigned gg)
{
    const unsigned lo = min(d, gg), hi = max(d, gg);
    return 2 * lo <= hi ? lo * DTF_LONG + (hi - 2 * lo) * DTF_HV : (hi - lo) * DTF_LONG + (2 * lo - hi) * DTF_DIAG;
}

__global__ void __launch_bounds__(256) dtf_row_kernel(const uint16_t* __restrict__ g, int w, const int* __restrict__ any_zero,
                                                      float* __restrict__ dist, unsigned* __restrict__ d_max_bits)
{
    const int x = blockIdx.x * 256 + threadIdx.x, y = blockIdx.y;
    float out = 0.f;
    if (x < w) {
        unsigned best;
        if (!*any_zero) best = DTF_DIST_MAX;
        else {
            const uint16_t* __restrict__ row = g + (size_t)y * w;
            const unsigned g0 = row[x];
            best = g0 == DTF_NONE ? 0xFFFFFFFFu : g0 * DTF_HV;
            const int reach = max(x, w - 1 - x);
            for (int d = 1; d <= reach && (unsigned)d * DTF_HV < best; d++) {
                if (d <= x) {
                    const unsigned gl = __ldg(row + x - d);
                    if (gl != DTF_NONE) best = min(best, dtf_cost((unsigned)d, gl));
                }
                if (x + d < w) {
                    const unsigned gr = __ldg(row + x + d);
                    if (gr != DTF_NONE) best = min(best, dtf_cost((unsigned)d, gr));
                }
            }
            if (best > DTF_DIST_MAX) best = DTF_DIST_MAX;
        }
        out = __fmul_rn(__uint2float_rn(best), 1.f / 65536);
        dist[(size_t)y * w + x] = out;
    }
    unsigned m = __float_as_uint(out);              // non-negative floats order like their bit patterns
#pragma unroll
    for (int o = 16; o; o >>= 1) m = max(m, __shfl_xor_sync(0xffffffffu, m, o));
    if ((threadIdx.x & 31) == 0 && m) atomicMax(d_max_bits, m);
}

}  // namespace

int k_distance_transform_fixed_max_dim() { return 16384; }

// d_dist: dense w*h floats; d_max (device float): maximum of the result.  Scratch: ctx->d_scratch.
int k_distance_transform_fixed(msg_ctx* ctx, const uint8_t* d_src, size_t sstep, float* d_dist, int w, int h, float* d_max)
{
    if (w > k_distance_transform_fixed_max_dim() || h > k_distance_transform_fixed_max_dim())
        return msg_fail(ctx, MSG_EINVAL, "distanceTransform (16.16 fixed-point mode) supports images up to %d x %d",
                        k_distance_transform_fixed_max_dim(), k_distance_transform_fixed_max_dim());
    const int nbands = (h + DTF_BAND - 1) / DTF_BAND;
    const size_t n = (size_t)w * h, tab = (size_t)nbands * w;
    const size_t off_first = (n * 2 + 255) & ~(size_t)255, off_last = off_first + ((tab * 2 + 255) & ~(size_t)255);
    const size_t off_flag = off_last + ((tab * 2 + 255) & ~(size_t)255);
    MSG_TRY(msg_reserve(ctx, (void**)&ctx->d_scratch, &ctx->d_scratch_cap, off_flag + 256));
    uint8_t* s = (uint8_t*)ctx->d_scratch;
    uint16_t* g = (uint16_t*)s;
    int16_t* first = (int16_t*)(s + off_first);
    int16_t* last = (int16_t*)(s + off_last);
    int* flag = (int*)(s + off_flag);
    cudaStream_t st = ctx->stream;
    MSG_CUDA(ctx, cudaMemsetAsync(flag, 0, sizeof(int), st));
    MSG_CUDA(ctx, cudaMemsetAsync(d_max, 0, sizeof(float), st));
    dim3 gridb((w + 127) / 128, nbands);
    dtf_band_kernel<<<gridb, 128, 0, st>>>(d_src, sstep, w, h, first, last, flag);
    MSG_LAUNCHED(ctx);
    dtf_vertical_kernel<<<gridb, 128, 0, st>>>(d_src, sstep, w, h, nbands, first, last, g);
    MSG_LAUNCHED(ctx);
    dim3 gridr((w + 255) / 256, h);
    dtf_row_kernel<<<gridr, 256, 0, st>>>(g, w, flag, d_dist, (unsigned*)d_max);
    MSG_LAUNCHED(ctx);
    MSG_CHECK_LAUNCH(ctx);
    return MSG_OK;
}
