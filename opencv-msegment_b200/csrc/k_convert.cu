// k_convert.cu -- layout conversion, Gaussian pyramid (cv::pyrDown / cv::pyrUp integer formulas,
// SURVEY.md App. A.3), the pyramid change-mask of cv::pyrMeanShiftFiltering (App. A.2), label
// rendering helpers and the synthetic-image generator.  All HBM-bound streaming kernels.
#include "msg_internal.h"

namespace {

__device__ __forceinline__ int reflect101(int p, int n)
{
    if (n == 1) return 0;
    // |p| < 2n always holds for the 5-tap kernels used here, but loop for safety on tiny n
    while (p < 0 || p >= n) p = p < 0 ? -p : 2 * (n - 1) - p;
    return p;
}

__device__ __forceinline__ int clampi(int v, int lo, int hi) { return v < lo ? lo : (v > hi ? hi : v); }

// ------------------------------------------------------------------ BGR (8UC3, byte step) -> plane
__global__ void __launch_bounds__(256) bgr_to_plane_kernel(const uint8_t* __restrict__ src, size_t step,
                                                           uint32_t* __restrict__ dst, int w, int rows, int pitch)
{
    int x = blockIdx.x * blockDim.x + threadIdx.x;
    int y = blockIdx.y;
    if (x >= w || y >= rows) return;
    const uint8_t* p = src + (size_t)y * step + 3 * (size_t)x;
    uint32_t v = (uint32_t)p[0] | ((uint32_t)p[1] << 8) | ((uint32_t)p[2] << 16) | 0x01000000u;
    dst[(size_t)y * pitch + x] = v;
}

__global__ void __launch_bounds__(256) plane_to_bgr_kernel(const uint32_t* __restrict__ src, int pitch, int w,
                                                           int nrows, uint8_t* __restrict__ dst, size_t step)
{
    int x = blockIdx.x * blockDim.x + threadIdx.x;
    int y = blockIdx.y;
    if (x >= w || y >= nrows) return;
    uint32_t v = src[(size_t)y * pitch + x];
    uint8_t* p = dst + (size_t)y * step + 3 * (size_t)x;
    p[0] = (uint8_t)v; p[1] = (uint8_t)(v >> 8); p[2] = (uint8_t)(v >> 16);
}

// Four pixels per thread: 12 bytes of BGR as three 32-bit words against one 16-byte access of packed words (rows whose width
// is a multiple of 4 and whose BGR side is 4-byte aligned; everything else takes the one-pixel kernels above).
__global__ void __launch_bounds__(256) bgr_to_plane_v4_kernel(const uint8_t* __restrict__ src, size_t step,
                                                              uint32_t* __restrict__ dst, int w4, int rows, int pitch)
{
    int q = blockIdx.x * blockDim.x + threadIdx.x;
    int y = blockIdx.y;
    if (q >= w4 || y >= rows) return;
    const uint32_t* p = (const uint32_t*)(src + (size_t)y * step) + 3 * (size_t)q;
    uint32_t w0 = __ldg(p), w1 = __ldg(p + 1), w2 = __ldg(p + 2);
    uint4 o;
    o.x = (w0 & 0x00FFFFFFu) | 0x01000000u;
    o.y = (w0 >> 24) | ((w1 & 0xFFFFu) << 8) | 0x01000000u;
    o.z = (w1 >> 16) | ((w2 & 0xFFu) << 16) | 0x01000000u;
    o.w = (w2 >> 8) | 0x01000000u;
    *(uint4*)(dst + (size_t)y * pitch + 4 * (size_t)q) = o;
}

__global__ void __launch_bounds__(256) plane_to_bgr_v4_kernel(const uint32_t* __restrict__ src, int pitch, int w4, int nrows,
                                                              uint8_t* __restrict__ dst, size_t step)
{
    int q = blockIdx.x * blockDim.x + threadIdx.x;
    int y = blockIdx.y;
    if (q >= w4 || y >= nrows) return;
    uint4 c = *(const uint4*)(src + (size_t)y * pitch + 4 * (size_t)q);
    uint32_t* p = (uint32_t*)(dst + (size_t)y * step) + 3 * (size_t)q;
    p[0] = (c.x & 0x00FFFFFFu) | (c.y << 24);
    p[1] = ((c.y >> 8) & 0xFFFFu) | (c.z << 16);
    p[2] = ((c.z >> 16) & 0xFFu) | (c.w << 8);
}

// ------------------------------------------------------------------ pyrDown: 5x5 [1 4 6 4 1]^2, (acc+128)>>8
// Separable in 16-bit lanes: the horizontal sums of a row are at most 16 * 255, the vertical weights add up to 16 again, so
// (B, R) and (G) stay below 2^16 in two registers and `(acc + 128) >> 8` is applied to both lanes at once.  The five reflected
// column indices are computed once per thread, not once per tap (the kernel was instruction-bound: ~250 instructions per pixel).
__global__ void __launch_bounds__(256) pyr_down_kernel(msg_plane s, msg_plane d)
{
    int x = blockIdx.x * blockDim.x + threadIdx.x;
    int r = blockIdx.y;  // stored row of d
    if (x >= d.w || r >= d.rows) return;
    int y = d.y0 + r;    // global row at the coarse level
    int xi[5];
#pragma unroll
    for (int b = 0; b < 5; b++) xi[b] = reflect101(2 * x + b - 2, s.w);
    uint32_t lo = 0, hi = 0;
#pragma unroll
    for (int a = -2; a <= 2; a++) {
        int yy = reflect101(2 * y + a, s.hfull) - s.y0;
        yy = clampi(yy, 0, s.rows - 1);  // rows outside a strip's halo: value unused by valid outputs
        const uint32_t* row = s.p + (size_t)yy * s.pitch;
        const uint32_t v0 = __ldg(row + xi[0]), v1 = __ldg(row + xi[1]), v2 = __ldg(row + xi[2]), v3 = __ldg(row + xi[3]),
                       v4 = __ldg(row + xi[4]);
        const uint32_t M = 0x00FF00FFu;
        const uint32_t hl = (v0 & M) + (v4 & M) + 4u * ((v1 & M) + (v3 & M)) + 6u * (v2 & M);
        const uint32_t hh = ((v0 >> 8) & 0xFFu) + ((v4 >> 8) & 0xFFu) + 4u * (((v1 >> 8) & 0xFFu) + ((v3 >> 8) & 0xFFu)) +
                            6u * ((v2 >> 8) & 0xFFu);
        const uint32_t wy = (a == 0) ? 6u : ((a == -1 || a == 1) ? 4u : 1u);
        lo += wy * hl;
        hi += wy * hh;
    }
    lo = ((lo + 0x00800080u) >> 8) & 0x00FF00FFu;        // (B, R)
    hi = ((hi + 128u) >> 8) & 0xFFu;                     // G
    d.p[(size_t)r * d.pitch + x] = lo | (hi << 8) | 0x01000000u;
}

// ------------------------------------------------------------------ change flags of D[l+1] (App. A.2)
// flag(i,j) = any 8-neighbour n with ||D(i,j) - D(n)||^2 >= isr22, for interior pixels; stored in byte3.
__global__ void __launch_bounds__(256) flag_kernel(msg_plane d, int isr22)
{
    int x = blockIdx.x * blockDim.x + threadIdx.x;
    int r = blockIdx.y;
    if (x >= d.w || r >= d.rows) return;
    int y = d.y0 + r;
    uint32_t c = d.p[(size_t)r * d.pitch + x] & 0x00FFFFFFu;
    uint32_t flag = 0;
    if (x >= 1 && x <= d.w - 2 && y >= 1 && y <= d.hfull - 2 && r >= 1 && r <= d.rows - 2) {
#pragma unroll
        for (int dy = -1; dy <= 1; dy++)
#pragma unroll
            for (int dx = -1; dx <= 1; dx++) {
                if (dx == 0 && dy == 0) continue;
                uint32_t n = d.p[(size_t)(r + dy) * d.pitch + (x + dx)] & 0x00FFFFFFu;
                uint32_t e = __vabsdiffu4(c, n);
                if ((int)__dp4a(e, e, 0u) >= isr22) flag = 1;
            }
    }
    // only byte3 changes; neighbours read bytes 0..2 (32-bit accesses are atomic, colour bits unchanged)
    d.p[(size_t)r * d.pitch + x] = c | (flag << 24);
}

// pyrUp taps along one axis for output index o (n = source length): source indices (i0, i1, i2) with weights
// even o: (1, 6, 1) on (i-1, i, i+1); odd o: (0, 4, 4) on (-, i, i+1); s[-1] := s[1] (s[0] if n == 1), s[n] := s[n-1].
__device__ __forceinline__ void up_taps(int o, int n, int& i0, int& i1, int& i2, int& w0, int& w1, int& w2)
{
    const int i = o >> 1;
    const bool odd = o & 1;
    i1 = i;
    i2 = (i + 1 < n) ? i + 1 : n - 1;
    i0 = (i - 1 >= 0) ? i - 1 : (n > 1 ? 1 : 0);
    w0 = odd ? 0 : 1;
    w1 = odd ? 4 : 6;
    w2 = odd ? 4 : 1;
}

// D[l] = pyrUp(D[l+1]); byte3 = dilate3x3(Mraw) where Mraw[2i+1][2j-1] = flag(i,j), 1<=i<=h1-2, 1<=j<=w1-2.
// Channel sums are kept in 16-bit lanes of two registers ((B,R) and (G,-)): the weights sum to 64, so a lane never exceeds
// 64 * 255, and `(v + 32) >> 6` is applied to both lanes at once.
__global__ void __launch_bounds__(256) pyr_up_mask_kernel(msg_plane s /*D[l+1] with flags*/, msg_plane d,
                                                          int* __restrict__ cell_count, int cells_x)
{
    int x = blockIdx.x * blockDim.x + threadIdx.x;
    int r = blockIdx.y;
    if (x >= d.w || r >= d.rows) return;
    int y = d.y0 + r;
    int y0i, y1i, y2i, wy0, wy1, wy2, x0i, x1i, x2i, wx0, wx1, wx2;
    up_taps(y, s.hfull, y0i, y1i, y2i, wy0, wy1, wy2);
    up_taps(x, s.w, x0i, x1i, x2i, wx0, wx1, wx2);
    uint32_t lo = 0, hi = 0;
    const int yi[3] = {y0i, y1i, y2i}, yw[3] = {wy0, wy1, wy2};
#pragma unroll
    for (int a = 0; a < 3; a++) {
        int rr = clampi(yi[a] - s.y0, 0, s.rows - 1);
        const uint32_t* row = s.p + (size_t)rr * s.pitch;
        uint32_t v0 = __ldg(row + x0i), v1 = __ldg(row + x1i), v2 = __ldg(row + x2i);
        // horizontal combination first (weights <= 6, lanes <= 8 * 255), then the vertical weight
        uint32_t hl = (uint32_t)wx0 * (v0 & 0x00FF00FFu) + (uint32_t)wx1 * (v1 & 0x00FF00FFu) + (uint32_t)wx2 * (v2 & 0x00FF00FFu);
        uint32_t hh = (uint32_t)wx0 * ((v0 >> 8) & 0xFFu) + (uint32_t)wx1 * ((v1 >> 8) & 0xFFu) + (uint32_t)wx2 * ((v2 >> 8) & 0xFFu);
        lo += (uint32_t)yw[a] * hl;
        hi += (uint32_t)yw[a] * hh;
    }
    lo = ((lo + 0x00200020u) >> 6) & 0x00FF00FFu;        // (B, R)
    hi = ((hi + 32u) >> 6) & 0xFFu;                      // G
    uint32_t col = lo | (hi << 8);
    // mask = OR of the flags whose raw position (2i+1, 2j-1) lies in the 3x3 neighbourhood of (y, x):
    //   rows: i = y>>1 (raw row 2i+1 in {y, y+1}) and, for even y, i = (y>>1) - 1 (raw row y-1)
    //   cols: j = (x+1)>>1 (raw col 2j-1 in {x, x-1}) and, for even x, j + 1 (raw col x+1)
    uint32_t m = 0;
    const int h1 = s.hfull, w1 = s.w;
    const int ihi = y >> 1, ilo = ihi - 1;
    const int jlo = (x + 1) >> 1, jhi = jlo + 1;
    const bool r_hi = (2 * ihi + 1 < d.hfull) && ihi >= 1 && ihi <= h1 - 2 && (ihi - s.y0) >= 0 && (ihi - s.y0) < s.rows;
    const bool r_lo = !(y & 1) && ilo >= 1 && ilo <= h1 - 2 && (ilo - s.y0) >= 0 && (ilo - s.y0) < s.rows;
    const bool c_lo = jlo >= 1 && jlo <= w1 - 2;
    const bool c_hi = !(x & 1) && (x + 1 < d.w) && jhi >= 1 && jhi <= w1 - 2;
    if (r_hi) {
        const uint32_t* row = s.p + (size_t)(ihi - s.y0) * s.pitch;
        if (c_lo) m |= __ldg(row + jlo) >> 24;
        if (c_hi) m |= __ldg(row + jhi) >> 24;
    }
    if (r_lo) {
        const uint32_t* row = s.p + (size_t)(ilo - s.y0) * s.pitch;
        if (c_lo) m |= __ldg(row + jlo) >> 24;
        if (c_hi) m |= __ldg(row + jhi) >> 24;
    }
    d.p[(size_t)r * d.pitch + x] = col | ((m ? 1u : 0u) << 24);
    if (cell_count) {   // active pixels per 32x32 cell (one atomic per warp): lets the mean-shift kernel run heavy tiles first
        unsigned bal = __ballot_sync(__activemask(), m != 0);
        if ((threadIdx.x & 31) == 0 && bal) atomicAdd(cell_count + (r / 32) * cells_x + (x / 32), __popc(bal));
    }
}

// The same, one 2 x 2 output quad per thread (d.y0 even): the four pixels (2i, 2j) .. (2i+1, 2j+1) read the same 3 x 3 source
// neighbourhood (rows i-1, i, i+1 x columns j-1, j, j+1 with pyrUp's border rule) and the four flags they need are byte 3 of
// words that are loaded anyway, so a quad costs 9 loads and ~100 instructions instead of 4 x (13 loads + ~150 instructions).
// (ncu launch list of the 4K step: 73 us per frame for a 41 MB pass -- instruction-bound, and not hidden by the other streams.)
__global__ void __launch_bounds__(256) pyr_up_mask_quad_kernel(msg_plane s /*D[l+1] with flags*/, msg_plane d,
                                                               int* __restrict__ cell_count, int cells_x)
{
    const int j = blockIdx.x * blockDim.x + threadIdx.x;       // source column; output columns 2j, 2j + 1
    const int r0 = 2 * blockIdx.y;                              // stored output rows r0, r0 + 1
    const int i = (d.y0 >> 1) + blockIdx.y;                     // source row (global, level l+1)
    const int h1 = s.hfull, w1 = s.w;
    const bool inx = 2 * j < d.w;
    int act = 0;
    if (inx && r0 < d.rows) {
        const int jm = (j - 1 >= 0) ? j - 1 : (w1 > 1 ? 1 : 0), jp = (j + 1 < w1) ? j + 1 : w1 - 1;
        const int jc = j < w1 ? j : w1 - 1;
        const int im = (i - 1 >= 0) ? i - 1 : (h1 > 1 ? 1 : 0), ip = (i + 1 < h1) ? i + 1 : h1 - 1;
        const int ic = i < h1 ? i : h1 - 1;
        const int rows3[3] = {im, ic, ip};
        const uint32_t M = 0x00FF00FFu;
        uint32_t El[3], Eh[3], Ol[3], Oh[3], F1[3], F2[3];
#pragma unroll
        for (int a = 0; a < 3; a++) {
            const int rr = clampi(rows3[a] - s.y0, 0, s.rows - 1);
            const uint32_t* row = s.p + (size_t)rr * s.pitch;
            const uint32_t v0 = __ldg(row + jm), v1 = __ldg(row + jc), v2 = __ldg(row + jp);
            const uint32_t a0 = v0 & M, a1 = v1 & M, a2 = v2 & M;
            const uint32_t g0 = (v0 >> 8) & 0xFFu, g1 = (v1 >> 8) & 0xFFu, g2 = (v2 >> 8) & 0xFFu;
            El[a] = a0 + 6u * a1 + a2;  Eh[a] = g0 + 6u * g1 + g2;          // even output column: (1, 6, 1) on (j-1, j, j+1)
            Ol[a] = 4u * (a1 + a2);     Oh[a] = 4u * (g1 + g2);              // odd output column: (4, 4) on (j, j+1)
            F1[a] = v1 >> 24;           F2[a] = v2 >> 24;
        }
        auto fin = [&](uint32_t lo, uint32_t hi) {
            return (((lo + 0x00200020u) >> 6) & M) | ((((hi + 32u) >> 6) & 0xFFu) << 8);
        };
        const uint32_t cEE = fin(El[0] + 6u * El[1] + El[2], Eh[0] + 6u * Eh[1] + Eh[2]);     // (2i, 2j)
        const uint32_t cEO = fin(Ol[0] + 6u * Ol[1] + Ol[2], Oh[0] + 6u * Oh[1] + Oh[2]);     // (2i, 2j+1)
        const uint32_t cOE = fin(4u * (El[1] + El[2]), 4u * (Eh[1] + Eh[2]));                 // (2i+1, 2j)
        const uint32_t cOO = fin(4u * (Ol[1] + Ol[2]), 4u * (Oh[1] + Oh[2]));                 // (2i+1, 2j+1)
        // mask = OR of the flags whose raw position (2 i' + 1, 2 j' - 1) lies in the 3 x 3 neighbourhood of the pixel (the rule of
        // pyr_up_mask_kernel above, specialised to the four parities)
        const bool RH = (2 * i + 1 < d.hfull) && i >= 1 && i <= h1 - 2 && (i - s.y0) >= 0 && (i - s.y0) < s.rows;
        const bool RL = (i - 1) >= 1 && (i - 1) <= h1 - 2 && (i - 1 - s.y0) >= 0 && (i - 1 - s.y0) < s.rows;
        const bool Cj = j >= 1 && j <= w1 - 2, Cj1 = (j + 1) >= 1 && (j + 1) <= w1 - 2;
        const bool CH = Cj1 && (2 * j + 1 < d.w);
        const uint32_t hi_j = (RH && Cj) ? F1[1] : 0u, hi_j1h = (RH && CH) ? F2[1] : 0u, hi_j1 = (RH && Cj1) ? F2[1] : 0u;
        const uint32_t lo_j = (RL && Cj) ? F1[0] : 0u, lo_j1h = (RL && CH) ? F2[0] : 0u, lo_j1 = (RL && Cj1) ? F2[0] : 0u;
        const uint32_t mEE = (hi_j | hi_j1h | lo_j | lo_j1h) ? 1u : 0u;
        const uint32_t mEO = (hi_j1 | lo_j1) ? 1u : 0u;
        const uint32_t mOE = (hi_j | hi_j1h) ? 1u : 0u;
        const uint32_t mOO = hi_j1 ? 1u : 0u;
        const bool two_cols = 2 * j + 1 < d.w, two_rows = r0 + 1 < d.rows;
        uint32_t* o0 = d.p + (size_t)r0 * d.pitch + 2 * j;
        if (two_cols) *reinterpret_cast<uint2*>(o0) = make_uint2(cEE | (mEE << 24), cEO | (mEO << 24));
        else o0[0] = cEE | (mEE << 24);
        act = (int)mEE + (two_cols ? (int)mEO : 0);
        if (two_rows) {
            uint32_t* o1 = o0 + d.pitch;
            if (two_cols) *reinterpret_cast<uint2*>(o1) = make_uint2(cOE | (mOE << 24), cOO | (mOO << 24));
            else o1[0] = cOE | (mOE << 24);
            act += (int)mOE + (two_cols ? (int)mOO : 0);
        }
    }
    if (cell_count) {   // active pixels per 32 x 32 cell: a warp covers 64 output columns = two cells of one cell row
        const int lane = threadIdx.x & 31;
        const int half = __reduce_add_sync(0xffffffffu, lane < 16 ? act : 0);
        const int other = __reduce_add_sync(0xffffffffu, lane < 16 ? 0 : act);
        if (lane == 0 && half) atomicAdd(cell_count + (r0 / 32) * cells_x + (2 * j) / 32, half);
        if (lane == 16 && other) atomicAdd(cell_count + (r0 / 32) * cells_x + (2 * j) / 32, other);
    }
}

// ------------------------------------------------------------------ synthetic image
__device__ __forceinline__ uint64_t splitmix64(uint64_t z)
{
    z += 0x9E3779B97F4A7C15ull;
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    return z ^ (z >> 31);
}
__device__ __forceinline__ uint64_t synth_hash(uint64_t sm, uint32_t a, uint32_t b, uint32_t c)
{
    return splitmix64(sm ^ (((uint64_t)a << 40) | ((uint64_t)b << 16) | (uint64_t)c));
}

__global__ void __launch_bounds__(256) synth_kernel(uint8_t* __restrict__ dst, size_t step, int w, int h, int row0, int rows,
                                                    uint64_t sm)
{
    int x = blockIdx.x * blockDim.x + threadIdx.x;
    int y = row0 + blockIdx.y;      // global row; the buffer holds rows [row0, row0 + rows)
    if (x >= w || (int)blockIdx.y >= rows || y >= h) return;
    int ncx = (w + 63) / 64, ncy = (h + 63) / 64;
    int cx0 = x / 64, cy0 = y / 64;
    long long best = -1;
    uint32_t bc = 0;
    for (int cy = cy0 - 1; cy <= cy0 + 1; cy++)
        for (int cx = cx0 - 1; cx <= cx0 + 1; cx++) {
            if (cx < 0 || cy < 0 || cx >= ncx || cy >= ncy) continue;
            int sx = 64 * cx + (int)(synth_hash(sm, cx, cy, 0) % 64);
            int sy = 64 * cy + (int)(synth_hash(sm, cx, cy, 1) % 64);
            long long ddx = x - sx, ddy = y - sy, dd = ddx * ddx + ddy * ddy;
            if (best < 0 || dd < best) { best = dd; bc = (uint32_t)(synth_hash(sm, cx, cy, 2) & 0xFFFFFFu); }
        }
    uint8_t* p = dst + (size_t)(y - row0) * step + 3 * (size_t)x;
    for (int c = 0; c < 3; c++) {
        int base = (int)((bc >> (8 * c)) & 0xFF);
        int nz = (int)(synth_hash(sm, x, y, 16u + c) % 13) + (int)(synth_hash(sm, x, y, 32u + c) % 13) - 12;
        int v = base + nz;
        p[c] = (uint8_t)(v < 0 ? 0 : (v > 255 ? 255 : v));
    }
}

uint64_t host_splitmix64(uint64_t z)
{
    z += 0x9E3779B97F4A7C15ull;
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    return z ^ (z >> 31);
}

}  // namespace

int k_bgr_to_plane(msg_ctx* ctx, const uint8_t* d_bgr, size_t step, msg_plane dst)
{
    if (dst.w % 4 == 0 && step % 4 == 0 && ((uintptr_t)d_bgr & 3) == 0 && ((uintptr_t)dst.p & 15) == 0 && dst.pitch % 4 == 0) {
        dim3 grid4((dst.w / 4 + 255) / 256, dst.rows);
        bgr_to_plane_v4_kernel<<<grid4, 256, 0, ctx->stream>>>(d_bgr, step, dst.p, dst.w / 4, dst.rows, dst.pitch);
        MSG_LAUNCHED(ctx);
        MSG_CHECK_LAUNCH(ctx);
        return MSG_OK;
    }
    dim3 grid((dst.w + 255) / 256, dst.rows);
    bgr_to_plane_kernel<<<grid, 256, 0, ctx->stream>>>(d_bgr, step, dst.p, dst.w, dst.rows, dst.pitch);
    MSG_LAUNCHED(ctx);
    MSG_CHECK_LAUNCH(ctx);
    return MSG_OK;
}

int k_plane_to_bgr(msg_ctx* ctx, msg_plane src, int row_first, int nrows, uint8_t* d_bgr, size_t step)
{
    if (src.w % 4 == 0 && step % 4 == 0 && ((uintptr_t)d_bgr & 3) == 0 && ((uintptr_t)src.p & 15) == 0 && src.pitch % 4 == 0) {
        dim3 grid4((src.w / 4 + 255) / 256, nrows);
        plane_to_bgr_v4_kernel<<<grid4, 256, 0, ctx->stream>>>(src.p + (size_t)row_first * src.pitch, src.pitch, src.w / 4, nrows,
                                                              d_bgr, step);
        MSG_LAUNCHED(ctx);
        MSG_CHECK_LAUNCH(ctx);
        return MSG_OK;
    }
    dim3 grid((src.w + 255) / 256, nrows);
    plane_to_bgr_kernel<<<grid, 256, 0, ctx->stream>>>(src.p + (size_t)row_first * src.pitch, src.pitch, src.w, nrows,
                                                      d_bgr, step);
    MSG_LAUNCHED(ctx);
    MSG_CHECK_LAUNCH(ctx);
    return MSG_OK;
}

int k_pyr_down(msg_ctx* ctx, msg_plane src, msg_plane dst)
{
    dim3 grid((dst.w + 255) / 256, dst.rows);
    pyr_down_kernel<<<grid, 256, 0, ctx->stream>>>(src, dst);
    MSG_LAUNCHED(ctx);
    MSG_CHECK_LAUNCH(ctx);
    return MSG_OK;
}

int k_pyr_up_mask(msg_ctx* ctx, msg_plane dsrc, msg_plane ddst, int isr22, int* d_cell_count, int cells_x)
{
    dim3 g1((dsrc.w + 255) / 256, dsrc.rows);
    flag_kernel<<<g1, 256, 0, ctx->stream>>>(dsrc, isr22);
    MSG_LAUNCHED(ctx);
    MSG_CHECK_LAUNCH(ctx);
    if ((ddst.y0 & 1) == 0 && ddst.pitch % 2 == 0) {       // 2 x 2 quads need even strip origins (they are multiples of 2^maxLevel)
        dim3 gq(((ddst.w + 1) / 2 + 255) / 256, (ddst.rows + 1) / 2);
        pyr_up_mask_quad_kernel<<<gq, 256, 0, ctx->stream>>>(dsrc, ddst, d_cell_count, cells_x);
        MSG_LAUNCHED(ctx);
        MSG_CHECK_LAUNCH(ctx);
        return MSG_OK;
    }
    dim3 g2((ddst.w + 255) / 256, ddst.rows);
    pyr_up_mask_kernel<<<g2, 256, 0, ctx->stream>>>(dsrc, ddst, d_cell_count, cells_x);
    MSG_LAUNCHED(ctx);
    MSG_CHECK_LAUNCH(ctx);
    return MSG_OK;
}

int k_synth(msg_ctx* ctx, uint8_t* d_bgr, size_t step, int w, int h, int row0, int rows, uint64_t seed)
{
    dim3 grid((w + 255) / 256, rows);
    synth_kernel<<<grid, 256, 0, ctx->stream>>>(d_bgr, step, w, h, row0, rows, host_splitmix64(seed));
    MSG_LAUNCHED(ctx);
    MSG_CHECK_LAUNCH(ctx);
    return MSG_OK;
}
