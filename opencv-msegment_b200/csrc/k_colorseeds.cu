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
#include <type_traits>

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
// The transform is a sequential recurrence: pixel (y, x) of a pass needs (y, x-1), (y-1, x-2 .. x+2) and (y-2, x-1), (y-2, x+1)
// of the same pass, and the float additions do not associate, so the order of the published two-pass algorithm is kept and
// only its wavefront parallelism is used.  Lane G of one CTA owns the P consecutive columns [G*P, (G+1)*P) and handles row r
// at step s = 2r + G: then its left neighbour has just finished row r (the carry of the in-row recurrence) and its right
// neighbour row r-1.  A lane keeps its last two rows in registers; the seven halo values it needs come from the neighbour
// lanes by shuffles (from the neighbour warps through a double-buffered shared-memory slot), one __syncthreads per step.
// Steps per pass: 2 * height + lanes in use.  The initial values of a lane's coming rows are loaded four steps ahead.
// Arithmetic: the backward pass is the plain float recurrence; the forward pass keeps the running value unrounded inside
// aligned groups of four columns (dt_wave_pass), which is what the IPP-backed cv2 build computes.
constexpr int DT_THREADS = 256, DT_WARPS = DT_THREADS / 32, DT_PMAXMAX = 32;
constexpr float DT_A = 1.0f, DT_B = 1.4f, DT_C = 2.1969f;

template <int P>
__device__ __forceinline__ float dt_pick(const float (&row)[P], int k, float l0, float l1, float r0, float r1)
{
    // k is a compile-time constant after unrolling: -2, -1 -> left halo; P, P+1 -> right halo
    return k < 0 ? (k == -1 ? l1 : l0) : (k >= P ? (k == P ? r0 : r1) : row[k < 0 ? 0 : (k >= P ? P - 1 : k)]);
}

struct dt_shared {
    float4 right[2][DT_WARPS];    // lane 31 of each warp: A[P-1], B[P-2], B[P-1], Clast
    float4 left[2][DT_WARPS];     // lane 0 of each warp: A[0], A[1], B[0]
};

// One pass.  DIR = +1: forward (top -> bottom, left -> right; initial value 0 on zero pixels of src, "infinite" elsewhere),
// DIR = -1: backward (bottom -> top, right -> left; initial value = the forward result in dist).  Coordinates are in pass
// order (mirrored for the backward pass).  Returns the lane's maximum of the values it stored.
template <int V> struct dt_vec;
template <> struct dt_vec<1> { typedef float type; };
template <> struct dt_vec<2> { typedef float2 type; };
template <> struct dt_vec<4> { typedef float4 type; };

// V = floats per memory access of a lane's chunk (rows must then be a multiple of V floats wide: aligned chunks).  Every lane
// works on a different row, so a warp-wide access touches up to 32 cache lines: wide accesses keep the load/store unit,
// which is what bounds this kernel, four times less busy.
template <int P, int DIR, int V>
__device__ __forceinline__ float dt_wave_pass(float* __restrict__ dist, int w, int h, dt_shared& sh)
{
    typedef typename dt_vec<V>::type vec_t;
    const int G = threadIdx.x, lane = G & 31, warp = G >> 5;
    const int m0 = G * P, cnt = max(0, min(w - m0, P));
    const int NL = (w + P - 1) / P;            // lanes in use
    const int nsteps = 2 * h + NL - 2;
    // lowest image column of the chunk (the chunk is [m0, m0 + P) in pass order, mirrored for the backward pass)
    const int xlo = DIR > 0 ? m0 : w - m0 - P;
    const bool full = cnt == P;                // only the last lane in use can hold a partial chunk
    // A = the lane's last finished row, B = the one before, Clast = last pixel of the one before that (pass order)
    float A[P], B[P], Clast = FLT_MAX, vmax = 0.f;
    // initial values of the lane's coming rows, loaded four steps (two rows) ahead.  The register set is chosen by the step
    // number modulo 4 (uniform over the warp), so a set is not touched between its load and its use.
    float q0[P], q1[P], q2[P], q3[P];
#pragma unroll
    for (int j = 0; j < P; j++) { A[j] = FLT_MAX; B[j] = FLT_MAX; q0[j] = q1[j] = q2[j] = q3[j] = FLT_MAX; }
    // element e (ascending image column xlo + e) of the chunk is pixel j = e (forward) or P - 1 - e (backward) in pass order
    auto load_row = [&](int it, float (&pre)[P]) {            // only called with cnt > 0
        const int y = DIR > 0 ? it : h - 1 - it;
        const float* p = dist + (size_t)y * w + xlo;
        if (full) {
#pragma unroll
            for (int e = 0; e < P; e += V) {
                vec_t v = __ldcg((const vec_t*)(p + e));
                const float* f = (const float*)&v;
#pragma unroll
                for (int i = 0; i < V; i++) pre[DIR > 0 ? e + i : P - 1 - (e + i)] = f[i];
            }
        } else {
#pragma unroll
            for (int j = 0; j < P; j++) {
                const int jj = min(j, cnt - 1);               // clamped: always inside the row
                float v = __ldcg(p + (DIR > 0 ? jj : P - 1 - jj));
                pre[j] = j < cnt ? v : FLT_MAX;
            }
        }
    };
    auto store_row = [&](int it, const float (&val)[P]) {
        const int y = DIR > 0 ? it : h - 1 - it;
        float* p = dist + (size_t)y * w + xlo;
        if (full) {
#pragma unroll
            for (int e = 0; e < P; e += V) {
                vec_t v;
                float* f = (float*)&v;
#pragma unroll
                for (int i = 0; i < V; i++) f[i] = val[DIR > 0 ? e + i : P - 1 - (e + i)];
                *(vec_t*)(p + e) = v;
            }
        } else {
#pragma unroll
            for (int j = 0; j < P; j++)
                if (j < cnt) p[DIR > 0 ? j : P - 1 - j] = val[j];
        }
    };
    if (cnt > 0) {                                            // row 0 is used at step G, row 1 at step G + 2
        if ((G & 3) == 0) { load_row(0, q0); if (h > 1) load_row(1, q2); }
        if ((G & 3) == 1) { load_row(0, q1); if (h > 1) load_row(1, q3); }
        if ((G & 3) == 2) { load_row(0, q2); if (h > 1) load_row(1, q0); }
        if ((G & 3) == 3) { load_row(0, q3); if (h > 1) load_row(1, q1); }
    }
    if (G < 2 * DT_WARPS) {
        const float4 inf = make_float4(FLT_MAX, FLT_MAX, FLT_MAX, FLT_MAX);
        (&sh.right[0][0])[G] = inf;
        (&sh.left[0][0])[G] = inf;
    }
    __syncthreads();
    auto step = [&](int s, float (&pre)[P]) {
        const int d = s - G;
        const bool active = cnt > 0 && d >= 0 && !(d & 1) && (d >> 1) < h;
        // halo: left lane finished row r one step ago (A = row r, B = r-1, Clast = r-2), right lane row r-1 (A = r-1, B = r-2)
        float Lc = __shfl_up_sync(0xffffffffu, A[P - 1], 1);
        float Lb0 = __shfl_up_sync(0xffffffffu, B[P - 2], 1);
        float Lb1 = __shfl_up_sync(0xffffffffu, B[P - 1], 1);
        float Lcl = __shfl_up_sync(0xffffffffu, Clast, 1);
        float Ra0 = __shfl_down_sync(0xffffffffu, A[0], 1);
        float Ra1 = __shfl_down_sync(0xffffffffu, A[1], 1);
        float Rb0 = __shfl_down_sync(0xffffffffu, B[0], 1);
        if (lane == 0) {
            float4 v = warp > 0 ? sh.right[(s + 1) & 1][warp - 1] : make_float4(FLT_MAX, FLT_MAX, FLT_MAX, FLT_MAX);
            Lc = v.x; Lb0 = v.y; Lb1 = v.z; Lcl = v.w;
        }
        if (lane == 31) {
            float4 v = warp + 1 < DT_WARPS ? sh.left[(s + 1) & 1][warp + 1] : make_float4(FLT_MAX, FLT_MAX, FLT_MAX, FLT_MAX);
            Ra0 = v.x; Ra1 = v.y; Rb0 = v.z;
        }
        if (active) {
            const int it = d >> 1;
            float t[P];
#pragma unroll
            for (int j = 0; j < P; j++) t[j] = pre[j];
            if (it + 2 < h) load_row(it + 2, pre);            // consumed four steps later, from the same register set
            if constexpr (DIR > 0) {
                // Forward pass.  Inside aligned groups of four columns (x < lim) the running value of the in-row recurrence is
                // NOT rounded to float between pixels: the candidates are exact float + float sums, the running value carries
                // that exact sum through up to three additions of 1, and only the stored pixel is rounded.  A group starts
                // (x % 4 == 0) from the stored, rounded left neighbour; the last columns (x >= lim) round at every pixel.
                // Exact sums fit a double.  (This is what the IPP-backed cv2 build computes: see the oracle.)
                double tu[P];
#pragma unroll
                for (int j = 0; j < P; j++) {
                    double a0 = (double)dt_pick<P>(B, j - 1, FLT_MAX, Lcl, Rb0, FLT_MAX) + (double)DT_C;
                    double a1 = (double)dt_pick<P>(B, j + 1, FLT_MAX, Lcl, Rb0, FLT_MAX) + (double)DT_C;
                    double a2 = (double)dt_pick<P>(A, j - 2, Lb0, Lb1, Ra0, Ra1) + (double)DT_C;
                    double a3 = (double)dt_pick<P>(A, j - 1, Lb0, Lb1, Ra0, Ra1) + (double)DT_B;
                    double a4 = (double)A[j] + (double)DT_A;
                    double a5 = (double)dt_pick<P>(A, j + 1, Lb0, Lb1, Ra0, Ra1) + (double)DT_B;
                    double a6 = (double)dt_pick<P>(A, j + 2, Lb0, Lb1, Ra0, Ra1) + (double)DT_C;
                    tu[j] = fmin(fmin(fmin(a0, a1), fmin(a2, a3)), fmin(fmin(a4, a5), fmin(a6, (double)t[j])));
                }
                const int lim = ((w - 2) / 4) * 4;
                double runv = (double)Lc;
#pragma unroll
                for (int j = 0; j < P; j++) {
                    const int x = m0 + j;
                    const double rin = ((x & 3) == 0 || x >= lim) ? (double)__double2float_rn(runv) : runv;
                    runv = fmin(tu[j], rin + 1.0);
                    t[j] = j < cnt ? __double2float_rn(runv) : FLT_MAX;     // outside the image: "infinite"
                }
            } else {
                // Backward pass: plain float.  Candidates that do not depend on the in-row recurrence are independent min trees
                // (a pixel whose value is <= 1 keeps it: every candidate is >= 1; FLT_MAX + metric rounds back to FLT_MAX)
#pragma unroll
                for (int j = 0; j < P; j++) {
                    float a0 = __fadd_rn(dt_pick<P>(B, j - 1, FLT_MAX, Lcl, Rb0, FLT_MAX), DT_C);
                    float a1 = __fadd_rn(dt_pick<P>(B, j + 1, FLT_MAX, Lcl, Rb0, FLT_MAX), DT_C);
                    float a2 = __fadd_rn(dt_pick<P>(A, j - 2, Lb0, Lb1, Ra0, Ra1), DT_C);
                    float a3 = __fadd_rn(dt_pick<P>(A, j - 1, Lb0, Lb1, Ra0, Ra1), DT_B);
                    float a4 = __fadd_rn(A[j], DT_A);
                    float a5 = __fadd_rn(dt_pick<P>(A, j + 1, Lb0, Lb1, Ra0, Ra1), DT_B);
                    float a6 = __fadd_rn(dt_pick<P>(A, j + 2, Lb0, Lb1, Ra0, Ra1), DT_C);
                    t[j] = fminf(fminf(fminf(a0, a1), fminf(a2, a3)), fminf(fminf(a4, a5), fminf(a6, t[j])));
                }
                // the recurrence along the row: two dependent operations per pixel
                float run = Lc;
#pragma unroll
                for (int j = 0; j < P; j++) {
                    run = fminf(t[j], __fadd_rn(run, DT_A));
                    t[j] = j < cnt ? run : FLT_MAX;           // outside the image
                    vmax = fmaxf(vmax, j < cnt ? run : 0.f);
                }
            }
            store_row(it, t);
            Clast = B[P - 1];
#pragma unroll
            for (int j = 0; j < P; j++) { B[j] = A[j]; A[j] = t[j]; }
        }
        if (lane == 31) sh.right[s & 1][warp] = make_float4(A[P - 1], B[P - 2], B[P - 1], Clast);
        if (lane == 0) sh.left[s & 1][warp] = make_float4(A[0], A[1], B[0], 0.f);
        __syncthreads();
    };
#pragma unroll 1
    for (int s = 0; s < nsteps; s += 4) {                     // nsteps is uniform over the CTA: every thread meets every barrier
        step(s, q0);
        if (s + 1 < nsteps) step(s + 1, q1);
        if (s + 2 < nsteps) step(s + 2, q2);
        if (s + 3 < nsteps) step(s + 3, q3);
    }
    __syncthreads();                                          // the pass' values (global) are visible to the whole CTA
    return vmax;
}

template <int P, int V>
__global__ void __launch_bounds__(DT_THREADS, 1) dt_wave_kernel(float* __restrict__ dist, int w, int h, float* d_max)
{
    __shared__ dt_shared sh;
    __shared__ float s_max[DT_WARPS];
    dt_wave_pass<P, 1, V>(dist, w, h, sh);
    float vmax = dt_wave_pass<P, -1, V>(dist, w, h, sh);
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
    for (int o = 16; o; o >>= 1) vmax = fmaxf(vmax, __shfl_xor_sync(0xffffffffu, vmax, o));
    if (lane == 0) s_max[warp] = vmax;
    __syncthreads();
    if (threadIdx.x == 0) {
        float v = s_max[0];
        for (int i = 1; i < DT_WARPS; i++) v = fmaxf(v, s_max[i]);
        *d_max = v;
    }
}

// ---------------------------------------------------------------- the same wavefront, second form (default)
// What bounded dt_wave_kernel (ncu, profiles/r01_dt_kernel_history.md): one SM's issue slots and its FP64 pipe, not the chain
// latency -- half of the lanes idle at every step (lane G works when s - G is even) and the forward pass spent ~25 double
// operations and ~10 conversions per pixel.  Here
//   * every thread owns TWO adjacent chunks ("virtual lanes" 2t and 2t+1) and works at every step: the even one at even
//     steps, the odd one at odd steps (virtual lane v handles row r at step s = 2r + v, exactly the schedule above), so the
//     halo of a step comes half from the thread's own registers and half from ONE neighbour thread (3-4 shuffles, not 7);
//   * candidates that share a metric are reduced in float first -- min(v1 + c, v2 + c) = min(v1, v2) + c holds for exact sums
//     and for rounded ones (rounding is monotone) -- so a pixel costs three additions and two minima instead of seven and
//     seven;
//   * float -> double is exact widening, done with two integer operations instead of the quarter-rate F2F (0 widens to
//     2^-127, which every use adds to a metric >= 1 where it vanishes in the rounding; zero pixels are selected explicitly);
//   * the rule "round at every pixel" of the last columns (x >= lim) is compiled separately and taken by the one or two
//     virtual lanes it concerns.
// Same arithmetic, same results (tests/test_color_seeds.py), 8192 columns at most.
constexpr int DT2_MAX_WARPS = 16;          // CTAs of 256 (rows up to 2048 pixels) or 512 threads, chunks of 4 / 8 columns

struct dt2_shared {
    float4 right[DT2_MAX_WARPS];      // lane 31's ODD chunk after an odd step: A[P-1], B[P-2], B[P-1], Clast (read at the next even step)
    float4 left[DT2_MAX_WARPS];       // lane 0's EVEN chunk after an even step: A[0], A[1], B[0] (read at the next odd step)
};

__device__ __forceinline__ double dt_widen(float f)        // exact for normal non-negative floats; 0 -> 2^-127
{
    const unsigned b = __float_as_uint(f);
    return __hiloint2double((int)((b >> 3) + 0x38000000u), (int)(b << 29));
}
__device__ __forceinline__ double dt_dmin(double a, double b) { return a < b ? a : b; }

// Row prefetch: the initial values of a chunk's row are fetched FOUR STEPS ahead with cp.async (LDGSTS) into the thread's own
// shared-memory slot of the step's set (s mod 4) and read back with one LDS when the row is processed.  (Prefetching into
// registers, as dt_wave_kernel does, made ptxas stage the load in scratch registers and move it "home" right behind the load:
// every step then waited a full L2 round trip -- 47 % of the stall samples of the first capture, profiles/r02_dt_wave2_ncu.md.)
// One commit group per step, so cp.async.wait_group 3 = "the group of four steps ago has landed".  The prefetch is
// unconditional with clamped coordinates (a lane that has not started yet fetches exactly its rows 0 and 1 during the four
// steps before its first, and anything fetched past the last row or chunk is never used), so there is no prologue.
template <int BYTES>
__device__ __forceinline__ void dt_cp_async(void* smem, const void* gmem)
{
    const unsigned sa = (unsigned)__cvta_generic_to_shared(smem);
    asm volatile("cp.async.ca.shared.global [%0], [%1], %2;" ::"r"(sa), "l"(gmem), "n"(BYTES) : "memory");
}

template <int NT, int P, int DIR, int V>
__device__ __forceinline__ float dt2_pass(float* __restrict__ dist, int w, int h, dt2_shared& sh, float* s_pre)
{
    typedef typename dt_vec<V>::type vec_t;
    const int T = threadIdx.x, lane = T & 31, warp = T >> 5;
    const int NVL = (w + P - 1) / P;                 // virtual lanes in use
    const int nsteps = 2 * h + NVL - 2;
    const int lim = ((w - 2) / 4) * 4;               // forward pass: columns >= lim round at every pixel
    float A[2][P], B[2][P], Cl[2] = {FLT_MAX, FLT_MAX}, vmax = 0.f;
#pragma unroll
    for (int j = 0; j < P; j++) A[0][j] = A[1][j] = B[0][j] = B[1][j] = FLT_MAX;
    // slot of (set, element group e / V) of this thread: consecutive threads are V floats apart -> conflict-free
    auto slot = [&](int set, int eg) { return s_pre + ((size_t)(set * (P / V) + eg) * NT + T) * V; };
    // element e (ascending image column xlo + e) of the chunk is pixel j = e (forward) or P - 1 - e (backward) in pass order
    auto prefetch_row = [&](int par, int it, int set) {
        const int vl = min(2 * T + par, NVL - 1);                        // clamped: always a chunk of the image
        const int m0 = vl * P, cnt = min(w - m0, P);
        const int y = DIR > 0 ? min(max(it, 0), h - 1) : h - 1 - min(max(it, 0), h - 1);
        const int xlo = DIR > 0 ? m0 : w - m0 - P;
        const float* p = dist + (size_t)y * w + xlo;
#pragma unroll
        for (int e = 0; e < P; e += V) {
            // groups outside the row (partial last chunk) re-fetch a valid group; their pixels are masked at use
            const int ec = DIR > 0 ? min(e, (cnt - 1) / V * V) : max(e, (P - cnt) / V * V);
            dt_cp_async<4 * V>(slot(set, e / V), p + ec);
        }
    };
    auto read_row = [&](int set, int cnt, float (&pre)[P]) {
#pragma unroll
        for (int e = 0; e < P; e += V) {
            const vec_t v = *(const vec_t*)slot(set, e / V);
            const float* f = (const float*)&v;
#pragma unroll
            for (int i = 0; i < V; i++) {
                const int j = DIR > 0 ? e + i : P - 1 - (e + i);
                pre[j] = j < cnt ? f[i] : FLT_MAX;
            }
        }
    };
    auto store_row = [&](int m0, int cnt, int it, const float (&val)[P]) {
        const int y = DIR > 0 ? it : h - 1 - it;
        const int xlo = DIR > 0 ? m0 : w - m0 - P;
        float* p = dist + (size_t)y * w + xlo;
        if (cnt == P) {
#pragma unroll
            for (int e = 0; e < P; e += V) {
                vec_t v;
                float* f = (float*)&v;
#pragma unroll
                for (int i = 0; i < V; i++) f[i] = val[DIR > 0 ? e + i : P - 1 - (e + i)];
                *(vec_t*)(p + e) = v;
            }
        } else {
#pragma unroll
            for (int j = 0; j < P; j++)
                if (j < cnt) p[DIR > 0 ? j : P - 1 - j] = val[j];
        }
    };
    // what steps -4 .. -1 would have prefetched: virtual lane v uses row 0 at step v, row 1 at step v + 2
#pragma unroll
    for (int s = -4; s < 0; s++) {
        const int par = s & 1;
        prefetch_row(par, ((s - (2 * T + par)) >> 1) + 2, s & 3);
        asm volatile("cp.async.commit_group;" ::: "memory");
    }
    if (T < NT / 32) {
        const float4 inf = make_float4(FLT_MAX, FLT_MAX, FLT_MAX, FLT_MAX);
        sh.right[T] = inf;
        sh.left[T] = inf;
    }
    __syncthreads();

    // one row of one chunk.  Ac / Bc: the chunk's last two rows; halo values as in dt_wave_pass
    auto row = [&](auto tail_tag, int m0, int cnt, float (&t)[P], const float (&Ac)[P], const float (&Bc)[P], float Lc, float Lb0,
                   float Lb1, float Lcl, float Ra0, float Ra1, float Rb0) {
        constexpr bool TAIL = decltype(tail_tag)::value;
        // candidates of one metric reduced in float
        float mC[P], mB[P];
#pragma unroll
        for (int j = 0; j < P; j++) {
            const float b0 = dt_pick<P>(Bc, j - 1, FLT_MAX, Lcl, Rb0, FLT_MAX), b1 = dt_pick<P>(Bc, j + 1, FLT_MAX, Lcl, Rb0, FLT_MAX);
            const float a0 = dt_pick<P>(Ac, j - 2, Lb0, Lb1, Ra0, Ra1), a1 = dt_pick<P>(Ac, j + 2, Lb0, Lb1, Ra0, Ra1);
            mC[j] = fminf(fminf(b0, b1), fminf(a0, a1));
            mB[j] = fminf(dt_pick<P>(Ac, j - 1, Lb0, Lb1, Ra0, Ra1), dt_pick<P>(Ac, j + 1, Lb0, Lb1, Ra0, Ra1));
        }
        if constexpr (DIR > 0) {
            double runv = dt_widen(Lc);
#pragma unroll
            for (int j = 0; j < P; j++) {
                const double cand = dt_dmin(dt_dmin(dt_widen(mC[j]) + (double)DT_C, dt_widen(mB[j]) + (double)DT_B),
                                            dt_widen(Ac[j]) + (double)DT_A);
                const double tu = t[j] == 0.f ? 0.0 : cand;                   // initial value: 0 on zero pixels, "infinite" elsewhere
                bool round_in = (j & 3) == 0 && j > 0;                        // j == 0: Lc is a stored (rounded) pixel already
                if (TAIL) round_in = round_in || (j > 0 && m0 + j >= lim);
                const double rin = round_in ? dt_widen(__double2float_rn(runv)) : runv;
                runv = dt_dmin(tu, rin + 1.0);
                t[j] = j < cnt ? __double2float_rn(runv) : FLT_MAX;           // outside the image: "infinite"
            }
        } else {
            float run = Lc;
#pragma unroll
            for (int j = 0; j < P; j++) {
                const float cand = fminf(fminf(__fadd_rn(mC[j], DT_C), __fadd_rn(mB[j], DT_B)), __fadd_rn(Ac[j], DT_A));
                run = fminf(fminf(t[j], cand), __fadd_rn(run, DT_A));
                t[j] = j < cnt ? run : FLT_MAX;
                vmax = fmaxf(vmax, j < cnt ? run : 0.f);
            }
        }
    };

    auto step = [&](auto par_tag, int s) {
        constexpr int PAR = decltype(par_tag)::value;
        const int vl = 2 * T + PAR, d = s - vl, it = d >> 1;
        const int m0 = vl * P, cnt = max(0, min(w - m0, P));
        const bool active = cnt > 0 && d >= 0 && it < h;
        float Lc, Lb0, Lb1, Lcl, Ra0, Ra1, Rb0;
        if (PAR == 0) {
            // left = the odd chunk of thread T-1 (row r done one step ago), right = my own odd chunk (row r-1)
            Lc = __shfl_up_sync(0xffffffffu, A[1][P - 1], 1);
            Lb0 = __shfl_up_sync(0xffffffffu, B[1][P - 2], 1);
            Lb1 = __shfl_up_sync(0xffffffffu, B[1][P - 1], 1);
            Lcl = __shfl_up_sync(0xffffffffu, Cl[1], 1);
            if (lane == 0) {
                const float4 v = warp > 0 ? sh.right[warp - 1] : make_float4(FLT_MAX, FLT_MAX, FLT_MAX, FLT_MAX);
                Lc = v.x; Lb0 = v.y; Lb1 = v.z; Lcl = v.w;
            }
            Ra0 = A[1][0]; Ra1 = A[1][1]; Rb0 = B[1][0];
        } else {
            // left = my own even chunk (row r done one step ago), right = the even chunk of thread T+1 (row r-1)
            Lc = A[0][P - 1]; Lb0 = B[0][P - 2]; Lb1 = B[0][P - 1]; Lcl = Cl[0];
            Ra0 = __shfl_down_sync(0xffffffffu, A[0][0], 1);
            Ra1 = __shfl_down_sync(0xffffffffu, A[0][1], 1);
            Rb0 = __shfl_down_sync(0xffffffffu, B[0][0], 1);
            if (lane == 31) {
                const float4 v = warp + 1 < NT / 32 ? sh.left[warp + 1] : make_float4(FLT_MAX, FLT_MAX, FLT_MAX, FLT_MAX);
                Ra0 = v.x; Ra1 = v.y; Rb0 = v.z;
            }
        }
        asm volatile("cp.async.wait_group 3;" ::: "memory");      // the prefetch of four steps ago (this step's set) has landed
        if (active) {
            float t[P];
            read_row(s & 3, cnt, t);
            if (DIR > 0 && m0 + P > lim)
                row(std::true_type(), m0, cnt, t, A[PAR], B[PAR], Lc, Lb0, Lb1, Lcl, Ra0, Ra1, Rb0);
            else
                row(std::false_type(), m0, cnt, t, A[PAR], B[PAR], Lc, Lb0, Lb1, Lcl, Ra0, Ra1, Rb0);
            store_row(m0, cnt, it, t);
            Cl[PAR] = B[PAR][P - 1];
#pragma unroll
            for (int j = 0; j < P; j++) { B[PAR][j] = A[PAR][j]; A[PAR][j] = t[j]; }
        }
        prefetch_row(PAR, it + 2, s & 3);                          // consumed at step s + 4 from the same slot
        asm volatile("cp.async.commit_group;" ::: "memory");
        if (PAR == 1 && lane == 31) sh.right[warp] = make_float4(A[1][P - 1], B[1][P - 2], B[1][P - 1], Cl[1]);
        if (PAR == 0 && lane == 0) sh.left[warp] = make_float4(A[0][0], A[0][1], B[0][0], 0.f);
        __syncthreads();
    };
    typedef std::integral_constant<int, 0> even_t;
    typedef std::integral_constant<int, 1> odd_t;
#pragma unroll 1
    for (int s = 0; s < nsteps; s += 4) {                     // nsteps is uniform over the CTA: every thread meets every barrier
        step(even_t(), s);
        if (s + 1 < nsteps) step(odd_t(), s + 1);
        if (s + 2 < nsteps) step(even_t(), s + 2);
        if (s + 3 < nsteps) step(odd_t(), s + 3);
    }
    asm volatile("cp.async.wait_group 0;" ::: "memory");      // no copy of this pass is still in flight when the next one reuses the slots
    __syncthreads();
    return vmax;
}

template <int NT, int P, int V>
__global__ void __launch_bounds__(NT, 1) dt_wave2_kernel(float* __restrict__ dist, int w, int h, float* d_max)
{
    __shared__ dt2_shared sh;
    __shared__ float s_max[DT2_MAX_WARPS];
    extern __shared__ __align__(16) float s_pre[];            // 4 sets x P floats per thread
    dt2_pass<NT, P, 1, V>(dist, w, h, sh, s_pre);
    float vmax = dt2_pass<NT, P, -1, V>(dist, w, h, sh, s_pre);
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
    for (int o = 16; o; o >>= 1) vmax = fmaxf(vmax, __shfl_xor_sync(0xffffffffu, vmax, o));
    if (lane == 0) s_max[warp] = vmax;
    __syncthreads();
    if (threadIdx.x == 0) {
        float v = s_max[0];
        for (int i = 1; i < NT / 32; i++) v = fmaxf(v, s_max[i]);
        *d_max = v;
    }
}

template <int NT, int P>
int dt2_launch(msg_ctx* ctx, cudaStream_t st, float* d_dist, int w, int h, float* d_max)
{
    const size_t smem = (size_t)4 * P * NT * sizeof(float);                // 16 / 32 / 64 KB
    if (w % 4 == 0 && ((uintptr_t)d_dist & 15) == 0) {
        MSG_TRY(msg_func_smem(ctx, (const void*)dt_wave2_kernel<NT, P, 4>, smem));
        dt_wave2_kernel<NT, P, 4><<<1, NT, smem, st>>>(d_dist, w, h, d_max);
    } else {
        MSG_TRY(msg_func_smem(ctx, (const void*)dt_wave2_kernel<NT, P, 1>, smem));
        dt_wave2_kernel<NT, P, 1><<<1, NT, smem, st>>>(d_dist, w, h, d_max);
    }
    return MSG_OK;
}

// initial values of the forward pass: 0 on the zero pixels of the source, "infinite" elsewhere
__global__ void __launch_bounds__(256) dt_init_kernel(const uint8_t* __restrict__ src, size_t sstep, float* __restrict__ dist, int w)
{
    int x = blockIdx.x * 256 + threadIdx.x, y = blockIdx.y;
    if (x < w) dist[(size_t)y * w + x] = src[(size_t)y * sstep + x] ? FLT_MAX : 0.f;
}

template <int P>
void dt_launch(cudaStream_t st, float* d_dist, int w, int h, float* d_max)
{
    constexpr int V = P < 4 ? P : 4;
    if (w % V == 0 && ((uintptr_t)d_dist & 15) == 0) dt_wave_kernel<P, V><<<1, DT_THREADS, 0, st>>>(d_dist, w, h, d_max);
    else dt_wave_kernel<P, 1><<<1, DT_THREADS, 0, st>>>(d_dist, w, h, d_max);
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

int k_distance_transform_max_width(msg_ctx* ctx)
{
    return ctx->tune.dt_fixed ? k_distance_transform_fixed_max_dim() : DT_THREADS * DT_PMAXMAX;
}

// d_dist: dense w*h floats; d_max (device float, may not be NULL): maximum of the result
int k_distance_transform(msg_ctx* ctx, const uint8_t* d_src, size_t sstep, float* d_dist, int w, int h, float* d_max)
{
    if (ctx->tune.dt_fixed) return k_distance_transform_fixed(ctx, d_src, sstep, d_dist, w, h, d_max);
    if (w > k_distance_transform_max_width(ctx))
        return msg_fail(ctx, MSG_EINVAL, "distanceTransform supports rows up to %d pixels", k_distance_transform_max_width(ctx));
    int need = (w + DT_THREADS - 1) / DT_THREADS;
    cudaStream_t st = ctx->stream;
    dim3 grid((w + 255) / 256, h);
    dt_init_kernel<<<grid, 256, 0, st>>>(d_src, sstep, d_dist, w);
    MSG_LAUNCHED(ctx);
    if (!ctx->tune.dt_legacy) {
        // threads x 2 chunks x P columns (a chunk is a whole number of 4-column groups); short chunks = short steps
        if (w <= 2048) MSG_TRY((dt2_launch<256, 4>(ctx, st, d_dist, w, h, d_max)));
        else if (w <= 4096) MSG_TRY((dt2_launch<512, 4>(ctx, st, d_dist, w, h, d_max)));
        else MSG_TRY((dt2_launch<512, 8>(ctx, st, d_dist, w, h, d_max)));
    } else if (need <= 4) dt_launch<4>(st, d_dist, w, h, d_max);
    else if (need <= 8) dt_launch<8>(st, d_dist, w, h, d_max);
    else if (need <= 16) dt_launch<16>(st, d_dist, w, h, d_max);
    else dt_launch<32>(st, d_dist, w, h, d_max);
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
        MSG_TRY(msg_func_smem(ctx, (const void*)bilateral_kernel<1>, smem));
        bilateral_kernel<1><<<grid, 256, smem, ctx->stream>>>(d_src, sstep, d_dst, dstep, w, h, radius, maxk, d_space_w,
                                                             (const short2*)d_space_ofs, d_color_w);
    } else {
        MSG_TRY(msg_func_smem(ctx, (const void*)bilateral_kernel<3>, smem));
        bilateral_kernel<3><<<grid, 256, smem, ctx->stream>>>(d_src, sstep, d_dst, dstep, w, h, radius, maxk, d_space_w,
                                                             (const short2*)d_space_ofs, d_color_w);
    }
    MSG_LAUNCHED(ctx);
    MSG_CHECK_LAUNCH(ctx);
    return MSG_OK;
}
