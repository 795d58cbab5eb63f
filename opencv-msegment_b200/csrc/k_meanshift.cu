// k_meanshift.cu -- K1: the mean-shift filtering kernels (one pyramid level per launch).
//
// Semantics: cv::pyrMeanShiftFiltering's per-level loop, SURVEY.md App. A.2 (bit-exact):
//   window  = [cvRound(x0-sp), cvRound(x0+sp)] x [cvRound(y0-sp), cvRound(y0+sp)] clamped to the image
//   in-range test ||t-c||^2 <= isr2 on 8-bit BGR; integer sums of colour and ABSOLUTE position;
//   new state = cvRound(sum * (1.0/count)) in IEEE double, round-half-even;
//   stop when the centre does not move or |dx|+|dy|+||dc||^2 <= eps, or after maxCount iterations.
//
// Design (B200, INT-ALU bound, DESIGN.md "K1"):
//   * one CTA per 64x32 (large planes) or 32x32 (small planes: shorter tail) pixel tile; tile + (radius + drift) halo staged once in shared memory as
//     packed BGRx words, out-of-image positions hold a sentinel that can never be in range, so the
//     window loops run unclamped and warp-convergent;
//   * work is a queue of (pixel, centre, colour) items held in shared memory and processed in
//     rounds, ONE mean-shift iteration per item per round; converged items retire, the rest are
//     compacted (ballot + one shared atomic per warp) into the next round's queue, so lanes never
//     idle on pixels that converged early (iterations vary 1..maxCount per pixel);
//   * per test: LDS + VABSDIFF4 + IDP4A + ISETP + predicated 16-bit-lane SIMD accumulates;
//   * an item whose next window would leave the staged rectangle is pushed to a global overflow
//     list and finished by the warp-cooperative generic kernel (reads HBM/L2 directly, explicit clamping).
#include <stdlib.h>

#include "msg_internal.h"

namespace {

constexpr int TH = 32;            // tile height (pixels); tile width TW is 64 (256 threads) or 32 (128 threads)
constexpr uint32_t SENTINEL = 0xFF000000u;  // byte3 = 255 vs 1 of real pixels: distance >= 254^2

__device__ __forceinline__ int rnd_f(float v) { return __float2int_rn(v); }  // cvRound(float): half-even

struct iter_result {
    int x1, y1;
    uint32_t c1;
    int count;
};

// One mean-shift iteration epilogue: exact OpenCV arithmetic (double reciprocal, then multiplies).
__device__ __forceinline__ iter_result ms_epilogue(int s0, int s1, int s2, long long sx, long long sy, int count)
{
    iter_result r;
    double icount = __ddiv_rn(1.0, (double)count);
    r.x1 = __double2int_rn(__dmul_rn((double)sx, icount));
    r.y1 = __double2int_rn(__dmul_rn((double)sy, icount));
    int n0 = __double2int_rn(__dmul_rn((double)s0, icount));
    int n1 = __double2int_rn(__dmul_rn((double)s1, icount));
    int n2 = __double2int_rn(__dmul_rn((double)s2, icount));
    r.c1 = (uint32_t)n0 | ((uint32_t)n1 << 8) | ((uint32_t)n2 << 16) | 0x01000000u;
    r.count = count;
    return r;
}

__device__ __forceinline__ bool ms_stop(int x0, int y0, uint32_t c0, const iter_result& r, int ieps)
{
    if (r.x1 == x0 && r.y1 == y0) return true;
    uint32_t e = __vabsdiffu4(c0, r.c1);  // byte3 equal (1) on both sides
    int dc2 = (int)__dp4a(e, e, 0u);
    return abs(r.x1 - x0) + abs(r.y1 - y0) + dc2 <= ieps;
}

// Window scan over the staged tile.  NX > 0: window is NX x NX for every item (integral sp).
// Per test: LDS + VABSDIFF4 + IDP4A + ISETP + 3 predicated IDP4A (B, G, R sums; FMA pipe) + 1 predicated add of the
// immediate (1 << 16 | xx) (hit count and sum of x in one register; ALU pipe) = 8 issue slots, 4 FMA : 3 ALU : 1 LSU.
__device__ __forceinline__ uint32_t dp4a_u(uint32_t a, uint32_t b, uint32_t c)
{
    uint32_t d;
    asm("dp4a.u32.u32 %0, %1, %2, %3;" : "=r"(d) : "r"(a), "r"(b), "r"(c));
    return d;
}

// In-range accumulate, branch free: the packed pixel is zeroed on a miss (1 SEL, ALU pipe) and the three colour sums
// are unconditional IDP4A (FMA pipe); hit count and sum of x share one register via a predicated immediate add.
// Per test: LDS + VABSDIFF4 + IDP4A + ISETP + SEL + 3 IDP4A + 1 add = 9 issue slots, 4 ALU : 4 FMA : 1 LSU.
__device__ __forceinline__ void accumulate_if_hit(uint32_t d2, int isr2, uint32_t t, uint32_t add_cx, uint32_t& a0,
                                                  uint32_t& a1, uint32_t& a2, uint32_t& cx)
{
    uint32_t tm;
    // one predicate, one select, one predicated add -- spelled in PTX so that the add stays predicated
    asm("{\n\t"
        ".reg .pred p;\n\t"
        "setp.le.s32 p, %2, %3;\n\t"
        "selp.u32 %0, %4, 0, p;\n\t"
        "@p add.u32 %1, %1, %5;\n\t"
        "}"
        : "=r"(tm), "+r"(cx)
        : "r"(d2), "r"(isr2), "r"(t), "r"(add_cx));
    a0 = dp4a_u(tm, 0x00000001u, a0);
    a1 = dp4a_u(tm, 0x00000100u, a1);
    a2 = dp4a_u(tm, 0x00010000u, a2);
}

// ACC = 1: IDP4A accumulates (above).  ACC = 0: 16-bit-lane SIMD accumulates (PRMT + predicated adds; 5 ALU : 3 FMA).
template <int NX, int ACC>
__device__ __forceinline__ void window_scan(const uint32_t* __restrict__ base, int swp, int nx, int ny, uint32_t c,
                                            int isr2, int& s0, int& s1, int& s2, int& sxr, int& syr, int& cnt)
{
    uint32_t a0 = 0, a1 = 0, a2 = 0;
    sxr = syr = cnt = 0;
    if (NX > 0) { nx = NX; ny = NX; }
    for (int yy = 0; yy < ny; ++yy) {
        const uint32_t* row = base + yy * swp;
        if (ACC == 1) {
            uint32_t cx = 0;   // (hits << 16) | sum of xx over the hits of this row (nx <= 241: both fields fit)
            if (NX > 0) {
#pragma unroll
                for (int xx = 0; xx < (NX > 0 ? NX : 1); ++xx) {
                    uint32_t t = row[xx];
                    uint32_t e = __vabsdiffu4(t, c);
                    accumulate_if_hit(dp4a_u(e, e, 0u), isr2, t, 0x10000u + (uint32_t)xx, a0, a1, a2, cx);
                }
            } else {
                for (int xx = 0; xx < nx; ++xx) {
                    uint32_t t = row[xx];
                    uint32_t e = __vabsdiffu4(t, c);
                    accumulate_if_hit(dp4a_u(e, e, 0u), isr2, t, 0x10000u + (uint32_t)xx, a0, a1, a2, cx);
                }
            }
            int rc = (int)(cx >> 16);
            cnt += rc;
            syr += yy * rc;
            sxr += (int)(cx & 0xFFFFu);
        } else {
            uint32_t lo = 0, hi = 0;
            int sx = 0;
            if (NX > 0 && NX <= 41) {
                // Issue-slot and pipe balance: a hit costs PRMT (B,0,R,0) + three plain adds -- `lo` (16-bit lanes B, R),
                // `st` (the whole packed word: B + G<<8 + R<<16 + flag<<24 with flag = 1, no carry out of 32 bits for
                // rows of <= 41 pixels) and `cx` (count << 16 | sum of xx).  G is recovered at the row flush from
                // st - B - (R << 16) - (count << 24).  Plain adds can issue on either integer pipe, so the 8 slots per
                // test split 4 : 4 between the ALU pipe (VABSDIFF4, ISETP, PRMT, one add) and the FMA pipe (IDP4A,
                // two adds as IMAD) instead of 5 : 3 with two PRMTs.
                uint32_t st = 0, cx = 0;
#pragma unroll
                for (int xx = 0; xx < (NX > 0 ? NX : 1); ++xx) {
                    uint32_t t = row[xx];
                    uint32_t e = __vabsdiffu4(t, c);
                    const uint32_t x = __byte_perm(t, 0u, 0x4240);     // (B, 0, R, 0), unconditional
                    // one predicate, three predicated adds -- spelled in PTX so that they stay predicated adds
                    asm("{\n\t"
                        ".reg .pred p;\n\t"
                        "setp.le.s32 p, %3, %4;\n\t"
                        "@p add.u32 %0, %0, %5;\n\t"
                        "@p add.u32 %1, %1, %6;\n\t"
                        "@p add.u32 %2, %2, %7;\n\t"
                        "}"
                        : "+r"(lo), "+r"(st), "+r"(cx)
                        : "r"(dp4a_u(e, e, 0u)), "r"(isr2), "r"(x), "r"(t), "r"(0x10000u + (uint32_t)xx));
                }
                const uint32_t rcn = cx >> 16;
                sx = (int)(cx & 0xFFFFu);
                const uint32_t g = (st - (lo & 0xFFFFu) - (lo & 0xFFFF0000u) - (rcn << 24)) >> 8;
                hi = g | (rcn << 16);
            } else if (NX > 0) {
#pragma unroll
                for (int xx = 0; xx < (NX > 0 ? NX : 1); ++xx) {
                    uint32_t t = row[xx];
                    uint32_t e = __vabsdiffu4(t, c);
                    if ((int)__dp4a(e, e, 0u) <= isr2) {
                        lo += __byte_perm(t, 0u, 0x4240);  // (B, 0, R, 0)
                        hi += __byte_perm(t, 0u, 0x4341);  // (G, 0, 1, 0)
                        sx += xx;
                    }
                }
            } else {
                for (int xx = 0; xx < nx; ++xx) {
                    uint32_t t = row[xx];
                    uint32_t e = __vabsdiffu4(t, c);
                    if ((int)__dp4a(e, e, 0u) <= isr2) {
                        lo += __byte_perm(t, 0u, 0x4240);
                        hi += __byte_perm(t, 0u, 0x4341);
                        sx += xx;
                    }
                }
            }
            // per-row flush of the 16-bit lanes (row length <= 257 keeps 255*n < 65536)
            int rc = (int)(hi >> 16);
            a0 += lo & 0xFFFFu;
            a2 += lo >> 16;
            a1 += hi & 0xFFFFu;
            cnt += rc;
            syr += yy * rc;
            sxr += sx;
        }
    }
    s0 = (int)a0; s1 = (int)a1; s2 = (int)a2;
}

struct tile_geom {
    int halo;     // radius + drift allowance
    int sw, sh;   // staged width / height
    int swp;      // staged row pitch (odd, to spread banks between rows)
    int tiles_x;
    const int* order;   // tile processing order (heaviest first) or nullptr = raster order
    int use_tma;  // interior tiles are staged with cp.async.bulk (requires halo, sw, swp multiples of 4 pixels)
};

template <int NX, int TW, int ACC>
__global__ void __launch_bounds__(TW * 4, 3) meanshift_tile_kernel(msg_plane S, msg_plane D, msg_ms_params prm, tile_geom g,
                                                                msg_ovf_item* __restrict__ ovf, int* __restrict__ ovf_count,
                                                                unsigned long long* __restrict__ active_count,
                                                                unsigned long long* __restrict__ work)
{
    constexpr int NT = TW * 4;        // threads per CTA
    constexpr int NPIX = TW * TH;     // <= 2048 -> 11 bits
    extern __shared__ __align__(16) uint32_t smem[];
    uint32_t* stage = smem;                                  // g.sh * g.swp
    uint2* q0 = reinterpret_cast<uint2*>(smem + ((g.sh * g.swp + 1) & ~1));
    uint2* q1 = q0 + NPIX;
    __shared__ int qn[2];

    const int tid = threadIdx.x;
    const int lane = tid & 31;
    const int tile = g.order ? g.order[blockIdx.x] : (int)blockIdx.x;
    const int tile_x = tile % g.tiles_x, tile_y = tile / g.tiles_x;
    const int tx0 = tile_x * TW;                 // global x of the tile origin
    const int ty0 = S.y0 + tile_y * TH;          // global y of the tile origin
    const int ox = tx0 - g.halo, oy = ty0 - g.halo;  // global coords of staged (0,0)

    if (tid < 2) qn[tid] = 0;

    // ---- stage tile + halo.  Interior tiles (staged rectangle entirely inside the stored plane, i.e. no sentinel needed):
    //      one TMA bulk copy (cp.async.bulk, global -> shared, SASS UBLKCP) per staged row, issued by lane 0 of each warp and tracked by an
    //      mbarrier transaction count; the rows are 16-byte aligned by construction (halo and pitch are multiples of 4 pixels).
    //      Border tiles: plain loads, with the out-of-image sentinel.
    const bool interior = g.use_tma && ox >= 0 && ox + g.sw <= S.w && oy >= 0 && oy + g.sh <= S.hfull &&
                          oy - S.y0 >= 0 && oy - S.y0 + g.sh <= S.rows;
    if (interior) {
        __shared__ __align__(8) unsigned long long mbar;
        const uint32_t mbar_a = (uint32_t)__cvta_generic_to_shared(&mbar);
        if (tid == 0) {
            asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(mbar_a));
            asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        }
        __syncthreads();
        if (tid == 0) {
            asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(mbar_a), "r"((uint32_t)(g.sh * g.sw * 4)) : "memory");
        }
        if (lane == 0) {   // one elected lane per warp issues that warp's share of the row copies
            const uint32_t row_bytes = (uint32_t)g.sw * 4u;
            for (int sy = tid / 32; sy < g.sh; sy += NT / 32) {
                const uint32_t* src = S.p + (size_t)(oy - S.y0 + sy) * S.pitch + ox;
                const uint32_t dst = (uint32_t)__cvta_generic_to_shared(stage + sy * g.swp);
                asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                             ::"r"(dst), "l"(src), "r"(row_bytes), "r"(mbar_a) : "memory");
            }
        }
        uint32_t done = 0;
        while (!done) {
            asm volatile("{\n\t.reg .pred p;\n\t"
                         "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], 0;\n\t"
                         "selp.u32 %0, 1, 0, p;\n\t}"
                         : "=r"(done) : "r"(mbar_a) : "memory");
        }
    } else {
        for (int sy = tid / 32; sy < g.sh; sy += NT / 32) {
            int gy = oy + sy;
            int r = gy - S.y0;
            bool row_ok = (gy >= 0) && (gy < S.hfull) && (r >= 0) && (r < S.rows);
            const uint32_t* srow = S.p + (size_t)(row_ok ? r : 0) * S.pitch;
            uint32_t* drow = stage + sy * g.swp;
            for (int sx = lane; sx < g.sw; sx += 32) {
                int gx = ox + sx;
                uint32_t v = SENTINEL;
                if (row_ok && gx >= 0 && gx < S.w) v = __ldg(srow + gx);
                drow[sx] = v;
            }
        }
    }
    __syncthreads();

    // ---- initial queue: active pixels of the tile.  A warp queues a block of RPW consecutive rows of one 32-pixel-wide
    //      column of the tile (one slot allocation per block), so 32 consecutive queue items come from few rows of one 32-column
    //      band: pixels of one row fall in distinct shared-memory banks, and the row pitch residue spreads consecutive rows
    //      (fewer bank conflicts among the 32 windows a warp reads at a time than with half-rows in arrival order).
    {
        constexpr int NW = NT / 32, RPW = TH / NW;             // 8 warps x 4 rows (TW = 64) or 4 warps x 8 rows (TW = 32)
        const int wq = tid / 32;
        for (int half = 0; half < TW / 32; ++half) {
            const int tx = half * 32 + lane;
            const int gx = tx0 + tx;
            unsigned actbits = 0;
            int total = 0;
#pragma unroll
            for (int k = 0; k < RPW; ++k) {
                const int ty = wq * RPW + k;
                const int gy = ty0 + ty, r = gy - S.y0;
                bool act = (gx < S.w) && (gy < S.hfull) && (r < S.rows);
                if (act && prm.use_mask) act = (D.p[(size_t)r * D.pitch + gx] >> 24) != 0;
                const unsigned m = __ballot_sync(0xffffffffu, act);
                if (act) actbits |= 1u << k;
                total += __popc(m);
            }
            int pos = 0;
            if (lane == 0 && total) pos = atomicAdd(&qn[0], total);
            pos = __shfl_sync(0xffffffffu, pos, 0);
            int run = 0;
#pragma unroll
            for (int k = 0; k < RPW; ++k) {
                const bool act = (actbits >> k) & 1u;
                const unsigned m = __ballot_sync(0xffffffffu, act);
                if (act) {
                    const int ty = wq * RPW + k;
                    const int p = ty * TW + tx;
                    const int slot = pos + run + __popc(m & ((1u << lane) - 1));
                    uint32_t w0 = (uint32_t)p | ((uint32_t)(tx + g.halo) << 12) | ((uint32_t)(ty + g.halo) << 21);
                    q0[slot] = make_uint2(w0, stage[(ty + g.halo) * g.swp + tx + g.halo]);
                }
                run += __popc(m);
            }
        }
    }
    __syncthreads();
    if (tid == 0 && active_count) atomicAdd(active_count, (unsigned long long)qn[0]);

    const float sp = prm.sp;
    const int R = prm.radius;
    unsigned wk_tests = 0, wk_hits = 0;   // algorithmic work of this thread (profiling only)
    uint2* qc = q0;
    uint2* qnx = q1;
    for (int it = 0; it < prm.max_count; ++it) {
        const int cur = it & 1;
        const int n = qn[cur];
        if (n == 0) break;
        __syncthreads();               // everyone has read qn[cur] / finished the previous round
        if (tid == 0) qn[cur ^ 1] = 0;
        __syncthreads();
        const bool last_round = (it == prm.max_count - 1);
        for (int ibase = (tid / 32) * 32; ibase < n; ibase += NT) {
            const int i = ibase + lane;
            const bool valid = i < n;
            uint2 item = valid ? qc[i] : make_uint2(0u, 0u);
            const int pix = item.x & 0xFFF;
            const int xs = (item.x >> 12) & 0x1FF, ys = (item.x >> 21) & 0x1FF;
            const uint32_t c = item.y;
            const int x0 = ox + xs, y0 = oy + ys;
            bool requeue = false, to_ovf = false;
            iter_result res;
            res.x1 = x0; res.y1 = y0; res.c1 = c; res.count = 0;
            if (valid) {
                int minx = rnd_f((float)x0 - sp), maxx = rnd_f((float)x0 + sp);
                int miny = rnd_f((float)y0 - sp), maxy = rnd_f((float)y0 + sp);
                const uint32_t* base = stage + (miny - oy) * g.swp + (minx - ox);
                int s0, s1, s2, sxr, syr, cnt;
                window_scan<NX, ACC>(base, g.swp, maxx - minx + 1, maxy - miny + 1, c, prm.isr2, s0, s1, s2, sxr, syr, cnt);
                if (work) {   // uniform branch; counts what the CPU oracle counts: clamped window area and hits
                    int cx = min(maxx, S.w - 1) - max(minx, 0) + 1, cy = min(maxy, S.hfull - 1) - max(miny, 0) + 1;
                    wk_tests += (unsigned)(cx * cy);
                    wk_hits += (unsigned)cnt;
                }
                bool fin = true;
                if (cnt > 0) {
                    long long sx = (long long)sxr + (long long)cnt * minx;   // absolute-coordinate sums
                    long long sy = (long long)syr + (long long)cnt * miny;
                    res = ms_epilogue(s0, s1, s2, sx, sy, cnt);
                    fin = ms_stop(x0, y0, c, res, prm.ieps) || last_round;
                }
                if (fin) {
                    int ty = pix / TW, tx = pix % TW;
                    D.p[(size_t)(ty0 - S.y0 + ty) * D.pitch + (tx0 + tx)] = res.c1;
                } else {
                    int nxs = res.x1 - ox, nys = res.y1 - oy;
                    bool inside = (nxs - R >= 0) && (nxs + R < g.sw) && (nys - R >= 0) && (nys + R < g.sh);
                    requeue = inside;
                    to_ovf = !inside;
                }
            }
            unsigned m = __ballot_sync(0xffffffffu, requeue);
            int pos = 0;
            if (lane == 0 && m) pos = atomicAdd(&qn[cur ^ 1], __popc(m));
            pos = __shfl_sync(0xffffffffu, pos, 0);
            if (requeue) {
                int slot = pos + __popc(m & ((1u << lane) - 1));
                uint32_t w0 = (uint32_t)pix | ((uint32_t)(res.x1 - ox) << 12) | ((uint32_t)(res.y1 - oy) << 21);
                qnx[slot] = make_uint2(w0, res.c1);
            }
            if (to_ovf) {
                int slot = atomicAdd(ovf_count, 1);
                int ty = pix / TW, tx = pix % TW;
                msg_ovf_item o;
                o.pix = (uint32_t)((size_t)(ty0 - S.y0 + ty) * D.pitch + (tx0 + tx));
                o.x0 = (int16_t)res.x1;
                o.y0rel = (int16_t)(res.y1 - S.y0);
                o.c = res.c1;
                o.iter = (uint32_t)(it + 1);
                ovf[slot] = o;
            }
        }
        __syncthreads();
        uint2* t = qc; qc = qnx; qnx = t;
    }
    if (work) {
#pragma unroll
        for (int o = 16; o; o >>= 1) {
            wk_tests += __shfl_xor_sync(0xffffffffu, wk_tests, o);
            wk_hits += __shfl_xor_sync(0xffffffffu, wk_hits, o);
        }
        if (lane == 0) {
            atomicAdd(work, (unsigned long long)wk_tests);
            atomicAdd(work + 1, (unsigned long long)wk_hits);
        }
    }
}

// Generic path, one WARP per item: lanes stride over the columns of the (explicitly clamped) window, rows are looped,
// partial sums are combined with shuffles and every lane evaluates the (identical) epilogue, so control flow stays
// warp-uniform.  Reads the plane through L1/L2.  Either finishes overflow items (items != nullptr) or processes every
// (active) pixel of the plane (items == nullptr: parameter corners the tile kernel does not serve).
__global__ void __launch_bounds__(128) meanshift_generic_kernel(msg_plane S, msg_plane D, msg_ms_params prm,
                                                                const msg_ovf_item* __restrict__ items,
                                                                const int* __restrict__ n_items,
                                                                unsigned long long* __restrict__ active_count,
                                                                unsigned long long* __restrict__ ovf_total,
                                                                unsigned long long* __restrict__ work)
{
    const long long total = items ? (long long)*n_items : (long long)S.rows * S.w;
    if (items && ovf_total && blockIdx.x == 0 && threadIdx.x == 0 && total > 0)
        atomicAdd(ovf_total, (unsigned long long)total);
    const float sp = prm.sp;
    const int lane = threadIdx.x & 31;
    const long long warp0 = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const long long nwarps = ((long long)gridDim.x * blockDim.x) >> 5;
    unsigned long long wk_tests = 0, wk_hits = 0;
    for (long long i = warp0; i < total; i += nwarps) {
        int x0, y0, it0;
        uint32_t c;
        size_t out;
        if (items) {
            msg_ovf_item o = items[i];
            x0 = o.x0; y0 = S.y0 + o.y0rel; c = o.c; it0 = (int)o.iter; out = o.pix;
        } else {
            int r = (int)(i / S.w), x = (int)(i % S.w);
            out = (size_t)r * D.pitch + x;
            if (prm.use_mask && (D.p[out] >> 24) == 0) continue;      // warp-uniform
            x0 = x; y0 = S.y0 + r; c = S.p[(size_t)r * S.pitch + x]; it0 = 0;
            if (active_count && lane == 0) atomicAdd(active_count, 1ull);
        }
        for (int it = it0; it < prm.max_count; ++it) {
            int minx = max(rnd_f((float)x0 - sp), 0), maxx = min(rnd_f((float)x0 + sp), S.w - 1);
            int miny = max(rnd_f((float)y0 - sp), 0), maxy = min(rnd_f((float)y0 + sp), S.hfull - 1);
            // stored-row guard (strip planes): rows outside the stored range do not exist here
            miny = max(miny, S.y0); maxy = min(maxy, S.y0 + S.rows - 1);
            int s0 = 0, s1 = 0, s2 = 0, cnt = 0;
            long long sx = 0, sy = 0;
            for (int y = miny; y <= maxy; ++y) {
                const uint32_t* row = S.p + (size_t)(y - S.y0) * S.pitch;
                int rc = 0, rsx = 0;
                for (int x = minx + lane; x <= maxx; x += 32) {
                    uint32_t t = __ldg(row + x);
                    uint32_t e = __vabsdiffu4(t, c);
                    if ((long long)__dp4a(e, e, 0u) <= (long long)prm.isr2) {
                        s0 += (int)(t & 0xFF); s1 += (int)((t >> 8) & 0xFF); s2 += (int)((t >> 16) & 0xFF);
                        rsx += x; rc++;
                    }
                }
                cnt += rc; sx += rsx; sy += (long long)y * rc;
            }
#pragma unroll
            for (int o = 16; o; o >>= 1) {
                s0 += __shfl_xor_sync(0xffffffffu, s0, o);
                s1 += __shfl_xor_sync(0xffffffffu, s1, o);
                s2 += __shfl_xor_sync(0xffffffffu, s2, o);
                cnt += __shfl_xor_sync(0xffffffffu, cnt, o);
                sx += __shfl_xor_sync(0xffffffffu, sx, o);
                sy += __shfl_xor_sync(0xffffffffu, sy, o);
            }
            if (work && lane == 0 && maxx >= minx && maxy >= miny) {
                wk_tests += (unsigned long long)(maxx - minx + 1) * (unsigned long long)(maxy - miny + 1);
                wk_hits += (unsigned long long)cnt;
            }
            if (cnt == 0) break;
            iter_result res = ms_epilogue(s0, s1, s2, sx, sy, cnt);
            bool stop = ms_stop(x0, y0, c, res, prm.ieps);
            x0 = res.x1; y0 = res.y1; c = res.c1;
            if (stop) break;
        }
        if (lane == 0) D.p[out] = c;
    }
    if (work && (wk_tests | wk_hits)) {
        atomicAdd(work + 2, wk_tests);
        atomicAdd(work + 3, wk_hits);
    }
}

// Heavy-first tile order.  CTAs are dispatched roughly in blockIdx order, so a launch whose tiles differ a lot in work (the
// masked levels: 0..2048 active pixels per tile) ends with a long tail if heavy tiles start late.  One small CTA sorts the
// tiles by active-pixel count, descending (counting sort over 2049 keys), from the 32x32-cell counts the pyrUp+mask kernel
// left; the tile kernel then maps blockIdx through this order.
__global__ void __launch_bounds__(256) tile_order_kernel(const int* __restrict__ cells, int cells_x, int cells_y, int tiles_x,
                                                         int tiles_y, int tw_cells, int* __restrict__ order)
{
    __shared__ int hist[2050];
    const int nt = tiles_x * tiles_y;
    for (int i = threadIdx.x; i < 2050; i += 256) hist[i] = 0;
    __syncthreads();
    auto weight = [&](int t) {
        int tx = t % tiles_x, ty = t / tiles_x, c = 0;
        for (int k = 0; k < tw_cells; k++) {
            int cx = tx * tw_cells + k;
            if (cx < cells_x && ty < cells_y) c += cells[ty * cells_x + cx];
        }
        return min(c, 2048);
    };
    for (int t = threadIdx.x; t < nt; t += 256) atomicAdd(&hist[2048 - weight(t)], 1);
    __syncthreads();
    {   // exclusive prefix over the 2049 keys: 9 consecutive keys per thread, then a block scan of the 256 partial sums
        __shared__ int wsum[8];
        const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
        const int k0 = threadIdx.x * 9;
        int local[9], sum = 0;
#pragma unroll
        for (int k = 0; k < 9; k++) { local[k] = (k0 + k <= 2048) ? hist[k0 + k] : 0; sum += local[k]; }
        int incl = sum;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            int v = __shfl_up_sync(0xffffffffu, incl, o);
            if (lane >= o) incl += v;
        }
        if (lane == 31) wsum[warp] = incl;
        __syncthreads();
        int woff = 0;
        for (int k = 0; k < warp; k++) woff += wsum[k];
        int run = woff + incl - sum;
#pragma unroll
        for (int k = 0; k < 9; k++) {
            if (k0 + k <= 2048) hist[k0 + k] = run;
            run += local[k];
        }
    }
    __syncthreads();
    for (int t = threadIdx.x; t < nt; t += 256) order[atomicAdd(&hist[2048 - weight(t)], 1)] = t;
}

template <int NX, int TW, int ACC>
cudaError_t launch_tile(msg_ctx* ctx, msg_plane S, msg_plane D, const msg_ms_params& prm, const tile_geom& g, int tiles,
                        size_t smem, int* ovf_count, unsigned long long* active, unsigned long long* work)
{
    // attribute is per function AND per device: issued once per (kernel, size) and context (msg_func_smem caches it)
    if (msg_func_smem(ctx, (const void*)meanshift_tile_kernel<NX, TW, ACC>, smem) != MSG_OK) return cudaErrorInvalidValue;
    meanshift_tile_kernel<NX, TW, ACC><<<tiles, TW * 4, smem, ctx->stream>>>(S, D, prm, g, ctx->d_ovf, ovf_count, active, work);
    return cudaGetLastError();
}

template <int TW, int ACC>
cudaError_t dispatch_tile(int nx, msg_ctx* ctx, msg_plane S, msg_plane D, const msg_ms_params& prm, const tile_geom& g,
                          int tiles, size_t smem, int* ovf_count, unsigned long long* active, unsigned long long* work)
{
    switch (nx) {   // integral sp -> every window is (2 sp + 1)^2: fully unrolled instantiations
        case 3: return launch_tile<3, TW, ACC>(ctx, S, D, prm, g, tiles, smem, ovf_count, active, work);
        case 5: return launch_tile<5, TW, ACC>(ctx, S, D, prm, g, tiles, smem, ovf_count, active, work);
        case 7: return launch_tile<7, TW, ACC>(ctx, S, D, prm, g, tiles, smem, ovf_count, active, work);
        case 11: return launch_tile<11, TW, ACC>(ctx, S, D, prm, g, tiles, smem, ovf_count, active, work);
        case 21: return launch_tile<21, TW, ACC>(ctx, S, D, prm, g, tiles, smem, ovf_count, active, work);
        case 41: return launch_tile<41, TW, ACC>(ctx, S, D, prm, g, tiles, smem, ovf_count, active, work);
        default: return launch_tile<0, TW, ACC>(ctx, S, D, prm, g, tiles, smem, ovf_count, active, work);
    }
}

}  // namespace

// d_counters layout: [0] overflow count of the current level, [2..3] active items (u64), [4..5] overflow total (u64)
int k_meanshift_level(msg_ctx* ctx, msg_plane S, msg_plane D, const msg_ms_params& prm, int level)
{
    // profiling: d_work[level] = {tile tests, tile hits, generic tests, generic hits} (u64 each)
    unsigned long long* work = ctx->profiling ? ctx->d_work + 4 * level : nullptr;
    int* ovf_count = ctx->d_counters;
    unsigned long long* active = reinterpret_cast<unsigned long long*>(ctx->d_counters + 2);
    unsigned long long* ovf_total = reinterpret_cast<unsigned long long*>(ctx->d_counters + 4);

    // tile width: 64 (256 threads) when the plane gives enough 64x32 tiles for several waves, else 32 (128 threads):
    // small planes would otherwise run 1-2 waves of unequal tiles and idle SMs in the tail
    const long long tiles64 = (long long)((S.w + 63) / 64) * ((S.rows + TH - 1) / TH);
    // Measured on B200 (tools/k1_matrix.py, profiles/r01_k1_matrix.md): 64-wide tiles win whenever they fill the SMs once;
    // IDP4A accumulates win for narrow windows (row flush amortised over few tests), 16-bit lanes for wide ones.
    int TWsel = tiles64 >= (long long)ctx->sm_count ? 64 : 32;
    int acc = (prm.sp == (float)(int)prm.sp && 2 * (int)prm.sp + 1 >= 21) ? 0 : 1;
    // tuning overrides (experiments only): MSG_TILE_W = 32 | 64, MSG_ACC = 0 | 1
    if (ctx->tune.tile_w == 32 || ctx->tune.tile_w == 64) TWsel = ctx->tune.tile_w;
    if (ctx->tune.acc >= 0) acc = ctx->tune.acc ? 1 : 0;

    // tile path limits: sentinel distance 254^2 must exceed isr2; staged tile must fit shared memory
    const int R = prm.radius;
    int drift = R + (R + 3) / 4;  // ~1.25 * radius (covers > 99 % of measured drift)
    if (drift < 4) drift = 4;
    tile_geom g;
    size_t smem = 0;
    bool tile_ok = prm.isr2 < 254 * 254 && R <= 120 && S.rows <= 32767 && S.w <= 32767;
    if (tile_ok) {
        for (;; drift = drift * 3 / 4) {
            g.halo = (R + drift + 2) & ~3;           // nearest multiple of 4 pixels: staged rows start 16-byte aligned
            if (g.halo < R + 1) g.halo += 4;
            g.sw = TWsel + 2 * g.halo;
            g.sh = TH + 2 * g.halo;
            // row pitch: a multiple of 4 words (TMA bulk copies need 16-byte aligned destinations) whose residue mod 32 banks
            // is +-8: the queue holds blocks of 4..8 consecutive rows of a 32-column band, so runs of active pixels on
            // neighbouring rows land 8 banks apart (measured at 1080p, sp = 10: residue 8 / 24: 0.342 ms, 16: 0.359, 0: 0.425)
            g.swp = g.sw;
            while (g.swp % 32 != 8 && g.swp % 32 != 24) g.swp += 4;
            if (ctx->tune.pitch_res >= 0) {     // experiment switch: force the pitch residue mod 32 banks
                const int want = ctx->tune.pitch_res & 28;
                g.swp = g.sw;
                while (g.swp % 32 != want) g.swp += 4;
            }
            smem = ((size_t)((g.sh * g.swp + 1) & ~1) + 4 * (size_t)(TWsel * TH)) * sizeof(uint32_t);
            if (smem <= (size_t)ctx->max_smem_optin - 1024 && g.sw < 512 && g.sh < 512) break;
            if (drift == 0) { tile_ok = false; break; }
        }
    }
    if (!tile_ok) {
        int blocks = ctx->sm_count * 16;
        meanshift_generic_kernel<<<blocks, 128, 0, ctx->stream>>>(S, D, prm, nullptr, nullptr, active, nullptr, work);
        MSG_LAUNCHED(ctx);
        MSG_CHECK_LAUNCH(ctx);
        return MSG_OK;
    }

    size_t need = (size_t)S.rows * S.w;
    MSG_TRY(msg_reserve(ctx, (void**)&ctx->d_ovf, &ctx->d_ovf_cap, need * sizeof(msg_ovf_item)));
    MSG_CUDA(ctx, cudaMemsetAsync(ovf_count, 0, sizeof(int), ctx->stream));
    g.tiles_x = (S.w + TWsel - 1) / TWsel;
    g.order = nullptr;
    g.use_tma = ctx->tune.use_tma;   // A/B switch (experiments)
    int tiles_y = (S.rows + TH - 1) / TH;
    int tiles = g.tiles_x * tiles_y;
    if (prm.use_mask && ctx->d_cells && !ctx->tune.no_order) {
        int cells_x = (S.w + 31) / 32, cells_y = (S.rows + 31) / 32;
        int* order = ctx->d_cells + (size_t)cells_x * cells_y;       // second half of the cell buffer
        tile_order_kernel<<<1, 256, 0, ctx->stream>>>(ctx->d_cells, cells_x, cells_y, g.tiles_x, tiles_y, TWsel / 32, order);
        MSG_LAUNCHED(ctx);
        g.order = order;
    }

    int nx = 0;
    if (prm.sp == (float)(int)prm.sp) nx = 2 * (int)prm.sp + 1;
    if (ctx->profiling) msg_prof_begin(ctx, level);
    cudaError_t e;
    if (TWsel == 64)
        e = acc ? dispatch_tile<64, 1>(nx, ctx, S, D, prm, g, tiles, smem, ovf_count, active, work)
                : dispatch_tile<64, 0>(nx, ctx, S, D, prm, g, tiles, smem, ovf_count, active, work);
    else
        e = acc ? dispatch_tile<32, 1>(nx, ctx, S, D, prm, g, tiles, smem, ovf_count, active, work)
                : dispatch_tile<32, 0>(nx, ctx, S, D, prm, g, tiles, smem, ovf_count, active, work);
    MSG_LAUNCHED(ctx);
    MSG_CUDA(ctx, e);
    if (ctx->profiling) msg_prof_end(ctx, level, 0);

    // finish the items that left their tile (count is read on the device: no host sync)
    meanshift_generic_kernel<<<ctx->sm_count * 4, 128, 0, ctx->stream>>>(S, D, prm, ctx->d_ovf, ovf_count, nullptr, ovf_total, work);
    if (ctx->profiling) msg_prof_end(ctx, level, 1);
    MSG_LAUNCHED(ctx);
    MSG_CHECK_LAUNCH(ctx);
    return MSG_OK;
}
