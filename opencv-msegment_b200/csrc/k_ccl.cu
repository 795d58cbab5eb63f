// k_ccl.cu -- K2a: connected-component labelling by union-find in HBM, plus canonical relabelling.
//
//   colour predicate, 4-connectivity  == OpenCV floodFill loop with loDiff = upDiff = d, floating range
//                                        (samples/cpp/meanshift_segmentation.cpp)
//   binary predicate, 4/8-connectivity == Imgproc.connectedComponents (PictureService.java:441-442)
//
// Passes (all HBM-bound streaming kernels, one thread per pixel, warps along rows):
//   1 rows    : one CTA per row; warp __ballot of "connected to left" bits per 32-pixel chunk, runs crossing chunk
//               borders resolved by a prefix-max over the chunks -> every pixel points at the exact start of its row run
//   2 merge   : union(run, run above) with atomicMin on roots and path halving; redundant unions inside an overlap of
//               two runs are skipped (only the first column of an overlap unites)
//   3 flatten + count : label = find(label) (root = smallest linear index of the component), roots flagged and ranked
//               inside 4096-pixel chunks in the same pass
//   4 scan of the chunk totals, 5 apply: label = chunk offset + local rank of the root + 1 (raster order of first pixel)
#include "msg_internal.h"

namespace {

constexpr int CCL_THREADS = 256;
constexpr int SCAN_CHUNK = 4096;   // pixels per block in the rank scan (256 threads x 16)

__device__ __forceinline__ bool color_close(uint32_t a, uint32_t b, int d)
{
    uint32_t e = __vabsdiffu4(a & 0x00FFFFFFu, b & 0x00FFFFFFu);
    return (int)(e & 0xFF) <= d && (int)((e >> 8) & 0xFF) <= d && (int)(e >> 16) <= d;
}

__device__ __forceinline__ int uf_find(const int32_t* L, int a)
{
    int p = __ldcg(L + a);
    while (p != a) { a = p; p = __ldcg(L + a); }
    return a;
}

// find with path halving: every visited node is re-pointed at its grandparent.  Racing writers only ever store
// ancestors (parents decrease monotonically towards the root), so the forest stays valid.
__device__ __forceinline__ int uf_find_halve(int32_t* L, int a)
{
    int p = __ldcg(L + a);
    while (p != a) {
        int g = __ldcg(L + p);
        if (g != p) L[a] = g;
        a = p; p = g;
    }
    return a;
}

__device__ __forceinline__ void uf_union(int32_t* L, int a, int b)
{
    for (;;) {
        a = uf_find_halve(L, a);
        b = uf_find_halve(L, b);
        if (a == b) return;
        if (a < b) { int t = a; a = b; b = t; }   // a > b: link the larger root under the smaller
        int old = atomicMin(L + a, b);
        if (old == a) return;                     // a was still a root
        a = old;                                  // somebody re-linked a meanwhile: keep merging
    }
}

// ---------------------------------------------------------------- pass 1: row runs
// PRED: 0 = colour (plane u32, pitch in pixels), 1 = binary mask (u8, step in bytes)
// One CTA per image row.  Phase 1: per 32-pixel chunk a warp __ballot of "connected to the left" bits (kept in shared
// memory).  Phase 2: one warp resolves runs that cross chunk borders with a prefix-max over the chunks ("start of the run
// that reaches the end of chunk c", undefined when the whole chunk is one open run).  Phase 3: every pixel is pointed at the
// exact start of its row run, so no pointer chains exist along rows and the merge pass only unites vertically.
constexpr int CCL_MAX_CHUNKS = 1024;   // rows up to 32768 pixels

template <int PRED>
__global__ void __launch_bounds__(CCL_THREADS) ccl_rows_kernel(const void* __restrict__ img, size_t pitch, int w, int h,
                                                               int d, int32_t* __restrict__ L)
{
    __shared__ unsigned s_bits[CCL_MAX_CHUNKS];    // connected-to-left bits per chunk
    __shared__ unsigned s_fg[CCL_MAX_CHUNKS];      // foreground bits per chunk (binary predicate)
    __shared__ int s_open[CCL_MAX_CHUNKS];         // global x of the run start that enters chunk c from the left
    const int y = blockIdx.x;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int nchunks = (w + 31) / 32;
    for (int c = warp; c < nchunks; c += CCL_THREADS / 32) {
        int x = c * 32 + lane;
        bool in = x < w, fg = in, cl = false;
        if (PRED == 0) {
            const uint32_t* row = (const uint32_t*)img + (size_t)y * pitch;
            if (in && x > 0) cl = color_close(__ldg(row + x), __ldg(row + x - 1), d);
        } else {
            const uint8_t* row = (const uint8_t*)img + (size_t)y * pitch;
            fg = in && row[x] != 0;
            if (fg && x > 0) cl = row[x - 1] != 0;
        }
        unsigned bits = __ballot_sync(0xffffffffu, cl);
        unsigned fgb = __ballot_sync(0xffffffffu, fg);
        if (lane == 0) { s_bits[c] = bits; s_fg[c] = fgb; }
    }
    __syncthreads();
    if (warp == 0) {
        // E[c] = start x of the run containing the last pixel of chunk c, or -1 if that run is open to the left
        int carry = -1;
        for (int base = 0; base < nchunks; base += 32) {
            int c = base + lane;
            int e = -1;
            if (c < nchunks) {
                unsigned bits = s_bits[c];
                if (bits != 0xffffffffu) e = c * 32 + (31 - __clz(~bits));     // highest lane whose bit is clear
            }
            int incl = e;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                int v = __shfl_up_sync(0xffffffffu, incl, o);
                if (lane >= o) incl = max(incl, v);
            }
            int excl = __shfl_up_sync(0xffffffffu, incl, 1);
            excl = lane == 0 ? carry : max(excl, carry);
            if (c < nchunks) s_open[c] = excl;
            carry = max(carry, __shfl_sync(0xffffffffu, incl, 31));
        }
    }
    __syncthreads();
    for (int c = warp; c < nchunks; c += CCL_THREADS / 32) {
        int x = c * 32 + lane;
        if (x >= w) continue;
        size_t p = (size_t)y * w + x;
        if (!((s_fg[c] >> lane) & 1u)) { L[p] = -1; continue; }
        unsigned bits = s_bits[c];
        unsigned starts = ~bits & (0xffffffffu >> (31 - lane));      // clear bits at lanes <= mine start a run
        int sx = starts ? c * 32 + (31 - __clz(starts)) : s_open[c];
        L[p] = y * w + sx;
    }
}

// ---------------------------------------------------------------- pass 2: merge runs
template <int PRED, int CONN>
__global__ void __launch_bounds__(CCL_THREADS) ccl_merge_kernel(const void* __restrict__ img, size_t pitch, int w, int h,
                                                                int d, int32_t* __restrict__ L)
{
    int x = blockIdx.x * CCL_THREADS + threadIdx.x;
    int y = blockIdx.y;
    if (x >= w) return;
    int p = y * w + x;
    if (PRED == 0) {
        const uint32_t* row = (const uint32_t*)img + (size_t)y * pitch;
        uint32_t c = __ldg(row + x);
        bool cl = x > 0 && color_close(c, __ldg(row + x - 1), d);
        if (y > 0) {
            const uint32_t* up = row - pitch;
            uint32_t cu = __ldg(up + x);
            if (color_close(c, cu, d)) {
                bool redundant = cl && color_close(cu, __ldg(up + x - 1), d) &&
                                 color_close(__ldg(row + x - 1), __ldg(up + x - 1), d);
                if (!redundant) uf_union(L, p, p - w);
            }
            if (CONN == 8) {   // floodFill with the 8-connectivity flag: the two upper diagonals are edges of their own
                if (x > 0 && color_close(c, __ldg(up + x - 1), d)) uf_union(L, p, p - w - 1);
                if (x + 1 < w && color_close(c, __ldg(up + x + 1), d)) uf_union(L, p, p - w + 1);
            }
        }
    } else {
        const uint8_t* row = (const uint8_t*)img + (size_t)y * pitch;
        if (!row[x]) return;
        bool left = x > 0 && row[x - 1];
        if (y > 0) {
            const uint8_t* up = row - pitch;
            bool u = up[x] != 0;
            bool ul = x > 0 && up[x - 1] != 0;
            if (u) {
                if (!(left && ul)) uf_union(L, p, p - w);
            } else if (CONN == 8) {
                bool ur = x + 1 < w && up[x + 1] != 0;
                bool right = x + 1 < w && row[x + 1] != 0;
                if (ul && !left) uf_union(L, p, p - w - 1);   // left pixel (if any) unites with ul itself
                if (ur && !right) uf_union(L, p, p - w + 1);  // right pixel (if any) unites with ur itself
            }
        }
    }
}

// ---------------------------------------------------------------- pass 3: flatten
__global__ void __launch_bounds__(CCL_THREADS) ccl_flatten_kernel(int32_t* __restrict__ L, size_t n)
{
    size_t i = (size_t)blockIdx.x * CCL_THREADS + threadIdx.x;
    if (i >= n) return;
    int v = L[i];
    if (v < 0) return;
    int r = uf_find(L, v);
    if (r != v) L[i] = r;
}

// ---------------------------------------------------------------- pass 4: canonical relabel
// MODE 0: labels are root pixel indices (>= 0), background < 0; a pixel is a "first" iff L[p] == p
// MODE 1: labels are arbitrary positive ids (<= n), <= 0 ignored; first[] holds min pixel index per id
template <int MODE>
__device__ __forceinline__ bool is_first(const int32_t* __restrict__ L, const int32_t* __restrict__ first, size_t i)
{
    int v = L[i];
    if (MODE == 0) return v == (int)i;
    return v > 0 && first[v] == (int)i;
}

__global__ void __launch_bounds__(CCL_THREADS) first_pixel_kernel(const int32_t* __restrict__ L, size_t n,
                                                                  int32_t* __restrict__ first)
{
    size_t i = (size_t)blockIdx.x * CCL_THREADS + threadIdx.x;
    int v = i < n ? L[i] : 0;
    // only the first lane of each run of equal labels inside a warp issues the atomic
    int prev = __shfl_up_sync(0xffffffffu, v, 1);
    bool head = (threadIdx.x & 31) == 0 || prev != v;
    if (i < n && v > 0 && head && first[v] > (int)i) atomicMin(first + v, (int)i);
}

__device__ __forceinline__ int block_exclusive_scan(int v, int* total)  // 256 threads
{
    __shared__ int warp_sums[CCL_THREADS / 32];
    int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    int incl = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        int t = __shfl_up_sync(0xffffffffu, incl, o);
        if (lane >= o) incl += t;
    }
    if (lane == 31) warp_sums[wid] = incl;
    __syncthreads();
    int ws = (lane < CCL_THREADS / 32) ? warp_sums[lane] : 0;
    int wincl = ws;
#pragma unroll
    for (int o = 1; o < CCL_THREADS / 32; o <<= 1) {
        int t = __shfl_up_sync(0xffffffffu, wincl, o);
        if (lane >= o) wincl += t;
    }
    int woff = __shfl_sync(0xffffffffu, wincl - ws, wid);
    if (total) *total = __shfl_sync(0xffffffffu, wincl, CCL_THREADS / 32 - 1);
    __syncthreads();
    return woff + incl - v;
}

template <int MODE>
__global__ void __launch_bounds__(CCL_THREADS) count_first_kernel(const int32_t* __restrict__ L,
                                                                  const int32_t* __restrict__ first, size_t n,
                                                                  int32_t* __restrict__ block_sums)
{
    size_t base = (size_t)blockIdx.x * SCAN_CHUNK;
    int cnt = 0;
    for (int k = threadIdx.x; k < SCAN_CHUNK; k += CCL_THREADS) {
        size_t i = base + k;
        if (i < n && is_first<MODE>(L, first, i)) cnt++;
    }
    int total;
    block_exclusive_scan(cnt, &total);
    if (threadIdx.x == 0) block_sums[blockIdx.x] = total;
}

__global__ void __launch_bounds__(CCL_THREADS) scan_block_sums_kernel(int32_t* __restrict__ block_sums, int nb,
                                                                      int32_t* __restrict__ total_out, int add_to_total)
{
    __shared__ int carry;
    if (threadIdx.x == 0) carry = 0;
    __syncthreads();
    for (int base = 0; base < nb; base += CCL_THREADS) {
        int i = base + threadIdx.x;
        int v = i < nb ? block_sums[i] : 0;
        int total;
        int ex = block_exclusive_scan(v, &total);
        int c = carry;
        if (i < nb) block_sums[i] = c + ex;
        __syncthreads();
        if (threadIdx.x == 0) carry = c + total;
        __syncthreads();
    }
    if (threadIdx.x == 0 && total_out) *total_out = carry + add_to_total;
}

// ================================================================ round 2: tile-local union-find
// The row-run passes above make five sweeps over HBM (rows: R+W, merge: R + atomics, flatten: R+W, apply: R+W).  The
// tile-local form does the bulk of the unions in shared memory and touches HBM twice:
//   A  ccl_tile_kernel    one CTA per 128 x 16 tile: pixels -> shared memory (exactly the tile, no halo: 4 B/px coalesced
//                         read), row runs by ballot, vertical (and diagonal) unions inside the tile with shared-memory
//                         atomicMin, flatten; writes L[p] = GLOBAL index of the tile-local root (4 B/px) and one bit per
//                         pixel "is a tile-local root" (1/8 B/px)
//   B  ccl_border_kernel  only the pixels on tile borders: the unions whose two pixels lie in different tiles, global
//                         atomicMin (the only global atomics of the stage); top rows are coalesced, left columns are not but
//                         there are 8 x fewer of them (tiles are 128 wide, 16 high)
//   C  ccl_roots_kernel   over the root BITMAP (n/32 words, not the pixels): tile-local roots that were linked away are
//                         pointed straight at their final root (so every chain is <= 2 links) and lose their bit; per-block
//                         counts of the surviving = global roots
//   D  scan of the block counts;  E  ccl_rank_kernel: global roots get L[root] = -(label + 1), label = 1 + rank in raster order
//   F  ccl_final_kernel   out[p] = label of L[p]'s root: one coalesced 16-byte read, <= 2 gathers that hit L1/L2 (neighbouring
//                         pixels share their tile-local root), one coalesced 16-byte write (4 + 4 B/px)
// Unions are decided per pixel exactly as in ccl_merge_kernel (same redundancy rules); a rule that cannot be evaluated inside
// the tile is simply not applied (an extra union is harmless, a missing one is not), and every cross-tile pair is visited by B.
constexpr int CT_W = 128, CT_H = 16, CT_THREADS = 256;
constexpr uint32_t CT_INVALID = 0xFF000000u;       // colour predicates: byte3 set = "no pixel"; binary: 0 = background

// PRED: 0 = packed plane (u32 per pixel, pitch in pixels), 1 = binary mask (u8, pitch in bytes), 2 = BGR bytes (pitch in bytes)
template <int PRED>
__device__ __forceinline__ uint32_t ct_load(const void* __restrict__ img, size_t pitch, int x, int y)
{
    if (PRED == 0) return __ldg((const uint32_t*)img + (size_t)y * pitch + x) & 0x00FFFFFFu;
    if (PRED == 1) return ((const uint8_t*)img)[(size_t)y * pitch + x] != 0 ? 1u : 0u;
    const uint8_t* p = (const uint8_t*)img + (size_t)y * pitch + 3 * (size_t)x;
    return (uint32_t)p[0] | ((uint32_t)p[1] << 8) | ((uint32_t)p[2] << 16);
}

template <int PRED>
__device__ __forceinline__ bool ct_fg(uint32_t v) { return PRED == 1 ? v != 0u : (v >> 24) == 0u; }

template <int PRED>
__device__ __forceinline__ bool ct_conn(uint32_t a, uint32_t b, int d)
{
    if (PRED == 1) return a != 0u && b != 0u;
    uint32_t e = __vabsdiffu4(a, b);                // byte3: 0 when both are pixels, 255 when exactly one is
    return (int)(e & 0xFF) <= d && (int)((e >> 8) & 0xFF) <= d && (int)((e >> 16) & 0xFF) <= d && (e >> 24) == 0u;
}

// the same test for two REAL pixels (byte3 = 0 on both sides) in 16-bit lanes: a lane of (|delta| + 255 - d) carries into
// bit 8 exactly when |delta| > d.  kd = 0x00FF00FF - d * 0x00010001 (d <= 255).
template <int PRED>
__device__ __forceinline__ bool ct_conn_px(uint32_t a, uint32_t b, uint32_t kd)
{
    if (PRED == 1) return a != 0u && b != 0u;
    const uint32_t e = __vabsdiffu4(a, b);
    return ((((e & 0x00FF00FFu) + kd) | (((e >> 8) & 0x00FF00FFu) + kd)) & 0x01000100u) == 0u;
}

__device__ __forceinline__ int uf_find_s(const int* lab, int a)
{
    int p = lab[a];
    while (p != a) { a = p; p = lab[a]; }
    return a;
}

__device__ __forceinline__ void uf_union_s(int* lab, int a, int b)     // shared-memory union-find, smaller index wins
{
    for (;;) {
        a = uf_find_s(lab, a);
        b = uf_find_s(lab, b);
        if (a == b) return;
        if (a < b) { int t = a; a = b; b = t; }
        int old = atomicMin(lab + a, b);
        if (old == a) return;
        a = old;
    }
}

// Per 32-pixel chunk the connectivity lives in two ballot words: clb = "connected to the left pixel" (run structure of the
// row) and vb = "connected to the pixel above".  A vertical union is needed only at the first column of an overlap of two
// runs: need = vb & ~(clb & clb_of_the_row_above & (vb << 1)) -- the ccl_merge_kernel redundancy rule in bit-parallel form
// (pixel x is redundant when x-1 is in its run, x-1's upper neighbour is in the upper run, and x-1 is connected upwards).
// Per-PIXEL work is the load, two predicate tests, two ballots and the final store; the union-find work (unions, find,
// conversion to the global index, root flags) is done once per RUN by the lane that starts it.  The kernel is bound by the
// instruction issue rate (ncu: 70 % of the issue slots at 14 % of the DRAM bandwidth; 162 thread-instructions per pixel in
// the first per-pixel form).  A variant that compacted the unions / finds into shared-memory task lists processed one task per
// thread was SLOWER (521 vs 380 us at 8192^2): the pointer chases then sit between block-wide barriers instead of
// overlapping other warps' streaming work (profiles/r02_ccl_history.md).
template <int PRED, int CONN>
__global__ void __launch_bounds__(CT_THREADS) ccl_tile_kernel(const void* __restrict__ img, size_t pitch, int w, int h, int d,
                                                              int32_t* __restrict__ L, uint32_t* __restrict__ bitmap, int wp)
{
    __shared__ uint32_t s_col[CT_H * CT_W];   // pixels; after the unions: global index of the root, stored at every run start
    __shared__ int s_lab[CT_H * CT_W];        // union-find parents (tile-local indices); a pixel points at the start of its run
    __shared__ unsigned s_clb[CT_H * (CT_W / 32)];
    constexpr int CHUNKS = CT_W / 32;                            // 4
    constexpr int ROWS_PER_WARP = CT_H / (CT_THREADS / 32);      // 2
    const int lane = threadIdx.x & 31, wq = threadIdx.x >> 5;
    const int tx0 = blockIdx.x * CT_W, ty0 = blockIdx.y * CT_H;
    const uint32_t kd = 0x00FF00FFu - (uint32_t)min(d, 255) * 0x00010001u;
    const unsigned le_mask = 0xffffffffu >> (31 - lane);
    uint32_t vreg[ROWS_PER_WARP][CHUNKS];
    unsigned clb[ROWS_PER_WARP][CHUNKS];
    int sreg[ROWS_PER_WARP][CHUNKS];          // tile-local index of the start of my run

    // ---- load + row runs
#pragma unroll
    for (int rr = 0; rr < ROWS_PER_WARP; rr++) {
        const int r = wq * ROWS_PER_WARP + rr, gy = ty0 + r;
        int carry = r * CT_W;                 // start of the run that reaches the end of the previous chunk
        uint32_t prev_last = 0u;
#pragma unroll
        for (int c = 0; c < CHUNKS; c++) {
            const int x = c * 32 + lane, gx = tx0 + x;
            const bool in = gx < w && gy < h;
            const uint32_t v = in ? ct_load<PRED>(img, pitch, gx, gy) : (PRED == 1 ? 0u : CT_INVALID);
            vreg[rr][c] = v;
            s_col[r * CT_W + x] = v;
            uint32_t left = __shfl_up_sync(0xffffffffu, v, 1);
            if (lane == 0) left = prev_last;
            // the left neighbour of a real pixel is a real pixel, except for the first column of the tile (no link inside the tile)
            const bool cl = ct_fg<PRED>(v) && x > 0 && ct_conn_px<PRED>(v, left, kd);
            const unsigned bits = __ballot_sync(0xffffffffu, cl);
            clb[rr][c] = bits;
            if (lane == 0) s_clb[r * CHUNKS + c] = bits;
            const unsigned starts = ~bits & le_mask;
            const int sp = starts ? r * CT_W + c * 32 + (31 - __clz(starts)) : carry;
            sreg[rr][c] = sp;
            s_lab[r * CT_W + x] = ct_fg<PRED>(v) ? sp : -1;
            if (bits != 0xffffffffu) carry = r * CT_W + c * 32 + (31 - __clz(~bits));
            prev_last = __shfl_sync(0xffffffffu, v, 31);
        }
    }
    __syncthreads();

    // ---- unions with the row above, inside the tile
#pragma unroll
    for (int rr = 0; rr < ROWS_PER_WARP; rr++) {
        const int r = wq * ROWS_PER_WARP + rr;
        if (r == 0) continue;                 // the tile's top row unites with the tile above in ccl_border_kernel
        unsigned vcarry = 0;                  // "connected upwards" of the last pixel of the previous chunk
#pragma unroll
        for (int c = 0; c < CHUNKS; c++) {
            const int x = c * 32 + lane, p = r * CT_W + x;
            const uint32_t v = vreg[rr][c];
            const uint32_t u = rr > 0 ? vreg[rr > 0 ? rr - 1 : 0][c] : s_col[p - CT_W];
            // the pixel above a real pixel is a real pixel (r >= 1)
            const bool vup = ct_fg<PRED>(v) && ct_conn_px<PRED>(v, u, kd);
            const unsigned vb = __ballot_sync(0xffffffffu, vup);
            const unsigned clup = rr > 0 ? clb[rr > 0 ? rr - 1 : 0][c] : s_clb[(r - 1) * CHUNKS + c];
            const unsigned need = vb & ~(clb[rr][c] & clup & ((vb << 1) | vcarry));
            vcarry = vb >> 31;
            if ((need >> lane) & 1u) uf_union_s(s_lab, sreg[rr][c], rr > 0 ? sreg[rr > 0 ? rr - 1 : 0][c] : s_lab[p - CT_W]);
            if (CONN == 8) {
                if (PRED != 1) {              // floodFill's 8-connectivity: both upper diagonals are edges of their own
                    if (ct_fg<PRED>(v)) {
                        if (x > 0 && ct_conn_px<PRED>(v, s_col[p - CT_W - 1], kd)) uf_union_s(s_lab, p, p - CT_W - 1);
                        if (x + 1 < CT_W && ct_conn<PRED>(v, s_col[p - CT_W + 1], d)) uf_union_s(s_lab, p, p - CT_W + 1);
                    }
                } else if (v && !u) {         // connectedComponents(8): diagonals matter only under a background pixel
                    const uint32_t lf = x > 0 ? s_col[p - 1] : 0u, ul = x > 0 ? s_col[p - CT_W - 1] : 0u;
                    const uint32_t rt = x + 1 < CT_W ? s_col[p + 1] : 0u, ur = x + 1 < CT_W ? s_col[p - CT_W + 1] : 0u;
                    if (ul && !lf) uf_union_s(s_lab, p, p - CT_W - 1);
                    if (ur && !rt) uf_union_s(s_lab, p, p - CT_W + 1);
                }
            }
        }
    }
    __syncthreads();

    // ---- per run: find the root, leave its GLOBAL index at the run start (s_col is free now), flag tile-local roots
#pragma unroll
    for (int rr = 0; rr < ROWS_PER_WARP; rr++) {
        const int r = wq * ROWS_PER_WARP + rr, gy = ty0 + r;
#pragma unroll
        for (int c = 0; c < CHUNKS; c++) {
            const int x = c * 32 + lane, p = r * CT_W + x;
            bool is_root = false;
            if (ct_fg<PRED>(vreg[rr][c]) && !((clb[rr][c] >> lane) & 1u)) {          // I start a run
                const int root = uf_find_s(s_lab, p);
                is_root = root == p;
                s_col[p] = (uint32_t)((ty0 + root / CT_W) * w + tx0 + (root % CT_W));
            }
            const unsigned bal = __ballot_sync(0xffffffffu, is_root);
            if (lane == 0 && gy < h && tx0 + c * 32 < w) bitmap[(size_t)gy * wp + (tx0 >> 5) + c] = bal;
        }
    }
    __syncthreads();

    // ---- per pixel: the label of my run
#pragma unroll
    for (int rr = 0; rr < ROWS_PER_WARP; rr++) {
        const int gy = ty0 + wq * ROWS_PER_WARP + rr;
        if (gy >= h) continue;
        int32_t* __restrict__ Lrow = L + (size_t)gy * w + tx0;
#pragma unroll
        for (int c = 0; c < CHUNKS; c++) {
            const int x = c * 32 + lane;
            if (tx0 + x < w) Lrow[x] = ct_fg<PRED>(vreg[rr][c]) ? (int32_t)s_col[sreg[rr][c]] : -1;
        }
    }
}

// ---- four pixels per lane (4-connectivity, colour predicates; rows that allow 16-byte / 4-byte aligned vector accesses).
// ccl_tile_kernel above spends ~160 thread-instructions per pixel (two ballots, a shuffle and the run bookkeeping PER PIXEL).
// Here one warp owns a whole 128-pixel tile row and a lane owns 4 consecutive pixels: the links between a lane's own pixels are
// plain register compares, the run structure of the row needs two ballots per ROW -- C = "my first pixel is linked to the lane
// on my left", T = "my four pixels are linked to each other": the run through my first pixel starts in the highest lane below
// me whose bit in (C & T) is clear, at that lane's last internal break -- one shuffle.  The vertical rule is the same
// (need = vb & ~(cl & cl_above & (vb << 1))) on 4-bit masks, with one shuffle for the carry between lanes.  Loads, parent
// stores and label stores are 16-byte accesses; the root bits of 8 lanes are one bitmap word (__reduce_or_sync).
// Produces exactly what ccl_tile_kernel<PRED, 4> produces (every tile-local component points at its smallest pixel).
template <int PRED>
__device__ __forceinline__ void ct4_load(const void* __restrict__ img, size_t pitch, int gx, int gy, uint32_t (&v)[4])
{
    if (PRED == 0) {
        const uint4 q = __ldg(reinterpret_cast<const uint4*>((const uint32_t*)img + (size_t)gy * pitch + gx));
        v[0] = q.x & 0x00FFFFFFu; v[1] = q.y & 0x00FFFFFFu; v[2] = q.z & 0x00FFFFFFu; v[3] = q.w & 0x00FFFFFFu;
    } else {
        const uint32_t* p = reinterpret_cast<const uint32_t*>((const uint8_t*)img + (size_t)gy * pitch + 3 * (size_t)gx);
        const uint32_t w0 = __ldg(p), w1 = __ldg(p + 1), w2 = __ldg(p + 2);
        v[0] = w0 & 0x00FFFFFFu;
        v[1] = (w0 >> 24) | ((w1 & 0xFFFFu) << 8);
        v[2] = (w1 >> 16) | ((w2 & 0xFFu) << 16);
        v[3] = w2 >> 8;
    }
}

template <int PRED>
__global__ void __launch_bounds__(CT_THREADS) ccl_tile4_kernel(const void* __restrict__ img, size_t pitch, int w, int h, int d,
                                                               int32_t* __restrict__ L, uint32_t* __restrict__ bitmap, int wp)
{
    __shared__ __align__(16) uint32_t s_col[CT_H * CT_W];   // pixels; after the unions: global index of the root at every run start
    __shared__ __align__(16) int s_lab[CT_H * CT_W];        // union-find parents (tile-local indices)
    __shared__ unsigned char s_hb[CT_H * 32];               // per lane: the four "linked to my left neighbour" bits
    constexpr int ROWS_PER_WARP = CT_H / (CT_THREADS / 32); // 2
    const int lane = threadIdx.x & 31, wq = threadIdx.x >> 5;
    const int tx0 = blockIdx.x * CT_W, ty0 = blockIdx.y * CT_H;
    const int gx = tx0 + 4 * lane;
    const uint32_t kd = 0x00FF00FFu - (uint32_t)min(d, 255) * 0x00010001u;
    const unsigned below = (1u << lane) - 1u;
    uint32_t vreg[ROWS_PER_WARP][4];
    unsigned hbr[ROWS_PER_WARP];              // bit j: pixel j is linked to the pixel on its left
    int st0[ROWS_PER_WARP];                   // tile-local index of the start of the run through my first pixel

    // ---- load + row runs
#pragma unroll
    for (int rr = 0; rr < ROWS_PER_WARP; rr++) {
        const int r = wq * ROWS_PER_WARP + rr, gy = ty0 + r;
        uint32_t (&v)[4] = vreg[rr];
        if (gx < w && gy < h) ct4_load<PRED>(img, pitch, gx, gy, v);
        else v[0] = v[1] = v[2] = v[3] = CT_INVALID;
        reinterpret_cast<uint4*>(s_col)[r * 32 + lane] = make_uint4(v[0], v[1], v[2], v[3]);
        const uint32_t left = __shfl_up_sync(0xffffffffu, v[3], 1);
        // the left neighbour of a real pixel is a real pixel, except for the first column of the tile (no link inside the tile)
        unsigned hb = (ct_fg<PRED>(v[0]) && lane > 0 && ct_conn_px<PRED>(v[0], left, kd)) ? 1u : 0u;
        hb |= (ct_fg<PRED>(v[1]) && ct_conn_px<PRED>(v[1], v[0], kd)) ? 2u : 0u;
        hb |= (ct_fg<PRED>(v[2]) && ct_conn_px<PRED>(v[2], v[1], kd)) ? 4u : 0u;
        hb |= (ct_fg<PRED>(v[3]) && ct_conn_px<PRED>(v[3], v[2], kd)) ? 8u : 0u;
        hbr[rr] = hb;
        s_hb[r * 32 + lane] = (unsigned char)hb;
        const unsigned C = __ballot_sync(0xffffffffu, hb & 1u);
        const unsigned T = __ballot_sync(0xffffffffu, (hb & 0xEu) == 0xEu);
        // start of the run through my last pixel, as an offset inside my four: 3, 2, 1, or 0 when the four are linked
        const int lb = !(hb & 8u) ? 3 : (!(hb & 4u) ? 2 : (!(hb & 2u) ? 1 : 0));
        const unsigned stop = ~(C & T) & below;                      // lanes below me where the run cannot pass through
        const int sl = (hb & 1u) ? 31 - __clz(stop) : lane;          // lane 0 never links left, so stop != 0 whenever bit 0 of hb is set
        const int slb = __shfl_sync(0xffffffffu, lb, sl);
        const int base = r * CT_W;
        const int s0 = (hb & 1u) ? base + 4 * sl + slb : base + 4 * lane;
        st0[rr] = s0;
        int4 par;
        par.x = s0;
        par.y = (hb & 2u) ? par.x : base + 4 * lane + 1;
        par.z = (hb & 4u) ? par.y : base + 4 * lane + 2;
        par.w = (hb & 8u) ? par.z : base + 4 * lane + 3;
        if (!ct_fg<PRED>(v[0])) par.x = -1;
        if (!ct_fg<PRED>(v[1])) par.y = -1;
        if (!ct_fg<PRED>(v[2])) par.z = -1;
        if (!ct_fg<PRED>(v[3])) par.w = -1;
        reinterpret_cast<int4*>(s_lab)[r * 32 + lane] = par;
    }
    __syncthreads();

    // ---- unions with the row above, inside the tile.  (All rows at once: the runs of a region that crosses the tile may link
    //      into a chain that later finds walk -- ncu: 23 % of the instructions are the find loop at 1.9 active lanes -- but
    //      closing the seams between the warps' row pairs as a binary tree, three more block barriers, measured SLOWER:
    //      297 -> 328 us at 8192^2, profiles/r02_ccl_quad.md.)
#pragma unroll
    for (int rr = 0; rr < ROWS_PER_WARP; rr++) {
        const int r = wq * ROWS_PER_WARP + rr;
        if (r == 0) continue;                 // the tile's top row unites with the tile above in ccl_border_kernel
        uint32_t u[4];
        unsigned hbu;
        if (rr > 0) {
#pragma unroll
            for (int j = 0; j < 4; j++) u[j] = vreg[rr > 0 ? rr - 1 : 0][j];
            hbu = hbr[rr > 0 ? rr - 1 : 0];
        } else {
            const uint4 q = reinterpret_cast<const uint4*>(s_col)[(r - 1) * 32 + lane];
            u[0] = q.x; u[1] = q.y; u[2] = q.z; u[3] = q.w;
            hbu = s_hb[(r - 1) * 32 + lane];
        }
        unsigned vb = 0;
#pragma unroll
        for (int j = 0; j < 4; j++)           // the pixel above a real pixel is a real pixel (r >= 1)
            vb |= (ct_fg<PRED>(vreg[rr][j]) && ct_conn_px<PRED>(vreg[rr][j], u[j], kd)) ? (1u << j) : 0u;
        const unsigned carry = __shfl_up_sync(0xffffffffu, vb >> 3, 1) & (lane > 0 ? 1u : 0u);
        unsigned need = vb & ~(hbr[rr] & hbu & ((vb << 1) | carry)) & 0xFu;
        const int p0 = r * CT_W + 4 * lane;
        while (need) {
            const int j = __ffs(need) - 1;
            need &= need - 1;
            uf_union_s(s_lab, p0 + j, p0 + j - CT_W);
        }
    }
    __syncthreads();

    // ---- per run: find the root, leave its GLOBAL index at the run start (s_col is free now), flag tile-local roots
#pragma unroll
    for (int rr = 0; rr < ROWS_PER_WARP; rr++) {
        const int r = wq * ROWS_PER_WARP + rr, gy = ty0 + r;
        const int p0 = r * CT_W + 4 * lane;
        unsigned roots = 0;
#pragma unroll
        for (int j = 0; j < 4; j++) {
            if (ct_fg<PRED>(vreg[rr][j]) && !((hbr[rr] >> j) & 1u)) {               // I start a run
                const int root = uf_find_s(s_lab, p0 + j);
                if (root == p0 + j) roots |= 1u << j;
                s_col[p0 + j] = (uint32_t)((ty0 + root / CT_W) * w + tx0 + (root % CT_W));
            }
        }
        const unsigned word = __reduce_or_sync(0xFFu << (lane & 24), roots << (4 * (lane & 7)));
        if ((lane & 7) == 0 && gy < h && tx0 + 4 * lane < w) bitmap[(size_t)gy * wp + (tx0 >> 5) + (lane >> 3)] = word;
    }
    __syncthreads();

    // ---- per pixel: the label of my run
#pragma unroll
    for (int rr = 0; rr < ROWS_PER_WARP; rr++) {
        const int r = wq * ROWS_PER_WARP + rr, gy = ty0 + r;
        if (gy >= h || gx >= w) continue;
        const int p0 = r * CT_W + 4 * lane;
        const unsigned hb = hbr[rr];
        int4 o;
        o.x = ct_fg<PRED>(vreg[rr][0]) ? (int32_t)s_col[st0[rr]] : -1;
        o.y = (hb & 2u) ? o.x : (ct_fg<PRED>(vreg[rr][1]) ? (int32_t)s_col[p0 + 1] : -1);
        o.z = (hb & 4u) ? o.y : (ct_fg<PRED>(vreg[rr][2]) ? (int32_t)s_col[p0 + 2] : -1);
        o.w = (hb & 8u) ? o.z : (ct_fg<PRED>(vreg[rr][3]) ? (int32_t)s_col[p0 + 3] : -1);
        *reinterpret_cast<int4*>(L + (size_t)gy * w + gx) = o;
    }
}

// unions across tile borders.  Threads [0, n_hb): pixels of the top rows of the tiles (y = k * CT_H, k >= 1), the per-pixel
// rules of ccl_merge_kernel with the row above; threads [n_hb, n_hb + n_vb): pixels of the left columns of the tiles
// (x = j * CT_W, j >= 1) against the column to their left (and its diagonals for 8-connectivity).
template <int PRED, int CONN>
__global__ void __launch_bounds__(CT_THREADS) ccl_border_kernel(const void* __restrict__ img, size_t pitch, int w, int h, int d,
                                                                int32_t* __restrict__ L, long long n_hb, long long n_vb)
{
    const long long idx = (long long)blockIdx.x * CT_THREADS + threadIdx.x;
    const uint32_t none = PRED == 1 ? 0u : CT_INVALID;
    auto get = [&](int x, int y) -> uint32_t { return (x >= 0 && x < w && y >= 0 && y < h) ? ct_load<PRED>(img, pitch, x, y) : none; };
    // top rows: the lane on my left holds the pixel on my left (and the one above it) unless I am the first lane of the warp or
    // the first pixel of a row, so the redundancy test costs two shuffles instead of two more (byte-wise, for BGR) loads
    const bool top = idx < n_hb;
    int x = 0, y = 0;
    uint32_t v = none, u = none;
    if (top) {
        x = (int)(idx % w); y = (int)(idx / w + 1) * CT_H;
        v = get(x, y); u = get(x, y - 1);
    }
    uint32_t lf = __shfl_up_sync(0xffffffffu, v, 1), ul = __shfl_up_sync(0xffffffffu, u, 1);
    if (top) {
        if (!ct_fg<PRED>(v)) return;
        if ((threadIdx.x & 31) == 0 || x == 0) { lf = get(x - 1, y); ul = get(x - 1, y - 1); }
        const int p = y * w + x;
        if (PRED != 1) {
            if (ct_conn<PRED>(v, u, d)) {
                const bool redundant = ct_conn<PRED>(v, lf, d) && ct_conn<PRED>(u, ul, d) && ct_conn<PRED>(lf, ul, d);
                if (!redundant) uf_union(L, p, p - w);
            }
            if (CONN == 8) {
                if (ct_conn<PRED>(v, ul, d)) uf_union(L, p, p - w - 1);
                if (ct_conn<PRED>(v, get(x + 1, y - 1), d)) uf_union(L, p, p - w + 1);
            }
        } else {
            if (u) {
                if (!(lf && ul)) uf_union(L, p, p - w);
            } else if (CONN == 8) {
                const uint32_t rt = get(x + 1, y), ur = get(x + 1, y - 1);
                if (ul && !lf) uf_union(L, p, p - w - 1);
                if (ur && !rt) uf_union(L, p, p - w + 1);
            }
        }
    } else if (idx < n_hb + n_vb) {
        const long long k = idx - n_hb;
        const int y = (int)(k % h), x = (int)(k / h + 1) * CT_W;
        const uint32_t v = get(x, y);
        if (!ct_fg<PRED>(v)) return;
        const int p = y * w + x;
        const uint32_t lf = get(x - 1, y);
        if (ct_conn<PRED>(v, lf, d)) {
            // the pixel above unites the same two tile-local components when it is connected to me, its left neighbour to mine,
            // and it to its left neighbour -- unless it sits in the tile row above (then my link to it is not made yet)
            bool redundant = false;
            if (y % CT_H != 0) {
                const uint32_t u = get(x, y - 1), ul = get(x - 1, y - 1);
                redundant = ct_conn<PRED>(v, u, d) && ct_conn<PRED>(lf, ul, d) && ct_conn<PRED>(u, ul, d);
            }
            if (!redundant) uf_union(L, p, p - 1);
        }
        if (CONN == 8) {
            if (ct_conn<PRED>(v, get(x - 1, y - 1), d)) uf_union(L, p, p - w - 1);
            if (ct_conn<PRED>(v, get(x - 1, y + 1), d)) uf_union(L, p, p + w - 1);
        }
    }
}

constexpr int RB_WORDS = 4;                               // bitmap words per thread in the root kernels
constexpr int RB_BLOCK = CT_THREADS * RB_WORDS;           // 1024 words = 32768 pixels per block

// C: which tile-local roots are still roots after the border pass?  One independent load per candidate (L[p] == p), no
// pointer chasing; the surviving bits go to a second bitmap, the per-block counts to block_sums.
__global__ void __launch_bounds__(CT_THREADS) ccl_roots_kernel(const int32_t* __restrict__ L, const uint32_t* __restrict__ bitmap,
                                                               uint32_t* __restrict__ rootmap, int w, int wp, size_t nwords,
                                                               int32_t* __restrict__ block_sums)
{
    const size_t base = (size_t)blockIdx.x * RB_BLOCK + (size_t)threadIdx.x * RB_WORDS;
    unsigned bits[RB_WORDS];
#pragma unroll
    for (int k = 0; k < RB_WORDS; k++) bits[k] = base + k < nwords ? bitmap[base + k] : 0u;
    int cnt = 0;
#pragma unroll
    for (int k = 0; k < RB_WORDS; k++) {
        const size_t wi = base + k;
        unsigned keep = 0;
        if (bits[k]) {
            const int p0 = (int)(wi / wp) * w + (int)(wi % wp) * 32;
            for (unsigned b = bits[k]; b;) {                      // four independent loads in flight
                int q[4], v[4];
#pragma unroll
                for (int j = 0; j < 4; j++) { q[j] = b ? __ffs(b) - 1 : -1; b &= b - 1; }
#pragma unroll
                for (int j = 0; j < 4; j++) v[j] = q[j] >= 0 ? __ldcg(L + p0 + q[j]) : -1;
#pragma unroll
                for (int j = 0; j < 4; j++) if (q[j] >= 0 && v[j] == p0 + q[j]) keep |= 1u << q[j];
            }
        }
        if (wi < nwords) rootmap[wi] = keep;
        cnt += __popc(keep);
    }
    int total;
    block_exclusive_scan(cnt, &total);
    if (threadIdx.x == 0) block_sums[blockIdx.x] = total;
}

// D: exclusive scan of up to 1024 * per entries by one CTA: every thread sums `per` consecutive entries, one block scan,
// then the entries are rewritten with their exclusive prefixes; *total_out = sum + add_to_total.
__global__ void __launch_bounds__(1024) scan_sums_kernel(int32_t* __restrict__ sums, int nb, int per, int32_t* __restrict__ total_out,
                                                         int add_to_total)
{
    __shared__ int wsum[32];
    const int lane = threadIdx.x & 31, wq = threadIdx.x >> 5;
    const int i0 = threadIdx.x * per;
    int s = 0;
    for (int k = 0; k < per; k++) s += i0 + k < nb ? sums[i0 + k] : 0;
    int incl = s;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        int t = __shfl_up_sync(0xffffffffu, incl, o);
        if (lane >= o) incl += t;
    }
    if (lane == 31) wsum[wq] = incl;
    __syncthreads();
    if (wq == 0) {
        int v = wsum[lane], inc2 = v;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            int t = __shfl_up_sync(0xffffffffu, inc2, o);
            if (lane >= o) inc2 += t;
        }
        wsum[lane] = inc2 - v;
        if (lane == 31 && total_out) *total_out = inc2 + add_to_total;
    }
    __syncthreads();
    int run = wsum[wq] + incl - s;
    for (int k = 0; k < per; k++) {
        if (i0 + k < nb) { int v = sums[i0 + k]; sums[i0 + k] = run; run += v; }
    }
}

// E: the roots get L[root] = -(label + 1), label = 1 + rank in raster order (-1 stays "background")
__global__ void __launch_bounds__(CT_THREADS) ccl_rank_kernel(int32_t* __restrict__ L, const uint32_t* __restrict__ rootmap, int w,
                                                              int wp, size_t nwords, const int32_t* __restrict__ block_offs)
{
    const size_t base = (size_t)blockIdx.x * RB_BLOCK + (size_t)threadIdx.x * RB_WORDS;
    unsigned bits[RB_WORDS];
    int cnt = 0;
#pragma unroll
    for (int k = 0; k < RB_WORDS; k++) {
        bits[k] = base + k < nwords ? rootmap[base + k] : 0u;
        cnt += __popc(bits[k]);
    }
    int label = block_exclusive_scan(cnt, nullptr) + block_offs[blockIdx.x] + 1;
#pragma unroll
    for (int k = 0; k < RB_WORDS; k++) {
        if (!bits[k]) continue;
        const size_t wi = base + k;
        const int p0 = (int)(wi / wp) * w + (int)(wi % wp) * 32;
        for (unsigned b = bits[k]; b; b &= b - 1) {
            L[p0 + __ffs(b) - 1] = -(label + 1);
            label++;
        }
    }
}

// E': every tile-local root that was linked away takes the encoded label of its final root, so that the per-pixel pass needs
// one gather only.  One thread per bitmap word (few candidates per thread: the chains of a warp run side by side).  A
// concurrent traversal that meets an already rewritten entry stops there with the same label.
__global__ void __launch_bounds__(CT_THREADS) ccl_spread_kernel(int32_t* __restrict__ L, const uint32_t* __restrict__ bitmap,
                                                                const uint32_t* __restrict__ rootmap, int w, int wp, size_t nwords)
{
    const size_t wi = (size_t)blockIdx.x * CT_THREADS + threadIdx.x;
    if (wi >= nwords) return;
    unsigned b = bitmap[wi] & ~rootmap[wi];
    if (!b) return;
    const int p0 = (int)(wi / wp) * w + (int)(wi % wp) * 32;
    for (; b; b &= b - 1) {
        const int p = p0 + __ffs(b) - 1;
        int v = __ldcg(L + p);
        while (v >= 0) v = __ldcg(L + v);
        L[p] = v;
    }
}

__device__ __forceinline__ int ccl_resolve(const int32_t* __restrict__ L, int v)
{
    while (v >= 0) v = __ldcg(L + v);         // own entry -> tile root (which holds the encoded label after ccl_spread_kernel)
    return -(v + 1);                          // -1 (background) -> 0
}

__global__ void __launch_bounds__(CT_THREADS) ccl_final_kernel(const int32_t* __restrict__ L, int32_t* __restrict__ out, size_t n)
{
    const size_t i4 = ((size_t)blockIdx.x * CT_THREADS + threadIdx.x) * 4;
    if (i4 + 3 < n && ((reinterpret_cast<uintptr_t>(L) | reinterpret_cast<uintptr_t>(out)) & 15) == 0) {
        const int4 v = __ldcs(reinterpret_cast<const int4*>(L + i4));
        int4 o;
        o.x = ccl_resolve(L, v.x); o.y = ccl_resolve(L, v.y); o.z = ccl_resolve(L, v.z); o.w = ccl_resolve(L, v.w);
        __stcs(reinterpret_cast<int4*>(out + i4), o);
    } else {
        for (size_t i = i4; i < n && i < i4 + 4; i++) out[i] = ccl_resolve(L, L[i]);
    }
}

// rank[] is indexed by pixel (MODE 0: at the root pixel) or by label id (MODE 1)
template <int MODE>
__global__ void __launch_bounds__(CCL_THREADS) assign_rank_kernel(const int32_t* __restrict__ L,
                                                                  const int32_t* __restrict__ first, size_t n,
                                                                  const int32_t* __restrict__ block_offs,
                                                                  int32_t* __restrict__ rank)
{
    // each thread owns 16 consecutive pixels of the chunk so that ranks follow raster order
    size_t base = (size_t)blockIdx.x * SCAN_CHUNK + (size_t)threadIdx.x * (SCAN_CHUNK / CCL_THREADS);
    unsigned flags = 0;
    int cnt = 0;
#pragma unroll
    for (int k = 0; k < SCAN_CHUNK / CCL_THREADS; k++) {
        size_t i = base + k;
        if (i < n && is_first<MODE>(L, first, i)) { flags |= 1u << k; cnt++; }
    }
    int ex = block_exclusive_scan(cnt, nullptr) + block_offs[blockIdx.x];
#pragma unroll
    for (int k = 0; k < SCAN_CHUNK / CCL_THREADS; k++) {
        if (flags & (1u << k)) {
            size_t i = base + k;
            if (MODE == 0) rank[i] = ex; else rank[L[i]] = ex;
            ex++;
        }
    }
}

__global__ void __launch_bounds__(CCL_THREADS) apply_rank_kernel(int32_t* __restrict__ L, size_t n,
                                                                 const int32_t* __restrict__ rank, int mode)
{
    size_t i = (size_t)blockIdx.x * CCL_THREADS + threadIdx.x;
    if (i >= n) return;
    int v = L[i];
    if (mode == 0) L[i] = v >= 0 ? rank[v] + 1 : 0;
    else if (v > 0) L[i] = rank[v] + 1;
}

// MODE 0 fast path: flatten + root flags + block-local ranks in ONE pass.  A block owns 4096 consecutive pixels and walks
// them in 16 coalesced sweeps of 256 (pixel = base + sweep * 256 + thread, i.e. raster order = sweep-major); the root flags
// of every (sweep, warp) are kept as ballots in shared memory, one warp turns their popcounts into exclusive prefixes and
// lrank[root pixel] = number of roots before it inside the block's 4096 pixels.
__global__ void __launch_bounds__(CCL_THREADS) flatten_count_kernel(int32_t* __restrict__ L, size_t n,
                                                                    int32_t* __restrict__ block_sums,
                                                                    int32_t* __restrict__ lrank)
{
    constexpr int SWEEPS = SCAN_CHUNK / CCL_THREADS;      // 16
    constexpr int WARPS = CCL_THREADS / 32;               // 8
    __shared__ unsigned s_ballot[SWEEPS * WARPS];
    __shared__ int s_prefix[SWEEPS * WARPS];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const size_t base = (size_t)blockIdx.x * SCAN_CHUNK;
    int v[SWEEPS];
#pragma unroll
    for (int k = 0; k < SWEEPS; k++) {
        size_t i = base + (size_t)k * CCL_THREADS + threadIdx.x;
        v[k] = i < n ? L[i] : -1;
    }
#pragma unroll
    for (int k = 0; k < SWEEPS; k++) {
        size_t i = base + (size_t)k * CCL_THREADS + threadIdx.x;
        bool is_root = false;
        if (v[k] >= 0) {
            int r = uf_find(L, v[k]);
            if (r != v[k]) L[i] = r;
            is_root = r == (int)i;
        }
        unsigned bal = __ballot_sync(0xffffffffu, is_root);
        if (lane == 0) s_ballot[k * WARPS + warp] = bal;
    }
    __syncthreads();
    if (warp == 0) {   // exclusive prefix over the 128 popcounts, in raster order (sweep-major, warp-minor)
        int carry = 0;
#pragma unroll
        for (int q = 0; q < SWEEPS * WARPS / 32; q++) {
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
    for (int k = 0; k < SWEEPS; k++) {
        unsigned bal = s_ballot[k * WARPS + warp];
        if ((bal >> lane) & 1u) {
            size_t i = base + (size_t)k * CCL_THREADS + threadIdx.x;
            lrank[i] = s_prefix[k * WARPS + warp] + __popc(bal & ((1u << lane) - 1));
        }
    }
}

__global__ void __launch_bounds__(CCL_THREADS) apply_lrank_kernel(int32_t* __restrict__ L, size_t n,
                                                                  const int32_t* __restrict__ block_offs,
                                                                  const int32_t* __restrict__ lrank)
{
    size_t i = (size_t)blockIdx.x * CCL_THREADS + threadIdx.x;
    if (i >= n) return;
    int v = L[i];
    L[i] = v >= 0 ? __ldg(block_offs + v / SCAN_CHUNK) + __ldg(lrank + v) + 1 : 0;
}

__global__ void __launch_bounds__(CCL_THREADS) fill_i32_kernel(int32_t* __restrict__ p, size_t n, int32_t v)
{
    size_t i = (size_t)blockIdx.x * CCL_THREADS + threadIdx.x;
    if (i < n) p[i] = v;
}

__global__ void __launch_bounds__(CCL_THREADS) copy_labels_2d_kernel(const int32_t* __restrict__ src, size_t sstep,
                                                                     int32_t* __restrict__ dst, size_t dstep, int w)
{
    int x = blockIdx.x * CCL_THREADS + threadIdx.x;
    int y = blockIdx.y;
    if (x >= w) return;
    const int32_t* s = (const int32_t*)((const char*)src + (size_t)y * sstep);
    int32_t* d = (int32_t*)((char*)dst + (size_t)y * dstep);
    d[x] = s[x];
}

__global__ void __launch_bounds__(CCL_THREADS) add_label_base_kernel(int32_t* __restrict__ L, size_t n, int w, int fullw,
                                                                     long long base)
{
    // strip labelling: local root index (row*w + x) -> 1 + global linear index
    size_t i = (size_t)blockIdx.x * CCL_THREADS + threadIdx.x;
    if (i >= n) return;
    int v = L[i];
    if (v < 0) { L[i] = 0; return; }
    L[i] = (int32_t)(base + (long long)(v / w) * fullw + (v % w) + 1);
}

__global__ void __launch_bounds__(CCL_THREADS) seam_pairs_kernel(const uint8_t* __restrict__ up_bgr,
                                                                 const int32_t* __restrict__ up_lab,
                                                                 const uint8_t* __restrict__ lo_bgr,
                                                                 const int32_t* __restrict__ lo_lab, int w, int d,
                                                                 int32_t* __restrict__ pairs, int32_t* __restrict__ count)
{
    int x = blockIdx.x * CCL_THREADS + threadIdx.x;
    if (x >= w) return;
    auto close = [&](const uint8_t* a, const uint8_t* b) {
        return abs((int)a[0] - (int)b[0]) <= d && abs((int)a[1] - (int)b[1]) <= d && abs((int)a[2] - (int)b[2]) <= d;
    };
    if (!close(up_bgr + 3 * x, lo_bgr + 3 * x)) return;
    int a = up_lab[x], b = lo_lab[x];
    if (a == b) return;
    // skip a pair identical to the one of the previous column (same two runs)
    if (x > 0 && up_lab[x - 1] == a && lo_lab[x - 1] == b && close(up_bgr + 3 * (x - 1), lo_bgr + 3 * (x - 1))) return;
    int slot = atomicAdd(count, 1);
    pairs[2 * slot] = a;
    pairs[2 * slot + 1] = b;
}

__global__ void __launch_bounds__(CCL_THREADS) apply_map_kernel(int32_t* __restrict__ labels, size_t lstep, int w,
                                                                const int32_t* __restrict__ from,
                                                                const int32_t* __restrict__ to, int n)
{
    int x = blockIdx.x * CCL_THREADS + threadIdx.x;
    int y = blockIdx.y;
    if (x >= w) return;
    int32_t* row = (int32_t*)((char*)labels + (size_t)y * lstep);
    int v = row[x];
    int lo = 0, hi = n - 1;
    while (lo <= hi) {
        int mid = (lo + hi) >> 1;
        int f = __ldg(from + mid);
        if (f == v) { row[x] = __ldg(to + mid); return; }
        if (f < v) lo = mid + 1; else hi = mid - 1;
    }
}

// ---------------------------------------------------------------- strip sharding: dense global numbering
// Labels of a strip after seam resolution are 1 + GLOBAL index of the component's first pixel.  A pixel is an owned root
// iff its label points at itself.  Ranks of owned roots inside 4096-pixel chunks (same layout as flatten_count_kernel).
__global__ void __launch_bounds__(CCL_THREADS) strip_rank_kernel(const int32_t* __restrict__ L, size_t lstep_words, int w,
                                                                 size_t n, long long base, int32_t* __restrict__ block_sums,
                                                                 int32_t* __restrict__ lrank)
{
    constexpr int SWEEPS = SCAN_CHUNK / CCL_THREADS;
    constexpr int WARPS = CCL_THREADS / 32;
    __shared__ unsigned s_ballot[SWEEPS * WARPS];
    __shared__ int s_prefix[SWEEPS * WARPS];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const size_t cbase = (size_t)blockIdx.x * SCAN_CHUNK;
#pragma unroll
    for (int k = 0; k < SWEEPS; k++) {
        size_t i = cbase + (size_t)k * CCL_THREADS + threadIdx.x;
        bool is_root = false;
        if (i < n) {
            int v = L[(i / w) * lstep_words + (i % w)];
            is_root = (long long)v == base + (long long)i + 1;
        }
        unsigned bal = __ballot_sync(0xffffffffu, is_root);
        if (lane == 0) s_ballot[k * WARPS + warp] = bal;
    }
    __syncthreads();
    if (warp == 0) {
        int carry = 0;
#pragma unroll
        for (int q = 0; q < SWEEPS * WARPS / 32; q++) {
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
    for (int k = 0; k < SWEEPS; k++) {
        unsigned bal = s_ballot[k * WARPS + warp];
        if ((bal >> lane) & 1u) {
            size_t i = cbase + (size_t)k * CCL_THREADS + threadIdx.x;
            lrank[i] = s_prefix[k * WARPS + warp] + __popc(bal & ((1u << lane) - 1));
        }
    }
}

// dense id of owned root labels: out[q] = offset + rank(label) + 1 (0 if the label is not an owned root position)
__global__ void __launch_bounds__(CCL_THREADS) strip_query_kernel(const int32_t* __restrict__ q, int nq, long long base, size_t n,
                                                                  int offset, const int32_t* __restrict__ block_offs,
                                                                  const int32_t* __restrict__ lrank, int32_t* __restrict__ out)
{
    int i = blockIdx.x * CCL_THREADS + threadIdx.x;
    if (i >= nq) return;
    long long loc = (long long)q[i] - 1 - base;
    out[i] = (loc >= 0 && loc < (long long)n) ? offset + block_offs[loc / SCAN_CHUNK] + lrank[loc] + 1 : 0;
}

__global__ void __launch_bounds__(CCL_THREADS) strip_apply_dense_kernel(int32_t* __restrict__ L, size_t lstep_words, int w, size_t n,
                                                                        long long base, int offset,
                                                                        const int32_t* __restrict__ block_offs,
                                                                        const int32_t* __restrict__ lrank,
                                                                        const int32_t* __restrict__ rlab,
                                                                        const int32_t* __restrict__ rdense, int nr)
{
    size_t i = (size_t)blockIdx.x * CCL_THREADS + threadIdx.x;
    if (i >= n) return;
    int32_t* p = L + (i / w) * lstep_words + (i % w);
    int v = *p;
    if (v <= 0) return;
    long long loc = (long long)v - 1 - base;
    if (loc >= 0 && loc < (long long)n) { *p = offset + block_offs[loc / SCAN_CHUNK] + lrank[loc] + 1; return; }
    int lo = 0, hi = nr - 1;                     // root owned by another strip: look its dense id up
    while (lo <= hi) {
        int mid = (lo + hi) >> 1;
        int f = __ldg(rlab + mid);
        if (f == v) { *p = __ldg(rdense + mid); return; }
        if (f < v) lo = mid + 1; else hi = mid - 1;
    }
}


// Seam equivalences with the strip-local ranks of both roots (single-exchange sharding): quads (A, B, rankA + 1, rankB + 1).
// up_rank1: rank + 1 of the upper labels' roots inside the upper strip (the rank above sends it with its boundary row);
// the lower labels are this strip's own: their ranks come from the tables msg_strip_rank_dev left in the workspace.
__global__ void __launch_bounds__(CCL_THREADS) seam_quads_kernel(const uint8_t* __restrict__ up_bgr, const int32_t* __restrict__ up_lab,
                                                                 const int32_t* __restrict__ up_rank1, const uint8_t* __restrict__ lo_bgr,
                                                                 const int32_t* __restrict__ lo_lab, int w, int d, long long base, size_t n,
                                                                 const int32_t* __restrict__ block_offs, const int32_t* __restrict__ lrank,
                                                                 int32_t* __restrict__ quads, int32_t* __restrict__ count)
{
    int x = blockIdx.x * CCL_THREADS + threadIdx.x;
    if (x >= w) return;
    auto close = [&](const uint8_t* a, const uint8_t* b) {
        return abs((int)a[0] - (int)b[0]) <= d && abs((int)a[1] - (int)b[1]) <= d && abs((int)a[2] - (int)b[2]) <= d;
    };
    if (!close(up_bgr + 3 * x, lo_bgr + 3 * x)) return;
    int a = up_lab[x], b = lo_lab[x];
    if (a == b) return;
    if (x > 0 && up_lab[x - 1] == a && lo_lab[x - 1] == b && close(up_bgr + 3 * (x - 1), lo_bgr + 3 * (x - 1))) return;
    long long loc = (long long)b - 1 - base;
    int rb = (loc >= 0 && loc < (long long)n) ? block_offs[loc / SCAN_CHUNK] + lrank[loc] + 1 : 0;
    int slot = atomicAdd(count, 1);
    quads[4 * slot] = a;
    quads[4 * slot + 1] = b;
    quads[4 * slot + 2] = up_rank1[x];
    quads[4 * slot + 3] = rb;
}

// Single pass from provisional strip labels (1 + global index of the strip-local root) to the dense global numbering:
// a label listed in `frm` (sorted; the labels that seam resolution merges into a smaller one) takes dense[j]; every other
// label is a surviving root of this strip: offset + its local rank - the number of removed roots of this strip before it.
__global__ void __launch_bounds__(CCL_THREADS) strip_finalize_kernel(int32_t* __restrict__ L, size_t lstep_words, int w, size_t n,
                                                                     long long base, int offset,
                                                                     const int32_t* __restrict__ block_offs,
                                                                     const int32_t* __restrict__ lrank,
                                                                     const int32_t* __restrict__ frm,
                                                                     const int32_t* __restrict__ dense, int nmap, int frm_lo)
{
    size_t i = (size_t)blockIdx.x * CCL_THREADS + threadIdx.x;
    if (i >= n) return;
    int32_t* p = L + (i / w) * lstep_words + (i % w);
    int v = *p;
    if (v <= 0) return;
    int lo = 0, hi = nmap;                       // lower bound of v in frm
    while (lo < hi) {
        int mid = (lo + hi) >> 1;
        if (__ldg(frm + mid) < v) lo = mid + 1; else hi = mid;
    }
    if (lo < nmap && __ldg(frm + lo) == v) { *p = __ldg(dense + lo); return; }
    long long loc = (long long)v - 1 - base;
    if (loc < 0 || loc >= (long long)n) return;   // not a label of this strip: left untouched (caller error)
    *p = offset + block_offs[loc / SCAN_CHUNK] + lrank[loc] - (lo - frm_lo) + 1;
}

inline unsigned blocks_for(size_t n, int per) { return (unsigned)((n + per - 1) / per); }

// dense int32 labels -> 16-bit labels, saturating (the caller reports MSG_ERANGE when the region count exceeds 65535)
__global__ void __launch_bounds__(CCL_THREADS) labels_to_u16_kernel(const int32_t* __restrict__ L, int w, uint16_t* __restrict__ dst,
                                                                   size_t dstep)
{
    int x = (blockIdx.x * CCL_THREADS + threadIdx.x) * 2;
    int y = blockIdx.y;
    if (x >= w) return;
    const int32_t* s = L + (size_t)y * w;
    uint16_t* d = (uint16_t*)((char*)dst + (size_t)y * dstep);
    int a = s[x];
    a = a < 0 ? 0 : (a > 65535 ? 65535 : a);
    if (x + 1 < w) {
        int b = s[x + 1];
        b = b < 0 ? 0 : (b > 65535 ? 65535 : b);
        if (((uintptr_t)(d + x) & 3) == 0) { *(uint32_t*)(d + x) = (uint32_t)a | ((uint32_t)b << 16); return; }
        d[x + 1] = (uint16_t)b;
    }
    d[x] = (uint16_t)a;
}

// ---------------------------------------------------------------- Canny hysteresis on top of the binary union-find
// cls: 0 none, 1 candidate, 2 strong candidate (k_seeds.cu).  L = union-find parents of the 8-connected candidate set.
__global__ void __launch_bounds__(CCL_THREADS) hyst_mark_kernel(const uint8_t* __restrict__ cls, const int32_t* __restrict__ L,
                                                                size_t n, uint8_t* __restrict__ flag)
{
    size_t i = (size_t)blockIdx.x * CCL_THREADS + threadIdx.x;
    if (i >= n || cls[i] != 2) return;
    flag[uf_find(L, (int)i)] = 1;
}

__global__ void __launch_bounds__(CCL_THREADS) hyst_out_kernel(const uint8_t* __restrict__ cls, const int32_t* __restrict__ L,
                                                               int w, const uint8_t* __restrict__ flag,
                                                               uint8_t* __restrict__ dst, size_t dstep)
{
    int x = blockIdx.x * CCL_THREADS + threadIdx.x, y = blockIdx.y;
    if (x >= w) return;
    size_t i = (size_t)y * w + x;
    dst[(size_t)y * dstep + x] = (cls[i] && flag[uf_find(L, (int)i)]) ? 255 : 0;
}

}  // namespace

// edges = candidates whose 8-connected component holds a strong candidate.  d_cls: w*h bytes (step w); d_labels: w*h int32
// scratch; d_flag: w*h bytes scratch.
int k_hysteresis(msg_ctx* ctx, const uint8_t* d_cls, int w, int h, int32_t* d_labels, uint8_t* d_flag, uint8_t* d_dst, size_t dstep)
{
    size_t n = (size_t)w * h;
    MSG_TRY(k_ccl_binary(ctx, d_cls, (size_t)w, w, h, 8, d_labels));
    MSG_CUDA(ctx, cudaMemsetAsync(d_flag, 0, n, ctx->stream));
    hyst_mark_kernel<<<(unsigned)((n + CCL_THREADS - 1) / CCL_THREADS), CCL_THREADS, 0, ctx->stream>>>(d_cls, d_labels, n, d_flag);
    MSG_LAUNCHED(ctx);
    dim3 grid((w + CCL_THREADS - 1) / CCL_THREADS, h);
    hyst_out_kernel<<<grid, CCL_THREADS, 0, ctx->stream>>>(d_cls, d_labels, w, d_flag, d_dst, dstep);
    MSG_LAUNCHED(ctx);
    MSG_CHECK_LAUNCH(ctx);
    return MSG_OK;
}

// scratch layout for relabel: [rank: n+1 ints][first: n+1 ints (MODE 1)][block_sums: nb ints]
static int relabel_impl(msg_ctx* ctx, int32_t* d_labels, size_t n, int mode, int32_t* d_n_out, int add_to_total)
{
    int nb = (int)blocks_for(n, SCAN_CHUNK);
    size_t need = ((n + 1) * (mode ? 2 : 1) + (size_t)nb + 64) * sizeof(int32_t);
    MSG_TRY(msg_reserve(ctx, (void**)&ctx->d_scratch, &ctx->d_scratch_cap, need));
    int32_t* rank = (int32_t*)ctx->d_scratch;
    int32_t* first = mode ? rank + (n + 1) : nullptr;
    int32_t* block_sums = rank + (n + 1) * (mode ? 2 : 1);
    cudaStream_t st = ctx->stream;
    if (!mode) {   // labels are union-find parents (pixel indices): flatten, count, rank, apply = 3 launches
        flatten_count_kernel<<<nb, CCL_THREADS, 0, st>>>(d_labels, n, block_sums, rank);
        MSG_LAUNCHED(ctx);
        scan_block_sums_kernel<<<1, CCL_THREADS, 0, st>>>(block_sums, nb, d_n_out, add_to_total);
        MSG_LAUNCHED(ctx);
        apply_lrank_kernel<<<blocks_for(n, CCL_THREADS), CCL_THREADS, 0, st>>>(d_labels, n, block_sums, rank);
        MSG_LAUNCHED(ctx);
        MSG_CHECK_LAUNCH(ctx);
        return MSG_OK;
    }
    fill_i32_kernel<<<blocks_for(n + 1, CCL_THREADS), CCL_THREADS, 0, st>>>(first, n + 1, 0x7fffffff);
    MSG_LAUNCHED(ctx);
    first_pixel_kernel<<<blocks_for(n, CCL_THREADS), CCL_THREADS, 0, st>>>(d_labels, n, first);
    MSG_LAUNCHED(ctx);
    count_first_kernel<1><<<nb, CCL_THREADS, 0, st>>>(d_labels, first, n, block_sums);
    MSG_LAUNCHED(ctx);
    scan_block_sums_kernel<<<1, CCL_THREADS, 0, st>>>(block_sums, nb, d_n_out, add_to_total);
    MSG_LAUNCHED(ctx);
    assign_rank_kernel<1><<<nb, CCL_THREADS, 0, st>>>(d_labels, first, n, block_sums, rank);
    MSG_LAUNCHED(ctx);
    apply_rank_kernel<<<blocks_for(n, CCL_THREADS), CCL_THREADS, 0, st>>>(d_labels, n, rank, mode);
    MSG_LAUNCHED(ctx);
    MSG_CHECK_LAUNCH(ctx);
    return MSG_OK;
}

int k_relabel_canonical(msg_ctx* ctx, int32_t* d_labels, int w, int h, int roots_are_pixels, int32_t* d_n_out,
                        int add_to_count)
{
    return relabel_impl(ctx, d_labels, (size_t)w * h, roots_are_pixels ? 0 : 1, d_n_out, add_to_count);
}

// ---- tile-local union-find launchers.  Scratch of the stage lives in ctx->d_ccl (callers keep their label arrays in
// d_scratch / d_labels): [L: n ints (canonical path only)][root bitmap: h * wp words][block sums]
struct ccl_ws { int32_t* L; uint32_t* bitmap; uint32_t* rootmap; int32_t* block_sums; int wp; size_t nwords; int nblocks; };

static int ccl_workspace(msg_ctx* ctx, int w, int h, bool with_labels, ccl_ws* ws)
{
    const size_t n = (size_t)w * h;
    ws->wp = (w + 31) / 32;
    ws->nwords = (size_t)ws->wp * h;
    ws->nblocks = (int)((ws->nwords + RB_BLOCK - 1) / RB_BLOCK);
    const size_t lab_bytes = with_labels ? ((n * 4 + 255) & ~(size_t)255) : 0;
    const size_t bm_bytes = (ws->nwords * 4 + 255) & ~(size_t)255;
    MSG_TRY(msg_reserve(ctx, (void**)&ctx->d_ccl, &ctx->d_ccl_cap, lab_bytes + 2 * bm_bytes + ((size_t)ws->nblocks + 64) * 4));
    ws->L = with_labels ? (int32_t*)ctx->d_ccl : nullptr;
    ws->bitmap = (uint32_t*)(ctx->d_ccl + lab_bytes);
    ws->rootmap = (uint32_t*)(ctx->d_ccl + lab_bytes + bm_bytes);
    ws->block_sums = (int32_t*)(ctx->d_ccl + lab_bytes + 2 * bm_bytes);
    return MSG_OK;
}

// phases A + B: union-find parents in L (roots = first pixels of the components, background -1), root bitmap in ws
template <int PRED>
static int ccl_tiles(msg_ctx* ctx, const void* img, size_t pitch, int w, int h, int d, int conn, int32_t* L, const ccl_ws& ws)
{
    cudaStream_t st = ctx->stream;
    dim3 grid((w + CT_W - 1) / CT_W, (h + CT_H - 1) / CT_H);
    const long long n_hb = (long long)(grid.y - 1) * w, n_vb = (long long)(grid.x - 1) * h;
    const unsigned bblocks = (unsigned)((n_hb + n_vb + CT_THREADS - 1) / CT_THREADS);
    // four pixels per lane when the rows allow aligned vector accesses (width a multiple of 4; plane: 16-byte aligned rows,
    // BGR: 4-byte aligned rows); option ccl_quad = 0 keeps the one-pixel-per-lane kernel (A/B hook)
    bool quad = PRED != 1 && conn == 4 && ctx->tune.ccl_quad && w % 4 == 0 && (reinterpret_cast<uintptr_t>(L) & 15) == 0;
    if (PRED == 0) quad = quad && pitch % 4 == 0 && (reinterpret_cast<uintptr_t>(img) & 15) == 0;
    if (PRED == 2) quad = quad && pitch % 4 == 0 && (reinterpret_cast<uintptr_t>(img) & 3) == 0;
    if (quad) ccl_tile4_kernel<PRED == 1 ? 0 : PRED><<<grid, CT_THREADS, 0, st>>>(img, pitch, w, h, d, L, ws.bitmap, ws.wp);
    else if (conn == 8) ccl_tile_kernel<PRED, 8><<<grid, CT_THREADS, 0, st>>>(img, pitch, w, h, d, L, ws.bitmap, ws.wp);
    else ccl_tile_kernel<PRED, 4><<<grid, CT_THREADS, 0, st>>>(img, pitch, w, h, d, L, ws.bitmap, ws.wp);
    MSG_LAUNCHED(ctx);
    if (bblocks) {
        if (conn == 8) ccl_border_kernel<PRED, 8><<<bblocks, CT_THREADS, 0, st>>>(img, pitch, w, h, d, L, n_hb, n_vb);
        else ccl_border_kernel<PRED, 4><<<bblocks, CT_THREADS, 0, st>>>(img, pitch, w, h, d, L, n_hb, n_vb);
        MSG_LAUNCHED(ctx);
    }
    MSG_CHECK_LAUNCH(ctx);
    return MSG_OK;
}

// phases C .. F: canonical labels 1..n (raster order of first pixel) from the forest in ws.L into d_out; *d_n = n + add
static int ccl_canonical(msg_ctx* ctx, int w, int h, const ccl_ws& ws, int32_t* d_out, int32_t* d_n, int add_to_count)
{
    cudaStream_t st = ctx->stream;
    const size_t n = (size_t)w * h;
    ccl_roots_kernel<<<ws.nblocks, CT_THREADS, 0, st>>>(ws.L, ws.bitmap, ws.rootmap, w, ws.wp, ws.nwords, ws.block_sums);
    MSG_LAUNCHED(ctx);
    scan_sums_kernel<<<1, 1024, 0, st>>>(ws.block_sums, ws.nblocks, (ws.nblocks + 1023) / 1024, d_n, add_to_count);
    MSG_LAUNCHED(ctx);
    ccl_rank_kernel<<<ws.nblocks, CT_THREADS, 0, st>>>(ws.L, ws.rootmap, w, ws.wp, ws.nwords, ws.block_sums);
    MSG_LAUNCHED(ctx);
    ccl_spread_kernel<<<blocks_for(ws.nwords, CT_THREADS), CT_THREADS, 0, st>>>(ws.L, ws.bitmap, ws.rootmap, w, ws.wp, ws.nwords);
    MSG_LAUNCHED(ctx);
    ccl_final_kernel<<<blocks_for((n + 3) / 4, CT_THREADS), CT_THREADS, 0, st>>>(ws.L, d_out, n);
    MSG_LAUNCHED(ctx);
    MSG_CHECK_LAUNCH(ctx);
    return MSG_OK;
}

// label_base < 0: union-find parents (the caller relabels); otherwise strip mode: 1 + global index of the strip-local root
int k_ccl_color(msg_ctx* ctx, const uint32_t* d_plane, int pitch, int w, int h, int d, int conn, int32_t* d_labels,
                int64_t label_base, int full_w)
{
    dim3 grid((w + CCL_THREADS - 1) / CCL_THREADS, h);
    size_t n = (size_t)w * h;
    cudaStream_t st = ctx->stream;
    if (w > CCL_MAX_CHUNKS * 32) return msg_fail(ctx, MSG_EINVAL, "labelling supports rows up to %d pixels", CCL_MAX_CHUNKS * 32);
    if (!ctx->tune.ccl_legacy) {
        ccl_ws ws;
        MSG_TRY(ccl_workspace(ctx, w, h, false, &ws));
        MSG_TRY(ccl_tiles<0>(ctx, d_plane, (size_t)pitch, w, h, d, conn, d_labels, ws));
    } else {
        ccl_rows_kernel<0><<<h, CCL_THREADS, 0, st>>>(d_plane, (size_t)pitch, w, h, d, d_labels);
        MSG_LAUNCHED(ctx);
        if (conn == 8) ccl_merge_kernel<0, 8><<<grid, CCL_THREADS, 0, st>>>(d_plane, (size_t)pitch, w, h, d, d_labels);
        else ccl_merge_kernel<0, 4><<<grid, CCL_THREADS, 0, st>>>(d_plane, (size_t)pitch, w, h, d, d_labels);
        MSG_LAUNCHED(ctx);
    }
    if (label_base >= 0) {      // strip mode: no canonical relabel follows, flatten here
        ccl_flatten_kernel<<<blocks_for(n, CCL_THREADS), CCL_THREADS, 0, st>>>(d_labels, n);
        MSG_LAUNCHED(ctx);
        add_label_base_kernel<<<blocks_for(n, CCL_THREADS), CCL_THREADS, 0, st>>>(d_labels, n, w, full_w, label_base);
        MSG_LAUNCHED(ctx);
    }
    MSG_CHECK_LAUNCH(ctx);
    return MSG_OK;
}

int k_ccl_binary(msg_ctx* ctx, const uint8_t* d_mask, size_t step, int w, int h, int conn, int32_t* d_labels)
{
    dim3 grid((w + CCL_THREADS - 1) / CCL_THREADS, h);
    cudaStream_t st = ctx->stream;
    if (w > CCL_MAX_CHUNKS * 32) return msg_fail(ctx, MSG_EINVAL, "labelling supports rows up to %d pixels", CCL_MAX_CHUNKS * 32);
    if (!ctx->tune.ccl_legacy) {
        ccl_ws ws;
        MSG_TRY(ccl_workspace(ctx, w, h, false, &ws));
        return ccl_tiles<1>(ctx, d_mask, step, w, h, 0, conn, d_labels, ws);
    }
    ccl_rows_kernel<1><<<h, CCL_THREADS, 0, st>>>(d_mask, step, w, h, 0, d_labels);
    MSG_LAUNCHED(ctx);
    if (conn == 8) ccl_merge_kernel<1, 8><<<grid, CCL_THREADS, 0, st>>>(d_mask, step, w, h, 0, d_labels);
    else ccl_merge_kernel<1, 4><<<grid, CCL_THREADS, 0, st>>>(d_mask, step, w, h, 0, d_labels);
    MSG_LAUNCHED(ctx);
    MSG_CHECK_LAUNCH(ctx);
    return MSG_OK;
}

// union-find parents -> roots (label = linear index of the component's first pixel, background stays negative)
int k_ccl_flatten(msg_ctx* ctx, int32_t* d_labels, size_t n)
{
    ccl_flatten_kernel<<<blocks_for(n, CCL_THREADS), CCL_THREADS, 0, ctx->stream>>>(d_labels, n);
    MSG_LAUNCHED(ctx);
    MSG_CHECK_LAUNCH(ctx);
    return MSG_OK;
}

int k_copy_labels_2d(msg_ctx* ctx, const int32_t* src, size_t sstep, int32_t* dst, size_t dstep, int w, int h)
{
    dim3 grid((w + CCL_THREADS - 1) / CCL_THREADS, h);
    copy_labels_2d_kernel<<<grid, CCL_THREADS, 0, ctx->stream>>>(src, sstep, dst, dstep, w);
    MSG_LAUNCHED(ctx);
    MSG_CHECK_LAUNCH(ctx);
    return MSG_OK;
}

int k_seam_pairs(msg_ctx* ctx, const uint8_t* up_bgr, const int32_t* up_lab, const uint8_t* lo_bgr,
                 const int32_t* lo_lab, int w, int d, int32_t* pairs, int32_t* count)
{
    MSG_CUDA(ctx, cudaMemsetAsync(count, 0, sizeof(int32_t), ctx->stream));
    seam_pairs_kernel<<<(w + CCL_THREADS - 1) / CCL_THREADS, CCL_THREADS, 0, ctx->stream>>>(up_bgr, up_lab, lo_bgr, lo_lab,
                                                                                           w, d, pairs, count);
    MSG_LAUNCHED(ctx);
    MSG_CHECK_LAUNCH(ctx);
    return MSG_OK;
}

int k_apply_map(msg_ctx* ctx, int32_t* labels, size_t lstep, int w, int rows, const int32_t* from, const int32_t* to,
                int n)
{
    if (n <= 0) return MSG_OK;
    dim3 grid((w + CCL_THREADS - 1) / CCL_THREADS, rows);
    apply_map_kernel<<<grid, CCL_THREADS, 0, ctx->stream>>>(labels, lstep, w, from, to, n);
    MSG_LAUNCHED(ctx);
    MSG_CHECK_LAUNCH(ctx);
    return MSG_OK;
}

// ---- strip dense numbering launchers (scratch layout: [lrank: n ints][block_sums: nb ints]; kept between the calls)
int k_strip_rank(msg_ctx* ctx, const int32_t* d_labels, size_t lstep, int w, int rows, long long base, int32_t* d_count)
{
    size_t n = (size_t)w * rows;
    int nb = (int)blocks_for(n, SCAN_CHUNK);
    MSG_TRY(msg_reserve(ctx, (void**)&ctx->d_scratch, &ctx->d_scratch_cap, (n + (size_t)nb + 64) * sizeof(int32_t)));
    int32_t* lrank = (int32_t*)ctx->d_scratch;
    int32_t* block_sums = lrank + n;
    strip_rank_kernel<<<nb, CCL_THREADS, 0, ctx->stream>>>(d_labels, lstep / 4, w, n, base, block_sums, lrank);
    MSG_LAUNCHED(ctx);
    scan_block_sums_kernel<<<1, CCL_THREADS, 0, ctx->stream>>>(block_sums, nb, d_count, 0);
    MSG_LAUNCHED(ctx);
    MSG_CHECK_LAUNCH(ctx);
    return MSG_OK;
}

int k_strip_query(msg_ctx* ctx, const int32_t* d_q, int nq, int w, int rows, long long base, int offset, int32_t* d_out)
{
    if (nq <= 0) return MSG_OK;
    size_t n = (size_t)w * rows;
    int32_t* lrank = (int32_t*)ctx->d_scratch;
    strip_query_kernel<<<blocks_for((size_t)nq, CCL_THREADS), CCL_THREADS, 0, ctx->stream>>>(d_q, nq, base, n, offset, lrank + n, lrank, d_out);
    MSG_LAUNCHED(ctx);
    MSG_CHECK_LAUNCH(ctx);
    return MSG_OK;
}

int k_strip_apply_dense(msg_ctx* ctx, int32_t* d_labels, size_t lstep, int w, int rows, long long base, int offset,
                        const int32_t* d_rlab, const int32_t* d_rdense, int nr)
{
    size_t n = (size_t)w * rows;
    int32_t* lrank = (int32_t*)ctx->d_scratch;
    strip_apply_dense_kernel<<<blocks_for(n, CCL_THREADS), CCL_THREADS, 0, ctx->stream>>>(d_labels, lstep / 4, w, n, base, offset,
                                                                                          lrank + n, lrank, d_rlab, d_rdense, nr);
    MSG_LAUNCHED(ctx);
    MSG_CHECK_LAUNCH(ctx);
    return MSG_OK;
}

int k_seam_quads(msg_ctx* ctx, const uint8_t* up_bgr, const int32_t* up_lab, const int32_t* up_rank1, const uint8_t* lo_bgr,
                 const int32_t* lo_lab, int w, int d, int rows, long long base, int32_t* quads, int32_t* count)
{
    size_t n = (size_t)w * rows;
    int32_t* lrank = (int32_t*)ctx->d_scratch;
    MSG_CUDA(ctx, cudaMemsetAsync(count, 0, sizeof(int32_t), ctx->stream));
    seam_quads_kernel<<<blocks_for((size_t)w, CCL_THREADS), CCL_THREADS, 0, ctx->stream>>>(up_bgr, up_lab, up_rank1, lo_bgr, lo_lab, w, d,
                                                                                          base, n, lrank + n, lrank, quads, count);
    MSG_LAUNCHED(ctx);
    MSG_CHECK_LAUNCH(ctx);
    return MSG_OK;
}

int k_strip_finalize(msg_ctx* ctx, int32_t* d_labels, size_t lstep, int w, int rows, long long base, int offset,
                     const int32_t* d_frm, const int32_t* d_dense, int nmap, int frm_lo)
{
    size_t n = (size_t)w * rows;
    int32_t* lrank = (int32_t*)ctx->d_scratch;
    strip_finalize_kernel<<<blocks_for(n, CCL_THREADS), CCL_THREADS, 0, ctx->stream>>>(d_labels, lstep / 4, w, n, base, offset, lrank + n,
                                                                                       lrank, d_frm, d_dense, nmap, frm_lo);
    MSG_LAUNCHED(ctx);
    MSG_CHECK_LAUNCH(ctx);
    return MSG_OK;
}

int k_labels_to_u16(msg_ctx* ctx, const int32_t* d_labels, int w, int h, uint16_t* d_dst, size_t dstep)
{
    dim3 grid(((w + 1) / 2 + CCL_THREADS - 1) / CCL_THREADS, h);
    labels_to_u16_kernel<<<grid, CCL_THREADS, 0, ctx->stream>>>(d_labels, w, d_dst, dstep);
    MSG_LAUNCHED(ctx);
    MSG_CHECK_LAUNCH(ctx);
    return MSG_OK;
}

// Colour-predicate labelling with canonical numbering in one call (the fused pipeline's label stage): phases A .. F.
// src_kind 0: packed plane (pitch in pixels); 2: BGR bytes (pitch in bytes) -- the stand-alone operator needs no plane.
int k_label_canonical_src(msg_ctx* ctx, const void* d_img, size_t pitch, int src_kind, int w, int h, int d, int conn,
                          int32_t* d_labels, int32_t* d_n)
{
    if (w > CCL_MAX_CHUNKS * 32) return msg_fail(ctx, MSG_EINVAL, "labelling supports rows up to %d pixels", CCL_MAX_CHUNKS * 32);
    ccl_ws ws;
    MSG_TRY(ccl_workspace(ctx, w, h, true, &ws));
    if (src_kind == 2) MSG_TRY(ccl_tiles<2>(ctx, d_img, pitch, w, h, d, conn, ws.L, ws));
    else MSG_TRY(ccl_tiles<0>(ctx, d_img, pitch, w, h, d, conn, ws.L, ws));
    return ccl_canonical(ctx, w, h, ws, d_labels, d_n, 0);
}

int k_label_canonical(msg_ctx* ctx, const uint32_t* d_plane, int pitch, int w, int h, int d, int conn, int32_t* d_labels,
                      int32_t* d_n)
{
    if (ctx->tune.ccl_legacy) {
        MSG_TRY(k_ccl_color(ctx, d_plane, pitch, w, h, d, conn, d_labels, -1, w));
        return k_relabel_canonical(ctx, d_labels, w, h, 1, d_n, 0);
    }
    return k_label_canonical_src(ctx, d_plane, (size_t)pitch, 0, w, h, d, conn, d_labels, d_n);
}

// Imgproc.connectedComponents: binary predicate, canonical numbering; *d_n = number of labels INCLUDING the background
int k_cc_canonical(msg_ctx* ctx, const uint8_t* d_mask, size_t step, int w, int h, int conn, int32_t* d_labels, int32_t* d_n)
{
    if (ctx->tune.ccl_legacy) {
        MSG_TRY(k_ccl_binary(ctx, d_mask, step, w, h, conn, d_labels));
        return k_relabel_canonical(ctx, d_labels, w, h, 1, d_n, 1);
    }
    if (w > CCL_MAX_CHUNKS * 32) return msg_fail(ctx, MSG_EINVAL, "labelling supports rows up to %d pixels", CCL_MAX_CHUNKS * 32);
    ccl_ws ws;
    MSG_TRY(ccl_workspace(ctx, w, h, true, &ws));
    MSG_TRY(ccl_tiles<1>(ctx, d_mask, step, w, h, 0, conn, ws.L, ws));
    return ccl_canonical(ctx, w, h, ws, d_labels, d_n, 1);
}
