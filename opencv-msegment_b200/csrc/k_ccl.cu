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

// label_base < 0: canonical labels 1..n (n -> d_n via caller's relabel); otherwise strip mode
int k_ccl_color(msg_ctx* ctx, const uint32_t* d_plane, int pitch, int w, int h, int d, int conn, int32_t* d_labels,
                int64_t label_base, int full_w)
{
    dim3 grid((w + CCL_THREADS - 1) / CCL_THREADS, h);
    size_t n = (size_t)w * h;
    cudaStream_t st = ctx->stream;
    if (w > CCL_MAX_CHUNKS * 32) return msg_fail(ctx, MSG_EINVAL, "labelling supports rows up to %d pixels", CCL_MAX_CHUNKS * 32);
    ccl_rows_kernel<0><<<h, CCL_THREADS, 0, st>>>(d_plane, (size_t)pitch, w, h, d, d_labels);
    MSG_LAUNCHED(ctx);
    if (conn == 8) ccl_merge_kernel<0, 8><<<grid, CCL_THREADS, 0, st>>>(d_plane, (size_t)pitch, w, h, d, d_labels);
    else ccl_merge_kernel<0, 4><<<grid, CCL_THREADS, 0, st>>>(d_plane, (size_t)pitch, w, h, d, d_labels);
    MSG_LAUNCHED(ctx);
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
    size_t n = (size_t)w * h;
    cudaStream_t st = ctx->stream;
    if (w > CCL_MAX_CHUNKS * 32) return msg_fail(ctx, MSG_EINVAL, "labelling supports rows up to %d pixels", CCL_MAX_CHUNKS * 32);
    (void)n;
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

int k_label_canonical(msg_ctx* ctx, const uint32_t* d_plane, int pitch, int w, int h, int d, int conn, int32_t* d_labels,
                      int32_t* d_n)
{
    MSG_TRY(k_ccl_color(ctx, d_plane, pitch, w, h, d, conn, d_labels, -1, w));
    return k_relabel_canonical(ctx, d_labels, w, h, 1, d_n, 0);
}
