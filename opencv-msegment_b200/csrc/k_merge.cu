// k_merge.cu -- K2b region statistics + merge rounds, K2c label rendering.
//
// K2b has no OpenCV counterpart; its specification is DESIGN.md "K2b" (restated by the oracle's
// orc_merge_regions): per round every participating region selects the 4-adjacent region with the
// smallest (dist2 of rounded mean colours, label) key, accepted selections are united
// simultaneously (union by smallest label); phase A = colour fuse, phase B = min-size prune.
// K2c = PictureService.colorByIndexes (PictureService.java:913-936).
//
// Launch sequence of one merge (the label count lives on the device, so every rounds kernel is always launched and the
// ones whose regime does not apply return at once):
//   merge_init_kernel             tables of the n_in labels
//   merge_stats_kernel            one pass over the pixels: area + colour sums per label, raw list of adjacent label pairs
//   merge_pair_set_kernel         (images of <= 2^24 pixels) the SET of adjacent pairs: global hash table + unique list
//   merge_rounds_small_kernel<12> <= 4095 labels (a 1080p frame): ONE CTA, every table in shared memory, __syncthreads
//                                 between the passes of a round, then the dense renumbering
//   merge_rounds_small_kernel<13> <= 8191 labels (a 4K frame): the same with 13-bit labels, 224 KB of shared memory
//   merge_rounds_large_kernel     more labels: cooperative grid, tables in L2/HBM, grid-wide barriers, same passes; builds
//                                 the pair set itself beyond 2^24 pixels and in the strip-sharded merge
//   merge_rewrite_kernel          one pass over the pixels: label -> final id
#include <cooperative_groups.h>
#include <stdlib.h>

#include "msg_internal.h"

namespace cg = cooperative_groups;

namespace {

constexpr int MT = 256;
constexpr int MERGE_MAX_ROUNDS = 64;
constexpr size_t MEDIUM_SMEM = (size_t)8192 * 7 * sizeof(uint32_t);
constexpr int MEDIUM_MAX_LABELS = 8191;         // medium path: 13-bit label field, every table of the CTA in 224 KB of shared memory
constexpr int SMALL_MAX_LABELS = 4095;          // small path: labels 1..4095 -> 12-bit label field in the 32-bit selection key
constexpr long long SMALL_MAX_PIXELS = 1ll << 24;   // small path: colour sums of a region fit 32 bits (255 * 2^24 < 2^32)
constexpr int ST = 1024;                        // threads of the small-path CTA

struct merge_tables {
    unsigned int* area;            // [nl]
    unsigned long long* sum;       // [3*nl]  B,G,R
    uint32_t* mean;                // [nl]   packed B | G<<8 | R<<16; reused as the final-id table after the rounds
    unsigned long long* best;      // [nl]   (dist2 << 32) | neighbour label
    int32_t* par;                  // [nl]
};

__global__ void __launch_bounds__(MT) max_label_kernel(const int32_t* __restrict__ L, size_t n, int32_t* __restrict__ out)
{
    int m = 0;
    for (size_t i = (size_t)blockIdx.x * MT + threadIdx.x; i < n; i += (size_t)gridDim.x * MT) m = max(m, L[i]);
#pragma unroll
    for (int o = 16; o; o >>= 1) m = max(m, __shfl_xor_sync(0xffffffffu, m, o));
    if ((threadIdx.x & 31) == 0 && m > 0) atomicMax(out, m);
}

__device__ __forceinline__ int pf_find(const int32_t* P, int a)
{
    int p = __ldcg(P + a);
    while (p != a) { a = p; p = __ldcg(P + a); }
    return a;
}

struct merge_persist_args {
    const uint32_t* plane; int pitch;
    int32_t* labels; int w, h;
    const int32_t* n_in;      // device: number of input labels (labels are 1..*n_in, numbered by first pixel)
    int32_t* n_saved;         // device: copy of *n_in taken by the init kernel (n_out may alias n_in)
    int cap;                  // table capacity in labels
    merge_tables t;
    int32_t* newid;           // [cap+1]
    int32_t* bsum;            // [gridDim.x + 1] of the large-path kernel
    int32_t* accepted;        // device counter
    int32_t* n_out;           // device: number of regions after the merge
    int32_t* rounds_out;      // device: rounds executed (statistics)
    int2* pairs;              // [pair_cap] adjacent (label, label) pairs found by the statistics pass
    int32_t* npairs;          // device counter
    long long pair_cap;
    int min_size, color_dist;
    int vec;                  // w % 4 == 0 and labels 16-byte aligned -> 16-byte loads in the statistics pass: 2 = column strips, 1 = row chunks
    int small_max;            // labels up to which the single-CTA rounds kernel with the shared-memory pair set is used
    int medium_max;           // labels up to which the single-CTA rounds kernel with the global pair set is used (0 / 0: large path)
    int nin_host;             // >= 0: the label count is known on the host (strip-sharded merge) and overrides *n_saved
    // The raw pair list holds one entry per boundary corner (10-15 x more than there are adjacent region pairs) and every round
    // walks the pairs, so the SET of pairs is built once in a global hash table: by merge_pair_set_kernel (whole GPU, set_ready = 1:
    // unsharded merge of a frame-sized image) or by the cooperative rounds kernel itself (images beyond 2^24 pixels, strip-sharded merge).
    unsigned long long* htab; // [1 << hbits] open-addressing set, key = min label << 32 | max label (0 = empty); zeroed by the host
    int hbits;
    int2* uniq;               // [1 << hbits] unique pairs in insertion order
    int32_t* nuniq;           // device: [0] entries of uniq, [1] overflow flag (table full: the rounds use the raw list)
    int set_ready;            // 1: merge_pair_set_kernel ran before the rounds kernels (unsharded merge of a frame-sized image)
    long long npairs_host;    // >= 0: number of valid entries of `pairs` (all-gathered list), overrides *npairs
};

__device__ __forceinline__ int merge_nin(const merge_persist_args& A)
{
    int nin = A.nin_host >= 0 ? A.nin_host : *A.n_saved;
    return nin > A.cap ? A.cap : nin;
}

__device__ __forceinline__ long long merge_npairs(const merge_persist_args& A)
{
    long long np = A.npairs_host >= 0 ? A.npairs_host : (long long)*A.npairs;
    return np > A.pair_cap ? A.pair_cap : np;
}

// Which rounds kernel serves this image: 0 = single CTA, <= 4095 labels, pair set in shared memory; 1 = single CTA, <= 8191
// labels, pair set in a global hash table (a 4K frame of the bench: 5.5 k regions); 2 = cooperative grid.  Uniform over a launch.
__device__ __forceinline__ int merge_regime(const merge_persist_args& A, int nin)
{
    if ((long long)A.w * A.h > SMALL_MAX_PIXELS) return 2;       // colour sums of a region must fit 32 bits
    if (nin <= A.small_max) return 0;
    if (nin <= A.medium_max) return 1;
    return 2;
}

__device__ __forceinline__ void persist_union(int32_t* par, int a, int b)
{
    for (;;) {
        a = pf_find(par, a);
        b = pf_find(par, b);
        if (a == b) return;
        if (a < b) { int s = a; a = b; b = s; }
        int old = atomicMin(par + a, b);
        if (old == a) return;
        a = old;
    }
}

// ---------------------------------------------------------------- init
__global__ void __launch_bounds__(MT) merge_init_kernel(merge_persist_args A)
{
    int nin = *A.n_in;
    if (nin > A.cap) nin = A.cap;
    const int nl = nin + 1;
    const long long gtid = (long long)blockIdx.x * MT + threadIdx.x, nthreads = (long long)gridDim.x * MT;
    for (long long i = gtid; i < nl; i += nthreads) {
        A.t.area[i] = 0;
        A.t.sum[3 * i] = 0; A.t.sum[3 * i + 1] = 0; A.t.sum[3 * i + 2] = 0;
        A.t.par[i] = (int)i;
    }
    if (gtid == 0) {
        *A.accepted = 0; *A.rounds_out = 0; *A.npairs = 0; *A.n_saved = nin;
        if (A.nuniq) { A.nuniq[0] = 0; A.nuniq[1] = 0; }
    }
}

// ---------------------------------------------------------------- statistics + adjacency pairs (one pass over the pixels)
// add one run (label, pixel count, colour sums) to the tables
__device__ __forceinline__ void stats_flush(const merge_tables& t, int lab, unsigned cnt, unsigned b, unsigned g, unsigned r)
{
    atomicAdd(t.area + lab, cnt);
    atomicAdd(t.sum + 3 * (size_t)lab, (unsigned long long)b);
    atomicAdd(t.sum + 3 * (size_t)lab + 1, (unsigned long long)g);
    atomicAdd(t.sum + 3 * (size_t)lab + 2, (unsigned long long)r);
}

// insert an adjacent pair into the global set; true if it was not there yet (the caller appends it to A.uniq)
__device__ __forceinline__ bool pair_set_insert(const merge_persist_args& A, int a, int b)
{
    const unsigned lo = (unsigned)min(a, b), hi = (unsigned)max(a, b);
    const unsigned long long key = ((unsigned long long)lo << 32) | hi;
    const unsigned long long hmask = (1ull << A.hbits) - 1;
    unsigned long long slot = (key * 0x9E3779B97F4A7C15ull) >> (64 - A.hbits);
    for (int probes = 0; probes < 64; probes++) {
        const unsigned long long old = atomicCAS(A.htab + slot, 0ull, key);
        if (old == 0ull) return true;
        if (old == key) return false;
        slot = (slot + 1) & hmask;
    }
    A.nuniq[1] = 1;                                        // table too full for this image: the rounds walk the raw list
    return false;
}

// Statistics + adjacency pass, 16-byte loads (requires w % 4 == 0 and 16-byte aligned label rows): a warp walks the image
// as a flat array in chunks of 128 pixels, 4 consecutive pixels per lane.  Runs of equal labels are summed inside the
// lane, then across lanes (segmented shuffle reduction of every lane's last run), so a region costs a few atomics per
// chunk.  Adjacent-pair list: a right pair is skipped when the row above holds the same pair, a down pair when the pixel
// to the left holds the same pair (the first occurrence is always emitted), which keeps roughly one entry per boundary
// corner instead of one per boundary pixel.
__device__ __forceinline__ void stats_pass_vec4(const merge_persist_args& A, const merge_tables& t, int nin, int lane,
                                                long long gwarp, long long nwarps)
{
    const int w = A.w, h = A.h;
    const long long n = (long long)w * h;
    const long long nchunks = (n + 127) / 128;
    const unsigned FULL = 0xffffffffu;
    for (long long c = gwarp; c < nchunks; c += nwarps) {
        const long long p = c * 128 + lane * 4;
        const bool in = p < n;
        int y = 0, x = 0;
        int4 L4 = make_int4(0, 0, 0, 0), D4 = L4, U4 = L4;
        uint4 C4 = make_uint4(0, 0, 0, 0);
        if (in) {
            y = (int)(p / w); x = (int)(p - (long long)y * w);
            L4 = *reinterpret_cast<const int4*>(A.labels + p);
            C4 = __ldg(reinterpret_cast<const uint4*>(A.plane + (size_t)y * A.pitch + x));
            if (y + 1 < h) D4 = *reinterpret_cast<const int4*>(A.labels + p + w);
            if (y > 0) U4 = *reinterpret_cast<const int4*>(A.labels + p - w);
        }
        int lab[4] = {L4.x, L4.y, L4.z, L4.w}, dn[4] = {D4.x, D4.y, D4.z, D4.w}, up[5] = {U4.x, U4.y, U4.z, U4.w, 0};
        const uint32_t col[4] = {C4.x, C4.y, C4.z, C4.w};
#pragma unroll
        for (int k = 0; k < 4; k++) {
            if (lab[k] > nin || lab[k] < 0) lab[k] = 0;
            if (dn[k] > nin || dn[k] < 0) dn[k] = 0;
        }
        // label right of my 4th pixel (and of the pixel above it): the next lane's first pixel, or one scalar load
        int nx = __shfl_down_sync(FULL, lab[0], 1), nu = __shfl_down_sync(FULL, up[0], 1);
        const bool has_right = in && x + 4 < w;
        if (lane == 31 && has_right) {
            nx = A.labels[p + 4];
            if (nx > nin || nx < 0) nx = 0;
            nu = y > 0 ? A.labels[p + 4 - w] : 0;
        }
        if (!has_right) nx = 0;
        up[4] = nu;
        // ---- sums: runs inside the lane, the last run continues into the next lanes
        int cur = lab[0];
        unsigned cnt = cur > 0, b = 0, g = 0, r = 0;
        if (cur > 0) { b = col[0] & 0xFF; g = (col[0] >> 8) & 0xFF; r = (col[0] >> 16) & 0xFF; }
        bool uniform = true;
#pragma unroll
        for (int k = 1; k < 4; k++) {
            if (lab[k] != cur) {
                if (cur > 0) stats_flush(t, cur, cnt, b, g, r);
                cur = lab[k]; cnt = 0; b = g = r = 0; uniform = false;
            }
            if (cur > 0) { cnt++; b += col[k] & 0xFF; g += (col[k] >> 8) & 0xFF; r += (col[k] >> 16) & 0xFF; }
        }
        int prev = __shfl_up_sync(FULL, cur, 1);
        bool head = lane == 0 || !uniform || prev != cur;
        unsigned heads = __ballot_sync(FULL, head);
        unsigned above = lane == 31 ? 0u : (heads >> (lane + 1));
        int seg_end = above ? lane + __ffs(above) - 1 : 31;
        // a chunk is 128 pixels: every sum of a segment stays below 2^16 (128 * 255), so two 16-bit lanes share a register
        unsigned bg = b | (g << 16), rc = r | (cnt << 16);
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const unsigned bg2 = __shfl_down_sync(FULL, bg, o);
            const unsigned rc2 = __shfl_down_sync(FULL, rc, o);
            if (lane + o <= seg_end) { bg += bg2; rc += rc2; }
        }
        if (head && cur > 0) stats_flush(t, cur, rc >> 16, bg & 0xFFFFu, bg >> 16, rc & 0xFFFFu);
        // ---- adjacency pairs
        int pl = __shfl_up_sync(FULL, lab[3], 1), pd = __shfl_up_sync(FULL, dn[3], 1);   // pixel left of my first one
        if (lane == 0) { pl = 0; pd = 0; }
        int rn[4] = {lab[1], lab[2], lab[3], nx};
        bool er[4], ed[4];
        int mine = 0;
#pragma unroll
        for (int k = 0; k < 4; k++) {
            er[k] = lab[k] > 0 && rn[k] > 0 && rn[k] != lab[k] && !(up[k] == lab[k] && up[k + 1] == rn[k]);
            int ll = k ? lab[k - 1] : pl, ld = k ? dn[k - 1] : pd;
            ed[k] = lab[k] > 0 && dn[k] > 0 && dn[k] != lab[k] && !(ll == lab[k] && ld == dn[k]);
            mine += (int)er[k] + (int)ed[k];
        }
        if (__any_sync(FULL, mine)) {
            int incl = mine;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                int v = __shfl_up_sync(FULL, incl, o);
                if (lane >= o) incl += v;
            }
            int tot = __shfl_sync(FULL, incl, 31);
            long long pos = 0;
            if (lane == 0) pos = (long long)atomicAdd(A.npairs, tot);
            pos = __shfl_sync(FULL, pos, 0) + incl - mine;
#pragma unroll
            for (int k = 0; k < 4; k++) {
                if (er[k]) { if (pos < A.pair_cap) A.pairs[pos] = make_int2(lab[k], rn[k]); pos++; }
                if (ed[k]) { if (pos < A.pair_cap) A.pairs[pos] = make_int2(lab[k], dn[k]); pos++; }
            }
        }
    }
}


// Statistics + adjacency pass, column strips (requires w % 4 == 0 and 16-byte aligned label rows).  A warp walks DOWN a strip
// of 128 columns (4 consecutive pixels per lane) for 16..64 rows: the labels of the row above and below are the registers of
// the previous / next iteration (one label load and one colour load per row instead of three and one, and the next row is
// in flight while this one is processed), and the sums stay in the lane: two register-resident entries {label, count,
// B, G, R} per lane take the pixels of the lane's four columns row after row and go to the tables (4 atomics) only when a
// third label shows up or the strip ends -- no segmented shuffle reduction, and a region costs a few atomics per lane and
// strip instead of a few per 128-pixel chunk of every row.  All sums are integer: the order of the additions is free.
// 16-bit lanes: an entry sees at most 64 rows x 4 pixels x 255 = 65 280 < 2^16 per channel.
// Adjacent-pair list: the rules of stats_pass_vec4 (a right pair is skipped when the row above holds the same pair, a down
// pair when the pixel to the left holds the same pair; the first pixel of a strip row never skips).
constexpr int SS_MIN_ROWS = 16, SS_MAX_ROWS = 64;
struct stat_entry { int lab; unsigned br, gc; };           // br = sum B | sum R << 16, gc = sum G | count << 16

__device__ __forceinline__ void stat_entry_flush(const merge_tables& t, const stat_entry& e)
{
    if (e.lab > 0 && e.gc) stats_flush(t, e.lab, e.gc >> 16, e.br & 0xFFFFu, e.gc & 0xFFFFu, e.br >> 16);
}

// The raw pair list is ONE array with ONE counter: an atomicAdd per warp and row on that counter is a same-address atomic every
// 128 pixels -- ~1 ns each, serialised in one L2 slice: 59 k of them per 4K frame, 476 k at 8192^2 were the duration of the
// kernel (57 us / 0.5 ms) whatever else it did.  A warp therefore collects its pairs in shared memory and claims list space
// once per SS_PAIR_BUF pairs (a row adds at most 32 lanes x 8 = 256).
constexpr int SS_PAIR_BUF = 256;

__device__ __forceinline__ void stat_pairs_flush(const merge_persist_args& A, const int2* buf, int nbuf, int lane)
{
    if (nbuf == 0) return;                                                        // warp-uniform
    long long base = 0;
    if (lane == 0) base = (long long)atomicAdd(A.npairs, nbuf);
    base = __shfl_sync(0xffffffffu, base, 0);
    for (int i = lane; i < nbuf; i += 32)
        if (base + i < A.pair_cap) A.pairs[base + i] = buf[i];
    __syncwarp();
}

__device__ __forceinline__ void stats_pass_strips(const merge_persist_args& A, const merge_tables& t, int nin, int lane,
                                                  long long gwarp, long long nwarps, int2* buf)
{
    int nbuf = 0;
    const int w = A.w, h = A.h;
    const int sx = (w + 127) / 128;
    const long long per = ((long long)h * sx + nwarps - 1) / nwarps;              // rows per unit: about one unit per warp
    const int ru = (int)(per < SS_MIN_ROWS ? SS_MIN_ROWS : (per > SS_MAX_ROWS ? SS_MAX_ROWS : per));
    const long long nunits = (long long)sx * ((h + ru - 1) / ru);
    const unsigned FULL = 0xffffffffu;
    const unsigned unin = (unsigned)nin;
    auto clean = [&](int v) { return (unsigned)v > unin ? 0 : v; };             // negative or beyond the table: background
    for (long long u = gwarp; u < nunits; u += nwarps) {
        const int x = (int)(u % sx) * 128 + lane * 4;
        const int y0 = (int)(u / sx) * ru, y1 = min(h, y0 + ru);
        const bool in = x < w;                                                    // w % 4 == 0: all four pixels or none
        const bool has_right = in && x + 4 < w;
        const bool edge = lane == 31 && has_right;                                // my right neighbour is in the next strip
        const int32_t* lrow = A.labels + (size_t)y0 * w + x;
        int up[5] = {0, 0, 0, 0, 0};                                              // row above: my four columns + my right neighbour's
        if (y0 > 0 && in) {
            const int4 U4 = *reinterpret_cast<const int4*>(lrow - w);
            up[0] = clean(U4.x); up[1] = clean(U4.y); up[2] = clean(U4.z); up[3] = clean(U4.w);
        }
        up[4] = __shfl_down_sync(FULL, up[0], 1);
        if (edge) up[4] = y0 > 0 ? clean(lrow[4 - w]) : 0;
        if (!has_right) up[4] = 0;
        int lab[4] = {0, 0, 0, 0};
        if (in) {
            const int4 L4 = *reinterpret_cast<const int4*>(lrow);
            lab[0] = clean(L4.x); lab[1] = clean(L4.y); lab[2] = clean(L4.z); lab[3] = clean(L4.w);
        }
        stat_entry e0 = {0, 0u, 0u}, e1 = {0, 0u, 0u};
        for (int y = y0; y < y1; y++, lrow += w) {
            int dn[4] = {0, 0, 0, 0};
            uint4 C4 = make_uint4(0, 0, 0, 0);
            int nx = 0;
            if (in) {
                if (y + 1 < h) {
                    const int4 D4 = *reinterpret_cast<const int4*>(lrow + w);
                    dn[0] = clean(D4.x); dn[1] = clean(D4.y); dn[2] = clean(D4.z); dn[3] = clean(D4.w);
                }
                C4 = __ldg(reinterpret_cast<const uint4*>(A.plane + (size_t)y * A.pitch + x));
                if (edge) nx = clean(lrow[4]);
            }
            {
                const int t0 = __shfl_down_sync(FULL, lab[0], 1);
                if (!edge) nx = has_right ? t0 : 0;
            }
            // ---- sums
            const uint32_t col[4] = {C4.x, C4.y, C4.z, C4.w};
            const bool uni = lab[0] == lab[1] && lab[1] == lab[2] && lab[2] == lab[3];
            if (uni && lab[0] == e1.lab && lab[0] != e0.lab) { const stat_entry s = e0; e0 = e1; e1 = s; }
            if (uni && lab[0] == e0.lab) {
                e0.br += (col[0] & 0x00FF00FFu) + (col[1] & 0x00FF00FFu) + (col[2] & 0x00FF00FFu) + (col[3] & 0x00FF00FFu);
                e0.gc += ((col[0] >> 8) & 0xFFu) + ((col[1] >> 8) & 0xFFu) + ((col[2] >> 8) & 0xFFu) + ((col[3] >> 8) & 0xFFu) + 0x40000u;
            } else {
#pragma unroll
                for (int k = 0; k < 4; k++) {
                    const unsigned abr = col[k] & 0x00FF00FFu, agc = ((col[k] >> 8) & 0xFFu) + 0x10000u;
                    if (lab[k] == e0.lab) { e0.br += abr; e0.gc += agc; }
                    else if (lab[k] == e1.lab) { e1.br += abr; e1.gc += agc; }
                    else { stat_entry_flush(t, e1); e1.lab = lab[k]; e1.br = abr; e1.gc = agc; }
                }
            }
            // ---- adjacency pairs
            int pl = __shfl_up_sync(FULL, lab[3], 1), pd = __shfl_up_sync(FULL, dn[3], 1);   // pixel left of my first one
            if (lane == 0) { pl = 0; pd = 0; }
            const int rn[4] = {lab[1], lab[2], lab[3], nx};
            bool er[4], ed[4];
            int mine = 0;
#pragma unroll
            for (int k = 0; k < 4; k++) {
                er[k] = lab[k] > 0 && rn[k] > 0 && rn[k] != lab[k] && !(up[k] == lab[k] && up[k + 1] == rn[k]);
                const int ll = k ? lab[k - 1] : pl, ld = k ? dn[k - 1] : pd;
                ed[k] = lab[k] > 0 && dn[k] > 0 && dn[k] != lab[k] && !(ll == lab[k] && ld == dn[k]);
                mine += (int)er[k] + (int)ed[k];
            }
            if (__any_sync(FULL, mine)) {
                int incl = mine;
#pragma unroll
                for (int o = 1; o < 32; o <<= 1) {
                    const int v = __shfl_up_sync(FULL, incl, o);
                    if (lane >= o) incl += v;
                }
                const int tot = __shfl_sync(FULL, incl, 31);                      // <= 256 = SS_PAIR_BUF
                if (nbuf + tot > SS_PAIR_BUF) { stat_pairs_flush(A, buf, nbuf, lane); nbuf = 0; }
                int pos = nbuf + incl - mine;
#pragma unroll
                for (int k = 0; k < 4; k++) {
                    if (er[k]) buf[pos++] = make_int2(lab[k], rn[k]);
                    if (ed[k]) buf[pos++] = make_int2(lab[k], dn[k]);
                }
                nbuf += tot;
                __syncwarp();
            }
            // ---- next row
#pragma unroll
            for (int k = 0; k < 4; k++) { up[k] = lab[k]; lab[k] = dn[k]; }
            up[4] = nx;
        }
        stat_entry_flush(t, e0);
        stat_entry_flush(t, e1);
    }
    stat_pairs_flush(A, buf, nbuf, lane);
}

__global__ void __launch_bounds__(MT) merge_stats_kernel(merge_persist_args A)
{
    const int lane = threadIdx.x & 31;
    const long long gtid = (long long)blockIdx.x * MT + threadIdx.x;
    const long long nthreads = (long long)gridDim.x * MT;
    const long long gwarp = gtid >> 5, nwarps = nthreads >> 5;
    const int nin = merge_nin(A);
    const int w = A.w, h = A.h;
    const merge_tables t = A.t;
    if (A.vec == 2) {
        __shared__ int2 s_pairs[MT / 32][SS_PAIR_BUF];
        stats_pass_strips(A, t, nin, lane, gwarp, nwarps, s_pairs[threadIdx.x >> 5]);
        return;
    }
    if (A.vec) { stats_pass_vec4(A, t, nin, lane, gwarp, nwarps); return; }
    const int cpr = (w + 31) / 32;                      // 32-pixel chunks per row
    const long long nchunks = (long long)cpr * h;
    for (long long c = gwarp; c < nchunks; c += nwarps) {
        int y = (int)(c / cpr), x = (int)(c % cpr) * 32 + lane;
        int lab = 0;
        unsigned cnt = 0, b = 0, g = 0, r = 0;
        if (x < w) {
            lab = A.labels[(size_t)y * w + x];
            if (lab > nin) lab = 0;
            if (lab > 0) {
                uint32_t col = __ldg(A.plane + (size_t)y * A.pitch + x);
                cnt = 1; b = col & 0xFF; g = (col >> 8) & 0xFF; r = (col >> 16) & 0xFF;
            }
        }
        int prev = __shfl_up_sync(0xffffffffu, lab, 1);
        bool head = lane == 0 || prev != lab;
        unsigned heads = __ballot_sync(0xffffffffu, head);
        unsigned above = lane == 31 ? 0u : (heads >> (lane + 1));
        int seg_end = above ? lane + __ffs(above) - 1 : 31;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            unsigned c2 = __shfl_down_sync(0xffffffffu, cnt, o);
            unsigned b2 = __shfl_down_sync(0xffffffffu, b, o);
            unsigned g2 = __shfl_down_sync(0xffffffffu, g, o);
            unsigned r2 = __shfl_down_sync(0xffffffffu, r, o);
            if (lane + o <= seg_end) { cnt += c2; b += b2; g += g2; r += r2; }
        }
        if (head && lab > 0) {
            atomicAdd(t.area + lab, cnt);
            atomicAdd(t.sum + 3 * (size_t)lab, (unsigned long long)b);
            atomicAdd(t.sum + 3 * (size_t)lab + 1, (unsigned long long)g);
            atomicAdd(t.sum + 3 * (size_t)lab + 2, (unsigned long long)r);
        }
        // region adjacency list: every 4-adjacent pixel pair with two different positive labels, found once here; the
        // rounds then iterate over this list (a few % of the pixels) instead of over the image
        int lr = 0, ld = 0;
        if (lab > 0) {
            if (x + 1 < w) { lr = A.labels[(size_t)y * w + x + 1]; if (lr > nin || lr == lab) lr = 0; }
            if (y + 1 < h) { ld = A.labels[(size_t)(y + 1) * w + x]; if (ld > nin || ld == lab) ld = 0; }
        }
        unsigned mr = __ballot_sync(0xffffffffu, lr > 0), md = __ballot_sync(0xffffffffu, ld > 0);
        int tot = __popc(mr) + __popc(md);
        if (tot) {
            long long pos = 0;
            if (lane == 0) pos = (long long)atomicAdd(A.npairs, tot);
            pos = __shfl_sync(0xffffffffu, pos, 0);
            unsigned lt = (1u << lane) - 1;
            if (lr > 0) { long long k = pos + __popc(mr & lt); if (k < A.pair_cap) A.pairs[k] = make_int2(lab, lr); }
            if (ld > 0) { long long k = pos + __popc(mr) + __popc(md & lt); if (k < A.pair_cap) A.pairs[k] = make_int2(lab, ld); }
        }
    }
}

// ---------------------------------------------------------------- the set of adjacent pairs (frame-sized images)
// One pass over the raw pair list with the whole GPU: insert into the global hash set, append the new ones to A.uniq.  (Doing
// this inside the statistics pass put an L2 atomic round trip into the critical path of its warps: 66 -> 89 us per 4K frame;
// as a kernel of its own the list is spread evenly over the threads and the latency is hidden: profiles/r02_launches_4k.md.)
__global__ void __launch_bounds__(MT) merge_pair_set_kernel(merge_persist_args A)
{
    const int lane = threadIdx.x & 31;
    const long long gtid = (long long)blockIdx.x * MT + threadIdx.x, nthreads = (long long)gridDim.x * MT;
    const long long npairs = merge_npairs(A);
    const int ucap = 1 << A.hbits;
    for (long long base = gtid - lane; base < npairs; base += nthreads) {          // warp-uniform trip count
        const long long i = base + lane;
        int2 pr = make_int2(0, 0);
        bool fresh = false;
        if (i < npairs) {
            pr = A.pairs[i];
            fresh = pr.x > 0 && pr.y > 0 && pr.x != pr.y && pair_set_insert(A, pr.x, pr.y);
        }
        const unsigned m = __ballot_sync(0xffffffffu, fresh);
        if (m) {
            int pos = 0;
            if (lane == 0) pos = atomicAdd(A.nuniq, __popc(m));
            pos = __shfl_sync(0xffffffffu, pos, 0) + __popc(m & ((1u << lane) - 1));
            if (fresh) {
                if (pos < ucap) A.uniq[pos] = pr; else A.nuniq[1] = 1;
            }
        }
    }
}

// ---------------------------------------------------------------- rounds, small path (one CTA, tables in shared memory)
// Selection key: (dist2 << 12) | neighbour label -- dist2 <= 3 * 255^2 < 2^18 and labels < 2^12, so 32 bits are enough
// and shared-memory atomicMin is native.  Colour sums fit 32 bits because the image has at most 2^24 pixels.
__device__ __forceinline__ int sm_find(volatile int* par, int a)
{
    int p = par[a];
    while (p != a) { a = p; p = par[a]; }
    return a;
}

__device__ __forceinline__ void sm_union(int* par, int a, int b)
{
    for (;;) {
        a = sm_find(par, a);
        b = sm_find(par, b);
        if (a == b) return;
        if (a < b) { int s = a; a = b; b = s; }
        int old = atomicMin(par + a, b);
        if (old == a) return;
        a = old;
    }
}

template <int LBITS>
__global__ void __launch_bounds__(ST, 1) merge_rounds_small_kernel(merge_persist_args A)
{
    extern __shared__ uint32_t sm_tab[];
    constexpr int NS = 1 << LBITS;                   // labels 0 .. NS - 1: LBITS-bit label field in the 32-bit selection key
    constexpr uint32_t LMASK = NS - 1;
    int* par = reinterpret_cast<int*>(sm_tab);
    uint32_t* area = sm_tab + NS;
    uint32_t* mean = sm_tab + 2 * NS;
    uint32_t* best = sm_tab + 3 * NS;          // reused for the dense ids after the rounds
    uint32_t* sum0 = sm_tab + 4 * NS;
    uint32_t* sum1 = sm_tab + 5 * NS;
    uint32_t* sum2 = sm_tab + 6 * NS;
    __shared__ int s_accepted;
    __shared__ int s_wsum[ST / 32];
    const int nin = merge_nin(A);
    if (merge_regime(A, nin) != (LBITS == 12 ? 0 : 1)) return;
    const int nl = nin + 1;
    const int tid = threadIdx.x, lane = tid & 31;
    // the pairs the rounds walk: the set built by the statistics pass (read coalesced from L2), or the raw list when there is
    // none (strip-sharded merge) or its table overflowed
    const int2* __restrict__ pairs = A.pairs;
    long long npairs = merge_npairs(A);
    if (npairs > A.pair_cap) npairs = A.pair_cap;
    if (A.set_ready && A.nuniq[1] == 0) { pairs = A.uniq; npairs = A.nuniq[0]; }
    for (int i = tid; i < nl; i += ST) {
        par[i] = i;
        area[i] = A.t.area[i];
        sum0[i] = (uint32_t)A.t.sum[3 * (size_t)i];
        sum1[i] = (uint32_t)A.t.sum[3 * (size_t)i + 1];
        sum2[i] = (uint32_t)A.t.sum[3 * (size_t)i + 2];
    }
    __syncthreads();
    auto edge = [&](int la, int lb, uint32_t size_thr) {    // both ends of an adjacent pair bid for each other
        int ra = par[la], rb = par[lb];
        if (ra == rb) return;
        bool pa = area[ra] < size_thr, pb = area[rb] < size_thr;
        if (!pa && !pb) return;
        uint32_t e = __vabsdiffu4(mean[ra], mean[rb]);
        uint32_t d2 = __dp4a(e, e, 0u) << LBITS;
        if (pa) atomicMin(best + ra, d2 | (uint32_t)rb);
        if (pb) atomicMin(best + rb, d2 | (uint32_t)ra);
    };
    int rounds = 0;
    for (int phase = 0; phase < 2; phase++) {
        uint32_t size_thr, dist_limit;
        if (phase == 0) { if (A.color_dist <= 0) continue; size_thr = 0xffffffffu; dist_limit = (uint32_t)min((long long)A.color_dist * A.color_dist, 1ll << 19); }
        else { if (A.min_size <= 0) continue; size_thr = (uint32_t)A.min_size; dist_limit = 0xffffffffu; }
        for (int round = 0; round < MERGE_MAX_ROUNDS; round++) {
            for (int i = tid; i < nl; i += ST) {          // rounded means of the live roots: floor(s / a + 1/2)
                uint32_t a = area[i], m = 0;
                if (i > 0 && a) {
                    uint32_t q0 = sum0[i] / a, q1 = sum1[i] / a, q2 = sum2[i] / a;
                    q0 += 2 * (sum0[i] - q0 * a) >= a;
                    q1 += 2 * (sum1[i] - q1 * a) >= a;
                    q2 += 2 * (sum2[i] - q2 * a) >= a;
                    m = q0 | (q1 << 8) | (q2 << 16);
                }
                mean[i] = m;
                best[i] = 0xffffffffu;
            }
            if (tid == 0) s_accepted = 0;
            __syncthreads();
            for (long long k = tid; k < npairs; k += ST) {
                const int2 pr = pairs[k];
                edge(pr.x, pr.y, size_thr);
            }
            __syncthreads();
            for (int i = tid; i < nl; i += ST) {
                if (i == 0) continue;
                uint32_t k = best[i];
                if (k == 0xffffffffu || (k >> LBITS) > dist_limit) continue;
                sm_union(par, i, (int)(k & LMASK));
                atomicAdd(&s_accepted, 1);
            }
            __syncthreads();
            rounds++;
            if (s_accepted == 0) break;                    // uniform; it is reset only after the barrier below
            for (int i = tid; i < nl; i += ST) {           // fold: flatten, move the statistics of absorbed labels to their roots
                if (i == 0) continue;
                int r = sm_find(par, i);
                if (r != i) {
                    par[i] = r;
                    uint32_t a = area[i];
                    if (a) {
                        atomicAdd(area + r, a);
                        atomicAdd(sum0 + r, sum0[i]); atomicAdd(sum1 + r, sum1[i]); atomicAdd(sum2 + r, sum2[i]);
                        area[i] = 0;
                    }
                }
            }
            __syncthreads();
        }
    }
    // dense renumbering of the surviving roots (ascending label == ascending first pixel); 4 consecutive labels per thread
    constexpr int PER = NS / ST;
    int alive[PER], cnt = 0;
#pragma unroll
    for (int k = 0; k < PER; k++) {
        int i = tid * PER + k;
        alive[k] = (i > 0 && i < nl && par[i] == i && area[i] > 0) ? 1 : 0;
        cnt += alive[k];
    }
    int incl = cnt;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        int v = __shfl_up_sync(0xffffffffu, incl, o);
        if (lane >= o) incl += v;
    }
    if (lane == 31) s_wsum[tid >> 5] = incl;
    __syncthreads();                                       // also: every read of best[] as a key is done
    int ws = s_wsum[lane], wincl = ws;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        int v = __shfl_up_sync(0xffffffffu, wincl, o);
        if (lane >= o) wincl += v;
    }
    int run = __shfl_sync(0xffffffffu, wincl - ws, tid >> 5) + incl - cnt;
    const int total = __shfl_sync(0xffffffffu, wincl, 31);
#pragma unroll
    for (int k = 0; k < PER; k++) {
        int i = tid * PER + k;
        if (i < nl) best[i] = (uint32_t)run;
        run += alive[k];
    }
    __syncthreads();
    int32_t* fin = reinterpret_cast<int32_t*>(A.t.mean);
    for (int i = tid; i < nl; i += ST) fin[i] = i > 0 ? (int)best[par[i]] + 1 : 0;
    if (tid == 0) { *A.n_out = total; *A.rounds_out = rounds; }
}

// ---------------------------------------------------------------- rounds, large path (cooperative grid)
// All rounds of both phases inside one kernel with grid-wide barriers: no host round trip, no per-round pixel
// passes for statistics (per-label sums are folded into the surviving roots) and the final renumbering is a scan
// over the label table (valid because input labels are numbered by first pixel and unions keep the smallest label).
__global__ void __launch_bounds__(MT) merge_rounds_large_kernel(merge_persist_args A)
{
    cg::grid_group grid = cg::this_grid();
    const int lane = threadIdx.x & 31;
    const long long gtid = (long long)blockIdx.x * MT + threadIdx.x;
    const long long nthreads = (long long)gridDim.x * MT;
    const int nin = merge_nin(A);
    if (merge_regime(A, nin) != 2) return;               // uniform over the grid: a single-CTA kernel did the work
    const int nl = nin + 1;
    merge_tables t = A.t;
    long long npairs = merge_npairs(A);
    const int2* __restrict__ pairs = A.pairs;

    // ---- the set of adjacent pairs (every round walks it; the raw list repeats a pair once per boundary corner)
    if (A.set_ready) {
        if (*((volatile int32_t*)A.nuniq + 1) == 0) { pairs = A.uniq; npairs = *((volatile int32_t*)A.nuniq); }
    } else if (A.htab) {
        const long long slots = 1ll << A.hbits;
        for (long long i = gtid; i < slots; i += nthreads) A.htab[i] = 0ull;
        if (gtid == 0) { A.nuniq[0] = 0; A.nuniq[1] = 0; }
        grid.sync();
        const unsigned long long hmask = (unsigned long long)slots - 1;
        for (long long base = (gtid - lane); base < npairs; base += nthreads) {     // warp-uniform trip count
            const long long i = base + lane;
            bool fresh = false;
            int2 pr = make_int2(0, 0);
            if (i < npairs) {
                pr = A.pairs[i];
                if (pr.x > 0 && pr.y > 0 && pr.x != pr.y) {
                    const unsigned lo = (unsigned)min(pr.x, pr.y), hi = (unsigned)max(pr.x, pr.y);
                    const unsigned long long key = ((unsigned long long)lo << 32) | hi;
                    unsigned long long slot = (key * 0x9E3779B97F4A7C15ull) >> (64 - A.hbits);
                    int probes = 0;
                    for (; probes < 64; probes++) {
                        const unsigned long long old = atomicCAS(A.htab + slot, 0ull, key);
                        if (old == 0ull) { fresh = true; break; }
                        if (old == key) break;
                        slot = (slot + 1) & hmask;
                    }
                    if (probes == 64) A.nuniq[1] = 1;                                  // table too full for this image
                }
            }
            const unsigned m = __ballot_sync(0xffffffffu, fresh);
            if (m) {
                int pos = 0;
                if (lane == 0) pos = atomicAdd(A.nuniq, __popc(m));
                pos = __shfl_sync(0xffffffffu, pos, 0) + __popc(m & ((1u << lane) - 1));
                if (fresh) {
                    if (pos < slots) A.uniq[pos] = pr; else A.nuniq[1] = 1;
                }
            }
        }
        grid.sync();
        if (*((volatile int32_t*)A.nuniq + 1) == 0) {
            pairs = A.uniq;
            npairs = *((volatile int32_t*)A.nuniq);
        }
    }

    const long long INF = 1ll << 40;
    int rounds = 0;
    for (int phase = 0; phase < 2; phase++) {
        long long size_thr, dist_limit;
        if (phase == 0) { if (A.color_dist <= 0) continue; size_thr = INF; dist_limit = (long long)A.color_dist * A.color_dist; }
        else { if (A.min_size <= 0) continue; size_thr = A.min_size; dist_limit = INF; }
        for (int round = 0; round < MERGE_MAX_ROUNDS; round++) {
            // means of the live roots, reset the selection keys
            for (long long i = gtid; i < nl; i += nthreads) {
                unsigned long long a = t.area[i];
                uint32_t m = 0;
                if (i > 0 && a) {
#pragma unroll
                    for (int c = 0; c < 3; c++) m |= (uint32_t)((2ull * t.sum[3 * i + c] + a) / (2ull * a)) << (8 * c);
                }
                t.mean[i] = m;
                t.best[i] = ~0ull;
            }
            if (gtid == 0) *A.accepted = 0;
            grid.sync();
            // region adjacency through the (flattened) parent table: both ends of every recorded pair
            for (long long i = gtid; i < npairs; i += nthreads) {
                int2 pr = pairs[i];
                int ra = __ldcg(t.par + pr.x), rb = __ldcg(t.par + pr.y);
                if (ra == rb) continue;
                bool pa = (long long)t.area[ra] < size_thr, pb = (long long)t.area[rb] < size_thr;
                if (!pa && !pb) continue;
                uint32_t e = __vabsdiffu4(t.mean[ra], t.mean[rb]);
                unsigned long long d2 = (unsigned long long)__dp4a(e, e, 0u) << 32;
                if (pa) { unsigned long long key = d2 | (unsigned)rb; if (key < t.best[ra]) atomicMin(t.best + ra, key); }
                if (pb) { unsigned long long key = d2 | (unsigned)ra; if (key < t.best[rb]) atomicMin(t.best + rb, key); }
            }
            grid.sync();
            // accepted selections are united (smallest label wins)
            for (long long i = gtid; i < nl; i += nthreads) {
                if (i == 0) continue;
                unsigned long long k = t.best[i];
                if (k == ~0ull || (long long)(k >> 32) > dist_limit) continue;
                persist_union(t.par, (int)i, (int)(k & 0xffffffffu));
                atomicAdd(A.accepted, 1);
            }
            grid.sync();
            rounds++;
            const int acc = *((volatile int32_t*)A.accepted);
            if (acc == 0) break;                              // uniform: every thread reads the same value
            // fold: flatten parents, move the statistics of absorbed labels into their roots
            for (long long i = gtid; i < nl; i += nthreads) {
                if (i == 0) continue;
                int r = pf_find(t.par, (int)i);
                if (r != (int)i) {
                    t.par[i] = r;
                    unsigned a = t.area[i];
                    if (a) {
                        atomicAdd(t.area + r, a);
                        atomicAdd(t.sum + 3 * (size_t)r, t.sum[3 * i]);
                        atomicAdd(t.sum + 3 * (size_t)r + 1, t.sum[3 * i + 1]);
                        atomicAdd(t.sum + 3 * (size_t)r + 2, t.sum[3 * i + 2]);
                        t.area[i] = 0;
                    }
                }
            }
            grid.sync();
        }
    }

    // ---- dense renumbering of the surviving roots (ascending label == ascending first pixel)
    const int per = (nl + gridDim.x - 1) / gridDim.x;          // labels per block
    {
        __shared__ int carry;
        if (threadIdx.x == 0) carry = 0;
        __syncthreads();
        const int lo = blockIdx.x * per, hi = min(nl, lo + per);
        for (int base = lo; base < hi; base += MT) {
            int i = base + threadIdx.x;
            int alive = (i < hi && i > 0 && t.par[i] == i && t.area[i] > 0) ? 1 : 0;
            // block exclusive scan of `alive`
            __shared__ int wsum[MT / 32];
            int incl = alive;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                int v = __shfl_up_sync(0xffffffffu, incl, o);
                if (lane >= o) incl += v;
            }
            if (lane == 31) wsum[threadIdx.x >> 5] = incl;
            __syncthreads();
            int ws = lane < MT / 32 ? wsum[lane] : 0, wincl = ws;
#pragma unroll
            for (int o = 1; o < MT / 32; o <<= 1) {
                int v = __shfl_up_sync(0xffffffffu, wincl, o);
                if (lane >= o) wincl += v;
            }
            int woff = __shfl_sync(0xffffffffu, wincl - ws, threadIdx.x >> 5);
            int total = __shfl_sync(0xffffffffu, wincl, MT / 32 - 1);
            int c0 = carry;
            if (i < hi) A.newid[i] = c0 + woff + incl - alive;
            __syncthreads();
            if (threadIdx.x == 0) carry = c0 + total;
            __syncthreads();
        }
        if (threadIdx.x == 0) A.bsum[blockIdx.x] = carry;
    }
    grid.sync();
    if (blockIdx.x == 0 && threadIdx.x == 0) {
        int run = 0;
        for (unsigned b = 0; b < gridDim.x; b++) { int v = A.bsum[b]; A.bsum[b] = run; run += v; }
        *A.n_out = run;
        *A.rounds_out = rounds;
    }
    grid.sync();
    // ---- final id of every input label (one gather per pixel in the rewrite pass); t.mean is free by now
    int32_t* fin = reinterpret_cast<int32_t*>(t.mean);
    for (long long i = gtid; i < nl; i += nthreads) {
        int r = __ldcg(t.par + i);
        fin[i] = i > 0 ? A.newid[r] + A.bsum[r / per] + 1 : 0;
    }
}

// ---------------------------------------------------------------- rewrite the pixels: label -> final id
// Labels outside 1..n_in are left as they are.  16-byte loads, two in flight per thread, when the buffer is aligned.
__global__ void __launch_bounds__(MT) merge_rewrite_kernel(merge_persist_args A)
{
    const int nin = merge_nin(A);
    const int32_t* __restrict__ fin = reinterpret_cast<const int32_t*>(A.t.mean);
    const long long gtid = (long long)blockIdx.x * MT + threadIdx.x, nthreads = (long long)gridDim.x * MT;
    const long long n = (long long)A.w * A.h;
    long long done = 0;
    if ((reinterpret_cast<uintptr_t>(A.labels) & 15) == 0) {
        int4* L4 = reinterpret_cast<int4*>(A.labels);
        const long long n4 = n >> 2;
        auto map4 = [&](int4 v) {
            int4 o = v;
            if (v.x > 0 && v.x <= nin) o.x = __ldg(fin + v.x);
            if (v.y > 0 && v.y <= nin) o.y = v.y == v.x ? o.x : __ldg(fin + v.y);
            if (v.z > 0 && v.z <= nin) o.z = v.z == v.y ? o.y : __ldg(fin + v.z);
            if (v.w > 0 && v.w <= nin) o.w = v.w == v.z ? o.z : __ldg(fin + v.w);
            return o;
        };
        long long i = gtid;
        for (; i + nthreads < n4; i += 2 * nthreads) {
            int4 a = L4[i], b2 = L4[i + nthreads];
            L4[i] = map4(a);
            L4[i + nthreads] = map4(b2);
        }
        if (i < n4) L4[i] = map4(L4[i]);
        done = n4 << 2;
    }
    for (long long p = done + gtid; p < n; p += nthreads) {
        int l0 = A.labels[p];
        if (l0 <= 0 || l0 > nin) continue;
        A.labels[p] = fin[l0];
    }
}

__global__ void __launch_bounds__(MT) render_kernel(const int32_t* __restrict__ L, size_t lstep, uint8_t* __restrict__ dst,
                                                    size_t dstep, int w, int depth, const uint8_t* __restrict__ colors)
{
    int x = blockIdx.x * MT + threadIdx.x;
    int y = blockIdx.y;
    if (x >= w) return;
    int v = ((const int32_t*)((const char*)L + (size_t)y * lstep))[x];
    uint8_t c0 = 0, c1 = 0, c2 = 0;
    if (v > 0 && v <= depth) {
        if (colors) { c0 = colors[3 * (size_t)(v - 1)]; c1 = colors[3 * (size_t)(v - 1) + 1]; c2 = colors[3 * (size_t)(v - 1) + 2]; }
        else c0 = c1 = c2 = 255;
    }
    uint8_t* p = dst + (size_t)y * dstep + 3 * (size_t)x;
    p[0] = c0; p[1] = c1; p[2] = c2;
}

// four labels per thread: one 16-byte load, 12 bytes of BGR as three 32-bit words
__global__ void __launch_bounds__(MT) render_v4_kernel(const int32_t* __restrict__ L, size_t lstep, uint8_t* __restrict__ dst,
                                                       size_t dstep, int w4, int depth, const uint8_t* __restrict__ colors)
{
    int q = blockIdx.x * MT + threadIdx.x;
    int y = blockIdx.y;
    if (q >= w4) return;
    int4 v = *(const int4*)((const char*)L + (size_t)y * lstep + 16 * (size_t)q);
    const int lab[4] = {v.x, v.y, v.z, v.w};
    uint32_t c[4];
#pragma unroll
    for (int i = 0; i < 4; i++) {
        uint32_t col = 0;
        if (lab[i] > 0 && lab[i] <= depth) {
            if (colors) {
                const uint8_t* t = colors + 3 * (size_t)(lab[i] - 1);
                col = (uint32_t)__ldg(t) | ((uint32_t)__ldg(t + 1) << 8) | ((uint32_t)__ldg(t + 2) << 16);
            } else col = 0x00FFFFFFu;
        }
        c[i] = col;
    }
    uint32_t* p = (uint32_t*)(dst + (size_t)y * dstep) + 3 * (size_t)q;
    p[0] = c[0] | (c[1] << 24);
    p[1] = (c[1] >> 8) | (c[2] << 16);
    p[2] = (c[2] >> 16) | (c[3] << 8);
}

}  // namespace

int k_render(msg_ctx* ctx, const int32_t* d_labels, size_t lstep, uint8_t* d_dst, size_t dstep, int w, int h, int depth,
             const uint8_t* d_colors)
{
    if (w % 4 == 0 && lstep % 16 == 0 && dstep % 4 == 0 && ((uintptr_t)d_labels & 15) == 0 && ((uintptr_t)d_dst & 3) == 0) {
        dim3 grid4((w / 4 + MT - 1) / MT, h);
        render_v4_kernel<<<grid4, MT, 0, ctx->stream>>>(d_labels, lstep, d_dst, dstep, w / 4, depth, d_colors);
        MSG_LAUNCHED(ctx);
        MSG_CHECK_LAUNCH(ctx);
        return MSG_OK;
    }
    dim3 grid((w + MT - 1) / MT, h);
    render_kernel<<<grid, MT, 0, ctx->stream>>>(d_labels, lstep, d_dst, dstep, w, depth, d_colors);
    MSG_LAUNCHED(ctx);
    MSG_CHECK_LAUNCH(ctx);
    return MSG_OK;
}

// d_n_in: device count of input labels, which must be canonical (1..n by first pixel).
// d_counters: [8] max label, [9] accepted, [10] rounds, [11] spare n_out, [13] pair count, [14] saved n_in
static int merge_launch(msg_ctx* ctx, const uint32_t* d_plane, int pitch, int32_t* d_labels, int w, int h, int min_size,
                        int color_dist, const int32_t* d_n_in, int32_t* d_n_out)
{
    size_t n = (size_t)w * h;
    int cap = (int)n;                                 // worst case: every pixel its own region
    if (ctx->merge_blocks_per_sm <= 0)       // occupancy of the cooperative kernel: queried once per context, not per launch
        MSG_CUDA(ctx, cudaOccupancyMaxActiveBlocksPerMultiprocessor(&ctx->merge_blocks_per_sm, merge_rounds_large_kernel, MT, 0));
    const int blocks_per_sm = ctx->merge_blocks_per_sm;
    if (blocks_per_sm < 1) return msg_fail(ctx, MSG_ECUDA, "merge: cooperative kernel does not fit");
    const int grid_max = ctx->sm_count * (blocks_per_sm < 2 ? blocks_per_sm : 2);
    // CTAs of the cooperative rounds kernel (option "merge_grid" overrides).  Alone, more CTAs are faster (merge stage at 4K /
    // 8192^2: 16 CTAs 0.52 / 4.79 ms, 32: 0.37 / 2.84, 64: 0.28 / 1.83, 148: 0.24 / 1.25, 296: 0.22 / 1.04: the passes are bound by
    // the work over the pair list, not by the grid barriers).  But a frame-sized image is one of many in flight: two CTAs of this
    // kernel per SM do not fit next to three mean-shift CTAs (registers), and a cooperative grid holds its slots on EVERY SM while
    // it mostly waits at barriers -- the 4K bench step runs 4663 Mpix/s with 296 CTAs, 4728 with 148, 4764 with 74, 4758 with 37
    // (tools/overlap_probe.py, profiles/r02_overlap_probe.txt).  So: half a CTA per SM up to 2^24 pixels, the full grid beyond
    // (a single very large image has the GPU to itself).
    int grid = ctx->tune.merge_grid > 0 ? ctx->tune.merge_grid : (n <= ((size_t)1 << 24) ? (ctx->sm_count + 1) / 2 : grid_max);
    if (grid > grid_max) grid = grid_max;
    size_t nl = (size_t)cap + 1;
    size_t pair_cap = 2 * n;                        // every pixel has at most a right and a down neighbour
    int hbits = 16;                                 // pair set of the large path: one slot per 16 pixels, at least 2^16
    while (((size_t)1 << hbits) < n / 16 && hbits < 28) hbits++;
    const size_t hslots = (size_t)1 << hbits;
    size_t bytes = nl * (24 + 8 + 4 + 4 + 4 + 4) + (size_t)(grid + 1) * 4 + 256 + pair_cap * sizeof(int2) + hslots * 16 + 64;
    MSG_TRY(msg_reserve(ctx, (void**)&ctx->d_ovf, &ctx->d_ovf_cap, bytes));
    char* base = (char*)ctx->d_ovf;
    merge_persist_args A;
    A.t.sum = (unsigned long long*)base;   base += nl * 24;
    A.t.best = (unsigned long long*)base;  base += nl * 8;
    A.t.area = (unsigned int*)base;        base += nl * 4;
    A.t.mean = (uint32_t*)base;            base += nl * 4;
    A.t.par = (int32_t*)base;              base += nl * 4;
    A.newid = (int32_t*)base;              base += nl * 4;
    A.bsum = (int32_t*)base;               base += ((size_t)(grid + 1) * 4 + 15) / 16 * 16;
    A.pairs = (int2*)base;                 base += pair_cap * sizeof(int2);
    A.htab = (unsigned long long*)base;    base += hslots * 8;
    A.uniq = (int2*)base;
    A.hbits = hbits;
    A.nuniq = ctx->d_counters + 18;
    A.pair_cap = (long long)pair_cap;
    A.npairs = ctx->d_counters + 13;
    A.plane = d_plane; A.pitch = pitch; A.labels = d_labels; A.w = w; A.h = h;
    A.n_in = d_n_in; A.cap = cap;
    A.n_saved = ctx->d_counters + 14;
    A.accepted = ctx->d_counters + 9;
    A.rounds_out = ctx->d_counters + 10;
    A.n_out = d_n_out ? d_n_out : ctx->d_counters + 11;
    A.min_size = min_size; A.color_dist = color_dist;
    A.nin_host = -1; A.npairs_host = -1;
    A.vec = (w % 4 == 0 && (reinterpret_cast<uintptr_t>(d_labels) & 15) == 0 && !ctx->tune.merge_scalar) ? (ctx->tune.merge_strips ? 2 : 1) : 0;
    // test hooks: merge_small_max = v caps both single-CTA regimes (0 forces the cooperative path), merge_medium_only = 1 sends
    // every image of <= 8191 labels through the global-pair-set kernel
    A.small_max = ctx->tune.merge_small_max >= 0 ? ctx->tune.merge_small_max : SMALL_MAX_LABELS;
    if (A.small_max > SMALL_MAX_LABELS) A.small_max = SMALL_MAX_LABELS;
    A.medium_max = ctx->tune.merge_small_max >= 0 ? ctx->tune.merge_small_max : MEDIUM_MAX_LABELS;
    if (A.medium_max > MEDIUM_MAX_LABELS) A.medium_max = MEDIUM_MAX_LABELS;
    if (ctx->tune.merge_medium_only) { A.small_max = -1; A.medium_max = MEDIUM_MAX_LABELS; }
    if (MEDIUM_SMEM + 1024 > (size_t)ctx->max_smem_optin) A.medium_max = -1;
    cudaStream_t st = ctx->stream;
    const int wide = ctx->sm_count * 8;                              // CTAs of the streaming kernels (grid-stride)
    auto blocks_for = [&](size_t items) { size_t b = (items + MT - 1) / MT; return (int)(b < 1 ? 1 : (b > (size_t)wide ? (size_t)wide : b)); };
    // Who builds the set of pairs: a kernel of its own for frame-sized images (the single-CTA rounds kernels need it ready), the
    // cooperative rounds kernel itself beyond 2^24 pixels (always the large regime: one launch less, same work).
    A.set_ready = n <= (size_t)SMALL_MAX_PIXELS ? 1 : 0;
    if (A.set_ready) MSG_CUDA(ctx, cudaMemsetAsync(A.htab, 0, hslots * sizeof(unsigned long long), st));
    merge_init_kernel<<<blocks_for(n < 65536 ? n : 65536), MT, 0, st>>>(A);
    MSG_LAUNCHED(ctx);
    size_t stat_threads = A.vec ? ((n + 127) / 128) * 32 : ((size_t)((w + 31) / 32) * h) * 32;   // one warp per chunk
    merge_stats_kernel<<<blocks_for(stat_threads), MT, 0, st>>>(A);
    MSG_LAUNCHED(ctx);
    if (A.set_ready) {
        merge_pair_set_kernel<<<ctx->sm_count * 4, MT, 0, st>>>(A);
        MSG_LAUNCHED(ctx);
    }
    const size_t small_smem = (size_t)(SMALL_MAX_LABELS + 1) * 7 * sizeof(uint32_t);
    MSG_TRY(msg_func_smem(ctx, (const void*)merge_rounds_small_kernel<12>, small_smem));
    merge_rounds_small_kernel<12><<<1, ST, small_smem, st>>>(A);
    MSG_LAUNCHED(ctx);
    if (A.medium_max > A.small_max) {                                                        // 224 KB: the whole SM
        MSG_TRY(msg_func_smem(ctx, (const void*)merge_rounds_small_kernel<13>, MEDIUM_SMEM));
        merge_rounds_small_kernel<13><<<1, ST, MEDIUM_SMEM, st>>>(A);
        MSG_LAUNCHED(ctx);
    }
    void* args[] = {&A};
    MSG_CUDA(ctx, cudaLaunchCooperativeKernel((void*)merge_rounds_large_kernel, dim3(grid), dim3(MT), args, 0, st));
    MSG_LAUNCHED(ctx);
    merge_rewrite_kernel<<<blocks_for((n / 4 + 1) / 2 + 1), MT, 0, st>>>(A);
    MSG_LAUNCHED(ctx);
    MSG_CHECK_LAUNCH(ctx);
    return MSG_OK;
}

// Merge entry: d_n_in != NULL promises canonical labels 1..*d_n_in (the fused pipeline); otherwise the labels are
// validated (<= w*h, needs one stream sync) and renumbered canonically first.
int k_merge(msg_ctx* ctx, const uint32_t* d_plane, int pitch, int32_t* d_labels, int w, int h, int min_size,
            int color_dist, const int32_t* d_n_in, int32_t* d_n_out)
{
    cudaStream_t st = ctx->stream;
    size_t n = (size_t)w * h;
    int32_t* d_n_tmp = ctx->d_counters + 12;
    if (!d_n_in) {
        int32_t* d_max = ctx->d_counters + 8;
        MSG_CUDA(ctx, cudaMemsetAsync(d_max, 0, sizeof(int32_t), st));
        max_label_kernel<<<ctx->sm_count * 8, MT, 0, st>>>(d_labels, n, d_max);
        MSG_LAUNCHED(ctx);
        MSG_CUDA(ctx, cudaMemcpyAsync(ctx->h_counters + 8, d_max, sizeof(int32_t), cudaMemcpyDeviceToHost, st));
        MSG_CUDA(ctx, cudaStreamSynchronize(st));
        long long maxl = ctx->h_counters[8];
        if (maxl > (long long)n) return msg_fail(ctx, MSG_EINVAL, "merge: labels must be <= width*height (max label %lld)", maxl);
        MSG_TRY(k_relabel_canonical(ctx, d_labels, w, h, 0, d_n_tmp, 0));
        d_n_in = d_n_tmp;
    }
    if (min_size <= 0 && color_dist <= 0) {      // identity after canonical renumbering
        if (d_n_out && d_n_out != d_n_in)
            MSG_CUDA(ctx, cudaMemcpyAsync(d_n_out, d_n_in, sizeof(int32_t), cudaMemcpyDeviceToDevice, st));
        return MSG_OK;
    }
    MSG_TRY(merge_launch(ctx, d_plane, pitch, d_labels, w, h, min_size, color_dist, d_n_in, d_n_out));
    MSG_CUDA(ctx, cudaMemcpyAsync(ctx->h_counters + 10, ctx->d_counters + 10, sizeof(int32_t), cudaMemcpyDeviceToHost, st));
    return MSG_OK;
}


// ================================================================ strip-sharded merge (one very large image over several GPUs)
// regions are global (dense labels 1..n_total after the seam resolution); every rank accumulates the statistics
// and the adjacency list of ITS strip into caller-owned tables, the host sums the tables over the ranks (all-reduce) and
// concatenates the pair lists (all-gather), and every rank then runs the same rounds on the same tables -- identical decisions
// without any further exchange -- and rewrites its strip.  Integer sums are order independent, duplicates in the pair list are
// harmless (the rounds take minima over it), so the result equals the unsharded merge bit for bit.
namespace {

__global__ void __launch_bounds__(MT) strip_merge_init_kernel(unsigned int* __restrict__ area, unsigned long long* __restrict__ sum,
                                                              int nl, int32_t* __restrict__ npairs)
{
    const long long gtid = (long long)blockIdx.x * MT + threadIdx.x, nthreads = (long long)gridDim.x * MT;
    for (long long i = gtid; i < nl; i += nthreads) { area[i] = 0; sum[3 * i] = 0; sum[3 * i + 1] = 0; sum[3 * i + 2] = 0; }
    if (gtid == 0) *npairs = 0;
}

// adjacency across the seam above this strip: last row of the strip above (dense labels) against my first row
__global__ void __launch_bounds__(MT) seam_adjacency_kernel(const int32_t* __restrict__ up, const int32_t* __restrict__ lo, int w,
                                                            int nin, int2* __restrict__ pairs, long long cap,
                                                            int32_t* __restrict__ npairs)
{
    const int x = blockIdx.x * MT + threadIdx.x;
    int a = 0, b = 0;
    if (x < w) { a = up[x]; b = lo[x]; }
    bool emit = a > 0 && b > 0 && a != b && a <= nin && b <= nin;
    if (emit && x > 0 && up[x - 1] == a && lo[x - 1] == b) emit = false;      // same pair as the column to the left
    const unsigned m = __ballot_sync(0xffffffffu, emit);
    if (!m) return;
    const int lane = threadIdx.x & 31;
    long long pos = 0;
    if (lane == 0) pos = (long long)atomicAdd(npairs, __popc(m));
    pos = __shfl_sync(0xffffffffu, pos, 0) + __popc(m & ((1u << lane) - 1));
    if (emit && pos < cap) pairs[pos] = make_int2(a, b);
}

__global__ void __launch_bounds__(MT) strip_merge_par_kernel(int32_t* __restrict__ par, int nl, int32_t* __restrict__ accepted,
                                                             int32_t* __restrict__ rounds)
{
    const long long gtid = (long long)blockIdx.x * MT + threadIdx.x, nthreads = (long long)gridDim.x * MT;
    for (long long i = gtid; i < nl; i += nthreads) par[i] = (int)i;
    if (gtid == 0) { *accepted = 0; *rounds = 0; }
}

}  // namespace

int k_strip_merge_stats(msg_ctx* ctx, const uint32_t* d_plane, int pitch, const int32_t* d_labels, int w, int rows,
                        const int32_t* d_up_row_labels, int n_total, unsigned int* d_area, unsigned long long* d_sum,
                        int32_t* d_pairs, long long pair_cap, int32_t* d_npairs)
{
    merge_persist_args A;
    memset(&A, 0, sizeof(A));
    A.t.area = d_area; A.t.sum = d_sum;
    A.pairs = (int2*)d_pairs; A.pair_cap = pair_cap; A.npairs = d_npairs;
    A.plane = d_plane; A.pitch = pitch; A.labels = const_cast<int32_t*>(d_labels); A.w = w; A.h = rows;
    A.cap = n_total; A.nin_host = n_total; A.npairs_host = -1;
    A.n_saved = ctx->d_counters + 14;
    A.vec = (w % 4 == 0 && (reinterpret_cast<uintptr_t>(d_labels) & 15) == 0 && !ctx->tune.merge_scalar) ? (ctx->tune.merge_strips ? 2 : 1) : 0;
    cudaStream_t st = ctx->stream;
    const size_t n = (size_t)w * rows;
    const int wide = ctx->sm_count * 8;
    auto blocks_for = [&](size_t items) { size_t b = (items + MT - 1) / MT; return (int)(b < 1 ? 1 : (b > (size_t)wide ? (size_t)wide : b)); };
    strip_merge_init_kernel<<<blocks_for((size_t)n_total + 1), MT, 0, st>>>(d_area, d_sum, n_total + 1, d_npairs);
    MSG_LAUNCHED(ctx);
    size_t stat_threads = A.vec ? ((n + 127) / 128) * 32 : ((size_t)((w + 31) / 32) * rows) * 32;
    merge_stats_kernel<<<blocks_for(stat_threads), MT, 0, st>>>(A);
    MSG_LAUNCHED(ctx);
    if (d_up_row_labels) {
        seam_adjacency_kernel<<<(w + MT - 1) / MT, MT, 0, st>>>(d_up_row_labels, d_labels, w, n_total, (int2*)d_pairs, pair_cap, d_npairs);
        MSG_LAUNCHED(ctx);
    }
    MSG_CHECK_LAUNCH(ctx);
    return MSG_OK;
}

int k_strip_merge_finish(msg_ctx* ctx, int32_t* d_labels, int w, int rows, long long full_pixels, int n_total, unsigned int* d_area,
                         unsigned long long* d_sum, const int32_t* d_all_pairs, long long n_all_pairs, int min_size, int color_dist,
                         int32_t* d_n_out)
{
    if (ctx->merge_blocks_per_sm <= 0)       // occupancy of the cooperative kernel: queried once per context, not per launch
        MSG_CUDA(ctx, cudaOccupancyMaxActiveBlocksPerMultiprocessor(&ctx->merge_blocks_per_sm, merge_rounds_large_kernel, MT, 0));
    const int blocks_per_sm = ctx->merge_blocks_per_sm;
    if (blocks_per_sm < 1) return msg_fail(ctx, MSG_ECUDA, "merge: cooperative kernel does not fit");
    const int grid_max = ctx->sm_count * (blocks_per_sm < 2 ? blocks_per_sm : 2);
    int grid = ctx->tune.merge_grid > 0 ? ctx->tune.merge_grid : grid_max;
    if (grid > grid_max) grid = grid_max;
    const size_t nl = (size_t)n_total + 1;
    int hbits = 16;                                 // pair set: 8 slots per region (a planar adjacency graph has < 3 edges per region)
    while (((size_t)1 << hbits) < 8 * nl && hbits < 28) hbits++;
    const size_t hslots = (size_t)1 << hbits;
    const size_t bsum_bytes = ((size_t)(grid + 1) * 4 + 255) / 256 * 256;
    const size_t bytes = nl * (8 + 4 + 4 + 4) + bsum_bytes + 256 + hslots * 16;
    MSG_TRY(msg_reserve(ctx, (void**)&ctx->d_ovf, &ctx->d_ovf_cap, bytes));
    char* base = (char*)ctx->d_ovf;
    merge_persist_args A;
    memset(&A, 0, sizeof(A));
    A.t.area = d_area; A.t.sum = d_sum;
    A.t.best = (unsigned long long*)base;  base += nl * 8;
    A.t.mean = (uint32_t*)base;            base += nl * 4;
    A.t.par = (int32_t*)base;              base += nl * 4;
    A.newid = (int32_t*)base;              base += nl * 4;
    A.bsum = (int32_t*)base;               base += bsum_bytes;
    base = (char*)(((uintptr_t)base + 15) & ~(uintptr_t)15);
    A.htab = (unsigned long long*)base;    base += hslots * 8;
    A.uniq = (int2*)base;
    A.hbits = hbits;
    A.nuniq = ctx->d_counters + 18;
    A.pairs = (int2*)const_cast<int32_t*>(d_all_pairs); A.pair_cap = n_all_pairs; A.npairs = ctx->d_counters + 13;
    A.labels = d_labels; A.w = w; A.h = rows;
    A.cap = n_total; A.nin_host = n_total; A.npairs_host = n_all_pairs;
    A.n_saved = ctx->d_counters + 14;
    A.accepted = ctx->d_counters + 9;
    A.rounds_out = ctx->d_counters + 10;
    A.n_out = d_n_out ? d_n_out : ctx->d_counters + 11;
    A.min_size = min_size; A.color_dist = color_dist;
    A.small_max = ctx->tune.merge_small_max >= 0 ? ctx->tune.merge_small_max : SMALL_MAX_LABELS;
    if (A.small_max > SMALL_MAX_LABELS) A.small_max = SMALL_MAX_LABELS;
    A.medium_max = ctx->tune.merge_small_max >= 0 ? ctx->tune.merge_small_max : MEDIUM_MAX_LABELS;
    if (A.medium_max > MEDIUM_MAX_LABELS) A.medium_max = MEDIUM_MAX_LABELS;
    if (ctx->tune.merge_medium_only) { A.small_max = -1; A.medium_max = MEDIUM_MAX_LABELS; }
    if (MEDIUM_SMEM + 1024 > (size_t)ctx->max_smem_optin) A.medium_max = -1;
    // the single-CTA paths keep 32-bit colour sums: WHOLE image <= 2^24 pixels (merge_regime only sees the strip)
    if (full_pixels > SMALL_MAX_PIXELS) { A.small_max = -1; A.medium_max = -1; }
    cudaStream_t st = ctx->stream;
    const size_t n = (size_t)w * rows;
    const int wide = ctx->sm_count * 8;
    auto blocks_for = [&](size_t items) { size_t b = (items + MT - 1) / MT; return (int)(b < 1 ? 1 : (b > (size_t)wide ? (size_t)wide : b)); };
    strip_merge_par_kernel<<<blocks_for(nl), MT, 0, st>>>(A.t.par, (int)nl, A.accepted, A.rounds_out);
    MSG_LAUNCHED(ctx);
    const size_t small_smem = (size_t)(SMALL_MAX_LABELS + 1) * 7 * sizeof(uint32_t);
    MSG_TRY(msg_func_smem(ctx, (const void*)merge_rounds_small_kernel<12>, small_smem));
    merge_rounds_small_kernel<12><<<1, ST, small_smem, st>>>(A);
    MSG_LAUNCHED(ctx);
    if (A.medium_max > A.small_max) {                                                        // 224 KB: the whole SM
        MSG_TRY(msg_func_smem(ctx, (const void*)merge_rounds_small_kernel<13>, MEDIUM_SMEM));
        merge_rounds_small_kernel<13><<<1, ST, MEDIUM_SMEM, st>>>(A);
        MSG_LAUNCHED(ctx);
    }
    void* args[] = {&A};
    MSG_CUDA(ctx, cudaLaunchCooperativeKernel((void*)merge_rounds_large_kernel, dim3(grid), dim3(MT), args, 0, st));
    MSG_LAUNCHED(ctx);
    merge_rewrite_kernel<<<blocks_for((n / 4 + 1) / 2 + 1), MT, 0, st>>>(A);
    MSG_LAUNCHED(ctx);
    MSG_CHECK_LAUNCH(ctx);
    return MSG_OK;
}
