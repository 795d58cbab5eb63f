// k_shard.cu -- strip sharding of one very large image over several GPUs (BASELINE.json config 5):
// the host-side plan and the DEVICE-side seam resolution.
//
//   msg_shard_plan_make          rows / halo rows of every strip (pure host function; any host language can call it)
//   k_strip_resolve              from the all-gathered seam quads (A, B, rankA + 1, rankB + 1) and root counts of all strips
//                                to the tables msg_strip_finalize_tables_dev consumes, entirely on the device (round 1 did this
//                                step with numpy / scipy on the host after a .cpu() of the gathered buffer: 3.4 of 11.6 ms at
//                                16384^2).  One CTA: the data is a few thousand quads.
//   k_strip_finalize_tables      one pass over the strip: provisional labels -> dense global ids 1..N (raster order of first
//                                pixel, identical to the unsharded call), offsets / counts read from the device tables.
#include <math.h>

#include "msg_internal.h"

extern "C" int msg_shard_plan_make(int width, int height, int n_strips, double sp, int max_level, int term_type, int max_count,
                                   msg_shard_plan* out)
{
    if (!out || width <= 0 || height <= 0 || n_strips < 1 || n_strips > MSG_MAX_STRIPS || max_level < 0 || max_level > 8)
        return MSG_EINVAL;
    const int a = 1 << max_level;
    const long long units = ((long long)height + a - 1) / a;
    if (units < n_strips) return MSG_EINVAL;                 // image too small for that many strips at this alignment
    const int halo = msg_meanshift_halo_rows(sp, max_level, term_type, max_count);
    if (halo < 0) return MSG_EINVAL;
    memset(out, 0, sizeof(*out));
    out->n_strips = n_strips; out->width = width; out->height = height; out->max_level = max_level; out->halo_rows = halo;
    for (int k = 0; k <= n_strips; k++) {
        long long b = k == n_strips ? height : (units * k / n_strips) * a;
        if (b > height) b = height;
        if (k < n_strips) out->row0[k] = (int)b;
        if (k > 0) out->row1[k - 1] = (int)b;
    }
    for (int k = 0; k < n_strips; k++) {
        int h0 = out->row0[k] - halo;
        if (h0 < 0) h0 = 0;
        h0 -= h0 % a;
        long long h1 = (long long)out->row1[k] + halo;
        if (h1 > height) h1 = height;
        out->halo0[k] = h0;
        out->halo1[k] = (int)h1;
    }
    return MSG_OK;
}

namespace {

constexpr int RT = 1024;                     // threads of the resolve CTA
constexpr int SORT_SMEM_KEYS = 16384;        // keys sorted in shared memory (128 KiB); longer lists are sorted in HBM

struct resolve_args {
    const int32_t* gathered;     // [n_strips][width + 1][4]: row 0 = (quad count, root count, 0, 0), rows 1.. = quads
    int n_strips, width;
    int row0[MSG_MAX_STRIPS + 1];
    int32_t* tables;             // MSG_SHARD_TABLE_HEADER ints, then frm[cap], dense[cap]
    int cap;                     // n_strips * width
    unsigned long long* keys;    // [P] scratch (P = padded 2 * quads)
    int32_t* ulab; int32_t* urank; int32_t* parent; int32_t* rootidx; int32_t* froot;   // [2 * cap] each
};

__device__ __forceinline__ int lower_bound_i32(const int32_t* a, int n, long long v)
{
    int lo = 0, hi = n;
    while (lo < hi) { int mid = (lo + hi) >> 1; if ((long long)a[mid] < v) lo = mid + 1; else hi = mid; }
    return lo;
}

__device__ int rs_find(int32_t* par, int a)
{
    int p = par[a];
    while (p != a) { int g = par[p]; if (g != p) par[a] = g; a = p; p = g; }
    return a;
}

// exclusive scan of per-thread counts over the CTA (1024 threads); returns the thread's offset, *total the sum
__device__ int cta_exclusive_scan(int v, int* total)
{
    __shared__ int wsum[32];
    __shared__ int s_total;
    const int lane = threadIdx.x & 31, wq = threadIdx.x >> 5;
    int incl = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) { int t = __shfl_up_sync(0xffffffffu, incl, o); if (lane >= o) incl += t; }
    if (lane == 31) wsum[wq] = incl;
    __syncthreads();
    if (wq == 0) {
        int x = wsum[lane], i2 = x;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) { int t = __shfl_up_sync(0xffffffffu, i2, o); if (lane >= o) i2 += t; }
        wsum[lane] = i2 - x;
        if (lane == 31) s_total = i2;
    }
    __syncthreads();
    const int off = wsum[wq] + incl - v;
    if (total) *total = s_total;
    __syncthreads();
    return off;
}

__global__ void __launch_bounds__(RT, 1) strip_resolve_kernel(resolve_args A)
{
    extern __shared__ unsigned long long s_keys[];
    __shared__ int s_qoff[MSG_MAX_STRIPS + 1], s_nroots[MSG_MAX_STRIPS], s_frm_lo[MSG_MAX_STRIPS + 1], s_off[MSG_MAX_STRIPS + 1];
    __shared__ int s_n, s_u, s_nmap;
    const int tid = threadIdx.x;
    const size_t stride = (size_t)(A.width + 1) * 4;
    int32_t* frm = A.tables + MSG_SHARD_TABLE_HEADER;
    int32_t* dense = frm + A.cap;
    if (tid == 0) {
        int run = 0;
        for (int s = 0; s < A.n_strips; s++) {
            int q = A.gathered[s * stride];
            if (q < 0) q = 0;
            if (q > A.width) q = A.width;
            s_qoff[s] = run; run += q;
            s_nroots[s] = A.gathered[s * stride + 1];
        }
        s_qoff[A.n_strips] = run;
        s_n = run;
    }
    __syncthreads();
    const int n = s_n;                                        // quads over all seams
    int nmap = 0;
    if (n > 0) {
        // ---- (label << 32 | rank + 1) of both sides of every quad, sorted: equal labels become neighbours
        int P = 1;
        while (P < 2 * n) P <<= 1;
        unsigned long long* K = P <= SORT_SMEM_KEYS ? s_keys : A.keys;
        for (int s = 0; s < A.n_strips; s++) {
            const int q0 = s_qoff[s], qn = s_qoff[s + 1] - q0;
            const int32_t* Q = A.gathered + s * stride + 4;
            for (int i = tid; i < qn; i += RT) {
                K[2 * (q0 + i)] = ((unsigned long long)(uint32_t)Q[4 * i] << 32) | (uint32_t)Q[4 * i + 2];
                K[2 * (q0 + i) + 1] = ((unsigned long long)(uint32_t)Q[4 * i + 1] << 32) | (uint32_t)Q[4 * i + 3];
            }
        }
        for (int i = 2 * n + tid; i < P; i += RT) K[i] = ~0ull;
        __syncthreads();
        for (int k = 2; k <= P; k <<= 1)
            for (int j = k >> 1; j > 0; j >>= 1) {
                for (int i = tid; i < P; i += RT) {
                    const int ixj = i ^ j;
                    if (ixj > i) {
                        const unsigned long long a = K[i], b = K[ixj];
                        if (((i & k) == 0) ? (a > b) : (a < b)) { K[i] = b; K[ixj] = a; }
                    }
                }
                __syncthreads();
            }
        // ---- distinct labels (ascending) with the strip-local rank of their root
        const int per = (2 * n + RT - 1) / RT;
        const int i0 = tid * per, i1 = min(2 * n, i0 + per);
        int cnt = 0;
        for (int i = i0; i < i1; i++) cnt += (i == 0 || (K[i] >> 32) != (K[i - 1] >> 32)) ? 1 : 0;
        int total;
        int pos = cta_exclusive_scan(cnt, &total);
        for (int i = i0; i < i1; i++)
            if (i == 0 || (K[i] >> 32) != (K[i - 1] >> 32)) {
                A.ulab[pos] = (int32_t)(K[i] >> 32);
                A.urank[pos] = (int32_t)(K[i] & 0xffffffffu) - 1;
                A.parent[pos] = pos;
                pos++;
            }
        if (tid == 0) s_u = total;
        __syncthreads();
        const int u = s_u;
        // ---- union-find over the distinct labels, smaller index = smaller label wins
        for (int s = 0; s < A.n_strips; s++) {
            const int qn = s_qoff[s + 1] - s_qoff[s];
            const int32_t* Q = A.gathered + s * stride + 4;
            for (int i = tid; i < qn; i += RT) {
                int a = lower_bound_i32(A.ulab, u, Q[4 * i]), b = lower_bound_i32(A.ulab, u, Q[4 * i + 1]);
                for (;;) {
                    a = rs_find(A.parent, a);
                    b = rs_find(A.parent, b);
                    if (a == b) break;
                    if (a < b) { int t = a; a = b; b = t; }
                    const int old = atomicMin(A.parent + a, b);
                    if (old == a) break;
                    a = old;
                }
            }
        }
        __syncthreads();
        // ---- labels that merge into a smaller one: frm (ascending, ulab is sorted), froot = index of the class' label
        const int per2 = (u + RT - 1) / RT;
        const int j0 = tid * per2, j1 = min(u, j0 + per2);
        int c2 = 0;
        for (int i = j0; i < j1; i++) {
            int r = i;
            while (A.parent[r] != r) r = A.parent[r];
            A.rootidx[i] = r;
            c2 += r != i;
        }
        int pos2 = cta_exclusive_scan(c2, &total);
        for (int i = j0; i < j1; i++)
            if (A.rootidx[i] != i) { frm[pos2] = A.ulab[i]; A.froot[pos2] = A.rootidx[i]; pos2++; }
        nmap = total;
    }
    if (tid == 0) s_nmap = nmap;
    __syncthreads();
    nmap = s_nmap;
    // ---- per strip: first entry of frm inside it, surviving roots above it; total
    if (tid <= A.n_strips)
        s_frm_lo[tid] = tid == A.n_strips ? nmap : lower_bound_i32(frm, nmap, (long long)A.row0[tid] * A.width + 1);
    __syncthreads();
    if (tid == 0) {
        int run = 0;
        for (int s = 0; s < A.n_strips; s++) {
            s_off[s] = run;
            run += s_nroots[s] - (s_frm_lo[s + 1] - s_frm_lo[s]);
        }
        s_off[A.n_strips] = run;
        A.tables[0] = nmap;
        A.tables[1] = run;                                    // regions of the whole image
        for (int s = 0; s < MSG_MAX_STRIPS; s++) {
            A.tables[2 + s] = s < A.n_strips ? s_off[s] : 0;
            A.tables[2 + MSG_MAX_STRIPS + s] = s <= A.n_strips ? s_frm_lo[s] : nmap;
        }
        A.tables[2 + 2 * MSG_MAX_STRIPS] = nmap;
    }
    __syncthreads();
    // ---- dense id of the class every merged label joins
    for (int k = tid; k < nmap; k += RT) {
        const int r = A.froot[k];
        const long long t = A.ulab[r];
        int strip = 0;
        while (strip + 1 < A.n_strips && (long long)A.row0[strip + 1] * A.width + 1 <= t) strip++;
        const int removed_before = lower_bound_i32(frm, nmap, t) - s_frm_lo[strip];
        dense[k] = s_off[strip] + A.urank[r] - removed_before + 1;
    }
}

__global__ void __launch_bounds__(256) strip_finalize_tables_kernel(int32_t* __restrict__ L, size_t lstep_words, int w, size_t n,
                                                                    long long base, const int32_t* __restrict__ block_offs,
                                                                    const int32_t* __restrict__ lrank,
                                                                    const int32_t* __restrict__ tables, int cap, int strip)
{
    const size_t i = (size_t)blockIdx.x * 256 + threadIdx.x;
    if (i >= n) return;
    const int nmap = __ldg(tables), offset = __ldg(tables + 2 + strip), frm_lo = __ldg(tables + 2 + MSG_MAX_STRIPS + strip);
    const int32_t* __restrict__ frm = tables + MSG_SHARD_TABLE_HEADER;
    const int32_t* __restrict__ dense = frm + cap;
    int32_t* p = L + (i / w) * lstep_words + (i % w);
    const int v = *p;
    if (v <= 0) return;
    int lo = 0, hi = nmap;
    while (lo < hi) { int mid = (lo + hi) >> 1; if (__ldg(frm + mid) < v) lo = mid + 1; else hi = mid; }
    if (lo < nmap && __ldg(frm + lo) == v) { *p = __ldg(dense + lo); return; }
    const long long loc = (long long)v - 1 - base;
    if (loc < 0 || loc >= (long long)n) return;
    *p = offset + block_offs[loc / 4096] + lrank[loc] - (lo - frm_lo) + 1;
}

}  // namespace

int k_strip_resolve(msg_ctx* ctx, const int32_t* d_gathered, int n_strips, int width, const int* row0, int32_t* d_tables)
{
    resolve_args A;
    memset(&A, 0, sizeof(A));
    A.gathered = d_gathered; A.n_strips = n_strips; A.width = width; A.tables = d_tables;
    A.cap = n_strips * width;
    for (int s = 0; s < n_strips; s++) A.row0[s] = row0[s];
    size_t P = 1;
    while (P < 2 * (size_t)A.cap) P <<= 1;
    const size_t need = P * 8 + 5 * (size_t)(2 * A.cap) * 4 + 256;
    MSG_TRY(msg_reserve(ctx, (void**)&ctx->d_ccl, &ctx->d_ccl_cap, need));
    char* b = (char*)ctx->d_ccl;
    A.keys = (unsigned long long*)b; b += P * 8;
    A.ulab = (int32_t*)b; b += (size_t)2 * A.cap * 4;
    A.urank = (int32_t*)b; b += (size_t)2 * A.cap * 4;
    A.parent = (int32_t*)b; b += (size_t)2 * A.cap * 4;
    A.rootidx = (int32_t*)b; b += (size_t)2 * A.cap * 4;
    A.froot = (int32_t*)b;
    const size_t smem = (size_t)SORT_SMEM_KEYS * 8;
    MSG_TRY(msg_func_smem(ctx, (const void*)strip_resolve_kernel, smem));
    strip_resolve_kernel<<<1, RT, smem, ctx->stream>>>(A);
    MSG_LAUNCHED(ctx);
    MSG_CHECK_LAUNCH(ctx);
    return MSG_OK;
}

// lrank / block sums of this strip are the ones msg_strip_rank_dev left in ctx->d_scratch
int k_strip_finalize_tables(msg_ctx* ctx, int32_t* d_labels, size_t lstep, int w, int rows, long long base, const int32_t* d_tables,
                            int cap, int strip)
{
    const size_t n = (size_t)w * rows;
    int32_t* lrank = (int32_t*)ctx->d_scratch;
    strip_finalize_tables_kernel<<<(unsigned)((n + 255) / 256), 256, 0, ctx->stream>>>(d_labels, lstep / 4, w, n, base, lrank + n, lrank,
                                                                                      d_tables, cap, strip);
    MSG_LAUNCHED(ctx);
    MSG_CHECK_LAUNCH(ctx);
    return MSG_OK;
}
