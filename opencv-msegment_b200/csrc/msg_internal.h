// msg_internal.h -- shared declarations of libmsegment_b200 (sm_100a).  Not part of the ABI.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>

#include "../../include/msegment.h"

#define MSG_MAX_LEVELS 9  // max_level in [0,8]

// A pyramid plane in HBM: one u32 per pixel, byte0=B byte1=G byte2=R byte3=flag.
//   S planes: byte3 = 1 for every pixel (so the packed word can be used directly as a
//             mean-shift "centre" and count accumulates through the same SIMD lane).
//   D planes: byte3 = change-mask of the level below maxLevel (0 = keep pyrUp value).
// `pitch` is in pixels (multiple of 32 -> rows start on 128-byte lines).
// A plane may hold only rows [y0, y0+rows) of a level whose full height is `hfull`
// (strip sharding); unsharded: y0 = 0, rows = hfull.
struct msg_plane {
    uint32_t* p;
    int w;       // level width
    int rows;    // rows stored
    int y0;      // global row of stored row 0
    int hfull;   // full level height
    int pitch;   // pixels per stored row
};

struct msg_ms_params {   // per level
    float sp;            // spatial radius of the level (float32, as OpenCV computes it)
    int radius;          // ceil(sp): max |offset| the window can reach
    int isr2;            // cvRound(sr*sr)
    int max_count;
    int ieps;            // floor(eps) clamped to int: (double)k <= eps  <=>  k <= ieps
    int use_mask;        // level < max_level: only pixels with D.byte3 != 0 run
};

struct msg_ovf_item {    // mean-shift item that left its staged tile; finished by the generic kernel
    uint32_t pix;        // plane-linear pixel index (row*pitch + x, stored rows)
    int16_t x0, y0rel;   // current window centre: x global, y relative to plane.y0 (fits: <= 32767 rows/plane)
    uint32_t c;          // packed colour, byte3 = 1
    uint32_t iter;       // iterations already done
};

// Tuning / experiment switches.  Read from the environment ONCE in msg_create (MSG_TILE_W, MSG_ACC, MSG_PITCH_RES, MSG_TMA,
// MSG_NO_ORDER, MSG_MERGE_SCALAR, MSG_MERGE_STRIPS, MSG_MERGE_SMALL_MAX, MSG_NO_GRAPH, MSG_GRAPH_DEBUG, MSG_CCL_LEGACY, MSG_CCL_QUAD) and changeable per context
// with msg_set_option; nothing on the launch path calls getenv.
struct msg_tuning {
    int tile_w;            // 0 = automatic, else 32 | 64
    int acc;               // -1 = automatic, else 0 | 1
    int pitch_res;         // -1 = automatic, else forced residue of the staged row pitch mod 32 banks
    int use_tma;           // 1
    int no_order;          // 0
    int merge_scalar;      // 0
    int merge_strips;      // 1 = statistics pass walks column strips with register-resident sums (0: row chunks + segmented shuffle reduction, A/B hook)
    int merge_small_max;   // -1 = compiled default
    int merge_medium_only; // test hook: 1 = images of <= 8191 labels use the single-CTA kernel with the global pair set
    int merge_grid;        // 0 = automatic, else CTAs of the cooperative large-path rounds kernel
    int no_graph, graph_debug;
    int ccl_legacy;        // 1 = row-run union-find of round 1 instead of the tile-local one
    int ccl_quad;          // 1 = the tile kernel with four pixels per lane where the rows allow it (0: one pixel per lane, A/B hook)
    int gray_compat;       // 0 = OpenCV 4.x 15-bit BGR2GRAY coefficients, 1 = OpenCV 3.4.2 14-bit ones
    int dt_legacy;         // 1 = the first wavefront kernel of the float distance transform (dt_wave_kernel) instead of dt_wave2_kernel
    int dt_fixed;          // 0 = IPP float chamfer arithmetic (cv2 4.13 build), 1 = OpenCV's own 16.16 fixed-point fallback
    int labels_canonical;  // 1 = msg_merge_regions_dev trusts that its input labels are canonical (1..*d_n, raster order of first pixel)
    int staging;           // 1 = pageable caller buffers go through the pinned staging ring (0: handed to cudaMemcpyAsync as is)
};

// Pinned staging ring for pageable caller buffers (SURVEY 8(b) "Ownership"): MSG_RING_CHUNKS chunks, each with the event of the
// DMA that last used it.
#define MSG_RING_CHUNKS 4
#define MSG_RING_CHUNK_BYTES ((size_t)4 << 20)

struct msg_ctx {
    int device;
    msg_tuning tune;
    // cudaFuncSetAttribute(MaxDynamicSharedMemorySize) is issued once per (kernel, size) and context, not per launch
    struct { const void* func; size_t smem; } attr_cache[64];
    int n_attr;
    int merge_blocks_per_sm;   // resident CTAs per SM of the cooperative merge kernel (0 = not queried yet)
    cudaStream_t own_stream;
    cudaStream_t stream;
    cudaEvent_t ev[8];
    char err[512];
    int cuda_failed;
    int sm_count;
    int max_smem_optin;

    // grow-only device workspace
    uint8_t* d_in;     size_t d_in_cap;      // raw BGR / mask upload
    uint8_t* d_out;    size_t d_out_cap;     // raw BGR download staging
    uint8_t* d_out2;   size_t d_out2_cap;    // rendered image staging
    int32_t* d_labels; size_t d_labels_cap;  // dense labels (w*h)
    uint32_t* d_planes; size_t d_planes_cap; // all S and D planes, carved per call
    msg_ovf_item* d_ovf; size_t d_ovf_cap;   // overflow queue (items)
    uint32_t* d_scratch; size_t d_scratch_cap; // CCL / scan / merge scratch (bytes)
    int32_t* d_counters;                     // 64 int32 device counters
    int32_t* h_counters;                     // pinned mirror
    uint8_t* d_colors; size_t d_colors_cap;
    uint8_t* d_aux;    size_t d_aux_cap;     // 8-bit planes of the seed generator (gray, blurred, classes, edges, ...)
    uint8_t* d_small;  size_t d_small_cap;   // small tables of the colour-seed generator / bilateral filter (histogram, spans, weights)
    uint8_t* d_ccl;    size_t d_ccl_cap;     // labelling stage: union-find forest (canonical path), root bitmap, block sums
    uint8_t* d_ws;     size_t d_ws_cap;      // watershed queues / per-image state (k_watershed.cu)
    int32_t* d_cells;  size_t d_cells_cap;   // active pixels per 32x32 cell of the current level + tile order (K1 scheduling)

    // pinned host staging for pageable caller buffers: a small ring for uploads and synchronous downloads
    uint8_t* h_ring;                          // MSG_RING_CHUNKS * MSG_RING_CHUNK_BYTES, allocated on first use
    cudaEvent_t ring_ev[MSG_RING_CHUNKS];
    int ring_busy[MSG_RING_CHUNKS];
    int ring_next;
    // explicitly registered caller ranges (msg_register_host): treated as pinned
    struct { const uint8_t* base; size_t bytes; } reg[16];
    int n_reg;

    uint64_t ws_epoch;        // bumped by every (re)allocation of a workspace buffer: invalidates captured graphs
    int no_events;            // 1 while enqueueing for the asynchronous path: no timing events (not meaningful there)

    msg_plane S[MSG_MAX_LEVELS], D[MSG_MAX_LEVELS];
    int last_levels;

    msg_timings tm;
    msg_stats st;

    // optional per-kernel profiling of the mean-shift levels (msg_set_profiling)
    int profiling;
    unsigned long long* d_work;          // [MSG_MAX_LEVELS][4] u64 work counters
    cudaEvent_t prof_ev[MSG_MAX_LEVELS][3];   // begin, after tile kernel, after overflow kernel
    int prof_pending[MSG_MAX_LEVELS];
    msg_kernel_profile prof;

    // asynchronous submissions (msg_submit_segment): upload, kernels and download of consecutive frames overlap on three
    // streams; every in-flight frame owns its device input / output buffers
    cudaStream_t h2d_stream, d2h_stream;
    struct pending {
        int used;
        int32_t* n_regions_host;  // pinned slot
        cudaEvent_t done;         // all downloads of the frame have landed
        cudaEvent_t ev_in, ev_core;   // upload finished / kernels finished
        uint8_t* d_in;   size_t d_in_cap;
        uint8_t* d_filt; size_t d_filt_cap;
        uint8_t* d_ren;  size_t d_ren_cap;
        int32_t* d_lab;  size_t d_lab_cap;
        uint16_t* d_lab16; size_t d_lab16_cap;   // labels_type = MSG_LABELS_16U
        int32_t* d_cnt;               // 16 device counters of this frame: [0] n_regions, [1] u16 overflow flag, [2..5] mean-shift stats
        int32_t* h_cnt;               // pinned mirror
        // deferred copies to pageable destinations: the frame's outputs land in pinned staging (h_out, grow-only) and are
        // copied to the caller's buffers by msg_wait
        uint8_t* h_out;  size_t h_out_cap;
        struct { void* dst; size_t dstep; size_t off; size_t row_bytes; int rows; } defer[3];
        int n_defer;
        int l16;                      // the frame's labels were requested as 16-bit
        // CUDA graph of the frame's kernel sequence (one launch call per frame instead of ~35); re-captured whenever
        // the geometry, the parameters, a buffer or the stream changes
        cudaGraphExec_t g_exec;
        int g_state;              // 0 nothing, 1 the configuration ran eagerly once (buffers sized), 2 g_exec valid, -1 disabled
        uint64_t g_epoch, g_launches;
        unsigned char g_key[160];
    } pend[MSG_MAX_INFLIGHT];
};

// ---------------------------------------------------------------- error helpers
int msg_fail(msg_ctx* ctx, int code, const char* fmt, ...);
#define MSG_CUDA(ctx, call)                                                                   \
    do {                                                                                      \
        cudaError_t e__ = (call);                                                             \
        if (e__ != cudaSuccess) {                                                             \
            /* argument / configuration / allocation errors are not sticky: the context stays usable */ \
            if (e__ != cudaErrorInvalidConfiguration && e__ != cudaErrorInvalidValue &&       \
                e__ != cudaErrorMemoryAllocation)                                             \
                (ctx)->cuda_failed = 1;                                                       \
            cudaGetLastError();                                                               \
            return msg_fail((ctx), MSG_ECUDA, "%s failed: %s (%s:%d)", #call,                 \
                            cudaGetErrorString(e__), __FILE__, __LINE__);                     \
        }                                                                                     \
    } while (0)
#define MSG_TRY(expr)                    \
    do {                                 \
        int rc__ = (expr);               \
        if (rc__ != MSG_OK) return rc__; \
    } while (0)

int msg_reserve(msg_ctx* ctx, void** p, size_t* cap, size_t bytes);
// opt-in dynamic shared memory of `func` is at least `smem` bytes on this context's device (cached per context)
int msg_func_smem(msg_ctx* ctx, const void* func, size_t smem);

static inline int msg_align_up(int v, int a) { return (v + a - 1) / a * a; }

// ---------------------------------------------------------------- kernel launchers (k_*.cu)
// conversions / pyramid
int k_bgr_to_plane(msg_ctx* ctx, const uint8_t* d_bgr, size_t step, msg_plane dst);          // all stored rows
int k_plane_to_bgr(msg_ctx* ctx, msg_plane src, int row_first, int nrows, uint8_t* d_bgr, size_t step);
int k_pyr_down(msg_ctx* ctx, msg_plane src, msg_plane dst);
int k_pyr_up_mask(msg_ctx* ctx, msg_plane dsrc /*D[l+1]*/, msg_plane ddst /*D[l]*/, int isr22, int* d_cell_count, int cells_x);
int k_synth(msg_ctx* ctx, uint8_t* d_bgr, size_t step, int w, int h, int row0, int rows, uint64_t seed);
// mean shift
int k_meanshift_level(msg_ctx* ctx, msg_plane S, msg_plane D, const msg_ms_params& prm, int level);
// profiling hooks (capi.cu): CUDA events around the tile kernel (slot 0) and the overflow kernel (slot 1) of a level
void msg_prof_begin(msg_ctx* ctx, int level);
void msg_prof_end(msg_ctx* ctx, int level, int slot);
// labelling
int k_ccl_color(msg_ctx* ctx, const uint32_t* d_plane, int pitch, int w, int h, int d, int conn, int32_t* d_labels,
                int64_t label_base, int lab_pitch);
int k_ccl_binary(msg_ctx* ctx, const uint8_t* d_mask, size_t step, int w, int h, int conn, int32_t* d_labels);
int k_relabel_canonical(msg_ctx* ctx, int32_t* d_labels, int w, int h, int roots_are_pixels,
                        int32_t* d_n_out /*device, may be NULL*/, int add_to_count);
int k_merge(msg_ctx* ctx, const uint32_t* d_plane, int pitch, int32_t* d_labels, int w, int h, int min_size,
            int color_dist, const int32_t* d_n_in /*NULL: labels not known canonical*/, int32_t* d_n_out);
int k_render(msg_ctx* ctx, const int32_t* d_labels, size_t lstep, uint8_t* d_dst, size_t dstep, int w, int h,
             int depth, const uint8_t* d_colors);
int k_copy_labels_2d(msg_ctx* ctx, const int32_t* src, size_t sstep, int32_t* dst, size_t dstep, int w, int h);
// colour-predicate labelling with canonical numbering (1..n in raster order of first pixel) in one call; *d_n = n
int k_label_canonical(msg_ctx* ctx, const uint32_t* d_plane, int pitch, int w, int h, int d, int conn, int32_t* d_labels,
                      int32_t* d_n);
int k_label_canonical_src(msg_ctx* ctx, const void* d_img, size_t pitch, int src_kind, int w, int h, int d, int conn,
                          int32_t* d_labels, int32_t* d_n);
int k_cc_canonical(msg_ctx* ctx, const uint8_t* d_mask, size_t step, int w, int h, int conn, int32_t* d_labels, int32_t* d_n);
// dense int32 labels -> 16-bit labels (saturating at 65535), dstep in bytes
int k_labels_to_u16(msg_ctx* ctx, const int32_t* d_labels, int w, int h, uint16_t* d_dst, size_t dstep);
int k_seam_pairs(msg_ctx* ctx, const uint8_t* up_bgr, const int32_t* up_lab, const uint8_t* lo_bgr,
                 const int32_t* lo_lab, int w, int d, int32_t* pairs, int32_t* count);
int k_apply_map(msg_ctx* ctx, int32_t* labels, size_t lstep, int w, int rows, const int32_t* from,
                const int32_t* to, int n);

int k_strip_rank(msg_ctx* ctx, const int32_t* d_labels, size_t lstep, int w, int rows, long long base, int32_t* d_count);
int k_strip_query(msg_ctx* ctx, const int32_t* d_q, int nq, int w, int rows, long long base, int offset, int32_t* d_out);
int k_strip_apply_dense(msg_ctx* ctx, int32_t* d_labels, size_t lstep, int w, int rows, long long base, int offset,
                        const int32_t* d_rlab, const int32_t* d_rdense, int nr);

int k_seam_quads(msg_ctx* ctx, const uint8_t* up_bgr, const int32_t* up_lab, const int32_t* up_rank1, const uint8_t* lo_bgr,
                 const int32_t* lo_lab, int w, int d, int rows, long long base, int32_t* quads, int32_t* count);
int k_strip_finalize(msg_ctx* ctx, int32_t* d_labels, size_t lstep, int w, int rows, long long base, int offset,
                     const int32_t* d_frm, const int32_t* d_dense, int nmap, int frm_lo);

// pre-filters (k_filters.cu)
int k_sharpen(msg_ctx* ctx, const uint8_t* d_src, size_t sstep, uint8_t* d_dst, size_t dstep, int w, int h, const int8_t* taps,
              int krows, int kcols);
int k_gray(msg_ctx* ctx, const uint8_t* d_src, size_t sstep, uint8_t* d_dst, size_t dstep, int w, int h);
int k_median(msg_ctx* ctx, const uint8_t* d_src, size_t sstep, uint8_t* d_dst, size_t dstep, int w, int h, int k);

// k_seeds.cu / k_ccl.cu: shape-method marker generator (8(f3))
int k_canny_nms(msg_ctx* ctx, const uint8_t* d_src, size_t sstep, uint8_t* d_cls, size_t cstep, int w, int h, int low, int high);
int k_hysteresis(msg_ctx* ctx, const uint8_t* d_cls, int w, int h, int32_t* d_labels, uint8_t* d_flag, uint8_t* d_dst, size_t dstep);
int k_dilate(msg_ctx* ctx, const uint8_t* d_src, size_t sstep, uint8_t* d_dst, size_t dstep, int w, int h, int kw, int kh);
int k_copy_masked(msg_ctx* ctx, const uint8_t* d_src, size_t sstep, const uint8_t* d_mask, size_t mstep, uint8_t* d_dst, size_t dstep,
                  int w, int h);
int k_subtract(msg_ctx* ctx, const uint8_t* d_a, size_t astep, const uint8_t* d_b, size_t bstep, uint8_t* d_dst, size_t dstep,
               int w, int h);

// k_colorseeds.cu / k_contours.cu: colour-method marker generator (8(f3), rows a6 / a4) and the bilateral pre-filter (a5)
int k_ccl_flatten(msg_ctx* ctx, int32_t* d_labels, size_t n);
int k_white_to_black(msg_ctx* ctx, const uint8_t* d_src, size_t sstep, uint8_t* d_dst, size_t dstep, int w, int h);
int k_otsu(msg_ctx* ctx, const uint8_t* d_src, size_t sstep, int w, int h, unsigned* d_hist, int32_t* d_thresh);
int k_threshold_u8(msg_ctx* ctx, const uint8_t* d_src, size_t sstep, uint8_t* d_dst, size_t dstep, int w, int h,
                   const int32_t* d_thresh, int thresh, int maxval);
int k_distance_transform_max_width(msg_ctx* ctx);
int k_distance_transform(msg_ctx* ctx, const uint8_t* d_src, size_t sstep, float* d_dist, int w, int h, float* d_max);
// k_dt_fixed.cu: the same call in OpenCV's own 16.16 fixed-point arithmetic (option "dt_fixed"); k_distance_transform dispatches
int k_distance_transform_fixed_max_dim();
int k_distance_transform_fixed(msg_ctx* ctx, const uint8_t* d_src, size_t sstep, float* d_dist, int w, int h, float* d_max);
int k_normalize_minmax_f32(msg_ctx* ctx, const float* d_src, float* d_dst, int w, int h, double alpha, double beta, float* d_mm);
int k_threshold_f32(msg_ctx* ctx, const float* d_src, float* d_dst, int w, int h, float thresh, float maxval);
int k_dilate_f32(msg_ctx* ctx, const float* d_src, float* d_dst, int w, int h, int kw, int kh);
int k_f32_to_u8(msg_ctx* ctx, const float* d_src, uint8_t* d_dst, size_t dstep, int w, int h);
int k_circle_filled_i32(msg_ctx* ctx, int32_t* d_img, size_t step, int w, int h, int cx, int cy, int radius, int32_t value,
                        int* d_spans, int* h_spans);
int k_bilateral(msg_ctx* ctx, const uint8_t* d_src, size_t sstep, uint8_t* d_dst, size_t dstep, int w, int h, int cn, int radius,
                int maxk, const float* d_space_w, const short* d_space_ofs, const float* d_color_w);
int k_contour_markers(msg_ctx* ctx, const uint8_t* d_mask, size_t step, int w, int h, int32_t* d_markers, size_t mstep,
                      int32_t* n_contours_host);

// k_shard.cu / k_merge.cu: device-side seam resolution and the strip-sharded merge
int k_strip_resolve(msg_ctx* ctx, const int32_t* d_gathered, int n_strips, int width, const int* row0, int32_t* d_tables);
int k_strip_finalize_tables(msg_ctx* ctx, int32_t* d_labels, size_t lstep, int w, int rows, long long base, const int32_t* d_tables,
                            int cap, int strip);
int k_strip_merge_stats(msg_ctx* ctx, const uint32_t* d_plane, int pitch, const int32_t* d_labels, int w, int rows,
                        const int32_t* d_up_row_labels, int n_total, unsigned int* d_area, unsigned long long* d_sum,
                        int32_t* d_pairs, long long pair_cap, int32_t* d_npairs);
int k_strip_merge_finish(msg_ctx* ctx, int32_t* d_labels, int w, int rows, long long full_pixels, int n_total, unsigned int* d_area,
                         unsigned long long* d_sum, const int32_t* d_all_pairs, long long n_all_pairs, int min_size, int color_dist,
                         int32_t* d_n_out);

// k_watershed.cu: exact cv::watershed, `count` images of one geometry per call
int k_watershed(msg_ctx* ctx, const uint8_t* d_bgr, size_t step, size_t image_stride, int32_t* d_markers, size_t mstep,
                size_t markers_stride, int w, int h, int count, unsigned long long* d_pops);

#define MSG_LAUNCHED(ctx) ((ctx)->st.kernel_launches++)
#define MSG_CHECK_LAUNCH(ctx) MSG_CUDA(ctx, cudaGetLastError())
