/*
 * msegment.h -- C ABI of libmsegment_b200.so: the B200-native (sm_100a CUDA) segmentation hot path
 * behind the OpenCV `Imgproc` calls that ShayHulud/opencv-msegment drives through OpenCV's Java
 * bindings.  Plain C: pointers, sizes, ints and doubles only, so that JNI, Panama FFI, ctypes and
 * cffi bind it directly (INTEGRATION.md shows the Java side).
 *
 * Reference interfaces replaced (paths relative to /root/reference; the JNI layer itself is the
 * un-vendored org.openpnp:opencv:3.4.2-1, pom.xml:39-43):
 *   Imgproc.connectedComponents(mask, markers, 8, CV_32S)
 *        src/main/java/ru/shayhulud/opencvcmsegment/service/PictureService.java:441-442
 *        -> msg_connected_components
 *   PictureService.colorByIndexes(markers, depth, colored)         PictureService.java:913-936
 *        -> msg_render_labels
 *   Imgproc.pyrMeanShiftFiltering(src, dst, sp, sr[, maxLevel, termcrit])   (no call site in the
 *        reference; named by BASELINE.json north_star subsystem 1)   -> msg_meanshift_filter
 *   Imgproc.floodFill(...) region-growing loop (north_star subsystem 2; canonical use: OpenCV
 *        samples/cpp/meanshift_segmentation.cpp)                     -> msg_label_regions
 *   colour-distance + min-size region merge (north_star subsystem 2, self-defined spec in
 *        DESIGN.md "K2b")                                            -> msg_merge_regions
 * The CLI contract of App.main (App.java:14-31: input folder, output root, file name) is kept by
 * the host program on top of this ABI (opencv-msegment_b200/host).
 *
 * Conventions
 *   - every function returns MSG_OK (0) or a negative msg_status; the message of the last failure
 *     of a context is msg_last_error(ctx).  Nothing aborts or throws across the ABI.
 *   - images are row-major with a row step in BYTES (cv::Mat::step): 8UC3 BGR interleaved,
 *     8UC1 masks, 32SC1 labels.  The caller owns all buffers; the library never retains them
 *     past the call (or past msg_wait for submitted work).
 *   - a msg_ctx is bound to one CUDA device and is NOT thread-safe; use one per thread / GPU.
 *   - there is no CPU fallback: without a CUDA device msg_create fails with MSG_ECUDA.
 */
#ifndef MSEGMENT_H
#define MSEGMENT_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define MSG_VERSION 200 /* 0.2.0: msg_segment_params.labels_type, options, host registration, watershed, sharding plan */

#if defined(__GNUC__)
#define MSG_API __attribute__((visibility("default")))
#else
#define MSG_API
#endif

typedef enum msg_status {
    MSG_OK = 0,
    MSG_EINVAL = -1,  /* bad argument (mirrors OpenCV's CV_Assert failures -> CvException) */
    MSG_ECUDA = -2,   /* CUDA runtime / driver error (sticky per context) */
    MSG_ENOMEM = -3,  /* host or device allocation failed */
    MSG_ESTATE = -4,  /* call sequence error (bad ticket, ...) */
    MSG_ERANGE = -5   /* result does not fit the requested output type (more than 65535 regions for 16-bit labels) */
} msg_status;

/* cv::TermCriteria type bits */
#define MSG_TERM_COUNT 1
#define MSG_TERM_EPS 2

typedef struct msg_ctx msg_ctx;

/* ---- lifecycle ------------------------------------------------------------------------- */
MSG_API int msg_version(void);
MSG_API int msg_device_count(void);                       /* number of CUDA devices, 0 if none / error */
MSG_API int msg_create(int device, msg_ctx** out);
MSG_API void msg_destroy(msg_ctx* ctx);
MSG_API const char* msg_last_error(const msg_ctx* ctx);   /* ctx may be NULL: last msg_create error */
/* Use an externally owned cudaStream_t for all work of this context (NULL = the context's own stream;
 * pass cudaStreamLegacy, (void*)1, for the legacy default stream). */
MSG_API int msg_set_stream(msg_ctx* ctx, void* cuda_stream);
MSG_API int msg_synchronize(msg_ctx* ctx);

/* Per-context options (integers).  Every option is initialised once in msg_create from the environment variable MSG_<NAME>
 * (upper case) when that is set; nothing on the launch path reads the environment.
 *   "gray_compat"     0 (default): cvtColor(BGR2GRAY) with the OpenCV 4.x coefficients (3735, 19235, 9798) >> 15 -- what the
 *                     cv2 4.13 oracle is pinned on; 1: the OpenCV 3.4.2 ones (1868, 9617, 4899) >> 14 -- the version the
 *                     reference binds (pom.xml:39-43); differs by 1 LSB on 0.26 % of colours (INTEGRATION.md "version skew")
 *   "dt_fixed"        0 (default): distanceTransform(L2, 5) in the float arithmetic of the IPP-backed cv2 4.13 build;
 *                     1: OpenCV's own 16.16 fixed-point chamfer (what a non-IPP build, e.g. the openpnp 3.4.2 natives, runs)
 *   "labels_canonical" 0 (default): msg_merge_regions_dev validates and renumbers arbitrary positive labels (two extra passes over
 *                     the pixels); 1: the caller vouches that the labels are canonical -- 1..n in raster order of first pixel, what
 *                     msg_label_regions / msg_connected_components write -- and passes n in *d_n_regions
 *   "dt_legacy"       1: the first wavefront kernel of the float distance transform (A/B hook)
 *   "staging"         1 (default): pageable caller buffers are staged through the context's pinned ring; 0: handed to
 *                     cudaMemcpyAsync as they are (driver staging, serialises the asynchronous path)
 *   "ccl_quad"        1 (default): the labelling's tile kernel handles four pixels per lane where the rows allow aligned vector
 *                     accesses (width a multiple of 4); 0: one pixel per lane everywhere (A/B hook, same labels)
 *   "merge_strips"    1 (default): the merge's statistics pass walks column strips with register-resident sums; 0: row chunks
 *                     with a segmented shuffle reduction (A/B hook, same tables)
 *   "merge_small_max", "merge_medium_only", "merge_grid", "tile_w", "acc", "pitch_res", "tma", "no_order", "merge_scalar", "no_graph", "ccl_legacy":
 *                     tuning / test hooks */
MSG_API int msg_set_option(msg_ctx* ctx, const char* name, int value);
MSG_API int msg_get_option(msg_ctx* ctx, const char* name, int* value);

/* Page-lock a caller-owned host range (a long-lived Mat: cv::Mat data is pageable fastMalloc memory) so that every later call
 * that passes a buffer inside it copies straight from / to it asynchronously (no staging copy).  The caller MUST unregister
 * before freeing the memory.  Up to 16 ranges per context.  Unregistered pageable buffers are always legal: they go through the
 * context's pinned staging ring (uploads: chunked memcpy + DMA overlapped; msg_submit_segment downloads: DMA into pinned staging,
 * copied to the caller's buffer inside msg_wait). */
MSG_API int msg_register_host(msg_ctx* ctx, void* ptr, size_t bytes);
MSG_API int msg_unregister_host(msg_ctx* ctx, void* ptr);

/* ---- drop-in operators on HOST buffers (synchronous; staging + H2D/D2H inside) ---------- */

/* Imgproc.pyrMeanShiftFiltering.  src/dst 8UC3 (dst may equal src).  OpenCV rules: max_level in
 * [0,8]; term_type without COUNT -> max_count=5, clamp to [1,100]; without EPS -> eps=1; eps>=0.
 * Java's 4-argument overload = (max_level=1, term_type=COUNT|EPS, max_count=5, eps=1). */
MSG_API int msg_meanshift_filter(msg_ctx* ctx, const uint8_t* src_bgr, size_t src_step, uint8_t* dst_bgr,
                         size_t dst_step, int width, int height, double sp, double sr, int max_level,
                         int term_type, int max_count, double eps);

/* floodFill-style region growing over the whole image: regions = connected components of
 * "adjacent (4- or 8-neighbourhood) and per-channel |delta| <= lo_diff"; lo_diff must equal up_diff (asymmetric ranges make
 * floodFill order dependent -> MSG_EINVAL); connectivity 4 or 8 (floodFill's flag).  labels 32SC1, numbered 1..n in
 * raster order of each region's first pixel (== the floodFill loop's numbering). */
MSG_API int msg_label_regions(msg_ctx* ctx, const uint8_t* bgr, size_t step, int32_t* labels, size_t labels_step,
                      int width, int height, int lo_diff, int up_diff, int connectivity,
                      int32_t* n_regions);

/* Region merge on (image, labels): colour fuse (color_dist > 0) then min-size prune (min_size > 0);
 * labels in/out, renumbered canonically.  Labels <= 0 are left untouched. */
MSG_API int msg_merge_regions(msg_ctx* ctx, const uint8_t* bgr, size_t step, int32_t* labels, size_t labels_step,
                      int width, int height, int min_size, int color_dist, int32_t* n_regions);

/* Imgproc.connectedComponents(image 8UC1, labels 32SC1, connectivity 4|8, CV_32S).
 * n_labels = number of labels including background 0 (OpenCV's return value).  Foreground labels
 * are numbered 1.. in raster order of first pixel (OpenCV's numbering is implementation defined). */
MSG_API int msg_connected_components(msg_ctx* ctx, const uint8_t* mask, size_t step, int32_t* labels,
                             size_t labels_step, int width, int height, int connectivity,
                             int32_t* n_labels);

/* PictureService.colorByIndexes: dst(8UC3) = (0 < label <= depth) ? colors[label-1] : (0,0,0);
 * colors_bgr = depth*3 bytes, or NULL for all white (the CLI path, colored=false). */
MSG_API int msg_render_labels(msg_ctx* ctx, const int32_t* labels, size_t labels_step, uint8_t* dst_bgr,
                      size_t dst_step, int width, int height, int depth, const uint8_t* colors_bgr);

/* Imgproc.watershed(image 8UC3, markers 32SC1 in/out) -- PictureService.java:908-911 (call sites :284, :372, :457, :852): the
 * reference's region-growing step.  EXACT emulation of cv::watershed: the border and every pixel where two
 * basins meet become -1, positive seeds flood the zero pixels in priority order of the max-channel difference to the pushing
 * neighbour, FIFO inside a level; unreachable pixels stay 0.  The flood is inherently sequential, so one image runs on one warp
 * (about the speed of one CPU core); msg_watershed_batch_dev floods `count` images of one geometry concurrently, which is where
 * the GPU pays (DESIGN.md "K4").  msg_get_timings().filter_ms = kernel time of the last msg_watershed call. */
MSG_API int msg_watershed(msg_ctx* ctx, const uint8_t* image_bgr, size_t step, int32_t* markers, size_t markers_step, int width,
                  int height);
MSG_API int msg_watershed_dev(msg_ctx* ctx, const uint8_t* d_image_bgr, size_t step, int32_t* d_markers, size_t markers_step,
                      int width, int height);
/* image k at d_images_bgr + k * image_stride bytes, its markers at d_markers + k * markers_stride bytes.  d_pops (device
 * uint64, may be NULL) receives the total number of queue pops = the length of the sequential chains. */
MSG_API int msg_watershed_batch_dev(msg_ctx* ctx, const uint8_t* d_images_bgr, size_t step, size_t image_stride, int32_t* d_markers,
                            size_t markers_step, size_t markers_stride, int width, int height, int count,
                            unsigned long long* d_pops);

/* ---- pre-filters the reference calls around its segmentation stage (exact integer forms) ------------- */
/* Laplacian sharpen chain of PictureService.java:323-333: filter2D(src, lap, CV_32F, K); src.convertTo(CV_32F);
 * subtract; convertTo(CV_8UC3)  ==  dst = saturate_u8(src - sum_K taps * src), BORDER_REFLECT_101, anchor at the kernel centre.
 * taps: krows*kcols small integers, row-major; the reference's MatOfFloat(1,1,1,1,-8,1,1,1,1) is 9x1 read literally and 3x3 as
 * intended -- the kernel is data, so either reading is one call.  krows, kcols odd, krows*kcols <= 1024. */
MSG_API int msg_laplacian_sharpen(msg_ctx* ctx, const uint8_t* src_bgr, size_t src_step, uint8_t* dst_bgr, size_t dst_step,
                          int width, int height, const int8_t* taps, int krows, int kcols);
/* Imgproc.cvtColor(src, dst, COLOR_BGR2GRAY) (PictureService.java:405, :940): (3735 B + 19235 G + 9798 R + 16384) >> 15,
 * the cv2 4.13 fixed-point form (OpenCV 3.4.2's scalar path differs by 1 LSB on 0.26 % of colours). */
MSG_API int msg_bgr2gray(msg_ctx* ctx, const uint8_t* src_bgr, size_t src_step, uint8_t* dst_gray, size_t dst_step, int width,
                 int height);
/* Imgproc.medianBlur(src 8UC1, dst, ksize) (PictureService.java:408, :436): odd ksize in [1, 127], BORDER_REPLICATE. */
MSG_API int msg_median_blur(msg_ctx* ctx, const uint8_t* src_gray, size_t src_step, uint8_t* dst_gray, size_t dst_step, int width,
                    int height, int ksize);

/* ---- shape-method marker generator (PictureService.java:404-442) ------------------------------ */
/* Imgproc.Canny(image 8UC1, edges, threshold1, threshold2) (PictureService.java:416): aperture 3, L1 gradient; thresholds are
 * swapped if threshold1 > threshold2 and floored, as cv::Canny does.  edges = 0 / 255. */
MSG_API int msg_canny(msg_ctx* ctx, const uint8_t* src_gray, size_t src_step, uint8_t* dst_edges, size_t dst_step, int width,
              int height, double threshold1, double threshold2);
/* Imgproc.dilate(src 8UC1, dst, Mat.ones(kh, kw)) (PictureService.java:428-429): anchor at the centre, border pixels ignored. */
MSG_API int msg_dilate(msg_ctx* ctx, const uint8_t* src, size_t src_step, uint8_t* dst, size_t dst_step, int width, int height,
               int kw, int kh);
/* Core.subtract(a, b, dst) on 8UC1 (PictureService.java:430): saturating. */
MSG_API int msg_subtract(msg_ctx* ctx, const uint8_t* a, size_t a_step, const uint8_t* b, size_t b_step, uint8_t* dst,
                 size_t dst_step, int width, int height);
/* src.copyTo(dst, mask) onto Mat.zeros (PictureService.java:417-418, the "borders" Result): dst = mask != 0 ? src : 0. */
MSG_API int msg_copy_masked(msg_ctx* ctx, const uint8_t* src_bgr, size_t src_step, const uint8_t* mask, size_t mask_step,
                    uint8_t* dst_bgr, size_t dst_step, int width, int height);
/* The whole generator, intermediates in HBM: cvtColor(BGR2GRAY) -> medianBlur(median_ksize) -> Canny(t1, t2) -> dilate 3x3 ->
 * dilate 5x5 -> subtract -> medianBlur 3 -> connectedComponents(8, CV_32S).  markers: 0 background, 1..n-1 in raster order of
 * first pixel; *n_labels counts the background like OpenCV.  stages (optional, NULL to skip): 4 planes of `height` rows of
 * stage_step bytes each, back to back: blurred gray, Canny edges, dilate-dilate-subtract band, its 3x3 median -- the images
 * the reference saves as "blured_by_KxK", "gray_borders", "dde_step", "dde_step_blurred_3x3". */
MSG_API int msg_shape_seeds(msg_ctx* ctx, const uint8_t* src_bgr, size_t src_step, int width, int height, int median_ksize,
                    double threshold1, double threshold2, int32_t* markers, size_t markers_step, int32_t* n_labels,
                    uint8_t* stages, size_t stage_step);
/* Device-resident form: all pointers are device memory on ctx's device; d_stages (optional) = 4 dense width*height planes. */
MSG_API int msg_shape_seeds_dev(msg_ctx* ctx, const uint8_t* d_bgr, size_t step, int width, int height, int median_ksize,
                        double threshold1, double threshold2, int32_t* d_markers, size_t markers_step, int32_t* d_n_labels,
                        uint8_t* d_stages);

/* ---- colour-method marker generator (PictureService.java:309-366, :938-943, :1018-1023) -- */
/* The Java loop of PictureService.java:309-318: pixels equal to (255,255,255) become (0,0,0). */
MSG_API int msg_white_to_black(msg_ctx* ctx, const uint8_t* src_bgr, size_t src_step, uint8_t* dst_bgr, size_t dst_step,
                       int width, int height);
/* Imgproc.threshold(src 8UC1, dst, thresh, maxval, type) (PictureService.java:941) for type = THRESH_BINARY (0), optionally
 * | THRESH_OTSU (8): dst = src > t ? maxval : 0 with t = floor(thresh), or the Otsu threshold (getThreshVal_Otsu_8u: the
 * sequential double-precision scan over the 256 bins).  *used_thresh (may be NULL) receives t, the value cv::threshold returns. */
#define MSG_THRESH_BINARY 0
#define MSG_THRESH_OTSU 8
MSG_API int msg_threshold(msg_ctx* ctx, const uint8_t* src_gray, size_t src_step, uint8_t* dst, size_t dst_step, int width,
                  int height, double thresh, double maxval, int type, double* used_thresh);
/* Imgproc.distanceTransform(src 8UC1, dst 32FC1, CV_DIST_L2, 5) (PictureService.java:1020): two-pass 5x5 chamfer with the
 * metrics 1 / 1.4 / 2.1969 accumulated in float32 exactly as the IPP-backed cv2 build the oracle is pinned on does it (forward
 * pass: running value unrounded inside aligned groups of four columns).  Only (dist_type 2, mask 5).
 * A source without zero pixels gives FLT_MAX everywhere, as cv2 does.  Rows up to msg_distance_transform_max_width().
 * msg_get_timings().filter_ms holds the kernel time of the last call (CUDA events). */
MSG_API int msg_distance_transform(msg_ctx* ctx, const uint8_t* src, size_t src_step, float* dst, size_t dst_step, int width,
                           int height, int dist_type, int mask_size);
MSG_API int msg_distance_transform_max_width(msg_ctx* ctx);
/* Core.normalize(src 32FC1, dst, alpha, beta, NORM_MINMAX) (PictureService.java:1021): scale / shift in double, applied in
 * float with one fused multiply-add; a constant image maps to min(alpha, beta). */
MSG_API int msg_normalize_minmax(msg_ctx* ctx, const float* src, size_t src_step, float* dst, size_t dst_step, int width,
                         int height, double alpha, double beta);
/* Imgproc.threshold on 32FC1, THRESH_BINARY (PictureService.java:348): dst = src > (float)thresh ? (float)maxval : 0. */
MSG_API int msg_threshold_f32(msg_ctx* ctx, const float* src, size_t src_step, float* dst, size_t dst_step, int width, int height,
                      double thresh, double maxval);
/* Imgproc.dilate(src 32FC1, dst, Mat.ones(kh, kw)) (PictureService.java:349-350). */
MSG_API int msg_dilate_f32(msg_ctx* ctx, const float* src, size_t src_step, float* dst, size_t dst_step, int width, int height,
                   int kw, int kh);
/* Mat.convertTo(dst, CV_8U) from 32FC1 (PictureService.java:355-356): saturate_cast<uchar>(cvRound(v)). */
MSG_API int msg_convert_f32_to_u8(msg_ctx* ctx, const float* src, size_t src_step, uint8_t* dst, size_t dst_step, int width,
                          int height);
/* findContours(mask, contours, hierarchy, RETR_CCOMP, CHAIN_APPROX_NONE) followed by
 * for (i = 0; i < contours.size(); i++) drawContours(markers, contours, i, Scalar.all(i + 1), -1, 8, hierarchy, INT_MAX, Point())
 * (PictureService.java:360-364, also :272-278) in one call, without tracing: markers (32SC1) is written completely (0 where
 * nothing is painted), *n_contours = contours.size() (the reference's `depth`, :365).  Equivalent rule: DESIGN.md section K3. */
MSG_API int msg_contour_markers(msg_ctx* ctx, const uint8_t* mask, size_t mask_step, int32_t* markers, size_t markers_step,
                        int width, int height, int32_t* n_contours);
/* Imgproc.circle(img 32SC1, (cx,cy), radius, Scalar(value), FILLED) (PictureService.java:366), in place. */
MSG_API int msg_circle_filled(msg_ctx* ctx, int32_t* img, size_t step, int width, int height, int cx, int cy, int radius,
                      int32_t value);
/* The whole generator, intermediates in HBM (PictureService.java:309-366): white->black, Laplacian sharpen (taps as in
 * msg_laplacian_sharpen), cvtColor(BGR2GRAY), threshold(40, 255, BINARY | OTSU), distanceTransform(L2, 5), normalize(0, 1,
 * MINMAX), threshold(peak_thresh, 1, BINARY), dilate 3x3, convertTo 8U, findContours / drawContours labelling,
 * circle((5,5), 3, 255, FILLED).  markers: 32SC1; *n_contours = contours.size().  Optional stage outputs (NULL to skip):
 * sharp (8UC3, sharp_step) = "laplassian_sharp", bw (8UC1) = "bw", dist (32FC1, normalised) = "distance_transform",
 * peaks (8UC1, 0/1) = "distance_peaks". */
MSG_API int msg_color_seeds(msg_ctx* ctx, const uint8_t* src_bgr, size_t src_step, int width, int height, const int8_t* taps,
                    int krows, int kcols, double peak_thresh, int32_t* markers, size_t markers_step, int32_t* n_contours,
                    uint8_t* sharp_bgr, size_t sharp_step, uint8_t* bw, size_t bw_step, float* dist, size_t dist_step,
                    uint8_t* peaks, size_t peaks_step);
/* Device-resident form: every pointer is device memory on ctx's device; the optional stage planes are dense (width*3,
 * width, width*4, width bytes per row).  *n_contours is a HOST int (the labelling reads its list sizes back). */
MSG_API int msg_color_seeds_dev(msg_ctx* ctx, const uint8_t* d_bgr, size_t step, int width, int height, const int8_t* taps,
                        int krows, int kcols, double peak_thresh, int32_t* d_markers, size_t markers_step,
                        int32_t* n_contours, uint8_t* d_sharp, uint8_t* d_bw, float* d_dist, uint8_t* d_peaks);

/* Imgproc.bilateralFilter(src, dst, d, sigmaColor, sigmaSpace) on 8UC1 / 8UC3 (PictureService.java:490),
 * BORDER_REFLECT_101: OpenCV's scalar loop (float accumulation in window order).  Floating point: equal to the oracle bit for
 * bit, within 1 LSB of cv2 (whose vector path fuses the multiply-adds).  channels = 1 or 3; window radius <= 32. */
MSG_API int msg_bilateral_filter(msg_ctx* ctx, const uint8_t* src, size_t src_step, uint8_t* dst, size_t dst_step, int width,
                         int height, int channels, int d, double sigma_color, double sigma_space);

/* ---- fused pipeline: filter -> label -> merge -> render, intermediates stay in HBM -------- */
typedef struct msg_segment_params {
    double sp, sr;
    int max_level, term_type, max_count;
    double eps;
    int lo_diff;      /* label stage; < 0 skips labelling (filter only) */
    int min_size;     /* merge stage; 0 with color_dist 0 skips it      */
    int color_dist;
    int render_depth; /* > 0: render with that depth; 0: depth = n_regions; < 0: skip */
    int connectivity; /* label stage: 4 (default, also for 0) or 8 */
    int labels_type;  /* MSG_LABELS_32S (0, CV_32SC1, what the reference's Mat expects) or MSG_LABELS_16U (CV_16UC1: the labels
                         pointer is a uint16_t buffer, 2 bytes per pixel over PCIe instead of 4; more than 65535 regions ->
                         MSG_ERANGE from msg_segment / msg_wait, pixels saturate at 65535) */
} msg_segment_params;
#define MSG_LABELS_32S 0
#define MSG_LABELS_16U 1

MSG_API void msg_segment_params_default(msg_segment_params* p); /* sp=sr=10, L1, (3,5,1), lo=2, no merge */

/* Any output pointer may be NULL: that product is then neither converted nor downloaded (the download mask of the call).
 * labels is int32_t* or, with labels_type = MSG_LABELS_16U, uint16_t* (labels_step in bytes either way). */
MSG_API int msg_segment(msg_ctx* ctx, const uint8_t* src_bgr, size_t src_step, int width, int height,
                const msg_segment_params* params, uint8_t* filtered_bgr, size_t filtered_step,
                void* labels, size_t labels_step, uint8_t* rendered_bgr, size_t rendered_step,
                int32_t* n_regions);

/* ---- asynchronous batch interface (stream ordered within a context) ----------------------- */
/* Buffers must stay valid until msg_wait(ticket); host buffers from msg_alloc_pinned make the
 * copies truly asynchronous.  Up to MSG_MAX_INFLIGHT submissions may be pending per context; consecutive submissions are
 * pipelined (upload, kernels and downloads run on three streams, each in-flight frame owns its device buffers), so the
 * upload of frame i+1 and the download of frame i-1 overlap the kernels of frame i. */
#define MSG_MAX_INFLIGHT 4
MSG_API int msg_submit_segment(msg_ctx* ctx, const uint8_t* src_bgr, size_t src_step, int width, int height,
                       const msg_segment_params* params, uint8_t* filtered_bgr, size_t filtered_step,
                       void* labels, size_t labels_step, uint8_t* rendered_bgr, size_t rendered_step,
                       int* ticket);
MSG_API int msg_wait(msg_ctx* ctx, int ticket, int32_t* n_regions);

MSG_API void* msg_alloc_pinned(size_t bytes);
MSG_API void msg_free_pinned(void* p);

/* ---- device-resident interface (pointers are CUDA device pointers on the context's device;
 *      work is enqueued on the context's stream and NOT synchronised) ------------------------ */
/* msg_segment on device buffers (any output may be NULL); d_n_regions: device int32 or NULL. */
MSG_API int msg_segment_dev(msg_ctx* ctx, const uint8_t* d_src_bgr, size_t src_step, int width, int height,
                    const msg_segment_params* params, uint8_t* d_filtered_bgr, size_t filtered_step,
                    void* d_labels, size_t labels_step, uint8_t* d_rendered_bgr, size_t rendered_step,
                    int32_t* d_n_regions);
MSG_API int msg_meanshift_filter_dev(msg_ctx* ctx, const uint8_t* d_src_bgr, size_t src_step, uint8_t* d_dst_bgr,
                             size_t dst_step, int width, int height, double sp, double sr, int max_level,
                             int term_type, int max_count, double eps);
/* d_n_regions: device int32 (may be NULL). */
MSG_API int msg_label_regions_dev(msg_ctx* ctx, const uint8_t* d_bgr, size_t step, int32_t* d_labels,
                          size_t labels_step, int width, int height, int lo_diff, int connectivity,
                          int32_t* d_n_regions);
MSG_API int msg_connected_components_dev(msg_ctx* ctx, const uint8_t* d_mask, size_t step, int32_t* d_labels,
                                 size_t labels_step, int width, int height, int connectivity,
                                 int32_t* d_n_labels);
/* merge iterates to a fixed point, so it synchronises the stream internally. */
MSG_API int msg_merge_regions_dev(msg_ctx* ctx, const uint8_t* d_bgr, size_t step, int32_t* d_labels,
                          size_t labels_step, int width, int height, int min_size, int color_dist,
                          int32_t* d_n_regions);
MSG_API int msg_render_labels_dev(msg_ctx* ctx, const int32_t* d_labels, size_t labels_step, uint8_t* d_dst_bgr,
                          size_t dst_step, int width, int height, int depth, const uint8_t* d_colors_bgr);
/* Deterministic synthetic test image generated on the device. */
MSG_API int msg_synth_bgr_dev(msg_ctx* ctx, uint8_t* d_dst_bgr, size_t step, int width, int height, uint64_t seed);
/* Rows [row0, row0+rows) of the same width x full_height image (strip-sharded generation). */
MSG_API int msg_synth_bgr_rows_dev(msg_ctx* ctx, uint8_t* d_dst_bgr, size_t step, int width, int full_height, int row0,
                           int rows, uint64_t seed);

/* ---- strip sharding support (one very large image over several GPUs; DESIGN.md "multi-GPU") */
/* Mean-shift filter of rows [row0,row1) of a width x full_height image, given device rows
 * [halo_row0, halo_row1) of the source (halo_row0 <= row0 < row1 <= halo_row1).  Coordinates are
 * global, so the result is bit-identical to the same rows of the unsharded call provided the halo
 * is wide enough: msg_meanshift_halo_rows().  row0 and halo_row0 must be multiples of
 * 2^max_level.  Output: rows [row0,row1) at d_dst_bgr (row 0 of the buffer = row0). */
MSG_API int msg_meanshift_halo_rows(double sp, int max_level, int term_type, int max_count);
MSG_API int msg_meanshift_filter_strip_dev(msg_ctx* ctx, const uint8_t* d_src_rows, size_t src_step,
                                   int halo_row0, int halo_row1, uint8_t* d_dst_bgr, size_t dst_step,
                                   int width, int full_height, int row0, int row1, double sp, double sr,
                                   int max_level, int term_type, int max_count, double eps);
/* Labels of a strip with GLOBAL provisional labels (1 + global linear index of the component's
 * first pixel inside the strip); seams are resolved by msg_seam_pairs_dev + msg_apply_label_map_dev. */
MSG_API int msg_label_strip_dev(msg_ctx* ctx, const uint8_t* d_bgr_rows, size_t step, int32_t* d_labels,
                        size_t labels_step, int width, int rows, int row0, int full_width, int lo_diff);
/* Equivalence pairs across one seam: upper strip's last row vs lower strip's first row.
 * Writes up to `width` (labelA,labelB) pairs to d_pairs (2*width int32), count to d_count. */
MSG_API int msg_seam_pairs_dev(msg_ctx* ctx, const uint8_t* d_upper_row_bgr, const int32_t* d_upper_row_labels,
                       const uint8_t* d_lower_row_bgr, const int32_t* d_lower_row_labels, int width,
                       int lo_diff, int32_t* d_pairs, int32_t* d_count);
/* labels[p] = map(labels[p]) where map is given as sorted (from,to) pairs (binary search). */
MSG_API int msg_apply_label_map_dev(msg_ctx* ctx, int32_t* d_labels, size_t labels_step, int width, int rows,
                            const int32_t* d_from_sorted, const int32_t* d_to, int n_map);

/* Dense global numbering of strip labels (after msg_apply_label_map_dev): regions are renumbered 1..N over the WHOLE image in
 * raster order of their first pixel, exactly as the unsharded call numbers them.  Three steps per strip (the context keeps
 * the ranks between them): (1) count the roots this strip owns -> *d_count; the caller exclusive-scans the counts over the
 * strips (an all-gather of one int per rank) to get `offset`; (2) dense ids of owned roots named in d_query (the labels other
 * strips were mapped onto; 0 for labels not owned) -> d_out, which the ranks all-gather into a sorted (label, dense) table;
 * (3) rewrite the strip: owned roots through the ranks, foreign roots through the table (binary search). */
MSG_API int msg_strip_rank_dev(msg_ctx* ctx, const int32_t* d_labels, size_t labels_step, int width, int rows, int row0,
                       int full_width, int32_t* d_count);
MSG_API int msg_strip_query_dense_dev(msg_ctx* ctx, const int32_t* d_query, int n_query, int width, int rows, int row0,
                              int full_width, int offset, int32_t* d_out);
MSG_API int msg_strip_apply_dense_dev(msg_ctx* ctx, int32_t* d_labels, size_t labels_step, int width, int rows, int row0,
                              int full_width, int offset, const int32_t* d_remote_labels_sorted,
                              const int32_t* d_remote_dense, int n_remote);

/* Single-exchange form of the seam resolution + dense numbering (one boundary-row transfer and ONE all-gather per image):
 *  (1) msg_strip_rank_dev on the PROVISIONAL strip labels (right after msg_label_strip_dev): *d_count = roots of the strip;
 *  (2) the rank above sends its last row of colours, labels and msg_strip_query_dense_dev(labels of that row, offset 0) =
 *      rank + 1 of every label's root; msg_seam_quads_dev emits one quad (A, B, rankA + 1, rankB + 1) per equivalence
 *      (up to `width` quads, count to d_count); B's rank comes from this context's tables of step (1);
 *  (3) the ranks all-gather (quad count, root count, quads); every rank resolves the equivalences (union-find over the quads)
 *      and derives, identically, the sorted list d_frm of labels that merge into a smaller one, the dense id d_dense[j] of the
 *      class frm[j] joins, this strip's offset (surviving roots of the strips above) and frm_lo (index of the first entry of
 *      d_frm that lies in this strip) -- opencv-msegment_b200/sharded.py:resolve_dense;
 *  (4) msg_strip_finalize_dense_dev rewrites the strip in one pass.  Result: identical to the unsharded call's labels. */
MSG_API int msg_seam_quads_dev(msg_ctx* ctx, const uint8_t* d_upper_row_bgr, const int32_t* d_upper_row_labels,
                       const int32_t* d_upper_row_rank1, const uint8_t* d_lower_row_bgr, const int32_t* d_lower_row_labels,
                       int width, int lo_diff, int rows, int row0, int full_width, int32_t* d_quads, int32_t* d_count);
MSG_API int msg_strip_finalize_dense_dev(msg_ctx* ctx, int32_t* d_labels, size_t labels_step, int width, int rows, int row0,
                                 int full_width, int offset, const int32_t* d_frm_sorted, const int32_t* d_dense, int n_map,
                                 int frm_lo);

/* ---- strip sharding as a library feature (round 2): plan, device-side seam resolution, sharded merge ------------------------
 * The collectives themselves stay with the host (NCCL: ncclSend / ncclRecv of halo rows and one boundary row per seam, ONE
 * ncclAllGather of the seam payloads, ncclAllReduce + ncclAllGather for the merge tables -- "NCCL only for that exchange",
 * BASELINE.json north_star); everything else is behind this ABI, so a Java / C host drives the same sequence the Python tool
 * tools/shard_large_image.py drives (INTEGRATION.md section 6 lists the calls in order). */
#define MSG_MAX_STRIPS 64
#define MSG_SHARD_TABLE_HEADER 192
typedef struct msg_shard_plan {
    int n_strips, width, height, max_level, halo_rows;
    int row0[MSG_MAX_STRIPS], row1[MSG_MAX_STRIPS];      /* rows [row0, row1) a rank filters and labels */
    int halo0[MSG_MAX_STRIPS], halo1[MSG_MAX_STRIPS];    /* source rows [halo0, halo1) it must hold to do so */
} msg_shard_plan;
/* Pure host function: contiguous strips whose first rows are multiples of 2^max_level (pyramid phase) + their halo ranges
 * (msg_meanshift_halo_rows).  MSG_EINVAL if the image is too small for n_strips. */
MSG_API int msg_shard_plan_make(int width, int height, int n_strips, double sp, int max_level, int term_type, int max_count,
                        msg_shard_plan* out);
/* Device-side form of step (3) of the single-exchange seam resolution above (no host round trip, no host union-find):
 * d_gathered = the all-gathered payloads, n_strips x (width + 1) x 4 int32, row 0 of every payload = (quad count, root count,
 * 0, 0), rows 1.. = the quads of msg_seam_quads_dev; row0[s] = first row of strip s.  d_tables (int32, at least
 * MSG_SHARD_TABLE_HEADER + 2 * n_strips * width entries) receives: [0] n_map, [1] total number of regions of the image,
 * [2 + s] offset of strip s, [2 + MSG_MAX_STRIPS + s] frm_lo of strip s, then frm[n_strips * width], dense[n_strips * width]. */
MSG_API int msg_strip_resolve_dense_dev(msg_ctx* ctx, const int32_t* d_gathered, int n_strips, int width, const int* row0,
                                int32_t* d_tables, size_t tables_capacity_ints);
/* Step (4) reading everything from d_tables: rewrites the strip with the dense global ids (identical to the unsharded call). */
MSG_API int msg_strip_finalize_tables_dev(msg_ctx* ctx, int32_t* d_labels, size_t labels_step, int width, int rows, int row0,
                                  int full_width, int strip, int n_strips, const int32_t* d_tables);
/* Region merge across strips: labels are the dense global ids 1..n_total.
 *  (1) every rank: msg_strip_merge_stats_dev accumulates the area / colour sums of ITS strip into the caller-owned tables
 *      d_area[n_total + 1] (uint32) and d_sum[3 * (n_total + 1)] (uint64, B G R per label; both are zeroed by the call) and
 *      appends the adjacent (label, label) pairs of the strip -- including the seam with the strip above when
 *      d_up_row_labels (the dense labels of that strip's last row) is given -- to d_pairs (2 int32 per pair, at most pair_cap
 *      pairs; *d_npairs = pairs found, the call fails later if it exceeded pair_cap);
 *  (2) host: all-reduce(sum) d_area and d_sum over the ranks, all-gather the pair lists;
 *  (3) every rank: msg_strip_merge_finish_dev runs the merge rounds on the (identical) tables -- identical decisions on every
 *      rank -- and rewrites its strip; *d_n_out = regions after the merge.  d_area / d_sum are modified.
 * The result equals msg_merge_regions_dev on the whole image bit for bit. */
MSG_API int msg_strip_merge_stats_dev(msg_ctx* ctx, const uint8_t* d_bgr_rows, size_t step, const int32_t* d_labels, size_t labels_step,
                              int width, int rows, const int32_t* d_up_row_labels, int n_total, uint32_t* d_area,
                              unsigned long long* d_sum, int32_t* d_pairs, long long pair_cap, int32_t* d_npairs);
MSG_API int msg_strip_merge_finish_dev(msg_ctx* ctx, int32_t* d_labels, size_t labels_step, int width, int rows, long long full_pixels,
                               int n_total, uint32_t* d_area, unsigned long long* d_sum, const int32_t* d_all_pairs,
                               long long n_all_pairs, int min_size, int color_dist, int32_t* d_n_out);

/* ---- introspection ------------------------------------------------------------------------- */
typedef struct msg_timings { /* milliseconds of the last host-buffer call, CUDA events */
    float h2d_ms, filter_ms, label_ms, merge_ms, render_ms, d2h_ms, total_ms;
} msg_timings;
MSG_API int msg_get_timings(msg_ctx* ctx, msg_timings* out);

typedef struct msg_stats {
    uint64_t kernel_launches;     /* kernels of this library launched by this context so far */
    uint64_t ms_overflow_items;   /* mean-shift items that left their staged tile (slow path), last call */
    uint64_t ms_active_items;     /* mean-shift pixels processed (all levels), last call */
    uint64_t merge_rounds;        /* rounds of the last merge call */
    uint64_t h2d_bytes, d2h_bytes; /* bytes copied by host-buffer calls so far */
    uint64_t staged_bytes;        /* of those, bytes that went through pinned staging because the caller's buffer was pageable */
} msg_stats;
MSG_API int msg_get_stats(msg_ctx* ctx, msg_stats* out);

/* Per-kernel profile of the mean-shift levels (bench.py's roofline): CUDA-event durations of the tile kernel
 * and of the overflow kernel, and the ALGORITHMIC work they executed, counted on the device exactly as the CPU
 * oracle counts it (window tests T = clamped window area per iteration, hits = in-range tests).  Profiling
 * serialises the host with every level; leave it off when timing throughput. */
typedef struct msg_kernel_profile {
    uint64_t launches[9];       /* tile-kernel launches per pyramid level since profiling was enabled */
    double tile_ms[9];          /* summed tile-kernel durations per level */
    double overflow_ms[9];      /* summed overflow (generic) kernel durations per level */
    uint64_t tile_tests[9], tile_hits[9];          /* work done by the tile kernel */
    uint64_t overflow_tests[9], overflow_hits[9];  /* work done by the generic kernel */
} msg_kernel_profile;
MSG_API int msg_set_profiling(msg_ctx* ctx, int enable);      /* enabling resets the accumulated profile */
MSG_API int msg_get_kernel_profile(msg_ctx* ctx, msg_kernel_profile* out);

/* Debug: copy pyramid plane (kind 0 = source S[level], 1 = result D[level]) of the last filter call
 * to a host BGRX buffer (4 bytes/pixel, dense).  Writes its size to *w,*h. */
MSG_API int msg_debug_get_plane(msg_ctx* ctx, int kind, int level, uint32_t* host_out, size_t capacity_pixels,
                        int* w, int* h);

#ifdef __cplusplus
}
#endif
#endif /* MSEGMENT_H */
