// GpuImgproc.hpp -- C++ host-side mirror of the reference's operator interface for the segmentation hot path.
//
// The reference is Java: static methods of org.opencv.imgproc.Imgproc on org.opencv.core.Mat
// (PictureService.java:441-442 connectedComponents, :913-936 colorByIndexes; pyrMeanShiftFiltering / floodFill-style
// labelling / merge per BASELINE.json north_star).  The build image has no JDK, so this header gives the same static
// signatures in C++ over a minimal Mat (rows, cols, type, byte step, data) and forwards to the C ABI
// (include/msegment.h).  Non-zero status -> CvException, exactly where the Java shim (java/GpuImgproc.java) throws
// org.opencv.core.CvException.  Header only; no CPU implementation of any operator.
#pragma once
#include <cstdint>
#include <cstring>
#include <stdexcept>
#include <string>
#include <vector>

#include "msegment.h"

namespace msegment {

enum MatType { CV_8UC1 = 0, CV_8UC3 = 16, CV_32SC1 = 4, CV_32FC1 = 5 };   // OpenCV's CvType codes

struct CvException : std::runtime_error {
    int status;
    CvException(int st, const std::string& msg) : std::runtime_error("msegment status " + std::to_string(st) + ": " + msg), status(st) {}
};

struct Mat {   // row-major, continuous (like every Mat the reference creates: imread / clone / zeros)
    int rows = 0, cols = 0, type = CV_8UC3;
    std::vector<uint8_t> buf;
    Mat() = default;
    Mat(int r, int c, int t) { create(r, c, t); }
    static int elemSize(int t) { return t == CV_8UC3 ? 3 : (t == CV_32SC1 || t == CV_32FC1 ? 4 : 1); }
    void create(int r, int c, int t) { rows = r; cols = c; type = t; buf.assign((size_t)r * c * elemSize(t), 0); }
    size_t step() const { return (size_t)cols * elemSize(type); }
    uint8_t* data() { return buf.data(); }
    const uint8_t* data() const { return buf.data(); }
    bool empty() const { return buf.empty(); }
};

struct TermCriteria {
    enum { COUNT = 1, EPS = 2 };
    int type = COUNT + EPS, maxCount = 5;
    double epsilon = 1.0;
    TermCriteria() = default;
    TermCriteria(int t, int c, double e) : type(t), maxCount(c), epsilon(e) {}
};

class GpuImgproc {
public:
    // one context per thread, like the Java shim's ThreadLocal
    static msg_ctx* ctx()
    {
        thread_local struct Holder {
            msg_ctx* c = nullptr;
            ~Holder() { if (c) msg_destroy(c); }
        } h;
        if (!h.c) {
            int rc = msg_create(0, &h.c);
            if (rc != MSG_OK) throw CvException(rc, msg_last_error(nullptr));
        }
        return h.c;
    }

    static void pyrMeanShiftFiltering(const Mat& src, Mat& dst, double sp, double sr, int maxLevel = 1,
                                      TermCriteria tc = TermCriteria())
    {
        require(src.type == CV_8UC3, "src must be CV_8UC3");
        dst.create(src.rows, src.cols, CV_8UC3);
        check(msg_meanshift_filter(ctx(), src.data(), src.step(), dst.data(), dst.step(), src.cols, src.rows, sp, sr, maxLevel,
                                   tc.type, tc.maxCount, tc.epsilon));
    }

    static int labelRegions(const Mat& image, Mat& labels, int loDiff = 2, int upDiff = 2, int connectivity = 4)
    {
        require(image.type == CV_8UC3, "image must be CV_8UC3");
        labels.create(image.rows, image.cols, CV_32SC1);
        int32_t n = 0;
        check(msg_label_regions(ctx(), image.data(), image.step(), (int32_t*)labels.data(), labels.step(), image.cols, image.rows,
                                loDiff, upDiff, connectivity, &n));
        return n;
    }

    static int mergeRegions(const Mat& image, Mat& labels, int minSize, int colorDist)
    {
        require(image.type == CV_8UC3 && labels.type == CV_32SC1 && labels.rows == image.rows && labels.cols == image.cols,
                "image CV_8UC3 and labels CV_32SC1 of equal size required");
        int32_t n = 0;
        check(msg_merge_regions(ctx(), image.data(), image.step(), (int32_t*)labels.data(), labels.step(), image.cols, image.rows,
                                minSize, colorDist, &n));
        return n;
    }

    // Imgproc.connectedComponents(image, labels, connectivity, ltype)  -- PictureService.java:441-442
    static int connectedComponents(const Mat& image, Mat& labels, int connectivity = 8, int ltype = CV_32SC1)
    {
        require(image.type == CV_8UC1 && ltype == CV_32SC1, "CV_8UC1 image and CV_32S labels only");
        labels.create(image.rows, image.cols, CV_32SC1);
        int32_t n = 0;
        check(msg_connected_components(ctx(), image.data(), image.step(), (int32_t*)labels.data(), labels.step(), image.cols,
                                       image.rows, connectivity, &n));
        return n;
    }

    // Imgproc.watershed(image, markers) -- PictureService.java:909; markers CV_32SC1, modified in place
    static void watershed(const Mat& image, Mat& markers)
    {
        require(image.type == CV_8UC3 && markers.type == CV_32SC1, "watershed: image CV_8UC3, markers CV_32SC1");
        require(image.rows == markers.rows && image.cols == markers.cols, "watershed: markers must have the size of image");
        check(msg_watershed(ctx(), image.data(), image.step(), (int32_t*)markers.data(), markers.step(), image.cols, image.rows));
    }

    // PictureService.watershed(src, markers, depth, colored) -- PictureService.java:908-911
    static Mat watershedAndColor(const Mat& src, Mat& markers, int depth, const uint8_t* colorsBgr = nullptr)
    {
        watershed(src, markers);
        return colorByIndexes(markers, depth, colorsBgr);
    }

    // src.copyTo(dst, mask) on a zero dst -- PictureService.java:417-418
    static void copyToMasked(const Mat& src, const Mat& mask, Mat& dst)
    {
        require(src.type == CV_8UC3 && mask.type == CV_8UC1, "copyTo: src CV_8UC3, mask CV_8UC1");
        dst.create(src.rows, src.cols, CV_8UC3);
        check(msg_copy_masked(ctx(), src.data(), src.step(), mask.data(), mask.step(), dst.data(), dst.step(), src.cols, src.rows));
    }

    // PictureService.colorByIndexes(markers, depth, colored) -- PictureService.java:913-936 (colors == nullptr: white)
    static Mat colorByIndexes(const Mat& markers, int depth, const uint8_t* colorsBgr = nullptr)
    {
        require(markers.type == CV_32SC1, "markers must be CV_32SC1");
        Mat dst(markers.rows, markers.cols, CV_8UC3);
        check(msg_render_labels(ctx(), (const int32_t*)markers.data(), markers.step(), dst.data(), dst.step(), markers.cols,
                                markers.rows, depth, colorsBgr));
        return dst;
    }

    // ---- pre-filters the reference calls around the segmentation stage
    // filter2D + convertTo + subtract + convertTo chain of PictureService.java:323-333; kernel = integer taps (krows x kcols)
    static void sharpenLaplacian(const Mat& src, Mat& dst, const int8_t* taps, int krows, int kcols)
    {
        require(src.type == CV_8UC3, "src must be CV_8UC3");
        dst.create(src.rows, src.cols, CV_8UC3);
        check(msg_laplacian_sharpen(ctx(), src.data(), src.step(), dst.data(), dst.step(), src.cols, src.rows, taps, krows, kcols));
    }

    // Imgproc.cvtColor(src, dst, COLOR_BGR2GRAY)
    static void cvtColorBGR2GRAY(const Mat& src, Mat& dst)
    {
        require(src.type == CV_8UC3, "src must be CV_8UC3");
        dst.create(src.rows, src.cols, CV_8UC1);
        check(msg_bgr2gray(ctx(), src.data(), src.step(), dst.data(), dst.step(), src.cols, src.rows));
    }

    // Imgproc.medianBlur(src, dst, ksize) on CV_8UC1
    static void medianBlur(const Mat& src, Mat& dst, int ksize)
    {
        require(src.type == CV_8UC1, "src must be CV_8UC1");
        Mat out(src.rows, src.cols, CV_8UC1);
        check(msg_median_blur(ctx(), src.data(), src.step(), out.data(), out.step(), src.cols, src.rows, ksize));
        dst = out;
    }

    // ---- shape-method marker generator (PictureService.java:404-442)
    static void Canny(const Mat& image, Mat& edges, double threshold1, double threshold2)
    {
        require(image.type == CV_8UC1, "image must be CV_8UC1");
        Mat out(image.rows, image.cols, CV_8UC1);
        check(msg_canny(ctx(), image.data(), image.step(), out.data(), out.step(), image.cols, image.rows, threshold1, threshold2));
        edges = out;
    }

    // Imgproc.dilate(src, dst, Mat.ones(krows, kcols, type))
    static void dilate(const Mat& src, Mat& dst, int krows, int kcols)
    {
        require(src.type == CV_8UC1, "src must be CV_8UC1");
        Mat out(src.rows, src.cols, CV_8UC1);
        check(msg_dilate(ctx(), src.data(), src.step(), out.data(), out.step(), src.cols, src.rows, kcols, krows));
        dst = out;
    }

    // Core.subtract(src1, src2, dst)
    static void subtract(const Mat& a, const Mat& b, Mat& dst)
    {
        require(a.type == CV_8UC1 && b.type == CV_8UC1 && a.rows == b.rows && a.cols == b.cols, "subtract: CV_8UC1 of equal size");
        Mat out(a.rows, a.cols, CV_8UC1);
        check(msg_subtract(ctx(), a.data(), a.step(), b.data(), b.step(), out.data(), out.step(), a.cols, a.rows));
        dst = out;
    }

    // PictureService.calculateSizeOfSquareBlurMask (PictureService.java:877-899)
    static int calculateSizeOfSquareBlurMask(int cols, int rows)
    {
        int m = cols <= rows ? cols : rows;
        if (m < 3) return 1;
        if (m <= 100) return 5;
        double scale = m <= 360 ? 0.025 : m <= 480 ? 0.02 : m <= 720 ? 0.015 : m <= 1080 ? 0.01 : 0.005;
        int r = (int)(m * scale);
        return r % 2 == 0 ? r + 1 : r;
    }

    // marker half of shapeAutoMarkerWatershed (:404-442); returns the label count (background included)
    static int shapeSeeds(const Mat& src, Mat& markers, double lowThreshold = 5, double ratio = 10)
    {
        require(src.type == CV_8UC3, "src must be CV_8UC3");
        markers.create(src.rows, src.cols, CV_32SC1);
        int32_t n = 0;
        check(msg_shape_seeds(ctx(), src.data(), src.step(), src.cols, src.rows, calculateSizeOfSquareBlurMask(src.cols, src.rows),
                              lowThreshold, lowThreshold * ratio, (int32_t*)markers.data(), markers.step(), &n, nullptr, 0));
        return n;
    }

    // ---- colour-method marker generator (PictureService.java:309-366, :938-943, :1018-1023)
    enum { THRESH_BINARY = 0, THRESH_OTSU = 8, CV_DIST_L2 = 2, NORM_MINMAX = 32 };

    // the Java loop of PictureService.java:309-318: (255,255,255) -> (0,0,0)
    static void whiteToBlack(const Mat& src, Mat& dst)
    {
        require(src.type == CV_8UC3, "src must be CV_8UC3");
        Mat out(src.rows, src.cols, CV_8UC3);
        check(msg_white_to_black(ctx(), src.data(), src.step(), out.data(), out.step(), src.cols, src.rows));
        dst = out;
    }

    // Imgproc.threshold(src, dst, thresh, maxval, type): CV_8UC1 (BINARY [| OTSU]) or CV_32FC1 (BINARY); returns the threshold used
    static double threshold(const Mat& src, Mat& dst, double thresh, double maxval, int type)
    {
        require(src.type == CV_8UC1 || src.type == CV_32FC1, "src must be CV_8UC1 or CV_32FC1");
        Mat out(src.rows, src.cols, src.type);
        double used = thresh;
        if (src.type == CV_8UC1)
            check(msg_threshold(ctx(), src.data(), src.step(), out.data(), out.step(), src.cols, src.rows, thresh, maxval, type, &used));
        else {
            require(type == THRESH_BINARY, "CV_32FC1: THRESH_BINARY only");
            check(msg_threshold_f32(ctx(), (const float*)src.data(), src.step(), (float*)out.data(), out.step(), src.cols, src.rows,
                                    thresh, maxval));
        }
        dst = out;
        return used;
    }

    // Imgproc.distanceTransform(src, dst, Imgproc.CV_DIST_L2, 5)
    static void distanceTransform(const Mat& src, Mat& dst, int distanceType, int maskSize)
    {
        require(src.type == CV_8UC1, "src must be CV_8UC1");
        Mat out(src.rows, src.cols, CV_32FC1);
        check(msg_distance_transform(ctx(), src.data(), src.step(), (float*)out.data(), out.step(), src.cols, src.rows, distanceType,
                                     maskSize));
        dst = out;
    }

    // Core.normalize(src, dst, alpha, beta, Core.NORM_MINMAX)
    static void normalize(const Mat& src, Mat& dst, double alpha, double beta, int normType)
    {
        require(src.type == CV_32FC1 && normType == NORM_MINMAX, "normalize: CV_32FC1, NORM_MINMAX");
        Mat out(src.rows, src.cols, CV_32FC1);
        check(msg_normalize_minmax(ctx(), (const float*)src.data(), src.step(), (float*)out.data(), out.step(), src.cols, src.rows,
                                   alpha, beta));
        dst = out;
    }

    // Imgproc.dilate on CV_32FC1 with Mat.ones(krows, kcols) (PictureService.java:349-350)
    static void dilateF32(const Mat& src, Mat& dst, int krows, int kcols)
    {
        require(src.type == CV_32FC1, "src must be CV_32FC1");
        Mat out(src.rows, src.cols, CV_32FC1);
        check(msg_dilate_f32(ctx(), (const float*)src.data(), src.step(), (float*)out.data(), out.step(), src.cols, src.rows, kcols,
                             krows));
        dst = out;
    }

    // Mat.convertTo(dst, CvType.CV_8U) from CV_32FC1
    static void convertToU8(const Mat& src, Mat& dst)
    {
        require(src.type == CV_32FC1, "src must be CV_32FC1");
        Mat out(src.rows, src.cols, CV_8UC1);
        check(msg_convert_f32_to_u8(ctx(), (const float*)src.data(), src.step(), out.data(), out.step(), src.cols, src.rows));
        dst = out;
    }

    // findContours(RETR_CCOMP, CHAIN_APPROX_NONE) + the drawContours loop of PictureService.java:360-364; returns contours.size()
    static int contourMarkers(const Mat& image, Mat& markers)
    {
        require(image.type == CV_8UC1, "image must be CV_8UC1");
        markers.create(image.rows, image.cols, CV_32SC1);
        int32_t n = 0;
        check(msg_contour_markers(ctx(), image.data(), image.step(), (int32_t*)markers.data(), markers.step(), image.cols,
                                  image.rows, &n));
        return n;
    }

    // Imgproc.circle(img, center, radius, color, -1) on CV_32SC1
    static void circle(Mat& img, int cx, int cy, int radius, int value)
    {
        require(img.type == CV_32SC1, "img must be CV_32SC1");
        check(msg_circle_filled(ctx(), (int32_t*)img.data(), img.step(), img.cols, img.rows, cx, cy, radius, value));
    }

    // marker half of colorAutoMarkerWatershed (:309-366) with the reference's literal 9 x 1 sharpen kernel; returns contours.size()
    static int colorSeeds(const Mat& src, Mat& markers, double peakThresh = 0.4)
    {
        require(src.type == CV_8UC3, "src must be CV_8UC3");
        static const int8_t taps[9] = {1, 1, 1, 1, -8, 1, 1, 1, 1};
        markers.create(src.rows, src.cols, CV_32SC1);
        int32_t n = 0;
        check(msg_color_seeds(ctx(), src.data(), src.step(), src.cols, src.rows, taps, 9, 1, peakThresh, (int32_t*)markers.data(),
                              markers.step(), &n, nullptr, 0, nullptr, 0, nullptr, 0, nullptr, 0));
        return n;
    }

    // Imgproc.bilateralFilter(src, dst, d, sigmaColor, sigmaSpace) (PictureService.java:490)
    static void bilateralFilter(const Mat& src, Mat& dst, int d, double sigmaColor, double sigmaSpace)
    {
        require(src.type == CV_8UC1 || src.type == CV_8UC3, "src must be CV_8UC1 or CV_8UC3");
        Mat out(src.rows, src.cols, src.type);
        check(msg_bilateral_filter(ctx(), src.data(), src.step(), out.data(), out.step(), src.cols, src.rows,
                                   src.type == CV_8UC3 ? 3 : 1, d, sigmaColor, sigmaSpace));
        dst = out;
    }

private:
    static void require(bool ok, const char* msg) { if (!ok) throw CvException(MSG_EINVAL, msg); }
    static void check(int rc) { if (rc != MSG_OK) throw CvException(rc, msg_last_error(ctx())); }
};

}  // namespace msegment
