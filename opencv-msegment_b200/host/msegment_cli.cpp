// msegment_cli.cpp -- the reference's console entry point (App.java:14-31) over the B200 path.
//
// Contract kept from App.main: exactly three positional arguments (input folder, output root, input file name);
// any other count prints "error parsing args" and returns normally; the arguments are echoed as "arg i: v".
// The reference then runs its two watershed pipelines and (at HEAD) writes nothing; its file writer
// (PictureService.saveResultsToFS, PictureService.java:194-234) names results
//   <imageDir>/<name>_output/<yyyyMMdd'T'HHmmss>/<SegMethod>_<name>_<step %05d>_<stepName>.png
// This program runs the mean-shift segmentation pass named by BASELINE.json (filter -> label -> merge -> render) and writes
// that batch with the same naming scheme and file type (.png, written by the small codec in png_io.hpp; MSG_OUT_FORMAT=pnm
// switches to binary P6/P5).  Inputs: PNG (non-interlaced, 8 bit or palette), baseline JPEG (jpeg_io.hpp: the reference's own
// sample images are baseline 4:2:0 files; pixels equal imread's) or binary PPM.  Optional environment: MSG_SP, MSG_SR, MSG_MIN_SIZE,
// MSG_COLOR_DIST, MSG_OUT_FORMAT.
#include <sys/stat.h>

#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <ctime>
#include <fstream>
#include <iostream>
#include <sstream>

#include "GpuImgproc.hpp"
#include "jpeg_io.hpp"
#include "png_io.hpp"

using namespace msegment;

static bool read_ppm(const std::string& path, Mat& img)
{
    std::ifstream f(path, std::ios::binary);
    if (!f) return false;
    std::string magic;
    f >> magic;
    if (magic != "P6") return false;
    auto next_int = [&](int& v) {
        for (;;) {
            f >> std::ws;
            if (f.peek() == '#') { std::string line; std::getline(f, line); continue; }
            break;
        }
        f >> v;
    };
    int w = 0, h = 0, maxv = 0;
    next_int(w); next_int(h); next_int(maxv);
    f.get();
    if (!f || w <= 0 || h <= 0 || maxv != 255) return false;
    std::vector<uint8_t> rgb((size_t)w * h * 3);
    f.read((char*)rgb.data(), (std::streamsize)rgb.size());
    if (!f) return false;
    img.create(h, w, CV_8UC3);                      // OpenCV order: BGR
    for (size_t i = 0; i < (size_t)w * h; i++) {
        img.buf[3 * i] = rgb[3 * i + 2]; img.buf[3 * i + 1] = rgb[3 * i + 1]; img.buf[3 * i + 2] = rgb[3 * i];
    }
    return true;
}

static bool read_image(const std::string& path, Mat& img)
{
    if (read_ppm(path, img)) return true;
    std::vector<uint8_t> bgr;
    int w = 0, h = 0;
    if (!png::read_bgr(path, bgr, w, h) && !jpeg::read_bgr(path, bgr, w, h)) return false;
    img.create(h, w, CV_8UC3);
    img.buf = bgr;
    return true;
}

// Core.multiply(m, multiplier) then the 8-bit saturation imwrite applies to CV_32S (PictureService.java:209-216)
static std::vector<uint8_t> labels_to_gray(const Mat& m, int multiplier)
{
    const int32_t* p = (const int32_t*)m.data();
    std::vector<uint8_t> g((size_t)m.rows * m.cols);
    for (size_t i = 0; i < g.size(); i++) {
        long long v = (long long)p[i] * multiplier;
        g[i] = (uint8_t)(v < 0 ? 0 : (v > 255 ? 255 : v));
    }
    return g;
}

// Core.multiply(m, Scalar(multiplier)) followed by imwrite's conversion to 8 bit (saturate_cast<uchar>(cvRound(v)))
static std::vector<uint8_t> float_to_gray(const Mat& m, int multiplier)
{
    const float* p = (const float*)m.data();
    std::vector<uint8_t> g((size_t)m.rows * m.cols);
    for (size_t i = 0; i < g.size(); i++) {
        double v = std::nearbyint((double)p[i] * multiplier);
        g[i] = (uint8_t)(v < 0 ? 0 : (v > 255 ? 255 : v));
    }
    return g;
}

static void write_png(const std::string& path, const Mat& m, int multiplier)
{
    bool ok;
    if (m.type == CV_8UC3) {
        std::vector<uint8_t> rgb(m.buf.size());
        for (size_t i = 0; i < (size_t)m.rows * m.cols; i++) {
            rgb[3 * i] = m.buf[3 * i + 2]; rgb[3 * i + 1] = m.buf[3 * i + 1]; rgb[3 * i + 2] = m.buf[3 * i];
        }
        ok = png::write(path, rgb.data(), (size_t)m.cols * 3, m.cols, m.rows, 3);
    } else if (m.type == CV_32SC1 || m.type == CV_32FC1) {
        std::vector<uint8_t> g = m.type == CV_32SC1 ? labels_to_gray(m, multiplier) : float_to_gray(m, multiplier);
        ok = png::write(path, g.data(), (size_t)m.cols, m.cols, m.rows, 1);
    } else {
        ok = png::write(path, m.data(), (size_t)m.cols, m.cols, m.rows, 1);
    }
    if (!ok) std::cerr << "cannot write " << path << std::endl;
}

static void write_pnm(const std::string& path, const Mat& m, int multiplier)
{
    std::ofstream f(path, std::ios::binary);
    if (m.type == CV_8UC3) {
        f << "P6\n" << m.cols << " " << m.rows << "\n255\n";
        std::vector<uint8_t> rgb(m.buf.size());
        for (size_t i = 0; i < (size_t)m.rows * m.cols; i++) {
            rgb[3 * i] = m.buf[3 * i + 2]; rgb[3 * i + 1] = m.buf[3 * i + 1]; rgb[3 * i + 2] = m.buf[3 * i];
        }
        f.write((const char*)rgb.data(), (std::streamsize)rgb.size());
    } else if (m.type == CV_32SC1 || m.type == CV_32FC1) {   // Core.multiply(m, multiplier) then saturate to 8 bit, as imwrite would
        f << "P5\n" << m.cols << " " << m.rows << "\n255\n";
        std::vector<uint8_t> g = m.type == CV_32SC1 ? labels_to_gray(m, multiplier) : float_to_gray(m, multiplier);
        f.write((const char*)g.data(), (std::streamsize)g.size());
    } else {
        f << "P5\n" << m.cols << " " << m.rows << "\n255\n";
        f.write((const char*)m.data(), (std::streamsize)m.buf.size());
    }
}

static double env_or(const char* k, double d) { const char* v = getenv(k); return v ? atof(v) : d; }

int main(int argc, char** argv)
{
    if (argc - 1 != 3) {                            // App.java:17-20
        std::cout << "error parsing args" << std::endl;
        return 0;
    }
    for (int i = 1; i < argc; i++) std::cout << "arg " << (i - 1) << ": " << argv[i] << std::endl;   // App.java:22-24
    const std::string dir = argv[1], out_root = argv[2], file = argv[3];
    (void)out_root;                                 // dead in the reference too (PictureService.java:118, :130)
    Mat src;
    if (!read_image(dir + "/" + file, src)) {         // readPicture: dataAddr()==0 -> IOException -> logged, pipeline returns null
        std::cerr << "There is an error with file stream processing: cannot read PNG / baseline JPEG / binary PPM " << dir << "/" << file << std::endl;
        return 0;
    }
    const std::string name = file.substr(0, file.find('.'));   // ImageInfo: text before the first '.'
    char stamp[32];
    std::time_t t = std::time(nullptr);
    std::strftime(stamp, sizeof(stamp), "%Y%m%dT%H%M%S", std::localtime(&t));
    const std::string odir1 = dir + "/" + name + "_output", odir = odir1 + "/" + stamp;
    mkdir(odir1.c_str(), 0755);
    mkdir(odir.c_str(), 0755);
    const char* fmt = getenv("MSG_OUT_FORMAT");
    const bool pnm = fmt && std::string(fmt) == "pnm";
    std::string method = "MEANSHIFT_METHOD";
    auto save = [&](int step, const std::string& step_name, const Mat& m, int multiplier) {
        char b[64];
        snprintf(b, sizeof(b), "%05d", step);
        const std::string stem = odir + "/" + method + "_" + name + "_" + b + "_" + step_name;
        if (pnm) write_pnm(stem + (m.type == CV_8UC3 ? ".ppm" : ".pgm"), m, multiplier);
        else write_png(stem + ".png", m, multiplier);
    };
    try {
        const double sp = env_or("MSG_SP", 10), sr = env_or("MSG_SR", 10);
        const int min_size = (int)env_or("MSG_MIN_SIZE", 50), color_dist = (int)env_or("MSG_COLOR_DIST", 10);
        int step = 0;
        Mat filtered, labels;
        GpuImgproc::pyrMeanShiftFiltering(src, filtered, sp, sr);
        save(++step, "meanshift_filtered", filtered, 1);
        int n = GpuImgproc::labelRegions(filtered, labels, 2, 2, 4);
        save(++step, "markers", labels, 1);
        std::cout << "regions after labelling: " << n << std::endl;
        n = GpuImgproc::mergeRegions(filtered, labels, min_size, color_dist);
        save(++step, "merged_markers", labels, 1);
        std::cout << "regions after merge: " << n << std::endl;
        Mat result = GpuImgproc::colorByIndexes(labels, n);           // colored=false -> white (CLI path, PictureService.java:293)
        save(++step, "result", result, 1);
        if (env_or("MSG_REFERENCE_MARKERS", 1) != 0) {
            // The reference's two pipelines, all 8 + 8 Results, step names and multipliers as it saves them
            // (PictureService.java:301-382 and :396-467), Imgproc.watershed included (msg_watershed: exact flood).
            method = "COLOR_METHOD";                                  // SegMethod.COLOR_METHOD
            step = 0;
            Mat black, sharp, gray, bw, dist, peaks, peaks8, markers;
            GpuImgproc::whiteToBlack(src, black);
            save(++step, "black_bg", black, 1);
            static const int8_t taps[9] = {1, 1, 1, 1, -8, 1, 1, 1, 1};
            GpuImgproc::sharpenLaplacian(black, sharp, taps, 9, 1);
            save(++step, "laplassian_sharp", sharp, 1);
            GpuImgproc::cvtColorBGR2GRAY(sharp, gray);
            GpuImgproc::threshold(gray, bw, 40, 255, GpuImgproc::THRESH_BINARY | GpuImgproc::THRESH_OTSU);
            save(++step, "bw", bw, 1);
            GpuImgproc::distanceTransform(bw, dist, GpuImgproc::CV_DIST_L2, 5);
            GpuImgproc::normalize(dist, dist, 0, 1., GpuImgproc::NORM_MINMAX);
            save(++step, "distance_transform", dist, 1000);
            GpuImgproc::threshold(dist, peaks, .4, 1., GpuImgproc::THRESH_BINARY);
            GpuImgproc::dilateF32(peaks, peaks, 3, 3);
            save(++step, "distance_peaks", peaks, 1000);
            GpuImgproc::convertToU8(peaks, peaks8);
            int depth = GpuImgproc::contourMarkers(peaks8, markers);
            GpuImgproc::circle(markers, 5, 5, 3, 255);
            save(++step, "markers", markers, 10000);                  // cloned BEFORE watershed (PictureService.java:369)
            std::cout << "colour-method contours: " << depth << std::endl;
            {
                Mat dst = GpuImgproc::watershedAndColor(sharp, markers, depth);    // :372: src is the sharpened image by now
                save(++step, "result", dst, 1);
                Mat bwres;
                GpuImgproc::cvtColorBGR2GRAY(dst, bwres);                          // :376-379
                save(++step, "bw_result", bwres, 1);
            }
            method = "SHAPE_METHOD";                                  // SegMethod.SHAPE_METHOD
            step = 0;
            Mat sgray, edges, d3, d5, dde, dde3, smarkers;
            const int k = GpuImgproc::calculateSizeOfSquareBlurMask(src.cols, src.rows);
            GpuImgproc::cvtColorBGR2GRAY(src, sgray);
            GpuImgproc::medianBlur(sgray, sgray, k);
            save(++step, "blured_by_" + std::to_string(k) + "x" + std::to_string(k), sgray, 1);
            GpuImgproc::Canny(sgray, edges, 5, 50);
            {
                Mat borders;
                GpuImgproc::copyToMasked(src, edges, borders);        // src.copyTo(zeros, edges), PictureService.java:417-418
                save(++step, "borders", borders, 1);
            }
            save(++step, "gray_borders", edges, 1);
            GpuImgproc::dilate(edges, d3, 3, 3);
            GpuImgproc::dilate(d3, d5, 5, 5);
            GpuImgproc::subtract(d5, d3, dde);
            save(++step, "dde_step", dde, 1);
            GpuImgproc::medianBlur(dde, dde3, 3);
            save(++step, "dde_step_blurred_3x3", dde3, 1);
            int nlab = GpuImgproc::connectedComponents(dde3, smarkers, 8);
            save(++step, "markers", smarkers, 10000);
            std::cout << "shape-method labels: " << nlab << std::endl;
            {
                // depth = contours.size() of findContours(markerMask, RETR_CCOMP): outer contours AND holes
                // (PictureService.java:450-455; empty -> the reference returns null, no further Results)
                Mat scratch;
                const int sdepth = GpuImgproc::contourMarkers(dde3, scratch);
                if (sdepth == 0) {
                    std::cout << "contours is empty" << std::endl;
                } else {
                    Mat dst = GpuImgproc::watershedAndColor(src, smarkers, sdepth);   // :457
                    save(++step, "result", dst, 1);
                    Mat bwres;
                    GpuImgproc::cvtColorBGR2GRAY(dst, bwres);                         // :461-464
                    save(++step, "bw_result", bwres, 1);
                }
            }
        }
        std::cout << "results written to " << odir << std::endl;
    } catch (const CvException& e) {                // the reference lets CvException propagate: uncaught -> non-zero exit
        std::cerr << "CvException: " << e.what() << std::endl;
        return 1;
    }
    return 0;
}
