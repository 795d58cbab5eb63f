package ru.shayhulud.opencvcmsegment.gpu;

import org.opencv.core.CvException;
import org.opencv.core.CvType;
import org.opencv.core.Mat;
import org.opencv.core.Point;
import org.opencv.core.Scalar;
import org.opencv.core.TermCriteria;

/**
 * Drop-in sibling of {@code org.opencv.imgproc.Imgproc} for the segmentation hot path, bound to
 * libmsegment_b200.so (include/msegment.h) through JNI (msegment_jni.c).  Same static signatures taking
 * {@code Mat}; swap {@code Imgproc.} for {@code GpuImgproc.} at PictureService.java:442 (connectedComponents)
 * and :913-936 (colorByIndexes), and call pyrMeanShiftFiltering / labelRegions / mergeRegions where the
 * mean-shift pipeline replaces the watershed stage.  NOT compiled in this repository (no JDK in the build image).
 */
public final class GpuImgproc {

	static {
		System.loadLibrary("msegment_jni"); // which links libmsegment_b200.so
	}

	private static final ThreadLocal<Long> CTX = ThreadLocal.withInitial(() -> check(nCreate(0)));

	private GpuImgproc() {
	}

	public static void pyrMeanShiftFiltering(Mat src, Mat dst, double sp, double sr) {
		pyrMeanShiftFiltering(src, dst, sp, sr, 1, new TermCriteria(TermCriteria.COUNT + TermCriteria.EPS, 5, 1));
	}

	public static void pyrMeanShiftFiltering(Mat src, Mat dst, double sp, double sr, int maxLevel, TermCriteria tc) {
		require(src.type() == CvType.CV_8UC3, "src must be CV_8UC3");
		dst.create(src.size(), src.type());
		status(nMeanshift(CTX.get(), src.dataAddr(), src.step1() * src.elemSize1(), dst.dataAddr(),
			dst.step1() * dst.elemSize1(), src.cols(), src.rows(), sp, sr, maxLevel, tc.type, tc.maxCount, tc.epsilon));
	}

	/** floodFill region-growing loop (loDiff == upDiff, 4-connectivity) as one call; returns the region count. */
	public static int labelRegions(Mat image, Mat labels, int loDiff, int upDiff, int connectivity) {
		require(image.type() == CvType.CV_8UC3, "image must be CV_8UC3");
		labels.create(image.size(), CvType.CV_32SC1);
		int[] n = new int[1];
		status(nLabelRegions(CTX.get(), image.dataAddr(), image.step1(), labels.dataAddr(), labels.step1() * 4,
			image.cols(), image.rows(), loDiff, upDiff, connectivity, n));
		return n[0];
	}

	public static int mergeRegions(Mat image, Mat labels, int minSize, int colorDist) {
		int[] n = new int[1];
		status(nMergeRegions(CTX.get(), image.dataAddr(), image.step1(), labels.dataAddr(), labels.step1() * 4,
			image.cols(), image.rows(), minSize, colorDist, n));
		return n[0];
	}

	/** Same contract as Imgproc.connectedComponents(image, labels, connectivity, ltype) -- PictureService.java:442. */
	public static int connectedComponents(Mat image, Mat labels, int connectivity, int ltype) {
		require(image.type() == CvType.CV_8UC1 && ltype == CvType.CV_32S, "CV_8UC1 image and CV_32S labels only");
		labels.create(image.size(), CvType.CV_32SC1);
		int[] n = new int[1];
		status(nConnectedComponents(CTX.get(), image.dataAddr(), image.step1(), labels.dataAddr(), labels.step1() * 4,
			image.cols(), image.rows(), connectivity, n));
		return n[0];
	}

	/** Same contract as Imgproc.watershed(image, markers) -- PictureService.java:909: markers CV_32SC1, modified in place. */
	public static void watershed(Mat image, Mat markers) {
		require(image.type() == CvType.CV_8UC3 && markers.type() == CvType.CV_32SC1, "image CV_8UC3, markers CV_32SC1");
		require(image.rows() == markers.rows() && image.cols() == markers.cols(), "markers must have the size of image");
		status(nWatershed(CTX.get(), image.dataAddr(), image.step1(), markers.dataAddr(), markers.step1() * 4, image.cols(),
			image.rows()));
	}

	/** src.copyTo(dst, mask) onto Mat.zeros -- PictureService.java:417-418 ("borders"). */
	public static void copyToMasked(Mat src, Mat mask, Mat dst) {
		require(src.type() == CvType.CV_8UC3 && mask.type() == CvType.CV_8UC1, "src CV_8UC3, mask CV_8UC1");
		dst.create(src.size(), src.type());
		status(nCopyMasked(CTX.get(), src.dataAddr(), src.step1(), mask.dataAddr(), mask.step1(), dst.dataAddr(), dst.step1(),
			src.cols(), src.rows()));
	}

	/**
	 * Fused mean shift + labelling + merge, intermediates in HBM (msg_segment).  labels is created as CV_32SC1, or as CV_16UC1
	 * when labels16 is set (half the download; CvException beyond 65535 regions).  filtered may be null (not downloaded).
	 * Returns the region count.
	 */
	public static int segment(Mat src, Mat filtered, Mat labels, double sp, double sr, int maxLevel, int loDiff, int minSize,
		int colorDist, boolean labels16) {
		require(src.type() == CvType.CV_8UC3, "src must be CV_8UC3");
		if (filtered != null) {
			filtered.create(src.size(), src.type());
		}
		labels.create(src.size(), labels16 ? CvType.CV_16UC1 : CvType.CV_32SC1);
		int[] n = new int[1];
		status(nSegment(CTX.get(), src.dataAddr(), src.step1(), src.cols(), src.rows(), sp, sr, maxLevel, loDiff, minSize, colorDist,
			labels16 ? 1 : 0, filtered == null ? 0 : filtered.dataAddr(), filtered == null ? 0 : filtered.step1(),
			labels.dataAddr(), labels.step1() * labels.elemSize1(), n));
		return n[0];
	}

	/** msg_set_option, e.g. ("gray_compat", 1) for the OpenCV 3.4.2 BGR2GRAY coefficients the reference's natives use. */
	public static void setOption(String name, int value) {
		status(nSetOption(CTX.get(), name, value));
	}

	/** Page-lock a long-lived Mat so that calls copy from / to it without staging; unregister before releasing it. */
	public static void registerMat(Mat m) {
		status(nRegisterHost(CTX.get(), m.dataAddr(), m.step1() * m.elemSize1() * m.rows()));
	}

	public static void unregisterMat(Mat m) {
		status(nUnregisterHost(CTX.get(), m.dataAddr()));
	}

	/** PictureService.colorByIndexes (PictureService.java:913-936); colors == null renders white. */
	public static Mat colorByIndexes(Mat markers, int depth, byte[] colorsBgr) {
		Mat dst = new Mat(markers.size(), CvType.CV_8UC3);
		status(nRender(CTX.get(), markers.dataAddr(), markers.step1() * 4, dst.dataAddr(), dst.step1(), markers.cols(),
			markers.rows(), depth, colorsBgr));
		return dst;
	}

	/** filter2D(CV_32F, kernel) + convertTo + subtract + convertTo(CV_8UC3) of PictureService.java:323-333 as one call. */
	public static void sharpenLaplacian(Mat src, Mat dst, byte[] taps, int krows, int kcols) {
		require(src.type() == CvType.CV_8UC3, "src must be CV_8UC3");
		dst.create(src.size(), src.type());
		status(nSharpen(CTX.get(), src.dataAddr(), src.step1(), dst.dataAddr(), dst.step1(), src.cols(), src.rows(), taps,
			krows, kcols));
	}

	/** Imgproc.cvtColor(src, dst, Imgproc.COLOR_BGR2GRAY). */
	public static void cvtColorBGR2GRAY(Mat src, Mat dst) {
		require(src.type() == CvType.CV_8UC3, "src must be CV_8UC3");
		dst.create(src.size(), CvType.CV_8UC1);
		status(nGray(CTX.get(), src.dataAddr(), src.step1(), dst.dataAddr(), dst.step1(), src.cols(), src.rows()));
	}

	/** Imgproc.medianBlur(src, dst, ksize) for CV_8UC1 (PictureService.java:408, :436). */
	public static void medianBlur(Mat src, Mat dst, int ksize) {
		require(src.type() == CvType.CV_8UC1, "src must be CV_8UC1");
		Mat out = new Mat(src.size(), CvType.CV_8UC1);
		status(nMedian(CTX.get(), src.dataAddr(), src.step1(), out.dataAddr(), out.step1(), src.cols(), src.rows(), ksize));
		out.copyTo(dst);
	}

	/** Imgproc.Canny(image, edges, threshold1, threshold2) (PictureService.java:416). */
	public static void Canny(Mat image, Mat edges, double threshold1, double threshold2) {
		require(image.type() == CvType.CV_8UC1, "image must be CV_8UC1");
		Mat out = new Mat(image.size(), CvType.CV_8UC1);
		status(nCanny(CTX.get(), image.dataAddr(), image.step1(), out.dataAddr(), out.step1(), image.cols(), image.rows(),
			threshold1, threshold2));
		out.copyTo(edges);
	}

	/** Imgproc.dilate(src, dst, Mat.ones(krows, kcols, type)) (PictureService.java:428-429). */
	public static void dilate(Mat src, Mat dst, int krows, int kcols) {
		require(src.type() == CvType.CV_8UC1, "src must be CV_8UC1");
		Mat out = new Mat(src.size(), CvType.CV_8UC1);
		status(nDilate(CTX.get(), src.dataAddr(), src.step1(), out.dataAddr(), out.step1(), src.cols(), src.rows(), kcols, krows));
		out.copyTo(dst);
	}

	/** Core.subtract(src1, src2, dst) on CV_8UC1 (PictureService.java:430). */
	public static void subtract(Mat src1, Mat src2, Mat dst) {
		require(src1.type() == CvType.CV_8UC1 && src2.type() == CvType.CV_8UC1 && src1.size().equals(src2.size()),
			"subtract: CV_8UC1 of equal size");
		Mat out = new Mat(src1.size(), CvType.CV_8UC1);
		status(nSubtract(CTX.get(), src1.dataAddr(), src1.step1(), src2.dataAddr(), src2.step1(), out.dataAddr(), out.step1(),
			src1.cols(), src1.rows()));
		out.copyTo(dst);
	}

	/**
	 * Marker half of PictureService.shapeAutoMarkerWatershed (PictureService.java:404-442) as one call, intermediates on
	 * the device: gray, medianBlur(medianKsize), Canny(low, low * ratio), dilate 3x3, dilate 5x5, subtract, medianBlur 3,
	 * connectedComponents(8).  Returns the label count, background included.
	 */
	public static int shapeSeeds(Mat src, Mat markers, int medianKsize, double lowThreshold, double ratio) {
		require(src.type() == CvType.CV_8UC3, "src must be CV_8UC3");
		markers.create(src.size(), CvType.CV_32SC1);
		int[] n = new int[1];
		status(nShapeSeeds(CTX.get(), src.dataAddr(), src.step1(), src.cols(), src.rows(), medianKsize, lowThreshold,
			lowThreshold * ratio, markers.dataAddr(), markers.step1() * 4, n));
		return n[0];
	}

	/** The Java loop of PictureService.java:309-318 as one kernel: (255,255,255) becomes (0,0,0). */
	public static void whiteToBlack(Mat src, Mat dst) {
		require(src.type() == CvType.CV_8UC3, "src must be CV_8UC3");
		Mat out = new Mat(src.size(), CvType.CV_8UC3);
		status(nWhiteToBlack(CTX.get(), src.dataAddr(), src.step1(), out.dataAddr(), out.step1(), src.cols(), src.rows()));
		out.copyTo(dst);
	}

	/** Imgproc.threshold(src, dst, thresh, maxval, type) (PictureService.java:941, :348): CV_8UC1 BINARY [| OTSU], CV_32FC1 BINARY. */
	public static double threshold(Mat src, Mat dst, double thresh, double maxval, int type) {
		require(src.type() == CvType.CV_8UC1 || src.type() == CvType.CV_32FC1, "src must be CV_8UC1 or CV_32FC1");
		Mat out = new Mat(src.size(), src.type());
		double[] used = new double[] {thresh};
		if (src.type() == CvType.CV_8UC1) {
			status(nThreshold(CTX.get(), src.dataAddr(), src.step1(), out.dataAddr(), out.step1(), src.cols(), src.rows(),
				thresh, maxval, type, used));
		} else {
			require(type == 0, "CV_32FC1: THRESH_BINARY only");
			status(nThresholdF32(CTX.get(), src.dataAddr(), src.step1() * 4, out.dataAddr(), out.step1() * 4, src.cols(),
				src.rows(), thresh, maxval));
		}
		out.copyTo(dst);
		return used[0];
	}

	/** Imgproc.distanceTransform(src, dst, Imgproc.CV_DIST_L2, 5) (PictureService.java:1020). */
	public static void distanceTransform(Mat src, Mat dst, int distanceType, int maskSize) {
		require(src.type() == CvType.CV_8UC1, "src must be CV_8UC1");
		Mat out = new Mat(src.size(), CvType.CV_32FC1);
		status(nDistanceTransform(CTX.get(), src.dataAddr(), src.step1(), out.dataAddr(), out.step1() * 4, src.cols(),
			src.rows(), distanceType, maskSize));
		out.copyTo(dst);
	}

	/** Core.normalize(src, dst, alpha, beta, Core.NORM_MINMAX) on CV_32FC1 (PictureService.java:1021). */
	public static void normalize(Mat src, Mat dst, double alpha, double beta, int normType) {
		require(src.type() == CvType.CV_32FC1 && normType == 32, "normalize: CV_32FC1, NORM_MINMAX");
		Mat out = new Mat(src.size(), CvType.CV_32FC1);
		status(nNormalize(CTX.get(), src.dataAddr(), src.step1() * 4, out.dataAddr(), out.step1() * 4, src.cols(), src.rows(),
			alpha, beta));
		out.copyTo(dst);
	}

	/** Imgproc.dilate on CV_32FC1 with Mat.ones(krows, kcols) (PictureService.java:349-350). */
	public static void dilateF32(Mat src, Mat dst, int krows, int kcols) {
		require(src.type() == CvType.CV_32FC1, "src must be CV_32FC1");
		Mat out = new Mat(src.size(), CvType.CV_32FC1);
		status(nDilateF32(CTX.get(), src.dataAddr(), src.step1() * 4, out.dataAddr(), out.step1() * 4, src.cols(), src.rows(),
			kcols, krows));
		out.copyTo(dst);
	}

	/** Mat.convertTo(dst, CvType.CV_8U) from CV_32FC1 (PictureService.java:355-356). */
	public static void convertToU8(Mat src, Mat dst) {
		require(src.type() == CvType.CV_32FC1, "src must be CV_32FC1");
		Mat out = new Mat(src.size(), CvType.CV_8UC1);
		status(nConvertU8(CTX.get(), src.dataAddr(), src.step1() * 4, out.dataAddr(), out.step1(), src.cols(), src.rows()));
		out.copyTo(dst);
	}

	/**
	 * findContours(image, contours, hierarchy, RETR_CCOMP, CHAIN_APPROX_NONE) and the drawContours(markers, contours, i,
	 * Scalar.all(i + 1), -1, 8, hierarchy, Integer.MAX_VALUE, new Point()) loop of PictureService.java:360-364 in one call.
	 * Returns contours.size() (the reference's depth, :365).
	 */
	public static int contourMarkers(Mat image, Mat markers) {
		require(image.type() == CvType.CV_8UC1, "image must be CV_8UC1");
		markers.create(image.size(), CvType.CV_32SC1);
		int[] n = new int[1];
		status(nContourMarkers(CTX.get(), image.dataAddr(), image.step1(), markers.dataAddr(), markers.step1() * 4,
			image.cols(), image.rows(), n));
		return n[0];
	}

	/** Imgproc.circle(img, center, radius, color, -1) on CV_32SC1 (PictureService.java:366). */
	public static void circle(Mat img, Point center, int radius, Scalar color) {
		require(img.type() == CvType.CV_32SC1, "img must be CV_32SC1");
		status(nCircle(CTX.get(), img.dataAddr(), img.step1() * 4, img.cols(), img.rows(), (int) center.x, (int) center.y,
			radius, (int) color.val[0]));
	}

	/**
	 * Marker half of PictureService.colorAutoMarkerWatershed (PictureService.java:309-366) as one call, intermediates on the
	 * device; kernel = the sharpen taps (the reference's MatOfFloat(1,1,1,1,-8,1,1,1,1) is 9 x 1).  Returns contours.size().
	 */
	public static int colorSeeds(Mat src, Mat markers, byte[] taps, int krows, int kcols, double peakThresh) {
		require(src.type() == CvType.CV_8UC3, "src must be CV_8UC3");
		markers.create(src.size(), CvType.CV_32SC1);
		int[] n = new int[1];
		status(nColorSeeds(CTX.get(), src.dataAddr(), src.step1(), src.cols(), src.rows(), taps, krows, kcols, peakThresh,
			markers.dataAddr(), markers.step1() * 4, n));
		return n[0];
	}

	/** Imgproc.bilateralFilter(src, dst, d, sigmaColor, sigmaSpace) on CV_8UC1 / CV_8UC3 (PictureService.java:490). */
	public static void bilateralFilter(Mat src, Mat dst, int d, double sigmaColor, double sigmaSpace) {
		require(src.type() == CvType.CV_8UC1 || src.type() == CvType.CV_8UC3, "src must be CV_8UC1 or CV_8UC3");
		Mat out = new Mat(src.size(), src.type());
		status(nBilateral(CTX.get(), src.dataAddr(), src.step1(), out.dataAddr(), out.step1(), src.cols(), src.rows(),
			src.channels(), d, sigmaColor, sigmaSpace));
		out.copyTo(dst);
	}

	private static void require(boolean ok, String msg) {
		if (!ok) {
			throw new CvException(msg);
		}
	}

	private static void status(int rc) {
		if (rc != 0) {
			throw new CvException("msegment status " + rc + ": " + nLastError(CTX.get()));
		}
	}

	private static long check(long handle) {
		if (handle == 0) {
			throw new CvException("msg_create failed: " + nLastError(0));
		}
		return handle;
	}

	private static native long nCreate(int device);
	private static native String nLastError(long ctx);
	private static native int nMeanshift(long ctx, long src, long sstep, long dst, long dstep, int w, int h, double sp,
		double sr, int maxLevel, int termType, int maxCount, double eps);
	private static native int nLabelRegions(long ctx, long img, long step, long labels, long lstep, int w, int h, int lo,
		int up, int conn, int[] n);
	private static native int nMergeRegions(long ctx, long img, long step, long labels, long lstep, int w, int h,
		int minSize, int colorDist, int[] n);
	private static native int nConnectedComponents(long ctx, long img, long step, long labels, long lstep, int w, int h,
		int conn, int[] n);
	private static native int nSharpen(long ctx, long src, long sstep, long dst, long dstep, int w, int h, byte[] taps,
		int krows, int kcols);
	private static native int nGray(long ctx, long src, long sstep, long dst, long dstep, int w, int h);
	private static native int nMedian(long ctx, long src, long sstep, long dst, long dstep, int w, int h, int ksize);
	private static native int nCanny(long ctx, long src, long sstep, long dst, long dstep, int w, int h, double t1, double t2);
	private static native int nDilate(long ctx, long src, long sstep, long dst, long dstep, int w, int h, int kw, int kh);
	private static native int nSubtract(long ctx, long a, long astep, long b, long bstep, long dst, long dstep, int w, int h);
	private static native int nShapeSeeds(long ctx, long src, long sstep, int w, int h, int ksize, double t1, double t2,
		long markers, long mstep, int[] n);
	private static native int nWhiteToBlack(long ctx, long src, long sstep, long dst, long dstep, int w, int h);
	private static native int nThreshold(long ctx, long src, long sstep, long dst, long dstep, int w, int h, double thresh,
		double maxval, int type, double[] used);
	private static native int nThresholdF32(long ctx, long src, long sstep, long dst, long dstep, int w, int h, double thresh,
		double maxval);
	private static native int nDistanceTransform(long ctx, long src, long sstep, long dst, long dstep, int w, int h,
		int distType, int maskSize);
	private static native int nNormalize(long ctx, long src, long sstep, long dst, long dstep, int w, int h, double alpha,
		double beta);
	private static native int nDilateF32(long ctx, long src, long sstep, long dst, long dstep, int w, int h, int kw, int kh);
	private static native int nConvertU8(long ctx, long src, long sstep, long dst, long dstep, int w, int h);
	private static native int nContourMarkers(long ctx, long img, long step, long markers, long mstep, int w, int h, int[] n);
	private static native int nCircle(long ctx, long img, long step, int w, int h, int cx, int cy, int radius, int value);
	private static native int nColorSeeds(long ctx, long src, long sstep, int w, int h, byte[] taps, int krows, int kcols,
		double peakThresh, long markers, long mstep, int[] n);
	private static native int nBilateral(long ctx, long src, long sstep, long dst, long dstep, int w, int h, int channels,
		int d, double sigmaColor, double sigmaSpace);
	private static native int nWatershed(long ctx, long img, long step, long markers, long mstep, int w, int h);
	private static native int nCopyMasked(long ctx, long src, long sstep, long mask, long mstep, long dst, long dstep, int w,
		int h);
	private static native int nSegment(long ctx, long src, long sstep, int w, int h, double sp, double sr, int maxLevel,
		int loDiff, int minSize, int colorDist, int labels16, long filtered, long fstep, long labels, long lstep, int[] n);
	private static native int nSetOption(long ctx, String name, int value);
	private static native int nRegisterHost(long ctx, long ptr, long bytes);
	private static native int nUnregisterHost(long ctx, long ptr);
	private static native int nRender(long ctx, long labels, long lstep, long dst, long dstep, int w, int h, int depth,
		byte[] colors);
}
