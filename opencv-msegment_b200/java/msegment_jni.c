/* msegment_jni.c -- JNI glue between GpuImgproc.java and libmsegment_b200.so.  Mechanical: every native method
 * forwards its arguments to the C ABI of include/msegment.h.  Build (on a machine with a JDK):
 *   gcc -shared -fPIC -I$JAVA_HOME/include -I$JAVA_HOME/include/linux -I../../include msegment_jni.c \
 *       -L.. -lmsegment_b200 -o libmsegment_jni.so
 * Not built in this repository (the build image has no JDK), but type-checked on every test run against tests/stubs/jni.h
 * (gcc -fsyntax-only -Wall -Werror, tests/test_abi.py). */
#include <jni.h>
#include <stdint.h>

#include "msegment.h"

#define J(name) Java_ru_shayhulud_opencvcmsegment_gpu_GpuImgproc_##name
#define P(x) ((void*)(intptr_t)(x))

JNIEXPORT jlong JNICALL J(nCreate)(JNIEnv* e, jclass c, jint device)
{
    msg_ctx* ctx = NULL;
    return msg_create(device, &ctx) == MSG_OK ? (jlong)(intptr_t)ctx : 0;
}

JNIEXPORT jstring JNICALL J(nLastError)(JNIEnv* e, jclass c, jlong ctx)
{
    return (*e)->NewStringUTF(e, msg_last_error((const msg_ctx*)P(ctx)));
}

JNIEXPORT jint JNICALL J(nMeanshift)(JNIEnv* e, jclass c, jlong ctx, jlong src, jlong sstep, jlong dst, jlong dstep, jint w,
                                     jint h, jdouble sp, jdouble sr, jint ml, jint tt, jint mc, jdouble eps)
{
    return msg_meanshift_filter((msg_ctx*)P(ctx), (const uint8_t*)P(src), (size_t)sstep, (uint8_t*)P(dst), (size_t)dstep, w, h,
                                sp, sr, ml, tt, mc, eps);
}

JNIEXPORT jint JNICALL J(nLabelRegions)(JNIEnv* e, jclass c, jlong ctx, jlong img, jlong step, jlong lab, jlong lstep, jint w,
                                        jint h, jint lo, jint up, jint conn, jintArray n)
{
    int32_t cnt = 0;
    int rc = msg_label_regions((msg_ctx*)P(ctx), (const uint8_t*)P(img), (size_t)step, (int32_t*)P(lab), (size_t)lstep, w, h, lo,
                               up, conn, &cnt);
    jint v = cnt;
    (*e)->SetIntArrayRegion(e, n, 0, 1, &v);
    return rc;
}

JNIEXPORT jint JNICALL J(nMergeRegions)(JNIEnv* e, jclass c, jlong ctx, jlong img, jlong step, jlong lab, jlong lstep, jint w,
                                        jint h, jint min_size, jint color_dist, jintArray n)
{
    int32_t cnt = 0;
    int rc = msg_merge_regions((msg_ctx*)P(ctx), (const uint8_t*)P(img), (size_t)step, (int32_t*)P(lab), (size_t)lstep, w, h,
                               min_size, color_dist, &cnt);
    jint v = cnt;
    (*e)->SetIntArrayRegion(e, n, 0, 1, &v);
    return rc;
}

JNIEXPORT jint JNICALL J(nConnectedComponents)(JNIEnv* e, jclass c, jlong ctx, jlong img, jlong step, jlong lab, jlong lstep,
                                               jint w, jint h, jint conn, jintArray n)
{
    int32_t cnt = 0;
    int rc = msg_connected_components((msg_ctx*)P(ctx), (const uint8_t*)P(img), (size_t)step, (int32_t*)P(lab), (size_t)lstep, w,
                                      h, conn, &cnt);
    jint v = cnt;
    (*e)->SetIntArrayRegion(e, n, 0, 1, &v);
    return rc;
}

JNIEXPORT jint JNICALL J(nRender)(JNIEnv* e, jclass c, jlong ctx, jlong lab, jlong lstep, jlong dst, jlong dstep, jint w, jint h,
                                  jint depth, jbyteArray colors)
{
    jbyte* col = colors ? (*e)->GetByteArrayElements(e, colors, NULL) : NULL;
    int rc = msg_render_labels((msg_ctx*)P(ctx), (const int32_t*)P(lab), (size_t)lstep, (uint8_t*)P(dst), (size_t)dstep, w, h,
                               depth, (const uint8_t*)col);
    if (col) (*e)->ReleaseByteArrayElements(e, colors, col, JNI_ABORT);
    return rc;
}

JNIEXPORT jint JNICALL J(nSharpen)(JNIEnv* e, jclass c, jlong ctx, jlong src, jlong sstep, jlong dst, jlong dstep, jint w, jint h,
                                   jbyteArray taps, jint krows, jint kcols)
{
    jbyte* t = (*e)->GetByteArrayElements(e, taps, NULL);
    int rc = msg_laplacian_sharpen((msg_ctx*)P(ctx), (const uint8_t*)P(src), (size_t)sstep, (uint8_t*)P(dst), (size_t)dstep, w, h,
                                   (const int8_t*)t, krows, kcols);
    (*e)->ReleaseByteArrayElements(e, taps, t, JNI_ABORT);
    return rc;
}

JNIEXPORT jint JNICALL J(nGray)(JNIEnv* e, jclass c, jlong ctx, jlong src, jlong sstep, jlong dst, jlong dstep, jint w, jint h)
{
    return msg_bgr2gray((msg_ctx*)P(ctx), (const uint8_t*)P(src), (size_t)sstep, (uint8_t*)P(dst), (size_t)dstep, w, h);
}

JNIEXPORT jint JNICALL J(nMedian)(JNIEnv* e, jclass c, jlong ctx, jlong src, jlong sstep, jlong dst, jlong dstep, jint w, jint h,
                                  jint ksize)
{
    return msg_median_blur((msg_ctx*)P(ctx), (const uint8_t*)P(src), (size_t)sstep, (uint8_t*)P(dst), (size_t)dstep, w, h, ksize);
}

JNIEXPORT jint JNICALL J(nCanny)(JNIEnv* e, jclass c, jlong ctx, jlong src, jlong sstep, jlong dst, jlong dstep, jint w, jint h,
                                 jdouble t1, jdouble t2)
{
    return msg_canny((msg_ctx*)P(ctx), (const uint8_t*)P(src), (size_t)sstep, (uint8_t*)P(dst), (size_t)dstep, w, h, t1, t2);
}

JNIEXPORT jint JNICALL J(nDilate)(JNIEnv* e, jclass c, jlong ctx, jlong src, jlong sstep, jlong dst, jlong dstep, jint w, jint h,
                                  jint kw, jint kh)
{
    return msg_dilate((msg_ctx*)P(ctx), (const uint8_t*)P(src), (size_t)sstep, (uint8_t*)P(dst), (size_t)dstep, w, h, kw, kh);
}

JNIEXPORT jint JNICALL J(nSubtract)(JNIEnv* e, jclass c, jlong ctx, jlong a, jlong astep, jlong b, jlong bstep, jlong dst,
                                    jlong dstep, jint w, jint h)
{
    return msg_subtract((msg_ctx*)P(ctx), (const uint8_t*)P(a), (size_t)astep, (const uint8_t*)P(b), (size_t)bstep,
                        (uint8_t*)P(dst), (size_t)dstep, w, h);
}

JNIEXPORT jint JNICALL J(nShapeSeeds)(JNIEnv* e, jclass c, jlong ctx, jlong src, jlong sstep, jint w, jint h, jint ksize,
                                      jdouble t1, jdouble t2, jlong markers, jlong mstep, jintArray n)
{
    int32_t count = 0;
    int rc = msg_shape_seeds((msg_ctx*)P(ctx), (const uint8_t*)P(src), (size_t)sstep, w, h, ksize, t1, t2, (int32_t*)P(markers),
                             (size_t)mstep, &count, NULL, 0);
    if (rc == 0 && n) { jint v = count; (*e)->SetIntArrayRegion(e, n, 0, 1, &v); }
    return rc;
}

/* ---- colour-method marker generator and bilateral filter ---- */
JNIEXPORT jint JNICALL J(nWhiteToBlack)(JNIEnv* e, jclass c, jlong ctx, jlong src, jlong sstep, jlong dst, jlong dstep, jint w, jint h)
{
    return msg_white_to_black((msg_ctx*)P(ctx), (const uint8_t*)P(src), (size_t)sstep, (uint8_t*)P(dst), (size_t)dstep, w, h);
}

JNIEXPORT jint JNICALL J(nThreshold)(JNIEnv* e, jclass c, jlong ctx, jlong src, jlong sstep, jlong dst, jlong dstep, jint w, jint h,
                                     jdouble thresh, jdouble maxval, jint type, jdoubleArray used)
{
    double u = 0;
    int rc = msg_threshold((msg_ctx*)P(ctx), (const uint8_t*)P(src), (size_t)sstep, (uint8_t*)P(dst), (size_t)dstep, w, h, thresh,
                           maxval, type, &u);
    if (rc == 0 && used) { jdouble v = u; (*e)->SetDoubleArrayRegion(e, used, 0, 1, &v); }
    return rc;
}

JNIEXPORT jint JNICALL J(nThresholdF32)(JNIEnv* e, jclass c, jlong ctx, jlong src, jlong sstep, jlong dst, jlong dstep, jint w,
                                        jint h, jdouble thresh, jdouble maxval)
{
    return msg_threshold_f32((msg_ctx*)P(ctx), (const float*)P(src), (size_t)sstep, (float*)P(dst), (size_t)dstep, w, h, thresh, maxval);
}

JNIEXPORT jint JNICALL J(nDistanceTransform)(JNIEnv* e, jclass c, jlong ctx, jlong src, jlong sstep, jlong dst, jlong dstep, jint w,
                                             jint h, jint dist_type, jint mask_size)
{
    return msg_distance_transform((msg_ctx*)P(ctx), (const uint8_t*)P(src), (size_t)sstep, (float*)P(dst), (size_t)dstep, w, h,
                                  dist_type, mask_size);
}

JNIEXPORT jint JNICALL J(nNormalize)(JNIEnv* e, jclass c, jlong ctx, jlong src, jlong sstep, jlong dst, jlong dstep, jint w, jint h,
                                     jdouble alpha, jdouble beta)
{
    return msg_normalize_minmax((msg_ctx*)P(ctx), (const float*)P(src), (size_t)sstep, (float*)P(dst), (size_t)dstep, w, h, alpha, beta);
}

JNIEXPORT jint JNICALL J(nDilateF32)(JNIEnv* e, jclass c, jlong ctx, jlong src, jlong sstep, jlong dst, jlong dstep, jint w, jint h,
                                     jint kw, jint kh)
{
    return msg_dilate_f32((msg_ctx*)P(ctx), (const float*)P(src), (size_t)sstep, (float*)P(dst), (size_t)dstep, w, h, kw, kh);
}

JNIEXPORT jint JNICALL J(nConvertU8)(JNIEnv* e, jclass c, jlong ctx, jlong src, jlong sstep, jlong dst, jlong dstep, jint w, jint h)
{
    return msg_convert_f32_to_u8((msg_ctx*)P(ctx), (const float*)P(src), (size_t)sstep, (uint8_t*)P(dst), (size_t)dstep, w, h);
}

JNIEXPORT jint JNICALL J(nContourMarkers)(JNIEnv* e, jclass c, jlong ctx, jlong img, jlong step, jlong markers, jlong mstep, jint w,
                                          jint h, jintArray n)
{
    int32_t count = 0;
    int rc = msg_contour_markers((msg_ctx*)P(ctx), (const uint8_t*)P(img), (size_t)step, (int32_t*)P(markers), (size_t)mstep, w, h,
                                 &count);
    if (rc == 0 && n) { jint v = count; (*e)->SetIntArrayRegion(e, n, 0, 1, &v); }
    return rc;
}

JNIEXPORT jint JNICALL J(nCircle)(JNIEnv* e, jclass c, jlong ctx, jlong img, jlong step, jint w, jint h, jint cx, jint cy, jint radius,
                                  jint value)
{
    return msg_circle_filled((msg_ctx*)P(ctx), (int32_t*)P(img), (size_t)step, w, h, cx, cy, radius, value);
}

JNIEXPORT jint JNICALL J(nColorSeeds)(JNIEnv* e, jclass c, jlong ctx, jlong src, jlong sstep, jint w, jint h, jbyteArray taps,
                                      jint krows, jint kcols, jdouble peak, jlong markers, jlong mstep, jintArray n)
{
    int32_t count = 0;
    jbyte* t = (*e)->GetByteArrayElements(e, taps, NULL);
    int rc = msg_color_seeds((msg_ctx*)P(ctx), (const uint8_t*)P(src), (size_t)sstep, w, h, (const int8_t*)t, krows, kcols, peak,
                             (int32_t*)P(markers), (size_t)mstep, &count, NULL, 0, NULL, 0, NULL, 0, NULL, 0);
    (*e)->ReleaseByteArrayElements(e, taps, t, JNI_ABORT);
    if (rc == 0 && n) { jint v = count; (*e)->SetIntArrayRegion(e, n, 0, 1, &v); }
    return rc;
}

JNIEXPORT jint JNICALL J(nBilateral)(JNIEnv* e, jclass c, jlong ctx, jlong src, jlong sstep, jlong dst, jlong dstep, jint w, jint h,
                                     jint channels, jint d, jdouble sigma_color, jdouble sigma_space)
{
    return msg_bilateral_filter((msg_ctx*)P(ctx), (const uint8_t*)P(src), (size_t)sstep, (uint8_t*)P(dst), (size_t)dstep, w, h,
                                channels, d, sigma_color, sigma_space);
}

JNIEXPORT jint JNICALL J(nWatershed)(JNIEnv* e, jclass c, jlong ctx, jlong img, jlong step, jlong markers, jlong mstep, jint w, jint h)
{
    return msg_watershed((msg_ctx*)P(ctx), (const uint8_t*)P(img), (size_t)step, (int32_t*)P(markers), (size_t)mstep, w, h);
}

JNIEXPORT jint JNICALL J(nCopyMasked)(JNIEnv* e, jclass c, jlong ctx, jlong src, jlong sstep, jlong mask, jlong mstep, jlong dst,
                                      jlong dstep, jint w, jint h)
{
    return msg_copy_masked((msg_ctx*)P(ctx), (const uint8_t*)P(src), (size_t)sstep, (const uint8_t*)P(mask), (size_t)mstep,
                           (uint8_t*)P(dst), (size_t)dstep, w, h);
}

JNIEXPORT jint JNICALL J(nSegment)(JNIEnv* e, jclass c, jlong ctx, jlong src, jlong sstep, jint w, jint h, jdouble sp, jdouble sr,
                                   jint max_level, jint lo_diff, jint min_size, jint color_dist, jint labels16, jlong filtered,
                                   jlong fstep, jlong labels, jlong lstep, jintArray n)
{
    msg_segment_params p;
    int32_t count = 0;
    msg_segment_params_default(&p);
    p.sp = sp; p.sr = sr; p.max_level = max_level; p.lo_diff = lo_diff; p.min_size = min_size; p.color_dist = color_dist;
    p.render_depth = -1;
    p.labels_type = labels16 ? MSG_LABELS_16U : MSG_LABELS_32S;
    int rc = msg_segment((msg_ctx*)P(ctx), (const uint8_t*)P(src), (size_t)sstep, w, h, &p, (uint8_t*)P(filtered), (size_t)fstep,
                         P(labels), (size_t)lstep, NULL, 0, &count);
    if (n) { jint v = count; (*e)->SetIntArrayRegion(e, n, 0, 1, &v); }
    return rc;
}

JNIEXPORT jint JNICALL J(nSetOption)(JNIEnv* e, jclass c, jlong ctx, jstring name, jint value)
{
    const char* s = (*e)->GetStringUTFChars(e, name, NULL);
    int rc = s ? msg_set_option((msg_ctx*)P(ctx), s, value) : MSG_EINVAL;
    if (s) (*e)->ReleaseStringUTFChars(e, name, s);
    return rc;
}

JNIEXPORT jint JNICALL J(nRegisterHost)(JNIEnv* e, jclass c, jlong ctx, jlong ptr, jlong bytes)
{
    return msg_register_host((msg_ctx*)P(ctx), P(ptr), (size_t)bytes);
}

JNIEXPORT jint JNICALL J(nUnregisterHost)(JNIEnv* e, jclass c, jlong ctx, jlong ptr)
{
    return msg_unregister_host((msg_ctx*)P(ctx), P(ptr));
}
