/* tests/stubs/jni.h -- a minimal stand-in for the JDK's <jni.h>, ONLY for type-checking opencv-msegment_b200/java/msegment_jni.c
 * in an image without a JDK (gcc -fsyntax-only; tests/test_abi.py).  It declares the JNI types and exactly the JNIEnv function
 * table entries the glue uses, with the signatures of the JNI specification (Java SE 8, chapter 4).  The entries are NOT at
 * their real table offsets: nothing compiled against this header may ever be linked or run. */
#ifndef MSEGMENT_TEST_STUB_JNI_H
#define MSEGMENT_TEST_STUB_JNI_H

#include <stdint.h>

typedef int32_t jint;
typedef int64_t jlong;
typedef int8_t jbyte;
typedef uint8_t jboolean;
typedef double jdouble;
typedef jint jsize;

struct _jobject;
typedef struct _jobject* jobject;
typedef jobject jclass;
typedef jobject jstring;
typedef jobject jarray;
typedef jarray jintArray;
typedef jarray jbyteArray;
typedef jarray jdoubleArray;

#define JNIEXPORT __attribute__((visibility("default")))
#define JNICALL
#define JNI_ABORT 2
#define JNI_COMMIT 1

struct JNINativeInterface_;
typedef const struct JNINativeInterface_* JNIEnv;

struct JNINativeInterface_ {
    jstring (*NewStringUTF)(JNIEnv* env, const char* utf);
    const char* (*GetStringUTFChars)(JNIEnv* env, jstring str, jboolean* isCopy);
    void (*ReleaseStringUTFChars)(JNIEnv* env, jstring str, const char* chars);
    jbyte* (*GetByteArrayElements)(JNIEnv* env, jbyteArray array, jboolean* isCopy);
    void (*ReleaseByteArrayElements)(JNIEnv* env, jbyteArray array, jbyte* elems, jint mode);
    void (*SetIntArrayRegion)(JNIEnv* env, jintArray array, jsize start, jsize len, const jint* buf);
    void (*SetDoubleArrayRegion)(JNIEnv* env, jdoubleArray array, jsize start, jsize len, const jdouble* buf);
};

#endif
