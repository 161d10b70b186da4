/*
 * jni_min.h -- the part of the JNI ABI libhadoop-4mc.so needs, written from the JNI specification
 * (function-table slot numbers) because this build image has no JDK.  SURVEY.md Appendix F lists
 * the slots the shipped reference library calls through (6, 14, 23, 94, 95, 100, 101, 109, 110,
 * 167, 222, 223, 230); they are the only named members, everything else is padding that keeps the
 * table layout of a real JVM.  With a JDK present, compile jni_shim.c with -DFOURMC_REAL_JNI and
 * <jni.h> instead.
 */
#ifndef FOURMC_JNI_MIN_H
#define FOURMC_JNI_MIN_H

#include <stdint.h>

typedef int32_t jint;
typedef int64_t jlong;
typedef int8_t jbyte;
typedef uint8_t jboolean;
typedef void *jobject;
typedef jobject jclass;
typedef jobject jstring;
typedef jobject jarray;
typedef jarray jbyteArray;
typedef jobject jthrowable;
typedef struct fm_jfieldID_ *jfieldID;

struct JNINativeInterface_;
typedef const struct JNINativeInterface_ *JNIEnv;

struct JNINativeInterface_ {
    void *pad0[6];
    jclass (*FindClass)(JNIEnv *, const char *);                                  /* 6 */
    void *pad7[7];
    jint (*ThrowNew)(JNIEnv *, jclass, const char *);                             /* 14 */
    void *pad15[8];
    void (*DeleteLocalRef)(JNIEnv *, jobject);                                    /* 23 */
    void *pad24[70];
    jfieldID (*GetFieldID)(JNIEnv *, jclass, const char *, const char *);         /* 94 */
    jobject (*GetObjectField)(JNIEnv *, jobject, jfieldID);                       /* 95 */
    void *pad96[4];
    jint (*GetIntField)(JNIEnv *, jobject, jfieldID);                             /* 100 */
    jlong (*GetLongField)(JNIEnv *, jobject, jfieldID);                           /* 101 */
    void *pad102[7];
    void (*SetIntField)(JNIEnv *, jobject, jfieldID, jint);                       /* 109 */
    void (*SetLongField)(JNIEnv *, jobject, jfieldID, jlong);                     /* 110 */
    void *pad111[56];
    jstring (*NewStringUTF)(JNIEnv *, const char *);                              /* 167 */
    void *pad168[54];
    void *(*GetPrimitiveArrayCritical)(JNIEnv *, jarray, jboolean *);             /* 222 */
    void (*ReleasePrimitiveArrayCritical)(JNIEnv *, jarray, void *, jint);        /* 223 */
    void *pad224[6];
    void *(*GetDirectBufferAddress)(JNIEnv *, jobject);                           /* 230 */
    void *pad231[4];
};

#define JNIEXPORT __attribute__((visibility("default")))
#define JNICALL

#endif
