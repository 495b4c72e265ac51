/*
 * Minimal, self-declared subset of the JNI ABI (JNI 1.6 function table layout).
 * No JDK / jni.h exists in this image, so the shim compiles against this header:
 * JNIEnv is a pointer to a pointer to a table of function pointers whose slot
 * indices are fixed by the JNI specification.  Only the slots the shim calls are
 * typed; all others are opaque padding so that the indices stay exact.
 */
#ifndef ROOMNET_JNI_MIN_H_
#define ROOMNET_JNI_MIN_H_
#include <stdint.h>

extern "C" {

typedef int32_t jint;
typedef int64_t jlong;
typedef float jfloat;
typedef uint8_t jboolean;
typedef jint jsize;
struct _jobject;
typedef _jobject* jobject;
typedef jobject jclass;
typedef jobject jstring;
typedef jobject jarray;
typedef jobject jobjectArray;
typedef jobject jfloatArray;
typedef jobject jintArray;
typedef jobject jthrowable;

struct JNINativeInterface_;
typedef const JNINativeInterface_* JNIEnv;

/* Slot numbers from the JNI specification ("Interface Function Table"). */
enum {
  kJniFindClass = 6,
  kJniThrowNew = 14,
  kJniDeleteLocalRef = 23,
  kJniGetStringUTFChars = 169,
  kJniReleaseStringUTFChars = 170,
  kJniGetArrayLength = 171,
  kJniGetObjectArrayElement = 173,
  kJniGetIntArrayElements = 187,
  kJniReleaseIntArrayElements = 195,
  kJniSetFloatArrayRegion = 213,
  kJniGetDirectBufferAddress = 230,
  kJniGetDirectBufferCapacity = 231,
  kJniTableSize = 232
};

struct JNINativeInterface_ {
  void* slot[kJniTableSize];
};

typedef jclass (*JniFindClassFn)(JNIEnv*, const char*);
typedef jint (*JniThrowNewFn)(JNIEnv*, jclass, const char*);
typedef void (*JniDeleteLocalRefFn)(JNIEnv*, jobject);
typedef const char* (*JniGetStringUTFCharsFn)(JNIEnv*, jstring, jboolean*);
typedef void (*JniReleaseStringUTFCharsFn)(JNIEnv*, jstring, const char*);
typedef jsize (*JniGetArrayLengthFn)(JNIEnv*, jarray);
typedef jobject (*JniGetObjectArrayElementFn)(JNIEnv*, jobjectArray, jsize);
typedef jint* (*JniGetIntArrayElementsFn)(JNIEnv*, jintArray, jboolean*);
typedef void (*JniReleaseIntArrayElementsFn)(JNIEnv*, jintArray, jint*, jint);
typedef void (*JniSetFloatArrayRegionFn)(JNIEnv*, jfloatArray, jsize, jsize, const jfloat*);
typedef void* (*JniGetDirectBufferAddressFn)(JNIEnv*, jobject);
typedef jlong (*JniGetDirectBufferCapacityFn)(JNIEnv*, jobject);

}  /* extern "C" */
#endif
