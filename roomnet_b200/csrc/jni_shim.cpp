// JNI shim over the C ABI (include/roomnet.h) for the mobile module's classifier
// entry point.  It stands in for the TFLite interpreter call inside
//   Classifier.runInference()  ->  tflite.run(imgData, labelProbArray)
//   (mobile/tf_image_classifier/.../tflite/ClassifierFloatMobileNet.java:96-98,
//    ClassifierQuantizedMobileNet.java:93-95), with the model opened in the
//   Classifier constructor (Classifier.java:175-200) and freed in close() (:291-301).
//
// Java side:
//   package org.tensorflow.lite.examples.classification.tflite;
//   final class RoomNetNative {
//     static native long create(String checkpointPrefix, int device, int imSide, int precision);
//     static native int run(long handle, java.nio.ByteBuffer imgData, float[][] labelProbArray);
//     static native int runArgb(long handle, int[] intValues, float[][] labelProbArray);
//     static native int runYuv(long handle, ByteBuffer y, ByteBuffer u, ByteBuffer v, int width, int height,
//                              int yRowStride, int uvRowStride, int uvPixelStride, int rotation, float[][] labelProbArray);
//     static native void close(long handle);
//   }
#include <cstring>
#include <string>

#include "jni_min.h"
#include "roomnet.h"

namespace {

template <typename Fn>
Fn Slot(JNIEnv* env, int idx) {
  return reinterpret_cast<Fn>((*env)->slot[idx]);
}

void Throw(JNIEnv* env, const char* cls_name, const std::string& msg) {
  jclass cls = Slot<JniFindClassFn>(env, kJniFindClass)(env, cls_name);
  if (cls) {
    Slot<JniThrowNewFn>(env, kJniThrowNew)(env, cls, msg.c_str());
    Slot<JniDeleteLocalRefFn>(env, kJniDeleteLocalRef)(env, cls);
  }
}

struct JniModel {
  rn_handle* h = nullptr;
  int im_side = 0;
  int num_classes = 6;
};

}  // namespace

extern "C" {

#define RN_JNI(name) Java_org_tensorflow_lite_examples_classification_tflite_RoomNetNative_##name

// Classifier(Activity, Device, int numThreads) loads the model file and throws IOException when
// that fails (Classifier.java:175-200); same contract here.
jlong RN_JNI(create)(JNIEnv* env, jclass, jstring checkpoint_prefix, jint device, jint im_side, jint precision) {
  if (!checkpoint_prefix) {
    Throw(env, "java/lang/NullPointerException", "checkpointPrefix");
    return 0;
  }
  const char* prefix = Slot<JniGetStringUTFCharsFn>(env, kJniGetStringUTFChars)(env, checkpoint_prefix, nullptr);
  if (!prefix) return 0;  // OutOfMemoryError already pending
  rn_config cfg;
  std::memset(&cfg, 0, sizeof(cfg));
  cfg.abi_version = RN_ABI_VERSION;
  cfg.im_side = im_side;
  cfg.num_classes = 6;
  cfg.precision = precision;
  cfg.n_devices = 1;
  cfg.devices[0] = device;
  cfg.max_batch = 1;  // the camera loop classifies one frame at a time (ClassifierActivity.java:103-135)
  JniModel* m = new JniModel;
  m->im_side = im_side;
  int rc = rn_create(&cfg, &m->h);
  std::string err;
  if (rc != RN_OK) {
    err = rn_last_error(nullptr);
  } else if ((rc = rn_load_tf_checkpoint(m->h, prefix)) != RN_OK) {
    err = rn_last_error(m->h);
    rn_destroy(m->h);
  }
  Slot<JniReleaseStringUTFCharsFn>(env, kJniReleaseStringUTFChars)(env, checkpoint_prefix, prefix);
  if (rc != RN_OK) {
    delete m;
    Throw(env, "java/io/IOException", "RoomNet: " + err);
    return 0;
  }
  return reinterpret_cast<jlong>(m);
}

// tflite.run(imgData, labelProbArray): imgData is the direct ByteBuffer filled by
// convertBitmapToByteBuffer (Classifier.java:226-243): 1 x S x S x 3 in R,G,B order, either
// float32 (p-127.5)/127.5 (ClassifierFloatMobileNet.java:74-78) or raw bytes
// (ClassifierQuantizedMobileNet.java:71-75).  The element type is inferred from the capacity.
jint RN_JNI(run)(JNIEnv* env, jclass, jlong handle, jobject img_data, jobjectArray label_prob_array) {
  JniModel* m = reinterpret_cast<JniModel*>(handle);
  if (!m || !m->h) {
    Throw(env, "java/lang/IllegalStateException", "RoomNet: classifier is closed");
    return RN_ERR_INVALID_ARG;
  }
  void* data = img_data ? Slot<JniGetDirectBufferAddressFn>(env, kJniGetDirectBufferAddress)(env, img_data) : nullptr;
  if (!data) {
    Throw(env, "java/lang/IllegalArgumentException", "RoomNet: imgData must be a direct ByteBuffer");
    return RN_ERR_INVALID_ARG;
  }
  const jlong cap = Slot<JniGetDirectBufferCapacityFn>(env, kJniGetDirectBufferCapacity)(env, img_data);
  const jlong px = static_cast<jlong>(m->im_side) * m->im_side * 3;
  if (!label_prob_array || Slot<JniGetArrayLengthFn>(env, kJniGetArrayLength)(env, label_prob_array) < 1) {
    Throw(env, "java/lang/IllegalArgumentException", "RoomNet: labelProbArray must be float[1][numLabels]");
    return RN_ERR_INVALID_ARG;
  }
  jobject row = Slot<JniGetObjectArrayElementFn>(env, kJniGetObjectArrayElement)(env, label_prob_array, 0);
  if (!row || Slot<JniGetArrayLengthFn>(env, kJniGetArrayLength)(env, row) < m->num_classes) {
    if (row) Slot<JniDeleteLocalRefFn>(env, kJniDeleteLocalRef)(env, row);
    Throw(env, "java/lang/IllegalArgumentException", "RoomNet: labelProbArray[0] is shorter than the label count");
    return RN_ERR_INVALID_ARG;
  }
  float probs[32];
  int rc;
  if (cap == px * 4) {
    rc = rn_infer_f32_rgb(m->h, static_cast<const float*>(data), 1, nullptr, probs, nullptr);
  } else if (cap == px) {
    rc = rn_infer_u8_rgb(m->h, static_cast<const uint8_t*>(data), 1, nullptr, probs, nullptr);
  } else {
    rc = RN_ERR_INVALID_ARG;
    Throw(env, "java/lang/IllegalArgumentException",
          "RoomNet: imgData capacity " + std::to_string(cap) + " is neither float32 nor uint8 1xSxSx3");
  }
  if (rc == RN_OK)
    Slot<JniSetFloatArrayRegionFn>(env, kJniSetFloatArrayRegion)(env, row, 0, m->num_classes, probs);
  else if (cap == px * 4 || cap == px)
    Throw(env, "java/lang/IllegalStateException", std::string("RoomNet: ") + rn_last_error(m->h));
  Slot<JniDeleteLocalRefFn>(env, kJniDeleteLocalRef)(env, row);  // the row reference is released on every path
  return rc;
}

// The demo's Bitmap -> ByteBuffer conversion folded into the native call: `int_values` is the int[] that
// convertBitmapToByteBuffer fills with bitmap.getPixels(...) (Classifier.java:226-231), S*S 0xAARRGGBB values; the
// per-pixel addPixelValue loop (:233-240, ClassifierFloatMobileNet.java:74-78) runs on the device instead.
jint RN_JNI(runArgb)(JNIEnv* env, jclass, jlong handle, jintArray int_values, jobjectArray label_prob_array) {
  JniModel* m = reinterpret_cast<JniModel*>(handle);
  if (!m || !m->h) {
    Throw(env, "java/lang/IllegalStateException", "RoomNet: classifier is closed");
    return RN_ERR_INVALID_ARG;
  }
  const jlong px = static_cast<jlong>(m->im_side) * m->im_side;
  if (!int_values || Slot<JniGetArrayLengthFn>(env, kJniGetArrayLength)(env, int_values) != px) {
    Throw(env, "java/lang/IllegalArgumentException", "RoomNet: intValues must hold imageSizeX * imageSizeY pixels");
    return RN_ERR_INVALID_ARG;
  }
  if (!label_prob_array || Slot<JniGetArrayLengthFn>(env, kJniGetArrayLength)(env, label_prob_array) < 1) {
    Throw(env, "java/lang/IllegalArgumentException", "RoomNet: labelProbArray must be float[1][numLabels]");
    return RN_ERR_INVALID_ARG;
  }
  jobject row = Slot<JniGetObjectArrayElementFn>(env, kJniGetObjectArrayElement)(env, label_prob_array, 0);
  if (!row || Slot<JniGetArrayLengthFn>(env, kJniGetArrayLength)(env, row) < m->num_classes) {
    if (row) Slot<JniDeleteLocalRefFn>(env, kJniDeleteLocalRef)(env, row);
    Throw(env, "java/lang/IllegalArgumentException", "RoomNet: labelProbArray[0] is shorter than the label count");
    return RN_ERR_INVALID_ARG;
  }
  jint* pixels = Slot<JniGetIntArrayElementsFn>(env, kJniGetIntArrayElements)(env, int_values, nullptr);
  if (!pixels) {  // OutOfMemoryError already pending
    Slot<JniDeleteLocalRefFn>(env, kJniDeleteLocalRef)(env, row);
    return RN_ERR_INTERNAL;
  }
  float probs[32];
  const int rc = rn_infer_argb8888(m->h, pixels, 1, nullptr, probs, nullptr);
  Slot<JniReleaseIntArrayElementsFn>(env, kJniReleaseIntArrayElements)(env, int_values, pixels, 2 /* JNI_ABORT */);
  if (rc == RN_OK)
    Slot<JniSetFloatArrayRegionFn>(env, kJniSetFloatArrayRegion)(env, row, 0, m->num_classes, probs);
  else
    Throw(env, "java/lang/IllegalStateException", std::string("RoomNet: ") + rn_last_error(m->h));
  Slot<JniDeleteLocalRefFn>(env, kJniDeleteLocalRef)(env, row);
  return rc;
}

// The camera front end folded into the native call: the three planes of the android.media.Image the camera hands
// to CameraActivity.onImageAvailable (direct ByteBuffers, Image.Plane.getBuffer()) with their strides and the
// sensor-to-screen rotation of ClassifierActivity.java:80.  Replaces ImageUtils.convertYUV420ToARGB8888
// (ImageUtils.java:131-151), rgbFrameBitmap.setPixels + canvas.drawBitmap(..., frameToCropTransform, ...)
// (ClassifierActivity.java:101-103) and convertBitmapToByteBuffer (Classifier.java:226-243).
jint RN_JNI(runYuv)(JNIEnv* env, jclass, jlong handle, jobject y_buf, jobject u_buf, jobject v_buf, jint width,
                    jint height, jint y_row_stride, jint uv_row_stride, jint uv_pixel_stride, jint rotation,
                    jobjectArray label_prob_array) {
  JniModel* m = reinterpret_cast<JniModel*>(handle);
  if (!m || !m->h) {
    Throw(env, "java/lang/IllegalStateException", "RoomNet: classifier is closed");
    return RN_ERR_INVALID_ARG;
  }
  jobject bufs[3] = {y_buf, u_buf, v_buf};
  void* ptr[3] = {nullptr, nullptr, nullptr};
  jlong cap[3] = {0, 0, 0};
  for (int k = 0; k < 3; ++k) {
    ptr[k] = bufs[k] ? Slot<JniGetDirectBufferAddressFn>(env, kJniGetDirectBufferAddress)(env, bufs[k]) : nullptr;
    if (!ptr[k]) {
      Throw(env, "java/lang/IllegalArgumentException", "RoomNet: the image planes must be direct ByteBuffers");
      return RN_ERR_INVALID_ARG;
    }
    cap[k] = Slot<JniGetDirectBufferCapacityFn>(env, kJniGetDirectBufferCapacity)(env, bufs[k]);
    if (cap[k] <= 0 || cap[k] > 0x7fffffff) {
      Throw(env, "java/lang/IllegalArgumentException", "RoomNet: bad plane capacity");
      return RN_ERR_INVALID_ARG;
    }
  }
  if (!label_prob_array || Slot<JniGetArrayLengthFn>(env, kJniGetArrayLength)(env, label_prob_array) < 1) {
    Throw(env, "java/lang/IllegalArgumentException", "RoomNet: labelProbArray must be float[1][numLabels]");
    return RN_ERR_INVALID_ARG;
  }
  jobject row = Slot<JniGetObjectArrayElementFn>(env, kJniGetObjectArrayElement)(env, label_prob_array, 0);
  if (!row || Slot<JniGetArrayLengthFn>(env, kJniGetArrayLength)(env, row) < m->num_classes) {
    if (row) Slot<JniDeleteLocalRefFn>(env, kJniDeleteLocalRef)(env, row);
    Throw(env, "java/lang/IllegalArgumentException", "RoomNet: labelProbArray[0] is shorter than the label count");
    return RN_ERR_INVALID_ARG;
  }
  float probs[32];
  const int rc = rn_infer_yuv420(m->h, static_cast<const uint8_t*>(ptr[0]), static_cast<const uint8_t*>(ptr[1]),
                                 static_cast<const uint8_t*>(ptr[2]), static_cast<int32_t>(cap[0]),
                                 static_cast<int32_t>(cap[1]), static_cast<int32_t>(cap[2]), width, height, y_row_stride,
                                 uv_row_stride, uv_pixel_stride, rotation, nullptr, probs, nullptr, nullptr);
  if (rc == RN_OK)
    Slot<JniSetFloatArrayRegionFn>(env, kJniSetFloatArrayRegion)(env, row, 0, m->num_classes, probs);
  else
    Throw(env, rc == RN_ERR_INVALID_ARG ? "java/lang/IllegalArgumentException" : "java/lang/IllegalStateException",
          std::string("RoomNet: ") + rn_last_error(m->h));
  Slot<JniDeleteLocalRefFn>(env, kJniDeleteLocalRef)(env, row);
  return rc;
}

// Classifier.close() (Classifier.java:291-301)
void RN_JNI(close)(JNIEnv*, jclass, jlong handle) {
  JniModel* m = reinterpret_cast<JniModel*>(handle);
  if (!m) return;
  if (m->h) rn_destroy(m->h);
  delete m;
}

// Latency statistics of the handle (p50/p99 of run() in ms), for the demo's per-frame latency read-out
// (ClassifierActivity.java:113-115,128).
jint RN_JNI(stats)(JNIEnv* env, jclass, jlong handle, jfloatArray out2) {
  JniModel* m = reinterpret_cast<JniModel*>(handle);
  if (!m || !m->h || !out2) return RN_ERR_INVALID_ARG;
  double p50 = 0, p99 = 0;
  int rc = rn_get_stats(m->h, &p50, &p99, nullptr, nullptr);
  float v[2] = {static_cast<float>(p50), static_cast<float>(p99)};
  Slot<JniSetFloatArrayRegionFn>(env, kJniSetFloatArrayRegion)(env, out2, 0, 2, v);
  return rc;
}

}  // extern "C"
