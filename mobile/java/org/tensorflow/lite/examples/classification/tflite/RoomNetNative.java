/*
 * JNI binding of roomnet_b200/csrc/jni_shim.cpp (libroomnet_jni.so -> libroomnet.so, include/roomnet.h).
 *
 * Drop this file and ClassifierRoomNet.java next to the reference's
 * mobile/tf_image_classifier/app/src/main/java/org/tensorflow/lite/examples/classification/tflite/Classifier.java.
 * The build image of this repository has no JDK: the sources are kept as files so that a maintainer can compile them
 * with the Android module; the native side is exercised by tests/jni_harness.cpp through a fake JNIEnv table.
 */
package org.tensorflow.lite.examples.classification.tflite;

import java.io.IOException;
import java.nio.ByteBuffer;

final class RoomNetNative {
  static {
    System.loadLibrary("roomnet_jni");
  }

  private RoomNetNative() {}

  /** Precision constants of rn_config.precision (include/roomnet.h). */
  static final int PRECISION_FP32 = 0;
  static final int PRECISION_FP16 = 1;
  static final int PRECISION_BF16 = 2;

  /** Opens the TF-V2 checkpoint `prefix` (.index / .data-00000-of-00001) on CUDA device `device`. */
  static native long create(String checkpointPrefix, int device, int imSide, int precision) throws IOException;

  /** tflite.run(imgData, labelProbArray): direct ByteBuffer, 1 x S x S x 3 float32 (p - 127.5) / 127.5 or raw RGB bytes. */
  static native int run(long handle, ByteBuffer imgData, float[][] labelProbArray);

  /** Bitmap.getPixels int[] (0xAARRGGBB), S * S values: the per-pixel conversion loop runs on the device. */
  static native int runArgb(long handle, int[] intValues, float[][] labelProbArray);

  /**
   * One YUV_420_888 camera frame: the three Image.Plane buffers (direct), their strides and the rotation of the
   * frame-to-crop transform. Colour conversion, crop/scale/rotate and the network run in one native call.
   */
  static native int runYuv(
      long handle,
      ByteBuffer yPlane,
      ByteBuffer uPlane,
      ByteBuffer vPlane,
      int width,
      int height,
      int yRowStride,
      int uvRowStride,
      int uvPixelStride,
      int rotationDegrees,
      float[][] labelProbArray);

  static native void close(long handle);

  /** p50 / p99 latency in milliseconds of the calls made through `handle`. */
  static native int stats(long handle, float[] p50p99Ms);
}
