/*
 * RoomNet behind the demo's Classifier contract (reference Classifier.java:308-377 abstract hooks): the adaptor that
 * lets mobile/tf_image_classifier serve the six RoomNet classes instead of MobileNet. The interpreter call
 * tflite.run(imgData, labelProbArray) (ClassifierFloatMobileNet.java:96-98) becomes RoomNetNative.run.
 *
 * To select it, add a Model.ROOMNET case to Classifier.create (Classifier.java:88-96):
 *     if (model == Model.ROOMNET) return new ClassifierRoomNet(activity, device, numThreads);
 * and ship assets/roomnet_labels.txt with the six labels of infer.py:22, one per line:
 *     Backyard, Bathroom, Bedroom, Frontyard, Kitchen, LivingRoom
 */
package org.tensorflow.lite.examples.classification.tflite;

import android.app.Activity;
import android.graphics.Bitmap;
import android.media.Image;
import java.io.IOException;
import java.nio.ByteBuffer;

public class ClassifierRoomNet extends Classifier {
  /** Where the TF-V2 bundle of final_model/ lives on the device that owns the GPU. */
  public static final String CHECKPOINT_PREFIX = "/data/local/tmp/final_model/roomnet";

  private static final int IMAGE_SIDE = 224; // infer.py:26
  private static final int NUM_LABELS = 6; // infer.py:22

  private final long handle;
  private final float[][] labelProbArray = new float[1][NUM_LABELS];
  private final int[] pixels = new int[IMAGE_SIDE * IMAGE_SIDE];

  public ClassifierRoomNet(Activity activity, Device device, int numThreads) throws IOException {
    super(activity, device, numThreads);
    handle = RoomNetNative.create(CHECKPOINT_PREFIX, 0, IMAGE_SIDE, RoomNetNative.PRECISION_FP16);
  }

  @Override
  public int getImageSizeX() {
    return IMAGE_SIDE;
  }

  @Override
  public int getImageSizeY() {
    return IMAGE_SIDE;
  }

  @Override
  protected String getModelPath() {
    // the base class maps a .tflite asset; RoomNet's weights are read natively from CHECKPOINT_PREFIX. Any small
    // placeholder asset keeps the base constructor (Classifier.java:175-200) working unchanged.
    return "roomnet_placeholder.tflite";
  }

  @Override
  protected String getLabelPath() {
    return "roomnet_labels.txt";
  }

  @Override
  protected int getNumBytesPerChannel() {
    return 4; // float feed, as ClassifierFloatMobileNet
  }

  @Override
  protected void addPixelValue(int pixelValue) {
    // same arithmetic as ClassifierFloatMobileNet.java:74-78: (p - 127.5) / 127.5 in R, G, B order
    imgData.putFloat((((pixelValue >> 16) & 0xFF) - 127.5f) / 127.5f);
    imgData.putFloat((((pixelValue >> 8) & 0xFF) - 127.5f) / 127.5f);
    imgData.putFloat(((pixelValue & 0xFF) - 127.5f) / 127.5f);
  }

  @Override
  protected float getProbability(int labelIndex) {
    return labelProbArray[0][labelIndex];
  }

  @Override
  protected void setProbability(int labelIndex, Number value) {
    labelProbArray[0][labelIndex] = value.floatValue();
  }

  @Override
  protected float getNormalizedProbability(int labelIndex) {
    return labelProbArray[0][labelIndex];
  }

  @Override
  protected int getNumLabels() {
    return NUM_LABELS;
  }

  /** Was: tflite.run(imgData, labelProbArray) (ClassifierFloatMobileNet.java:96-98). */
  @Override
  protected void runInference() {
    RoomNetNative.run(handle, imgData, labelProbArray);
  }

  /**
   * Fast path for recognizeImage (Classifier.java:246-288): skips convertBitmapToByteBuffer (:226-243, a 50 k
   * iteration Java loop); the 0xAARRGGBB ints are unpacked and normalised by the first kernel.
   */
  public float[] classifyBitmap(Bitmap croppedBitmap) {
    croppedBitmap.getPixels(pixels, 0, IMAGE_SIDE, 0, 0, IMAGE_SIDE, IMAGE_SIDE);
    RoomNetNative.runArgb(handle, pixels, labelProbArray);
    return labelProbArray[0];
  }

  /**
   * Fast path for the camera loop (CameraActivity.onImageAvailable -> ClassifierActivity.processImage): the planes of
   * the YUV_420_888 frame go straight to the device; ImageUtils.convertYUV420ToARGB8888, rgbFrameBitmap.setPixels and
   * canvas.drawBitmap(rgbFrameBitmap, frameToCropTransform, null) are not needed for classification any more.
   */
  public float[] classifyFrame(Image image, int sensorOrientation) {
    final Image.Plane[] planes = image.getPlanes();
    final ByteBuffer y = planes[0].getBuffer();
    final ByteBuffer u = planes[1].getBuffer();
    final ByteBuffer v = planes[2].getBuffer();
    RoomNetNative.runYuv(
        handle,
        y,
        u,
        v,
        image.getWidth(),
        image.getHeight(),
        planes[0].getRowStride(),
        planes[1].getRowStride(),
        planes[1].getPixelStride(),
        sensorOrientation,
        labelProbArray);
    return labelProbArray[0];
  }

  @Override
  public void close() {
    RoomNetNative.close(handle);
    super.close();
  }
}
