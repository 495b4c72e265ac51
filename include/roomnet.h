/*
 * libroomnet — C ABI of the B200-native RoomNet inference path.
 *
 * This header is the drop-in boundary.  Every entry point replaces one piece of
 * the reference's Python→TensorFlow (or Java→TFLite) seam; the reference
 * interface each one stands in for is cited as  <file>:<line>  relative to the
 * upstream repository (ironhide23586/RoomNet).
 *
 * Conventions: extern "C", opaque handle, int status (0 = RN_OK), no exceptions
 * cross the boundary, all buffers are caller-owned HOST memory unless the name
 * ends in _device.  A handle may be used from several threads; calls on one
 * handle are serialised internally (tf.Session.run is thread-safe too, but the
 * reference only ever calls it from one thread: infer.py:79-82).
 * There is NO CPU fallback: every inference entry point fails with
 * RN_ERR_CUDA when no sm_100 device is usable.
 */
#ifndef ROOMNET_H_
#define ROOMNET_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define RN_ABI_VERSION 2
#define RN_MAX_DEVICES 16

typedef struct rn_handle rn_handle;

enum rn_status {
  RN_OK = 0,
  RN_ERR_INVALID_ARG = 1,   /* bad pointer / size / enum (TF: InvalidArgumentError on a shape mismatch) */
  RN_ERR_IO = 2,            /* checkpoint file missing or unreadable (TF: NotFoundError in Saver.restore) */
  RN_ERR_FORMAT = 3,        /* checkpoint corrupt, CRC mismatch, tensor missing or wrong shape */
  RN_ERR_NOT_LOADED = 4,    /* inference before weights were loaded (TF: FailedPrecondition, uninitialised variable) */
  RN_ERR_CUDA = 5,          /* CUDA runtime / driver error, or no usable device */
  RN_ERR_INTERNAL = 6
};

enum rn_precision {
  RN_PREC_FP32 = 0,  /* fp32 FMA on CUDA cores end to end; budget: max|dlogit| <= 1e-3 */
  RN_PREC_FP16 = 1,  /* 16-bit tensor-core path: fp16 operands, fp32 accumulate (tcgen05 kind::f16); budget 2e-2 */
  RN_PREC_BF16 = 2,  /* same kernels with bf16 operands; measured to MISS the 2e-2 budget on flat images */
  RN_PREC_FP32_TC = 3, /* fp32-class path on the tensor cores: activations and weights as hi + lo fp16 pairs, three
                         products per MAC (xh*wh + xl*wh + xh*wl) into fp32 accumulators, fp32 epilogues; budget 1e-3 */
  RN_PREC_BF16X3 = 4  /* BF16 operands that DO meet the 2e-2 budget: the three-product scheme of RN_PREC_FP32_TC with bf16
                         halves (hi + lo = 16 mantissa bits per operand), a third of the single-product rate */
};

enum rn_flags {
  /* Run the 16-bit path layer by layer (one kernel per conv layer, every activation written to HBM) instead of
   * the fused residual-block kernel.  Same arithmetic; exists so that rn_debug_activation can return the
   * intermediate tensors of a fused block (tf.Session.run can fetch any graph node, network.py:131) and for A/B
   * timing.  Not the benchmarked configuration. */
  RN_FLAG_LAYERWISE = 1,
  RN_FLAG_JPEG_HOST_HUFFMAN = 2  /* rn_infer_jpeg: Huffman decoding on host threads for every file (default: on the device) */
};

/* Replaces the constructor arguments of the reference model object:
 *   RoomNet(num_classes, im_side, ..., optimized_inference=True)   network.py:21-48
 * plus what TensorFlow decided implicitly (device placement, network.py:89). */
typedef struct rn_config {
  int32_t abi_version;              /* RN_ABI_VERSION */
  int32_t im_side;                  /* network.py:21  (224 for the shipped checkpoint, infer.py:26) */
  int32_t num_classes;              /* network.py:21  (6, infer.py:22) */
  int32_t precision;                /* enum rn_precision */
  int32_t n_devices;                /* >=1; images of one call are split contiguously over the replicas.
                                       0 = host-only handle (load + fold only; inference returns RN_ERR_CUDA) */
  int32_t devices[RN_MAX_DEVICES];  /* CUDA ordinals */
  int32_t max_batch;                /* per-replica micro-batch held resident on the device (0 = default) */
  int32_t flags;                    /* bit set of rn_flags (0 = default) */
} rn_config;

/* network.py:21-48 (graph construction) + network.py:87-91 (session creation). */
int rn_create(const rn_config* cfg, rn_handle** out);

/* Classifier.close()                                    mobile/.../tflite/Classifier.java:291-301
 * (tf.Session teardown on the Python side; the reference never closes its session). */
int rn_destroy(rn_handle* h);

/* Saver.restore(sess, model_path)                      network.py:122
 * Reads <prefix>.index / <prefix>.data-00000-of-00001 (TF V2 bundle), verifies the
 * per-tensor CRC32C, folds every frozen BatchNorm forward into the next
 * conv/dense layer (fp64), packs and uploads the result to every replica. */
int rn_load_tf_checkpoint(rn_handle* h, const char* prefix);

/* Same restore from caller-held arrays (variable name → fp32 data), for callers
 * that already hold the variables (network.py:46 vars_to_keep).  shapes[i] has
 * ranks[i] entries. */
int rn_load_tensors(rn_handle* h, int32_t n, const char* const* names, const float* const* data,
                    const int64_t* const* shapes, const int32_t* ranks);

/* dense/kernel override for im_side != 224: the shipped dense/kernel is [64,32]
 * and only fits im_side 224 (network.py:231-234).  Call before rn_load_*. */
int rn_set_dense0(rn_handle* h, const float* kernel /* [flat_len,32] */, int32_t flat_len);

/* RoomNet.infer(im_in)                                  network.py:128-135
 * n pre-sized images, NHWC uint8 BGR [n, S, S, 3].  Normalisation
 * ((x[..., ::-1]/255)*2-1, network.py:129) happens on the device.
 * Outputs (any may be NULL): top1 int64[n] (= tf.argmax of the softmax,
 * network.py:45, first maximum wins), probs f32[n,C] (network.py:44),
 * logits f32[n,C] (= out_op, ReLU6-clipped, network.py:43/214). */
int rn_infer_u8_bgr(rn_handle* h, const uint8_t* nhwc, int32_t n, int64_t* top1, float* probs, float* logits);

/* Asynchronous form of rn_infer_u8_bgr for callers that stream batches (classify_im_dir, infer.py:65-100, feeds
 * one image after the other; a batch scheduler on this ABI keeps two or three calls in flight).  rn_submit_u8_bgr
 * enqueues the host->device copies and the kernels of one call (the last kernel stores the results into mapped host
 * memory) and returns a ticket; it blocks only while all eight staging slots of a replica (one per micro-batch of at
 * most max_batch images) are still busy.  The output buffers (and, when `nhwc` is page-locked
 * memory, the input) must stay valid until rn_wait(h, ticket) has returned: results are written by a later
 * rn_submit_u8_bgr / rn_wait on the calling thread.  rn_wait(h, 0) waits for everything submitted so far; the
 * synchronous entry points wait for earlier submissions first.  Same arithmetic, same results as rn_infer_u8_bgr. */
int rn_submit_u8_bgr(rn_handle* h, const uint8_t* nhwc, int32_t n, int64_t* top1, float* probs, float* logits,
                     uint64_t* ticket);
int rn_wait(rn_handle* h, uint64_t ticket);

/* sess.run(outs_final, {x_tensor: im})                  network.py:131-134, :155
 * and Interpreter.run(imgData, labelProbArray)          ClassifierFloatMobileNet.java:97
 * Raw feed: NHWC float32 RGB in [-1,1], [n, S, S, 3]. */
int rn_infer_f32_rgb(rn_handle* h, const float* nhwc, int32_t n, int64_t* top1, float* probs, float* logits);

/* Interpreter.run on the quantised demo path            ClassifierQuantizedMobileNet.java:94
 * NHWC uint8 RGB (Bitmap channel order, Classifier.java:226-243). */
int rn_infer_u8_rgb(rn_handle* h, const uint8_t* nhwc, int32_t n, int64_t* top1, float* probs, float* logits);

/* Classifier.convertBitmapToByteBuffer                   Classifier.java:226-243
 * bitmap.getPixels(intValues, ...) followed by the per-pixel addPixelValue loop (ClassifierFloatMobileNet.java:74-78:
 * ((v >> 16) & 0xFF, (v >> 8) & 0xFF, v & 0xFF) -> (p - 127.5) / 127.5) - the 50 k-iteration Java loop moves into the
 * library: `pixels` is the int[] of Bitmap.getPixels, [n, S, S] 0xAARRGGBB values (alpha ignored); the bytes of
 * each int are B,G,R,A in memory, so this is rn_infer_u8_bgr with a 4-byte pixel pitch (bit-identical results). */
int rn_infer_argb8888(rn_handle* h, const int32_t* pixels, int32_t n, int64_t* top1, float* probs, float* logits);

/* Device-resident variant of rn_infer_u8_bgr on replica 0: d_nhwc, d_top1 (int64),
 * d_probs, d_logits are device pointers on devices[0]; work is enqueued on
 * `cuda_stream` (a cudaStream_t, NULL = the replica's own stream) and NOT
 * synchronised.  Used to time the kernels without PCIe in the loop. */
int rn_infer_u8_bgr_device(rn_handle* h, const void* d_nhwc, int32_t n, void* d_top1, void* d_probs,
                           void* d_logits, void* cuda_stream);

/* RoomNet.center_crop + cv2.resize(im, (S, S))          network.py:137-146, :149-152
 * One BGR/RGB uint8 image [H, W, 3] of any size -> [S, S, 3] on the device: centre crop (with the reference's
 * floor-division offset) + OpenCV's INTER_LINEAR uint8 fixed-point bilinear (11-bit coefficients; exact 2x shrink
 * = 2x2 box), bit-identical to cv2.resize.  `out` receives S*S*3 bytes. */
int rn_preprocess_u8(rn_handle* h, const uint8_t* img, int32_t H, int32_t W, uint8_t* out);

/* RoomNet.infer_optimized(im_in)                        network.py:148-156
 * The whole per-image call of infer.py:81-82 on the device: rn_preprocess_u8 + rn_infer_u8_bgr for one image. */
int rn_infer_image_u8_bgr(rn_handle* h, const uint8_t* img, int32_t H, int32_t W, int64_t* top1, float* probs,
                          float* logits);

/* The loop of classify_im_dir                            infer.py:79-82
 * n BGR uint8 photos of arbitrary sizes (imgs[i] is heights[i] x widths[i] x 3, packed rows): the centre crop and the
 * cv2-identical resize of a whole micro-batch run as ONE kernel that writes the network's input tensor, followed by the
 * forward pass - nothing returns to the host between preprocessing and inference.  The list is split over the handle's
 * devices like any other batch.  Per image the result is bit-identical to rn_infer_image_u8_bgr. */
int rn_infer_images_u8_bgr(rn_handle* h, const uint8_t* const* imgs, const int32_t* heights, const int32_t* widths,
                           int32_t n, int64_t* top1, float* probs, float* logits);

/* cv2.imread(fpath) + RoomNet.infer_optimized(im)       infer.py:81-82
 * n files as they sit on disk (files[i] = sizes[i] encoded bytes).  For baseline JPEG files (8-bit, Huffman, grey or
 * YCbCr with 4:4:4 / 4:2:2 / 4:2:0 sampling, any EXIF orientation) the whole decoder runs on the device: the host only
 * strips the byte stuffing while it gathers the files (on `threads` threads, 0 = as many as the machine has, at most
 * 16); Huffman decoding (self-synchronising parallel decode; files with several scans fall back to host threads),
 * dequantisation, inverse DCT, chroma upsampling, colour conversion and EXIF orientation are CUDA kernels with the
 * integer arithmetic of the decoder cv2 links
 * (libjpeg-turbo defaults), so the decoded image - and everything after it - is bit-identical to cv2.imread's, and the
 * decoded photo never exists in host memory.  status[i] (required) receives RN_JPEG_OK, RN_JPEG_UNSUPPORTED
 * (progressive, arithmetic, CMYK, 12-bit, other samplings, not a JPEG at all) or RN_JPEG_CORRUPT (damaged or truncated
 * stream); the outputs of entries that are not RN_JPEG_OK are left untouched - decode those with cv2.imread and pass
 * them to rn_infer_images_u8_bgr, which is what roomnet_b200.network.RoomNet.infer_files does. */
enum rn_jpeg_status { RN_JPEG_OK = 0, RN_JPEG_UNSUPPORTED = 1, RN_JPEG_CORRUPT = 2 };
int rn_infer_jpeg(rn_handle* h, const uint8_t* const* files, const uint64_t* sizes, int32_t n, int32_t threads,
                  int64_t* top1, float* probs, float* logits, int32_t* status);

/* cv2.imread(fpath) alone                                infer.py:81
 * Decodes one file on the device and returns the BGR image (height x width x 3, EXIF orientation applied).  With
 * out == NULL only height / width / status are filled in.  *status as for rn_infer_jpeg. */
int rn_decode_jpeg_u8_bgr(rn_handle* h, const uint8_t* file, uint64_t size, uint8_t* out, uint64_t out_capacity,
                          int32_t* height, int32_t* width, int32_t* status);

/* Host-only helpers of the JPEG front end (no device needed).  rn_jpeg_info: info = {status, width, height,
 * components, luma h sampling, luma v sampling, EXIF orientation, number of int16 coefficients}.
 * rn_jpeg_coefficients: the entropy-decoded, still quantised DCT coefficients, per component [block rows][block
 * columns][64] in natural order, components back to back (what the device kernels consume); returns the rn_jpeg_status. */
/* How many files of rn_infer_jpeg / rn_decode_jpeg_u8_bgr had their Huffman decoding done on the device, and how many
 * on host threads (several scans, damaged streams, RN_FLAG_JPEG_HOST_HUFFMAN), since the handle was created. */
int rn_get_jpeg_counters(rn_handle* h, int64_t* device_huffman_files, int64_t* host_huffman_files);
int rn_jpeg_info(const uint8_t* file, uint64_t size, int64_t info[8]);
/* Host-only: what the host hands to the device Huffman decoder for this file - the scan with the byte stuffing and the
 * restart markers removed, every restart segment padded to a multiple of 128 bytes (stream), and the restart segment of
 * every 128-byte subsequence (sub_seg, capacity / 128 entries).  info = {stream bytes, restart segments, blocks per MCU,
 * total blocks}.  Returns the rn_jpeg_status (RN_JPEG_UNSUPPORTED = this file takes the host Huffman decoder). */
int rn_jpeg_prepare_scan(const uint8_t* file, uint64_t size, uint8_t* stream, uint64_t capacity, int32_t* sub_seg,
                         int64_t info[4]);
int rn_jpeg_coefficients(const uint8_t* file, uint64_t size, int16_t* coefs, uint64_t capacity);

/* CameraActivity.onImageAvailable -> ImageUtils.convertYUV420ToARGB8888 -> ClassifierActivity.processImage
 *   mobile/.../env/ImageUtils.java:131-151 (+ YUV2RGB :100-129), getTransformationMatrix :168-225,
 *   ClassifierActivity.java:89-106 (frameToCropTransform, canvas.drawBitmap), Classifier.java:226-243
 * One YUV_420_888 camera frame (the three android.media.Image planes with their strides) -> class probabilities: the
 * colour conversion (the reference's integer arithmetic, bit-exact), the frame-to-crop transform (rotation by a
 * multiple of 90 degrees, uniform scale max(S/w, S/h), nearest frame pixel as an unfiltered drawBitmap samples it) and
 * the network run in one device call; `rgb_out` (optional) receives the S*S*3 R,G,B bytes the network saw (what
 * croppedBitmap would hold).  Same results as rn_infer_u8_rgb on those bytes. */
int rn_infer_yuv420(rn_handle* h, const uint8_t* y, const uint8_t* u, const uint8_t* v, int32_t y_size, int32_t u_size,
                    int32_t v_size, int32_t width, int32_t height, int32_t y_row_stride, int32_t uv_row_stride,
                    int32_t uv_pixel_stride, int32_t rotation_degrees, int64_t* top1, float* probs, float* logits,
                    uint8_t* rgb_out);

/* Host-side geometry helper: writes the crop rectangle the reference would take (network.py:137-146). */
int rn_center_crop_rect(int32_t h, int32_t w, int32_t* y0, int32_t* x0, int32_t* side);

/* Introspection used by the parity tests (no reference analogue). */
int rn_flat_len(const rn_handle* h);                       /* network.py:231-232 */
int rn_num_kernel_launches(const rn_handle* h);            /* launches enqueued by the last inference call */
int rn_get_folded(rn_handle* h, const char* name, float* out, int64_t capacity, int64_t* size);
int rn_debug_activation(rn_handle* h, int32_t layer, float* out, int64_t capacity, int64_t* size, int32_t dims[4]);
int rn_get_stats(rn_handle* h, double* p50_ms, double* p99_ms, int64_t* calls, int64_t* images);
int rn_reset_stats(rn_handle* h);
/* Per-kernel device times of replica 0 (CUDA events recorded between launches on the launching
 * stream) accumulated since profiling was enabled / last read; used by bench.py for the roofline. */
int rn_set_profiling(rn_handle* h, int32_t enabled);
int rn_get_profile(rn_handle* h, int32_t capacity, int32_t* count, char names[][32], double* ms, int32_t* launches);

/* Error text of the last failing call on this handle (or of rn_create when h is NULL). */
const char* rn_last_error(const rn_handle* h);
const char* rn_version(void);

#ifdef __cplusplus
}
#endif
#endif /* ROOMNET_H_ */
