"""Drop-in ``RoomNet`` for the reference's inference surface (reference network.py).

Same constructor keywords, same ``load`` / ``init`` / ``infer`` /
``infer_optimized`` / ``center_crop`` methods and return types as the reference
class (network.py:19-156); everything that used to be a ``tf.Session.run``
(network.py:131-134, :155) is one call into libroomnet.so through ctypes.
Training-only members (train_step, save, the loss/optimizer graph,
network.py:49-85, :93-103, :158-170) are out of scope and raise.

Extra, behaviour-preserving keywords: ``precision`` ('fp16' tensor-core path |
'fp32' | 'bf16'), ``devices`` (CUDA ordinals, images of a batch are split over
them), ``max_batch``, ``gpu_preprocess`` (crop + resize on the device, bit-identical to cv2).
"""
from __future__ import annotations

import numpy as np

from . import _capi

DEFAULT_MODEL_PATH = './final_model/roomnet'  # reference infer.py:24


class RoomNet:

    def __init__(self, num_classes, im_side=600, compute_bn_mean_var=True, start_step=0, dropout_enabled=False,
                 learn_rate=1e-4, l2_regularizer_coeff=1e-2, num_steps=10000, dropout_rate=.2,
                 update_batchnorm_means_vars=True, optimized_inference=False,
                 precision='fp16', devices=(0,), max_batch=0, dense0_kernel=None, gpu_preprocess=False):
        if compute_bn_mean_var:
            # reference network.py:193: training=True would use batch statistics; the product
            # implements the frozen-statistics inference path only (infer.py:104 passes False).
            raise NotImplementedError("compute_bn_mean_var=True (batch-statistics BN) is a training mode; "
                                      "pass compute_bn_mean_var=False as infer.py:104 does")
        if dropout_enabled and not optimized_inference:
            raise NotImplementedError("dropout is a training feature (reference network.py:204-206)")
        self.num_classes = num_classes
        self.im_side = im_side
        self.compute_bn_mean_var = compute_bn_mean_var
        self.optimized_inference = optimized_inference
        self.dropout_enabled = False
        self.start_step = start_step
        self.step = start_step
        self.learn_rate = learn_rate
        self.precision = precision
        self.devices = tuple(devices)
        self.max_batch = max_batch
        self._dense0_kernel = dense0_kernel
        # center_crop + cv2.resize on the device (bit-identical to cv2 for uint8 images) instead of on the host
        self.gpu_preprocess = gpu_preprocess
        self.sess = None  # the libroomnet handle plays the role of tf.Session

    # reference network.py:87-91
    def init(self):
        if not self.sess:
            self.sess = _capi.Handle(im_side=self.im_side, num_classes=self.num_classes, precision=self.precision,
                                     devices=self.devices, max_batch=self.max_batch)
            if self._dense0_kernel is not None:
                self.sess.set_dense0(self._dense0_kernel)

    # reference network.py:105-126 (explicit-path restore; the "latest training checkpoint" search
    # of :109-118 belongs to the training driver)
    def load(self, model_path=None):
        if not self.sess:
            self.init()
        if model_path is None:
            print('No model found to restore from, initializing random weights')
            raise NotImplementedError("random-weight initialisation is a training feature; pass model_path")
        self.sess.load_tf_checkpoint(model_path)
        print('Model restored from', model_path)

    def _require(self):
        if not self.sess:
            raise _capi.RoomNetError(_capi.RN_ERR_NOT_LOADED, "call load() first")

    # reference network.py:128-135 — batch of pre-sized BGR uint8 images [N,S,S,3]
    def infer(self, im_in):
        self._require()
        im_in = np.asarray(im_in)
        if im_in.dtype == np.uint8:
            top1, probs = self.sess.infer_u8_bgr(im_in)
        else:
            # non-uint8 input: same arithmetic as the reference line 129, then the raw float feed
            im = ((im_in[:, :, :, [2, 1, 0]] / 255.) * 2) - 1
            top1, probs = self.sess.infer_f32_rgb(im.astype(np.float32))
        if self.optimized_inference:
            return top1, probs   # outs_final = (argmax, softmax)   network.py:45
        return top1              # outs_final = argmax              network.py:72

    # reference network.py:137-146
    def center_crop(self, x):
        h, w, _ = x.shape
        offset = abs((w - h) // 2)
        if h < w:
            x_pp = x[:, offset:offset + h, :]
        elif w < h:
            x_pp = x[offset:offset + w, :, :]
        else:
            x_pp = x.copy()
        return x_pp

    def preprocess(self, im_in):
        """center_crop + cv2.resize exactly as reference network.py:149-152 (host side)."""
        im = self.center_crop(im_in)
        h, w, _ = im.shape
        if h != self.im_side or w != self.im_side:
            import cv2
            im = cv2.resize(im, (self.im_side, self.im_side))
        return im

    # reference network.py:148-156 — one BGR uint8 image of any size
    def infer_optimized(self, im_in):
        self._require()
        if self.gpu_preprocess and im_in.dtype == np.uint8:
            return self.sess.infer_image_u8_bgr(im_in)
        im = self.preprocess(im_in)
        if im.dtype == np.uint8:
            out_label_idx, out_label_conf = self.sess.infer_u8_bgr(im[None])
        else:
            x = ((im[:, :, [2, 1, 0]] / 255.) * 2) - 1
            out_label_idx, out_label_conf = self.sess.infer_f32_rgb(np.expand_dims(x, 0).astype(np.float32))
        return out_label_idx, out_label_conf

    def infer_optimized_batch(self, ims):
        """Batched form of infer_optimized used by classify_im_dir: list of BGR images of any size."""
        self._require()
        if self.gpu_preprocess and all(im.dtype == np.uint8 for im in ims):
            # crop + resize of the whole list in one launch, straight into the network input (rn_infer_images_u8_bgr)
            return self.sess.infer_images_u8_bgr(ims)
        batch = np.stack([self.preprocess(im) for im in ims])
        return self.sess.infer_u8_bgr(batch)

    def infer_files(self, blobs, decode_threads=0):
        """cv2.imread + infer_optimized of reference infer.py:81-82 for a list of files given as their encoded bytes.

        Baseline JPEG files are decoded on the device (entropy decoding on host threads, inverse DCT / upsampling /
        colour conversion / EXIF orientation as CUDA kernels, bit-identical to cv2.imread), everything else (PNG,
        progressive JPEG, ...) through cv2.imdecode on the host and the batched photo call.  An undecodable file
        raises AttributeError like the reference does when cv2.imread returns None (network.py:138)."""
        self._require()
        import cv2
        top1, probs, status = self.sess.infer_jpeg(blobs, threads=decode_threads)
        rest = [i for i, st in enumerate(status) if st != 0]
        if rest:
            ims = [cv2.imdecode(np.frombuffer(blobs[i], dtype=np.uint8), cv2.IMREAD_COLOR) for i in rest]
            for i, im in zip(rest, ims):
                if im is None:
                    raise AttributeError("'NoneType' object has no attribute 'shape' (file %d of the list is not an "
                                         "image cv2 can read)" % i)
            t, p = self.infer_optimized_batch(ims)
            top1[rest] = t
            probs[rest] = p
        return top1, probs

    def train_step(self, x_in, y):
        raise NotImplementedError("training is out of scope of the inference hot path (reference network.py:158-170)")

    def save(self, suffix=None):
        raise NotImplementedError("checkpoint writing is out of scope (reference network.py:93-103)")

    def close(self):
        if self.sess:
            self.sess.close()
            self.sess = None
