"""Drop-in ``classify_im_dir`` over the B200 path.

Behavioural contract taken from the reference (infer.py:65-100): given a model
and a directory, write ``<dir>_classified/<Label>/<file>`` (the image with two
overlaid text lines, or a plain copy when ``overlay=False``) and
``<dir>_classified_results.xls`` (header IMAGE_NAME / PREDICTED_LABEL, one row
per file: name, label, confidence as a string), and return the xls path.

What is different: files go through the network in batches of ``BATCH`` images
(one C-ABI call each) instead of one ``Session.run`` per file (reference
infer.py:79-82), and the disk work either side of the network runs on a thread
pool: the next batch is decoded while the current one is classified, and the
overlay / copy writes of finished batches proceed in the background (cv2 drops
the GIL).  The artefacts and the order of the xls rows are unchanged.
"""
from __future__ import annotations

import os
import shutil
import struct
from collections import deque
from concurrent.futures import ThreadPoolExecutor
from glob import glob

import cv2

from .network import RoomNet

CLASS_LABELS = ['Backyard', 'Bathroom', 'Bedroom', 'Frontyard', 'Kitchen', 'LivingRoom']  # reference infer.py:22
INPUT_MODEL_PATH = './final_model/roomnet'       # reference infer.py:24
INPUT_IMAGES_DIR = './test_images/set2/images'   # reference infer.py:25
IMG_SIDE = 224                                   # reference infer.py:26
BATCH = 256
BATCH_BYTES = 512 << 20           # decoded pixels per device call (bounds host memory on large photos)
FILE_BATCH_BYTES = 256 << 20      # encoded bytes per device call on the file path (overlay=False)
IO_THREADS = min(32, os.cpu_count() or 1)
DECODE_AHEAD = 2 * IO_THREADS     # files being decoded ahead of the device
MAX_PENDING_WRITES = 64           # overlay / copy writes in flight


class _Biff2Sheet:
    """Single-sheet BIFF2 ``.xls`` writer for when xlwt is absent (it is, in this image).

    File = BOF(0x0009) · one LABEL(0x0004) record per cell · EOF(0x000A).
    Offers the two calls the results table needs: write(row, col, value), save(path).
    """

    def __init__(self):
        self._cells = []

    def write(self, row, col, value):
        self._cells.append((row, col, str(value)))

    def save(self, path):
        with open(path, 'wb') as f:
            f.write(struct.pack('<HHHH', 0x0009, 4, 2, 0x10))
            for row, col, text in self._cells:
                payload = text.encode('latin-1', 'replace')[:255]
                f.write(struct.pack('<HHHH3sB', 0x0004, 8 + len(payload), row, col, b'\0\0\0', len(payload)))
                f.write(payload)
            f.write(struct.pack('<HH', 0x000A, 0))


def _results_table():
    """(workbook, sheet) — xlwt when importable (reference infer.py:75-76), BIFF2 writer otherwise."""
    try:
        import xlwt
    except ImportError:
        sheet = _Biff2Sheet()
        return sheet, sheet
    book = xlwt.Workbook()
    return book, book.add_sheet('classification_results')


def _annotate(image, label, confidence):
    """The two overlay lines of reference infer.py:87-92 (positions, font scale and colours kept)."""
    height, width = image.shape[:2]
    scale = (height / 720.) * .85
    lines = (("Predicted Class: " + label, .90, (0, 255, 0)),
             ("Confidence: " + str(round(confidence * 100, 2)) + " %", .95, (255, 0, 0)))
    for text, rel_y, colour in lines:
        cv2.putText(image, text, (int(.5 * width), int(rel_y * height)), cv2.FONT_HERSHEY_SIMPLEX, scale, colour, 1,
                    cv2.LINE_AA)
    return image


def _read_image(path):
    image = cv2.imread(path)
    if image is None:
        # the reference dies in center_crop (network.py:138) on an unreadable file; keep the exception type
        raise AttributeError("'NoneType' object has no attribute 'shape' (cv2.imread failed on %s)" % path)
    return image


def _read_bytes(path):
    with open(path, 'rb') as f:
        return f.read()


def _emit(path, image, target_dir, name, label, confidence, overlay):
    """Output side of one file (reference infer.py:86-95)."""
    if overlay:
        cv2.imwrite(target_dir + os.sep + name, _annotate(image, label, confidence))
    else:
        shutil.copy(path, target_dir)


def classify_im_dir(nn, imgs_dir, overlay=True):
    print('Classifying images in', imgs_dir)
    paths = glob(imgs_dir + '/*')
    out_dir = imgs_dir + '_classified'
    xl_fpath = out_dir + '_results.xls'
    for label in CLASS_LABELS:
        os.makedirs(os.path.join(out_dir, label), exist_ok=True)
    print('Beginning inference..')
    workbook, sheet = _results_table()
    sheet.write(0, 0, 'IMAGE_NAME')
    sheet.write(0, 1, 'PREDICTED_LABEL')
    row = 1
    with ThreadPoolExecutor(max_workers=IO_THREADS) as pool:
        writes = deque()

        def flush(batch):
            """One device call for the images collected so far, then their output side on the pool."""
            nonlocal row
            if not batch:
                return
            top1, probs = nn.infer_optimized_batch([image for _, image in batch])
            for (path, image), cls, prob in zip(batch, top1, probs):
                label, confidence = CLASS_LABELS[cls], prob[cls]
                name = path.split(os.sep)[-1]
                print(path, '--->', label, confidence)
                # a plain copy does not need the pixels: do not keep them alive in the write queue
                writes.append(pool.submit(_emit, path, image if overlay else None, out_dir + os.sep + label, name,
                                          label, confidence, overlay))
                sheet.write(row, 0, name)
                sheet.write(row, 1, label)
                sheet.write(row, 2, str(confidence))
                row += 1
            while len(writes) > MAX_PENDING_WRITES:  # bounded: finished images are released as they are written
                writes.popleft().result()

        if not overlay and hasattr(nn, 'infer_files'):
            # plain copies do not need the pixels on the host: hand the encoded files to the library, which decodes
            # baseline JPEGs on the device (bit-identical to cv2.imread) and falls back to cv2 for the rest
            reading = deque()
            nxt = 0
            group, group_bytes = [], 0

            def flush_files(group):
                nonlocal row
                if not group:
                    return
                top1, probs = nn.infer_files([blob for _, blob in group])
                for (path, _), cls, prob in zip(group, top1, probs):
                    label, confidence = CLASS_LABELS[cls], prob[cls]
                    name = path.split(os.sep)[-1]
                    print(path, '--->', label, confidence)
                    writes.append(pool.submit(_emit, path, None, out_dir + os.sep + label, name, label, confidence,
                                              False))
                    sheet.write(row, 0, name)
                    sheet.write(row, 1, label)
                    sheet.write(row, 2, str(confidence))
                    row += 1
                while len(writes) > MAX_PENDING_WRITES:
                    writes.popleft().result()

            try:
                while nxt < len(paths) or reading:
                    while nxt < len(paths) and len(reading) < DECODE_AHEAD:
                        reading.append((paths[nxt], pool.submit(_read_bytes, paths[nxt])))
                        nxt += 1
                    path, fut = reading.popleft()
                    blob = fut.result()
                    group.append((path, blob))
                    group_bytes += len(blob)
                    if len(group) >= BATCH or group_bytes >= FILE_BATCH_BYTES:
                        flush_files(group)
                        group, group_bytes = [], 0
                flush_files(group)
            finally:
                for w in writes:
                    w.result()
                workbook.save(xl_fpath)
            return xl_fpath

        # decode a bounded window ahead of the device; a batch is closed at BATCH images or BATCH_BYTES of pixels,
        # whichever comes first (a directory of multi-megapixel photos must not hold hundreds of them in memory)
        decoding = deque()
        nxt = 0
        batch, batch_bytes = [], 0
        try:
            while nxt < len(paths) or decoding:
                while nxt < len(paths) and len(decoding) < DECODE_AHEAD:
                    decoding.append((paths[nxt], pool.submit(_read_image, paths[nxt])))
                    nxt += 1
                path, fut = decoding.popleft()
                image = fut.result()
                batch.append((path, image))
                batch_bytes += image.nbytes
                if len(batch) >= BATCH or batch_bytes >= BATCH_BYTES:
                    flush(batch)
                    batch, batch_bytes = [], 0
        except AttributeError:
            # unreadable file: like the reference (which processes file by file), everything before it is still
            # classified and written, then the error surfaces
            flush(batch)
            for w in writes:
                w.result()
            workbook.save(xl_fpath)
            raise
        flush(batch)
        for w in writes:
            w.result()  # surface I/O errors before the table is saved
    workbook.save(xl_fpath)
    return xl_fpath


if __name__ == '__main__':
    model = RoomNet(num_classes=len(CLASS_LABELS), im_side=IMG_SIDE, compute_bn_mean_var=False,
                    optimized_inference=True)
    model.load(INPUT_MODEL_PATH)
    classify_im_dir(model, INPUT_IMAGES_DIR)
