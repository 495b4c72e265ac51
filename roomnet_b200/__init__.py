"""roomnet_b200 — B200-native (sm_100a) RoomNet inference behind the reference's API.

Public surface (mirrors the reference): ``RoomNet`` (network.py) and
``classify_im_dir`` / ``CLASS_LABELS`` (infer.py).  All arithmetic happens in
``libroomnet.so`` (hand-written CUDA, C ABI in include/roomnet.h); importing this
package fails if that library has not been built — there is no CPU fallback.
"""
from . import _capi
from .infer import CLASS_LABELS, classify_im_dir
from .network import RoomNet

__all__ = ["RoomNet", "classify_im_dir", "CLASS_LABELS", "_capi"]
