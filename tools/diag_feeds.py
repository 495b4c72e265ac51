import sys, numpy as np
sys.path.insert(0,'/root/repo')
from oracle.roomnet_oracle import RoomNetOracle, synthetic_suite
from oracle.tf_bundle import default_checkpoint_prefix
from roomnet_b200 import _capi
imgs = synthetic_suite(8)
h = _capi.Handle(precision='fp16'); h.load_tf_checkpoint(default_checkpoint_prefix())
o = RoomNetOracle(dtype=np.float32, conv_backend='torch').load()
x = o.normalise(imgs).astype(np.float32)
ref = o.forward(x)['logits']
t0,p0,l0 = h.infer_u8_bgr(imgs, want_logits=True)
A = [h.debug_activation(i) for i in range(10)]
t2,p2,l2 = h.infer_f32_rgb(x, want_logits=True)
B = [h.debug_activation(i) for i in range(10)]
print('u8 vs ref', np.abs(l0-ref).max(axis=1))
print('f32 vs ref', np.abs(l2-ref).max(axis=1))
for i in range(10):
    d = np.abs(A[i]-B[i]); print('layer',i,'max diff',d.max(), 'rel', d.max()/np.abs(A[i]).max(), 'argmax', np.unravel_index(d.argmax(), d.shape))
