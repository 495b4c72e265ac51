"""ORACLE / TEST INFRASTRUCTURE — not part of the product path.

Executes the reference's SHIPPED inference graph (final_model/roomnet.meta, a
TensorFlow-1.13.1 ``MetaGraphDef``) node by node on the CPU, using the NumPy op
restatements in ``tf_ops``.  Structure, op order, strides, ksizes, paddings,
epsilons and resize sizes all come from the protobuf itself, not from a reading
of network.py, so this pins the *structure* of ``roomnet_oracle`` to the
reference's own artefact.  (The op *kernels* remain restatements — TensorFlow is
not available.)

Needs ``tensorboard`` (ships TF's graph protos) and /root/reference; only usable
in the build container.  Used by tests/test_oracle.py and
tests/golden/make_fixtures.py.
"""
from __future__ import annotations

import numpy as np

from . import tf_ops as ops


def load_graph_def(meta_path: str):
    from tensorboard.compat.proto import meta_graph_pb2
    m = meta_graph_pb2.MetaGraphDef()
    with open(meta_path, "rb") as f:
        m.ParseFromString(f.read())
    return m.graph_def


def _const_value(node):
    from tensorboard.util import tensor_util
    return tensor_util.make_ndarray(node.attr["value"].tensor)


class GraphInterpreter:
    def __init__(self, graph_def, variables: dict, dtype=np.float32):
        self.nodes = {n.name: n for n in graph_def.node}
        self.vars = variables
        self.dtype = np.dtype(dtype)
        self.executed_ops = []

    def run(self, fetches, feed: dict):
        cache = {k: np.asarray(v).astype(self.dtype) for k, v in feed.items()}
        return [self._eval(f, cache) for f in fetches]

    def _eval(self, name, cache):
        name = name.split(":")[0].lstrip("^")
        if name in cache:
            return cache[name]
        node = self.nodes[name]
        op = node.op
        ins = [i for i in node.input if not i.startswith("^")]
        a = lambda k: self._eval(ins[k], cache)  # noqa: E731
        dt = self.dtype
        if op == "Const":
            v = _const_value(node)
            out = v.astype(dt) if v.dtype.kind == "f" else v
        elif op == "VariableV2":
            out = self.vars[name].astype(dt)
        elif op == "Identity":
            out = a(0)
        elif op == "Conv2D":
            at = node.attr
            assert list(at["strides"].list.i) == [1, 1, 1, 1]
            assert at["padding"].s == b"VALID"
            assert at["data_format"].s in (b"NHWC", b"")
            dil = list(at["dilations"].list.i)
            assert dil in ([1, 1, 1, 1], [])
            out = ops.conv2d_valid(a(0), a(1))
        elif op == "Relu6":
            out = ops.relu6(a(0))
        elif op == "AvgPool":
            at = node.attr
            ks = list(at["ksize"].list.i)
            st = list(at["strides"].list.i)
            assert at["padding"].s == b"VALID"
            assert ks[0] == ks[3] == 1 and ks[1] == ks[2] and st[1] == st[2]
            out = ops.avg_pool_valid(a(0), ks[1], st[1])
        elif op == "FusedBatchNorm":
            at = node.attr
            assert at["is_training"].b is False
            eps = np.float32(at["epsilon"].f)
            out = ops.batch_norm_inference(a(0), a(1), a(2), a(3), a(4), eps)
        elif op == "ResizeBilinear":
            assert node.attr["align_corners"].b is False
            assert "half_pixel_centers" not in node.attr or not node.attr["half_pixel_centers"].b
            size = a(1)
            out = ops.resize_bilinear_legacy(a(0), int(size[0]), int(size[1]))
        elif op == "Add":
            out = a(0) + a(1)
        elif op == "Sub":
            out = a(0) - a(1)
        elif op == "Mul":
            out = a(0) * a(1)
        elif op == "Rsqrt":
            out = dt.type(1) / np.sqrt(a(0))
        elif op == "Reshape":
            out = a(0).reshape([int(v) for v in a(1)])
        elif op == "MatMul":
            assert not node.attr["transpose_a"].b and not node.attr["transpose_b"].b
            out = a(0) @ a(1)
        elif op == "BiasAdd":
            out = a(0) + a(1)
        elif op == "Softmax":
            out = ops.softmax(a(0))
        elif op == "ArgMax":
            assert int(a(1)) in (-1, a(0).ndim - 1)
            out = ops.argmax_first(a(0))
        else:
            raise NotImplementedError("op %s (%s) is not on the inference path" % (op, name))
        self.executed_ops.append((op, name))
        cache[name] = out
        return out


def run_reference_graph(x_rgb_float, variables, meta_path, dtype=np.float32,
                        fetches=("ArgMax", "Softmax", "Relu6_3", "dense_3/BiasAdd")):
    """``sess.run(outs_final, {x_tensor: x})`` over the shipped graph (reference network.py:155)."""
    g = load_graph_def(meta_path)
    interp = GraphInterpreter(g, variables, dtype)
    outs = interp.run(list(fetches), {"input_x_tensor": x_rgb_float})
    return dict(zip(fetches, outs)), interp
