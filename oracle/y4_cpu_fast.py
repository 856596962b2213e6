"""ORACLE — CPU BASELINE INFRASTRUCTURE ONLY (bench.py's cpu_baseline / --impl reference legs and tests/ import this;
the product never does).

The reference's hot path on the HOST with the best CPU kernels this image offers, as a stand-in for the tf.keras CPU
forward that cannot be installed here (no TensorFlow wheel, no network): the conv stack through torch's CPU backend
(oneDNN / MKL-DNN, channels_last, fp32, all host threads -- the same library family TF's CPU Conv2D uses), the head decode in
numpy and combined_non_max_suppression in compiled C (oracle/nms_ref.c, OpenMP over (image, class) like TF's kernel).
Same op order as the reference (conv -> BatchNorm(eps 1e-3) -> mish / leaky, custom_layers.py:5-31) and the same netlist
(oracle/netspec.py).  tests/test_oracle.py holds it to the numpy oracle.  kind = "port"."""
import ctypes as C
import os

import numpy as np

from netspec import build_netlist
import y4_oracle as O

_HERE = os.path.dirname(os.path.abspath(__file__))
_lib = None


def nms_lib():
    """oracle/libnmsref.so, built by __graft_entry__.build() (gcc -O2 -ffp-contract=off -fopenmp)."""
    global _lib
    if _lib is None:
        path = os.path.join(_HERE, 'libnmsref.so')
        if not os.path.exists(path):
            raise ImportError(f'{path} missing: run `python -c "import __graft_entry__ as g; g.build()"`')
        _lib = C.CDLL(path)
        _lib.y4ref_combined_nms.argtypes = [C.c_void_p, C.c_void_p, C.c_int32, C.c_int32, C.c_int32, C.c_float, C.c_float,
                                            C.c_int32, C.c_int32, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
    return _lib


def combined_nms_c(boxes, scores, iou_threshold=O.IOU_THRESHOLD, score_threshold=O.SCORE_THRESHOLD,
                   max_per_class=O.MAX_BOXES, max_total=O.MAX_BOXES):
    boxes = np.ascontiguousarray(boxes, np.float32)
    scores = np.ascontiguousarray(scores, np.float32)
    B, N, Cn = scores.shape
    T = max_total
    ob = np.empty((B, T, 4), np.float32); osc = np.empty((B, T), np.float32); ocl = np.empty((B, T), np.float32)
    ov = np.empty((B,), np.int32); oi = np.empty((B, T), np.int32)
    p = lambda a: a.ctypes.data_as(C.c_void_p)  # noqa: E731
    nms_lib().y4ref_combined_nms(p(boxes), p(scores), B, N, Cn, iou_threshold, score_threshold, max_per_class, T,
                                 p(ob), p(osc), p(ocl), p(ov), p(oi))
    return ob, osc, ocl, ov, oi


class TorchNet:
    """The 110 convs as torch CPU ops (weights converted once: HWIO -> OIHW, channels_last)."""

    def __init__(self, W):
        import torch
        self.torch = torch
        self.ops, self.heads = build_netlist(W.num_classes)
        self.params = {}
        for o in self.ops:
            if o.kind != 'conv':
                continue
            p = W.p[o.idx]
            w = torch.from_numpy(np.ascontiguousarray(p['w'].transpose(3, 2, 0, 1))).contiguous(memory_format=torch.channels_last)
            if o.bn:
                self.params[o.idx] = (w, tuple(torch.from_numpy(p[k].copy()) for k in ('gamma', 'beta', 'mean', 'var')))
            else:
                self.params[o.idx] = (w, torch.from_numpy(p['bias'].copy()))

    def forward(self, imgs):
        """imgs (B,S,S,3) float32 NHWC in [0,1] -> three heads (B,g,g,255) float32 NHWC (numpy)."""
        torch = self.torch
        F = torch.nn.functional
        last = {}
        for i, o in enumerate(self.ops):
            for n in o.ins:
                last[n] = i
        with torch.no_grad():
            t = {'img': torch.from_numpy(np.ascontiguousarray(imgs, np.float32)).permute(0, 3, 1, 2).contiguous(memory_format=torch.channels_last)}
            for i, o in enumerate(self.ops):
                if o.kind == 'conv':
                    x = t[o.ins[0]]
                    w, rest = self.params[o.idx]
                    if o.stride == 2:
                        x = F.pad(x, (1, 0, 1, 0))                     # ZeroPadding2D(((1,0),(1,0))) (custom_layers.py:10)
                        y = F.conv2d(x, w, None, stride=2)
                    else:
                        y = F.conv2d(x, w, None, padding=o.k // 2)
                    if o.bn:
                        g, b, m, v = rest
                        y = F.batch_norm(y, m, v, g, b, training=False, eps=O.BN_EPS)
                    else:
                        y = y + rest.view(1, -1, 1, 1)
                    if o.act == 'mish':
                        y = y * torch.tanh(F.softplus(y))             # custom_layers.py:6-7
                    elif o.act == 'leaky':
                        y = F.leaky_relu(y, 0.1)
                    t[o.out] = y
                elif o.kind == 'add':
                    t[o.out] = t[o.ins[0]] + t[o.ins[1]]
                elif o.kind == 'concat':
                    t[o.out] = torch.cat([t[n] for n in o.ins], dim=1)
                elif o.kind == 'maxpool':
                    t[o.out] = F.max_pool2d(t[o.ins[0]], o.pool, stride=1, padding=o.pool // 2)
                elif o.kind == 'upsample':
                    t[o.out] = F.interpolate(t[o.ins[0]], scale_factor=2, mode='nearest')
                for n in o.ins:
                    if last.get(n) == i and n not in self.heads and n != 'img':
                        t.pop(n, None)
            return [t[h].permute(0, 2, 3, 1).contiguous().numpy() for h in self.heads]

    def predict(self, imgs):
        """inference_model.predict on the host: forward (oneDNN) + decode (numpy) + combined NMS (C)."""
        heads = self.forward(imgs)
        boxes, scores = O.decode_heads(heads, imgs.shape[1], num_classes=(heads[0].shape[-1] // 3) - 5)
        return combined_nms_c(boxes, scores)
