"""Golden vectors produced by EXECUTING THE REFERENCE'S OWN SOURCE for the wiring of the hot path.

TensorFlow cannot be installed here, so the reference cannot run as shipped.  What can run is its Python: this script puts
an eager numpy/torch stand-in for the handful of `tensorflow` symbols the reference uses on sys.path and then imports
/root/reference/custom_layers.py and runs, unmodified:
    yolov4_neck()  -> cspdarknet53() -> conv / residual_block / csp_block      (custom_layers.py:5-198)  graph order, concat
                                                                               order, add-after-activation, SPP order
    load_weights()                                                             (utils.py:12-53)  darknet file order, BN
                                                                               row permutation, OIHW -> HWIO transpose
    yolov4_head() -> get_boxes(),  nms() up to the TF op                       (custom_layers.py:201-298)  decode formulas,
                                                                               anchor / stride / xyscale use, flattening order
The per-op arithmetic behind the stand-in (convolution, batch norm in inference mode with eps 1e-3, leaky / mish, max-pool
'same', nearest up-sampling, sigmoid, exp) is a restatement of the documented TF semantics -- so is the oracle's; what these
vectors pin is everything the reference's own code decides.  tf.image.combined_non_max_suppression itself is NOT available:
its inputs are captured (and pinned), its semantics stay restated.
    python tests/golden/make_golden_refgraph.py     (needs /root/reference; run in the build container; ~1 min)"""
import ast
import hashlib
import os
import sys
import textwrap
import types

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path[:0] = [ROOT, os.path.join(ROOT, 'oracle')]
import y4_oracle as O  # noqa: E402

# ---------------------------------------------------------------------------------------------------------------------
# eager stand-in for the tensorflow symbols custom_layers.py touches
# ---------------------------------------------------------------------------------------------------------------------
REG = {'order': [], 'by_name': {}, 'count': {}, 'replay': None, 'cursor': 0}


def _register(obj, base):
    if REG['replay'] is not None:                      # second pass: hand back the layer created in the first one
        o = REG['order'][REG['cursor']]
        REG['cursor'] += 1
        assert type(o) is type(obj), (type(o), type(obj))
        return o
    n = REG['count'].get(base, 0)
    REG['count'][base] = n + 1
    obj.name = base if n == 0 else f'{base}_{n}'       # Keras' automatic layer names
    REG['order'].append(obj)
    REG['by_name'][obj.name] = obj
    return obj


class _Layer:
    def __new__(cls, *a, **k):
        return _register(object.__new__(cls), cls.base)

    def __init__(self, *a, **k):
        if not hasattr(self, '_made'):
            self._made = True
            self.setup(*a, **k)


class Conv2D(_Layer):
    base = 'conv2d'

    def setup(self, filters, kernel_size, strides=1, padding='valid', use_bias=True, kernel_initializer=None):
        self.filters, self.kernel_size, self.strides, self.padding, self.use_bias = filters, (kernel_size, kernel_size), strides, padding, use_bias
        self.w = None; self.b = None; self.input_shape = None

    def set_weights(self, ws):
        self.w = np.asarray(ws[0], np.float32)
        assert self.w.shape == (self.kernel_size[0], self.kernel_size[0], self.input_shape[-1], self.filters)
        self.b = np.asarray(ws[1], np.float32) if len(ws) > 1 else None

    def __call__(self, x):
        self.input_shape = (None,) + x.shape[1:]
        k = self.kernel_size[0]
        w = self.w if self.w is not None else np.zeros((k, k, x.shape[-1], self.filters), np.float32)
        pad = k // 2 if self.padding == 'same' else 0
        y = torch.nn.functional.conv2d(torch.from_numpy(np.ascontiguousarray(x.transpose(0, 3, 1, 2))),
                                       torch.from_numpy(np.ascontiguousarray(w.transpose(3, 2, 0, 1))),
                                       bias=None if self.b is None else torch.from_numpy(self.b), stride=self.strides, padding=pad)
        return np.ascontiguousarray(y.numpy().transpose(0, 2, 3, 1))


class BatchNormalization(_Layer):
    base = 'batch_normalization'

    def setup(self):
        self.p = None

    def set_weights(self, ws):
        self.p = [np.asarray(a, np.float32) for a in ws]       # Keras order: gamma, beta, moving_mean, moving_variance

    def __call__(self, x):
        if self.p is None:
            return x
        g, b, m, v = self.p
        inv = g / np.sqrt(v + np.float32(1e-3))                # Keras default epsilon
        return (x * inv + (b - m * inv)).astype(np.float32)


class LeakyReLU(_Layer):
    base = 'leaky_re_lu'

    def setup(self, alpha=0.3):
        self.alpha = np.float32(alpha)

    def __call__(self, x):
        return np.where(x >= 0, x, self.alpha * x).astype(np.float32)


class ZeroPadding2D(_Layer):
    base = 'zero_padding2d'

    def setup(self, padding):
        self.padding = padding

    def __call__(self, x):
        (t, b), (l, r) = self.padding
        return np.pad(x, ((0, 0), (t, b), (l, r), (0, 0)))


class Add(_Layer):
    base = 'add'

    def setup(self):
        pass

    def __call__(self, xs):
        return (xs[0] + xs[1]).astype(np.float32)


class Concatenate(_Layer):
    base = 'concatenate'

    def setup(self):
        pass

    def __call__(self, xs):
        return np.concatenate(xs, axis=-1)


class MaxPooling2D(_Layer):
    base = 'max_pooling2d'

    def setup(self, pool_size, strides, padding):
        assert strides == 1 and padding == 'same'
        self.k = pool_size

    def __call__(self, x):
        y = torch.nn.functional.max_pool2d(torch.from_numpy(np.ascontiguousarray(x.transpose(0, 3, 1, 2))), self.k, 1, self.k // 2)
        return np.ascontiguousarray(y.numpy().transpose(0, 2, 3, 1))


class UpSampling2D(_Layer):
    base = 'up_sampling2d'

    def setup(self):
        pass

    def __call__(self, x):
        return x.repeat(2, axis=1).repeat(2, axis=2)


CAPTURE = {}


def _sigmoid(x):
    return (1.0 / (1.0 + np.exp(-x.astype(np.float32)))).astype(np.float32)


def _combined_nms(boxes, scores, max_output_size_per_class, max_total_size, iou_threshold, score_threshold):
    CAPTURE['nms_boxes'] = np.asarray(boxes)
    CAPTURE['nms_scores'] = np.asarray(scores)
    CAPTURE['nms_args'] = (max_output_size_per_class, max_total_size, float(iou_threshold), float(score_threshold))
    return None, None, None, None


tf = types.ModuleType('tensorflow')
tf.math = types.SimpleNamespace(tanh=lambda x: np.tanh(x).astype(np.float32), softplus=lambda x: np.logaddexp(np.float32(0), x).astype(np.float32))
tf.float32 = np.float32
tf.reshape = lambda x, s: np.reshape(x, tuple(int(v) for v in s))
tf.shape = lambda x: x.shape
tf.split = lambda x, sizes, axis=-1: np.split(x, np.cumsum(sizes)[:-1], axis=axis)
tf.sigmoid = _sigmoid
tf.exp = lambda x: np.exp(x).astype(np.float32)
tf.concat = lambda xs, axis=-1: np.concatenate(list(xs), axis=axis)
tf.range = lambda n: np.arange(n, dtype=np.int32)
tf.meshgrid = lambda a, b: list(np.meshgrid(a, b))                 # default 'xy' indexing in both libraries
tf.stack = lambda xs, axis=0: np.stack(xs, axis=axis)
tf.expand_dims = lambda x, axis: np.expand_dims(x, axis)
tf.cast = lambda x, dtype: np.asarray(x).astype(dtype)
tf.zeros = lambda shape: np.zeros(tuple(int(v) for v in shape), np.float32)
tf.image = types.SimpleNamespace(combined_non_max_suppression=_combined_nms)
keras = types.ModuleType('tensorflow.keras')
layers = types.ModuleType('tensorflow.keras.layers')
for c in (Conv2D, BatchNormalization, LeakyReLU, ZeroPadding2D, Add, Concatenate, MaxPooling2D, UpSampling2D):
    setattr(layers, c.__name__, c)
layers.Input = lambda shape: np.zeros((1,) + tuple(shape), np.float32)
initializers = types.ModuleType('tensorflow.keras.initializers')
initializers.RandomNormal = lambda mean=0.0, stddev=0.01: None
models = types.ModuleType('tensorflow.keras.models')


class _M:                                                  # models.Model(input, outputs): the neck reads .output
    def __init__(self, i, o):
        self.output = o


models.Model = _M
keras.layers, keras.initializers, keras.models = layers, initializers, models
tf.keras = keras
for name, mod in (('tensorflow', tf), ('tensorflow.keras', keras), ('tensorflow.keras.layers', layers),
                  ('tensorflow.keras.initializers', initializers), ('tensorflow.keras.models', models)):
    sys.modules[name] = mod
sys.path.insert(0, '/root/reference')
import custom_layers as CL  # noqa: E402  -- the reference's own file

# ---------------------------------------------------------------------------------------------------------------------
# pass 1: build the layers (tiny input); load the darknet file with the reference's load_weights; pass 2: run at 416
# ---------------------------------------------------------------------------------------------------------------------
NC, S = 80, 416
CL.yolov4_neck(np.zeros((1, 32, 32, 3), np.float32), NC)
n_conv = REG['count']['conv2d']
assert n_conv == 110 and REG['count']['batch_normalization'] == 107, REG['count']

W = O.synth_weights(seed=1)
blob = W.to_darknet_bytes()
path = '/tmp/y4_synth.weights'
open(path, 'wb').write(blob)


def cut(pth, name):
    src = open(pth).read()
    fn = next(n for n in ast.parse(src).body if isinstance(n, ast.FunctionDef) and n.name == name)
    return textwrap.dedent(ast.get_source_segment(src, fn))


npx = types.SimpleNamespace(**{k: getattr(np, k) for k in ('fromfile', 'int32', 'float32')}, product=np.prod)   # numpy 2 dropped np.product
ns = {'np': npx}
exec(cut('/root/reference/utils.py', 'load_weights'), ns)
model = types.SimpleNamespace(get_layer=lambda name: REG['by_name'][name])
ns['load_weights'](model, path)

# per-layer digests of what the reference's loader handed to Keras (HWIO kernels, [gamma, beta, mean, var], head biases)
loader = {}
for i in range(110):
    c = REG['by_name']['conv2d' if i == 0 else f'conv2d_{i}']
    h = hashlib.sha256(c.w.tobytes())
    if c.b is not None:
        h.update(c.b.tobytes())
    loader[f'conv{i}'] = h.hexdigest()
bn_i = 0
for i in range(110):
    if i in (93, 101, 109):
        continue
    b = REG['by_name']['batch_normalization' if bn_i == 0 else f'batch_normalization_{bn_i}']
    loader[f'bn{i}'] = hashlib.sha256(np.stack(b.p).tobytes()).hexdigest()
    bn_i += 1

REG['replay'] = True; REG['cursor'] = 0
img = O.synth_images(0, 0, 1, S).astype(np.float32)
heads = CL.yolov4_neck(img, NC)
assert [h.shape for h in heads] == [(1, 52, 52, 255), (1, 26, 26, 255), (1, 13, 13, 255)]

# decode + nms glue on the reference's code, from (a) these heads and (b) a sparse crafted head set that compresses well
cfg = {}
exec(open('/root/reference/config.py').read(), cfg)
yc = cfg['yolo_config']
anchors = np.array(yc['anchors']).reshape((3, 3, 2))
xyscale = yc['xyscale']


def ref_decode(hs):
    out = CL.yolov4_head(hs, NC, anchors, xyscale)
    CL.nms(out, yc['img_size'], NC, iou_threshold=yc['iou_threshold'], score_threshold=yc['score_threshold'])
    return out, CAPTURE['nms_boxes'].copy(), CAPTURE['nms_scores'].copy()


out_a, nb_a, ns_a = ref_decode(heads)
rng = np.random.default_rng(99)
sparse = []
for g in (52, 26, 13):
    h = np.zeros((1, g, g, 3, 85), np.float32)
    h[..., 4] = -8.0; h[..., 5:] = -6.0
    for _ in range(60):
        r, c, a = int(rng.integers(0, g)), int(rng.integers(0, g)), int(rng.integers(0, 3))
        h[0, r, c, a, :4] = rng.normal(0, 1, 4).astype(np.float32) * np.float32([1, 1, 0.4, 0.4])
        h[0, r, c, a, 4] = np.float32(rng.uniform(-1, 6))
        h[0, r, c, a, 5 + int(rng.integers(0, NC))] = np.float32(rng.uniform(-1, 6))
    sparse.append(h.reshape(1, g, g, 255))
out_b, nb_b, ns_b = ref_decode(sparse)

pick = rng.choice(heads[0].size, 1500, replace=False), rng.choice(heads[1].size, 800, replace=False), rng.choice(heads[2].size, 400, replace=False)
fix = {'img_size': S, 'weights_sha256': hashlib.sha256(blob).hexdigest(), 'nms_args': np.array(CAPTURE['nms_args'], np.float64),
       'loader_keys': np.array(sorted(loader)), 'loader_sha256': np.array([loader[k] for k in sorted(loader)])}
for i, h in enumerate(heads):
    fix[f'head{i}_idx'] = pick[i].astype(np.int64)
    fix[f'head{i}_val'] = h.reshape(-1)[pick[i]]
    fix[f'head{i}_absmax'] = np.float32(np.abs(h).max())
    fix[f'head{i}_sum'] = np.float64(h.astype(np.float64).sum())
# decode of the real heads: values at the 2000 highest-scoring (box, class) pairs + the candidate count above the threshold
sc = ns_a[0]
top = np.argsort(-sc.reshape(-1), kind='stable')[:2000]
fix['a_top_flat'] = top.astype(np.int64)
fix['a_top_scores'] = sc.reshape(-1)[top]
fix['a_top_boxes'] = nb_a[0, top // NC, 0, :]
fix['a_count_above_thr'] = np.int64((sc > yc['score_threshold']).sum())
for i in range(3):
    fix[f'sparse_head{i}'] = sparse[i]
fix['b_nms_boxes'] = nb_b[0, :, 0, :]                      # (10647, 4): compresses (background cells decode to few distinct values)
fix['b_nms_scores_nonzero_idx'] = np.nonzero(ns_b[0].reshape(-1) > 1e-4)[0].astype(np.int64)
fix['b_nms_scores_nonzero'] = ns_b[0].reshape(-1)[fix['b_nms_scores_nonzero_idx']]
fix['b_nms_scores_sha256'] = hashlib.sha256(ns_b[0].tobytes()).hexdigest()
fix['b_xywh0'] = out_b[3][0, :4, :4]                       # pred_box_xywh corner of scale 0 (unused by inference, shape check)
np.savez_compressed(os.path.join(HERE, 'refgraph_416.npz'), **fix)
print('written', os.path.getsize(os.path.join(HERE, 'refgraph_416.npz')), 'bytes; candidates above thr:', int(fix['a_count_above_thr']),
      'head absmax', [float(np.abs(h).max()) for h in heads])
