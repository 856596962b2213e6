"""ORACLE — CPU restatement of the reference hot path.  TEST INFRASTRUCTURE ONLY.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
import this file.  The product (yolo-v4-tf.keras_b200/) never does.

PARITY UNPINNED: the reference executes inside TensorFlow/Keras, which is not installed in this
image (and cannot be: no network), and the only recorded output of the reference
(notebook/Inference.ipynb cell 5) needs the real yolov4.weights, which is not on disk.  The
functions below therefore follow the reference source line by line plus TensorFlow's documented
op semantics; they are cross-checked against independent implementations
(torch.nn.functional on CPU, brute-force NMS) in tests/test_oracle.py.

Everything is NHWC float32 (float64 when dtype=np.float64), numpy only.
"""
import os
import struct
import numpy as np

from netspec import build_netlist, conv_ops, darknet_file_floats

# /root/reference/config.py:1-17
ANCHORS = np.array([12, 16, 19, 36, 40, 28, 36, 75, 76, 55, 72, 146, 142, 110, 192, 243, 459, 401],
                   dtype=np.float32).reshape(3, 3, 2)          # models.py:29
STRIDES = (8, 16, 32)
XYSCALE = (1.2, 1.1, 1.05)
MAX_BOXES = 100
IOU_THRESHOLD = 0.413
SCORE_THRESHOLD = 0.3
BN_EPS = 1e-3           # Keras BatchNormalization default (custom_layers.py:26), NOT darknet's 1e-5


# ----------------------------------------------------------------------------------------------
# elementwise pieces
# ----------------------------------------------------------------------------------------------
def mish(x):
    """custom_layers.py:6-7  x * tanh(softplus(x))"""
    sp = np.logaddexp(0, x).astype(x.dtype)
    return (x * np.tanh(sp)).astype(x.dtype)


def leaky(x):
    """custom_layers.py:30  LeakyReLU(alpha=0.1)"""
    return np.where(x > 0, x, x * x.dtype.type(0.1)).astype(x.dtype)


def sigmoid(x):
    return (1.0 / (1.0 + np.exp(-x))).astype(x.dtype)


# ----------------------------------------------------------------------------------------------
# conv / pool / upsample  (custom_layers.py:5-31, :130-132, :147)
# ----------------------------------------------------------------------------------------------
def conv2d_raw(x, w_hwio, stride, kperm=None):
    """Cross-correlation, NHWC x HWIO.  stride 1: 'same'.  stride 2: ZeroPadding2D(((1,0),(1,0)))
    followed by a 'valid' stride-2 conv (custom_layers.py:9-15).
    kperm: optional numpy Generator; the K = k*k*Cin terms of every dot product are then summed in a random order
    (same products, different rounding) -- a second, equally valid floating-point evaluation of the same conv (TF's own
    summation order is unspecified), used to find test inputs whose detections do not hinge on round-off."""
    B, H, W, Cin = x.shape
    k = w_hwio.shape[0]
    Cout = w_hwio.shape[3]
    if stride == 1:
        p = k // 2
        xp = np.pad(x, ((0, 0), (p, p), (p, p), (0, 0))) if p else x
        OH, OW = H, W
    else:
        assert k == 3 and stride == 2
        xp = np.pad(x, ((0, 0), (1, 0), (1, 0), (0, 0)))
        OH, OW = (H + 1 - k) // 2 + 1, (W + 1 - k) // 2 + 1
    wm = np.ascontiguousarray(w_hwio.reshape(k * k * Cin, Cout))
    out = np.empty((B, OH, OW, Cout), dtype=x.dtype)
    for b in range(B):                      # per image: bounds the im2col buffer
        if k == 1:
            cols = xp[b].reshape(OH * OW, Cin)
        else:
            taps = [xp[b, kh:kh + stride * (OH - 1) + 1:stride, kw:kw + stride * (OW - 1) + 1:stride, :]
                    for kh in range(k) for kw in range(k)]
            cols = np.concatenate(taps, axis=-1).reshape(OH * OW, k * k * Cin)
        if kperm is not None:
            perm = kperm.permutation(cols.shape[1])
            out[b] = (np.ascontiguousarray(cols[:, perm]) @ np.ascontiguousarray(wm[perm])).reshape(OH, OW, Cout)
        else:
            out[b] = (cols @ wm).reshape(OH, OW, Cout)
    return out


def maxpool_same(x, size):
    """MaxPooling2D(pool_size=size, strides=1, padding='same'): padded cells ignored (== -inf)."""
    B, H, W, C = x.shape
    p = size // 2
    neg = np.full((B, H + 2 * p, W + 2 * p, C), -np.inf, dtype=x.dtype)
    neg[:, p:p + H, p:p + W] = x
    out = np.full_like(x, -np.inf)
    for dy in range(size):
        for dx in range(size):
            np.maximum(out, neg[:, dy:dy + H, dx:dx + W], out=out)
    return out


def upsample2x(x):
    """UpSampling2D() nearest: out[i,j] = in[i//2, j//2]"""
    return np.repeat(np.repeat(x, 2, axis=1), 2, axis=2)


# ----------------------------------------------------------------------------------------------
# weights: synthetic generator + darknet file format (utils.py:12-53)
# ----------------------------------------------------------------------------------------------
class Weights:
    """Per conv idx: 'w' (k,k,Cin,Cout) HWIO fp32; BN convs: gamma,beta,mean,var; heads: bias."""

    def __init__(self, num_classes=80):
        self.num_classes = num_classes
        self.convs = conv_ops(num_classes)
        self.p = [dict() for _ in self.convs]

    # --- darknet serialisation, file order == conv idx order (utils.py:19-47)
    def to_darknet_bytes(self) -> bytes:
        chunks = [struct.pack('<5i', 0, 2, 5, 0, 0)]          # major, minor, revision, seen, _ (utils.py:16)
        for o, p in zip(self.convs, self.p):
            if o.bn:       # darknet order [beta, gamma, mean, variance]  (utils.py:29-32)
                chunks.append(np.stack([p['beta'], p['gamma'], p['mean'], p['var']]).astype('<f4').tobytes())
            else:
                chunks.append(p['bias'].astype('<f4').tobytes())
            # darknet (out, in, h, w)  <-  HWIO.transpose(3,2,0,1)   (inverse of utils.py:42)
            chunks.append(np.ascontiguousarray(p['w'].transpose(3, 2, 0, 1)).astype('<f4').tobytes())
        return b''.join(chunks)

    def save_darknet(self, path):
        with open(path, 'wb') as f:
            f.write(self.to_darknet_bytes())

    @classmethod
    def from_darknet_bytes(cls, buf: bytes, num_classes=80):
        self = cls(num_classes)
        need = 20 + 4 * darknet_file_floats(num_classes)
        if len(buf) != need:                                    # mirror of utils.py:50-53
            raise ValueError(f'darknet weights: expected {need} bytes, got {len(buf)}')
        off = 20
        for o, p in zip(self.convs, self.p):
            if o.bn:
                bn = np.frombuffer(buf, '<f4', 4 * o.cout, off).reshape(4, o.cout); off += 16 * o.cout
                p['beta'], p['gamma'], p['mean'], p['var'] = (bn[i].copy() for i in range(4))
            else:
                p['bias'] = np.frombuffer(buf, '<f4', o.cout, off).copy(); off += 4 * o.cout
            n = o.cout * o.cin * o.k * o.k
            w = np.frombuffer(buf, '<f4', n, off).reshape(o.cout, o.cin, o.k, o.k); off += 4 * n
            p['w'] = np.ascontiguousarray(w.transpose(2, 3, 1, 0))  # -> HWIO (utils.py:42)
        assert off == len(buf)
        return self

    @classmethod
    def load_darknet(cls, path, num_classes=80):
        with open(path, 'rb') as f:
            return cls.from_darknet_bytes(f.read(), num_classes)

    def folded(self, idx, dtype=np.float32):
        """(w', b') with inference BN folded in: w' = w*gamma/sqrt(var+eps), b' = beta - mean*gamma/sqrt(var+eps)."""
        o, p = self.convs[idx], self.p[idx]
        if not o.bn:
            return p['w'].astype(dtype), p['bias'].astype(dtype)
        s = (p['gamma'].astype(np.float64) / np.sqrt(p['var'].astype(np.float64) + BN_EPS))
        return (p['w'].astype(np.float64) * s).astype(dtype), \
               (p['beta'].astype(np.float64) - p['mean'].astype(np.float64) * s).astype(dtype)


def synth_weights(seed=1, num_classes=80, calib_size=416, calib_batch=1,
                  obj_bias=-4.5, cls_bias=-4.0, logit_std=2.0, wh_std=0.35):
    """Seeded synthetic weights.  The reference's own init (RandomNormal(0, 0.01) + identity BN,
    custom_layers.py:22) collapses activations to 0 over 110 layers (every score = 0.25 < 0.3 ->
    no detections), so BN statistics are *calibrated* like a trained net's: a float64 forward on a
    seeded calibration batch sets each BN layer's (mean, var) to the measured per-channel moments of
    its conv output, and scales each head channel to a target logit spread.  float64 moments are
    rounded to float32, which makes the result reproducible across hosts/BLAS builds."""
    rng = np.random.default_rng(seed)
    W = Weights(num_classes)
    ops, heads = build_netlist(num_classes)
    x0 = rng.random((calib_batch, calib_size, calib_size, 3))      # float64 in [0,1)
    t = {'img': x0}
    for o in ops:
        if o.kind == 'conv':
            fan_in = o.k * o.k * o.cin
            w = rng.standard_normal((o.k, o.k, o.cin, o.cout)) / np.sqrt(fan_in)
            w32 = w.astype(np.float32)
            y = conv2d_raw(t[o.ins[0]], w32.astype(np.float64), o.stride)
            p = W.p[o.idx]
            if o.bn:
                # NOT the centred moments: BN that removes the per-channel mean of the signal (but not of a
                # perturbation) makes a random deep net expansive (~1.1x per layer; fp32 round-off then reaches
                # 1e-4 at the heads and no two fp32 evaluations agree to the parity tolerance).  A small random
                # mean + the second moment about it keeps signal and perturbation gains equal.
                rms = np.sqrt((y * y).mean(axis=(0, 1, 2)))
                mean = (0.1 * rms * rng.standard_normal(o.cout)).astype(np.float32)
                var = ((y - mean.astype(np.float64)) ** 2).mean(axis=(0, 1, 2)).astype(np.float32)
                gamma = rng.uniform(0.8, 1.2, o.cout).astype(np.float32)
                beta = (0.2 * rng.standard_normal(o.cout)).astype(np.float32)
                p.update(w=w32, gamma=gamma, beta=beta, mean=mean, var=var)
                s = gamma.astype(np.float64) / np.sqrt(var.astype(np.float64) + BN_EPS)
                y = (y - mean.astype(np.float64)) * s + beta.astype(np.float64)
            else:
                # head: per output channel f = c % (5+nc): 0,1 xy | 2,3 wh | 4 obj | 5.. cls
                f = np.arange(o.cout) % (5 + num_classes)
                tgt_std = np.where((f == 2) | (f == 3), wh_std, np.where(f < 2, 1.0, logit_std))
                tgt_mean = np.where(f == 4, obj_bias, np.where(f >= 5, cls_bias, 0.0))
                sd = y.std(axis=(0, 1, 2)) + 1e-12
                scale = (tgt_std / sd)
                w32 = (w32.astype(np.float64) * scale).astype(np.float32)
                bias = (tgt_mean - y.mean(axis=(0, 1, 2)) * scale).astype(np.float32)
                p.update(w=w32, bias=bias)
                y = y * scale + bias.astype(np.float64)
            y = mish(y) if o.act == 'mish' else leaky(y) if o.act == 'leaky' else y
            t[o.out] = y
        else:
            t[o.out] = _apply_simple(o, t)
    return W


def _apply_simple(o, t):
    if o.kind == 'add':
        return t[o.ins[0]] + t[o.ins[1]]
    if o.kind == 'concat':
        return np.concatenate([t[i] for i in o.ins], axis=-1)
    if o.kind == 'maxpool':
        return maxpool_same(t[o.ins[0]], o.pool)
    if o.kind == 'upsample':
        return upsample2x(t[o.ins[0]])
    raise ValueError(o.kind)


# ----------------------------------------------------------------------------------------------
# forward: 110 convs -> 3 raw heads   (models.py:50-52 yolo_model)
# ----------------------------------------------------------------------------------------------
def forward(imgs, W: Weights, dtype=np.float32, keep=None, fold_bn=False, quant=None, kperm_seed=None):
    """imgs (B,S,S,3) in [0,1].  Returns [head_s, head_m, head_l] each (B,g,g,3*(5+nc)).
    keep: optional dict filled with every named intermediate (layer-by-layer parity).
    fold_bn=False follows the reference op order (conv -> BN -> act); fold_bn=True uses folded
    weights (what the engine computes) — used to measure the folding error itself.
    quant: optional callable applied to every conv output after activation/add (simulates fp16 storage).
    kperm_seed: sum every conv's K terms in a seeded random order (conv2d_raw): another round-off realisation."""
    ops, heads = build_netlist(W.num_classes)
    kperm = None if kperm_seed is None else np.random.default_rng(kperm_seed)
    t = {'img': np.asarray(imgs).astype(dtype)}          # Keras casts float64 input to float32
    live_until = {}
    for i, o in enumerate(ops):
        for n in o.ins:
            live_until[n] = i
    for i, o in enumerate(ops):
        if o.kind == 'conv':
            p = W.p[o.idx]
            x = t[o.ins[0]]
            if fold_bn:
                w, b = W.folded(o.idx, dtype)
                y = conv2d_raw(x, w, o.stride, kperm) + b
            else:
                y = conv2d_raw(x, p['w'].astype(dtype), o.stride, kperm)
                if o.bn:   # FusedBatchNorm inference: (x-mean)*gamma/sqrt(var+eps)+beta, in dtype
                    inv = (p['gamma'].astype(dtype) / np.sqrt(p['var'].astype(dtype) + dtype(BN_EPS))).astype(dtype)
                    y = ((y - p['mean'].astype(dtype)) * inv + p['beta'].astype(dtype)).astype(dtype)
                else:
                    y = (y + p['bias'].astype(dtype)).astype(dtype)
            y = mish(y) if o.act == 'mish' else leaky(y) if o.act == 'leaky' else y
            if quant is not None and o.bn:
                y = quant(y)
            t[o.out] = y
        else:
            y = _apply_simple(o, t)
            if quant is not None and o.kind == 'add':
                y = quant(y)
            t[o.out] = y
        if keep is not None:
            keep[o.out] = t[o.out]
        else:
            for n in o.ins:            # free dead intermediates
                if live_until.get(n) == i and n not in heads and n in t and n != 'img':
                    del t[n]
    return [t[h] for h in heads]


# ----------------------------------------------------------------------------------------------
# head decode  (custom_layers.py:201-258, generalised grid = S // stride)
# ----------------------------------------------------------------------------------------------
def get_boxes(pred, anchors, num_classes, grid_size, stride, xyscale):
    """custom_layers.py:221-258.  pred (B,g,g,3*(5+nc)) -> x1y1x2y2 (B,g,g,3,4) px, obj (..,1), cls (..,nc)."""
    f32 = np.float32
    B = pred.shape[0]
    pred = pred.astype(f32).reshape(B, grid_size, grid_size, 3, 5 + num_classes)
    box_xy, box_wh, obj, cls = pred[..., 0:2], pred[..., 2:4], pred[..., 4:5], pred[..., 5:]
    box_xy, obj, cls = sigmoid(box_xy), sigmoid(obj), sigmoid(cls)
    gx, gy = np.meshgrid(np.arange(grid_size), np.arange(grid_size))   # xy indexing: [...,0]=col, [...,1]=row
    grid = np.stack([gx, gy], axis=-1)[:, :, None, :].astype(f32)       # (g,g,1,2)
    xs, st = f32(xyscale), f32(stride)
    # :251  `0.5 * (xyscale - 1)` is Python float (double) arithmetic in the reference; TF casts the constant to float32
    box_xy = ((box_xy * xs) - f32(0.5 * (float(xyscale) - 1.0)) + grid) * st
    box_wh = np.exp(box_wh) * anchors.astype(f32)                       # :253
    x1y1 = box_xy - box_wh / f32(2)
    x2y2 = box_xy + box_wh / f32(2)
    return np.concatenate([x1y1, x2y2], axis=-1).astype(f32), obj, cls


def decode_heads(heads, img_size, num_classes=80, anchors=ANCHORS, strides=STRIDES, xyscale=XYSCALE):
    """yolov4_head + the flattening part of nms() (custom_layers.py:201-218, :270-284).
    Returns boxes (B,N,4) NORMALISED by img_size (x1,y1,x2,y2) and scores (B,N,nc) = obj*cls.
    Flat index n = off_scale + (row*g + col)*3 + a, scale order S, M, L."""
    bx, sc = [], []
    for i, h in enumerate(heads):
        g = img_size // strides[i]
        B = h.shape[0]
        b, o, c = get_boxes(h, anchors[i], num_classes, g, strides[i], xyscale[i])
        bx.append(b.reshape(B, -1, 4))
        sc.append((o * c).reshape(B, -1, num_classes))
    boxes = np.concatenate(bx, axis=1) / np.float32(img_size)
    return boxes.astype(np.float32), np.concatenate(sc, axis=1).astype(np.float32)


# ----------------------------------------------------------------------------------------------
# tf.image.combined_non_max_suppression  (custom_layers.py:290-297; semantics SURVEY App. D.8)
# ----------------------------------------------------------------------------------------------
def _iou(a, b):
    """TF's IOU on float32: coords re-ordered with min/max; area<=0 -> 0."""
    f32 = np.float32
    ya0, xa0, ya1, xa1 = min(a[0], a[2]), min(a[1], a[3]), max(a[0], a[2]), max(a[1], a[3])
    yb0, xb0, yb1, xb1 = min(b[0], b[2]), min(b[1], b[3]), max(b[0], b[2]), max(b[1], b[3])
    area_a = f32(f32(ya1 - ya0) * f32(xa1 - xa0))
    area_b = f32(f32(yb1 - yb0) * f32(xb1 - xb0))
    if area_a <= 0 or area_b <= 0:
        return f32(0)
    iy0, ix0, iy1, ix1 = max(ya0, yb0), max(xa0, xb0), min(ya1, yb1), min(xa1, xb1)
    inter = f32(max(f32(iy1 - iy0), f32(0)) * max(f32(ix1 - ix0), f32(0)))
    return f32(inter / f32(f32(area_a + area_b) - inter))


def combined_nms(boxes, scores, iou_threshold=IOU_THRESHOLD, score_threshold=SCORE_THRESHOLD,
                 max_per_class=MAX_BOXES, max_total=MAX_BOXES, margins=None):
    """boxes (B,N,4) normalised; scores (B,N,C).  Returns nmsed_boxes (B,T,4) clipped to [0,1],
    nmsed_scores (B,T), nmsed_classes (B,T) float32, valid (B,) int32, cand_idx (B,T) int32 (-1 pad).
    Order rules made explicit (TF leaves ties implementation-defined): within a class candidates are
    visited by (score desc, box index asc); the final merge sorts by (score desc, class asc, box asc).
    margins: optional dict collecting min |score-thr| and min |iou-thr| seen (test-data hygiene)."""
    f32 = np.float32
    B, N, C = scores.shape
    T = max_total
    ob = np.zeros((B, T, 4), f32); osc = np.zeros((B, T), f32); ocl = np.zeros((B, T), f32)
    ov = np.zeros((B,), np.int32); oidx = np.full((B, T), -1, np.int32)
    thr_s, thr_i = f32(score_threshold), f32(iou_threshold)
    for b in range(B):
        picked = []                                   # (score, class, box)
        bb = boxes[b]
        cand_n, cand_c = np.nonzero(scores[b] > thr_s)     # strict >
        if margins is not None and scores[b].size:
            margins['score'] = min(margins.get('score', 1e9), float(np.min(np.abs(scores[b] - thr_s))))
        for c in np.unique(cand_c):
            ns = cand_n[cand_c == c]
            s = scores[b, ns, c]
            order = np.lexsort((ns, -s.astype(np.float64)))     # score desc, then box idx asc
            sel = []
            for j in order:
                n = ns[j]
                keep = True
                for m in reversed(sel):                     # newest selected first, early exit
                    v = _iou(bb[n], bb[m])
                    if margins is not None:
                        margins['iou'] = min(margins.get('iou', 1e9), abs(float(v) - float(thr_i)))
                    if v > thr_i:                           # strict >
                        keep = False
                        break
                if keep:
                    sel.append(n)
                    picked.append((float(s[j]), int(c), int(n)))
                    if len(sel) >= max_per_class:
                        break
        picked.sort(key=lambda t: (-t[0], t[1], t[2]))
        picked = picked[:T]
        ov[b] = len(picked)
        for k, (s, c, n) in enumerate(picked):
            ob[b, k] = np.clip(bb[n], 0, 1)                 # clip_boxes=True, applied on output only
            osc[b, k] = s; ocl[b, k] = c; oidx[b, k] = n
    return ob, osc, ocl, ov, oidx


def decode_nms(heads, img_size, num_classes=80, iou_threshold=IOU_THRESHOLD,
               score_threshold=SCORE_THRESHOLD, margins=None):
    boxes, scores = decode_heads(heads, img_size, num_classes)
    return combined_nms(boxes, scores, iou_threshold, score_threshold, margins=margins)


def predict(imgs, W: Weights, margins=None, dtype=np.float32):
    """inference_model.predict (models.py:68-73,113): imgs (B,S,S,3) -> boxes, scores, classes, valid, cand_idx."""
    S = imgs.shape[1]
    return decode_nms(forward(imgs, W, dtype), S, W.num_classes, margins=margins)


# ----------------------------------------------------------------------------------------------
# host-side pieces of the API  (models.py:95-98, utils.py:56-78)
# ----------------------------------------------------------------------------------------------
def resize_linear_u8(img_u8, W, H):
    """cv2.resize(img, (W, H)) for 8-bit images, interpolation INTER_LINEAR (the default the reference uses,
    models.py:96), restated from OpenCV's fixed-point path (opencv/modules/imgproc/src/resize.cpp: resizeGeneric_ /
    HResizeLinear / VResizeLinear<uchar,int,short>, INTER_RESIZE_COEF_BITS = 11).  OpenCV is a third-party dependency of the
    reference (unpinned; 4.13.0 in this image); tests/test_oracle.py pins this restatement bit-exactly against cv2.resize on
    the reference's own images."""
    img = np.ascontiguousarray(img_u8, dtype=np.uint8)
    h, w = img.shape[:2]
    dx = np.arange(W, dtype=np.float64)
    fx = ((dx + 0.5) * (w / W) - 0.5).astype(np.float32)
    sx = np.floor(fx).astype(np.int64)
    fx = (fx - sx.astype(np.float32)).astype(np.float32)
    lo, hi = sx < 0, sx >= w - 1
    fx[lo] = 0; sx[lo] = 0
    fx[hi] = 0; sx[hi] = w - 1
    a1 = np.rint(fx * np.float32(2048)).astype(np.int64)                 # saturate_cast<short>: round half to even
    a0 = np.rint((np.float32(1) - fx) * np.float32(2048)).astype(np.int64)
    sx1 = np.minimum(sx + 1, w - 1)
    dy = np.arange(H, dtype=np.float64)
    fy = ((dy + 0.5) * (h / H) - 0.5).astype(np.float32)
    sy = np.floor(fy).astype(np.int64)
    fy = (fy - sy.astype(np.float32)).astype(np.float32)
    b1 = np.rint(fy * np.float32(2048)).astype(np.int64)
    b0 = np.rint((np.float32(1) - fy) * np.float32(2048)).astype(np.int64)
    sy0, sy1 = np.clip(sy, 0, h - 1), np.clip(sy + 1, 0, h - 1)
    src = img.astype(np.int64).reshape(h, w, -1)

    def hpass(rows):
        r = src[rows]
        return r[:, sx, :] * a0[None, :, None] + r[:, sx1, :] * a1[None, :, None]
    t0, t1 = hpass(sy0), hpass(sy1)
    out = (((b0[:, None, None] * (t0 >> 4)) >> 16) + ((b1[:, None, None] * (t1 >> 4)) >> 16) + 2) >> 2
    return np.clip(out, 0, 255).astype(np.uint8).reshape((H, W) + img.shape[2:])


def preprocess_img(img_u8, size):
    """models.py:95-98: cv2.resize(img, (W, H)) / 255.  -> float64 (Keras casts to float32 at predict)."""
    return resize_linear_u8(img_u8, size, size) / 255.0      # no letterbox: aspect ratio is not preserved


def detection_table(raw_hw, outputs, class_names):
    """utils.py:56-78 as plain lists: rows of (x1,y1,x2,y2,class_name,score,w,h); image 0 only."""
    boxes, scores, classes, valid = outputs[:4]
    n = int(valid[0]); h, w = raw_hw
    rows = []
    for i in range(n):
        x1, y1, x2, y2 = boxes[0, i]
        X1, X2 = int(np.int64(x1 * w)), int(np.int64(x2 * w))
        Y1, Y2 = int(np.int64(y1 * h)), int(np.int64(y2 * h))
        rows.append((X1, Y1, X2, Y2, class_names[int(classes[0, i])], float(scores[0, i]), X2 - X1, Y2 - Y1))
    return rows


# ----------------------------------------------------------------------------------------------
# synthetic inputs (BASELINE.json configs)
# ----------------------------------------------------------------------------------------------
def synth_images(seed, first_index, batch, size):
    """Image i depends only on (seed, global index i): results are independent of GPU count.
    Uses the same integer hash as the engine's device-side generator (y4_synth_fill)."""
    idx = (np.arange(first_index, first_index + batch, dtype=np.uint64)[:, None]
           * np.uint64(size * size * 3) + np.arange(size * size * 3, dtype=np.uint64)[None, :])
    return hash_uniform(idx, seed).reshape(batch, size, size, 3)


def hash_uniform(idx_u64, seed):
    """splitmix64-style hash -> float32 in [0,1) with 24 random bits (bit-identical on CPU and GPU)."""
    z = (idx_u64 + np.uint64(seed) * np.uint64(0x9E3779B97F4A7C15)) & np.uint64(0xFFFFFFFFFFFFFFFF)
    with np.errstate(over='ignore'):
        z = (z + np.uint64(0x9E3779B97F4A7C15))
        z = (z ^ (z >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)
        z = (z ^ (z >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)
        z = z ^ (z >> np.uint64(31))
    return ((z >> np.uint64(40)).astype(np.float32) * np.float32(1.0 / 16777216.0)).astype(np.float32)


def synth_heads(seed, batch, img_size, num_classes=80, n_clusters=150, per_cluster=7):
    """BASELINE config 4: head tensors with ~n_clusters*per_cluster candidates/img above 0.3.
    Background obj logit ~ N(-7,1); seeded cluster centres with neighbouring cells/anchors of the
    same class, obj & class logits U(1,6), xy ~ N(0,1), wh ~ N(0,0.4)."""
    rng = np.random.default_rng(seed)
    C = 5 + num_classes
    heads = []
    for s in STRIDES:
        g = img_size // s
        h = rng.standard_normal((batch, g, g, 3, C)).astype(np.float32)
        h[..., 2:4] *= 0.4
        h[..., 4] = h[..., 4] - 7.0
        h[..., 5:] = h[..., 5:] * 1.0 - 6.0
        heads.append(h)
    for b in range(batch):
        for _ in range(n_clusters):
            si = rng.integers(0, 3)
            g = img_size // STRIDES[si]
            r0, c0 = rng.integers(0, g, 2)
            cls = rng.integers(0, num_classes)
            for _ in range(per_cluster):
                r = int(np.clip(r0 + rng.integers(-1, 2), 0, g - 1))
                c = int(np.clip(c0 + rng.integers(-1, 2), 0, g - 1))
                a = rng.integers(0, 3)
                heads[si][b, r, c, a, 4] = rng.uniform(1, 6)
                heads[si][b, r, c, a, 5 + cls] = rng.uniform(1, 6)
    return [h.reshape(h.shape[0], h.shape[1], h.shape[2], 3 * C) for h in heads]
