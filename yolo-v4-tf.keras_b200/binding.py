"""ctypes binding of liby4.so (C-ABI declared in include/y4.h).  No CPU fallback: if the shared library is
missing, or no sm_100 device is usable, the calls raise."""
import ctypes as C
import os

import numpy as np

PREC_FP32, PREC_FP16, PREC_FP16_SIMT, PREC_FP16X3 = 0, 1, 2, 3

_HERE = os.path.dirname(os.path.abspath(__file__))


def lib_path():
    return os.path.join(_HERE, 'liby4.so')


class Y4Error(RuntimeError):
    def __init__(self, code, msg):
        super().__init__(f'liby4 error {code}: {msg}')
        self.code = code


class Y4Config(C.Structure):
    _fields_ = [('img_size', C.c_int32), ('num_classes', C.c_int32), ('max_batch', C.c_int32),
                ('precision', C.c_int32), ('device', C.c_int32), ('max_boxes', C.c_int32),
                ('strides', C.c_int32 * 3), ('reserved_', C.c_int32),
                ('anchors', C.c_float * 18), ('xyscale', C.c_double * 3),
                ('iou_threshold', C.c_float), ('score_threshold', C.c_float)]


class Y4LayerInfo(C.Structure):
    _fields_ = [(n, C.c_int32) for n in ('idx', 'cin', 'cout', 'ksize', 'stride', 'batch_norm', 'activation',
                                         'out_hw', 'kernel_kind', 'tile_n')] + [('flops', C.c_int64), ('out_name', C.c_char * 16)] + \
               [(n, C.c_int32) for n in ('tc_mode', 'tc_epilogue', 'tc_stages', 'tc_group', 'tc_ctas_per_sm', 'tc_bk', 'tc_epi_warps', 'tc_resident_w')]


EXPORTS = [
    'y4_default_config', 'y4_create', 'y4_destroy', 'y4_last_error', 'y4_load_darknet',
    'y4_load_darknet_from_memory', 'y4_predict', 'y4_predict_u8', 'y4_preprocess_u8', 'y4_submit', 'y4_submit_u8', 'y4_collect', 'y4_forward_heads', 'y4_decode_nms', 'y4_synth_fill',
    'y4_run_resident', 'y4_run_forward_resident', 'y4_run_decode_nms_resident', 'y4_upload_heads',
    'y4_fetch_results', 'y4_sync', 'y4_timer_begin', 'y4_timer_end', 'y4_flush_l2', 'y4_launch_count',
    'y4_profile_layers', 'y4_host_alloc', 'y4_host_free', 'y4_num_layers', 'y4_describe_layer', 'y4_num_steps', 'y4_describe_step', 'y4_num_boxes',
    'y4_debug_get_tensor', 'y4_debug_run_conv', 'y4_debug_trace_conv', 'y4_comm_unique_id', 'y4_comm_init', 'y4_allgather_results',
]

_lib = None


def load_library():
    """dlopen liby4.so once and declare argument types.  Raises if the extension has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    path = lib_path()
    if not os.path.exists(path):
        raise ImportError(f'{path} not found: build it with `python -c "import __graft_entry__ as g; g.build()"` '
                          '(nvcc, sm_100a). There is no CPU fallback.')
    lib = C.CDLL(path)
    fp, ip, vp = C.POINTER(C.c_float), C.POINTER(C.c_int32), C.c_void_p
    lib.y4_default_config.argtypes = [C.POINTER(Y4Config)]
    lib.y4_create.argtypes = [C.POINTER(vp), C.POINTER(Y4Config)]
    lib.y4_destroy.argtypes = [vp]; lib.y4_destroy.restype = None
    lib.y4_last_error.argtypes = [vp]; lib.y4_last_error.restype = C.c_char_p
    lib.y4_load_darknet.argtypes = [vp, C.c_char_p]
    lib.y4_load_darknet_from_memory.argtypes = [vp, vp, C.c_size_t]
    lib.y4_predict.argtypes = [vp, vp, C.c_int32, vp, vp, vp, vp, vp]
    lib.y4_submit.argtypes = [vp, vp, C.c_int32]
    lib.y4_predict_u8.argtypes = [vp, vp, vp, vp, C.c_int32, C.c_int32, vp, vp, vp, vp, vp]
    lib.y4_preprocess_u8.argtypes = [vp, vp, vp, vp, C.c_int32, C.c_int32, vp]
    lib.y4_submit_u8.argtypes = [vp, vp, vp, vp, C.c_int32, C.c_int32]
    lib.y4_collect.argtypes = [vp, C.c_int32, vp, vp, vp, vp, vp]
    lib.y4_forward_heads.argtypes = [vp, vp, C.c_int32, vp, vp, vp]
    lib.y4_decode_nms.argtypes = [vp, vp, vp, vp, C.c_int32, C.c_float, C.c_float, vp, vp, vp, vp, vp]
    lib.y4_synth_fill.argtypes = [vp, C.c_uint64, C.c_int64, C.c_int32]
    for n in ('y4_run_resident', 'y4_run_forward_resident', 'y4_run_decode_nms_resident'):
        getattr(lib, n).argtypes = [vp, C.c_int32]
    lib.y4_upload_heads.argtypes = [vp, vp, vp, vp, C.c_int32]
    lib.y4_fetch_results.argtypes = [vp, C.c_int32, vp, vp, vp, vp, vp]
    lib.y4_sync.argtypes = [vp]
    lib.y4_timer_begin.argtypes = [vp]
    lib.y4_timer_end.argtypes = [vp, fp]
    lib.y4_flush_l2.argtypes = [vp]
    lib.y4_launch_count.argtypes = [vp]; lib.y4_launch_count.restype = C.c_int64
    lib.y4_profile_layers.argtypes = [vp, C.c_int32, vp, C.c_int32]
    lib.y4_host_alloc.argtypes = [C.c_size_t]; lib.y4_host_alloc.restype = vp
    lib.y4_host_free.argtypes = [vp]; lib.y4_host_free.restype = None
    lib.y4_num_layers.argtypes = [vp]
    lib.y4_describe_layer.argtypes = [vp, C.c_int32, C.POINTER(Y4LayerInfo)]
    lib.y4_num_steps.argtypes = [vp]
    lib.y4_describe_step.argtypes = [vp, C.c_int32, C.POINTER(Y4LayerInfo)]
    lib.y4_num_boxes.argtypes = [vp]; lib.y4_num_boxes.restype = C.c_int64
    lib.y4_debug_get_tensor.argtypes = [vp, C.c_char_p, C.c_int32, vp, C.c_int64]
    lib.y4_debug_get_tensor.restype = C.c_int64
    lib.y4_debug_run_conv.argtypes = [vp, C.c_int32, C.c_int32, C.c_int32]
    lib.y4_debug_trace_conv.argtypes = [vp, C.c_int32, C.c_int32, vp, C.c_int32]
    lib.y4_comm_unique_id.argtypes = [vp]
    lib.y4_comm_init.argtypes = [vp, C.c_int32, C.c_int32, vp]
    lib.y4_allgather_results.argtypes = [vp, C.c_int32, vp, vp, vp, vp, vp]
    _lib = lib
    return lib


def _ptr(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


def pinned_array(shape, dtype=np.float32):
    """numpy array over cudaHostAlloc'ed memory (for async H2D/D2H); keep a reference to .base alive."""
    lib = load_library()
    n = int(np.prod(shape)) * np.dtype(dtype).itemsize
    p = lib.y4_host_alloc(n)
    if not p:
        raise Y4Error(-2, 'cudaHostAlloc failed')
    buf = (C.c_char * n).from_address(p)
    arr = np.frombuffer(buf, dtype=dtype).reshape(shape)
    return arr


class Engine:
    """One engine = one GPU + one stream.  Mirrors the Keras objects the reference's Yolov4 holds:
    predict() == inference_model.predict, forward_heads() == yolo_model.predict,
    decode_nms() == yolov4_head + nms with runtime thresholds."""

    def __init__(self, img_size=416, num_classes=80, max_batch=1, precision=PREC_FP16, device=0,
                 anchors=None, strides=None, xyscale=None, max_boxes=None, iou_threshold=None, score_threshold=None):
        self._lib = load_library()
        self._h = C.c_void_p()
        cfg = Y4Config()
        self._lib.y4_default_config(C.byref(cfg))
        cfg.img_size, cfg.num_classes, cfg.max_batch, cfg.precision, cfg.device = img_size, num_classes, max_batch, precision, device
        if anchors is not None:
            cfg.anchors = (C.c_float * 18)(*[float(a) for a in np.asarray(anchors).reshape(-1)])
        if strides is not None:
            cfg.strides = (C.c_int32 * 3)(*[int(s) for s in strides])
        if xyscale is not None:
            cfg.xyscale = (C.c_double * 3)(*[float(s) for s in xyscale])
        if max_boxes is not None:
            cfg.max_boxes = int(max_boxes)
        if iou_threshold is not None:
            cfg.iou_threshold = float(iou_threshold)
        if score_threshold is not None:
            cfg.score_threshold = float(score_threshold)
        self.cfg = cfg
        rc = self._lib.y4_create(C.byref(self._h), C.byref(cfg))
        if rc != 0:
            msg = self._lib.y4_last_error(None).decode()
            self._h = C.c_void_p()
            raise Y4Error(rc, msg)
        self.img_size, self.num_classes, self.max_batch, self.max_boxes = img_size, num_classes, max_batch, cfg.max_boxes
        self.grids = [img_size // s for s in cfg.strides]
        self.num_boxes = int(self._lib.y4_num_boxes(self._h))

    def close(self):
        if getattr(self, '_h', None) is not None and self._h:
            self._lib.y4_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _chk(self, rc):
        if rc < 0:
            raise Y4Error(rc, self._lib.y4_last_error(self._h).decode())
        return rc

    # ---- weights (utils.py:12-53)
    def load_darknet(self, path):
        self._chk(self._lib.y4_load_darknet(self._h, os.fsencode(path)))

    def load_darknet_bytes(self, data: bytes):
        buf = np.frombuffer(data, dtype=np.uint8)
        self._chk(self._lib.y4_load_darknet_from_memory(self._h, _ptr(buf), buf.size))

    # ---- outputs
    def _alloc_out(self, batch):
        mb = self.max_boxes
        return (np.zeros((batch, mb, 4), np.float32), np.zeros((batch, mb), np.float32),
                np.zeros((batch, mb), np.float32), np.zeros((batch,), np.int32), np.zeros((batch, mb), np.int32))

    def _imgs(self, imgs):
        imgs = np.ascontiguousarray(imgs, dtype=np.float32)       # Keras casts float64 -> float32 (models.py:113)
        S = self.img_size
        if imgs.ndim != 4 or imgs.shape[1:] != (S, S, 3):
            raise ValueError(f'imgs must be (batch,{S},{S},3), got {imgs.shape}')
        return imgs

    def predict(self, imgs, with_indices=False):
        imgs = self._imgs(imgs)
        b = imgs.shape[0]
        boxes, scores, classes, valid, idx = self._alloc_out(b)
        self._chk(self._lib.y4_predict(self._h, _ptr(imgs), b, _ptr(boxes), _ptr(scores), _ptr(classes), _ptr(valid), _ptr(idx)))
        return (boxes, scores, classes, valid, idx) if with_indices else [boxes, scores, classes, valid]

    # ---- raw 8-bit images (GPU preprocess: cv2.resize INTER_LINEAR semantics + /255)
    def _u8_table(self, imgs):
        imgs = [np.ascontiguousarray(im, dtype=np.uint8) for im in imgs]
        for im in imgs:
            if im.ndim != 3 or im.shape[2] != 3:
                raise ValueError(f'raw images must be (h, w, 3) uint8, got {im.shape}')
        n = len(imgs)
        ptrs = (C.c_void_p * n)(*[im.ctypes.data for im in imgs])
        hs = np.array([im.shape[0] for im in imgs], np.int32)
        ws = np.array([im.shape[1] for im in imgs], np.int32)
        return imgs, ptrs, hs, ws

    def preprocess_u8(self, imgs, reverse_channels=False):
        """(batch, S, S, 3) float32 network input the engine builds from raw uint8 images (== preprocess_img, models.py:95-98)."""
        imgs, ptrs, hs, ws = self._u8_table(imgs)
        out = np.zeros((len(imgs), self.img_size, self.img_size, 3), np.float32)
        self._chk(self._lib.y4_preprocess_u8(self._h, ptrs, _ptr(hs), _ptr(ws), len(imgs), int(reverse_channels), _ptr(out)))
        return out

    def predict_u8(self, imgs, reverse_channels=False, with_indices=False):
        imgs, ptrs, hs, ws = self._u8_table(imgs)
        b = len(imgs)
        boxes, scores, classes, valid, idx = self._alloc_out(b)
        self._chk(self._lib.y4_predict_u8(self._h, ptrs, _ptr(hs), _ptr(ws), b, int(reverse_channels),
                                          _ptr(boxes), _ptr(scores), _ptr(classes), _ptr(valid), _ptr(idx)))
        return (boxes, scores, classes, valid, idx) if with_indices else [boxes, scores, classes, valid]

    def submit_u8(self, imgs, reverse_channels=False):
        """Pipelined form of predict_u8 (keep the arrays alive until the matching collect())."""
        imgs, ptrs, hs, ws = self._u8_table(imgs)
        self._chk(self._lib.y4_submit_u8(self._h, ptrs, _ptr(hs), _ptr(ws), len(imgs), int(reverse_channels)))
        self._inflight = getattr(self, '_inflight', []) + [np.empty((len(imgs), 0))]
        self._keep = getattr(self, '_keep', [])[-4:] + [(imgs, ptrs, hs, ws)]

    def submit(self, imgs):
        """Pipelined path: enqueue one batch (keep `imgs` alive and unmodified until the matching collect())."""
        imgs = self._imgs(imgs)
        self._chk(self._lib.y4_submit(self._h, _ptr(imgs), imgs.shape[0]))
        self._inflight = getattr(self, '_inflight', []) + [imgs]

    def collect(self, with_indices=False):
        inflight = getattr(self, '_inflight', [])
        if not inflight:                      # nothing submitted: let the engine report it (Y4_ERR_ARG -> Y4Error)
            boxes, scores, classes, valid, idx = self._alloc_out(1)
            self._chk(self._lib.y4_collect(self._h, 1, _ptr(boxes), _ptr(scores), _ptr(classes), _ptr(valid), _ptr(idx)))
            raise Y4Error(-4, 'collect() without a matching submit()')
        imgs = inflight.pop(0)
        b = imgs.shape[0]
        boxes, scores, classes, valid, idx = self._alloc_out(b)
        self._chk(self._lib.y4_collect(self._h, b, _ptr(boxes), _ptr(scores), _ptr(classes), _ptr(valid), _ptr(idx)))
        return (boxes, scores, classes, valid, idx) if with_indices else [boxes, scores, classes, valid]

    def forward_heads(self, imgs):
        imgs = self._imgs(imgs)
        b = imgs.shape[0]
        ch = 3 * (5 + self.num_classes)
        heads = [np.zeros((b, g, g, ch), np.float32) for g in self.grids]
        self._chk(self._lib.y4_forward_heads(self._h, _ptr(imgs), b, *[_ptr(h) for h in heads]))
        return heads

    def decode_nms(self, heads, iou_threshold=None, score_threshold=None, with_indices=False):
        heads = [np.ascontiguousarray(h, dtype=np.float32) for h in heads]
        b = heads[0].shape[0]
        iou = self.cfg.iou_threshold if iou_threshold is None else iou_threshold
        sc = self.cfg.score_threshold if score_threshold is None else score_threshold
        boxes, scores, classes, valid, idx = self._alloc_out(b)
        self._chk(self._lib.y4_decode_nms(self._h, *[_ptr(h) for h in heads], b, iou, sc,
                                          _ptr(boxes), _ptr(scores), _ptr(classes), _ptr(valid), _ptr(idx)))
        return (boxes, scores, classes, valid, idx) if with_indices else [boxes, scores, classes, valid]

    # ---- device-resident path
    def synth_fill(self, seed, first_index, batch):
        self._chk(self._lib.y4_synth_fill(self._h, seed, first_index, batch))

    def run_resident(self, batch):
        self._chk(self._lib.y4_run_resident(self._h, batch))

    def run_forward_resident(self, batch):
        self._chk(self._lib.y4_run_forward_resident(self._h, batch))

    def run_decode_nms_resident(self, batch):
        self._chk(self._lib.y4_run_decode_nms_resident(self._h, batch))

    def upload_heads(self, heads):
        heads = [np.ascontiguousarray(h, dtype=np.float32) for h in heads]
        self._chk(self._lib.y4_upload_heads(self._h, *[_ptr(h) for h in heads], heads[0].shape[0]))

    def fetch_results(self, batch):
        out = self._alloc_out(batch)
        self._chk(self._lib.y4_fetch_results(self._h, batch, *[_ptr(o) for o in out]))
        return out

    def sync(self):
        self._chk(self._lib.y4_sync(self._h))

    def timer_begin(self):
        self._chk(self._lib.y4_timer_begin(self._h))

    def timer_end(self):
        ms = C.c_float()
        self._chk(self._lib.y4_timer_end(self._h, C.byref(ms)))
        return ms.value

    def flush_l2(self):
        self._chk(self._lib.y4_flush_l2(self._h))

    def launch_count(self):
        return int(self._lib.y4_launch_count(self._h))

    def profile_layers(self, batch):
        ms = np.zeros(256, np.float32)
        n = self._chk(self._lib.y4_profile_layers(self._h, batch, _ptr(ms), ms.size))
        return ms[:n].copy()

    # ---- introspection
    def layers(self):
        out = []
        for i in range(self._lib.y4_num_layers(self._h)):
            li = Y4LayerInfo()
            self._chk(self._lib.y4_describe_layer(self._h, i, C.byref(li)))
            d = {n: getattr(li, n) for n, _ in Y4LayerInfo._fields_}
            d['out_name'] = d['out_name'].decode()
            out.append(d)
        return out

    def steps(self):
        """The launch schedule of one forward, in order (what profile_layers() times): convs, fused sibling pairs, SPP."""
        out = []
        for i in range(self._lib.y4_num_steps(self._h)):
            li = Y4LayerInfo()
            self._chk(self._lib.y4_describe_step(self._h, i, C.byref(li)))
            d = {n: getattr(li, n) for n, _ in Y4LayerInfo._fields_}
            d['out_name'] = d['out_name'].decode()
            out.append(d)
        return out

    def get_tensor(self, name, batch):
        n = self._chk(self._lib.y4_debug_get_tensor(self._h, name.encode(), batch, None, 0))
        flat = np.zeros(n, np.float32)
        self._chk(self._lib.y4_debug_get_tensor(self._h, name.encode(), batch, _ptr(flat), n))
        return flat

    def run_conv(self, idx, batch, use_tc):
        self._chk(self._lib.y4_debug_run_conv(self._h, idx, batch, int(use_tc)))

    def trace_conv(self, idx, batch, max_ctas=4096):
        out = np.zeros((max_ctas, 16), np.int64)
        self._chk(self._lib.y4_debug_trace_conv(self._h, idx, batch, _ptr(out), max_ctas))
        return out

    # ---- multi-GPU
    def comm_unique_id(self):
        uid = np.zeros(128, np.uint8)
        self._chk(self._lib.y4_comm_unique_id(_ptr(uid)))
        return uid

    def comm_init(self, rank, nranks, uid):
        uid = np.ascontiguousarray(uid, dtype=np.uint8)
        self._chk(self._lib.y4_comm_init(self._h, rank, nranks, _ptr(uid)))
        self.nranks = nranks

    def allgather_results(self, batch, fetch=True):
        if not fetch:        # enqueue the NCCL all-gather only (device-resident, async on the engine stream)
            self._chk(self._lib.y4_allgather_results(self._h, batch, None, None, None, None, None))
            return None
        mb, R = self.max_boxes, self.nranks
        out = (np.zeros((R * batch, mb, 4), np.float32), np.zeros((R * batch, mb), np.float32),
               np.zeros((R * batch, mb), np.float32), np.zeros((R * batch,), np.int32), np.zeros((R * batch, mb), np.int32))
        self._chk(self._lib.y4_allgather_results(self._h, batch, *[_ptr(o) for o in out]))
        return out
