"""B200-native YOLOv4 inference hot path behind the reference's call surface.

  binding.py   ctypes binding of liby4.so (include/y4.h) — the only place device work happens
  config.py    yolo_config              (reference: config.py)
  utils.py     load_weights, get_detection_data, draw_bbox   (reference: utils.py:12-118)
  models.py    class Yolov4             (reference: models.py:17-179, 509-529)
  evaluate.py  eval_map, voc_ap         (reference: models.py:182-507, utils.py:311-356)
  dp.py        image sharding helpers for one-process-per-GPU serving
The reference's custom_layers.py (graph builders, decode, NMS) has no Python counterpart: it IS the engine
(csrc/, reached through binding.Engine.forward_heads / decode_nms / predict).
"""
from .binding import Engine, Y4Error, lib_path, PREC_FP32, PREC_FP16, PREC_FP16_SIMT, PREC_FP16X3  # noqa: F401
from .config import yolo_config  # noqa: F401
from .models import Yolov4  # noqa: F401
