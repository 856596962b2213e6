"""B200-native YOLOv4 inference hot path behind the reference's call surface.

  binding.py   ctypes binding of liby4.so (include/y4.h) — the only place device work happens
  config.py    yolo_config              (reference: config.py)
  utils.py     load_weights, get_detection_data, draw_bbox   (reference: utils.py:12-118)
  models.py    class Yolov4             (reference: models.py:17-127, 509-529)
  custom_layers.py  yolov4_head / nms as engine calls (reference: custom_layers.py:201-298)
"""
from .binding import Engine, Y4Error, lib_path, PREC_FP32, PREC_FP16, PREC_FP16_SIMT, PREC_FP16X3  # noqa: F401
from .config import yolo_config  # noqa: F401
from .models import Yolov4  # noqa: F401
