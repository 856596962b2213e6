"""Inference configuration — same keys, same defaults as the reference's `yolo_config` (config.py:1-17).
Unlike the reference (models.py:26-37 reads the module-level dict even when another config is passed),
`Yolov4(config=...)` here honours the dict it is given; the defaults are identical."""

_ANCHOR_PAIRS = (
    (12, 16), (19, 36), (40, 28),        # stride 8
    (36, 75), (76, 55), (72, 146),       # stride 16
    (142, 110), (192, 243), (459, 401),  # stride 32
)

yolo_config = dict(
    img_size=(416, 416, 3),
    anchors=[v for pair in _ANCHOR_PAIRS for v in pair],
    strides=[8, 16, 32],
    xyscale=[1.2, 1.1, 1.05],
    # training-only keys of the reference, kept so that user code reading them does not break
    iou_loss_thresh=0.5,
    batch_size=8,
    num_gpu=1,
    # inference
    max_boxes=100,
    iou_threshold=0.413,
    score_threshold=0.3,
    # engine extensions (absent from the reference)
    precision='fp16',      # 'fp16' (tcgen05) | 'fp32' (CUDA-core parity mode)
    max_batch=32,
    device=0,
)
