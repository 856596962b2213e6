"""class Yolov4 — the reference's inference surface (models.py:17-127, 141-179, 509-529) over the B200 engine.

Kept: ctor signature, predict / predict_img / predict_raw / predict_nonms / preprocess_img /
export_prediction, the BGR/RGB behaviour of each method, and the DataFrame result shape.
Dropped (out of scope): fit, Keras save/load, training model.  `max_boxes` (config.py) sets both NMS sizes
(per class and total), which the reference hard-codes to 100 = its default max_boxes (custom_layers.py:293-294).
Generalised: grid = img_size // stride (the reference hard-codes 52/26/13, custom_layers.py:204-212).
"""
import os

import numpy as np

from .binding import Engine, PREC_FP16, PREC_FP16X3, PREC_FP16_SIMT, PREC_FP32
from .config import yolo_config
from .utils import draw_bbox, get_detection_data, load_weights


class Yolov4(object):
    def __init__(self, weight_path=None, class_name_path='coco_classes.txt', config=yolo_config):
        assert config['img_size'][0] == config['img_size'][1], 'not support yet'
        assert config['img_size'][0] % config['strides'][-1] == 0, 'must be a multiple of last stride'
        with open(class_name_path) as f:
            self.class_names = [line.strip() for line in f.readlines()]
        self.config = dict(config)
        self.img_size = tuple(config['img_size'])
        self.num_classes = len(self.class_names)
        assert self.num_classes > 0, 'no classes detected!'
        self.weight_path = weight_path
        self.anchors = np.array(config['anchors']).reshape((3, 3, 2))
        self.xyscale = config['xyscale']
        self.strides = config['strides']
        self.output_sizes = [self.img_size[0] // s for s in self.strides]
        self.class_color = {name: list(np.random.random(size=3) * 255) for name in self.class_names}
        self.max_boxes = config['max_boxes']
        self.build_model(load_pretrained=bool(self.weight_path))

    def build_model(self, load_pretrained=True):
        precisions = {'fp32': PREC_FP32, 'fp16': PREC_FP16, 'fp16x3': PREC_FP16X3, 'fp16_simt': PREC_FP16_SIMT}
        name = self.config.get('precision', 'fp16')
        assert name in precisions, f'unknown precision {name!r} (one of {sorted(precisions)})'
        prec = precisions[name]
        # one engine plays both Keras models: yolo_model (heads) and inference_model (heads + decode + NMS)
        self.engine = Engine(img_size=self.img_size[0], num_classes=self.num_classes,
                             max_batch=int(self.config.get('max_batch', 32)), precision=prec,
                             device=int(self.config.get('device', 0)), anchors=self.config['anchors'],
                             strides=self.strides, xyscale=self.xyscale, max_boxes=self.max_boxes,
                             iou_threshold=self.config['iou_threshold'], score_threshold=self.config['score_threshold'])
        self.yolo_model = self.inference_model = self.engine
        print(f"nms iou: {self.config['iou_threshold']} score: {self.config['score_threshold']}")
        if load_pretrained and self.weight_path and self.weight_path.endswith('.weights'):
            load_weights(self.yolo_model, self.weight_path)
            print(f'load from {self.weight_path}')

    def preprocess_img(self, img):
        import cv2
        img = cv2.resize(img, self.img_size[:2])     # bilinear, aspect ratio NOT preserved (models.py:96)
        return img / 255.

    def _predict_batches(self, imgs):
        """inference_model.predict(imgs): Keras splits into batches (default 32); here max_batch."""
        mb = self.engine.max_batch
        parts = [self.engine.predict(imgs[i:i + mb]) for i in range(0, len(imgs), mb)]
        return [np.concatenate([p[k] for p in parts], axis=0) for k in range(4)]

    def _predict_raw_u8(self, raws, reverse_channels=False):
        """preprocess_img + inference_model.predict for raw uint8 images, both on the GPU (y4_predict_u8)."""
        mb = self.engine.max_batch
        parts = [self.engine.predict_u8(raws[i:i + mb], reverse_channels=reverse_channels) for i in range(0, len(raws), mb)]
        return [np.concatenate([p[k] for p in parts], axis=0) for k in range(4)]

    # raw_img: RGB
    def predict_img(self, raw_img, random_color=True, plot_img=True, figsize=(10, 10), show_text=True, return_output=False):
        print('img shape: ', raw_img.shape)
        if raw_img.dtype == np.uint8:                    # the normal case (cv2.imread): resize + /255 run on the GPU
            pred_output = self._predict_raw_u8([raw_img])
        else:
            img = self.preprocess_img(raw_img)
            pred_output = self._predict_batches(np.expand_dims(img, axis=0))
        detections = get_detection_data(img=raw_img, model_outputs=pred_output, class_names=self.class_names)
        output_img = draw_bbox(raw_img, detections, cmap=self.class_color, random_color=random_color,
                               figsize=figsize, show_text=show_text, show_img=plot_img)
        return (output_img, detections) if return_output else detections

    def predict(self, img_path, random_color=True, plot_img=True, figsize=(10, 10), show_text=True):
        import cv2
        raw_img = cv2.imread(img_path)[:, :, ::-1]       # BGR -> RGB (models.py:126)
        return self.predict_img(raw_img, random_color, plot_img, figsize, show_text)

    def predict_raw(self, img_path):
        import cv2
        raw_img = cv2.imread(img_path)                    # no RGB flip, as in the reference (models.py:510)
        print('img shape: ', raw_img.shape)
        imgs = np.expand_dims(self.preprocess_img(raw_img), axis=0)
        return self.engine.forward_heads(imgs)

    def predict_nonms(self, img_path, iou_threshold=0.413, score_threshold=0.1):
        import cv2
        raw_img = cv2.imread(img_path)
        print('img shape: ', raw_img.shape)
        imgs = np.expand_dims(self.preprocess_img(raw_img), axis=0)
        heads = self.engine.forward_heads(imgs)
        print(f'nms iou: {iou_threshold} score: {score_threshold}')
        pred_output = self.engine.decode_nms(heads, iou_threshold, score_threshold)
        detections = get_detection_data(img=raw_img, model_outputs=pred_output, class_names=self.class_names)
        draw_bbox(raw_img, detections, cmap=self.class_color, random_color=True)
        return detections

    def export_gt(self, annotation_path, gt_folder_path):
        """models.py:129-139: annotation lines `<img> x1,y1,x2,y2,cls ...` -> one `<class> <x1> <y1> <x2> <y2>` file per image
        (the ground-truth side of eval_map; host only)."""
        with open(annotation_path) as file:
            for line in file:
                parts = line.split(' ')
                stem = parts[0].split(os.sep)[-1].split('.')[0]
                with open(os.path.join(gt_folder_path, stem + '.txt'), 'w') as out:
                    for obj in parts[1:]:
                        x_min, y_min, x_max, y_max, class_id = [float(o) for o in obj.strip().split(',')]
                        out.write(f'{self.class_names[int(class_id)]} {x_min} {y_min} {x_max} {y_max}\n')

    def export_prediction(self, annotation_path, pred_folder_path, img_folder_path, bs=2):
        """Batched caller (models.py:141-179): `<class> <score> <x1> <y1> <x2> <y2>` per detection, raw-image px.
        Pipelined: while the GPU works on batch i (H2D of the raw bytes, resize, forward, decode, NMS), the host decodes the
        files of batch i+1 and writes the text files of batch i-1 (y4_submit_u8 / y4_collect, two batches in flight)."""
        import cv2
        with open(annotation_path) as file:
            img_paths = [os.path.join(img_folder_path, line.split(' ')[0].split(os.sep)[-1]) for line in file]
        bs = max(1, min(int(bs), self.engine.max_batch))

        def write(paths, raws, outs):
            b_boxes, b_scores, b_classes, b_valid = outs[:4]
            for k, path in enumerate(paths):
                n = int(b_valid[k])
                boxes = b_boxes[k, :n].copy()
                boxes[:, [0, 2]] *= raws[k].shape[1]
                boxes[:, [1, 3]] *= raws[k].shape[0]
                stem = path.split(os.sep)[-1].split('.')[0]
                with open(os.path.join(pred_folder_path, stem + '.txt'), 'w') as out:
                    for i in range(n):
                        bx = boxes[i]
                        out.write(f'{self.class_names[int(b_classes[k, i])]} {b_scores[k, i]} {bx[0]} {bx[1]} {bx[2]} {bx[3]}\n')

        pending = []
        for start in range(0, len(img_paths), bs):
            paths = img_paths[start:start + bs]
            raws = [cv2.imread(path) for path in paths]   # BGR kept (no flip), models.py:153
            self.engine.submit_u8(raws, reverse_channels=False)
            pending.append((paths, raws))
            if len(pending) == 2:
                p0 = pending.pop(0)
                write(p0[0], p0[1], self.engine.collect())
        while pending:
            p0 = pending.pop(0)
            write(p0[0], p0[1], self.engine.collect())

    def eval_map(self, gt_folder_path, pred_folder_path, temp_json_folder_path, output_files_path):
        """models.py:182-507 without the per-detection JSON round trips and plots; same arguments, same output.txt."""
        from .evaluate import eval_map
        return eval_map(gt_folder_path, pred_folder_path, temp_json_folder_path, output_files_path)
