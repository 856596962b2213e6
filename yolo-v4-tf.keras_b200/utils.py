"""Host-side helpers with the reference's names and argument meaning (reference: utils.py:12-118).

load_weights(model, path)          -> feeds the darknet file to the engine (BN folded, eps 1e-3)   utils.py:12-53
get_detection_data(img, outs, names)-> DataFrame[x1,y1,x2,y2,class_name,score,w,h], image 0 only  utils.py:56-78
draw_bbox(...)                     -> cv2 rectangles + labels; matplotlib only if show_img         utils.py:88-118
"""
import numpy as np

from .evaluate import voc_ap  # noqa: F401  (utils.py:311-356)


def load_weights(model, weights_file_path):
    """`model` is a binding.Engine (the stand-in for the reference's Keras yolo_model).  The engine checks the
    exact byte count the way utils.py:50-53 reports unread weights, but fails instead of printing."""
    model.load_darknet(weights_file_path)
    print('all weights read')


def get_detection_data(img, model_outputs, class_names):
    import pandas as pd
    boxes, scores, classes, valid = model_outputs[:4]
    count = int(valid[0])
    h, w = img.shape[:2]
    b = np.asarray(boxes[0][:count], dtype=np.float32)
    # (x * w).astype('int64'): float32 box times python int -> float32 product, truncated toward zero
    xs = (b[:, [0, 2]] * w).astype('int64')
    ys = (b[:, [1, 3]] * h).astype('int64')
    names = np.array(class_names)[np.asarray(classes[0][:count]).astype('int64')] if count else np.array([], dtype=object)
    df = pd.DataFrame({'x1': xs[:, 0], 'y1': ys[:, 0], 'x2': xs[:, 1], 'y2': ys[:, 1],
                       'class_name': names, 'score': np.asarray(scores[0][:count], dtype=np.float32)})
    df['w'] = df['x2'] - df['x1']
    df['h'] = df['y2'] - df['y1']
    print(f'# of bboxes: {count}')
    return df


def draw_bbox(img, detections, cmap, random_color=True, figsize=(10, 10), show_img=True, show_text=True):
    import cv2
    canvas = np.array(img)
    scale = max(canvas.shape[0:2]) / 416
    line_width = int(2 * scale)
    font = cv2.FONT_HERSHEY_DUPLEX
    font_scale = max(0.3 * scale, 0.3)
    thickness = max(int(1 * scale), 1)
    rng = np.random.default_rng()
    for x1, y1, x2, y2, name, score in zip(detections['x1'], detections['y1'], detections['x2'], detections['y2'],
                                           detections['class_name'], detections['score']):
        x1, y1, x2, y2 = int(x1), int(y1), int(x2), int(y2)
        color = [float(v) for v in (rng.random(3) * 255 if random_color else cmap[name])]
        cv2.rectangle(canvas, (x1, y1), (x2, y2), color, line_width)
        if show_text:
            label = f'{name} {score:.2f}'
            (tw, th), _ = cv2.getTextSize(label, font, fontScale=font_scale, thickness=thickness)
            cv2.rectangle(canvas, (x1 - line_width // 2, y1 - th), (x1 + tw, y1), color, cv2.FILLED)
            cv2.putText(canvas, label, (x1, y1), font, font_scale, (255, 255, 255), thickness, cv2.LINE_AA)
    if show_img:
        try:
            import matplotlib.pyplot as plt  # lazy: matplotlib is optional in this image
        except ImportError:
            print('draw_bbox: matplotlib is not installed, the annotated image is returned but not shown')
            return canvas
        plt.figure(figsize=figsize)
        plt.imshow(canvas)
        plt.show()
    return canvas
