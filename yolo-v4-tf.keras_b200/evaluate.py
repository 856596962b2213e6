"""mAP evaluation of exported predictions -- the reference's `Yolov4.eval_map` (models.py:182-507) and `voc_ap`
(utils.py:311-356) as array code: no per-detection JSON files re-read from disk, no matplotlib requirement.

Semantics kept exactly (tests/test_host.py pins them against the reference's own functions, run from its source):
  * ground truth `<class> <left> <top> <right> <bottom>` per line, predictions `<class> <conf> <l> <t> <r> <b>`;
  * classes = alphabetically sorted ground-truth classes; detections of a class from all files (files sorted by path),
    ordered by decreasing confidence with a STABLE sort (ties keep file / line order, as list.sort does);
  * a detection matches the ground-truth box of its file and class with the largest IoU (pixel-inclusive: +1 on widths and
    heights; the first maximum wins), true positive if IoU >= 0.5 and the box is still unused, false positive otherwise;
  * AP = area under the monotone precision envelope (VOC2012), mAP = mean over ground-truth classes;
  * `output.txt` has the same lines the reference writes.
"""
import os
from glob import glob

import numpy as np


def voc_ap(rec, prec):
    """utils.py:311-356.  Returns (ap, mrec, mpre) with mrec / mpre as lists, like the reference."""
    mrec = np.concatenate([[0.0], np.asarray(rec, dtype=np.float64), [1.0]])
    mpre = np.concatenate([[0.0], np.asarray(prec, dtype=np.float64), [0.0]])
    mpre = np.maximum.accumulate(mpre[::-1])[::-1]
    idx = np.nonzero(mrec[1:] != mrec[:-1])[0] + 1
    ap = 0.0
    for i in idx:                                   # same left-to-right accumulation order as the reference's loop
        ap += (mrec[i] - mrec[i - 1]) * mpre[i]
    return float(ap), mrec.tolist(), mpre.tolist()


def _read_lines(path):
    with open(path) as f:
        return [x.strip() for x in f.readlines()]


def eval_map(gt_folder_path, pred_folder_path, temp_json_folder_path=None, output_files_path=None, min_overlap=0.5):
    """Returns {'ap': {class: ap}, 'mAP': float, 'gt_counter_per_class': {...}, 'det_counter_per_class': {...},
    'count_true_positives': {...}}; writes <output_files_path>/output.txt when a path is given."""
    gt_files = sorted(glob(gt_folder_path + '/*.txt'))
    assert len(gt_files) > 0, 'no ground truth file'
    gt = {}                                         # file_id -> (classes list, boxes (n,4) float64)
    gt_counter, img_counter = {}, {}
    for txt in gt_files:
        file_id = os.path.basename(os.path.normpath(txt.split('.txt', 1)[0]))
        pred_path = os.path.join(pred_folder_path, file_id + '.txt')
        assert os.path.exists(pred_path), 'Error. File not found: {}\n'.format(pred_path)
        names, boxes = [], []
        for line in _read_lines(txt):
            class_name, left, top, right, bottom = line.split()
            names.append(class_name)
            boxes.append([float(left), float(top), float(right), float(bottom)])
            gt_counter[class_name] = gt_counter.get(class_name, 0) + 1
        for c in dict.fromkeys(names):
            img_counter[c] = img_counter.get(c, 0) + 1
        gt[file_id] = (names, np.asarray(boxes, dtype=np.float64).reshape(-1, 4))
    gt_classes = sorted(gt_counter.keys())

    dr_files = sorted(glob(os.path.join(pred_folder_path, '*.txt')))
    det = {}                                        # class -> list of (confidence, file_id, box)
    det_counter = {}
    for txt in dr_files:
        file_id = os.path.basename(os.path.normpath(txt.split('.txt', 1)[0]))
        for line in _read_lines(txt):
            if not line:
                continue
            name, conf, left, top, right, bottom = line.split()
            det.setdefault(name, []).append((float(conf), file_id, (float(left), float(top), float(right), float(bottom))))
            det_counter[name] = det_counter.get(name, 0) + 1

    ap_dictionary, count_tp = {}, {}
    sum_ap = 0.0
    for class_name in gt_classes:
        dets = det.get(class_name, [])
        order = np.argsort(-np.asarray([d[0] for d in dets], dtype=np.float64), kind='stable') if dets else []
        used = {fid: np.zeros(len(v[0]), bool) for fid, v in gt.items()}
        nd = len(dets)
        tp = np.zeros(nd, np.int64)
        fp = np.zeros(nd, np.int64)
        for k, j in enumerate(order):
            _, file_id, bb = dets[j]
            names, boxes = gt.get(file_id, ([], np.zeros((0, 4))))
            sel = np.nonzero(np.asarray([n == class_name for n in names], bool))[0] if names else np.zeros(0, np.int64)
            ovmax, match = -1.0, -1
            if len(sel):
                g = boxes[sel]
                iw = np.minimum(bb[2], g[:, 2]) - np.maximum(bb[0], g[:, 0]) + 1
                ih = np.minimum(bb[3], g[:, 3]) - np.maximum(bb[1], g[:, 1]) + 1
                ua = (bb[2] - bb[0] + 1) * (bb[3] - bb[1] + 1) + (g[:, 2] - g[:, 0] + 1) * (g[:, 3] - g[:, 1] + 1) - iw * ih
                ov = np.where((iw > 0) & (ih > 0), iw * ih / ua, -np.inf)
                a = int(np.argmax(ov))              # first maximum, as the strict `ov > ovmax` scan keeps
                if ov[a] > ovmax:
                    ovmax, match = float(ov[a]), int(sel[a])
            if ovmax >= min_overlap and not used[file_id][match]:
                tp[k] = 1
                used[file_id][match] = True
            else:
                fp[k] = 1
        count_tp[class_name] = int(tp.sum())
        ctp, cfp = np.cumsum(tp), np.cumsum(fp)
        rec = [float(t) / gt_counter[class_name] for t in ctp]
        prec = [float(t) / (f + t) for t, f in zip(ctp, cfp)]
        ap, _, _ = voc_ap(rec, prec)
        sum_ap += ap
        ap_dictionary[class_name] = ap
        print('{0:.2f}%'.format(ap * 100) + ' = ' + class_name + ' AP ')
    mAP = sum_ap / len(gt_classes)
    text = 'mAP = {0:.2f}%'.format(mAP * 100)
    print(text)
    if output_files_path:
        with open(os.path.join(output_files_path, 'output.txt'), 'w') as f:
            f.write('# AP and precision/recall per class\n')
            f.write('\n# mAP of all classes\n')
            f.write(text + '\n')
    for c in det_counter:
        count_tp.setdefault(c, 0)
    return {'ap': ap_dictionary, 'mAP': mAP, 'gt_counter_per_class': gt_counter, 'counter_images_per_class': img_counter,
            'det_counter_per_class': det_counter, 'count_true_positives': count_tp}
