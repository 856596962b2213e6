"""class Yolov4 over the engine: the reference's public calls (models.py:109-179) end to end on raw uint8 images --
GPU preprocess, forward, decode, NMS, DataFrame / text-file result formats -- against the oracle."""
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


@pytest.fixture(scope='module')
def model(weights, tmp_path_factory):
    import y4b200
    W, blob = weights
    d = tmp_path_factory.mktemp('y4')
    wpath = str(d / 'synthetic.weights')
    with open(wpath, 'wb') as f:
        f.write(blob)
    names = str(d / 'names.txt')
    with open(names, 'w') as f:
        f.write('\n'.join(f'class{i}' for i in range(80)) + '\n')
    cfg = dict(y4b200.yolo_config)
    cfg.update(img_size=(160, 160, 3), precision='fp32', max_batch=4)
    m = y4b200.Yolov4(weight_path=wpath, class_name_path=names, config=cfg)
    return m, W, d


def _raw(seed, h, w):
    import y4_oracle as O
    # smooth synthetic picture (the network's detections on white noise sit too close to the thresholds)
    base = O.synth_images(seed, 0, 1, 32)[0]
    import cv2
    return np.ascontiguousarray((cv2.resize(base, (w, h)) * 255).astype(np.uint8))


def test_predict_img_dataframe_matches_oracle(model):
    import y4_oracle as O
    m, W, _ = model
    raw = _raw(11, 120, 200)
    df = m.predict_img(raw, plot_img=False)
    pre = O.preprocess_img(raw, 160)[None].astype(np.float32)
    ref = O.predict(pre, W)
    rows = O.detection_table(raw.shape[:2], ref, m.class_names)
    assert list(df.columns) == ['x1', 'y1', 'x2', 'y2', 'class_name', 'score', 'w', 'h']
    assert len(df) == len(rows)
    for i, r in enumerate(rows):
        got = df.iloc[i]
        assert got['class_name'] == r[4]
        assert abs(float(got['score']) - r[5]) <= 2e-4
        for k, name in enumerate(('x1', 'y1', 'x2', 'y2')):
            assert abs(int(got[name]) - r[k]) <= 1, (i, name)         # int64 truncation of coords that may differ by 1e-4


def _street_png(d):
    """The reference's own test image (img/street.jpeg as cv2.imread decodes it, stored in the golden fixture because
    /root/reference does not exist on the GPU box), written losslessly so that cv2.imread(path) returns the same pixels."""
    import cv2
    g = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), 'golden', 'preprocess_street.npz'))
    path = str(d / 'street.png')
    cv2.imwrite(path, g['raw'])
    return path, g['raw']


def _assert_table(df, rows):
    assert list(df.columns) == ['x1', 'y1', 'x2', 'y2', 'class_name', 'score', 'w', 'h']
    assert len(df) == len(rows)
    for i, r in enumerate(rows):
        got = df.iloc[i]
        assert got['class_name'] == r[4], i
        assert abs(float(got['score']) - r[5]) <= 2e-4, i
        for k, name in enumerate(('x1', 'y1', 'x2', 'y2')):
            assert abs(int(got[name]) - r[k]) <= 1, (i, name)         # int64 truncation of coords that may differ by 1e-4


def test_predict_path_flips_bgr_to_rgb(model, tmp_path):
    """Yolov4.predict(path) (models.py:125-127): cv2.imread -> [:, :, ::-1] -> predict_img."""
    import y4_oracle as O
    m, W, _ = model
    path, bgr = _street_png(tmp_path)
    df = m.predict(path, plot_img=False)
    rgb = np.ascontiguousarray(bgr[:, :, ::-1])
    ref = O.predict(O.preprocess_img(rgb, 160)[None].astype(np.float32), W)
    _assert_table(df, O.detection_table(bgr.shape[:2], ref, m.class_names))


def test_predict_raw_and_predict_nonms(model, tmp_path):
    """predict_raw (models.py:509-514): the three raw heads, NO channel flip; predict_nonms (models.py:516-529): decode +
    NMS with caller thresholds (defaults 0.413 / 0.1) on those heads."""
    import y4_oracle as O
    m, W, _ = model
    path, bgr = _street_png(tmp_path)
    pre = O.preprocess_img(bgr, 160)[None].astype(np.float32)          # BGR kept, as the reference does
    want = O.forward(pre, W)
    heads = m.predict_raw(path)
    assert [h.shape for h in heads] == [(1, 20, 20, 255), (1, 10, 10, 255), (1, 5, 5, 255)]
    for a, b in zip(heads, want):
        assert float(np.abs(a - b).max() / np.abs(b).max()) < 1e-4
    for kw in (dict(), dict(iou_threshold=0.5, score_threshold=0.25)):
        df = m.predict_nonms(path, **kw)
        ref = O.decode_nms(want, 160, iou_threshold=kw.get('iou_threshold', 0.413), score_threshold=kw.get('score_threshold', 0.1))
        _assert_table(df, O.detection_table(bgr.shape[:2], ref, m.class_names))


def test_precision_names(weights, tmp_path):
    """config['precision'] selects every engine mode, including the tensor-core parity mode 'fp16x3'."""
    import y4b200
    W, blob = weights
    wpath, names = str(tmp_path / 'w.weights'), str(tmp_path / 'names.txt')
    open(wpath, 'wb').write(blob)
    open(names, 'w').write('\n'.join(f'class{i}' for i in range(80)) + '\n')
    for name, code in (('fp16x3', y4b200.PREC_FP16X3), ('fp16', y4b200.PREC_FP16)):
        cfg = dict(y4b200.yolo_config)
        cfg.update(img_size=(64, 64, 3), precision=name, max_batch=1)
        m = y4b200.Yolov4(weight_path=wpath, class_name_path=names, config=cfg)
        assert m.engine.cfg.precision == code
        m.engine.close()
    cfg.update(precision='bf16')
    with pytest.raises(AssertionError):
        y4b200.Yolov4(weight_path=wpath, class_name_path=names, config=cfg)


def test_export_prediction_files(model, tmp_path):
    """Batched caller: one txt per image, `<class> <score> <x1> <y1> <x2> <y2>` in raw-image pixels (models.py:141-179),
    against the ORACLE's predict on the oracle-preprocessed (BGR, unflipped) images."""
    import cv2
    import y4_oracle as O
    m, W, d = model
    img_dir, pred_dir = tmp_path / 'imgs', tmp_path / 'pred'
    img_dir.mkdir(); pred_dir.mkdir()
    names = []
    for i, (h, w) in enumerate([(120, 200), (90, 90), (200, 140), (64, 300), (150, 151)]):
        cv2.imwrite(str(img_dir / f'im{i}.png'), _raw(20 + i, h, w))
        names.append(f'im{i}.png')
    ann = tmp_path / 'ann.txt'
    ann.write_text('\n'.join(f'{n} 1,2,3,4,0' for n in names) + '\n')
    m.export_prediction(str(ann), str(pred_dir), str(img_dir), bs=2)
    for n in names:
        raw = cv2.imread(str(img_dir / n))
        ref = O.predict(O.preprocess_img(raw, 160)[None].astype(np.float32), W)
        lines = (pred_dir / (n.split('.')[0] + '.txt')).read_text().splitlines()
        assert len(lines) == int(ref[3][0])
        for i, line in enumerate(lines):
            cls, score, x1, y1, x2, y2 = line.split(' ')
            assert cls == m.class_names[int(ref[2][0, i])]
            assert abs(float(score) - float(ref[1][0, i])) <= 2e-4
            for v, k, dim in ((x1, 0, 1), (y1, 1, 0), (x2, 2, 1), (y2, 3, 0)):
                assert abs(float(v) - float(ref[0][0, i, k]) * raw.shape[dim]) <= 2e-4 * raw.shape[dim] + 1e-3


def test_export_gt_format(model, tmp_path):
    """export_gt (models.py:129-139): host-only; the ground-truth side of the mAP workflow."""
    m, _, _ = model
    ann = tmp_path / 'ann.txt'
    ann.write_text('a/b/img1.jpg 10,20,30,40,2 1,2,3,4,0\nimg2.jpg 5,6,7,8,1\n')
    gt = tmp_path / 'gt'
    gt.mkdir()
    m.export_gt(str(ann), str(gt))
    assert (gt / 'img1.txt').read_text() == 'class2 10.0 20.0 30.0 40.0\nclass0 1.0 2.0 3.0 4.0\n'
    assert (gt / 'img2.txt').read_text() == 'class1 5.0 6.0 7.0 8.0\n'
