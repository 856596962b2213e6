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


def test_export_prediction_files(model, tmp_path):
    """Batched caller: one txt per image, `<class> <score> <x1> <y1> <x2> <y2>` in raw-image pixels (models.py:170-179)."""
    import cv2
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
        out = m.engine.predict_u8([raw])
        lines = (pred_dir / (n.split('.')[0] + '.txt')).read_text().splitlines()
        assert len(lines) == int(out[3][0])
        for i, line in enumerate(lines):
            cls, score, x1, y1, x2, y2 = line.split(' ')
            assert cls == m.class_names[int(out[2][0, i])]
            assert float(score) == pytest.approx(float(out[1][0, i]), rel=1e-6)
            assert float(x1) == pytest.approx(float(out[0][0, i, 0]) * raw.shape[1], rel=1e-5, abs=1e-4)
            assert float(y2) == pytest.approx(float(out[0][0, i, 3]) * raw.shape[0], rel=1e-5, abs=1e-4)
