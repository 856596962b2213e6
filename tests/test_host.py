"""CPU tests: the C-ABI library loads and exports every symbol include/y4.h declares; host-side mirror of the
reference API (config, get_detection_data) behaves like the reference code paths it replaces."""
import os
import re

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_header_symbols_are_exported():
    from y4b200 import binding
    lib = binding.load_library()
    hdr = open(os.path.join(ROOT, 'include', 'y4.h')).read()
    declared = set(re.findall(r'\b(y4_[a-z0-9_]+)\s*\(', hdr))
    assert len(declared) >= 30
    for name in declared:
        assert hasattr(lib, name), name
    assert declared == set(binding.EXPORTS), declared ^ set(binding.EXPORTS)


def test_no_gpu_fails_loudly():
    """No CPU fallback: without a usable sm_100 device y4_create must fail with Y4_ERR_CUDA."""
    import ctypes as C
    import y4b200
    from y4b200 import binding
    lib = binding.load_library()
    cfg = binding.Y4Config()
    assert lib.y4_default_config(C.byref(cfg)) == 0
    assert (cfg.img_size, cfg.num_classes, cfg.max_boxes) == (416, 80, 100)
    assert list(cfg.strides) == [8, 16, 32] and list(cfg.xyscale) == [1.2, 1.1, 1.05]
    assert abs(cfg.iou_threshold - 0.413) < 1e-7 and abs(cfg.score_threshold - 0.3) < 1e-7
    have_gpu = os.path.exists('/dev/nvidia0')
    if not have_gpu:
        with pytest.raises(y4b200.Y4Error) as ei:
            y4b200.Engine()
        assert ei.value.code == -2
    cfg.img_size = 400                                  # not a multiple of 32 (models.py:24)
    h = C.c_void_p()
    assert lib.y4_create(C.byref(h), C.byref(cfg)) == -1
    assert b'multiple' in lib.y4_last_error(None)


def test_config_matches_reference_defaults():
    from y4b200 import yolo_config
    assert yolo_config['img_size'] == (416, 416, 3)
    assert yolo_config['anchors'] == [12, 16, 19, 36, 40, 28, 36, 75, 76, 55, 72, 146, 142, 110, 192, 243, 459, 401]
    assert yolo_config['strides'] == [8, 16, 32] and yolo_config['xyscale'] == [1.2, 1.1, 1.05]
    assert (yolo_config['max_boxes'], yolo_config['iou_threshold'], yolo_config['score_threshold']) == (100, 0.413, 0.3)


def test_get_detection_data_matches_oracle_table():
    """utils.py:56-78: first image only, int64 truncation, DataFrame columns."""
    import y4_oracle as O
    from y4b200.utils import get_detection_data
    boxes = np.zeros((2, 100, 4), np.float32); scores = np.zeros((2, 100), np.float32)
    classes = np.zeros((2, 100), np.float32); valid = np.array([2, 0], np.int32)
    boxes[0, 0] = [0.1017, 0.2049, 0.5551, 0.7999]; boxes[0, 1] = [0.0, 0.3333, 0.9999, 1.0]
    scores[0, :2] = [0.91, 0.42]; classes[0, :2] = [2, 0]
    names = ['person', 'bicycle', 'car']
    img = np.zeros((185, 273, 3), np.uint8)
    df = get_detection_data(img, [boxes, scores, classes, valid], names)
    assert list(df.columns) == ['x1', 'y1', 'x2', 'y2', 'class_name', 'score', 'w', 'h']
    rows = O.detection_table((185, 273), (boxes, scores, classes, valid), names)
    assert len(df) == 2
    for i, r in enumerate(rows):
        assert [int(df.iloc[i][c]) for c in ('x1', 'y1', 'x2', 'y2')] == list(r[:4])
        assert df.iloc[i]['class_name'] == r[4] and int(df.iloc[i]['w']) == r[6] and int(df.iloc[i]['h']) == r[7]
    assert df['x1'].dtype == np.int64


def test_preprocess_contract():
    """models.py:95-98: cv2.resize to (S,S) (no letterbox) then /255 -> float64 in [0,1]."""
    import y4_oracle as O
    img = (np.arange(30 * 50 * 3) % 256).astype(np.uint8).reshape(30, 50, 3)
    out = O.preprocess_img(img, 64)
    assert out.shape == (64, 64, 3) and out.dtype == np.float64 and 0 <= out.min() and out.max() <= 1


def test_eval_map_matches_reference_golden(tmp_path):
    """mAP evaluator vs golden results produced by running the reference's own eval_map / voc_ap source
    (tests/golden/make_golden_map.py) on a seeded ground-truth / prediction folder with ties, duplicates, wrong classes."""
    import json
    import os
    import y4b200
    from y4b200.evaluate import eval_map, voc_ap
    g = json.load(open(os.path.join(os.path.dirname(os.path.abspath(__file__)), 'golden', 'map_case.json')))
    gt_dir, pr_dir, out_dir = tmp_path / 'gt', tmp_path / 'pred', tmp_path / 'out'
    for d in (gt_dir, pr_dir, out_dir):
        d.mkdir()
    for k, v in g['files'].items():
        (gt_dir / (k + '.txt')).write_text('\n'.join(v['gt']) + '\n')
        (pr_dir / (k + '.txt')).write_text('\n'.join(v['pred']) + ('\n' if v['pred'] else ''))
    res = eval_map(str(gt_dir), str(pr_dir), None, str(out_dir))
    assert sorted(res['ap']) == g['classes_sorted']
    for c, want in zip(g['classes_sorted'], g['ap_in_class_order']):
        assert res['ap'][c] == pytest.approx(want, abs=1e-12), c
    assert res['mAP'] == pytest.approx(g['mAP'], abs=1e-12)
    assert (out_dir / 'output.txt').read_text() == g['output_txt']
    for case in g['voc_ap_cases']:
        ap, mrec, mpre = voc_ap(case['rec'], case['prec'])
        assert ap == pytest.approx(case['ap'], abs=1e-15)
        assert mrec == case['mrec'] and mpre == case['mpre']
    assert y4b200.Yolov4.eval_map is not None


def test_get_detection_data_matches_reference_golden():
    """utils.py:56-78 run from the reference's source (tests/golden/make_golden_detdata.py): same columns, dtypes, rows."""
    import json
    import os
    from y4b200.utils import get_detection_data
    g = json.load(open(os.path.join(os.path.dirname(os.path.abspath(__file__)), 'golden', 'detdata_case.json')))
    n = g['n']
    boxes = np.zeros((2, 100, 4), np.float32); boxes[0, :n] = np.array(g['boxes'], np.float32)
    scores = np.zeros((2, 100), np.float32); scores[0, :n] = np.array(g['scores'], np.float32)
    classes = np.zeros((2, 100), np.float32); classes[0, :n] = np.array(g['classes'], np.float32)
    valid = np.array([n, 0], np.int32)
    img = np.zeros(tuple(g['img_hw']) + (3,), np.uint8)
    df = get_detection_data(img, [boxes, scores, classes, valid], g['names'])
    assert list(df.columns) == g['columns']
    assert [str(t) for t in df.dtypes] == g['dtypes']
    rows = json.loads(df.to_json(orient='values'))
    assert rows == g['rows']


def test_bench_flow_with_stub_engine():
    """bench.py's `ours` arm end to end against a stub engine: exactly ONE line on stdout (native libraries such as NCCL print
    banners to fd 1; bench.py must keep them off it), valid JSON with every key the contract names."""
    import json
    import subprocess
    import sys
    here = os.path.dirname(os.path.abspath(__file__))
    r = subprocess.run([sys.executable, os.path.join(here, 'bench_stub_run.py'), '--steps', '2', '--warmup', '3', '--size', '64', '--batch', '2',
                        '--no-cpu-baseline'], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [l for l in r.stdout.splitlines() if l.strip()]
    assert len(lines) == 1, r.stdout
    d = json.loads(lines[0])
    for k in ('metric', 'value', 'unit', 'n_gpus', 'steps', 'warmup', 'ms_per_step', 'higher_is_better', 'scaling', 'vs_baseline', 'dtype', 'data',
              'config', 'gpu_launches', 'clocks', 'e2e', 'roofline'):
        assert k in d, k
    assert d['steps'] == 2 and d['warmup'] == 3 and d['n_gpus'] == 1 and d['gpu_launches'] == 2 * 116
    for k in ('bound', 'achieved', 'peak', 'unit', 'frac', 'traffic', 'kernel', 'frac_of_burst_peak', 'frac_of_sustained_peak'):
        assert k in d['roofline'], k
    for k in ('bound', 'achieved', 'peak', 'unit', 'frac', 'us_per_img'):
        assert k in d['decode_nms'], k
    assert set(d['parity_modes']) >= {'fp16x3', 'fp32'}
    for k in ('value', 'unit', 'h2d_bytes_per_step', 'd2h_bytes_per_step'):
        assert k in d['e2e'], k
    assert d['e2e']['h2d_bytes_per_step'] == 2 * 64 * 64 * 3


class _FacadeStubEngine:
    """Stand-in for binding.Engine: lets the host-side façade logic (batching, formats, flips) run without a GPU."""
    max_batch = 2

    def __init__(self, **kw):
        self.calls = []
        self.queue = []

    def load_darknet(self, path):
        self.calls.append(('load', path))

    def _out(self, n):
        boxes = np.zeros((n, 100, 4), np.float32); scores = np.zeros((n, 100), np.float32)
        classes = np.zeros((n, 100), np.float32); valid = np.full((n,), 2, np.int32)
        boxes[:, 0] = [0.1, 0.2, 0.5, 0.75]; boxes[:, 1] = [0.0, 0.0, 1.0, 1.0]
        scores[:, 0] = 0.9; scores[:, 1] = 0.4; classes[:, 1] = 1
        return [boxes, scores, classes, valid]

    def predict_u8(self, raws, reverse_channels=False, with_indices=False):
        self.calls.append(('predict_u8', len(raws), reverse_channels))
        return self._out(len(raws))

    def submit_u8(self, raws, reverse_channels=False):
        assert len(self.queue) < 2, 'more than two batches in flight'
        self.calls.append(('submit_u8', len(raws), reverse_channels))
        self.queue.append(len(raws))

    def collect(self, with_indices=False):
        return self._out(self.queue.pop(0))


def test_facade_host_logic_with_stub_engine(tmp_path, monkeypatch):
    """predict_img on a raw uint8 image goes through the GPU-preprocess entry point and yields the reference's DataFrame;
    export_prediction keeps at most two batches in flight, never flips channels (models.py:153) and writes
    `<class> <score> <x1> <y1> <x2> <y2>` in raw-image pixels (models.py:170-179)."""
    cv2 = pytest.importorskip('cv2')
    import y4b200
    from y4b200 import models
    monkeypatch.setattr(models, 'Engine', _FacadeStubEngine)
    names = tmp_path / 'names.txt'
    names.write_text('person\ncar\n')
    wfile = tmp_path / 'w.weights'
    wfile.write_bytes(b'')
    m = y4b200.Yolov4(weight_path=str(wfile), class_name_path=str(names))
    assert ('load', str(wfile)) in m.engine.calls
    raw = np.zeros((200, 400, 3), np.uint8)
    df = m.predict_img(raw, plot_img=False)
    assert ('predict_u8', 1, False) in m.engine.calls
    assert list(df['class_name']) == ['person', 'car'] and list(df['x2']) == [200, 400] and list(df['h']) == [int(0.75 * 200) - int(0.2 * 200), 200]
    img_dir, pred_dir = tmp_path / 'imgs', tmp_path / 'pred'
    img_dir.mkdir(); pred_dir.mkdir()
    for i in range(5):
        cv2.imwrite(str(img_dir / f'a{i}.png'), np.zeros((100 + i, 50, 3), np.uint8))
    ann = tmp_path / 'ann.txt'
    ann.write_text(''.join(f'a{i}.png 1,2,3,4,0\n' for i in range(5)))
    m.engine.calls.clear()
    m.export_prediction(str(ann), str(pred_dir), str(img_dir), bs=2)
    assert [c for c in m.engine.calls if c[0] == 'submit_u8'] == [('submit_u8', 2, False), ('submit_u8', 2, False), ('submit_u8', 1, False)]
    lines = (pred_dir / 'a3.txt').read_text().splitlines()
    assert len(lines) == 2
    cls, score, x1, y1, x2, y2 = lines[0].split(' ')
    assert cls == 'person' and float(score) == pytest.approx(0.9) and float(x2) == pytest.approx(0.5 * 50) and float(y2) == pytest.approx(0.75 * 103)


def test_library_sass_is_tcgen05_tma_and_packed_fp32():
    """The shipped liby4.so is sm_100a code whose conv kernels issue tcgen05.mma (UTCHMMA, incl. the cta_group::2 form), read
    TMEM (LDTM), move tiles by TMA (UTMALDG / UTMASTG) and run the mish epilogue on packed fp32 (FFMA2); no legacy mma.sync
    (HMMA) anywhere.  Static check with cuobjdump (skipped where the CUDA toolkit is absent)."""
    import shutil
    import subprocess
    cuobjdump = shutil.which('cuobjdump') or '/usr/local/cuda/bin/cuobjdump'
    if not os.path.exists(cuobjdump):
        pytest.skip('cuobjdump not available')
    import y4b200
    sass = subprocess.run([cuobjdump, '-sass', y4b200.lib_path()], capture_output=True, text=True, check=True).stdout
    assert 'sm_100a' in sass or 'SM100a' in sass or 'sm_100' in sass
    count = {k: len(re.findall(r'\b' + k + r'\b', sass)) for k in ('UTCHMMA', 'LDTM', 'UTMALDG', 'UTMASTG', 'FFMA2', 'HMMA')}
    count['UTCHMMA.2CTA'] = sass.count('UTCHMMA.2CTA')
    assert count['UTCHMMA'] > 500 and count['UTCHMMA.2CTA'] > 100 and count['LDTM'] > 20, count
    assert count['UTMALDG'] > 100 and count['UTMASTG'] > 10 and count['FFMA2'] > 1000, count
    assert count['HMMA'] == 0, count
