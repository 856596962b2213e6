"""Golden vectors for eval_map / voc_ap produced by RUNNING THE REFERENCE'S OWN CODE: the two functions are cut out of
/root/reference/{models,utils}.py with `ast` (the modules themselves do not import here: tensorflow / matplotlib are
absent) and executed with inert stand-ins for `plt` and `draw_plot_func`, on a seeded synthetic ground-truth / prediction
folder.  The dataset text and the reference's results are committed as tests/golden/map_case.json.
    python tests/golden/make_golden_map.py        (needs /root/reference; run in the build container)"""
import ast
import json
import os
import tempfile
import types
from glob import glob

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))


def cut(path, name, cls=None):
    src = open(path).read()
    tree = ast.parse(src)
    body = tree.body
    if cls:
        body = next(n for n in tree.body if isinstance(n, ast.ClassDef) and n.name == cls).body
    fn = next(n for n in body if isinstance(n, ast.FunctionDef) and n.name == name)
    import textwrap
    return textwrap.dedent(ast.get_source_segment(src, fn))


class _Inert:
    def __getattr__(self, k):
        return _Inert()

    def __call__(self, *a, **k):
        return _Inert()


ns = {'os': os, 'json': json, 'glob': glob, 'np': np, 'plt': _Inert(), 'draw_plot_func': lambda *a, **k: None}
exec(cut('/root/reference/utils.py', 'voc_ap'), ns)
exec(cut('/root/reference/utils.py', 'read_txt_to_list'), ns)
exec(cut('/root/reference/models.py', 'eval_map', cls='Yolov4'), ns)

rng = np.random.default_rng(2024)
classes = ['person', 'car', 'dog', 'bicycle', 'traffic_light']
files = {}
for i in range(12):
    n = int(rng.integers(1, 7))
    gts, prs = [], []
    for _ in range(n):
        c = classes[int(rng.integers(0, len(classes) - (1 if i < 3 else 0)))]
        x1, y1 = rng.integers(0, 300, 2); w, h = rng.integers(20, 200, 2)
        gts.append(f'{c} {x1} {y1} {x1 + w} {y1 + h}')
        r = rng.random()
        if r < 0.75:                                   # a detection near the object (sometimes the wrong class, sometimes twice)
            jit = rng.normal(0, 0.12, 4) * [w, h, w, h]
            pc = c if rng.random() < 0.85 else classes[int(rng.integers(0, len(classes)))]
            conf = float(np.round(rng.uniform(0.3, 0.99), 2))      # two decimals: ties happen
            prs.append(f'{pc} {conf} {x1 + jit[0]} {y1 + jit[1]} {x1 + w + jit[2]} {y1 + h + jit[3]}')
            if rng.random() < 0.2:
                prs.append(f'{pc} {float(np.round(rng.uniform(0.3, 0.99), 2))} {x1 + 2} {y1 - 1} {x1 + w} {y1 + h + 3}')
    for _ in range(int(rng.integers(0, 3))):           # spurious detections
        x1, y1 = rng.integers(0, 300, 2); w, h = rng.integers(20, 200, 2)
        prs.append(f'{classes[int(rng.integers(0, len(classes)))]} {float(np.round(rng.uniform(0.3, 0.9), 2))} {x1} {y1} {x1 + w} {y1 + h}')
    files[f'img{i:03d}'] = (gts, prs)

with tempfile.TemporaryDirectory() as d:
    gt_dir, pr_dir, tmp_dir, out_dir = (os.path.join(d, x) for x in ('gt', 'pred', 'tmp', 'out'))
    for x in (gt_dir, pr_dir, tmp_dir, out_dir):
        os.makedirs(x)
    for k, (g, p) in files.items():
        open(os.path.join(gt_dir, k + '.txt'), 'w').write('\n'.join(g) + '\n')
        open(os.path.join(pr_dir, k + '.txt'), 'w').write('\n'.join(p) + ('\n' if p else ''))
    captured = {}
    real_voc = ns['voc_ap']

    def spy(rec, prec):
        out = real_voc(rec, prec)
        captured.setdefault('aps', []).append(out[0])
        return out
    ns['voc_ap'] = spy
    ns['eval_map'](types.SimpleNamespace(), gt_dir, pr_dir, tmp_dir, out_dir)
    output_txt = open(os.path.join(out_dir, 'output.txt')).read()
# voc_ap on its own
cases = []
for _ in range(6):
    n = int(rng.integers(1, 30))
    tp = rng.integers(0, 2, n); fp = 1 - tp
    ctp, cfp = np.cumsum(tp), np.cumsum(fp)
    rec = [float(t) / (tp.sum() + 3) for t in ctp]
    prec = [float(t) / (f + t) for t, f in zip(ctp, cfp)]
    ap, mrec, mpre = real_voc(rec[:], prec[:])
    cases.append({'rec': rec, 'prec': prec, 'ap': ap, 'mrec': mrec, 'mpre': mpre})
gt_classes = sorted({l.split()[0] for g, _ in files.values() for l in g})
json.dump({'files': {k: {'gt': g, 'pred': p} for k, (g, p) in files.items()}, 'classes_sorted': gt_classes,
           'ap_in_class_order': captured['aps'], 'mAP': float(np.sum(captured['aps']) / len(gt_classes)) if False else sum(captured['aps']) / len(gt_classes),
           'output_txt': output_txt, 'voc_ap_cases': cases},
          open(os.path.join(HERE, 'map_case.json'), 'w'))
print('classes', gt_classes, 'aps', captured['aps'], '\n' + output_txt)
