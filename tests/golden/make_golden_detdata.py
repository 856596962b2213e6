"""Golden vector for get_detection_data (utils.py:56-78) produced by running the reference's own function source
(cut out with ast; utils.py itself does not import here) on seeded model outputs.
    python tests/golden/make_golden_detdata.py"""
import ast
import json
import os
import textwrap

import numpy as np
import pandas as pd

HERE = os.path.dirname(os.path.abspath(__file__))
src = open('/root/reference/utils.py').read()
fn = next(n for n in ast.parse(src).body if isinstance(n, ast.FunctionDef) and n.name == 'get_detection_data')
ns = {'np': np, 'pd': pd}
exec(textwrap.dedent(ast.get_source_segment(src, fn)), ns)
rng = np.random.default_rng(5)
n = 7
boxes = np.zeros((2, 100, 4), np.float32)
xy = np.sort(rng.random((n, 2, 2)).astype(np.float32), axis=1)
boxes[0, :n] = np.concatenate([xy[:, 0], xy[:, 1]], axis=1)
boxes[0, 0] = [0.0, 0.0, 1.0, 1.0]
scores = np.zeros((2, 100), np.float32); scores[0, :n] = np.sort(rng.uniform(0.3, 1, n).astype(np.float32))[::-1]
classes = np.zeros((2, 100), np.float32); classes[0, :n] = rng.integers(0, 5, n)
valid = np.array([n, 0], np.int32)
names = ['person', 'bicycle', 'car', 'motorbike', 'aeroplane']
img = np.zeros((273, 185, 3), np.uint8)
df = ns['get_detection_data'](img, [boxes, scores, classes, valid], names)
json.dump({'boxes': boxes[0, :n].tolist(), 'scores': scores[0, :n].tolist(), 'classes': classes[0, :n].tolist(), 'n': n, 'img_hw': [273, 185],
           'names': names, 'columns': list(df.columns), 'rows': json.loads(df.to_json(orient='values')), 'dtypes': [str(t) for t in df.dtypes]},
          open(os.path.join(HERE, 'detdata_case.json'), 'w'))
print(df)
