"""Regenerates tests/golden/*.npz from the oracle (the reference itself cannot run here: TensorFlow is not
installable, see DESIGN.md §2 — these fixtures pin the oracle and the engine against regressions, they are NOT
outputs of the reference).   python tests/golden/make_golden.py"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path[:0] = [ROOT, os.path.join(ROOT, 'oracle')]
import y4_oracle as O  # noqa: E402

# 1. decode + NMS: crafted heads at 96x96, 2 images, 12 classes (small enough to commit)
nc, S = 12, 96
heads = O.synth_heads(seed=42, batch=2, img_size=S, num_classes=nc, n_clusters=12)
m = {}
boxes, scores, classes, valid, idx = O.decode_nms(heads, S, num_classes=nc, margins=m)
assert m['score'] > 1e-5 and m.get('iou', 1) > 1e-4, m
np.savez_compressed(os.path.join(HERE, 'decode_nms_96.npz'), head_s=heads[0].astype(np.float16), head_m=heads[1].astype(np.float16),
                    head_l=heads[2].astype(np.float16), boxes=boxes, scores=scores, classes=classes, valid=valid, idx=idx,
                    num_classes=nc, img_size=S)
# heads are stored as float16 to keep the fixture small: recompute the expected outputs from the ROUNDED heads
heads16 = [h.astype(np.float16).astype(np.float32) for h in heads]
m = {}
boxes, scores, classes, valid, idx = O.decode_nms(heads16, S, num_classes=nc, margins=m)
assert m['score'] > 1e-5 and m.get('iou', 1) > 1e-4, m
np.savez_compressed(os.path.join(HERE, 'decode_nms_96.npz'), head_s=heads16[0].astype(np.float16), head_m=heads16[1].astype(np.float16),
                    head_l=heads16[2].astype(np.float16), boxes=boxes, scores=scores, classes=classes, valid=valid, idx=idx,
                    num_classes=nc, img_size=S)
# 2. forward: checksum-style fixture of the seeded network (weights are 257 MB, so only digests are committed)
W = O.synth_weights(seed=1)
blob = W.to_darknet_bytes()
imgs = O.synth_images(0, 0, 1, 96)
hs = O.forward(imgs, W)
import hashlib
np.savez_compressed(os.path.join(HERE, 'forward_96.npz'), weights_sha256=hashlib.sha256(blob).hexdigest(),
                    head_s=hs[0], head_m=hs[1], head_l=hs[2])
print('golden written', [f for f in os.listdir(HERE) if f.endswith('.npz')])
