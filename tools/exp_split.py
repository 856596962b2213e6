"""Experiment: accuracy of Y4_PREC_FP16X3 as a function of the tensor-core accumulation chunk (Y4_SPLIT_CHUNK k-blocks per
TMEM partial; 1 = every 64-deep k-block drained and summed round-to-nearest, 1000 = whole K inside the tensor core) and its
throughput.  Usage (GPU box): python tools/exp_split.py [chunks ...]  ->  gpurun_out/exp_split.jsonl"""
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, 'oracle'), os.path.join(ROOT, 'tests')):
    sys.path.insert(0, p)
import y4b200  # noqa: E402
import y4_oracle as O  # noqa: E402


def rel(a, b):
    return float(np.abs(a - b).max() / (np.abs(b).max() + 1e-12))


def out(**kv):
    os.makedirs(os.path.join(ROOT, 'gpurun_out'), exist_ok=True)
    with open(os.path.join(ROOT, 'gpurun_out', 'exp_split.jsonl'), 'a') as f:
        f.write(json.dumps(kv, default=float) + '\n')
    print(json.dumps(kv, default=float)[:400], flush=True)


def main():
    # args: chunk[:comp] ...   (comp = Y4_SPLIT_COMP multiplier of the modelled truncation bias, conv_tc.cuh)
    specs = [a.split(':') for a in sys.argv[1:]] or [['1000'], ['9'], ['2'], ['1']]
    chunks = [(int(sp[0]), sp[1] if len(sp) > 1 else None) for sp in specs]
    W = O.synth_weights(seed=1)
    blob = W.to_darknet_bytes()
    size, batch = 160, 2
    imgs = O.synth_images(0, 0, batch, size)
    k32, k64 = {}, {}
    heads32 = O.forward(imgs, W, keep=k32)
    O.forward(imgs, W, np.float64, keep=k64)
    from test_gpu_forward import _stable_case
    cases = {s: _stable_case(W, s) for s in (256, 416)}
    for ch, comp in chunks:
        os.environ['Y4_SPLIT_CHUNK'] = str(ch)
        os.environ.pop('Y4_SPLIT_COMP', None)
        if comp is not None:
            os.environ['Y4_SPLIT_COMP'] = comp
        ch = f'{ch}:{comp}'
        eng = y4b200.Engine(img_size=size, max_batch=batch, precision=y4b200.PREC_FP16X3)
        eng.load_darknet_bytes(blob)
        got_heads = eng.forward_heads(imgs)
        rows = {}
        for name, ref in k64.items():
            try:
                got = eng.get_tensor(name, batch).reshape(ref.shape)
            except y4b200.Y4Error:
                continue
            rows[name] = (rel(got, ref), rel(k32[name], ref))
        pick = ['c0', 'c1', 'c2', 'r1', 'c7', 'c8', 'r3', 'c16', 'c17', 'r11', 'c37', 'c58', 'c77', 'c91', 'c99', 'c107', 'c108', 'c109']
        out(chunk=ch, what='layerwise_160', rows={n: rows[n] for n in pick if n in rows},
            heads_vs_fp32=[rel(a, b) for a, b in zip(got_heads, heads32)])
        eng.close()
        for s, (im, ref, noise, first) in cases.items():
            eng = y4b200.Engine(img_size=s, max_batch=1, precision=y4b200.PREC_FP16X3)
            eng.load_darknet_bytes(blob)
            got = eng.predict(im, with_indices=True)
            out(chunk=ch, what=f'predict_{s}', noise=noise, valid=int(ref[3][0]), got_valid=int(got[3][0]),
                idx_equal=bool(np.array_equal(got[4], ref[4])), cls_equal=bool(np.array_equal(got[2], ref[2])),
                box_err=float(np.abs(got[0] - ref[0]).max()), score_err=float(np.abs(got[1] - ref[1]).max()))
            eng.close()
    for ch, comp in ([chunks[-1], chunks[0]] if os.environ.get('EXP_TP') else []):
        os.environ['Y4_SPLIT_CHUNK'] = str(ch)
        S, B = 608, 32
        t0 = time.time()
        eng = y4b200.Engine(img_size=S, max_batch=B, precision=y4b200.PREC_FP16X3)
        eng.load_darknet_bytes(blob)
        t_create = time.time() - t0
        eng.synth_fill(0, 0, B)
        for _ in range(3):
            eng.run_resident(B)
        eng.sync()
        eng.timer_begin()
        for _ in range(5):
            eng.run_resident(B)
        ms = eng.timer_end() / 5
        prof = eng.profile_layers(B) if hasattr(eng, 'profile_layers') else None
        out(chunk=ch, what='throughput_608_b32', ms_per_step=ms, img_s=B / ms * 1e3, create_s=t_create,
            layer_ms=[round(float(x), 4) for x in prof] if prof is not None else None)
        eng.close()


if __name__ == '__main__':
    main()
