"""Per-layer CUDA-event timing of the forward (y4_profile_layers) -> gpurun_out/layers_<tag>.json.
usage: python tools/profile_layers.py [size] [batch] [tag]"""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, 'oracle')]
import numpy as np  # noqa: E402
import y4b200  # noqa: E402
import y4_oracle as O  # noqa: E402

size = int(sys.argv[1]) if len(sys.argv) > 1 else 608
batch = int(sys.argv[2]) if len(sys.argv) > 2 else 32
tag = sys.argv[3] if len(sys.argv) > 3 else 'r01'
prec = {'fp16': y4b200.PREC_FP16, 'fp16x3': y4b200.PREC_FP16X3}[os.environ.get('Y4_PRECISION', 'fp16')]
eng = y4b200.Engine(img_size=size, max_batch=batch, precision=prec)
eng.load_darknet_bytes(O.synth_weights(seed=1).to_darknet_bytes())
eng.synth_fill(0, 0, batch)
for _ in range(3):
    eng.run_forward_resident(batch)
runs = np.stack([eng.profile_layers(batch) for _ in range(5)])
ms = np.median(runs, axis=0)
rows = []
for t, l in zip(ms, eng.steps()):                 # one entry per launch, in schedule order (fused sibling convs: name 'c2+c3')
    if l['kernel_kind'] == 5:
        rows.append({'name': 'spp', 'ms': float(t)})
        continue
    gf = l['flops'] * batch / 1e9
    rows.append({'name': l['out_name'] if l['idx'] >= 110 else f"c{l['idx']}", 'cin': l['cin'], 'cout': l['cout'], 'k': l['ksize'], 's': l['stride'], 'hw': l['out_hw'],
                 'kind': l['kernel_kind'], 'bn': l['tile_n'], 'mode': l['tc_mode'], 'epi': l['tc_epilogue'], 'st': l['tc_stages'], 'grp': l['tc_group'], 'cps': l['tc_ctas_per_sm'], 'nepi': l['tc_epi_warps'], 'bres': l['tc_resident_w'], 'ms': float(t), 'gflop': float(gf), 'tflops': float(gf / t) if t > 0 else 0.0})
tot = float(ms.sum())
out = {'size': size, 'batch': batch, 'total_ms': tot, 'img_per_s': float(batch / tot * 1e3), 'layers': rows}
os.makedirs(os.path.join(ROOT, 'gpurun_out'), exist_ok=True)
with open(os.path.join(ROOT, 'gpurun_out', f'layers_{tag}.json'), 'w') as f:
    json.dump(out, f, indent=0)
print(f'total {tot:.3f} ms / {batch} img  -> {out["img_per_s"]:.0f} img/s (sum of per-layer event times)')
for r in sorted(rows, key=lambda r: -r['ms'])[:25]:
    print(r)
