"""BASELINE config 5: input-size sweep 320/416/512/608/768 at batch 16 on one B200, images/s vs the conv-FLOP roofline.
usage: python tools/bench_sweep.py [batch] -> JSON lines (also gpurun_out/sweep.jsonl)"""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, 'oracle')]
import numpy as np  # noqa: E402
import netspec  # noqa: E402
import y4b200  # noqa: E402
import y4_oracle as O  # noqa: E402

batch = int(sys.argv[1]) if len(sys.argv) > 1 else 16
peak = 1440.6
try:
    peak = json.load(open(os.path.join(ROOT, 'MEASURED_PEAKS.json')))['bf16_tflops_sustained']
except Exception:
    pass
blob = O.synth_weights(seed=1).to_darknet_bytes()
os.makedirs(os.path.join(ROOT, 'gpurun_out'), exist_ok=True)
out = open(os.path.join(ROOT, 'gpurun_out', 'sweep.jsonl'), 'w')
for size in (320, 416, 512, 608, 768):
    eng = y4b200.Engine(img_size=size, max_batch=batch, precision=y4b200.PREC_FP16)
    eng.load_darknet_bytes(blob)
    eng.synth_fill(0, 0, batch)
    for _ in range(3):
        eng.run_resident(batch)
    eng.sync()
    steps = 20
    eng.timer_begin()
    for _ in range(steps):
        eng.run_resident(batch)
    ms = eng.timer_end() / steps
    gf = netspec.conv_gflop(size)
    line = {'workload': f'configs[4]: size sweep, batch {batch}', 'size': size, 'images_per_s': batch / ms * 1e3, 'ms_per_step': ms,
            'conv_gflop_per_image': gf, 'achieved_tflops': gf * batch / ms, 'peak_tflops_measured_sustained': peak,
            'frac_of_conv_flop_roofline': gf * batch / ms / peak}
    print(json.dumps(line), flush=True)
    out.write(json.dumps(line) + '\n')
    eng.close()
out.close()
