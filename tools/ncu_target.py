"""Target for ncu --profile-from-start off: build + autotune + warm up outside the profiled range, then one full step
(110 convs + SPP + decode + NMS) inside cudaProfilerStart/Stop.   usage: ncu ... python tools/ncu_target.py [size] [batch]"""
import ctypes
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, 'oracle')]
import y4b200  # noqa: E402
import y4_oracle as O  # noqa: E402

size = int(sys.argv[1]) if len(sys.argv) > 1 else 608
batch = int(sys.argv[2]) if len(sys.argv) > 2 else 32
os.environ['Y4_GRAPH'] = '0'
rt = ctypes.CDLL('libcudart.so.12')
eng = y4b200.Engine(img_size=size, max_batch=batch, precision=y4b200.PREC_FP16)
eng.load_darknet_bytes(O.synth_weights(seed=1).to_darknet_bytes())
eng.synth_fill(0, 0, batch)
for _ in range(3):
    eng.run_resident(batch)
eng.sync()
rt.cudaProfilerStart()
eng.run_resident(batch)
eng.sync()
rt.cudaProfilerStop()
print('profiled one step; launches so far', eng.launch_count())
