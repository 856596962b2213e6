"""Runs bench.main() against a stub engine (no GPU): exercises the whole control flow of the `ours` arm -- argument handling,
stdout protection, clock sampler, roofline grouping, e2e loops -- so that a typo cannot reach the round-end GPU run untested.
Invoked by tests/test_host.py::test_bench_flow_with_stub_engine as a subprocess; prints exactly what bench.py would."""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, 'oracle')]
import y4b200  # noqa: E402
from y4b200 import binding  # noqa: E402
import y4_oracle as O  # noqa: E402


class StubEngine:
    max_boxes, num_boxes, num_classes = 100, 22743, 80

    def __init__(self, img_size=416, max_batch=1, precision=0, device=0, **kw):
        self.S, self.B, self.n = img_size, max_batch, 0
        self._t = 0.0

    def load_darknet_bytes(self, b): pass
    def comm_unique_id(self): return np.zeros(128, np.uint8)
    def comm_init(self, *a): pass
    def sync(self): pass
    def synth_fill(self, *a): pass
    def run_resident(self, b): self.n += 116; time.sleep(0.002)
    def run_forward_resident(self, b): self.n += 111; time.sleep(0.002)
    def run_decode_nms_resident(self, b): self.n += 4
    def allgather_results(self, b, fetch=True): pass
    def launch_count(self): return self.n
    def timer_begin(self): self._t = time.perf_counter()
    def timer_end(self): return 1e3 * (time.perf_counter() - self._t)

    def layers(self):
        out = []
        for i in range(110):
            kind = 4 if i == 0 else (2 if i in (1, 8) else 1)
            out.append({'idx': i, 'kernel_kind': kind, 'tile_n': 256 if i % 3 else 64, 'tc_epi_warps': 44 if i % 5 == 0 else (8 if i % 2 else 4),
                        'tc_mode': 4 if i % 4 == 0 and i else 1, 'tc_bk': 64, 'flops': 10 ** 9})
        return out

    def steps(self):
        out = self.layers()
        out.insert(75, {'idx': -1, 'kernel_kind': 5, 'tile_n': 0, 'tc_epi_warps': 0, 'tc_mode': 0, 'tc_bk': 0, 'flops': 0})
        return out

    def profile_layers(self, b): return np.full(111, 0.05, np.float32)

    def _out(self, b):
        return [np.zeros((b, 100, 4), np.float32), np.zeros((b, 100), np.float32), np.zeros((b, 100), np.float32), np.full((b,), 7, np.int32)]

    def predict(self, imgs, with_indices=False): return self._out(len(imgs))
    def submit(self, imgs): self._b = len(imgs)
    def submit_u8(self, imgs, reverse_channels=False): self._b = len(imgs)
    def collect(self, with_indices=False): return self._out(self._b)
    def close(self): pass


class StubWeights:
    def to_darknet_bytes(self): return b''


y4b200.Engine = StubEngine
binding.pinned_array = lambda shape, dtype=np.float32: np.zeros(shape, dtype)
O.synth_weights = lambda seed=1: StubWeights()
O.synth_images = lambda seed, first, batch, size: np.zeros((batch, size, size, 3), np.float32)
sys.argv = ['bench.py'] + sys.argv[1:]
import bench  # noqa: E402
bench.main()
