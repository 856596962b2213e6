import json
import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, 'oracle')):
    if p not in sys.path:
        sys.path.insert(0, p)


def pytest_configure(config):
    config.addinivalue_line('markers', 'gpu: needs a real B200 (run with -m gpu under gpurun)')


@pytest.fixture(scope='session')
def weights():
    """Seeded synthetic darknet weights (oracle object + exact file bytes)."""
    import y4_oracle as O
    W = O.synth_weights(seed=1)
    return W, W.to_darknet_bytes()


def report(name, **kv):
    """Append a JSON line of measured deviations to gpurun_out/test_report.jsonl (read back after a GPU run)."""
    d = os.path.join(ROOT, 'gpurun_out')
    try:
        os.makedirs(d, exist_ok=True)
        with open(os.path.join(d, 'test_report.jsonl'), 'a') as f:
            f.write(json.dumps({'test': name, **kv}, default=float) + '\n')
    except OSError:
        pass
