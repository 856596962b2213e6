"""Y4_PREC_FP16X3 — the tensor-core PARITY mode: activations and weights as fp16 hi+lo pairs (products exact to 2^-22),
three tcgen05 MMAs per k-step, and CHUNKED accumulation: the tensor core's fp32 accumulate rounds toward zero (measured
on B200: ~0.4 ulp of the running sum lost per MMA, always in the same direction: 1.7e-3 of max at conv 108 when the
whole K is accumulated in TMEM), so the tensor core only accumulates one 64-deep k-block (cross terms first, then four
hi*hi MMAs) into a fresh TMEM partial and the epilogue warps sum the partials round-to-nearest in registers.
Measured (tools/exp_split.py, profiles/r02_split_chunk.md): conv 108 within 1.0e-4 of the float64 evaluation where the
fp32 oracle itself is at 5.9e-5 and the CUDA-core fp32 kernel at 7.0e-5; heads within 9e-5 of the fp32 oracle.
The tests hold this mode to the same bars as Y4_PREC_FP32 (tests/test_gpu_forward.py)."""
import numpy as np
import pytest

from conftest import report
from test_gpu_forward import STABLE_START, _rel, _stable_case

pytestmark = pytest.mark.gpu


def test_split_layer_by_layer(weights):
    """Every materialised tensor: as close to the float64 evaluation as the fp32 oracle is (x4 + 2e-6, the bar of the fp32
    CUDA-core mode); heads within 1.5e-4 of the fp32 oracle."""
    import y4b200
    import y4_oracle as O
    W, blob = weights
    size, batch = 160, 2
    imgs = O.synth_images(0, 0, batch, size)
    k32, k64 = {}, {}
    heads = O.forward(imgs, W, keep=k32)
    O.forward(imgs, W, np.float64, keep=k64)
    eng = y4b200.Engine(img_size=size, max_batch=batch, precision=y4b200.PREC_FP16X3)
    eng.load_darknet_bytes(blob)
    got_heads = eng.forward_heads(imgs)
    kinds = [l['kernel_kind'] for l in eng.layers()]
    assert sum(k in (1, 2) for k in kinds) >= 100, kinds          # it IS the tcgen05 path
    rows = []
    for name, ref in k64.items():
        try:
            got = eng.get_tensor(name, batch).reshape(ref.shape)
        except y4b200.Y4Error:
            continue
        rows.append((name, _rel(got, ref), _rel(k32[name], ref)))
    worst = sorted(rows, key=lambda r: -r[1])[:6]
    early = [(n, a, b) for n, a, b in rows if n in ('c0', 'c1', 'c2', 'r1', 'c7', 'c8', 'r3', 'c16', 'c17', 'r11', 'c37', 'c58', 'c77')]
    report('split_layerwise', n=len(rows), worst=worst, early=early, heads_vs_fp32=[_rel(a, b) for a, b in zip(got_heads, heads)])
    assert len(rows) >= 90
    for name, e_eng, e_ora in rows:
        assert e_eng <= 4 * e_ora + 2e-6, (name, e_eng, e_ora)
    for a, b in zip(got_heads, heads):
        assert _rel(a, b) < 1.5e-4
    eng.close()


@pytest.mark.parametrize('size', [256, 416, 608])
def test_split_predict_end_to_end_on_roundoff_stable_image(weights, size):
    """The north-star gate on the tensor-core path: bit-exact indices / classes / valid, boxes and scores within
    max(1e-4, 3 x the oracle's own round-off), on a round-off-stable image (test_gpu_forward._stable_case)."""
    import y4b200
    W, blob = weights
    imgs, ref, noise, first = _stable_case(W, size, start=STABLE_START[size])
    tol = max(1e-4, 3 * noise)
    eng = y4b200.Engine(img_size=size, max_batch=1, precision=y4b200.PREC_FP16X3)
    eng.load_darknet_bytes(blob)
    kinds = [l['kernel_kind'] for l in eng.layers()]
    assert sum(k in (1, 2) for k in kinds) >= 100, kinds
    got = eng.predict(imgs, with_indices=True)
    report(f'split_predict_{size}', first=first, oracle_noise=noise, tol=tol, valid=ref[3].tolist(), got_valid=got[3].tolist(),
           box_err=float(np.abs(got[0] - ref[0]).max()), score_err=float(np.abs(got[1] - ref[1]).max()),
           idx_equal=bool(np.array_equal(got[4], ref[4])))
    assert np.array_equal(got[3], ref[3])
    assert np.array_equal(got[4], ref[4])
    assert np.array_equal(got[2], ref[2])
    assert np.abs(got[0] - ref[0]).max() <= tol
    assert np.abs(got[1] - ref[1]).max() <= tol
    eng.close()
