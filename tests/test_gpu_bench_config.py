"""Parity ON THE BENCHMARKED CONFIGURATION (BASELINE.json configs[1]; SURVEY §8(d) cfg 2): batch 32 of 608x608 synthetic
images through the engine, >= 2 of the 32 images compared end to end with the oracle -- in the tensor-core parity mode
(Y4_PREC_FP16X3), in the CUDA-core fp32 mode, and (loosely: it is fp16) in the mode bench.py times by default."""
import numpy as np
import pytest

from conftest import compare_detections, report

pytestmark = pytest.mark.gpu

S, B = 608, 32
# images of the batch (global synthetic indices) that are round-off stable in the sense of test_gpu_forward._stable_case:
# four oracle evaluations (fp32, fp64, two other fp32 summation orders) agree on every index.  Found by that search over
# indices 0..12 (4 and 8 qualify); the test re-verifies the property before relying on it.
STABLE = (4, 8)
OTHERS = (0, 17)          # two arbitrary images: compared tie-aware (conftest.compare_detections)


@pytest.fixture(scope='module')
def oracle_refs(weights):
    import y4_oracle as O
    W, _ = weights
    refs = {}
    for i in STABLE + OTHERS:
        im = O.synth_images(0, i, 1, S)
        r32 = O.decode_nms(O.forward(im, W), S)
        noise = 0.0
        stable = True
        for kw in ((dict(dtype=np.float64), dict(kperm_seed=1), dict(kperm_seed=2)) if i in STABLE else (dict(dtype=np.float64),)):
            r = O.decode_nms([h.astype(np.float32) for h in O.forward(im, W, **kw)], S)
            stable = stable and np.array_equal(r[4], r32[4])
            if np.array_equal(r[3], r32[3]):
                noise = max(noise, float(np.abs(np.sort(r[1]) - np.sort(r32[1])).max()))
        refs[i] = (r32, noise, stable)
    return refs


def _run(weights, precision):
    import y4b200
    import y4_oracle as O
    W, blob = weights
    eng = y4b200.Engine(img_size=S, max_batch=B, precision=precision)
    eng.load_darknet_bytes(blob)
    got = eng.predict(O.synth_images(0, 0, B, S), with_indices=True)
    # the resident path bench.py times (device-side generator, CUDA graph) must give the same bits as the host-input path
    eng.synth_fill(0, 0, B)
    eng.run_resident(B)
    res = eng.fetch_results(B)
    kinds = [l['kernel_kind'] for l in eng.layers()]
    eng.close()
    for a, b in zip(got, res):
        assert np.array_equal(a, b)
    return got, kinds


@pytest.mark.parametrize('mode', ['fp16x3', 'fp32'])
def test_batch32_608_parity_modes_vs_oracle(weights, oracle_refs, mode):
    import y4b200
    got, kinds = _run(weights, {'fp16x3': y4b200.PREC_FP16X3, 'fp32': y4b200.PREC_FP32}[mode])
    if mode == 'fp16x3':
        assert sum(k in (1, 2) for k in kinds) >= 100, kinds          # the tcgen05 path
    for i in STABLE:
        ref, noise, stable = oracle_refs[i]
        assert stable, f'image {i} is no longer round-off stable: regenerate STABLE'
        tol = max(1e-4, 3 * noise)
        g = [a[i:i + 1] for a in got]
        report(f'cfg2_{mode}_img{i}', tol=tol, idx_equal=bool(np.array_equal(g[4], ref[4])),
               box_err=float(np.abs(g[0] - ref[0]).max()), score_err=float(np.abs(g[1] - ref[1]).max()))
        if mode == 'fp16x3':           # the north-star gate, on the tensor-core path
            assert np.array_equal(g[3], ref[3]) and np.array_equal(g[4], ref[4]) and np.array_equal(g[2], ref[2])
            assert np.abs(g[0] - ref[0]).max() <= tol and np.abs(g[1] - ref[1]).max() <= tol
        else:
            # the CUDA-core fp32 kernel sums K sequentially in fp32: its round-off is LARGER than that of the chunked tensor-core
            # mode (measured: it swaps two detections 1e-5 apart on image 8), so it is held to the tie-aware comparison
            d = compare_detections(ref, g, tol)[0]
            assert d['unexplained'] == 0 and d['max_score_err'] <= tol and d['max_box_err'] <= tol, d
    for i in OTHERS:
        ref, noise, _ = oracle_refs[i]
        tol = max(1e-4, 3 * noise)
        d = compare_detections(ref, [a[i:i + 1] for a in got], tol)[0]
        report(f'cfg2_{mode}_img{i}_tie_aware', tol=tol, **d)
        assert d['unexplained'] == 0, d
        assert d['max_score_err'] <= tol and d['max_box_err'] <= tol, d


def test_batch32_608_fp16x3_agrees_with_fp32_engine_on_all_images(weights):
    """All 32 images: the tensor-core parity mode against the CUDA-core fp32 mode of the same engine (tie-aware: two fp32-grade
    evaluations may order near-equal scores differently)."""
    import y4b200
    a, _ = _run(weights, y4b200.PREC_FP16X3)
    b, _ = _run(weights, y4b200.PREC_FP32)
    rows = compare_detections(b, a, 3e-4)
    report('cfg2_fp16x3_vs_fp32_all32', exact=sum(r['exact'] for r in rows), moved=sum(r['moved'] for r in rows),
           swapped_in=sum(r['swapped_in'] for r in rows), unexplained=sum(r['unexplained'] for r in rows),
           max_score_err=max(r['max_score_err'] for r in rows), max_box_err=max(r['max_box_err'] for r in rows))
    assert sum(r['unexplained'] for r in rows) <= 2, rows        # an NMS decision within round-off of the IoU threshold may flip
    assert max(r['max_score_err'] for r in rows) <= 3e-4


def test_batch32_608_fp16_default_mode_is_close(weights, oracle_refs):
    """The mode bench.py times by default (fp16 operands AND activations, BASELINE configs[1] 'fp16 tensor cores') cannot meet an
    fp32 tolerance; what it does deliver is measured and bounded: most of the oracle's detections are re-found."""
    import y4b200
    from test_gpu_forward import _match_detections
    got, kinds = _run(weights, y4b200.PREC_FP16)
    assert sum(k in (1, 2, 4) for k in kinds) >= 109, kinds
    fr = []
    for i in STABLE + OTHERS:
        ref = oracle_refs[i][0]
        fr.append(_match_detections(ref, [a[i:i + 1] for a in got]))
    report('cfg2_fp16_detections_refound', fractions=fr)
    assert min(fr) >= 0.5, fr
