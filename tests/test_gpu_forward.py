"""GPU parity of the 110-conv forward through the C-ABI vs the oracle (custom_layers.py:5-198)."""
import numpy as np
import pytest

from conftest import report

pytestmark = pytest.mark.gpu


def _rel(a, b):
    return float(np.abs(a - b).max() / (np.abs(b).max() + 1e-12))


@pytest.mark.parametrize('size,batch', [(160, 2)])
def test_fp32_layer_by_layer(weights, size, batch):
    """fp32 CUDA-core mode, every materialised tensor.  A 110-layer fp32 evaluation is only reproducible to the
    round-off it accumulates (oracle fp32 vs oracle fp64: ~6e-5 at the last layers on this net), so the engine is
    held to the same distance from the float64 evaluation as the fp32 oracle itself (x4 + 2e-6 slack), and the
    heads additionally to 1e-4 of the fp32 oracle."""
    import y4b200
    import y4_oracle as O
    W, blob = weights
    imgs = O.synth_images(0, 0, batch, size)
    k32, k64 = {}, {}
    heads = O.forward(imgs, W, keep=k32)
    O.forward(imgs, W, np.float64, keep=k64)
    eng = y4b200.Engine(img_size=size, max_batch=batch, precision=y4b200.PREC_FP32)
    eng.load_darknet_bytes(blob)
    got_heads = eng.forward_heads(imgs)
    rows = []
    for name, ref in k64.items():
        try:
            got = eng.get_tensor(name, batch).reshape(ref.shape)
        except y4b200.Y4Error:
            continue                      # fused away (conv feeding a residual add / upsample)
        rows.append((name, _rel(got, ref), _rel(k32[name], ref)))
    worst = sorted(rows, key=lambda r: -r[1])[:5]
    report('fp32_layerwise', n=len(rows), worst=worst, heads_vs_fp32=[_rel(a, b) for a, b in zip(got_heads, heads)])
    assert len(rows) >= 90
    for name, e_eng, e_ora in rows:
        assert e_eng <= 4 * e_ora + 2e-6, (name, e_eng, e_ora)
    for a, b in zip(got_heads, heads):
        assert _rel(a, b) < 1e-4


_STABLE_CACHE = {}


def _stable_case(W, size, tries=12, start=0):
    """Find a seeded image whose detections do not hinge on fp32 round-off.  TF's summation order is unspecified and the
    outputs hold ~100 scores whose closest pair is typically 1e-5 apart, so for most images two equally valid fp32
    evaluations of the reference already disagree on the ORDER of two detections (SURVEY §7 'hard parts': ties and
    near-threshold cases are implementation-defined in TF too).  'Bit-exact indices' is therefore tested on images where
    four evaluations of the oracle agree on every index -- fp32, fp64, and fp32 with the K terms of every dot product
    summed in two other (seeded random) orders -- i.e. on images that are round-off stable for ANY correct fp32
    implementation, not for ours in particular.  `noise` = the largest score / coordinate difference among them."""
    import y4_oracle as O
    if (size, start) in _STABLE_CACHE:
        return _STABLE_CACHE[(size, start)]
    for first in range(start, start + tries):
        imgs = O.synth_images(0, first, 1, size)
        r32 = O.decode_nms(O.forward(imgs, W), size)
        if r32[3][0] == 0:
            continue
        noise, ok = 0.0, True
        for kw in (dict(dtype=np.float64), dict(kperm_seed=1), dict(kperm_seed=2)):
            r = O.decode_nms([h.astype(np.float32) for h in O.forward(imgs, W, **kw)], size)
            if not np.array_equal(r[4], r32[4]):
                ok = False
                break
            noise = max(noise, float(np.abs(r[0] - r32[0]).max()), float(np.abs(r[1] - r32[1]).max()))
        if ok:
            _STABLE_CACHE[(size, start)] = (imgs, r32, noise, first)
            return imgs, r32, noise, first
    raise AssertionError('no round-off-stable synthetic image found')


# first image index to try per size (found with the same _stable_case search; the test re-verifies the stability)
STABLE_START = {256: 3, 416: 0, 608: 4}


@pytest.mark.parametrize('size', [256, 416, 608])
def test_fp32_predict_end_to_end_on_roundoff_stable_image(weights, size):
    """inference_model.predict parity (BASELINE config 1 at 416, the bench resolution 608): bit-exact indices / classes /
    valid on a round-off-stable image (_stable_case); boxes and scores within 1e-4 (north_star tolerance), widened only
    if fp32 round-off of the oracle itself exceeds it."""
    import y4b200
    import y4_oracle as O
    W, blob = weights
    imgs, ref, noise, first = _stable_case(W, size, start=STABLE_START[size])
    tol = max(1e-4, 3 * noise)
    eng = y4b200.Engine(img_size=size, max_batch=1, precision=y4b200.PREC_FP32)
    eng.load_darknet_bytes(blob)
    got = eng.predict(imgs, with_indices=True)
    report(f'fp32_predict_{size}', first=first, oracle_noise=noise, tol=tol, valid=ref[3].tolist(), got_valid=got[3].tolist(),
           box_err=float(np.abs(got[0] - ref[0]).max()), score_err=float(np.abs(got[1] - ref[1]).max()),
           idx_equal=bool(np.array_equal(got[4], ref[4])))
    assert np.array_equal(got[3], ref[3])
    assert np.array_equal(got[4], ref[4])
    assert np.array_equal(got[2], ref[2])
    assert np.abs(got[0] - ref[0]).max() <= tol
    assert np.abs(got[1] - ref[1]).max() <= tol
    eng.close()


def _match_detections(ref, got, iou_thr=0.9, score_tol=0.05):
    """fraction of reference detections (image 0) matched by a same-class engine detection with IoU > iou_thr."""
    rb, rs, rc, rv = ref[0][0], ref[1][0], ref[2][0], int(ref[3][0])
    gb, gs, gc, gv = got[0][0], got[1][0], got[2][0], int(got[3][0])
    hit = 0
    for i in range(rv):
        best = 0.0
        for j in range(gv):
            if gc[j] != rc[i] or abs(gs[j] - rs[i]) > score_tol:
                continue
            x1, y1 = max(rb[i, 0], gb[j, 0]), max(rb[i, 1], gb[j, 1])
            x2, y2 = min(rb[i, 2], gb[j, 2]), min(rb[i, 3], gb[j, 3])
            inter = max(x2 - x1, 0) * max(y2 - y1, 0)
            ua = (rb[i, 2] - rb[i, 0]) * (rb[i, 3] - rb[i, 1]) + (gb[j, 2] - gb[j, 0]) * (gb[j, 3] - gb[j, 1]) - inter
            best = max(best, inter / ua if ua > 0 else 0.0)
        hit += best > iou_thr
    return hit / max(rv, 1)


def test_fp16_heads_close(weights):
    """fp16 tensor-core mode (fp16 operands/activations, fp32 accumulate): it cannot meet the 1e-4 fp32 tolerance -
    no fp16 pipeline can - so its deviation from the fp32 oracle is MEASURED (gpurun_out/test_report.jsonl)
    and bounded: relative RMS of the raw heads < 8e-2 and >= 60 % of the oracle's detections re-found (same class,
    IoU > 0.9, |score diff| < 0.05).  Measured on B200 at 416: fp16 ACTIVATIONS alone (CUDA-core kernels, fp32 weights
    and math) give 1.6-3.8 % / 88 %; the tensor-core path also rounds the WEIGHTS to fp16: 2.6-6.1 % / 73 %.  That is
    the price of fp16 on this random, un-trained 110-layer network; Y4_PREC_FP32 is the parity mode."""
    import y4b200
    import y4_oracle as O
    W, blob = weights
    size, batch = 416, 1
    imgs = O.synth_images(0, 0, batch, size)
    heads = O.forward(imgs, W)
    ref = O.decode_nms(heads, size)
    for prec, tag in ((y4b200.PREC_FP16_SIMT, 'fp16_simt'), (y4b200.PREC_FP16, 'fp16_tc')):
        eng = y4b200.Engine(img_size=size, max_batch=batch, precision=prec)
        eng.load_darknet_bytes(blob)
        got = eng.forward_heads(imgs)
        rms = [float(np.sqrt(np.mean((a - b) ** 2)) / np.sqrt(np.mean(b ** 2))) for a, b in zip(got, heads)]
        mx = [_rel(a, b) for a, b in zip(got, heads)]
        det = eng.predict(imgs)
        agree = _match_detections(ref, det)
        report(tag + '_heads', rel_rms=rms, rel_max=mx, detections_refound=agree, valid=int(det[3][0]), ref_valid=int(ref[3][0]))
        assert max(rms) < 8e-2, rms
        assert agree >= 0.6, agree
        eng.close()


def test_synth_fill_matches_oracle(weights):
    """Device-side generator == oracle generator, bit for bit: conv 0 of the resident path equals conv 0 of the
    host path fed with the oracle's images."""
    import y4b200
    import y4_oracle as O
    W, blob = weights
    S, B = 64, 4
    eng = y4b200.Engine(img_size=S, max_batch=B, precision=y4b200.PREC_FP32)
    eng.load_darknet_bytes(blob)
    eng.synth_fill(11, 5, B)
    eng.run_forward_resident(B)
    a = eng.get_tensor('c0', B)
    eng.forward_heads(O.synth_images(11, 5, B, S))
    b = eng.get_tensor('c0', B)
    assert np.array_equal(a, b)
    eng.close()


def test_error_paths(weights):
    import y4b200
    W, blob = weights
    eng = y4b200.Engine(img_size=64, max_batch=1, precision=y4b200.PREC_FP32)
    imgs = np.zeros((1, 64, 64, 3), np.float32)
    with pytest.raises(y4b200.Y4Error) as ei:
        eng.predict(imgs)
    assert ei.value.code == -4                          # weights not loaded
    with pytest.raises(y4b200.Y4Error) as ei:
        eng.load_darknet_bytes(blob[:-4])
    assert ei.value.code == -3                          # wrong byte count (utils.py:50-53)
    eng.load_darknet_bytes(blob)
    with pytest.raises(y4b200.Y4Error):
        eng.predict(np.zeros((2, 64, 64, 3), np.float32))   # batch > max_batch
    eng.close()


def test_submit_collect_matches_predict(weights):
    """Pipelined host path (y4_submit / y4_collect, depth 2) returns exactly what the blocking y4_predict returns,
    in submission order."""
    import y4b200
    import y4_oracle as O
    W, blob = weights
    S, B = 160, 2
    eng = y4b200.Engine(img_size=S, max_batch=B, precision=y4b200.PREC_FP16)
    eng.load_darknet_bytes(blob)
    batches = [O.synth_images(0, i * B, B, S) for i in range(4)]
    ref = [eng.predict(x, with_indices=True) for x in batches]
    got = []
    eng.submit(batches[0])
    for i in range(1, 4):
        eng.submit(batches[i])
        got.append(eng.collect(with_indices=True))
    got.append(eng.collect(with_indices=True))
    with pytest.raises(y4b200.Y4Error):
        eng.collect()
    for r, g in zip(ref, got):
        for a, b in zip(r, g):
            assert np.array_equal(a, b)
    eng.close()


def test_preprocess_u8_matches_cv2_golden_and_oracle():
    """GPU preprocess (resize + /255) == cv2.resize golden vectors of the reference's image, bit for bit, and == the
    oracle on random images of mixed sizes in one batch; reverse_channels is predict()'s BGR->RGB flip."""
    import os
    import y4b200
    import y4_oracle as O
    g = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), 'golden', 'preprocess_street.npz'))
    raw = g['raw']
    eng = y4b200.Engine(img_size=160, max_batch=4, precision=y4b200.PREC_FP16)
    got = eng.preprocess_u8([raw])
    want = (g['resize_160'] / 255.).astype(np.float32)
    assert np.array_equal(got[0], want)
    rng = np.random.default_rng(3)
    imgs = [rng.integers(0, 256, (h, w, 3), dtype=np.uint8) for (h, w) in [(37, 53), (333, 517), (700, 300), (160, 160)]]
    got = eng.preprocess_u8(imgs)
    for i, im in enumerate(imgs):
        assert np.array_equal(got[i], O.preprocess_img(im, 160).astype(np.float32)), i
    flipped = eng.preprocess_u8(imgs, reverse_channels=True)
    assert np.array_equal(flipped, got[..., ::-1])
    eng.close()


def test_predict_u8_equals_predict_on_preprocessed(weights):
    """y4_predict_u8 / y4_submit_u8 == y4_predict on the oracle-preprocessed float images."""
    import y4b200
    import y4_oracle as O
    W, blob = weights
    S, B = 160, 2
    eng = y4b200.Engine(img_size=S, max_batch=B, precision=y4b200.PREC_FP16)
    eng.load_darknet_bytes(blob)
    rng = np.random.default_rng(5)
    raws = [rng.integers(0, 256, (120, 200, 3), dtype=np.uint8), rng.integers(0, 256, (300, 180, 3), dtype=np.uint8)]
    ref = eng.predict(np.stack([O.preprocess_img(r, S) for r in raws]), with_indices=True)
    got = eng.predict_u8(raws, with_indices=True)
    for a, b in zip(got, ref):
        assert np.array_equal(a, b)
    eng.submit_u8(raws)
    got2 = eng.collect(with_indices=True)
    for a, b in zip(got2, ref):
        assert np.array_equal(a, b)
    eng.close()
