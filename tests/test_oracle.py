"""CPU tests (-m "not gpu"): the oracle against independent implementations and the reference's structural
known-answers (the reference has no tests and no replayable golden vector -> parity unpinned)."""
import os

import numpy as np
import pytest

import netspec
import y4_oracle as O

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_netlist_known_answers():
    # utils.py:13-14: conv_layer_size = 110, conv_output_idxs = [93, 101, 109]
    cs = netspec.conv_ops()
    assert len(cs) == 110
    assert [c.idx for c in cs if not c.bn] == [93, 101, 109]
    assert all(c.act == 'linear' and c.cout == 255 for c in cs if not c.bn)
    # darknet yolov4.weights size implied by utils.py:16-41
    assert 20 + 4 * netspec.darknet_file_floats() == 257717640
    # census: 70 mish / 37 leaky / 3 linear; 66 1x1, 37 3x3 s1, 7 3x3 s2; convs 0,1 leaky
    acts = [c.act for c in cs]
    assert (acts.count('mish'), acts.count('leaky'), acts.count('linear')) == (70, 37, 3)
    ks = [(c.k, c.stride) for c in cs]
    assert (ks.count((1, 1)), ks.count((3, 1)), ks.count((3, 2))) == (66, 37, 7)
    assert cs[0].act == cs[1].act == 'leaky'
    for s, gf in ((320, 35.565), (416, 60.105), (512, 91.046), (608, 128.389), (768, 204.854)):
        assert abs(netspec.conv_gflop(s) - gf) < 1e-3
    ops, heads = netspec.build_netlist()
    cat = {o.out: o.ins for o in ops if o.kind == 'concat'}
    assert cat['cat1'] == ['c6', 'c2'] and cat['cat6'] == ['mp13', 'mp9', 'mp5', 'c74']      # main first, route second
    assert cat['cat7'] == ['c79', 'up_c78'] and cat['cat10'] == ['c102', 'c77']
    assert heads == ['c93', 'c101', 'c109']


def test_conv_against_torch():
    torch = pytest.importorskip('torch')
    F = torch.nn.functional
    rng = np.random.default_rng(0)
    for (cin, cout, k, stride, hw) in ((3, 8, 3, 1, 9), (16, 8, 1, 1, 6), (8, 16, 3, 2, 10), (8, 8, 3, 2, 12)):
        x = rng.standard_normal((2, hw, hw, cin)).astype(np.float32)
        w = rng.standard_normal((k, k, cin, cout)).astype(np.float32)
        got = O.conv2d_raw(x, w, stride)
        xt = torch.from_numpy(x).permute(0, 3, 1, 2)
        wt = torch.from_numpy(w).permute(3, 2, 0, 1)
        if stride == 1:
            ref = F.conv2d(xt, wt, padding=k // 2)
        else:                                           # ZeroPadding2D(((1,0),(1,0))) + valid stride 2
            ref = F.conv2d(F.pad(xt, (1, 0, 1, 0)), wt, stride=2)
        ref = ref.permute(0, 2, 3, 1).numpy()
        assert got.shape == ref.shape
        assert np.abs(got - ref).max() < 1e-4


def test_pool_upsample_act_against_torch():
    torch = pytest.importorskip('torch')
    F = torch.nn.functional
    rng = np.random.default_rng(1)
    x = rng.standard_normal((2, 13, 13, 6)).astype(np.float32)
    xt = torch.from_numpy(x).permute(0, 3, 1, 2)
    for k in (5, 9, 13):
        ref = F.max_pool2d(xt, k, stride=1, padding=k // 2).permute(0, 2, 3, 1).numpy()
        assert np.array_equal(O.maxpool_same(x, k), ref)
    assert np.array_equal(O.maxpool_same(O.maxpool_same(x, 5), 5), O.maxpool_same(x, 9))     # cascade identity
    ref = F.interpolate(xt, scale_factor=2, mode='nearest').permute(0, 2, 3, 1).numpy()
    assert np.array_equal(O.upsample2x(x), ref)
    v = np.linspace(-30, 30, 601).astype(np.float32)
    assert np.abs(O.mish(v) - F.mish(torch.from_numpy(v)).numpy()).max() < 1e-5
    assert np.abs(O.leaky(v) - F.leaky_relu(torch.from_numpy(v), 0.1).numpy()).max() == 0


def test_darknet_roundtrip_and_size_check(weights):
    W, blob = weights
    assert len(blob) == 257717640
    W2 = O.Weights.from_darknet_bytes(blob)
    for a, b in zip(W.p, W2.p):
        assert a.keys() == b.keys()
        for k in a:
            assert np.array_equal(a[k], b[k])
    with pytest.raises(ValueError):
        O.Weights.from_darknet_bytes(blob[:-4])
    # darknet order [beta, gamma, mean, var] then OIHW (utils.py:29-42): first conv's BN block sits right after the header
    beta0 = np.frombuffer(blob, '<f4', 32, 20)
    assert np.array_equal(beta0, W.p[0]['beta'])
    w0 = np.frombuffer(blob, '<f4', 32 * 3 * 9, 20 + 4 * 32 * 4).reshape(32, 3, 3, 3)
    assert np.array_equal(w0.transpose(2, 3, 1, 0), W.p[0]['w'])


def test_decode_matches_reference_formulas():
    """get_boxes (custom_layers.py:221-258) re-derived per element in float64."""
    rng = np.random.default_rng(2)
    g, nc, stride, xs = 4, 3, 16, 1.1
    pred = rng.standard_normal((1, g, g, 3 * (5 + nc))).astype(np.float32)
    box, obj, cls = O.get_boxes(pred, O.ANCHORS[1], nc, g, stride, xs)
    p = pred.reshape(1, g, g, 3, 5 + nc).astype(np.float64)
    for r in range(g):
        for c in range(g):
            for a in range(3):
                sx, sy = 1 / (1 + np.exp(-p[0, r, c, a, 0])), 1 / (1 + np.exp(-p[0, r, c, a, 1]))
                bx = (sx * xs - 0.5 * (xs - 1) + c) * stride          # grid[...,0] = column
                by = (sy * xs - 0.5 * (xs - 1) + r) * stride
                bw = np.exp(p[0, r, c, a, 2]) * O.ANCHORS[1][a, 0]
                bh = np.exp(p[0, r, c, a, 3]) * O.ANCHORS[1][a, 1]
                ref = [bx - bw / 2, by - bh / 2, bx + bw / 2, by + bh / 2]
                assert np.abs(box[0, r, c, a] - ref).max() < 1e-3
    assert np.abs(obj[..., 0] - 1 / (1 + np.exp(-p[..., 4]))).max() < 1e-6
    # flat index order: n = (row*g + col)*3 + a  (custom_layers.py:274-280)
    boxes, scores = O.decode_heads([np.zeros((1, 8, 8, 24), np.float32), pred, np.zeros((1, 2, 2, 24), np.float32)], 64, nc)
    n = 3 * 64 + (2 * g + 1) * 3 + 2
    assert np.allclose(boxes[0, n] * 64, box[0, 2, 1, 2], atol=1e-4)
    assert np.allclose(scores[0, n], obj[0, 2, 1, 2] * cls[0, 2, 1, 2])


def _brute_nms(boxes, scores, iou_thr, score_thr, max_per_class, max_total):
    """Independent restatement: vectorised IoU matrix + greedy scan."""
    out = []
    B, N, C = scores.shape
    for b in range(B):
        bb = boxes[b].astype(np.float32)
        y0, y1 = np.minimum(bb[:, 0], bb[:, 2]), np.maximum(bb[:, 0], bb[:, 2])
        x0, x1 = np.minimum(bb[:, 1], bb[:, 3]), np.maximum(bb[:, 1], bb[:, 3])
        area = ((y1 - y0) * (x1 - x0)).astype(np.float32)
        picked = []
        for c in range(C):
            idx = np.nonzero(scores[b, :, c] > np.float32(score_thr))[0]
            idx = idx[np.lexsort((idx, -scores[b, idx, c].astype(np.float64)))]
            keep = []
            for i in idx:
                ok = True
                for j in keep:
                    ih = np.float32(max(np.float32(min(y1[i], y1[j]) - max(y0[i], y0[j])), 0))
                    iw = np.float32(max(np.float32(min(x1[i], x1[j]) - max(x0[i], x0[j])), 0))
                    inter = np.float32(ih * iw)
                    iou = np.float32(0) if area[i] <= 0 or area[j] <= 0 else np.float32(inter / np.float32(np.float32(area[i] + area[j]) - inter))
                    if iou > np.float32(iou_thr):
                        ok = False
                        break
                if ok:
                    keep.append(i)
                    if len(keep) == max_per_class:
                        break
            picked += [(-float(scores[b, i, c]), c, int(i)) for i in keep]
        picked.sort()
        out.append(picked[:max_total])
    return out


def test_nms_against_bruteforce():
    for seed, size in ((1, 128), (2, 160)):
        heads = O.synth_heads(seed, 2, size, n_clusters=25)
        boxes, scores = O.decode_heads(heads, size)
        rb, rs, rc, rv, ri = O.combined_nms(boxes, scores)
        ref = _brute_nms(boxes, scores, O.IOU_THRESHOLD, O.SCORE_THRESHOLD, 100, 100)
        for b in range(2):
            assert rv[b] == len(ref[b])
            assert [int(i) for i in ri[b, :rv[b]]] == [t[2] for t in ref[b]]
            assert [int(c) for c in rc[b, :rv[b]]] == [t[1] for t in ref[b]]
            assert np.all(np.diff(rs[b, :rv[b]]) <= 0)
            assert (rb[b] >= 0).all() and (rb[b] <= 1).all()
            assert not rb[b, rv[b]:].any() and (ri[b, rv[b]:] == -1).all()


def test_nms_semantics_edge_cases():
    # two identical boxes, same class: second suppressed (iou 1 > thr); different class: both kept
    boxes = np.array([[[0.1, 0.1, 0.5, 0.5], [0.1, 0.1, 0.5, 0.5], [0.6, 0.6, 0.9, 0.9]]], np.float32)
    scores = np.zeros((1, 3, 2), np.float32)
    scores[0, 0, 0], scores[0, 1, 0], scores[0, 2, 1] = 0.9, 0.8, 0.7
    rb, rs, rc, rv, ri = O.combined_nms(boxes, scores)
    assert rv[0] == 2 and ri[0, :2].tolist() == [0, 2] and rc[0, :2].tolist() == [0, 1]
    scores[0, 1, 0], scores[0, 1, 1] = 0, 0.8
    assert O.combined_nms(boxes, scores)[3][0] == 3
    # strict '>' on the score threshold; clip on output only; degenerate (zero-area) boxes never suppress
    scores[:] = 0; scores[0, 0, 0] = np.float32(0.3)
    assert O.combined_nms(boxes, scores)[3][0] == 0
    boxes2 = np.array([[[-0.2, 0.2, 1.3, 0.8], [0.4, 0.4, 0.4, 0.9]]], np.float32)
    sc2 = np.array([[[0.9], [0.8]]], np.float32)
    rb, rs, rc, rv, ri = O.combined_nms(boxes2, sc2)
    assert rv[0] == 2 and rb[0, 0].tolist() == pytest.approx([0, 0.2, 1, 0.8])
    # empty input
    rv = O.combined_nms(np.zeros((1, 0, 4), np.float32), np.zeros((1, 0, 3), np.float32))[3]
    assert rv[0] == 0


def test_synth_images_shard_invariance():
    a = O.synth_images(3, 0, 4, 32)
    b = np.concatenate([O.synth_images(3, 0, 2, 32), O.synth_images(3, 2, 2, 32)])
    assert np.array_equal(a, b) and a.dtype == np.float32 and 0 <= a.min() and a.max() < 1


def test_oracle_predict_small(weights):
    W, _ = weights
    imgs = O.synth_images(0, 0, 1, 96)
    m = {}
    boxes, scores, classes, valid, idx = O.predict(imgs, W, margins=m)
    assert boxes.shape == (1, 100, 4) and scores.shape == (1, 100) and valid.shape == (1,)
    n = 3 * (12 * 12 + 6 * 6 + 3 * 3)
    assert (idx[0, :valid[0]] < n).all()
    heads = O.forward(imgs, W)
    heads_f = O.forward(imgs, W, fold_bn=True)
    for a, b in zip(heads, heads_f):                    # BN folding is exact up to fp32 round-off
        assert np.abs(a - b).max() / np.abs(a).max() < 1e-4


# ---- preprocess_img: oracle restatement pinned to the reference's dependency (cv2.resize) -------------------------
def _golden_pre():
    import os
    return np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), 'golden', 'preprocess_street.npz'))


def test_resize_restatement_matches_cv2_golden():
    """Golden vectors produced by cv2.resize on the reference's img/street.jpeg (tests/golden/make_golden_preprocess.py):
    the fixed-point restatement is bit-exact, for down-, up-scaling and odd sizes."""
    import hashlib
    import y4_oracle as O
    g = _golden_pre()
    raw = g['raw']
    assert np.array_equal(O.resize_linear_u8(raw, 160, 160), g['resize_160'])
    for S in (320, 416, 608):
        r = O.resize_linear_u8(raw, S, S)
        assert hashlib.sha256(r.tobytes()).hexdigest() == str(g[f'sha256_{S}'])
        pre = O.preprocess_img(raw, S)
        assert pre.dtype == np.float64 and pre.shape == (S, S, 3)
        assert abs(float(pre.astype(np.float32).astype(np.float64).sum()) - float(g[f'pre_f32_sum_{S}'])) < 1e-6
    assert np.array_equal(O.resize_linear_u8(g['crop'], 64, 96), g['crop_resize_64x96'])
    assert np.array_equal(O.resize_linear_u8(g['crop'], 301, 257), g['crop_resize_301x257'])


def test_resize_restatement_matches_cv2_live():
    """Where cv2 is importable (this image), compare against it directly on random images and sizes."""
    cv2 = pytest.importorskip('cv2')
    import y4_oracle as O
    rng = np.random.default_rng(7)
    for (h, w, S) in [(37, 53, 64), (333, 517, 416), (700, 300, 320), (64, 64, 608), (5, 9, 96)]:
        img = rng.integers(0, 256, (h, w, 3), dtype=np.uint8)
        assert np.array_equal(O.resize_linear_u8(img, S, S), cv2.resize(img, (S, S))), (h, w, S)


# ---- the oracle vs the reference's OWN source (tests/golden/make_golden_refgraph.py) -------------------------------
@pytest.fixture(scope='module')
def refgraph():
    import os
    return np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), 'golden', 'refgraph_416.npz'))


def test_loader_matches_reference_load_weights(refgraph, weights):
    """utils.py:12-53 executed from the reference's source on the synthetic darknet file: every array it hands to Keras
    (HWIO kernel, head bias, [gamma, beta, mean, var]) equals what the oracle's loader parses, layer by layer."""
    import hashlib
    W, blob = weights
    assert hashlib.sha256(blob).hexdigest() == str(refgraph['weights_sha256'])
    P = O.Weights.from_darknet_bytes(blob)
    want = dict(zip([str(k) for k in refgraph['loader_keys']], [str(v) for v in refgraph['loader_sha256']]))
    assert len(want) == 110 + 107
    for i, (o, p) in enumerate(zip(P.convs, P.p)):
        h = hashlib.sha256(np.ascontiguousarray(p['w'], dtype=np.float32).tobytes())
        if not o.bn:
            h.update(np.ascontiguousarray(p['bias'], dtype=np.float32).tobytes())
        assert h.hexdigest() == want[f'conv{i}'], i
        if o.bn:
            bn = np.stack([p['gamma'], p['beta'], p['mean'], p['var']]).astype(np.float32)
            assert hashlib.sha256(bn.tobytes()).hexdigest() == want[f'bn{i}'], i


def test_forward_matches_reference_graph(refgraph, weights):
    """custom_layers.py:5-198 executed from the reference's source (eager numpy/torch stand-in for the Keras layers) at
    416x416: the oracle's heads agree at 2,700 sampled positions and in their sums.  Pins conv order, concat order, SPP order,
    add-after-activation, the asymmetric stride-2 padding and which convs are leaky / mish / linear."""
    W, _ = weights
    S = int(refgraph['img_size'])
    heads = O.forward(O.synth_images(0, 0, 1, S), W)
    for i, h in enumerate(heads):
        idx, val, amax = refgraph[f'head{i}_idx'], refgraph[f'head{i}_val'], float(refgraph[f'head{i}_absmax'])
        got = h.reshape(-1)[idx]
        assert np.abs(got - val).max() <= 2e-4 * amax, (i, float(np.abs(got - val).max()), amax)
        assert abs(float(h.astype(np.float64).sum()) - float(refgraph[f'head{i}_sum'])) <= 1e-5 * amax * h.size


def test_decode_matches_reference_get_boxes_and_flattening(refgraph, weights):
    """custom_layers.py:201-298 executed from the reference's source up to the TF NMS op.  (a) identical sparse heads in:
    the tensors the reference passes to combined_non_max_suppression (boxes (N,1,4) / img_size, scores = conf * cls (N,80))
    equal the oracle's, including the flat order n = off_scale + (row*g+col)*3 + a.  (b) the network's own heads: the 2,000
    best (box, class) pairs and the candidate count above the threshold."""
    W, _ = weights
    S = int(refgraph['img_size'])
    mo, mt, iou, thr = refgraph['nms_args']
    assert (int(mo), int(mt)) == (O.MAX_BOXES, O.MAX_BOXES) and abs(iou - O.IOU_THRESHOLD) < 1e-7 and abs(thr - O.SCORE_THRESHOLD) < 1e-7
    sparse = [refgraph[f'sparse_head{i}'] for i in range(3)]
    boxes, scores = O.decode_heads(sparse, S)
    assert boxes.shape == (1, 10647, 4) and scores.shape == (1, 10647, 80)
    assert np.abs(boxes[0] - refgraph['b_nms_boxes']).max() <= 2e-6
    nz = refgraph['b_nms_scores_nonzero_idx']
    got_nz = np.nonzero(scores[0].reshape(-1) > 1e-4)[0]
    assert np.array_equal(got_nz, nz)
    assert np.abs(scores[0].reshape(-1)[nz] - refgraph['b_nms_scores_nonzero']).max() <= 1e-6
    heads = O.forward(O.synth_images(0, 0, 1, S), W)
    boxes, scores = O.decode_heads(heads, S)
    flat = refgraph['a_top_flat']
    assert np.abs(scores[0].reshape(-1)[flat] - refgraph['a_top_scores']).max() <= 2e-4
    assert np.abs(boxes[0][flat // 80] - refgraph['a_top_boxes']).max() <= 2e-4
    cnt = int((scores > O.SCORE_THRESHOLD).sum())
    assert abs(cnt - int(refgraph['a_count_above_thr'])) <= 3, (cnt, int(refgraph['a_count_above_thr']))


def test_nms_against_torchvision():
    """Independent library cross-check of the restated combined_non_max_suppression: torchvision.ops.nms (greedy, suppress when
    IoU > thr, the same rule TF uses) per class on the candidates above the score threshold, then the same top-100 merge.
    torchvision computes IoU in a different floating-point order, so the crafted heads keep every IoU away from the threshold
    (margins checked); the selected (box, class) lists must then be identical."""
    tv = pytest.importorskip('torchvision')
    import torch
    S = 160
    heads = O.synth_heads(seed=11, batch=2, img_size=S, n_clusters=40)
    boxes, scores = O.decode_heads(heads, S)
    m = {}
    ob, osc, ocl, ov, oidx = O.combined_nms(boxes, scores, margins=m)
    assert m['iou'] > 1e-5 and m['score'] > 1e-6, m
    for b in range(2):
        picked = []
        for c in range(scores.shape[2]):
            n = np.nonzero(scores[b, :, c] > np.float32(O.SCORE_THRESHOLD))[0]
            if n.size == 0:
                continue
            s = scores[b, n, c]
            order = np.lexsort((n, -s.astype(np.float64)))          # torchvision visits by score; make ties explicit the same way
            n, s = n[order], s[order]
            bb = boxes[b, n]
            bb = np.stack([np.minimum(bb[:, 0], bb[:, 2]), np.minimum(bb[:, 1], bb[:, 3]), np.maximum(bb[:, 0], bb[:, 2]), np.maximum(bb[:, 1], bb[:, 3])], 1)
            keep = tv.ops.nms(torch.from_numpy(bb.astype(np.float64)), torch.from_numpy(s.astype(np.float64)), float(np.float32(O.IOU_THRESHOLD))).numpy()
            keep = np.sort(keep)[:O.MAX_BOXES]                      # indices into the score-ordered list
            picked += [(float(s[k]), c, int(n[k])) for k in keep]
        picked.sort(key=lambda t: (-t[0], t[1], t[2]))
        picked = picked[:O.MAX_BOXES]
        assert ov[b] == len(picked)
        assert [p[2] for p in picked] == oidx[b, :ov[b]].tolist()
        assert [p[1] for p in picked] == ocl[b, :ov[b]].astype(int).tolist()


def test_c_nms_restatement_equals_python_oracle():
    """oracle/nms_ref.c (compiled CPU baseline of combined_non_max_suppression) == y4_oracle.combined_nms, bit for bit,
    including runtime thresholds, ties (all-equal scores), the empty case and the top-100 cut."""
    import subprocess
    src, lib = os.path.join(ROOT, 'oracle', 'nms_ref.c'), os.path.join(ROOT, 'oracle', 'libnmsref.so')
    if not os.path.exists(lib) or os.path.getmtime(lib) < os.path.getmtime(src):
        subprocess.run(['gcc', '-O2', '-fPIC', '-shared', '-ffp-contract=off', '-fopenmp', src, '-o', lib, '-lm'], check=True)
    import y4_cpu_fast as F
    for seed, S, clusters, iou, thr in ((3, 320, 150, 0.413, 0.3), (4, 160, 40, 0.6, 0.1), (5, 160, 400, 0.2, 0.5)):
        heads = O.synth_heads(seed=seed, batch=2, img_size=S, n_clusters=clusters)
        boxes, scores = O.decode_heads(heads, S)
        ref = O.combined_nms(boxes, scores, iou, thr)
        got = F.combined_nms_c(boxes, scores, iou, thr)
        for a, b in zip(ref, got):
            assert np.array_equal(a, b)
    hot = [np.full((1, 64 // s, 64 // s, 255), 10.0, np.float32) for s in (8, 16, 32)]
    cold = [np.full_like(h, -20.0) for h in hot]
    for heads in (hot, cold):
        boxes, scores = O.decode_heads(heads, 64)
        for a, b in zip(O.combined_nms(boxes, scores), F.combined_nms_c(boxes, scores)):
            assert np.array_equal(a, b)


def test_torch_cpu_forward_matches_numpy_oracle():
    """oracle/y4_cpu_fast.TorchNet (oneDNN convs; bench.py's CPU baseline) follows the same netlist and op order as the
    numpy oracle: heads agree to fp32 round-off of a 110-layer evaluation."""
    import y4_cpu_fast as F
    W = O.synth_weights(seed=1)
    imgs = O.synth_images(0, 0, 2, 96)
    want = O.forward(imgs, W)
    got = F.TorchNet(W).forward(imgs)
    for a, b in zip(got, want):
        assert a.shape == b.shape
        assert float(np.abs(a - b).max() / np.abs(b).max()) < 3e-4
