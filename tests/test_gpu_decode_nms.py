"""GPU parity: decode + score filter + per-class NMS through the C-ABI (y4_decode_nms) vs the oracle
(custom_layers.py:201-298 + TF CombinedNonMaxSuppression semantics).  Bit-exact indices / classes / valid,
coordinates and scores within 1e-4 (north_star tolerance)."""
import numpy as np
import pytest

from conftest import report

pytestmark = pytest.mark.gpu
TOL = 1e-4


def _check(got, ref, tag):
    gb, gs, gc, gv, gi = got
    rb, rs, rc, rv, ri = ref
    report(tag, valid=rv.tolist(), got_valid=gv.tolist(), max_box_err=float(np.abs(gb - rb).max()),
           max_score_err=float(np.abs(gs - rs).max()), idx_equal=bool(np.array_equal(gi, ri)))
    assert np.array_equal(gv, rv)
    assert np.array_equal(gi, ri)
    assert np.array_equal(gc, rc)
    assert np.abs(gb - rb).max() <= TOL
    assert np.abs(gs - rs).max() <= TOL


@pytest.mark.parametrize('size,batch,clusters', [(416, 2, 150), (608, 1, 150), (320, 3, 40)])
def test_decode_nms_config4(size, batch, clusters):
    import y4b200
    import y4_oracle as O
    heads = O.synth_heads(seed=7 + size, batch=batch, img_size=size, n_clusters=clusters)
    m = {}
    ref = O.decode_nms(heads, size, margins=m)
    assert m['score'] > 1e-6 and m['iou'] > 1e-5, m          # test-data hygiene (SURVEY §8d cfg 4)
    eng = y4b200.Engine(img_size=size, max_batch=batch)
    got = eng.decode_nms(heads, with_indices=True)
    _check(got, ref, f'decode_nms_{size}')
    # resident path (padded head buffers) must agree with the packed-user-heads path
    eng.upload_heads(heads)
    eng.run_decode_nms_resident(batch)
    got2 = eng.fetch_results(batch)
    for a, b in zip(got, got2):
        assert np.array_equal(a, b)
    eng.close()


def test_runtime_thresholds_and_empty():
    import y4b200
    import y4_oracle as O
    S = 320
    heads = O.synth_heads(seed=5, batch=2, img_size=S, n_clusters=30)
    eng = y4b200.Engine(img_size=S, max_batch=2)
    for iou, sc in ((0.413, 0.1), (0.6, 0.5), (0.2, 0.9)):
        ref = O.decode_nms(heads, S, iou_threshold=iou, score_threshold=sc)
        got = eng.decode_nms(heads, iou, sc, with_indices=True)
        _check(got, ref, f'thr_{iou}_{sc}')
    # nothing above threshold -> valid = 0, all-zero outputs, idx = -1
    empty = [np.full_like(h, -20.0) for h in heads]
    got = eng.decode_nms(empty, with_indices=True)
    assert got[3].tolist() == [0, 0] and not got[0].any() and not got[1].any() and (got[4] == -1).all()
    eng.close()


def test_capacity_error_is_loud():
    import y4b200
    S = 320
    eng = y4b200.Engine(img_size=S, max_batch=1)
    hot = [np.full((1, S // s, S // s, 255), 10.0, np.float32) for s in (8, 16, 32)]   # every (box,class) passes
    with pytest.raises(y4b200.Y4Error) as ei:
        eng.decode_nms(hot)
    assert ei.value.code == -5
    eng.close()
