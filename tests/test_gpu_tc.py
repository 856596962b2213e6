"""tcgen05/TMA conv kernel vs the CUDA-core kernel on IDENTICAL device inputs, layer by layer
(y4_debug_run_conv), then vs the oracle at the heads.  Both kernels read the same fp16 activations and
fp16 weights and accumulate in fp32, so they may differ only by summation order and one fp16 rounding."""
import numpy as np
import pytest

from conftest import report

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize('size,batch', [(160, 2), (608, 1)])
def test_tc_vs_simt_per_layer(weights, size, batch):
    import y4b200
    import y4_oracle as O
    W, blob = weights
    imgs = O.synth_images(0, 0, batch, size)
    eng = y4b200.Engine(img_size=size, max_batch=batch, precision=y4b200.PREC_FP16)
    eng.load_darknet_bytes(blob)
    eng.forward_heads(imgs)
    layers = eng.layers()
    kinds = [l['kernel_kind'] for l in layers]
    assert sum(k > 0 for k in kinds) >= 100, kinds          # the hot path must actually be the tcgen05 path
    bad = []
    errs = {}
    for l in layers:
        if l['kernel_kind'] == 0:
            continue
        idx, name = l['idx'], l['out_name']
        eng.run_conv(idx, batch, 1)
        a = eng.get_tensor(name, batch)
        eng.run_conv(idx, batch, 0)
        b = eng.get_tensor(name, batch)
        scale = float(np.abs(b).max()) + 1e-12
        e = float(np.abs(a - b).max()) / scale
        errs[idx] = e
        if not np.isfinite(e) or e > 4e-3:
            bad.append((idx, name, l['kernel_kind'], l['cin'], l['cout'], l['ksize'], l['stride'], e))
    worst = sorted(errs.items(), key=lambda kv: -kv[1])[:8]
    report(f'tc_vs_simt_{size}', worst=worst, bad=bad, n=len(errs))
    assert not bad, bad
    eng.close()


def test_spp_is_exact_max(weights):
    """SPP slices (custom_layers.py:130-134) of the fp16 engine == numpy max-pools of the engine's own input slice."""
    import y4b200
    import y4_oracle as O
    W, blob = weights
    size, batch = 416, 2
    eng = y4b200.Engine(img_size=size, max_batch=batch, precision=y4b200.PREC_FP16)
    eng.load_darknet_bytes(blob)
    eng.forward_heads(O.synth_images(0, 0, batch, size))
    g = size // 32
    x = eng.get_tensor('c74', batch).reshape(batch, g, g, 512)
    for k in (13, 9, 5):
        got = eng.get_tensor(f'mp{k}', batch).reshape(batch, g, g, 512)
        assert np.array_equal(got, O.maxpool_same(x, k)), k
    cat = eng.get_tensor('cat6', batch).reshape(batch, g, g, 2048)
    assert np.array_equal(cat[..., 1536:], x)
    eng.close()
