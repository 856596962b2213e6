"""Determinism across plans, devices and GPU counts.

DESIGN.md claims that every tcgen05 plan of a layer (single CTA / CTA pair, any N tile, ring depth, epilogue variant)
accumulates K in the same order and rounds once, so the autotuner -- which is free to pick different plans on every
engine creation and on every GPU -- never changes a bit of the output.  The data-parallel gather relies on exactly that
(SURVEY §8(d) cfg 3: gathered results identical to the 1-GPU run).  These tests force plan families and compare bits."""
import glob
import os
import subprocess
import sys

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
NGPU = len(glob.glob('/dev/nvidia[0-9]*'))

# Y4_FORCE = "bn,smemKB,patch,group,epi,nepi,bres,gw,cta2,lean,pairx" (y4_engine.cu): that plan wherever it applies
FAMILIES = {
    'autotuned': None,
    'single_cta_bn128_slab4': '128,224,0,1,1,4,0,32,0,0,0',
    'cta_pair_bn256_slab8': '256,224,0,2,1,8,0,32,1,0,0',
    'cta_pair_bn128_gw64': '128,224,0,3,1,4,0,64,1,0,0',
    'cta_pair_patch_bn128': '128,224,1,1,1,8,0,32,1,0,0',
    'cta_pair_bn128_16_epilogue_warps': '128,224,0,1,1,16,0,32,1,0,0',
    'cta_pair_patch_bn256_16_epilogue_warps': '256,224,1,1,1,16,0,32,1,0,0',
    'cta_pair_patch_bn256': '256,224,1,1,1,4,0,32,1,0,0',
    'bn64_per_thread_stores': '64,112,0,1,0,4,0,32,0,0,0',
    'lean_resident_w_gw64': '64,75,0,1,1,4,1,64,0,1,0',
}


def _heads_with(force, precision, blob, imgs, size, batch):
    import y4b200
    old = os.environ.pop('Y4_FORCE', None)
    if force:
        os.environ['Y4_FORCE'] = force
    try:
        eng = y4b200.Engine(img_size=size, max_batch=batch, precision=precision)
    finally:
        os.environ.pop('Y4_FORCE', None)
        if old is not None:
            os.environ['Y4_FORCE'] = old
    eng.load_darknet_bytes(blob)
    heads = eng.forward_heads(imgs)
    plans = sorted({(l['tc_mode'], l['tile_n'], l['tc_epilogue'], l['tc_epi_warps'], l['tc_resident_w']) for l in eng.layers() if l['kernel_kind'] in (1, 2)})
    mid = [eng.get_tensor(n, batch) for n in ('c1', 'r3', 'c37', 'c77', 'cat9')]
    eng.close()
    return heads, mid, plans


@pytest.mark.parametrize('precision_name', ['fp16', 'fp16x3'])
def test_plan_families_give_identical_bits(weights, precision_name):
    import y4b200
    import y4_oracle as O
    W, blob = weights
    size, batch = 160, 2
    imgs = O.synth_images(0, 0, batch, size)
    prec = {'fp16': y4b200.PREC_FP16, 'fp16x3': y4b200.PREC_FP16X3}[precision_name]
    fams = FAMILIES if precision_name == 'fp16' else {'autotuned': None, 'bn64': '64,224,0,1,0,4,0,32,0,0,0', 'bn128_group2': '128,224,0,2,0,4,0,32,0,0,0'}
    ref = None
    seen = set()
    for name, force in fams.items():
        heads, mid, plans = _heads_with(force, prec, blob, imgs, size, batch)
        seen.update(plans)
        if ref is None:
            ref = (heads, mid)
            continue
        for a, b in zip(heads + mid, ref[0] + ref[1]):
            assert np.array_equal(a, b), name
    assert len(seen) >= (6 if precision_name == 'fp16' else 2), seen       # the families really are different kernels


@pytest.mark.skipif(NGPU < 2, reason='needs two GPUs in one box')
def test_two_engines_on_two_devices_in_one_process(weights):
    """y4.h: independent engines (one per GPU) may be used concurrently from one process.  448x448 so that the SPP kernel
    needs its > 48 KB dynamic shared memory opt-in on BOTH devices (per-device function attributes)."""
    import y4b200
    import y4_oracle as O
    W, blob = weights
    size, batch = 448, 1
    imgs = O.synth_images(0, 0, batch, size)
    engs = [y4b200.Engine(img_size=size, max_batch=batch, precision=y4b200.PREC_FP16, device=d) for d in (0, 1)]
    outs = []
    for e in engs:
        e.load_darknet_bytes(blob)
        outs.append(e.predict(imgs, with_indices=True))
    for a, b in zip(*outs):
        assert np.array_equal(a, b)
    for e in engs:
        e.close()


@pytest.mark.skipif(NGPU < 2, reason='needs two GPUs in one box')
def test_dp_gather_equals_single_gpu_bitwise():
    """SURVEY §8(d) cfg 3 at the box's GPU count: tools/check_dp.py under torchrun, NCCL all-gather of the result records,
    rank 0 compares with its own single-engine run of the whole batch."""
    n = min(NGPU, 8)
    cmd = [sys.executable, '-m', 'torch.distributed.run', '--nnodes=1', f'--nproc-per-node={n}', '--master-addr', '127.0.0.1',
           '--master-port', str(29600 + os.getpid() % 300), os.path.join(ROOT, 'tools', 'check_dp.py'), '416', '4']
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=900)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    assert 'bitwise: True' in r.stdout


def test_sibling_fusion_is_bit_identical(weights):
    """csp_block's route / main 1x1 convs (custom_layers.py:59-60) run as one GEMM with two destinations; every output bit must
    equal the two separate convs (Y4_SIBLING=0)."""
    import y4b200
    import y4_oracle as O
    W, blob = weights
    size, batch = 160, 2
    imgs = O.synth_images(0, 0, batch, size)
    names = ['c2', 'c3', 'c9', 'c10', 'c18', 'c19', 'c39', 'c40', 'c60', 'c61', 'cat1', 'cat5']
    res = {}
    for tag, env in (('fused', None), ('separate', '0')):
        os.environ.pop('Y4_SIBLING', None)
        if env:
            os.environ['Y4_SIBLING'] = env
        try:
            eng = y4b200.Engine(img_size=size, max_batch=batch, precision=y4b200.PREC_FP16)
        finally:
            os.environ.pop('Y4_SIBLING', None)
        eng.load_darknet_bytes(blob)
        heads = eng.forward_heads(imgs)
        steps = eng.steps()
        res[tag] = (heads + [eng.get_tensor(n, batch) for n in names], steps)
        eng.close()
    fused_steps = [s for s in res['fused'][1] if s['idx'] >= 110]
    assert sorted(s['out_name'] for s in fused_steps) == ['c18+c19', 'c2+c3', 'c39+c40', 'c60+c61', 'c9+c10']
    assert len(res['fused'][1]) == len(res['separate'][1]) - 5
    assert all(s['kernel_kind'] == 1 and s['tc_epilogue'] in (32, 64) for s in fused_steps)
    for a, b in zip(res['fused'][0], res['separate'][0]):
        assert np.array_equal(a, b)


@pytest.mark.parametrize('size,batch', [(160, 2), (608, 2)])
def test_stem_fusion_is_bit_identical(weights, size, batch):
    """conv 0 + conv 1 in one kernel (stem_tc.cuh; conv 0's output never reaches HBM) == conv0_tc_kernel followed by conv 1's
    tcgen05 plan (Y4_STEM=0), bit for bit: c1, a few tensors downstream, the heads."""
    import y4b200
    import y4_oracle as O
    W, blob = weights
    imgs = O.synth_images(0, 0, batch, size)
    res = {}
    for tag, env in (('stem', None), ('separate', '0')):
        os.environ.pop('Y4_STEM', None)
        if env:
            os.environ['Y4_STEM'] = env
        try:
            eng = y4b200.Engine(img_size=size, max_batch=batch, precision=y4b200.PREC_FP16)
        finally:
            os.environ.pop('Y4_STEM', None)
        eng.load_darknet_bytes(blob)
        heads = eng.forward_heads(imgs)
        res[tag] = (heads + [eng.get_tensor(n, batch) for n in ('c1', 'r1', 'c8')], eng.steps())
        eng.close()
    assert [s['out_name'] for s in res['stem'][1]][:1] == ['c0+c1'] and res['stem'][1][0]['kernel_kind'] == 6
    assert len(res['stem'][1]) == len(res['separate'][1]) - 1
    c1a, c1b = res['stem'][0][3], res['separate'][0][3]
    assert c1a.shape == c1b.shape and np.abs(c1b).max() > 0
    assert np.array_equal(c1a, c1b), float(np.abs(c1a - c1b).max())
    for a, b in zip(res['stem'][0], res['separate'][0]):
        assert np.array_equal(a, b)


@pytest.mark.parametrize('size,batch', [(160, 2), (608, 1)])
def test_chain_fusion_is_bit_identical(weights, size, batch):
    """A 1x1 conv that reads exactly what the previous launch wrote (residual_block's first conv on r_k, custom_layers.py:38;
    the first residual after csp_block's main conv) runs inside that launch, on the output tile in shared memory (CTA-pair
    kernel, second MMA into the drained TMEM stage).  Every tensor must equal the separate launches (Y4_CHAIN=0) bit for bit."""
    import y4b200
    import y4_oracle as O
    W, blob = weights
    imgs = O.synth_images(0, 0, batch, size)
    names = ['c4', 'r1', 'c11', 'r2', 'c13', 'r3', 'c15', 'c20', 'r4', 'c22', 'r11', 'c36', 'cat3', 'r12', 'c43', 'c58']
    res = {}
    for tag, env in (('chain', '1'), ('separate', None)):          # opt-in: measured slower than the separate launches (y4_engine.cu)
        os.environ.pop('Y4_CHAIN', None)
        if env:
            os.environ['Y4_CHAIN'] = env
        try:
            eng = y4b200.Engine(img_size=size, max_batch=batch, precision=y4b200.PREC_FP16)
        finally:
            os.environ.pop('Y4_CHAIN', None)
        eng.load_darknet_bytes(blob)
        heads = eng.forward_heads(imgs)
        res[tag] = (heads + [eng.get_tensor(n, batch) for n in names], eng.steps())
        eng.close()
    chained = [s['out_name'] for s in res['chain'][1] if '>' in s['out_name']]
    assert len(chained) >= 10, chained
    assert len(res['chain'][1]) == len(res['separate'][1]) - len(chained)
    for n, a, b in zip(['hs', 'hm', 'hl'] + names, res['chain'][0], res['separate'][0]):
        assert np.array_equal(a, b), (n, float(np.abs(a - b).max()))
