"""Target of tools/sanitize.sh: one small forward (64x64, batch 2) + decode/NMS (fast path, long class segment, overflow path)
through the C-ABI, for compute-sanitizer.  usage: python tools/sanitize_target.py fp16|fp16x3|fp32|decode
(decode: no forward, only the decode / NMS kernels on crafted heads at 128x128: every regime of nms_image_kernel -- class
segments <= 64 (warp path), 65..704 (CTA-wide bit matrix), longer (sequential) -- the staged filter, and the overflow path)"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, 'oracle')]
import numpy as np  # noqa: E402
import y4b200  # noqa: E402
import y4_oracle as O  # noqa: E402

if sys.argv[1] == 'decode':
    S, B = 128, 3
    eng = y4b200.Engine(img_size=S, max_batch=B, precision=y4b200.PREC_FP16)
    heads = O.synth_heads(seed=2, batch=B, img_size=S, n_clusters=60, num_classes=80)
    r1 = eng.decode_nms(heads, with_indices=True)
    ref = O.decode_nms(heads, S)
    assert all(np.array_equal(r1[i], ref[i]) for i in (2, 3, 4)), 'decode/NMS differs from the oracle under the sanitizer'
    res = [r1[3].tolist()]
    for n_hot in (50, 150, 720):                                   # one class: warp path / bit matrix / sequential scan
        one = [h.copy() for h in heads]
        v = one[0].reshape(B, 16, 16, 3, 85)
        idx = np.random.default_rng(n_hot).choice(16 * 16 * 3, n_hot, replace=False)
        r, c, a = np.unravel_index(idx, (16, 16, 3))
        v[:, r, c, a, 4] = 4.0
        v[:, r, c, a, 5 + 3] = 4.0 + np.random.default_rng(1).uniform(0, 1, n_hot).astype(np.float32)
        v[:, r, c, a, 2:4] += 1.0
        got = eng.decode_nms(one, with_indices=True)
        ref = O.decode_nms(one, S)
        assert all(np.array_equal(got[i], ref[i]) for i in (2, 3, 4)), f'one-class segment of {n_hot} differs from the oracle'
        res.append(got[3].tolist())
    hot = [np.full_like(h, 3.0) for h in heads]                    # every (box, class) a candidate -> overflow path
    res.append(eng.decode_nms(hot)[3].tolist())
    res.append(eng.decode_nms(heads, 0.413, 0.999)[3].tolist())    # nothing passes
    print('SANITIZE TARGET OK decode valid', res, flush=True)
    eng.close()
    sys.exit(0)
prec = {'fp16': y4b200.PREC_FP16, 'fp16x3': y4b200.PREC_FP16X3, 'fp32': y4b200.PREC_FP32}[sys.argv[1]]
S, B = 64, 2
W = O.synth_weights(seed=1, calib_size=64)
eng = y4b200.Engine(img_size=S, max_batch=B, precision=prec)
eng.load_darknet_bytes(W.to_darknet_bytes())
out = eng.predict(O.synth_images(0, 0, B, S), with_indices=True)
raw = [(O.synth_images(1, i, 1, 40)[0] * 255).astype(np.uint8) for i in range(B)]
out8 = eng.predict_u8(raw)
eng.submit_u8(raw); eng.submit_u8(raw); eng.collect(); eng.collect()
heads = O.synth_heads(seed=2, batch=B, img_size=S, n_clusters=10)
r1 = eng.decode_nms(heads)
hot = [np.full_like(h, 3.0) for h in heads]                     # every (box, class) a candidate: 20,160 per image -> overflow path
r2 = eng.decode_nms(hot)
one = [h.copy() for h in heads]
one[0].reshape(B, 8, 8, 3, 85)[..., 4] = 4.0                     # 192 candidates of one class ...
one[0].reshape(B, 8, 8, 3, 85)[..., 5 + 3] = 4.0
r3 = eng.decode_nms(one, 0.413, 0.05)
kinds = sorted({(l['kernel_kind'], l['tc_mode'], l['tile_n'], l['tc_epilogue'], l['tc_epi_warps']) for l in eng.layers()})
print('SANITIZE TARGET OK', sys.argv[1], 'valid', out[3].tolist(), out8[3].tolist(), r1[3].tolist(), r2[3].tolist(), r3[3].tolist(), 'plans', kinds, flush=True)
eng.close()
