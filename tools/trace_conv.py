"""CTA timeline of tcgen05 conv layers.  usage: python tools/trace_conv.py size batch idx [idx...]"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, 'oracle')]
import numpy as np
import y4b200, y4_oracle as O
size, batch = int(sys.argv[1]), int(sys.argv[2])
eng = y4b200.Engine(img_size=size, max_batch=batch, precision=y4b200.PREC_FP16)
eng.load_darknet_bytes(O.synth_weights(seed=1).to_darknet_bytes())
eng.synth_fill(0, 0, batch); eng.run_forward_resident(batch); eng.sync()
L = eng.layers()
names = ['setup', 'tma0_issued', 'tma_all_issued', 'first_landed', 'last_mma_issued', 'acc_ready', 'epi_done', 'exit']
for idx in map(int, sys.argv[3:]):
    l = L[idx]
    t = eng.trace_conv(idx, batch)
    t = t[t[:, 0] != 0]
    rel = (t[:, 1:9] - t[:, [0]]).astype(np.float64)
    gt = t[:, 14] - t[:, 14].min()
    print(f"\nconv {idx}: {l['cin']}->{l['cout']} k{l['ksize']} s{l['stride']} hw{l['out_hw']} kind{l['kernel_kind']} bn{l['tile_n']}  ctas traced {len(t)}  span {gt.max()/1e3:.1f} us")
    print('  cycles from CTA entry (median / p90): ' + ', '.join(f'{n} {np.median(rel[:, i]):.0f}/{np.percentile(rel[:, i], 90):.0f}' for i, n in enumerate(names)))
    print(f'  MMA thread: loop {np.median(t[:, 11]):.0f} cyc, waiting on full(data) {np.median(t[:, 9]):.0f}, on patch {np.median(t[:, 13]):.0f}, on tempty(epilogue) {np.median(t[:, 10]):.0f};  producer waiting on empty(slots) {np.median(t[:, 12]):.0f}; tiles/CTA ~{l["flops"] and 0}')
    order = np.argsort(t[:, 14])
    sm = t[:, 15]
    s0 = sm[order[0]]
    mine = [i for i in order if sm[i] == s0][:8]
    print('  CTAs on SM', s0, 'start(us):', [round(float(gt[i]) / 1e3, 1) for i in mine], 'life(cyc):', [int(t[i, 8] - t[i, 0]) for i in mine])
