"""Compare per-layer times of several profile_layers runs.  usage: python tools/cmp_layers.py tag0 tag1 ... [--sel c2,c3]"""
import json, sys
args = [a for a in sys.argv[1:] if not a.startswith('--')]
sel = None
for a in sys.argv[1:]:
    if a.startswith('--sel='):
        sel = a[6:].split(',')
D = {t: json.load(open(f'gpurun_out/layers_{t}.json')) for t in args}
names = [r['name'] for r in D[args[0]]['layers']]
if sel is None:
    sel = ['c1','c2','c4','c5','c7','c8','c9','c11','c12','c20','c21','c22','c40','c41','c42','c60','c63','c64','c72','c73','c80','c81','c88','c92','c94','c100']
print('layer   shape                ', ' '.join(f'{t:>7s}' for t in args))
for n in sel:
    i = names.index(n)
    r = D[args[0]]['layers'][i]
    if 'hw' not in r:
        continue
    print(f"{n:5s} {r['hw']:3d} k{r['k']} s{r['s']} {r['cin']:4d}->{r['cout']:4d} ", ' '.join(f"{D[t]['layers'][i]['ms']*1000:7.1f}" for t in args),
          f" {args[0]}:bn{r['bn']} m{r['mode']} e{r['epi']} S{r['st']} g{r['grp']} c{r['cps']} w{r.get('nepi',4)} r{r.get('bres',0)}")
