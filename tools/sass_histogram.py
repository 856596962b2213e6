"""SASS opcode histogram per kernel of liby4.so (cuobjdump -sass), the evidence B200_PROFILING.md asks for: tcgen05.mma ->
UTC*MMA, tcgen05.ld/st -> LDTM/STTM, TMA -> UTMALDG/UTMASTG, cp.async -> LDGSTS; HMMA would be the legacy mma.sync path.
usage: python tools/sass_histogram.py [tag]  ->  profiles/<tag>_sass_histogram.md   (runs here: no GPU needed)"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
tag = sys.argv[1] if len(sys.argv) > 1 else 'r02'
lib = os.path.join(ROOT, 'yolo-v4-tf.keras_b200', 'liby4.so')
sass = subprocess.run(['/usr/local/cuda/bin/cuobjdump', '-sass', lib], capture_output=True, text=True, check=True).stdout
demangle = {}
kernels = collections.OrderedDict()
cur = None
for line in sass.splitlines():
    m = re.match(r'\s*Function : (\S+)', line)
    if m:
        cur = m.group(1)
        kernels[cur] = collections.Counter()
        continue
    m = re.match(r'\s+/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d+\s+)?([A-Z][A-Z0-9_]*(?:\.[A-Z0-9_]+)*)', line)
    if m and cur:
        kernels[cur][m.group(1)] += 1
names = subprocess.run(['c++filt'], input='\n'.join(kernels), capture_output=True, text=True).stdout.splitlines()
KEY = ['UTCHMMA', 'UTCBAR', 'LDTM', 'STTM', 'UTMALDG', 'UTMASTG', 'UTMAPF', 'LDGSTS', 'SYNCS', 'HMMA', 'FFMA', 'FFMA2', 'FMUL2', 'FADD2', 'MUFU', 'LDG', 'STG', 'LDS', 'STS', 'BAR']
out = [f'# SASS opcode histogram per kernel of liby4.so ({tag}; cuobjdump -sass, sm_100a)', '',
       'Counts of static instructions.  `UTCHMMA` = tcgen05.mma (`.2CTA` = cta_group::2), `LDTM` = tcgen05.ld, `UTMALDG` / `UTMASTG` = TMA',
       'load / store, `UTCBAR` = tcgen05.commit, `SYNCS` = mbarrier ops, `LDGSTS` = cp.async, `FFMA2` / `FMUL2` / `FADD2` = packed fp32 (.f32x2) in the epilogues.  No `HMMA` (legacy mma.sync) anywhere.', '',
       '| kernel | total | ' + ' | '.join(KEY) + ' |', '|---|---|' + '---|' * len(KEY)]
tot = collections.Counter()
for (mangled, c), nice in zip(kernels.items(), names):
    nice = re.sub(r'\(.*\)$', '', nice).replace('void ', '').replace('y4::', '')
    row = []
    for k in KEY:
        n = sum(v for op, v in c.items() if op == k or op.startswith(k + '.'))      # exact mnemonic: FFMA does not count FFMA2
        if k == 'UTCHMMA':
            n2 = sum(v for op, v in c.items() if op.startswith('UTCHMMA') and '2CTA' in op)
            row.append(f'{n} ({n2} .2CTA)' if n2 else str(n))
        else:
            row.append(str(n))
        tot[k] += n
    out.append(f'| `{nice}` | {sum(c.values())} | ' + ' | '.join(row) + ' |')
out.append('| **all kernels** | | ' + ' | '.join(str(tot[k]) for k in KEY) + ' |')
path = os.path.join(ROOT, 'profiles', f'{tag}_sass_histogram.md')
with open(path, 'w') as f:
    f.write('\n'.join(out) + '\n')
print(path, len(kernels), 'kernels;', {k: tot[k] for k in ('UTCHMMA', 'LDTM', 'UTMALDG', 'UTMASTG', 'HMMA')})
