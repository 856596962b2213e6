"""Several .ncu-rep captures (--set full) -> one markdown table: pipes, DRAM bytes, registers, issue rate and warp-stall shares per launch.
usage: python tools/ncu_kernels_md.py out.md "title" rep1.ncu-rep [rep2 ...]"""
import csv
import io
import re
import subprocess
import sys

out_path, title, reps = sys.argv[1], sys.argv[2], sys.argv[3:]
cols = [('us', 'gpu__time_duration.sum'), ('grid', 'launch__grid_size'), ('tensor%', 'sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active'),
        ('L2%', 'lts__throughput.avg.pct_of_peak_sustained_elapsed'), ('DRAM rd', 'dram__bytes_read.sum'), ('DRAM wr', 'dram__bytes_write.sum'),
        ('regs', 'launch__registers_per_thread'), ('occ%', 'sm__warps_active.avg.pct_of_peak_sustained_active'), ('IPC/SM', 'sm__inst_executed.avg.per_cycle_active'),
        ('issue%', 'smsp__issue_active.avg.pct_of_peak_sustained_active'), ('warp inst', 'smsp__inst_executed.sum')]
stalls = ['long_scoreboard', 'wait', 'short_scoreboard', 'mio_throttle', 'barrier', 'math_pipe_throttle', 'not_selected', 'selected', 'lg_throttle',
          'dispatch_stall', 'branch_resolving', 'no_instruction', 'membar', 'sleeping']
lines = [f'# {title}', '', 'stall columns = share of warp-stall samples (smsp__pcsamp_warps_issue_stalled_*); per-launch times under ncu are cold-cache and serialised.', '',
         '| kernel (launch) | ' + ' | '.join(c for c, _ in cols) + ' | ' + ' | '.join('st:' + s for s in stalls[:9]) + ' |', '|' + '---|' * (len(cols) + 10)]
for rep in reps:
    txt = subprocess.run(['ncu', '-i', rep, '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(txt)))
    h, u = rows[0], rows[1]
    ki = h.index('Kernel Name')
    for li, r in enumerate(rows[2:]):
        name = re.sub(r'\(.*', '', re.sub(r'^void y4::|^y4::|^void ', '', r[ki].replace('(int)', '')))
        vals = []
        for _, m in cols:
            v = ''
            if m in h:
                i = h.index(m)
                v = r[i]
                try:
                    f = float(v.replace(',', ''))
                    v = ('%.1f' % f if f < 1e4 else '%.3g' % f) + (u[i] if u[i] in ('Mbyte', 'Kbyte', 'byte', 'Gbyte') else '')
                except ValueError:
                    pass
            vals.append(v)
        st, tot = {}, 0.0
        for s in stalls:
            m = 'smsp__pcsamp_warps_issue_stalled_' + s
            if m in h:
                try:
                    st[s] = float(r[h.index(m)].replace(',', ''))
                except ValueError:
                    st[s] = 0.0
                tot += st[s]
        sv = [('%d%%' % round(100 * st.get(s, 0) / tot) if tot else '') for s in stalls[:9]]
        lines.append(f'| `{name}` ({li}) | ' + ' | '.join(vals) + ' | ' + ' | '.join(sv) + ' |')
open(out_path, 'w').write('\n'.join(lines) + '\n')
print('\n'.join(lines[4:]))
