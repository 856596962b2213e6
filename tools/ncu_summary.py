"""Summarise an .ncu-rep (raw page) into one line per kernel launch.  usage: python tools/ncu_summary.py rep [out.md]"""
import csv, subprocess, sys
rep = sys.argv[1]
raw = subprocess.run(['ncu', '-i', rep, '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
hdr, units = rows[0], rows[1]
col = {h: i for i, h in enumerate(hdr)}
def g(r, name, default=''):
    i = col.get(name); return r[i] if i is not None else default
want = [('gpu__time_duration.sum', 'us'), ('launch__grid_size', 'grid'),
        ('sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active', 'tensor%'),
        ('dram__bytes_read.sum', 'dramR'), ('dram__bytes_write.sum', 'dramW'),
        ('gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed', 'dram%'),
        ('lts__t_sector_hit_rate.pct', 'L2hit%'), ('lts__t_bytes.sum', 'L2bytes'),
        ('l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed', 'smemLSU%'),
        ('sm__warps_active.avg.pct_of_peak_sustained_active', 'occ%'),
        ('launch__occupancy_limit_shared_mem', 'ctaSmem'), ('launch__registers_per_thread', 'regs'),
        ('sm__throughput.avg.pct_of_peak_sustained_elapsed', 'sm%'),
        ('lts__throughput.avg.pct_of_peak_sustained_elapsed', 'lts%')]
lines = ['| # | kernel | ' + ' | '.join(n for _, n in want) + ' |', '|' + '---|' * (len(want) + 2)]
for k, r in enumerate(rows[2:]):
    vals = []
    for m, n in want:
        v = g(r, m)
        u = units[col[m]] if m in col else ''
        try:
            v = f'{float(v):.1f}' + (u if u in ('Mbyte', 'Gbyte', 'Kbyte', 'byte') else '')
        except ValueError:
            pass
        vals.append(v)
    lines.append(f'| {k} | {g(r, "Kernel Name")[:34]} | ' + ' | '.join(vals) + ' |')
out = '\n'.join(lines)
print(out)
if len(sys.argv) > 2:
    open(sys.argv[2], 'w').write(out + '\n')
