"""ncu CSV (one row per launch x metric, e.g. --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum,
sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active) -> profiles/traffic.json with per-kernel means per launch.
usage: python tools/ncu_traffic.py gpurun_out/metrics.csv [profiles/traffic.json]"""
import collections
import csv
import json
import re
import sys

src = sys.argv[1]
dst = sys.argv[2] if len(sys.argv) > 2 else 'profiles/traffic.json'
rows = [r for r in csv.reader(open(src)) if len(r) > 10]
hdr = rows[0]
ki, mi, ui, vi, idi = (hdr.index(c) for c in ('Kernel Name', 'Metric Name', 'Metric Unit', 'Metric Value', 'ID'))
scale = {'byte': 1, 'Kbyte': 1e3, 'Mbyte': 1e6, 'Gbyte': 1e9, 'ns': 1e-3, 'us': 1, 'usecond': 1, 'ms': 1e3, 'msecond': 1e3, 'nsecond': 1e-3, '%': 1}


def label(name):
    name = re.sub(r'^void ', '', name)
    name = re.sub(r'\(.*$', '', name)
    m = re.match(r'conv_tc_kernel<(\d+), (\d+), (\d+), (\d+)(?:, (\d+))?>', name)     # default template argument LEAN = 0 is not printed
    if m:
        name = f'conv_tc_kernel<{m.group(1)}, {m.group(2)}, {m.group(3)}, {m.group(4)}, {m.group(5) or 0}>'
    return name


per = collections.defaultdict(lambda: collections.defaultdict(dict))     # kernel -> launch id -> metric -> value
for r in rows[1:]:
    try:
        v = float(r[vi].replace(',', '')) * scale.get(r[ui], 1)
    except ValueError:
        continue
    per[label(r[ki])][r[idi]][r[mi]] = v
out = {'source': src, 'note': 'means per launch over one step; dram bytes in bytes, duration in us (cold-cache, serialised under ncu)', 'kernels': {}}
for k, launches in per.items():
    n = len(launches)
    rd = sum(l.get('dram__bytes_read.sum', 0.0) for l in launches.values()) / n
    wr = sum(l.get('dram__bytes_write.sum', 0.0) for l in launches.values()) / n
    us = sum(l.get('gpu__time_duration.sum', 0.0) for l in launches.values()) / n
    tens = [l[m] for l in launches.values() for m in l if m.startswith('sm__pipe_tensor_cycles_active')]
    out['kernels'][k] = {'launches': n, 'dram_read_bytes_per_launch': rd, 'dram_write_bytes_per_launch': wr,
                         'dram_bytes_per_launch': rd + wr, 'duration_us_per_launch': us,
                         'tensor_pipe_active_pct': sum(tens) / len(tens) if tens else None}
json.dump(out, open(dst, 'w'), indent=1)
for k, v in sorted(out['kernels'].items(), key=lambda kv: -kv[1]['duration_us_per_launch'] * kv[1]['launches']):
    print(f"{k:44s} n={v['launches']:3d}  {v['duration_us_per_launch']:8.1f} us  dram {v['dram_bytes_per_launch']/1e6:8.1f} MB/launch  tensor {v['tensor_pipe_active_pct']}")
