"""Turns ncu CSV exports into the short markdown summaries kept under profiles/.
  python tools/ncu_summary.py launches gpurun_out/launches.csv            -> per-kernel time shares
  python tools/ncu_summary.py full gpurun_out/prof.ncu-rep [regex]        -> key metrics per profiled launch
"""
import collections
import csv
import io
import re
import subprocess
import sys


def launches(path):
  with open(path) as f:
    lines = [l for l in f if not l.startswith('==')]
  agg = collections.defaultdict(lambda: [0, 0.0])
  for row in csv.DictReader(lines):
    v = float(row['Metric Value'].replace(',', ''))
    unit = row['Metric Unit']
    v = v / 1000 if unit in ('ns', 'nsecond') else v * 1000 if unit in ('ms', 'msecond') else v
    name = re.sub(r'\(.*', '', row['Kernel Name'])
    agg[name][0] += 1
    agg[name][1] += v
  tot = sum(v[1] for v in agg.values())
  print(f'| kernel | launches | total us | share | mean us |\n|---|---|---|---|---|')
  for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print(f'| `{k[:80]}` | {v[0]} | {v[1]:.1f} | {100 * v[1] / tot:.1f}% | {v[1] / v[0]:.2f} |')
  print(f'\ntotal {tot:.1f} us over {sum(v[0] for v in agg.values())} launches (cold-cache, serialised: compare shares)')


KEYS = ['gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum',
        'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed', 'sm__throughput.avg.pct_of_peak_sustained_elapsed',
        'sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active', 'sm__inst_executed_pipe_tensor.sum',
        'sm__warps_active.avg.pct_of_peak_sustained_active', 'launch__registers_per_thread', 'launch__grid_size',
        'launch__block_size', 'launch__shared_mem_per_block_dynamic', 'l1tex__data_bank_conflicts_pipe_lsu.sum',
        'lts__t_sector_hit_rate.pct', 'smsp__cycles_active.avg', 'sm__cycles_elapsed.max']


def full(path, pattern=None):
  out = subprocess.run(['ncu', '-i', path, '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
  rows = list(csv.reader(io.StringIO(out)))
  hdr, units = rows[0], rows[1]
  for row in rows[2:]:
    name = row[hdr.index('Kernel Name')]
    if pattern and not re.search(pattern, name):
      continue
    print(f'### `{name[:100]}`\n')
    print('| metric | value | unit |\n|---|---|---|')
    for k in KEYS:
      if k in hdr:
        i = hdr.index(k)
        print(f'| {k} | {row[i]} | {units[i]} |')
    print()


if __name__ == '__main__':
  {'launches': launches, 'full': full}[sys.argv[1]](*sys.argv[2:])
