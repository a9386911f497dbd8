"""Prints the per-launch times of the LAST training step in an ncu launch list (gpu__time_duration.sum CSV)."""
import csv
import re
import sys


def main(path, per_step):
  with open(path) as f:
    lines = [l for l in f if not l.startswith('==')]
  rows = []
  for row in csv.DictReader(lines):
    v = float(row['Metric Value'].replace(',', ''))
    unit = row['Metric Unit']
    v = v / 1000 if unit in ('ns', 'nsecond') else v * 1000 if unit in ('ms', 'msecond') else v
    rows.append((re.sub(r'\(.*', '', row['Kernel Name']).replace('void ', '').replace('ngb::', ''), v))
  step = rows[-per_step:]
  for n, v in step:
    print(f'{v:8.1f}  {n[:90]}')
  print(f'{sum(v for _, v in step):8.1f}  total of {len(step)} launches')


if __name__ == '__main__':
  main(sys.argv[1], int(sys.argv[2]))
