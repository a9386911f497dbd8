"""Hot spots of one kernel in an ncu report (SASS level): python tools/ncu_hot.py report.ncu-rep kernel_regex [top]"""
import csv, io, subprocess, sys, collections, re
rep, pat = sys.argv[1], sys.argv[2]
top = int(sys.argv[3]) if len(sys.argv) > 3 else 25
out = subprocess.run(['ncu', '-i', rep, '--page', 'source', '--csv', '--kernel-name', 'regex:' + pat], capture_output=True, text=True).stdout
lines = out.splitlines()
# first kernel only
start = next(i for i, l in enumerate(lines) if l.startswith('"Address"'))
end = next((i for i in range(start + 1, len(lines)) if lines[i].startswith('"Kernel Name"')), len(lines))
rows = list(csv.DictReader(io.StringIO('\n'.join(lines[start:end]))))
tot = sum(int(r['# Samples']) for r in rows)
inst = sum(int(r['Instructions Executed']) for r in rows)
print(f'{len(rows)} SASS lines, {tot} samples, {inst} warp instructions executed')
cls = collections.Counter(); clsi = collections.Counter()
for r in rows:
  op = r['Source'].split()[0] if not r['Source'].strip().startswith('@') else r['Source'].split()[1]
  op = op.split('.')[0]
  cls[op] += int(r['# Samples']); clsi[op] += int(r['Instructions Executed'])
print('by opcode (samples %, instr %):')
for op, n in cls.most_common(14):
  print(f'  {op:10s} {100*n/max(tot,1):5.1f}%  {100*clsi[op]/max(inst,1):5.1f}%')
print('top lines:')
for r in sorted(rows, key=lambda r: -int(r['# Samples']))[:top]:
  print(f"  {int(r['# Samples']):6d} {100*int(r['# Samples'])/max(tot,1):5.1f}%  exec {int(r['Instructions Executed']):8d}  {r['Source'].strip()[:90]}")
