"""Warp-stall breakdown of the kernels in an ncu report: python tools/ncu_stalls.py report.ncu-rep [kernel_regex]"""
import csv, io, re, subprocess, sys
rep = sys.argv[1]
pat = sys.argv[2] if len(sys.argv) > 2 else '.'
out = subprocess.run(['ncu', '-i', rep, '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
hdr = rows[0]
ki = hdr.index('Kernel Name')
for r in rows[2:]:
  if not re.search(pat, r[ki]):
    continue
  print(r[ki][:70])
  vals = []
  for i, h in enumerate(hdr):
    if 'issue_stalled' in h and h.endswith('_per_issue_active.ratio') and 'not_issued' not in h:
      try:
        vals.append((float(r[i]), h))
      except ValueError:
        pass
  for v, h in sorted(vals, reverse=True)[:10]:
    print(f'   {v:6.2f} warps/issue  {h.replace("smsp__average_warps_issue_stalled_", "").replace("_per_issue_active.ratio", "")}')
