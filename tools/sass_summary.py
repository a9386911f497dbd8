"""Per-kernel SASS evidence of the Blackwell-native instructions (profiles/r02_sass_summary.md):
  python tools/sass_summary.py > profiles/r02_sass_summary.md
Counts, per kernel of libngb200.so, the mnemonics that prove tcgen05 / TMEM / TMA (UTCHMMA = tcgen05.mma kind::tf32, LDTM =
tcgen05.ld, UTMALDG / UTMASTG / UTMAREDG = TMA load / store / reduce-store, UTCBAR = tcgen05.commit, SYNCS = mbarrier) and
the legacy tensor-core (HMMA = mma.sync) and FFMA instructions."""
import collections, os, re, subprocess, sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
lib = os.path.join(ROOT, 'neograd_b200', 'lib', 'libngb200.so')
out = subprocess.run(['cuobjdump', '-sass', lib], capture_output=True, text=True).stdout
KEYS = ['UTCHMMA', 'UTCBAR', 'LDTM', 'UTMALDG', 'UTMASTG', 'UTMAREDG', 'SYNCS', 'HMMA', 'FFMA', 'LDS', 'STS', 'LDG', 'STG', 'SHFL', 'ATOM', 'RED']
cur, counts = None, collections.OrderedDict()
for line in out.splitlines():
  m = re.search(r'Function : (\S+)', line)
  if m:
    cur = m.group(1)
    counts[cur] = collections.Counter()
    continue
  if cur is None:
    continue
  m = re.match(r'\s+/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)', line)
  if m:
    op = m.group(1)
    counts[cur]['_total'] += 1
    for k in KEYS:
      if op.startswith(k):
        counts[cur][k] += 1
        break
dem = subprocess.run(['cu++filt'] + list(counts), capture_output=True, text=True).stdout.splitlines()
print('# r02 -- SASS summary of libngb200.so (cuobjdump -sass, sm_100a)\n')
print('`UTCHMMA` = tcgen05.mma (kind::tf32), `UTCBAR` = tcgen05.commit, `LDTM` = tcgen05.ld, `UTMALDG` / `UTMASTG` / `UTMAREDG` = TMA '
      'load / store / reduce-add store, `SYNCS` = mbarrier ops, `HMMA` = legacy mma.sync.  Kernels without any of the first seven are '
      'CUDA-core kernels (bandwidth / latency bound by design, DESIGN.md section 4).\n')
print('| kernel | instr | ' + ' | '.join(KEYS) + ' |\n|---|---|' + '---|' * len(KEYS))
rows = []
for (name, c), d in zip(counts.items(), dem):
  short = re.sub(r'\((?!anonymous).*$', '', re.sub(r'\((?:int|bool|unsigned int)\)', '', d)).replace('ngb::', '').replace('(anonymous namespace)::', '').replace('<unnamed>::', '')
  rows.append((short, c))
rows.sort(key=lambda r: (-r[1]['UTCHMMA'], -r[1]['HMMA'], r[0]))
for short, c in rows:
  print(f'| `{short[:90]}` | {c["_total"]} | ' + ' | '.join(str(c[k]) if c[k] else '' for k in KEYS) + ' |')
