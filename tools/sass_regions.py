"""Groups the SASS view of an ncu report (ncu -i X --page source --csv) into runs of instructions with the same
execution count: shows which loop of a kernel the instructions (and the stall samples) go to."""
import collections
import csv
import sys


def main(path, thresh=0.005):
  rows = list(csv.reader(open(path)))
  h = next(i for i, r in enumerate(rows) if 'Instructions Executed' in r)
  hdr = rows[h]
  ie, isrc, ist = hdr.index('Instructions Executed'), hdr.index('Source'), hdr.index('# Samples')
  data = [r for r in rows[h + 1:] if len(r) > ie and r[ie].isdigit()]
  tot = sum(int(r[ie]) for r in data)
  stot = sum(int(r[ist]) for r in data)
  print('total warp instructions', tot, 'stall samples', stot)
  prev, start, acc, samp, ops = None, 0, 0, 0, collections.Counter()

  def flush(i):
    if acc > tot * thresh or samp > stot * thresh:
      print(f'[{start:5d}-{i:5d}] n={i - start:4d} execs={prev:9d} inst={100 * acc / tot:5.1f}% samples={100 * samp / max(stot, 1):5.1f}%',
            dict(ops.most_common(7)))

  for i, r in enumerate(data):
    c = int(r[ie])
    if prev is None or abs(c - prev) > 0.02 * max(prev, 1):
      if prev is not None:
        flush(i)
      start, acc, samp, ops, prev = i, 0, 0, collections.Counter(), c
    acc += c
    samp += int(r[ist])
    toks = r[isrc].split()
    op = toks[1] if toks[0].startswith('@') else toks[0]
    ops[op.split('.')[0]] += 1
  flush(len(data))


if __name__ == '__main__':
  main(sys.argv[1], float(sys.argv[2]) if len(sys.argv) > 2 else 0.005)
