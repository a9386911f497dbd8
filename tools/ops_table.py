"""jsonl records of tools/bench_ops.py -> the markdown table kept under profiles/."""
import json
import sys


def main(path):
  recs = [json.loads(l) for l in open(path) if l.strip()]
  pk = next((r['peaks'] for r in recs if 'peaks' in r), None)
  cub = next((r for r in recs if r.get('op', '').startswith('cublas')), None)
  if pk:
    print(f"Denominators: HBM {pk['hbm_gbs']:.1f} GB/s ({pk['source']}, MEASURED_PEAKS.json); TF32 {pk['tf32_tflops']:.1f} TFLOP/s = half the measured "
          f"BF16 {pk['bf16_tflops']:.0f} TFLOP/s.", end='')
  if cub:
    print(f"  cuBLAS TF32 8192^3 matmul timed in the same run: {cub['tflops']:.0f} TFLOP/s (best {cub.get('tflops_best', 0):.0f}).", end='')
  print('\n')
  print('| op | shape | mean us | TFLOP/s | % of TF32 peak | algorithmic GB/s | % of HBM peak |')
  print('|---|---|---|---|---|---|---|')
  for r in recs:
    if 'shape' not in r:
      continue
    if 'error' in r:
      print(f"| {r['op']} | {r['shape']} | error: {r['error']} | | | | |")
      continue
    gbs = r['bytes'] / r['us'] / 1e3
    hb = 100 * gbs / pk['hbm_gbs'] if pk else 0
    if 'tflops' in r:
      print(f"| {r['op']} | {r['shape']} | {r['us']:.1f} | {r['tflops']:.1f} | {100 * r['frac_tf32_peak']:.1f} | {gbs:.0f} | {hb:.1f} |")
    else:
      print(f"| {r['op']} | {r['shape']} | {r['us']:.1f} |  |  | {gbs:.0f} | {hb:.1f} |")


if __name__ == '__main__':
  main(sys.argv[1])
