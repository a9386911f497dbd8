"""Times the pooled conv-gradient entry points (ngb_conv_{w,d,b}grad_pooled) on the LeNet-B layer shapes, each kernel alone,
L2 flushed between launches.  Inputs come from a real forward pass (conv -> ReLU+pool) so masks and indices are realistic.

  python tools/bench_pooled.py [N]"""
import ctypes
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from neograd_b200 import _backend as B  # noqa: E402
from tools.bench_ops import timeit, P, S  # noqa: E402


def main():
  N = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
  B.ensure_runtime()
  lib = B.lib
  for (C, H, O, K) in ((1, 28, 6, 5), (6, 12, 16, 5)):
    Ho = H - K + 1
    Hp = Ho // 2
    x = torch.randn(N, C, H, H, device='cuda')
    w = torch.randn(O, C, K, K, device='cuda') * 0.2
    b = torch.randn(O, device='cuda') * 0.1
    z = torch.empty(N, O, Ho, Ho, device='cuda')
    B.check(lib.ngb_conv_fprop(P(x), P(w), P(b), P(z), N, C, H, H, O, K, K, 0, 1, S()))
    y = torch.empty(N, O, Hp, Hp, device='cuda')
    idx = torch.empty(N, O, Hp, Hp, device='cuda', dtype=torch.uint8)
    B.check(lib.ngb_maxpool_relu_fwd(P(z), P(y), P(idx), N * O, Ho, Ho, 2, 2, 0, 2, S()))
    ug = torch.randn(N, O, Hp, Hp, device='cuda')
    dw, dx, db = torch.empty_like(w), torch.empty_like(x), torch.empty_like(b)
    calls = {
        'wgrad_pooled': lambda: B.check(lib.ngb_conv_wgrad_pooled(P(ug), P(y), P(idx), P(z), P(x), P(dw), N, C, H, H, O, K, K, 0, 1, 0, S())),
        'dgrad_pooled': lambda: B.check(lib.ngb_conv_dgrad_pooled(P(ug), P(y), P(idx), P(z), P(w), P(dx), N, C, H, H, O, K, K, 0, 1, 0, S())),
        'wbgrad_pooled': lambda: B.check(lib.ngb_conv_wbgrad_pooled(P(ug), P(y), P(idx), P(z), P(x), P(dw), P(db), N, C, H, H, O, K, K, 0, 1, 0, 0, S())),
        'bgrad_pooled_idx': lambda: B.check(lib.ngb_conv_bgrad_pooled_idx(P(ug), P(idx), P(db), N, O, Hp * Hp, 0, S())),
        'bgrad_pooled': lambda: B.check(lib.ngb_conv_bgrad_pooled(P(ug), P(y), P(z), P(db), N, O, Ho, Ho, 0, S())),
    }
    # the register-prefetched wgrad against the plain staging path (different CTA partition -> fp32 re-association only)
    if os.environ.get('BENCH_POOLED_SKIP_CMP'):
      for name, fn in calls.items():
        if not (name == 'dgrad_pooled' and C == 1):
          mean, best = timeit(fn, reps=5)
          print(name, C, round(mean * 1e6, 2), flush=True)
      continue
    calls['wgrad_pooled']()
    d1 = dw.clone()
    os.environ['NGB_WGRAD_NO_PF'] = '1'
    calls['wgrad_pooled']()
    t_nopf = timeit(calls['wgrad_pooled'], reps=20)
    del os.environ['NGB_WGRAD_NO_PF']
    torch.cuda.synchronize()
    err = float((d1 - dw).abs().max() / dw.abs().max())
    print(json.dumps({'op': 'wgrad_pooled', 'shape': [N, C, H, H, O, K], 'pf_vs_plain_rel_err': err,
                      'plain_us': round(t_nopf[0] * 1e6, 2)}), flush=True)
    assert err < 1e-5, err
    if lib.ngb_conv_wbgrad_pooled_is_fused(N, C, H, H, O, K, K, 0, 1) == 1:
      calls['wbgrad_pooled']()
      ds, bs = dw.clone(), db.clone()
      os.environ['NGB_NO_SPARSE_WGRAD'] = '1'
      calls['wbgrad_pooled']()
      t_mma = timeit(calls['wbgrad_pooled'], reps=20)
      del os.environ['NGB_NO_SPARSE_WGRAD']
      torch.cuda.synchronize()
      print(json.dumps({'op': 'wbgrad_pooled', 'shape': [N, C, H, H, O, K], 'sparse_vs_mma_rel_err': float((ds - dw).abs().max() / dw.abs().max()),
                        'db_rel_err': float((bs - db).abs().max() / db.abs().max()), 'mma_plus_bgrad_us': round(t_mma[0] * 1e6, 2)}), flush=True)
    for name, fn in calls.items():
      if name == 'dgrad_pooled' and C == 1:
        continue
      mean, best = timeit(fn, reps=20)
      print(json.dumps({'op': name, 'shape': [N, C, H, H, O, K], 'us': round(mean * 1e6, 2), 'us_best': round(best * 1e6, 2),
                        'env': {k: v for k, v in os.environ.items() if k.startswith('NGB_')}}), flush=True)


if __name__ == '__main__':
  main()
