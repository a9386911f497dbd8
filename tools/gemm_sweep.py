"""Tile-width / K-split sweep of the tcgen05 GEMM at the GAN shapes (NGB_GEMM_BN / NGB_GEMM_SPLITS overrides):
  python tools/gemm_sweep.py > gpurun_out/gemm_sweep.txt
Prints mean time of each (pass, shape, BN, splits) with an L2 flush between launches; 'auto' = the library's own choice."""
import ctypes
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from neograd_b200 import _backend as B  # noqa: E402

P = lambda t: ctypes.c_void_p(t.data_ptr())
S = lambda: ctypes.c_void_p(B.stream())
B.ensure_runtime()
flush = torch.empty(256 << 20, device='cuda', dtype=torch.uint8)


def timeit(fn, reps=6):
  for _ in range(2):
    if fn() != 0:
      return None
  torch.cuda.synchronize()
  ts = []
  for _ in range(reps):
    flush.zero_()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    fn()
    e1.record()
    torch.cuda.synchronize()
    ts.append(e0.elapsed_time(e1) * 1e3)
  return float(np.mean(ts))


rnd = lambda *s: torch.randn(*s, device='cuda', dtype=torch.float32)
for M, K, N in ((8192, 784, 512), (8192, 512, 256), (8192, 256, 784), (4096, 256, 120), (4096, 120, 84)):
  A, Bm, U = rnd(M, K), rnd(K, N), rnd(M, N)
  C, dA, dB, db = torch.empty(M, N, device='cuda'), torch.empty(M, K, device='cuda'), torch.empty(K, N, device='cuda'), torch.empty(N, device='cuda')
  calls = {
    'fwd': lambda: B.lib.ngb_gemm(0, 0, 0, M, N, K, P(A), K, P(Bm), N, P(C), N, 0, S()),
    'dgrad': lambda: B.lib.ngb_gemm(0, 0, 1, M, K, N, P(U), N, P(Bm), N, P(dA), K, 0, S()),
    'wgrad': lambda: B.lib.ngb_gemm(0, 1, 0, K, N, M, P(A), K, P(U), N, P(dB), N, 0, S()),
    'wgrad+db': lambda: B.lib.ngb_linear_wgrad(P(A), P(U), P(dB), P(db), M, K, N, 0, 0, S()),
    'linear_fwd': lambda: B.lib.ngb_linear_fwd(P(A), P(Bm), P(db), P(C), P(U), 2, 0.0, M, K, N, S()),
  }
  for name, call in calls.items():
    os.environ.pop('NGB_GEMM_BN', None)
    os.environ.pop('NGB_GEMM_SPLITS', None)
    res = [('auto', timeit(call))]
    for bn in (64, 128, 256):
      for sp in ((1,) if name in ('fwd', 'dgrad', 'linear_fwd') and K <= 512 and name != 'dgrad' else (1, 2, 3, 4, 6, 8, 12, 16)):
        if name in ('fwd', 'linear_fwd') and sp > 1 and K < 512:
          continue
        os.environ['NGB_GEMM_BN'] = str(bn)
        os.environ['NGB_GEMM_SPLITS'] = str(sp)
        t = timeit(call, 4)
        if t is not None:
          res.append((f'bn{bn}/s{sp}', t))
    os.environ.pop('NGB_GEMM_BN', None)
    os.environ.pop('NGB_GEMM_SPLITS', None)
    best = min((r for r in res if r[1] is not None), key=lambda r: r[1])
    print(f'{name:10s} [{M},{K},{N}] auto {res[0][1]:.1f} us  best {best[0]} {best[1]:.1f} us  |  ' +
          '  '.join(f'{k}={v:.1f}' for k, v in res[1:] if v is not None), flush=True)
  del A, Bm, U, C, dA, dB
