"""Which part of the fused Linear epilogue costs what: ngb_linear_fwd / _dgrad variants timed at two shapes."""
import ctypes
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from neograd_b200 import _backend as B  # noqa: E402

P = lambda t: ctypes.c_void_p(t.data_ptr()) if t is not None else None
S = lambda: ctypes.c_void_p(B.stream())
B.ensure_runtime()
flush = torch.empty(256 << 20, device='cuda', dtype=torch.uint8)
only = sys.argv[1] if len(sys.argv) > 1 else None


def timeit(fn, reps=6):
  for _ in range(2):
    assert fn() == 0, B.lib.ngb_last_error()
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
for M, K, N in ((8192, 784, 512), (8192, 256, 784), (4096, 256, 120)):
  x, W, b, ug, zp = rnd(M, K), rnd(K, N), rnd(N), rnd(M, N), rnd(M, K)
  z, a, dx = torch.empty(M, N, device='cuda'), torch.empty(M, N, device='cuda'), torch.empty(M, K, device='cuda')
  variants = {
    'z': lambda: B.lib.ngb_linear_fwd(P(x), P(W), None, P(z), None, -1, 0.0, M, K, N, S()),
    'z+bias': lambda: B.lib.ngb_linear_fwd(P(x), P(W), P(b), P(z), None, -1, 0.0, M, K, N, S()),
    'a(relu)': lambda: B.lib.ngb_linear_fwd(P(x), P(W), None, None, P(a), 2, 0.0, M, K, N, S()),
    'z+a': lambda: B.lib.ngb_linear_fwd(P(x), P(W), None, P(z), P(a), 2, 0.0, M, K, N, S()),
    'z+a+bias': lambda: B.lib.ngb_linear_fwd(P(x), P(W), P(b), P(z), P(a), 2, 0.0, M, K, N, S()),
    'dgrad': lambda: B.lib.ngb_linear_dgrad(P(ug), P(W), None, P(dx), -1, 0.0, M, K, N, 0, S()),
    'dgrad+mask': lambda: B.lib.ngb_linear_dgrad(P(ug), P(W), P(zp), P(dx), 2, 0.0, M, K, N, 0, S()),
  }
  for name, fn in variants.items():
    if only and name != only:
      continue
    print(f'[{M},{K},{N}] {name:12s} {timeit(fn):7.1f} us', flush=True)
