"""CTA-pair (cta_group::2) GEMM check: correctness vs float64 and timing vs the single-CTA kernel, per shape and pass.
Run twice: NGB_GEMM_2CTA=0 and =1 (the switch is read once per process)."""
import ctypes, os, sys
import numpy as np
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from neograd_b200 import _backend as B
P = lambda t: ctypes.c_void_p(t.data_ptr())
S = lambda: ctypes.c_void_p(B.stream())
B.ensure_runtime()
flush = torch.empty(256 << 20, device='cuda', dtype=torch.uint8)

def timeit(fn, reps=8):
  for _ in range(3):
    assert fn() == 0, B.lib.ngb_last_error()
  torch.cuda.synchronize()
  ts = []
  for _ in range(reps):
    flush.zero_()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); fn(); e1.record(); torch.cuda.synchronize()
    ts.append(e0.elapsed_time(e1) * 1e3)
  return float(np.mean(ts))

def dist(a, b):
  a, b = a.double().flatten(), b.double().flatten()
  return float((a - b).norm() / (a.norm() + b.norm() + 1e-300))

print('NGB_GEMM_2CTA =', os.environ.get('NGB_GEMM_2CTA'))
for M, K, N in ((8192, 784, 512), (8192, 512, 256), (8192, 256, 784), (4096, 4096, 4096), (640, 200, 264), (300, 64, 136)):
  g = torch.Generator(device='cuda'); g.manual_seed(M + K + N)
  A = torch.randn(M, K, device='cuda', generator=g); Bm = torch.randn(K, N, device='cuda', generator=g) / K ** 0.5
  U = torch.randn(M, N, device='cuda', generator=g)
  C, dA, dB = (torch.full(s, float('nan'), device='cuda') for s in ((M, N), (M, K), (K, N)))
  calls = {
    'fwd': (lambda: B.lib.ngb_gemm(0, 0, 0, M, N, K, P(A), K, P(Bm), N, P(C), N, 0, S()), C, lambda: A.double() @ Bm.double()),
    'dgrad': (lambda: B.lib.ngb_gemm(0, 0, 1, M, K, N, P(U), N, P(Bm), N, P(dA), K, 0, S()), dA, lambda: U.double() @ Bm.double().T),
    'wgrad': (lambda: B.lib.ngb_gemm(0, 1, 0, K, N, M, P(A), K, P(U), N, P(dB), N, 0, S()), dB, lambda: A.double().T @ U.double()),
  }
  for name, (fn, out, ref) in calls.items():
    rc = fn()
    torch.cuda.synchronize()
    if rc != 0:
      print(f'[{M},{K},{N}] {name}: rc={rc} {B.lib.ngb_last_error()}'); continue
    d = dist(out, ref())
    t = timeit(fn)
    print(f'[{M},{K},{N}] {name:6s} dist {d:.2e}  {t:7.1f} us', flush=True)
  rc = B.lib.ngb_debug_check_pipelines()
  if rc != 0:
    print('PIPELINE TIMEOUT:', B.lib.ngb_last_error()); break
