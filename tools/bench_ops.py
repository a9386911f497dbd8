"""Op micro-benchmarks (BASELINE.json config 5 and the bandwidth-bound kernels): times single C-ABI entry points with
CUDA events and reports achieved TFLOP/s or GB/s against the measured peaks.

  python tools/bench_ops.py [gemm] [conv] [bw] [--quick] > gpurun_out/ops.jsonl

Timing: 3 warm-up launches, then `reps` launches each bracketed by its own event pair with a 256 MiB L2 flush in
between; the mean is reported (min also kept).  Algorithmic work per launch as in bench.py:call_work / DESIGN.md."""
import ctypes
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from neograd_b200 import _backend as B  # noqa: E402
import bench  # noqa: E402

PK = bench.peaks()
P = lambda t: ctypes.c_void_p(t.data_ptr())
S = lambda: ctypes.c_void_p(B.stream())
flush = None


def timeit(fn, reps=10):
  global flush
  if flush is None:
    flush = torch.empty(256 << 20, device='cuda', dtype=torch.uint8)
  for _ in range(3):
    fn()
  torch.cuda.synchronize()
  ts = []
  for _ in range(reps):
    flush.zero_()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    fn()
    e1.record()
    torch.cuda.synchronize()
    ts.append(e0.elapsed_time(e1) * 1e-3)
  return float(np.mean(ts)), float(np.min(ts))


def emit(name, shape, by, fl, mean, best, bound, extra=None):
  rec = {'op': name, 'shape': shape, 'us': mean * 1e6, 'us_best': best * 1e6, 'bytes': by, 'flops': fl}
  if bound == 'tensor':
    rec.update(tflops=fl / mean / 1e12, frac_tf32_peak=fl / mean / 1e12 / PK['tf32_tflops'], peak=PK['tf32_tflops'])
  else:
    rec.update(gbs=by / mean / 1e9, frac_hbm_peak=by / mean / 1e9 / PK['hbm_gbs'], peak=PK['hbm_gbs'])
  if extra:
    rec.update(extra)
  print(json.dumps(rec), flush=True)


def rnd(*shape):
  return torch.randn(*shape, device='cuda', dtype=torch.float32)


def bench_gemm(quick):
  shapes = [(8192, 784, 512), (8192, 512, 256), (8192, 256, 784), (4096, 4096, 4096), (8192, 8192, 8192),
            (65536, 256, 256), (262144, 128, 128), (1048576, 64, 64), (1048576, 256, 256)]
  if quick:
    shapes = shapes[:4]
  # cuBLAS TF32 reference point (the denominator BASELINE.md asks to measure)
  torch.backends.cuda.matmul.allow_tf32 = True
  a, b = rnd(8192, 8192), rnd(8192, 8192)
  m, best = timeit(lambda: torch.matmul(a, b), 10)
  print(json.dumps({'op': 'cublas_tf32_matmul_8192', 'us': m * 1e6, 'tflops': 2 * 8192**3 / m / 1e12,
                    'tflops_best': 2 * 8192**3 / best / 1e12}), flush=True)
  del a, b
  for M, K, N in shapes:
    A, Bm, U = rnd(M, K), rnd(K, N), rnd(M, N)
    C, dA, dB = torch.empty(M, N, device='cuda'), torch.empty(M, K, device='cuda'), torch.empty(K, N, device='cuda')
    by = 4 * (M * K + K * N + M * N)
    fl = 2 * M * N * K
    for name, call in (
        ('dot_fwd', lambda: B.lib.ngb_gemm(0, 0, 0, M, N, K, P(A), K, P(Bm), N, P(C), N, 0, S())),
        ('dot_dgrad', lambda: B.lib.ngb_gemm(0, 0, 1, M, K, N, P(U), N, P(Bm), N, P(dA), K, 0, S())),
        ('dot_wgrad', lambda: B.lib.ngb_gemm(0, 1, 0, K, N, M, P(A), K, P(U), N, P(dB), N, 0, S()))):
      assert call() == 0, B.lib.ngb_last_error()
      m, best = timeit(call, 5 if M * N * K > 1e12 else 10)
      emit(name, [M, K, N], by, fl, m, best, 'tensor')
    del A, Bm, U, C, dA, dB


def bench_conv(quick):
  cases = [(256, 64, 32), (256, 128, 32), (256, 256, 32), (256, 64, 64), (256, 128, 64), (256, 256, 64), (256, 64, 128),
           (256, 128, 128), (256, 256, 128)]
  if quick:
    cases = [(256, 64, 32), (256, 128, 64)]
  for N, C, H in cases:
    O, W, KH, KW, pad, stride = C, H, 3, 3, 1, 1
    x, w, b = rnd(N, C, H, W), rnd(O, C, KH, KW) / np.sqrt(9 * C), torch.zeros(O, device='cuda')
    y, ug = torch.empty(N, O, H, W, device='cuda'), rnd(N, O, H, W)
    dx, dw, db = torch.empty_like(x), torch.empty_like(w), torch.empty(O, device='cuda')
    by = 4 * (x.numel() + w.numel() + y.numel())
    fl = 2 * N * O * C * KH * KW * H * W
    g = (N, C, H, W, O, KH, KW, pad, stride)
    for name, call in (
        ('conv_fprop', lambda: B.lib.ngb_conv_fprop(P(x), P(w), P(b), P(y), *g, S())),
        ('conv_dgrad', lambda: B.lib.ngb_conv_dgrad(P(ug), P(w), P(dx), *g, 0, S())),
        ('conv_wgrad', lambda: B.lib.ngb_conv_wgrad(P(ug), P(x), P(dw), *g, 0, S()))):
      rc = call()
      if rc != 0:
        print(json.dumps({'op': name, 'shape': list(g), 'error': B.lib.ngb_last_error().decode()}), flush=True)
        continue
      m, best = timeit(call, 3 if fl > 1e12 else 5)
      emit(name, list(g), by, fl, m, best, 'tensor')
    call = lambda: B.lib.ngb_conv_bgrad(P(ug), P(db), N, O, H * W, 0, S())
    call()
    m, best = timeit(call, 5)
    emit('conv_bgrad', [N, O, H * W], 4 * ug.numel(), ug.numel(), m, best, 'hbm')
    del x, w, y, ug, dx, dw


def bench_bw(quick):
  n = 1 << (24 if quick else 28)
  a, b2, out = rnd(n), rnd(n), torch.empty(n, device='cuda')
  sh = (ctypes.c_int64 * 1)(n)
  for name, by, call in (
      ('ew_add_flat', 12 * n, lambda: B.lib.ngb_ew_binary_fwd(B.ADD, P(a), 0.0, sh, 1, P(b2), 0.0, sh, 1, P(out), S())),
      ('ew_mul_scalar', 8 * n, lambda: B.lib.ngb_ew_binary_fwd(B.MUL, P(a), 0.0, sh, 1, None, 2.0, sh, 0, P(out), S())),
      ('relu_fwd', 8 * n, lambda: B.lib.ngb_ew_unary_fwd(B.RELU, 0.0, P(a), P(out), n, S())),
      ('relu_bwd', 12 * n, lambda: B.lib.ngb_ew_unary_bwd(B.RELU, 0.0, P(a), P(b2), P(out), n, 0, S())),
      ('sigmoid_fwd', 8 * n, lambda: B.lib.ngb_ew_unary_fwd(B.SIGMOID, 0.0, P(a), P(out), n, S())),
      ('fill', 4 * n, lambda: B.lib.ngb_fill(P(out), 1.0, n, S())),
      ('axpy', 12 * n, lambda: B.lib.ngb_axpy(P(out), P(a), 0.5, n, S())),
      ('sgd_step', 12 * n, lambda: B.lib.ngb_sgd_step(P(out), P(a), n, 1e-3, 1.0, S()))):
    assert call() == 0
    m, best = timeit(call)
    emit(name, [n], by, n, m, best, 'hbm')
  na = n // 2
  p, g, mm, vv = rnd(na), rnd(na), torch.zeros(na, device='cuda'), torch.zeros(na, device='cuda')
  call = lambda: B.lib.ngb_adam_step(P(p), P(g), P(mm), P(vv), na, 1e-3, 0.9, 0.999, 1e-8, 1, 1.0, S())
  call()
  m, best = timeit(call)
  emit('adam_step', [na], 28 * na, 12 * na, m, best, 'hbm')
  del p, g, mm, vv
  # bias add + its unbroadcast (column reduction) at Dot output shapes
  for M, Nn in ((8192, 784), (65536, 256), (1048576, 256)) if not quick else ((8192, 784),):
    A, bias, o2, gb = rnd(M, Nn), rnd(1, Nn), torch.empty(M, Nn, device='cuda'), torch.empty(1, Nn, device='cuda')
    sa, sb = (ctypes.c_int64 * 2)(M, Nn), (ctypes.c_int64 * 2)(1, Nn)
    call = lambda: B.lib.ngb_ew_binary_fwd(B.ADD, P(A), 0.0, sa, 2, P(bias), 0.0, sb, 2, P(o2), S())
    call()
    m, best = timeit(call)
    emit('bias_add', [M, Nn], 4 * (2 * M * Nn + Nn), M * Nn, m, best, 'hbm')
    call = lambda: B.lib.ngb_ew_binary_bwd(B.ADD, 1, P(A), P(A), 0.0, sa, 2, P(bias), 0.0, sb, 2, P(gb), 0, S())
    call()
    m, best = timeit(call)
    emit('bias_grad_unbroadcast', [M, Nn], 4 * (M * Nn + Nn), M * Nn, m, best, 'hbm')
    del A, o2
  # maxpool k2 s2 on (256, C, H, W)
  for C, H in ((64, 64), (128, 128)) if not quick else ((64, 32),):
    outer = 256 * C
    x = rnd(outer, H, H)
    y, idx = torch.empty(outer, H // 2, H // 2, device='cuda'), torch.empty(outer, H // 2, H // 2, device='cuda', dtype=torch.uint8)
    dxx = torch.empty_like(x)
    by = 4 * x.numel() + 5 * y.numel()
    call = lambda: B.lib.ngb_maxpool_fwd(P(x), P(y), P(idx), outer, H, H, 2, 2, 0, 2, S())
    call()
    m, best = timeit(call)
    emit('maxpool_fwd', [outer, H, H], by, x.numel(), m, best, 'hbm')
    call = lambda: B.lib.ngb_maxpool_bwd(P(y), P(idx), P(dxx), outer, H, H, 2, 2, 0, 2, 0, S())
    call()
    m, best = timeit(call)
    emit('maxpool_bwd', [outer, H, H], by, x.numel(), m, best, 'hbm')
    del x, y, idx, dxx


if __name__ == '__main__':
  B.ensure_runtime()
  args = sys.argv[1:]
  quick = '--quick' in args
  which = [a for a in args if not a.startswith('--')] or ['gemm', 'conv', 'bw']
  print(json.dumps({'peaks': PK}), flush=True)
  if 'gemm' in which:
    bench_gemm(quick)
  if 'bw' in which:
    bench_bw(quick)
  if 'conv' in which:
    bench_conv(quick)
