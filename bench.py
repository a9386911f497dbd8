#!/usr/bin/env python
"""bench.py -- training throughput of the neograd hot path on B200 (BASELINE.json: "CNN train samples/s at 1/2/4/8 B200;
Conv2D/Dot fwd+bwd TFLOP/s vs peak").

  python bench.py --gpus N --steps K --warmup W            # this repo's CUDA path (one rank per GPU under torchrun)
  python bench.py --impl reference --steps K --warmup W     # the reference's own CPU implementation on the host cores

A "step" is one full training iteration of the hot path over one synthetic batch, through the public neograd API:
  lenet_b (headline; BASELINE config 3, LeNet-5 channels): zero_grad, forward, SoftmaxCE, backward, (gradient exchange,)
          fused Adam on 4096 synthetic 28x28x1 images PER GPU (weak scaling) -- README.md:109-115 of the reference
  lenet_a (config 3 read literally: neograd's single-channel Conv2D / MaxPool2D)
  gan784  (config 4): generator step + discriminator step of examples/gans/basic.ipynb cells 9-10 at 784-512-256,
          8192 real rows per GPU, two Adam optimizers

ONE JSON line on stdout.  The headline keys describe `--workload` (default lenet_b):
  value        whole-job samples/s with the batch resident in HBM; every step is the CUDA-graph replay of the captured
               step (all kernels run every step; nothing is cached), timed with CUDA events per step, L2 flushed
               between timed steps, max over ranks
  e2e          same metric through the public API with HOST inputs: per step a pinned H2D copy of the batch, the step,
               and a D2H read of the loss; wall clock between synchronised barriers
  eager        same step dispatched op by op from Python (no graph capture), for comparison
  roofline     the launch site (entry point + shape) with the most device time in the step AS DISPATCHED IN THE TIMED
               REGION (backward lanes on): algorithmic bytes or flops / its mean CUDA-event time, events
               recorded on the launching stream behind a queue backlog; `group` = the same for the dominant entry point
  cpu_baseline the unmodified reference (baseline/_ref, kind "reference") or, when it is absent, the oracle port, on a
               bounded sample, rank 0, N=1
Extra keys: `workloads` (the other workloads' value / ms_per_step / roofline, and the strong-scaling LeNet-B figure:
global batch 4096 split over the N GPUs), `ops` (BASELINE config 5: Dot and Conv3D fwd / data-grad / weight-grad TFLOP/s
and the bandwidth kernels' GB/s against the peaks, N=1 only), `peaks` (incl. a cuBLAS TF32 8192^3 matmul timed in this
process: the measured TF32 denominator).
"""
import argparse
import ctypes
import importlib.util
import json
import os
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
  sys.path.insert(0, ROOT)

E2E_REPEATS = 5
WORKLOADS = ('lenet_b', 'lenet_a', 'gan784')
BATCH = {'lenet_b': 4096, 'lenet_a': 4096, 'gan784': 8192}     # samples per GPU (weak scaling)
METRIC = {'lenet_b': 'cnn_train_samples_per_s', 'lenet_a': 'cnn_train_samples_per_s', 'gan784': 'gan_train_samples_per_s'}
DESCR = {
  'lenet_b': 'lenet_b: LeNet-style CNN (Conv3D 1-6-16, MaxPool3D, Linear 256-120-84-10), synthetic 28x28x1, fwd+bwd+Adam(1e-3), SoftmaxCE',
  'lenet_a': 'lenet_a: single-channel Conv2D / MaxPool2D LeNet variant, synthetic 28x28, fwd+bwd+Adam(1e-3), SoftmaxCE',
  'gan784': 'gan784: MLP GAN 784-512-256 (generator step + discriminator step, BCE, two Adam(5e-3)), synthetic 784-d rows',
}


def peaks():
  p = {'hbm_gbs': 6650.0, 'bf16_tflops': 1590.0, 'source': 'fallback'}
  try:
    with open(os.path.join(ROOT, 'MEASURED_PEAKS.json')) as f:
      m = json.load(f)
    p.update(hbm_gbs=float(m['hbm_gbs']), bf16_tflops=float(m['bf16_tflops']), source='measured')
  except Exception:
    pass
  # TF32 dense rate is half the BF16 rate on B200; run_device() replaces this with a cuBLAS TF32 matmul timed in-process
  p['tf32_tflops'] = p['bf16_tflops'] / 2
  p['tf32_source'] = 'half of the BF16 peak'
  return p


def measure_tf32_peak(pk):
  """cuBLAS TF32 8192^3 matmul (torch.matmul, allow_tf32) timed with CUDA events: best of 10 (burst) and back to back
  for ~1 s (sustained) -- the method MEASURED_PEAKS.json uses for BF16."""
  import torch
  prev = torch.backends.cuda.matmul.allow_tf32
  torch.backends.cuda.matmul.allow_tf32 = True
  try:
    n = 8192
    a = torch.randn(n, n, device='cuda')
    b = torch.randn(n, n, device='cuda')
    c = torch.empty(n, n, device='cuda')
    for _ in range(3):
      torch.matmul(a, b, out=c)
    best = 1e9
    for _ in range(10):
      e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
      e0.record()
      torch.matmul(a, b, out=c)
      e1.record()
      e1.synchronize()
      best = min(best, e0.elapsed_time(e1))
    reps = max(10, int(1000.0 / best))
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
      torch.matmul(a, b, out=c)
    e1.record()
    e1.synchronize()
    sustained = e0.elapsed_time(e1) / reps
    fl = 2.0 * n ** 3
    pk['tf32_half_bf16_tflops'] = pk['bf16_tflops'] / 2
    pk['tf32_cublas_burst_tflops'] = fl / (best * 1e-3) / 1e12
    pk['tf32_cublas_sustained_tflops'] = fl / (sustained * 1e-3) / 1e12
    # denominator for every tensor-bound fraction: the larger of the two estimates (the stricter one)
    if pk['tf32_cublas_burst_tflops'] > pk['tf32_tflops']:
      pk['tf32_tflops'] = pk['tf32_cublas_burst_tflops']
      pk['tf32_source'] = 'cuBLAS TF32 8192^3 matmul timed in this process (best of 10)'
    else:
      pk['tf32_source'] = 'half of the BF16 peak (the cuBLAS TF32 matmul timed in this process is lower)'
    del a, b, c
  finally:
    torch.backends.cuda.matmul.allow_tf32 = prev
  return pk


# --------------------------------------------------------------------------------------------- reference arm (CPU)
def load_reference():
  """The UNMODIFIED reference under the alias `neograd_ref`: $NEOGRAD_REF, /root/reference (build container), or
  baseline/_ref (pip --target install made by __graft_entry__.build(); travels to the GPU box).  None if absent."""
  for root in (os.environ.get('NEOGRAD_REF'), '/root/reference', os.path.join(ROOT, 'baseline', '_ref')):
    if root and os.path.exists(os.path.join(root, 'neograd', '__init__.py')):
      try:
        spec = importlib.util.spec_from_file_location('neograd_ref', os.path.join(root, 'neograd', '__init__.py'),
                                                      submodule_search_locations=[os.path.join(root, 'neograd')])
        m = importlib.util.module_from_spec(spec)
        sys.modules['neograd_ref'] = m
        spec.loader.exec_module(m)
        return m, root
      except Exception as e:   # noqa: BLE001 (e.g. dill missing)
        print(f'bench.py: reference at {root} failed to import ({type(e).__name__}: {e})', file=sys.stderr)
        sys.modules.pop('neograd_ref', None)
  return None, None


def _timed_steps(step, steps, warmup, budget_s):
  """`warmup` untimed + up to `steps` timed calls, cut short when the wall-clock budget is used up (>= 2 timed)."""
  t_start = time.perf_counter()
  for _ in range(warmup):
    step()
  times = []
  for i in range(steps):
    t0 = time.perf_counter()
    step()
    times.append(time.perf_counter() - t0)
    if i >= 1 and time.perf_counter() - t_start + times[-1] > budget_s:
      break
  return times


def _workload_defs():
  """neograd_b200/workload_defs.py loaded BY PATH: the user-level model / data / loop definitions, with no import of
  this repo's package -- the reference arm runs none of our engine."""
  spec = importlib.util.spec_from_file_location('ngb_workload_defs', os.path.join(ROOT, 'neograd_b200', 'workload_defs.py'))
  m = importlib.util.module_from_spec(spec)
  spec.loader.exec_module(m)
  return m


def reference_steps(workload, batch, steps, warmup, budget_s):
  """The reference's own implementation (the real package, its public API and stock code path) of one training step,
  timed.  None if the package is absent."""
  ref, root = load_reference()
  if ref is None:
    return None
  defs = _workload_defs()
  classes = defs.model_classes(ref.nn)
  import types
  from neograd_ref.nn import optim as _roptim, loss as _rloss     # the reference's users import these submodules explicitly
  rnn = types.SimpleNamespace(optim=_roptim, loss=_rloss)
  if workload.startswith('lenet'):
    kind = workload[-1]
    x, y = defs.lenet_data(kind, batch, 0)
    model = (classes['LeNetA'] if kind == 'a' else classes['LeNetB'])()
    defs.set_params(model, defs.synthetic_params([p.shape for p in model.parameters()], 1, 0.1))
    X, Y = ref.tensor(x.astype(np.float64)), ref.tensor(y.astype(np.float64))
    optim = rnn.optim.Adam(model.parameters(), 1e-3)
    loss_fn = rnn.loss.SoftmaxCE(axis=1)
    step = lambda: defs.train_step(model, optim, loss_fn, X, Y)
  else:
    x, ng_, nd_ = defs.gan_data(batch)
    model = classes['GAN'](784, 512, 256)
    defs.set_params(model, defs.synthetic_params([p.shape for p in model.parameters()], 1, 0.05))
    og, od = rnn.optim.Adam(model.parameters(), 5e-3), rnn.optim.Adam(model.parameters(), 5e-3)
    loss_fn = rnn.loss.BCE()
    REAL, NG, ND = (ref.tensor(a.astype(np.float64)) for a in (x, ng_, nd_))
    ones, zeros = ref.tensor(np.ones((batch, 1))), ref.tensor(np.zeros((batch, 1)))
    step = lambda: defs.gan_iteration(model, og, od, loss_fn, NG, ND, REAL, ones, zeros)
  times = _timed_steps(step, steps, warmup, budget_s)
  return {'times': times, 'root': root}


def port_steps(workload, batch, steps, warmup, budget_s):
  """The oracle port (float64 NumPy restatement of the reference algorithm, oracle/) of one training step, timed."""
  from oracle import models
  rng = np.random.default_rng(0)
  if workload.startswith('lenet'):
    kind = workload[-1]
    x = rng.standard_normal((batch, 28, 28) if kind == 'a' else (batch, 1, 28, 28))
    y = np.eye(10)[rng.integers(0, 10, batch)]
    net = (models.lenet_a if kind == 'a' else models.lenet_b)(np.random.default_rng(1))
    opt = models.Adam(net, 1e-3)
    step = lambda: models.train_step(net, opt, 'softmax_ce', x, y)
  else:
    gen, disc = models.gan_nets(np.random.default_rng(1))
    og, od = models.Adam(gen, 5e-3), models.Adam(disc, 5e-3)
    real, ng_, nd_ = (rng.standard_normal((batch, 784)) for _ in range(3))
    step = lambda: models.gan_iteration(gen, disc, og, od, ng_, nd_, real)
  return {'times': _timed_steps(step, steps, warmup, budget_s)}


def run_reference(args):
  """The reference arm: the reference's own CPU implementation of the step on the box's host cores, SAME batch as the
  device arm's per-GPU batch; rank 0 alone.  The unmodified reference when it can be imported (kind "reference"), the
  oracle port otherwise (kind "port"); the port is also reported as a second, stronger baseline (`strong_baseline`)."""
  rank = int(os.environ.get('RANK', '0'))
  if rank != 0:
    return
  wl = args.workload
  batch = BATCH[wl]
  cores = len(os.sched_getaffinity(0))
  budget = float(os.environ.get('NGB_REF_BUDGET_S', '150'))
  r = reference_steps(wl, batch, args.steps, 1, budget)
  kind = 'reference'
  if r is None:
    kind = 'port'
    r = port_steps(wl, batch, args.steps, max(1, min(args.warmup, 2)), budget)
  times = r['times']
  per = float(np.median(times))
  v = batch / per
  strong = None
  if kind == 'reference':
    ps = port_steps(wl, batch, 3, 1, 30.0)['times']
    strong = {'value': batch / float(np.median(ps)), 'unit': 'samples/s', 'kind': 'port', 'steps': len(ps),
              'note': 'oracle port (sliding-window + einsum restatement, float64): a faster CPU formulation of the same arithmetic'}
  what = ('the UNMODIFIED reference package (pranftw/neograd 0.0.4, float64 NumPy) from ' + str(r.get('root'))) if kind == 'reference' \
      else 'oracle port of the reference CPU path (the reference package is not importable here)'
  sample = (f'{len(times)} timed steps (1 warm-up; --steps {args.steps} requested, bounded by a {budget:.0f} s budget) of batch {batch} '
            f'(= the device arm\'s per-GPU batch) of {wl}, OpenBLAS default threads')
  print(json.dumps({
    'impl': 'reference', 'metric': METRIC[wl], 'value': v, 'unit': 'samples/s', 'n_gpus': args.gpus,
    'steps': len(times), 'steps_requested': args.steps, 'warmup': 1, 'ms_per_step': per * 1e3, 'higher_is_better': True,
    'scaling': 'weak', 'vs_baseline': None, 'dtype': 'f64', 'data': 'synthetic',
    'config': {'workload': DESCR[wl], 'batch_per_step': batch, 'same_config': True, 'note': what},
    'cpu_baseline': {'value': v, 'unit': 'samples/s', 'cores': cores, 'kind': kind, 'sample': sample},
    'strong_baseline': strong,
    'e2e': {'value': v, 'unit': 'samples/s', 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0},
  }))


# --------------------------------------------------------------------------------------------- clocks
class ClockSampler(threading.Thread):
  def __init__(self, index):
    super().__init__(daemon=True)
    self.index, self.samples, self.reasons, self.stop_flag, self.max_mhz = index, [], set(), False, None
    try:
      import pynvml
      pynvml.nvmlInit()
      self.nv = pynvml
      self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
      self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
    except Exception:
      self.nv = None

  def run(self):
    if self.nv is None:
      return
    nv = self.nv
    names = {'hw_slowdown': getattr(nv, 'nvmlClocksEventReasonHwSlowdown', 0x8),
             'hw_thermal_slowdown': getattr(nv, 'nvmlClocksEventReasonHwThermalSlowdown', 0x40),
             'sw_thermal_slowdown': getattr(nv, 'nvmlClocksEventReasonSwThermalSlowdown', 0x20),
             'sw_power_cap': getattr(nv, 'nvmlClocksEventReasonSwPowerCap', 0x4)}
    self.names = names
    while not self.stop_flag:
      self.sample()
      time.sleep(0.005)

  def sample(self):
    nv = self.nv
    try:
      self.samples.append(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM))
      try:
        r = nv.nvmlDeviceGetCurrentClocksEventReasons(self.h)
      except Exception:
        r = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
      for k, bit in self.names.items():
        if r & bit:
          self.reasons.add(k)
    except Exception:
      pass

  def result(self):
    """Called right after the timed region's closing synchronize: one more sample is taken here so that a region shorter
    than an NVML query (20 steps of 0.2 ms) is still bracketed by a sample on each side."""
    self.stop_flag = True
    if self.nv is not None and hasattr(self, 'names'):
      self.sample()
    if not self.samples:
      return {'sm_mhz': None, 'sm_max_mhz': self.max_mhz, 'reasons': []}
    return {'sm_mhz': float(np.median(self.samples)), 'sm_max_mhz': self.max_mhz, 'reasons': sorted(self.reasons),
            'samples': len(self.samples)}


# --------------------------------------------------------------------------------------------- algorithmic work per call
def _numel(shape_arr, rank):
  n = 1
  for i in range(rank):
    n *= shape_arr[i]
  return n


def call_work(name, a):
  """(algorithmic bytes, flops) of one C-ABI call, from its arguments (formulas: SURVEY section 8d / DESIGN.md)."""
  if name == 'ngb_gemm':
    M, N, K = a[3], a[4], a[5]
    return 4 * (M * K + K * N + M * N), 2 * M * N * K
  if name == 'ngb_linear_fwd':
    # Dot + bias (+ activation) in one kernel: reads x, W, b; writes the pre-activation and / or the activation
    M, K, N = a[7], a[8], a[9]                      # (x, W, b, z, a, act, alpha, batch, n_in, n_out, stream)
    outs = (1 if a[3] else 0) + (1 if a[4] else 0)
    return 4 * (M * K + K * N + N + outs * M * N), 2 * M * N * K
  if name == 'ngb_linear_dgrad':
    # dx = (ug . W^T) * act'(z_prev): reads ug, W and (if masked) z_prev; writes dx
    M, K, N = a[6], a[7], a[8]                      # (ug, W, zprev, dx, act, alpha, batch, n_in, n_out, accumulate, stream)
    return 4 * (M * N + K * N + (2 if a[2] else 1) * M * K), 2 * M * N * K
  if name == 'ngb_linear_wgrad':
    # dW = x^T . ug and db = column sums of ug from the same pass
    M, K, N = a[4], a[5], a[6]                      # (x, ug, dW, db, batch, n_in, n_out, acc_w, acc_b, stream)
    return 4 * (M * K + M * N + K * N + N), 2 * M * N * K + M * N
  if name == 'ngb_linear_softmax_ce':
    M, K, N = a[7], a[8], a[9]                      # (x, W, b, t, z, loss, dz, batch, n_in, n_out, ...)
    return 4 * (M * K + K * N + N + 3 * M * N), 2 * M * N * K + 12 * M * N
  if name in ('ngb_bce_fwd', 'ngb_bce_bwd'):
    n = a[2] if name == 'ngb_bce_fwd' else a[3]
    return (8 if name == 'ngb_bce_fwd' else 12) * n, 8 * n
  if name == 'ngb_xgpu_exchange_update':
    n, world = a[7], a[3]                           # push to world-1 peers, read world slots, 28 B/element of Adam traffic
    return (4 * (world - 1) + 4 * world + 28) * n, 12 * n
  if name == 'ngb_prep_u8_affine':
    return 5 * a[2], 2 * a[2]
  if name == 'ngb_prep_one_hot':
    return 4 * a[2] + 4 * a[2] * a[3], a[2] * a[3]
  if name == 'ngb_randn':
    return 4 * a[1], 40 * a[1]
  if name in ('ngb_conv_fprop', 'ngb_conv_dgrad', 'ngb_conv_wgrad'):
    o = 4 if name == 'ngb_conv_fprop' else 3
    N, C, H, W, O, KH, KW, pad, stride = a[o:o + 9]
    Ho, Wo = (H + 2 * pad - KH) // stride + 1, (W + 2 * pad - KW) // stride + 1
    return 4 * (N * C * H * W + O * C * KH * KW + N * O * Ho * Wo), 2 * N * O * C * KH * KW * Ho * Wo
  if name in ('ngb_conv_dgrad_pooled', 'ngb_conv_wgrad_pooled'):
    # the conv gradient taken from the 2x2-pooled upstream gradient: per pooling window the kernel needs ug_pool (4 B) and
    # the argmax byte (cell + ReLU flag) instead of the 4 full-resolution gradient values (16 B)
    N, C, H, W, O, KH, KW, pad, stride = a[6:15]
    Ho, Wo = (H + 2 * pad - KH) // stride + 1, (W + 2 * pad - KW) // stride + 1
    win = N * O * (Ho // 2) * (Wo // 2)
    return 4 * (N * C * H * W + O * C * KH * KW) + 5 * win, 2 * N * O * C * KH * KW * Ho * Wo
  if name == 'ngb_conv_relu_pool_fwd':
    # conv + bias + ReLU + 2x2 pool in one pass: reads x and the filter bank, writes the pooled map and one argmax byte per window
    N, C, H, W, O, KH, KW, pad, stride = a[5:14]
    Ho, Wo = (H + 2 * pad - KH) // stride + 1, (W + 2 * pad - KW) // stride + 1
    win = N * O * (Ho // 2) * (Wo // 2)
    return 4 * (N * C * H * W + O * C * KH * KW + O) + 5 * win, 2 * N * O * C * KH * KW * Ho * Wo
  if name == 'ngb_conv_wbgrad_pooled':
    # weight + bias gradient in one pass: reads x, the pooled gradient and the argmax bytes; one window cell carries gradient
    N, C, H, W, O, KH, KW, pad, stride = a[7:16]
    Ho, Wo = (H + 2 * pad - KH) // stride + 1, (W + 2 * pad - KW) // stride + 1
    win = N * O * (Ho // 2) * (Wo // 2)
    return 4 * (N * C * H * W + O * C * KH * KW + O) + 5 * win, 2 * win * C * KH * KW + win
  if name == 'ngb_conv_bgrad_pooled_idx':
    win = a[3] * a[4] * a[5]
    return 5 * win + 4 * a[4], win
  if name == 'ngb_conv_bgrad_pooled':
    win = a[4] * a[5] * (a[6] // 2) * (a[7] // 2)
    return 12 * win + 4 * a[5], win
  if name == 'ngb_maxpool_relu_mask':
    win = a[4] * (a[5] // 2) * (a[6] // 2)
    return 16 * win, win
  if name == 'ngb_maxpool_relu_fwd':
    outer, H, W, KH, KW, pad, stride = a[3:10]
    Ho, Wo = (H + 2 * pad - KH) // stride + 1, (W + 2 * pad - KW) // stride + 1
    return 4 * outer * H * W + 5 * outer * Ho * Wo, outer * Ho * Wo * KH * KW
  if name == 'ngb_maxpool_relu_bwd':
    outer, H, W, KH, KW, pad, stride = a[4:11]
    Ho, Wo = (H + 2 * pad - KH) // stride + 1, (W + 2 * pad - KW) // stride + 1
    return 8 * outer * H * W + 5 * outer * Ho * Wo, outer * Ho * Wo * KH * KW
  if name == 'ngb_conv_bgrad':
    return 4 * (a[2] * a[3] * a[4] + a[3]), a[2] * a[3] * a[4]
  if name in ('ngb_maxpool_fwd', 'ngb_maxpool_bwd'):
    outer, H, W, KH, KW, pad, stride = a[3:10]
    Ho, Wo = (H + 2 * pad - KH) // stride + 1, (W + 2 * pad - KW) // stride + 1
    return 4 * outer * H * W + 5 * outer * Ho * Wo, outer * Ho * Wo * KH * KW
  if name == 'ngb_ew_unary_fwd':
    return 8 * a[4], a[4]
  if name == 'ngb_ew_unary_bwd':
    return 12 * a[5], a[5]
  if name == 'ngb_ew_binary_fwd':
    na, nb = (_numel(a[3], a[4]) if a[1] else 0), (_numel(a[7], a[8]) if a[5] else 0)
    sa = [a[3][i] for i in range(a[4])]
    sb = [a[7][i] for i in range(a[8])]
    no = int(np.prod(np.broadcast_shapes(tuple(sa), tuple(sb)))) if (sa or sb) else 1
    return 4 * (na + nb + no), no
  if name == 'ngb_ew_binary_bwd':
    sa = [a[5][i] for i in range(a[6])]
    sb = [a[9][i] for i in range(a[10])]
    nu = int(np.prod(np.broadcast_shapes(tuple(sa), tuple(sb)))) if (sa or sb) else 1
    ng_ = int(np.prod((sa, sb)[a[1]])) if (sa, sb)[a[1]] else 1
    other = 0
    if a[0] in (2, 3, 4):   # mul / div / pow read the operands
      other = (int(np.prod(sa)) if a[3] else 0) + (int(np.prod(sb)) if a[7] else 0)
    return 4 * (nu + other + ng_), nu
  if name in ('ngb_adam_step', 'ngb_adam_step_dev'):
    return 28 * a[4], 12 * a[4]
  if name == 'ngb_sgd_step':
    return 12 * a[2], 2 * a[2]
  if name in ('ngb_softmax_ce_fwd', 'ngb_softmax_ce_bwd'):
    o = 2 if name == 'ngb_softmax_ce_fwd' else 3
    n = a[o] * a[o + 1] * a[o + 2]
    return (8 if name == 'ngb_softmax_ce_fwd' else 12) * n, 10 * n
  if name in ('ngb_fill',):
    return 4 * a[2], 0
  if name in ('ngb_axpy',):
    return 12 * a[3], 2 * a[3]
  if name == 'ngb_copy':
    return 8 * a[2], 0
  if name == 'ngb_reduce_sum':
    return 4 * _numel(a[1], a[2]), _numel(a[1], a[2])
  return 0, 0


def _roof(name, args, mean_s, pk):
  by, fl = call_work(name, args)
  ridge = pk['tf32_tflops'] * 1e12 / (pk['hbm_gbs'] * 1e9)
  if fl / max(by, 1) > ridge:
    r = {'kernel': name, 'bound': 'tensor', 'achieved': fl / mean_s / 1e12, 'peak': pk['tf32_tflops'], 'unit': 'TFLOP/s',
         'peak_source': pk['tf32_source']}
  else:
    r = {'kernel': name, 'bound': 'hbm', 'achieved': by / mean_s / 1e9, 'peak': pk['hbm_gbs'], 'unit': 'GB/s',
         'peak_source': pk['source'] + ' (MEASURED_PEAKS.json hbm_gbs)'}
  r.update(frac=r['achieved'] / r['peak'], algorithmic_bytes=by, algorithmic_flops=fl, mean_launch_us=mean_s * 1e6)
  return r


# --------------------------------------------------------------------------------------------- workloads (device arm)
class LeNetWorkload:
  """README.md:109-115 on the LeNet configs.  `tensors` are the step's device inputs; `make_step(inputs)` returns the
  step closure over them (the e2e pipeline builds two, one per input buffer)."""

  def __init__(self, ng, name, batch, rank):
    from neograd_b200.configs import lenet_synthetic, train_step
    self.name, self.batch = name, batch
    self.model, x, y = lenet_synthetic(name[-1], batch, seed=100 + rank)
    self.host_inputs = [x, y]
    self.optim = None
    self.ng, self.train_step = ng, train_step
    self.loss_fn = ng.nn.loss.SoftmaxCE(axis=1)

  def params(self):
    return self.model.parameters()

  def setup_optim(self):
    self.optim = self.ng.nn.optim.Adam(self.model.parameters(), 1e-3)

  def make_step(self, tensors):
    X, Y = tensors
    model, optim, loss_fn, train_step = self.model, self.optim, self.loss_fn, self.train_step
    return lambda: (train_step(model, optim, loss_fn, X, Y),)

  E2E_INPUT = 'uint8 28x28 images + int32 class ids in pinned host memory; normalisation and one-hot on the device (nn.DeviceLoader)'

  def e2e_loader(self, nbatches, rank):
    """End-to-end input: images as they sit on disk / in host RAM (uint8 pixels, integer class ids), streamed through the
    library's DeviceLoader (pinned H2D on a copy stream, double buffered; cast + normalise + one-hot as device kernels)."""
    rng = np.random.default_rng(1000 + rank)
    shape = (28, 28) if self.name[-1] == 'a' else (1, 28, 28)
    imgs = rng.integers(0, 256, (nbatches * self.batch,) + shape, dtype=np.uint8)
    labels = rng.integers(0, 10, nbatches * self.batch).astype(np.int32)
    return self.ng.nn.DeviceLoader(imgs, labels, self.batch, normalize=(1.0 / 73.9, -127.5 / 73.9), num_classes=10,
                                   prepare_ahead=True)

  def make_e2e_step(self, slot_tensors):
    return self.make_step(slot_tensors)


class GanWorkload:
  """examples/gans/basic.ipynb cells 9-10 at 784-512-256: one generator step + one discriminator step per iteration."""

  def __init__(self, ng, name, batch, rank):
    from neograd_b200.configs import gan_synthetic, gan_iteration
    self.name, self.batch = name, batch
    self.model, real, noise_g, noise_d = gan_synthetic(batch, seed=100 + rank)
    self.host_inputs = [real, noise_g, noise_d]
    self.ng, self.gan_iteration = ng, gan_iteration
    self.loss_fn = ng.nn.loss.BCE()
    self.ones, self.zeros = ng.tensor(np.ones((batch, 1))), ng.tensor(np.zeros((batch, 1)))

  def params(self):
    return self.model.parameters()

  def setup_optim(self):
    self.og = self.ng.nn.optim.Adam(self.model.parameters(), 5e-3)
    self.od = self.ng.nn.optim.Adam(self.model.parameters(), 5e-3)

  def make_step(self, tensors):
    REAL, NG, ND = tensors
    it, model, og, od, lf, ones, zeros = self.gan_iteration, self.model, self.og, self.od, self.loss_fn, self.ones, self.zeros
    return lambda: it(model, og, od, lf, NG, ND, REAL, ones, zeros)

  E2E_INPUT = ('uint8 784-d real rows in pinned host memory, normalised on the device (nn.DeviceLoader); the two noise batches are '
               'drawn on the device inside the step (nn.randn, Philox), as np.random.randn draws them on the host in the notebook')

  def e2e_loader(self, nbatches, rank):
    rng = np.random.default_rng(1000 + rank)
    rows = rng.integers(0, 256, (nbatches * self.batch, 784), dtype=np.uint8)
    return self.ng.nn.DeviceLoader(rows, None, self.batch, normalize=(1.0 / 73.9, -127.5 / 73.9), prepare_ahead=True)

  def make_e2e_step(self, slot_tensors):
    import torch
    ng = self.ng
    REAL = slot_tensors[0]
    NG, ND = ng.tensor(np.zeros((self.batch, 784))), ng.tensor(np.zeros((self.batch, 784)))
    counter = torch.zeros(1, dtype=torch.int32, device='cuda')
    inner = self.make_step((REAL, NG, ND))

    def step():
      ng.nn.randn(None, seed=11, out=NG, counter=counter)       # fresh noise on every replay: the counter lives on the device
      ng.nn.randn(None, seed=12, out=ND, counter=counter)
      return inner()
    return step


def make_workload(ng, name, batch, rank):
  return (GanWorkload if name == 'gan784' else LeNetWorkload)(ng, name, batch, rank)


class Harness:
  """Timing utilities shared by every workload of one process."""

  def __init__(self, torch, world, rank):
    self.torch, self.world, self.rank = torch, world, rank
    self.flush_buf = torch.empty(256 << 20, device='cuda', dtype=torch.uint8)   # 256 MiB > 126 MB L2

  def barrier(self):
    t = self.torch
    t.cuda.synchronize()
    if self.world > 1:
      t.distributed.barrier()
    t.cuda.synchronize()

  def max_over_ranks(self, v):
    if self.world == 1:
      return v
    t = self.torch
    x = t.tensor([v], device='cuda', dtype=t.float64)
    t.distributed.all_reduce(x, op=t.distributed.ReduceOp.MAX)
    return float(x.item())

  def timed(self, fn, steps):
    """sum of per-step CUDA-event times (ms), L2 flushed before every step"""
    t = self.torch
    evs = []
    self.barrier()
    for _ in range(steps):
      self.flush_buf.zero_()
      e0, e1 = t.cuda.Event(enable_timing=True), t.cuda.Event(enable_timing=True)
      e0.record()
      fn()
      e1.record()
      evs.append((e0, e1))
    self.barrier()
    return sum(a.elapsed_time(b) for a, b in evs)


def measure_workload(h, ng, B, name, batch, args, pk, full=True, clocks_index=None):
  """Graph-replayed `value` (+ clocks), and with `full`: eager, e2e, per-call roofline / breakdown.  Every rank runs
  everything (the steps contain the gradient exchange); rank 0 keeps the numbers."""
  torch = h.torch
  from neograd_b200 import distributed as dist
  from neograd_b200.capture import CapturedStep
  from neograd_b200.device import DeviceArray
  world, rank = h.world, h.rank
  wl = make_workload(ng, name, batch, rank)
  if world > 1:
    dist.broadcast_params(wl.params())
  wl.setup_optim()
  tensors = [ng.tensor(a) for a in wl.host_inputs]
  step = wl.make_step(tensors)
  warm = max(args.warmup, 3)

  # ---- eager warm-up, then capture the whole step in a CUDA graph
  for _ in range(warm):
    step()
  torch.cuda.synchronize()
  n0 = B.lib.ngb_launch_count()
  step()
  launches_eager = int(B.lib.ngb_launch_count() - n0)
  captured = CapturedStep(step, warmup=2)
  for _ in range(warm):
    captured()

  clocks = ClockSampler(clocks_index) if clocks_index is not None else None
  if clocks:
    clocks.start()
  for _ in range(64):                     # the same load while the sampler takes its first readings -- a FIXED count on every
    captured()                            # rank: a replay contains the gradient exchange, so ranks must stay in step
  graph_ms = h.max_over_ranks(h.timed(captured, args.steps))
  clk = clocks.result() if clocks else None
  total_samples = batch * world * args.steps
  out = {'value': total_samples / (graph_ms * 1e-3), 'unit': 'samples/s', 'ms_per_step': graph_ms / args.steps,
         'batch_per_gpu': batch, 'global_batch': batch * world, 'launches_per_step': {'graph': captured.launches, 'eager': launches_eager},
         'gpu_launches': captured.launches * args.steps}
  if clk is not None:
    out['clocks'] = clk
  if not full:
    return out
  eager_ms = h.max_over_ranks(h.timed(step, args.steps))
  out['eager'] = {'value': total_samples / (eager_ms * 1e-3), 'unit': 'samples/s', 'ms_per_step': eager_ms / args.steps}

  # ---- e2e: host inputs -> pinned H2D copy -> on-device preparation -> step -> D2H loss, wall clock, through the library's
  # public input API (nn.DeviceLoader): the batch crosses PCIe in its source format and is expanded by device kernels; the
  # copy of batch i+1 (copy stream) overlaps step i.  Double buffered at the INPUT level: one captured step per loader slot
  # over the same model / optimizer; every step copies its full batch from host memory and reads its loss(es) back; the
  # loss of step i is consumed on the host before step i+2 is issued.
  loader = wl.e2e_loader(8, rank)
  for _ in range(2):                # both slots hold a real batch before the steps over them are captured
    loader.release(loader.next()[2])
  torch.cuda.synchronize()

  # cast + normalise + one-hot of a batch run on the loader's copy stream right behind its H2D copy (prepare_ahead), i.e. beside
  # the previous step's kernels: 1-2 launches per step on top of the captured step's
  steps2 = [CapturedStep(wl.make_e2e_step([t_ for t_ in loader.slots[i] if t_ is not None]), warmup=2) for i in range(2)]
  n0 = B.lib.ngb_launch_count()
  loader.prepare(0)
  e2e_launches = steps2[0].launches + int(B.lib.ngb_launch_count() - n0)
  n_loss = len(steps2[0].result)
  loss_host = [torch.empty(n_loss, dtype=torch.float32).pin_memory() for _ in range(2)]
  loss_dev = [torch.empty(n_loss, device='cuda', dtype=torch.float32) for _ in range(2)]
  step_done = [torch.cuda.Event() for _ in range(2)]
  loss_read = [torch.cuda.Event() for _ in range(2)]
  loss_stream = torch.cuda.Stream()
  main = torch.cuda.current_stream()

  def e2e_run(nsteps):
    loader.reset()
    for i in range(nsteps):
      item = loader.next(prepare=False)
      if item is None:
        loader.reset()
        item = loader.next(prepare=False)
      slot = item[2]
      losses = steps2[slot]()
      loader.release(slot)
      b_ = i & 1
      # the loss read-back rides on its own stream behind the step (each of the two captured steps owns its loss buffers, so
      # the copy of step i is long finished when step i+2 overwrites them): the next step does not queue behind a D2H copy
      step_done[b_].record(main)
      loss_stream.wait_event(step_done[b_])
      with torch.cuda.stream(loss_stream):
        if n_loss == 1:
          loss_host[b_].copy_(losses[0].data.t.view(1), non_blocking=True)
        else:
          for j, l in enumerate(losses):
            loss_dev[b_][j:j + 1].copy_(l.data.t.view(1), non_blocking=True)
          loss_host[b_].copy_(loss_dev[b_], non_blocking=True)
        loss_read[b_].record(loss_stream)
      if i >= 1:
        loss_read[(i - 1) & 1].synchronize()              # loss of the previous step is now on the host
        float(loss_host[(i - 1) & 1][0])
    loss_stream.synchronize()
    main.synchronize()
    return [float(v) for v in loss_host[(nsteps - 1) & 1]]

  # host-side hygiene any training script may do: long-lived objects leave the cyclic GC's working set, so a gen-2
  # collection (several ms with torch imported) cannot land inside a sub-millisecond step
  import gc
  gc.collect()
  gc.freeze()
  # untimed warm-up until steady state: after the host-bound eager phase above the copy path / clocks need a few hundred
  # steps to come back to speed (chunks of K steps are repeated until two successive chunks agree within 3 %)
  prev_t, agree = None, 0
  for _ in range(30):
    h.barrier()
    t0 = time.perf_counter()
    e2e_run(max(args.steps, 20))
    h.barrier()
    dt = time.perf_counter() - t0
    agree = agree + 1 if (prev_t is not None and abs(dt - prev_t) <= 0.03 * prev_t) else 0
    prev_t = dt
    if world > 1:      # every rank must take the same number of chunks
      flag = torch.tensor([1.0 if agree >= 2 else 0.0], device='cuda')
      torch.distributed.all_reduce(flag, op=torch.distributed.ReduceOp.MIN)
      if flag.item() > 0:
        break
    elif agree >= 2:
      break
  e2e_samples = []
  last_loss = None
  for _ in range(E2E_REPEATS):      # the host of a shared box jitters: median of a few K-step measurements
    h.barrier()
    t0 = time.perf_counter()
    last_loss = e2e_run(args.steps)
    h.barrier()
    e2e_samples.append(h.max_over_ranks(time.perf_counter() - t0))
  e2e_s = float(np.median(e2e_samples))

  # the copy this pipeline hides behind the step (measured warm, right after the pipelined loop)
  h2d_ev = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
  src_x = loader._x_all[:loader.batch_size] if loader._whole else loader._pin_x[0]
  with torch.cuda.stream(loader._copy_stream):
    for _ in range(3):
      loader._raw_x[0].copy_(src_x, non_blocking=True)
    h2d_ev[0].record(loader._copy_stream)
    for _ in range(10):
      loader._raw_x[0].copy_(src_x, non_blocking=True)
    h2d_ev[1].record(loader._copy_stream)
  torch.cuda.synchronize()
  h2d_ms = h2d_ev[0].elapsed_time(h2d_ev[1]) / 10
  out['e2e'] = {'value': total_samples / e2e_s, 'unit': 'samples/s', 'h2d_bytes_per_step': int(loader.bytes_per_batch),
                'd2h_bytes_per_step': 4 * n_loss, 'ms_per_step': e2e_s / args.steps * 1e3, 'last_loss': last_loss,
                'input': wl.E2E_INPUT, 'gpu_launches_per_step': e2e_launches,
                'method': f'median of {E2E_REPEATS} wall-clock measurements of {args.steps} steps each',
                'h2d_ms_per_batch': h2d_ms,   # PCIe time of one batch alone: the floor of the pipelined end-to-end step
                'ms_per_step_all': [t / args.steps * 1e3 for t in e2e_samples]}

  # ---- per-call breakdown of the eager step AS DISPATCHED in the timed region (backward lanes on): CUDA events around
  # every C-ABI call on its launching stream.  A ~10 ms spin kernel goes first, so the step's launches queue up behind
  # it and run back to back: an event pair brackets the kernel(s) of the call, not the host's launch latency.
  reps = 5
  with B.profile_calls() as prof:
    for _ in range(reps):
      torch.cuda._sleep(20_000_000)
      step()
  h.barrier()
  if rank == 0:
    summ = prof.summary()
    total_ms = sum(d['ms'] for d in summ.values())
    out['kernel_breakdown'] = {k: {'calls_per_step': d['calls'] // reps, 'ms_per_step': d['ms'] / reps, 'share': d['ms'] / total_ms}
                               for k, d in summ.items()}
    # dominant launch site: the (entry point, shape arguments) group with the most device time in the step
    groups = {}
    for nm, d in summ.items():
      for a, ms in d['items']:
        key = (nm, tuple(v for v in a if isinstance(v, int)))
        g = groups.setdefault(key, {'ms': 0.0, 'n': 0, 'args': a})
        g['ms'] += ms
        g['n'] += 1
    (top, gkey), g = max(groups.items(), key=lambda kv: kv[1]['ms'])
    roofline = _roof(top, g['args'], g['ms'] / g['n'] * 1e-3, pk)
    traffic = None
    try:   # DRAM bytes per launch of this kernel at this shape from the committed ncu --set full capture, if there is one
      with open(os.path.join(ROOT, 'profiles', 'ncu_traffic.json')) as f:
        traffic = json.load(f).get(top + ':' + ','.join(str(v) for v in gkey))
    except Exception:
      pass
    roofline.update(traffic=traffic, shape_args=list(gkey), share_of_step_kernel_time=g['ms'] / total_ms,
                    dispatch='as timed: backward lanes on, events on the launching stream')
    # dominant entry-point GROUP (all shapes of one entry point): algorithmic work summed over its calls / their summed time
    gname, gd = max(summ.items(), key=lambda kv: kv[1]['ms'])
    gby = sum(call_work(gname, a)[0] for a, _ in gd['items'])
    gfl = sum(call_work(gname, a)[1] for a, _ in gd['items'])
    gs = gd['ms'] * 1e-3
    ridge = pk['tf32_tflops'] * 1e12 / (pk['hbm_gbs'] * 1e9)
    if gfl / max(gby, 1) > ridge:
      grp = {'entry': gname, 'bound': 'tensor', 'achieved': gfl / gs / 1e12, 'peak': pk['tf32_tflops'], 'unit': 'TFLOP/s'}
    else:
      grp = {'entry': gname, 'bound': 'hbm', 'achieved': gby / gs / 1e9, 'peak': pk['hbm_gbs'], 'unit': 'GB/s'}
    grp.update(frac=grp['achieved'] / grp['peak'], calls_per_step=gd['calls'] // reps, share_of_step_kernel_time=gd['ms'] / total_ms)
    roofline['group'] = grp
    out['roofline'] = roofline
  # the same launch site timed ALONE (backward lanes off: nothing co-runs with it), for comparison with the as-dispatched
  # figure above -- a kernel that shares the SMs with another lane's kernel is charged its co-run duration there
  prev_lanes = ng.set_backward_lanes(0)
  with B.profile_calls() as prof2:
    for _ in range(reps):
      torch.cuda._sleep(20_000_000)
      step()
  h.barrier()
  ng.set_backward_lanes(prev_lanes)
  if rank == 0:
    ms_alone = [ms for nm, d in prof2.summary().items() if nm == top for a, ms in d['items']
                if tuple(v for v in a if isinstance(v, int)) == gkey]
    if ms_alone:
      alone = _roof(top, g['args'], float(np.mean(ms_alone)) * 1e-3, pk)
      out['roofline']['alone'] = {k: alone[k] for k in ('achieved', 'frac', 'mean_launch_us')}
      out['roofline']['alone']['dispatch'] = 'backward lanes off: the kernel has the GPU to itself'
  return out


# --------------------------------------------------------------------------------------------- ops block (config 5)
def run_ops(torch, B, pk, budget_s=45.0):
  """BASELINE config 5 inside the bench line: Dot fwd / dgrad / wgrad at the three GAN shapes and 4096^3, Conv3D 3x3 p1
  fprop / dgrad / wgrad at N=256, C=O in {64,128,256}, H=W=32, and the bandwidth kernels (MaxPool 2x2, conv bias
  gradient, bias-gradient unbroadcast, Adam).  2 warm-up + 5 timed launches each, CUDA events, 256 MiB L2 flush between
  launches; `frac` = achieved / peak of the bound that applies (TF32 tensor pipe or HBM; SURVEY 8d work formulas)."""
  P = lambda t: ctypes.c_void_p(t.data_ptr())
  S = lambda: ctypes.c_void_p(B.stream())
  flush = torch.empty(256 << 20, device='cuda', dtype=torch.uint8)
  t_begin = time.perf_counter()
  out = []

  def timeit(fn, reps=5):
    for _ in range(2):
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

  def emit(name, shape, by, fl, mean, best, bound):
    rec = {'op': name, 'shape': shape, 'us': mean * 1e6, 'us_best': best * 1e6, 'bound': bound}
    if bound == 'tensor':
      rec.update(tflops=fl / mean / 1e12, frac=fl / mean / 1e12 / pk['tf32_tflops'])
      hb = by / mean / 1e9 / pk['hbm_gbs']
      if hb > rec['frac']:        # skinny shapes sit on the HBM roofline: report the bound that actually applies
        rec.update(bound='hbm', gbs=by / mean / 1e9, frac=hb)
    else:
      rec.update(gbs=by / mean / 1e9, frac=by / mean / 1e9 / pk['hbm_gbs'])
    out.append(rec)

  def over():
    return time.perf_counter() - t_begin > budget_s

  rnd = lambda *s: torch.randn(*s, device='cuda', dtype=torch.float32)
  for M, K, N in ((8192, 784, 512), (8192, 512, 256), (8192, 256, 784), (4096, 4096, 4096)):
    if over():
      break
    A, Bm, U = rnd(M, K), rnd(K, N), rnd(M, N)
    C, dA, dB = torch.empty(M, N, device='cuda'), torch.empty(M, K, device='cuda'), torch.empty(K, N, device='cuda')
    by, fl = 4 * (M * K + K * N + M * N), 2 * M * N * K
    for nm, call in (('dot_fwd', lambda: B.lib.ngb_gemm(0, 0, 0, M, N, K, P(A), K, P(Bm), N, P(C), N, 0, S())),
                     ('dot_dgrad', lambda: B.lib.ngb_gemm(0, 0, 1, M, K, N, P(U), N, P(Bm), N, P(dA), K, 0, S())),
                     ('dot_wgrad', lambda: B.lib.ngb_gemm(0, 1, 0, K, N, M, P(A), K, P(U), N, P(dB), N, 0, S()))):
      if call() != 0:
        out.append({'op': nm, 'shape': [M, K, N], 'error': B.lib.ngb_last_error().decode()})
        continue
      m, best = timeit(call)
      emit(nm, [M, K, N], by, fl, m, best, 'tensor')
    del A, Bm, U, C, dA, dB
  for Nn, C_, H in ((256, 64, 32), (256, 128, 32), (256, 256, 32)):
    if over():
      break
    O, W, KH, KW, pad, stride = C_, H, 3, 3, 1, 1
    x, w, b = rnd(Nn, C_, H, W), rnd(O, C_, KH, KW) / float(np.sqrt(9 * C_)), torch.zeros(O, device='cuda')
    y, ug = torch.empty(Nn, O, H, W, device='cuda'), rnd(Nn, O, H, W)
    dx, dw, db = torch.empty_like(x), torch.empty_like(w), torch.empty(O, device='cuda')
    by, fl = 4 * (x.numel() + w.numel() + y.numel()), 2 * Nn * O * C_ * KH * KW * H * W
    g = (Nn, C_, H, W, O, KH, KW, pad, stride)
    for nm, call in (('conv_fprop', lambda: B.lib.ngb_conv_fprop(P(x), P(w), P(b), P(y), *g, S())),
                     ('conv_dgrad', lambda: B.lib.ngb_conv_dgrad(P(ug), P(w), P(dx), *g, 0, S())),
                     ('conv_wgrad', lambda: B.lib.ngb_conv_wgrad(P(ug), P(x), P(dw), *g, 0, S()))):
      if call() != 0:
        out.append({'op': nm, 'shape': list(g), 'error': B.lib.ngb_last_error().decode()})
        continue
      m, best = timeit(call)
      emit(nm, list(g), by, fl, m, best, 'tensor')
    call = lambda: B.lib.ngb_conv_bgrad(P(ug), P(db), Nn, O, H * W, 0, S())
    call()
    m, best = timeit(call)
    emit('conv_bgrad', [Nn, O, H * W], 4 * ug.numel() + 4 * O, ug.numel(), m, best, 'hbm')
    if C_ == 256:
      outer = Nn * O
      yp = torch.empty(outer, H // 2, W // 2, device='cuda')
      idx = torch.empty(outer, H // 2, W // 2, device='cuda', dtype=torch.uint8)
      byp = 4 * ug.numel() + 5 * yp.numel()
      call = lambda: B.lib.ngb_maxpool_fwd(P(ug), P(yp), P(idx), outer, H, W, 2, 2, 0, 2, S())
      call()
      m, best = timeit(call)
      emit('maxpool_fwd', [outer, H, W], byp, ug.numel(), m, best, 'hbm')
      call = lambda: B.lib.ngb_maxpool_bwd(P(yp), P(idx), P(y), outer, H, W, 2, 2, 0, 2, 0, S())
      call()
      m, best = timeit(call)
      emit('maxpool_bwd', [outer, H, W], byp, ug.numel(), m, best, 'hbm')
      del yp, idx
    del x, w, y, ug, dx, dw
  if not over():
    M, Nn = 1 << 20, 256
    A, bias, gb = rnd(M, Nn), rnd(1, Nn), torch.empty(1, Nn, device='cuda')
    sa, sb = (ctypes.c_int64 * 2)(M, Nn), (ctypes.c_int64 * 2)(1, Nn)
    call = lambda: B.lib.ngb_ew_binary_bwd(B.ADD, 1, P(A), P(A), 0.0, sa, 2, P(bias), 0.0, sb, 2, P(gb), 0, S())
    call()
    m, best = timeit(call)
    emit('bias_grad_unbroadcast', [M, Nn], 4 * (M * Nn + Nn), M * Nn, m, best, 'hbm')
    o2 = torch.empty(M, Nn, device='cuda')
    call = lambda: B.lib.ngb_ew_binary_fwd(B.ADD, P(A), 0.0, sa, 2, P(bias), 0.0, sb, 2, P(o2), S())
    call()
    m, best = timeit(call)
    emit('bias_add', [M, Nn], 4 * (2 * M * Nn + Nn), M * Nn, m, best, 'hbm')
    del A, o2
    na = 1 << 27
    p, g_, mm, vv = rnd(na), rnd(na), torch.zeros(na, device='cuda'), torch.zeros(na, device='cuda')
    call = lambda: B.lib.ngb_adam_step(P(p), P(g_), P(mm), P(vv), na, 1e-3, 0.9, 0.999, 1e-8, 1, 1.0, S())
    call()
    m, best = timeit(call)
    emit('adam_step', [na], 28 * na, 12 * na, m, best, 'hbm')
    del p, g_, mm, vv
  del flush
  torch.cuda.empty_cache()
  return out


# --------------------------------------------------------------------------------------------- device arm
def run_device(args):
  # stdout carries exactly ONE JSON line: everything else that writes to fd 1 (NCCL prints a version banner there at
  # communicator creation) goes to stderr; the result is written to the saved descriptor at the end.
  sys.stdout.flush()
  json_fd = os.dup(1)
  os.dup2(2, 1)
  import torch
  import neograd_b200 as ng
  from neograd_b200 import _backend as B
  from neograd_b200 import distributed as dist

  world = int(os.environ.get('WORLD_SIZE', '1'))
  rank = int(os.environ.get('RANK', '0'))
  local = int(os.environ.get('LOCAL_RANK', '0'))
  torch.cuda.set_device(local)
  if world > 1:
    dist.init('nccl')
  B.ensure_runtime()
  pk = peaks()
  if not args.no_extra:
    measure_tf32_peak(pk)
  h = Harness(torch, world, rank)
  name = args.workload

  main = measure_workload(h, ng, B, name, BATCH[name], args, pk, full=True, clocks_index=local)

  extra = {}
  if not args.no_extra:
    # the other BASELINE training configs (graph-replayed value + launches only) ...
    for other in ('gan784', 'lenet_b'):
      if other != name:
        try:
          extra[other] = measure_workload(h, ng, B, other, BATCH[other], args, pk, full=False)
          extra[other]['scaling'] = 'weak'
        except Exception as e:   # noqa: BLE001
          extra[other] = {'error': f'{type(e).__name__}: {e}'}
    # ... and the strong-scaling figure of the headline config: global batch fixed, split over the N GPUs
    if world > 1 and BATCH[name] % world == 0:
      try:
        s = measure_workload(h, ng, B, name, BATCH[name] // world, args, pk, full=False)
        s['scaling'] = 'strong'
        extra[name + '_strong'] = s
      except Exception as e:   # noqa: BLE001
        extra[name + '_strong'] = {'error': f'{type(e).__name__}: {e}'}

  ops = None
  if rank == 0 and world == 1 and not args.no_extra:
    try:
      ops = run_ops(torch, B, pk)
    except Exception as e:   # noqa: BLE001
      ops = [{'error': f'{type(e).__name__}: {e}'}]

  cpu = None
  if rank == 0 and world == 1 and not args.no_cpu_baseline:
    cores = len(os.sched_getaffinity(0))
    b = min(1024, BATCH[name])
    r = reference_steps(name, b, 3, 1, 25.0)
    if r is not None:
      per = float(np.median(r['times']))
      cpu = {'value': b / per, 'unit': 'samples/s', 'cores': cores, 'kind': 'reference',
             'sample': f'{len(r["times"])} steps (1 warm-up) of batch {b} (of {BATCH[name]}) of {name}: the unmodified reference package, float64 NumPy/OpenBLAS default threads'}
    else:
      b = 256
      ps = port_steps(name, b, 5, 1, 25.0)['times']
      cpu = {'value': b / float(np.median(ps)), 'unit': 'samples/s', 'cores': cores, 'kind': 'port',
             'sample': f'{len(ps)} steps of batch {b} (of {BATCH[name]}) of {name}: oracle port, float64 NumPy/OpenBLAS default threads'}

  if rank == 0:
    out = {
      'metric': METRIC[name], 'value': main['value'], 'unit': 'samples/s', 'n_gpus': world,
      'steps': args.steps, 'warmup': max(args.warmup, 3), 'ms_per_step': main['ms_per_step'], 'higher_is_better': True,
      'scaling': 'weak', 'vs_baseline': None, 'dtype': 'f32 (TF32 tensor-core operands where a Dot / Conv qualifies)',
      'data': 'synthetic',
      'config': {'workload': DESCR[name], 'batch_per_gpu': BATCH[name], 'global_batch': BATCH[name] * world,
                 'parallelism': f'dp{world}', 'l2': 'flushed (256 MiB memset) between timed steps',
                 'step_dispatch': 'CUDA-graph replay of the captured step'},
      'clocks': main.get('clocks'),
      'e2e': main['e2e'],
      'gpu_launches': main['gpu_launches'],   # timed region of `value`; every e2e step replays the same graph
      'launches_per_step': main['launches_per_step'],
      'eager': main['eager'],
      'roofline': main.get('roofline'), 'cpu_baseline': cpu, 'kernel_breakdown': main.get('kernel_breakdown'),
      'workloads': extra, 'ops': ops, 'peaks': pk,
    }
    sys.stdout.flush()
    os.write(json_fd, (json.dumps(out) + '\n').encode())
  if world > 1:
    # NCCL teardown after a graph-captured collective was seen to hang on exit: synchronise, then leave without running
    # the communicator destructors (the process is ending anyway).
    h.barrier()
    sys.stdout.flush()
    sys.stderr.flush()
    os._exit(0)


def main():
  ap = argparse.ArgumentParser()
  ap.add_argument('--gpus', type=int, default=1)
  ap.add_argument('--steps', type=int, default=20)
  ap.add_argument('--warmup', type=int, default=3)
  ap.add_argument('--impl', default='ours', choices=['ours', 'reference'])
  ap.add_argument('--workload', default='lenet_b', choices=WORKLOADS)
  ap.add_argument('--no-cpu-baseline', action='store_true')
  ap.add_argument('--no-extra', action='store_true', help='skip the secondary workloads, the ops block and the TF32 peak')
  args = ap.parse_args()
  if args.impl == 'reference':
    run_reference(args)
  else:
    run_device(args)


if __name__ == '__main__':
  main()
