"""CPU tests of bench.py's host logic: the reference arm (the unmodified reference driven through the user-level model /
loop definitions of neograd_b200/workload_defs.py, loaded by path) against the oracle, the algorithmic-work formulas of the
new entry points, and the GEMM heuristics' documentation constants.  Skips the reference part where neither
/root/reference nor baseline/_ref exists."""
import ctypes
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
from oracle import models  # noqa: E402


def test_reference_arm_runs_the_unmodified_reference_and_matches_the_oracle():
  ref, root = bench.load_reference()
  if ref is None:
    pytest.skip('no reference tree here')
  defs = bench._workload_defs()
  assert 'neograd_b200' not in getattr(defs, '__name__', '') and not hasattr(defs, 'ng')
  classes = defs.model_classes(ref.nn)
  from neograd_ref.nn import optim as roptim, loss as rloss
  # LeNet-B, batch 8: one README.md:109-115 step through the reference == the oracle's step
  x, y = defs.lenet_data('b', 8, 0)
  model = classes['LeNetB']()
  init = defs.synthetic_params([p.shape for p in model.parameters()], 1, 0.1)
  defs.set_params(model, init)
  loss = defs.train_step(model, roptim.Adam(model.parameters(), 1e-3), rloss.SoftmaxCE(axis=1), ref.tensor(x.astype(np.float64)),
                         ref.tensor(y.astype(np.float64)))
  net = models.lenet_b(np.random.default_rng(0))
  net.set_params(init)
  want = models.train_step(net, models.Adam(net, 1e-3), 'softmax_ce', x.astype(np.float64), y.astype(np.float64))
  assert abs(float(loss.data) - want) < 1e-10 * abs(want)
  for p, r in zip(model.parameters(), net.params()):
    assert np.allclose(np.asarray(p.data, dtype=float).reshape(r.shape), r, rtol=0, atol=1e-12)
  # GAN (notebook dims), one iteration of basic.ipynb cells 9-10 through the reference == the oracle's gan_iteration
  rng = np.random.default_rng(3)
  gan = classes['GAN'](64, 50, 25)
  ginit = [0.05 * rng.standard_normal(p.shape) for p in gan.parameters()]
  defs.set_params(gan, ginit)
  real, ng_, nd_ = (rng.standard_normal((40, 64)) for _ in range(3))
  T = ref.tensor
  lg, ld = defs.gan_iteration(gan, roptim.Adam(gan.parameters(), 1e-3), roptim.Adam(gan.parameters(), 1e-3), rloss.BCE(), T(ng_), T(nd_),
                              T(real), T(np.ones((40, 1))), T(np.zeros((40, 1))))
  gen, disc = models.gan_nets(np.random.default_rng(0), 64, 50, 25)
  gen.set_params(ginit[:6])
  disc.set_params(ginit[6:])
  og, od = models.Adam(gen, 1e-3), models.Adam(disc, 1e-3)
  wg, wd = models.gan_iteration(gen, disc, og, od, ng_, nd_, real)
  assert abs(float(lg.data) - wg) < 1e-10 * abs(wg) and abs(float(ld.data) - wd) < 1e-10 * abs(wd)


def test_timed_steps_respects_the_budget():
  calls = []
  times = bench._timed_steps(lambda: calls.append(1), steps=50, warmup=2, budget_s=0.0)
  assert len(times) == 2 and len(calls) == 4         # warm-up + the two timed steps every budget allows


def test_call_work_of_the_fused_entry_points():
  P = ctypes.c_void_p(1)
  by, fl = bench.call_work('ngb_linear_fwd', (P, P, P, P, P, 2, 0.0, 4096, 256, 120, None))
  assert fl == 2 * 4096 * 256 * 120 and by == 4 * (4096 * 256 + 256 * 120 + 120 + 2 * 4096 * 120)
  by, fl = bench.call_work('ngb_linear_fwd', (P, P, P, P, None, -1, 0.0, 4096, 256, 120, None))
  assert by == 4 * (4096 * 256 + 256 * 120 + 120 + 4096 * 120)
  by, fl = bench.call_work('ngb_linear_dgrad', (P, P, P, P, 2, 0.0, 8192, 784, 512, 0, None))
  assert fl == 2 * 8192 * 784 * 512
  by, fl = bench.call_work('ngb_linear_wgrad', (P, P, P, P, 8192, 784, 512, 0, 0, None))
  assert fl == 2 * 8192 * 784 * 512 + 8192 * 512
