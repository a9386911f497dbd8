"""GPU tests of the optimizer arena's host logic added in round 2: exact resume from a checkpoint that carries optimizer
state (reference: nn/checkpoint.py:100-148, nn/optim.py:181-197), `Adam.iter` as a live property, optimizers
over a sub-list of another optimizer's parameters (allowed by the reference), and release of arenas that nobody references."""
import gc
import json

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


@pytest.fixture(scope='module')
def ng():
  import torch
  if not torch.cuda.is_available():
    pytest.skip('needs a CUDA device')
  import neograd_b200
  return neograd_b200


def _setup(ng, seed=0):
  from neograd_b200.configs import MLPCircles
  np.random.seed(seed)
  model = MLPCircles()
  rng = np.random.default_rng(1)
  X, y = ng.tensor(rng.standard_normal((128, 2))), ng.tensor(rng.integers(0, 2, (128, 1)).astype(float))
  return model, X, y


def test_checkpoint_with_optimizer_state_resumes_exactly(ng, tmp_path):
  from neograd_b200.configs import train_step
  loss_fn = ng.nn.loss.BCE()
  model, X, y = _setup(ng)
  optim = ng.nn.optim.Adam(model.parameters(), 0.05)
  for _ in range(3):
    train_step(model, optim, loss_fn, X, y)
  ckpt = ng.Checkpoint(model, str(tmp_path / 'ck'), optimizer=optim)
  ckpt.add(step=3)
  fname = list(json.load(open(tmp_path / 'ck' / 'checkpoints.json'))[ckpt.session])[-1]
  assert (tmp_path / 'ck' / ckpt.session / (fname + '.optim')).exists()
  want = [float(train_step(model, optim, loss_fn, X, y).data) for _ in range(2)]
  want_params = [p.data.numpy() for p in model.parameters()]
  assert optim.iter == 5
  # a fresh process would rebuild model + optimizer and load: parameters, moments and the iteration count come back
  model2, _, _ = _setup(ng, seed=7)
  optim2 = ng.nn.optim.Adam(model2.parameters(), 0.01)
  ck2 = ng.Checkpoint(model2, str(tmp_path / 'ck'), optimizer=optim2)
  ck2.specify_session(ckpt.session)
  meta = ck2.load(fname)
  assert meta['step'] == 3 and optim2.iter == 3 and optim2.lr == 0.05
  got = [float(train_step(model2, optim2, loss_fn, X, y).data) for _ in range(2)]
  assert got == want                                                 # bit-identical continuation
  for a, p in zip(want_params, model2.parameters()):
    assert np.array_equal(a, p.data.numpy())
  # without the keyword the class writes exactly the reference's files
  ck3 = ng.Checkpoint(model2, str(tmp_path / 'plain'))
  ck3.add(step=1)
  assert not list((tmp_path / 'plain').rglob('*.optim'))


def test_adam_iter_is_live_and_assignable(ng):
  from neograd_b200.capture import CapturedStep
  from neograd_b200.configs import train_step
  loss_fn = ng.nn.loss.BCE()
  model, X, y = _setup(ng)
  optim = ng.nn.optim.Adam(model.parameters(), 0.05)
  step = CapturedStep(lambda: train_step(model, optim, loss_fn, X, y), warmup=2)
  before = optim.iter
  for _ in range(4):
    step()
  assert optim.iter == before + 4                                    # replays advance the device counter; the property reads it
  optim.iter = 100                                                   # resuming: takes effect on the next step (optim.py:169-178)
  p0 = model.parameters()[0].data.numpy().copy()
  g = None
  loss = loss_fn(model(X), y)
  optim.zero_grad()
  loss.backward()
  g = model.parameters()[0].grad.numpy().copy()
  m0, v0 = model.parameters()[0].momentum_grad.numpy().copy(), model.parameters()[0].rms_grad.numpy().copy()
  optim.step()
  assert optim.iter == 101
  m1, v1 = 0.9 * m0 + 0.1 * g, 0.999 * v0 + 0.001 * g * g
  want = p0 - 0.05 * (m1 / (1 - 0.9 ** 101)) / (np.sqrt(v1 / (1 - 0.999 ** 101)) + 1e-8)
  assert np.allclose(model.parameters()[0].data.numpy(), want, rtol=1e-5, atol=1e-6)
  optim.reset_iter()
  assert optim.iter == 0


def test_sub_list_optimizer_shares_the_arena(ng):
  """optim.py:5-33: nothing in the reference stops a second optimizer over part of the parameters."""
  loss_fn = ng.nn.loss.BCE()
  model, X, y = _setup(ng)
  params = model.parameters()
  adam = ng.nn.optim.Adam(params, 0.05)
  sgd = ng.nn.optim.GD(params[:2], 0.1)                              # first Linear only
  assert sgd.arena is adam.arena
  adam.zero_grad()
  loss_fn(model(X), y).backward()
  before = [p.data.numpy().copy() for p in params]
  grads = [p.grad.numpy().copy() for p in params]
  sgd.step()
  after = [p.data.numpy() for p in params]
  for i in range(2):
    assert np.allclose(after[i], before[i] - 0.1 * grads[i], rtol=1e-6, atol=1e-7)
  for i in range(2, len(params)):
    assert np.array_equal(after[i], before[i])                       # the sub-list optimizer leaves the rest alone
  other, _, _ = _setup(ng, seed=3)
  with pytest.raises(ValueError):
    ng.nn.optim.GD([params[0], other.parameters()[0]], 0.1)          # parameters of two arenas cannot be mixed


def test_arenas_are_released(ng):
  import torch
  from neograd_b200.nn import optim as O
  gc.collect()
  n0 = len(O._ARENAS)
  for _ in range(3):
    model, X, y = _setup(ng)
    opt = ng.nn.optim.Adam(model.parameters(), 0.05)
    assert len(O._ARENAS) == n0 + 1
    del model, opt
    gc.collect()
    assert len(O._ARENAS) == n0
