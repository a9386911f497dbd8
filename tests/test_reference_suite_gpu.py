"""The reference's own test-suite, re-stated against the device package (reference: tests/test_basics.py,
test_layers.py, test_activations.py, test_loss.py, tests/_setup.py:22-42 and the save/load script tests/test.py).

Same technique as the reference -- finite-difference gradient checking through `fn_grad_check`, same operand shapes
and geometries -- with two changes: inputs are seeded (the reference's are not), and the probe width / tolerance are
the device's calibrated ones (eps 1e-3, dist < 1e-2) instead of 1e-7, which is below fp32 resolution."""
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
TOL = 1e-2


@pytest.fixture(scope='module')
def ng():
  import torch
  if not torch.cuda.is_available():
    pytest.skip('needs a CUDA device')
  import neograd_b200
  return neograd_b200


def execute(ng, fn, inputs, params=None, targets=None, loss_fn=None, tolerance=TOL, **kwargs):
  """tests/_setup.py:22-42: operands -> requires_grad tensors, fn_grad_check, assert on the distance."""
  tens = [ng.tensor(np.asarray(a, dtype=float), requires_grad=True) for a in inputs]
  params = tens if params is None else params
  d = ng.fn_grad_check(fn, tens, params, targets=targets, loss_fn=loss_fn, print_vals=False, **kwargs)
  assert d < tolerance, d
  return d


RNG = np.random.default_rng(42)
A = lambda *s: RNG.standard_normal(s) if s else float(RNG.standard_normal())
PAIRS = [(np.array(3.0), A(3)), (np.array(2.0), A(2, 3, 4)), (A(3), A(2, 3)), (A(3, 4), A(2, 3, 4)), (A(4, 1), A(4, 5))]


@pytest.mark.parametrize('i', range(len(PAIRS)))
@pytest.mark.parametrize('op', ['add', 'sub', 'mul', 'div'])
def test_basics_binary(ng, op, i):
  """test_basics.py:7-31 -- scalar (+) vector, scalar (+) 3-D, vector (*) matrix, matrix (*) 3-D."""
  a, b = PAIRS[i]
  if op == 'div':
    b = np.abs(b) + 1.0
  execute(ng, getattr(ng, op), [a, b])
  execute(ng, getattr(ng, op), [b, a] if op != 'div' else [a, b])


def test_basics_pow_exp_log_dot_sum(ng):
  execute(ng, ng.pow, [np.abs(A(3, 4)) + 0.5, np.abs(A(3, 4)) + 0.5])
  execute(ng, ng.exp, [A(2, 3) * 0.5])
  execute(ng, ng.log, [np.abs(A(2, 3)) + 1.0])
  execute(ng, ng.dot, [A(2, 2), A(2, 3)])                      # test_basics.py:41
  for axis in (None, 0, 1):                                    # test_basics.py:61-63
    execute(ng, lambda t, axis=axis: ng.sum(t, axis=axis), [A(3, 4)])
  execute(ng, ng.transpose, [A(3, 4)])
  execute(ng, ng.flatten, [A(3, 4)])
  execute(ng, lambda t: ng.reshape(t, (4, 3)), [A(3, 4)])


@pytest.mark.parametrize('ks,pad,stride', [((3, 3), 0, 1), ((2, 2), 2, 1), ((4, 4), 0, 2), ((3, 3), 1, 1)])
def test_layers_conv2d_conv3d(ng, ks, pad, stride):
  """test_layers.py:35-56 -- Conv2D on (2,12,15), Conv3D on (2,2,12,13) with 2 -> 5 channels, the four geometries."""
  np.random.seed(1)
  conv2 = ng.nn.Conv2D(ks, pad, stride)
  execute(ng, conv2, [A(2, 12, 15)], params=conv2.parameters())
  conv3 = ng.nn.Conv3D(2, 5, ks, pad, stride)
  execute(ng, conv3, [A(2, 2, 12, 13)], params=conv3.parameters())


def test_layers_maxpool_linear_dropout(ng):
  np.random.seed(2)
  x3 = ng.tensor(A(2, 12, 15), requires_grad=True)
  d = ng.fn_grad_check(ng.nn.MaxPool2D((3, 3), stride=3), [x3], [x3], print_vals=False)       # test_layers.py:61-63
  assert d < TOL, d
  x4 = ng.tensor(A(2, 3, 12, 15), requires_grad=True)
  d = ng.fn_grad_check(ng.nn.MaxPool3D((3, 3), stride=3), [x4], [x4], print_vals=False)       # test_layers.py:68-70
  assert d < TOL, d
  lin = ng.nn.Linear(5, 3)
  execute(ng, lin, [A(2, 5)], params=lin.parameters())                                          # test_layers.py:28-30
  drop = ng.nn.Dropout(0.5)
  execute(ng, drop, [A(4, 6)])                                                                  # identity


def test_activations(ng):
  """test_activations.py: ReLU / Sigmoid / Tanh / Softmax(axis 0, 1, 2) / LeakyReLU."""
  x = A(3, 4)
  x[np.abs(x) < 0.05] = 0.3   # keep the probes away from the ReLU kink
  for act in (ng.nn.ReLU(), ng.nn.Sigmoid(), ng.nn.Tanh(), ng.nn.LeakyReLU()):
    execute(ng, act, [x])
  for axis in (0, 1, 2):
    execute(ng, ng.nn.Softmax(axis), [A(2, 3, 4)])


def test_softmax_values_and_jacobian(ng):
  """Forward values against the closed form and the backward against the explicit Jacobian (activations.py:131-174)."""
  x = RNG.standard_normal((5, 7)).astype(np.float32).astype(np.float64)
  ug = RNG.standard_normal((5, 7)).astype(np.float32).astype(np.float64)
  t = ng.tensor(x, requires_grad=True)
  y = ng.nn.Softmax(1)(t)
  e = np.exp(x - x.max(1, keepdims=True))
  s = e / e.sum(1, keepdims=True)
  assert np.allclose(y.data.numpy(), s, atol=1e-6)
  y.backward(ug)
  ref = np.stack([(np.diag(si) - np.outer(si, si)) @ gi for si, gi in zip(s, ug)])
  assert np.allclose(t.grad.numpy(), ref, atol=1e-6)


def test_losses(ng):
  """test_loss.py: MSE / BCE / CE / SoftmaxCE."""
  nn = ng.nn
  out = RNG.uniform(0.1, 0.9, (6, 4))
  tgt = RNG.uniform(0.1, 0.9, (6, 4))
  T = ng.tensor(tgt)
  for loss in (nn.loss.MSE(), nn.loss.BCE(), nn.loss.CE()):
    execute(ng, lambda o: o, [out], targets=T, loss_fn=loss)
  onehot = np.eye(4)[RNG.integers(0, 4, 6)]
  execute(ng, lambda o: o, [RNG.standard_normal((6, 4))], targets=ng.tensor(onehot), loss_fn=nn.loss.SoftmaxCE(axis=1))
  with pytest.raises(AssertionError):                                # loss.py:172
    l = nn.loss.SoftmaxCE(axis=1)(ng.tensor(out, requires_grad=True), ng.tensor(onehot, requires_grad=True))
    l.backward()
  ng._NG_GRAPH.reset_graph()


def test_save_load_and_checkpoint(ng, tmp_path):
  """tests/test.py:22-65 (model.save / load round trip) and nn/checkpoint.py."""
  from neograd_b200.configs import MLPCircles
  np.random.seed(5)
  model = MLPCircles()
  X = ng.tensor(RNG.standard_normal((16, 2)))
  with model.eval():
    before = model(X).data.numpy()
  f = str(tmp_path / 'params.pkl')
  model.save(f)
  saved = [p.data.numpy().copy() for p in model.parameters()]
  for p in model.parameters():
    p.data = np.zeros(p.shape)
  model.load(f)
  assert all(np.array_equal(a, p.data.numpy()) for a, p in zip(saved, model.parameters()))
  with model.eval():
    assert np.array_equal(model(X).data.numpy(), before)
  # the file holds the reference's nested {attr: [ {weights, bias}, {}, ... ]} structure of float64 ndarrays
  import pickle
  try:
    import dill as pickler
  except ImportError:
    pickler = pickle
  blob = pickler.load(open(f, 'rb'))
  assert set(blob) == {'stack'} and blob['stack'][0]['weights'].dtype == np.float64 and blob['stack'][1] == {}
  ckpt = ng.Checkpoint(model, str(tmp_path / 'ckpts'))
  ckpt.add(epoch=1, loss=0.5)
  import json
  sessions = json.load(open(tmp_path / 'ckpts' / 'checkpoints.json'))
  fname = list(sessions[ckpt.session])[-1]
  assert sessions[ckpt.session][fname]['epoch'] == 1
  for p in model.parameters():
    p.data = np.ones(p.shape)
  meta = ckpt.load(fname + '.hkl')
  assert meta['loss'] == 0.5 and all(np.array_equal(a, p.data.numpy()) for a, p in zip(saved, model.parameters()))
  with pytest.raises(ValueError):
    ckpt.add(bad=np.zeros(2))
  # whole-model pickle (nn/utils.py:5-36): structure travels, weights are saved separately (layers/__init__.py:160-173)
  mf = str(tmp_path / 'model.pkl')
  ng.save(mf, model)
  m2 = ng.load(mf)
  assert [p.shape for p in m2.parameters()] == [p.shape for p in model.parameters()]
  with pytest.raises(TypeError):
    ng.save(mf, object())


def test_optimizers_against_oracle(ng):
  """GD / Momentum / RMSProp / Adam through the public API vs the float64 oracle on the same gradients."""
  from oracle import ops
  for name in ('GD', 'Momentum', 'RMSProp', 'Adam'):
    lin = ng.nn.Linear(7, 5)
    p0 = [p.data.numpy().copy() for p in lin.parameters()]
    opt = getattr(ng.nn.optim, name)(lin.parameters(), 0.01)
    X, Y = ng.tensor(RNG.standard_normal((9, 7))), ng.tensor(RNG.standard_normal((9, 5)))
    ref, s1, s2 = [p.copy() for p in p0], [0.0, 0.0], [0.0, 0.0]
    for step in range(3):
      opt.zero_grad()
      loss = ng.nn.loss.MSE()(lin(X), Y)
      loss.backward()
      grads = [p.grad.numpy().copy() for p in lin.parameters()]
      opt.step()
      for i in range(2):
        if name == 'GD':
          ref[i] = ops.sgd_step(ref[i], grads[i], 0.01)
        elif name == 'Momentum':
          ref[i], s1[i] = ops.momentum_step(ref[i], grads[i], s1[i], 0.01)
        elif name == 'RMSProp':
          ref[i], s1[i] = ops.rmsprop_step(ref[i], grads[i], s1[i], 0.01)
        else:
          ref[i], s1[i], s2[i] = ops.adam_step(ref[i], grads[i], s1[i], s2[i], step + 1, 0.01)
    for i, p in enumerate(lin.parameters()):
      assert ops.dist(p.data.numpy(), ref[i]) < 1e-5, (name, i)
