"""Pins the CPU oracle (oracle/ops.py, oracle/models.py) to outputs of the unmodified reference.

Fixtures in tests/golden/ were produced by oracle/gen_golden.py from /root/reference.
Tolerance: both sides are float64; only summation order differs -> dist <= 1e-12
(dist = ||a-b||/(||a||+||b||), the reference's own metric, neograd/autograd/utils.py:153).
MaxPool values, argmax-driven gradients are compared bit-exactly.
"""
import numpy as np
import pytest

from conftest import golden_names, load_golden
from oracle import ops, models

TOL = 1e-12


def close(a, b, tol=TOL):
  a, b = np.asarray(a, float), np.asarray(b, float)
  assert a.shape == b.shape, (a.shape, b.shape)
  d = ops.dist(a, b)
  assert d <= tol, d


@pytest.mark.parametrize('name', golden_names('binary_'))
def test_binary(name):
  g = load_golden(name)
  op = str(g['op'])
  close(ops.binary_fwd(op, g['a'], g['b']), g['out'])
  close(ops.binary_bwd(op, 0, g['a'], g['b'], g['ug']), g['ga'])
  close(ops.binary_bwd(op, 1, g['a'], g['b'], g['ug']), g['gb'])


@pytest.mark.parametrize('name', golden_names('unary_'))
def test_unary(name):
  g = load_golden(name)
  op = str(g['op'])
  close(ops.unary_fwd(op, g['x'], float(g['alpha'])), g['out'])
  close(ops.unary_bwd(op, g['x'], g['ug'], float(g['alpha'])), g['gx'])


@pytest.mark.parametrize('name', golden_names('sum_'))
def test_sum(name):
  g = load_golden(name)
  axis = None if int(g['axis']) == -99 else int(g['axis'])
  close(ops.sum_fwd(g['x'], axis), g['out'])
  close(ops.sum_bwd(g['ug'], g['x'].shape, axis), g['gx'])


@pytest.mark.parametrize('name', golden_names('dot_'))
def test_dot(name):
  g = load_golden(name)
  close(ops.dot_fwd(g['a'], g['b']), g['out'])
  close(ops.dot_dgrad(g['ug'], g['b']), g['ga'])
  close(ops.dot_wgrad(g['a'], g['ug']), g['gb'])


def test_linear():
  g = load_golden('linear_5_3')
  close(ops.dot_fwd(g['x'], g['w']) + g['b'], g['out'])
  close(ops.dot_dgrad(g['ug'], g['w']), g['gx'])
  close(ops.dot_wgrad(g['x'], g['ug']), g['gw'])
  close(ops.unbroadcast(g['ug'], g['b'].shape, g['ug'].shape), g['gb'])


@pytest.mark.parametrize('name', golden_names('conv2d_'))
def test_conv2d(name):
  g = load_golden(name)
  p, s = int(g['pad']), int(g['stride'])
  close(ops.conv2d_fwd(g['x'], g['w'], g['b'], p, s), g['out'])
  close(ops.conv2d_dgrad(g['ug'], g['w'], g['x'].shape, p, s), g['gx'])
  close(ops.conv2d_wgrad(g['ug'], g['x'], g['w'].shape, p, s), g['gw'])
  close(ops.conv2d_bgrad(g['ug']), g['gb'])


@pytest.mark.parametrize('name', golden_names('conv3d_'))
def test_conv3d(name):
  g = load_golden(name)
  p, s = int(g['pad']), int(g['stride'])
  close(ops.conv3d_fwd(g['x'], g['w'], g['b'], p, s), g['out'])
  close(ops.conv3d_dgrad(g['ug'], g['w'], g['x'].shape, p, s), g['gx'])
  close(ops.conv3d_wgrad(g['ug'], g['x'], g['w'].shape, p, s), g['gw'])
  close(ops.conv3d_bgrad(g['ug']), g['gb'])


@pytest.mark.parametrize('name', golden_names('maxpool'))
def test_maxpool_bit_exact(name):
  g = load_golden(name)
  ks, p, s = (int(g['kh']), int(g['kw'])), int(g['pad']), int(g['stride'])
  y, idx = ops.maxpool_fwd(g['x'], ks, p, s)
  assert np.array_equal(y, g['out'])
  assert idx.dtype == np.uint8 and idx.shape == y.shape
  gx = ops.maxpool_bwd(g['ug'], g['x'], ks, p, s)
  assert np.array_equal(np.where(g['covered'], gx, 0.0), g['gx'])


@pytest.mark.parametrize('name', golden_names('sce_'))
def test_softmax_ce(name):
  g = load_golden(name)
  axis = int(g['axis'])
  close(ops.softmax_ce_fwd(g['logits'], g['targets'], axis), g['out'])
  close(ops.softmax_ce_bwd(g['logits'], g['targets'], axis), g['gx'])


@pytest.mark.parametrize('name,fn', [('loss_bce', models.bce), ('loss_mse', models.mse)])
def test_composite_losses(name, fn):
  g = load_golden(name)
  v, d = fn(g['outputs'], g['targets'])
  close(v, g['out'])
  close(d, g['gx'])


@pytest.mark.parametrize('name', ['adam', 'gd', 'momentum', 'rmsprop'])
def test_optimizers(name):
  g = load_golden('optim_' + name)
  ps = [g[f'p0_{i}'] for i in range(3)]
  st1 = [0.0] * 3
  st2 = [0.0] * 3
  for step in range(4):
    for i in range(3):
      gr = g[f'g{step}_{i}']
      if name == 'adam':
        ps[i], st1[i], st2[i] = ops.adam_step(ps[i], gr, st1[i], st2[i], step + 1, 0.05)
      elif name == 'gd':
        ps[i] = ops.sgd_step(ps[i], gr, 0.1)
      elif name == 'momentum':
        ps[i], st1[i] = ops.momentum_step(ps[i], gr, st1[i], 0.1)
      else:
        ps[i], st1[i] = ops.rmsprop_step(ps[i], gr, st1[i], 0.01)
      close(ps[i], g[f'p{step + 1}_{i}'])


def _set_init(net, g):
  n = len(net.params())
  net.set_params([g[f'init_{i}'] for i in range(n)])
  return n


def test_step_c1_mlp():
  g = load_golden('step_c1_mlp')
  net = models.mlp_circles(np.random.default_rng(0))
  n = _set_init(net, g)
  opt = models.Adam(net, 0.05)
  losses = [models.train_step(net, opt, 'bce', g['X'], g['y']) for _ in range(50)]
  close(losses, g['losses'], 1e-9)
  for i in range(n):
    close(net.params()[i], g[f'final_{i}'], 1e-8)


def test_step_c2_digits():
  g = load_golden('step_c2_digits')
  net = models.digits_cnn(np.random.default_rng(0))
  n = _set_init(net, g)
  opt = models.Adam(net, 5e-3)
  losses = []
  for _ in range(5):
    for s in range(0, 400, 200):
      losses.append(models.train_step(net, opt, 'softmax_ce', g['X'][s:s + 200], g['Y'][s:s + 200]))
  close(losses, g['losses'], 1e-9)
  for i in range(n):
    close(net.params()[i], g[f'final_{i}'], 1e-8)


@pytest.mark.parametrize('kind', ['a', 'b'])
def test_step_c3_lenet(kind):
  g = load_golden(f'step_c3{kind}_lenet')
  net = (models.lenet_a if kind == 'a' else models.lenet_b)(np.random.default_rng(0))
  n = _set_init(net, g)
  out = net.forward(g['X'])
  _, dout = models.softmax_ce(out, g['Y'])
  grads = net.backward(dout)
  for i in range(n):
    close(np.reshape(grads[i], g[f'grad0_{i}'].shape), g[f'grad0_{i}'], 1e-11)
  opt = models.Adam(net, 1e-3)
  losses = [models.train_step(net, opt, 'softmax_ce', g['X'], g['Y']) for _ in range(4)]
  close(losses, g['losses'], 1e-9)
  for i in range(n):
    close(net.params()[i], g[f'final_{i}'], 1e-7)


def test_step_c4_gan784():
  """Config C4 at the BASELINE dims (784-512-256), batch 512: the oracle's GAN iteration against the unmodified reference's
  recorded run (oracle/gen_golden_gan784.py): first-step gradients of all twelve parameters, three iterations of losses,
  final parameters (each compared on its L2 norm and every 97th element)."""
  g = load_golden('step_c4_gan784')
  B, iters, lr = int(g['batch']), int(g['iters']), float(g['lr'])
  params, real, noise_g, noise_d = models.gan784_inputs(B, iters)
  gen, disc = models.gan_nets(np.random.default_rng(0))
  gen.set_params(params[:6])
  disc.set_params(params[6:])
  opt_g, opt_d = models.Adam(gen, lr), models.Adam(disc, lr)
  lg, ld = [], []
  for it in range(iters):
    a, b, gg, gd = models.gan_iteration(gen, disc, opt_g, opt_d, noise_g[it], noise_d[it], real, return_grads=True)
    lg.append(a)
    ld.append(b)
    if it == 0:
      for i, gr in enumerate(gg + gd):
        close(models.sample(gr), g[f'grad_{i}'], 1e-11)
  close(lg, g['losses_g'], 1e-10)
  close(ld, g['losses_d'], 1e-10)
  for i, p in enumerate(gen.params() + disc.params()):
    close(models.sample(p), g[f'final_{i}'], 1e-8)


def test_geometry_and_errors():
  assert ops.out_hw(12, 15, 4, 4, 0, 2) == (5, 6)
  assert ops.out_hw(12, 15, 2, 2, 2, 1) == (15, 18)
  with pytest.raises(AssertionError):
    ops.out_hw(5, 5, 3, 3, 0, 0)
  with pytest.raises(AssertionError):
    ops.out_hw(5, 5, 3, 3, -1, 1)
  with pytest.raises(ValueError):
    ops.conv3d_fwd(np.zeros((2, 5, 5)), np.zeros((1, 1, 3, 3)), np.zeros(1))
  with pytest.raises(ValueError):
    ops.conv2d_fwd(np.zeros((2, 1, 5, 5)), np.zeros((3, 3)), 0.0)
  x = np.arange(2 * 3 * 4, dtype=float).reshape(2, 3, 4)
  assert np.array_equal(ops.unscramble(ops.scramble(x)), x)
  assert ops.unbroadcast_axes((1, 3), (5, 3)) == (0,)
  assert ops.unbroadcast_axes((), (2, 3)) == (0, 1)
  assert ops.unbroadcast_axes((3,), (2, 2, 3)) == (0, 1)
