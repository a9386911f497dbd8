"""Step-level oracle: sequential nets built from oracle.ops with explicit backward + optimizer
(TEST INFRASTRUCTURE, see oracle/__init__.py).

A net is a list of layer specs; parameters live in a flat list in the order neograd's
`Model.parameters()` would return them (layers/__init__.py:112-129: weights then bias per layer).
Reference call stack reproduced: README.md:109-115 (zero_grad / forward / loss / backward / step).
"""
import numpy as np
from . import ops

F64 = np.float64


class SeqNet:
  """layers: list of tuples
       ('linear', W(in,out), b(1,out))        layers/misc.py:37-63
       ('conv2d', w(KH,KW), b(), pad, stride) layers/conv.py:6-45
       ('conv3d', w(O,C,KH,KW), b(O,), pad, stride)  layers/conv.py:48-91
       ('maxpool', (KH,KW), pad, stride)      layers/conv.py:94-163
       ('relu',) ('leaky_relu', alpha) ('sigmoid',) ('tanh',)   activations.py
       ('reshape', shape_without_batch)       -> x.reshape((B,)+shape)
  """

  def __init__(self, layers):
    self.layers = [list(l) for l in layers]
    for l in self.layers:
      for i, v in enumerate(l):
        if isinstance(v, np.ndarray) or (l[0] in ('conv2d',) and i == 2):
          l[i] = np.array(v, dtype=F64)

  # parameters in Model.parameters() order
  def params(self):
    out = []
    for l in self.layers:
      if l[0] in ('linear', 'conv2d', 'conv3d'):
        out += [l[1], l[2]]
    return out

  def set_params(self, new):
    it = iter(new)
    for l in self.layers:
      if l[0] in ('linear', 'conv2d', 'conv3d'):
        l[1] = np.array(next(it), dtype=F64)
        l[2] = np.array(next(it), dtype=F64)

  def forward(self, x):
    x = np.asarray(x, F64)
    self.cache = []
    for l in self.layers:
      kind = l[0]
      self.cache.append(x)
      if kind == 'linear':
        x = ops.dot_fwd(x, l[1]) + l[2]
      elif kind == 'conv2d':
        x = ops.conv2d_fwd(x, l[1], l[2], l[3], l[4])
      elif kind == 'conv3d':
        x = ops.conv3d_fwd(x, l[1], l[2], l[3], l[4])
      elif kind == 'maxpool':
        x = ops.maxpool_fwd(x, l[1], l[2], l[3])[0]
      elif kind in ('relu', 'sigmoid', 'tanh'):
        x = ops.unary_fwd(kind, x)
      elif kind == 'leaky_relu':
        x = ops.unary_fwd(kind, x, l[1])
      elif kind == 'reshape':
        x = x.reshape((x.shape[0],) + tuple(l[1]))
      else:
        raise KeyError(kind)
    return x

  def backward(self, ug):
    """Returns grads in params() order."""
    grads = []
    for l, x in zip(reversed(self.layers), reversed(self.cache)):
      kind = l[0]
      if kind == 'linear':
        gw = ops.dot_wgrad(x, ug)
        gb = ops.unbroadcast(ug, l[2].shape, ug.shape)
        ug = ops.dot_dgrad(ug, l[1])
        grads += [gb, gw]
      elif kind == 'conv2d':
        gw = ops.conv2d_wgrad(ug, x, l[1].shape, l[3], l[4])
        gb = ops.conv2d_bgrad(ug)
        ug = ops.conv2d_dgrad(ug, l[1], x.shape, l[3], l[4])
        grads += [gb, gw]
      elif kind == 'conv3d':
        gw = ops.conv3d_wgrad(ug, x, l[1].shape, l[3], l[4])
        gb = ops.conv3d_bgrad(ug)
        ug = ops.conv3d_dgrad(ug, l[1], x.shape, l[3], l[4])
        grads += [gb, gw]
      elif kind == 'maxpool':
        ug = ops.maxpool_bwd(ug, x, l[1], l[2], l[3])
      elif kind in ('relu', 'sigmoid', 'tanh'):
        ug = ops.unary_bwd(kind, x, ug)
      elif kind == 'leaky_relu':
        ug = ops.unary_bwd(kind, x, ug, l[1])
      elif kind == 'reshape':
        ug = ug.reshape(x.shape)
    self.input_grad = ug
    return grads[::-1]


# ----------------------------------------------------------------------------- losses (value, d/d outputs)
def bce(outputs, targets, epsilon=1e-9):
  """loss.py:69-85 -- note the reference's swapped roles of outputs/targets."""
  outputs, targets = np.asarray(outputs, F64), np.asarray(targets, F64)
  n = ops.num_examples(outputs.shape)
  lt, l1t = np.log(targets + epsilon), np.log(1 - targets + epsilon)
  cost = (-1 / n) * np.sum(outputs * lt + (1 - outputs) * l1t)
  return cost, (-1 / n) * (lt - l1t) * np.ones(outputs.shape)


def mse(outputs, targets):
  """loss.py:44-56."""
  outputs, targets = np.asarray(outputs, F64), np.asarray(targets, F64)
  n = ops.num_examples(outputs.shape)
  d = outputs - targets
  return (1 / (2 * n)) * np.sum(d ** 2), (1 / n) * d


def softmax_ce(outputs, targets, axis=1):
  """loss.py:142-170."""
  return ops.softmax_ce_fwd(outputs, targets, axis), ops.softmax_ce_bwd(outputs, targets, axis)


LOSSES = {'bce': bce, 'mse': mse, 'softmax_ce': softmax_ce}


class Adam:
  """optim.py:150-208."""

  def __init__(self, net, lr, beta1=0.9, beta2=0.999, epsilon=1e-8):
    self.net, self.lr, self.b1, self.b2, self.eps, self.iter = net, lr, beta1, beta2, epsilon, 0
    self.m = [0.0 for _ in net.params()]
    self.v = [0.0 for _ in net.params()]

  def step(self, grads):
    self.iter += 1
    new = []
    for i, (p, g) in enumerate(zip(self.net.params(), grads)):
      p2, self.m[i], self.v[i] = ops.adam_step(p, np.asarray(g, F64).reshape(p.shape), self.m[i], self.v[i],
                                               self.iter, self.lr, self.b1, self.b2, self.eps)
      new.append(p2)
    self.net.set_params(new)


class GD:
  """optim.py:36-47."""

  def __init__(self, net, lr):
    self.net, self.lr = net, lr

  def step(self, grads):
    self.net.set_params([ops.sgd_step(p, np.asarray(g, F64).reshape(p.shape), self.lr)
                         for p, g in zip(self.net.params(), grads)])


def train_step(net, optim, loss_name, x, y):
  """One README.md:109-115 iteration; returns the loss (float)."""
  out = net.forward(x)
  loss, dout = LOSSES[loss_name](out, y)
  optim.step(net.backward(dout))
  return float(loss)


# ----------------------------------------------------------------------------- the BASELINE configs
def lenet_a(rng, scale=0.1):
  """neograd Conv2D/MaxPool2D (single channel), 13 106 parameters."""
  r = lambda *s: scale * rng.standard_normal(s)
  return SeqNet([
    ('conv2d', r(5, 5), np.array(0.0), 0, 1), ('relu',), ('maxpool', (2, 2), 0, 2),
    ('conv2d', r(5, 5), np.array(0.0), 0, 1), ('relu',), ('maxpool', (2, 2), 0, 2),
    ('reshape', (16,)),
    ('linear', r(16, 120), np.zeros((1, 120))), ('relu',),
    ('linear', r(120, 84), np.zeros((1, 84))), ('relu',),
    ('linear', r(84, 10), np.zeros((1, 10)))])


def lenet_b(rng, scale=0.1):
  """neograd Conv3D(1,6,5x5)/MaxPool3D/Conv3D(6,16,5x5), 44 426 parameters."""
  r = lambda *s: scale * rng.standard_normal(s)
  return SeqNet([
    ('conv3d', r(6, 1, 5, 5), np.zeros(6), 0, 1), ('relu',), ('maxpool', (2, 2), 0, 2),
    ('conv3d', r(16, 6, 5, 5), np.zeros(16), 0, 1), ('relu',), ('maxpool', (2, 2), 0, 2),
    ('reshape', (256,)),
    ('linear', r(256, 120), np.zeros((1, 120))), ('relu',),
    ('linear', r(120, 84), np.zeros((1, 84))), ('relu',),
    ('linear', r(84, 10), np.zeros((1, 10)))])


def mlp_circles(rng):
  """README.md:95-106: Linear(2,100) ReLU Linear(100,1) Sigmoid."""
  return SeqNet([('linear', rng.standard_normal((2, 100)), np.zeros((1, 100))), ('relu',),
                 ('linear', rng.standard_normal((100, 1)), np.zeros((1, 1))), ('sigmoid',)])


def digits_cnn(rng):
  """examples/digits.ipynb cell 7: Conv2D((3,3)) ReLU reshape(B,36) Linear(36,10)."""
  return SeqNet([('conv2d', rng.standard_normal((3, 3)), np.array(0.0), 0, 1), ('relu',), ('reshape', (36,)),
                 ('linear', rng.standard_normal((36, 10)), np.zeros((1, 10)))])


# ----------------------------------------------------------------------------- config C4: the MLP GAN
def gan_nets(rng, data_dim=784, h1=512, h2=256, scale=0.05, leak=0.01):
  """examples/gans/basic.ipynb cell 7 with the dims of SURVEY 8d C4: generator data_dim-h1-h2-data_dim, discriminator
  data_dim-h1-h2-1-Sigmoid, LeakyReLU between the Linears.  Draw order = Model.parameters() order (generator first)."""
  r = lambda *s: scale * rng.standard_normal(s)
  def stack(dims, tail):
    layers = []
    for i, (a, b) in enumerate(zip(dims[:-1], dims[1:])):
      layers.append(('linear', r(a, b), np.zeros((1, b))))
      if i + 2 < len(dims):
        layers.append(('leaky_relu', leak))
    return SeqNet(layers + tail)
  gen = stack([data_dim, h1, h2, data_dim], [])
  disc = stack([data_dim, h1, h2, 1], [('sigmoid',)])
  return gen, disc


def gan_iteration(gen, disc, opt_g, opt_d, noise_g, noise_d, real, return_grads=False):
  """One iteration of basic.ipynb cells 9-10: generator step (discriminator frozen, BCE(D(G(z)), 1)), then
  discriminator step (generator frozen, BCE(D(real), 1) + BCE(D(G(z')), 0)).  BCE keeps the reference's swapped
  argument roles (loss.py:83).  A frozen sub-model's parameters have requires_grad=False, so its Adam moments and
  values are untouched by the other optimizer's step (optim.py:173-174): two independent Adam states, each with its
  own `iter`.  Returns (loss_g, loss_d[, generator grads, discriminator grads])."""
  n = noise_g.shape[0]
  ones, zeros = np.ones((n, 1)), np.zeros((n, 1))
  out = disc.forward(gen.forward(noise_g))
  loss_g, dout = bce(out, ones)
  disc.backward(dout)
  g_gen = gen.backward(disc.input_grad)
  opt_g.step(g_gen)
  fake = gen.forward(noise_d)
  l_real, d_real = bce(disc.forward(real), np.ones((real.shape[0], 1)))
  g_real = disc.backward(d_real)
  l_fake, d_fake = bce(disc.forward(fake), zeros)
  g_fake = disc.backward(d_fake)
  g_disc = [a + b for a, b in zip(g_real, g_fake)]
  opt_d.step(g_disc)
  if return_grads:
    return float(loss_g), float(l_real + l_fake), g_gen, g_disc
  return float(loss_g), float(l_real + l_fake)


GAN784_DIMS = (784, 512, 256)
GAN784_STRIDE = 97


def _f32(x):
  return np.asarray(x, np.float32).astype(np.float64)


def gan784_inputs(batch, iters, seed=7):
  """Seeded, float32-representable parameters and data shared by the generator script and the tests."""
  d, h1, h2 = GAN784_DIMS
  rng = np.random.default_rng(seed)
  shapes = [(d, h1), (1, h1), (h1, h2), (1, h2), (h2, d), (1, d), (d, h1), (1, h1), (h1, h2), (1, h2), (h2, 1), (1, 1)]
  params = [_f32(0.05 * rng.standard_normal(s)) for s in shapes]
  real = _f32(rng.standard_normal((batch, d)))
  noise_g = _f32(rng.standard_normal((iters, batch, d)))
  noise_d = _f32(rng.standard_normal((iters, batch, d)))
  return params, real, noise_g, noise_d


def sample(a):
  a = np.asarray(a, dtype=np.float64).reshape(-1)
  return np.concatenate([[np.sqrt(np.sum(a * a))], a[::GAN784_STRIDE]])
