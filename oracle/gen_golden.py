"""Generate tests/golden/*.npz by running the UNMODIFIED reference (TEST INFRASTRUCTURE).

Run in the build container only (the GPU box has no /root/reference):
    python oracle/gen_golden.py            # rewrites tests/golden/
The reference is imported read-only under the alias `neograd_ref`.  All inputs are
seeded and float32-representable, so the fp32 device path sees bit-identical inputs.
Each fixture stores inputs, the op/geometry, and the reference's forward value and per-operand
gradients (as accumulated by Tensor._backward, tensor.py:93-100).
"""
import importlib.util
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
OUT = os.path.join(ROOT, 'tests', 'golden')


def load_reference(root=None):
  root = root or os.environ.get('NEOGRAD_REF', '/root/reference')
  spec = importlib.util.spec_from_file_location(
    'neograd_ref', f'{root}/neograd/__init__.py', submodule_search_locations=[f'{root}/neograd'])
  m = importlib.util.module_from_spec(spec)
  sys.modules['neograd_ref'] = m
  spec.loader.exec_module(m)
  return m


def f32(x):
  return np.asarray(x, np.float32).astype(np.float64)


def main():
  ng = load_reference()
  nn = ng.nn
  from neograd_ref.nn import optim as roptim, loss as rloss
  os.makedirs(OUT, exist_ok=True)
  rng = np.random.default_rng(20261019)
  manifest = {}

  def save(name, **arrs):
    np.savez_compressed(os.path.join(OUT, name + '.npz'), **arrs)
    manifest[name] = sorted(arrs.keys())

  def T(x, rg=True):
    return ng.tensor(np.array(x, dtype=float), requires_grad=rg)

  def grad_of(t):
    return np.array(t.grad, dtype=float).reshape(t.shape)

  # ------------------------------------------------------------------ binary ops (test_basics.py:7-31 operands + random)
  a = np.array(3.); b = np.array([1., 2, 3]); c = np.array([[3., 4, 5], [6, 7, 8]])
  d = np.array([[[9., 8, 7], [6, 5, 4]], [[1, 2, 3], [4, 5, 6]]])
  fpos = f32(np.abs(rng.standard_normal((4, 1, 5))) + 0.5)
  gpos = f32(np.abs(rng.standard_normal((3, 1))) + 0.5)
  pairs = [(a, b), (a, d), (b, c), (c, d), (c, c + 1), (f32(rng.standard_normal((7, 5))), f32(rng.standard_normal((1, 5)))),
           (fpos, gpos), (f32(rng.standard_normal((6, 1)) + 3), f32(rng.standard_normal((1, 4)) + 3))]
  fns = {'add': ng.add, 'sub': ng.sub, 'mul': ng.mul, 'div': ng.div, 'pow': ng.pow}
  k = 0
  for opname, fn in fns.items():
    for (x, y) in pairs:
      if opname == 'pow':
        x = np.abs(x) + 0.5
      if opname == 'div':
        y = np.where(np.abs(y) < 0.3, 0.7, y)
      with ng.new_graph():
        tx, ty = T(x), T(y)
        out = fn(tx, ty)
        ug = f32(rng.standard_normal(out.shape))
        out.backward(ug)
        save(f'binary_{k:02d}_{opname}', op=np.array(opname), a=x, b=y, ug=ug, out=out.data, ga=grad_of(tx), gb=grad_of(ty))
      k += 1

  # ------------------------------------------------------------------ unary ops
  xs = f32(rng.standard_normal((5, 7)))
  xs[0, :3] = 0.0                                   # exact zeros: relu/leaky grad is 1 at 0 (activations.py:31,216)
  un = {'exp': ng.exp, 'log': ng.log, 'relu': nn.ReLU(), 'leaky_relu': nn.LeakyReLU(), 'sigmoid': nn.Sigmoid(), 'tanh': nn.Tanh()}
  for opname, fn in un.items():
    x = np.abs(xs) + 0.25 if opname == 'log' else xs
    with ng.new_graph():
      tx = T(x)
      out = fn(tx)
      ug = f32(rng.standard_normal(out.shape))
      out.backward(ug)
      save(f'unary_{opname}', op=np.array(opname), x=x, ug=ug, out=out.data, gx=grad_of(tx), alpha=np.array(0.01))

  # ------------------------------------------------------------------ sum / transpose / flatten / reshape
  x3 = f32(rng.standard_normal((4, 3, 5)))
  for axis in (None, 0, 1, 2):
    with ng.new_graph():
      tx = T(x3)
      out = ng.sum(tx, axis=axis)
      ug = f32(rng.standard_normal(out.shape))
      out.backward(ug)
      save(f'sum_axis{axis}', x=x3, axis=np.array(-99 if axis is None else axis), ug=ug, out=out.data, gx=grad_of(tx))

  # ------------------------------------------------------------------ dot (test_basics.py:41 + Linear shapes)
  for i, (m_, k_, n_) in enumerate([(2, 2, 3), (7, 5, 3), (33, 17, 9), (64, 36, 10), (128, 64, 32), (5, 1, 4)]):
    A, B = f32(rng.standard_normal((m_, k_))), f32(rng.standard_normal((k_, n_)))
    with ng.new_graph():
      ta, tb = T(A), T(B)
      out = ng.dot(ta, tb)
      ug = f32(rng.standard_normal(out.shape))
      out.backward(ug)
      save(f'dot_{i}', a=A, b=B, ug=ug, out=out.data, ga=grad_of(ta), gb=grad_of(tb))

  # ------------------------------------------------------------------ Linear (test_layers.py:28-30)
  np.random.seed(5)
  lin = nn.Linear(5, 3)
  W0, b0 = f32(lin.weights.data), f32(rng.standard_normal((1, 3)))
  lin.weights.data, lin.bias.data = W0, b0
  x = f32(rng.standard_normal((2, 5)))
  with ng.new_graph():
    tx = T(x)
    out = lin(tx)
    ug = f32(rng.standard_normal(out.shape))
    out.backward(ug)
    save('linear_5_3', x=x, w=W0, b=b0, ug=ug, out=out.data, gx=grad_of(tx), gw=grad_of(lin.weights), gb=grad_of(lin.bias))

  # ------------------------------------------------------------------ Conv2D (test_layers.py:35-43 + square / LeNet / digits geometry)
  conv2d_cases = [((2, 12, 15), (3, 3), 0, 1), ((2, 12, 15), (2, 2), 2, 1), ((2, 12, 15), (4, 4), 0, 2), ((2, 12, 15), (3, 3), 1, 1),
                  ((3, 8, 8), (3, 3), 0, 1), ((2, 28, 28), (5, 5), 0, 1), ((2, 11, 14), (2, 3), 2, 2), ((1, 9, 7), (3, 3), 1, 3)]
  for i, (xs_, ks_, p_, s_) in enumerate(conv2d_cases):
    layer = nn.Conv2D(ks_, padding=p_, stride=s_)
    w0, bb = f32(rng.standard_normal(ks_)), float(f32(rng.standard_normal()))
    layer.weights.data, layer.bias.data = w0, bb
    x = f32(rng.standard_normal(xs_))
    with ng.new_graph():
      tx = T(x)
      out = layer(tx)
      ug = f32(rng.standard_normal(out.shape))
      out.backward(ug)
      save(f'conv2d_{i}', x=x, w=w0, b=np.array(bb), pad=np.array(p_), stride=np.array(s_), ug=ug, out=out.data,
           gx=grad_of(tx), gw=grad_of(layer.weights), gb=grad_of(layer.bias))

  # ------------------------------------------------------------------ Conv3D (test_layers.py:48-56 + extras)
  conv3d_cases = [((2, 2, 12, 13), 5, (3, 3), 0, 1), ((2, 2, 12, 13), 5, (2, 2), 2, 1), ((2, 2, 12, 13), 5, (4, 4), 0, 2),
                  ((2, 2, 12, 13), 5, (3, 3), 1, 1), ((3, 2, 11, 14), 4, (2, 3), 2, 2), ((2, 1, 28, 28), 6, (5, 5), 0, 1),
                  ((2, 6, 12, 12), 16, (5, 5), 0, 1), ((2, 8, 9, 9), 8, (3, 3), 1, 1)]
  for i, (xs_, o_, ks_, p_, s_) in enumerate(conv3d_cases):
    layer = nn.Conv3D(xs_[1], o_, ks_, padding=p_, stride=s_)
    w0, bb = f32(rng.standard_normal((o_, xs_[1]) + ks_)), f32(rng.standard_normal(o_))
    layer.weights.data, layer.bias.data = w0, bb
    x = f32(rng.standard_normal(xs_))
    with ng.new_graph():
      tx = T(x)
      out = layer(tx)
      ug = f32(rng.standard_normal(out.shape))
      out.backward(ug)
      save(f'conv3d_{i}', x=x, w=w0, b=bb, pad=np.array(p_), stride=np.array(s_), ug=ug, out=out.data,
           gx=grad_of(tx), gw=grad_of(layer.weights), gb=grad_of(layer.bias))

  # ------------------------------------------------------------------ MaxPool (test_layers.py:61-70 + ties / padding / overlap / partial cover)
  def pool_case(name, x, ks_, p_, s_, three_d):
    layer = (nn.MaxPool3D if three_d else nn.MaxPool2D)(ks_, padding=p_, stride=s_)
    with ng.new_graph():
      tx = T(x)
      out = layer(tx)
      ug = f32(rng.standard_normal(out.shape))
      out.backward(ug)
      # cells covered by at least one window (the reference leaves the rest uninitialised, conv.py:419/518)
      H, W = x.shape[-2:]
      ho, wo = out.shape[-2:]
      cov = np.zeros((H + 2 * p_, W + 2 * p_), bool)
      for pi in range(ho):
        for qi in range(wo):
          cov[pi * s_:pi * s_ + ks_[0], qi * s_:qi * s_ + ks_[1]] = True
      cov = cov[p_:p_ + H, p_:p_ + W]
      save(name, x=x, kh=np.array(ks_[0]), kw=np.array(ks_[1]), pad=np.array(p_), stride=np.array(s_), ug=ug,
           out=out.data, gx=np.where(cov, grad_of(tx), 0.0), covered=cov)

  pool_case('maxpool2d_0', f32(rng.standard_normal((2, 12, 15))), (3, 3), 0, 3, False)
  pool_case('maxpool3d_0', f32(rng.standard_normal((2, 3, 12, 15))), (3, 3), 0, 3, True)
  pool_case('maxpool2d_1', f32(rng.standard_normal((3, 12, 12))), (2, 2), 0, 2, False)
  pool_case('maxpool3d_1', f32(rng.standard_normal((2, 6, 24, 24))), (2, 2), 0, 2, True)
  ties = np.maximum(0, f32(rng.standard_normal((2, 4, 8, 8))))          # post-ReLU zeros -> ties
  ties[0, 0] = 1.0                                                        # all-equal windows
  ties[1, 1] = -1.0                                                       # all-negative: still argmax 0
  pool_case('maxpool3d_ties', ties, (2, 2), 0, 2, True)
  pool_case('maxpool2d_pad', f32(rng.standard_normal((2, 6, 6)) - 1.0), (2, 2), 1, 2, False)   # zero padding wins
  pool_case('maxpool2d_overlap', f32(rng.standard_normal((2, 7, 7))), (3, 3), 0, 2, False)    # later windows overwrite
  pool_case('maxpool2d_partial', f32(rng.standard_normal((2, 5, 5))), (2, 2), 0, 2, False)    # uncovered last row/col
  pool_case('maxpool3d_rect', f32(rng.standard_normal((2, 2, 6, 9))), (2, 3), 0, 1, True)

  # ------------------------------------------------------------------ SoftmaxCE
  for name, shape, axis in (('sce_axis1', (6, 10), 1), ('sce_axis0', (10, 6), 0)):
    logits = f32(3 * rng.standard_normal(shape))
    tgt = np.zeros(shape)
    if axis == 1:
      tgt[np.arange(shape[0]), rng.integers(0, shape[1], shape[0])] = 1
    else:
      tgt[rng.integers(0, shape[0], shape[1]), np.arange(shape[1])] = 1
    with ng.new_graph():
      tl = T(logits)
      loss = rloss.SoftmaxCE(axis=axis)(tl, ng.tensor(tgt))
      loss.backward()
      save(name, logits=logits, targets=tgt, axis=np.array(axis), out=np.array(loss.data), gx=grad_of(tl))

  # ------------------------------------------------------------------ composite losses (BCE swapped-arg quirk, MSE, CE)
  o = f32(rng.uniform(0.05, 0.95, (9, 1))); t = f32(rng.integers(0, 2, (9, 1)))
  for name, L in (('loss_bce', rloss.BCE()), ('loss_mse', rloss.MSE()), ('loss_ce', rloss.CE())):
    with ng.new_graph():
      to = T(o)
      loss = L(to, ng.tensor(t))
      loss.backward()
      save(name, outputs=o, targets=t, out=np.array(loss.data), gx=grad_of(to))

  # ------------------------------------------------------------------ optimizers: 4 steps with given grads
  p0 = [f32(rng.standard_normal((4, 3))), f32(rng.standard_normal((1, 3))), f32(rng.standard_normal(()))]
  gs = [[f32(rng.standard_normal(p.shape)) for p in p0] for _ in range(4)]
  for name, mk in (('adam', lambda ps: roptim.Adam(ps, 0.05)), ('gd', lambda ps: roptim.GD(ps, 0.1)),
                   ('momentum', lambda ps: roptim.Momentum(ps, 0.1)), ('rmsprop', lambda ps: roptim.RMSProp(ps, 0.01))):
    ps = [nn.layers.Param(np.array(p), requires_grad=True) for p in p0]
    opt = mk(ps)
    arrs = {f'p0_{i}': p for i, p in enumerate(p0)}
    for s, g in enumerate(gs):
      for prm, gg in zip(ps, g):
        prm.grad = np.array(gg)
      opt.step()
      for i, prm in enumerate(ps):
        arrs[f'p{s + 1}_{i}'] = np.array(prm.data)
        arrs[f'g{s}_{i}'] = g[i]
    save(f'optim_{name}', **arrs)

  # ------------------------------------------------------------------ step level: C1 README MLP (README.md:80-115), 50 Adam iters
  from sklearn.datasets import make_circles
  from sklearn.model_selection import train_test_split
  X, y = make_circles(n_samples=1000, noise=0.05, random_state=100)
  Xtr, _, ytr, _ = train_test_split(X, y, random_state=0)
  Xtr, ytr = f32(Xtr[:750]), f32(ytr[:750].reshape(750, 1))

  class NN(nn.Model):
    def __init__(self):
      self.stack = nn.Sequential(nn.Linear(2, 100), nn.ReLU(), nn.Linear(100, 1), nn.Sigmoid())

    def forward(self, inputs):
      return self.stack(inputs)

  np.random.seed(0)
  model = NN()
  for prm in model.parameters():
    prm.data = f32(prm.data)
  init = [np.array(prm.data) for prm in model.parameters()]
  opt = roptim.Adam(model.parameters(), 0.05)
  lossf = rloss.BCE()
  tx, ty = ng.tensor(Xtr), ng.tensor(ytr)
  losses = []
  for _ in range(50):
    opt.zero_grad()
    loss = lossf(model(tx), ty)
    loss.backward()
    opt.step()
    losses.append(float(loss.data))
  arrs = {'X': Xtr.astype(np.float32), 'y': ytr.astype(np.float32), 'losses': np.array(losses)}
  for i, (p_i, p_f) in enumerate(zip(init, model.parameters())):
    arrs[f'init_{i}'] = p_i
    arrs[f'final_{i}'] = np.array(p_f.data)
  save('step_c1_mlp', **arrs)

  # ------------------------------------------------------------------ step level: C2 digits CNN (digits.ipynb cells 5-10), 2 batches x 5 epochs
  from sklearn.datasets import load_digits
  Xd, yd = load_digits(return_X_y=True)
  Xd, _, yd, _ = train_test_split(Xd, yd, random_state=0)
  Xd = Xd[:400]; yd = yd[:400]
  Xd = (Xd - Xd.mean(axis=1, keepdims=True)) / Xd.std(axis=1, keepdims=True)
  Xd = f32(Xd.reshape(-1, 8, 8))
  Yd = np.eye(10)[yd]

  class CNN(nn.Model):
    def __init__(self):
      self.conv = nn.Sequential(nn.Conv2D((3, 3)), nn.ReLU())
      self.stack = nn.Sequential(nn.Linear(36, 10))

    def forward(self, inputs):
      c = self.conv(inputs)
      return self.stack(c.reshape((inputs.shape[0], 36)))

  np.random.seed(0)
  model = CNN()
  for prm in model.parameters():
    prm.data = f32(prm.data)
  init = [np.array(prm.data) for prm in model.parameters()]
  opt = roptim.Adam(model.parameters(), 5e-3)
  lossf = rloss.SoftmaxCE(axis=1)
  losses = []
  for _ in range(5):
    for bx, by in nn.utils.get_batches(ng.tensor(Xd), ng.tensor(Yd), 200):
      opt.zero_grad()
      loss = lossf(model(bx), by)
      loss.backward()
      opt.step()
      losses.append(float(loss.data))
  arrs = {'X': Xd.astype(np.float32), 'Y': Yd.astype(np.float32), 'losses': np.array(losses)}
  for i, (p_i, p_f) in enumerate(zip(init, model.parameters())):
    arrs[f'init_{i}'] = p_i
    arrs[f'final_{i}'] = np.array(p_f.data)
  save('step_c2_digits', **arrs)

  # ------------------------------------------------------------------ step level: C3A / C3B LeNet, batch 16, 4 Adam(1e-3) steps
  def lenet(kind):
    class LeNet(nn.Model):
      def __init__(self):
        if kind == 'a':
          self.features = nn.Sequential(nn.Conv2D((5, 5)), nn.ReLU(), nn.MaxPool2D((2, 2), stride=2),
                                        nn.Conv2D((5, 5)), nn.ReLU(), nn.MaxPool2D((2, 2), stride=2))
          self.flat = 16
        else:
          self.features = nn.Sequential(nn.Conv3D(1, 6, (5, 5)), nn.ReLU(), nn.MaxPool3D((2, 2), stride=2),
                                        nn.Conv3D(6, 16, (5, 5)), nn.ReLU(), nn.MaxPool3D((2, 2), stride=2))
          self.flat = 256
        self.head = nn.Sequential(nn.Linear(self.flat, 120), nn.ReLU(), nn.Linear(120, 84), nn.ReLU(), nn.Linear(84, 10))

      def forward(self, inputs):
        f = self.features(inputs)
        return self.head(f.reshape((inputs.shape[0], self.flat)))
    return LeNet()

  for kind in ('a', 'b'):
    r2 = np.random.default_rng(1)
    np.random.seed(0)
    model = lenet(kind)
    for prm in model.parameters():
      prm.data = f32(0.1 * r2.standard_normal(prm.shape)) if prm.shape != () else float(f32(0.1 * r2.standard_normal()))
    init = [np.array(prm.data) for prm in model.parameters()]
    B = 16
    r0 = np.random.default_rng(0)
    Xl = f32(r0.standard_normal((B, 28, 28) if kind == 'a' else (B, 1, 28, 28)))
    Yl = np.eye(10)[r0.integers(0, 10, B)]
    opt = roptim.Adam(model.parameters(), 1e-3)
    lossf = rloss.SoftmaxCE(axis=1)
    losses = []
    g0 = None
    for it in range(4):
      opt.zero_grad()
      loss = lossf(model(ng.tensor(Xl)), ng.tensor(Yl))
      loss.backward()
      if it == 0:
        g0 = [np.array(prm.grad, dtype=float).reshape(prm.shape) for prm in model.parameters()]
      opt.step()
      losses.append(float(loss.data))
    arrs = {'X': Xl.astype(np.float32), 'Y': Yl.astype(np.float32), 'losses': np.array(losses)}
    for i, (p_i, p_f) in enumerate(zip(init, model.parameters())):
      arrs[f'init_{i}'] = p_i
      arrs[f'final_{i}'] = np.array(p_f.data)
      arrs[f'grad0_{i}'] = g0[i]
    save(f'step_c3{kind}_lenet', **arrs)

  with open(os.path.join(OUT, 'MANIFEST.json'), 'w') as fp:
    json.dump({'generator': 'oracle/gen_golden.py', 'reference': 'pranftw/neograd v0.0.4 (unmodified, /root/reference)',
               'numpy': np.__version__, 'fixtures': manifest}, fp, indent=1, sort_keys=True)
  print(f'wrote {len(manifest)} fixtures to {OUT}')


if __name__ == '__main__':
  main()
