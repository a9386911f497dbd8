"""Optimizers (reference: neograd/nn/optim.py): GD, Momentum, RMSProp, Adam -- same constructors, `step()`,
`zero_grad(all_members)`, per-Param state attributes (`momentum_grad`, `rms_grad`) and frozen-param skipping.

B200 design: the optimizer moves its parameters into ONE flat fp32 arena (data, grad and state buffers laid out in
`Model.parameters()` order, every Param's `data` / `grad` becoming a pinned view into it), so that
  * `step()` is one fused, 128-bit-vectorised launch over the whole model instead of a Python loop with ~10 NumPy
    temporaries per parameter (optim.py:173-179);
  * data-parallel training needs a single NCCL allreduce of the flat gradient buffer, issued here between
    `loss.backward()` and the update (the 1/world averaging rides inside the update kernel as `grad_scale`).
State is keyed by the parameter set, not by the optimizer object, because the reference stores it ON the Param
(optim.py:188-189, 196-197): two optimizers over the same parameters share -- and the second constructor re-zeroes --
the moments, while each keeps its own `iter` (examples/gans/basic.ipynb relies on this; SURVEY hard part 8).
"""
import ctypes

from ..autograd.utils import get_graph
from ..device import DeviceArray, torch, _st
from .. import device as D
from .. import _backend as B
from .. import distributed as dist

_ARENAS = {}   # tuple(id(param)) -> Arena   (params keep the arena alive through their pinned views)


class Arena:
  """Flat storage for a parameter list.  Offsets are padded to 4 floats so every view stays 16-byte aligned."""

  def __init__(self, params):
    self.params = list(params)
    self.offsets = []
    n = 0
    for p in self.params:
      self.offsets.append(n)
      n += (p.data.size + 3) & ~3
    self.numel = n
    self.data = DeviceArray.zeros((max(n, 1),))
    # data parallel on one node: the gradient buffer lives in symmetric (peer-mapped) memory and the ranks sum each
    # other's buffers over NVLink themselves (distributed.PeerExchange); otherwise a plain buffer + NCCL allreduce
    self.exchange = dist.PeerExchange.create(n)
    if self.exchange is not None:
      self.grad = DeviceArray(self.exchange.grad[:max(n, 1)])
      self.grad_sum = DeviceArray(self.exchange.summed[:max(n, 1)])
    else:
      self.grad = DeviceArray.zeros((max(n, 1),))
      self.grad_sum = self.grad
    self.state = {}
    for p, off in zip(self.params, self.offsets):
      size, shape = p.data.size, p.shape
      view = self.data[off:off + size].reshape(shape)
      view.assign(p.data)
      gview = self.grad[off:off + size].reshape(shape)
      if p._grad_live:
        gview.assign(p._grad)
      p._data, p._pinned = view, True
      p._grad, p._grad_pinned = gview, True

  def state_buffer(self, name, attr):
    """Flat state buffer `name`, (re)zeroed, with a view installed on every trainable Param as `attr`."""
    buf = self.state.get(name)
    if buf is None:
      buf = self.state[name] = DeviceArray.empty((max(self.numel, 1),))
    buf.fill(0.0)
    for p, off in zip(self.params, self.offsets):
      if p.requires_grad:
        setattr(p, attr, buf[off:off + p.data.size].reshape(p.shape))
    return buf

  def ranges(self):
    """Contiguous [begin, end) element ranges covering the trainable params (one range when nothing is frozen)."""
    out = []
    for p, off in zip(self.params, self.offsets):
      if not p.requires_grad:
        continue
      end = off + ((p.data.size + 3) & ~3)
      if out and out[-1][1] == off:
        out[-1][1] = end
      else:
        out.append([off, end])
    return out

  def prepare_grads(self):
    """A trainable param that received no gradient holds the reference's `0.`: make its slice of the arena say so."""
    for p in self.params:
      if p.requires_grad and not p._grad_live:
        p._grad.fill(0.0)


def _arena_for(params):
  key = tuple(id(p) for p in params)
  arena = _ARENAS.get(key)
  if arena is None:
    for p in params:
      if p._pinned:
        raise ValueError('a parameter is already registered with an optimizer over a different parameter list; '
                         'construct optimizers over the same list (e.g. model.parameters()) and use freeze()/unfreeze()')
    arena = _ARENAS[key] = Arena(params)
  return arena


def _off(arr, elems):
  return ctypes.c_void_p(arr.t.data_ptr() + 4 * elems)


class Optimizer:
  """optim.py:5-33."""

  def __init__(self, params, lr):
    self.params = params
    self.lr = lr
    self.arena = _arena_for(params)

  def zero_grad(self, all_members=False):
    if all_members:
      get_graph().zero_grad()
    for param in self.params:
      param.zero_grad()

  def _sync_grads(self):
    """Data parallel: sum the flat gradient buffer over ranks -- peer-memory exchange over NVLink (the summed gradient is
    then `arena.grad_sum`), or one NCCL allreduce in place when peer memory is unavailable; returns grad_scale."""
    a = self.arena
    D.flush_lazy()           # deferred conv outputs somebody still holds are produced before the weights change
    a.prepare_grads()
    if dist.is_initialized() and dist.world_size() > 1:
      if a.exchange is not None:
        a.exchange.run(a.numel, _st())
      else:
        dist.all_reduce_sum(a.grad)
      return 1.0 / dist.world_size()
    return 1.0


class GD(Optimizer):
  """optim.py:36-53: p -= lr * g."""

  def __init__(self, params, lr):
    super().__init__(params, lr)

  def step(self):
    gs = self._sync_grads()
    a = self.arena
    for b, e in a.ranges():
      B.check(B.lib.ngb_sgd_step(_off(a.data, b), _off(a.grad_sum, b), e - b, self.lr, gs, _st()))

  def __repr__(self):
    return f'GD(params={self.params}, lr={self.lr})'

  __str__ = __repr__


class Momentum(Optimizer):
  """optim.py:56-99."""

  def __init__(self, params, lr, beta=0.9):
    super().__init__(params, lr)
    self.beta = beta
    self.init_momentum_grads()

  def init_momentum_grads(self):
    self._m = self.arena.state_buffer('momentum', 'momentum_grad')

  def step(self):
    gs = self._sync_grads()
    a = self.arena
    for b, e in a.ranges():
      B.check(B.lib.ngb_momentum_step(_off(a.data, b), _off(a.grad_sum, b), _off(self._m, b), e - b, self.lr, self.beta, gs, _st()))

  def __repr__(self):
    return f'Momentum(params={self.params}, lr={self.lr}, beta={self.beta})'

  __str__ = __repr__


class RMSProp(Optimizer):
  """optim.py:102-147."""

  def __init__(self, params, lr, beta=0.9, epsilon=1e-8):
    super().__init__(params, lr)
    self.beta = beta
    self.epsilon = epsilon
    self.init_rms_grads()

  def init_rms_grads(self):
    self._r = self.arena.state_buffer('rms', 'rms_grad')

  def step(self):
    gs = self._sync_grads()
    a = self.arena
    for b, e in a.ranges():
      B.check(B.lib.ngb_rmsprop_step(_off(a.data, b), _off(a.grad_sum, b), _off(self._r, b), e - b, self.lr, self.beta,
                                     self.epsilon, gs, _st()))

  def __repr__(self):
    return f'RMSProp(params={self.params}, lr={self.lr}, beta={self.beta}, epsilon={self.epsilon})'

  __str__ = __repr__


class Adam(Optimizer):
  """optim.py:150-208.  `iter` also lives in device memory and is advanced by the step's own launch, so a training
  step captured in a CUDA graph keeps the correct bias correction on every replay."""

  def __init__(self, params, lr, beta1=0.9, beta2=0.999, epsilon=1e-8):
    super().__init__(params, lr)
    self._t_dev = torch().zeros(2, device='cuda', dtype=torch().int32)   # [steps taken, scratch counter of the update kernel]
    self.iter = 0
    self.beta1, self.beta2 = beta1, beta2
    self.epsilon = epsilon
    self.init_adam_grads()

  def init_adam_grads(self):
    self._m = self.arena.state_buffer('momentum', 'momentum_grad')
    self._v = self.arena.state_buffer('rms', 'rms_grad')

  def step(self):
    self.iter += 1
    gs = self._sync_grads()
    a = self.arena
    t_ptr = ctypes.c_void_p(self._t_dev.data_ptr())
    ranges = a.ranges() or [[0, 0]]          # (nothing trainable: the step is still counted, optim.py:169)
    for i, (b, e) in enumerate(ranges):
      tick = 1 if i == len(ranges) - 1 else 2      # the last launch of the step records the new iteration count
      B.check(B.lib.ngb_adam_step_dev(_off(a.data, b), _off(a.grad_sum, b), _off(self._m, b), _off(self._v, b), e - b, self.lr,
                                      self.beta1, self.beta2, self.epsilon, t_ptr, tick, gs, _st()))

  def reset_iter(self):
    self.iter = 0
    self._t_dev.zero_()

  def __repr__(self):
    return f'Adam(params={self.params}, lr={self.lr}, beta1={self.beta1}, beta2={self.beta2}, epsilon={self.epsilon})'

  __str__ = __repr__
