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

import weakref

import numpy as np

# tuple(id(param)) -> Arena, WEAKLY: an arena lives as long as an optimizer over it (Optimizer.arena) or one of its
# parameters (Param._arena) does, and its flat buffers, state and peer-memory exchange are released with the last of them.
_ARENAS = weakref.WeakValueDictionary()


class Arena:
  """Flat storage for a parameter list.  Offsets are padded to 4 floats so every view stays 16-byte aligned."""

  def __init__(self, params):
    params = list(params)
    self._param_refs = [weakref.ref(p) for p in params]      # (params point at the arena, not the other way round)
    self.offsets = []
    n = 0
    for p in params:
      self.offsets.append(n)
      n += (p.data.size + 3) & ~3
    self.numel = n
    self.data = DeviceArray.zeros((max(n, 1),))
    # data parallel on one node: the ranks push their gradients into each other's receive slots over NVLink and sum them
    # inside the update kernel (distributed.PeerExchange); otherwise one NCCL allreduce of the flat buffer, then the update
    self.exchange = dist.PeerExchange.create(n)
    self.grad = DeviceArray.zeros((max(n, 1),))
    self.grad_sum = self.grad
    self.state = {}
    for p, off in zip(params, self.offsets):
      size, shape = p.data.size, p.shape
      view = self.data[off:off + size].reshape(shape)
      view.assign(p.data)
      gview = self.grad[off:off + size].reshape(shape)
      if p._grad_live:
        gview.assign(p._grad)
      p._data, p._pinned = view, True
      p._grad, p._grad_pinned = gview, True
      p._arena = self

  @property
  def params(self):
    return [r() for r in self._param_refs]

  def state_buffer(self, name, attr, only=None):
    """Flat state buffer `name` with a view installed as `attr` on every trainable Param (of `only`, a set of ids, when the
    optimizer covers part of the arena), (re)zeroed for those Params -- the reference's init_*_grads, optim.py:181-189."""
    buf = self.state.get(name)
    fresh = buf is None
    if fresh:
      buf = self.state[name] = DeviceArray.empty((max(self.numel, 1),))
    if fresh or only is None:
      buf.fill(0.0)
    for p, off in zip(self.params, self.offsets):
      if p is None or not p.requires_grad or (only is not None and id(p) not in only):
        continue
      view = buf[off:off + p.data.size].reshape(p.shape)
      if only is not None and not fresh:
        view.fill(0.0)
      setattr(p, attr, view)
    return buf

  def ranges(self):
    """Contiguous [begin, end) element ranges covering the trainable params (one range when nothing is frozen)."""
    out = []
    for p, off in zip(self.params, self.offsets):
      if p is None or not p.requires_grad:
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
      if p is not None and p.requires_grad and not p._grad_live:
        p._grad.fill(0.0)


def _arena_for(params):
  """The arena holding exactly this parameter list, created on first use.  A list that is a SUBSET of an existing arena's
  parameters (e.g. GD(model.parameters()[:2]) after Adam(model.parameters()), which the reference allows) is served by
  that arena: the optimizer then updates only its own parameters' ranges (Optimizer._ranges)."""
  key = tuple(id(p) for p in params)
  arena = _ARENAS.get(key)
  if arena is not None:
    return arena
  owners = {id(getattr(p, '_arena', None)) for p in params if getattr(p, '_pinned', False)}
  if owners:
    host = getattr(params[0], '_arena', None) if getattr(params[0], '_pinned', False) else None
    if len(owners) == 1 and host is not None and all(getattr(p, '_arena', None) is host for p in params):
      return host
    raise ValueError('the parameters are registered with different optimizer arenas (or only some of them are registered); '
                     'construct optimizers over lists drawn from one model.parameters()')
  arena = Arena(params)
  _ARENAS[key] = arena
  return arena


def _off(arr, elems):
  return ctypes.c_void_p(arr.t.data_ptr() + 4 * elems)


class Optimizer:
  """optim.py:5-33."""

  def __init__(self, params, lr):
    self.params = params
    self.lr = lr
    self.arena = _arena_for(params)
    mine = {id(p) for p in params}
    self._whole = all(id(p) in mine for p in self.arena.params if p is not None)

  def _only(self):
    return None if self._whole else {id(p) for p in self.params}

  def _ranges(self):
    """[begin, end) element ranges of the arena this optimizer updates: its own trainable parameters."""
    if self._whole:
      return self.arena.ranges()
    mine = {id(p) for p in self.params}
    out = []
    for p, off in zip(self.arena.params, self.arena.offsets):
      if p is None or id(p) not in mine or not p.requires_grad:
        continue
      end = off + ((p.data.size + 3) & ~3)
      if out and out[-1][1] == off:
        out[-1][1] = end
      else:
        out.append([off, end])
    return out

  # ---------------------------------------------------------------- state for checkpoints (new)
  _STATE_ATTRS = ()       # Param attributes holding this optimizer's per-parameter state
  _HYPER = ('lr',)

  def state_dict(self):
    """Everything needed to resume exactly: hyper-parameters, iteration count, per-parameter state as float64 host arrays
    (same nesting as the reference keeps it: ON the Params, optim.py:188-189 / 196-197)."""
    state = {'class': type(self).__name__, 'hyper': {k: getattr(self, k) for k in self._HYPER}}
    if hasattr(self, 'iter'):
      state['iter'] = int(self.iter)
    state['params'] = [{a: (getattr(p, a).numpy() if p.requires_grad and hasattr(p, a) and hasattr(getattr(p, a), 'numpy') else None)
                        for a in self._STATE_ATTRS} for p in self.params]
    return state

  def load_state_dict(self, state):
    if state.get('class') != type(self).__name__:
      raise ValueError(f"optimizer state of a {state.get('class')} cannot be loaded into a {type(self).__name__}")
    if len(state['params']) != len(self.params):
      raise ValueError('optimizer state does not match the parameter list')
    for k, v in state['hyper'].items():
      setattr(self, k, v)
    for p, st in zip(self.params, state['params']):
      for a, val in st.items():
        if val is not None and hasattr(p, a):
          getattr(p, a).assign(D.DeviceArray.from_host(np.asarray(val).reshape(p.shape)))
    if 'iter' in state and hasattr(self, 'iter'):
      self.iter = state['iter']

  _polls = 0

  @staticmethod
  def _poll_health():
    """Every 16th optimizer step: non-blocking check that no tcgen05 pipeline gave up and no peer exchange failed."""
    Optimizer._polls += 1
    if (Optimizer._polls & 15) == 0:
      B.check(B.lib.ngb_runtime_status())

  def zero_grad(self, all_members=False):
    if all_members:
      get_graph().zero_grad()
    for param in self.params:
      param.zero_grad()

  def _sync_grads(self):
    """Data parallel without peer memory: one NCCL allreduce of the flat gradient buffer in place; returns grad_scale.
    (With peer memory the exchange happens inside the update kernel: _fused_exchange.)"""
    a = self.arena
    self._poll_health()
    D.flush_lazy()           # deferred arrays somebody still holds are produced before the weights change
    a.prepare_grads()
    if dist.is_initialized() and dist.world_size() > 1:
      if a.exchange is None:
        dist.all_reduce_sum(a.grad)
      return 1.0 / dist.world_size()
    return 1.0

  def _fused_exchange(self, kind, s1=None, s2=None, beta1=0.0, beta2=0.0, epsilon=0.0, t_dev=None):
    """Data parallel over NVLink peer memory: ONE kernel per parameter range pushes / gathers the gradients, sums them in
    rank order, scales by 1/world and applies the update (csrc/xgpu.cu).  Returns False when this arena has no exchange."""
    a = self.arena
    if a.exchange is None or not (dist.is_initialized() and dist.world_size() > 1):
      return False
    gs = self._sync_grads()
    ranges = self._ranges()
    if not ranges and kind == 'adam':
      ranges = [[0, 0]]
    for i, (b, e) in enumerate(ranges):
      tick = (1 if i == len(ranges) - 1 else 2) if kind == 'adam' else 0
      a.exchange.update(kind, b, e - b, a.grad.t, a.data.t, s1.t if s1 is not None else None, s2.t if s2 is not None else None,
                        self.lr, beta1, beta2, epsilon, t_dev, tick, gs, _st())
    return True


class GD(Optimizer):
  """optim.py:36-53: p -= lr * g."""

  def __init__(self, params, lr):
    super().__init__(params, lr)

  def step(self):
    if self._fused_exchange('sgd'):
      return
    gs = self._sync_grads()
    a = self.arena
    for b, e in self._ranges():
      B.check(B.lib.ngb_sgd_step(_off(a.data, b), _off(a.grad_sum, b), e - b, self.lr, gs, _st()))

  def __repr__(self):
    return f'GD(params={self.params}, lr={self.lr})'

  __str__ = __repr__


class Momentum(Optimizer):
  """optim.py:56-99."""
  _STATE_ATTRS = ('momentum_grad',)
  _HYPER = ('lr', 'beta')

  def __init__(self, params, lr, beta=0.9):
    super().__init__(params, lr)
    self.beta = beta
    self.init_momentum_grads()

  def init_momentum_grads(self):
    self._m = self.arena.state_buffer('momentum', 'momentum_grad', self._only())

  def step(self):
    if self._fused_exchange('momentum', self._m, None, beta1=self.beta):
      return
    gs = self._sync_grads()
    a = self.arena
    for b, e in self._ranges():
      B.check(B.lib.ngb_momentum_step(_off(a.data, b), _off(a.grad_sum, b), _off(self._m, b), e - b, self.lr, self.beta, gs, _st()))

  def __repr__(self):
    return f'Momentum(params={self.params}, lr={self.lr}, beta={self.beta})'

  __str__ = __repr__


class RMSProp(Optimizer):
  """optim.py:102-147."""
  _STATE_ATTRS = ('rms_grad',)
  _HYPER = ('lr', 'beta', 'epsilon')

  def __init__(self, params, lr, beta=0.9, epsilon=1e-8):
    super().__init__(params, lr)
    self.beta = beta
    self.epsilon = epsilon
    self.init_rms_grads()

  def init_rms_grads(self):
    self._r = self.arena.state_buffer('rms', 'rms_grad', self._only())

  def step(self):
    if self._fused_exchange('rmsprop', self._r, None, beta1=self.beta, epsilon=self.epsilon):
      return
    gs = self._sync_grads()
    a = self.arena
    for b, e in self._ranges():
      B.check(B.lib.ngb_rmsprop_step(_off(a.data, b), _off(a.grad_sum, b), _off(self._r, b), e - b, self.lr, self.beta,
                                     self.epsilon, gs, _st()))

  def __repr__(self):
    return f'RMSProp(params={self.params}, lr={self.lr}, beta={self.beta}, epsilon={self.epsilon})'

  __str__ = __repr__


class Adam(Optimizer):
  """optim.py:150-208.  `iter` also lives in device memory and is advanced by the step's own launch, so a training
  step captured in a CUDA graph keeps the correct bias correction on every replay."""

  _STATE_ATTRS = ('momentum_grad', 'rms_grad')
  _HYPER = ('lr', 'beta1', 'beta2', 'epsilon')

  def __init__(self, params, lr, beta1=0.9, beta2=0.999, epsilon=1e-8):
    super().__init__(params, lr)
    self._t_dev = torch().zeros(2, device='cuda', dtype=torch().int32)   # [steps taken, scratch counter of the update kernel]
    self.beta1, self.beta2 = beta1, beta2
    self.epsilon = epsilon
    self.init_adam_grads()

  def init_adam_grads(self):
    self._m = self.arena.state_buffer('momentum', 'momentum_grad', self._only())
    self._v = self.arena.state_buffer('rms', 'rms_grad', self._only())

  @property
  def iter(self):
    """optim.py:160: iterations taken so far.  The count lives on the device (a captured step advances it on every replay);
    reading it here synchronises, assigning to it (e.g. when resuming from a checkpoint) takes effect on the next step."""
    return int(self._t_dev[0].item())

  @iter.setter
  def iter(self, value):
    self._t_dev[0] = int(value)

  def step(self):
    if self._fused_exchange('adam', self._m, self._v, self.beta1, self.beta2, self.epsilon, self._t_dev):
      return
    gs = self._sync_grads()
    a = self.arena
    t_ptr = ctypes.c_void_p(self._t_dev.data_ptr())
    ranges = self._ranges() or [[0, 0]]          # (nothing trainable: the step is still counted, optim.py:169)
    for i, (b, e) in enumerate(ranges):
      tick = 1 if i == len(ranges) - 1 else 2      # the last launch of the step records the new iteration count
      B.check(B.lib.ngb_adam_step_dev(_off(a.data, b), _off(a.grad_sum, b), _off(self._m, b), _off(self._v, b), e - b, self.lr,
                                      self.beta1, self.beta2, self.epsilon, t_ptr, tick, gs, _st()))

  def reset_iter(self):
    self._t_dev.zero_()

  def __repr__(self):
    return f'Adam(params={self.params}, lr={self.lr}, beta1={self.beta1}, beta2={self.beta2}, epsilon={self.epsilon})'

  __str__ = __repr__
