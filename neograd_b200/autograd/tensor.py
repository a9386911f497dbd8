"""Tensor: neograd's user-facing array-with-gradient, with `data` and `grad` resident in B200 HBM
(reference: neograd/autograd/tensor.py).  The public surface is the reference's; the differences are internal:

  * `data` / `grad` are `DeviceArray`s (fp32, device) instead of float64 ndarrays;
  * Python-scalar operands stay scalars (kernel immediates) until something asks for their `.data`;
  * gradients of the built-in ops are written by ONE kernel that fuses local-gradient x upstream-gradient, the
    `unbroadcast` reduction (utils.py:34-78), the reshape (tensor.py:99) and the `+=` accumulation (tensor.py:301)
    straight into the gradient buffer (see `GradFn.into`); user-defined Operations that return a plain array from
    their grad_fn take the reference's generic path;
  * a parameter's `data` / `grad` may be pinned views into an optimizer's flat arena (nn/optim.py here): assigning to
    them copies into place so the arena (one allreduce, one fused optimizer launch) stays valid.
"""
import numpy as np

from .._backend import lanes as _lanes
from ..device import ConstScalar, DeviceArray
from .utils import get_graph, process_data, unbroadcast_data


class GradFn:
  """A grad_fn the engine can fuse.  `into(ug, dst, accumulate)` must write (accumulate=False) or add
  (accumulate=True) the operand-shaped gradient into `dst`; `dst=None` means allocate.  Calling the object like the
  reference's lambdas (`grad_fn(ug) -> array`) still works."""
  __slots__ = ('into', 'alias', 'lazy', 'lazy_ok', 'joint')

  def __init__(self, into, alias=None, lazy=None, lazy_ok=False, joint=None):
    self.into = into
    # joint = (other_leaf, fn, usable[, primary]): fn(ug, dst, accumulate, other_dst, other_accumulate) produces the PRIMARY
    # operand's gradient (dst) AND its partner's (other_dst) in one go (weight + bias from one pass over the upstream
    # gradient) when usable(ug).  primary (default True) says whether this operand is fn's first; a partner may carry the
    # mirrored entry with primary = False so that whichever leaf is visited first triggers the kernel.  Used by the backward
    # lanes when both are leaves waiting on the same upstream gradient, and by the single-stream visit otherwise
    self.joint = joint
    # lazy_ok: `into` understands a still-deferred upstream gradient (device.LazyArray) and need not have it materialised
    self.lazy_ok = lazy_ok
    # lazy(ug) -> a device.LazyArray standing for this operand's gradient (produced on first use), or None.  Same
    # conditions as `alias`; lets the operand's own backward fuse with the producing kernel (MaxPool -> ReLU).
    self.lazy = lazy
    # alias(ug) -> a view of ug that IS this operand's gradient (Add / reshape pass-through), or None.  The engine uses
    # it instead of a copy kernel when the operand has a single consumer and no gradient yet (buffers then become
    # copy-on-write, see Tensor._grad_shared).
    self.alias = alias

  def __call__(self, ug):
    return self.into(ug, None, False)


class Tensor:
  def __init__(self, data, requires_grad=False, requires_broadcasting=True):
    self._data = None
    self._scalar = None
    self._pinned = False        # data buffer is a view into an optimizer arena: assignments copy into place
    self._grad = None           # gradient buffer (may be a pinned arena view)
    self._grad_live = False     # whether _grad currently holds the gradient (False <=> reference's `0.`)
    self._grad_pinned = False
    self._grad_shared = False   # _grad aliases another tensor's gradient buffer: copy before accumulating in place
    self.data = data
    self.requires_grad = requires_grad
    self.requires_broadcasting = requires_broadcasting
    self.grad_fn = None

  # ------------------------------------------------------------------ data
  @property
  def data(self):
    if self._data is None:                       # lazy upload of a Python-scalar operand
      self._data = DeviceArray.full((), self._scalar)
    return self._data

  @data.setter
  def data(self, data):
    """tensor.py:312-322: every assignment is coerced by process_data (TypeError on unsupported types)."""
    if type(data) in (int, float) and self._data is None and not self._pinned:
      self._scalar = float(data)
      return
    new = process_data(data)
    self._scalar = None
    if self._pinned and self._data is not None:
      if new is not self._data:
        if new.shape != self._data.shape:
          raise ValueError(f'cannot assign data of shape {new.shape} to a parameter of shape {self._data.shape} '
                           'that is registered with an optimizer')
        self._data.assign(new)
    else:
      self._data = new

  @property
  def shape(self):
    return () if self._data is None else self._data.shape

  # ------------------------------------------------------------------ grad
  @property
  def grad(self):
    if self._grad_live:
      return self._grad
    return 0. if self.requires_grad else None     # tensor.py:36, 42

  @grad.setter
  def grad(self, value):
    if value is None or (isinstance(value, (int, float)) and value == 0):
      self._grad_live = False
      if not self._grad_pinned:
        self._grad = None
      return
    if isinstance(value, (int, float)):
      value = DeviceArray.full(self.shape, value)
    value = process_data(value)
    if self._grad_pinned:
      self._grad.assign(value)
    else:
      self._grad = value
    self._grad_live = True

  def zero_grad(self):
    """tensor.py:39-42."""
    self._grad_live = False
    self._grad_shared = False
    if not self._grad_pinned:
      self._grad = None

  def accumulate_grad(self, grad):
    """tensor.py:291-301: `self.grad += grad` (the first accumulation replaces the Python 0.)."""
    if self._grad_live:
      self._unshare_grad()
      self._grad.add_(grad)
    elif self._grad_pinned:
      self._grad.assign(grad)
      self._grad_live = True
    else:
      self._grad = grad.copy()   # the producer may hand back the upstream buffer itself (Add): never alias it
      self._grad_live = True

  def set_grad_fn(self, grad_fn):
    """tensor.py:106-114."""
    self.grad_fn = grad_fn if self.requires_grad else None

  # ------------------------------------------------------------------ backward
  def backward(self, upper_grad=1., retain_graph=False):
    """tensor.py:44-73."""
    if not self.requires_grad:
      raise ValueError("Only tensors who requires_grad can call backward")
    graph = get_graph()
    if type(upper_grad) in (int, float):
      if self.shape != ():
        raise ValueError("Shapes of grad and Tensor data must match!")
      upper_grad = ConstScalar(upper_grad)      # stays a host-known immediate until a kernel needs it in memory
      if not self._grad_live and not self._grad_pinned:
        self._grad, self._grad_live, self._grad_shared = upper_grad, True, False   # our own fresh buffer: no copy needed
      else:
        self.accumulate_grad(upper_grad)
    else:
      upper_grad = process_data(upper_grad)
      if self.shape != upper_grad.shape:
        raise ValueError("Shapes of grad and Tensor data must match!")
      self.accumulate_grad(upper_grad)
    node = graph.get_node(self)
    try:
      node.backward(retain_graph)
    finally:
      _lanes.join()               # weight / bias gradients issued on side streams are complete for whatever follows
    if not retain_graph:
      graph.reset_graph()

  def _backward(self, node, retain_graph, calculate_grads=True):
    """tensor.py:75-104.  For every child (an op that consumed this tensor): install the grad_fns, run this operand's
    grad_fn on the child's gradient, unbroadcast, reshape, accumulate.

    Backward lanes (new): once this tensor's own gradient is final, the gradients of the LEAF operands of the op that
    produced it (weights, biases) are issued right away on side streams -- nothing else in the pass consumes them, so
    they run beside the data-gradient chain instead of after it.  The leaf's own visit then finds that contribution
    already accumulated."""
    graph = get_graph()
    for child in node.children:
      if self.requires_grad and calculate_grads:
        if child.lane_done is not None and child.lane_done.get(id(self)) == 'lane':
          del child.lane_done[id(self)]              # accumulated on a backward lane when the child's gradient became final
        else:
          if child.defer_leaves and not node.parents:
            if child.lane_done is None:
              child.lane_done = {}
            child.lane_done[id(self)] = 'main'       # visited before the deferred issue: that one must skip this leaf
          child.backward_fn(*[n.tens for n in child.parents])
          ct = child.tens
          upper_grad = ct._grad if ct._grad_live else DeviceArray.zeros(ct.shape)
          gf = self.grad_fn
          if isinstance(gf, GradFn):
            view = None
            single = not self._grad_live and not self._grad_pinned and len(node.children) == 1
            if gf.alias is not None and single:
              view = gf.alias(upper_grad)
            deferred = gf.lazy(upper_grad) if (view is None and gf.lazy is not None and single and node.parents) else None
            if view is not None:                     # pass-through gradient: share the buffer, no copy kernel
              self._grad, self._grad_live, self._grad_shared = view, True, True
              ct._grad_shared = True
            elif deferred is not None:               # produced when (and if) somebody reads it
              self._grad, self._grad_live, self._grad_shared = deferred, True, False
            else:
              self._ensure_grad_buffer()
              if self._grad_live:
                self._unshare_grad()
              other = self._joint_partner(gf, child, node, upper_grad)
              if other is not None:
                # one kernel yields this leaf's gradient AND its partner's (weight + bias from one pass over the upstream
                # gradient): the partner's own visit then finds its contribution from this op already accumulated
                other._ensure_grad_buffer()
                if other._grad_live:
                  other._unshare_grad()
                mine_first = gf.joint[3] if len(gf.joint) > 3 else True
                if mine_first:
                  gf.joint[1](upper_grad, self._grad, self._grad_live, other._grad, other._grad_live)
                else:
                  gf.joint[1](upper_grad, other._grad, other._grad_live, self._grad, self._grad_live)
                other._grad_live = True
                if child.lane_done is None:
                  child.lane_done = {}
                child.lane_done[id(other)] = 'lane'
              else:
                gf.into(upper_grad, self._grad, self._grad_live)
              self._grad_live = True
            if child.defer_leaves:                   # the op's data gradient is on its way: now its weight / bias gradients
              child.defer_leaves = False
              ct._issue_leaf_grads(child)
          else:
            grad = gf(upper_grad)
            grad = unbroadcast_data(grad, self.shape, child.parent_broadcast_shape)
            grad = process_data(grad) if not isinstance(grad, DeviceArray) else grad
            self.accumulate_grad(grad.reshape(self.shape))
      if not retain_graph and child.are_parents_visited():
        graph.remove_tensor(child.tens)
    if node.parents and self.requires_grad and _lanes.active():
      # If the op also has a data operand that needs a gradient, that kernel is the one the rest of the pass waits on:
      # it is issued first (at the operand's visit) and the leaf gradients follow it, so they overlap the NEXT kernels
      # of the chain instead of competing with this one for the SMs.
      if any(p.parents and p.tens.requires_grad for p in node.parents):
        node.defer_leaves = True
      else:
        self._issue_leaf_grads(node, corun=False)   # end of a chain: nothing of this op runs on the main stream
    if not retain_graph and node.are_parents_visited():
      graph.remove_tensor(node.tens)

  def _joint_partner(self, gf, child, node, upper_grad):
    """The leaf whose gradient from `child` can come out of the same kernel as this leaf's (GradFn.joint), if it is still
    pending there; None otherwise.  Only for leaves visited on the main stream (the lanes pair leaves themselves)."""
    if gf.joint is None or node.parents or _lanes.active():
      return None
    other = gf.joint[0]
    if other is self or not other.requires_grad or not isinstance(other.grad_fn, GradFn):
      return None
    onode = None
    for p in child.parents:
      if p.tens is other:
        onode = p
        break
    if onode is None or onode.parents or onode.visited:       # (visited: its contribution from this op is already in)
      return None
    if child.lane_done is not None and id(other) in child.lane_done:
      return None
    return other if gf.joint[2](upper_grad) else None

  def _issue_leaf_grads(self, node, corun=True):
    leaves = [p for p in node.parents if not p.parents and p.tens.requires_grad and p.tens is not self]
    if not leaves:
      return
    node.backward_fn(*[n.tens for n in node.parents])
    upper_grad = self._grad if self._grad_live else DeviceArray.zeros(self.shape)
    if getattr(upper_grad, 'pending', False):
      # a deferred gradient is produced HERE, on the main stream, once -- or, when every consumer below understands its
      # deferred form, only that form's shared pre-computation (the lanes must not race to create it)
      prep = getattr(upper_grad.args, 'prep', None)
      if prep is not None and all(isinstance(p.tens.grad_fn, GradFn) and p.tens.grad_fn.lazy_ok for p in leaves):
        prep()
      else:
        upper_grad.t
    ready = None
    # End of a chain (corun False): nothing else is left for the main stream, so the largest leaf's gradient (the weight
    # gradient) runs THERE -- on the high-priority capture stream it is not queued behind the other layers' side-lane
    # kernels, which became ready at the same moment -- and the small ones go to the lanes.
    on_main = None
    if not corun:
      on_main = max(leaves, key=lambda p: p.tens.data.size).tens
      leaves = [p for p in leaves if p.tens is not on_main] + [p for p in leaves if p.tens is on_main]
    def pending(t):
      return isinstance(t.grad_fn, GradFn) and not (node.lane_done is not None and id(t) in node.lane_done)

    # leaves whose gradient comes out of another leaf's joint kernel (conv bias beside the conv weight)
    by_tensor = {id(p.tens): p.tens for p in leaves}
    partner_of = {}
    for p in leaves:
      gf = p.tens.grad_fn
      if isinstance(gf, GradFn) and gf.joint is not None and pending(p.tens) and (len(gf.joint) < 4 or gf.joint[3]):
        other = by_tensor.get(id(gf.joint[0]))
        if (other is not None and other is not p.tens and pending(other) and _lanes.can_share(p.tens, other)
            and gf.joint[2](upper_grad)):
          partner_of[id(p.tens)] = other
    produced_by_partner = {id(o) for o in partner_of.values()}

    def prepare(t):
      t._ensure_grad_buffer()
      if t._grad_live:
        t._unshare_grad()
      return t._grad, t._grad_live

    def mark(t):
      t._grad_live = True
      if node.lane_done is None:
        node.lane_done = {}
      node.lane_done[id(t)] = 'lane'

    for p in leaves:
      t = p.tens
      gf = t.grad_fn
      if node.lane_done is not None and node.lane_done.get(id(t)) == 'main':
        del node.lane_done[id(t)]                    # already accumulated at the leaf's own visit
        continue
      if not isinstance(gf, GradFn) or (node.lane_done is not None and id(t) in node.lane_done):
        continue                                     # user-defined grad_fn: the reference's path at the leaf's own visit
      if id(t) in produced_by_partner:
        continue
      if ready is None:
        ready = _lanes.mark_ready()                  # main stream: this tensor's gradient is complete
      dst, live = prepare(t)
      other = partner_of.get(id(t))
      if other is not None:
        odst, olive = prepare(other)
        fn = lambda gf=gf, dst=dst, live=live, odst=odst, olive=olive: gf.joint[1](upper_grad, dst, live, odst, olive)
      else:
        fn = lambda gf=gf, dst=dst, live=live: gf.into(upper_grad, dst, live)
      if t is on_main:
        fn()                                         # (issued last: the lanes' waits on `ready` are already queued)
      else:
        _lanes.run(t, fn, (upper_grad, gf, node), ready, corun, also=other)
      mark(t)
      if other is not None:
        mark(other)

  def _ensure_grad_buffer(self):
    if self._grad is None:
      self._grad = DeviceArray.empty(self.shape)
      self._grad_live = False
      self._grad_shared = False

  def _unshare_grad(self):
    if self._grad_shared:
      self._grad = self._grad.copy()
      self._grad_shared = False

  # ------------------------------------------------------------------ operators (tensor.py:116-289)
  def __add__(self, other):
    from .ops import add
    return add(self, other)

  def __radd__(self, other):
    from .ops import add
    return add(other, self)

  def __sub__(self, other):
    from .ops import sub
    return sub(self, other)

  def __rsub__(self, other):
    from .ops import sub
    return sub(other, self)

  def __mul__(self, other):
    from .ops import mul
    return mul(self, other)

  def __rmul__(self, other):
    from .ops import mul
    return mul(other, self)

  def __truediv__(self, other):
    from .ops import div
    return div(self, other)

  def __rtruediv__(self, other):
    from .ops import div
    return div(other, self)

  def __pow__(self, other):
    from .ops import pow as _pow
    return _pow(self, other)

  def __rpow__(self, other):
    from .ops import pow as _pow
    return _pow(other, self)

  def __pos__(self):
    from .ops import mul
    return mul(1, self)          # tensor.py:226-232

  def __neg__(self):
    from .ops import mul
    return mul(-1, self)         # tensor.py:234-240

  def dot(self, other):
    from .ops import dot
    return dot(self, other)

  def sum(self, axis=None):
    from .ops import sum as _sum
    return _sum(self, axis)

  def exp(self):
    from .ops import exp
    return exp(self)

  def flatten(self):
    from .ops import flatten
    return flatten(self)

  def reshape(self, new_shape):
    from .ops import reshape
    return reshape(self, new_shape)

  @property
  def T(self):
    from .ops import transpose
    return transpose(self)

  def __getitem__(self, *indices):
    """tensor.py:342-355: one int or slice on the leading axis -> an UNTRACKED tensor (a view of the same HBM)."""
    index = indices[0]
    if isinstance(index, (int, slice)):
      return Tensor(self.data[index], requires_grad=self.requires_grad)
    raise TypeError(f"Expected index of int or slice instead got {type(index)}")

  def __repr__(self):
    return f'Tensor({self.data.numpy()},\n requires_grad={self.requires_grad},\n requires_broadcasting={self.requires_broadcasting})'

  def __str__(self):
    return f'Tensor( {self.data.numpy()},\n requires_grad={self.requires_grad},\n grad_fn={self.grad_fn},\n shape={self.shape},\n requires_broadcasting={self.requires_broadcasting} )\n'
