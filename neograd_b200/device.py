"""DeviceArray -- the value of `Tensor.data` / `Tensor.grad` in neograd_b200: an fp32, C-contiguous array resident in
B200 HBM with the slice of NumPy's surface that neograd's users and its own engine touch
(reference: everything `Tensor.data` is asked to do in neograd/autograd/tensor.py, autograd/utils.py:163-192
(`param.data[idx] += eps`), README.md:119 (`np.where(out.data >= 0.5, ...)`), examples/digits.ipynb
(`np.argmax(probs.data, axis=1)`, `loss.data.tolist()`)).

PyTorch is used here only as the device-memory allocator (its caching allocator makes every op's result buffer a
sub-microsecond allocation and lets a whole training step be captured in a CUDA graph); all arithmetic is done by
libngb200.so through the C ABI.  Reads that must produce host values (`__array__`, comparisons, `tolist`, ...) copy
device -> host and return float64 NumPy data, which is what the reference's float64 arrays would have given.
"""
import ctypes

import numpy as np

from . import _backend as B

_torch = None


def torch():
  global _torch
  if _torch is None:
    import torch as t
    _torch = t
  return _torch


_EMPTY = None      # (torch.empty, device, dtype) resolved once: the per-op result allocation is on the eager hot path


def _alloc(shape):
  global _EMPTY
  if _EMPTY is None:
    B.ensure_runtime()
    t = torch()
    _EMPTY = (t.empty, t.device('cuda', t.cuda.current_device()), t.float32)
  empty, dev, f32 = _EMPTY
  return empty(tuple(shape), device=dev, dtype=f32)


def _ptr(t):
  return ctypes.c_void_p(t.data_ptr())


def _st():
  return ctypes.c_void_p(B.stream())


_HOST_TYPES = (int, float, list, np.ndarray)


class DeviceArray:
  __slots__ = ('t', '_shp')
  __array_priority__ = 1000   # ndarray (op) DeviceArray -> our reflected dunder

  def __init__(self, t):
    self.t = t   # torch.Tensor, cuda, float32, contiguous
    self._shp = None

  # ------------------------------------------------------------------ construction
  @staticmethod
  def from_host(data):
    """int / float / list / ndarray -> device (fp64 -> fp32 rounding happens here, once)."""
    B.ensure_runtime()
    a = np.ascontiguousarray(np.asarray(data, dtype=np.float64).astype(np.float32))
    return DeviceArray(torch().from_numpy(a).to('cuda', non_blocking=False))

  @staticmethod
  def empty(shape):
    return DeviceArray(_alloc(shape))

  @staticmethod
  def full(shape, value):
    out = DeviceArray(_alloc(shape))
    out.fill(value)
    return out

  @staticmethod
  def zeros(shape):
    return DeviceArray.full(shape, 0.0)

  # ------------------------------------------------------------------ metadata
  @property
  def shape(self):
    s = self._shp
    if s is None:
      s = self._shp = tuple(self.t.shape)     # (the buffer of a DeviceArray is never re-shaped in place)
    return s

  @property
  def ndim(self):
    return self.t.dim()

  @property
  def size(self):
    return self.t.numel()

  @property
  def dtype(self):
    return np.dtype(np.float32)

  @property
  def ptr(self):
    return _ptr(self.t)

  def __len__(self):
    if self.t.dim() == 0:
      raise TypeError('len() of unsized object')
    return self.t.shape[0]

  # ------------------------------------------------------------------ host reads
  def numpy(self):
    """device -> host, float64 (synchronises the stream)."""
    return self.t.detach().cpu().numpy().astype(np.float64)

  def __array__(self, dtype=None, copy=None):
    a = self.numpy()
    return a if dtype is None else a.astype(dtype)

  def tolist(self):
    return self.numpy().tolist()

  def item(self):
    return float(self.numpy().reshape(-1)[0]) if self.size == 1 else self.numpy().item()

  def __float__(self):
    return self.item()

  def __repr__(self):
    return 'DeviceArray(' + repr(self.numpy())[len('array('):]

  __str__ = lambda self: str(self.numpy())

  def astype(self, dtype):
    if dtype in (float, np.float64, np.float32, 'float', 'float32', 'float64'):
      return self
    return self.numpy().astype(dtype)

  # comparisons / predicates give host boolean arrays (they steer host control flow)
  def __ge__(self, o): return self.numpy() >= np.asarray(o)
  def __gt__(self, o): return self.numpy() > np.asarray(o)
  def __le__(self, o): return self.numpy() <= np.asarray(o)
  def __lt__(self, o): return self.numpy() < np.asarray(o)
  def __eq__(self, o): return self.numpy() == np.asarray(o)
  def __ne__(self, o): return self.numpy() != np.asarray(o)
  __hash__ = None

  # ------------------------------------------------------------------ views / copies
  def reshape(self, *shape):
    if len(shape) == 1 and not isinstance(shape[0], (int, np.integer)):
      shape = tuple(shape[0])
    return DeviceArray(self.t.view(tuple(int(s) for s in shape)))   # metadata only, same storage

  def flatten(self):
    return self.reshape(self.size)

  def copy(self):
    out = DeviceArray(_alloc(self.shape))
    out.assign(self)
    return out

  @property
  def T(self):
    return self.transpose()

  def transpose(self):
    """np.transpose semantics (all axes reversed); materialised (results are always C-contiguous)."""
    if self.ndim < 2:
      return self
    out = DeviceArray(_alloc(self.shape[::-1]))
    if self.size == 0:
      return out
    if self.ndim == 2:
      B.check(B.lib.ngb_transpose2d(self.ptr, out.ptr, self.shape[0], self.shape[1], _st()))
    else:
      sh, r = B.shape_arg(self.shape)
      perm = (ctypes.c_int * r)(*range(r - 1, -1, -1))
      B.check(B.lib.ngb_permute(self.ptr, out.ptr, sh, perm, r, _st()))
    return out

  def __getitem__(self, idx):
    """int / slice on axis 0 -> view of the same storage (tensor.py:342-355 supports exactly these); anything else is
    served from a host copy."""
    if isinstance(idx, (int, np.integer)) and self.ndim >= 1:
      return DeviceArray(self.t[int(idx)])
    if isinstance(idx, slice) and self.ndim >= 1 and idx.step in (None, 1):
      return DeviceArray(self.t[idx])
    if isinstance(idx, tuple) and all(isinstance(i, (int, np.integer)) for i in idx) and len(idx) == self.ndim:
      return float(self.t[tuple(int(i) for i in idx)].item())
    return self.numpy()[idx]

  def __setitem__(self, idx, value):
    """Element / slice pokes (grad_check's `param.data[idx] += eps`, utils.py:185-190).  Off the hot path: host round trip."""
    a = self.numpy()
    a[idx] = np.asarray(value)
    self.assign(DeviceArray.from_host(a))

  # ------------------------------------------------------------------ in-place kernels
  def fill(self, value):
    if self.size:
      flush_readers_of(self)
      B.check(B.lib.ngb_fill(self.ptr, float(value), self.size, _st()))
    return self

  def assign(self, other):
    """self[...] = other (same number of elements)."""
    if not isinstance(other, DeviceArray):
      other = DeviceArray.from_host(np.broadcast_to(np.asarray(other, dtype=np.float64), self.shape))
    if other.size != self.size:
      raise ValueError(f'could not assign array of shape {other.shape} to shape {self.shape}')
    if self.size and other.t.data_ptr() != self.t.data_ptr():
      flush_readers_of(self)
      B.check(B.lib.ngb_copy(self.ptr, other.ptr, self.size, _st()))
    return self

  def add_(self, other, alpha=1.0):
    """self += alpha * other  (tensor.py:301 grad accumulation)."""
    if self.size:
      flush_readers_of(self)
      B.check(B.lib.ngb_axpy(self.ptr, other.ptr, float(alpha), self.size, _st()))
    return self

  # ------------------------------------------------------------------ arithmetic (NumPy broadcasting, on device)
  def _binary(self, op, other, reflected=False):
    a, b = (other, self) if reflected else (self, other)
    return binary(op, a, b)

  def __add__(self, o): return self._binary(B.ADD, o)
  def __radd__(self, o): return self._binary(B.ADD, o, True)
  def __sub__(self, o): return self._binary(B.SUB, o)
  def __rsub__(self, o): return self._binary(B.SUB, o, True)
  def __mul__(self, o): return self._binary(B.MUL, o)
  def __rmul__(self, o): return self._binary(B.MUL, o, True)
  def __truediv__(self, o): return self._binary(B.DIV, o)
  def __rtruediv__(self, o): return self._binary(B.DIV, o, True)
  def __pow__(self, o): return self._binary(B.POW, o)
  def __rpow__(self, o): return self._binary(B.POW, o, True)
  def __neg__(self): return self._binary(B.MUL, -1.0)
  def __pos__(self): return self

  def __iadd__(self, o):
    return self.assign(self._binary(B.ADD, o)) if not isinstance(o, DeviceArray) or o.shape != self.shape else self.add_(o)

  def __isub__(self, o):
    return self.assign(self._binary(B.SUB, o)) if not isinstance(o, DeviceArray) or o.shape != self.shape else self.add_(o, -1.0)

  def __imul__(self, o): return self.assign(self._binary(B.MUL, o))
  def __itruediv__(self, o): return self.assign(self._binary(B.DIV, o))

  def sum(self, axis=None, keepdims=False):
    return reduce_sum(self, axis, keepdims)


def _operand(x):
  """-> (DeviceArray or None, scalar immediate, shape)"""
  if isinstance(x, DeviceArray):
    return x, 0.0, x.shape
  if isinstance(x, (int, float, np.integer, np.floating)):
    return None, float(x), ()
  d = DeviceArray.from_host(x)
  return d, 0.0, d.shape


def binary(op, a, b):
  """out = a (op) b with NumPy broadcasting (basics.py:20, 63, 107, 150, 313).  Python scalars travel as immediates."""
  A, a_s, ash = _operand(a)
  Bm, b_s, bsh = _operand(b)
  try:
    out_shape = np.broadcast_shapes(ash, bsh)
  except ValueError:
    raise ValueError(f'operands could not be broadcast together with shapes {ash} {bsh}') from None
  out = DeviceArray(_alloc(out_shape))
  sa, ra = B.shape_arg(ash)
  sb, rb = B.shape_arg(bsh)
  B.check(B.lib.ngb_ew_binary_fwd(op, A.ptr if A is not None else None, a_s, sa, ra,
                                  Bm.ptr if Bm is not None else None, b_s, sb, rb, out.ptr, _st()))
  return out


def broadcast_to(x, shape):
  """np.broadcast_to(x, shape), materialised (NGB_SECOND with a shape-only first operand: nothing but x is read)."""
  out = DeviceArray(_alloc(shape))
  sa, ra = B.shape_arg(shape)
  sb, rb = B.shape_arg(x.shape)
  B.check(B.lib.ngb_ew_binary_fwd(B.SECOND, None, 0.0, sa, ra, x.ptr, 0.0, sb, rb, out.ptr, _st()))
  return out


def binary_bwd(op, side, ug, a, b, grad_out=None, accumulate=False):
  """Fused local-gradient x upstream-gradient + unbroadcast (utils.py:34-78) [+ accumulate].  Returns the operand-shaped grad."""
  A, a_s, ash = _operand(a)
  Bm, b_s, bsh = _operand(b)
  shape = (ash, bsh)[side]
  if grad_out is None:
    grad_out = DeviceArray(_alloc(shape))
    accumulate = False
  sa, ra = B.shape_arg(ash)
  sb, rb = B.shape_arg(bsh)
  B.check(B.lib.ngb_ew_binary_bwd(op, side, ug.ptr, A.ptr if A is not None else None, a_s, sa, ra,
                                  Bm.ptr if Bm is not None else None, b_s, sb, rb, grad_out.ptr, int(accumulate), _st()))
  return grad_out


class LazyArray(DeviceArray):
  """A DeviceArray whose buffer is produced on first use (`thunk() -> DeviceArray`).

  Used where the next op can consume the PRODUCER's inputs directly with a fused kernel and the array itself would
  only be written to HBM and read back once (ReLU in front of MaxPool; MaxPool's input gradient in front of ReLU's
  backward).  `tag` / `args` tell such a consumer what the array stands for; anything else that touches `.t` / `.ptr`
  simply triggers the producing kernel, so the rest of the engine treats it as an ordinary array.  Metadata (shape,
  ndim, size) never materialises it."""
  __slots__ = ('_shape', '_thunk', 'tag', 'args', '__weakref__')

  def __init__(self, shape, thunk, tag, args):
    self._shape = tuple(int(s) for s in shape)
    self._thunk = thunk
    self.tag = tag
    self.args = args

  @property
  def pending(self):
    return self._thunk is not None

  @property
  def t(self):
    if self._thunk is not None:
      thunk, self._thunk = self._thunk, None
      DeviceArray.t.__set__(self, thunk().t)
      self.args = None
    return DeviceArray.t.__get__(self)

  @t.setter
  def t(self, value):
    self._thunk, self.args = None, None
    DeviceArray.t.__set__(self, value)

  @property
  def shape(self):
    return self._shape

  @property
  def ndim(self):
    return len(self._shape)

  @property
  def size(self):
    n = 1
    for s in self._shape:
      n *= s
    return n

  def __len__(self):
    if not self._shape:
      raise TypeError('len() of unsized object')
    return self._shape[0]


# Deferred arrays read their producer's inputs when they are finally produced, not at forward time.  To keep that
# invisible (the reference's results hold forward-time values) every deferred array is registered with the buffers it will
# read: the optimizers flush all pending ones before they touch the parameters (flush_lazy), and every in-place write to a
# DeviceArray first produces the pending arrays that read the written range (flush_readers_of).
_pending = []          # [(weakref to the LazyArray, ((data_ptr, nbytes), ...))]


def _extent(arr):
  """(address, bytes) of a materialised DeviceArray; None for one that is itself still deferred (its own entry covers it)"""
  if arr is None or (isinstance(arr, LazyArray) and arr.pending):
    return None
  t = arr.t
  return (t.data_ptr(), t.numel() * 4)


def register_lazy(arr, sources=()):
  import weakref
  if len(_pending) >= 256:     # forward-only loops never reach an optimizer: keep the list to the live ones
    _pending[:] = [e for e in _pending if e[0]() is not None and e[0]().pending]
  _pending.append((weakref.ref(arr), tuple(x for x in (_extent(s) for s in sources) if x is not None)))


def flush_lazy():
  if not _pending:
    return
  entries, _pending[:] = list(_pending), []
  for r, _ in entries:
    arr = r()
    if arr is not None and arr.pending:
      arr.t


def flush_readers_of(dst):
  """Produce the pending deferred arrays that will read any part of `dst`'s buffer (called before in-place writes)."""
  if not _pending:
    return
  ext = _extent(dst)
  if ext is None:
    return
  lo, hi = ext[0], ext[0] + ext[1]
  hit = False
  for r, srcs in _pending:
    for (p, n) in srcs:
      if p < hi and lo < p + n:
        hit = True
        arr = r()
        if arr is not None and arr.pending:
          arr.t
        break
  if hit:
    _pending[:] = [e for e in _pending if e[0]() is not None and e[0]().pending]


class ConstScalar(LazyArray):
  """A 0-d array whose value the host knows (the `1.` that Tensor.backward() seeds, tensor.py:44): consumers that can use
  the value as an immediate never make it exist on the device; anything else materialises it with one fill."""
  __slots__ = ('value',)

  def __init__(self, value):
    value = float(value)
    LazyArray.__init__(self, (), lambda: DeviceArray.full((), value), 'const', None)
    self.value = value


def const_value(arr):
  """the host-known value of a ConstScalar, else None"""
  return arr.value if isinstance(arr, ConstScalar) else None


def opt_ptr(arr):
  """Pointer argument for an operand the kernel does not read: NULL instead of producing a still-deferred array."""
  if arr is None or (isinstance(arr, LazyArray) and arr.pending):
    return None
  return arr.ptr


class PoolReluGrad:
  """What a deferred ReLU+MaxPool(2x2, stride 2) backward stands for: the pooled upstream gradient `pug`, the pooling
  argmax indices `idx`, the pool's output `ypool` and the ReLU's input `x`.  The convolution's weight gradient consumes
  these directly (ngb_conv_wgrad_pooled); `gm()` is the compact masked gradient pug * relu'(x at the argmax) for its
  bias gradient, computed once (ngb_maxpool_relu_mask) and cached -- by its only consumer, on that consumer's stream."""
  __slots__ = ('pug', 'idx', 'ypool', 'x', 'outer', 'H', 'W', '_gm')

  def __init__(self, pug, idx, ypool, x, outer, H, W):
    self.pug, self.idx, self.ypool, self.x, self.outer, self.H, self.W = pug, idx, ypool, x, outer, H, W
    self._gm = None

  def gm(self):
    if self._gm is None:
      out = DeviceArray(_alloc(self.pug.shape))
      B.check(B.lib.ngb_maxpool_relu_mask(self.pug.ptr, self.ypool.ptr, self.x.ptr, out.ptr, self.outer, self.H, self.W, _st()))
      self._gm = out
    return self._gm

  def prep(self):
    """Nothing is shared between the consumers (autograd/tensor.py:_issue_leaf_grads calls this before the lanes start)."""


import os as _os
POOLED_CONV_GRADS = _os.environ.get('NGB_NO_POOLED') != '1'   # conv weight / bias gradients straight from a pooled gradient
POOLED_CONV_DGRAD = _os.environ.get('NGB_NO_POOLED_DGRAD') != '1'
LAZY_FUSION = True   # ReLU -> MaxPool pairs run as fused kernels (set False to materialise every op's output eagerly)
FUSED_CONV_POOL = _os.environ.get('NGB_NO_FUSED_CONV_POOL') != '1'   # first-layer conv -> ReLU -> MaxPool blocks in one kernel
FUSED_LINEAR = _os.environ.get('NGB_NO_FUSED_LINEAR') != '1'   # Linear (+ activation / + SoftmaxCE) as one kernel per pass


def unary(op, x, alpha=0.0):
  out = DeviceArray(_alloc(x.shape))
  B.check(B.lib.ngb_ew_unary_fwd(op, float(alpha), x.ptr, out.ptr, x.size, _st()))
  return out


def unary_bwd(op, x, ug, alpha=0.0):
  out = DeviceArray(_alloc(x.shape))
  B.check(B.lib.ngb_ew_unary_bwd(op, float(alpha), x.ptr, ug.ptr, out.ptr, x.size, 0, _st()))
  return out


def reduce_sum(x, axis=None, keepdims=False, out=None, accumulate=False):
  """np.sum(x, axis) (basics.py:369, utils.py:75)."""
  nd = x.ndim
  if axis is None:
    axes = tuple(range(nd))
  elif isinstance(axis, (int, np.integer)):
    axes = (int(axis) % nd if nd else 0,)
  else:
    axes = tuple(int(a) % nd for a in axis)
  mask = 0
  for a in axes:
    mask |= 1 << a
  out_shape = tuple((1 if (mask >> i) & 1 else s) for i, s in enumerate(x.shape)) if keepdims else \
      tuple(s for i, s in enumerate(x.shape) if not (mask >> i) & 1)
  if out is None:
    out = DeviceArray(_alloc(out_shape))
    accumulate = False
  sh, r = B.shape_arg(x.shape)
  B.check(B.lib.ngb_reduce_sum(x.ptr, sh, r, mask, out.ptr, int(accumulate), _st()))
  return out


def gemm(a, b, ta=False, tb=False, out=None, accumulate=False):
  """op(a) @ op(b) for 2-D device arrays (np.dot at basics.py:194, 206, 207)."""
  M, K = (a.shape[1], a.shape[0]) if ta else a.shape
  K2, N = (b.shape[1], b.shape[0]) if tb else b.shape
  if K != K2:
    raise ValueError(f'shapes {a.shape} and {b.shape} not aligned: {K} (dim 1) != {K2} (dim 0)')
  if out is None:
    out = DeviceArray(_alloc((M, N)))
    accumulate = False
  B.check(B.lib.ngb_gemm(B.GEMM_AUTO, int(ta), int(tb), M, N, K, a.ptr, a.shape[1], b.ptr, b.shape[1], out.ptr, N,
                         int(accumulate), _st()))
  return out


def as_device(data):
  return data if isinstance(data, DeviceArray) else DeviceArray.from_host(data)


def synchronize():
  torch().cuda.synchronize()
