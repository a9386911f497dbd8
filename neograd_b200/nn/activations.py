"""Activations (reference: neograd/nn/activations.py).  Each is a Layer and an Operation at once, as in the reference;
forward and backward are single elementwise kernels (backward recomputes sigmoid / tanh from the saved input inside
the kernel instead of keeping a second activation-sized buffer)."""
from .layers import Layer
from ..autograd.ops.operation import Operation
from ..autograd.ops.basics import _Unary
from ..autograd.tensor import GradFn
from .. import _backend as B
from .. import device as D


class ReLU(Layer, _Unary):
  """activations.py:7-37: max(0, x); gradient mask x >= 0 (so 1 at exactly 0, line 31)."""
  OP = B.RELU

  def __repr__(self): return 'ReLU()'
  def __str__(self): return 'ReLU'


class Sigmoid(Layer, _Unary):
  """activations.py:41-69."""
  OP = B.SIGMOID

  def __repr__(self): return 'Sigmoid()'
  def __str__(self): return 'Sigmoid'


class Tanh(Layer, _Unary):
  """activations.py:73-101."""
  OP = B.TANH

  def __repr__(self): return 'Tanh()'
  def __str__(self): return 'Tanh'


class LeakyReLU(Layer, _Unary):
  """activations.py:184-222: where(x >= 0, x, leak * x), leak = 0.01 by default."""
  OP = B.LEAKY_RELU

  def __init__(self, leak=0.01):
    self.leak = leak
    self.alpha = leak

  def __repr__(self): return f'LeakyReLU(leak={self.leak})'
  def __str__(self): return 'LeakyReLU'


def _split_axis(shape, axis):
  """(outer, L, inner) view of a C-contiguous array for a softmax along `axis`"""
  axis = axis % len(shape)
  outer = 1
  for s in shape[:axis]:
    outer *= s
  inner = 1
  for s in shape[axis + 1:]:
    inner *= s
  return outer, shape[axis], inner


class Softmax(Layer, Operation):
  """activations.py:105-180: max-shifted softmax along `axis`.  The reference builds a Jacobian per slice
  (apply_along_axis, 131-158); Jacobian(s) . ug collapses to s * (ug - sum(ug * s)), one kernel here."""

  def __init__(self, axis):
    self.axis = axis

  def forward(self, inputs):
    inputs = self.get_tensors(inputs)
    return self.get_result_tensor(self.calc_softmax(inputs.data, axis=self.axis), inputs)

  def backward(self, inputs):
    x, axis = inputs.data, self.axis

    def into(ug, dst, acc):
      outer, L, inner = _split_axis(x.shape, axis)
      if dst is None:
        dst, acc = D.DeviceArray.empty(x.shape), False
      B.check(B.lib.ngb_softmax_bwd(x.ptr, ug.ptr, outer, L, inner, dst.ptr, int(acc), D._st()))
      return dst
    inputs.set_grad_fn(GradFn(into))

  @staticmethod
  def calc_softmax(arr, axis=None):
    """activations.py:161-174."""
    arr = D.as_device(arr)
    if axis is None:
      outer, L, inner = 1, arr.size, 1
    else:
      outer, L, inner = _split_axis(arr.shape, axis)
    out = D.DeviceArray.empty(arr.shape)
    B.check(B.lib.ngb_softmax_fwd(arr.ptr, outer, L, inner, out.ptr, D._st()))
    return out

  def __repr__(self): return f'Softmax(axis={self.axis})'
  def __str__(self): return 'Softmax'
