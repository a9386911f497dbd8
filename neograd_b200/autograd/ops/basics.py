"""Elementwise / Dot / Sum / shape ops (reference: neograd/autograd/ops/basics.py).

Forward bodies call one kernel of libngb200.so; every grad_fn is a `GradFn` whose `into` issues ONE kernel that
computes local-gradient x upstream-gradient, reduces over the broadcast axes of that operand (utils.py:34-78) and
writes / accumulates into the operand's gradient buffer."""
import ctypes

import numpy as np

from ... import _backend as B
from ... import device as D
from ..tensor import GradFn
from .operation import Operation


def _val(t):
  """operand as the kernels want it: the DeviceArray, or the Python scalar if it has not been uploaded"""
  return t._scalar if t._data is None else t._data


class _Binary(Operation):
  OP = None

  def forward(self, tens1, tens2):
    tens1, tens2 = self.get_tensors(tens1, tens2)
    return self.get_result_tensor(D.binary(self.OP, _val(tens1), _val(tens2)), tens1, tens2)

  def backward(self, tens1, tens2):
    op = self.OP

    def make(side):
      def into(ug, dst, accumulate):
        return D.binary_bwd(op, side, ug, _val(tens1), _val(tens2), dst, accumulate)
      alias = None
      if op == B.ADD or (op == B.SUB and side == 0):     # gradient == ug when the operand was not broadcast
        shape = (tens1, tens2)[side].shape
        alias = lambda ug: ug if ug.shape == shape else None
      return GradFn(into, alias)
    tens1.set_grad_fn(make(0))
    tens2.set_grad_fn(make(1))


class Add(_Binary):
  """basics.py:6-33: grads (ug, ug)."""
  OP = B.ADD


class Sub(_Binary):
  """basics.py:49-77: grads (ug, -ug)."""
  OP = B.SUB


class Mul(_Binary):
  """basics.py:93-120: grads (b*ug, a*ug)."""
  OP = B.MUL


class Div(_Binary):
  """basics.py:136-164: grads (ug/b, -a/b^2*ug)."""
  OP = B.DIV


class Pow(_Binary):
  """basics.py:299-327: grads (b*a^(b-1)*ug, a^b*ln(a)*ug)."""
  OP = B.POW


def add(tens1, tens2): return Add().forward(tens1, tens2)
def sub(tens1, tens2): return Sub().forward(tens1, tens2)
def mul(tens1, tens2): return Mul().forward(tens1, tens2)
def div(tens1, tens2): return Div().forward(tens1, tens2)
def pow(tens1, tens2): return Pow().forward(tens1, tens2)


class Dot(Operation):
  """basics.py:180-207: C = A.B; dA = ug.B^T; dB = A^T.ug -- three GEMMs (tcgen05 TF32 for tensor-core sized
  shapes, fp32 SIMT otherwise; ngb_gemm chooses), each able to accumulate into the gradient buffer in its epilogue."""

  def forward(self, tens1, tens2):
    tens1, tens2 = self.get_tensors(tens1, tens2)
    a, b = tens1.data, tens2.data
    if a.ndim != 2 or b.ndim != 2:
      raise ValueError(f"dot: only 2-D operands are supported on the device, got shapes {a.shape} and {b.shape}")
    return self.get_result_tensor(D.gemm(a, b), tens1, tens2)

  def backward(self, tens1, tens2):
    a, b = tens1.data, tens2.data
    tens1.set_grad_fn(GradFn(lambda ug, dst, acc: D.gemm(ug, b, tb=True, out=dst, accumulate=acc)))   # basics.py:206
    tens2.set_grad_fn(GradFn(lambda ug, dst, acc: D.gemm(a, ug, ta=True, out=dst, accumulate=acc)))   # basics.py:207


def dot(tens1, tens2): return Dot().forward(tens1, tens2)


class _Unary(Operation):
  OP = None
  alpha = 0.0

  def forward(self, tens):
    tens = self.get_tensors(tens)
    x, op, alpha = tens.data, self.OP, self.alpha
    if (op in _LINEAR_ACTS and D.LAZY_FUSION and D.FUSED_LINEAR and isinstance(x, D.LazyArray) and x.pending
        and x.tag == 'linear'):
      # the input is a Linear layer's output nobody has read yet: z = x.W + b and act(z) come out of ONE kernel (the GEMM's
      # epilogue writes both) when this activation is first used
      out = D.LazyArray(x.shape, x.args.act_thunk(x, op, alpha), 'linear_act', None)
      D.register_lazy(out, x.args.sources())
      if x.args.act_out is None:
        x.args.act_out = (op, alpha, out)           # (backward may take its mask from this output, see below)
    elif op == B.RELU and D.LAZY_FUSION:
      # produced on first use: a MaxPool that follows reads x itself (maxpool(relu(x)) in one kernel, conv.py here)
      out = D.LazyArray(x.shape, lambda: D.unary(op, x, alpha), 'relu', x)
      D.register_lazy(out, (x,))
    else:
      out = D.unary(op, x, alpha)
    return self.get_result_tensor(out, tens)

  def backward(self, tens):
    op, alpha, x = self.OP, self.alpha, tens.data
    if op == B.LEAKY_RELU and alpha > 0 and isinstance(x, D.LazyArray) and x.pending and x.tag == 'linear':
      # the fused Linear + LeakyReLU kernel wrote only the activation: sign(y) == sign(x) and y == 0 <=> x == 0, so the
      # reference's mask x >= 0 (activations.py:216) is y >= 0 -- the deferred pre-activation is never produced
      fused = x.args.act_out
      if fused is not None and fused[0] == op and fused[1] == alpha and not fused[2].pending:
        x = fused[2]

    def into(ug, dst, accumulate):
      if op in _LINEAR_ACTS and isinstance(ug, D.LazyArray) and ug.pending and ug.tag == 'linear_dgrad':
        # upstream gradient = data gradient of the Linear layer that consumed this activation, not computed yet:
        # (ug0 . W^T) * act'(x) in one kernel (the mask rides in the GEMM's epilogue)
        if dst is None:
          dst, accumulate = D.DeviceArray.empty(x.shape), False
        ug.args.run(dst, accumulate, x, op, alpha)
        return dst
      if op == B.RELU and isinstance(ug, D.LazyArray) and ug.pending and ug.tag == 'maxpool_bwd':
        # upstream gradient = input gradient of a MaxPool that has not been written yet: route + mask in one kernel
        pug, idx, outer, H, W, KH, KW, pad, stride, _ = ug.args
        if dst is None:
          dst, accumulate = D.DeviceArray.empty(x.shape), False
        rc = B.lib.ngb_maxpool_relu_bwd(pug.ptr, ctypes.c_void_p(idx.data_ptr()), x.ptr, dst.ptr, outer, H, W, KH, KW, pad, stride,
                                        int(accumulate), D._st())
        if rc == 0:
          return dst
        if rc != B.ERR_UNSUPPORTED:
          B.check(rc)
      if dst is None:
        return D.unary_bwd(op, x, ug, alpha)
      B.check(B.lib.ngb_ew_unary_bwd(op, float(alpha), x.ptr, ug.ptr, dst.ptr, x.size, int(accumulate), D._st()))
      return dst

    def lazy(ug):
      # ReLU behind a still-deferred 2x2 / stride 2 MaxPool backward: stay deferred.  A convolution that consumes this
      # gradient takes its compact pooled form (conv.py here: ngb_conv_wgrad_pooled, bias gradient of the pooled map);
      # anything else materialises it with the fused route-and-mask kernel above.
      if not (op == B.RELU and D.LAZY_FUSION and D.POOLED_CONV_GRADS and isinstance(ug, D.LazyArray) and ug.pending
              and ug.tag == 'maxpool_bwd'):
        return None
      pug, idx, outer, H, W, KH, KW, pad, stride, ypool = ug.args
      if (KH, KW, stride, pad) != (2, 2, 2, 0) or H % 2 or W % 2 or ypool is None:
        return None
      return D.LazyArray(x.shape, lambda: into(ug, None, False), 'pool_relu_bwd', D.PoolReluGrad(pug, idx, ypool, x, outer, H, W))
    tens.set_grad_fn(GradFn(into, None, lazy))


_LINEAR_ACTS = (B.RELU, B.LEAKY_RELU, B.SIGMOID, B.TANH)   # activations the Linear kernels' epilogues implement


class Exp(_Unary):
  """basics.py:223-246."""
  OP = B.EXP


class Log(_Unary):
  """basics.py:261-284."""
  OP = B.LOG


def exp(tens): return Exp().forward(tens)
def log(tens): return Log().forward(tens)


def _copy_into(src, dst, accumulate, fresh=False):
  """dst (+)= src for equally sized arrays (the gradient of a pure relabelling op).  `fresh`: src is a new buffer
  nobody else holds, so it can be handed out without a copy."""
  if dst is None:
    return src if fresh else src.copy()
  if accumulate:
    dst.add_(src)
  else:
    dst.assign(src)
  return dst


class Sum(Operation):
  """basics.py:343-386: np.sum(x, axis); backward broadcasts ug back over the summed axes."""

  def __init__(self, axis=None):
    self.axis = axis

  def forward(self, tens):
    tens = self.get_tensors(tens)
    return self.get_result_tensor(D.reduce_sum(tens.data, self.axis), tens)

  def backward(self, tens):
    shape, axis = tens.shape, self.axis

    def into(ug, dst, accumulate):
      # ones(shape) * expand_dims(ug, axis)  (basics.py:381-386): one broadcast-copy kernel
      if axis is None:
        ugk = ug.reshape(())
      else:
        axes = (axis,) if isinstance(axis, (int, np.integer)) else tuple(axis)
        axes = tuple(a % len(shape) for a in axes)
        ugk = ug.reshape(tuple(1 if i in axes else s for i, s in enumerate(shape)))
      return _copy_into(D.broadcast_to(ugk, shape), dst, accumulate, fresh=True)
    tens.set_grad_fn(GradFn(into))


def sum(tens, axis=None): return Sum(axis).forward(tens)


class Transpose(Operation):
  """basics.py:403-437: np.transpose (all axes reversed); gradient is the transpose of ug."""

  def forward(self, tens):
    tens = self.get_tensors(tens)
    return self.get_result_tensor(tens.data.transpose(), tens)

  def backward(self, tens):
    tens.set_grad_fn(GradFn(lambda ug, dst, acc: _copy_into(ug.transpose(), dst, acc, fresh=ug.ndim >= 2)))


def transpose(tens): return Transpose().forward(tens)


class Flatten(Operation):
  """basics.py:441-476: (n,1) column view."""

  def forward(self, tens):
    tens = self.get_tensors(tens)
    return self.get_result_tensor(tens.data.reshape((tens.data.size, 1)), tens)

  def backward(self, tens):
    shape = tens.shape
    tens.set_grad_fn(GradFn(lambda ug, dst, acc: _copy_into(ug.reshape(shape), dst, acc), lambda ug: ug.reshape(shape)))


def flatten(tens): return Flatten().forward(tens)


class Reshape(Operation):
  """basics.py:480-514: metadata-only view of the same HBM."""

  def forward(self, tens, new_shape):
    tens = self.get_tensors(tens)
    return self.get_result_tensor(tens.data.reshape(new_shape), tens)

  def backward(self, tens):
    shape = tens.shape
    tens.set_grad_fn(GradFn(lambda ug, dst, acc: _copy_into(ug.reshape(shape), dst, acc), lambda ug: ug.reshape(shape)))


def reshape(tens, new_shape): return Reshape().forward(tens, new_shape)
