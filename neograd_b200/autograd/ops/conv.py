"""Conv2D / Conv3D / MaxPool2D / MaxPool3D (reference: neograd/autograd/ops/conv.py).

One kernel launch per forward and per operand gradient (the reference runs a Python loop over output positions).
Results and upstream gradients use the reference's output order -- position (p,q) of the textbook (Ho,Wo) map is
stored at flat index q*Ho+p (conv.py:39-47 vs 149-151) -- which the kernels produce / consume directly.
neograd's single-channel Conv2D (conv.py:121-215) runs through the same entry points with C = O = 1."""
import ctypes

import numpy as np

from ... import _backend as B
from ... import device as D
from ..tensor import GradFn
from .operation import Operation


class Conv:
  """Geometry shared by the four ops (conv.py:6-117)."""

  def check_geometry(self):
    assert self.stride >= 1, 'Stride must be greater than or equal to 1'          # conv.py:33
    assert self.padding >= 0, 'Padding must be greater than or equal to zero'     # conv.py:34

  def get_result_shape(self, inputs_shape, kernel_shape):
    """conv.py:49-68: floor((dim + 2p - k)/s + 1) on the last two dims."""
    self.check_geometry()
    ho = int((inputs_shape[-2] + 2 * self.padding - kernel_shape[-2]) / self.stride + 1)
    wo = int((inputs_shape[-1] + 2 * self.padding - kernel_shape[-1]) / self.stride + 1)
    return ho, wo


def _conv_forward(op, x4, w4, b1, out_shape):
  N, C, H, W = x4.shape
  O, _, KH, KW = w4.shape
  pad, stride = op.padding, op.stride

  def run():
    y = D.DeviceArray.empty(out_shape)
    B.check(B.lib.ngb_conv_fprop(x4.ptr, w4.ptr, b1.ptr, y.ptr, N, C, H, W, O, KH, KW, pad, stride, D._st()))
    return y

  if (D.LAZY_FUSION and D.FUSED_CONV_POOL and len(out_shape) == 4 and
      B.lib.ngb_conv_relu_pool_fwd_serves(N, C, H, W, O, KH, KW, pad, stride) == 1):
    # a ReLU -> MaxPool(2x2, s2) behind this convolution runs with it as ONE kernel (_MaxPool.forward) and this output is
    # never written; anything else that reads it produces it here
    out = D.LazyArray(out_shape, run, 'conv', (x4, w4, b1, (N, C, H, W, O, KH, KW, pad, stride)))
    D.register_lazy(out, (x4, w4, b1))
    return out
  return run()


def _conv_grad_fns(op, inputs, kernel, bias, x4shape, w4shape):
  """dgrad / wgrad / bgrad closures (conv.py:299-327, 179-199); each can accumulate into the gradient buffer."""
  N, C, H, W = x4shape
  O, _, KH, KW = w4shape
  pad, stride = op.padding, op.stride
  x, w = inputs.data, kernel.data
  in_shape, k_shape, b_shape = inputs.shape, kernel.shape, bias.shape   # (closures must not hold the tensors: cycles)

  def pooled(ug):
    """The upstream gradient as (masked pooled gradient, argmax indices) if it is a still-deferred ReLU+MaxPool backward
    (device.PoolReluGrad): three quarters of the full-resolution gradient are zeros, the kernels below never build it."""
    if isinstance(ug, D.LazyArray) and ug.pending and ug.tag == 'pool_relu_bwd':
      return ug.args
    return None

  def dgrad(ug, dst, acc):
    if dst is None:
      dst, acc = D.DeviceArray.empty(in_shape), False
    pg = pooled(ug) if D.POOLED_CONV_DGRAD else None
    if pg is not None:
      rc = B.lib.ngb_conv_dgrad_pooled(pg.pug.ptr, None, ctypes.c_void_p(pg.idx.data_ptr()), None, w.ptr, dst.ptr,
                                       N, C, H, W, O, KH, KW, pad, stride, int(acc), D._st())
      if rc == 0:
        return dst
      if rc != B.ERR_UNSUPPORTED:
        B.check(rc)
    B.check(B.lib.ngb_conv_dgrad(ug.ptr, w.ptr, dst.ptr, N, C, H, W, O, KH, KW, pad, stride, int(acc), D._st()))
    return dst

  def wgrad(ug, dst, acc):
    if dst is None:
      dst, acc = D.DeviceArray.empty(k_shape), False
    pg = pooled(ug)
    if pg is not None:
      rc = B.lib.ngb_conv_wgrad_pooled(pg.pug.ptr, None, ctypes.c_void_p(pg.idx.data_ptr()), None, x.ptr, dst.ptr,
                                       N, C, H, W, O, KH, KW, pad, stride, int(acc), D._st())
      if rc == 0:
        return dst
      if rc != B.ERR_UNSUPPORTED:
        B.check(rc)
    B.check(B.lib.ngb_conv_wgrad(ug.ptr, x.ptr, dst.ptr, N, C, H, W, O, KH, KW, pad, stride, int(acc), D._st()))
    return dst

  def bgrad(ug, dst, acc):
    if dst is None:
      dst, acc = D.DeviceArray.empty(b_shape), False
    pg = pooled(ug)
    if pg is not None:      # the sum over a plane of the scattered gradient is the sum over its masked pooled form
      B.check(B.lib.ngb_conv_bgrad_pooled_idx(pg.pug.ptr, ctypes.c_void_p(pg.idx.data_ptr()), dst.ptr, N, O,
                                              (pg.H // 2) * (pg.W // 2), int(acc), D._st()))
      return dst
    B.check(B.lib.ngb_conv_bgrad(ug.ptr, dst.ptr, N, O, ug.size // max(N * O, 1), int(acc), D._st()))
    return dst

  def wbgrad(ug, dst_w, acc_w, dst_b, acc_b):
    """Weight and bias gradient from ONE pass over a pooled upstream gradient (ngb_conv_wbgrad_pooled); otherwise the two
    kernels above, one after the other."""
    pg = pooled(ug)
    if pg is not None:
      rc = B.lib.ngb_conv_wbgrad_pooled(pg.pug.ptr, None, ctypes.c_void_p(pg.idx.data_ptr()), None, x.ptr, dst_w.ptr,
                                        dst_b.ptr, N, C, H, W, O, KH, KW, pad, stride, int(acc_w), int(acc_b), D._st())
      if rc == 0:
        return
      if rc != B.ERR_UNSUPPORTED:
        B.check(rc)
    wgrad(ug, dst_w, acc_w)
    bgrad(ug, dst_b, acc_b)

  inputs.set_grad_fn(GradFn(dgrad))
  fused = []

  def joint_ok(ug):          # only when ONE kernel yields both: otherwise the two run beside each other on two lanes
    if pooled(ug) is None:
      return False
    if not fused:
      fused.append(B.lib.ngb_conv_wbgrad_pooled_is_fused(N, C, H, W, O, KH, KW, pad, stride) == 1)
    return fused[0]

  kernel.set_grad_fn(GradFn(wgrad, lazy_ok=True, joint=(bias, wbgrad, joint_ok)))
  bias.set_grad_fn(GradFn(bgrad, lazy_ok=True))


class Conv2D(Operation, Conv):
  """conv.py:121-215: (N,H,W) (*) (KH,KW) + scalar bias -> (N,Ho,Wo)."""

  def __init__(self, padding=0, stride=1):
    self.padding = padding
    self.stride = stride

  def forward(self, inputs, kernel, bias):
    inputs, kernel, bias = self.get_tensors(inputs, kernel, bias)
    self.validate_inputs(inputs)
    ho, wo = self.get_result_shape(inputs.shape, kernel.shape)
    N, H, W = inputs.shape
    y = _conv_forward(self, inputs.data.reshape((N, 1, H, W)), kernel.data.reshape((1, 1) + kernel.shape),
                      bias.data.reshape((1,)), (N, ho, wo))
    return self.get_result_tensor(y, inputs, kernel, bias)

  def backward(self, inputs, kernel, bias):
    N, H, W = inputs.shape
    _conv_grad_fns(self, inputs, kernel, bias, (N, 1, H, W), (1, 1) + kernel.shape)

  def validate_inputs(self, inputs):
    if len(inputs.shape) != 3:
      raise ValueError("Only 3D inputs, with 0th dim as number of examples are supported!")


def conv2d(inputs, kernel, bias, padding, stride):
  return Conv2D(padding, stride).forward(inputs, kernel, bias)


class Conv3D(Operation, Conv):
  """conv.py:234-343: (N,C,H,W) (*) (O,C,KH,KW) + bias (O,) -> (N,O,Ho,Wo)."""

  def __init__(self, padding=0, stride=1):
    self.padding = padding
    self.stride = stride

  def forward(self, inputs, kernel, bias):
    inputs, kernel, bias = self.get_tensors(inputs, kernel, bias)
    self.validate_inputs(inputs)
    if len(kernel.shape) != 4 or kernel.shape[1] != inputs.shape[1]:
      raise ValueError(f"operands could not be broadcast together with shapes {inputs.shape} {kernel.shape}")
    ho, wo = self.get_result_shape(inputs.shape, kernel.shape)
    y = _conv_forward(self, inputs.data, kernel.data, bias.data.reshape((kernel.shape[0],)),
                      (inputs.shape[0], kernel.shape[0], ho, wo))
    return self.get_result_tensor(y, inputs, kernel, bias)

  def backward(self, inputs, kernel, bias):
    _conv_grad_fns(self, inputs, kernel, bias, inputs.shape, kernel.shape)

  def validate_inputs(self, inputs):
    if len(inputs.shape) != 4:
      raise ValueError("Only 4D inputs, with 0th dim as number of examples, 1st dim as number of channels are supported!")


def conv3d(inputs, kernel, bias, padding, stride):
  return Conv3D(padding, stride).forward(inputs, kernel, bias)


class _MaxPool(Operation, Conv):
  """conv.py:362-542.  Forward also stores the uint8 argmax of every window (first maximum, row-major in the
  zero-padded window: np.argmax at conv.py:424/523); backward routes ug through those indices instead of re-running
  argmax over the input."""
  NDIM = None
  ERR = None

  def __init__(self, kernel_shape, padding=0, stride=1):
    self.kernel_shape = kernel_shape
    self.padding = padding
    self.stride = stride
    self._idx = None
    self._y = None

  def forward(self, inputs):
    inputs = self.get_tensors(inputs)
    self.validate_inputs(inputs)
    if len(self.kernel_shape) != 2:
      raise ValueError("Only 2D kernels are allowed!")                    # conv.py:377-380, 476-479
    ho, wo = self.get_result_shape(inputs.shape, self.kernel_shape)
    lead = inputs.shape[:-2]
    H, W = inputs.shape[-2:]
    outer = int(np.prod(lead))
    y = D.DeviceArray.empty(lead + (ho, wo))
    idx = D.torch().empty(lead + (ho, wo), device='cuda', dtype=D.torch().uint8)
    xin, rc = inputs.data, -1
    if isinstance(xin, D.LazyArray) and xin.pending and xin.tag == 'relu':
      src = xin.args
      if (isinstance(src, D.LazyArray) and src.pending and src.tag == 'conv' and tuple(self.kernel_shape) == (2, 2)
          and self.stride == 2 and self.padding == 0):
        # ... and the ReLU's input is a convolution output nobody has read either: conv + bias + ReLU + pool in one kernel
        x4, w4, b1, (cN, cC, cH, cW, cO, cKH, cKW, cpad, cstride) = src.args
        rc = B.lib.ngb_conv_relu_pool_fwd(x4.ptr, w4.ptr, b1.ptr, y.ptr, ctypes.c_void_p(idx.data_ptr()), cN, cC, cH, cW, cO,
                                          cKH, cKW, cpad, cstride, D._st())
        if rc not in (0, B.ERR_UNSUPPORTED):
          B.check(rc)
    if rc != 0 and isinstance(xin, D.LazyArray) and xin.pending and xin.tag == 'relu':
      # the input is a ReLU output nobody has read yet: pool max(0, x) of the ReLU's own input, relu(x) is never written
      rc = B.lib.ngb_maxpool_relu_fwd(xin.args.ptr, y.ptr, ctypes.c_void_p(idx.data_ptr()), outer, H, W,
                                      self.kernel_shape[0], self.kernel_shape[1], self.padding, self.stride, D._st())
      if rc not in (0, B.ERR_UNSUPPORTED):
        B.check(rc)
    if rc != 0:
      B.check(B.lib.ngb_maxpool_fwd(xin.ptr, y.ptr, ctypes.c_void_p(idx.data_ptr()), outer, H, W,
                                    self.kernel_shape[0], self.kernel_shape[1], self.padding, self.stride, D._st()))
    # only the fused kernel records the ReLU gradient in bit 2 of the argmax bytes, which the pooled conv-gradient kernels read
    self._idx, self._y = idx, (y if rc == 0 else None)
    return self.get_result_tensor(y, inputs)

  def backward(self, inputs):
    idx, ypool, (KH, KW), pad, stride = self._idx, self._y, self.kernel_shape, self.padding, self.stride
    in_shape = inputs.shape          # (the closures below must not hold `inputs`: tensor <-> grad_fn cycles outlive the pass)
    lead = inputs.shape[:-2]
    H, W = inputs.shape[-2:]
    outer = int(np.prod(lead))

    def into(ug, dst, acc):
      if dst is None:
        dst, acc = D.DeviceArray.empty(in_shape), False
      B.check(B.lib.ngb_maxpool_bwd(ug.ptr, ctypes.c_void_p(idx.data_ptr()), dst.ptr, outer, H, W, KH, KW, pad, stride,
                                    int(acc), D._st()))
      return dst

    def lazy(ug):
      # the input gradient as a deferred array: a ReLU backward that consumes it fuses the routing with its mask
      if not D.LAZY_FUSION:
        return None
      return D.LazyArray(in_shape, lambda: into(ug, None, False), 'maxpool_bwd', (ug, idx, outer, H, W, KH, KW, pad, stride, ypool))
    inputs.set_grad_fn(GradFn(into, None, lazy))

  def validate_inputs(self, inputs):
    if len(inputs.shape) != self.NDIM:
      raise ValueError(self.ERR)


class MaxPool2D(_MaxPool):
  NDIM = 3
  ERR = "Only 3D inputs, with 0th dim as number of examples are supported!"


class MaxPool3D(_MaxPool):
  NDIM = 4
  ERR = "Only 4D inputs, with 0th dim as number of examples, 1st dim as number of channels are supported!"


def maxpool2d(inputs, kernel_shape, padding, stride):
  return MaxPool2D(kernel_shape, padding, stride).forward(inputs)


def maxpool3d(inputs, kernel_shape, padding, stride):
  return MaxPool3D(kernel_shape, padding, stride).forward(inputs)
