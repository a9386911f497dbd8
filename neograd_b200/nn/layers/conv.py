"""Conv2D / Conv3D / MaxPool2D / MaxPool3D layers (reference: neograd/nn/layers/conv.py): thin callers of the ops."""
import numpy as np

from ..layers import Layer, Param
from ...autograd.ops import conv2d, conv3d, maxpool2d, maxpool3d


def _check_kernel_shape(kernel_shape):
  if len(kernel_shape) != 2:
    raise ValueError("Kernel shape can only have 2 dims")


class Conv2D(Layer):
  """layers/conv.py:6-45: single-channel, single-filter; weights ~ randn(KH, KW), scalar bias 0."""

  def __init__(self, kernel_shape, padding=0, stride=1):
    self.padding = padding
    self.stride = stride
    _check_kernel_shape(kernel_shape)
    self.weights = Param(np.random.randn(*kernel_shape), requires_grad=True, requires_broadcasting=False)
    self.bias = Param(0, requires_grad=True, requires_broadcasting=False)

  def forward(self, inputs):
    return conv2d(inputs, self.weights, self.bias, self.padding, self.stride)

  def __repr__(self):
    return f'Conv2D(kernel_shape={self.weights.shape}, padding={self.padding}, stride={self.stride})'

  __str__ = __repr__


class Conv3D(Layer):
  """layers/conv.py:48-91: weights ~ randn(O, C, KH, KW), bias zeros(O)."""

  def __init__(self, in_channels, out_channels, kernel_shape, padding=0, stride=1):
    self.padding = padding
    self.stride = stride
    _check_kernel_shape(kernel_shape)
    self.weights = Param(np.random.randn(out_channels, in_channels, *kernel_shape), requires_grad=True,
                         requires_broadcasting=False)
    self.bias = Param(np.zeros(out_channels), requires_grad=True, requires_broadcasting=False)

  def forward(self, inputs):
    return conv3d(inputs, self.weights, self.bias, self.padding, self.stride)

  def __repr__(self):
    ks = self.weights.shape
    return f'Conv3D(out_channels={ks[0]}, in_channels={ks[1]}, kernel_shape={ks[2:]}, padding={self.padding}, stride={self.stride})'

  __str__ = __repr__


class MaxPool2D(Layer):
  """layers/conv.py:94-127."""

  def __init__(self, kernel_shape, padding=0, stride=1):
    self.padding = padding
    self.stride = stride
    _check_kernel_shape(kernel_shape)
    self.kernel_shape = kernel_shape

  def forward(self, inputs):
    return maxpool2d(inputs, self.kernel_shape, self.padding, self.stride)

  def __repr__(self):
    return f'MaxPool2D(kernel_shape={self.kernel_shape}, padding={self.padding}, stride={self.stride})'

  __str__ = __repr__


class MaxPool3D(Layer):
  """layers/conv.py:130-163."""

  def __init__(self, kernel_shape, padding=0, stride=1):
    self.padding = padding
    self.stride = stride
    _check_kernel_shape(kernel_shape)
    self.kernel_shape = kernel_shape

  def forward(self, inputs):
    return maxpool3d(inputs, self.kernel_shape, self.padding, self.stride)

  def __repr__(self):
    return f'MaxPool3D(kernel_shape={self.kernel_shape}, padding={self.padding}, stride={self.stride})'

  __str__ = __repr__
