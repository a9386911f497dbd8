"""Sequential / Linear / Dropout (reference: neograd/nn/layers/misc.py)."""
import numpy as np

from ..layers import Container, Layer, Param
from ...autograd import dot
from ...autograd.ops.operation import Operation
from ...autograd.tensor import GradFn
from ...autograd.ops.basics import _copy_into
from ...autograd.ops import linear as _linear


class Sequential(Container):
  """misc.py:7-34."""

  def __init__(self, *args):
    self.layers = args

  def forward(self, inputs):
    for layer in self.layers:
      output = layer(inputs)
      inputs = output
    return output

  def __repr__(self):
    return f'Sequential({super().__repr__()})'

  def __str__(self):
    return f'Sequential({super().__str__()})'


class Linear(Layer):
  """misc.py:37-69: dot(inputs, W) + b with W ~ randn(in, out), b = zeros((1, out)) drawn from the global NumPy RNG at
  construction (so a seeded script builds the same model as with the reference)."""

  def __init__(self, num_in, num_out):
    self.num_in = num_in
    self.num_out = num_out
    self.weights = Param(np.random.randn(num_in, num_out), requires_grad=True)
    self.bias = Param(np.zeros((1, num_out)), requires_grad=True)

  def forward(self, inputs):
    if _linear.fusable(inputs, self.weights, self.bias):
      return _linear.linear(inputs, self.weights, self.bias)      # same arithmetic, one Operation (autograd/ops/linear.py)
    return dot(inputs, self.weights) + self.bias                  # misc.py:63

  def __repr__(self):
    return f'Linear(num_in={self.num_in}, num_out={self.num_out})'

  def __str__(self):
    return f'Linear in:{self.num_in} out:{self.num_out}'


class Dropout(Layer, Operation):
  """misc.py:72-128.  The reference's Dropout is an identity in both directions: forward returns
  `get_result_tensor(inputs.data, ...)` ignoring the masked `result` (misc.py:103-108) and backward's second
  `set_grad_fn` overrides the first (misc.py:120-122).  Results must match the reference, so this is reproduced --
  including the draw from NumPy's global RNG in training mode, which keeps seeded scripts on the same random stream."""

  def __init__(self, prob):
    Layer.__init__(self)
    assert prob > 0 and prob <= 1, 'Probability should be between 0 and 1'
    self.prob = prob

  def forward(self, inputs):
    inputs = self.get_tensors(inputs)
    if not self.eval:
      np.random.random(inputs.shape)
    return self.get_result_tensor(inputs.data, inputs)

  def backward(self, inputs, *_):
    inputs.set_grad_fn(GradFn(lambda ug, dst, acc: _copy_into(ug, dst, acc)))

  def __repr__(self):
    return f'Dropout(prob={self.prob})'

  def __str__(self):
    return f'Dropout(prob={self.prob})'
