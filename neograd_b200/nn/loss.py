"""Losses (reference: neograd/nn/loss.py).  MSE / BCE / CE are composites of the elementwise ops exactly as in the
reference (so BCE keeps its swapped arguments, loss.py:83 / SURVEY A.7); SoftmaxCE is one fused Operation."""
from ..autograd import sum as _sum, log
from ..autograd.ops.operation import Operation
from ..autograd.tensor import GradFn
from .. import _backend as B
from .. import device as D
from .activations import _split_axis


class Loss:
  def __call__(self, outputs, targets):
    return self.forward(outputs, targets)

  def get_num_examples(self, outputs_shape):
    """loss.py:22-37."""
    return 1 if len(outputs_shape) in (0, 1) else outputs_shape[0]


class MSE(Loss):
  """loss.py:41-62."""

  def forward(self, outputs, targets):
    num_examples = self.get_num_examples(outputs.shape)
    return (1 / (2 * num_examples)) * _sum((outputs - targets) ** 2)

  def __repr__(self): return 'MSE()'
  def __str__(self): return 'MeanSquaredError'


class _BCEOp(Operation):
  """The reference's BCE expression (loss.py:83-84: eleven elementwise ops and a sum) as ONE Operation: one reduction
  kernel forward, one elementwise kernel backward -- same arithmetic per element, argument roles as the reference has them."""

  def __init__(self, num_examples, epsilon):
    self.n, self.eps = num_examples, epsilon

  def forward(self, outputs, targets):
    outputs, targets = self.get_tensors(outputs, targets)
    loss = D.DeviceArray.empty(())
    B.check(B.lib.ngb_bce_fwd(outputs.data.ptr, targets.data.ptr, outputs.data.size, 1.0 / self.n, self.eps, loss.ptr, D._st()))
    return self.get_result_tensor(loss, outputs, targets)

  def get_broadcast_shape(self, *tensors):
    return None

  def backward(self, outputs, targets):
    o, t, inv_n, eps = outputs.data, targets.data, 1.0 / self.n, self.eps

    def make(side):
      def into(ug, dst, acc):
        if dst is None:
          dst, acc = D.DeviceArray.empty(o.shape), False
        ugp = None if D.const_value(ug) == 1.0 else ug.ptr          # loss.backward()'s host-known 1. stays an immediate
        B.check(B.lib.ngb_bce_bwd(o.ptr, t.ptr, ugp, o.size, inv_n, eps, dst.ptr if side == 0 else None, int(acc),
                                  dst.ptr if side == 1 else None, int(acc), D._st()))
        return dst
      return GradFn(into)
    outputs.set_grad_fn(make(0))
    targets.set_grad_fn(make(1))


class BCE(Loss):
  """loss.py:66-91."""

  def forward(self, outputs, targets, epsilon=1e-9):
    from ..autograd.tensor import Tensor
    num_examples = self.get_num_examples(outputs.shape)
    if (D.FUSED_LINEAR and isinstance(outputs, Tensor) and isinstance(targets, Tensor) and outputs.shape == targets.shape
        and outputs._data is not None and targets._data is not None and outputs.data.size >= 1):
      return _BCEOp(num_examples, epsilon).forward(outputs, targets)
    entropy = _sum((outputs * log(targets + epsilon)) + ((1 - outputs) * (log(1 - targets + epsilon))))
    return (-1 / num_examples) * entropy

  def __repr__(self): return 'BCE()'
  def __str__(self): return 'BinaryCrossEntropy'


class CE(Loss):
  """loss.py:95-120."""

  def forward(self, outputs, targets, epsilon=1e-9):
    num_examples = self.get_num_examples(outputs.shape)
    entropy = _sum(targets * log(outputs + epsilon))
    return (-1 / num_examples) * entropy

  def __repr__(self): return 'CE()'
  def __str__(self): return 'CrossEntropy'


class SoftmaxCE(Operation, Loss):
  """loss.py:124-172: loss = -(1/N) sum t*log(softmax(z)+eps); dz = (ug/N)(softmax(z) - t)."""

  def __init__(self, axis):
    self.axis = axis

  def forward(self, outputs, targets, epsilon=1e-9):
    outputs, targets = self.get_tensors(outputs, targets)
    if outputs.shape != targets.shape:
      # loss.py:155 / 170 multiply and subtract with NumPy broadcasting (the reference's own test passes one target row for
      # the whole batch): targets that broadcast TO the logits' shape are expanded on the device
      import numpy as np
      try:
        ok = np.broadcast_shapes(outputs.shape, targets.shape) == tuple(outputs.shape)
      except ValueError:
        ok = False
      if not ok:
        raise ValueError(f"operands could not be broadcast together with shapes {outputs.shape} {targets.shape}")
      from ..autograd.tensor import Tensor
      targets = Tensor(D.broadcast_to(targets.data, outputs.shape), requires_grad=targets.requires_grad)
    n = self.get_num_examples(outputs.shape)
    outer, L, inner = _split_axis(outputs.shape, self.axis)
    loss = D.DeviceArray.empty(())
    self._dz0 = None
    z = outputs.data
    if (D.LAZY_FUSION and D.FUSED_LINEAR and isinstance(z, D.LazyArray) and z.pending and z.tag == 'linear'
        and len(outputs.shape) == 2 and inner == 1 and targets._data is not None
        and B.lib.ngb_linear_softmax_ce_serves(outer, z.args.n_in, L) == 1):
      # the logits are a Linear layer's output nobody has read yet (a classifier head): logits, loss and the loss gradient
      # for an upstream gradient of 1 -- (softmax(z) - t) / n, loss.py:170 -- come out of ONE kernel
      pend = z.args
      zbuf = D.DeviceArray.empty(outputs.shape)
      dz0 = D.DeviceArray.empty(outputs.shape)
      B.check(B.lib.ngb_linear_softmax_ce(pend.x.ptr, pend.W.ptr, pend.b.ptr, targets.data.ptr, zbuf.ptr, loss.ptr, dz0.ptr,
                                          pend.M, pend.n_in, pend.n_out, 1.0 / n, epsilon, D._st()))
      z.t = zbuf.t
      self._dz0 = dz0
    else:
      B.check(B.lib.ngb_softmax_ce_fwd(z.ptr, targets.data.ptr, outer, L, inner, 1.0 / n, epsilon, loss.ptr, D._st()))
    return self.get_result_tensor(loss, outputs, targets)

  def backward(self, outputs, targets):
    z, t = outputs.data, targets.data
    n = self.get_num_examples(outputs.shape)
    outer, L, inner = _split_axis(outputs.shape, self.axis)

    dz0 = getattr(self, '_dz0', None)

    def into(ug, dst, acc):
      if dst is None:
        dst, acc = D.DeviceArray.empty(z.shape), False
      B.check(B.lib.ngb_softmax_ce_bwd(z.ptr, t.ptr, ug.ptr, outer, L, inner, 1.0 / n, dst.ptr, int(acc), D._st()))
      return dst

    def alias(ug):
      # upstream gradient is the host-known 1. of loss.backward(): the gradient the forward kernel already wrote IS the result
      return dz0 if (dz0 is not None and D.const_value(ug) == 1.0) else None
    outputs.set_grad_fn(GradFn(into, alias))
    assert targets.requires_grad is False, 'Targets Tensor should have requires_grad=False'

  def __repr__(self): return f'SoftmaxCE(axis={self.axis})'
  def __str__(self): return 'SoftmaxCrossEntropy'
