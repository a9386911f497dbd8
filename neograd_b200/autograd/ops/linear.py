"""Linear as ONE Operation (new; the reference writes `dot(inputs, W) + b`, neograd/nn/layers/misc.py:63, i.e. a Dot node
and an Add node -- basics.py:180-207 and 6-33 -- whose arithmetic this reproduces).

Why one node: the kernels of a Linear layer fuse across the Dot / Add / activation boundaries (csrc/linear.cu):
  forward   z = x.W + b  and  act(z)  from one GEMM (bias and activation in the epilogue);
  backward  the bias gradient is the column sum of the SAME upstream gradient the weight gradient multiplies, so one
            kernel yields both (GradFn.joint); the data gradient is deferred so that the backward of the activation in
            front of this layer can apply its mask in that GEMM's epilogue.
Results equal the two-node form: z, dW = x^T.ug, db = unbroadcast(ug) = ug.sum(axis=0, keepdims=True), dx = ug.W^T.
Anything the fused kernels do not cover (operands that are not 2-D, a bias that is not (1, out)) takes the reference's
`dot(...) + b` path in nn/layers/misc.py."""
from ... import _backend as B
from ... import device as D
from ..tensor import GradFn, Tensor
from ..utils import get_graph
from .operation import Operation


class LinearPending:
  """What a still-deferred Linear output stands for (args of a LazyArray tagged 'linear')."""
  __slots__ = ('x', 'W', 'b', 'M', 'n_in', 'n_out', 'track', 'act_out')

  def __init__(self, x, W, b, track):
    self.x, self.W, self.b, self.track = x, W, b, track
    self.act_out = None                           # (op, alpha, deferred act(z)) of the activation fused onto this layer
    self.M, self.n_in = x.shape
    self.n_out = W.shape[1]

  def sources(self):
    return (self.x, self.W, self.b)

  def launch(self, z, a, act, alpha):
    B.check(B.lib.ngb_linear_fwd(self.x.ptr, self.W.ptr, self.b.ptr, z.ptr if z is not None else None,
                                 a.ptr if a is not None else None, act, float(alpha), self.M, self.n_in, self.n_out, D._st()))

  def z_thunk(self):
    def thunk():
      z = D.DeviceArray.empty((self.M, self.n_out))
      self.launch(z, None, -1, 0.0)
      return z
    return thunk

  def act_thunk(self, zlazy, op, alpha):
    """Producer of act(z) for an activation layer that found z still deferred: one kernel writes z (kept for the
    activation's backward when the graph is tracking) and act(z)."""
    def thunk():
      if not zlazy.pending:                       # somebody read z in the meantime: plain elementwise kernel
        return D.unary(op, zlazy, alpha)
      shape = (self.M, self.n_out)
      # LeakyReLU with a positive slope preserves the sign (and the zero) of its input, so its backward mask z >= 0 can be
      # read off the OUTPUT: the pre-activation is then not written at all (it stays deferred for whoever asks for it)
      need_z = self.track and not (op == B.LEAKY_RELU and alpha > 0)
      z = D.DeviceArray.empty(shape) if need_z else None
      a = D.DeviceArray.empty(shape)
      self.launch(z, a, op, alpha)
      if z is not None:
        zlazy.t = z.t                             # z now exists: the deferred array is fulfilled by the same launch
      return a
    return thunk


class LinearDgradPending:
  """A still-deferred data gradient ug.W^T (args of a LazyArray tagged 'linear_dgrad')."""
  __slots__ = ('ug', 'W', 'M', 'n_in', 'n_out')

  def __init__(self, ug, W, M, n_in, n_out):
    self.ug, self.W, self.M, self.n_in, self.n_out = ug, W, M, n_in, n_out

  def run(self, dst, accumulate, zprev=None, act=-1, alpha=0.0):
    B.check(B.lib.ngb_linear_dgrad(self.ug.ptr, self.W.ptr, zprev.ptr if zprev is not None else None, dst.ptr, act,
                                   float(alpha), self.M, self.n_in, self.n_out, int(accumulate), D._st()))
    return dst


class Linear(Operation):
  def forward(self, inputs, weights, bias):
    inputs, weights, bias = self.get_tensors(inputs, weights, bias)
    x, W, b = inputs.data, weights.data, bias.data
    pend = LinearPending(x, W, b, get_graph().track)
    if D.LAZY_FUSION:
      out = D.LazyArray((pend.M, pend.n_out), pend.z_thunk(), 'linear', pend)
      D.register_lazy(out, pend.sources())
    else:
      out = pend.z_thunk()()
    return self.get_result_tensor(out, inputs, weights, bias)

  def get_broadcast_shape(self, *tensors):
    return None          # (the bias broadcast is handled inside the kernels: db is already bias-shaped)

  def backward(self, inputs, weights, bias):
    x, W = inputs.data, weights.data
    M, n_in = x.shape
    n_out = W.shape[1]
    x_shape, w_shape, b_shape = inputs.shape, weights.shape, bias.shape

    def dgrad(ug, dst, acc):                                           # basics.py:206
      if dst is None:
        dst, acc = D.DeviceArray.empty(x_shape), False
      return LinearDgradPending(ug, W, M, n_in, n_out).run(dst, acc)

    def dgrad_lazy(ug):
      if not (D.LAZY_FUSION and D.FUSED_LINEAR):
        return None
      pend = LinearDgradPending(ug, W, M, n_in, n_out)
      return D.LazyArray(x_shape, lambda: pend.run(D.DeviceArray.empty(x_shape), False), 'linear_dgrad', pend)

    def wgrad(ug, dst, acc):                                           # basics.py:207
      if dst is None:
        dst, acc = D.DeviceArray.empty(w_shape), False
      B.check(B.lib.ngb_linear_wgrad(x.ptr, ug.ptr, dst.ptr, None, M, n_in, n_out, int(acc), 0, D._st()))
      return dst

    def bgrad(ug, dst, acc):                                           # unbroadcast of the bias add: utils.py:75
      if dst is None:
        dst, acc = D.DeviceArray.empty(b_shape), False
      return D.reduce_sum(ug, axis=0, keepdims=True, out=dst, accumulate=acc)

    def wbgrad(ug, dst_w, acc_w, dst_b, acc_b):                        # both from one pass over ug
      B.check(B.lib.ngb_linear_wgrad(x.ptr, ug.ptr, dst_w.ptr, dst_b.ptr, M, n_in, n_out, int(acc_w), int(acc_b), D._st()))

    inputs.set_grad_fn(GradFn(dgrad, None, dgrad_lazy))
    weights.set_grad_fn(GradFn(wgrad, joint=(bias, wbgrad, lambda ug: True)))
    bias.set_grad_fn(GradFn(bgrad, joint=(weights, wbgrad, lambda ug: True, False)))


def fusable(inputs, weights, bias):
  """Whether `dot(inputs, weights) + bias` is the plain Linear case the fused kernels serve."""
  if not D.FUSED_LINEAR or not isinstance(inputs, Tensor) or inputs._data is None:
    return False
  xs, ws, bs = inputs.shape, weights.shape, bias.shape
  return (len(xs) == 2 and len(ws) == 2 and xs[1] == ws[0] and bs == (1, ws[1]) and xs[0] >= 1
          and inputs.requires_broadcasting and bias.requires_broadcasting and weights.requires_broadcasting)


def linear(inputs, weights, bias):
  return Linear().forward(inputs, weights, bias)
