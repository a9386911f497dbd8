"""Engine utilities (reference: neograd/autograd/utils.py): data coercion, unbroadcast, graph context managers and
finite-difference gradient checking."""
import numpy as np

from ..device import DeviceArray, reduce_sum
from .graph import Graph


def process_data(data):
  """utils.py:6-32.  Accepts exactly int / float / list / np.ndarray (and, new, a DeviceArray, which is already
  resident); anything else -- including NumPy scalar types -- raises TypeError as in the reference.
  Host data is rounded float64 -> float32 once and uploaded."""
  if isinstance(data, DeviceArray):
    return data
  supported = (int, float, list, np.ndarray)
  if type(data) in supported:
    try:
      return DeviceArray.from_host(data)
    except (ValueError, TypeError):
      raise TypeError(f'Elements of data should be of type float or be typecastable to float')
  raise TypeError(f'Expected data of types {supported} instead got {type(data)}')


def unbroadcast_axes(orig_shape, broadcast_shape):
  """utils.py:49-71: right-aligned; a missing leading dim or a dim that differs from the broadcast dim is summed."""
  nb, no = len(broadcast_shape), len(orig_shape)
  return tuple(d for d in range(nb) if d - (nb - no) < 0 or orig_shape[d - (nb - no)] != broadcast_shape[d])


def unbroadcast_data(data, orig_data_shape, broadcasted_shape):
  """utils.py:34-78: sum `data` over the axes along which the operand was broadcast (generic path, used for
  gradients produced by user-defined Operations; the built-in ops fuse this reduction into their backward kernel)."""
  if broadcasted_shape is None:
    return data
  axes = unbroadcast_axes(tuple(orig_data_shape), tuple(broadcasted_shape))
  if not axes:
    return data
  if not isinstance(data, DeviceArray):
    data = DeviceArray.from_host(data)
  return reduce_sum(data, axes)


_GLOBAL_GRAPH = None     # neograd_b200._NG_GRAPH, looked up once (this function runs several times per op)


def get_graph():
  """utils.py:80-93: the graph of the innermost `new_graph()` block, else the global one."""
  g = Graph.graph
  if g is not None:
    return g
  global _GLOBAL_GRAPH
  if _GLOBAL_GRAPH is None:
    from .. import _NG_GRAPH
    _GLOBAL_GRAPH = _NG_GRAPH
  return _GLOBAL_GRAPH


class new_graph:
  """utils.py:96-110."""

  def __enter__(self):
    self._prev = Graph.graph
    Graph.graph = Graph()

  def __exit__(self, exc_type, exc_value, exc_traceback):
    Graph.graph = self._prev if getattr(self, '_prev', None) is not None else None


class no_track:
  """utils.py:113-133: operations inside the block create no nodes."""

  def __init__(self):
    self.graph = get_graph()

  def __enter__(self):
    self.graph.track = False

  def __exit__(self, exc_type, exc_value, exc_traceback):
    self.graph.track = True


# The device computes in fp32, so the reference's default probe width of 1e-7 (utils.py:195) is below the arithmetic's
# resolution.  Measured on a B200 (README MLP, MSE): eps 1e-1 -> dist 8e-2 (ReLU kinks crossed), 1e-2 -> 2e-2,
# 1e-3 -> 3e-3, 1e-4 -> 3e-2 (fp32 rounding of the loss).  Defaults: central differences with eps = 1e-3 and a pass
# threshold of 1e-2 on the reference's distance ||a-n|| / (||a|| + ||n||) (utils.py:153); the check runs with the
# exact-fp32 kernels (set_precision('fp32')) so TF32 rounding does not enter the finite differences.
DEFAULT_EPSILON = 1e-3
DEFAULT_TOLERANCE = 1e-2
MIN_EPSILON = 1e-4     # below this a central difference of an fp32 loss is rounding noise


def _device_epsilon(epsilon):
  """A probe width a reference script passes explicitly (grad_check(..., epsilon=1e-7), the reference's own default) is
  below fp32 resolution: it is widened to the device default, with a one-line notice, instead of reporting a bogus FAILED."""
  if epsilon < MIN_EPSILON:
    import warnings
    warnings.warn(f'neograd_b200: grad_check epsilon {epsilon:g} is below fp32 resolution; using {DEFAULT_EPSILON:g} '
                  '(values on the device are float32, the reference computes in float64)', stacklevel=3)
    return DEFAULT_EPSILON
  return epsilon


def _evaluate_grad_check(analytical_grads, calculated_grads, tolerance, print_vals):
  """utils.py:136-160."""
  a = np.asarray(analytical_grads, dtype=np.float64)
  c = np.asarray(calculated_grads, dtype=np.float64)
  dist = np.linalg.norm(a - c) / (np.linalg.norm(a) + np.linalg.norm(c))
  if print_vals:
    print('Gradient Check Distance:', dist)
    print('Gradient Check PASSED' if dist < tolerance else 'Gradient Check FAILED')
  return dist


def _wiggle_params(analytical_grads, calculated_grads, params, get_loss, epsilon):
  """utils.py:163-192: central difference for every scalar element of every param."""
  for param in params:
    if not param.requires_grad:
      continue
    if not isinstance(param.grad, DeviceArray):
      param.grad = DeviceArray.zeros(param.shape)
    host = param.data.numpy()
    grad = param.grad.numpy()
    for idx in np.ndindex(param.shape):
      with no_track():
        saved = host[idx]
        host[idx] = saved + epsilon
        param.data = host.copy()
        loss1 = get_loss()
        host[idx] = saved - epsilon
        param.data = host.copy()
        loss2 = get_loss()
        host[idx] = saved
        param.data = host.copy()
      calculated_grads.append((float(loss1.data) - float(loss2.data)) / (2 * epsilon))
      analytical_grads.append(float(grad[idx]))
    param.zero_grad()


def grad_check(model, inputs, targets, loss_fn, epsilon=DEFAULT_EPSILON, print_vals=True, tolerance=DEFAULT_TOLERANCE):
  """utils.py:195-234.  Returns the distance between analytic and finite-difference gradients."""
  from .._backend import set_precision
  epsilon = _device_epsilon(epsilon)
  prev = set_precision('fp32')
  try:
    return _grad_check(model, inputs, targets, loss_fn, epsilon, print_vals, tolerance)
  finally:
    set_precision(prev)


def _grad_check(model, inputs, targets, loss_fn, epsilon, print_vals, tolerance):
  params = model.parameters()
  analytical_grads, calculated_grads = [], []
  for param in params:
    param.zero_grad()

  def get_loss():
    return loss_fn(model(inputs), targets)

  with new_graph():
    loss = get_loss()
    loss.backward()
    _wiggle_params(analytical_grads, calculated_grads, params, get_loss, epsilon)
  return _evaluate_grad_check(analytical_grads, calculated_grads, tolerance, print_vals)


def fn_grad_check(fn, inputs, params, targets=None, loss_fn=None, epsilon=DEFAULT_EPSILON, print_vals=True,
                  tolerance=DEFAULT_TOLERANCE, **kwargs):
  """utils.py:237-280: same for a bare function of tensors; default loss is MSE against all ones."""
  from .._backend import set_precision
  epsilon = _device_epsilon(epsilon)
  prev = set_precision('fp32')
  try:
    return _fn_grad_check(fn, inputs, params, targets, loss_fn, epsilon, print_vals, tolerance, **kwargs)
  finally:
    set_precision(prev)


def _fn_grad_check(fn, inputs, params, targets, loss_fn, epsilon, print_vals, tolerance, **kwargs):
  if loss_fn is None:
    from ..nn.loss import MSE
    loss_fn = MSE()
  analytical_grads, calculated_grads = [], []
  for param in params:
    param.zero_grad()

  def get_loss(targets=targets):
    outputs = fn(*inputs, **kwargs)
    if targets is None:
      from .tensor import Tensor
      targets = Tensor(np.ones(outputs.shape))
    return loss_fn(outputs, targets)

  with new_graph():
    loss = get_loss()
    loss.backward()
    _wiggle_params(analytical_grads, calculated_grads, params, get_loss, epsilon)
  return _evaluate_grad_check(analytical_grads, calculated_grads, tolerance, print_vals)
