"""Operation: the extension protocol of neograd (reference: neograd/autograd/ops/operation.py, README.md:44-52).

  forward():  tensors = self.get_tensors(*operands) -> compute on `tens.data` (DeviceArrays; kernels from
              libngb200.so) -> `return self.get_result_tensor(result_array, *tensors)`
  backward(): called by the engine with the operand tensors; installs one grad_fn per operand via
              `tens.set_grad_fn(...)`.  Built-in ops install `GradFn` objects (fused unbroadcast + accumulate);
              user ops may install any callable `ug -> grad array` exactly as with the reference.
"""
import numpy as np

from ..node import Node
from ..utils import get_graph

_TENSOR = None


def _tensor_cls():
  global _TENSOR
  if _TENSOR is None:
    from ..tensor import Tensor
    _TENSOR = Tensor
  return _TENSOR


class Operation:
  def process_operands(self, operands):
    Tensor = _tensor_cls()
    return tuple(o if isinstance(o, Tensor) else Tensor(o) for o in operands)   # operation.py:13-27

  def get_tensors(self, *operands):
    tensors = self.process_operands(operands)
    if len(tensors) == 0:
      return None
    if len(tensors) == 1:
      return tensors[0]
    return tensors

  def get_broadcast_shape(self, *tensors):
    """operation.py:46-66: None when any operand opts out of broadcasting or the shapes are incompatible."""
    for tens in tensors:
      if not tens.requires_broadcasting:
        return None
    try:
      return np.broadcast_shapes(*(tens.shape for tens in tensors))
    except ValueError:
      return None

  def result_requires_grad(self, tensors):
    return any(tens.requires_grad for tens in tensors)

  def get_result_tensor(self, result, *tensors):
    """operation.py:83-109 (without the object-dtype round trip: `result` already lives on the device)."""
    Tensor = _tensor_cls()
    graph = get_graph()
    result_tensor = Tensor(result, self.result_requires_grad(tensors))
    if graph.track:
      result_node = Node(result_tensor)
      result_node.backward_fn = self.backward
      result_node.parent_broadcast_shape = self.get_broadcast_shape(*tensors)
      graph.add_edge(result_node, tensors)
    return result_tensor

  def backward(self, *args):
    raise NotImplementedError(f"Backward method not implemented for Operation {self}")
