"""Layer / Container / Param shells (reference: neograd/nn/layers/__init__.py).  No arithmetic lives here: these
classes discover parameters (in `__dict__` order, which fixes the order of `Model.parameters()` and therefore the
layout of the optimizer's flat arena), toggle eval mode, freeze / unfreeze and guard against re-assigning a Param."""
import numpy as np

from ...autograd.tensor import Tensor


class Param(Tensor):
  """layers/__init__.py:192-229: a Tensor that layers report as trainable.  Always materialised on the device."""

  def __init__(self, data, requires_grad=False, requires_broadcasting=True):
    super().__init__(data, requires_grad, requires_broadcasting)
    self.data  # noqa: B018  (forces the upload of scalar params such as Conv2D's bias, layers/conv.py:28)
    self.__frozen = False

  def freeze(self):
    self.requires_grad = False
    self.__frozen = True

  def unfreeze(self):
    if self.__frozen:
      self.requires_grad = True
    self.__frozen = False

  def __str__(self):
    return f'Param({super().__str__()})'

  def __repr__(self):
    return f'Param({super().__repr__()})'


class Container:
  """layers/__init__.py:5-88."""

  def __init__(self):
    self.eval = False
    self.layers = None

  def __call__(self, inputs):
    return self.forward(inputs)

  def parameters(self, as_dict=False):
    params = []
    for layer in self.layers:
      if as_dict:
        params.append(layer.parameters(as_dict))
      else:
        params += layer.parameters(as_dict)
    return params

  def set_eval(self, eval):
    self.eval = eval
    for layer in self.layers:
      layer.set_eval(eval)

  def set_params(self, container_params):
    for layer_param, layer in zip(container_params, self.layers):
      layer.set_params(layer_param)

  def freeze(self):
    for layer in self.layers:
      layer.freeze()

  def unfreeze(self):
    for layer in self.layers:
      layer.unfreeze()

  def __repr__(self):
    return ', '.join(str(layer) for layer in self.layers)

  def __str__(self):
    return ', '.join(repr(layer) for layer in self.layers)


class Layer:
  """layers/__init__.py:91-189."""

  def __init__(self):
    self.eval = False

  def __call__(self, inputs):
    return self.forward(inputs)

  def parameters(self, as_dict=False):
    """Params in attribute-definition order; `as_dict` gives {attr: float64 ndarray} (the checkpoint format,
    model.py:69-78), which costs a device -> host copy per param."""
    params = {}
    for attr, val in self.__dict__.items():
      if isinstance(val, Param):
        params[attr] = val.data.numpy() if as_dict else val
    return params if as_dict else list(params.values())

  def set_eval(self, eval):
    self.eval = eval

  def set_params(self, layer_params):
    for attr, param_data in layer_params.items():
      param = self.__getattribute__(attr)
      param.data = np.asarray(param_data, dtype=np.float64) if not isinstance(param_data, (int, float, list)) else param_data

  def freeze(self):
    for param in self.parameters(as_dict=False):
      param.freeze()

  def unfreeze(self):
    for param in self.parameters(as_dict=False):
      param.unfreeze()

  def __getstate__(self):
    """layers/__init__.py:160-173: a pickled layer carries its structure, not its weights (they are saved separately
    by Model.save); Params are replaced by zero-valued stand-ins."""
    state = dict(self.__dict__)
    for attr, val in self.__dict__.items():
      if isinstance(val, Param):
        state[attr] = _ParamStub(val.shape, val.requires_grad, val.requires_broadcasting)
    return state

  def __setstate__(self, state):
    for attr, val in state.items():
      if isinstance(val, _ParamStub):
        val = Param(np.zeros(val.shape), val.requires_grad, val.requires_broadcasting)
      object.__setattr__(self, attr, val)

  def __setattr__(self, attr, val):
    if isinstance(val, Param) and attr in self.__dict__:
      raise AttributeError(f"Attribute {attr} has already been defined, it cannot be defined again for a Param")
    object.__setattr__(self, attr, val)


class _ParamStub:
  def __init__(self, shape, requires_grad, requires_broadcasting):
    self.shape, self.requires_grad, self.requires_broadcasting = shape, requires_grad, requires_broadcasting


# after Layer / Container / Param exist (circular import, as in the reference)
from .misc import Sequential, Linear, Dropout  # noqa: E402
from .conv import Conv2D, Conv3D, MaxPool2D, MaxPool3D  # noqa: E402
