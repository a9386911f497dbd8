"""Model shell (reference: neograd/nn/model.py): layer discovery, parameters(), eval mode, save / load."""
import pickle
from itertools import chain

from .layers import Container, Layer
from ..autograd.utils import get_graph

try:
  import dill as _pickler   # the reference's serializer (model.py:1); files are interchangeable
except ImportError:          # pragma: no cover
  _pickler = pickle


class Model:
  def __call__(self, inputs):
    return self.forward(inputs)

  def eval(self, no_track=True):
    return EvalMode(self, no_track)

  def get_layers(self):
    return {attr: val for attr, val in self.__dict__.items() if isinstance(val, (Container, Layer))}

  def parameters(self, as_dict=False):
    """model.py:45-56: flat list in layer-definition order, or the nested {attr: ...} dict of float64 ndarrays that
    save() writes."""
    params = {attr: layer.parameters(as_dict) for attr, layer in self.get_layers().items()}
    return params if as_dict else list(chain(*params.values()))

  def set_eval(self, eval):
    for layer in self.get_layers().values():
      layer.set_eval(eval)

  def save(self, fpath):
    """model.py:69-78: device -> host, float64, same nested-dict file format as the reference."""
    params = self.parameters(as_dict=True)
    with open(fpath, 'wb') as fp:
      _pickler.dump(params, fp)
    print(f"\nPARAMS SAVED at {fpath}\n")

  def load(self, fpath):
    """model.py:80-91 (host -> device through the Param.data setter, so optimizer arenas stay valid)."""
    with open(fpath, 'rb') as fp:
      params = _pickler.load(fp)
    for attr, param in params.items():
      self.__getattribute__(attr).set_params(param)
    print(f"\nPARAMS LOADED from {fpath}\n")

  def __setattr__(self, attr, val):
    if isinstance(val, (Container, Layer)) and attr in self.__dict__:
      raise AttributeError(f"Attribute {attr} has already been defined, it cannot be defined again for a Container/Layer")
    object.__setattr__(self, attr, val)

  def __repr__(self):
    return f'Model( {[str(layer) for layer in self.get_layers().values()]} )'

  __str__ = __repr__


class EvalMode:
  """model.py:114-148."""

  def __init__(self, model, no_track):
    self.model = model
    self.no_track = no_track
    self.graph = get_graph()

  def __enter__(self):
    if self.no_track:
      self.graph.track = False
    self.model.set_eval(True)

  def __exit__(self, exc_type, exc_value, exc_traceback):
    if self.no_track:
      self.graph.track = True
    self.model.set_eval(False)
