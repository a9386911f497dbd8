"""Model: the user-facing container of a network (public surface of neograd/nn/model.py: `__call__`, `eval()`,
`parameters()`, `get_layers()`, `set_eval()`, `save()`, `load()`, re-assignment guard; file format of save / load is the
reference's nested dict of float64 ndarrays, model.py:69-91, so files are interchangeable).

A model owns no arithmetic: it finds its layers among its attributes (definition order fixes the order of
`parameters()`, which is also the layout of the optimizer's flat device arena) and moves parameter VALUES between the
device and files."""
import pickle

from .layers import Container, Layer
from ..autograd.utils import get_graph

try:
  import dill as _pickler   # the reference serialises with dill (model.py:1)
except ImportError:          # pragma: no cover
  _pickler = pickle

_LAYER_TYPES = (Container, Layer)


class Model:
  def __call__(self, inputs):
    return self.forward(inputs)

  # ---------------------------------------------------------------- structure
  def get_layers(self):
    """{attribute name: layer} for every Container / Layer attribute, in definition order."""
    found = {}
    for name, value in vars(self).items():
      if isinstance(value, _LAYER_TYPES):
        found[name] = value
    return found

  def parameters(self, as_dict=False):
    """Flat list of Params (default) or {attribute: nested host values} as written by save()."""
    layers = self.get_layers()
    if as_dict:
      return {name: layer.parameters(True) for name, layer in layers.items()}
    flat = []
    for layer in layers.values():
      flat.extend(layer.parameters(False))
    return flat

  def __setattr__(self, attr, val):
    already = attr in vars(self)
    if already and isinstance(val, _LAYER_TYPES):
      raise AttributeError(f"Attribute {attr} has already been defined, it cannot be defined again for a Container/Layer")
    object.__setattr__(self, attr, val)

  # ---------------------------------------------------------------- modes
  def set_eval(self, eval):
    for layer in self.get_layers().values():
      layer.set_eval(eval)

  def eval(self, no_track=True):
    """`with model.eval():` -- layers in eval mode and (by default) no graph tracking inside the block."""
    return EvalMode(self, no_track)

  # ---------------------------------------------------------------- persistence
  def save(self, fpath):
    """Device -> host (float64) -> file."""
    with open(fpath, 'wb') as fp:
      _pickler.dump(self.parameters(as_dict=True), fp)
    print(f"\nPARAMS SAVED at {fpath}\n")

  def load(self, fpath):
    """File -> device, through each layer's set_params (i.e. the Param.data setter: parameters registered with an
    optimizer are overwritten in place, so its arena stays valid)."""
    with open(fpath, 'rb') as fp:
      stored = _pickler.load(fp)
    for name, values in stored.items():
      getattr(self, name).set_params(values)
    print(f"\nPARAMS LOADED from {fpath}\n")

  def __repr__(self):
    return f'Model( {[str(layer) for layer in self.get_layers().values()]} )'

  __str__ = __repr__


class EvalMode:
  """Context manager returned by Model.eval(): restores training mode (and tracking) on exit."""

  def __init__(self, model, no_track):
    self.model, self.no_track = model, no_track
    self.graph = get_graph()

  def __enter__(self):
    self.model.set_eval(True)
    if self.no_track:
      self.graph.track = False

  def __exit__(self, exc_type, exc_value, exc_traceback):
    self.model.set_eval(False)
    if self.no_track:
      self.graph.track = True
