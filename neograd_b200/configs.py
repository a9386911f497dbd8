"""The BASELINE.json configurations bound to this package (definitions: workload_defs.py, shared with the reference arm
of bench.py).  Used by bench.py, the parity tests and smoke(); nothing here is special-cased by the backend."""
import numpy as np  # noqa: F401

from . import nn
from .workload_defs import (model_classes, set_params, synthetic_params, lenet_data, gan_data, train_step,  # noqa: F401
                            gan_iteration)

_C = model_classes(nn)
for _cls in _C.values():       # module-level citizens of THIS module, so pickle / dill store them by reference (ng.save(model))
  _cls.__module__, _cls.__qualname__ = __name__, _cls.__name__
MLPCircles, DigitsCNN, LeNetA, LeNetB, GAN = _C['MLPCircles'], _C['DigitsCNN'], _C['LeNetA'], _C['LeNetB'], _C['GAN']


def lenet_synthetic(kind, batch, seed=0, scale=0.1):
  """X ~ N(0,1), one-hot labels, weights 0.1*randn (seed+1), biases 0."""
  x, y = lenet_data(kind, batch, seed)
  model = (LeNetA if kind == 'a' else LeNetB)()
  set_params(model, synthetic_params([p.shape for p in model.parameters()], seed + 1, scale))
  return model, x, y


def gan_synthetic(batch, data_dim=784, h1=512, h2=256, seed=0, scale=0.05):
  """real rows ~ N(0,1), two noise batches, weights 0.05*randn (seed+1), biases 0."""
  real, noise_g, noise_d = gan_data(batch, data_dim, seed)
  model = GAN(data_dim, h1, h2)
  set_params(model, synthetic_params([p.shape for p in model.parameters()], seed + 1, scale))
  return model, real, noise_g, noise_d
