"""save_model / load_model / get_batches (reference: neograd/nn/utils.py)."""
from .model import Model, _pickler


def save_model(fpath, model):
  """Pickles a whole Model object (structure only: layers pickle their Params as zero-valued stand-ins, see
  layers/__init__.py; weights travel through Model.save).  TypeError for anything that is not a Model (nn/utils.py:5-21)."""
  if not isinstance(model, Model):
    raise TypeError(f'Expected Model object, instead got {type(model)}')
  with open(fpath, 'wb') as fp:
    _pickler.dump(model, fp)
  print(f'MODEL SAVED at {fpath}')


def load_model(fpath):
  """Inverse of save_model (nn/utils.py:24-36)."""
  with open(fpath, 'rb') as fp:
    restored = _pickler.load(fp)
  print(f'MODEL LOADED from {fpath}')
  return restored


def get_batches(inputs, targets=None, batch_size=None):
  """Generator over mini-batches along axis 0 (public behaviour of nn/utils.py:39-81: the last batch may be short; a batch
  size of None means one batch with everything; ValueError for sizes that are zero, negative or larger than the data).
  Each batch is a leading-axis slice of the device tensors -- a view of the same HBM, no copy."""
  total = inputs.shape[0]
  if targets is not None:
    assert total == targets.shape[0], '0th dim should be number of examples and should match'
  if batch_size is None:
    batch_size = total
  else:
    for bad, why in ((batch_size > total, "Batch size cannot be greater than the number of examples"),
                     (batch_size < 0, "Batch size cannot be negative"), (batch_size == 0, "Batch size cannot be zero")):
      if bad:
        raise ValueError(why)
  for begin in range(0, total, batch_size):
    rows = slice(begin, min(begin + batch_size, total))
    yield (inputs[rows], targets[rows]) if targets is not None else inputs[rows]
