"""save_model / load_model / get_batches (reference: neograd/nn/utils.py)."""
from .model import Model, _pickler


def save_model(fpath, model):
  if not isinstance(model, Model):
    raise TypeError(f'Expected Model object, instead got {type(model)}')
  with open(fpath, 'wb') as fp:
    _pickler.dump(model, fp)
    print(f'MODEL SAVED at {fpath}')


def load_model(fpath):
  with open(fpath, 'rb') as fp:
    model = _pickler.load(fp)
    print(f'MODEL LOADED from {fpath}')
  return model


def get_batches(inputs, targets=None, batch_size=None):
  """nn/utils.py:39-81.  Slices are views of the same HBM (Tensor.__getitem__): batching a device-resident dataset
  moves no data."""
  if targets is not None:
    assert inputs.shape[0] == targets.shape[0], '0th dim should be number of examples and should match'
  num_examples = inputs.shape[0]
  if batch_size is not None:
    if batch_size > num_examples:
      raise ValueError("Batch size cannot be greater than the number of examples")
    elif batch_size < 0:
      raise ValueError("Batch size cannot be negative")
    elif batch_size == 0:
      raise ValueError("Batch size cannot be zero")
  else:
    batch_size = num_examples
  start = 0
  while start < num_examples:
    end = start + batch_size if start + batch_size < num_examples else num_examples
    if targets is not None:
      yield inputs[start:end], targets[start:end]
    else:
      yield inputs[start:end]
    start = end
