import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
  sys.path.insert(0, ROOT)

GOLDEN = os.path.join(ROOT, 'tests', 'golden')


def pytest_configure(config):
  config.addinivalue_line('markers', 'gpu: needs a CUDA device (run on the B200 box with -m gpu)')


def golden_names(prefix):
  return sorted(f[:-4] for f in os.listdir(GOLDEN) if f.startswith(prefix) and f.endswith('.npz'))


def load_golden(name):
  import numpy as np
  with np.load(os.path.join(GOLDEN, name + '.npz'), allow_pickle=False) as z:
    return {k: z[k] for k in z.files}


@pytest.fixture(scope='session')
def has_cuda():
  import torch
  return torch.cuda.is_available()
