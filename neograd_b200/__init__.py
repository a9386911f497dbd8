"""neograd_b200 -- neograd's tensor-op hot path on B200 (sm_100a).

Same public surface as pranftw/neograd (`import neograd_b200 as ng`: ng.tensor, ng.nn.Model, ng.nn.Linear / Conv2D /
Conv3D / MaxPool2D / MaxPool3D, ng.nn.loss.*, ng.nn.optim.*, ng.new_graph, ng.no_track, ...; reference:
neograd/__init__.py:1-10) with `Tensor.data` resident in HBM and every op body a hand-written CUDA kernel from
libngb200.so (C ABI: include/ngb200.h).  Importing fails loudly when the library has not been built; there is no CPU
fallback.
"""
from . import _backend  # noqa: F401  (loads libngb200.so and checks every symbol of include/ngb200.h)
from .autograd.graph import Graph

_NG_GRAPH = Graph()   # the global graph (neograd/__init__.py:9-10)

from . import autograd, nn, distributed  # noqa: E402
from .nn import Checkpoint  # noqa: E402
from .autograd import tensor, new_graph, no_track  # noqa: E402
from .autograd import add, sub, mul, div, pow, exp, log, dot, sum, transpose, flatten, reshape  # noqa: E402
from .autograd.utils import grad_check, fn_grad_check  # noqa: E402
from .nn.utils import load_model as load, save_model as save  # noqa: E402
from .device import DeviceArray, synchronize  # noqa: E402
from ._backend import set_precision, set_backward_lanes  # noqa: E402
