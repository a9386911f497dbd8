from .tensor import Tensor as tensor
from .ops import add, sub, mul, div, pow, exp, log, dot, sum, transpose, flatten, reshape
from .utils import new_graph, no_track
