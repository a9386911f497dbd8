from .basics import add, sub, mul, div, dot, exp, log, pow, sum, transpose, flatten, reshape
from .conv import conv2d, conv3d, maxpool2d, maxpool3d
