from .model import Model
from .layers import Sequential, Dropout, Linear, Conv2D, Conv3D, MaxPool2D, MaxPool3D
from .activations import ReLU, Sigmoid, Tanh, Softmax, LeakyReLU
from .checkpoint import Checkpoint
from .utils import save_model, load_model, get_batches
from .data import DeviceLoader, randn
from . import loss, optim
