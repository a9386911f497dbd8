"""The BASELINE.json configurations as a neograd USER would write them (README.md:95-115, examples/digits.ipynb cell 7,
SURVEY 8d C3A / C3B, examples/gans/basic.ipynb cells 7-11), written against ANY neograd-style package: nothing here
imports neograd_b200, so bench.py's reference arm can load this file by path and build the very same models, data and
training loops on the unmodified reference."""
import numpy as np


def model_classes(nn):
  """The five model definitions written against a neograd-style `nn` namespace -- this package's, or the unmodified
  reference's (bench.py --impl reference builds the very same classes on `neograd.nn`)."""

  class MLPCircles(nn.Model):
    """C1: Linear(2,100) ReLU Linear(100,1) Sigmoid."""

    def __init__(self):
      self.stack = nn.Sequential(nn.Linear(2, 100), nn.ReLU(), nn.Linear(100, 1), nn.Sigmoid())

    def forward(self, inputs):
      return self.stack(inputs)

  class DigitsCNN(nn.Model):
    """C2: Conv2D((3,3)) ReLU reshape(B,36) Linear(36,10)."""

    def __init__(self):
      self.conv = nn.Sequential(nn.Conv2D((3, 3)), nn.ReLU())
      self.stack = nn.Sequential(nn.Linear(36, 10))

    def forward(self, inputs):
      conv_outputs = self.conv(inputs)
      return self.stack(conv_outputs.reshape((inputs.shape[0], 36)))

  class LeNetA(nn.Model):
    """C3A: neograd's single-channel Conv2D / MaxPool2D on (B,28,28); 13 106 parameters."""

    def __init__(self):
      self.conv = nn.Sequential(nn.Conv2D((5, 5)), nn.ReLU(), nn.MaxPool2D((2, 2), stride=2),
                                nn.Conv2D((5, 5)), nn.ReLU(), nn.MaxPool2D((2, 2), stride=2))
      self.stack = nn.Sequential(nn.Linear(16, 120), nn.ReLU(), nn.Linear(120, 84), nn.ReLU(), nn.Linear(84, 10))

    def forward(self, inputs):
      return self.stack(self.conv(inputs).reshape((inputs.shape[0], 16)))

  class LeNetB(nn.Model):
    """C3B: LeNet-5 channels with Conv3D(1,6,5x5) / MaxPool3D / Conv3D(6,16,5x5) on (B,1,28,28); 44 426 parameters."""

    def __init__(self):
      self.conv = nn.Sequential(nn.Conv3D(1, 6, (5, 5)), nn.ReLU(), nn.MaxPool3D((2, 2), stride=2),
                                nn.Conv3D(6, 16, (5, 5)), nn.ReLU(), nn.MaxPool3D((2, 2), stride=2))
      self.stack = nn.Sequential(nn.Linear(256, 120), nn.ReLU(), nn.Linear(120, 84), nn.ReLU(), nn.Linear(84, 10))

    def forward(self, inputs):
      return self.stack(self.conv(inputs).reshape((inputs.shape[0], 256)))

  class GAN(nn.Model):
    """C4: MLP generator + discriminator (examples/gans/basic.ipynb cell 7 with the dims of SURVEY 8d)."""

    def __init__(self, data_dim=784, h1=512, h2=256):
      self.data_dim = data_dim
      self.generator = nn.Sequential(nn.Linear(data_dim, h1), nn.LeakyReLU(), nn.Linear(h1, h2), nn.LeakyReLU(),
                                     nn.Linear(h2, data_dim))
      self.discriminator = nn.Sequential(nn.Linear(data_dim, h1), nn.LeakyReLU(), nn.Linear(h1, h2), nn.LeakyReLU(),
                                         nn.Linear(h2, 1), nn.Sigmoid())

    def forward(self, noise):
      return self.discriminator(self.generator(noise))

  return {'MLPCircles': MLPCircles, 'DigitsCNN': DigitsCNN, 'LeNetA': LeNetA, 'LeNetB': LeNetB, 'GAN': GAN}


def set_params(model, arrays):
  """Assign host arrays to model.parameters() in order, through the Param.data setter."""
  params = model.parameters()
  assert len(params) == len(arrays), (len(params), len(arrays))
  for p, a in zip(params, arrays):
    p.data = np.asarray(a, dtype=np.float64).reshape(p.shape)


def synthetic_params(shapes, seed, scale):
  """weights scale*randn, biases 0, drawn in Model.parameters() order (weights, bias per layer)"""
  wrng = np.random.default_rng(seed)
  return [scale * wrng.standard_normal(s) if i % 2 == 0 else np.zeros(s) for i, s in enumerate(shapes)]


def lenet_data(kind, batch, seed=0):
  """SURVEY 8d C3 inputs: X ~ N(0,1) (B,28,28) / (B,1,28,28), one-hot labels (B,10)."""
  rng = np.random.default_rng(seed)
  x = rng.standard_normal((batch, 28, 28) if kind == 'a' else (batch, 1, 28, 28)).astype(np.float32)
  y = np.eye(10, dtype=np.float32)[rng.integers(0, 10, batch)]
  return x, y


def train_step(model, optim, loss_fn, X, Y):
  """README.md:109-115."""
  optim.zero_grad()
  loss = loss_fn(model(X), Y)
  loss.backward()
  optim.step()
  return loss


def gan_iteration(model, optim_g, optim_d, loss_fn, noise_g, noise_d, real, ones, zeros):
  """examples/gans/basic.ipynb cells 9-10, verbatim choreography: generator step with the discriminator frozen, then
  discriminator step with the generator frozen.  All arguments are Tensors; returns (loss_g, loss_d)."""
  model.discriminator.freeze()
  optim_g.zero_grad(all_members=True)
  loss_g = loss_fn(model(noise_g), ones)
  loss_g.backward()
  optim_g.step()
  model.discriminator.unfreeze()
  model.generator.freeze()
  optim_d.zero_grad(all_members=True)
  fake = model.generator(noise_d)
  loss_d = loss_fn(model.discriminator(real), ones) + loss_fn(model.discriminator(fake), zeros)
  loss_d.backward()
  optim_d.step()
  model.generator.unfreeze()
  return loss_g, loss_d


def gan_data(batch, data_dim=784, seed=0):
  """SURVEY 8d C4 inputs: real rows ~ N(0,1) and the two noise batches of one iteration."""
  rng = np.random.default_rng(seed)
  return tuple(rng.standard_normal((batch, data_dim)).astype(np.float32) for _ in range(3))
