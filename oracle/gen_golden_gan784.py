"""Golden fixture for config C4 at the BASELINE dims (784-512-256) with the UNMODIFIED reference, at a batch
the reference still finishes in seconds (512): three generator + discriminator iterations of examples/gans/basic.ipynb
cells 9-10 (freeze / unfreeze, zero_grad(all_members=True), BCE, two Adam optimizers).  TEST INFRASTRUCTURE; build
container only.

  python oracle/gen_golden_gan784.py   # writes tests/golden/step_c4_gan784.npz and adds it to MANIFEST.json

The 1 268 241 parameters and the inputs are regenerated from seeds by the tests (`gan784_inputs`), so the fixture keeps
only what the reference computed: the losses, and for every parameter its first-step gradient and its final value,
each as (L2 norm, every `STRIDE`-th element).

Learning rate: the notebook's Adam(5e-3) moves every weight by 5e-3 in its gradient's sign direction per step; with 784
coherent inputs the discriminator's sigmoid saturates to exactly 1.0 in float32 after ONE step (the reference's own
float64 run gives loss_g = 5.7e-6 at the second iteration, i.e. mean(1 - D) = 2.7e-7, below float32 resolution), so a
multi-step trajectory at that rate compares nothing but saturation.  The first-step gradients and losses recorded here do
not depend on the rate; the 3-iteration trajectory uses LR = 3e-4, which keeps the losses O(10)."""
import json
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle.gen_golden import OUT, load_reference  # noqa: E402
from oracle.models import GAN784_DIMS as DIMS, GAN784_STRIDE as STRIDE, gan784_inputs, sample  # noqa: E402

LR = 3e-4


def main():
  ng = load_reference()
  nn = ng.nn
  from neograd_ref.nn import optim as roptim, loss as rloss
  d, h1, h2 = DIMS

  class GAN(nn.Model):
    def __init__(self):
      self.generator = nn.Sequential(nn.Linear(d, h1), nn.LeakyReLU(), nn.Linear(h1, h2), nn.LeakyReLU(), nn.Linear(h2, d))
      self.discriminator = nn.Sequential(nn.Linear(d, h1), nn.LeakyReLU(), nn.Linear(h1, h2), nn.LeakyReLU(),
                                         nn.Linear(h2, 1), nn.Sigmoid())

    def forward(self, noise):
      return self.discriminator(self.generator(noise))

  B, iters = 512, 3
  params, real, noise_g, noise_d = gan784_inputs(B, iters)
  np.random.seed(0)
  model = GAN()
  for p, v in zip(model.parameters(), params):
    p.data = v
  optim_g, optim_d = roptim.Adam(model.parameters(), LR), roptim.Adam(model.parameters(), LR)
  loss_fn = rloss.BCE()
  ones, zeros = ng.tensor(np.ones((B, 1))), ng.tensor(np.zeros((B, 1)))
  arrs = {'batch': np.array(B), 'iters': np.array(iters), 'stride': np.array(STRIDE), 'lr': np.array(LR)}
  losses_g, losses_d = [], []
  for it in range(iters):
    model.discriminator.freeze()
    optim_g.zero_grad(all_members=True)
    loss_g = loss_fn(model(ng.tensor(noise_g[it])), ones)
    loss_g.backward()
    if it == 0:
      for i, p in enumerate(model.parameters()[:6]):
        arrs[f'grad_{i}'] = sample(p.grad)
    optim_g.step()
    model.discriminator.unfreeze()
    model.generator.freeze()
    optim_d.zero_grad(all_members=True)
    fake = model.generator(ng.tensor(noise_d[it]))
    loss_d = loss_fn(model.discriminator(ng.tensor(real)), ones) + loss_fn(model.discriminator(fake), zeros)
    loss_d.backward()
    if it == 0:
      for i, p in enumerate(model.parameters()[6:]):
        arrs[f'grad_{i + 6}'] = sample(p.grad)
    optim_d.step()
    model.generator.unfreeze()
    losses_g.append(float(loss_g.data))
    losses_d.append(float(loss_d.data))
  arrs['losses_g'], arrs['losses_d'] = np.array(losses_g), np.array(losses_d)
  for i, p in enumerate(model.parameters()):
    arrs[f'final_{i}'] = sample(p.data)
  np.savez_compressed(os.path.join(OUT, 'step_c4_gan784.npz'), **arrs)
  mpath = os.path.join(OUT, 'MANIFEST.json')
  with open(mpath) as fp:
    man = json.load(fp)
  man['fixtures']['step_c4_gan784'] = sorted(arrs.keys())
  man['generator_c4_784'] = 'oracle/gen_golden_gan784.py'
  with open(mpath, 'w') as fp:
    json.dump(man, fp, indent=1, sort_keys=True)
  print('wrote step_c4_gan784:', losses_g, losses_d)


if __name__ == '__main__':
  main()
