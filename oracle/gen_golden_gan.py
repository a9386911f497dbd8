"""Golden fixture for config C4 (examples/gans/basic.ipynb cells 7-10 of the reference, at notebook scale 64-50-25,
batch 200): generator / discriminator steps with freeze / unfreeze, zero_grad(all_members=True), BCE and two Adam
optimizers, run with the UNMODIFIED reference in the build container (test infrastructure only).

  python oracle/gen_golden_gan.py      # writes tests/golden/step_c4_gan.npz and adds it to tests/golden/MANIFEST.json
"""
import json
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle.gen_golden import OUT, f32, load_reference  # noqa: E402


def main():
  ng = load_reference()
  nn = ng.nn
  from neograd_ref.nn import optim as roptim, loss as rloss

  class GAN(nn.Model):
    def __init__(self):
      self.generator = nn.Sequential(nn.Linear(64, 50), nn.LeakyReLU(), nn.Linear(50, 25), nn.LeakyReLU(), nn.Linear(25, 64))
      self.discriminator = nn.Sequential(nn.Linear(64, 50), nn.LeakyReLU(), nn.Linear(50, 25), nn.LeakyReLU(),
                                         nn.Linear(25, 1), nn.Sigmoid())

    def forward(self, noise):
      return self.discriminator(self.generator(noise))

  np.random.seed(0)
  model = GAN()
  rng = np.random.default_rng(4)
  for p in model.parameters():
    p.data = f32(0.05 * rng.standard_normal(p.shape))       # plain randn saturates the sigmoid
  init = [np.array(p.data) for p in model.parameters()]
  B, iters = 200, 3
  real = f32(rng.standard_normal((B, 64)))
  noise_g = f32(rng.standard_normal((iters, B, 64)))
  noise_d = f32(rng.standard_normal((iters, B, 64)))
  optim_g, optim_d = roptim.Adam(model.parameters(), 5e-3), roptim.Adam(model.parameters(), 5e-3)
  loss_fn = rloss.BCE()
  ones, zeros = ng.tensor(np.ones((B, 1))), ng.tensor(np.zeros((B, 1)))
  losses_g, losses_d = [], []
  for it in range(iters):
    # generator step (cell 9)
    model.discriminator.freeze()
    optim_g.zero_grad(all_members=True)
    loss_g = loss_fn(model(ng.tensor(noise_g[it])), ones)
    loss_g.backward()
    optim_g.step()
    model.discriminator.unfreeze()
    # discriminator step (cell 10)
    model.generator.freeze()
    optim_d.zero_grad(all_members=True)
    fake = model.generator(ng.tensor(noise_d[it]))
    loss_d = loss_fn(model.discriminator(ng.tensor(real)), ones) + loss_fn(model.discriminator(fake), zeros)
    loss_d.backward()
    optim_d.step()
    model.generator.unfreeze()
    losses_g.append(float(loss_g.data))
    losses_d.append(float(loss_d.data))
  arrs = {'real': real.astype(np.float32), 'noise_g': noise_g.astype(np.float32), 'noise_d': noise_d.astype(np.float32),
          'losses_g': np.array(losses_g), 'losses_d': np.array(losses_d)}
  for i, (p_i, p_f) in enumerate(zip(init, model.parameters())):
    arrs[f'init_{i}'] = p_i
    arrs[f'final_{i}'] = np.array(p_f.data)
  np.savez_compressed(os.path.join(OUT, 'step_c4_gan.npz'), **arrs)
  mpath = os.path.join(OUT, 'MANIFEST.json')
  with open(mpath) as fp:
    man = json.load(fp)
  man['fixtures']['step_c4_gan'] = sorted(arrs.keys())
  man['generator_c4'] = 'oracle/gen_golden_gan.py'
  with open(mpath, 'w') as fp:
    json.dump(man, fp, indent=1, sort_keys=True)
  print('wrote step_c4_gan:', losses_g, losses_d)


if __name__ == '__main__':
  main()
