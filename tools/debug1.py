import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import neograd_b200 as ng
from neograd_b200.configs import GAN, MLPCircles
nn = ng.nn
np.random.seed(0)
model = GAN(64, 50, 25)
rng = np.random.default_rng(0)
real = ng.tensor(rng.standard_normal((200, 64)))
optim_g, optim_d = nn.optim.Adam(model.parameters(), 5e-3), nn.optim.Adam(model.parameters(), 5e-3)
loss_fn = nn.loss.BCE()
ones, zeros = ng.tensor(np.ones((200, 1))), ng.tensor(np.zeros((200, 1)))
before = [p.data.numpy().copy() for p in model.parameters()]
model.discriminator.freeze()
optim_g.zero_grad(all_members=True)
noise = ng.tensor(rng.standard_normal((200, 64)))
out = model(noise)
loss_g = loss_fn(out, ones)
print('out rg', out.requires_grad, 'loss', float(loss_g.data))
loss_g.backward()
print('live', [p._grad_live for p in model.parameters()], 'rg', [p.requires_grad for p in model.parameters()])
print('gnorm', [float(np.abs(p._grad.numpy()).sum()) for p in model.parameters()])
print('ranges', optim_g.arena.ranges())
optim_g.step()
after = [p.data.numpy() for p in model.parameters()]
print('changed', [bool(np.any(a != b)) for a, b in zip(after, before)])

for eps in (1e-1, 1e-2, 1e-3, 1e-4):
  np.random.seed(3)
  m = MLPCircles()
  r2 = np.random.default_rng(0)
  X, y = ng.tensor(r2.standard_normal((64, 2))), ng.tensor(r2.integers(0, 2, (64, 1)).astype(float))
  print('eps', eps, 'd', ng.grad_check(m, X, y, nn.loss.MSE(), epsilon=eps, print_vals=False))
