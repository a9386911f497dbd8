import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import neograd_b200 as ng
from neograd_b200 import device as D
from neograd_b200.configs import GAN, set_params
from oracle import models, ops

B = int(sys.argv[1]) if len(sys.argv) > 1 else 8192
params, real, noise_g, noise_d = models.gan784_inputs(B, 1)
gen, disc = models.gan_nets(np.random.default_rng(0))
gen.set_params(params[:6]); disc.set_params(params[6:])
out = disc.forward(gen.forward(noise_g[0]))
lg, dout = models.bce(out, np.ones((B, 1)))
disc.backward(dout)
ogg = gen.backward(disc.input_grad)
for prec in ('fp32', 'tf32'):
  ng.set_precision(prec)
  for fused in (False, True):
    for lanes in (0, 3):
      D.FUSED_LINEAR = fused
      ng.set_backward_lanes(lanes)
      np.random.seed(0)
      model = GAN(784, 512, 256)
      set_params(model, params)
      opt = ng.nn.optim.Adam(model.parameters(), 3e-4)
      model.discriminator.freeze()
      opt.zero_grad(all_members=True)
      loss = ng.nn.loss.BCE()(model(ng.tensor(noise_g[0])), ng.tensor(np.ones((B, 1))))
      loss.backward()
      ds = [ops.dist(p.grad.numpy().ravel(), np.asarray(r).ravel()) for p, r in zip(model.parameters()[:6], ogg)]
      print(prec, 'fused' if fused else 'plain', 'lanes', lanes, 'loss', float(loss.data), lg, ' '.join(f'{d:.2e}' for d in ds), flush=True)
      model.discriminator.unfreeze()
      ng._NG_GRAPH.reset_graph()
