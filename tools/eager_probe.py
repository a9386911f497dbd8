import sys, time
sys.path.insert(0, '/root/repo')
import torch
import neograd_b200 as ng
from neograd_b200.capture import CapturedStep
from neograd_b200.configs import lenet_synthetic, train_step
model, x, y = lenet_synthetic('b', 4096, seed=100)
X, Y = ng.tensor(x), ng.tensor(y)
optim = ng.nn.optim.Adam(model.parameters(), 1e-3)
loss_fn = ng.nn.loss.SoftmaxCE(axis=1)
step = lambda: train_step(model, optim, loss_fn, X, Y)
for _ in range(3): step()
torch.cuda.synchronize()
cap = CapturedStep(step, warmup=2)
for _ in range(3): cap()
torch.cuda.synchronize()
ts = []
for i in range(12):
  t0 = time.perf_counter(); step(); t1 = time.perf_counter(); torch.cuda.synchronize(); t2 = time.perf_counter()
  ts.append((round((t1-t0)*1e3,2), round((t2-t0)*1e3,2)))
print('eager after capture (issue ms, total ms):', ts)
print(torch.cuda.memory_stats()['num_alloc_retries'], torch.cuda.memory_reserved()>>20)
