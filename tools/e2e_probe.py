"""Timeline of the end-to-end pipeline of bench.py (H2D copy stream + step graph): where does a step wait?"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from torch.profiler import ProfilerActivity, profile
import neograd_b200 as ng
from neograd_b200.capture import CapturedStep
from neograd_b200.configs import lenet_synthetic, train_step
model, x, y = lenet_synthetic('b', 4096, seed=100)
X, Y = ng.tensor(x), ng.tensor(y)
optim = ng.nn.optim.Adam(model.parameters(), 1e-3)
loss_fn = ng.nn.loss.SoftmaxCE(axis=1)
cap = CapturedStep(lambda: train_step(model, optim, loss_fn, X, Y), warmup=3)
xh, yh = torch.from_numpy(x).pin_memory(), torch.from_numpy(y).pin_memory()
copy_stream = torch.cuda.Stream()
stage = [(torch.empty_like(X.data.t), torch.empty_like(Y.data.t)) for _ in range(2)]
h2d_done = [torch.cuda.Event() for _ in range(2)]
step_done = [torch.cuda.Event() for _ in range(2)]
loss_host = [torch.empty((), dtype=torch.float32).pin_memory() for _ in range(2)]
main = torch.cuda.current_stream()
def issue(i):
  with torch.cuda.stream(copy_stream):
    copy_stream.wait_event(step_done[i & 1])
    stage[i & 1][0].view(-1).copy_(xh.view(-1), non_blocking=True)
    stage[i & 1][1].view(-1).copy_(yh.view(-1), non_blocking=True)
    h2d_done[i & 1].record(copy_stream)
def run(n):
  for b in range(2): step_done[b].record(main)
  issue(0)
  for i in range(n):
    if i + 1 < n: issue(i + 1)
    main.wait_event(h2d_done[i & 1])
    X.data.t.copy_(stage[i & 1][0], non_blocking=True)
    Y.data.t.copy_(stage[i & 1][1], non_blocking=True)
    loss = cap()
    step_done[i & 1].record(main)
    loss_host[i & 1].copy_(loss.data.t, non_blocking=True)
    if i >= 1: step_done[(i - 1) & 1].synchronize()
  main.synchronize()
for _ in range(10): run(20)
t0 = time.perf_counter(); run(40); print('ms/step', (time.perf_counter() - t0) / 40 * 1e3)
with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
  run(8)
evs = [e for e in prof.events() if e.device_type == torch.autograd.DeviceType.CUDA]
evs.sort(key=lambda e: e.time_range.start)
t0 = evs[0].time_range.start
last_end = {}
for e in evs:
  nm = e.name
  if 'Memcpy' in nm or 'memcpy' in nm or 'conv_gather_mma_kernel<4, 1' in nm or 'adam_kernel' in nm:
    print(f'{e.time_range.start - t0:9.1f} {e.time_range.end - e.time_range.start:8.1f}  {nm[:60]}')
