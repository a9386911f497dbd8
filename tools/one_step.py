"""A few eagerly dispatched training steps of one workload (single stream: every kernel runs alone) -- the short command
ncu profiles:  python tools/one_step.py [lenet_b|gan784] [steps]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import neograd_b200 as ng
import bench

name = sys.argv[1] if len(sys.argv) > 1 else 'lenet_b'
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 3
ng.set_backward_lanes(0)
wl = bench.make_workload(ng, name, bench.BATCH[name], 0)
wl.setup_optim()
step = wl.make_step([ng.tensor(a) for a in wl.host_inputs])
for _ in range(steps):
  out = step()
torch.cuda.synchronize()
print('ok', [float(o.data) for o in out])
