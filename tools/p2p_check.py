"""2+ GPU check of the peer-memory gradient exchange: trains LeNet-B for a few steps with the NVLink peer exchange and with
the NCCL allreduce (NGB_NO_P2P=1 in a second run), prints a checksum of the parameters and the graph-replayed step time.
  torchrun --nproc-per-node 2 tools/p2p_check.py"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import neograd_b200 as ng
from neograd_b200 import distributed as dist, _backend as B
from neograd_b200.capture import CapturedStep
from neograd_b200.configs import lenet_synthetic, train_step
local = int(os.environ.get('LOCAL_RANK', '0'))
torch.cuda.set_device(local)
dist.init('nccl')
rank, world = dist.rank(), dist.world_size()
model, x, y = lenet_synthetic('b', 1024, seed=100 + rank)
dist.broadcast_params(model.parameters())
X, Y = ng.tensor(x), ng.tensor(y)
optim = ng.nn.optim.Adam(model.parameters(), 1e-3)
loss_fn = ng.nn.loss.SoftmaxCE(axis=1)
mode = 'p2p' if optim.arena.exchange is not None else 'nccl'
for _ in range(3):
  loss = train_step(model, optim, loss_fn, X, Y)
cap = CapturedStep(lambda: train_step(model, optim, loss_fn, X, Y), warmup=2)
for _ in range(5):
  cap()
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
torch.distributed.barrier(); torch.cuda.synchronize()
e0.record()
for _ in range(50):
  cap()
e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / 50
chk = float(sum(np.abs(p.data.numpy()).sum() for p in model.parameters()))
all_chk = [None] * world
torch.distributed.all_gather_object(all_chk, chk)
rc = B.lib.ngb_xgpu_check()
if rank == 0:
  print(f'mode={mode} world={world} ms_per_step={ms:.4f} loss={float(cap.result.data):.6f} checksums={all_chk} identical={len(set(all_chk)) == 1} xgpu_check={rc}')
torch.distributed.barrier()
sys.stdout.flush()
os._exit(0)
