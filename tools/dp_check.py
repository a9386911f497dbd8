"""Data-parallel checks on N >= 2 GPUs (torchrun --nproc-per-node N tools/dp_check.py [--timeout-test]):
  1. the gradient a DP step applies == the full-batch gradient (fp32 mode, 2e-6): every loss divides by the LOCAL example
     count (nn/loss.py:22-37), so the 1/world-scaled sum over equal shards is the full-batch gradient;
  2. replicas are BIT-identical after 3 Adam steps (rank-ordered sum in every rank);
  3. graph-replayed step time with the fused peer-memory exchange + update, and with NCCL allreduce + update (NGB_NO_P2P=1);
  4. --timeout-test: one rank stalls longer than NGB_XGPU_TIMEOUT_S before its step: every rank must RAISE (no update is
     applied, nothing hangs).
Prints one JSON line per check on rank 0."""
import json, os, sys, time, hashlib
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import neograd_b200 as ng
from neograd_b200 import distributed as dist, _backend as B
from neograd_b200.capture import CapturedStep
from neograd_b200.configs import lenet_synthetic, train_step, LeNetB, set_params, synthetic_params, lenet_data

local = int(os.environ.get('LOCAL_RANK', '0'))
torch.cuda.set_device(local)
dist.init('nccl')
rank, world = dist.rank(), dist.world_size()
say = lambda **kw: rank == 0 and print(json.dumps(kw), flush=True)
loss_fn = ng.nn.loss.SoftmaxCE(axis=1)


def fresh_model():
  np.random.seed(0)
  m = LeNetB()
  set_params(m, synthetic_params([p.shape for p in m.parameters()], 1, 0.1))
  return m


if '--timeout-test' in sys.argv:
  model = fresh_model()
  x, y = lenet_data('b', 256, seed=rank)
  X, Y = ng.tensor(x), ng.tensor(y)
  optim = ng.nn.optim.Adam(model.parameters(), 1e-3)
  assert optim.arena.exchange is not None
  train_step(model, optim, loss_fn, X, Y)
  torch.cuda.synchronize()
  before = [p.data.numpy().copy() for p in model.parameters()]
  if rank == world - 1:
    time.sleep(float(os.environ.get('NGB_XGPU_TIMEOUT_S', '60')) + 3.0)      # a stalled rank (eval, checkpoint, GC pause ...)
  t0 = time.time()
  raised = None
  try:
    for _ in range(40):                       # the health word is polled every 16 exchanges
      train_step(model, optim, loss_fn, X, Y)
    torch.cuda.synchronize()
    dist.PeerExchange.check()
  except Exception as e:   # noqa: BLE001
    raised = f'{type(e).__name__}: {e}'
  torch.cuda.synchronize()
  after = [p.data.numpy() for p in model.parameters()]
  res = {'rank': rank, 'raised': raised, 'seconds': round(time.time() - t0, 1)}
  allres = [None] * world
  torch.distributed.all_gather_object(allres, res)
  say(check='timeout', ok=all(r['raised'] is not None for r in allres), results=allres)
  os._exit(0)

# ---- 1. DP gradient == full-batch gradient (fp32 mode)
ng.set_precision('fp32')
G = 64 * world
xf, yf = lenet_data('b', G, seed=5)
model = fresh_model()
loss_fn(model(ng.tensor(xf)), ng.tensor(yf)).backward()
full = [p.grad.numpy().copy() for p in model.parameters()]
lr = 16.0
optim = ng.nn.optim.GD(model.parameters(), lr)
optim.zero_grad()
mode = 'p2p' if optim.arena.exchange is not None else 'nccl'
b, e = dist.shard(G)
before = [p.data.numpy().copy() for p in model.parameters()]
train_step(model, optim, loss_fn, ng.tensor(xf[b:e]), ng.tensor(yf[b:e]))
after = [p.data.numpy() for p in model.parameters()]
errs = []
for p0, p1, g in zip(before, after, full):
  got = (p0 - p1) / lr
  errs.append(float(np.linalg.norm(got - g) / (np.linalg.norm(got) + np.linalg.norm(g) + 1e-30)))
allerr = [None] * world
torch.distributed.all_gather_object(allerr, max(errs))
say(check='dp_gradient_equals_full_batch', mode=mode, world=world, max_dist=max(allerr), ok=max(allerr) < 2e-6, per_param=errs)
ng.set_precision('tf32')

# ---- 2. replicas bit-identical after 3 Adam steps (default precision), 3. step time
model, x, y = lenet_synthetic('b', 1024, seed=100 + rank)
dist.broadcast_params(model.parameters())
X, Y = ng.tensor(x), ng.tensor(y)
optim = ng.nn.optim.Adam(model.parameters(), 1e-3)
for _ in range(3):
  loss = train_step(model, optim, loss_fn, X, Y)
digest = hashlib.sha256(b''.join(p.data.t.detach().cpu().numpy().tobytes() for p in model.parameters())).hexdigest()
alld = [None] * world
torch.distributed.all_gather_object(alld, digest)
say(check='replicas_bit_identical_after_3_adam_steps', mode=mode, ok=len(set(alld)) == 1, digests=[d[:12] for d in alld], iter=optim.iter)

cap = CapturedStep(lambda: train_step(model, optim, loss_fn, X, Y), warmup=2)
for _ in range(10):
  cap()
torch.cuda.synchronize()
torch.distributed.barrier(); torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(200):
  cap()
e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / 200
t = torch.tensor([ms], device='cuda', dtype=torch.float64)
torch.distributed.all_reduce(t, op=torch.distributed.ReduceOp.MAX)
digest = hashlib.sha256(b''.join(p.data.t.detach().cpu().numpy().tobytes() for p in model.parameters())).hexdigest()
torch.distributed.all_gather_object(alld, digest)
rc = B.lib.ngb_xgpu_check()
say(check='captured_step', mode=mode, world=world, batch_per_gpu=1024, ms_per_step=float(t.item()), launches=cap.launches,
    replicas_identical=len(set(alld)) == 1, xgpu_check=rc, loss=float(cap.result.data))
torch.distributed.barrier()
sys.stdout.flush()
os._exit(0)
