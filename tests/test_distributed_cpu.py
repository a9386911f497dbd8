"""World-size-2 gloo test of the data-parallel plumbing (neograd_b200/distributed.py) on CPU: sharding the batch,
summing shard gradients with one allreduce and scaling by 1/world reproduces the full-batch gradient, because every
neograd loss divides by the LOCAL example count (nn/loss.py:22-37).  The arithmetic here is the CPU oracle's (this is a
test of the host-side exchange logic; the CUDA path is covered by the -m gpu tests and bench.py --gpus N)."""
import os
import socket
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
  s = socket.socket()
  s.bind(('127.0.0.1', 0))
  p = s.getsockname()[1]
  s.close()
  return p


def _worker(rank, world, port, q):
  try:
    _worker_body(rank, world, port, q)
  except Exception as e:   # surface failures immediately instead of letting the parent wait for its timeout
    q.put(f'rank {rank}: {type(e).__name__}: {e}')
    raise


def _worker_body(rank, world, port, q):
  sys.path.insert(0, ROOT)
  os.environ.update(RANK=str(rank), WORLD_SIZE=str(world), LOCAL_RANK=str(rank), MASTER_ADDR='127.0.0.1', MASTER_PORT=str(port))
  import torch
  from neograd_b200 import distributed as dist
  from oracle import models
  dist.init('gloo')
  assert dist.world_size() == world and dist.rank() == rank
  rng = np.random.default_rng(0)
  n = 64
  x, y = rng.standard_normal((n, 28, 28)), np.eye(10)[rng.integers(0, 10, n)]
  net = models.lenet_a(np.random.default_rng(1))
  b, e = dist.shard(n)
  assert e - b == n // world
  out = net.forward(x[b:e])
  _, dout = models.softmax_ce(out, y[b:e])
  grads = net.backward(dout)
  flat = torch.from_numpy(np.concatenate([np.ravel(g) for g in grads]))
  dist.all_reduce_sum(flat)                # the one exchange step per iteration
  flat /= world                            # grad_scale = 1/world (folded into the update kernel on the GPU)
  if rank == 0:
    full = models.lenet_a(np.random.default_rng(1))
    _, dfull = models.softmax_ce(full.forward(x), y)
    ref = np.concatenate([np.ravel(g) for g in full.backward(dfull)])
    q.put(float(np.linalg.norm(flat.numpy() - ref) / np.linalg.norm(ref)))
  dist.destroy()


def test_dp_gradient_equals_full_batch_gloo_world2():
  import torch.multiprocessing as mp
  ctx = mp.get_context('spawn')
  q = ctx.Queue()
  port = _free_port()
  procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
  for p in procs:
    p.start()
  err = q.get(timeout=180)
  for p in procs:
    p.join(timeout=60)
    assert p.exitcode == 0
  assert isinstance(err, float), err
  assert err < 1e-12, err


def test_shard_bounds():
  from neograd_b200 import distributed as dist
  assert [dist.shard(4096, r, 8) for r in (0, 7)] == [(0, 512), (3584, 4096)]
  assert dist.shard(10, 2, 3) == (6, 10)
