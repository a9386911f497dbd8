"""Data-parallel plumbing (new: the reference is single-process).

One process per GPU.  Every rank holds a full replica of the parameters and optimizer state and sees its shard of the
global batch; `Optimizer.step()` sums the flat gradient arena over ranks with ONE allreduce (NCCL over NVLink /
NVSwitch on GPUs, gloo in the CPU tests) and applies 1/world inside the update kernel.  Because every loss in
neograd divides by the LOCAL example count (nn/loss.py:22-37), equal shards make the averaged gradient identical to
the full-batch one up to fp32 reassociation.

torch.distributed is used only as the process-group / NCCL communicator plumbing.
"""
import os

_state = {'pg': False}


def init(backend=None):
  """Initialise from the torchrun environment (RANK / WORLD_SIZE / LOCAL_RANK / MASTER_ADDR / MASTER_PORT)."""
  import torch
  import torch.distributed as td
  if td.is_initialized():
    _state['pg'] = True
    return
  if int(os.environ.get('WORLD_SIZE', '1')) <= 1 and backend is None:
    return
  os.environ.setdefault('MASTER_ADDR', '127.0.0.1')
  os.environ.setdefault('MASTER_PORT', '29500')
  if backend is None:
    backend = 'nccl' if torch.cuda.is_available() else 'gloo'
  if backend == 'nccl':
    torch.cuda.set_device(int(os.environ.get('LOCAL_RANK', '0')))
    td.init_process_group(backend, device_id=torch.device('cuda', int(os.environ.get('LOCAL_RANK', '0'))))
  else:
    td.init_process_group(backend)
  _state['pg'] = True


def is_initialized():
  if not _state['pg']:
    return False
  import torch.distributed as td
  return td.is_initialized()


def world_size():
  import torch.distributed as td
  return td.get_world_size() if is_initialized() else 1


def rank():
  import torch.distributed as td
  return td.get_rank() if is_initialized() else 0


def all_reduce_sum(arr):
  """In-place sum over ranks of a flat DeviceArray (or any object with a torch tensor in `.t`, or a torch tensor)."""
  import torch.distributed as td
  import torch
  td.all_reduce(arr if isinstance(arr, torch.Tensor) else arr.t, op=td.ReduceOp.SUM)


def shard(n, r=None, w=None):
  """[begin, end) of rank r's contiguous slice of n samples (equal shards; n must divide evenly for exact parity)."""
  r = rank() if r is None else r
  w = world_size() if w is None else w
  per = n // w
  return r * per, (r + 1) * per if r < w - 1 else n


def broadcast_params(params, src=0):
  """Make every replica start from rank `src`'s parameters."""
  import torch.distributed as td
  if not is_initialized():
    return
  for p in params:
    td.broadcast(p.data.t, src=src)


def destroy():
  import torch.distributed as td
  if is_initialized():
    td.destroy_process_group()
  _state['pg'] = False
