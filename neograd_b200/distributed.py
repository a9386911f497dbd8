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


class PeerExchange:
  """Gradient exchange over NVLink peer memory (csrc/xgpu.cu) for one flat arena of `numel` floats.

  `grad` is this rank's gradient buffer, allocated in symmetric memory so every GPU of the node can load it; `summed` is
  the local result.  `run(stream)` issues barrier -> reduce (each GPU sums the `world` peer buffers in rank order, so all
  replicas get bit-identical sums) -> barrier.  Construction is collective (every rank must build its exchanges in the
  same order).  `create` returns None when symmetric memory is unavailable (gloo / CPU, single process, NGB_NO_P2P=1):
  the caller then uses the NCCL allreduce."""

  def __init__(self, grad, summed, peer_ptrs, flags, flag_ptrs, epoch, handles):
    self.grad, self.summed = grad, summed
    self._peer_ptrs, self._flags, self._flag_ptrs, self._epoch, self._handles = peer_ptrs, flags, flag_ptrs, epoch, handles
    self.rank, self.world = rank(), world_size()

  @staticmethod
  def create(numel):
    import ctypes
    import torch
    import torch.distributed as td
    # measured on 8 x B200 (tools/p2p_check.py): 3.6 us per step faster than the NCCL allreduce at 8 GPUs, 5 us slower at 2
    # (two operands: NCCL's low-latency protocol wins), so the exchange is used from 4 GPUs up (NGB_P2P=1 / NGB_NO_P2P=1 force)
    force = os.environ.get('NGB_P2P') == '1'
    if not is_initialized() or world_size() <= 1 or os.environ.get('NGB_NO_P2P') == '1' or (world_size() < 4 and not force):
      return None
    if td.get_backend() != 'nccl' or not torch.cuda.is_available():
      return None
    ok = torch.ones(1, device='cuda')
    ex = None
    try:
      import torch.distributed._symmetric_memory as sm
      group = td.group.WORLD
      try:
        sm.enable_symm_mem_for_group(group.group_name)
      except Exception:
        pass
      n = max((int(numel) + 3) & ~3, 4)
      grad = sm.empty(n, dtype=torch.float32, device='cuda')
      grad.zero_()
      hg = sm.rendezvous(grad, group)
      flags = sm.empty(64, dtype=torch.int32, device='cuda')
      flags.zero_()
      hf = sm.rendezvous(flags, group)
      peer_ptrs = torch.tensor([int(p) for p in hg.buffer_ptrs], dtype=torch.int64, device='cuda')
      flag_ptrs = torch.tensor([int(p) for p in hf.buffer_ptrs], dtype=torch.int64, device='cuda')
      ex = PeerExchange(grad, torch.zeros(n, dtype=torch.float32, device='cuda'), peer_ptrs, flags, flag_ptrs,
                        torch.zeros(1, dtype=torch.int32, device='cuda'), (hg, hf))
    except Exception as e:   # noqa: BLE001 -- any failure means "no peer memory here"
      sys_msg = f'neograd_b200: peer-memory gradient exchange unavailable ({type(e).__name__}: {e}); using NCCL'
      if rank() == 0:
        import sys
        print(sys_msg, file=sys.stderr)
      ok.zero_()
    td.all_reduce(ok, op=td.ReduceOp.MIN)    # all ranks or none (also orders the zero-fill of the flags before first use)
    torch.cuda.synchronize()
    td.barrier()
    return ex if ok.item() > 0 else None

  def run(self, n, stream):
    """summed[:n] = sum over ranks of grad[:n]; everything is stream-ordered and graph-capturable."""
    import ctypes
    from . import _backend as B
    fp, pp = ctypes.c_void_p(self._flag_ptrs.data_ptr()), ctypes.c_void_p(self._peer_ptrs.data_ptr())
    ep = ctypes.c_void_p(self._epoch.data_ptr())
    B.check(B.lib.ngb_xgpu_barrier(fp, self.rank, self.world, ep, stream))          # every rank's gradients are final
    B.check(B.lib.ngb_xgpu_reduce_sum(pp, ctypes.c_void_p(self.summed.data_ptr()), int(n), self.world, stream))
    B.check(B.lib.ngb_xgpu_barrier(fp, self.rank, self.world, ep, stream))          # every rank has read every buffer


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
