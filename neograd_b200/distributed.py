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
  """Gradient exchange FUSED with the optimizer update over NVLink peer memory (csrc/xgpu.cu) for one flat arena of `numel`
  floats: one kernel per optimizer step pushes this rank's gradients into every peer's receive slots, waits for the peers'
  pushes, sums the `world` slots in rank order (bit-identical replicas), scales by 1/world and applies the update.

  `recv` is this rank's receive buffer [2 parities][world][stride] in symmetric memory (torch.distributed._symmetric_memory:
  plumbing only), `flags` the per-(source, block) arrival words.  Construction is collective (every rank must build its
  exchanges in the same order).  `create` returns None when symmetric memory is unavailable (gloo / CPU, single process,
  NGB_NO_P2P=1): the caller then uses the NCCL allreduce followed by the plain update kernel."""
  MAX_BLOCKS = 148
  LL_MAX_FLOATS = 148 * 256 * 4      # kLLMaxFloats in csrc/xgpu.cu
  KIND = {'adam': 0, 'sgd': 1, 'momentum': 2, 'rmsprop': 3}

  def __init__(self, stride, recv, recv_ptrs, flags, flag_ptrs, epoch, handles):
    self.stride, self.recv, self._recv_ptrs, self._flags, self._flag_ptrs, self._epoch, self._handles = (
      stride, recv, recv_ptrs, flags, flag_ptrs, epoch, handles)
    self.rank, self.world = rank(), world_size()
    self.timeout_s = float(os.environ.get('NGB_XGPU_TIMEOUT_S', '60'))
    self._calls = 0

  @staticmethod
  def create(numel):
    import torch
    import torch.distributed as td
    if not is_initialized() or world_size() <= 1 or os.environ.get('NGB_NO_P2P') == '1':
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
      w = world_size()
      stride = max((int(numel) + 3) & ~3, 4)
      if stride <= PeerExchange.LL_MAX_FLOATS:
        stride *= 2          # room for {value, epoch} words: small arenas use the low-latency protocol (csrc/xgpu.cu)
      recv = sm.empty(2 * w * stride, dtype=torch.float32, device='cuda')
      recv.zero_()
      hr = sm.rendezvous(recv, group)
      flags = sm.empty(w * PeerExchange.MAX_BLOCKS, dtype=torch.int32, device='cuda')
      flags.zero_()
      hf = sm.rendezvous(flags, group)
      recv_ptrs = torch.tensor([int(p) for p in hr.buffer_ptrs], dtype=torch.int64, device='cuda')
      flag_ptrs = torch.tensor([int(p) for p in hf.buffer_ptrs], dtype=torch.int64, device='cuda')
      ex = PeerExchange(stride, recv, recv_ptrs, flags, flag_ptrs, torch.zeros(2, dtype=torch.int32, device='cuda'), (hr, hf))
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

  def update(self, kind, off, n, grad, params, state1, state2, lr, beta1, beta2, epsilon, t_dev, tick, grad_scale, stream):
    """One launch: exchange the range [off, off + n) of `grad` and update params / state in place (stream-ordered and
    graph-capturable; the epoch that selects the receive-slot parity lives in device memory)."""
    import ctypes
    from . import _backend as B
    P = lambda t: ctypes.c_void_p(t.data_ptr()) if t is not None else None
    self._calls += 1
    if (self._calls & 15) == 0:
      self.check()
    B.check(B.lib.ngb_xgpu_exchange_update(P(self._recv_ptrs), P(self._flag_ptrs), self.rank, self.world, self.MAX_BLOCKS,
                                           self.stride, int(off), int(n), P(grad), P(params), P(state1), P(state2),
                                           P(self._epoch), self.KIND[kind], lr, beta1, beta2, epsilon, P(t_dev), int(tick),
                                           grad_scale, self.timeout_s, stream))

  @staticmethod
  def check():
    """Raises if an exchange gave up on a peer (the kernels stop updating from then on); a cheap host-memory read."""
    from . import _backend as B
    B.check(B.lib.ngb_xgpu_status())


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
