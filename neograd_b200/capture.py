"""CUDA-graph capture of a whole training step (new; the reference has no notion of it).

A neograd step is ~100 small kernel launches driven by Python (SURVEY hard part 4): at LeNet sizes the GPU finishes
each kernel long before the interpreter has issued the next one.  `CapturedStep` runs the user's step function a few
times eagerly, records it once into a CUDA graph and then replays it: every kernel, the NCCL gradient allreduce and the
fused optimizer update run back to back with no host involvement.  What makes this possible: all intermediate buffers
come from a stream-ordered pool, parameters / gradients / optimizer state live at fixed addresses in the optimizer
arena, Adam's iteration count lives in device memory, and no op synchronises or touches host memory.

Contract for `fn`: it must not read device values on the host (no float(loss.data) inside) nor create tensors from
host data; inputs are refreshed by copying into the tensors `fn` closes over (`Tensor.data.assign(...)` /
`copy_from_host`).  Whatever `fn` returns (e.g. the loss tensor) stays valid and is updated in place by each replay.
"""
from .device import torch


class CapturedStep:
  def __init__(self, fn, warmup=3):
    t = torch()
    self.fn = fn
    side = t.cuda.Stream()
    side.wait_stream(t.cuda.current_stream())
    with t.cuda.stream(side):
      for _ in range(warmup):
        self.result = fn()
    t.cuda.current_stream().wait_stream(side)
    t.cuda.synchronize()
    from . import _backend as B
    n0 = B.lib.ngb_launch_count()
    self.graph = t.cuda.CUDAGraph()
    # Captured on a HIGH-priority stream: the backward lanes (default priority) carry the weight / bias gradients, and
    # when one of those and the next data-gradient kernel become ready together the chain that everything waits on
    # must get the SMs first (kernel nodes inherit the priority of the stream they were captured on).
    with t.cuda.graph(self.graph, stream=t.cuda.Stream(priority=-1)):
      self.result = fn()
    self.launches = int(B.lib.ngb_launch_count() - n0)   # kernels of libngb200 recorded in the graph
    self._replays = 0

  def __call__(self):
    self.graph.replay()
    self._replays += 1
    if (self._replays & 63) == 0:
      # a captured step runs with no host involvement: poll the library's health words now and then (host-memory reads), so a
      # rank that lost a peer, or a tcgen05 pipeline that gave up, raises instead of replaying wrong steps forever
      from . import _backend as B
      B.check(B.lib.ngb_runtime_status())
    return self.result

  replay = __call__


def copy_from_host(dst_tensor, pinned_host_tensor):
  """Async H2D copy of a pinned torch host tensor into a neograd Tensor's device buffer (same element count)."""
  from .device import flush_readers_of
  flush_readers_of(dst_tensor.data)      # deferred arrays that still read the old batch are produced first
  dst_tensor.data.t.view(-1).copy_(pinned_host_tensor.view(-1), non_blocking=True)
