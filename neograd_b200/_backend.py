"""ctypes binding of libngb200.so (C ABI declared in include/ngb200.h).

There is no CPU fallback: if the shared library is missing the import of `neograd_b200` fails loudly,
and any kernel call without a CUDA device raises.  PyTorch is used only for device memory, streams and
torch.distributed (plumbing); every arithmetic kernel on the hot path lives in the shared library.
"""
import ctypes
import os
from ctypes import c_double, c_float, c_int, c_int64, c_uint32, c_uint64, c_void_p, c_char_p, POINTER

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, 'lib', 'libngb200.so')

# enums of include/ngb200.h
ADD, SUB, MUL, DIV, POW, SECOND = range(6)
EXP, LOG, RELU, LEAKY_RELU, SIGMOID, TANH = range(6)
GEMM_AUTO, GEMM_SIMT, GEMM_TCGEN05 = range(3)
OK, ERR_INVALID, ERR_CUDA, ERR_UNSUPPORTED, ERR_SCRATCH = range(5)


class NgbError(RuntimeError):
  pass


def _load():
  if not os.path.exists(LIB_PATH):
    raise ImportError(
      f"neograd_b200: {LIB_PATH} is missing. Build it with `make -C neograd_b200/csrc` "
      "(or `python -c 'import __graft_entry__ as g; g.build()'`). There is no CPU fallback.")
  return ctypes.CDLL(LIB_PATH)


_cdll = _load()


class _Lib:
  """Namespace holding one callable per exported function (plain ctypes functions; swapped for timing wrappers by
  `profile_calls`)."""


lib = _Lib()

_i64p = POINTER(c_int64)
_P = c_void_p
_S = c_void_p  # stream

# name -> argtypes, exactly the prototypes of include/ngb200.h (checked by tests/test_abi.py)
SIGNATURES = {
  'ngb_abi_version': ([], c_int),
  'ngb_last_error': ([], c_char_p),
  'ngb_init': ([c_int64], c_int),
  'ngb_shutdown': ([], c_int),
  'ngb_debug_check_pipelines': ([], c_int),
  'ngb_runtime_status': ([], c_int),
  'ngb_debug_wgrad_layout': ([c_int64] * 7 + [_P], c_int),
  'ngb_workspace_reserve': ([c_int64], c_int),
  'ngb_set_default_engine': ([c_int], c_int),
  'ngb_set_lane': ([c_int], c_int),
  'ngb_set_corun': ([c_int], c_int),
  'ngb_prep_u8_affine': ([_P, _P, c_int64, c_float, c_float, _S], c_int),
  'ngb_prep_standardize_rows': ([_P, c_int, _P, c_int64, c_int64, _S], c_int),
  'ngb_prep_one_hot': ([_P, _P, c_int64, c_int64, _S], c_int),
  'ngb_randn': ([_P, c_int64, c_uint64, _P, c_uint32, _S], c_int),
  'ngb_xgpu_exchange_update': ([_P, _P, c_int, c_int, c_int, c_int64, c_int64, c_int64, _P, _P, _P, _P, _P, c_int] +
                               [c_double] * 4 + [_P, c_int, c_double, c_double, _S], c_int),
  'ngb_xgpu_blocks': ([c_int64, c_int], c_int),
  'ngb_xgpu_status': ([], c_int),
  'ngb_xgpu_check': ([], c_int),
  'ngb_xgpu_reset': ([], c_int),
  'ngb_gemm_query': ([c_int, c_int, c_int64, c_int64, c_int64, c_int64, c_int64, c_int64], c_int),
  'ngb_launch_count': ([], c_int64),
  'ngb_gemm': ([c_int, c_int, c_int, c_int64, c_int64, c_int64, _P, c_int64, _P, c_int64, _P, c_int64, c_int, _S], c_int),
  'ngb_linear_fwd': ([_P, _P, _P, _P, _P, c_int, c_float, c_int64, c_int64, c_int64, _S], c_int),
  'ngb_linear_dgrad': ([_P, _P, _P, _P, c_int, c_float, c_int64, c_int64, c_int64, c_int, _S], c_int),
  'ngb_linear_wgrad': ([_P, _P, _P, _P, c_int64, c_int64, c_int64, c_int, c_int, _S], c_int),
  'ngb_linear_softmax_ce': ([_P, _P, _P, _P, _P, _P, _P, c_int64, c_int64, c_int64, c_float, c_float, _S], c_int),
  'ngb_linear_softmax_ce_serves': ([c_int64, c_int64, c_int64], c_int),
  'ngb_conv_fprop': ([_P, _P, _P, _P] + [c_int64] * 9 + [_S], c_int),
  'ngb_conv_dgrad': ([_P, _P, _P] + [c_int64] * 9 + [c_int, _S], c_int),
  'ngb_conv_wgrad': ([_P, _P, _P] + [c_int64] * 9 + [c_int, _S], c_int),
  'ngb_conv_bgrad': ([_P, _P, c_int64, c_int64, c_int64, c_int, _S], c_int),
  'ngb_maxpool_fwd': ([_P, _P, _P] + [c_int64] * 7 + [_S], c_int),
  'ngb_maxpool_bwd': ([_P, _P, _P] + [c_int64] * 7 + [c_int, _S], c_int),
  'ngb_maxpool_relu_fwd': ([_P, _P, _P] + [c_int64] * 7 + [_S], c_int),
  'ngb_maxpool_relu_bwd': ([_P, _P, _P, _P] + [c_int64] * 7 + [c_int, _S], c_int),
  'ngb_maxpool_relu_mask': ([_P, _P, _P, _P] + [c_int64] * 3 + [_S], c_int),
  'ngb_conv_bgrad_pooled': ([_P, _P, _P, _P] + [c_int64] * 4 + [c_int, _S], c_int),
  'ngb_conv_dgrad_pooled': ([_P, _P, _P, _P, _P, _P] + [c_int64] * 9 + [c_int, _S], c_int),
  'ngb_conv_wgrad_pooled': ([_P, _P, _P, _P, _P, _P] + [c_int64] * 9 + [c_int, _S], c_int),
  'ngb_conv_wbgrad_pooled': ([_P, _P, _P, _P, _P, _P, _P] + [c_int64] * 9 + [c_int, c_int, _S], c_int),
  'ngb_conv_wbgrad_pooled_is_fused': ([c_int64] * 9, c_int),
  'ngb_conv_relu_pool_fwd': ([_P, _P, _P, _P, _P] + [c_int64] * 9 + [_S], c_int),
  'ngb_conv_relu_pool_fwd_serves': ([c_int64] * 9, c_int),
  'ngb_conv_bgrad_pooled_idx': ([_P, _P, _P, c_int64, c_int64, c_int64, c_int, _S], c_int),
  'ngb_ew_binary_fwd': ([c_int, _P, c_float, _i64p, c_int, _P, c_float, _i64p, c_int, _P, _S], c_int),
  'ngb_ew_binary_bwd': ([c_int, c_int, _P, _P, c_float, _i64p, c_int, _P, c_float, _i64p, c_int, _P, c_int, _S], c_int),
  'ngb_ew_unary_fwd': ([c_int, c_float, _P, _P, c_int64, _S], c_int),
  'ngb_ew_unary_bwd': ([c_int, c_float, _P, _P, _P, c_int64, c_int, _S], c_int),
  'ngb_reduce_sum': ([_P, _i64p, c_int, c_uint32, _P, c_int, _S], c_int),
  'ngb_fill': ([_P, c_float, c_int64, _S], c_int),
  'ngb_axpy': ([_P, _P, c_float, c_int64, _S], c_int),
  'ngb_transpose2d': ([_P, _P, c_int64, c_int64, _S], c_int),
  'ngb_copy': ([_P, _P, c_int64, _S], c_int),
  'ngb_permute': ([_P, _P, _i64p, POINTER(c_int), c_int, _S], c_int),
  'ngb_softmax_fwd': ([_P, c_int64, c_int64, c_int64, _P, _S], c_int),
  'ngb_softmax_bwd': ([_P, _P, c_int64, c_int64, c_int64, _P, c_int, _S], c_int),
  'ngb_softmax_ce_fwd': ([_P, _P, c_int64, c_int64, c_int64, c_float, c_float, _P, _S], c_int),
  'ngb_softmax_ce_bwd': ([_P, _P, _P, c_int64, c_int64, c_int64, c_float, _P, c_int, _S], c_int),
  'ngb_bce_fwd': ([_P, _P, c_int64, c_float, c_float, _P, _S], c_int),
  'ngb_bce_bwd': ([_P, _P, _P, c_int64, c_float, c_float, _P, c_int, _P, c_int, _S], c_int),
  'ngb_adam_step': ([_P, _P, _P, _P, c_int64] + [c_double] * 4 + [c_int64, c_double, _S], c_int),
  'ngb_adam_step_dev': ([_P, _P, _P, _P, c_int64] + [c_double] * 4 + [_P, c_int, c_double, _S], c_int),
  'ngb_sgd_step': ([_P, _P, c_int64, c_double, c_double, _S], c_int),
  'ngb_momentum_step': ([_P, _P, _P, c_int64, c_double, c_double, c_double, _S], c_int),
  'ngb_rmsprop_step': ([_P, _P, _P, c_int64, c_double, c_double, c_double, c_double, _S], c_int),
}

for _name, (_args, _res) in SIGNATURES.items():
  _fn = getattr(_cdll, _name)   # AttributeError here = the library does not export what the header declares
  _fn.argtypes = _args
  _fn.restype = _res
  setattr(lib, _name, _fn)

if lib.ngb_abi_version() != 1:
  raise ImportError(f"neograd_b200: ABI version mismatch ({lib.ngb_abi_version()} != 1); rebuild libngb200.so")

_EXC = {ERR_INVALID: ValueError, ERR_UNSUPPORTED: NotImplementedError}


def check(rc):
  if rc != OK:
    msg = lib.ngb_last_error().decode('utf-8', 'replace')
    raise _EXC.get(rc, NgbError)(msg)


def shape_arg(shape):
  """tuple -> (int64*, rank)"""
  n = len(shape)
  return (c_int64 * max(n, 1))(*shape), n


_runtime_ready = False


def ensure_runtime():
  """Allocates the library's scratch arena once (must happen before any CUDA-graph capture)."""
  global _runtime_ready
  if not _runtime_ready:
    torch = _t()
    if not torch.cuda.is_available():
      raise NgbError("neograd_b200 needs a CUDA device (B200, sm_100a); there is no CPU fallback")
    check(lib.ngb_init(0))
    _runtime_ready = True


_torch = None
_raw_stream = None
_dev_index = None


def _t():
  global _torch
  if _torch is None:
    import torch as _tm
    _torch = _tm
  return _torch


def stream():
  """cudaStream_t of torch's current stream (so torch.cuda.graph / stream contexts apply to our launches).  One process
  drives one GPU: the device index is looked up once."""
  global _raw_stream, _dev_index
  if _raw_stream is None:
    torch = _t()
    _raw_stream = torch._C._cuda_getCurrentRawStream
    _dev_index = torch.cuda.current_device()
  return _raw_stream(_dev_index)


class _Lanes:
  """Side streams for the leaf gradients of Tensor.backward() (tensor.py:93-100 of the reference).

  The gradient of a graph LEAF (a Param: weight / bias gradient) is consumed by nothing else in the backward pass, so
  its kernels need not sit on the data-gradient chain.  `run` issues them on one of `n` side streams (after the
  upstream gradient is complete on the main stream), `join` makes the main stream wait for them at the end of
  backward().  Under CUDA-graph capture the fork/join becomes parallel branches of the graph.  Each side stream is a
  library lane (ngb_set_lane: private scratch arena).  Buffers the side kernels read are kept referenced until the
  join, so the stream-ordered allocator cannot hand them to a later main-stream allocation."""

  def __init__(self):
    self.n = 3
    self.streams = None
    self.used = set()
    self.keep = []
    self.assign = {}
    self.rr = 0

  def active(self):
    return self.n > 0

  def _init(self):
    torch = _t()
    ensure_runtime()
    self.streams = {}
    for lane in range(1, self.n + 1):
      self.streams[lane] = torch.cuda.Stream()
      prev = lib.ngb_set_lane(lane)          # allocates the lane's arena now (not inside a later capture)
      if prev < 0:
        raise NgbError(lib.ngb_last_error().decode('utf-8', 'replace'))
      lib.ngb_set_lane(prev)

  def mark_ready(self):
    """Event on the current (main) stream: everything issued so far -- in particular the gradient of the tensor the
    engine has just finished -- is complete.  Side-lane work that consumes that gradient waits on THIS event, not on the
    main stream's tail at the time it is issued (leaves are visited last, long after their upstream gradient exists)."""
    torch = _t()
    ev = torch.cuda.Event()
    ev.record()
    return ev

  def can_share(self, a, b):
    """Whether one kernel may write both tensors' gradients: their accumulations must not already sit on different lanes."""
    la, lb = self.assign.get(id(a)), self.assign.get(id(b))
    return la is None or lb is None or la == lb

  def run(self, key, fn, keep, ready=None, corun=True, also=None):
    torch = _t()
    if self.streams is None or len(self.streams) != self.n:
      self._init()
    lane = self.assign.get(id(key))
    if lane is None and also is not None:
      lane = self.assign.get(id(also))
    if lane is None:                         # all accumulations into one tensor stay on one stream (ordered)
      lane = 1 + self.rr % self.n
      self.rr += 1
    self.assign[id(key)] = lane
    if also is not None:
      self.assign[id(also)] = lane
    side = self.streams[lane]
    if ready is not None:
      side.wait_event(ready)
    else:
      side.wait_stream(torch.cuda.current_stream())
    prev = lib.ngb_set_lane(lane)
    prev_corun = lib.ngb_set_corun(1 if corun else 0)
    main = torch.cuda.current_stream()
    try:
      torch.cuda.set_stream(side)            # (the `with torch.cuda.stream(...)` context manager costs several us per use)
      fn()
    finally:
      torch.cuda.set_stream(main)
      lib.ngb_set_lane(prev)
      lib.ngb_set_corun(prev_corun)
    self.used.add(lane)
    self.keep.append(keep)

  def join(self):
    if not self.used:
      return
    torch = _t()
    main = torch.cuda.current_stream()
    for lane in sorted(self.used):
      main.wait_stream(self.streams[lane])
    self.used.clear()
    self.keep.clear()
    self.assign.clear()
    self.rr = 0


lanes = _Lanes()


def set_backward_lanes(n):
  """Number of side streams Tensor.backward() may use for weight / bias gradients (0 = everything on one stream, the
  reference's order of execution; default 3).  Returns the previous setting."""
  n = int(n)
  if not 0 <= n <= 3:
    raise ValueError('set_backward_lanes: n must be 0..3')
  lanes.join()
  prev, lanes.n = lanes.n, n
  return prev


def set_precision(mode):
  """'tf32' (default): Dot / multi-channel conv run on the tcgen05 tensor cores when the shape qualifies;
  'fp32': exact-fp32 CUDA-core kernels everywhere (parity runs, grad_check).  Returns the previous mode."""
  prev = lib.ngb_set_default_engine({'tf32': GEMM_AUTO, 'fp32': GEMM_SIMT}[mode])
  return 'fp32' if prev == GEMM_SIMT else 'tf32'


class profile_calls:
  """Context manager: brackets every kernel-launching C-ABI call with CUDA events on the launching stream.
  `records` is a list of (name, args, start_event, end_event); call `summary()` after the block (it synchronises).
  Used by bench.py for the per-kernel roofline numbers; costs two event records per call, so never leave it on."""
  SKIP = ('ngb_abi_version', 'ngb_last_error', 'ngb_init', 'ngb_shutdown', 'ngb_gemm_query', 'ngb_launch_count',
          'ngb_set_default_engine', 'ngb_workspace_reserve', 'ngb_debug_check_pipelines', 'ngb_runtime_status', 'ngb_set_lane', 'ngb_set_corun', 'ngb_xgpu_check', 'ngb_xgpu_status', 'ngb_xgpu_reset', 'ngb_xgpu_blocks',
          'ngb_debug_wgrad_layout', 'ngb_conv_relu_pool_fwd_serves', 'ngb_conv_wbgrad_pooled_is_fused', 'ngb_linear_softmax_ce_serves')

  def __enter__(self):
    torch = _t()
    self.records = []
    self._saved = {}
    for name in SIGNATURES:
      if name in self.SKIP:
        continue
      fn = getattr(lib, name)
      self._saved[name] = fn

      def wrapped(*args, _fn=fn, _name=name):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        rc = _fn(*args)
        e1.record()
        self.records.append((_name, args, e0, e1))
        return rc
      setattr(lib, name, wrapped)
    return self

  def __exit__(self, *exc):
    for name, fn in self._saved.items():
      setattr(lib, name, fn)

  def summary(self):
    """{name: {'calls': n, 'ms': total, 'items': [(args, ms), ...]}} sorted by total time."""
    torch = _t()
    torch.cuda.synchronize()
    out = {}
    for name, args, e0, e1 in self.records:
      ms = e0.elapsed_time(e1)
      d = out.setdefault(name, {'calls': 0, 'ms': 0.0, 'items': []})
      d['calls'] += 1
      d['ms'] += ms
      d['items'].append((args, ms))
    return dict(sorted(out.items(), key=lambda kv: -kv[1]['ms']))
