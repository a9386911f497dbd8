"""Input pipeline on the device (new; SURVEY 8f row 4 -- the reference's notebooks prepare every batch in NumPy and hand
`ng.tensor` a float64 array: examples/digits.ipynb cells 5-6, examples/gans/basic.ipynb cell 9).

`DeviceLoader` streams mini-batches from HOST arrays in their source format -- uint8 / float pixels, integer class ids --
through pinned staging buffers and a copy stream, double buffered, and expands them in HBM with bandwidth kernels
(csrc/input.cu): affine or per-sample normalisation, one-hot targets.  A LeNet batch of 4096 crosses PCIe as 3.2 MB of
pixels + 16 KB of labels instead of 13 MB of fp32.  Batches come back as ordinary neograd Tensors (device resident), so
training loops read exactly like `get_batches` loops:

    loader = nn.DeviceLoader(images_u8, labels, batch_size=4096, normalize=(1 / 73.9, -127.5 / 73.9), num_classes=10)
    for X, Y in loader:                      # one epoch
      ...

The two device slots are fixed tensors (`loader.slots`), so a step can be captured once per slot (capture.CapturedStep)
and replayed on the slot `loader.next()` reports.  `randn` fills a tensor with device-generated N(0,1) noise."""
import ctypes

import numpy as np

from .. import _backend as B
from ..autograd.tensor import Tensor
from ..device import DeviceArray, torch, _st


def randn(shape, seed=0, out=None, counter=None, offset=0):
  """N(0,1) noise generated on the device (Philox; csrc/input.cu) -- the role of `ng.tensor(np.random.randn(...))`.
  `out`: an existing Tensor to refill in place.  `counter`: a 1-element int32 device tensor that numbers the calls and is
  advanced by the launch itself (fresh noise on every replay of a captured step); otherwise `offset` selects the draw."""
  B.ensure_runtime()
  if out is None:
    out = Tensor(DeviceArray.empty(tuple(shape)))
  arr = out.data
  cptr = ctypes.c_void_p(counter.data_ptr()) if counter is not None else None
  B.check(B.lib.ngb_randn(arr.ptr, arr.size, int(seed), cptr, int(offset), _st()))
  return out


class DeviceLoader:
  """Double-buffered host -> device mini-batch stream with on-device preparation.

  inputs      host array (N, ...), uint8 or float32/float64 (floats are staged as float32)
  targets     host array (N,) of integer class ids (-> one-hot rows when num_classes is given), or a float array (N, ...)
              copied as is, or None
  normalize   None | (scale, shift): x * scale + shift | 'per_sample': (x - mean) / std over each sample (digits.ipynb)
  The last batch of an epoch may be short, as with get_batches (nn/utils.py:39-81)."""

  PIN_LIMIT = 2 << 30     # datasets up to this many bytes are page-locked as a whole: batches then go host -> device straight
                          # from the dataset (no per-batch staging copy on the host)

  def __init__(self, inputs, targets=None, batch_size=None, normalize=None, num_classes=None, drop_last=False,
               prepare_ahead=False):
    # prepare_ahead: the normalisation / one-hot kernels of a batch run on the COPY stream right behind its host -> device
    # copy, i.e. beside the previous batch's compute, instead of on the consumer's stream when the batch is asked for
    self.prepare_ahead = bool(prepare_ahead)
    t = torch()
    B.ensure_runtime()
    inputs = np.asarray(inputs)
    self.n = inputs.shape[0]
    self.batch_size = self.n if batch_size is None else int(batch_size)
    if self.batch_size <= 0 or self.batch_size > self.n:
      raise ValueError("Batch size must be between 1 and the number of examples")
    self.sample_shape = tuple(inputs.shape[1:])
    self.sample_size = int(np.prod(self.sample_shape)) if self.sample_shape else 1
    self.u8 = inputs.dtype == np.uint8
    self._x_host = np.ascontiguousarray(inputs if self.u8 else inputs.astype(np.float32))
    self.normalize = normalize
    self.drop_last = drop_last
    self.num_classes = num_classes
    self._labels = None
    self._y_host = None
    if targets is not None:
      targets = np.asarray(targets)
      assert targets.shape[0] == self.n, '0th dim should be number of examples and should match'
      if num_classes is not None:
        self._labels = np.ascontiguousarray(targets.astype(np.int32).reshape(self.n))
        self.target_shape = (int(num_classes),)
      else:
        self._y_host = np.ascontiguousarray(targets.astype(np.float32))
        self.target_shape = tuple(targets.shape[1:])
    bs = self.batch_size
    xdt = t.uint8 if self.u8 else t.float32
    ybytes = self._labels.nbytes if self._labels is not None else (self._y_host.nbytes if self._y_host is not None else 0)
    self._whole = self._x_host.nbytes + ybytes <= self.PIN_LIMIT
    if self._whole:
      self._x_all = t.from_numpy(self._x_host.reshape(self.n, self.sample_size)).pin_memory()
      ysrc = self._labels if self._labels is not None else (self._y_host.reshape(self.n, -1) if self._y_host is not None else None)
      self._y_all = t.from_numpy(ysrc).pin_memory() if ysrc is not None else None
    self._pin_x = None if self._whole else [t.empty((bs, self.sample_size), dtype=xdt).pin_memory() for _ in range(2)]
    self._raw_x = [t.empty((bs, self.sample_size), dtype=xdt, device='cuda') for _ in range(2)]
    needs_prep = self.u8 or normalize is not None
    self._x = [DeviceArray(t.empty((bs,) + self.sample_shape, dtype=t.float32, device='cuda')) if needs_prep
               else DeviceArray(self._raw_x[i].view((bs,) + self.sample_shape)) for i in range(2)]
    self._needs_prep = needs_prep
    self._pin_y = self._raw_y = self._y = None
    if self._labels is not None:
      self._pin_y = None if self._whole else [t.empty((bs,), dtype=t.int32).pin_memory() for _ in range(2)]
      self._raw_y = [t.empty((bs,), dtype=t.int32, device='cuda') for _ in range(2)]
      self._y = [DeviceArray(t.empty((bs, int(num_classes)), dtype=t.float32, device='cuda')) for _ in range(2)]
    elif self._y_host is not None:
      ysz = int(np.prod(self.target_shape)) if self.target_shape else 1
      self._pin_y = None if self._whole else [t.empty((bs, ysz), dtype=t.float32).pin_memory() for _ in range(2)]
      self._raw_y = [t.empty((bs, ysz), dtype=t.float32, device='cuda') for _ in range(2)]
      self._y = [DeviceArray(self._raw_y[i].view((bs,) + self.target_shape)) for i in range(2)]
    # the two fixed device slots as Tensors (full-batch views; a short last batch is returned as a leading slice)
    self.slots = [(Tensor(self._x[i]), Tensor(self._y[i]) if self._y is not None else None) for i in range(2)]
    self._copy_stream = t.cuda.Stream()
    self._copied = [t.cuda.Event() for _ in range(2)]      # H2D of the slot's raw buffers finished
    self._consumed = [t.cuda.Event() for _ in range(2)]    # the consumer's work on the slot (as of release()) finished
    self._staged = [t.cuda.Event() for _ in range(2)]      # pinned buffer has been read by the copy engine
    self._staged_valid = [False, False]
    self._cursor = 0            # next sample to stage
    self._issued = []           # [(slot, rows)] copies in flight, oldest first
    self._turn = 0
    self.bytes_per_batch = int(bs * self.sample_size * (1 if self.u8 else 4) +
                               (self._raw_y[0].numel() * 4 if self._raw_y is not None else 0))

  # ---------------------------------------------------------------- staging
  def _stage(self):
    """Host gather of the next batch into the pinned buffers of the next slot + async H2D on the copy stream."""
    t = torch()
    start = self._cursor
    if start >= self.n:
      return False
    rows = min(self.batch_size, self.n - start)
    if rows < self.batch_size and self.drop_last:
      self._cursor = self.n
      return False
    slot = self._turn & 1
    self._turn += 1
    if self._whole:
      src_x = self._x_all[start:start + rows]
      src_y = self._y_all[start:start + rows] if self._y_all is not None else None
    else:
      if self._staged_valid[slot]:
        self._staged[slot].synchronize()                    # the copy engine has drained this pinned buffer
      self._pin_x[slot][:rows].copy_(t.from_numpy(self._x_host[start:start + rows].reshape(rows, self.sample_size)))
      if self._labels is not None:
        self._pin_y[slot][:rows].copy_(t.from_numpy(self._labels[start:start + rows]))
      elif self._y_host is not None:
        self._pin_y[slot][:rows].copy_(t.from_numpy(self._y_host[start:start + rows].reshape(rows, -1)))
      src_x = self._pin_x[slot][:rows]
      src_y = self._pin_y[slot][:rows] if self._pin_y is not None else None
    with t.cuda.stream(self._copy_stream):
      self._copy_stream.wait_event(self._consumed[slot])    # whoever used this slot last is done with it
      self._raw_x[slot][:rows].copy_(src_x, non_blocking=True)
      if src_y is not None:
        self._raw_y[slot][:rows].copy_(src_y, non_blocking=True)
      self._staged[slot].record(self._copy_stream)
      if self.prepare_ahead:
        self.prepare(slot, rows)                            # (current stream = the copy stream here)
      self._copied[slot].record(self._copy_stream)
    self._staged_valid[slot] = True
    self._cursor = start + rows
    self._issued.append((slot, rows))
    return True

  def prepare(self, slot, rows=None):
    """Expands the slot's raw batch (as copied from the host) into its fp32 tensors: the normalisation / one-hot kernels,
    on the current stream.  `next()` calls it; a CAPTURED step calls it itself as its first statement (the slot's
    buffers are fixed, so the kernels replay with the graph) and asks `next(prepare=False)` for the batches."""
    rows = self.batch_size if rows is None else rows
    st = _st()
    if self._needs_prep:
      src = ctypes.c_void_p(self._raw_x[slot].data_ptr())
      dst = self._x[slot].ptr
      if self.normalize == 'per_sample':
        B.check(B.lib.ngb_prep_standardize_rows(src, int(self.u8), dst, rows, self.sample_size, st))
      elif self.u8:
        scale, shift = self.normalize if self.normalize is not None else (1.0, 0.0)
        B.check(B.lib.ngb_prep_u8_affine(src, dst, rows * self.sample_size, float(scale), float(shift), st))
      else:
        scale, shift = self.normalize
        from .. import device as D
        tmp = D.binary(B.MUL, DeviceArray(self._raw_x[slot][:rows]), float(scale))
        self._x[slot].reshape((self.batch_size, self.sample_size))[0:rows].assign(D.binary(B.ADD, tmp, float(shift)))
    if self._labels is not None:
      B.check(B.lib.ngb_prep_one_hot(ctypes.c_void_p(self._raw_y[slot].data_ptr()), self._y[slot].ptr, rows, int(self.num_classes), st))

  # ---------------------------------------------------------------- consumer API
  def reset(self):
    """Start a new epoch (pending copies are dropped)."""
    self._cursor, self._issued = 0, []

  def next(self, prepare=True):
    """-> (X, Y, slot) of the next batch, or None at the end of the epoch.  The batch after it is already on its way.
    X / Y stay valid until the batch after the NEXT one is requested; call `release(slot)` once the work that reads the
    slot has been enqueued (iteration does it for you).  prepare=False: only make the current stream wait for the slot's
    host -> device copy (the caller's captured step runs `prepare(slot)` itself)."""
    if not self._issued and not self._stage():
      return None
    slot, rows = self._issued.pop(0)
    self._stage()                                            # prefetch: overlaps this batch's compute
    torch().cuda.current_stream().wait_event(self._copied[slot])
    if prepare and not self.prepare_ahead:
      self.prepare(slot, rows)
    X, Y = self.slots[slot]
    if rows < self.batch_size:
      X, Y = X[0:rows], (Y[0:rows] if Y is not None else None)
    return X, Y, slot

  def release(self, slot):
    """Marks the point on the current stream after which the slot's tensors may be overwritten."""
    self._consumed[slot].record(torch().cuda.current_stream())

  def __iter__(self):
    self.reset()
    prev = None
    while True:
      if prev is not None:
        self.release(prev)           # the loop body of the previous batch has been enqueued by now
      item = self.next()
      if item is None:
        return
      X, Y, prev = item
      yield (X, Y) if Y is not None else X

  def __len__(self):
    full, rem = divmod(self.n, self.batch_size)
    return full + (1 if rem and not self.drop_last else 0)
