"""Float64 NumPy restatement of neograd's hot-path arithmetic (TEST INFRASTRUCTURE, see oracle/__init__.py).

Every function cites the reference file:line (relative to /root/reference) whose result it
reproduces.  The formulation is the textbook one (sliding windows + einsum) plus the reference's
output-ordering identity, not the reference's fragment loops; `tests/test_oracle_golden.py`
pins it to outputs of the unmodified reference (fixtures made by oracle/gen_golden.py).

Layout conventions (all C-contiguous float64):
  Conv3D : x (N,C,H,W), w (O,C,KH,KW), b (O,)  -> y (N,O,Ho,Wo)      [neograd/autograd/ops/conv.py:234-343]
  Conv2D : x (N,H,W),   w (KH,KW),     b ()    -> y (N,Ho,Wo)        [conv.py:121-215]
  MaxPool: x (N,[C,]H,W) -> y (N,[C,]Ho,Wo)                          [conv.py:362-542]
"neograd order" = the reference emits conv/pool outputs with the spatial positions enumerated
W-outer/H-inner (conv.py:39-47) but stored row-major (conv.py:149-151, 264-268, 400-401, 499-500).
"""
import numpy as np
from numpy.lib.stride_tricks import sliding_window_view

F64 = np.float64


# ----------------------------------------------------------------------------- geometry
def out_hw(H, W, KH, KW, pad, stride):
  """conv.py:49-68 -- floor((dim + 2*pad - k)/stride + 1)."""
  assert stride >= 1, 'Stride must be greater than or equal to 1'        # conv.py:33
  assert pad >= 0, 'Padding must be greater than or equal to zero'       # conv.py:34
  ho = (H + 2 * pad - KH) // stride + 1
  wo = (W + 2 * pad - KW) // stride + 1
  return int(ho), int(wo)


def scramble(std):
  """textbook (...,Ho,Wo) -> neograd order.  Position (p,q) lands at flat index q*Ho+p."""
  *lead, ho, wo = std.shape
  return np.ascontiguousarray(np.swapaxes(std, -1, -2)).reshape(*lead, ho, wo)


def unscramble(ng):
  """neograd order (...,Ho,Wo) -> textbook order (inverse of scramble)."""
  *lead, ho, wo = ng.shape
  return np.ascontiguousarray(np.swapaxes(ng.reshape(*lead, wo, ho), -1, -2))


def _pad_hw(x, pad):
  """conv.py:70-85 -- zero padding of the last two dims."""
  if pad == 0:
    return x
  return np.pad(x, [(0, 0)] * (x.ndim - 2) + [(pad, pad), (pad, pad)])


def _windows(xp, KH, KW, stride):
  """(..., Hp, Wp) -> (..., Ho, Wo, KH, KW) view of all visited fragments."""
  win = sliding_window_view(xp, (KH, KW), axis=(-2, -1))
  return win[..., ::stride, ::stride, :, :]


# ----------------------------------------------------------------------------- Conv3D (multi-channel)
def conv3d_fwd(x, w, b, pad=0, stride=1):
  """conv.py:245-269."""
  x, w, b = np.asarray(x, F64), np.asarray(w, F64), np.asarray(b, F64)
  if x.ndim != 4:
    raise ValueError("Only 4D inputs, with 0th dim as number of examples, 1st dim as number of channels are supported!")
  KH, KW = w.shape[-2:]
  out_hw(x.shape[2], x.shape[3], KH, KW, pad, stride)
  win = _windows(_pad_hw(x, pad), KH, KW, stride)            # n c p q i j
  std = np.einsum('ncpqij,ocij->nopq', win, w, optimize=True) + b.reshape(1, -1, 1, 1)
  return scramble(std)


def conv3d_dgrad(ug, w, x_shape, pad=0, stride=1):
  """conv.py:299-309 (inputs_backward) -- result in natural (N,C,H,W) layout."""
  ug, w = np.asarray(ug, F64), np.asarray(w, F64)
  N, C, H, W = x_shape
  O, _, KH, KW = w.shape
  ho, wo = out_hw(H, W, KH, KW, pad, stride)
  g = unscramble(ug.reshape(N, O, ho, wo))
  dxp = np.zeros((N, C, H + 2 * pad, W + 2 * pad), F64)
  for i in range(KH):
    for j in range(KW):
      dxp[:, :, i:i + stride * ho:stride, j:j + stride * wo:stride] += np.einsum('nopq,oc->ncpq', g, w[:, :, i, j])
  return np.ascontiguousarray(dxp[:, :, pad:pad + H, pad:pad + W])


def conv3d_wgrad(ug, x, w_shape, pad=0, stride=1):
  """conv.py:311-321 (kernel_backward)."""
  ug, x = np.asarray(ug, F64), np.asarray(x, F64)
  N, C, H, W = x.shape
  O, _, KH, KW = w_shape
  ho, wo = out_hw(H, W, KH, KW, pad, stride)
  g = unscramble(ug.reshape(N, O, ho, wo))
  win = _windows(_pad_hw(x, pad), KH, KW, stride)
  return np.einsum('nopq,ncpqij->ocij', g, win, optimize=True)


def conv3d_bgrad(ug):
  """conv.py:323-327 -> (O,1,1); tensor.py:99 reshapes to (O,)."""
  return np.asarray(ug, F64).sum(axis=(0, 2, 3))


# ----------------------------------------------------------------------------- Conv2D (single channel)
def conv2d_fwd(x, w, b, pad=0, stride=1):
  """conv.py:131-152 -- single-channel, single-filter, scalar bias."""
  x = np.asarray(x, F64)
  if x.ndim != 3:
    raise ValueError("Only 3D inputs, with 0th dim as number of examples are supported!")
  w = np.asarray(w, F64)
  y = conv3d_fwd(x[:, None], w[None, None], np.asarray(b, F64).reshape(1), pad, stride)
  return y[:, 0]


def conv2d_dgrad(ug, w, x_shape, pad=0, stride=1):
  """conv.py:179-188."""
  N, H, W = x_shape
  w = np.asarray(w, F64)
  return conv3d_dgrad(np.asarray(ug, F64)[:, None], w[None, None], (N, 1, H, W), pad, stride)[:, 0]


def conv2d_wgrad(ug, x, w_shape, pad=0, stride=1):
  """conv.py:190-197."""
  return conv3d_wgrad(np.asarray(ug, F64)[:, None], np.asarray(x, F64)[:, None], (1, 1) + tuple(w_shape), pad, stride)[0, 0]


def conv2d_bgrad(ug):
  """conv.py:198-199 -- np.sum(ug)."""
  return np.asarray(ug, F64).sum()


# ----------------------------------------------------------------------------- MaxPool2D / MaxPool3D
def maxpool_fwd(x, kernel_shape, pad=0, stride=1):
  """conv.py:384-402 / 483-501.  Returns (y, idx): y in neograd order, idx = uint8 position of the
  first maximum inside the (zero-padded) window, row-major in the window (np.argmax, conv.py:424/523),
  stored in the same neograd order as y."""
  x = np.asarray(x, F64)
  if len(kernel_shape) != 2:
    raise ValueError("Only 2D kernels are allowed!")
  KH, KW = kernel_shape
  out_hw(x.shape[-2], x.shape[-1], KH, KW, pad, stride)
  win = _windows(_pad_hw(x, pad), KH, KW, stride)
  flat = win.reshape(*win.shape[:-2], KH * KW)
  idx = np.argmax(flat, axis=-1)
  y = np.take_along_axis(flat, idx[..., None], axis=-1)[..., 0]
  return scramble(y), scramble(idx).astype(np.uint8)


def maxpool_bwd(ug, x, kernel_shape, pad=0, stride=1):
  """conv.py:404-431 / 503-530.  The reference ASSIGNS each window's one-hot block (conv.py:427/526)
  in fragment order (W outer, H inner), so with stride < kernel later windows overwrite earlier ones;
  that is reproduced.  Cells no window covers are garbage (np.empty, conv.py:419/518) in the
  reference and are defined as 0 here (documented deviation; untested in the reference)."""
  ug, x = np.asarray(ug, F64), np.asarray(x, F64)
  KH, KW = kernel_shape
  H, W = x.shape[-2:]
  ho, wo = out_hw(H, W, KH, KW, pad, stride)
  xp = _pad_hw(x, pad)
  g = unscramble(ug.reshape(*x.shape[:-2], ho, wo))
  win = _windows(xp, KH, KW, stride)
  idx = np.argmax(win.reshape(*win.shape[:-2], KH * KW), axis=-1)
  eye = np.eye(KH * KW)
  dxp = np.zeros(xp.shape, F64)
  for q in range(wo):              # fragment order: W outer ...
    for p in range(ho):            # ... H inner (conv.py:39-47)
      block = eye[idx[..., p, q]].reshape(*idx.shape[:-2], KH, KW) * g[..., p, q][..., None, None]
      dxp[..., p * stride:p * stride + KH, q * stride:q * stride + KW] = block
  return np.ascontiguousarray(dxp[..., pad:pad + H, pad:pad + W])


# ----------------------------------------------------------------------------- Dot
def dot_fwd(a, b):
  """basics.py:194."""
  return np.dot(np.asarray(a, F64), np.asarray(b, F64))


def dot_dgrad(ug, b):
  """basics.py:206 -- ug . B^T."""
  return np.dot(np.asarray(ug, F64), np.asarray(b, F64).T)


def dot_wgrad(a, ug):
  """basics.py:207 -- A^T . ug."""
  return np.dot(np.asarray(a, F64).T, np.asarray(ug, F64))


# ----------------------------------------------------------------------------- broadcasting elementwise
BINARY = ('add', 'sub', 'mul', 'div', 'pow')


def binary_fwd(op, a, b):
  """basics.py:20 (add), 63 (sub), 107 (mul), 150 (div), 313 (pow)."""
  a, b = np.asarray(a, F64), np.asarray(b, F64)
  if op == 'add': return a + b
  if op == 'sub': return a - b
  if op == 'mul': return a * b
  if op == 'div': return a / b
  if op == 'pow': return np.power(a, b)
  raise KeyError(op)


def binary_local_grad(op, side, a, b, ug):
  """The grad_fn lambdas: basics.py:32-33, 76-77, 119-120, 163-164, 325-327 (broadcast shape)."""
  a, b, ug = np.asarray(a, F64), np.asarray(b, F64), np.asarray(ug, F64)
  if op == 'add': return ug * np.ones(np.broadcast_shapes(a.shape, b.shape))
  if op == 'sub': return (ug if side == 0 else -ug) * np.ones(np.broadcast_shapes(a.shape, b.shape))
  if op == 'mul': return (b if side == 0 else a) * ug
  if op == 'div': return (1 / b) * ug if side == 0 else ((-1 * a) / np.power(b, 2)) * ug
  if op == 'pow':
    return (np.power(a, b - 1) * b) * ug if side == 0 else (np.power(a, b) * np.log(a)) * ug
  raise KeyError(op)


def unbroadcast_axes(orig_shape, broadcast_shape):
  """utils.py:49-71 -- right-aligned comparison; missing leading dims count as broadcast."""
  nb, no = len(broadcast_shape), len(orig_shape)
  axes = []
  for d in range(nb):
    od = d - (nb - no)
    if od < 0 or orig_shape[od] != broadcast_shape[d]:
      axes.append(d)
  return tuple(axes)


def unbroadcast(data, orig_shape, broadcast_shape):
  """utils.py:34-78 followed by the reshape of tensor.py:99."""
  data = np.asarray(data, F64)
  if broadcast_shape is not None:
    data = np.sum(data, axis=unbroadcast_axes(orig_shape, broadcast_shape))
  return data.reshape(orig_shape)


def binary_bwd(op, side, a, b, ug):
  """local grad + unbroadcast + reshape: what Tensor._backward (tensor.py:93-100) accumulates."""
  a, b = np.asarray(a, F64), np.asarray(b, F64)
  g = binary_local_grad(op, side, a, b, ug)
  return unbroadcast(g, (a.shape, b.shape)[side], np.broadcast_shapes(a.shape, b.shape))


UNARY = ('exp', 'log', 'relu', 'leaky_relu', 'sigmoid', 'tanh')


def unary_fwd(op, x, alpha=0.01):
  """basics.py:236 (exp), 274 (log); activations.py:20 (relu), 54 (sigmoid), 86 (tanh), 205 (leaky)."""
  x = np.asarray(x, F64)
  if op == 'exp': return np.exp(x)
  if op == 'log': return np.log(x)
  if op == 'relu': return np.maximum(0, x)
  if op == 'leaky_relu': return np.where(x >= 0, x, alpha * x)
  if op == 'sigmoid': return 1 / (1 + np.exp(-x))
  if op == 'tanh': return np.tanh(x)
  raise KeyError(op)


def unary_bwd(op, x, ug, alpha=0.01):
  """basics.py:246, 284; activations.py:31 (mask x>=0 -> 1 at 0), 62-63, 94-95, 216."""
  x, ug = np.asarray(x, F64), np.asarray(ug, F64)
  if op == 'exp': return np.exp(x) * ug
  if op == 'log': return (1 / x) * ug
  if op == 'relu': return np.where(x >= 0, 1, 0) * ug
  if op == 'leaky_relu': return np.where(x >= 0, 1, alpha) * ug
  if op == 'sigmoid':
    s = 1 / (1 + np.exp(-x))
    return (s * (1 - s)) * ug
  if op == 'tanh': return (1 - np.power(np.tanh(x), 2)) * ug
  raise KeyError(op)


def sum_fwd(x, axis=None):
  """basics.py:369."""
  return np.sum(np.asarray(x, F64), axis=axis)


def sum_bwd(ug, x_shape, axis=None):
  """basics.py:381-386."""
  ug = np.asarray(ug, F64)
  if axis is not None:
    ug = np.expand_dims(ug, axis=axis)
  return np.ones(x_shape) * ug


# ----------------------------------------------------------------------------- SoftmaxCE
def softmax(x, axis):
  """activations.py:161-174 (max-shifted)."""
  x = np.asarray(x, F64)
  e = np.exp(x - np.max(x, axis=axis, keepdims=True))
  return e / np.sum(e, axis=axis, keepdims=True)


def num_examples(shape):
  """loss.py:22-37."""
  return 1 if len(shape) in (0, 1) else shape[0]


def softmax_ce_fwd(logits, targets, axis, epsilon=1e-9):
  """loss.py:142-157."""
  p = softmax(logits, axis)
  return (-1 / num_examples(np.shape(logits))) * np.sum(np.asarray(targets, F64) * np.log(p + epsilon))


def softmax_ce_bwd(logits, targets, axis, ug=1.0):
  """loss.py:167-170."""
  p = softmax(logits, axis)
  return (np.asarray(ug, F64) / num_examples(np.shape(logits))) * (p - np.asarray(targets, F64))


# ----------------------------------------------------------------------------- optimizers
def adam_step(p, g, m, v, t, lr, beta1=0.9, beta2=0.999, epsilon=1e-8):
  """optim.py:168-179, 191-197.  `t` is the iteration count AFTER the increment of optim.py:169."""
  p, g = np.asarray(p, F64), np.asarray(g, F64)
  m = beta1 * m + (1 - beta1) * g
  v = beta2 * v + (1 - beta2) * np.square(g)
  mhat = m / (1 - beta1 ** t)
  vhat = v / (1 - beta2 ** t)
  return p - lr * (mhat / (np.sqrt(vhat) + epsilon)), m, v


def sgd_step(p, g, lr):
  """optim.py:42-47 (GD)."""
  return np.asarray(p, F64) - lr * np.asarray(g, F64)


def momentum_step(p, g, m, lr, beta=0.9):
  """optim.py:63-70, 86-92."""
  m = beta * m + (1 - beta) * np.asarray(g, F64)
  return np.asarray(p, F64) - lr * m, m


def rmsprop_step(p, g, r, lr, beta=0.9, epsilon=1e-8):
  """optim.py:112-118, 134-140."""
  g = np.asarray(g, F64)
  r = beta * r + (1 - beta) * np.square(g)
  return np.asarray(p, F64) - lr * (g / (np.sqrt(r) + epsilon)), r


# ----------------------------------------------------------------------------- metric
def dist(a, b):
  """The reference's own acceptance metric, utils.py:153: ||a-b|| / (||a|| + ||b||)."""
  a, b = np.asarray(a, F64).ravel(), np.asarray(b, F64).ravel()
  den = np.linalg.norm(a) + np.linalg.norm(b)
  return 0.0 if den == 0 else float(np.linalg.norm(a - b) / den)


def tf32_round(x):
  """Round-to-nearest-even to TF32 (10 explicit mantissa bits); used to calibrate tolerances."""
  b = np.asarray(x, np.float32).view(np.uint32).astype(np.uint64)
  b = (b + 0xFFF + ((b >> 13) & 1)) & 0xFFFFE000
  return b.astype(np.uint32).view(np.float32)
