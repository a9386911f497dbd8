"""Helpers for the `-m gpu` parity tests: call libngb200.so through its C ABI (include/ngb200.h) on torch CUDA
buffers and hand back NumPy arrays.  torch is used only for device memory and the stream."""
import ctypes

import numpy as np
import torch

from neograd_b200 import _backend as B

lib = B.lib


def dev(a, dtype=torch.float32):
  """host array -> contiguous device tensor"""
  return torch.as_tensor(np.ascontiguousarray(a)).to(device='cuda', dtype=dtype).contiguous()


def host(t):
  torch.cuda.synchronize()
  return t.detach().cpu().numpy()


def ptr(t):
  return ctypes.c_void_p(t.data_ptr()) if t is not None else None


def st():
  return ctypes.c_void_p(B.stream())


def call(name, *args):
  B.ensure_runtime()
  B.check(getattr(lib, name)(*args))


class precision:
  """with precision('fp32'): ... -- exact-fp32 CUDA-core kernels (ngb_set_default_engine, include/ngb200.h); 'tf32' = default."""

  def __init__(self, mode):
    self.engine = {'tf32': B.GEMM_AUTO, 'fp32': B.GEMM_SIMT}[mode]

  def __enter__(self):
    self.prev = lib.ngb_set_default_engine(self.engine)

  def __exit__(self, *exc):
    lib.ngb_set_default_engine(self.prev)


def empty(*shape):
  return torch.empty(shape, device='cuda', dtype=torch.float32)


def f32(a):
  """fp32-representable float64 copy: both sides of every parity test see the same values."""
  return np.asarray(a, np.float32).astype(np.float64)


def gemm(a, b, ta=False, tb=False, engine=B.GEMM_AUTO, accumulate=None):
  """returns op(a) @ op(b) computed by ngb_gemm; a, b are host arrays as stored."""
  A, Bm = dev(a), dev(b)
  M = a.shape[1] if ta else a.shape[0]
  K = a.shape[0] if ta else a.shape[1]
  N = b.shape[0] if tb else b.shape[1]
  if accumulate is not None:
    C = dev(accumulate)
    acc = 1
  else:
    C = torch.full((M, N), float('nan'), device='cuda', dtype=torch.float32)
    acc = 0
  call('ngb_gemm', engine, int(ta), int(tb), M, N, K, ptr(A), a.shape[1], ptr(Bm), b.shape[1], ptr(C), N, acc, st())
  return host(C)


ACT = {None: -1, 'relu': B.RELU, 'leaky_relu': B.LEAKY_RELU, 'sigmoid': B.SIGMOID, 'tanh': B.TANH}


def linear_fwd(x, w, b, act=None, alpha=0.0, want_z=True, want_a=True):
  """(z, a) of ngb_linear_fwd; an output that is not requested comes back as None"""
  M, K = x.shape
  N = w.shape[1]
  X, W, Bv = dev(x), dev(w), (dev(np.reshape(b, -1)) if b is not None else None)
  z = torch.full((M, N), float('nan'), device='cuda') if want_z else None
  a = torch.full((M, N), float('nan'), device='cuda') if want_a else None
  call('ngb_linear_fwd', ptr(X), ptr(W), ptr(Bv), ptr(z), ptr(a), ACT[act], float(alpha), M, K, N, st())
  return (host(z) if want_z else None), (host(a) if want_a else None)


def linear_dgrad(ug, w, zprev=None, act=None, alpha=0.0, into=None):
  M, N = ug.shape
  K = w.shape[0]
  U, W, Z = dev(ug), dev(w), (dev(zprev) if zprev is not None else None)
  dx = torch.full((M, K), float('nan'), device='cuda') if into is None else dev(into)
  call('ngb_linear_dgrad', ptr(U), ptr(W), ptr(Z), ptr(dx), ACT[act], float(alpha), M, K, N, 0 if into is None else 1, st())
  return host(dx)


def linear_wgrad(x, ug, want_db=True, into_w=None, into_b=None):
  M, K = x.shape
  N = ug.shape[1]
  X, U = dev(x), dev(ug)
  dw = torch.full((K, N), float('nan'), device='cuda') if into_w is None else dev(into_w)
  db = None
  if want_db:
    db = torch.full((N,), float('nan'), device='cuda') if into_b is None else dev(np.reshape(into_b, -1))
  call('ngb_linear_wgrad', ptr(X), ptr(U), ptr(dw), ptr(db), M, K, N, 0 if into_w is None else 1, 0 if into_b is None else 1, st())
  return host(dw), (host(db) if want_db else None)


def linear_softmax_ce(x, w, b, t, eps=1e-9):
  """(z, loss, dz) of ngb_linear_softmax_ce"""
  M, K = x.shape
  N = w.shape[1]
  X, W, Bv, T = dev(x), dev(w), dev(np.reshape(b, -1)), dev(t)
  z = torch.full((M, N), float('nan'), device='cuda')
  dz = torch.full((M, N), float('nan'), device='cuda')
  loss = torch.full((1,), float('nan'), device='cuda')
  call('ngb_linear_softmax_ce', ptr(X), ptr(W), ptr(Bv), ptr(T), ptr(z), ptr(loss), ptr(dz), M, K, N, 1.0 / M, eps, st())
  return host(z), float(host(loss)[0]), host(dz)


def bce(o, t, ug=None, eps=1e-9, want=(True, True)):
  """(loss, d_outputs, d_targets) of ngb_bce_fwd / ngb_bce_bwd; num_examples = rows of o (loss.py:22-37)"""
  O, T = dev(o), dev(t)
  n = o.shape[0] if o.ndim >= 2 else 1
  loss = torch.full((1,), float('nan'), device='cuda')
  call('ngb_bce_fwd', ptr(O), ptr(T), O.numel(), 1.0 / n, eps, ptr(loss), st())
  do = torch.full(O.shape, float('nan'), device='cuda') if want[0] else None
  dt = torch.full(O.shape, float('nan'), device='cuda') if want[1] else None
  U = dev(np.reshape(ug, (1,))) if ug is not None else None
  call('ngb_bce_bwd', ptr(O), ptr(T), ptr(U), O.numel(), 1.0 / n, eps, ptr(do), 0, ptr(dt), 0, st())
  return float(host(loss)[0]), (host(do) if want[0] else None), (host(dt) if want[1] else None)


def conv_fprop(x, w, b, pad, stride):
  N, C, H, W = x.shape
  O, _, KH, KW = w.shape
  Ho, Wo = (H + 2 * pad - KH) // stride + 1, (W + 2 * pad - KW) // stride + 1
  y = torch.full((N, O, Ho, Wo), float('nan'), device='cuda')
  X, Wt, Bs = dev(x), dev(w), dev(b)
  call('ngb_conv_fprop', ptr(X), ptr(Wt), ptr(Bs), ptr(y), N, C, H, W, O, KH, KW, pad, stride, st())
  return host(y)


def conv_dgrad(ug, w, xshape, pad, stride):
  N, C, H, W = xshape
  O, _, KH, KW = w.shape
  dx = torch.full((N, C, H, W), float('nan'), device='cuda')
  U, Wt = dev(ug), dev(w)
  call('ngb_conv_dgrad', ptr(U), ptr(Wt), ptr(dx), N, C, H, W, O, KH, KW, pad, stride, 0, st())
  return host(dx)


def conv_wgrad(ug, x, wshape, pad, stride, into=None):
  """`into`: an existing gradient the kernel must accumulate onto (accumulate = 1)"""
  N, C, H, W = x.shape
  O, _, KH, KW = wshape
  dw = torch.full(tuple(wshape), float('nan'), device='cuda') if into is None else dev(into)
  U, X = dev(ug), dev(x)
  call('ngb_conv_wgrad', ptr(U), ptr(X), ptr(dw), N, C, H, W, O, KH, KW, pad, stride, 0 if into is None else 1, st())
  return host(dw)


def conv_bgrad(ug):
  N, O = ug.shape[:2]
  hw = int(np.prod(ug.shape[2:]))
  db = torch.full((O,), float('nan'), device='cuda')
  U = dev(ug)
  call('ngb_conv_bgrad', ptr(U), ptr(db), N, O, hw, 0, st())
  return host(db)


def maxpool_fwd(x, k, pad, stride):
  H, W = x.shape[-2:]
  outer = int(np.prod(x.shape[:-2]))
  Ho, Wo = (H + 2 * pad - k[0]) // stride + 1, (W + 2 * pad - k[1]) // stride + 1
  X = dev(x)
  y = torch.full(x.shape[:-2] + (Ho, Wo), float('nan'), device='cuda')
  idx = torch.full(x.shape[:-2] + (Ho, Wo), 255, device='cuda', dtype=torch.uint8)
  call('ngb_maxpool_fwd', ptr(X), ptr(y), ptr(idx), outer, H, W, k[0], k[1], pad, stride, st())
  return host(y), host(idx)


def maxpool_bwd(ug, idx, xshape, k, pad, stride):
  H, W = xshape[-2:]
  outer = int(np.prod(xshape[:-2]))
  U, I = dev(ug), dev(idx, torch.uint8)
  dx = torch.full(tuple(xshape), float('nan'), device='cuda')
  call('ngb_maxpool_bwd', ptr(U), ptr(I), ptr(dx), outer, H, W, k[0], k[1], pad, stride, 0, st())
  return host(dx)


def maxpool_relu_fwd(x, k, pad, stride):
  """fused maxpool(relu(x)); returns None when the geometry is not supported by the fused kernel"""
  H, W = x.shape[-2:]
  outer = int(np.prod(x.shape[:-2]))
  Ho, Wo = (H + 2 * pad - k[0]) // stride + 1, (W + 2 * pad - k[1]) // stride + 1
  X = dev(x)
  y = torch.full(x.shape[:-2] + (Ho, Wo), float('nan'), device='cuda')
  idx = torch.full(x.shape[:-2] + (Ho, Wo), 255, device='cuda', dtype=torch.uint8)
  B.ensure_runtime()
  rc = lib.ngb_maxpool_relu_fwd(ptr(X), ptr(y), ptr(idx), outer, H, W, k[0], k[1], pad, stride, st())
  if rc == B.ERR_UNSUPPORTED:
    return None
  B.check(rc)
  return host(y), host(idx)


def maxpool_relu_bwd(ug, idx, xm, k, pad, stride):
  H, W = xm.shape[-2:]
  outer = int(np.prod(xm.shape[:-2]))
  U, I, X = dev(ug), dev(idx, torch.uint8), dev(xm)
  dx = torch.full(tuple(xm.shape), float('nan'), device='cuda')
  B.ensure_runtime()
  rc = lib.ngb_maxpool_relu_bwd(ptr(U), ptr(I), ptr(X), ptr(dx), outer, H, W, k[0], k[1], pad, stride, 0, st())
  if rc == B.ERR_UNSUPPORTED:
    return None
  B.check(rc)
  return host(dx)


def maxpool_relu_mask(ug_pool, y_pool, x):
  H, W = x.shape[-2:]
  outer = int(np.prod(x.shape[:-2]))
  U, Y, X = dev(ug_pool), dev(y_pool), dev(x)
  gm = torch.full(tuple(ug_pool.shape), float('nan'), device='cuda')
  call('ngb_maxpool_relu_mask', ptr(U), ptr(Y), ptr(X), ptr(gm), outer, H, W, st())
  return host(gm)


def conv_bgrad_pooled(ug_pool, y_pool, z):
  N, O, Ho, Wo = z.shape
  db = torch.full((O,), float('nan'), device='cuda')
  G, Y, Z = dev(ug_pool), dev(y_pool), dev(z)
  call('ngb_conv_bgrad_pooled', ptr(G), ptr(Y), ptr(Z), ptr(db), N, O, Ho, Wo, 0, st())
  return host(db)


def conv_wbgrad_pooled(ug_pool, y_pool, idx, z, x, wshape, pad, stride, dw0=None, db0=None):
  """(dw, db, fused) from ngb_conv_wbgrad_pooled; dw0 / db0: accumulate into these; None when the geometry is unsupported"""
  N, C, H, W = x.shape
  O, _, KH, KW = wshape
  dw = torch.full(tuple(wshape), float('nan'), device='cuda') if dw0 is None else dev(dw0)
  db = torch.full((O,), float('nan'), device='cuda') if db0 is None else dev(db0)
  G, Y, I, Z, X = dev(ug_pool), dev(y_pool), dev(idx, torch.uint8), dev(z), dev(x)
  B.ensure_runtime()
  fused = lib.ngb_conv_wbgrad_pooled_is_fused(N, C, H, W, O, KH, KW, pad, stride) == 1
  rc = lib.ngb_conv_wbgrad_pooled(ptr(G), ptr(Y), ptr(I), ptr(Z), ptr(X), ptr(dw), ptr(db), N, C, H, W, O, KH, KW, pad, stride,
                                  int(dw0 is not None), int(db0 is not None), st())
  if rc == B.ERR_UNSUPPORTED:
    return None
  B.check(rc)
  return host(dw), host(db), fused


def conv_relu_pool_fwd(x, w, b, pad, stride):
  """(y_pool, idx) of the one-kernel conv + ReLU + MaxPool(2x2, s2) forward; None when the geometry is not served"""
  N, C, H, W = x.shape
  O, _, KH, KW = w.shape
  Ho, Wo = (H + 2 * pad - KH) // stride + 1, (W + 2 * pad - KW) // stride + 1
  X, Wt, Bs = dev(x), dev(w), dev(b)
  y = torch.full((N, O, Ho // 2, Wo // 2), float('nan'), device='cuda')
  idx = torch.full((N, O, Ho // 2, Wo // 2), 255, device='cuda', dtype=torch.uint8)
  B.ensure_runtime()
  rc = lib.ngb_conv_relu_pool_fwd(ptr(X), ptr(Wt), ptr(Bs), ptr(y), ptr(idx), N, C, H, W, O, KH, KW, pad, stride, st())
  if rc == B.ERR_UNSUPPORTED:
    return None
  B.check(rc)
  return host(y), host(idx)


def conv_bgrad_pooled_idx(ug_pool, idx):
  N, O, Hp, Wp = ug_pool.shape
  db = torch.full((O,), float('nan'), device='cuda')
  G, I = dev(ug_pool), dev(idx, torch.uint8)
  call('ngb_conv_bgrad_pooled_idx', ptr(G), ptr(I), ptr(db), N, O, Hp * Wp, 0, st())
  return host(db)


def conv_dgrad_pooled(ug_pool, y_pool, idx, z, w, xshape, pad, stride):
  """returns None when the geometry is not served by the pooled-gradient kernel"""
  N, C, H, W = xshape
  O, _, KH, KW = w.shape
  dx = torch.full((N, C, H, W), float('nan'), device='cuda')
  G, Y, I, Z, Wt = dev(ug_pool), dev(y_pool), dev(idx, torch.uint8), dev(z), dev(w)
  B.ensure_runtime()
  rc = lib.ngb_conv_dgrad_pooled(ptr(G), ptr(Y), ptr(I), ptr(Z), ptr(Wt), ptr(dx), N, C, H, W, O, KH, KW, pad, stride, 0, st())
  if rc == B.ERR_UNSUPPORTED:
    return None
  B.check(rc)
  return host(dx)


def conv_wgrad_pooled(ug_pool, y_pool, idx, z, x, wshape, pad, stride):
  """returns None when the geometry is not served by the pooled-gradient kernel"""
  N, C, H, W = x.shape
  O, _, KH, KW = wshape
  dw = torch.full(tuple(wshape), float('nan'), device='cuda')
  G, Y, I, Z, X = dev(ug_pool), dev(y_pool), dev(idx, torch.uint8), dev(z), dev(x)
  B.ensure_runtime()
  rc = lib.ngb_conv_wgrad_pooled(ptr(G), ptr(Y), ptr(I), ptr(Z), ptr(X), ptr(dw), N, C, H, W, O, KH, KW, pad, stride, 0, st())
  if rc == B.ERR_UNSUPPORTED:
    return None
  B.check(rc)
  return host(dw)


def binary_fwd(op, a, b):
  a, b = np.asarray(a), np.asarray(b)
  out_shape = np.broadcast_shapes(a.shape, b.shape)
  A, Bm = dev(a), dev(b)
  out = torch.full(out_shape, float('nan'), device='cuda')
  ash, ar = B.shape_arg(a.shape)
  bsh, br = B.shape_arg(b.shape)
  call('ngb_ew_binary_fwd', op, ptr(A), 0.0, ash, ar, ptr(Bm), 0.0, bsh, br, ptr(out), st())
  return host(out)


def binary_bwd(op, side, a, b, ug, accumulate=None):
  a, b = np.asarray(a), np.asarray(b)
  A, Bm, U = dev(a), dev(b), dev(ug)
  shape = (a.shape, b.shape)[side]
  if accumulate is not None:
    g = dev(accumulate)
    acc = 1
  else:
    g = torch.full(shape, float('nan'), device='cuda')
    acc = 0
  ash, ar = B.shape_arg(a.shape)
  bsh, br = B.shape_arg(b.shape)
  call('ngb_ew_binary_bwd', op, side, ptr(U), ptr(A), 0.0, ash, ar, ptr(Bm), 0.0, bsh, br, ptr(g), acc, st())
  return host(g)


def unary_fwd(op, x, alpha=0.01):
  X = dev(x)
  out = torch.full(X.shape, float('nan'), device='cuda')
  call('ngb_ew_unary_fwd', op, alpha, ptr(X), ptr(out), X.numel(), st())
  return host(out)


def unary_bwd(op, x, ug, alpha=0.01):
  X, U = dev(x), dev(ug)
  out = torch.full(X.shape, float('nan'), device='cuda')
  call('ngb_ew_unary_bwd', op, alpha, ptr(X), ptr(U), ptr(out), X.numel(), 0, st())
  return host(out)


def reduce_sum(x, axes):
  x = np.asarray(x)
  mask = 0
  for ax in axes:
    mask |= 1 << (ax % x.ndim)
  out_shape = tuple(s for i, s in enumerate(x.shape) if not (mask >> i) & 1)
  X = dev(x)
  out = torch.full(out_shape, float('nan'), device='cuda')
  sh, r = B.shape_arg(x.shape)
  call('ngb_reduce_sum', ptr(X), sh, r, mask, ptr(out), 0, st())
  return host(out)


def softmax_ce(logits, targets, axis):
  sh = logits.shape
  outer = int(np.prod(sh[:axis]))
  L = sh[axis]
  inner = int(np.prod(sh[axis + 1:]))
  n = 1 if len(sh) in (0, 1) else sh[0]
  Z, T = dev(logits), dev(targets)
  loss = torch.full((1,), float('nan'), device='cuda')
  dz = torch.full(sh, float('nan'), device='cuda')
  call('ngb_softmax_ce_fwd', ptr(Z), ptr(T), outer, L, inner, 1.0 / n, 1e-9, ptr(loss), st())
  call('ngb_softmax_ce_bwd', ptr(Z), ptr(T), None, outer, L, inner, 1.0 / n, ptr(dz), 0, st())
  return float(host(loss)[0]), host(dz)
