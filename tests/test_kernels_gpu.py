"""GPU parity tests, op level: every entry point of include/ngb200.h against the CPU oracle (oracle/ops.py, pinned
to the reference by tests/test_oracle_golden.py) and against the committed reference fixtures in tests/golden/.

Tolerances (dist = ||a-b||/(||a||+||b||), the reference's own metric, neograd/autograd/utils.py:153):
  * fp32 SIMT kernels (elementwise, unbroadcast, sums, SoftmaxCE, optimizers, direct conv, SIMT GEMM): 2e-6
    (the reference is float64; inputs are fp32-representable, so only fp32 rounding of the arithmetic remains);
  * tcgen05 TF32 kernels (Dot / conv implicit GEMM): 5e-4 (TF32 operand rounding, fp32 accumulate);
  * MaxPool values, argmax indices and pool backward: bit-exact.
"""
import ctypes

import numpy as np
import pytest

from conftest import golden_names, load_golden
from oracle import ops

pytestmark = pytest.mark.gpu

TOL32 = 2e-6
TOLTF32 = 5e-4

BOPS = {'add': 0, 'sub': 1, 'mul': 2, 'div': 3, 'pow': 4}
UOPS = {'exp': 0, 'log': 1, 'relu': 2, 'leaky_relu': 3, 'sigmoid': 4, 'tanh': 5}


@pytest.fixture(scope='module')
def U():
  import torch
  if not torch.cuda.is_available():
    pytest.skip('needs a CUDA device')
  import abi_util
  return abi_util


def close(a, b, tol):
  a, b = np.asarray(a, float), np.asarray(b, float)
  assert a.shape == b.shape, (a.shape, b.shape)
  assert np.all(np.isfinite(a)), 'non-finite values in the device result'
  d = ops.dist(a, b)
  assert d <= tol, d


def f32(a):
  return np.asarray(a, np.float32).astype(np.float64)


# ------------------------------------------------------------------------------------------------- elementwise
@pytest.mark.parametrize('name', golden_names('binary_'))
def test_binary_golden(U, name):
  g = load_golden(name)
  op = str(g['op'])
  a, b, ug = f32(g['a']), f32(g['b']), f32(g['ug'])
  close(U.binary_fwd(BOPS[op], a, b), ops.binary_fwd(op, a, b), TOL32)
  close(U.binary_bwd(BOPS[op], 0, a, b, ug), ops.binary_bwd(op, 0, a, b, ug), 4 * TOL32)
  close(U.binary_bwd(BOPS[op], 1, a, b, ug), ops.binary_bwd(op, 1, a, b, ug), 4 * TOL32)
  # and against the reference's own outputs (inputs differ from fp32 by rounding only)
  close(U.binary_fwd(BOPS[op], a, b), g['out'], 1e-5)


@pytest.mark.parametrize('shape_a,shape_b', [
  ((4096, 784), (1, 784)),     # bias add / bias grad: column reduction
  ((8192, 1), (1,)),           # (N,1) op scalar-like: full reduction
  ((513, 77), (513, 1)),       # row reduction, odd sizes (no 128-bit path)
  ((3, 5, 7, 9), (5, 1, 9)),   # general pattern
  ((6, 1, 4), (1, 5, 1)),      # both sides broadcast
  ((), (33, 3)),               # python-scalar operand
  ((1000003,), (1000003,)),    # flat, odd length (vector body + scalar tail)
])
@pytest.mark.parametrize('op', ['add', 'sub', 'mul', 'div'])
def test_binary_shapes(U, shape_a, shape_b, op):
  rng = np.random.default_rng(sum(shape_a) * 31 + sum(shape_b) * 7 + BOPS[op])
  a = f32(rng.standard_normal(shape_a))
  b = f32(rng.standard_normal(shape_b) + (3.0 if op == 'div' else 0.0))
  ug = f32(rng.standard_normal(np.broadcast_shapes(shape_a, shape_b)))
  close(U.binary_fwd(BOPS[op], a, b), ops.binary_fwd(op, a, b), TOL32)
  for side in (0, 1):
    close(U.binary_bwd(BOPS[op], side, a, b, ug), ops.binary_bwd(op, side, a, b, ug), 4 * TOL32)


def test_binary_bwd_accumulates(U):
  rng = np.random.default_rng(5)
  a, b = f32(rng.standard_normal((300, 40))), f32(rng.standard_normal((1, 40)))
  ug, g0 = f32(rng.standard_normal((300, 40))), f32(rng.standard_normal((1, 40)))
  close(U.binary_bwd(BOPS['mul'], 1, a, b, ug, accumulate=g0), g0 + ops.binary_bwd('mul', 1, a, b, ug), 4 * TOL32)


def test_binary_incompatible_shapes_raise(U):
  with pytest.raises(ValueError):
    U.binary_fwd(0, np.zeros((2, 3)), np.zeros((4, 5)))


@pytest.mark.parametrize('name', golden_names('unary_'))
def test_unary_golden(U, name):
  g = load_golden(name)
  op = str(g['op'])
  x, ug, alpha = f32(g['x']), f32(g['ug']), float(g['alpha'])
  close(U.unary_fwd(UOPS[op], x, alpha), ops.unary_fwd(op, x, alpha), TOL32)
  close(U.unary_bwd(UOPS[op], x, ug, alpha), ops.unary_bwd(op, x, ug, alpha), TOL32)
  close(U.unary_fwd(UOPS[op], x, alpha), g['out'], 1e-5)


@pytest.mark.parametrize('op', ['relu', 'leaky_relu', 'sigmoid', 'exp', 'tanh'])
def test_unary_large_and_relu_at_zero(U, op):
  rng = np.random.default_rng(3)
  x = f32(rng.standard_normal((4096, 257)))
  x[::7, ::3] = 0.0                      # activations.py:31 / 216: the gradient at exactly 0 is 1
  ug = f32(rng.standard_normal(x.shape))
  close(U.unary_fwd(UOPS[op], x), ops.unary_fwd(op, x), TOL32)
  got = U.unary_bwd(UOPS[op], x, ug)
  close(got, ops.unary_bwd(op, x, ug), TOL32)
  if op in ('relu', 'leaky_relu'):
    assert np.array_equal(got[::7, ::3], ug[::7, ::3].astype(np.float32))


@pytest.mark.parametrize('name', golden_names('sum_'))
def test_sum_golden(U, name):
  g = load_golden(name)
  axis = None if int(g['axis']) == -99 else int(g['axis'])
  x = f32(g['x'])
  axes = tuple(range(x.ndim)) if axis is None else (axis,)
  close(U.reduce_sum(x, axes), ops.sum_fwd(x, axis), TOL32)


@pytest.mark.parametrize('shape,axes', [((8192, 10), (0,)), ((8192, 10), (1,)), ((1 << 20,), (0,)), ((33, 65, 17), (0, 2)),
                                        ((256, 16, 8, 8), (0, 2, 3))])
def test_reduce_sum_shapes(U, shape, axes):
  x = f32(np.random.default_rng(9).standard_normal(shape) + 1.0)   # mean 1: no catastrophic cancellation in the reference sum
  close(U.reduce_sum(x, axes), x.sum(axis=axes), 4 * TOL32)


# ------------------------------------------------------------------------------------------------- Dot
@pytest.mark.parametrize('name', golden_names('dot_'))
def test_dot_golden(U, name):
  g = load_golden(name)
  a, b, ug = f32(g['a']), f32(g['b']), f32(g['ug'])
  for eng, tol in ((1, TOL32), (0, TOLTF32)):
    close(U.gemm(a, b, engine=eng), ops.dot_fwd(a, b), tol)
    close(U.gemm(ug, b, tb=True, engine=eng), ops.dot_dgrad(ug, b), tol)
    close(U.gemm(a, ug, ta=True, engine=eng), ops.dot_wgrad(a, ug), tol)


GEMM_SHAPES = [   # (M, K, N): the Linear layers of configs C1-C4 + tile-edge cases
  (750, 2, 100), (750, 100, 1), (200, 36, 10), (4096, 16, 120), (4096, 120, 84), (4096, 84, 10), (4096, 256, 120),
  (8192, 784, 512), (8192, 512, 256), (8192, 256, 784), (8192, 256, 1),
  (128, 32, 64), (129, 33, 65), (1000, 200, 300), (257, 1024, 96),
]


@pytest.mark.parametrize('M,K,N', GEMM_SHAPES)
def test_gemm_simt(U, M, K, N):
  rng = np.random.default_rng(M * 7 + K * 3 + N)
  a, b, ug = f32(rng.standard_normal((M, K))), f32(rng.standard_normal((K, N)) / np.sqrt(K)), f32(rng.standard_normal((M, N)))
  close(U.gemm(a, b, engine=1), a @ b, TOL32)
  close(U.gemm(ug, b, tb=True, engine=1), ug @ b.T, TOL32)
  close(U.gemm(a, ug, ta=True, engine=1), a.T @ ug, TOL32)


@pytest.mark.parametrize('M,K,N', [s for s in GEMM_SHAPES if s[1] % 4 == 0 and s[2] % 4 == 0 and s[1] >= 8])
def test_gemm_tcgen05(U, M, K, N):
  """The three Dot passes on the tensor cores (forced engine), incl. split-K wgrad and ragged tiles."""
  rng = np.random.default_rng(M + K * 5 + N * 11)
  a, b, ug = f32(rng.standard_normal((M, K))), f32(rng.standard_normal((K, N)) / np.sqrt(K)), f32(rng.standard_normal((M, N)))
  close(U.gemm(a, b, engine=2), a @ b, TOLTF32)                 # basics.py:194
  if N % 4 == 0:
    close(U.gemm(ug, b, tb=True, engine=2), ug @ b.T, TOLTF32)  # basics.py:206
  close(U.gemm(a, ug, ta=True, engine=2), a.T @ ug, TOLTF32)    # basics.py:207


def test_gemm_tcgen05_exact_on_tf32_inputs(U):
  """With operands already representable in TF32 and small integer-ish values the tensor-core result is exact."""
  rng = np.random.default_rng(1)
  a = rng.integers(-4, 5, (256, 64)).astype(np.float64)
  b = rng.integers(-4, 5, (64, 128)).astype(np.float64)
  assert np.array_equal(U.gemm(a, b, engine=2), (a @ b).astype(np.float32))
  assert np.array_equal(U.gemm(a, a, ta=True, engine=2), (a.T @ a).astype(np.float32))
  assert np.array_equal(U.gemm(b, b, tb=True, engine=2), (b @ b.T).astype(np.float32))


def test_gemm_accumulate(U):
  rng = np.random.default_rng(2)
  a, b, c = f32(rng.standard_normal((300, 64))), f32(rng.standard_normal((64, 128))), f32(rng.standard_normal((300, 128)))
  close(U.gemm(a, b, engine=1, accumulate=c), c + a @ b, TOL32)
  close(U.gemm(a, b, engine=2, accumulate=c), c + a @ b, TOLTF32)


def test_gemm_linearity_large(U):
  """Size-independent property at a C5-sweep size: (A1+A2).B == A1.B + A2.B within TF32 tolerance."""
  rng = np.random.default_rng(4)
  M, K, N = 65536, 256, 256
  a1, a2 = f32(rng.standard_normal((M, K))), f32(rng.standard_normal((M, K)))
  b = f32(rng.standard_normal((K, N)) / 16)
  close(U.gemm(a1 + a2, b), U.gemm(a1, b).astype(float) + U.gemm(a2, b), 2 * TOLTF32)


# ------------------------------------------------------------------------------------------------- Conv
@pytest.mark.parametrize('name', golden_names('conv3d_'))
def test_conv3d_golden(U, name):
  g = load_golden(name)
  p, s = int(g['pad']), int(g['stride'])
  x, w, b, ug = f32(g['x']), f32(g['w']), f32(g['b']), f32(g['ug'])
  with U.precision('fp32'):
    close(U.conv_fprop(x, w, b, p, s), ops.conv3d_fwd(x, w, b, p, s), TOL32)
    close(U.conv_dgrad(ug, w, x.shape, p, s), ops.conv3d_dgrad(ug, w, x.shape, p, s), TOL32)
    close(U.conv_wgrad(ug, x, w.shape, p, s), ops.conv3d_wgrad(ug, x, w.shape, p, s), TOL32)
    close(U.conv_bgrad(ug), ops.conv3d_bgrad(ug), TOL32)
    close(U.conv_fprop(x, w, b, p, s), g['out'], 1e-5)
    close(U.conv_wgrad(ug, x, w.shape, p, s), g['gw'], 1e-5)
  # default precision: 2 -> 5 channels goes to the TF32 mma.sync kernels
  close(U.conv_fprop(x, w, b, p, s), g['out'], TOLTF32)
  close(U.conv_dgrad(ug, w, x.shape, p, s), ops.conv3d_dgrad(ug, w, x.shape, p, s), TOLTF32)
  close(U.conv_wgrad(ug, x, w.shape, p, s), g['gw'], TOLTF32)


@pytest.mark.parametrize('name', golden_names('conv2d_'))
def test_conv2d_golden(U, name):
  """neograd's single-channel Conv2D (conv.py:121-215) = the C = O = 1 case of the same entry points."""
  g = load_golden(name)
  p, s = int(g['pad']), int(g['stride'])
  x, w, b, ug = f32(g['x']), f32(g['w']), f32(g['b']), f32(g['ug'])
  x4, w4, b1, ug4 = x[:, None], w[None, None], b.reshape(1), ug[:, None]
  close(U.conv_fprop(x4, w4, b1, p, s)[:, 0], ops.conv2d_fwd(x, w, b, p, s), TOL32)
  close(U.conv_dgrad(ug4, w4, x4.shape, p, s)[:, 0], ops.conv2d_dgrad(ug, w, x.shape, p, s), TOL32)
  close(U.conv_wgrad(ug4, x4, w4.shape, p, s)[0, 0], ops.conv2d_wgrad(ug, x, w.shape, p, s), TOL32)
  close(U.conv_bgrad(ug4)[0], ops.conv2d_bgrad(ug), TOL32)
  close(U.conv_fprop(x4, w4, b1, p, s)[:, 0], g['out'], 1e-5)


CONV_CASES = [   # N, C, H, W, O, KH, KW, pad, stride
  (64, 1, 28, 28, 6, 5, 5, 0, 1),      # LeNet-B conv1
  (64, 6, 12, 12, 16, 5, 5, 0, 1),     # LeNet-B conv2
  (32, 1, 28, 28, 1, 5, 5, 0, 1),      # LeNet-A conv1 (neograd Conv2D)
  (200, 1, 8, 8, 1, 3, 3, 0, 1),       # digits
  (3, 2, 11, 14, 3, 2, 3, 2, 2),       # rectangular filter, pad 2, stride 2
  (2, 3, 9, 9, 4, 3, 3, 1, 1),
  (2, 5, 17, 13, 9, 4, 4, 0, 3),       # stride leftovers: unvisited cells get zero gradient (conv.py:40-42, 300)
  (4, 16, 16, 16, 32, 3, 3, 1, 1),     # multi-channel 3x3 p1 (C5 family, small)
]


@pytest.mark.parametrize('precision', ['fp32', 'tf32'])
@pytest.mark.parametrize('N,C,H,W,O,KH,KW,pad,stride', CONV_CASES)
def test_conv_direct(U, N, C, H, W, O, KH, KW, pad, stride, precision):
  """fp32 mode: the CUDA-core kernels (exact fp32).  tf32 mode: whatever the dispatcher picks for the geometry (the
  small-channel mma.sync kernels for LeNet-B, CUDA cores for single-channel layers, tcgen05 for 32-multiples)."""
  rng = np.random.default_rng(N + C * 3 + H * 5 + O * 7 + KH)
  x = f32(rng.standard_normal((N, C, H, W)))
  w = f32(rng.standard_normal((O, C, KH, KW)) / np.sqrt(C * KH * KW))
  b = f32(rng.standard_normal(O))
  y = ops.conv3d_fwd(x, w, b, pad, stride)
  ug = f32(rng.standard_normal(y.shape))
  tol = TOL32 if precision == 'fp32' else TOLTF32
  with U.precision(precision):
    close(U.conv_fprop(x, w, b, pad, stride), y, tol)
    close(U.conv_dgrad(ug, w, x.shape, pad, stride), ops.conv3d_dgrad(ug, w, x.shape, pad, stride), tol)
    close(U.conv_wgrad(ug, x, w.shape, pad, stride), ops.conv3d_wgrad(ug, x, w.shape, pad, stride), 2 * tol)
    close(U.conv_bgrad(ug), ops.conv3d_bgrad(ug), 2 * tol)


SMALL_MMA_CASES = [   # N, C, H, W, O, KH, KW, pad, stride: small-channel layers on the mma.sync TF32 kernels (conv_small_mma.cu)
  (4096, 6, 12, 12, 16, 5, 5, 0, 1),   # LeNet-B conv2 at the bench's batch
  (512, 1, 28, 28, 6, 5, 5, 0, 1),     # LeNet-B conv1
  (37, 3, 11, 14, 5, 3, 3, 1, 1),      # ragged everything: batch not a multiple of the images per pass, F % 16 != 0, padding
  (9, 5, 13, 9, 12, 2, 3, 1, 2),       # rectangular filter, stride 2 (fprop / wgrad on tensor cores, dgrad on CUDA cores)
  (5, 8, 10, 10, 8, 3, 3, 2, 1),       # "full" padding = K-1: no halo in the data gradient
  (3, 4, 16, 16, 32, 3, 3, 1, 1),      # O = 32: four channel tiles
  (130, 16, 8, 8, 4, 3, 3, 0, 1),      # C = 16: two channel tiles in the data gradient, 9 filter m-tiles in wgrad
  (7, 4, 9, 11, 12, 5, 5, 1, 1),       # odd W (32-bit shift-add in the row-decomposed dgrad), padding clips rows
  (10, 7, 8, 8, 20, 3, 3, 0, 1),       # odd C (half-empty channel pair), O = 20 -> three k-blocks of output channels
]


@pytest.mark.parametrize('N,C,H,W,O,KH,KW,pad,stride', SMALL_MMA_CASES)
def test_conv_small_mma(U, N, C, H, W, O, KH, KW, pad, stride):
  rng = np.random.default_rng(N * 5 + C * 3 + H + O * 7 + KH)
  x = f32(rng.standard_normal((N, C, H, W)))
  w = f32(rng.standard_normal((O, C, KH, KW)) / np.sqrt(C * KH * KW))
  b = f32(rng.standard_normal(O))
  y = ops.conv3d_fwd(x, w, b, pad, stride)
  ug = f32(rng.standard_normal(y.shape))
  close(U.conv_fprop(x, w, b, pad, stride), y, TOLTF32)
  close(U.conv_dgrad(ug, w, x.shape, pad, stride), ops.conv3d_dgrad(ug, w, x.shape, pad, stride), TOLTF32)
  close(U.conv_wgrad(ug, x, w.shape, pad, stride), ops.conv3d_wgrad(ug, x, w.shape, pad, stride), TOLTF32)


@pytest.mark.parametrize('N,C,H,W,O,KH,KW,pad,stride', SMALL_MMA_CASES[2:])
def test_conv_small_mma_exact_on_small_integers(U, N, C, H, W, O, KH, KW, pad, stride):
  """TF32-representable inputs -> the mma.sync kernels must be bit-exact (any fragment / offset-table mix-up shows), incl.
  accumulation into an existing gradient."""
  rng = np.random.default_rng(7 + N)
  x = rng.integers(-3, 4, (N, C, H, W)).astype(np.float64)
  w = rng.integers(-2, 3, (O, C, KH, KW)).astype(np.float64)
  b = rng.integers(-5, 6, O).astype(np.float64)
  y = ops.conv3d_fwd(x, w, b, pad, stride)
  ug = rng.integers(-3, 4, y.shape).astype(np.float64)
  assert np.array_equal(U.conv_fprop(x, w, b, pad, stride), y.astype(np.float32))
  assert np.array_equal(U.conv_dgrad(ug, w, x.shape, pad, stride), ops.conv3d_dgrad(ug, w, x.shape, pad, stride).astype(np.float32))
  assert np.array_equal(U.conv_wgrad(ug, x, w.shape, pad, stride), ops.conv3d_wgrad(ug, x, w.shape, pad, stride).astype(np.float32))


IGEMM_CASES = [   # N, C, H, W, O, KH, KW, pad  (stride 1, C % 32 == O % 32 == 0: tcgen05 implicit-GEMM path)
  (16, 32, 16, 16, 64, 3, 3, 1),       # smallest eligible
  (4, 64, 20, 24, 32, 3, 3, 1),        # non-square, partial tiles in both directions
  (4, 32, 18, 18, 32, 3, 3, 0),        # no padding
  (2, 32, 28, 32, 96, 5, 5, 2),        # 5x5, O not a power of two (ragged channel tile)
  (2, 128, 32, 32, 128, 3, 3, 1),      # C5 family
  (1, 64, 64, 64, 256, 3, 3, 1),       # BN = 256
  (3, 32, 16, 20, 32, 3, 2, 1),        # rectangular filter
  (256, 64, 32, 32, 64, 3, 3, 1),      # a full BASELINE C5 sweep point (N = 256, C = O = 64, H = W = 32)
]


@pytest.mark.parametrize('N,C,H,W,O,KH,KW,pad', IGEMM_CASES)
def test_conv_igemm_tcgen05(U, N, C, H, W, O, KH, KW, pad):
  """Conv3D forward and data gradient on the tensor cores vs the float64 oracle (conv.py:245-269, 299-309)."""
  rng = np.random.default_rng(N * 3 + C + H * 7 + O * 11 + KH)
  x = f32(rng.standard_normal((N, C, H, W)))
  w = f32(rng.standard_normal((O, C, KH, KW)) / np.sqrt(C * KH * KW))
  b = f32(rng.standard_normal(O))
  y = ops.conv3d_fwd(x, w, b, pad, 1)
  ug = f32(rng.standard_normal(y.shape))
  close(U.conv_fprop(x, w, b, pad, 1), y, TOLTF32)
  close(U.conv_dgrad(ug, w, x.shape, pad, 1), ops.conv3d_dgrad(ug, w, x.shape, pad, 1), TOLTF32)
  close(U.conv_wgrad(ug, x, w.shape, pad, 1), ops.conv3d_wgrad(ug, x, w.shape, pad, 1), TOLTF32)


def test_conv_igemm_exact_on_small_integers(U):
  """TF32-representable inputs -> the tensor-core convolution must be exact (catches any tap / layout mix-up)."""
  rng = np.random.default_rng(12)
  x = rng.integers(-3, 4, (16, 32, 16, 16)).astype(np.float64)
  w = rng.integers(-2, 3, (32, 32, 3, 3)).astype(np.float64)
  b = rng.integers(-5, 6, 32).astype(np.float64)
  y = ops.conv3d_fwd(x, w, b, 1, 1)
  assert np.array_equal(U.conv_fprop(x, w, b, 1, 1), y.astype(np.float32))
  ug = rng.integers(-3, 4, y.shape).astype(np.float64)
  assert np.array_equal(U.conv_dgrad(ug, w, x.shape, 1, 1), ops.conv3d_dgrad(ug, w, x.shape, 1, 1).astype(np.float32))
  # weight gradient (conv_wgrad_igemm_kernel, split over positions + deterministic reduce): |sum| <= 9 * 4096 < 2^24, so
  # any summation order is exact in fp32
  assert np.array_equal(U.conv_wgrad(ug, x, w.shape, 1, 1), ops.conv3d_wgrad(ug, x, w.shape, 1, 1).astype(np.float32))
  # ... and accumulating into an existing gradient buffer
  dw0 = rng.integers(-4, 5, w.shape).astype(np.float64)
  assert np.array_equal(U.conv_wgrad(ug, x, w.shape, 1, 1, into=dw0),
                        (dw0 + ops.conv3d_wgrad(ug, x, w.shape, 1, 1)).astype(np.float32))


# ---------------------------------------------------------------------------------------------- Linear (fused) entry points
LINEAR_CASES = [   # batch, n_in, n_out : which engine serves it in 'tf32' mode
  (4096, 256, 120),    # LeNet FC1: tcgen05, ragged N tile (120 = 3 x 32 + 24), TMA-store epilogue
  (4096, 120, 84),     # LeNet FC2: tcgen05, K = 120 (partial k-block), N = 84
  (4096, 84, 10),      # LeNet FC3: small-N kernels
  (8192, 784, 512),    # GAN: tcgen05 BN = 256
  (1024, 256, 784),    # GAN last generator layer
  (2048, 256, 1),      # GAN discriminator head: small-N, one output
  (8192, 256, 1),      # ... at the BASELINE batch: the slab-parallel weight gradient (x >= 4 MB)
  (16384, 84, 10),     # a classifier head large enough for the slab kernel (three rows in flight per block)
  (200, 36, 10),       # digits CNN head
  (750, 2, 100),       # README MLP first layer: SIMT (K = 2)
  (333, 50, 25),       # notebook-scale GAN layer: SIMT (unaligned leading dimensions)
  (65, 33, 32),        # small-N with the widest tile
]


def _act_ref(name, z, alpha):
  return z if name is None else ops.unary_fwd(name, z, alpha) if name == 'leaky_relu' else ops.unary_fwd(name, z)


def _act_grad_ref(name, z, alpha):
  one = np.ones_like(z)
  return one if name is None else (ops.unary_bwd(name, z, one, alpha) if name == 'leaky_relu' else ops.unary_bwd(name, z, one))


@pytest.mark.parametrize('precision', ['tf32', 'fp32'])
@pytest.mark.parametrize('M,K,N', LINEAR_CASES)
def test_linear_fused_vs_oracle(U, M, K, N, precision):
  """ngb_linear_fwd / _dgrad / _wgrad against the float64 oracle of the two-node form (misc.py:63: dot + broadcast add;
  basics.py:206-207; utils.py:75), with every activation the epilogues implement."""
  rng = np.random.default_rng(M + 3 * K + 7 * N)
  x, w = f32(rng.standard_normal((M, K))), f32(rng.standard_normal((K, N)) / np.sqrt(K))
  b, ug = f32(rng.standard_normal((1, N))), f32(rng.standard_normal((M, N)))
  zprev = f32(rng.standard_normal((M, K)))
  zprev[0, : min(3, K)] = 0.0                      # exact zeros: relu' = 1 at 0 (activations.py:31)
  tc = precision == 'tf32' and U.lib.ngb_gemm_query(0, 0, M, N, K, K, N, N) == 2 and N > 32
  tol = TOLTF32 if tc else TOL32
  zr = ops.dot_fwd(x, w) + b
  with U.precision(precision):
    for act, alpha in ((None, 0.0), ('relu', 0.0), ('leaky_relu', 0.01), ('sigmoid', 0.0), ('tanh', 0.0)):
      z, a = U.linear_fwd(x, w, b, act, alpha)
      close(z, zr, tol)
      close(a, _act_ref(act, zr, alpha), tol)
      dx = U.linear_dgrad(ug, w, zprev, act, alpha)
      close(dx, ops.dot_dgrad(ug, w) * _act_grad_ref(act, zprev, alpha), tol)
    z_only, none = U.linear_fwd(x, w, b, 'relu', want_a=False)
    assert none is None
    close(z_only, zr, tol)
    none, a_only = U.linear_fwd(x, w, None, 'relu', want_z=False)
    assert none is None
    close(a_only, np.maximum(ops.dot_fwd(x, w), 0), tol)
    close(U.linear_dgrad(ug, w), ops.dot_dgrad(ug, w), tol)
    dx0 = f32(rng.standard_normal((M, K)))
    close(U.linear_dgrad(ug, w, zprev, 'relu', into=dx0), dx0 + ops.dot_dgrad(ug, w) * (zprev >= 0), tol)
    dw, db = U.linear_wgrad(x, ug)
    close(dw, ops.dot_wgrad(x, ug), tol)
    close(db, ug.sum(axis=0), tol)
    dw2, none = U.linear_wgrad(x, ug, want_db=False)
    assert none is None
    close(dw2, ops.dot_wgrad(x, ug), tol)
    dw0, db0 = f32(rng.standard_normal((K, N))), f32(rng.standard_normal(N))
    dw3, db3 = U.linear_wgrad(x, ug, into_w=dw0, into_b=db0)
    close(dw3, dw0 + ops.dot_wgrad(x, ug), tol)
    close(db3, db0 + ug.sum(axis=0), tol)


@pytest.mark.parametrize('M,K,N', [(4096, 256, 120), (512, 128, 256), (300, 64, 96)])
def test_linear_fused_matches_unfused_kernels_bitwise(U, M, K, N):
  """Same accumulators, same operations in the same order: the fused forward equals ngb_gemm followed by the bias add and
  the activation kernel BIT FOR BIT; the masked data gradient equals ngb_gemm followed by the activation's backward kernel;
  the weight gradient equals ngb_gemm's."""
  rng = np.random.default_rng(M + K + N)
  x, w = f32(rng.standard_normal((M, K))), f32(rng.standard_normal((K, N)) / np.sqrt(K))
  b, ug = f32(rng.standard_normal((1, N))), f32(rng.standard_normal((M, N)))
  zprev = f32(rng.standard_normal((M, K)))
  z, a = U.linear_fwd(x, w, b, 'relu')
  z2 = U.binary_fwd(BOPS['add'], U.gemm(x, w), b)
  assert np.array_equal(z, z2)
  assert np.array_equal(a, U.unary_fwd(UOPS['relu'], z2))
  dx = U.linear_dgrad(ug, w, zprev, 'leaky_relu', 0.01)
  assert np.array_equal(dx, U.unary_bwd(UOPS['leaky_relu'], zprev, U.gemm(ug, w, tb=True), 0.01))
  dw, _ = U.linear_wgrad(x, ug)
  assert np.array_equal(dw, U.gemm(x, ug, ta=True))


def test_linear_tcgen05_exact_on_small_integers(U):
  """TF32-representable integers: products and sums are exact, so the tensor-core Linear kernels -- including the
  ones-tile MMA that yields the bias gradient and its split-K partials -- must equal the oracle exactly."""
  rng = np.random.default_rng(5)
  for M, K, N in ((4096, 256, 128), (8192, 784, 512), (640, 120, 84)):
    x = rng.integers(-3, 4, (M, K)).astype(np.float64)
    w = rng.integers(-2, 3, (K, N)).astype(np.float64)
    b = rng.integers(-5, 6, (1, N)).astype(np.float64)
    ug = rng.integers(-2, 3, (M, N)).astype(np.float64)
    z, a = U.linear_fwd(x, w, b, 'relu')
    zr = x @ w + b
    assert np.array_equal(z, zr.astype(np.float32)) and np.array_equal(a, np.maximum(zr, 0).astype(np.float32))
    assert np.array_equal(U.linear_dgrad(ug, w, zr @ w.T, 'relu'), ((ug @ w.T) * ((zr @ w.T) >= 0)).astype(np.float32))
    dw, db = U.linear_wgrad(x, ug)
    assert np.array_equal(dw, (x.T @ ug).astype(np.float32))
    assert np.array_equal(db, ug.sum(axis=0).astype(np.float32))


@pytest.mark.parametrize('M,K,N', [(4096, 84, 10), (200, 36, 10), (77, 16, 3), (512, 120, 32)])
def test_linear_softmax_ce_head(U, M, K, N):
  """Classifier head in one kernel: logits, SoftmaxCE loss (loss.py:153-157) and its gradient for an upstream 1 (loss.py:170)."""
  rng = np.random.default_rng(M + N)
  x, w, b = f32(rng.standard_normal((M, K))), f32(rng.standard_normal((K, N)) / np.sqrt(K)), f32(rng.standard_normal((1, N)))
  t = np.eye(N)[rng.integers(0, N, M)]
  z, loss, dz = U.linear_softmax_ce(x, w, b, t)
  zr = ops.dot_fwd(x, w) + b
  close(z, zr, TOL32)
  lr = ops.softmax_ce_fwd(zr, t, 1)
  assert abs(loss - lr) <= 2e-6 * abs(lr), (loss, lr)
  close(dz, ops.softmax_ce_bwd(zr, t, 1), 5e-6)
  # run to run bit-identical (fixed-order loss reduction), and the blocks-done counter is left at zero
  z2, loss2, dz2 = U.linear_softmax_ce(x, w, b, t)
  assert loss2 == loss and np.array_equal(dz2, dz)


@pytest.mark.parametrize('shape', [(8192, 1), (750, 1), (33, 7), (1000003, 1)])
def test_bce_single_pass(U, shape):
  """ngb_bce_fwd / _bwd against the reference's composite expression (loss.py:83-84) evaluated in float64, with the
  reference's argument roles (outputs multiply the logs of the targets) and gradients for both operands."""
  rng = np.random.default_rng(shape[0])
  o = f32(rng.uniform(0.02, 0.98, shape))
  t = f32(rng.uniform(0.05, 0.95, shape))
  eps, n = 1e-9, shape[0]
  want = (-1 / n) * np.sum(o * np.log(t + eps) + (1 - o) * np.log(1 - t + eps))
  loss, do, dt = U.bce(o, t, ug=0.5)
  assert abs(loss - want) <= 2e-6 * abs(want), (loss, want)
  close(do, 0.5 * (-1 / n) * (np.log(t + eps) - np.log(1 - t + eps)), 4 * TOL32)
  close(dt, 0.5 * (-1 / n) * (o / (t + eps) - (1 - o) / (1 - t + eps)), 4 * TOL32)
  # hard targets (the GAN's ones / zeros): log(0 + eps) is finite, run-to-run bit-identical
  ones = np.ones(shape)
  l1, d1, _ = U.bce(o, ones, want=(True, False))
  l2, d2, _ = U.bce(o, ones, want=(True, False))
  assert l1 == l2 and np.array_equal(d1, d2) and np.isfinite(l1)
  assert abs(l1 - (-1 / n) * np.sum(o * np.log(1 + eps) + (1 - o) * np.log(eps))) <= 1e-5 * abs(l1)


def test_no_pipeline_timeouts(U):
  """Runs last in this file's conv / gemm group: no tcgen05 warp gave up on a barrier during the tests above."""
  U.call('ngb_debug_check_pipelines')


@pytest.mark.parametrize('shape', [(64, 6, 24, 24), (33, 16, 8, 8), (5, 12, 20), (3, 2, 10, 36)])
def test_maxpool_relu_fused_is_bit_exact(U, shape):
  """ngb_maxpool_relu_fwd / _bwd (ReLU applied on load / ReLU mask applied on store) against the separate kernels and
  the oracle: values, argmax indices (ties between clamped zeros included) and input gradients are bit-identical."""
  rng = np.random.default_rng(sum(shape))
  x = f32(rng.standard_normal(shape))
  x[0] = -np.abs(x[0])                      # whole windows of negatives: relu makes them all-zero ties
  x[..., ::3, ::2] = 0.0                    # exact zeros: relu gradient is 1 there (activations.py:31)
  got = U.maxpool_relu_fwd(x, (2, 2), 0, 2)
  assert got is not None
  y, idx = got
  yr, idxr = U.maxpool_fwd(np.maximum(x, 0.0), (2, 2), 0, 2)
  assert np.array_equal(y, yr) and np.array_equal(idx & 3, idxr)
  # bit 2 of the fused kernel's index bytes: relu'(x at the argmax) = (x_argmax >= 0); windows are stored (p', q') at q'*Hp + p'
  H2, W2 = x.shape[-2] // 2, x.shape[-1] // 2
  win = x.reshape(x.shape[:-2] + (H2, 2, W2, 2)).swapaxes(-3, -2).reshape(x.shape[:-2] + (H2, W2, 4))
  arg = (idx & 3).reshape(x.shape[:-2] + (W2, H2)).swapaxes(-1, -2).astype(int)
  x_arg = np.take_along_axis(win, arg[..., None], axis=-1)[..., 0]
  flag = (idx >> 2).reshape(x.shape[:-2] + (W2, H2)).swapaxes(-1, -2)
  assert np.array_equal(flag, (x_arg >= 0).astype(flag.dtype))
  yo = ops.maxpool_fwd(np.maximum(x, 0.0), (2, 2), 0, 2)
  yo = yo[0] if isinstance(yo, tuple) else yo
  assert np.array_equal(y, yo.astype(np.float32))
  ug = f32(rng.standard_normal(y.shape))
  dx = U.maxpool_relu_bwd(ug, idx, x, (2, 2), 0, 2)
  assert dx is not None
  ref = U.maxpool_bwd(ug, idx, x.shape, (2, 2), 0, 2) * (x >= 0)
  assert np.array_equal(dx, ref.astype(np.float32))


@pytest.mark.parametrize('N,C,H,W,O,K', [(64, 1, 28, 28, 6, 5), (33, 6, 12, 12, 16, 5), (5, 3, 10, 14, 8, 3)])
def test_conv_wgrad_from_pooled_gradient(U, N, C, H, W, O, K):
  """Conv -> ReLU -> MaxPool(2x2, s2) backward without the full-resolution gradient: ngb_maxpool_relu_mask keeps the
  conv-output gradient in pooled form, ngb_conv_wgrad_pooled scatters it inside the kernel.  Against the separate kernels:
  the scattered compact gradient is bit-identical to ngb_maxpool_relu_bwd's, the weight gradient is bit-identical to
  ngb_conv_wgrad on that full gradient (same operands, same summation order), the bias gradient agrees to fp32 rounding."""
  rng = np.random.default_rng(N + C + H + O)
  xin = f32(rng.standard_normal((N, C, H, W)))
  Ho, Wo = H - K + 1, W - K + 1
  z = f32(rng.standard_normal((N, O, Ho, Wo)))             # the conv output = ReLU input
  z[0] = -np.abs(z[0]) - 0.5                              # fully clamped windows
  z[1, :, ::2, ::2] = 0.0                                   # exact zeros at window origins: relu'(0) = 1
  y, idx = U.maxpool_relu_fwd(z, (2, 2), 0, 2)
  ugp = f32(rng.standard_normal(y.shape))
  gm = U.maxpool_relu_mask(ugp, y, z)
  full = U.maxpool_relu_bwd(ugp, idx, z, (2, 2), 0, 2)
  # scatter gm by idx on the host: pooled (p', q') sits at q'*Hp + p'; window cell (kh, kw) -> (2p'+kh, 2q'+kw)
  Hp, Wp = Ho // 2, Wo // 2
  gmr, idr = gm.reshape(N, O, Wp, Hp), idx.reshape(N, O, Wp, Hp)
  scat = np.zeros((N, O, Ho, Wo), np.float32)
  for qq in range(Wp):
    for pp in range(Hp):
      a = (idr[:, :, qq, pp] & 3).astype(int)
      for kh in range(2):
        for kw in range(2):
          m = a == kh * 2 + kw
          scat[:, :, 2 * pp + kh, 2 * qq + kw] = np.where(m, gmr[:, :, qq, pp], scat[:, :, 2 * pp + kh, 2 * qq + kw])
  assert np.array_equal(scat, full)
  dw = U.conv_wgrad_pooled(ugp, y, idx, z, xin, (O, C, K, K), 0, 1)
  assert dw is not None
  got = U.conv_wbgrad_pooled(ugp, y, idx, z, xin, (O, C, K, K), 0, 1)
  assert got is not None
  dw2, db2, fused = got
  assert np.array_equal(dw, dw2)                           # same kernel behind both entry points
  oracle_dw = ops.conv3d_wgrad(full.astype(np.float64), xin, (O, C, K, K), 0, 1)
  if fused:                                                # first layers: sparse fp32 kernel (no TF32 rounding)
    close(dw, oracle_dw, 4 * TOL32)
  else:
    ref = U.conv_wgrad(full.astype(np.float64), xin, (O, C, K, K), 0, 1)
    assert np.array_equal(dw, ref)
    close(dw, oracle_dw, TOLTF32)
  close(db2, full.astype(np.float64).sum(axis=(0, 2, 3)), 4 * TOL32)
  close(U.conv_bgrad(gm.astype(np.float64)), U.conv_bgrad(full.astype(np.float64)), 4 * TOL32)
  close(U.conv_bgrad_pooled(ugp, y, z), full.astype(np.float64).sum(axis=(0, 2, 3)), 4 * TOL32)
  close(U.conv_bgrad_pooled_idx(ugp, idx), full.astype(np.float64).sum(axis=(0, 2, 3)), 4 * TOL32)
  if C >= 3:                                               # data gradient: the row-decomposed kernel serves 3 <= C <= 8
    wt = f32(rng.standard_normal((O, C, K, K)) / np.sqrt(C * K * K))
    dxp = U.conv_dgrad_pooled(ugp, y, idx, z, wt, xin.shape, 0, 1)
    assert dxp is not None
    assert np.array_equal(dxp, U.conv_dgrad(full.astype(np.float64), wt, xin.shape, 0, 1))


@pytest.mark.parametrize('N,C,H,W,O,K,stride', [(700, 1, 28, 28, 6, 5, 1), (333, 1, 12, 14, 4, 3, 1), (129, 3, 10, 14, 8, 3, 1),
                                                 (65, 1, 14, 18, 12, 3, 2), (9, 1, 28, 28, 7, 5, 1)])
def test_conv_wbgrad_pooled_single_pass(U, N, C, H, W, O, K, stride):
  """ngb_conv_wbgrad_pooled on first-layer geometries (one fp32 kernel for dw and db: only the argmax cell of a pooling window
  carries gradient) against the float64 oracle on the materialised gradient; several image groups per CTA, every argmax
  position, clamped windows, exact zeros, accumulate on and off."""
  rng = np.random.default_rng(N + C + H + O + K)
  xin = f32(rng.standard_normal((N, C, H, W)))
  Ho, Wo = (H - K) // stride + 1, (W - K) // stride + 1
  assert Ho % 2 == 0 and Wo % 2 == 0
  z = f32(rng.standard_normal((N, O, Ho, Wo)))
  z[0] = -np.abs(z[0]) - 0.5
  z[1, :, ::2, ::2] = 0.0
  y, idx = U.maxpool_relu_fwd(z, (2, 2), 0, 2)
  assert set(np.unique(idx & 3)) == {0, 1, 2, 3}
  ugp = f32(rng.standard_normal(y.shape))
  full = U.maxpool_relu_bwd(ugp, idx, z, (2, 2), 0, 2).astype(np.float64)
  got = U.conv_wbgrad_pooled(ugp, y, idx, z, xin, (O, C, K, K), 0, stride)
  assert got is not None
  dw, db, fused = got
  assert fused
  ref_w = ops.conv3d_wgrad(full, xin, (O, C, K, K), 0, stride)
  ref_b = full.sum(axis=(0, 2, 3))
  close(dw, ref_w, 4 * TOL32)
  close(db, ref_b, 4 * TOL32)
  dw0, db0 = f32(rng.standard_normal(dw.shape)), f32(rng.standard_normal(db.shape))
  dwa, dba, _ = U.conv_wbgrad_pooled(ugp, y, idx, z, xin, (O, C, K, K), 0, stride, dw0=dw0, db0=db0)
  close(dwa, ref_w + dw0, 4 * TOL32)
  close(dba, ref_b + db0, 4 * TOL32)
  again = U.conv_wbgrad_pooled(ugp, y, idx, z, xin, (O, C, K, K), 0, stride)
  assert np.array_equal(again[0], dw) and np.array_equal(again[1], db)       # fixed summation order


@pytest.mark.parametrize('N,H,O,K', [(300, 28, 6, 5), (37, 12, 6, 5), (130, 10, 4, 3), (77, 18, 8, 3), (3, 28, 6, 5)])
def test_conv_relu_pool_forward_in_one_kernel(U, N, H, O, K):
  """ngb_conv_relu_pool_fwd (first-layer block, the conv output is never written) against the separate kernels.  Small
  integer operands make every product and sum exact in fp32, so pooled values, argmax bytes (ties between clamped zeros
  included) and the ReLU flag in bit 2 must be IDENTICAL to ngb_conv_fprop (fp32 mode) -> ngb_maxpool_relu_fwd; real-valued
  operands agree with the float64 oracle to fp32 rounding."""
  rng = np.random.default_rng(N + H + O + K)
  xi = rng.integers(-3, 4, (N, 1, H, H)).astype(np.float64)
  wi = rng.integers(-2, 3, (O, 1, K, K)).astype(np.float64)
  bi = rng.integers(-2, 3, (O,)).astype(np.float64)
  xi[0] = 0.0                                             # z == bias everywhere: all-equal windows, exact zeros for bias 0
  got = U.conv_relu_pool_fwd(xi, wi, bi, 0, 1)
  assert got is not None
  y, idx = got
  with U.precision('fp32'):
    z = U.conv_fprop(xi, wi, bi, 0, 1)
  yr, idxr = U.maxpool_relu_fwd(z, (2, 2), 0, 2)
  assert np.array_equal(y, yr) and np.array_equal(idx, idxr)
  zo = ops.conv3d_fwd(xi, wi, bi, 0, 1)
  yo = ops.maxpool_fwd(np.maximum(zo, 0.0), (2, 2), 0, 2)
  yo = yo[0] if isinstance(yo, tuple) else yo
  assert np.array_equal(y, yo.astype(np.float32))
  xr, wr, br = f32(rng.standard_normal(xi.shape)), f32(rng.standard_normal(wi.shape) / K), f32(rng.standard_normal(O))
  y2, _ = U.conv_relu_pool_fwd(xr, wr, br, 0, 1)
  yo2 = ops.maxpool_fwd(np.maximum(ops.conv3d_fwd(xr, wr, br, 0, 1), 0.0), (2, 2), 0, 2)
  yo2 = yo2[0] if isinstance(yo2, tuple) else yo2
  close(y2, yo2, 4 * TOL32)


@pytest.mark.parametrize('N,C,H,O,K,pad', [(64, 6, 12, 16, 5, 0), (37, 6, 12, 16, 5, 0), (300, 4, 10, 8, 3, 0), (5, 8, 8, 32, 3, 1),
                                           (4096, 6, 12, 16, 5, 0)])
def test_conv_relu_pool_forward_small_channel_layers(U, N, C, H, O, K, pad):
  """The second-layer block (LeNet conv2: 6 -> 16 channels, 8 x 8 output plane) in one kernel: the tensor-core gather kernel
  with the pooling epilogue.  TF32-representable integer operands make the MMA sums exact, so pooled values, argmax bytes
  (ties between clamped zeros included) and the ReLU flag in bit 2 must be IDENTICAL to ngb_conv_fprop -> ngb_maxpool_relu_fwd;
  real-valued operands are bit-identical too (same kernel, same accumulation order) and agree with the oracle to TF32 rounding."""
  rng = np.random.default_rng(N + C + H + O + K)
  xi = rng.integers(-3, 4, (N, C, H, H)).astype(np.float64)
  wi = rng.integers(-2, 3, (O, C, K, K)).astype(np.float64)
  bi = rng.integers(-2, 3, (O,)).astype(np.float64)
  xi[0] = 0.0                                             # z == bias everywhere: all-equal windows, exact zeros for bias 0
  got = U.conv_relu_pool_fwd(xi, wi, bi, pad, 1)
  assert got is not None
  y, idx = got
  z = U.conv_fprop(xi, wi, bi, pad, 1)
  yr, idxr = U.maxpool_relu_fwd(z, (2, 2), 0, 2)
  assert np.array_equal(y, yr) and np.array_equal(idx, idxr)
  zo = ops.conv3d_fwd(xi, wi, bi, pad, 1)
  yo = ops.maxpool_fwd(np.maximum(zo, 0.0), (2, 2), 0, 2)
  yo = yo[0] if isinstance(yo, tuple) else yo
  assert np.array_equal(y, yo.astype(np.float32))
  xr, wr, br = f32(rng.standard_normal(xi.shape)), f32(rng.standard_normal(wi.shape) / K), f32(rng.standard_normal(O))
  y2, i2 = U.conv_relu_pool_fwd(xr, wr, br, pad, 1)
  yr2, ir2 = U.maxpool_relu_fwd(U.conv_fprop(xr, wr, br, pad, 1), (2, 2), 0, 2)
  assert np.array_equal(y2, yr2) and np.array_equal(i2, ir2)
  yo2 = ops.maxpool_fwd(np.maximum(ops.conv3d_fwd(xr, wr, br, pad, 1), 0.0), (2, 2), 0, 2)
  yo2 = yo2[0] if isinstance(yo2, tuple) else yo2
  close(y2, yo2, TOLTF32)


def test_conv_relu_pool_forward_rejects_other_geometries(U):
  x, w, b = np.zeros((2, 1, 12, 12)), np.zeros((6, 1, 5, 5)), np.zeros(6)
  assert U.conv_relu_pool_fwd(x, w, b, 1, 1) is None                                   # padding
  assert U.conv_relu_pool_fwd(np.zeros((2, 1, 13, 13)), w, b, 0, 1) is None            # odd output plane
  assert U.conv_relu_pool_fwd(np.zeros((2, 3, 14, 14)), np.zeros((6, 3, 5, 5)), b, 0, 1) is None   # 10 x 10 plane: neither kernel


def test_maxpool_relu_fused_rejects_other_geometries(U):
  x = np.zeros((2, 3, 9, 9))
  assert U.maxpool_relu_fwd(x, (3, 3), 0, 3) is None
  assert U.maxpool_relu_fwd(np.zeros((2, 3, 8, 6)), (2, 2), 0, 2) is None     # rows not 16-byte multiples


def test_conv_errors(U):
  x, w, b = np.zeros((1, 1, 5, 5)), np.zeros((1, 1, 3, 3)), np.zeros(1)
  import torch
  X, W, Bs, y = U.dev(x), U.dev(w), U.dev(b), torch.zeros(1, 1, 3, 3, device='cuda')
  with pytest.raises(ValueError, match='Stride'):       # conv.py:33 (stride >= 1)
    U.call('ngb_conv_fprop', U.ptr(X), U.ptr(W), U.ptr(Bs), U.ptr(y), 1, 1, 5, 5, 1, 3, 3, 0, 0, U.st())
  with pytest.raises(ValueError, match='Padding'):      # conv.py:34 (padding >= 0)
    U.call('ngb_conv_fprop', U.ptr(X), U.ptr(W), U.ptr(Bs), U.ptr(y), 1, 1, 5, 5, 1, 3, 3, -1, 1, U.st())


# ------------------------------------------------------------------------------------------------- MaxPool
@pytest.mark.parametrize('name', golden_names('maxpool'))
def test_maxpool_golden_bit_exact(U, name):
  g = load_golden(name)
  ks, p, s = (int(g['kh']), int(g['kw'])), int(g['pad']), int(g['stride'])
  x, ug = f32(g['x']), f32(g['ug'])
  y, idx = U.maxpool_fwd(x, ks, p, s)
  y_ref, idx_ref = ops.maxpool_fwd(x, ks, p, s)
  assert np.array_equal(y, y_ref.astype(np.float32))
  assert np.array_equal(idx, idx_ref)                   # bit-exact argmax, first maximum, zero padding participates
  gx = U.maxpool_bwd(ug, idx, x.shape, ks, p, s)
  assert np.array_equal(gx, ops.maxpool_bwd(ug, x, ks, p, s).astype(np.float32))
  if np.array_equal(x, g['x']):                         # fixture inputs representable in fp32 -> reference outputs too
    assert np.array_equal(y, g['out'].astype(np.float32))
    assert np.array_equal(np.where(g['covered'], gx, 0.0), g['gx'].astype(np.float32))


@pytest.mark.parametrize('shape,k,pad,stride', [
  ((512, 24, 24), (2, 2), 0, 2), ((512, 8, 8), (2, 2), 0, 2), ((64, 6, 24, 24), (2, 2), 0, 2), ((64, 16, 8, 8), (2, 2), 0, 2),
  ((2, 3, 12, 15), (3, 3), 0, 3), ((3, 9, 10), (3, 2), 1, 2), ((2, 7, 7), (3, 3), 0, 1), ((5, 13, 11), (2, 2), 0, 2),
])
def test_maxpool_bit_exact_with_ties(U, shape, k, pad, stride):
  rng = np.random.default_rng(sum(shape) + k[0])
  x = f32(rng.standard_normal(shape))
  x = np.maximum(x, 0.0)                 # post-ReLU zeros: many exact ties, incl. all-zero windows
  x[0] = 1.0                             # an all-equal image
  x[-1] = -np.abs(x[-1]) - 1.0           # all negative: with padding the zero pad wins (conv.py:399,416)
  y_ref, idx_ref = ops.maxpool_fwd(x, k, pad, stride)
  y, idx = U.maxpool_fwd(x, k, pad, stride)
  assert np.array_equal(y, y_ref.astype(np.float32))
  assert np.array_equal(idx, idx_ref)
  ug = f32(rng.standard_normal(y.shape))
  assert np.array_equal(U.maxpool_bwd(ug, idx, shape, k, pad, stride), ops.maxpool_bwd(ug, x, k, pad, stride).astype(np.float32))


# ------------------------------------------------------------------------------------------------- SoftmaxCE
@pytest.mark.parametrize('name', golden_names('sce_'))
def test_softmax_ce_golden(U, name):
  g = load_golden(name)
  axis = int(g['axis'])
  z, t = f32(g['logits']), f32(g['targets'])
  loss, dz = U.softmax_ce(z, t, axis)
  close(loss, ops.softmax_ce_fwd(z, t, axis), TOL32)
  close(dz, ops.softmax_ce_bwd(z, t, axis), TOL32)
  close(loss, g['out'], 1e-5)


def test_softmax_ce_batch(U):
  rng = np.random.default_rng(6)
  z = f32(rng.standard_normal((4096, 10)) * 3)
  t = np.eye(10)[rng.integers(0, 10, 4096)]
  loss, dz = U.softmax_ce(z, t, 1)
  close(loss, ops.softmax_ce_fwd(z, t, 1), TOL32)
  close(dz, ops.softmax_ce_bwd(z, t, 1), TOL32)


# ------------------------------------------------------------------------------------------------- optimizers
@pytest.mark.parametrize('name', ['adam', 'gd', 'momentum', 'rmsprop'])
def test_optimizers_golden(U, name):
  """4 steps on the reference's own trajectories (fixtures from the unmodified nn/optim.py)."""
  import torch
  g = load_golden('optim_' + name)
  for i in range(3):
    p = U.dev(np.atleast_1d(f32(g[f'p0_{i}'])).ravel())
    s1, s2 = torch.zeros_like(p), torch.zeros_like(p)
    ref = f32(g[f'p0_{i}'])
    r1 = r2 = 0.0
    for step in range(4):
      gr = f32(g[f'g{step}_{i}'])
      G = U.dev(np.atleast_1d(gr).ravel())
      n = p.numel()
      if name == 'adam':
        t = step + 1
        U.call('ngb_adam_step', U.ptr(p), U.ptr(G), U.ptr(s1), U.ptr(s2), n, 0.05, 0.9, 0.999, 1e-8, t, 1.0, U.st())
        ref, r1, r2 = ops.adam_step(ref, gr, r1, r2, t, 0.05)
      elif name == 'gd':
        U.call('ngb_sgd_step', U.ptr(p), U.ptr(G), n, 0.1, 1.0, U.st())
        ref = ops.sgd_step(ref, gr, 0.1)
      elif name == 'momentum':
        U.call('ngb_momentum_step', U.ptr(p), U.ptr(G), U.ptr(s1), n, 0.1, 0.9, 1.0, U.st())
        ref, r1 = ops.momentum_step(ref, gr, r1, 0.1)
      else:
        U.call('ngb_rmsprop_step', U.ptr(p), U.ptr(G), U.ptr(s1), n, 0.01, 0.9, 1e-8, 1.0, U.st())
        ref, r1 = ops.rmsprop_step(ref, gr, r1, 0.01)
      close(U.host(p).reshape(np.shape(ref)), ref, 1e-5)
      close(U.host(p).reshape(np.shape(ref)), g[f'p{step + 1}_{i}'], 1e-5)


def test_adam_large_flat_arena(U):
  import torch
  rng = np.random.default_rng(8)
  n = 1268241                               # C4 GAN parameter count (odd -> scalar tail)
  p0, g0 = f32(rng.standard_normal(n)), f32(rng.standard_normal(n) * 0.1)
  p, G = U.dev(p0), U.dev(g0)
  m, v = torch.zeros_like(p), torch.zeros_like(p)
  U.call('ngb_adam_step', U.ptr(p), U.ptr(G), U.ptr(m), U.ptr(v), n, 5e-3, 0.9, 0.999, 1e-8, 1, 1.0, U.st())
  ref, rm, rv = ops.adam_step(p0, g0, 0.0, 0.0, 1, 5e-3)
  close(U.host(p), ref, TOL32)
  close(U.host(m), rm, TOL32)
  close(U.host(v), rv, TOL32)
  # graph-friendly variant: iteration count in device memory, incremented by the launch itself
  p2, m2, v2 = U.dev(p0), torch.zeros_like(p), torch.zeros_like(p)
  t_dev = torch.zeros(2, device='cuda', dtype=torch.int32)     # [steps taken, scratch counter]
  half = (n // 2) & ~3
  for _ in range(3):
    # one step = two launches over two ranges of the arena (frozen parameters split it): tick 2 (same step number, not
    # recorded) then tick 1 (recorded by the launch's last block)
    U.call('ngb_adam_step_dev', U.ptr(p2), U.ptr(G), U.ptr(m2), U.ptr(v2), half, 5e-3, 0.9, 0.999, 1e-8, U.ptr(t_dev), 2, 1.0, U.st())
    off = ctypes.c_void_p
    U.call('ngb_adam_step_dev', off(p2.data_ptr() + 4 * half), off(G.data_ptr() + 4 * half), off(m2.data_ptr() + 4 * half),
           off(v2.data_ptr() + 4 * half), n - half, 5e-3, 0.9, 0.999, 1e-8, U.ptr(t_dev), 1, 1.0, U.st())
  ref, rm, rv = p0, 0.0, 0.0
  for t in (1, 2, 3):
    ref, rm, rv = ops.adam_step(ref, g0, rm, rv, t, 5e-3)
  assert t_dev.tolist() == [3, 0]
  close(U.host(p2), ref, TOL32)
