"""GPU tests of the device-side input pipeline (csrc/input.cu, neograd_b200/nn/data.py): the NumPy preprocessing of the
reference's notebooks -- examples/digits.ipynb cell 5 (per-sample standardisation) and cell 6 (one-hot), examples/gans
(np.random.randn noise) -- as device kernels behind a double-buffered host -> device loader."""
import ctypes

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


@pytest.fixture(scope='module')
def ng():
  import torch
  if not torch.cuda.is_available():
    pytest.skip('needs a CUDA device')
  import neograd_b200
  return neograd_b200


def test_prep_kernels_match_numpy(ng):
  import torch
  from neograd_b200 import _backend as B
  B.ensure_runtime()
  P = lambda t: ctypes.c_void_p(t.data_ptr())
  S = lambda: ctypes.c_void_p(B.stream())
  rng = np.random.default_rng(0)
  for n in (1, 15, 16, 4096 * 784 + 5):
    u8 = rng.integers(0, 256, n, dtype=np.uint8)
    src, dst = torch.from_numpy(u8).cuda(), torch.full((n,), float('nan'), device='cuda')
    B.check(B.lib.ngb_prep_u8_affine(P(src), P(dst), n, 1 / 73.9, -127.5 / 73.9, S()))
    want = u8.astype(np.float32) * np.float32(1 / 73.9) + np.float32(-127.5 / 73.9)
    assert np.allclose(dst.cpu().numpy(), want, rtol=0, atol=2e-7)
  for rows, cols, u8 in ((1347, 64, False), (200, 64, True), (33, 784, True), (5, 7, False)):
    x = rng.integers(0, 17, (rows, cols)).astype(np.uint8 if u8 else np.float32)
    if not u8:
      x = (x + rng.standard_normal((rows, cols))).astype(np.float32)
    src, dst = torch.from_numpy(x).cuda(), torch.full((rows, cols), float('nan'), device='cuda')
    B.check(B.lib.ngb_prep_standardize_rows(P(src), int(u8), P(dst), rows, cols, S()))
    xf = x.astype(np.float64)
    want = (xf - xf.mean(axis=1, keepdims=True)) / xf.std(axis=1, keepdims=True)          # digits.ipynb cell 5
    assert np.allclose(dst.cpu().numpy(), want, rtol=1e-5, atol=1e-5)
  labels = rng.integers(0, 10, 4097).astype(np.int32)
  lab, oh = torch.from_numpy(labels).cuda(), torch.full((4097, 10), float('nan'), device='cuda')
  B.check(B.lib.ngb_prep_one_hot(P(lab), P(oh), 4097, 10, S()))
  assert np.array_equal(oh.cpu().numpy(), np.eye(10, dtype=np.float32)[labels])             # digits.ipynb cell 6


def test_randn_is_standard_normal_and_counter_based(ng):
  import torch
  a = ng.nn.randn((8192, 784), seed=3).data.numpy()
  assert abs(a.mean()) < 2e-3 and abs(a.std() - 1) < 2e-3
  assert abs((a ** 3).mean()) < 1e-2 and abs((a ** 4).mean() - 3) < 3e-2                     # skewness 0, kurtosis 3
  assert np.abs(a).max() < 7 and np.isfinite(a).all()
  b = ng.nn.randn((8192, 784), seed=3).data.numpy()
  assert np.array_equal(a, b)                                                                # same (seed, offset): same draw
  c = ng.nn.randn((8192, 784), seed=3, offset=1).data.numpy()
  d = ng.nn.randn((8192, 784), seed=4).data.numpy()
  assert abs(np.corrcoef(a.ravel(), c.ravel())[0, 1]) < 5e-3 and abs(np.corrcoef(a.ravel(), d.ravel())[0, 1]) < 5e-3
  # device-resident call counter: consecutive calls draw consecutive offsets, also under graph replay
  cnt = torch.zeros(1, dtype=torch.int32, device='cuda')
  t = ng.nn.randn((1000,), seed=3, counter=cnt)
  first = t.data.numpy().copy()
  ng.nn.randn(None, seed=3, out=t, counter=cnt)
  assert int(cnt.item()) == 2 and not np.array_equal(first, t.data.numpy())
  assert np.array_equal(first, ng.nn.randn((1000,), seed=3, offset=0).data.numpy())
  assert np.array_equal(t.data.numpy(), ng.nn.randn((1000,), seed=3, offset=1).data.numpy())
  odd = ng.nn.randn((7,), seed=9).data.numpy()                                                # n not a multiple of 4
  assert odd.shape == (7,) and np.isfinite(odd).all()


@pytest.mark.parametrize('ahead', [False, True])
def test_device_loader_epochs(ng, ahead):
  """ahead: the normalisation / one-hot kernels run on the copy stream behind each batch's H2D copy (prepare_ahead)."""
  rng = np.random.default_rng(1)
  n, bs = 1000, 256
  imgs = rng.integers(0, 256, (n, 1, 28, 28), dtype=np.uint8)
  labels = rng.integers(0, 10, n)
  loader = ng.nn.DeviceLoader(imgs, labels, bs, normalize=(1 / 255.0, -0.5), num_classes=10, prepare_ahead=ahead)
  assert len(loader) == 4 and loader.bytes_per_batch == bs * 784 + bs * 4
  for _ in range(2):                                   # two epochs over the same loader
    seen = 0
    for X, Y in loader:
      rows = X.shape[0]
      want = imgs[seen:seen + rows].astype(np.float32) / np.float32(255.0) - np.float32(0.5)
      assert X.shape == (rows, 1, 28, 28) and Y.shape == (rows, 10)
      assert np.allclose(X.data.numpy(), want, atol=1e-6)
      assert np.array_equal(Y.data.numpy(), np.eye(10)[labels[seen:seen + rows]])
      seen += rows
    assert seen == n                                   # the last batch is short (1000 = 3 * 256 + 232), as with get_batches
  # per-sample standardisation of float inputs, float targets copied as they are, no one-hot (digits.ipynb cells 5-6)
  xs, ys = rng.standard_normal((300, 8, 8)) * 3 + 1, rng.standard_normal((300, 10))
  loader = ng.nn.DeviceLoader(xs, ys, 128, normalize='per_sample', prepare_ahead=ahead)
  flat = xs.reshape(300, 64)
  want = ((flat - flat.mean(axis=1, keepdims=True)) / flat.std(axis=1, keepdims=True)).reshape(300, 8, 8)
  seen = 0
  for X, Y in loader:
    rows = X.shape[0]
    assert np.allclose(X.data.numpy(), want[seen:seen + rows], atol=2e-5)
    assert np.allclose(Y.data.numpy(), ys[seen:seen + rows].astype(np.float32), atol=0)
    seen += rows
  assert seen == 300
  with pytest.raises(ValueError):
    ng.nn.DeviceLoader(xs, ys, 0)


def test_device_loader_trains_digits_like_model(ng):
  """The loader feeds an ordinary training loop (README.md:109-115) and gives the same losses as preparing the batches in
  NumPy and handing them to ng.tensor, as the reference's notebooks do."""
  from neograd_b200.configs import DigitsCNN, train_step
  rng = np.random.default_rng(2)
  x = rng.integers(0, 17, (400, 8, 8)).astype(np.uint8)
  y = rng.integers(0, 10, 400)
  losses = []
  for use_loader in (False, True):
    np.random.seed(0)
    model = DigitsCNN()
    optim = ng.nn.optim.Adam(model.parameters(), 5e-3)
    loss_fn = ng.nn.loss.SoftmaxCE(axis=1)
    run = []
    if use_loader:
      for X, Y in ng.nn.DeviceLoader(x, y, 200, normalize='per_sample', num_classes=10):
        run.append(float(train_step(model, optim, loss_fn, X, Y).data))
    else:
      flat = x.reshape(400, 64).astype(np.float64)
      xn = ((flat - flat.mean(axis=1, keepdims=True)) / flat.std(axis=1, keepdims=True)).reshape(400, 8, 8)
      for s in range(0, 400, 200):
        run.append(float(train_step(model, optim, loss_fn, ng.tensor(xn[s:s + 200]), ng.tensor(np.eye(10)[y[s:s + 200]])).data))
    losses.append(run)
  assert np.allclose(losses[0], losses[1], rtol=2e-5), losses
