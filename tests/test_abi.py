"""CPU checks of the drop-in boundary: the library loads, exports every symbol include/ngb200.h declares, the ctypes
table matches the header, and the product package never reaches into the oracle (no CPU fallback)."""
import os
import re

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def header_functions():
  text = open(os.path.join(ROOT, 'include', 'ngb200.h')).read()
  text = re.sub(r'/\*.*?\*/', '', text, flags=re.S)
  decls = re.findall(r'\b(?:int|int64_t|const char\*)\s+(ngb_\w+)\s*\(([^;]*?)\)\s*;', text, flags=re.S)
  return {name: [a.strip() for a in args.split(',')] if args.strip() != 'void' else [] for name, args in decls}


def test_library_exports_every_declared_symbol():
  import ctypes
  from neograd_b200 import _backend as B
  fns = header_functions()
  assert len(fns) >= 30
  cdll = ctypes.CDLL(B.LIB_PATH)
  for name in fns:
    assert hasattr(cdll, name), f'{name} is declared in include/ngb200.h but not exported by libngb200.so'
  assert cdll.ngb_abi_version() == 1


def test_ctypes_table_matches_header():
  from neograd_b200 import _backend as B
  fns = header_functions()
  assert set(fns) == set(B.SIGNATURES), set(fns) ^ set(B.SIGNATURES)
  for name, args in fns.items():
    assert len(args) == len(B.SIGNATURES[name][0]), (name, args, B.SIGNATURES[name][0])


def test_no_compute_without_gpu_and_error_reporting():
  import torch
  from neograd_b200 import _backend as B
  if torch.cuda.is_available():
    return
  try:
    B.ensure_runtime()
  except B.NgbError as e:
    assert 'no CPU fallback' in str(e)
  else:
    raise AssertionError('ensure_runtime() must fail loudly without a CUDA device')


def test_product_never_imports_oracle():
  bad = []
  for dp, _, files in os.walk(os.path.join(ROOT, 'neograd_b200')):
    for f in files:
      if f.endswith(('.py', '.cu', '.cuh', '.h')):
        src = open(os.path.join(dp, f)).read()
        if re.search(r'^\s*(from|import)\s+oracle\b', src, flags=re.M) or 'oracle/' in src and f.endswith('.py') and 'import' in src and re.search(r'import.*oracle', src):
          bad.append(os.path.join(dp, f))
  assert not bad, bad


def test_neograd_alias_package():
  """`import neograd as ng` (the reference's import name) resolves to the device backend, sub-modules included."""
  import neograd as ng
  from neograd import nn
  from neograd.nn import loss, optim
  from neograd.autograd.utils import grad_check, fn_grad_check
  import neograd_b200
  assert ng is neograd_b200 and nn is neograd_b200.nn and loss.SoftmaxCE is neograd_b200.nn.loss.SoftmaxCE
  assert optim.Adam is neograd_b200.nn.optim.Adam and callable(grad_check) and callable(fn_grad_check)
  for name in ('tensor', 'new_graph', 'no_track', 'add', 'sub', 'mul', 'div', 'pow', 'exp', 'log', 'dot', 'sum', 'transpose',
               'flatten', 'reshape', 'load', 'save', 'Checkpoint', 'nn', 'autograd'):   # neograd/__init__.py:1-6
    assert hasattr(ng, name), name
  for name in ('Model', 'Sequential', 'Dropout', 'Linear', 'Conv2D', 'Conv3D', 'MaxPool2D', 'MaxPool3D', 'ReLU', 'Sigmoid', 'Tanh',
               'Softmax', 'LeakyReLU', 'Checkpoint', 'save_model', 'load_model'):       # neograd/nn/__init__.py:1-5
    assert hasattr(nn, name), name


def test_wgrad_tile_layout_search_is_conflict_free_for_lenet():
  """Host-side logic of the small-channel weight-gradient kernel (csrc/conv_small_mma.cu:search_x_layout): lane (gq, t) of
  an A-gather load reads word moff[tap gq] + poff[position t]; the searched strides / plane pitch / tap order must bring
  both LeNet-B layers to ONE shared-memory wavefront per load (the natural pitch needs ~2), keep images 16-byte multiples
  and leave the tile large enough for every tap of every position."""
  import ctypes
  from neograd_b200 import _backend as B
  for (C, H, W, K) in ((1, 28, 28, 5), (6, 12, 12, 5), (3, 10, 14, 3), (4, 9, 11, 3)):
    out = (ctypes.c_int * 6)()
    B.check(B.lib.ngb_debug_wgrad_layout(C, H, W, K, K, 0, 1, ctypes.cast(out, ctypes.c_void_p)))
    sr, sc, plane, khfast, cost, natural = list(out)
    assert (C * plane) % 4 == 0 and khfast in (0, 1) and 1 in (sr, sc)
    assert (H - 1) * sr + (W - 1) * sc < plane                      # every pixel of a plane fits
    assert cost <= natural, (C, H, W, K, cost, natural)
    if (C, H) in ((1, 28), (6, 12)):
      assert cost == 1000 and natural >= 1700, (cost, natural)
