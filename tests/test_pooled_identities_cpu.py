"""CPU checks (float64 oracle only) of the identities the fused Conv -> ReLU -> MaxPool(2x2, s2) kernels rely on
(DESIGN.md sections 4 and 5).  They say nothing about the device code -- the `-m gpu` tests do that -- but they pin the
algebra: if one of these failed, the kernels would be computing the wrong thing by construction.

  * the ReLU-gradient flag carried in bit 2 of the pool's argmax byte: relu'(z at the argmax) == (y > 0) or (z_first >= 0)
  * conv weight / bias gradient from the POOLED gradient (one cell per window carries gradient: ngb_conv_wbgrad_pooled)
    == conv.py:311-327 applied to the materialised full-resolution gradient
  * the layout of that cell: pooled element (p', q') is stored at q'*Hp + p'; its window covers stored rows 2p'.., columns
    2q'.. of the conv output plane, whose stored cell (r, c) is conv position (p, q) = (c, r)   
"""
import os
import sys

import numpy as np
import pytest

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import ops  # noqa: E402


def _block(seed, N, C, H, O, K):
  rng = np.random.default_rng(seed)
  x = rng.standard_normal((N, C, H, H))
  w = rng.standard_normal((O, C, K, K)) / K
  b = rng.standard_normal(O)
  z = ops.conv3d_fwd(x, w, b)                     # (N, O, Ho, Wo) in neograd (stored) order
  z[0] = -np.abs(z[0]) - 0.5                      # fully clamped windows
  z[1, :, ::2, ::2] = 0.0                         # exact zeros at window origins: relu'(0) = 1 (activations.py:31)
  r = np.maximum(z, 0.0)
  y, idx = ops.maxpool_fwd(r, (2, 2), 0, 2)       # stored order: (p', q') at q'*Hp + p'
  return rng, x, w, z, y, idx


@pytest.mark.parametrize('N,C,H,O,K', [(5, 1, 12, 6, 5), (4, 3, 10, 4, 3), (3, 1, 28, 6, 5)])
def test_relu_flag_rule_and_pooled_gradients(N, C, H, O, K):
  rng, x, w, z, y, idx = _block(N + H + O, N, C, H, O, K)
  Ho = H - K + 1
  Hp = Ho // 2
  # ---- stored planes as 2-D arrays: z_s[r][c] (pool input), pooled arrays P[q'][p']
  z_s = z.reshape(N, O, Ho, Ho)
  yP, iP = y.reshape(N, O, Hp, Hp), idx.reshape(N, O, Hp, Hp).astype(int)
  kh, kw = iP >> 1, iP & 1                        # argmax cell inside the window, row-major (kh, kw)
  qq, pp = np.meshgrid(np.arange(Hp), np.arange(Hp), indexing='ij')   # P[q'][p']
  r_arg, c_arg = 2 * pp + kh, 2 * qq + kw        # stored cell of the argmax
  nn, oo = np.arange(N)[:, None, None, None], np.arange(O)[None, :, None, None]
  z_arg = z_s[nn, oo, r_arg, c_arg]
  z_first = z_s[nn, oo, 2 * pp, 2 * qq]
  assert np.array_equal(np.maximum(z_arg, 0.0), yP)                     # the layout above is the pool's
  flag = (yP > 0) | (z_first >= 0)                                      # what the fused forward kernels store in bit 2
  assert np.array_equal(flag, z_arg >= 0)                               # == relu'(z at the argmax)

  # ---- backward through pool and ReLU, materialised (the reference's path) ...
  ugp = rng.standard_normal(y.shape)
  full = ops.maxpool_bwd(ugp, np.maximum(z, 0.0), (2, 2), 0, 2) * (z >= 0)
  dw_ref = ops.conv3d_wgrad(full, x, w.shape)
  db_ref = ops.conv3d_bgrad(full).reshape(-1)
  # ... and straight from the pooled gradient: gm * x[patch at the argmax cell's conv position (p, q) = (c, r)]
  gm = ugp.reshape(N, O, Hp, Hp) * flag
  dw = np.zeros_like(w)
  xs = np.moveaxis(x, 1, -1)                                            # (N, H, W, C)
  for a in range(K):
    for b_ in range(K):
      px = xs[nn, c_arg + a, r_arg + b_]                                # (N, O, Hp, Hp, C): x[n, :, p + a, q + b]
      dw[:, :, a, b_] = np.einsum('nopq,nopqc->oc', gm, px)
  assert ops.dist(dw, dw_ref) < 1e-12
  assert ops.dist(gm.sum(axis=(0, 2, 3)), db_ref) < 1e-12


def test_argmax_ties_pick_the_first_cell():
  """All-equal windows (post-ReLU zeros): np.argmax returns cell 0, so the flag rule reads z at the window's first cell."""
  z = -np.ones((1, 1, 4, 4))
  y, idx = ops.maxpool_fwd(np.maximum(z, 0.0), (2, 2), 0, 2)
  assert np.all(idx == 0) and np.all(y == 0)
