"""CPU oracle for the neograd tensor-op hot path -- TEST INFRASTRUCTURE ONLY.

This package is a float64 NumPy restatement of the arithmetic that pranftw/neograd performs on
its hot path (Dot, Conv2D/Conv3D, MaxPool2D/MaxPool3D, broadcasting elementwise ops with
`unbroadcast`, SoftmaxCE, Adam/GD/Momentum/RMSProp).  It exists so that the CUDA path in
`neograd_b200` can be checked on a GPU box where the reference tree is absent.

Rules (enforced by tests/test_no_oracle_in_product.py):
  * only `tests/`, `__graft_entry__.smoke()` and `bench.py`'s cpu_baseline / `--impl reference`
    legs may import anything from here;
  * nothing under `neograd_b200/` imports it -- the product path has no CPU fallback.

Parity status: PINNED.  The reference publishes no golden vectors, so the oracle is
pinned against outputs of the unmodified reference itself, generated in the build container by
`oracle/gen_golden.py` (imports /root/reference read-only) and committed under `tests/golden/`.
`tests/test_oracle_golden.py` replays every fixture against this restatement.
"""
from . import ops, models  # noqa: F401
