"""neograd_b200 -- B200 (sm_100a) device backend for neograd's tensor-op hot path.

`_backend` binds libngb200.so (C ABI in include/ngb200.h).  Importing this package fails loudly when the
library has not been built; there is no CPU fallback.
"""
from . import _backend  # noqa: F401
