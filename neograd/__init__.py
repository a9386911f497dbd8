"""Drop-in alias: `import neograd as ng` resolves to the B200 backend (`neograd_b200`), so scripts written against
pranftw/neograd (README.md:67-125, examples/*.ipynb) run unchanged with `Tensor.data` resident on the GPU.
Sub-modules are aliased too (`neograd.nn`, `neograd.nn.loss`, `neograd.autograd.utils`, ...)."""
import sys

import neograd_b200 as _impl

_prefix = _impl.__name__
for _name, _mod in list(sys.modules.items()):
  if _name == _prefix or _name.startswith(_prefix + '.'):
    sys.modules['neograd' + _name[len(_prefix):]] = _mod
sys.modules[__name__] = _impl
