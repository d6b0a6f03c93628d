"""Importable alias of the package directory ``tensor-cuda-fft-_b200/`` (whose name is not a Python identifier).

``import tensor_cuda_fft_b200`` executes ``tensor-cuda-fft-_b200/__init__.py`` with this module's ``__path__``
pointing at that directory, so ``tensor_cuda_fft_b200.spectral_layers`` etc. resolve to the real files.
"""
import os as _os

_REAL = _os.path.join(_os.path.dirname(_os.path.dirname(_os.path.abspath(__file__))), "tensor-cuda-fft-_b200")
__path__ = [_REAL]
with open(_os.path.join(_REAL, "__init__.py")) as _f:
    exec(compile(_f.read(), _os.path.join(_REAL, "__init__.py"), "exec"))
del _f
