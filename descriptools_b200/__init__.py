"""descriptools_b200 -- B200-native (sm_100a CUDA) implementation of descriptools' per-cell
terrain-descriptor path, behind the reference's own module functions and NumPy signatures.

Like the reference (`descriptools/__init__.py` is empty, callers write
`import descriptools.slope as slope`, Example/example.py:11-16), submodules are imported
explicitly:

    import descriptools_b200.slope as slope
    import descriptools_b200.flowhand as flowhand
    ...

Importing the package loads libdtb200.so and raises if it is missing (no CPU fallback).
"""
from . import _lib  # noqa: F401  (fail loudly at import time when the CUDA library is absent)

__version__ = "0.1.0"
