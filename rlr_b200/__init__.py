"""Importable alias for the hyphenated package directory.

The framework lives in ``defending-against-backdoors-with-robust-learning-rate_b200/`` (the name the
build spec asks for); Python cannot import a hyphenated name, so ``import rlr_b200`` points its
``__path__`` at that directory and executes its ``__init__``.
"""
import os as _os

_REAL = _os.path.join(_os.path.dirname(_os.path.dirname(_os.path.abspath(__file__))),
                      "defending-against-backdoors-with-robust-learning-rate_b200")
__path__ = [_REAL]
__file__ = _os.path.join(_REAL, "__init__.py")
with open(__file__) as _f:
    exec(compile(_f.read(), __file__, "exec"))
