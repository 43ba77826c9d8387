"""Importable alias for the package directory ``street-gaussians-ns_b200/``.

The product lives in ``street-gaussians-ns_b200/`` (the name the build contract fixes); a hyphen
is not importable, so this stub points ``__path__`` at that directory and executes its
``__init__``.  ``import street_gaussians_ns_b200 as sgn`` is the supported spelling.
"""
import os as _os

_real = _os.path.join(_os.path.dirname(_os.path.dirname(_os.path.abspath(__file__))), "street-gaussians-ns_b200")
__path__ = [_real]
with open(_os.path.join(_real, "__init__.py")) as _f:
    exec(compile(_f.read(), _os.path.join(_real, "__init__.py"), "exec"))
del _os, _f
