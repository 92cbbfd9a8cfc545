"""Importable alias for the product package.

The product lives in the directory the build contract names,
``physics-informed-neural-network-for-explainable-fault-diagnosis-in-fuel-cells_b200/``,
whose name is not a valid Python identifier.  This shim makes it importable as
``b200pinn`` by pointing ``__path__`` at that directory and executing its
``__init__.py`` in this module's namespace.
"""
import os as _os

_PKG_DIR = _os.path.join(
    _os.path.dirname(_os.path.dirname(_os.path.abspath(__file__))),
    "physics-informed-neural-network-for-explainable-fault-diagnosis-in-fuel-cells_b200",
)
__path__ = [_PKG_DIR]
with open(_os.path.join(_PKG_DIR, "__init__.py")) as _f:
    exec(compile(_f.read(), _os.path.join(_PKG_DIR, "__init__.py"), "exec"))
del _f
