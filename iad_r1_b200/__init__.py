"""Importable alias of the ``iad-r1_b200/`` package directory (a hyphen cannot appear in a Python module name).

All code lives in ``iad-r1_b200/``; this shim only points ``__path__`` there and runs the real package init.
"""
import os as _os

_real = _os.path.join(_os.path.dirname(_os.path.dirname(_os.path.abspath(__file__))), "iad-r1_b200")
__path__.insert(0, _real)
with open(_os.path.join(_real, "__init__.py")) as _f:
    exec(compile(_f.read(), _os.path.join(_real, "__init__.py"), "exec"))
