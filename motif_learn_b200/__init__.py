"""Importable alias of the ``motif-learn_b200/`` source directory.

The hyphen in the project name is not a valid Python identifier, so this stub package points
its ``__path__`` at ``motif-learn_b200/`` and executes that directory's ``__init__``:
``import motif_learn_b200`` imports the real sources, nothing is duplicated.
"""
import os as _os

_real = _os.path.join(_os.path.dirname(_os.path.dirname(_os.path.abspath(__file__))), "motif-learn_b200")
__path__ = [_real]
__file__ = _os.path.join(_real, "__init__.py")
with open(__file__, "r", encoding="utf-8") as _f:
    exec(compile(_f.read(), __file__, "exec"))
del _f, _os
