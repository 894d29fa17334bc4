"""Import shim for the LIVE reference (mtflearn) -- test infrastructure only.

TEST INFRASTRUCTURE.  Nothing under ``oracle/`` is imported by the product
package; only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s
cpu-baseline / reference arm may use it.

The reference package cannot be imported as shipped in this image because
``mtflearn/features/__init__.py:5`` pulls ``_window_size.py:2`` which imports
matplotlib (absent); skimage and h5py are absent too.
This module installs a ``sys.meta_path`` finder that answers those three
top-level names with ``MagicMock`` modules and then imports the real package
from ``/root/reference``.  It is used ONLY in this container (the GPU box has no
``/root/reference``) to (i) generate the committed golden vectors
(``oracle/make_goldens.py``) and (ii) validate the numpy restatement
(``oracle/zernike_oracle.py``) against the real code.
"""
from __future__ import annotations

import importlib.abc
import importlib.machinery
import os
import sys
from unittest import mock

REFERENCE_ROOT = os.environ.get("MTFLEARN_REFERENCE_ROOT", "/root/reference")
_STUBBED = ("matplotlib", "skimage", "h5py", "numba", "tqdm")


class _StubLoader(importlib.abc.Loader):
    def create_module(self, spec):
        m = mock.MagicMock(name=spec.name)
        m.__name__ = spec.name
        m.__path__ = []          # behave like a package so sub-imports resolve
        m.__spec__ = spec
        m.__loader__ = self
        return m

    def exec_module(self, module):
        return None


class _StubFinder(importlib.abc.MetaPathFinder):
    def find_spec(self, fullname, path=None, target=None):
        top = fullname.split(".")[0]
        if top in _STUBBED:
            try:
                # prefer the real module when it exists (e.g. tqdm)
                sys.meta_path.remove(self)
                try:
                    real = importlib.util.find_spec(fullname) if "." not in fullname else None
                finally:
                    sys.meta_path.insert(0, self)
                if real is not None:
                    return None
            except Exception:
                pass
            return importlib.machinery.ModuleSpec(fullname, _StubLoader(), is_package=True)
        return None


def available() -> bool:
    return os.path.isdir(os.path.join(REFERENCE_ROOT, "mtflearn"))


def load_reference():
    """Return the real ``mtflearn`` package (raises if /root/reference is absent)."""
    if not available():
        raise RuntimeError(f"reference tree not present at {REFERENCE_ROOT}")
    if "mtflearn" in sys.modules:
        return sys.modules["mtflearn"]
    import importlib.util  # noqa: F401  (used by the finder)
    if not any(isinstance(f, _StubFinder) for f in sys.meta_path):
        sys.meta_path.insert(0, _StubFinder())
    if REFERENCE_ROOT not in sys.path:
        sys.path.insert(0, REFERENCE_ROOT)
    import mtflearn  # noqa: E402
    return mtflearn
