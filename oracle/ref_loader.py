"""TEST INFRASTRUCTURE ONLY -- load the UNMODIFIED reference (``/root/reference/bask``)
in the build container.

The reference cannot be imported as-is here: ``skopt``, ``emcee`` and ``arviz`` are not
installed (no network) and ``bask/__init__.py:10`` asks importlib.metadata for an
installed distribution.  ``load_reference()`` therefore
  1. puts ``oracle/ref_shims`` (thin packages re-exporting oracle/skopt_port.py and
     oracle/emcee_port.py) first on ``sys.path``;
  2. registers a bare package object named ``bask`` whose ``__path__`` is the read-only
     reference tree, which skips ``bask/__init__.py`` but imports every other reference
     module verbatim from ``/root/reference/bask``.
Nothing is copied.  /root/reference does not exist on the GPU box: callers must guard
with ``reference_available()``.
"""
import importlib
import os
import sys
import types

REFERENCE_ROOT = "/root/reference"
_SHIMS = os.path.join(os.path.dirname(os.path.abspath(__file__)), "ref_shims")
_REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def reference_available():
    return os.path.isdir(os.path.join(REFERENCE_ROOT, "bask"))


def load_reference():
    """Returns the ``bask`` package object with bayesgpr/acquisition/optimizer/utils/priors
    imported from the reference tree."""
    if not reference_available():
        raise RuntimeError("/root/reference is not present on this machine")
    sys.dont_write_bytecode = True  # the reference tree is read-only
    for p in (_REPO, _SHIMS):
        if p not in sys.path:
            sys.path.insert(0, p)
    if "bask" not in sys.modules or getattr(sys.modules["bask"], "__graft_ref__", False) is False:
        pkg = types.ModuleType("bask")
        pkg.__path__ = [os.path.join(REFERENCE_ROOT, "bask")]
        pkg.__graft_ref__ = True
        sys.modules["bask"] = pkg
    pkg = sys.modules["bask"]
    for name in ("priors", "init", "utils", "bayesgpr", "acquisition", "optimizer"):
        setattr(pkg, name, importlib.import_module("bask." + name))
    return pkg
