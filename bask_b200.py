"""Import shim: the product package lives in the directory ``bayes-skopt_b200/`` (not a valid
Python identifier), and is importable as ``bask_b200``."""
import importlib.util
import os
import sys

_dir = os.path.join(os.path.dirname(os.path.abspath(__file__)), "bayes-skopt_b200")
_spec = importlib.util.spec_from_file_location(
    "bask_b200", os.path.join(_dir, "__init__.py"), submodule_search_locations=[_dir])
_mod = importlib.util.module_from_spec(_spec)
sys.modules["bask_b200"] = _mod
_spec.loader.exec_module(_mod)
