"""Import shim: the product package lives in the directory ``r3det-pytorch_b200/`` (not a valid Python
identifier), and is importable as ``r3det_b200``.  ``import r3det_b200`` loads that directory as a regular
package (submodules: ``r3det_b200.rbbox_geo``, ``r3det_b200.rnms``, ...)."""
import importlib.util
import os
import sys

_dir = os.path.join(os.path.dirname(os.path.abspath(__file__)), "r3det-pytorch_b200")
_spec = importlib.util.spec_from_file_location(
    "r3det_b200", os.path.join(_dir, "__init__.py"), submodule_search_locations=[_dir])
_mod = importlib.util.module_from_spec(_spec)
sys.modules["r3det_b200"] = _mod
_spec.loader.exec_module(_mod)
