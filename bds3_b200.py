"""Import shim: ``import bds3_b200`` loads the package that lives in the
hyphenated directory ``bds-3-b1c-b2a-sdr-receiver_b200/`` (not a valid Python
identifier, so it cannot be imported by name)."""
import importlib.util
import os
import sys

_dir = os.path.join(os.path.dirname(os.path.abspath(__file__)), "bds-3-b1c-b2a-sdr-receiver_b200")
_spec = importlib.util.spec_from_file_location(
    "bds3_b200", os.path.join(_dir, "__init__.py"), submodule_search_locations=[_dir])
_mod = importlib.util.module_from_spec(_spec)
sys.modules["bds3_b200"] = _mod
_spec.loader.exec_module(_mod)
