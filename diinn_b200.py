"""Import shim: exposes the package directory ``dual-interactive-implicit-neural-network_b200/`` (whose name is
not a valid Python identifier) as the importable package ``diinn_b200``."""
import importlib.util
import os
import sys

_dir = os.path.join(os.path.dirname(os.path.abspath(__file__)), "dual-interactive-implicit-neural-network_b200")
_spec = importlib.util.spec_from_file_location(
    "diinn_b200", os.path.join(_dir, "__init__.py"), submodule_search_locations=[_dir])
_mod = importlib.util.module_from_spec(_spec)
sys.modules["diinn_b200"] = _mod
_spec.loader.exec_module(_mod)
