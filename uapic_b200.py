"""Import shim: the package directory is `uapic.jl_b200/` (a dot cannot be part of a Python package name), so this
module loads it under the name `uapic_b200`.  `import uapic_b200` from the repo root gives the package."""
import importlib.util
import os
import sys

_dir = os.path.join(os.path.dirname(os.path.abspath(__file__)), "uapic.jl_b200")
_spec = importlib.util.spec_from_file_location("uapic_b200", os.path.join(_dir, "__init__.py"), submodule_search_locations=[_dir])
_mod = importlib.util.module_from_spec(_spec)
sys.modules["uapic_b200"] = _mod
_spec.loader.exec_module(_mod)
