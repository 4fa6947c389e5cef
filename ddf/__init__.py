"""Import shim: makes ``import ddf.jl_b200`` resolve to the package directory
``ddf.jl_b200/`` at the repo root (a directory name with a dot cannot be imported
by the normal path finder)."""
import importlib.util
import os
import sys

_root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
_pkg = os.path.join(_root, "ddf.jl_b200")

if "ddf.jl_b200" not in sys.modules:
    _spec = importlib.util.spec_from_file_location(
        "ddf.jl_b200", os.path.join(_pkg, "__init__.py"), submodule_search_locations=[_pkg])
    _mod = importlib.util.module_from_spec(_spec)
    sys.modules["ddf.jl_b200"] = _mod
    _spec.loader.exec_module(_mod)
jl_b200 = sys.modules["ddf.jl_b200"]
