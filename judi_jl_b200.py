"""Import shim: the package lives in the directory `judi.jl_b200/` (a dot is not importable),
so `import judi_jl_b200` loads that directory as a regular package under this name."""
import importlib.util
import os
import sys

_root = os.path.join(os.path.dirname(os.path.abspath(__file__)), "judi.jl_b200")
_spec = importlib.util.spec_from_file_location(
    __name__, os.path.join(_root, "__init__.py"), submodule_search_locations=[_root])
_mod = importlib.util.module_from_spec(_spec)
sys.modules[__name__] = _mod
_spec.loader.exec_module(_mod)
