"""Import alias: the package directory is named `ode.jl_b200` (not a valid Python
identifier), so `import ode_jl_b200` loads it from that directory."""
import importlib.util
import os
import sys

_dir = os.path.join(os.path.dirname(os.path.abspath(__file__)), "ode.jl_b200")
_spec = importlib.util.spec_from_file_location("ode_jl_b200", os.path.join(_dir, "__init__.py"),
                                               submodule_search_locations=[_dir])
_mod = importlib.util.module_from_spec(_spec)
sys.modules["ode_jl_b200"] = _mod
_spec.loader.exec_module(_mod)
