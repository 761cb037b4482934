"""Import shim: `import bsm_b200` loads the package directory blocksparsematrices.jl_b200/ (its name
is not a valid Python identifier) under the module name bsm_b200."""
import importlib.util
import sys
from pathlib import Path

_pkg = Path(__file__).resolve().parent / "blocksparsematrices.jl_b200"
_spec = importlib.util.spec_from_file_location("bsm_b200", _pkg / "__init__.py",
                                               submodule_search_locations=[str(_pkg)])
_mod = importlib.util.module_from_spec(_spec)
sys.modules["bsm_b200"] = _mod
_spec.loader.exec_module(_mod)
