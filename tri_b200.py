"""Importable alias of the package directory `3d-reconstruction-triangulation_b200/` (its name is not a
Python identifier).  `import tri_b200` gives the ctypes binding of libtri_b200.so."""
import importlib.util
import os
import sys

_here = os.path.dirname(os.path.abspath(__file__))
_pkg = os.path.join(_here, "3d-reconstruction-triangulation_b200")
_spec = importlib.util.spec_from_file_location("tri_b200", os.path.join(_pkg, "__init__.py"),
                                               submodule_search_locations=[_pkg])
_mod = importlib.util.module_from_spec(_spec)
sys.modules["tri_b200"] = _mod
_spec.loader.exec_module(_mod)
