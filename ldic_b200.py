"""Import shim: the package directory is named after the reference repository
(`learning-driven-image-compression-algorithm_b200/`, not a valid Python identifier)."""
import importlib.util as _u
import os as _os
import sys as _sys

_d = _os.path.join(_os.path.dirname(_os.path.abspath(__file__)), "learning-driven-image-compression-algorithm_b200")
_spec = _u.spec_from_file_location("ldic_b200", _os.path.join(_d, "__init__.py"), submodule_search_locations=[_d])
_mod = _u.module_from_spec(_spec)
_sys.modules["ldic_b200"] = _mod
_spec.loader.exec_module(_mod)
