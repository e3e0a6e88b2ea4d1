"""Import shim: the package directory is named ``colibri-core_b200`` (with a hyphen, as the project layout asks),
which Python cannot import by name.  ``import colibri_core_b200`` loads that directory as this module."""
import importlib.util
import os
import sys

_dir = os.path.join(os.path.dirname(os.path.abspath(__file__)), "colibri-core_b200")
_spec = importlib.util.spec_from_file_location("colibri_core_b200", os.path.join(_dir, "__init__.py"), submodule_search_locations=[_dir])
_mod = importlib.util.module_from_spec(_spec)
sys.modules["colibri_core_b200"] = _mod
_spec.loader.exec_module(_mod)
