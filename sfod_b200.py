"""Import shim: exposes the package directory ``simple-sfod_b200/`` (not a valid Python identifier) as the
module ``sfod_b200``.  ``import sfod_b200`` executes this file, which replaces itself in ``sys.modules`` by
the real package so that ``sfod_b200.ops`` / ``from sfod_b200.modeling import ...`` resolve normally."""
import importlib.util
import os
import sys

_pkg_dir = os.path.join(os.path.dirname(os.path.abspath(__file__)), "simple-sfod_b200")
_spec = importlib.util.spec_from_file_location("sfod_b200", os.path.join(_pkg_dir, "__init__.py"),
                                               submodule_search_locations=[_pkg_dir])
_mod = importlib.util.module_from_spec(_spec)
sys.modules["sfod_b200"] = _mod
_spec.loader.exec_module(_mod)
