"""Import alias: ``import hamt_b200`` loads the package that lives in ``vln-hamt_b200/``.

(The package directory name is fixed by the build contract and contains a hyphen, so it cannot be
imported by name; this module replaces itself in ``sys.modules`` with the real package.)
"""
import importlib.util
import os
import sys

_here = os.path.dirname(os.path.abspath(__file__))
_pkg_dir = os.path.join(_here, "vln-hamt_b200")
_spec = importlib.util.spec_from_file_location("hamt_b200", os.path.join(_pkg_dir, "__init__.py"),
                                               submodule_search_locations=[_pkg_dir])
_mod = importlib.util.module_from_spec(_spec)
sys.modules["hamt_b200"] = _mod
_spec.loader.exec_module(_mod)
