"""Locates the B200 package for the script-directory-style modules in this folder.

The reference's modules import each other by bare name from the script directory
(``from process import DataProcess``, pre/read_MSRA.py:27).  The drop-in modules here
keep that convention; this helper loads the enclosing package (whose directory name is
not a Python identifier) under the alias ``hpe3d_b200``.
"""
import importlib.util
import os
import sys

_ALIAS = "hpe3d_b200"


def package():
    mod = sys.modules.get(_ALIAS)
    if mod is not None:
        return mod
    pkg_dir = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    spec = importlib.util.spec_from_file_location(
        _ALIAS, os.path.join(pkg_dir, "__init__.py"), submodule_search_locations=[pkg_dir])
    mod = importlib.util.module_from_spec(spec)
    sys.modules[_ALIAS] = mod
    spec.loader.exec_module(mod)
    return mod
