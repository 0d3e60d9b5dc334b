"""Importable alias: the package directory is named ``wavesandeigenvalues.jl_b200`` (with a dot), which the
import system cannot address directly.  ``import wae_b200`` loads that directory as the package ``wae_b200``."""
import importlib.util
import os
import sys

_dir = os.path.join(os.path.dirname(os.path.abspath(__file__)), "wavesandeigenvalues.jl_b200")
_spec = importlib.util.spec_from_file_location("wae_b200", os.path.join(_dir, "__init__.py"), submodule_search_locations=[_dir])
_mod = importlib.util.module_from_spec(_spec)
sys.modules["wae_b200"] = _mod
_spec.loader.exec_module(_mod)
