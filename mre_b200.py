"""Import shim: the product package lives in ``movie-recommendation-engine_b200/`` (the
directory name the build contract prescribes, which is not a valid Python identifier).
``import mre_b200`` loads that directory as the package ``mre_b200``."""
import importlib.util
import os
import sys

_dir = os.path.join(os.path.dirname(os.path.abspath(__file__)), "movie-recommendation-engine_b200")
_spec = importlib.util.spec_from_file_location(
    "mre_b200", os.path.join(_dir, "__init__.py"), submodule_search_locations=[_dir])
_mod = importlib.util.module_from_spec(_spec)
sys.modules["mre_b200"] = _mod
_spec.loader.exec_module(_mod)
