"""Import helper: the package directory is `cuda-cholesky_b200/` (dash, as the repo layout asks),
which Python cannot import by name.  load() registers it as module `cuda_cholesky_b200`."""
import importlib.util
import os
import sys

ROOT = os.path.dirname(os.path.abspath(__file__))
PKG_DIR = os.path.join(ROOT, "cuda-cholesky_b200")


def load():
    name = "cuda_cholesky_b200"
    if name in sys.modules:
        return sys.modules[name]
    spec = importlib.util.spec_from_file_location(name, os.path.join(PKG_DIR, "__init__.py"),
                                                  submodule_search_locations=[PKG_DIR])
    mod = importlib.util.module_from_spec(spec)
    sys.modules[name] = mod
    spec.loader.exec_module(mod)
    return mod
