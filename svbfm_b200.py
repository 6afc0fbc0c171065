"""Import shim: the package directory name contains hyphens (`scalable-variational-bayesian-factorization-machine_b200`),
so it is loaded by path and exposed as the module `svbfm_b200`."""
import importlib.util
import os
import sys

_PKG_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "scalable-variational-bayesian-factorization-machine_b200")
_NAME = "svbfm_b200_pkg"

if _NAME not in sys.modules:
    _spec = importlib.util.spec_from_file_location(_NAME, os.path.join(_PKG_DIR, "__init__.py"),
                                                   submodule_search_locations=[_PKG_DIR])
    _mod = importlib.util.module_from_spec(_spec)
    sys.modules[_NAME] = _mod
    _spec.loader.exec_module(_mod)
_pkg = sys.modules[_NAME]
globals().update({k: v for k, v in vars(_pkg).items() if not k.startswith("__")})
PKG_DIR = _PKG_DIR


def submodule(name):
    """Load `<package>/<name>.py` (e.g. synth, dist)."""
    full = f"{_NAME}.{name}"
    if full not in sys.modules:
        spec = importlib.util.spec_from_file_location(full, os.path.join(_PKG_DIR, name + ".py"))
        mod = importlib.util.module_from_spec(spec)
        sys.modules[full] = mod
        spec.loader.exec_module(mod)
    return sys.modules[full]
