"""Import shim: the package directory is `upsp-processing_b200/` (hyphenated, mirroring the
reference project's name), which Python cannot import by name.  `import upsp_b200` loads it
under the module name `upsp_processing_b200` and re-exports it."""
import importlib.util
import os
import sys

_PKG_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "upsp-processing_b200")
_NAME = "upsp_processing_b200"


def load():
    if _NAME in sys.modules:
        return sys.modules[_NAME]
    spec = importlib.util.spec_from_file_location(
        _NAME, os.path.join(_PKG_DIR, "__init__.py"), submodule_search_locations=[_PKG_DIR])
    mod = importlib.util.module_from_spec(spec)
    sys.modules[_NAME] = mod
    spec.loader.exec_module(mod)
    return mod


pkg = load()
synth = importlib.import_module(_NAME + ".synth")
build = importlib.import_module(_NAME + ".build")
globals().update({k: getattr(pkg, k) for k in dir(pkg) if not k.startswith("__")})
