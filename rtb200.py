"""Import shim: the package directory `win32-ray-tracing-demo_b200/` is not a valid Python
identifier, so `import rtb200` loads it under this name."""
import importlib.util
import os
import sys

_PKG = os.path.join(os.path.dirname(os.path.abspath(__file__)), "win32-ray-tracing-demo_b200")
_spec = importlib.util.spec_from_file_location("rtb200", os.path.join(_PKG, "__init__.py"),
                                               submodule_search_locations=[_PKG])
_mod = importlib.util.module_from_spec(_spec)
sys.modules["rtb200"] = _mod
_spec.loader.exec_module(_mod)
