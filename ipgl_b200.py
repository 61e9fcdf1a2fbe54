"""Import shim: the package directory is named `image-processing-graph-laplacian_b200` (not an identifier),
so `import ipgl_b200` loads it from there."""
import importlib.util
import os
import sys

_dir = os.path.join(os.path.dirname(os.path.abspath(__file__)), "image-processing-graph-laplacian_b200")
_spec = importlib.util.spec_from_file_location("ipgl_b200", os.path.join(_dir, "__init__.py"),
                                               submodule_search_locations=[_dir])
_mod = importlib.util.module_from_spec(_spec)
sys.modules[__name__] = _mod
_spec.loader.exec_module(_mod)
