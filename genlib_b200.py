"""Import shim: the package directory is named `genlib.jl_b200/` (after the
reference, GenLib.jl), which is not a valid dotted module name, so it is
loaded here under the module name `genlib_jl_b200` and re-exported.

    import genlib_b200 as gen
"""
import importlib.util as _u
import os as _os
import sys as _sys

_NAME = "genlib_jl_b200"
if _NAME not in _sys.modules:
    _dir = _os.path.join(_os.path.dirname(_os.path.abspath(__file__)), "genlib.jl_b200")
    _spec = _u.spec_from_file_location(_NAME, _os.path.join(_dir, "__init__.py"),
                                       submodule_search_locations=[_dir])
    _mod = _u.module_from_spec(_spec)
    _sys.modules[_NAME] = _mod
    _spec.loader.exec_module(_mod)
_mod = _sys.modules[_NAME]
globals().update({k: v for k, v in vars(_mod).items() if not k.startswith("__")})
