"""Import alias: the product package lives in ./neurips-cellseg_b200/ (a hyphen is
not importable), so `import cellseg_b200` loads that directory as a package."""
import importlib.util as _u
import os as _os
import sys as _sys

_dir = _os.path.join(_os.path.dirname(_os.path.abspath(__file__)), "neurips-cellseg_b200")
_spec = _u.spec_from_file_location("cellseg_b200", _os.path.join(_dir, "__init__.py"),
                                   submodule_search_locations=[_dir])
_mod = _u.module_from_spec(_spec)
_sys.modules["cellseg_b200"] = _mod
_spec.loader.exec_module(_mod)
