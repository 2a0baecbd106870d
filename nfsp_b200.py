"""Import alias: `import nfsp_b200` loads the package directory
`neural-ficititious-self-play-in-imperfect-information-games_b200/` (whose name is not a valid
Python identifier) under the module name `nfsp_b200`."""
import importlib.util as _u
import os as _os
import sys as _sys

_dir = _os.path.join(_os.path.dirname(_os.path.abspath(__file__)),
                     "neural-ficititious-self-play-in-imperfect-information-games_b200")
_spec = _u.spec_from_file_location("nfsp_b200", _os.path.join(_dir, "__init__.py"), submodule_search_locations=[_dir])
_mod = _u.module_from_spec(_spec)
_sys.modules["nfsp_b200"] = _mod
_spec.loader.exec_module(_mod)
