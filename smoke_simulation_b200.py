"""Import shim: exposes the directory ``smoke-simulation_b200/`` as the package ``smoke_simulation_b200``
(a hyphen is not a valid identifier).  The real package is loaded from that directory and takes this
module's place in ``sys.modules``."""
import importlib.util as _ilu
import os as _os
import sys as _sys

_dir = _os.path.join(_os.path.dirname(_os.path.abspath(__file__)), "smoke-simulation_b200")
_spec = _ilu.spec_from_file_location(__name__, _os.path.join(_dir, "__init__.py"), submodule_search_locations=[_dir])
_mod = _ilu.module_from_spec(_spec)
_sys.modules[__name__] = _mod
_spec.loader.exec_module(_mod)
