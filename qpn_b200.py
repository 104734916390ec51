"""Import shim: loads the package that lives in `quadraticprogramnetworks.jl_b200/`
(a directory name Python cannot import because of the dot) under the name `qpn_b200`."""
import importlib.util
import os
import sys

_dir = os.path.join(os.path.dirname(os.path.abspath(__file__)), "quadraticprogramnetworks.jl_b200")
_spec = importlib.util.spec_from_file_location("qpn_b200", os.path.join(_dir, "__init__.py"),
                                               submodule_search_locations=[_dir])
_mod = importlib.util.module_from_spec(_spec)
sys.modules["qpn_b200"] = _mod
_spec.loader.exec_module(_mod)
