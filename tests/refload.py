"""Loads the UNMODIFIED reference build (baseline/_ref, installed from /root/reference by
__graft_entry__.build()) under the alias ``diso_ref``; returns None when it is not available."""
import importlib.util
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def load_reference():
    if "diso_ref" in sys.modules:
        return sys.modules["diso_ref"]
    pkg = os.path.join(ROOT, "baseline", "_ref", "diso")
    if not os.path.exists(os.path.join(pkg, "_C.so")):
        return None
    try:
        spec = importlib.util.spec_from_file_location("diso_ref", os.path.join(pkg, "__init__.py"),
                                                      submodule_search_locations=[pkg])
        mod = importlib.util.module_from_spec(spec)
        sys.modules["diso_ref"] = mod
        spec.loader.exec_module(mod)
        return mod
    except Exception:
        sys.modules.pop("diso_ref", None)
        return None
