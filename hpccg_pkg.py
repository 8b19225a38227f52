"""Imports the package directory `hpccg-sycl_b200/` (a hyphen is not a valid module name) as the
module `hpccg_sycl_b200`.  Used by tests/, bench.py and __graft_entry__.py."""
import importlib.util
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parent
PKG_DIR = ROOT / "hpccg-sycl_b200"
MODULE_NAME = "hpccg_sycl_b200"


def load():
    if MODULE_NAME in sys.modules:
        return sys.modules[MODULE_NAME]
    spec = importlib.util.spec_from_file_location(MODULE_NAME, PKG_DIR / "__init__.py",
                                                  submodule_search_locations=[str(PKG_DIR)])
    mod = importlib.util.module_from_spec(spec)
    sys.modules[MODULE_NAME] = mod
    try:
        spec.loader.exec_module(mod)
    except BaseException:
        sys.modules.pop(MODULE_NAME, None)
        raise
    return mod
