"""Import helper: the package directory is named `moshi.cpp_b200` (not a valid Python identifier),
so it is loaded explicitly and registered as `moshi_cpp_b200`."""
import importlib.util
import os
import sys

ROOT = os.path.dirname(os.path.abspath(__file__))
PKG_DIR = os.path.join(ROOT, "moshi.cpp_b200")


def load():
    if "moshi_cpp_b200" in sys.modules:
        return sys.modules["moshi_cpp_b200"]
    spec = importlib.util.spec_from_file_location(
        "moshi_cpp_b200", os.path.join(PKG_DIR, "__init__.py"), submodule_search_locations=[PKG_DIR])
    mod = importlib.util.module_from_spec(spec)
    sys.modules["moshi_cpp_b200"] = mod
    spec.loader.exec_module(mod)
    return mod
