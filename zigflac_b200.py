"""Alias: `import zigflac_b200` loads the package directory `zig-flac_b200/` (a hyphen cannot be
written in an import statement)."""
import importlib.util
import os
import sys

_dir = os.path.join(os.path.dirname(os.path.abspath(__file__)), "zig-flac_b200")
_spec = importlib.util.spec_from_file_location("zigflac_b200", os.path.join(_dir, "__init__.py"),
                                               submodule_search_locations=[_dir])
_mod = importlib.util.module_from_spec(_spec)
sys.modules["zigflac_b200"] = _mod
sys.modules["zig-flac_b200"] = _mod
_spec.loader.exec_module(_mod)
