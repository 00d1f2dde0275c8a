"""Import shim: loads the package directory `recommendation-system-maang-nvidia-_b200/` (whose name is not
a Python identifier) as the package `recsys_b200`, so `import recsys_b200` works from the repo root."""

import importlib.util
import sys
from pathlib import Path

_pkg_dir = Path(__file__).resolve().parent / "recommendation-system-maang-nvidia-_b200"
_spec = importlib.util.spec_from_file_location("recsys_b200", _pkg_dir / "__init__.py",
                                               submodule_search_locations=[str(_pkg_dir)])
_mod = importlib.util.module_from_spec(_spec)
sys.modules["recsys_b200"] = _mod
_spec.loader.exec_module(_mod)
