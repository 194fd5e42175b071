"""Diagnostics only: D3IL_VARIANT=<name> makes the profiling scripts load profiles/build/<name>/libd3il.so
(built by profiles/build_variant.sh) instead of the product library.  Import before the first BatchedEnv."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
from d3il_b200 import lib  # noqa: E402

_v = os.environ.get("D3IL_VARIANT")
if _v:
    lib.SO_PATH = os.path.join(ROOT, "profiles", "build", _v, "libd3il.so")
    assert os.path.exists(lib.SO_PATH), lib.SO_PATH
