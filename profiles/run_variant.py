"""Diagnostics: run a repo script (e.g. bench.py) against a diagnostic build: D3IL_VARIANT=<name> python profiles/run_variant.py bench.py --steps 60"""
import os
import runpy
import sys

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import _variant  # noqa: F401,E402

script = sys.argv[1]
sys.argv = sys.argv[1:]
runpy.run_path(script, run_name="__main__")
