"""Runs selected `-m gpu` test files unmodified inside the dry-run subprocess (see tests/dry_torch_patch.py):

    LD_PRELOAD=<shim> DRY_SHIM_EMULATE=1 python dry_pytest_runner.py <libdcvc_dry.so> <pytest args...>"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import dry_torch_patch  # noqa: E402

dry_torch_patch.apply()

from dcvc_b200 import _lib  # noqa: E402

_lib.LIB_PATH = sys.argv[1]          # the relinked copy; the shipped library's static runtime cannot be interposed

import pytest  # noqa: E402

sys.exit(pytest.main(["-p", "no:cacheprovider"] + sys.argv[2:]))
