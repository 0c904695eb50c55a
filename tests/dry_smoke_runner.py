"""__graft_entry__.smoke() inside the dry-run subprocess (tests/test_host_dry_run.py): python dry_smoke_runner.py <libdcvc_dry.so>"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import dry_torch_patch  # noqa: E402

dry_torch_patch.apply()

from dcvc_b200 import _lib  # noqa: E402

_lib.LIB_PATH = sys.argv[1]

import __graft_entry__ as entry  # noqa: E402

entry.smoke()
print("smoke-under-emulation ok")
