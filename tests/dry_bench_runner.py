"""bench.py's control flow inside the dry-run subprocess (tests/test_host_dry_run.py):
    python dry_bench_runner.py <libdcvc_dry.so> [bench args]
The timings are made up (the shim's events), the size is tiny (DCVC_B200_BENCH_TEST_SIZE): what is checked is that the
file the driver runs unattended at round end still executes end to end and prints the JSON line of the contract."""
import os
import runpy
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import torch  # noqa: E402

import dry_torch_patch  # noqa: E402

dry_torch_patch.apply()
torch.cuda.set_stream = lambda *_: None
torch.Tensor.pin_memory = lambda self, *a, **k: self
torch.cuda.Event.elapsed_time = lambda self, other: 0.5
_orig_copy = torch.Tensor.copy_
torch.Tensor.copy_ = lambda self, src, non_blocking=False: _orig_copy(self, src)

from dcvc_b200 import _lib  # noqa: E402

_lib.LIB_PATH = sys.argv[1]
sys.argv = [os.path.join(ROOT, "bench.py")] + sys.argv[2:]
runpy.run_path(os.path.join(ROOT, "bench.py"), run_name="__main__")
