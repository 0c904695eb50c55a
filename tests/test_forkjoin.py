"""Host logic: the fork-join pool of the rANS coder (busy-wait then block) must neither lose wake-ups nor run a task
twice — checked with a compiled stress program under three spin settings."""
import os
import shutil
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.skipif(shutil.which("g++") is None, reason="g++ not available")
def test_forkjoin_stress(tmp_path):
    exe = str(tmp_path / "forkjoin_stress")
    subprocess.run(["g++", "-O2", "-std=c++17", "-I", os.path.join(ROOT, "dcvc_b200", "csrc"),
                    os.path.join(ROOT, "tests", "cpp", "forkjoin_stress.cpp"),
                    os.path.join(ROOT, "dcvc_b200", "csrc", "rans_host.cpp"), "-o", exe, "-lpthread"], check=True)
    for spin_us in ("0", "100", "2000"):
        env = dict(os.environ, DCVC_B200_RANS_SPIN_US=spin_us)
        r = subprocess.run([exe], env=env, capture_output=True, text=True, timeout=120)
        assert r.returncode == 0 and " OK " in r.stdout, f"spin_us={spin_us}: {r.stdout} {r.stderr}"
