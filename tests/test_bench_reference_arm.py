"""`bench.py --impl reference` — the reference's CPU path (oracle port + the reference's own rANS coder) timed on the
host cores.  It needs no GPU, so the contract of its JSON line is checked here: same metric / unit / config keys as
the product arm, `impl`, a `cpu_baseline` describing the run, an `e2e` with zero copy bytes; ranks other than 0 of a
torchrun launch print nothing and exit 0."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _run(extra_env=None):
    env = dict(os.environ)
    env.update(extra_env or {})
    return subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "0"],
                          env=env, capture_output=True, text=True, timeout=600, cwd=ROOT)


def test_reference_arm_json_line():
    r = _run()
    assert r.returncode == 0, r.stderr[-3000:]
    lines = [l for l in r.stdout.splitlines() if l.startswith("{")]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["metric"] == "1080p_yuv_decode_fps" and d["unit"] == "frames/s"
    assert d["higher_is_better"] is True and d["steps"] == 1 and d["value"] > 0
    assert d["config"]["resolution"] == [1080, 1920] and "workload" in d["config"]
    cb = d["cpu_baseline"]
    assert cb["kind"] in ("port", "reference") and cb["cores"] >= 1 and cb["value"] == d["value"] and cb["sample"]
    assert d["e2e"] == {"value": d["value"], "unit": "frames/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}


def test_reference_arm_other_ranks_stay_silent():
    r = _run({"RANK": "1", "WORLD_SIZE": "2", "LOCAL_RANK": "1"})
    assert r.returncode == 0 and r.stdout.strip() == ""
