#!/bin/bash
# the tree as committed at the end of round 2: whole GPU suite, smoke(), default bench line
cd ${GRAFT_REPO_ROOT:-.}
mkdir -p gpurun_out
O=gpurun_out
echo "== pytest -m gpu"
S=$(date +%s)
timeout 1500 python -m pytest tests/ -m gpu -x -q -s > $O/r2c25_pytest.log 2>&1; echo "rc=$? ($(( $(date +%s) - S )) s)"; tail -3 $O/r2c25_pytest.log
grep -a "synthesis on identical latents\|parity vs oracle" $O/r2c25_pytest.log | cut -c1-300
echo "== smoke"
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > $O/r2c25_smoke.log 2>&1; echo "rc=$?"; tail -2 $O/r2c25_smoke.log
echo "== bench (default flags)"
S=$(date +%s)
timeout 900 python bench.py > $O/r2c25_bench.json 2> $O/r2c25_bench.err; echo "rc=$? ($(( $(date +%s) - S )) s)"
python - <<'PY'
import json
d = json.loads(open("gpurun_out/r2c25_bench.json").read().strip().splitlines()[-1])
r = d["roofline"]
print("value %.2f e2e %.2f enc %.1f gpu-only %.3f | pipelined %s | roofline %s frac %.3f tensor %.3f whole %.3f | parity %s" % (
    d["value"], d["e2e"]["value"], d["encode_fps"], d["gpu_only_ms_per_decode"], (d.get("pipelined") or {}).get("decode_fps"), r["kernel"], r["frac"], r["tensor_frac"], r["whole_decode_frac"], d.get("parity")))
print("hts %s/%s ld %s/%s htl %s/%s 4K %s/%s | speedups %s" % (d["hts"]["decode_fps"], d["hts"]["encode_fps"], d["ld"]["decode_fps"], d["ld"]["encode_fps"],
      d["htl"]["decode_fps"], d["htl"]["encode_fps"], d["hts_extra"]["decode_fps"], d["hts_extra"]["encode_fps"], d.get("speedup_vs_reference_cuda")))
PY
