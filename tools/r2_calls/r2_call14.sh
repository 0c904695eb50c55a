#!/bin/bash
cd ${GRAFT_REPO_ROOT:-.}
mkdir -p gpurun_out
O=gpurun_out
echo "== reference-surface + reference-cuda tests"
timeout 2400 python -m pytest tests/test_reference_surface_gpu.py tests/test_reference_cuda_gpu.py tests/test_codec_gpu.py -m gpu -q -x > $O/r2c14_pytest.log 2>&1
echo "rc=$?"; tail -4 $O/r2c14_pytest.log
echo "== the default bench line (everything on), timed"
T0=$(date +%s); timeout 1200 python bench.py > $O/r2c14_bench_default.json 2> $O/r2c14_bench_default.err; RC=$?; echo "rc=$RC elapsed $(( $(date +%s) - T0 )) s"
tail -c 600 $O/r2c14_bench_default.err
python - <<PY
import json
d = json.loads(open("$O/r2c14_bench_default.json").read().strip().splitlines()[-1])
for k in ("value","ms_per_step","e2e","encode_fps","gpu_only_ms_per_decode","roofline","cpu_baseline","seq8","hts","ld","htl","hts_extra","reference_cuda","parity","speedup_vs_reference_cuda","clocks"):
    print(k, "=", json.dumps(d.get(k))[:900])
PY
