#!/bin/bash
# two ranks after the job deal changed (rotated blocks: every rank gets a mix of rate points): headline + configs[3] job list
cd ${GRAFT_REPO_ROOT:-.}
mkdir -p gpurun_out
O=gpurun_out
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29541 \
    bench.py --gpus 2 --steps 20 --warmup 5 > $O/r2c28_bench_n2.json 2> $O/r2c28_bench_n2.err; echo "rc=$?"
python - <<'PY'
import json
try:
    d = json.loads(open("gpurun_out/r2c28_bench_n2.json").read().strip().splitlines()[-1])
    s = d.get("seq8") or {}
    print("N=2: value %.1f (%.3f ms/step) | seq8 %s" % (d["value"], d["ms_per_step"], {k: s.get(k) for k in ("jobs_per_rank", "protocol_decode_fps", "protocol_encode_fps", "aggregate_decode_fps", "aggregate_encode_fps", "error")}))
except Exception as e:
    print("no line:", e); print(open("gpurun_out/r2c28_bench_n2.err").read()[-1500:])
PY
