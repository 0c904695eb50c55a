#!/bin/bash
# eight GPUs of one box, launched as the driver does: headline (weak scaling) + the configs[3] job list dealt to 8 ranks
cd ${GRAFT_REPO_ROOT:-.}
mkdir -p gpurun_out
O=gpurun_out
nvidia-smi -L | wc -l
S=$(date +%s)
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29531 \
    bench.py --gpus 8 --steps 20 --warmup 5 > $O/r2c27_bench_n8.json 2> $O/r2c27_bench_n8.err; echo "rc=$? ($(( $(date +%s) - S )) s)"
python - <<'PY'
import json
try:
    d = json.loads(open("gpurun_out/r2c27_bench_n8.json").read().strip().splitlines()[-1])
    s = d.get("seq8") or {}
    print("N=8: value %.1f (%.3f ms/step) e2e %.1f enc %.1f gpu-only %.3f | seq8 protocol %s %s aggregate %s %s jobs/rank %s | clocks %s host %s" % (
        d["value"], d["ms_per_step"], d["e2e"]["value"], d["encode_fps"], d["gpu_only_ms_per_decode"], s.get("protocol_decode_fps"), s.get("protocol_encode_fps"),
        s.get("aggregate_decode_fps"), s.get("aggregate_encode_fps"), s.get("jobs_per_rank"), d.get("clocks"), d.get("host")))
except Exception as e:
    print("no N=8 line:", e); print(open("gpurun_out/r2c27_bench_n8.err").read()[-1500:])
PY
