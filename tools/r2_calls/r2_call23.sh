#!/bin/bash
# two GPUs of one box, launched as the driver does: the headline (weak scaling by frame), the sharded configs[3] job list,
# and the reference arm's rank gating
cd ${GRAFT_REPO_ROOT:-.}
mkdir -p gpurun_out
O=gpurun_out
nvidia-smi -L | head -4
S=$(date +%s)
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 \
    bench.py --gpus 2 --steps 20 --warmup 5 > $O/r2c23_bench_n2.json 2> $O/r2c23_bench_n2.err; echo "rc=$? ($(( $(date +%s) - S )) s)"
python - <<'PY'
import json
try:
    d = json.loads(open("gpurun_out/r2c23_bench_n2.json").read().strip().splitlines()[-1])
    s = d.get("seq8") or {}
    print("N=2: value %.1f e2e %.1f enc %.1f ms/step %.3f gpu-only %.3f | hts %s %s | ld %s | htl %s | seq8 protocol %s %s aggregate %s %s jobs/rank %s" % (
        d["value"], d["e2e"]["value"], d["encode_fps"], d["ms_per_step"], d["gpu_only_ms_per_decode"], d["hts"].get("decode_fps"), d["hts"].get("encode_fps"),
        d["ld"].get("decode_fps"), d["htl"].get("decode_fps"), s.get("protocol_decode_fps"), s.get("protocol_encode_fps"), s.get("aggregate_decode_fps"),
        s.get("aggregate_encode_fps"), s.get("jobs_per_rank")))
    print("clocks", d.get("clocks"), "host", d.get("host"))
except Exception as e:
    print("no N=2 line:", e); print(open("gpurun_out/r2c23_bench_n2.err").read()[-1500:])
PY
S=$(date +%s)
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 \
    bench.py --impl reference --gpus 2 --steps 2 --warmup 1 > $O/r2c23_ref_n2.json 2> $O/r2c23_ref_n2.err; echo "rc=$? ($(( $(date +%s) - S )) s)"
tail -c 700 $O/r2c23_ref_n2.json
