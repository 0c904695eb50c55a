#!/bin/bash
# two ranks on one box: what the 6 % weak-scaling loss of the Intra headline is made of (core placement / spin budget)
cd ${GRAFT_REPO_ROOT:-.}
mkdir -p gpurun_out
O=gpurun_out
lscpu | grep -E "^CPU\(s\)|Thread|Core|Socket|NUMA" | head -8
cat /sys/bus/pci/devices/*/numa_node 2>/dev/null | sort | uniq -c | head -3
run() {  # name, env...
    local NAME=$1; shift
    env "$@" timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29521 \
        bench.py --gpus 2 --steps 40 --warmup 5 --no-hts --no-seq8 --no-cpu-baseline --no-reference-cuda > $O/r2c24_$NAME.json 2> $O/r2c24_$NAME.err
    python - <<PY
import json
try:
    d = json.loads(open("$O/r2c24_$NAME.json").read().strip().splitlines()[-1])
    print("$NAME: value %.1f (%.3f ms/step) e2e %.1f enc %.1f gpu-only %.3f pinned %s" % (d["value"], d["ms_per_step"], d["e2e"]["value"], d["encode_fps"], d["gpu_only_ms_per_decode"], d["host"]))
except Exception as e:
    print("$NAME: no line (%s)" % e); print(open("$O/r2c24_$NAME.err").read()[-600:])
PY
}
for REP in 1 2; do
    run unpinned_$REP DCVC_B200_PIN=0
    run pinned_$REP DCVC_B200_PIN=1
done
run pinned_spin1000 DCVC_B200_PIN=1 DCVC_B200_RANS_SPIN_US=1000
run unpinned_spin2000 DCVC_B200_PIN=0 DCVC_B200_RANS_SPIN_US=2000
# one rank alone on the same box for the denominator
env CUDA_VISIBLE_DEVICES=0 timeout 300 python bench.py --steps 40 --warmup 5 --no-hts --no-seq8 --no-cpu-baseline --no-reference-cuda --no-pipelined > $O/r2c24_n1.json 2> $O/r2c24_n1.err
python - <<PY
import json
d = json.loads(open("gpurun_out/r2c24_n1.json").read().strip().splitlines()[-1])
print("N=1 on this box: value %.1f (%.3f ms/step) e2e %.1f enc %.1f gpu-only %.3f" % (d["value"], d["ms_per_step"], d["e2e"]["value"], d["encode_fps"], d["gpu_only_ms_per_decode"]))
PY
