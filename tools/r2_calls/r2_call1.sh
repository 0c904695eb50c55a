#!/bin/bash
# Round 2, first device call: everything that only ran under emulation in round 1 (HT-L, capture lanes, one-sync decode),
# then one A/B per host-side switch and the ingest / tensor / epilogue split of the 384x384 GEMM.
cd ${GRAFT_REPO_ROOT:-.}
mkdir -p gpurun_out
O=gpurun_out
export DCVC_B200_EXPERIMENTAL_HTL=1
echo "== 1. pytest -m gpu (HT-L + lanes enabled)"
DCVC_B200_TEST_LANES=1 timeout 900 python -m pytest tests -m gpu -q > $O/r2_pytest_gpu.log 2>&1
echo "rc=$?"; tail -15 $O/r2_pytest_gpu.log
echo "== 2. switches A/B"
for V in "base:1:0:1:0" "lanes2:2:0:1:0" "lanes4:4:0:1:0" "lanes4nopdl:4:0:0:0" "split:1:1:1:0" "splitnopdl:1:1:0:0" "onesync:1:0:1:1"; do
    IFS=: read NAME LN SP PDL OS <<< "$V"
    DCVC_B200_DECODE_ONE_SYNC=$OS DCVC_B200_LANES_PDL=$PDL DCVC_B200_HEAD_LANES=$LN DCVC_B200_SPLIT_P8=$SP timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline \
        > $O/r2_bench_$NAME.json 2> $O/r2_bench_$NAME.err
    python - <<PY
import json
try:
    d = json.loads(open("$O/r2_bench_$NAME.json").read().strip().splitlines()[-1])
    h = d["hts"]
    print("$NAME: hts dec %.1f FPS gpu-only %.3f ms/chunk enc %.1f | intra dec %.1f e2e %.1f gpu-only %.3f ms enc %.1f | htl %s" %
          (h["decode_fps"], h["gpu_only_ms_per_chunk_decode"], h["encode_fps"], d["value"], d["e2e"]["value"], d["gpu_only_ms_per_decode"], d["encode_fps"], d.get("htl")))
except Exception as e:
    print("$NAME: no result (%s)" % e)
PY
done
echo "== 3. GEMM steady state"
for variant in "ARES=0" "ARES=1 PAIR=1"; do
    for dbg in 0 1 2 3; do
        envs="DCVC_B200_GEMM_DBG=$dbg"
        for kv in $variant; do envs="$envs DCVC_B200_GEMM_$kv"; done
        echo -n "[$variant dbg=$dbg] "
        env $envs timeout 120 python tools/gemm_micro.py 136 240 384 384 0 0 1 2>&1 | tail -1
    done
done | tee $O/r2_gemm_steady_state.log
