#!/bin/bash
# The first gpurun call of the next round, in one script (about 50 minutes of box time on one B200; the steps are
# independent — cut the list when the budget is tight):
#   /usr/local/graft/bin/gpurun --timeout 3600 -- 'bash tools/round2_first_call.sh'
# Every step runs under its own timeout and writes to gpurun_out/r2_*.{log,json,csv}, so a hang costs one step and
# whatever finished before it comes back.  What it answers (DESIGN.md §0, round-2 plan):
#   1. are the device cases that only ran under emulation green (4K HT-S / LD, sequence driver, recon-head lanes)?
#   2. does the experimental HT-L path pass its first device run (tools/validate_htl.sh)?
#   3. do parallel graph branches pay: recon-head lanes (DCVC_B200_HEAD_LANES) and half-picture lanes (DCVC_B200_SPLIT_P8)?
#   4. what is the throughput of two concurrent Intra decodes per GPU (bench.py --pipelined)?
#   5. where does the steady state of the N = K = 384 GEMM go: operand ingest, tensor pipe or epilogue
#      (DCVC_B200_GEMM_DBG 0 / 1 / 2 / 3 on the streaming, A-resident and CTA-pair kernels)?
#   6. the round's ncu evidence (tools/profile_round.sh).
cd ${GRAFT_REPO_ROOT:-.}
mkdir -p gpurun_out
O=gpurun_out

echo "== 1. pytest -m gpu"
DCVC_B200_TEST_LANES=1 timeout 1500 python -m pytest tests -m gpu -q -x > $O/r2_pytest_gpu.log 2>&1
echo "rc=$?"; tail -3 $O/r2_pytest_gpu.log

echo "== 2. HT-L first device run"
timeout 1500 bash tools/validate_htl.sh > $O/r2_validate_htl.log 2>&1
tail -12 $O/r2_validate_htl.log

echo "== 3. capture lanes: recon-head branches (HT-S) and half-picture branches (Intra synthesis, HT-S P8 chains)"
# PDL matters here: a dependent kernel that was launched early sits in griddepcontrol.wait ON an SM — it may take the SMs a
# finished CTA frees before the other branch's ready CTAs get them.  So every variant is measured with PDL on and off
# inside the lane regions (DCVC_B200_LANES_PDL; the serial parts of the graphs keep PDL either way).
for V in "lanes1:1:0:1" "lanes2:2:0:1" "lanes4:4:0:1" "lanes2nopdl:2:0:0" "lanes4nopdl:4:0:0" "split:1:1:1" "splitnopdl:1:1:0" "split4:1:4:1" "split4nopdl:1:4:0"; do
    IFS=: read NAME LN SP PDL <<< "$V"
    DCVC_B200_LANES_PDL=$PDL DCVC_B200_HEAD_LANES=$LN DCVC_B200_SPLIT_P8=$SP timeout 400 python bench.py --steps 6 --warmup 3 --no-cpu-baseline \
        > $O/r2_bench_$NAME.json 2> $O/r2_bench_$NAME.err
    python - <<EOF
import json
try:
    d = json.loads(open("$O/r2_bench_$NAME.json").read().strip().splitlines()[-1])
    h = d["hts"]
    print("$NAME: hts decode %.1f FPS gpu-only %.3f ms/chunk encode %.1f FPS | intra decode %.1f FPS gpu-only %.3f ms encode %.1f FPS" %
          (h["decode_fps"], h["gpu_only_ms_per_chunk_decode"], h["encode_fps"], d["value"], d["gpu_only_ms_per_decode"], d["encode_fps"]))
except Exception as e:
    print("$NAME: no result (%s)" % e)
EOF
done

echo "== 3b. HT-S at 4K (configs[4]; published reference CUTLASS build on B200: 424.0 / 289.5 FPS)"
timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --hts-size 2160x3840 > $O/r2_bench_4k.json 2> $O/r2_bench_4k.err
python -c "
import json
d = json.loads(open('$O/r2_bench_4k.json').read().strip().splitlines()[-1])
print('hts 4K:', d.get('hts_extra'))" 2>&1 | tail -1

echo "== 4. two concurrent Intra decodes"
timeout 400 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-hts --pipelined > $O/r2_bench_pipelined.json 2> $O/r2_bench_pipelined.err
python -c "
import json
d = json.loads(open('$O/r2_bench_pipelined.json').read().strip().splitlines()[-1])
print('single', d['value'], 'FPS; pipelined', d.get('pipelined'))" 2>&1 | tail -1

echo "== 4a. one host wait per prior step in the Intra decoder"
DCVC_B200_DECODE_ONE_SYNC=1 timeout 400 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-hts > $O/r2_bench_onesync.json 2> $O/r2_bench_onesync.err
python -c "
import json
d = json.loads(open('$O/r2_bench_onesync.json').read().strip().splitlines()[-1])
print('one-sync: decode', d['value'], 'e2e', d['e2e']['value'])" 2>&1 | tail -1

echo "== 4b. host threads pinned to the GPU's NUMA node"
DCVC_B200_NUMA_PIN=1 timeout 400 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-hts > $O/r2_bench_numa.json 2> $O/r2_bench_numa.err
python -c "
import json
d = json.loads(open('$O/r2_bench_numa.json').read().strip().splitlines()[-1])
print('numa-pinned: decode', d['value'], 'e2e', d['e2e']['value'], 'encode', d['encode_fps'], d.get('host'))" 2>&1 | tail -1

echo "== 5. where the steady state of the 384x384 GEMM goes (us per launch, M = 32640)"
for variant in "ARES=0" "ARES=1 PAIR=0" "ARES=1 PAIR=1"; do
    for dbg in 0 1 2 3; do
        envs="DCVC_B200_GEMM_DBG=$dbg"
        for kv in $variant; do envs="$envs DCVC_B200_GEMM_$kv"; done
        echo -n "[$variant dbg=$dbg] "
        env $envs timeout 120 python tools/gemm_micro.py 136 240 384 384 0 0 1 2>&1 | tail -1
    done
done | tee $O/r2_gemm_steady_state.log
for shape in "136 240 384 1536 1 1 0" "136 240 512 256 0 0 1" "68 120 512 512 1 0 0"; do
    for dbg in 0 3; do
        echo -n "[stream dbg=$dbg] "
        DCVC_B200_GEMM_DBG=$dbg timeout 120 python tools/gemm_micro.py $shape 2>&1 | tail -1
    done
done | tee -a $O/r2_gemm_steady_state.log

echo "== 6. ncu evidence"
timeout 1500 bash tools/profile_round.sh r2a > $O/r2_profile_round.log 2>&1
tail -8 $O/r2_profile_round.log
