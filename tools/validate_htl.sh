#!/bin/bash
# First device run of the experimental HT-L path (run under gpurun on one B200): the two GEMM-kind tests HT-L adds
# (3x3 + pixel_shuffle(2); 2x2 transposed conv with bias), then the HT-L codec tests (state identity, oracle, stream
# bytes), each under its own timeout so that a hang costs one step, not the box.  Logs go to gpurun_out/htl_*.log.
cd ${GRAFT_REPO_ROOT:-.}
mkdir -p gpurun_out
export DCVC_B200_EXPERIMENTAL_HTL=1
timeout 300 python -m pytest tests/test_ops_gpu.py -q -x -k "ps2 or tconv2x2_with_bias" > gpurun_out/htl_ops.log 2>&1
echo "ops rc=$?"; tail -3 gpurun_out/htl_ops.log
timeout 600 python -m pytest tests/test_htl_gpu.py -q -x -k "64-64 or oracle" > gpurun_out/htl_small.log 2>&1
echo "small rc=$?"; tail -3 gpurun_out/htl_small.log
timeout 900 python -m pytest tests/test_htl_gpu.py -q > gpurun_out/htl_all.log 2>&1
echo "all rc=$?"; tail -5 gpurun_out/htl_all.log
timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/htl_bench.json 2> gpurun_out/htl_bench.err
echo "bench rc=$?"; python -c "import json;d=json.loads(open('gpurun_out/htl_bench.json').read().strip().splitlines()[-1]);print(d.get('htl'))"
