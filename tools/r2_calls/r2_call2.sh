#!/bin/bash
# Round 2, second device call: first runs of the fused DepthConvBlock-tail kernel (each case in its own process under a
# timeout: a hang costs one case), then the reference-surface tests.
cd ${GRAFT_REPO_ROOT:-.}
mkdir -p gpurun_out
O=gpurun_out
for K in "16-16-128" "16-16-384" "17-30-384" "32-40-256" "68-120-512" "136-240-384-384-384" "136-240-384-384-0" "pitched" "declines"; do
    echo "== dcb_tail $K"
    timeout 150 python -m pytest tests/test_dcb_tail_gpu.py -x -q -s -k "$K" > $O/r2_dcbtail_$K.log 2>&1
    echo "rc=$?"; grep -E "dcb_tail\]|passed|failed|mismatch|Error|error" $O/r2_dcbtail_$K.log | head -8
done
echo "== reference surface"
timeout 1200 python -m pytest tests/test_reference_surface_gpu.py -q -x > $O/r2_refsurface.log 2>&1
echo "rc=$?"; tail -30 $O/r2_refsurface.log
