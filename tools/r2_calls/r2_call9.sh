#!/bin/bash
# the reference's own CUDA extension (compiled here for sm_100a) on the box: does it run, how fast, how close are we
cd ${GRAFT_REPO_ROOT:-.}
mkdir -p gpurun_out
O=gpurun_out
echo "== reference CUDA: intra only"
timeout 600 python baseline/run_ref_cuda.py --steps 5 --warmup 2 --models intra > $O/r2_refcuda_intra.json 2> $O/r2_refcuda_intra.err
echo "rc=$?"; tail -c 1500 $O/r2_refcuda_intra.json; tail -c 1500 $O/r2_refcuda_intra.err
echo "== reference CUDA: all models"
timeout 900 python baseline/run_ref_cuda.py --steps 10 --warmup 3 > $O/r2_refcuda_all.json 2> $O/r2_refcuda_all.err
echo "rc=$?"; tail -c 2500 $O/r2_refcuda_all.json; tail -c 800 $O/r2_refcuda_all.err
echo "== parity vs reference CUDA"
timeout 1200 python -m pytest tests/test_reference_cuda_gpu.py -q -s > $O/r2_parity_refcuda.log 2>&1
echo "rc=$?"; grep -E "parity vs|passed|failed|Error" $O/r2_parity_refcuda.log | head -20
